// K4, grouped evaluation: (query, leaf) pairs are bucketed by leaf, and one CTA solves up to
// QB = 8 queries (32 right-hand sides) against one leaf, so a leaf's factor is streamed from
// L2/HBM once per 8 queries instead of once per query.
//
// The record stores, for every off-diagonal tile, G(i,j) = L(i,j) * inv(L(j,j)) and, per block
// row, Dinv(j) = inv(L(j,j)). With U = the right-hand sides after block elimination,
//      U_i = B_i - sum_{j<i} G(i,j) U_j            (no per-block triangular solve on the chain)
//      V_j = Dinv(j) U_j,  var = prior - sum V^2    (off the dependency chain)
// which is algebraically L^{-1} B (OnGPIS.cpp:199-213) evaluated block-wise.
//
// Per CTA: U (npad x 32 floats) lives in shared memory for the whole solve; warp w owns block rows
// i = w, w+8, ...; for block column j each warp pulls its G(i,j) tiles through its own
// double-buffered TMA bulk-copy pipeline (4 KB tiles, mbarrier completion) and applies a 32x32x32
// register-tiled FMA update. Bound: FP32 FMA pipe (2 n^2 MACs per query-evaluation); L2->SM
// traffic is 4 n^2 / 8 bytes per evaluation.
#pragma once
#include <string>

#include "common.cuh"
#include "leaf_train.cuh"
#include "query.cuh"

namespace gpis {

#define QB 8
#define EVAL2_WARPS 8
#define EVAL2_THREADS (EVAL2_WARPS * 32)

struct SortBufs {
    int32_t* count;    // nslots
    int32_t* start;    // nslots + 1   (exclusive prefix of count)
    int32_t* istart;   // nslots + 1   (exclusive prefix of ceil(count/QB))
    int32_t* cursor;   // nslots
    int2* sorted;      // npairs
    int4* items;       // work items of the 8-query class: slot, first sorted pair, count, unused
    int4* itemsB;      // work items of the 4-query class (leaves too large for 32 right-hand sides in smem)
    int4* itemsM;      // work items of the 6-query class
    int2* pairsC;      // pairs of leaves too large for either (one CTA per pair, k_eval_v1)
    int32_t* totals;   // [0] = #items (legacy list), [1] = #itemsA, [2] = #itemsB, [3] = #pairsC, [4] = #itemsM
};

__global__ void k_pair_hist(const int2* __restrict__ pairs, int npairs, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) atomicAdd(&count[pairs[i].y & 0x0fffffff], 1);
}

// single-block exclusive scans over the slots
__global__ void __launch_bounds__(1024) k_slot_scan(SortBufs S, int nslots) {
    __shared__ int s_cnt[1024], s_itm[1024];
    const int t = threadIdx.x;
    const int per = (nslots + 1023) / 1024;
    const int lo = t * per, hi = min(nslots, lo + per);
    int c = 0, it = 0;
    for (int i = lo; i < hi; ++i) { c += S.count[i]; it += (S.count[i] + QB - 1) / QB; }
    s_cnt[t] = c; s_itm[t] = it;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int a = 0, b = 0;
        if (t >= off) { a = s_cnt[t - off]; b = s_itm[t - off]; }
        __syncthreads();
        s_cnt[t] += a; s_itm[t] += b;
        __syncthreads();
    }
    int pc = s_cnt[t] - c, pi = s_itm[t] - it;
    for (int i = lo; i < hi; ++i) {
        S.start[i] = pc; S.istart[i] = pi; S.cursor[i] = pc;
        pc += S.count[i]; pi += (S.count[i] + QB - 1) / QB;
    }
    if (t == 1023) { S.start[nslots] = s_cnt[1023]; S.istart[nslots] = s_itm[1023]; S.totals[0] = s_itm[1023]; }
}

__global__ void k_pair_scatter(const int2* __restrict__ pairs, int npairs, SortBufs S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int2 p = pairs[i];
    const int pos = atomicAdd(&S.cursor[p.y & 0x0fffffff], 1);
    S.sorted[pos] = p;
}

__global__ void k_make_items(SortBufs S, int nslots) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int c = S.count[s];
    if (c == 0) return;
    int it = S.istart[s];
    for (int o = 0; o < c; o += QB) S.items[it++] = make_int4(s, S.start[s] + o, min(QB, c - o), 0);
}

// Work lists by leaf size class (nb = 32-row blocks of the leaf system). A leaf's items are contiguous so that
// CTAs running at the same time share its tiles in L2.
__global__ void k_make_items_classed(SortBufs S, LeafTable T, int nslots, int nbA, int nbM, int nbB) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int c = S.count[s];
    if (c == 0) return;
    const int nb = T.meta[s].w;
    if (nb <= nbA) {
        int it = atomicAdd(&S.totals[1], (c + 7) / 8);
        for (int o = 0; o < c; o += 8) S.items[it++] = make_int4(s, S.start[s] + o, min(8, c - o), 0);
    } else if (nb <= nbM) {
        int it = atomicAdd(&S.totals[4], (c + 5) / 6);
        for (int o = 0; o < c; o += 6) S.itemsM[it++] = make_int4(s, S.start[s] + o, min(6, c - o), 0);
    } else if (nb <= nbB) {
        int it = atomicAdd(&S.totals[2], (c + 3) / 4);
        for (int o = 0; o < c; o += 4) S.itemsB[it++] = make_int4(s, S.start[s] + o, min(4, c - o), 0);
    } else {
        int it = atomicAdd(&S.totals[3], c);
        for (int o = 0; o < c; ++o) S.pairsC[it++] = S.sorted[S.start[s] + o];
    }
}

// Algorithmic work of one evaluation pass, accumulated over leaves with work:
//   acc[0] flops            sum over evaluations of 4n^2 + 16n + 80N
//   acc[1] gather bytes     sum over evaluations of 16N + 4n + 2n(n+1)
//   acc[2] compulsory bytes sum over DISTINCT leaves touched of the same record size
__global__ void k_query_stats(SortBufs S, LeafTable T, int nslots, double* acc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double f = 0, g = 0, c = 0;
    if (s < nslots && S.count[s] > 0) {
        const int4 m = T.meta[s];
        const double N = m.x, n = m.z, cnt = S.count[s];
        const double rec = 16.0 * N + 4.0 * n + 2.0 * n * (n + 1.0);
        f = cnt * (4.0 * n * n + 16.0 * n + 80.0 * N);
        g = cnt * rec;
        c = rec;
    }
    for (int o = 16; o > 0; o >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0 && (f != 0 || c != 0)) { atomicAdd(acc, f); atomicAdd(acc + 1, g); atomicAdd(acc + 2, c); }
}

struct Eval2Smem {
    static constexpr int off_bar = 0;                                         // EVAL2_WARPS * 2 mbarriers
    static constexpr int off_stage = 256;                                     // per warp 2 x 4 KB
    static constexpr int off_red = off_stage + EVAL2_WARPS * 2 * GPIS_TILE_BYTES;  // 2 x 32 x EVAL2_WARPS floats
    static constexpr int off_U = off_red + 2 * 32 * EVAL2_WARPS * 4;
    static int total(int nbmax) { return off_U + nbmax * 32 * 32 * 4; }
};

// C[i][j] (4x8 per lane, rows 4rg.., cols 8cg..) of a row-major [32][32] smem tile
__device__ __forceinline__ void ctile_load(const float* U, float (&acc)[4][8], int rg, int cg) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(U + (4 * rg + i) * 32 + 8 * cg);
        const float4 b = *reinterpret_cast<const float4*>(U + (4 * rg + i) * 32 + 8 * cg + 4);
        acc[i][0] = a.x; acc[i][1] = a.y; acc[i][2] = a.z; acc[i][3] = a.w;
        acc[i][4] = b.x; acc[i][5] = b.y; acc[i][6] = b.z; acc[i][7] = b.w;
    }
}
__device__ __forceinline__ void ctile_store(float* U, const float (&acc)[4][8], int rg, int cg) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(U + (4 * rg + i) * 32 + 8 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(U + (4 * rg + i) * 32 + 8 * cg + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

__global__ void __launch_bounds__(EVAL2_THREADS, 1)
k_eval_v2(const float* __restrict__ x, LeafTable T, QueryParams P, QueryWork W, SortBufs S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int4 item = S.items[blockIdx.x];
    const int slot = item.x, first = item.y, cnt = item.z;
    const int dim = P.dim, w = 1 + dim;
    const unsigned char* rec = reinterpret_cast<const unsigned char*>(T.rec[slot]);
    const int4 meta = T.meta[slot];
    const int N = meta.x, ng = meta.y, n = meta.z, nb = meta.w;
    const int npad = nb * 32;
    const float4* pts = reinterpret_cast<const float4*>(rec + rec_off_pts());
    const float* alpha = reinterpret_cast<const float*>(rec + rec_off_alpha(N));
    const float* dinv = reinterpret_cast<const float*>(rec + rec_off_dinv(N, nb));
    const float* tiles = reinterpret_cast<const float*>(rec + rec_off_tiles(N, nb));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = lane >> 2, cg = lane & 3;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Eval2Smem::off_bar) + warp * 2;
    float* stg[2] = {reinterpret_cast<float*>(smem_raw + Eval2Smem::off_stage) + (warp * 2 + 0) * GPIS_TILE_ELEMS,
                     reinterpret_cast<float*>(smem_raw + Eval2Smem::off_stage) + (warp * 2 + 1) * GPIS_TILE_ELEMS};
    float* red = reinterpret_cast<float*>(smem_raw + Eval2Smem::off_red);   // [2][EVAL2_WARPS][32]
    float* U = reinterpret_cast<float*>(smem_raw + Eval2Smem::off_U);       // [npad][32]

    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }

    // ---- right-hand sides: k* of every query of the item (covFnc.cpp:282-311 / 425-446)
    {
        float4* U4 = reinterpret_cast<float4*>(U);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < npad * 8; i += EVAL2_THREADS) U4[i] = z4;
    }
    __syncthreads();
    for (int idx = tid; idx < N * QB; idx += EVAL2_THREADS) {
        const int qi = idx % QB, k = idx / QB;
        if (qi >= cnt) continue;
        const int q = S.sorted[first + qi].x;
        float xq[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < dim; ++c) xq[c] = x[(int64_t)q * dim + c];
        const float4 p = pts[k];
        const int g = __float_as_int(p.w);
        const float xs[3] = {p.x, p.y, p.z};
        float d[3], s2 = 0.f;
        for (int c = 0; c < dim; ++c) { d[c] = xs[c] - xq[c]; s2 = (c == 0) ? d[c] * d[c] : s2 + d[c] * d[c]; }
        const float r = sqrtf(s2);
        const DF e = exp_df(-P.a * r);
        float* col = U + 4 * qi;
        col[k * 32] = kf_val(r, P.a, e);
        float k1[3];
        for (int c = 0; c < dim; ++c) { k1[c] = kf1_val(d[c], P.a, e); col[k * 32 + 1 + c] = k1[c]; }
        if (g >= 0) {
            for (int c = 0; c < dim; ++c) {
                const int row = N + c * ng + g;
                col[row * 32] = -k1[c];
                for (int e2 = 0; e2 < dim; ++e2) {
                    const int c0 = min(c, e2), e0 = max(c, e2);
                    col[row * 32 + 1 + e2] = kf2_val(r, d[c0], d[e0], c == e2 ? 1.f : 0.f, P.a, e);
                }
            }
        }
    }
    __syncthreads();

    // ---- mean: U^T alpha (OnGPIS.cpp:187); lane = column, warps split the rows
    {
        float mu = 0.f;
        for (int i = warp; i < n; i += EVAL2_WARPS) mu = fmaf(U[i * 32 + lane], __ldg(alpha + i), mu);
        red[warp * 32 + lane] = mu;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
            for (int ww = 0; ww < EVAL2_WARPS; ++ww) s += red[ww * 32 + lane];
            const int qi = lane >> 2, c = lane & 3;
            if (qi < cnt && c < w) {
                const int2 pr = S.sorted[first + qi];
                W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + c] = s;
            }
        }
        __syncthreads();
    }

    // ---- block elimination: for column j, warp w updates its rows i > j, i = w (mod 8)
    uint32_t ph = 0;  // parity bits of this warp's two barriers
    for (int j = 0; j + 1 < nb; ++j) {
        const float* Uj = U + j * 32 * 32;
        int i0 = j + 1 + ((warp - (j + 1)) % EVAL2_WARPS + EVAL2_WARPS) % EVAL2_WARPS;  // first row >= j+1 owned by this warp
        int s = 0;
        if (i0 < nb && lane == 0) {
            mbar_expect_tx(&bars[0], GPIS_TILE_BYTES);
            tma_load_1d(stg[0], tiles + (size_t)tile_index(i0, j, nb) * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[0]);
        }
        for (int i = i0; i < nb; i += EVAL2_WARPS, s ^= 1) {
            const int inext = i + EVAL2_WARPS;
            if (inext < nb && lane == 0) {
                mbar_expect_tx(&bars[s ^ 1], GPIS_TILE_BYTES);
                tma_load_1d(stg[s ^ 1], tiles + (size_t)tile_index(inext, j, nb) * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[s ^ 1]);
            }
            mbar_wait(&bars[s], (ph >> s) & 1u);
            ph ^= (1u << s);
            float acc[4][8];
            float* Ui = U + i * 32 * 32;
            ctile_load(Ui, acc, rg, cg);
            tile_mma_sub(acc, stg[s], Uj, rg, cg);
            ctile_store(Ui, acc, rg, cg);
            __syncwarp();   // all lanes done reading stg[s] before it is refilled two iterations later
        }
        __syncthreads();    // U_{j+1} is final before anyone uses it as an operand
    }

    // ---- V_j = Dinv(j) U_j and the column sums of V^2 (OnGPIS.cpp:200-201)
    float ss[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) ss[c] = 0.f;
    {
        int s = 0;
        const int j0 = warp;
        if (j0 < nb && lane == 0) {
            mbar_expect_tx(&bars[0], GPIS_TILE_BYTES);
            tma_load_1d(stg[0], dinv + (size_t)j0 * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[0]);
        }
        for (int j = j0; j < nb; j += EVAL2_WARPS, s ^= 1) {
            const int jn = j + EVAL2_WARPS;
            if (jn < nb && lane == 0) {
                mbar_expect_tx(&bars[s ^ 1], GPIS_TILE_BYTES);
                tma_load_1d(stg[s ^ 1], dinv + (size_t)jn * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[s ^ 1]);
            }
            mbar_wait(&bars[s], (ph >> s) & 1u);
            ph ^= (1u << s);
            float v[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) v[a][b] = 0.f;
            tile_mma_add(v, stg[s], U + j * 32 * 32, rg, cg);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) ss[b] = fmaf(v[a][b], v[a][b], ss[b]);
            __syncwarp();
        }
    }
    // reduce over the 8 row groups of the warp (lane bits 2..4), then over warps
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float t = ss[c];
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        ss[c] = t;
    }
    if (rg == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) red[warp * 32 + 8 * cg + c] = ss[c];
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int ww = 0; ww < EVAL2_WARPS; ++ww) s += red[ww * 32 + lane];
        const int qi = lane >> 2, c = lane & 3;
        if (qi < cnt && c < w) {
            const int2 pr = S.sorted[first + qi];
            const double prior = (c == 0) ? (double)P.prior_f : P.prior_g;   // OnGPIS.cpp:203-212 / 235-237
            W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + w + c] = (float)(prior - (double)s);
        }
    }
}

}  // namespace gpis
