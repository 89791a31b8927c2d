// K4, grouped evaluation — the production kernel.
//
// (query, leaf) pairs are bucketed by leaf; one CTA solves up to QBT queries (4*QBT right-hand
// sides) against one leaf. The record stores G(i,j) = L(i,j) inv(L(j,j)) for off-diagonal 32x32
// tiles and Dinv(j) = inv(L(j,j)); with U = right-hand sides after block elimination,
//      U_i = B_i - sum_{j<i} G(i,j) U_j,   V_j = Dinv(j) U_j,   var = prior - sum V^2
// which is L^{-1} k* (OnGPIS.cpp:199-213) evaluated block-wise.
//
// Mapping (B200): 8 warps per CTA, one CTA per SM. U (npad x 4*QBT floats) stays in shared memory.
// Block rows are dealt to warps modulo 8 and processed in waves of 32 block rows: inside a wave a
// warp keeps its (up to) 4 block rows as a 16x8-per-lane register accumulator, so one k-step costs
// 24 shared-memory words per lane for 128 FMAs — the FMA pipe, not the shared-memory pipe, is the
// limiter (a 4x8 lane tile was measured smem-bound: profiles/r01). G tiles stream L2 -> shared memory as
// 1 KB quarter-tiles through per-warp double-buffered TMA bulk copies (mbarrier completion), issued one
// step ahead and running across column and wave boundaries (G is read-only). A row's tile of U is written
// to shared memory exactly once, when it becomes final. One __syncthreads per block column.
#pragma once
#include <string>

#include "common.cuh"
#include "query.cuh"
#include "query_v2.cuh"

namespace gpis {

#define E3_WARPS 8
#define E3_THREADS (E3_WARPS * 32)
#define E3_R 4                      // block rows per warp per wave
#define E3_WAVE (E3_WARPS * E3_R)   // 32 block rows per wave

struct Eval3Smem {
    static constexpr int off_bar = 0;                                // per warp 2 mbarriers
    static constexpr int off_stage = 256;                            // per warp 2 stages x 4 slots x 1 KB
    static constexpr int off_red = off_stage + E3_WARPS * 2 * 4096;  // E3_WARPS x 32 floats
    static constexpr int off_U = off_red + E3_WARPS * 32 * 4;
    static int total(int nbmax, int ncol) { return off_U + nbmax * 32 * ncol * 4; }
};

// acc[r][i][j] -= sum_{k<8} A_r[k][4rg+i] * B[k][CPL*cg+j] for r in [R0, R0+R);  A_r: quarter tile [8][32] at
// As + r*256, B: [8][NCOL] rows of U. R0 and R are compile-time so the accumulators stay in registers.
template <int R0, int R, int CPL, int NCOL>
__device__ __forceinline__ void qmma_sub(float (&acc)[E3_R][4][CPL], const float* __restrict__ As,
                                         const float* __restrict__ Bq, int rg, int cg) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float bv[CPL];
#pragma unroll
        for (int v = 0; v < CPL / 4; ++v) {
            const float4 b = *reinterpret_cast<const float4*>(Bq + k * NCOL + CPL * cg + 4 * v);
            bv[4 * v] = b.x; bv[4 * v + 1] = b.y; bv[4 * v + 2] = b.z; bv[4 * v + 3] = b.w;
        }
#pragma unroll
        for (int r = R0; r < R0 + R; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(As + r * 256 + k * 32 + 4 * rg);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < CPL; ++j) acc[r][i][j] = fmaf(-av[i], bv[j], acc[r][i][j]);
        }
    }
}
// rows r0 .. R-1 of this warp are active (a suffix): dispatch to the compile-time variant
template <int CPL, int NCOL>
__device__ __forceinline__ void qmma_dispatch(int r0, int R, float (&acc)[E3_R][4][CPL], const float* As, const float* Bq, int rg, int cg) {
    const int code = r0 * 8 + (R - r0);
    switch (code) {
        case 0 * 8 + 4: qmma_sub<0, 4, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 0 * 8 + 3: qmma_sub<0, 3, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 0 * 8 + 2: qmma_sub<0, 2, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 0 * 8 + 1: qmma_sub<0, 1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 1 * 8 + 3: qmma_sub<1, 3, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 1 * 8 + 2: qmma_sub<1, 2, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 1 * 8 + 1: qmma_sub<1, 1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 2 * 8 + 2: qmma_sub<2, 2, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 2 * 8 + 1: qmma_sub<2, 1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        case 3 * 8 + 1: qmma_sub<3, 1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        default: break;
    }
}

template <int QBT>
__global__ void __launch_bounds__(E3_THREADS, 1)
k_eval_v3(const float* __restrict__ x, LeafTable T, QueryParams P, QueryWork W, SortBufs S, const int4* __restrict__ items) {
    constexpr int NCOL = 4 * QBT;      // right-hand sides per CTA
    constexpr int CPL = NCOL / 4;      // columns per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int4 item = items[blockIdx.x];
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

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Eval3Smem::off_bar) + warp * 2;
    float* stg = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_stage) + warp * 2048;   // [2][4][256]
    float* red = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_red);
    float* U = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_U);                       // [npad][NCOL]

    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }

    // ---- right-hand sides: k* of every query of the item (covFnc.cpp:282-311 / 425-446)
    {
        float4* U4 = reinterpret_cast<float4*>(U);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < npad * NCOL / 4; i += E3_THREADS) U4[i] = z4;
    }
    __syncthreads();
    for (int idx = tid; idx < N * QBT; idx += E3_THREADS) {
        const int qi = idx % QBT, k = idx / QBT;
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
        const double e = exp((double)(-P.a * r));
        float* col = U + 4 * qi;
        col[k * NCOL] = kf_val(r, P.a, e);
        float k1[3];
        for (int c = 0; c < dim; ++c) { k1[c] = kf1_val(d[c], P.a, e); col[k * NCOL + 1 + c] = k1[c]; }
        if (g >= 0) {
            for (int c = 0; c < dim; ++c) {
                const int row = N + c * ng + g;
                col[row * NCOL] = -k1[c];
                for (int e2 = 0; e2 < dim; ++e2) {
                    const int c0 = min(c, e2), e0 = max(c, e2);
                    col[row * NCOL + 1 + e2] = kf2_val(r, d[c0], d[e0], c == e2 ? 1.f : 0.f, P.a, e);
                }
            }
        }
    }
    __syncthreads();

    // ---- mean: U^T alpha (OnGPIS.cpp:187); lane = column, warps split the rows
    {
        float mu = 0.f;
        if (lane < NCOL)
            for (int i = warp; i < n; i += E3_WARPS) mu = fmaf(U[i * NCOL + lane], __ldg(alpha + i), mu);
        red[warp * 32 + lane] = mu;
        __syncthreads();
        if (warp == 0 && lane < NCOL) {
            float s = 0.f;
            for (int ww = 0; ww < E3_WARPS; ++ww) s += red[ww * 32 + lane];
            const int qi = lane >> 2, c = lane & 3;
            if (qi < cnt && c < w) {
                const int2 pr = S.sorted[first + qi];
                W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + c] = s;
            }
        }
        __syncthreads();
    }

    // ---- block elimination in waves of 32 block rows
    // Step cursor of this warp: (wave c, column j, quarter q). Active rows at column j: i_r = 32c + warp + 8r
    // with i_r < nb and i_r > j.
    struct Cur { int c, j, q; };
    const int nwaves = (nb + E3_WAVE - 1) / E3_WAVE;
    auto rows_in_wave = [&](int c) { const int lo = E3_WAVE * c + warp; return lo < nb ? min(E3_R, (nb - lo + 7) / 8) : 0; };
    auto last_j = [&](int c) { return E3_WAVE * c + warp + 8 * (rows_in_wave(c) - 1) - 1; };   // last column with an active row
    auto cur_valid = [&](const Cur& s) { return s.c < nwaves; };
    auto normalize = [&](Cur& s) {   // move to the first step at or after s that has work
        while (s.c < nwaves) {
            if (rows_in_wave(s.c) > 0 && s.j <= last_j(s.c)) return;
            ++s.c; s.j = 0; s.q = 0;
        }
    };
    auto advance = [&](Cur s) { if (++s.q == 4) { s.q = 0; ++s.j; } normalize(s); return s; };
    auto issue = [&](const Cur& s, int st) {
        if (lane != 0) return;
        const int base = E3_WAVE * s.c + warp;
        const int R = rows_in_wave(s.c);
        int nact = 0;
        for (int r = 0; r < R; ++r) nact += (base + 8 * r > s.j) ? 1 : 0;
        mbar_expect_tx(&bars[st], (uint32_t)nact * 1024u);
        for (int r = 0; r < R; ++r) {
            const int i = base + 8 * r;
            if (i > s.j)
                tma_load_1d(stg + (st * 4 + r) * 256, tiles + (size_t)tile_index(i, s.j, nb) * GPIS_TILE_ELEMS + s.q * 256, 1024u, &bars[st]);
        }
    };

    float acc[E3_R][4][CPL];
    uint32_t ph = 0;
    int st = 0;
    Cur cur{0, 0, 0};
    normalize(cur);
    if (cur_valid(cur)) issue(cur, 0);

#define E3_STEP_BEGIN                                                                             \
    const Cur nxt = advance(cur);                                                                  \
    if (cur_valid(nxt)) issue(nxt, st ^ 1);                                                        \
    mbar_wait(&bars[st], (ph >> st) & 1u);                                                         \
    ph ^= (1u << st);                                                                              \
    const float* Bq = U + (size_t)(cur.j * 32 + cur.q * 8) * NCOL;                                 \
    const float* As = stg + st * 1024;
#define E3_STEP_END                                                                                \
    __syncwarp();                                                                                  \
    cur = nxt;                                                                                     \
    st ^= 1;

    for (int c = 0; c < nwaves; ++c) {
        const int wbase = E3_WAVE * c;
        const int base = wbase + warp;
        const int R = rows_in_wave(c);
        // accumulators <- right-hand sides of this warp's rows
#pragma unroll
        for (int r = 0; r < E3_R; ++r) {
            if (r < R) {
                const float* Ui = U + (size_t)(base + 8 * r) * 32 * NCOL;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int v = 0; v < CPL / 4; ++v) {
                        const float4 t = *reinterpret_cast<const float4*>(Ui + (4 * rg + i) * NCOL + CPL * cg + 4 * v);
                        acc[r][i][4 * v] = t.x; acc[r][i][4 * v + 1] = t.y; acc[r][i][4 * v + 2] = t.z; acc[r][i][4 * v + 3] = t.w;
                    }
            }
        }
        // phase 1: columns of earlier waves, all rows active, no block-level synchronisation
        if (R > 0) {
            for (int j = 0; j < wbase; ++j)
                for (int q = 0; q < 4; ++q) {
                    E3_STEP_BEGIN
                    qmma_dispatch<CPL, NCOL>(0, R, acc, As, Bq, rg, cg);
                    E3_STEP_END
                }
        }
        // phase 2: the wave's own columns; the owner of row j publishes U_j, then later rows consume it
        const int tcount = min(E3_WAVE, nb - wbase);
        for (int t = 0; t < tcount; ++t) {
            const int j = wbase + t;
            if ((t & 7) == warp) {
                const int ro = t >> 3;   // which of this warp's rows is row j
                float* Uj = U + (size_t)j * 32 * NCOL;
#pragma unroll
                for (int r = 0; r < E3_R; ++r) {
                    if (r == ro) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int v = 0; v < CPL / 4; ++v)
                                *reinterpret_cast<float4*>(Uj + (4 * rg + i) * NCOL + CPL * cg + 4 * v) =
                                    make_float4(acc[r][i][4 * v], acc[r][i][4 * v + 1], acc[r][i][4 * v + 2], acc[r][i][4 * v + 3]);
                    }
                }
            }
            __syncthreads();   // U_j is final and visible
            // rows of this warp below j: r >= r0
            int r0 = 0;
            while (r0 < R && base + 8 * r0 <= j) ++r0;
            const int nact = R - r0;
            if (nact > 0) {
                for (int q = 0; q < 4; ++q) {
                    E3_STEP_BEGIN
                    qmma_dispatch<CPL, NCOL>(r0, R, acc, As, Bq, rg, cg);   // TMA filled slots r0..R-1
                    E3_STEP_END
                }
            }
        }
    }
#undef E3_STEP_BEGIN
#undef E3_STEP_END
    __syncthreads();

    // ---- V_j = Dinv(j) U_j and the column sums of V^2 (OnGPIS.cpp:200-201); warp w takes j = w, w+8, ...
    float ss[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) ss[c] = 0.f;
    {
        auto issue_d = [&](int j, int s2) {
            if (lane == 0) {
                mbar_expect_tx(&bars[s2], GPIS_TILE_BYTES);
                tma_load_1d(stg + s2 * 1024, dinv + (size_t)j * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[s2]);
            }
        };
        if (warp < nb) issue_d(warp, st);
        for (int j = warp; j < nb; j += E3_WARPS) {
            if (j + E3_WARPS < nb) issue_d(j + E3_WARPS, st ^ 1);
            mbar_wait(&bars[st], (ph >> st) & 1u);
            ph ^= (1u << st);
            float v[E3_R][4][CPL];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < CPL; ++jj) v[0][i][jj] = 0.f;
            const float* Uj = U + (size_t)j * 32 * NCOL;
            // a full tile is four consecutive quarter tiles: slot q of this stage
#pragma unroll
            for (int q = 0; q < 4; ++q) qmma_sub<0, 1, CPL, NCOL>(v, stg + st * 1024 + q * 1024 / 4, Uj + q * 8 * NCOL, rg, cg);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < CPL; ++jj) ss[jj] = fmaf(v[0][i][jj], v[0][i][jj], ss[jj]);   // (-V)^2 = V^2
            __syncwarp();
            st ^= 1;
        }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        float t = ss[c];
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        ss[c] = t;
    }
    if (rg == 0) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) red[warp * 32 + CPL * cg + c] = ss[c];
    }
    __syncthreads();
    if (warp == 0 && lane < NCOL) {
        float s = 0.f;
        for (int ww = 0; ww < E3_WARPS; ++ww) s += red[ww * 32 + lane];
        const int qi = lane >> 2, c = lane & 3;
        if (qi < cnt && c < w) {
            const int2 pr = S.sorted[first + qi];
            const double prior = (c == 0) ? (double)P.prior_f : P.prior_g;   // OnGPIS.cpp:203-212 / 235-237
            W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + w + c] = (float)(prior - (double)s);
        }
    }
}

// ------------------------------------------------------------------ host side
#define E3_NB_A 40   // 8 queries per CTA: U = nb * 4 KB  (n <= 1280)
#define E3_NB_B 80   // 4 queries per CTA: U = nb * 2 KB  (n <= 2560); larger leaves go to k_eval_v1

static inline int query_eval_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_eval_v3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval_v3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_eval_*): ") + cudaGetErrorString(e); return -2; }
    return 0;
}

// Buckets the npairs work items in W.pairs by leaf and evaluates them. *d_sort is a grow-only device
// buffer owned by the context. version: 1 = one CTA per pair, 2 = first grouped kernel (4x8 lane tile),
// 3 = production kernel.
static inline int query_eval(cudaStream_t st, const float* d_x, const LeafTable& T, const QueryParams& P,
                             const QueryWork& W, int npairs, int nslots, int max_nb, int32_t** d_sort,
                             int64_t* sort_cap, int64_t* launches, std::string& err, double* d_acc, int version) {
#define CK2(call)                                                                 \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return -2; } \
    } while (0)
    const int64_t max_items = npairs / 4 + nslots + 8;
    const int64_t need = (int64_t)nslots * 4 + 64 + (int64_t)npairs * 4 + max_items * 8 + 64;
    if (*sort_cap < need) {
        if (*d_sort) CK2(cudaFree(*d_sort));
        *d_sort = nullptr; *sort_cap = 0;
        CK2(cudaMalloc(d_sort, sizeof(int32_t) * (need + need / 4)));
        *sort_cap = need + need / 4;
    }
    SortBufs S;
    int32_t* p = *d_sort;
    S.totals = p; p += 16;
    S.count = p; p += nslots;
    S.start = p; p += nslots + 1;
    S.istart = p; p += nslots + 1;
    S.cursor = p; p += nslots;
    p += (4 - (((uintptr_t)p >> 2) & 3)) & 3;   // 16-byte align for int4
    S.items = reinterpret_cast<int4*>(p); p += max_items * 4;
    S.itemsB = reinterpret_cast<int4*>(p); p += max_items * 4;
    S.sorted = reinterpret_cast<int2*>(p); p += (int64_t)npairs * 2;
    S.pairsC = reinterpret_cast<int2*>(p);
    CK2(cudaMemsetAsync(S.totals, 0, sizeof(int32_t) * 16, st));
    CK2(cudaMemsetAsync(S.count, 0, sizeof(int32_t) * nslots, st));
    k_pair_hist<<<(npairs + 255) / 256, 256, 0, st>>>(W.pairs, npairs, S.count);
    k_slot_scan<<<1, 1024, 0, st>>>(S, nslots);
    k_pair_scatter<<<(npairs + 255) / 256, 256, 0, st>>>(W.pairs, npairs, S);
    *launches += 3;
    if (d_acc) { k_query_stats<<<(nslots + 255) / 256, 256, 0, st>>>(S, T, nslots, d_acc); *launches += 1; }
    const int v1_smem = (1 + P.dim) * max_nb * 32 * (int)sizeof(float);
    if (version == 1) {
        k_eval_v1<<<npairs, EVAL1_THREADS, v1_smem, st>>>(d_x, T, P, W, S.sorted);
        *launches += 1;
        CK2(cudaGetLastError());
        return 0;
    }
    if (version == 2 && Eval2Smem::total(max_nb) <= 227 * 1024) {
        k_make_items<<<(nslots + 255) / 256, 256, 0, st>>>(S, nslots);
        *launches += 1;
        int32_t nitems = 0;
        CK2(cudaMemcpyAsync(&nitems, S.totals, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK2(cudaStreamSynchronize(st));
        if (nitems > 0) {
            k_eval_v2<<<nitems, EVAL2_THREADS, Eval2Smem::total(max_nb), st>>>(d_x, T, P, W, S);
            *launches += 1;
            CK2(cudaGetLastError());
        }
        return 0;
    }
    k_make_items_classed<<<(nslots + 255) / 256, 256, 0, st>>>(S, T, nslots, E3_NB_A, E3_NB_B);
    *launches += 1;
    int32_t tot[4] = {0, 0, 0, 0};
    CK2(cudaMemcpyAsync(tot, S.totals, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
    CK2(cudaStreamSynchronize(st));
    if (tot[1] > 0) {
        const int nbm = max_nb < E3_NB_A ? max_nb : E3_NB_A;
        k_eval_v3<8><<<tot[1], E3_THREADS, Eval3Smem::total(nbm, 32), st>>>(d_x, T, P, W, S, S.items);
        *launches += 1;
    }
    if (tot[2] > 0) {
        const int nbm = max_nb < E3_NB_B ? max_nb : E3_NB_B;
        k_eval_v3<4><<<tot[2], E3_THREADS, Eval3Smem::total(nbm, 16), st>>>(d_x, T, P, W, S, S.itemsB);
        *launches += 1;
    }
    if (tot[3] > 0) {
        k_eval_v1<<<tot[3], EVAL1_THREADS, v1_smem, st>>>(d_x, T, P, W, S.pairsC);
        *launches += 1;
    }
    CK2(cudaGetLastError());
#undef CK2
    return 0;
}

}  // namespace gpis
