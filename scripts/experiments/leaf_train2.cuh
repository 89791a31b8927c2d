// K1 — batched leaf-GP training, second design (production): covFnc (train) + OnGPIS::train as one persistent
// kernel. Replaces cpp/src/covFnc.cpp:142-256 / 317-402 and cpp/src/OnGPIS.cpp:34-149.
//
// What changed against the first design (leaf_train.cuh, kept for A/B runs behind gpis_set_train_version):
//   * 64-column panels instead of 32-column block columns, and a 64x64 output tile per warp (8x16 accumulators
//     per lane): one k-step needs 24 shared-memory words per 128 FMAs instead of 12 per 32, and every staged A tile
//     is used against two block columns, which halves the re-read traffic of the left-looking update.
//   * 4 warps per CTA, 2 CTAs per SM, persistent CTAs that pull leaves from a counter: a CTA in the narrow tail of
//     its factorization shares the SM with one that is in the wide middle of another leaf.
//   * a 4-stage TMA (cp.async.bulk) ring of 16-k half tiles with full/empty mbarriers: three stages in flight
//     cover the HBM latency that the one-stage look-ahead of the first design exposed (30 % of warp time).
//   * the factor is written twice: L into a per-CTA workspace (operand of the later panels and of the backward
//     solve) and the query form G(i,j) = L(i,j) inv(L(j,j)) straight into the record while the tile is still on
//     chip, so the separate conversion sweep (phase E, 10 % of the first design) is gone.
//   * z and alpha come from proper triangular substitutions with the 32x32 diagonal factors (like the reference's
//     solves, OnGPIS.cpp:139-143), not from products with their inverses. Off-diagonal tiles are still finished as
//     products with D = inv(L(j,j)) (L = Y D^T, then G = L D): a row-per-lane substitution variant was
//     built and measured — same parity statistics, but its unrolled code thrashed the instruction cache
//     (profiles/r02_history.md).
//
// Per leaf (N samples, ng with a usable normal, n = N + dim*ng unknowns, nb = ceil(n/32)):
//   A. samples -> shared memory, gradflag rule (OnGPIS.cpp:63-66, 122-125), y
//   B. K (lower triangle) straight into the workspace tiles — one exp per ordered sample pair, entries rounded to
//      float exactly like the reference (covFnc.cpp:29-33)
//   C. for every panel p (block columns 2p, 2p+1): left-looking update of the panel's tiles in passes of 4 warp
//      tiles, then the panel's 64x64 diagonal block (two in-warp 32x32 Choleskys), then the triangular solves of
//      the rows below; the forward solve L z = y rides along on the staged tiles of the panel's own rows
//   D. backward solve L^T alpha = z by block columns
// Bound: FP32 FMA pipe (n^3/3 flops vs ~2n^2 compulsory bytes). No tensor cores: independent factorizations in
// fp32 with a 1e-4 parity contract.
#pragma once
#include <string>
#include <type_traits>

#include "common.cuh"
#include "leaf_train.cuh"

namespace gpis {

#define T2_WARPS 4
#define T2_THREADS (T2_WARPS * 32)
#define T2_STAGES 4
#define T2_HALF_FLOATS 512                      // half tile: [16 k][32 rows]
#define T2_HALF_BYTES 2048
#define T2_STAGE_HALVES (2 * T2_WARPS + 2)      // A: two block rows per warp, B: the panel's two block rows
#define T2_STAGE_FLOATS (T2_STAGE_HALVES * T2_HALF_FLOATS)

struct Train2Smem {
    static constexpr int off_stage = 0;                                        // T2_STAGES x 20 KB; aliased by the per-warp
    static constexpr int stage_bytes = T2_STAGE_FLOATS * 4;                    // 16 KB scratch and by the samples of phase A/B
    static constexpr int off_diag = T2_STAGES * stage_bytes;                   // D0, D0t, L10, D1, D1t
    static constexpr int off_bar = off_diag + 5 * GPIS_TILE_BYTES;             // full[4], empty[4]
    static constexpr int off_misc = off_bar + 64;                              // job index, wtot[64], part[4][32], z[npad]
    static int total(int nbmax) { return off_misc + 16 + 64 * 4 + T2_WARPS * 32 * 4 + nbmax * 32 * 4; }
};
static_assert(T2_WARPS * 4 * GPIS_TILE_BYTES <= T2_STAGES * Train2Smem::stage_bytes, "scratch must fit the stage area");

// acc[rt][ct][i][j] -= sum_{k<16} A_rt[k][4rg+i] * B_ct[k][8cg+j]; A_*, B_*: half tiles [16][32] in shared memory.
// Packed FFMA2 along j, operands of step k+1 loaded before the FMAs of step k issue.
__device__ __forceinline__ void mma64_half(float (&acc)[2][2][4][8], const float* __restrict__ A0, const float* __restrict__ A1,
                                           const float* __restrict__ B0, const float* __restrict__ B1, int rg, int cg) {
    const float* ap[2] = {A0 + 4 * rg, A1 + 4 * rg};
    const float* bp[2] = {B0 + 8 * cg, B1 + 8 * cg};
    float4 a[2][2], b[2][2][2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        a[0][t] = *reinterpret_cast<const float4*>(ap[t]);
        b[0][t][0] = *reinterpret_cast<const float4*>(bp[t]);
        b[0][t][1] = *reinterpret_cast<const float4*>(bp[t] + 4);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int cu = k & 1, nx = cu ^ 1;
        if (k + 1 < 16) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                a[nx][t] = *reinterpret_cast<const float4*>(ap[t] + (k + 1) * 32);
                b[nx][t][0] = *reinterpret_cast<const float4*>(bp[t] + (k + 1) * 32);
                b[nx][t][1] = *reinterpret_cast<const float4*>(bp[t] + (k + 1) * 32 + 4);
            }
        }
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const float av[4] = {a[cu][rt].x, a[cu][rt].y, a[cu][rt].z, a[cu][rt].w};
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const float bv[8] = {b[cu][ct][0].x, b[cu][ct][0].y, b[cu][ct][0].z, b[cu][ct][0].w,
                                     b[cu][ct][1].x, b[cu][ct][1].y, b[cu][ct][1].z, b[cu][ct][1].w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; j += 2) fma2_sub(acc[rt][ct][i][j], acc[rt][ct][i][j + 1], bv[j], bv[j + 1], av[i]);
            }
        }
    }
}

// Tiles in shared memory are k-major: element (r, c) at c*32 + r.
// Solves L z = t for one 32-vector spread over the lanes (lane = row). L k-major in shared memory.
__device__ __forceinline__ float warp_solve_fwd(float t, const float* __restrict__ L, int lane) {
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float v = __shfl_sync(0xffffffffu, t, c) / L[c * 32 + c];
        if (lane == c) z = v;
        if (lane > c) t = fmaf(-v, L[c * 32 + lane], t);
    }
    return z;
}
// Solves L^T a = t (lane = row of a). L k-major in shared memory: lane i reads L(c, i) at i*32 + c (a strided,
// bank-conflicting read; 32 steps per block row of the backward solve, negligible).
__device__ __forceinline__ float warp_solve_bwd_t(float t, const float* __restrict__ L, int lane) {
    float a = 0.f;
#pragma unroll
    for (int c = 31; c >= 0; --c) {
        const float v = __shfl_sync(0xffffffffu, t, c) / L[c * 32 + c];
        if (lane == c) a = v;
        if (lane < c) t = fmaf(-v, L[lane * 32 + c], t);
    }
    return a;
}
// coalesced copy of one 4 KB tile by a warp (any of global / shared)
__device__ __forceinline__ void warp_copy_tile(float* __restrict__ dst, const float* __restrict__ src, int lane) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[lane + 32 * i] = s4[lane + 32 * i];
}
// dst (row-major image) <- src (k-major), both in shared memory
__device__ __forceinline__ void warp_transpose_tile(float* __restrict__ dst, const float* __restrict__ src, int lane) {
#pragma unroll
    for (int c = 0; c < 32; ++c) dst[lane * 32 + c] = src[c * 32 + lane];
}
#ifdef T2_TIMING
__device__ long long g_t2_timing[16];   // clock64 ticks of CTA 0, warp 0, by phase: A+B, pass prologue, k-loop, drain+scratch, diagonal, rows, D
#define T2_T(i) { if (blockIdx.x == 0 && tid == 0) { const long long t_ = clock64(); g_t2_timing[i] += t_ - t2_t0; t2_t0 = t_; } }
#else
#define T2_T(i)
#endif

__global__ void __launch_bounds__(T2_THREADS, 2)
k_leaf_train2(const TrainJob* __restrict__ jobs, int njobs, const float* __restrict__ samples, TrainParams P,
              int32_t* __restrict__ status, float* __restrict__ ws_base, uint64_t ws_stride, int32_t* __restrict__ counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = lane >> 2, cg = lane & 3;
    const int dim = P.dim, w9 = 2 * dim + 3;

    float* stage_f = reinterpret_cast<float*>(smem_raw + Train2Smem::off_stage);
    float* scratch = stage_f + warp * 4 * GPIS_TILE_ELEMS;                         // 4 tiles per warp: Y00 Y01 Y10 Y11
    float* sD0 = reinterpret_cast<float*>(smem_raw + Train2Smem::off_diag);   // inv(L00) k-major; its row-major image; L10; inv(L11); image
    float* sD0t = sD0 + GPIS_TILE_ELEMS;
    float* sL10 = sD0 + 2 * GPIS_TILE_ELEMS;
    float* sD1 = sD0 + 3 * GPIS_TILE_ELEMS;
    float* sD1t = sD0 + 4 * GPIS_TILE_ELEMS;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Train2Smem::off_bar);
    uint64_t* empty = full + T2_STAGES;
    int* s_job = reinterpret_cast<int*>(smem_raw + Train2Smem::off_misc);
    int* wtot = s_job + 4;
    float* part = reinterpret_cast<float*>(wtot + 64);
    float* zv = part + T2_WARPS * 32;
    float* ws = ws_base + (size_t)blockIdx.x * ws_stride;

    if (tid == 0) {
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], T2_WARPS); }
        fence_mbar_init();
    }
    uint32_t gstep = 0;   // k-steps this CTA has pushed through the ring so far (stage = gstep % S, round = gstep / S)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_job[0] = atomicAdd(counter, 1);
        __syncthreads();
        const int ji = s_job[0];
        if (ji >= njobs) break;
        const TrainJob job = jobs[ji];
        const int N = job.N, ng = job.ng, n = job.n, nb = job.nb;
#ifdef T2_TIMING
        long long t2_t0 = clock64();
#endif
        unsigned char* rec = reinterpret_cast<unsigned char*>(job.rec);
        float4* rec_pts = reinterpret_cast<float4*>(rec + rec_off_pts());
        float* rec_alpha = reinterpret_cast<float*>(rec + rec_off_alpha(N));
        float* rec_dinv = reinterpret_cast<float*>(rec + rec_off_dinv(N, nb));
        float* rec_tiles = reinterpret_cast<float*>(rec + rec_off_tiles(N, nb));
        const int ntiles = nb * (nb + 1) / 2;
        // phase A/B view of the stage area
        float4* pts = reinterpret_cast<float4*>(smem_raw + Train2Smem::off_stage);
        float* sigx = reinterpret_cast<float*>(pts + N);
        float* sigg = sigx + N;

        // ---------------------------------------------------------------- A. samples, gradflag, y (into zv)
        const float* smp = samples + (size_t)job.sample_off * w9;
        for (int base = warp * 32; base < N; base += T2_WARPS * 32) {
            const int k = base + lane;
            bool valid = false;
            if (k < N) {
                const float* s = smp + (size_t)k * w9;
                const float sg = s[2 * dim + 2];
                bool allsmall = true;
                for (int c = 0; c < dim; ++c) allsmall = allsmall && (fabs((double)s[dim + c]) < 1e-6);
                valid = !((double)sg > 0.1001 || allsmall);
            }
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) wtot[base >> 5] = __popc(m);
        }
        for (int i = tid; i < nb * 32; i += T2_THREADS) zv[i] = 0.f;
        __syncthreads();
        for (int base = warp * 32; base < N; base += T2_WARPS * 32) {
            const int k = base + lane;
            int before = 0;
            for (int c = 0; c < (base >> 5); ++c) before += wtot[c];
            bool valid = false;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            float sx = 0.f, sg = 0.f, f = 0.f, g[3] = {0.f, 0.f, 0.f};
            if (k < N) {
                const float* s = smp + (size_t)k * w9;
                p.x = s[0]; p.y = s[1]; p.z = (dim == 3) ? s[2] : 0.f;
                for (int c = 0; c < dim; ++c) g[c] = s[dim + c];
                f = s[2 * dim]; sx = s[2 * dim + 1]; sg = s[2 * dim + 2];
                bool allsmall = true;
                for (int c = 0; c < dim; ++c) allsmall = allsmall && (fabs((double)g[c]) < 1e-6);
                valid = !((double)sg > 0.1001 || allsmall);
            }
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (k < N) {
                int gi = -1;
                if (valid) gi = before + __popc(m & ((1u << lane) - 1u));
                else sx = 2.0f;  // OnGPIS.cpp:65,124
                p.w = __int_as_float(gi);
                pts[k] = p; sigx[k] = sx; sigg[k] = sg;
                rec_pts[k] = p;
                zv[k] = f;
                if (gi >= 0)
                    for (int c = 0; c < dim; ++c) zv[N + c * ng + gi] = g[c];
            }
        }
        {   // zero the workspace tiles (padding rows/cols, structural zeros, strict upper parts of diagonal tiles)
            float4* t4 = reinterpret_cast<float4*>(ws);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < ntiles * (GPIS_TILE_ELEMS / 4); i += T2_THREADS) t4[i] = z4;
        }
        __syncthreads();

        // ---------------------------------------------------------------- B. covariance -> workspace tiles
#define KSTORE(row, col, v) ws[(size_t)tile_index((row) >> 5, (col) >> 5, nb) * GPIS_TILE_ELEMS + ((col) & 31) * 32 + ((row) & 31)] = (v)
        auto build_cov = [&](auto dimc) {
            constexpr int DIM = decltype(dimc)::value;
            const int Npad = (N + 31) & ~31;
            const int parts = max(1, T2_THREADS / Npad);
            const int a = (Npad >= T2_THREADS) ? tid : tid % Npad;
            const int prt = (Npad >= T2_THREADS) ? 0 : tid / Npad;
            const int astride = (Npad >= T2_THREADS) ? T2_THREADS : (1 << 30);
            for (int a0 = a; a0 < N && prt < parts; a0 += astride) {
                const float4 pa = pts[a0];
                const int ga = __float_as_int(pa.w);
                const float xa[3] = {pa.x, pa.y, pa.z};
#pragma unroll 2
                for (int b = prt; b < N; b += parts) {
                    if (b == a0) {
                        KSTORE(a0, a0, (float)(1.0 + (double)sigx[a0]));  // covFnc.cpp:173
                        if (ga >= 0) {
#pragma unroll
                            for (int c = 0; c < DIM; ++c) {
                                const int rc = N + c * ng + ga;
                                float v = P.a2 + sigg[a0];  // :182-190 / :355
                                if (DIM == 2 && c == 0) v = (float)((double)P.a2 + sqrt((double)(sigx[a0] * sigg[a0])));  // :352
                                KSTORE(rc, rc, v);
                            }
                        }
                        continue;
                    }
                    const float4 pb = pts[b];
                    const int gb = __float_as_int(pb.w);
                    if (ga < 0 && a0 < b) continue;  // nothing of this ordered pair lies in the lower triangle
                    const float xb[3] = {pb.x, pb.y, pb.z};
                    const bool fwd = a0 < b;   // differences oriented like the reference: first index = smaller sample index
                    float d[DIM], s2 = 0.f;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        d[c] = fwd ? (xa[c] - xb[c]) : (xb[c] - xa[c]);
                        s2 = (c == 0) ? d[c] * d[c] : s2 + d[c] * d[c];
                    }
                    const float r = sqrtf(s2);
                    const DF e = exp_df(-P.a * r);
                    if (a0 > b) KSTORE(a0, b, kf_val(r, P.a, e));
                    if (ga >= 0) {
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            const float k1 = kf1_val(d[c], P.a, e);   // covFnc.cpp:198-203 / 239-249
                            KSTORE(N + c * ng + ga, b, fwd ? -k1 : k1);
                        }
                        if (gb >= 0) {
                            float k2[DIM][DIM];
#pragma unroll
                            for (int c = 0; c < DIM; ++c)
#pragma unroll
                                for (int e2 = c; e2 < DIM; ++e2) k2[c][e2] = k2[e2][c] = kf2_val(r, d[c], d[e2], c == e2 ? 1.f : 0.f, P.a, e);
#pragma unroll
                            for (int c = 0; c < DIM; ++c)
#pragma unroll
                                for (int e2 = 0; e2 < DIM; ++e2) {
                                    const int row = N + c * ng + ga, col = N + e2 * ng + gb;
                                    if (row > col) KSTORE(row, col, k2[c][e2]);
                                }
                        }
                    }
                }
            }
            for (int i = n + tid; i < nb * 32; i += T2_THREADS) KSTORE(i, i, 1.0f);   // identity on the padded diagonal
        };
        if (dim == 3) build_cov(std::integral_constant<int, 3>{});
        else build_cov(std::integral_constant<int, 2>{});
#undef KSTORE

        T2_T(0)
        // ---------------------------------------------------------------- C. panels
        int bad_total = 0;
        const int npanels = (nb + 1) / 2;
        for (int p = 0; p < npanels; ++p) {
            const int c0 = 2 * p, c1 = c0 + 1;
            const bool has1 = c1 < nb;
            const int wtiles = (nb - c0 + 1) / 2;          // warp tiles (two block rows each) of this panel
            const int npass = (wtiles + T2_WARPS - 1) / T2_WARPS;
            const int H = 2 * c0;                           // half-k steps: block columns 0 .. c0-1
            for (int q = 0; q < npass; ++q) {
                // the workspace tiles written so far (phase B, earlier panels) must be visible to the TMA reads of this
                // pass, and nobody may still be using the stage area (scratch of the previous pass, samples of phase B)
                __threadfence_block();
                fence_proxy_async();
                __syncthreads();
                const int t = T2_WARPS * q + warp;
                const int bi0 = c0 + 2 * t, bi1 = bi0 + 1;
                const bool v0 = bi0 < nb, v1 = bi1 < nb;
                auto produce = [&](int h) {   // thread 0: half-k step h of this pass into the ring
                    const uint32_t g = gstep + (uint32_t)h;
                    const int s = (int)(g % T2_STAGES);
                    const uint32_t round = g / T2_STAGES;
                    if (round > 0) mbar_wait(&empty[s], (round - 1u) & 1u);
                    const int bk = h >> 1, hh = h & 1;
                    float* st = stage_f + s * T2_STAGE_FLOATS;
                    int ncopies = has1 ? 2 : 1;
                    for (int wq = 0; wq < 2 * T2_WARPS; ++wq) ncopies += (c0 + 2 * T2_WARPS * q + wq < nb) ? 1 : 0;
                    mbar_expect_tx(&full[s], (uint32_t)ncopies * T2_HALF_BYTES);
                    for (int wq = 0; wq < 2 * T2_WARPS; ++wq) {
                        const int bi = c0 + 2 * T2_WARPS * q + wq;
                        if (bi < nb)
                            tma_load_1d(st + wq * T2_HALF_FLOATS, ws + (size_t)tile_index(bi, bk, nb) * GPIS_TILE_ELEMS + hh * T2_HALF_FLOATS,
                                        T2_HALF_BYTES, &full[s]);
                    }
                    tma_load_1d(st + (2 * T2_WARPS) * T2_HALF_FLOATS, ws + (size_t)tile_index(c0, bk, nb) * GPIS_TILE_ELEMS + hh * T2_HALF_FLOATS,
                                T2_HALF_BYTES, &full[s]);
                    if (has1)
                        tma_load_1d(st + (2 * T2_WARPS + 1) * T2_HALF_FLOATS, ws + (size_t)tile_index(c1, bk, nb) * GPIS_TILE_ELEMS + hh * T2_HALF_FLOATS,
                                    T2_HALF_BYTES, &full[s]);
                };
                if (tid == 0)
                    for (int h = 0; h < H && h < T2_STAGES - 1; ++h) produce(h);

                // accumulators <- K tiles of this warp's 2x2 blocks (lower triangle only)
                float acc[2][2][4][8];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) {
                        const int bi = bi0 + rt, bc = c0 + ct;
                        if (bi < nb && bc < nb && bi >= bc) tile_load(ws + (size_t)tile_index(bi, bc, nb) * GPIS_TILE_ELEMS, acc[rt][ct], rg, cg);
                        else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 8; ++j) acc[rt][ct][i][j] = 0.f;
                        }
                    }
                // forward-solve rider (pass 0, last warp): t_c0, t_c1 -= L(c, bk) z_bk, lane = row
                const bool rider = (q == 0 && warp == T2_WARPS - 1);
                float tz0 = 0.f, tz1 = 0.f;
                if (rider) { tz0 = zv[c0 * 32 + lane]; tz1 = has1 ? zv[c1 * 32 + lane] : 0.f; }
                T2_T(1)

                for (int h = 0; h < H; ++h) {
                    if (tid == 0 && h + T2_STAGES - 1 < H) produce(h + T2_STAGES - 1);
                    const uint32_t g = gstep + (uint32_t)h;
                    const int s = (int)(g % T2_STAGES);
                    mbar_wait(&full[s], (g / T2_STAGES) & 1u);
                    const float* st = stage_f + s * T2_STAGE_FLOATS;
                    const float* B0 = st + (2 * T2_WARPS) * T2_HALF_FLOATS;
                    const float* B1 = B0 + T2_HALF_FLOATS;
                    if (v0) mma64_half(acc, st + (2 * warp) * T2_HALF_FLOATS, st + (2 * warp + 1) * T2_HALF_FLOATS, B0, B1, rg, cg);
                    if (rider) {
                        const float* zk = zv + (h >> 1) * 32 + (h & 1) * 16;
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            s0 = fmaf(B0[k * 32 + lane], zk[k], s0);
                            s1 = fmaf(B1[k * 32 + lane], zk[k], s1);
                        }
                        tz0 -= s0;
                        if (has1) tz1 -= s1;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                gstep += (uint32_t)H;
                T2_T(2)
                if (rider) { zv[c0 * 32 + lane] = tz0; if (has1) zv[c1 * 32 + lane] = tz1; }
                __syncthreads();   // every warp is done with the stages: the area becomes per-warp scratch
                if (v0) {
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct) tile_store(scratch + (rt * 2 + ct) * GPIS_TILE_ELEMS, acc[rt][ct], rg, cg);
                }
                __syncwarp();
                T2_T(3)

                if (q == 0) {
                    if (warp == 0) {
                        // ---- diagonal 64x64 block [L00 0; L10 L11]; scratch tiles: Y00, (free), Y10, Y11
                        float* Y10 = scratch + 2 * GPIS_TILE_ELEMS;
                        float z0 = 0.f;
#pragma unroll 1
                        for (int dd = 0; dd < (has1 ? 2 : 1); ++dd) {
                            float* Yd = scratch + (dd ? 3 : 0) * GPIS_TILE_ELEMS;
                            float* sD = dd ? sD1 : sD0;
                            float* sDt = dd ? sD1t : sD0t;
                            const int cd = c0 + dd;
                            if (dd) {   // T11 = Y11 - L10 L10^T
                                float tt[4][8];
                                tile_load(Yd, tt, rg, cg);
                                tile_mma_sub(tt, sL10, sL10, rg, cg);
                                __syncwarp();
                                tile_store(Yd, tt, rg, cg);
                                __syncwarp();
                            }
                            bad_total += warp_chol_inv(Yd, sD, lane);
                            __syncwarp();
                            warp_copy_tile(rec_tiles + (size_t)tile_index(cd, cd, nb) * GPIS_TILE_ELEMS, Yd, lane);
                            warp_copy_tile(rec_dinv + (size_t)cd * GPIS_TILE_ELEMS, sD, lane);
                            warp_transpose_tile(sDt, sD, lane);
                            // forward solve of this block row: z0 = L00^-1 t0;  z1 = L11^-1 (t1 - L10 z0)
                            float tv = zv[cd * 32 + lane];
                            if (dd) {
#pragma unroll 8
                                for (int k = 0; k < 32; ++k) tv = fmaf(-sL10[k * 32 + lane], __shfl_sync(0xffffffffu, z0, k), tv);
                            }
                            const float zz = warp_solve_fwd(tv, Yd, lane);
                            zv[cd * 32 + lane] = zz;
                            if (!dd) z0 = zz;
                            __syncwarp();
                            if (!dd && has1) {   // L10 = Y10 D0^T (workspace + shared memory), G10 = L10 D0 (record)
                                float l[4][8], g[4][8];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { l[i][j] = 0.f; g[i][j] = 0.f; }
                                tile_mma_add(l, Y10, sD0, rg, cg);
                                tile_store(sL10, l, rg, cg);
                                tile_store(ws + (size_t)tile_index(c1, c0, nb) * GPIS_TILE_ELEMS, l, rg, cg);
                                __syncwarp();
                                tile_mma_add(g, sL10, sD0t, rg, cg);
                                tile_store(rec_tiles + (size_t)tile_index(c1, c0, nb) * GPIS_TILE_ELEMS, g, rg, cg);
                            }
                        }
                    }
                    __syncthreads();   // D0, D0t, L10, D1, D1t visible to every warp (they stay valid for the later passes)
                }
                T2_T(4)
                // ---- rows below the diagonal block: L = Y D^T into the workspace, G = Y M into the record
                if (v0) {
#pragma unroll 1
                    for (int rt = 0; rt < 2; ++rt) {
                        const int bi = bi0 + rt;
                        if (bi >= nb || bi <= c1) continue;            // the diagonal block's rows were done by warp 0
#pragma unroll 1
                        for (int ct = 0; ct < (has1 ? 2 : 1); ++ct) {
                            float* Y = scratch + (rt * 2 + ct) * GPIS_TILE_ELEMS;
                            const float* sD = ct ? sD1 : sD0;
                            const float* sDt = ct ? sD1t : sD0t;
                            const int bc = c0 + ct;
                            if (ct) {   // Y1 -= L(bi,c0) L10^T; L(bi,c0) sits in the first column's scratch tile
                                float tt[4][8];
                                tile_load(Y, tt, rg, cg);
                                tile_mma_sub(tt, scratch + (rt * 2) * GPIS_TILE_ELEMS, sL10, rg, cg);
                                __syncwarp();
                                tile_store(Y, tt, rg, cg);
                                __syncwarp();
                            }
                            float l[4][8], g[4][8];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 8; ++j) { l[i][j] = 0.f; g[i][j] = 0.f; }
                            tile_mma_add(l, Y, sD, rg, cg);            // L = Y D^T
                            tile_store(ws + (size_t)tile_index(bi, bc, nb) * GPIS_TILE_ELEMS, l, rg, cg);
                            __syncwarp();
                            tile_store(Y, l, rg, cg);                  // L replaces Y in the scratch tile
                            __syncwarp();
                            tile_mma_add(g, Y, sDt, rg, cg);           // G = L D
                            tile_store(rec_tiles + (size_t)tile_index(bi, bc, nb) * GPIS_TILE_ELEMS, g, rg, cg);
                        }
                    }
                }
                T2_T(5)
            }
        }
        __threadfence_block();
        __syncthreads();

        // ---------------------------------------------------------------- D. backward solve L^T alpha = z
        // alpha_j = inv(Ljj)^T ( z_j - sum_{i>j} L(i,j)^T alpha_i ), j = nb-1 .. 0; zv is overwritten block by block.
        for (int bj = nb - 1; bj >= 0; --bj) {
            float s = 0.f;
            for (int bi = bj + 1 + warp; bi < nb; bi += T2_WARPS) {
                const float* T = ws + (size_t)tile_index(bi, bj, nb) * GPIS_TILE_ELEMS + lane * 32;  // column `lane`
                const float* al = zv + bi * 32;
#pragma unroll
                for (int r4 = 0; r4 < 8; ++r4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(T + 4 * r4);
                    s = fmaf(t4.x, al[4 * r4 + 0], s);
                    s = fmaf(t4.y, al[4 * r4 + 1], s);
                    s = fmaf(t4.z, al[4 * r4 + 2], s);
                    s = fmaf(t4.w, al[4 * r4 + 3], s);
                }
            }
            part[warp * 32 + lane] = s;
            if (warp == 0) warp_copy_tile(sD0, rec_tiles + (size_t)tile_index(bj, bj, nb) * GPIS_TILE_ELEMS, lane);   // L(bj,bj)
            __syncthreads();
            if (warp == 0) {
                float t = zv[bj * 32 + lane];
                for (int w = 0; w < T2_WARPS; ++w) t -= part[w * 32 + lane];
                zv[bj * 32 + lane] = warp_solve_bwd_t(t, sD0, lane);
            }
            __syncthreads();
        }
        for (int i = tid; i < nb * 32; i += T2_THREADS) rec_alpha[i] = (i < n) ? zv[i] : 0.f;
        T2_T(6)

        if (tid == 0) {
            LeafHeader* h = reinterpret_cast<LeafHeader*>(rec);
            h->N = N; h->ng = ng; h->n = n; h->nb = nb; h->dim = dim; h->chol_fail = 0; h->slot = job.slot;
            h->bytes = rec_bytes(N, nb);
            h->key = cell_key(job.cell[0], job.cell[1], dim == 3 ? job.cell[2] : 0);
            for (int c = 0; c < 3; ++c) { h->cell[c] = job.cell[c]; h->centre[c] = job.centre[c]; h->lo[c] = job.lo[c]; h->hi[c] = job.hi[c]; }
            h->cell[3] = 0; h->centre[3] = 0.f; h->lo[3] = 0.f; h->hi[3] = 0.f;
        }
        // non-positive pivots are counted by warp 0 only
        if (warp == 0 && lane == 0) {
            reinterpret_cast<LeafHeader*>(rec)->chol_fail = bad_total;
            if (status) status[ji] = bad_total;
        }
    }
}

// Persistent launch: 2 CTAs per SM (or fewer for small batches); *ws is a grow-only per-CTA workspace.
static inline int launch_leaf_train2(cudaStream_t st, const TrainJob* d_jobs, int njobs, const float* d_samples,
                                     const TrainParams& P, int32_t* d_status, int maxnb, int num_sms, float** ws,
                                     uint64_t* ws_floats, int32_t* d_counter, std::string& err) {
    if (njobs <= 0) return 0;
    const int grid = njobs < 2 * num_sms ? njobs : 2 * num_sms;
    const uint64_t stride = (uint64_t)maxnb * (maxnb + 1) / 2 * GPIS_TILE_ELEMS;
    const uint64_t need = stride * (uint64_t)grid;
    cudaError_t e = cudaSuccess;
    if (*ws_floats < need) {
        if (*ws) e = cudaFree(*ws);
        *ws = nullptr; *ws_floats = 0;
        const uint64_t want = need + need / 8;
        if (e == cudaSuccess) e = cudaMalloc(ws, want * sizeof(float));
        if (e == cudaSuccess) *ws_floats = want;
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(d_counter, 0, sizeof(int32_t), st);
    if (e == cudaSuccess) {
        k_leaf_train2<<<grid, T2_THREADS, Train2Smem::total(maxnb), st>>>(d_jobs, njobs, d_samples, P, d_status, *ws, stride, d_counter);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { err = std::string("k_leaf_train2: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

}  // namespace gpis
