// K2 — observation GPs (Ornstein-Uhlenbeck kernel), batched over tiles / groups.
// Replaces GPou::train/test (cpp/src/ObsGP.cpp:32-62), ornstein_uhlenbeck (cpp/src/covFnc.cpp:47-109),
// ObsGP2D::trainValidPoints / test_kernel (cpp/src/ObsGP.cpp:280-329, 352-408) and
// ObsGP1D::train/test (cpp/src/ObsGP.cpp:85-187). The partition tables (computePartition,
// ObsGP.cpp:204-265, and the 1-D group ranges, :91-137) are tiny and built on the host side of the
// C ABI; they are passed in as index ranges + boundary values.
//
// Arithmetic order. These are <= 64x64 systems whose outputs decide which samples enter the map, so the
// kernels follow the operation order of the oracle's dense LA exactly (oracle/eigen_shim/Eigen/Dense: LLT with
// ascending-p updates, column-oriented forward substitution, dot-form back substitution, sequential
// K^T alpha and squared column sums; separate multiply and add/subtract, no FMA contraction — the library
// is built with -fmad=false). With that, update() through the GPU yields bit-identical samples.
//
//   k_obs_train          one 64-thread CTA per tile: gather valid pixels in the reference's order, K, in-smem
//                        Cholesky, alpha.
//   k_obs_locate/scan/scatter   bucket the test points by tile (the reference's margin / boundary rules).
//   k_obs_test_grouped   one CTA per (tile, chunk of points): the tile's factor, inputs and alpha are staged in
//                        shared memory once; every thread solves one test point with a register-blocked
//                        (16 rows) forward substitution. Latency/issue bound, not bandwidth bound.
#pragma once
#include "common.cuh"

namespace gpis {

#define OBS_MAXP 64
#define OBS_TEST_THREADS 128
#define OBS_TEST_CHUNKS 4     // CTAs per tile (grid.y); a CTA strides over its tile's chunks of 128 points
struct ObsTileDesc {   // one per tile, built on the host from the partition
    int32_t i0, i1, j0, j1;   // inclusive index ranges (1-D: j0 = j1 = 0)
};
// Trained tile record: fixed stride for direct indexing.
struct ObsTile {
    int32_t p;                 // valid samples (0 = untrained)
    int32_t pad[3];
    float x[2 * OBS_MAXP];     // sample coordinates (d floats each)
    float alpha[OBS_MAXP];
    float L[OBS_MAXP * OBS_MAXP];  // column-major lower factor: L(r,c) at c*64 + r; rows/cols >= p: identity
};

struct ObsParams {
    int d;              // 1 or 2
    int ni;             // 2-D: fast dimension of the pixel grid
    float a;            // 1/scale              covFnc.cpp:51
    float diag;         // (float)(1.0 + noise) covFnc.cpp:57
    float var_prior;    // 1 + noise            ObsGP.cpp:61
    float margin;
    int nb0, nb1;       // number of boundary values (Val_i / Val_j or range)
    int ng0, ntiles;
};

__global__ void __launch_bounds__(OBS_MAXP)
k_obs_train(const float* __restrict__ xin, const float* __restrict__ fin, const ObsTileDesc* __restrict__ desc,
            ObsTile* __restrict__ tiles, ObsParams P) {
    __shared__ float Ks[OBS_MAXP * (OBS_MAXP + 1)];   // row-padded to dodge bank conflicts
    __shared__ float xs[2 * OBS_MAXP];
    __shared__ float ys[OBS_MAXP];
    __shared__ int idx[OBS_MAXP];
    __shared__ int cnt;
    const int t = threadIdx.x;
    const ObsTileDesc D = desc[blockIdx.x];
    ObsTile* out = tiles + blockIdx.x;
    // gather valid samples in the reference's order: j outer, i inner (ObsGP.cpp:301-310)
    const int wi = D.i1 - D.i0 + 1, wj = D.j1 - D.j0 + 1;
    const int tot = wi * wj;   // <= 64
    int ind = -1;
    bool valid = false;
    if (t < tot) {
        const int j = D.j0 + t / wi, i = D.i0 + t % wi;
        ind = (P.d == 2) ? j * P.ni + i : i;
        valid = (P.d == 2) ? (fin[ind] > 0.f) : true;   // 1-D groups take every beam (ObsGP.cpp:104-108)
    }
    // ordered compaction over 64 threads (2 warps)
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    __shared__ int wcount[2];
    if ((t & 31) == 0) wcount[t >> 5] = __popc(m);
    __syncthreads();
    const int pos = ((t >> 5) ? wcount[0] : 0) + __popc(m & ((1u << (t & 31)) - 1u));
    if (t == 0) cnt = wcount[0] + wcount[1];
    if (valid) idx[pos] = ind;
    __syncthreads();
    const int p = cnt;
    if (p < 1) { if (t == 0) out->p = 0; return; }   // ObsGP.cpp:313: needs >= 1 valid pixel
    if (t < p) {
        const int id = idx[t];
        for (int c = 0; c < P.d; ++c) xs[t * P.d + c] = xin[(size_t)id * P.d + c];
        ys[t] = fin[id];
    }
    __syncthreads();
    // K = OU(x) + noise on the diagonal (covFnc.cpp:54-66); thread t builds row t
    if (t < p) {
        for (int j = 0; j < p; ++j) {
            float v;
            if (j == t) v = P.diag;
            else {
                float s2 = 0.f;
                for (int c = 0; c < P.d; ++c) { const float dd = xs[t * P.d + c] - xs[j * P.d + c]; s2 = (c == 0) ? dd * dd : s2 + dd * dd; }
                // the reference computes the (k<j) entry and mirrors it; |d| is symmetric so both agree
                v = df_round(exp_df(-P.a * sqrtf(s2)));
            }
            Ks[t * (OBS_MAXP + 1) + j] = v;
        }
    }
    __syncthreads();
    // Cholesky, thread t owns row t. Element (t, c) receives its updates in ascending j, each one a rounded
    // product followed by a rounded subtraction: the order of the oracle's LLT (Eigen/Dense shim: factor_panel +
    // trailing update; blocked or not, every element sees p = 0..c-1 in that order).
    for (int j = 0; j < p; ++j) {
        const float d = sqrtf(Ks[j * (OBS_MAXP + 1) + j]);
        __syncthreads();
        if (t == j) Ks[j * (OBS_MAXP + 1) + j] = d;
        if (t > j && t < p) Ks[t * (OBS_MAXP + 1) + j] = Ks[t * (OBS_MAXP + 1) + j] / d;
        __syncthreads();
        if (t > j && t < p) {
            const float l = Ks[t * (OBS_MAXP + 1) + j];
            for (int c = j + 1; c <= t; ++c) {
                const float prod = l * Ks[c * (OBS_MAXP + 1) + j];
                Ks[t * (OBS_MAXP + 1) + c] = Ks[t * (OBS_MAXP + 1) + c] - prod;
            }
        }
        __syncthreads();
    }
    // alpha, forward: column-oriented (axpy form), b[j] /= L[j][j]; b[i] -= b[j] * L[i][j] (ObsGP.cpp:42-43)
    for (int j = 0; j < p; ++j) {
        if (t == j) ys[j] = ys[j] / Ks[j * (OBS_MAXP + 1) + j];
        __syncthreads();
        if (t > j && t < p) {
            const float prod = ys[j] * Ks[t * (OBS_MAXP + 1) + j];
            ys[t] = ys[t] - prod;
        }
        __syncthreads();
    }
    // alpha, backward: dot form, s = b[i] - L[i+1][i] b[i+1] - ... in ascending order, then / L[i][i]
    // (ObsGP.cpp:44). The order makes the chain serial; one thread walks it (2,016 steps at p = 64).
    if (t == 0) {
        for (int i = p - 1; i >= 0; --i) {
            float s = ys[i];
            for (int j = i + 1; j < p; ++j) {
                const float prod = Ks[j * (OBS_MAXP + 1) + i] * ys[j];
                s = s - prod;
            }
            ys[i] = s / Ks[i * (OBS_MAXP + 1) + i];
        }
    }
    __syncthreads();
    if (t == 0) out->p = p;
    for (int c = 0; c < P.d; ++c) out->x[t * P.d + c] = (t < p) ? xs[t * P.d + c] : 0.f;
    out->alpha[t] = (t < p) ? ys[t] : 0.f;
    // L column-major: thread t writes row t of each column -> coalesced. Padded to 64 with the identity so the
    // test kernel's fixed-size blocks need no bounds (a padded row contributes exact zeros).
    for (int c = 0; c < OBS_MAXP; ++c) {
        float v = 0.f;
        if (t < p && c < p) v = (t >= c) ? Ks[t * (OBS_MAXP + 1) + c] : 0.f;
        else if (t == c) v = 1.f;
        out->L[c * OBS_MAXP + t] = v;
    }
}

// ------------------------------------------------------------------ test: bucketing by tile
// Tile of a test point with the reference's margin rules (ObsGP.cpp:363-377, 152-186); -1 = not evaluated.
__device__ __forceinline__ int obs_tile_of(float x0, float x1, const float* __restrict__ b0, const float* __restrict__ b1,
                                           const ObsParams& P) {
    int tile = -1;
    if (P.d == 2) {
        if (!(x0 < b0[0] + P.margin) && !(x0 > b0[P.nb0 - 1] - P.margin) && !(x1 < b1[0] + P.margin) &&
            !(x1 > b1[P.nb1 - 1] - P.margin)) {
            int n = 0, mm = 0;
            for (int i = 1; i < P.nb0; ++i, ++n) if (x0 < b0[i]) break;
            for (int i = 1; i < P.nb1; ++i, ++mm) if (x1 < b1[i]) break;
            const int id = mm * P.ng0 + n;
            if (id < P.ntiles) tile = id;
        }
    } else {
        const float liml = b0[0] + P.margin, limr = b0[P.nb0 - 1] - P.margin;
        if (!(x0 < liml) && !(x0 > limr)) {
            for (int j = 0; j + 1 < P.nb0; ++j)
                if (x0 > b0[j] && x0 < b0[j + 1]) { if (j < P.ntiles) tile = j; break; }
        }
    }
    return tile;
}

// var = 1e6 and val untouched where the reference would not evaluate (ObsGP.cpp:363-377, 396-403, 152-186)
__global__ void __launch_bounds__(256)
k_obs_locate(const float* __restrict__ xt, int m, const float* __restrict__ b0, const float* __restrict__ b1,
             const ObsTile* __restrict__ tiles, ObsParams P, int32_t* __restrict__ tile_of, int32_t* __restrict__ count,
             float* __restrict__ var) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const float x0 = xt[(size_t)q * P.d], x1 = (P.d == 2) ? xt[(size_t)q * P.d + 1] : 0.f;
    int tile = obs_tile_of(x0, x1, b0, b1, P);
    if (tile >= 0 && tiles[tile].p < 1) tile = -1;
    tile_of[q] = tile;
    if (tile >= 0) atomicAdd(count + tile, 1);
    else var[q] = 1e6f;
}
// exclusive scan of count[0..n) -> start[0..n], cursor := start   (one 1024-thread block, n of any size)
__global__ void __launch_bounds__(1024)
k_obs_scan(const int32_t* __restrict__ count, int n, int32_t* __restrict__ start, int32_t* __restrict__ cursor) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int b = t * per, e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += count[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = (t == 0) ? 0 : part[t - 1];
    for (int i = b; i < e; ++i) { start[i] = run; cursor[i] = run; run += count[i]; }
    if (t == 1023) start[n] = part[1023];
}
__global__ void __launch_bounds__(256)
k_obs_scatter(const int32_t* __restrict__ tile_of, int m, int32_t* __restrict__ cursor, int32_t* __restrict__ order) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int tile = tile_of[q];
    if (tile >= 0) order[atomicAdd(cursor + tile, 1)] = q;
}

// One CTA per (tile, chunk): thread = one test point. GPou::test (ObsGP.cpp:50-62) in the oracle's order:
//   k_i = exp(-a |x_i - x|);  f = sum_i k_i alpha_i (ascending);  forward substitution in axpy form (every k_i takes
//   its updates in ascending j);  var = (1 + noise) - sum_j v_j^2 (ascending).
// The substitution is blocked by 16 rows held in registers; finished v_j are parked in shared memory (one
// column per thread, conflict-free) and re-read by the later blocks; L is read as 16-byte broadcasts.
__global__ void __launch_bounds__(OBS_TEST_THREADS)
k_obs_test_grouped(const float* __restrict__ xt, const int32_t* __restrict__ start, const int32_t* __restrict__ order,
                   const ObsTile* __restrict__ tiles, ObsParams P, float* __restrict__ val, float* __restrict__ var) {
    const int tile = blockIdx.x;
    const int first = start[tile], npts = start[tile + 1] - first;
    if ((int)blockIdx.y * OBS_TEST_THREADS >= npts) return;
    __shared__ __align__(16) float Ls[OBS_MAXP * OBS_MAXP];
    __shared__ float xs[2 * OBS_MAXP];
    __shared__ float al[OBS_MAXP];
    __shared__ float vs[(OBS_MAXP - 16) * OBS_TEST_THREADS];   // finished v_j of the first 48 rows, one column per thread
    const int tid = threadIdx.x;
    const ObsTile* T = tiles + tile;
    const int p = T->p;
    {
        const float4* src = reinterpret_cast<const float4*>(T->L);
        float4* dst = reinterpret_cast<float4*>(Ls);
        for (int i = tid; i < OBS_MAXP * OBS_MAXP / 4; i += OBS_TEST_THREADS) dst[i] = src[i];
        for (int i = tid; i < 2 * OBS_MAXP; i += OBS_TEST_THREADS) xs[i] = T->x[i];
        for (int i = tid; i < OBS_MAXP; i += OBS_TEST_THREADS) al[i] = T->alpha[i];
    }
    __syncthreads();
    const int nblk = (p + 15) >> 4;
    for (int base = blockIdx.y * OBS_TEST_THREADS; base < npts; base += OBS_TEST_CHUNKS * OBS_TEST_THREADS) {
        const int k = base + tid;
        if (k >= npts) continue;   // no block-wide barrier below: every thread only touches its own vs column
        const int q = order[first + k];
        const float x0 = xt[(size_t)q * P.d], x1 = (P.d == 2) ? xt[(size_t)q * P.d + 1] : 0.f;
        // k* (covFnc.cpp:93-109) is built block by block straight into registers; the mean K^T alpha
        // (ObsGP.cpp:54) accumulates in ascending i along the way
        float f = 0.f, ss = 0.f;
        for (int ib = 0; ib < nblk; ++ib) {
            float kk[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int i = 16 * ib + r;
                float ki = 0.f;
                if (i < p) {
                    const float d0 = xs[i * P.d] - x0;
                    float s2 = d0 * d0;
                    if (P.d == 2) { const float d1 = xs[i * P.d + 1] - x1; s2 = s2 + d1 * d1; }
                    ki = df_round(exp_df(-P.a * sqrtf(s2)));
                    const float prod = ki * al[i];
                    f = f + prod;
                }
                kk[r] = ki;
            }
            for (int j = 0; j < 16 * ib; ++j) {
                const float vj = vs[j * OBS_TEST_THREADS + tid];
                const float4* Lc = reinterpret_cast<const float4*>(Ls + j * OBS_MAXP + 16 * ib);
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const float4 l = Lc[r4];
                    kk[4 * r4 + 0] = kk[4 * r4 + 0] - vj * l.x;
                    kk[4 * r4 + 1] = kk[4 * r4 + 1] - vj * l.y;
                    kk[4 * r4 + 2] = kk[4 * r4 + 2] - vj * l.z;
                    kk[4 * r4 + 3] = kk[4 * r4 + 3] - vj * l.w;
                }
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float* Lc = Ls + (16 * ib + jj) * OBS_MAXP + 16 * ib;
                const float vj = kk[jj] / Lc[jj];
                kk[jj] = vj;
                ss = ss + vj * vj;
#pragma unroll
                for (int r = jj + 1; r < 16; ++r) kk[r] = kk[r] - vj * Lc[r];
            }
            if (ib + 1 < nblk) {   // the last block's v is never re-read
#pragma unroll
                for (int r = 0; r < 16; ++r) vs[(16 * ib + r) * OBS_TEST_THREADS + tid] = kk[r];
            }
        }
        val[q] = f;
        var[q] = P.var_prior - ss;
    }
}

}  // namespace gpis
