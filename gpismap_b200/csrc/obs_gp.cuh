// K2 — observation GPs (Ornstein-Uhlenbeck kernel), batched over tiles / groups.
// Replaces GPou::train/test (cpp/src/ObsGP.cpp:32-62), ornstein_uhlenbeck (cpp/src/covFnc.cpp:47-109),
// ObsGP2D::trainValidPoints / test_kernel (cpp/src/ObsGP.cpp:280-329, 352-408) and
// ObsGP1D::train/test (cpp/src/ObsGP.cpp:85-187). The partition tables (computePartition,
// ObsGP.cpp:204-265, and the 1-D group ranges, :91-137) are tiny and built on the host side of the
// C ABI; they are passed in as index ranges + boundary values.
//
// Each tile holds at most 64 samples ((5+3)^2 pixels, or <= 26 beams), so one 64-thread CTA trains
// one tile entirely in shared memory: K (64x64), in-place Cholesky, alpha. Latency/occupancy
// bound (thousands of tiny factorizations), not bandwidth or FMA bound.
#pragma once
#include "common.cuh"

namespace gpis {

#define OBS_MAXP 64
struct ObsTileDesc {   // one per tile, built on the host from the partition
    int32_t i0, i1, j0, j1;   // inclusive index ranges (1-D: j0 = j1 = 0)
};
// Trained tile record: fixed stride for direct indexing.
struct ObsTile {
    int32_t p;                 // valid samples (0 = untrained)
    int32_t pad[3];
    float x[2 * OBS_MAXP];     // sample coordinates (d floats each)
    float alpha[OBS_MAXP];
    float L[OBS_MAXP * OBS_MAXP];  // column-major lower factor: L(r,c) at c*64 + r
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
    // right-looking Cholesky, thread t owns row t
    for (int j = 0; j < p; ++j) {
        const float d = sqrtf(Ks[j * (OBS_MAXP + 1) + j]);
        __syncthreads();
        if (t == j) Ks[j * (OBS_MAXP + 1) + j] = d;
        if (t > j && t < p) Ks[t * (OBS_MAXP + 1) + j] /= d;
        __syncthreads();
        if (t > j && t < p) {
            const float l = Ks[t * (OBS_MAXP + 1) + j];
            for (int c = j + 1; c <= t; ++c) Ks[t * (OBS_MAXP + 1) + c] = fmaf(-l, Ks[c * (OBS_MAXP + 1) + j], Ks[t * (OBS_MAXP + 1) + c]);
        }
        __syncthreads();
    }
    // alpha: forward then backward substitution (ObsGP.cpp:42-44), column-oriented, thread t = row t
    for (int j = 0; j < p; ++j) {
        if (t == j) ys[j] = ys[j] / Ks[j * (OBS_MAXP + 1) + j];
        __syncthreads();
        if (t > j && t < p) ys[t] = fmaf(-Ks[t * (OBS_MAXP + 1) + j], ys[j], ys[t]);
        __syncthreads();
    }
    for (int j = p - 1; j >= 0; --j) {
        if (t == j) ys[j] = ys[j] / Ks[j * (OBS_MAXP + 1) + j];
        __syncthreads();
        if (t < j) ys[t] = fmaf(-Ks[j * (OBS_MAXP + 1) + t], ys[j], ys[t]);
        __syncthreads();
    }
    if (t == 0) out->p = p;
    if (t < p) {
        for (int c = 0; c < P.d; ++c) out->x[t * P.d + c] = xs[t * P.d + c];
        out->alpha[t] = ys[t];
    }
    // L column-major: thread t writes row t of each column -> coalesced
    for (int c = 0; c < p; ++c)
        if (t < p) out->L[c * OBS_MAXP + t] = (t >= c) ? Ks[t * (OBS_MAXP + 1) + c] : 0.f;
}

// One warp per test point. val/var are read-modify-write: var = 1e6 and val untouched where the
// reference would not evaluate (ObsGP.cpp:363-377, 396-403, 152-186).
__global__ void __launch_bounds__(256)
k_obs_test(const float* __restrict__ xt, int m, const float* __restrict__ b0, const float* __restrict__ b1,
           const ObsTile* __restrict__ tiles, ObsParams P, float* __restrict__ val, float* __restrict__ var) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= m) return;
    float x0 = xt[(size_t)gw * P.d], x1 = (P.d == 2) ? xt[(size_t)gw * P.d + 1] : 0.f;
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
    int p = 0;
    const ObsTile* T = nullptr;
    if (tile >= 0) { T = tiles + tile; p = T->p; }
    if (p < 1) { if (lane == 0) var[gw] = 1e6f; return; }
    // k* (covFnc.cpp:93-109): lane handles samples lane and lane+32
    float k[2] = {0.f, 0.f};
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < p) {
            float s2;
            const float d0 = T->x[i * P.d] - x0;
            s2 = d0 * d0;
            if (P.d == 2) { const float d1 = T->x[i * P.d + 1] - x1; s2 = s2 + d1 * d1; }
            k[h] = df_round(exp_df(-P.a * sqrtf(s2)));
        }
    }
    float f = 0.f;
    for (int h = 0; h < 2; ++h) { const int i = lane + 32 * h; if (i < p) f = fmaf(k[h], T->alpha[i], f); }
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    // forward substitution, column-oriented: v_j = k_j / L_jj, then k_r -= L_rj v_j for r > j
    float ss = 0.f;
    for (int j = 0; j < p; ++j) {
        const float kj = __shfl_sync(0xffffffffu, (j < 32) ? k[0] : k[1], j & 31);
        const float vj = kj / T->L[j * OBS_MAXP + j];
        ss = fmaf(vj, vj, ss);
        const float* col = T->L + j * OBS_MAXP;
        if (lane > j && lane < p) k[0] = fmaf(-col[lane], vj, k[0]);
        if (lane + 32 > j && lane + 32 < p) k[1] = fmaf(-col[lane + 32], vj, k[1]);
    }
    if (lane == 0) { val[gw] = f; var[gw] = P.var_prior - ss; }
}

}  // namespace gpis
