// SURVEY.md 8 f-3 + f-1 — per-frame sensor pipeline of GPisMap3::update on the device.
//   f-3  preprocData (cpp/src/GPisMap3.cpp:125-216): sub-sampled depth -> validity, inverse depth, back-projection,
//        local -> global transform, range_obs_max; valid measurements compacted in the reference's order
//        (columns outer, rows inner).
//   f-1  evalPoints numerics (cpp/src/GPisMap3.cpp:580-696): for every valid measurement the observation-GP test at
//        its own pixel and at six finite-difference probes (batched through K2's grouped test kernel), then the
//        occupancy values, surface normal, position / normal noise. The host loop that follows only mutates the tree.
//        reEvalPoints numerics (cpp/src/GPisMap3.cpp:321-534) for the in-view samples: projection, first test,
//        occupancy gate + the 10-step walk (which re-tests the same pixel), six probes, fused normal / position /
//        noises (k_reeval_stage1, k_reeval_numerics).
// Every expression keeps the reference's scalar types (float storage, double where the reference's unqualified
// exp/acos/cos/sin and double literals promote), evaluated without FMA contraction (-fmad=false), so the samples
// the host inserts are the reference's, bit for bit (tests: whole-map SHA-256 against the reference's own mapping).
// HBM-bound elementwise work plus K2's test kernel; a few hundred microseconds per frame.
#pragma once
#include "common.cuh"

namespace gpis {

struct FrameParams {
    int width, height, skip, n, m;     // n = width/skip sub-sampled columns, m = height/skip rows
    int N;                             // pixels in the depth image handed in
    float R[9], t[3];                  // pose: R column-major, local -> global (GPisMap3.cpp:141-142)
    float delx, obs_var_thre, min_position_noise, min_grad_noise;
    double max_range, min_range;       // params.h:77-78
};

// occ_test (GPisMap3.cpp:38-41): float arguments, double inside, float result
__device__ __forceinline__ float occ_test_dev(float rinv, float rinv0, float a) {
    return (float)(2.0 * (1.0 / (1.0 + exp((double)(-a * (rinv - rinv0)))) - 0.5));
}
__device__ __forceinline__ float saturate_dev(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// pass 1: validity + inverse depth per sub-sampled pixel (index g = m * n_ + m_, the layout of vu_grid / zinv)
__global__ void __launch_bounds__(256)
k_frame_valid(const float* __restrict__ depth, FrameParams P, float* __restrict__ zinv, int32_t* __restrict__ flag) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n * P.m) return;
    const int n_ = g / P.m, m_ = g % P.m;
    const int k = (n_ * P.skip) * P.height + m_ * P.skip;             // GPisMap3.cpp:183: dataz[col*height + row]
    bool ok = false;
    float z = 0.f;
    if (k < P.N) { z = depth[k]; ok = ((double)z < P.max_range) && ((double)z > P.min_range); }   // isRangeValid, :33-36
    zinv[g] = ok ? (float)(1.0 / (double)z) : -1.0f;                   // :188 / :208
    flag[g] = ok ? 1 : 0;
}
// pass 2: compaction in g order + back-projection (GPisMap3.cpp:186-206). start = exclusive scan of flag.
__global__ void __launch_bounds__(256)
k_frame_project(const float* __restrict__ depth, const float* __restrict__ vu, FrameParams P, const int32_t* __restrict__ flag,
                const int32_t* __restrict__ start, float* __restrict__ vu_valid, float* __restrict__ xyz_local,
                float* __restrict__ xyz_global, int32_t* __restrict__ range_max_bits) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n * P.m || !flag[g]) return;
    const int n_ = g / P.m, m_ = g % P.m;
    const float z = depth[(n_ * P.skip) * P.height + m_ * P.skip];
    const int kk = start[g];
    const float v = vu[2 * g], u = vu[2 * g + 1];
    vu_valid[2 * kk] = v; vu_valid[2 * kk + 1] = u;                   // test inputs are [v, u] (GPisMap3.cpp:597-598)
    const float xloc = u * z, yloc = v * z;
    xyz_local[3 * kk] = xloc; xyz_local[3 * kk + 1] = yloc; xyz_local[3 * kk + 2] = z;
    xyz_global[3 * kk] = P.R[0] * xloc + P.R[3] * yloc + P.R[6] * z + P.t[0];
    xyz_global[3 * kk + 1] = P.R[1] * xloc + P.R[4] * yloc + P.R[7] * z + P.t[1];
    xyz_global[3 * kk + 2] = P.R[2] * xloc + P.R[5] * yloc + P.R[8] * z + P.t[2];
    atomicMax(range_max_bits, __float_as_int(z));                     // z > 0: float order = integer order
}
// six probes around every valid measurement (GPisMap3.cpp:633-641), test inputs [Y/Z, X/Z]
__global__ void __launch_bounds__(256)
k_frame_probes(const float* __restrict__ xyz_local, int K, float delx, float* __restrict__ vu_probe) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * K) return;
    const int kk = t / 6, i = t % 6;
    const float Xp = (i == 0) ? 1.f : (i == 1) ? -1.f : 0.f;
    const float Yp = (i == 2) ? 1.f : (i == 3) ? -1.f : 0.f;
    const float Zp = (i == 4) ? 1.f : (i == 5) ? -1.f : 0.f;
    const float X = xyz_local[3 * kk] + delx * Xp, Y = xyz_local[3 * kk + 1] + delx * Yp, Z = xyz_local[3 * kk + 2] + delx * Zp;
    vu_probe[2 * t] = Y / Z;
    vu_probe[2 * t + 1] = X / Z;
}
// numerics of evalPoints for one measurement (GPisMap3.cpp:599-686). status: 0 = centre test rejected (skip),
// 1 = a probe was rejected (the sample is inserted and removed again, :652-655), 2 = ok.
__global__ void __launch_bounds__(256)
k_frame_numerics(const float* __restrict__ xyz_local, int K, FrameParams P, const float* __restrict__ var_c,
                 const float* __restrict__ rinv0_p, const float* __restrict__ var_p, int32_t* __restrict__ status,
                 float* __restrict__ grad_out, float* __restrict__ noise_out, float* __restrict__ grad_noise_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (var_c[k] > P.obs_var_thre) { status[k] = 0; return; }
    const float w = (float)(1.0 / 6.0);
    float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
    float occ_mean = 0.0f;
    const float* xl = xyz_local + 3 * k;
    const float* r0 = rinv0_p + 6 * k;
    const float* vr = var_p + 6 * k;
    bool failed = false;
    for (int i = 0; i < 6; i++) {
        if (vr[i] > P.obs_var_thre) { failed = true; break; }
        const float Zp = (i == 4) ? 1.f : (i == 5) ? -1.f : 0.f;
        const float Z = xl[2] + P.delx * Zp;
        occ[i] = occ_test_dev((float)(1.0 / (double)Z), r0[i], (float)((double)Z * 30.0));
        occ_mean += w * occ[i];
    }
    if (failed) { status[k] = 1; return; }
    float noise = 100.0f, grad_noise = 1.00f;
    float grad[3];
    grad[0] = (occ[0] - occ[1]) / P.delx;
    grad[1] = (occ[2] - occ[3]) / P.delx;
    grad[2] = (occ[4] - occ[5]) / P.delx;
    float norm_grad = grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2];
    if ((double)norm_grad > 1e-6) {
        norm_grad = sqrtf(norm_grad);
        const float glx = grad[0] / norm_grad, gly = grad[1] / norm_grad, glz = grad[2] / norm_grad;
        grad[0] = P.R[0] * glx + P.R[3] * gly + P.R[6] * glz;
        grad[1] = P.R[1] * glx + P.R[4] * gly + P.R[7] * glz;
        grad[2] = P.R[2] * glx + P.R[5] * gly + P.R[8] * glz;
        const float dist = sqrtf(xl[0] * xl[0] + xl[1] * xl[1] + xl[2] * xl[2]);
        noise = P.min_position_noise * (saturate_dev(dist, 1.0f, noise));
        grad_noise = saturate_dev(fabsf(occ_mean), P.min_grad_noise, grad_noise);
        const float view_ang = fmaxf(-(xl[0] * glx + xl[1] * gly + xl[2] * glz) / dist, (float)1e-1);
        const float view_ang2 = view_ang * view_ang;
        const float view_noise = (float)((double)P.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
        noise += view_noise;
    }
    grad_out[3 * k] = grad[0]; grad_out[3 * k + 1] = grad[1]; grad_out[3 * k + 2] = grad[2];
    noise_out[k] = noise; grad_noise_out[k] = grad_noise;
    status[k] = 2;
}

}  // namespace gpis
