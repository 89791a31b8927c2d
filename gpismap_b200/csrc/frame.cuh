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

namespace gpis {

// ------------------------------------------------------------------ reEvalPoints on the device (f-1, second half)
struct ReevalParams {
    float R[9], t[3];
    float delx, obs_var_thre, min_position_noise, min_grad_noise, map_noise_param;
};

// quat2dcm (GPisMap3.cpp:48-64)
__device__ __forceinline__ void quat2dcm_dev(const float q[4], float dcm[9]) {
    dcm[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
    dcm[1] = (float)(2.0 * (double)(q[1] * q[2] + q[0] * q[3]));
    dcm[2] = (float)(2.0 * (double)(q[1] * q[3] - q[0] * q[2]));
    dcm[3] = (float)(2.0 * (double)(q[1] * q[2] - q[0] * q[3]));
    dcm[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
    dcm[5] = (float)(2.0 * (double)(q[0] * q[1] + q[2] * q[3]));
    dcm[6] = (float)(2.0 * (double)(q[1] * q[3] + q[0] * q[2]));
    dcm[7] = (float)(2.0 * (double)(q[2] * q[3] - q[0] * q[1]));
    dcm[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}

// samples: 8 floats each [pos(3), grad(3), pose_sig, grad_sig]. Projection into the camera (GPisMap3.cpp:334-346);
// a sample behind the camera gets a test input far outside the image (the test then reports var = 1e6).
__global__ void __launch_bounds__(256)
k_reeval_project(const float* __restrict__ smp, int n, ReevalParams P, float* __restrict__ loc, float* __restrict__ vu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = smp + 8 * (size_t)i;
    const float d0 = s[0] - P.t[0], d1 = s[1] - P.t[1], d2 = s[2] - P.t[2];
    const float l0 = P.R[0] * d0 + P.R[1] * d1 + P.R[2] * d2;
    const float l1 = P.R[3] * d0 + P.R[4] * d1 + P.R[5] * d2;
    const float l2 = P.R[6] * d0 + P.R[7] * d1 + P.R[8] * d2;
    loc[3 * i] = l0; loc[3 * i + 1] = l1; loc[3 * i + 2] = l2;
    const bool front = !((double)l2 < 0.0);
    vu[2 * i] = front ? l1 / l2 : 1.0e30f;
    vu[2 * i + 1] = front ? l0 / l2 : 1.0e30f;
}
// occupancy gate + the 10-step walk (GPisMap3.cpp:349-411; the re-test uses the original pixel), then the
// six probes around x_new (:413-430). alive[i] = 1 when the sample goes on to the numerics.
__global__ void __launch_bounds__(256)
k_reeval_walk(const float* __restrict__ smp, int n, ReevalParams P, const float* __restrict__ loc, const float* __restrict__ rinv0,
              const float* __restrict__ var, int32_t* __restrict__ alive, float* __restrict__ xnew, float* __restrict__ absoc,
              float* __restrict__ vu_probe) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x_loc = loc[3 * i], y_loc = loc[3 * i + 1], z_loc = loc[3 * i + 2];
    bool ok = !((double)z_loc < 0.0) && !(var[i] > P.obs_var_thre);
    float oc = 0.f, abs_oc = 0.f;
    float x_new[3] = {x_loc, y_loc, z_loc};
    if (ok) {
        const float rinv = (float)(1.0 / (double)z_loc);
        oc = occ_test_dev(rinv, rinv0[i], (float)((double)z_loc * 30.0));
        if ((double)oc < -0.02) ok = false;
    }
    if (ok) {
        const float* s = smp + 8 * (size_t)i;
        float gl[3];
        gl[0] = P.R[0] * s[3] + P.R[1] * s[4] + P.R[2] * s[5];
        gl[1] = P.R[3] * s[3] + P.R[4] * s[4] + P.R[5] * s[5];
        gl[2] = P.R[6] * s[3] + P.R[7] * s[4] + P.R[8] * s[5];
        abs_oc = (float)fabs((double)oc);
        float dx = P.delx;
        for (int it = 0; it < 10 && (double)abs_oc > 0.02; it++) {
            if (oc < 0) { x_new[0] += gl[0] * dx; x_new[1] += gl[1] * dx; x_new[2] += gl[2] * dx; }
            else        { x_new[0] -= gl[0] * dx; x_new[1] -= gl[1] * dx; x_new[2] -= gl[2] * dx; }
            const float r_new = z_loc;
            const float oc_new = occ_test_dev((float)(1.0 / (double)r_new), rinv0[i], (float)((double)r_new * 30.0));
            const float abs_oc_new = (float)fabs((double)oc_new);
            if ((double)abs_oc_new < 0.02 || (double)oc < -0.02) break;
            else if ((double)(oc * oc_new) < 0.0) dx = (float)(0.5 * (double)dx);
            else dx = (float)(1.1 * (double)dx);
            abs_oc = abs_oc_new;
            oc = oc_new;
        }
    }
    alive[i] = ok ? 1 : 0;
    xnew[3 * i] = x_new[0]; xnew[3 * i + 1] = x_new[1]; xnew[3 * i + 2] = x_new[2];
    absoc[i] = abs_oc;
    for (int p = 0; p < 6; ++p) {
        float a = 1.0e30f, b = 1.0e30f;
        if (ok) {
            const float Xp = (p == 0) ? 1.f : (p == 1) ? -1.f : 0.f;
            const float Yp = (p == 2) ? 1.f : (p == 3) ? -1.f : 0.f;
            const float Zp = (p == 4) ? 1.f : (p == 5) ? -1.f : 0.f;
            const float X = x_new[0] + P.delx * Xp, Y = x_new[1] + P.delx * Yp, Z = x_new[2] + P.delx * Zp;
            a = Y / Z; b = X / Z;
        }
        vu_probe[2 * (6 * (size_t)i + p)] = a;
        vu_probe[2 * (6 * (size_t)i + p) + 1] = b;
    }
}
// The rest of one reEvalPoints iteration (GPisMap3.cpp:413-534). action: -1 = not re-evaluated, 0 = nothing,
// 1 = double both noises (:451-454), 2 = replace by (pos_new, grad_new, noise, grad_noise).
__global__ void __launch_bounds__(128)
k_reeval_numerics(const float* __restrict__ smp, int n, ReevalParams P, const float* __restrict__ loc, const int32_t* __restrict__ alive,
                  const float* __restrict__ xnew, const float* __restrict__ absoc, const float* __restrict__ rinv0p,
                  const float* __restrict__ varp, int32_t* __restrict__ action, float* __restrict__ pos_out,
                  float* __restrict__ grad_out, float* __restrict__ noise_out, float* __restrict__ gnoise_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!alive[i]) { action[i] = -1; return; }
    const float* rinv0 = rinv0p + 6 * (size_t)i;
    const float* var = varp + 6 * (size_t)i;
    const float* xn = xnew + 3 * (size_t)i;
    const float* s = smp + 8 * (size_t)i;
    const float w = (float)(1.0 / 6.0);
    float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
    float occ_mean = 0.0f, r0_mean = 0.0f, r0_sqr_sum = 0.0f;
    float r_new = loc[3 * (size_t)i + 2];
    float last_var = 0.f;
    for (int p = 0; p < 6; p++) {
        const float Zp = (p == 4) ? 1.f : (p == 5) ? -1.f : 0.f;
        const float Z = xn[2] + P.delx * Zp;
        r_new = Z;
        last_var = var[p];
        if (var[p] > P.obs_var_thre) break;
        occ[p] = occ_test_dev((float)(1.0 / (double)r_new), rinv0[p], (float)((double)r_new * 30.0));
        occ_mean += w * occ[p];
        const float r0 = (float)(1.0 / (double)rinv0[p]);
        r0_sqr_sum += r0 * r0;
        r0_mean += w * r0;
    }
    if (last_var > P.obs_var_thre) { action[i] = 0; return; }
    const float pos[3] = {s[0], s[1], s[2]};
    const float grad[3] = {s[3], s[4], s[5]};
    float gnl[3];
    gnl[0] = (occ[0] - occ[1]) / P.delx;
    gnl[1] = (occ[2] - occ[3]) / P.delx;
    gnl[2] = (occ[4] - occ[5]) / P.delx;
    const float norm_grad_new = sqrtf(gnl[0] * gnl[0] + gnl[1] * gnl[1] + gnl[2] * gnl[2]);
    if ((double)norm_grad_new < 1e-3) { action[i] = 1; return; }
    float r_var = (float)((double)r0_sqr_sum / 5.0 - (double)(r0_mean * r0_mean) * 6.0 / 5.0);
    r_var /= P.delx;
    float noise = 100.0f, grad_noise = 1.0f;
    if ((double)norm_grad_new > 1e-6) {
        gnl[0] = gnl[0] / norm_grad_new; gnl[1] = gnl[1] / norm_grad_new; gnl[2] = gnl[2] / norm_grad_new;
        noise = P.min_position_noise * saturate_dev(r_new * r_new, 1.0f, noise);
        grad_noise = saturate_dev(fabsf(occ_mean) + r_var, P.min_grad_noise, grad_noise);
    } else {
        noise = P.min_position_noise * noise;
    }
    const float dist = sqrtf(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
    const float view_ang = fmaxf(-(xn[0] * gnl[0] + xn[1] * gnl[1] + xn[2] * gnl[2]) / dist, (float)1e-1);
    const float view_ang2 = view_ang * view_ang;
    const float view_noise = (float)((double)P.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
    noise += view_noise + absoc[i];
    grad_noise = (float)((double)grad_noise + 0.1 * (double)view_noise);

    float pos_new[3], grad_new[3];
    pos_new[0] = P.R[0] * xn[0] + P.R[3] * xn[1] + P.R[6] * xn[2] + P.t[0];
    pos_new[1] = P.R[1] * xn[0] + P.R[4] * xn[1] + P.R[7] * xn[2] + P.t[1];
    pos_new[2] = P.R[2] * xn[0] + P.R[5] * xn[1] + P.R[8] * xn[2] + P.t[2];
    grad_new[0] = P.R[0] * gnl[0] + P.R[3] * gnl[1] + P.R[6] * gnl[2];
    grad_new[1] = P.R[1] * gnl[0] + P.R[4] * gnl[1] + P.R[7] * gnl[2];
    grad_new[2] = P.R[2] * gnl[0] + P.R[5] * gnl[1] + P.R[8] * gnl[2];

    const float noise_old = s[6], grad_noise_old = s[7];
    const float pos_noise_sum = noise_old + noise;
    const float grad_noise_sum = grad_noise_old + grad_noise;
    if ((double)grad_noise_old > 0.5 || (double)grad_noise_old > 0.6) {
        ;
    } else {
        pos_new[0] = (noise * pos[0] + noise_old * pos_new[0]) / pos_noise_sum;
        pos_new[1] = (noise * pos[1] + noise_old * pos_new[1]) / pos_noise_sum;
        pos_new[2] = (noise * pos[2] + noise_old * pos_new[2]) / pos_noise_sum;
        const float d2 = (pos[0] - pos_new[0]) * (pos[0] - pos_new[0]) + (pos[1] - pos_new[1]) * (pos[1] - pos_new[1]) +
                         (pos[2] - pos_new[2]) * (pos[2] - pos_new[2]);
        const float dist2 = (float)(0.5 * (double)sqrtf(d2));
        float axis[3];
        axis[0] = grad_new[1] * grad[2] - grad_new[2] * grad[1];
        axis[1] = -grad_new[0] * grad[2] + grad_new[2] * grad[0];
        axis[2] = grad_new[0] * grad[1] - grad_new[1] * grad[0];
        float ang = (float)acos((double)(grad_new[0] * grad[0] + grad_new[1] * grad[1] + grad_new[2] * grad[2]));
        ang = ang * noise / pos_noise_sum;
        float q[4] = {1.0f, 0.0f, 0.0f, 0.0f};
        if (ang > 1 - 6) {   // sic: always true unless NaN (GPisMap3.cpp:515)
            q[0] = (float)cos((double)ang / 2.0);
            const float sina = (float)sin((double)ang / 2.0);
            q[1] = axis[0] * sina; q[2] = axis[1] * sina; q[3] = axis[2] * sina;
        }
        float Rot[9];
        quat2dcm_dev(q, Rot);
        grad_new[0] = Rot[0] * grad[0] + Rot[1] * grad[1] + Rot[2] * grad[2];
        grad_new[1] = Rot[3] * grad[0] + Rot[4] * grad[1] + Rot[5] * grad[2];
        grad_new[2] = Rot[6] * grad[0] + Rot[7] * grad[1] + Rot[8] * grad[2];
        grad_noise = fminf((float)1.0, fmaxf(grad_noise * grad_noise_old / grad_noise_sum + dist2, P.map_noise_param));
        noise = fmaxf((noise * noise_old / pos_noise_sum + dist2), P.map_noise_param);
    }
    action[i] = 2;
    pos_out[3 * (size_t)i] = pos_new[0]; pos_out[3 * (size_t)i + 1] = pos_new[1]; pos_out[3 * (size_t)i + 2] = pos_new[2];
    grad_out[3 * (size_t)i] = grad_new[0]; grad_out[3 * (size_t)i + 1] = grad_new[1]; grad_out[3 * (size_t)i + 2] = grad_new[2];
    noise_out[i] = noise; gnoise_out[i] = grad_noise;
}

}  // namespace gpis
