// Drop-in GPisMap3 (include/gpismap/GPisMap3.h). Host-side sensor pipeline with the reference's
// semantics (cpp/src/GPisMap3.cpp); every dense-LA / GP step goes through the C ABI to the GPU:
//   regressObs      -> gpis_obs_train_2d                       (ObsGP2D::train)
//   gpo->test(...)  -> gpis_obs_test, batched per frame         (ObsGP2D::test, one call per point in the reference)
//   updateGPs       -> gpis_leaves_update                       (OnGPIS::train per dirty leaf)
//   test            -> gpis_query                               (test_kernel)
// The serial loops that mutate the octree consume pre-computed batches of observation-GP
// predictions: the inputs of those predictions do not depend on the tree state.
#include "gpismap/GPisMap3.h"

#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "map_core.hpp"

using namespace gpismap_host;
namespace gd = gpismap_defaults;

namespace {
inline bool isRangeValid(float r) { return ((double)r < gd::kMaxRange3) && ((double)r > gd::kMinRange3); }   // GPisMap3.cpp:33-36

// quat2dcm (GPisMap3.cpp:48-64)
std::array<float, 9> quat2dcm(const float q[4]) {
    std::array<float, 9> dcm;
    dcm[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
    dcm[1] = (float)(2.0 * (double)(q[1] * q[2] + q[0] * q[3]));
    dcm[2] = (float)(2.0 * (double)(q[1] * q[3] - q[0] * q[2]));
    dcm[3] = (float)(2.0 * (double)(q[1] * q[2] - q[0] * q[3]));
    dcm[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
    dcm[5] = (float)(2.0 * (double)(q[0] * q[1] + q[2] * q[3]));
    dcm[6] = (float)(2.0 * (double)(q[1] * q[3] + q[0] * q[2]));
    dcm[7] = (float)(2.0 * (double)(q[2] * q[3] - q[0] * q[1]));
    dcm[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    return dcm;
}
}  // namespace

struct GPisMap3::Impl {
    GPisMap3Param setting;
    camParam cam;
    MapCore<3> core;
    GPisMap3Timing timing{};
    GPisMap3Tuning tuning{};
    float u_obs_limit[2] = {0.f, 0.f}, v_obs_limit[2] = {0.f, 0.f};
    std::vector<float> vu_grid;
    std::vector<float> obs_valid_u, obs_valid_v, obs_zinv, obs_valid_xyzlocal, obs_valid_xyzglobal;
    std::vector<float> pose_tr, pose_R;
    int obs_numdata = 0;
    float range_obs_max = 0.f;
    bool obs_ready = false;

    Impl(const GPisMap3Param& p, const camParam& c)
        : setting(p), cam(c),
          core(TreeParam((float)gd::kTree3MinHalf, (float)gd::kTree3MaxHalf, (float)gd::kTree3InitRootHalf,
                         (float)gd::kTree3ClusterHalf, 1e-6f, false),
               (float)gd::kRtimes3, 0),
          pose_tr(3), pose_R(9) {}

    bool ensure_ctx() {
        gpis_config cfg;
        gpis_config_default(&cfg, 3);
        cfg.map_scale = setting.map_scale_param;
        cfg.map_noise = setting.map_noise_param;
        cfg.cluster_half = tuning.tree_cluster_half;
        cfg.search_half = tuning.tree_cluster_half * 3.0;      // GPisMap3.cpp:811  C_leng*3.0
        const bool fresh = core.ctx == nullptr;
        if (!core.ensure_ctx(cfg)) return false;
        if (fresh) {
            // Leaf training overlaps the next frame's host work (include/gpis_b200.h, gpis_set_train_mode): nothing in
            // update() reads the leaf GPs, and test() waits for them inside gpis_query. GPIS_TRAIN_MODE=0 restores the
            // reference's timing (update() returns when every GP is trained).
            const char* e = std::getenv("GPIS_TRAIN_MODE");
            const int mode = e ? std::atoi(e) : (device_frame ? 3 : 1);
            train_mode = mode;
            if (gpis_set_train_mode(core.ctx, mode) != GPIS_OK)
                std::fprintf(stderr, "gpismap_b200: gpis_set_train_mode(%d) failed: %s\n", mode, gpis_last_error(core.ctx));
        }
        return true;
    }

    // f-3 + f-1 on the device (gpis_frame_eval): preprocData + regressObs + the numerics of evalPoints in one call;
    // GPIS_HOST_FRAME=1 keeps the host implementation of those steps (identical results)
    int train_mode = 0;
    bool device_frame = std::getenv("GPIS_HOST_FRAME") == nullptr || std::atoi(std::getenv("GPIS_HOST_FRAME")) == 0;
    std::vector<int32_t> pre_status;
    std::vector<float> pre_grad, pre_noise, pre_gnoise;
    bool frameEval(float* dataz, int N, std::vector<float>& pose);
    bool preprocData(float* dataz, int N, std::vector<float>& pose);
    bool regressObs();
    void updateMapPoints();
    void evalPoints();

    // ---- batched ObsGP2D::test
    void obs_test(const std::vector<float>& vu, std::vector<float>& val, std::vector<float>& var) {
        const int m = (int)vu.size() / 2;
        val.assign(m, 0.f);
        var.assign(m, 0.f);
        if (m > 0) { ProfScope ps(0); gpis_obs_test(core.ctx, vu.data(), 2, m, val.data(), var.data()); }
    }

    struct ReEval {            // per-sample state between the two observation batches
        int sample;
        bool alive1;           // survived the first test + occupancy gate
        float loc[3], oc, abs_oc, x_new[3];
    };
    void reeval_stage1(const std::vector<int>& ids, std::vector<ReEval>& st);
    void reeval_stage2(std::vector<ReEval>& st, std::vector<float>& rinv0, std::vector<float>& var);
    struct ReOut { int action; float pos_new[3], grad_new[3], noise, grad_noise; };   // 0 nothing, 1 double the sigmas, 2 replace
    ReOut reeval_compute(const ReEval& e, const float* rinv0, const float* var) const;
    int reeval_commit(const ReEval& e, const ReOut& o);   // id of the fused sample it inserted, -1 if none
    std::vector<int> scratch_freed, scratch_touched;   // reeval_commit runs once per in-view sample
    std::vector<int> pre_index_buf;
    // second-generation re-evaluations computed ahead of the serial pass (updateMapPoints)
    std::vector<int> spec_of_pre;      // first-batch index -> index into spec_*, -1 = none
    std::vector<ReEval> spec_st;
    std::vector<ReOut> spec_out;
    void reeval_apply(const ReEval& e, const float* rinv0, const float* var);
};

// ------------------------------------------------------------------ preprocData (GPisMap3.cpp:125-216)
bool GPisMap3::Impl::preprocData(float* dataz, int N, std::vector<float>& pose) {
    if (dataz == 0 || N < 1) return false;
    obs_valid_xyzlocal.clear(); obs_valid_xyzglobal.clear();
    obs_valid_u.clear(); obs_valid_v.clear(); obs_zinv.clear();
    range_obs_max = 0.0f;
    if (pose.size() != 12) return false;
    std::copy(pose.begin(), pose.begin() + 3, pose_tr.begin());
    std::copy(pose.begin() + 3, pose.end(), pose_R.begin());

    const int n = cam.width / setting.obs_skip;
    const int m = cam.height / setting.obs_skip;
    if (vu_grid.size() == 0) {
        if (cam.width * cam.height != N) {
            std::cout << "Error: The dimensions do not match!" << std::endl;
            return false;
        }
        vu_grid.resize(2 * (size_t)n * m);
        int col = 0, row = 0;
        for (int n_ = 0; n_ < n; n_++) {
            col = n_ * setting.obs_skip;
            for (int m_ = 0; m_ < m; m_++) {
                row = m_ * setting.obs_skip;
                const int j = 2 * (m * n_ + m_);
                vu_grid[j] = (float(row) - cam.cy) / cam.fy;
                vu_grid[j + 1] = (float(col) - cam.cx) / cam.fx;
            }
        }
        u_obs_limit[0] = -cam.cx / cam.fx;
        u_obs_limit[1] = (float(col) - cam.cx) / cam.fx;
        v_obs_limit[0] = -cam.cy / cam.fy;
        v_obs_limit[1] = (float(row) - cam.cy) / cam.fy;
    }
    obs_numdata = 0;
    for (int n_ = 0; n_ < n; n_++) {
        const int col = n_ * setting.obs_skip;
        for (int m_ = 0; m_ < m; m_++) {
            const int row = m_ * setting.obs_skip;
            const int k = col * cam.height + row;
            if ((k < N) && isRangeValid(dataz[k])) {
                const int j = 2 * (m * n_ + m_);
                if (range_obs_max < dataz[k]) range_obs_max = dataz[k];
                obs_zinv.push_back((float)(1.0 / (double)dataz[k]));
                const float u = vu_grid[j + 1], v = vu_grid[j];
                obs_valid_u.push_back(u);
                obs_valid_v.push_back(v);
                const float xloc = u * dataz[k], yloc = v * dataz[k];
                obs_valid_xyzlocal.push_back(xloc);
                obs_valid_xyzlocal.push_back(yloc);
                obs_valid_xyzlocal.push_back(dataz[k]);
                obs_valid_xyzglobal.push_back(pose_R[0] * xloc + pose_R[3] * yloc + pose_R[6] * dataz[k] + pose_tr[0]);
                obs_valid_xyzglobal.push_back(pose_R[1] * xloc + pose_R[4] * yloc + pose_R[7] * dataz[k] + pose_tr[1]);
                obs_valid_xyzglobal.push_back(pose_R[2] * xloc + pose_R[5] * yloc + pose_R[8] * dataz[k] + pose_tr[2]);
                obs_numdata++;
            } else {
                obs_zinv.push_back(-1.0f);
            }
        }
    }
    return obs_numdata > 1;
}

// ------------------------------------------------------------------ device path of Steps 0, 1 and the numerics of Step 3
bool GPisMap3::Impl::frameEval(float* dataz, int N, std::vector<float>& pose) {
    if (dataz == 0 || N < 1) return false;
    obs_valid_xyzlocal.clear(); obs_valid_xyzglobal.clear();
    obs_valid_u.clear(); obs_valid_v.clear(); obs_zinv.clear();
    range_obs_max = 0.0f;
    obs_numdata = 0;
    if (pose.size() != 12) return false;
    std::copy(pose.begin(), pose.begin() + 3, pose_tr.begin());
    std::copy(pose.begin() + 3, pose.end(), pose_R.begin());
    const int n = cam.width / setting.obs_skip;
    const int m = cam.height / setting.obs_skip;
    if (vu_grid.size() == 0) {          // GPisMap3.cpp:147-176
        if (cam.width * cam.height != N) {
            std::cout << "Error: The dimensions do not match!" << std::endl;
            return false;
        }
        vu_grid.resize(2 * (size_t)n * m);
        int col = 0, row = 0;
        for (int n_ = 0; n_ < n; n_++) {
            col = n_ * setting.obs_skip;
            for (int m_ = 0; m_ < m; m_++) {
                row = m_ * setting.obs_skip;
                const int j = 2 * (m * n_ + m_);
                vu_grid[j] = (float(row) - cam.cy) / cam.fy;
                vu_grid[j + 1] = (float(col) - cam.cx) / cam.fx;
            }
        }
        u_obs_limit[0] = -cam.cx / cam.fx;
        u_obs_limit[1] = (float(col) - cam.cx) / cam.fx;
        v_obs_limit[0] = -cam.cy / cam.fy;
        v_obs_limit[1] = (float(row) - cam.cy) / cam.fy;
    }
    gpis_frame_params fp{};
    fp.width = cam.width; fp.height = cam.height; fp.skip = setting.obs_skip;
    for (int i = 0; i < 12; ++i) fp.pose[i] = pose[i];
    fp.delx = setting.delx; fp.obs_var_thre = setting.obs_var_thre;
    fp.min_position_noise = setting.min_position_noise; fp.min_grad_noise = setting.min_grad_noise;
    fp.max_range = gd::kMaxRange3; fp.min_range = gd::kMinRange3;
    const int cap = n * m;
    obs_valid_xyzglobal.resize(3 * (size_t)cap);
    pre_status.resize(cap); pre_grad.resize(3 * (size_t)cap); pre_noise.resize(cap); pre_gnoise.resize(cap);
    int32_t K = 0;
    const int rc = gpis_frame_eval(core.ctx, dataz, N, vu_grid.data(), &fp, &K, &range_obs_max, cap, obs_valid_xyzglobal.data(),
                                   pre_status.data(), pre_grad.data(), pre_noise.data(), pre_gnoise.data());
    if (rc != GPIS_OK) {
        std::fprintf(stderr, "gpismap_b200: gpis_frame_eval failed (%d): %s\n", rc, gpis_last_error(core.ctx));
        return false;
    }
    obs_numdata = K;
    obs_ready = K > 1;
    return K > 1;
}

// ------------------------------------------------------------------ regressObs (GPisMap3.cpp:239-256)
bool GPisMap3::Impl::regressObs() {
    if (2 * obs_zinv.size() != vu_grid.size()) return false;
    const int ni = cam.height / setting.obs_skip, nj = cam.width / setting.obs_skip;
    if (ni <= 0 || nj <= 0) return false;
    if (gpis_obs_train_2d(core.ctx, vu_grid.data(), obs_zinv.data(), ni, nj) != GPIS_OK) {
        std::fprintf(stderr, "gpismap_b200: gpis_obs_train_2d failed: %s\n", gpis_last_error(core.ctx));
        return false;
    }
    obs_ready = true;
    return true;
}

// ------------------------------------------------------------------ reEvalPoints (GPisMap3.cpp:321-569)
// Stage 1: project, first observation test, occupancy gate, and the 10-step walk (which re-tests
// the SAME pixel in the reference, GPisMap3.cpp:390-393, so it needs no further GP call).
void GPisMap3::Impl::reeval_stage1(const std::vector<int>& ids, std::vector<ReEval>& st) {
    st.clear();
    std::vector<float> vu, rinv0, var;
    std::vector<int> who;
    for (int s : ids) {
        ReEval e{};
        e.sample = s; e.alive1 = false;
        const Sample<3>& sm = core.tree->sample(s);
        const float* pos = sm.pos;
        e.loc[0] = pose_R[0] * (pos[0] - pose_tr[0]) + pose_R[1] * (pos[1] - pose_tr[1]) + pose_R[2] * (pos[2] - pose_tr[2]);
        e.loc[1] = pose_R[3] * (pos[0] - pose_tr[0]) + pose_R[4] * (pos[1] - pose_tr[1]) + pose_R[5] * (pos[2] - pose_tr[2]);
        e.loc[2] = pose_R[6] * (pos[0] - pose_tr[0]) + pose_R[7] * (pos[1] - pose_tr[1]) + pose_R[8] * (pos[2] - pose_tr[2]);
        st.push_back(e);
        if ((double)e.loc[2] < 0.0) continue;
        vu.push_back(e.loc[1] / e.loc[2]);
        vu.push_back(e.loc[0] / e.loc[2]);
        who.push_back((int)st.size() - 1);
    }
    obs_test(vu, rinv0, var);
    for (size_t i = 0; i < who.size(); ++i) {
        ReEval& e = st[who[i]];
        if (var[i] > setting.obs_var_thre) continue;
        const float x_loc = e.loc[0], y_loc = e.loc[1], z_loc = e.loc[2];
        const float rinv = (float)(1.0 / (double)z_loc);
        float oc = occ_test(rinv, rinv0[i], (float)((double)z_loc * 30.0));
        if ((double)oc < -0.02) continue;
        const Sample<3>& sm = core.tree->sample(e.sample);
        float grad_loc[3];
        grad_loc[0] = pose_R[0] * sm.grad[0] + pose_R[1] * sm.grad[1] + pose_R[2] * sm.grad[2];
        grad_loc[1] = pose_R[3] * sm.grad[0] + pose_R[4] * sm.grad[1] + pose_R[5] * sm.grad[2];
        grad_loc[2] = pose_R[6] * sm.grad[0] + pose_R[7] * sm.grad[1] + pose_R[8] * sm.grad[2];
        float abs_oc = (float)std::fabs((double)oc);
        float dx = setting.delx;
        float x_new[3] = {x_loc, y_loc, z_loc};
        for (int it = 0; it < 10 && (double)abs_oc > 0.02; it++) {
            if (oc < 0) { x_new[0] += grad_loc[0] * dx; x_new[1] += grad_loc[1] * dx; x_new[2] += grad_loc[2] * dx; }
            else        { x_new[0] -= grad_loc[0] * dx; x_new[1] -= grad_loc[1] * dx; x_new[2] -= grad_loc[2] * dx; }
            // the re-test uses the original pixel: same rinv0 / var as above
            const float r_new = z_loc;
            const float oc_new = occ_test((float)(1.0 / (double)r_new), rinv0[i], (float)((double)r_new * 30.0));
            const float abs_oc_new = (float)std::fabs((double)oc_new);
            if ((double)abs_oc_new < 0.02 || (double)oc < -0.02) break;
            else if ((double)(oc * oc_new) < 0.0) dx = (float)(0.5 * (double)dx);
            else dx = (float)(1.1 * (double)dx);
            abs_oc = abs_oc_new;
            oc = oc_new;
        }
        e.alive1 = true; e.oc = oc; e.abs_oc = abs_oc;
        e.x_new[0] = x_new[0]; e.x_new[1] = x_new[1]; e.x_new[2] = x_new[2];
    }
}

// Stage 2: the six finite-difference probes around x_new (GPisMap3.cpp:413-441), one batch.
void GPisMap3::Impl::reeval_stage2(std::vector<ReEval>& st, std::vector<float>& rinv0, std::vector<float>& var) {
    static const float Xp[6] = {1.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    static const float Yp[6] = {0.0f, 0.0f, 1.0f, -1.0f, 0.0f, 0.0f};
    static const float Zp[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f, -1.0f};
    std::vector<float> vu;
    vu.reserve(st.size() * 12);
    for (const ReEval& e : st) {
        if (!e.alive1) continue;
        for (int i = 0; i < 6; i++) {
            const float X = e.x_new[0] + setting.delx * Xp[i];
            const float Y = e.x_new[1] + setting.delx * Yp[i];
            const float Z = e.x_new[2] + setting.delx * Zp[i];
            vu.push_back(Y / Z);
            vu.push_back(X / Z);
        }
    }
    obs_test(vu, rinv0, var);
}

// The rest of one reEvalPoints iteration (GPisMap3.cpp:413-567) for one sample, in two halves: the numerics
// (a pure function of the observation tests, the pose and the sample's own fields, which nobody else touches
// before its commit) and the tree mutation. updateMapPoints runs the first half for all in-view samples on a
// few host threads and the second half serially in the reference's order.
GPisMap3::Impl::ReOut GPisMap3::Impl::reeval_compute(const ReEval& e, const float* rinv0, const float* var) const {
    ReOut out{};
    out.action = 0;
    static const float Zp[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f, -1.0f};
    auto* tree = core.tree;
    const float w = (float)(1.0 / 6.0);
    float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
    float occ_mean = 0.0f, r0_mean = 0.0f, r0_sqr_sum = 0.0f;
    float r_new = e.loc[2];
    float last_var = 0.f;
    for (int i = 0; i < 6; i++) {
        const float Z = e.x_new[2] + setting.delx * Zp[i];
        r_new = Z;
        last_var = var[i];
        if (var[i] > setting.obs_var_thre) break;
        occ[i] = occ_test((float)(1.0 / (double)r_new), rinv0[i], (float)((double)r_new * 30.0));
        occ_mean += w * occ[i];
        const float r0 = (float)(1.0 / (double)rinv0[i]);
        r0_sqr_sum += r0 * r0;
        r0_mean += w * r0;
    }
    if (last_var > setting.obs_var_thre) return out;

    const Sample<3>& old = tree->sample(e.sample);
    const float pos[3] = {old.pos[0], old.pos[1], old.pos[2]};
    const float grad[3] = {old.grad[0], old.grad[1], old.grad[2]};
    float gnl[3];
    gnl[0] = (occ[0] - occ[1]) / setting.delx;
    gnl[1] = (occ[2] - occ[3]) / setting.delx;
    gnl[2] = (occ[4] - occ[5]) / setting.delx;
    const float norm_grad_new = std::sqrt(gnl[0] * gnl[0] + gnl[1] * gnl[1] + gnl[2] * gnl[2]);
    if ((double)norm_grad_new < 1e-3) {   // uncertainty increased (GPisMap3.cpp:451-454): applied at commit
        out.action = 1;
        return out;
    }
    float r_var = (float)((double)r0_sqr_sum / 5.0 - (double)(r0_mean * r0_mean) * 6.0 / 5.0);
    r_var /= setting.delx;
    float noise = 100.0f, grad_noise = 1.0f;
    if ((double)norm_grad_new > 1e-6) {
        gnl[0] = gnl[0] / norm_grad_new; gnl[1] = gnl[1] / norm_grad_new; gnl[2] = gnl[2] / norm_grad_new;
        noise = setting.min_position_noise * saturate(r_new * r_new, 1.0f, noise);
        grad_noise = saturate(std::fabs(occ_mean) + r_var, setting.min_grad_noise, grad_noise);
    } else {
        noise = setting.min_position_noise * noise;
    }
    const float* xn = e.x_new;
    const float dist = std::sqrt(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
    const float view_ang = std::max(-(xn[0] * gnl[0] + xn[1] * gnl[1] + xn[2] * gnl[2]) / dist, (float)1e-1);
    const float view_ang2 = view_ang * view_ang;
    const float view_noise = (float)((double)setting.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
    noise += view_noise + e.abs_oc;
    grad_noise = (float)((double)grad_noise + 0.1 * (double)view_noise);

    float pos_new[3], grad_new[3];
    pos_new[0] = pose_R[0] * xn[0] + pose_R[3] * xn[1] + pose_R[6] * xn[2] + pose_tr[0];
    pos_new[1] = pose_R[1] * xn[0] + pose_R[4] * xn[1] + pose_R[7] * xn[2] + pose_tr[1];
    pos_new[2] = pose_R[2] * xn[0] + pose_R[5] * xn[1] + pose_R[8] * xn[2] + pose_tr[2];
    grad_new[0] = pose_R[0] * gnl[0] + pose_R[3] * gnl[1] + pose_R[6] * gnl[2];
    grad_new[1] = pose_R[1] * gnl[0] + pose_R[4] * gnl[1] + pose_R[7] * gnl[2];
    grad_new[2] = pose_R[2] * gnl[0] + pose_R[5] * gnl[1] + pose_R[8] * gnl[2];

    const float noise_old = old.pose_sig, grad_noise_old = old.grad_sig;
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
        const float dist2 = (float)(0.5 * (double)std::sqrt(d2));
        float axis[3];
        axis[0] = grad_new[1] * grad[2] - grad_new[2] * grad[1];
        axis[1] = -grad_new[0] * grad[2] + grad_new[2] * grad[0];
        axis[2] = grad_new[0] * grad[1] - grad_new[1] * grad[0];
        float ang = (float)std::acos((double)(grad_new[0] * grad[0] + grad_new[1] * grad[1] + grad_new[2] * grad[2]));
        ang = ang * noise / pos_noise_sum;
        float q[4] = {1.0f, 0.0f, 0.0f, 0.0f};
        if (ang > 1 - 6) {   // sic: always true (GPisMap3.cpp:515)
            q[0] = (float)std::cos((double)ang / 2.0);
            const float sina = (float)std::sin((double)ang / 2.0);
            q[1] = axis[0] * sina; q[2] = axis[1] * sina; q[3] = axis[2] * sina;
        }
        const std::array<float, 9> Rot = quat2dcm(q);
        grad_new[0] = Rot[0] * grad[0] + Rot[1] * grad[1] + Rot[2] * grad[2];
        grad_new[1] = Rot[3] * grad[0] + Rot[4] * grad[1] + Rot[5] * grad[2];
        grad_new[2] = Rot[6] * grad[0] + Rot[7] * grad[1] + Rot[8] * grad[2];
        grad_noise = std::min((float)1.0, std::max(grad_noise * grad_noise_old / grad_noise_sum + dist2, setting.map_noise_param));
        noise = std::max((noise * noise_old / pos_noise_sum + dist2), setting.map_noise_param);
    }
    out.action = 2;
    for (int a = 0; a < 3; ++a) { out.pos_new[a] = pos_new[a]; out.grad_new[a] = grad_new[a]; }
    out.noise = noise; out.grad_noise = grad_noise;
    return out;
}

int GPisMap3::Impl::reeval_commit(const ReEval& e, const ReOut& o) {
    auto* tree = core.tree;
    if (o.action == 0) return -1;
    if (o.action == 1) {
        Sample<3>& old = tree->sample(e.sample);
        old.pose_sig = (float)(2.0 * (double)old.pose_sig);
        old.grad_sig = (float)(2.0 * (double)old.grad_sig);
        tree->touch(e.sample);
        return -1;
    }
    const float noise = o.noise, grad_noise = o.grad_noise;
    const float* pos_new = o.pos_new;
    const float* grad_new = o.grad_new;
    // remove the old sample (GPisMap3.cpp:536), then try the fused one
    std::vector<int>& freed = scratch_freed;
    freed.clear();
    tree->remove_tracked(e.sample, freed);
    core.drop_freed(freed);
    if ((double)noise > 1.0 && (double)grad_noise > 0.61) return -1;
    std::vector<int>& touched = scratch_touched;
    const int s = core.try_insert(pos_new, touched);
    if (s < 0) return -1;
    Sample<3>& sm = tree->sample(s);
    sm.val = -setting.fbias; sm.pose_sig = noise; sm.grad_sig = grad_noise;
    sm.grad[0] = grad_new[0]; sm.grad[1] = grad_new[1]; sm.grad[2] = grad_new[2];
    core.activate(touched);
    return s;
}

void GPisMap3::Impl::reeval_apply(const ReEval& e, const float* rinv0, const float* var) {
    reeval_commit(e, reeval_compute(e, rinv0, var));
}

// ------------------------------------------------------------------ updateMapPoints (GPisMap3.cpp:258-319)
void GPisMap3::Impl::updateMapPoints() {
    if (!core.tree || !obs_ready) return;
    auto* tree = core.tree;
    std::vector<int> oc;
    double tp0 = now_s();
    tree->query_clusters(pose_tr.data(), range_obs_max, oc);
    if (oc.empty()) return;
    const float r2 = range_obs_max * range_obs_max;
    // in-view leaves, in DFS order (range + frustum culling, GPisMap3.cpp:268-305)
    std::vector<LeafHandle> inview;
    for (int cid : oc) {
        const auto& n = tree->cell(cid);
        const float l = n.half;
        const float sqr_range = (n.c[0] - pose_tr[0]) * (n.c[0] - pose_tr[0]) + (n.c[1] - pose_tr[1]) * (n.c[1] - pose_tr[1]) +
                                (n.c[2] - pose_tr[2]) * (n.c[2] - pose_tr[2]);
        if (sqr_range > (r2 + 2 * l * l)) continue;
        // corners in the reference's order NWF,NEF,SWF,SEF,NWB,NEB,SWB,SEB (octree.h:79-86)
        int within_angle = 0;
        for (int k = 0; k < 8; ++k) {
            const float ex = (k & 1) ? n.hi[0] : n.lo[0];
            const float ey = (k & 2) ? n.lo[1] : n.hi[1];
            const float ez = (k & 4) ? n.lo[2] : n.hi[2];
            const float x_loc = pose_R[0] * (ex - pose_tr[0]) + pose_R[1] * (ey - pose_tr[1]) + pose_R[2] * (ez - pose_tr[2]);
            const float y_loc = pose_R[3] * (ex - pose_tr[0]) + pose_R[4] * (ey - pose_tr[1]) + pose_R[5] * (ez - pose_tr[2]);
            const float z_loc = pose_R[6] * (ex - pose_tr[0]) + pose_R[7] * (ey - pose_tr[1]) + pose_R[8] * (ez - pose_tr[2]);
            if (z_loc > 0) {
                const float xv = x_loc / z_loc, yv = y_loc / z_loc;
                // sic: '=' keeps only the last in-front corner's verdict (GPisMap3.cpp:298)
                within_angle = int((xv > u_obs_limit[0]) && (xv < u_obs_limit[1]) && (yv > v_obs_limit[0]) && (yv < v_obs_limit[1]));
            }
        }
        if (within_angle == 0) continue;
        inview.push_back(LeafHandle{cid, n.gen});
    }
    g_prof_s[1] += now_s() - tp0; g_prof_n[1] += 1; tp0 = now_s();
    // Pre-compute both observation batches for every sample currently under an in-view leaf.
    std::vector<int> ids_all;
    for (const LeafHandle& h : inview) tree->collect_samples(h.cell, ids_all);
    std::vector<ReEval> st;
    std::vector<float> rinv0, var;
    std::vector<ReOut> pre_out;
    // sample id -> index into the pre-computed batch; a member that only ever grows, its entries are set for the
    // in-view samples and cleared again at the end of this call (sample ids are never reused, so the array would
    // otherwise be re-created at 4 bytes per sample EVER inserted, every frame)
    std::vector<int>& pre_index = pre_index_buf;
    if (pre_index.size() < tree->num_samples()) pre_index.resize(tree->num_samples() + tree->num_samples() / 4, -1);
    std::vector<int> pre_probe;
    struct ClearMarks {
        std::vector<int>& idx; const std::vector<int>& ids;
        ~ClearMarks() { for (int s : ids) idx[s] = -1; }
    } clear_marks{pre_index, ids_all};
    if (device_frame) {
        // projection, both observation batches and the numerics of every in-view sample on the device (gpis_reeval)
        const int n = (int)ids_all.size();
        std::vector<float> smp8(8 * (size_t)n), pos_new(3 * (size_t)n), grad_new(3 * (size_t)n), noise(n), gnoise(n);
        std::vector<int32_t> act(n, -1);
        for (int i = 0; i < n; ++i) {
            const Sample<3>& sm = tree->sample(ids_all[i]);
            float* o = &smp8[8 * (size_t)i];
            o[0] = sm.pos[0]; o[1] = sm.pos[1]; o[2] = sm.pos[2];
            o[3] = sm.grad[0]; o[4] = sm.grad[1]; o[5] = sm.grad[2];
            o[6] = sm.pose_sig; o[7] = sm.grad_sig;
        }
        gpis_frame_params fp{};
        for (int i = 0; i < 3; ++i) fp.pose[i] = pose_tr[i];
        for (int i = 0; i < 9; ++i) fp.pose[3 + i] = pose_R[i];
        fp.delx = setting.delx; fp.obs_var_thre = setting.obs_var_thre;
        fp.min_position_noise = setting.min_position_noise; fp.min_grad_noise = setting.min_grad_noise;
        {
            ProfScope ps(0);
            if (n > 0 && gpis_reeval(core.ctx, n, smp8.data(), &fp, setting.map_noise_param, act.data(), pos_new.data(), grad_new.data(),
                                     noise.data(), gnoise.data()) != GPIS_OK) {
                std::fprintf(stderr, "gpismap_b200: gpis_reeval failed: %s\n", gpis_last_error(core.ctx));
                return;
            }
        }
        st.resize(n);
        pre_out.resize(n);
        for (int i = 0; i < n; ++i) {
            st[i] = ReEval{};
            st[i].sample = ids_all[i];
            st[i].alive1 = act[i] >= 0;
            pre_index[ids_all[i]] = i;
            ReOut& o = pre_out[i];
            o.action = act[i] > 0 ? act[i] : 0;
            for (int a = 0; a < 3; ++a) { o.pos_new[a] = pos_new[3 * (size_t)i + a]; o.grad_new[a] = grad_new[3 * (size_t)i + a]; }
            o.noise = noise[i]; o.grad_noise = gnoise[i];
        }
        // Second generation. A fused sample that lands in a leaf the serial pass below has not reached yet is
        // re-evaluated when that leaf comes up (the reference walks the tree as it is by then). Its re-evaluation is
        // a function of its own fields and the frame only, and those fields are known now: the fused samples whose
        // cluster-level cell changes are sent through gpis_reeval once more, in one batch, instead of a few tiny
        // observation-test calls in the middle of the serial pass (which also land beside K1 of the previous frame).
        {
            const double pitch = 2.0 * (double)core.tparam.cluster_half;
            std::vector<int> src;
            for (int i = 0; i < n; ++i) {
                if (act[i] != 2 || ((double)noise[i] > 1.0 && (double)gnoise[i] > 0.61)) continue;
                bool moved = false;
                for (int a = 0; a < 3; ++a)
                    moved = moved || std::floor((double)smp8[8 * (size_t)i + a] / pitch) != std::floor((double)pos_new[3 * (size_t)i + a] / pitch);
                if (moved) src.push_back(i);
            }
            const int n2 = (int)src.size();
            spec_of_pre.assign(n, -1);
            spec_st.assign(n2, ReEval{});
            spec_out.assign(n2, ReOut{});
            if (n2 > 0) {
                std::vector<float> s8(8 * (size_t)n2), p2(3 * (size_t)n2), g2(3 * (size_t)n2), no2(n2), gn2(n2);
                std::vector<int32_t> a2(n2, -1);
                for (int j = 0; j < n2; ++j) {
                    const int i = src[j];
                    float* o = &s8[8 * (size_t)j];
                    for (int a = 0; a < 3; ++a) { o[a] = pos_new[3 * (size_t)i + a]; o[3 + a] = grad_new[3 * (size_t)i + a]; }
                    o[6] = noise[i]; o[7] = gnoise[i];
                }
                ProfScope ps(0);
                if (gpis_reeval(core.ctx, n2, s8.data(), &fp, setting.map_noise_param, a2.data(), p2.data(), g2.data(), no2.data(), gn2.data()) == GPIS_OK) {
                    for (int j = 0; j < n2; ++j) {
                        spec_of_pre[src[j]] = j;
                        spec_st[j].alive1 = a2[j] >= 0;
                        ReOut& o = spec_out[j];
                        o.action = a2[j] > 0 ? a2[j] : 0;
                        for (int a = 0; a < 3; ++a) { o.pos_new[a] = p2[3 * (size_t)j + a]; o.grad_new[a] = g2[3 * (size_t)j + a]; }
                        o.noise = no2[j]; o.grad_noise = gn2[j];
                    }
                }   // on failure nothing is speculated: the serial pass evaluates on demand
            }
        }
        if (train_mode == 3) gpis_train_kick(core.ctx);   // this frame's batched device work is done: K1 of the previous frame may start
        g_prof_s[2] += now_s() - tp0; g_prof_n[2] += 1;
    } else {
    reeval_stage1(ids_all, st);
    reeval_stage2(st, rinv0, var);
    pre_probe.assign(st.size(), -1);
    {
        int probe = 0;
        for (size_t i = 0; i < st.size(); ++i) {
            pre_index[st[i].sample] = (int)i;
            if (st[i].alive1) { pre_probe[i] = probe; probe += 6; }
        }
    }
    g_prof_s[2] += now_s() - tp0; g_prof_n[2] += 1;
    // numerics of every pre-existing in-view sample, in parallel (see reeval_compute)
    pre_out.resize(st.size());
    {
        const int nthr = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, (int)st.size() / 2048 + 1}));
        auto work = [&](int t) {
            const size_t b = st.size() * t / nthr, e2 = st.size() * (t + 1) / nthr;
            for (size_t i = b; i < e2; ++i)
                if (st[i].alive1) pre_out[i] = reeval_compute(st[i], &rinv0[pre_probe[i]], &var[pre_probe[i]]);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    }
    ProfScope ps3(3);
    // Serial pass in the reference's order. Samples created during this pass (a fused point that
    // landed in a leaf not yet visited) are evaluated on demand, leaf by leaf.
    std::vector<int> ids, fresh;
    std::vector<ReEval> st2;
    std::vector<float> rinv0b, varb;
    const int base_id = (int)tree->num_samples();          // samples created below get ids from here on
    std::vector<int> spec_of_new;                           // (id - base_id) -> index into spec_*, -1 = none
    auto spec_of = [&](int s) { return (s >= base_id && (size_t)(s - base_id) < spec_of_new.size()) ? spec_of_new[s - base_id] : -1; };
    const bool have_spec = device_frame && !spec_of_pre.empty();
    for (const LeafHandle& h : inview) {
        if (!tree->cell_alive(h.cell, h.gen)) continue;   // freed by a collapse; dangling pointer in the reference
        ids.clear();
        tree->collect_samples(h.cell, ids);
        fresh.clear();
        for (int s : ids) if ((s >= (int)pre_index.size() || pre_index[s] < 0) && spec_of(s) < 0) fresh.push_back(s);
        st2.clear(); rinv0b.clear(); varb.clear();
        if (!fresh.empty()) { reeval_stage1(fresh, st2); reeval_stage2(st2, rinv0b, varb); }
        size_t fi = 0; int fprobe = 0;
        for (int s : ids) {
            if (s < (int)pre_index.size() && pre_index[s] >= 0) {
                const int i = pre_index[s];
                if (!st[i].alive1) continue;
                const int ns = reeval_commit(st[i], pre_out[i]);
                if (have_spec && ns >= 0 && spec_of_pre[i] >= 0) {      // remember what is already known about the new sample
                    if ((size_t)(ns - base_id) >= spec_of_new.size()) spec_of_new.resize((size_t)(ns - base_id) + 64, -1);
                    spec_of_new[ns - base_id] = spec_of_pre[i];
                }
            } else if (spec_of(s) >= 0) {
                const int j = spec_of(s);
                if (!spec_st[j].alive1) continue;
                ReEval e = spec_st[j];
                e.sample = s;
                reeval_commit(e, spec_out[j]);
            } else {
                const ReEval& e = st2[fi++];
                if (e.alive1) { reeval_apply(e, &rinv0b[fprobe], &varb[fprobe]); fprobe += 6; }
            }
        }
    }
}

// ------------------------------------------------------------------ evalPoints (GPisMap3.cpp:580-696)
void GPisMap3::Impl::evalPoints() {
    if (!core.tree || obs_numdata < 1) return;
    auto* tree = core.tree;
    if (device_frame) {
        // the numerics came from gpis_frame_eval; what is left is the reference's serial loop over the tree
        std::vector<int> touched, freed;
        ProfScope ps4(4);
        for (int k = 0; k < obs_numdata; k++) {
            if (pre_status[k] == 0) continue;                                    // GPisMap3.cpp:600-601
            const int s = core.try_insert(&obs_valid_xyzglobal[3 * (size_t)k], touched);
            if (s < 0) continue;
            if (pre_status[k] == 1) {                                            // GPisMap3.cpp:652-655
                freed.clear();
                tree->remove_plain(s, freed);
                core.drop_freed(freed);
                continue;
            }
            Sample<3>& sm = tree->sample(s);
            sm.val = -setting.fbias; sm.pose_sig = pre_noise[k]; sm.grad_sig = pre_gnoise[k];
            sm.grad[0] = pre_grad[3 * (size_t)k]; sm.grad[1] = pre_grad[3 * (size_t)k + 1]; sm.grad[2] = pre_grad[3 * (size_t)k + 2];
            core.activate(touched);
        }
        return;
    }
    static const float Xp[6] = {1.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    static const float Yp[6] = {0.0f, 0.0f, 1.0f, -1.0f, 0.0f, 0.0f};
    static const float Zp[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f, -1.0f};
    const float w = (float)(1.0 / 6.0);
    const int K = obs_numdata;
    // batch 1: the centre pixel of every valid measurement
    std::vector<float> vu(2 * (size_t)K), rinv0c, varc;
    for (int k = 0; k < K; ++k) { vu[2 * k] = obs_valid_v[k]; vu[2 * k + 1] = obs_valid_u[k]; }
    obs_test(vu, rinv0c, varc);
    // batch 2: six probes for every measurement that passed batch 1
    std::vector<int> probe_of(K, -1);
    std::vector<float> vup;
    int np = 0;
    for (int k = 0; k < K; ++k) {
        if (varc[k] > setting.obs_var_thre) continue;
        probe_of[k] = np; np += 6;
        const int k3 = 3 * k;
        for (int i = 0; i < 6; i++) {
            const float X = obs_valid_xyzlocal[k3] + setting.delx * Xp[i];
            const float Y = obs_valid_xyzlocal[k3 + 1] + setting.delx * Yp[i];
            const float Z = obs_valid_xyzlocal[k3 + 2] + setting.delx * Zp[i];
            vup.push_back(Y / Z);
            vup.push_back(X / Z);
        }
    }
    std::vector<float> rinv0p, varp;
    obs_test(vup, rinv0p, varp);

    // numerics of every measurement that passed batch 1 (occupancy probes, normal, noise terms: a pure function
    // of the observation tests and the pose), on a few host threads; the serial loop below only touches the tree
    struct NewOut { bool failed; float grad[3], noise, grad_noise; };
    std::vector<NewOut> pre(K);
    {
        auto compute = [&](int k) {
            NewOut o{};
            const int k3 = 3 * k;
            float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
            float occ_mean = 0.0f;
            const float* r0 = &rinv0p[probe_of[k]];
            const float* vr = &varp[probe_of[k]];
            o.failed = false;
            for (int i = 0; i < 6; i++) {
                if (vr[i] > setting.obs_var_thre) { o.failed = true; break; }
                const float Z = obs_valid_xyzlocal[k3 + 2] + setting.delx * Zp[i];
                occ[i] = occ_test((float)(1.0 / (double)Z), r0[i], (float)((double)Z * 30.0));
                occ_mean += w * occ[i];
            }
            if (o.failed) return o;   // GPisMap3.cpp:652-655
            float noise = 100.0f, grad_noise = 1.00f;
            float grad[3];
            grad[0] = (occ[0] - occ[1]) / setting.delx;
            grad[1] = (occ[2] - occ[3]) / setting.delx;
            grad[2] = (occ[4] - occ[5]) / setting.delx;
            float norm_grad = grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2];
            if ((double)norm_grad > 1e-6) {
                norm_grad = std::sqrt(norm_grad);
                const float glx = grad[0] / norm_grad, gly = grad[1] / norm_grad, glz = grad[2] / norm_grad;
                grad[0] = pose_R[0] * glx + pose_R[3] * gly + pose_R[6] * glz;
                grad[1] = pose_R[1] * glx + pose_R[4] * gly + pose_R[7] * glz;
                grad[2] = pose_R[2] * glx + pose_R[5] * gly + pose_R[8] * glz;
                const float* xl = &obs_valid_xyzlocal[k3];
                const float dist = std::sqrt(xl[0] * xl[0] + xl[1] * xl[1] + xl[2] * xl[2]);
                noise = setting.min_position_noise * (saturate(dist, 1.0f, noise));
                grad_noise = saturate(std::fabs(occ_mean), setting.min_grad_noise, grad_noise);
                const float view_ang = std::max(-(xl[0] * glx + xl[1] * gly + xl[2] * glz) / dist, (float)1e-1);
                const float view_ang2 = view_ang * view_ang;
                const float view_noise = (float)((double)setting.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
                noise += view_noise;
            }
            o.grad[0] = grad[0]; o.grad[1] = grad[1]; o.grad[2] = grad[2];
            o.noise = noise; o.grad_noise = grad_noise;
            return o;
        };
        const int nthr = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, K / 4096 + 1}));
        auto work = [&](int t) {
            const int b = (int)((int64_t)K * t / nthr), e = (int)((int64_t)K * (t + 1) / nthr);
            for (int k = b; k < e; ++k)
                if (!(varc[k] > setting.obs_var_thre)) pre[k] = compute(k);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }

    std::vector<int> touched, freed;
    ProfScope ps4(4);
    for (int k = 0; k < K; k++) {
        const int k3 = 3 * k;
        if (varc[k] > setting.obs_var_thre) continue;
        const int s = core.try_insert(&obs_valid_xyzglobal[k3], touched);
        if (s < 0) continue;
        const NewOut& o = pre[k];
        if (o.failed) {   // GPisMap3.cpp:652-655
            freed.clear();
            tree->remove_plain(s, freed);
            core.drop_freed(freed);
            continue;
        }
        Sample<3>& sm = tree->sample(s);
        sm.val = -setting.fbias; sm.pose_sig = o.noise; sm.grad_sig = o.grad_noise;
        sm.grad[0] = o.grad[0]; sm.grad[1] = o.grad[1]; sm.grad[2] = o.grad[2];
        core.activate(touched);
    }
}

// ------------------------------------------------------------------ public API
GPisMap3::GPisMap3() : d(new Impl(GPisMap3Param(), camParam())) {}
GPisMap3::GPisMap3(GPisMap3Param par) : d(new Impl(par, camParam())) {}
GPisMap3::GPisMap3(GPisMap3Param par, camParam c) : d(new Impl(par, c)) {}
GPisMap3::~GPisMap3() { delete d; }

void GPisMap3::reset() {
    d->core.reset();
    d->obs_numdata = 0;
    d->obs_ready = false;
}
void GPisMap3::resetCam(camParam c) {   // GPisMap3.cpp:117-123
    d->cam = c;
    d->vu_grid.clear();
}
void GPisMap3::setDevice(int dev) { d->core.device_ = dev; }
bool GPisMap3::setTuning(const GPisMap3Tuning& t) {
    if (d->core.tree || d->core.ctx) return false;   // the constants are baked into the tree and the device context
    d->tuning = t;
    d->core.tparam = TreeParam(t.tree_min_half, t.tree_max_half, t.tree_init_root_half, t.tree_cluster_half, 1e-6f, false);
    d->core.rtimes_ = t.rtimes;
    return true;
}
const GPisMap3Timing& GPisMap3::lastTiming() const { return d->timing; }
void* GPisMap3::cabiContext() { d->ensure_ctx(); return d->core.ctx; }

void GPisMap3::update(float* dataz, int N, std::vector<float>& pose) {
    GPisMap3Timing& T = d->timing;
    T = GPisMap3Timing{};
    double t0 = now_s();
    double t1, t2;
    if (d->device_frame) {
        if (!d->ensure_ctx()) return;
        const bool ok = d->frameEval(dataz, N, pose);    // Steps 0 + 1 and the numerics of Step 3, on the device
        t1 = t2 = now_s();
        T.phase[0] = t1 - t0;
        T.valid_pixels = d->obs_numdata;
        if (!ok) return;
    } else {
        const bool ok = d->preprocData(dataz, N, pose);
        t1 = now_s();
        T.phase[0] = t1 - t0;
        T.valid_pixels = d->obs_numdata;
        if (!ok) return;
        if (!d->ensure_ctx()) return;
        const bool reg = d->regressObs();          // Step 1
        t2 = now_s();
        T.phase[1] = t2 - t1;
        if (!reg) return;
    }
    d->updateMapPoints();                      // Step 2
    double t3 = now_s();
    T.phase[2] = t3 - t2;
    d->core.ensure_tree();                     // Step 3 (addNewMeas, GPisMap3.cpp:571-578)
    d->evalPoints();
    double t4 = now_s();
    T.phase[3] = t4 - t3;
    T.active_leaves = (int)d->core.active.size();
    d->core.train_active();                    // Step 4
    T.phase[4] = now_s() - t4;
    T.trained_leaves = d->core.last_trained;
    T.train_kernel_ms = d->core.last_train_ms;
}

bool GPisMap3::test(float* x, int dim, int leng, float* res) {
    if (x == 0 || dim != 3 || leng < 1) return false;   // GPisMap3.cpp:905-906
    return d->core.query(x, leng, res);
}

void GPisMap3::getAllPoints(std::vector<float>& pos) { d->core.all_points(pos); }
void GPisMap3::getAllSamples(std::vector<float>& s) { d->core.all_samples(s); }
void GPisMap3::getLeaves(std::vector<float>& c, std::vector<int>& n) { d->core.leaves(c, n); }
int GPisMap3::numLeaves() { std::vector<float> c; std::vector<int> n; d->core.leaves(c, n); return (int)n.size(); }
int GPisMap3::insertSamples(const float* s, int n) {
    if (!d->ensure_ctx()) return 0;
    return d->core.insert_samples(s, n);
}
int GPisMap3::activateAll() { return d->core.activate_all(); }
int GPisMap3::trainActive() {
    if (!d->ensure_ctx()) return 0;
    const int n = (int)d->core.active.size();
    d->core.train_active();
    return n;
}
