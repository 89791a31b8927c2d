// Drop-in GPisMap (2D, include/gpismap/GPisMap.h). Host-side laser-scan pipeline with the
// reference's semantics (cpp/src/GPisMap.cpp); GP work goes through the C ABI:
//   regressObs -> gpis_obs_train_1d, gpo->test -> gpis_obs_test (batched per step of the surface
//   walk), updateGPs -> gpis_leaves_update, test -> gpis_query.
#include "gpismap/GPisMap.h"

#include <cmath>
#include <cstdio>

#include "map_core.hpp"

using namespace gpismap_host;
namespace gd = gpismap_defaults;

namespace {
inline bool isRangeValid(float r) { return ((double)r < gd::kMaxRange2) && ((double)r > gd::kMinRange2); }   // GPisMap.cpp:34-37
inline void polar2Cart(float a, float r, float& x, float& y) {   // GPisMap.cpp:49-54
    x = (float)((double)r * std::cos((double)a));
    y = (float)((double)r * std::sin((double)a));
}
inline void cart2polar(float x, float y, float& a, float& r) {   // GPisMap.cpp:55-60
    a = (float)std::atan2((double)y, (double)x);
    r = (float)std::sqrt((double)(x * x + y * y));
}
}  // namespace

struct GPisMap::Impl {
    GPisMapParam setting;
    MapCore<2> core;
    GPisMapTiming timing{};
    std::vector<float> obs_theta, obs_range, obs_f, obs_xylocal, obs_xyglobal;
    std::vector<float> pose_tr, pose_R;
    int obs_numdata = 0;
    float range_obs_max = 0.f;
    bool obs_ready = false;

    explicit Impl(const GPisMapParam& p)
        : setting(p),
          core(TreeParam((float)gd::kTree2MinHalf, (float)gd::kTree2MaxHalf, (float)gd::kTree2InitRootHalf,
                         (float)gd::kTree2ClusterHalf, 1e-3f, true),
               (float)gd::kRtimes2, 0),
          pose_tr(2), pose_R(4) {}

    bool ensure_ctx() {
        gpis_config cfg;
        gpis_config_default(&cfg, 2);
        cfg.map_scale = setting.map_scale_param;
        cfg.map_noise = setting.map_noise_param;
        cfg.search_half = (float)((double)setting.map_scale_param * 4.0);   // GPisMap.cpp:680
        return core.ensure_ctx(cfg);
    }
    void obs_test(const std::vector<float>& a, std::vector<float>& val, std::vector<float>& var) {
        const int m = (int)a.size();
        val.assign(m, 0.f);
        var.assign(m, 0.f);
        if (m > 0) gpis_obs_test(core.ctx, a.data(), 1, m, val.data(), var.data());
    }
    bool preproData(float* datax, float* dataf, int N, std::vector<float>& pose);
    bool regressObs();
    void updateMapPoints();
    void evalPoints();

    struct ReEval {
        int sample;
        bool alive1, walking;
        float oc, abs_oc, dx, x_new[2], r_new, grad_loc[2];
    };
    void reeval_run(const std::vector<int>& ids, std::vector<ReEval>& st, std::vector<float>& rinv0, std::vector<float>& var);
    void reeval_apply(const ReEval& e, const float* rinv0, const float* var);
};

// ------------------------------------------------------------------ preproData (GPisMap.cpp:105-149)
bool GPisMap::Impl::preproData(float* datax, float* dataf, int N, std::vector<float>& pose) {
    if (datax == 0 || dataf == 0 || N < 1) return false;
    obs_theta.clear(); obs_range.clear(); obs_f.clear(); obs_xylocal.clear(); obs_xyglobal.clear();
    range_obs_max = 0.0f;
    if (pose.size() != 6) return false;
    std::copy(pose.begin(), pose.begin() + 2, pose_tr.begin());
    std::copy(pose.begin() + 2, pose.end(), pose_R.begin());
    obs_numdata = 0;
    for (int k = 0; k < N; k++) {
        float xloc = 0.0f, yloc = 0.0f;
        if (isRangeValid(dataf[k])) {
            if (range_obs_max < dataf[k]) range_obs_max = dataf[k];
            obs_theta.push_back(datax[k]);
            obs_range.push_back(dataf[k]);
            obs_f.push_back((float)(1.0 / (double)std::sqrt(dataf[k])));
            polar2Cart(datax[k], dataf[k], xloc, yloc);
            obs_xylocal.push_back(xloc);
            obs_xylocal.push_back(yloc);
            xloc += setting.sensor_offset[0];
            yloc += setting.sensor_offset[1];
            obs_xyglobal.push_back(pose_R[0] * xloc + pose_R[2] * yloc + pose_tr[0]);
            obs_xyglobal.push_back(pose_R[1] * xloc + pose_R[3] * yloc + pose_tr[1]);
            obs_numdata++;
        }
    }
    return obs_numdata > 1;
}

bool GPisMap::Impl::regressObs() {   // GPisMap.cpp:169-179
    if (gpis_obs_train_1d(core.ctx, obs_theta.data(), obs_f.data(), obs_numdata) != GPIS_OK) {
        std::fprintf(stderr, "gpismap_b200: gpis_obs_train_1d failed: %s\n", gpis_last_error(core.ctx));
        return false;
    }
    obs_ready = true;
    return true;
}

// ------------------------------------------------------------------ reEvalPoints (GPisMap.cpp:235-455)
// The observation tests of the reference's per-sample loop, regrouped into batches: one for the
// projection, one per step of the surface walk (2D re-tests at the MOVED point, GPisMap.cpp:294-296),
// one for the four finite-difference probes. Inputs never depend on the tree state.
void GPisMap::Impl::reeval_run(const std::vector<int>& ids, std::vector<ReEval>& st, std::vector<float>& rinv0p,
                               std::vector<float>& varp) {
    st.clear();
    std::vector<float> a, rinv0, var;
    std::vector<float> rr;
    for (int s : ids) {
        ReEval e{};
        e.sample = s;
        const Sample<2>& sm = core.tree->sample(s);
        float x_loc = pose_R[0] * (sm.pos[0] - pose_tr[0]) + pose_R[1] * (sm.pos[1] - pose_tr[1]);
        float y_loc = pose_R[2] * (sm.pos[0] - pose_tr[0]) + pose_R[3] * (sm.pos[1] - pose_tr[1]);
        x_loc -= setting.sensor_offset[0];
        y_loc -= setting.sensor_offset[1];
        float ang, r;
        cart2polar(x_loc, y_loc, ang, r);
        e.x_new[0] = x_loc; e.x_new[1] = y_loc; e.r_new = r;
        st.push_back(e);
        a.push_back(ang);
        rr.push_back(r);
    }
    obs_test(a, rinv0, var);
    for (size_t i = 0; i < st.size(); ++i) {
        ReEval& e = st[i];
        if (var[i] > setting.obs_var_thre) continue;
        const float r = rr[i];
        const float oc = occ_test((float)(1.0 / (double)std::sqrt(r)), rinv0[i], (float)((double)r * 30.0));
        if ((double)oc < -0.1) continue;
        const Sample<2>& sm = core.tree->sample(e.sample);
        e.grad_loc[0] = pose_R[0] * sm.grad[0] + pose_R[1] * sm.grad[1];
        e.grad_loc[1] = pose_R[2] * sm.grad[0] + pose_R[3] * sm.grad[1];
        e.alive1 = true;
        e.oc = oc; e.abs_oc = (float)std::fabs((double)oc); e.dx = setting.delx;
        e.walking = (double)e.abs_oc > 0.02;
    }
    // surface walk, GPisMap.cpp:275-313
    for (int it = 0; it < 10; ++it) {
        std::vector<int> who;
        a.clear();
        for (size_t i = 0; i < st.size(); ++i) {
            ReEval& e = st[i];
            if (!e.alive1 || !e.walking) continue;
            if (e.oc < 0) { e.x_new[0] += e.grad_loc[0] * e.dx; e.x_new[1] += e.grad_loc[1] * e.dx; }
            else          { e.x_new[0] -= e.grad_loc[0] * e.dx; e.x_new[1] -= e.grad_loc[1] * e.dx; }
            float ang;
            cart2polar(e.x_new[0], e.x_new[1], ang, e.r_new);
            a.push_back(ang);
            who.push_back((int)i);
        }
        if (who.empty()) break;
        obs_test(a, rinv0, var);
        for (size_t j = 0; j < who.size(); ++j) {
            ReEval& e = st[who[j]];
            if (var[j] > setting.obs_var_thre) { e.walking = false; continue; }
            const float oc_new = occ_test((float)(1.0 / (double)std::sqrt(e.r_new)), rinv0[j], (float)((double)e.r_new * 30.0));
            const float abs_oc_new = (float)std::fabs((double)oc_new);
            if ((double)abs_oc_new < 0.02 || (double)e.oc < -0.1) { e.walking = false; continue; }
            else if ((double)(e.oc * oc_new) < 0.0) e.dx = (float)(0.5 * (double)e.dx);
            else e.dx = (float)(1.1 * (double)e.dx);
            e.abs_oc = abs_oc_new;
            e.oc = oc_new;
            if (!((double)e.abs_oc > 0.02)) e.walking = false;
        }
    }
    // four finite-difference probes, GPisMap.cpp:315-343
    static const float Xp[4] = {1.0f, -1.0f, 0.0f, 0.0f};
    static const float Yp[4] = {0.0f, 0.0f, 1.0f, -1.0f};
    a.clear();
    for (const ReEval& e : st) {
        if (!e.alive1) continue;
        for (int i = 0; i < 4; i++) {
            const float X = e.x_new[0] + setting.delx * Xp[i];
            const float Y = e.x_new[1] + setting.delx * Yp[i];
            float ang, r_;
            cart2polar(X, Y, ang, r_);
            a.push_back(ang);
        }
    }
    obs_test(a, rinv0p, varp);
}

void GPisMap::Impl::reeval_apply(const ReEval& e, const float* rinv0, const float* var) {
    static const float Xp[4] = {1.0f, -1.0f, 0.0f, 0.0f};
    static const float Yp[4] = {0.0f, 0.0f, 1.0f, -1.0f};
    auto* tree = core.tree;
    float occ[4] = {-1.0f, -1.0f, -1.0f, -1.0f};
    float occ_mean = 0.0f, r0_mean = 0.0f, r0_sqr_sum = 0.0f;
    float last_var = 0.f;
    for (int i = 0; i < 4; i++) {
        const float X = e.x_new[0] + setting.delx * Xp[i];
        const float Y = e.x_new[1] + setting.delx * Yp[i];
        float ang, r_;
        cart2polar(X, Y, ang, r_);
        last_var = var[i];
        if (var[i] > setting.obs_var_thre) break;
        occ[i] = occ_test((float)(1.0 / (double)std::sqrt(r_)), rinv0[i], (float)((double)r_ * 30.0));
        occ_mean = (float)((double)occ_mean + 0.25 * (double)occ[i]);
        const float r0 = (float)(1.0 / (double)(rinv0[i] * rinv0[i]));
        r0_sqr_sum += r0 * r0;
        r0_mean = (float)((double)r0_mean + 0.25 * (double)r0);
    }
    if (last_var > setting.obs_var_thre) return;

    Sample<2>& old = tree->sample(e.sample);
    const float pos[2] = {old.pos[0], old.pos[1]};
    const float grad[2] = {old.grad[0], old.grad[1]};
    float gnl[2];
    gnl[0] = (occ[0] - occ[1]) / setting.delx;
    gnl[1] = (occ[2] - occ[3]) / setting.delx;
    const float norm_grad_new = std::sqrt(gnl[0] * gnl[0] + gnl[1] * gnl[1]);
    if ((double)norm_grad_new < 1e-3) {
        old.pose_sig = (float)(2.0 * (double)old.pose_sig);
        old.grad_sig = (float)(2.0 * (double)old.grad_sig);
        tree->touch(e.sample);
        return;
    }
    float r_var = (float)((double)r0_sqr_sum / 3.0 - (double)(r0_mean * r0_mean) * 4.0 / 3.0);
    r_var /= setting.delx;
    float noise = 100.0f, grad_noise = 1.0f;
    const float r_new = e.r_new;
    if ((double)norm_grad_new > 1e-6) {
        gnl[0] = gnl[0] / norm_grad_new; gnl[1] = gnl[1] / norm_grad_new;
        noise = setting.min_position_noise * saturate(r_new * r_new, 1.0f, noise);
        grad_noise = saturate(std::fabs(occ_mean) + r_var, setting.min_grad_noise, grad_noise);
    } else {
        noise = setting.min_position_noise * noise;
    }
    float xn[2] = {e.x_new[0], e.x_new[1]};
    const float dist = std::sqrt(xn[0] * xn[0] + xn[1] * xn[1]);
    const float view_ang = std::max(-(xn[0] * gnl[0] + xn[1] * gnl[1]) / dist, (float)1e-1);
    const float view_ang2 = view_ang * view_ang;
    const float view_noise = (float)((double)setting.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
    noise += view_noise + e.abs_oc;
    grad_noise = (float)((double)grad_noise + 0.1 * (double)view_noise);

    float pos_new[2], grad_new[2];
    xn[0] += setting.sensor_offset[0];
    xn[1] += setting.sensor_offset[1];
    pos_new[0] = pose_R[0] * xn[0] + pose_R[2] * xn[1] + pose_tr[0];
    pos_new[1] = pose_R[1] * xn[0] + pose_R[3] * xn[1] + pose_tr[1];
    grad_new[0] = pose_R[0] * gnl[0] + pose_R[2] * gnl[1];
    grad_new[1] = pose_R[1] * gnl[0] + pose_R[3] * gnl[1];

    const float noise_old = old.pose_sig, grad_noise_old = old.grad_sig;
    const float pos_noise_sum = noise_old + noise;
    const float grad_noise_sum = grad_noise_old + grad_noise;
    if ((double)grad_noise_old > 0.5 || (double)grad_noise_old > 0.6) {
        ;
    } else {
        pos_new[0] = (noise * pos[0] + noise_old * pos_new[0]) / pos_noise_sum;
        pos_new[1] = (noise * pos[1] + noise_old * pos_new[1]) / pos_noise_sum;
        const float d2 = (pos[0] - pos_new[0]) * (pos[0] - pos_new[0]) + (pos[1] - pos_new[1]) * (pos[1] - pos_new[1]);
        const float dist2 = (float)(0.5 * (double)std::sqrt(d2));
        // rotate the old normal toward the new one by a noise-weighted angle (GPisMap.cpp:407-415)
        const float tx = grad[0] * grad_new[0] + grad[1] * grad_new[1];
        const float ty = -grad[1] * grad_new[0] + grad[0] * grad_new[1];
        const float ang_dist = (float)(std::atan2((double)ty, (double)tx) * (double)noise / (double)pos_noise_sum);
        const float sina = (float)std::sin((double)ang_dist);
        const float cosa = (float)std::cos((double)ang_dist);
        grad_new[0] = cosa * grad[0] - sina * grad[1];
        grad_new[1] = sina * grad[0] + cosa * grad[1];
        grad_noise = std::min((float)1.0, std::max(grad_noise * grad_noise_old / grad_noise_sum + dist2, setting.map_noise_param));
        noise = std::max((noise * noise_old / pos_noise_sum + dist2), setting.map_noise_param);
    }
    std::vector<int> freed;
    tree->remove_tracked(e.sample, freed);
    core.drop_freed(freed);
    if ((double)noise > 1.0 && (double)grad_noise > 0.61) return;
    std::vector<int> touched;
    const int s = core.try_insert(pos_new, touched);
    if (s < 0) return;
    Sample<2>& sm = tree->sample(s);
    sm.val = -setting.fbias; sm.pose_sig = noise; sm.grad_sig = grad_noise;
    sm.grad[0] = grad_new[0]; sm.grad[1] = grad_new[1];
    core.activate(touched);
}

// ------------------------------------------------------------------ updateMapPoints (GPisMap.cpp:181-233)
void GPisMap::Impl::updateMapPoints() {
    if (!core.tree || !obs_ready) return;
    auto* tree = core.tree;
    std::vector<int> quads;
    tree->query_clusters(pose_tr.data(), range_obs_max, quads);
    if (quads.empty()) return;
    const float r2 = range_obs_max * range_obs_max;
    std::vector<LeafHandle> inview;
    for (int cid : quads) {
        const auto& n = tree->cell(cid);
        const float l = n.half;
        const float sqr_range = (n.c[0] - pose_tr[0]) * (n.c[0] - pose_tr[0]) + (n.c[1] - pose_tr[1]) * (n.c[1] - pose_tr[1]);
        if (sqr_range > (r2 + 2 * l * l)) continue;
        int within_angle = 0;
        for (int k = 0; k < 4; ++k) {   // NW, NE, SW, SE (quadtree.h:61-64)
            const float ex = (k & 1) ? n.hi[0] : n.lo[0];
            const float ey = (k & 2) ? n.lo[1] : n.hi[1];
            float x_loc = pose_R[0] * (ex - pose_tr[0]) + pose_R[1] * (ey - pose_tr[1]);
            float y_loc = pose_R[2] * (ex - pose_tr[0]) + pose_R[3] * (ey - pose_tr[1]);
            x_loc -= setting.sensor_offset[0];
            y_loc -= setting.sensor_offset[1];
            float ang = 0.0f, r = 0.0f;
            cart2polar(x_loc, y_loc, ang, r);
            within_angle += int((ang > setting.angle_obs_limit[0]) && (ang < setting.angle_obs_limit[1]));
        }
        if (within_angle == 0) continue;
        inview.push_back(LeafHandle{cid, n.gen});
    }
    std::vector<int> ids_all;
    for (const LeafHandle& h : inview) tree->collect_samples(h.cell, ids_all);
    std::vector<ReEval> st, st2;
    std::vector<float> rinv0, var, rinv0b, varb;
    reeval_run(ids_all, st, rinv0, var);
    std::vector<int> pre_index(tree->num_samples(), -1), pre_probe(st.size(), -1);
    {
        int probe = 0;
        for (size_t i = 0; i < st.size(); ++i) {
            pre_index[st[i].sample] = (int)i;
            if (st[i].alive1) { pre_probe[i] = probe; probe += 4; }
        }
    }
    std::vector<int> ids, fresh;
    for (const LeafHandle& h : inview) {
        if (!tree->cell_alive(h.cell, h.gen)) continue;
        ids.clear();
        tree->collect_samples(h.cell, ids);
        fresh.clear();
        for (int s : ids) if (s >= (int)pre_index.size() || pre_index[s] < 0) fresh.push_back(s);
        st2.clear();
        if (!fresh.empty()) reeval_run(fresh, st2, rinv0b, varb);
        size_t fi = 0; int fprobe = 0;
        for (int s : ids) {
            if (s < (int)pre_index.size() && pre_index[s] >= 0) {
                const int i = pre_index[s];
                if (st[i].alive1) reeval_apply(st[i], &rinv0[pre_probe[i]], &var[pre_probe[i]]);
            } else {
                const ReEval& e = st2[fi++];
                if (e.alive1) { reeval_apply(e, &rinv0b[fprobe], &varb[fprobe]); fprobe += 4; }
            }
        }
    }
}

// ------------------------------------------------------------------ evalPoints (GPisMap.cpp:466-572)
void GPisMap::Impl::evalPoints() {
    if (!core.tree || obs_numdata < 1) return;
    auto* tree = core.tree;
    static const float Xp[4] = {1.0f, -1.0f, 0.0f, 0.0f};
    static const float Yp[4] = {0.0f, 0.0f, 1.0f, -1.0f};
    const int K = obs_numdata;
    std::vector<float> rinv0c, varc, ap, rinv0p, varp;
    obs_test(obs_theta, rinv0c, varc);
    std::vector<int> probe_of(K, -1);
    std::vector<float> rp;
    int np = 0;
    for (int k = 0; k < K; ++k) {
        if (varc[k] > setting.obs_var_thre) continue;
        probe_of[k] = np; np += 4;
        for (int i = 0; i < 4; i++) {
            const float X = obs_xylocal[2 * k] + setting.delx * Xp[i];
            const float Y = obs_xylocal[2 * k + 1] + setting.delx * Yp[i];
            float a, r;
            cart2polar(X, Y, a, r);
            ap.push_back(a);
            rp.push_back(r);
        }
    }
    obs_test(ap, rinv0p, varp);
    std::vector<int> touched, freed;
    for (int k = 0; k < K; k++) {
        const int k2 = 2 * k;
        if (varc[k] > setting.obs_var_thre) continue;
        const int s = core.try_insert(&obs_xyglobal[k2], touched);
        if (s < 0) continue;
        float occ[4] = {-1.0f, -1.0f, -1.0f, -1.0f};
        float occ_mean = 0.0f;
        const int p0 = probe_of[k];
        bool failed = false;
        for (int i = 0; i < 4; i++) {
            if (varp[p0 + i] > setting.obs_var_thre) { failed = true; break; }
            const float r = rp[p0 + i];
            occ[i] = occ_test((float)(1.0 / (double)std::sqrt(r)), rinv0p[p0 + i], (float)((double)r * 30.0));
            occ_mean = (float)((double)occ_mean + 0.25 * (double)occ[i]);
        }
        if (failed) {
            freed.clear();
            tree->remove_plain(s, freed);
            core.drop_freed(freed);
            continue;
        }
        float noise = 100.0f, grad_noise = 1.00f;
        float grad[2];
        grad[0] = (occ[0] - occ[1]) / setting.delx;
        grad[1] = (occ[2] - occ[3]) / setting.delx;
        float norm_grad = grad[0] * grad[0] + grad[1] * grad[1];
        if ((double)norm_grad > 1e-6) {
            norm_grad = std::sqrt(norm_grad);
            const float glx = grad[0] / norm_grad, gly = grad[1] / norm_grad;
            grad[0] = pose_R[0] * glx + pose_R[2] * gly;
            grad[1] = pose_R[1] * glx + pose_R[3] * gly;
            noise = setting.min_position_noise * (saturate(obs_range[k] * obs_range[k], 1.0f, noise));
            grad_noise = saturate(std::fabs(occ_mean), setting.min_grad_noise, grad_noise);
            const float dist = std::sqrt(obs_xylocal[k2] * obs_xylocal[k2] + obs_xylocal[k2 + 1] * obs_xylocal[k2 + 1]);
            const float view_ang = std::max(-(obs_xylocal[k2] * glx + obs_xylocal[k2 + 1] * gly) / dist, (float)1e-1);
            const float view_ang2 = view_ang * view_ang;
            const float view_noise = (float)((double)setting.min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
            noise += view_noise;
        }
        Sample<2>& sm = tree->sample(s);
        sm.val = -setting.fbias; sm.pose_sig = noise; sm.grad_sig = grad_noise;
        sm.grad[0] = grad[0]; sm.grad[1] = grad[1];
        core.activate(touched);
    }
}

// ------------------------------------------------------------------ public API
GPisMap::GPisMap() : d(new Impl(GPisMapParam())) {}
GPisMap::GPisMap(GPisMapParam par) : d(new Impl(par)) {}
GPisMap::~GPisMap() { delete d; }
void GPisMap::reset() { d->core.reset(); d->obs_numdata = 0; d->obs_ready = false; }
void GPisMap::setDevice(int dev) { d->core.device_ = dev; }
const GPisMapTiming& GPisMap::lastTiming() const { return d->timing; }
void* GPisMap::cabiContext() { d->ensure_ctx(); return d->core.ctx; }

void GPisMap::update(float* datax, float* dataf, int N, std::vector<float>& pose) {
    GPisMapTiming& T = d->timing;
    T = GPisMapTiming{};
    double t0 = now_s();
    const bool ok = d->preproData(datax, dataf, N, pose);
    double t1 = now_s();
    T.phase[0] = t1 - t0;
    T.valid_beams = d->obs_numdata;
    if (!ok) return;
    if (!d->ensure_ctx()) return;
    const bool reg = d->regressObs();
    double t2 = now_s();
    T.phase[1] = t2 - t1;
    if (!reg) return;
    d->updateMapPoints();
    double t3 = now_s();
    T.phase[2] = t3 - t2;
    d->core.ensure_tree();   // addNewMeas, GPisMap.cpp:457-464
    d->evalPoints();
    double t4 = now_s();
    T.phase[3] = t4 - t3;
    T.active_leaves = (int)d->core.active.size();
    d->core.train_active();   // the 2D reference divides by zero on an empty update set (GPisMap.cpp:625-632); not replicated
    T.phase[4] = now_s() - t4;
    T.trained_leaves = d->core.last_trained;
    T.train_kernel_ms = d->core.last_train_ms;
}

bool GPisMap::test(float* x, int dim, int leng, float* res) {
    if (x == 0 || dim != 2 || leng < 1) return false;   // GPisMap.cpp:766-767
    return d->core.query(x, leng, res);
}
void GPisMap::getAllPoints(std::vector<float>& pos) { d->core.all_points(pos); }
void GPisMap::getAllSamples(std::vector<float>& s) { d->core.all_samples(s); }
void GPisMap::getLeaves(std::vector<float>& c, std::vector<int>& n) { d->core.leaves(c, n); }
int GPisMap::numLeaves() { std::vector<float> c; std::vector<int> n; d->core.leaves(c, n); return (int)n.size(); }
int GPisMap::insertSamples(const float* s, int n) {
    if (!d->ensure_ctx()) return 0;
    return d->core.insert_samples(s, n);
}
int GPisMap::activateAll() { return d->core.activate_all(); }
int GPisMap::trainActive() {
    if (!d->ensure_ctx()) return 0;
    const int n = (int)d->core.active.size();
    d->core.train_active();
    return n;
}
