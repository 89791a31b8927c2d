// TEST SCAFFOLDING ONLY. A CPU stand-in for libgpis_b200.so that implements the same C ABI
// (include/gpis_b200.h) on top of the plain-C oracle (oracle/gpis_oracle.c). It exists so the
// host-side logic of the drop-in classes (tree, evalPoints / reEvalPoints heuristics, dirty-leaf
// CSR, table sync) can be validated against oracle/_ref in a container without a GPU. It is built
// into tests/_mock/ by tests/mockbuild.py, never into gpismap_b200/, and nothing in the product
// path can load it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <cstdlib>
#include <vector>

#include "gpis_b200.h"
#include "gpis_oracle.h"

struct MockLeaf {
    int32_t cell[3];
    float centre[3];
    float lo[3], hi[3];
    bool box_set = false;
    gpo_gp* gp = nullptr;
    int N = 0;
    int index = 0;
    std::vector<float> own;   // the leaf's own samples (gpis_samples_set)
};
struct gpis_ctx {
    gpis_config cfg;
    std::string err;
    std::map<std::vector<int32_t>, MockLeaf> leaves;
    int next_index = 0;
    int32_t root_min[3] = {-(1 << 19), -(1 << 19), -(1 << 19)};
    int levels = 20;
    gpo_obs* obs = nullptr;
    int obs_d = 0;
    gpis_stats st{};
};

static uint64_t spread3(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
static uint64_t dfs_key(const gpis_ctx* c, const int32_t* cell) {
    const uint32_t mask = (1u << c->levels) - 1u;
    const uint32_t lx = (uint32_t)(cell[0] - c->root_min[0]) & mask;
    const uint32_t ly = (~(uint32_t)(cell[1] - c->root_min[1])) & mask;
    const uint32_t lz = c->cfg.dim == 3 ? ((~(uint32_t)(cell[2] - c->root_min[2])) & mask) : 0u;
    return spread3(lx) | (spread3(ly) << 1) | (spread3(lz) << 2);
}

extern "C" {

int gpis_config_default(gpis_config* cfg, int dim) {
    if (!cfg || (dim != 2 && dim != 3)) return GPIS_ERR_ARG;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->dim = dim;
    if (dim == 3) { cfg->map_scale = 0.04f; cfg->map_noise = 5e-3f; cfg->cluster_half = (float)0.025; cfg->search_half = (float)0.025 * 3.0; cfg->var_thre = 0.5f; }
    else { cfg->map_scale = 1.2f; cfg->map_noise = 1e-2f; cfg->cluster_half = (float)0.8; cfg->search_half = 1.2f * 4.0; cfg->var_thre = 0.4f; }
    cfg->obs_scale = 0.5f; cfg->obs_noise = 0.01f; cfg->max_leaves = 1 << 16; cfg->arena_chunk_bytes = 1ull << 30;
    return GPIS_OK;
}
int gpis_create(gpis_ctx** out, const gpis_config* cfg) { *out = new gpis_ctx(); (*out)->cfg = *cfg; return GPIS_OK; }
static void drop_all(gpis_ctx* c) {
    for (auto& kv : c->leaves) gpo_gp_free(kv.second.gp);
    c->leaves.clear();
    if (c->obs) gpo_obs_free(c->obs);
    c->obs = nullptr;
}
void gpis_destroy(gpis_ctx* c) { if (c) { drop_all(c); delete c; } }
int gpis_reset(gpis_ctx* c) { drop_all(c); c->next_index = 0; return GPIS_OK; }
const char* gpis_last_error(const gpis_ctx* c) { return c ? c->err.c_str() : ""; }
int gpis_device(const gpis_ctx*) { return -1; }

int gpis_leaves_update(gpis_ctx* c, int n, const int32_t* cells, const float* centres, const int32_t* offsets,
                       const float* samples, int32_t* status) {
    const int dim = c->cfg.dim, w = 2 * dim + 3;
    c->st.last_train_leaves = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<int32_t> key(cells + (size_t)i * dim, cells + (size_t)(i + 1) * dim);
        MockLeaf& L = c->leaves[key];
        if (L.N == 0 && L.gp == nullptr && L.index == 0 && c->leaves.size() > (size_t)c->next_index) L.index = c->next_index++;
        for (int a = 0; a < dim; ++a) { L.cell[a] = key[a]; L.centre[a] = centres[(size_t)i * dim + a]; }
        const int N = offsets[i + 1] - offsets[i];
        if (status) status[i] = 0;
        if (N <= 0) continue;
        gpo_gp_free(L.gp);
        L.gp = gpo_gp_train(dim, samples + (size_t)offsets[i] * w, N, c->cfg.map_scale, c->cfg.map_noise);
        L.N = N;
        if (status) status[i] = gpo_gp_chol_fail(L.gp);
        c->st.last_train_leaves++;
    }
    return GPIS_OK;
}
// f-2 on the CPU, written independently of the CUDA kernels: flat scan over all leaves instead of a hash lookup.
static void box_of(const gpis_ctx* c, const MockLeaf& L, float* lo, float* hi) {
    for (int a = 0; a < c->cfg.dim; ++a) {
        lo[a] = L.box_set ? L.lo[a] : L.centre[a] - c->cfg.cluster_half;
        hi[a] = L.box_set ? L.hi[a] : L.centre[a] + c->cfg.cluster_half;
    }
}
static bool touches(const gpis_ctx* c, const MockLeaf& L, const float* centre, float radius) {
    float lo[3], hi[3];
    box_of(c, L, lo, hi);
    for (int a = 0; a < c->cfg.dim; ++a) {
        const float qlo = centre[a] - radius, qhi = centre[a] + radius;
        if (qhi < lo[a] || qlo > hi[a]) return false;
    }
    return true;
}
int gpis_samples_set(gpis_ctx* c, int n, const int32_t* cells, const float* centres, const int32_t* offsets, const float* samples) {
    const int dim = c->cfg.dim, w = 2 * dim + 3;
    for (int i = 0; i < n; ++i) {
        std::vector<int32_t> key(cells + (size_t)i * dim, cells + (size_t)(i + 1) * dim);
        if (!c->leaves.count(key)) {
            MockLeaf& L = c->leaves[key];
            L.index = c->next_index++;
            for (int a = 0; a < dim; ++a) { L.cell[a] = key[a]; L.centre[a] = centres[(size_t)i * dim + a]; }
        }
        MockLeaf& L = c->leaves[key];
        L.own.assign(samples + (size_t)offsets[i] * w, samples + (size_t)offsets[i + 1] * w);
    }
    return GPIS_OK;
}
int gpis_leaves_train_dirty(gpis_ctx* c, int n_active, const int32_t* active, float radius, int32_t* n_trained) {
    const int dim = c->cfg.dim, w = 2 * dim + 3;
    c->st.last_train_leaves = 0;
    std::vector<MockLeaf*> dirty;
    for (int i = 0; i < n_active; ++i) {
        std::vector<int32_t> key(active + (size_t)i * dim, active + (size_t)(i + 1) * dim);
        auto it = c->leaves.find(key);
        if (it == c->leaves.end()) continue;
        for (auto& kv : c->leaves)
            if (&kv.second == &it->second || touches(c, kv.second, it->second.centre, radius))
                if (std::find(dirty.begin(), dirty.end(), &kv.second) == dirty.end()) dirty.push_back(&kv.second);
    }
    const float r2 = radius * radius;
    for (MockLeaf* D : dirty) {
        std::vector<std::pair<uint64_t, const MockLeaf*>> nb;
        for (auto& kv : c->leaves)
            if (touches(c, kv.second, D->centre, radius)) nb.push_back({dfs_key(c, kv.second.cell), &kv.second});
        std::sort(nb.begin(), nb.end(), [](const std::pair<uint64_t, const MockLeaf*>& a, const std::pair<uint64_t, const MockLeaf*>& b) { return a.first < b.first; });
        std::vector<float> ball;
        for (auto& e : nb)
            for (size_t k = 0; k + w <= e.second->own.size(); k += w) {
                const float* s = e.second->own.data() + k;
                float sq = 0.f;
                for (int a = 0; a < dim; ++a) { const float d = s[a] - D->centre[a]; sq = (a == 0) ? d * d : sq + d * d; }
                if (sq < r2) ball.insert(ball.end(), s, s + w);
            }
        const int N = (int)(ball.size() / w);
        if (N <= 0) continue;
        if (std::getenv("GPIS_MOCK_SKIP_TRAIN")) { c->st.last_train_leaves++; continue; }   // host-profile runs: the samples do not depend on the leaf GPs
        gpo_gp_free(D->gp);
        D->gp = gpo_gp_train(dim, ball.data(), N, c->cfg.map_scale, c->cfg.map_noise);
        D->N = N;
        c->st.last_train_leaves++;
    }
    if (n_trained) *n_trained = (int32_t)c->st.last_train_leaves;
    return GPIS_OK;
}
int gpis_set_train_mode(gpis_ctx* c, int mode) { return (c && mode >= 0 && mode <= 3) ? GPIS_OK : GPIS_ERR_ARG; }   // the CPU stand-in always trains at once
int gpis_train_wait(gpis_ctx* c) { return c ? GPIS_OK : GPIS_ERR_ARG; }
int gpis_train_kick(gpis_ctx* c) { return c ? GPIS_OK : GPIS_ERR_ARG; }
int gpis_leaves_mark(gpis_ctx* c, int n, const int32_t* cells, const float* centres) {
    const int dim = c->cfg.dim;
    for (int i = 0; i < n; ++i) {
        std::vector<int32_t> key(cells + (size_t)i * dim, cells + (size_t)(i + 1) * dim);
        if (c->leaves.count(key)) continue;
        MockLeaf& L = c->leaves[key];
        L.index = c->next_index++;
        for (int a = 0; a < dim; ++a) { L.cell[a] = key[a]; L.centre[a] = centres[(size_t)i * dim + a]; }
    }
    return GPIS_OK;
}
int gpis_leaves_set_boxes(gpis_ctx* c, int n, const int32_t* cells, const float* boxes) {
    const int dim = c->cfg.dim;
    for (int i = 0; i < n; ++i) {
        std::vector<int32_t> key(cells + (size_t)i * dim, cells + (size_t)(i + 1) * dim);
        auto it = c->leaves.find(key);
        if (it == c->leaves.end()) continue;
        for (int a = 0; a < dim; ++a) { it->second.lo[a] = boxes[(size_t)i * 2 * dim + a]; it->second.hi[a] = boxes[(size_t)i * 2 * dim + dim + a]; }
        it->second.box_set = true;
    }
    return GPIS_OK;
}
int gpis_leaves_erase(gpis_ctx* c, int n, const int32_t* cells) {
    const int dim = c->cfg.dim;
    for (int i = 0; i < n; ++i) {
        std::vector<int32_t> key(cells + (size_t)i * dim, cells + (size_t)(i + 1) * dim);
        auto it = c->leaves.find(key);
        if (it == c->leaves.end()) continue;
        gpo_gp_free(it->second.gp);
        c->leaves.erase(it);
    }
    return GPIS_OK;
}
int gpis_rebase(gpis_ctx* c, const int32_t* rm, int levels) {
    for (int a = 0; a < 3; ++a) c->root_min[a] = a < c->cfg.dim ? rm[a] : 0;
    c->levels = levels;
    return GPIS_OK;
}
int gpis_leaf_index(gpis_ctx* c, const int32_t* cell) {
    std::vector<int32_t> key(cell, cell + c->cfg.dim);
    auto it = c->leaves.find(key);
    return it == c->leaves.end() ? -1 : it->second.index;
}
int gpis_leaf_get(gpis_ctx* c, const int32_t* cell, int32_t* N, int32_t* ng, float* alpha, float* L, float* gradflag, int cap_n) {
    std::vector<int32_t> key(cell, cell + c->cfg.dim);
    auto it = c->leaves.find(key);
    if (it == c->leaves.end() || !it->second.gp) return 0;
    const int n = gpo_gp_n(it->second.gp);
    if (N) *N = it->second.N;
    if (ng) *ng = gpo_gp_ng(it->second.gp);
    if (n <= cap_n) gpo_gp_get(it->second.gp, alpha, L, gradflag);
    return n;
}
int gpis_query(gpis_ctx* c, const float* x, int64_t n, float* res) {
    const int dim = c->cfg.dim;
    std::vector<std::pair<uint64_t, const MockLeaf*>> order;
    for (auto& kv : c->leaves) order.push_back({dfs_key(c, kv.second.cell), &kv.second});
    std::sort(order.begin(), order.end(), [](const std::pair<uint64_t, const MockLeaf*>& a, const std::pair<uint64_t, const MockLeaf*>& b) { return a.first < b.first; });
    std::vector<float> centres, boxes;
    std::vector<gpo_gp*> gps;
    for (auto& o : order) {
        for (int a = 0; a < dim; ++a) centres.push_back(o.second->centre[a]);
        for (int a = 0; a < dim; ++a) boxes.push_back(o.second->box_set ? o.second->lo[a] : o.second->centre[a] - c->cfg.cluster_half);
        for (int a = 0; a < dim; ++a) boxes.push_back(o.second->box_set ? o.second->hi[a] : o.second->centre[a] + c->cfg.cluster_half);
        gps.push_back(o.second->gp);
    }
    gpo_map* m = gpo_map_create(dim, (int)gps.size(), centres.data(), c->cfg.cluster_half, gps.data(), c->cfg.search_half, c->cfg.var_thre, c->cfg.map_noise, boxes.data());
    gpo_map_test(m, x, (int)n, res, nullptr, nullptr);
    gpo_map_free(m);
    return GPIS_OK;
}
int gpis_query_device(gpis_ctx*, const float*, int64_t, float*) { return GPIS_ERR_STATE; }
int gpis_query_debug(gpis_ctx* c, const float* x, int64_t n, float* res, int32_t*, int32_t*) { return gpis_query(c, x, n, res); }

int gpis_obs_train_2d(gpis_ctx* c, const float* vu, const float* zinv, int ni, int nj) {
    gpo_obs* prev = (c->obs && c->obs_d == 2) ? c->obs : nullptr;
    gpo_obs* now = gpo_obs2d_retrain(vu, zinv, ni, nj, prev);
    if (c->obs) gpo_obs_free(c->obs);
    c->obs = now;
    c->obs_d = 2;
    return GPIS_OK;
}
int gpis_obs_train_1d(gpis_ctx* c, const float* th, const float* f, int n) {
    if (c->obs) gpo_obs_free(c->obs);
    c->obs = gpo_obs1d_train(th, f, n);
    c->obs_d = 1;
    return GPIS_OK;
}
int gpis_obs_test(gpis_ctx* c, const float* xt, int d, int m, float* val, float* var) {
    if (!c->obs || d != c->obs_d) return GPIS_OK;
    gpo_obs_test(c->obs, xt, m, val, var);
    return GPIS_OK;
}
// f-3 + f-1 on the CPU: the host implementation the drop-in classes used before these steps moved to the device
// (validated bit for bit against oracle/_ref), kept here as the checker of the host-side plumbing.
static float occ_test_h(float rinv, float rinv0, float a) {
    return (float)(2.0 * (1.0 / (1.0 + std::exp((double)(-a * (rinv - rinv0)))) - 0.5));
}
static float sat_h(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
int gpis_frame_eval(gpis_ctx* c, const float* dataz, int N, const float* vu_grid, const gpis_frame_params* fp, int32_t* n_valid,
                    float* range_obs_max, int32_t cap, float* xyzg, int32_t* status, float* grad_o, float* noise_o, float* gnoise_o) {
    const int n = fp->width / fp->skip, m = fp->height / fp->skip;
    const float* R = fp->pose + 3;
    const float* t = fp->pose;
    std::vector<float> zinv, vv, uu, xl;
    float rmax = 0.f;
    int K = 0;
    for (int n_ = 0; n_ < n; n_++)
        for (int m_ = 0; m_ < m; m_++) {
            const int k = (n_ * fp->skip) * fp->height + m_ * fp->skip;
            if ((k < N) && ((double)dataz[k] < fp->max_range) && ((double)dataz[k] > fp->min_range)) {
                const int j = 2 * (m * n_ + m_);
                if (rmax < dataz[k]) rmax = dataz[k];
                zinv.push_back((float)(1.0 / (double)dataz[k]));
                const float u = vu_grid[j + 1], v = vu_grid[j];
                uu.push_back(u); vv.push_back(v);
                const float xloc = u * dataz[k], yloc = v * dataz[k];
                xl.push_back(xloc); xl.push_back(yloc); xl.push_back(dataz[k]);
                if (K < cap && xyzg) {
                    xyzg[3 * K] = R[0] * xloc + R[3] * yloc + R[6] * dataz[k] + t[0];
                    xyzg[3 * K + 1] = R[1] * xloc + R[4] * yloc + R[7] * dataz[k] + t[1];
                    xyzg[3 * K + 2] = R[2] * xloc + R[5] * yloc + R[8] * dataz[k] + t[2];
                }
                ++K;
            } else zinv.push_back(-1.0f);
        }
    *n_valid = K; *range_obs_max = rmax;
    if (K <= 1) return GPIS_OK;
    if (K > cap) return GPIS_ERR_ARG;
    gpis_obs_train_2d(c, vu_grid, zinv.data(), m, n);
    static const float Xp[6] = {1.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f}, Yp[6] = {0.0f, 0.0f, 1.0f, -1.0f, 0.0f, 0.0f}, Zp[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f, -1.0f};
    std::vector<float> vu(2 * (size_t)K), r0c(K, 0.f), vc(K, 0.f);
    for (int k = 0; k < K; ++k) { vu[2 * k] = vv[k]; vu[2 * k + 1] = uu[k]; }
    gpis_obs_test(c, vu.data(), 2, K, r0c.data(), vc.data());
    std::vector<float> vup(12 * (size_t)K), r0p(6 * (size_t)K, 0.f), vp(6 * (size_t)K, 0.f);
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < 6; ++i) {
            const float X = xl[3 * k] + fp->delx * Xp[i], Y = xl[3 * k + 1] + fp->delx * Yp[i], Z = xl[3 * k + 2] + fp->delx * Zp[i];
            vup[2 * (6 * k + i)] = Y / Z; vup[2 * (6 * k + i) + 1] = X / Z;
        }
    gpis_obs_test(c, vup.data(), 2, 6 * K, r0p.data(), vp.data());
    const float w = (float)(1.0 / 6.0);
    for (int k = 0; k < K; ++k) {
        if (vc[k] > fp->obs_var_thre) { status[k] = 0; continue; }
        float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f}, occ_mean = 0.0f;
        bool failed = false;
        for (int i = 0; i < 6; i++) {
            if (vp[6 * k + i] > fp->obs_var_thre) { failed = true; break; }
            const float Z = xl[3 * k + 2] + fp->delx * Zp[i];
            occ[i] = occ_test_h((float)(1.0 / (double)Z), r0p[6 * k + i], (float)((double)Z * 30.0));
            occ_mean += w * occ[i];
        }
        if (failed) { status[k] = 1; continue; }
        float noise = 100.0f, grad_noise = 1.00f, grad[3];
        grad[0] = (occ[0] - occ[1]) / fp->delx; grad[1] = (occ[2] - occ[3]) / fp->delx; grad[2] = (occ[4] - occ[5]) / fp->delx;
        float norm_grad = grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2];
        if ((double)norm_grad > 1e-6) {
            norm_grad = std::sqrt(norm_grad);
            const float glx = grad[0] / norm_grad, gly = grad[1] / norm_grad, glz = grad[2] / norm_grad;
            grad[0] = R[0] * glx + R[3] * gly + R[6] * glz;
            grad[1] = R[1] * glx + R[4] * gly + R[7] * glz;
            grad[2] = R[2] * glx + R[5] * gly + R[8] * glz;
            const float* p = &xl[3 * k];
            const float dist = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            noise = fp->min_position_noise * (sat_h(dist, 1.0f, noise));
            grad_noise = sat_h(std::fabs(occ_mean), fp->min_grad_noise, grad_noise);
            const float view_ang = std::max(-(p[0] * glx + p[1] * gly + p[2] * glz) / dist, (float)1e-1);
            const float view_ang2 = view_ang * view_ang;
            const float view_noise = (float)((double)fp->min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
            noise += view_noise;
        }
        grad_o[3 * k] = grad[0]; grad_o[3 * k + 1] = grad[1]; grad_o[3 * k + 2] = grad[2];
        noise_o[k] = noise; gnoise_o[k] = grad_noise; status[k] = 2;
    }
    return GPIS_OK;
}
static void quat2dcm_h(const float q[4], float dcm[9]) {
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
// reEvalPoints numerics on the CPU (the host implementation that preceded gpis_reeval; bit-exact against oracle/_ref)
int gpis_reeval(gpis_ctx* c, int n, const float* smp, const gpis_frame_params* fp, float map_noise_param, int32_t* action,
                float* pos_o, float* grad_o, float* noise_o, float* gnoise_o) {
    const float* R = fp->pose + 3;
    const float* t = fp->pose;
    static const float Xp[6] = {1.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f}, Yp[6] = {0.0f, 0.0f, 1.0f, -1.0f, 0.0f, 0.0f}, Zp[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f, -1.0f};
    for (int i = 0; i < n; ++i) {
        action[i] = -1;
        const float* s = smp + 8 * (size_t)i;
        float loc[3];
        loc[0] = R[0] * (s[0] - t[0]) + R[1] * (s[1] - t[1]) + R[2] * (s[2] - t[2]);
        loc[1] = R[3] * (s[0] - t[0]) + R[4] * (s[1] - t[1]) + R[5] * (s[2] - t[2]);
        loc[2] = R[6] * (s[0] - t[0]) + R[7] * (s[1] - t[1]) + R[8] * (s[2] - t[2]);
        if ((double)loc[2] < 0.0) continue;
        float vu[2] = {loc[1] / loc[2], loc[0] / loc[2]}, rinv0 = 0.f, var = 0.f;
        gpis_obs_test(c, vu, 2, 1, &rinv0, &var);
        if (var > fp->obs_var_thre) continue;
        const float z_loc = loc[2];
        float oc = occ_test_h((float)(1.0 / (double)z_loc), rinv0, (float)((double)z_loc * 30.0));
        if ((double)oc < -0.02) continue;
        float gl[3];
        gl[0] = R[0] * s[3] + R[1] * s[4] + R[2] * s[5];
        gl[1] = R[3] * s[3] + R[4] * s[4] + R[5] * s[5];
        gl[2] = R[6] * s[3] + R[7] * s[4] + R[8] * s[5];
        float abs_oc = (float)std::fabs((double)oc), dx = fp->delx;
        float xn[3] = {loc[0], loc[1], loc[2]};
        for (int it = 0; it < 10 && (double)abs_oc > 0.02; it++) {
            if (oc < 0) { xn[0] += gl[0] * dx; xn[1] += gl[1] * dx; xn[2] += gl[2] * dx; }
            else        { xn[0] -= gl[0] * dx; xn[1] -= gl[1] * dx; xn[2] -= gl[2] * dx; }
            const float r_new = z_loc;
            const float oc_new = occ_test_h((float)(1.0 / (double)r_new), rinv0, (float)((double)r_new * 30.0));
            const float abs_oc_new = (float)std::fabs((double)oc_new);
            if ((double)abs_oc_new < 0.02 || (double)oc < -0.02) break;
            else if ((double)(oc * oc_new) < 0.0) dx = (float)(0.5 * (double)dx);
            else dx = (float)(1.1 * (double)dx);
            abs_oc = abs_oc_new;
            oc = oc_new;
        }
        float vup[12], r0p[6] = {0}, vp[6] = {0};
        for (int p = 0; p < 6; ++p) {
            const float X = xn[0] + fp->delx * Xp[p], Y = xn[1] + fp->delx * Yp[p], Z = xn[2] + fp->delx * Zp[p];
            vup[2 * p] = Y / Z; vup[2 * p + 1] = X / Z;
        }
        gpis_obs_test(c, vup, 2, 6, r0p, vp);
        action[i] = 0;
        const float w = (float)(1.0 / 6.0);
        float occ[6] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f}, occ_mean = 0.0f, r0_mean = 0.0f, r0_sqr_sum = 0.0f, r_new = loc[2], last_var = 0.f;
        for (int p = 0; p < 6; p++) {
            const float Z = xn[2] + fp->delx * Zp[p];
            r_new = Z; last_var = vp[p];
            if (vp[p] > fp->obs_var_thre) break;
            occ[p] = occ_test_h((float)(1.0 / (double)r_new), r0p[p], (float)((double)r_new * 30.0));
            occ_mean += w * occ[p];
            const float r0 = (float)(1.0 / (double)r0p[p]);
            r0_sqr_sum += r0 * r0; r0_mean += w * r0;
        }
        if (last_var > fp->obs_var_thre) continue;
        const float pos[3] = {s[0], s[1], s[2]}, grad[3] = {s[3], s[4], s[5]};
        float gnl[3] = {(occ[0] - occ[1]) / fp->delx, (occ[2] - occ[3]) / fp->delx, (occ[4] - occ[5]) / fp->delx};
        const float norm = std::sqrt(gnl[0] * gnl[0] + gnl[1] * gnl[1] + gnl[2] * gnl[2]);
        if ((double)norm < 1e-3) { action[i] = 1; continue; }
        float r_var = (float)((double)r0_sqr_sum / 5.0 - (double)(r0_mean * r0_mean) * 6.0 / 5.0);
        r_var /= fp->delx;
        float noise = 100.0f, grad_noise = 1.0f;
        if ((double)norm > 1e-6) {
            gnl[0] = gnl[0] / norm; gnl[1] = gnl[1] / norm; gnl[2] = gnl[2] / norm;
            noise = fp->min_position_noise * sat_h(r_new * r_new, 1.0f, noise);
            grad_noise = sat_h(std::fabs(occ_mean) + r_var, fp->min_grad_noise, grad_noise);
        } else noise = fp->min_position_noise * noise;
        const float dist = std::sqrt(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
        const float view_ang = std::max(-(xn[0] * gnl[0] + xn[1] * gnl[1] + xn[2] * gnl[2]) / dist, (float)1e-1);
        const float view_ang2 = view_ang * view_ang;
        const float view_noise = (float)((double)fp->min_position_noise * ((1.0 - (double)view_ang2) / (double)view_ang2));
        noise += view_noise + abs_oc;
        grad_noise = (float)((double)grad_noise + 0.1 * (double)view_noise);
        float pn[3], gn[3];
        pn[0] = R[0] * xn[0] + R[3] * xn[1] + R[6] * xn[2] + t[0];
        pn[1] = R[1] * xn[0] + R[4] * xn[1] + R[7] * xn[2] + t[1];
        pn[2] = R[2] * xn[0] + R[5] * xn[1] + R[8] * xn[2] + t[2];
        gn[0] = R[0] * gnl[0] + R[3] * gnl[1] + R[6] * gnl[2];
        gn[1] = R[1] * gnl[0] + R[4] * gnl[1] + R[7] * gnl[2];
        gn[2] = R[2] * gnl[0] + R[5] * gnl[1] + R[8] * gnl[2];
        const float noise_old = s[6], gno = s[7], pns = noise_old + noise, gns = gno + grad_noise;
        if ((double)gno > 0.5 || (double)gno > 0.6) {
            ;
        } else {
            pn[0] = (noise * pos[0] + noise_old * pn[0]) / pns;
            pn[1] = (noise * pos[1] + noise_old * pn[1]) / pns;
            pn[2] = (noise * pos[2] + noise_old * pn[2]) / pns;
            const float d2 = (pos[0] - pn[0]) * (pos[0] - pn[0]) + (pos[1] - pn[1]) * (pos[1] - pn[1]) + (pos[2] - pn[2]) * (pos[2] - pn[2]);
            const float dist2 = (float)(0.5 * (double)std::sqrt(d2));
            float axis[3];
            axis[0] = gn[1] * grad[2] - gn[2] * grad[1];
            axis[1] = -gn[0] * grad[2] + gn[2] * grad[0];
            axis[2] = gn[0] * grad[1] - gn[1] * grad[0];
            float ang = (float)std::acos((double)(gn[0] * grad[0] + gn[1] * grad[1] + gn[2] * grad[2]));
            ang = ang * noise / pns;
            float q[4] = {1.0f, 0.0f, 0.0f, 0.0f};
            if (ang > 1 - 6) {
                q[0] = (float)std::cos((double)ang / 2.0);
                const float sina = (float)std::sin((double)ang / 2.0);
                q[1] = axis[0] * sina; q[2] = axis[1] * sina; q[3] = axis[2] * sina;
            }
            float Rot[9];
            quat2dcm_h(q, Rot);
            gn[0] = Rot[0] * grad[0] + Rot[1] * grad[1] + Rot[2] * grad[2];
            gn[1] = Rot[3] * grad[0] + Rot[4] * grad[1] + Rot[5] * grad[2];
            gn[2] = Rot[6] * grad[0] + Rot[7] * grad[1] + Rot[8] * grad[2];
            grad_noise = std::min((float)1.0, std::max(grad_noise * gno / gns + dist2, map_noise_param));
            noise = std::max((noise * noise_old / pns + dist2), map_noise_param);
        }
        action[i] = 2;
        for (int a = 0; a < 3; ++a) { pos_o[3 * (size_t)i + a] = pn[a]; grad_o[3 * (size_t)i + a] = gn[a]; }
        noise_o[i] = noise; gnoise_o[i] = grad_noise;
    }
    return GPIS_OK;
}
int gpis_get_stats(gpis_ctx* c, gpis_stats* out) {
    c->st.leaves = (int64_t)c->leaves.size();
    *out = c->st;
    return GPIS_OK;
}
int gpis_set_eval_version(gpis_ctx*, int) { return GPIS_OK; }
int gpis_comm_unique_id(void*) { return GPIS_ERR_STATE; }
int gpis_comm_init(gpis_ctx*, int, int, const void*) { return GPIS_ERR_STATE; }
int gpis_replicate(gpis_ctx*, int) { return GPIS_ERR_STATE; }
int gpis_snapshot_save(gpis_ctx*, const char*) { return GPIS_ERR_STATE; }
int gpis_snapshot_load(gpis_ctx*, const char*) { return GPIS_ERR_STATE; }
int gpis_debug_program(int, int, int32_t*, int) { return -1; }

}  // extern "C"
