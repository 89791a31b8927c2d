// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// extern "C" harness around the UNMODIFIED reference classes (compiled from
// /root/reference/cpp/src/*.cpp against oracle/eigen_shim) so that Python tests and the
// bench's cpu_baseline leg can drive GPisMap / GPisMap3 exactly like the MATLAB demos do
// (mex/mexGPisMap3.cpp:43-169, mex/mexGPisMap.cpp:30-134) and can read back the internal
// state that defines the kernel-boundary golden vectors (training sets in QueryRange
// order, alpha, L, candidate lists in DFS order).
//
// Built only by oracle/Makefile into oracle/_ref/libgpisref.so (git-ignored).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <thread>
#include <unordered_set>
#include <vector>
#include <Eigen/Dense>

// The harness needs protected/private members (tree root, active set, OnGPIS factors).
// Access specifiers do not change object layout with this compiler, and the reference
// translation units themselves are compiled without this define.
#define private public
#define class struct
#define protected public
#include "GPisMap.h"
#include "GPisMap3.h"
#include "covFnc.h"
#undef private
#undef class
#undef protected

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Samples3 {
    static std::shared_ptr<Node3> make(const float* s) {
        // layout: pos[3], grad[3], val, pose_sig, grad_sig
        return std::make_shared<Node3>(Point3<float>(s[0], s[1], s[2]), s[6], s[7],
                                       Point3<float>(s[3], s[4], s[5]), s[8], NODE_TYPE::HIT);
    }
    static void dump(const std::shared_ptr<Node3>& n, float* s) {
        s[0] = n->getPosX(); s[1] = n->getPosY(); s[2] = n->getPosZ();
        s[3] = n->getGradX(); s[4] = n->getGradY(); s[5] = n->getGradZ();
        s[6] = n->getVal(); s[7] = n->getPosNoise(); s[8] = n->getGradNoise();
    }
};
struct Samples2 {
    static std::shared_ptr<Node> make(const float* s) {
        // layout: pos[2], grad[2], val, pose_sig, grad_sig
        return std::make_shared<Node>(Point<float>(s[0], s[1]), s[4], s[5], Point<float>(s[2], s[3]), s[6],
                                      NODE_TYPE::HIT);
    }
    static void dump(const std::shared_ptr<Node>& n, float* s) {
        s[0] = n->getPosX(); s[1] = n->getPosY();
        s[2] = n->getGradX(); s[3] = n->getGradY();
        s[4] = n->getVal(); s[5] = n->getPosNoise(); s[6] = n->getGradNoise();
    }
};

const float HUGE_HALF = 1.0e6f;

}  // namespace

extern "C" {

// The sort the reference runs on its candidate lists (GPisMap3.cpp:826-829: std::sort of an index array with the
// comparator sqdst[i1] < sqdst[i2]), on its own: the pin for the oracle's move-by-move replay.
void ref_std_sort_indices(const float* keys, int n, int* idx) {
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx, idx + n, [&](int a, int b) { return keys[a] < keys[b]; });
}
// McIlroy's adversary ("A killer adversary for quicksort", 1999) run against this libstdc++'s std::sort: produces
// keys for which the introsort's partitions are as bad as possible, so that its depth limit trips and the heapsort
// fallback runs. keys_out receives a permutation-like array of n distinct values.
void ref_std_sort_killer(int n, int ties, float* keys_out) {   // ties > 1: the comparator sees val / ties, so exact ties occur
    std::vector<int> val(n), idx(n);
    if (ties < 1) ties = 1;
    const int gas = n;
    int nsolid = 0, candidate = 0;
    for (int i = 0; i < n; ++i) { val[i] = gas; idx[i] = i; }
    std::sort(idx.begin(), idx.end(), [&](int x, int y) {
        if (val[x] == gas && val[y] == gas) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
        if (val[x] == gas) candidate = x;
        else if (val[y] == gas) candidate = y;
        return val[x] / ties < val[y] / ties;
    });
    for (int i = 0; i < n; ++i) keys_out[i] = (float)((val[i] == gas ? nsolid++ : val[i]) / ties);
}
int ref_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

// ------------------------------------------------------------------ covariance functions
// covFnc.cpp:111-139 (dispatch). x is dim x N column-major, K is (N+dim*ng)^2 column-major.
int ref_matern_train(int dim, const float* x, const float* gradflag, int N, float scale, const float* sigx,
                     const float* siggrad, float* Kout, int cap) {
    EMatrixX X(dim, N);
    std::memcpy(X.data(), x, sizeof(float) * dim * N);
    std::vector<float> gf(gradflag, gradflag + N);
    EVectorX sx(N), sg(N);
    std::memcpy(sx.data(), sigx, sizeof(float) * N);
    std::memcpy(sg.data(), siggrad, sizeof(float) * N);
    EMatrixX K = matern32_sparse_deriv1(X, gf, scale, sx, sg);
    int n = K.rows();
    if (Kout && cap >= n * n) std::memcpy(Kout, K.data(), sizeof(float) * n * n);
    return n;
}

// covFnc.cpp:126-139. xt is dim x m; K is (N+dim*ng) x m(1+dim) column-major.
int ref_matern_test(int dim, const float* x, const float* gradflag, int N, const float* xt, int m, float scale,
                    float* Kout, int cap) {
    EMatrixX X(dim, N), XT(dim, m);
    std::memcpy(X.data(), x, sizeof(float) * dim * N);
    std::memcpy(XT.data(), xt, sizeof(float) * dim * m);
    std::vector<float> gf(gradflag, gradflag + N);
    EMatrixX K = matern32_sparse_deriv1(X, gf, XT, scale);
    int sz = K.rows() * K.cols();
    if (Kout && cap >= sz) std::memcpy(Kout, K.data(), sizeof(float) * sz);
    return K.rows();
}

// covFnc.cpp:47-68 / 93-109
void ref_ou_train(int dim, const float* x, int N, float scale, float sig, float* Kout) {
    EMatrixX X(dim, N);
    std::memcpy(X.data(), x, sizeof(float) * dim * N);
    EMatrixX K = ornstein_uhlenbeck(X, scale, sig);
    std::memcpy(Kout, K.data(), sizeof(float) * N * N);
}

// ------------------------------------------------------------------ stand-alone leaf GP (OnGPIS)
struct RefGP {
    OnGPIS gp;
    int dim;
    RefGP(float s, float n, int d) : gp(s, n), dim(d) {}
};

// OnGPIS.cpp:91-149 (3D), :34-89 (2D). samples: 9 floats (3D) / 7 floats (2D) per sample.
void* ref_gp_train(int dim, const float* samples, int N, float scale, float noise) {
    RefGP* h = new RefGP(scale, noise, dim);
    if (dim == 3) {
        vecNode3 v;
        for (int i = 0; i < N; ++i) v.push_back(Samples3::make(samples + 9 * i));
        h->gp.train(v);
    } else {
        vecNode v;
        for (int i = 0; i < N; ++i) v.push_back(Samples2::make(samples + 7 * i));
        h->gp.train(v);
    }
    return h;
}
void ref_gp_free(void* h) { delete (RefGP*)h; }
int ref_gp_n(void* h) { return ((RefGP*)h)->gp.alpha.size(); }
int ref_gp_nsamples(void* h) { return ((RefGP*)h)->gp.nSamples; }
// alpha: n floats; L: n*n column-major dense lower; gradflag: N floats (0/1)
void ref_gp_get(void* h_, float* alpha, float* L, float* gradflag) {
    RefGP* h = (RefGP*)h_;
    int n = h->gp.alpha.size();
    if (alpha) std::memcpy(alpha, h->gp.alpha.data(), sizeof(float) * n);
    if (L) std::memcpy(L, h->gp.L.data(), sizeof(float) * (size_t)n * n);
    if (gradflag) std::memcpy(gradflag, h->gp.gradflag.data(), sizeof(float) * h->gp.gradflag.size());
}
// OnGPIS.cpp:177-216 (3D) / :218-239 (2D). res row: 2(1+dim) floats, read-modify-write.
void ref_gp_test(void* h_, const float* x, int m, float* res) {
    RefGP* h = (RefGP*)h_;
    int dim = h->dim, w = 2 * (1 + dim);
    for (int i = 0; i < m; ++i) {
        EVectorX xt(dim);
        for (int c = 0; c < dim; ++c) xt(c) = x[dim * i + c];
        float* r = res + (size_t)w * i;
        if (dim == 3) h->gp.testSinglePoint(xt, r[0], &r[1], &r[4]);
        else h->gp.test2Dpoint(xt, r[0], r[1], r[2], r[3], r[4], r[5]);
    }
}

// ------------------------------------------------------------------ observation GPs
void* ref_obs2d_create() { return new ObsGP2D(); }
void ref_obs2d_free(void* h) { delete (ObsGP2D*)h; }
// GPisMap3.cpp:239-256 calls the BASE reset() then train(): reproduce that exact sequence.
int ref_obs2d_train(void* h_, float* vu, float* zinv, int ni, int nj) {
    ObsGP* h = (ObsGP2D*)h_;
    int dim[2] = {ni, nj};
    h->reset();
    h->train(vu, zinv, dim);
    return h->isTrained() ? 1 : 0;
}
// ObsGP.cpp:410-463; xt is 2 x m [v;u]; val/var are read-modify-write. Each point is tested
// by its own call, as every reference call site does (GPisMap3.cpp:349,393,430,599,641).
void ref_obs2d_test(void* h_, const float* xt, int m, float* val, float* var) {
    ObsGP2D* h = (ObsGP2D*)h_;
    EMatrixX vu(2, 1);
    EVectorX f(1), v(1);
    for (int i = 0; i < m; ++i) {
        vu(0) = xt[2 * i]; vu(1) = xt[2 * i + 1];
        f(0) = val[i]; v(0) = var[i];
        h->test(vu, f, v);
        val[i] = f(0); var[i] = v(0);
    }
}
int ref_obs2d_tiles(void* h_, int* npts, int cap) {
    std::vector<int> n;
    ((ObsGP2D*)h_)->getNumValidPoints(n);
    for (int i = 0; i < (int)n.size() && i < cap; ++i) npts[i] = n[i];
    return (int)n.size();
}
int ref_obs2d_partition(void* h_, float* val_i, float* val_j, int cap) {
    ObsGP2D* h = (ObsGP2D*)h_;
    for (int i = 0; i < (int)h->Val_i.size() && i < cap; ++i) val_i[i] = h->Val_i[i];
    for (int i = 0; i < (int)h->Val_j.size() && i < cap; ++i) val_j[i] = h->Val_j[i];
    return (int)h->Val_i.size() * 65536 + (int)h->Val_j.size();
}

void* ref_obs1d_create() { return new ObsGP1D(); }
void ref_obs1d_free(void* h) { delete (ObsGP1D*)h; }
int ref_obs1d_train(void* h_, float* theta, float* f, int n) {
    ObsGP* h = (ObsGP1D*)h_;
    int N[2] = {n, 0};
    h->reset();
    h->train(theta, f, N);
    return h->isTrained() ? 1 : 0;
}
void ref_obs1d_test(void* h_, const float* xt, int m, float* val, float* var) {
    ObsGP1D* h = (ObsGP1D*)h_;
    EMatrixX a(1, 1);
    EVectorX f(1), v(1);
    for (int i = 0; i < m; ++i) {
        a(0) = xt[i];
        f(0) = val[i]; v(0) = var[i];
        h->test(a, f, v);
        val[i] = f(0); var[i] = v(0);
    }
}
int ref_obs1d_ranges(void* h_, float* range, int cap) {
    ObsGP1D* h = (ObsGP1D*)h_;
    for (int i = 0; i < (int)h->range.size() && i < cap; ++i) range[i] = h->range[i];
    return (int)h->range.size();
}

// ------------------------------------------------------------------ GPisMap3 (3D)
void* ref3_create() { return new GPisMap3(); }
void* ref3_create_cam(float fx, float fy, float cx, float cy, int w, int h) {
    GPisMap3Param p;
    camParam c(fx, fy, cx, cy, w, h);
    return new GPisMap3(p, c);
}
void ref3_destroy(void* m) { delete (GPisMap3*)m; }
void ref3_reset(void* m) { ((GPisMap3*)m)->reset(); }
void ref3_set_cam(void* m, float fx, float fy, float cx, float cy, int w, int h) {
    camParam c(fx, fy, cx, cy, w, h);
    ((GPisMap3*)m)->resetCam(c);
}
void ref3_update(void* m, float* depth, int N, const float* pose12) {
    std::vector<float> pose(pose12, pose12 + 12);
    ((GPisMap3*)m)->update(depth, N, pose);
}
// Same steps as GPisMap3::update (GPisMap3.cpp:218-237) with wall-clock per phase:
// phases = {preproc, regressObs, updateMapPoints, addNewMeas, updateGPs} seconds; counts =
// {valid pixels, active clusters before updateGPs}.
void ref3_update_timed(void* m_, float* depth, int N, const float* pose12, double* phases, int* counts) {
    GPisMap3* m = (GPisMap3*)m_;
    std::vector<float> pose(pose12, pose12 + 12);
    for (int i = 0; i < 5; ++i) phases[i] = 0;
    counts[0] = counts[1] = 0;
    double t0 = now_s();
    bool ok = m->preprocData(depth, N, pose);
    double t1 = now_s();
    phases[0] = t1 - t0;
    counts[0] = m->obs_numdata;
    if (!ok) return;
    ok = m->regressObs();
    double t2 = now_s();
    phases[1] = t2 - t1;
    if (!ok) return;
    m->updateMapPoints();
    double t3 = now_s();
    phases[2] = t3 - t2;
    m->addNewMeas();
    double t4 = now_s();
    phases[3] = t4 - t3;
    counts[1] = (int)m->activeSet.size();
    m->updateGPs();
    phases[4] = now_s() - t4;
}
// GPisMap3::update without its last step: the samples of the map depend only on the observation GP and the tree
// logic (updateMapPoints / addNewMeas never read a leaf GP), so a long sequence can be mapped without paying for
// updateGPs after every frame; train once at the end with ref3_activate + ref3_update_gps.
void ref3_update_nogp(void* m_, float* depth, int N, const float* pose12) {
    GPisMap3* m = (GPisMap3*)m_;
    std::vector<float> pose(pose12, pose12 + 12);
    if (!m->preprocData(depth, N, pose)) return;
    if (!m->regressObs()) return;
    m->updateMapPoints();
    m->addNewMeas();
    m->activeSet.clear();
}
// Put every non-empty cluster whose centre lies inside [lo,hi] into the active set (for ref3_update_gps).
int ref3_activate(void* m_, const float* lo, const float* hi) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<OcTree*> all;
    m->t->QueryNonEmptyLevelC(AABB3(0.f, 0.f, 0.f, HUGE_HALF), all);
    int n = 0;
    for (auto q : all) {
        Point3<float> c = q->getCenter();
        if (c.x >= lo[0] && c.x <= hi[0] && c.y >= lo[1] && c.y <= hi[1] && c.z >= lo[2] && c.z <= hi[2]) {
            m->activeSet.insert(q);
            ++n;
        }
    }
    return n;
}
// ---- exact tree snapshot. The final state of a mapped octree cannot be rebuilt by re-inserting its samples (the
// minimum-spacing rule, octree.cpp:331-333, rejects pairs that ended up closer than that across cell boundaries), so
// long runs are frozen as a pre-order dump and rebuilt node by node with the reference's own Subdivide().
//   flags: one byte per visited node, bit 0 = leaf, bit 1 = holds a sample (its 9 floats follow in `samples`).
// Subtrees whose box does not touch [lo,hi] are recorded as empty leaves. Returns the node count; sample count in *ns.
namespace {
void dump_rec(OcTree* t, const float* lo, const float* hi, std::vector<unsigned char>& flags, std::vector<float>& smp) {
    Point3<float> c = t->getCenter();
    const float h = t->getHalfLength();
    const bool touch = !(c.x + h < lo[0] || c.x - h > hi[0] || c.y + h < lo[1] || c.y - h > hi[1] || c.z + h < lo[2] || c.z - h > hi[2]);
    if (!touch) { flags.push_back(1); return; }
    unsigned char f = (t->leaf ? 1 : 0) | (t->node != nullptr ? 2 : 0);
    flags.push_back(f);
    if (t->node != nullptr) { float s[9]; Samples3::dump(t->node, s); smp.insert(smp.end(), s, s + 9); }
    if (!t->leaf) {
        OcTree* ch[8] = {t->northWestFront, t->northEastFront, t->southWestFront, t->southEastFront,
                         t->northWestBack, t->northEastBack, t->southWestBack, t->southEastBack};
        for (int i = 0; i < 8; ++i) dump_rec(ch[i], lo, hi, flags, smp);
    }
}
int load_rec(OcTree* t, const unsigned char* flags, size_t& fi, const float* smp, size_t& si) {
    const unsigned char f = flags[fi++];
    int cnt = 0;
    if (f & 2) { t->node = Samples3::make(smp + 9 * si); ++si; cnt = 1; }
    if (!(f & 1)) {
        t->Subdivide();
        t->leaf = false;
        OcTree* ch[8] = {t->northWestFront, t->northEastFront, t->southWestFront, t->southEastFront,
                         t->northWestBack, t->northEastBack, t->southWestBack, t->southEastBack};
        for (int i = 0; i < 8; ++i) cnt += load_rec(ch[i], flags, fi, smp, si);
    }
    t->numNodes = cnt;
    return cnt;
}
}  // namespace
int ref3_tree_dump(void* m_, const float* lo, const float* hi, unsigned char* flags, int cap_flags, float* samples,
                   int cap_samples, float* root_c_half, int* ns) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<unsigned char> f;
    std::vector<float> s;
    dump_rec(m->t, lo, hi, f, s);
    Point3<float> c = m->t->getCenter();
    root_c_half[0] = c.x; root_c_half[1] = c.y; root_c_half[2] = c.z; root_c_half[3] = m->t->getHalfLength();
    *ns = (int)(s.size() / 9);
    if ((int)f.size() <= cap_flags && *ns <= cap_samples) {
        std::memcpy(flags, f.data(), f.size());
        std::memcpy(samples, s.data(), s.size() * sizeof(float));
    }
    return (int)f.size();
}
int ref3_tree_load(void* m_, const unsigned char* flags, int nflags, const float* samples, int ns, const float* root_c_half) {
    GPisMap3* m = (GPisMap3*)m_;
    m->reset();
    m->t = new OcTree(AABB3(Point3<float>(root_c_half[0], root_c_half[1], root_c_half[2]), root_c_half[3]), (OcTree*)0);
    size_t fi = 0, si = 0;
    const int cnt = load_rec(m->t, flags, fi, samples, si);
    return (fi == (size_t)nflags && si == (size_t)ns) ? cnt : -1;
}
int ref3_test(void* m_, float* x, int n, float* res) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;  // the reference would dereference a null tree (GPisMap3.cpp:814)
    return m->test(x, 3, n, res) ? 1 : 0;
}
int ref3_get_all_points(void* m, float* out, int cap) {
    std::vector<float> pos;
    ((GPisMap3*)m)->getAllPoints(pos);
    int n = (int)pos.size() / 3;
    if (out && cap >= n) std::memcpy(out, pos.data(), sizeof(float) * pos.size());
    return n;
}
// All non-empty cluster-level nodes in DFS order (octree.cpp:829-859).
int ref3_clusters(void* m_, float* centres, int* nsamples, int* trained, int cap) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<OcTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB3(0.f, 0.f, 0.f, HUGE_HALF), oc);
    for (int i = 0; i < (int)oc.size() && i < cap; ++i) {
        Point3<float> c = oc[i]->getCenter();
        if (centres) { centres[3 * i] = c.x; centres[3 * i + 1] = c.y; centres[3 * i + 2] = c.z; }
        if (nsamples) nsamples[i] = oc[i]->getNodeCount();
        if (trained) trained[i] = (oc[i]->getGP() != nullptr) ? oc[i]->getGP()->nSamples : -1;
    }
    return (int)oc.size();
}
// Effective box of every non-empty cluster (DFS order): intersection of the float AABBs of the
// cluster and all its ancestors — what the level-by-level pruning of QueryNonEmptyLevelC
// (octree.cpp:864-867) amounts to. boxes: n x 6 floats (lo xyz, hi xyz).
int ref3_cluster_boxes(void* m_, float* boxes, int cap) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<OcTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB3(0.f, 0.f, 0.f, HUGE_HALF), oc);
    for (int i = 0; i < (int)oc.size() && i < cap; ++i) {
        float lo[3] = {-3e38f, -3e38f, -3e38f}, hi[3] = {3e38f, 3e38f, 3e38f};
        for (OcTree* q = oc[i]; q != 0; q = q->getParent()) {
            lo[0] = std::max(lo[0], q->boundary.getXMinbound()); hi[0] = std::min(hi[0], q->getXMaxbound());
            lo[1] = std::max(lo[1], q->getYMinbound()); hi[1] = std::min(hi[1], q->getYMaxbound());
            lo[2] = std::max(lo[2], q->getZMinbound()); hi[2] = std::min(hi[2], q->getZMaxbound());
        }
        for (int c = 0; c < 3; ++c) { boxes[6 * i + c] = lo[c]; boxes[6 * i + 3 + c] = hi[c]; }
    }
    return (int)oc.size();
}
int ref2_cluster_boxes(void* m_, float* boxes, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<QuadTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB(0.f, 0.f, HUGE_HALF), oc);
    for (int i = 0; i < (int)oc.size() && i < cap; ++i) {
        float lo[2] = {-3e38f, -3e38f}, hi[2] = {3e38f, 3e38f};
        for (QuadTree* q = oc[i]; q != 0; q = q->getParent()) {
            lo[0] = std::max(lo[0], q->boundary.getXMinbound()); hi[0] = std::min(hi[0], q->boundary.getXMaxbound());
            lo[1] = std::max(lo[1], q->boundary.getYMinbound()); hi[1] = std::min(hi[1], q->boundary.getYMaxbound());
        }
        for (int c = 0; c < 2; ++c) { boxes[4 * i + c] = lo[c]; boxes[4 * i + 2 + c] = hi[c]; }
    }
    return (int)oc.size();
}
void ref3_root(void* m_, float* c_half) {
    GPisMap3* m = (GPisMap3*)m_;
    c_half[0] = c_half[1] = c_half[2] = c_half[3] = 0.f;
    if (m->t == 0) return;
    Point3<float> c = m->t->getCenter();
    c_half[0] = c.x; c_half[1] = c.y; c_half[2] = c.z; c_half[3] = m->t->getHalfLength();
}
// Training set of the cluster centred at `centre`, exactly as updateGPs_kernel gathers it
// (GPisMap3.cpp:705-709): ball of radius Rtimes*l in QueryRange DFS order. 9 floats/sample.
int ref3_train_set(void* m_, const float* centre, float half, float rtimes, float* samples, int cap) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    AABB3 bb(centre[0], centre[1], centre[2], half * rtimes);
    std::vector<std::shared_ptr<Node3> > res;
    m->t->QueryRange(bb, res);
    for (int i = 0; i < (int)res.size() && i < cap; ++i) Samples3::dump(res[i], samples + 9 * i);
    return (int)res.size();
}
static OcTree* find_cluster3(GPisMap3* m, const float* centre) {
    std::vector<OcTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB3(centre[0], centre[1], centre[2], 1e-4f), oc);
    for (auto q : oc) {
        Point3<float> c = q->getCenter();
        if (std::fabs(c.x - centre[0]) < 1e-4f && std::fabs(c.y - centre[1]) < 1e-4f &&
            std::fabs(c.z - centre[2]) < 1e-4f)
            return q;
    }
    return nullptr;
}
// Trained factors the map holds for a cluster. Returns n (0 if untrained); fills N.
int ref3_cluster_gp(void* m_, const float* centre, float* alpha, float* L, int* N, int cap_n) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    OcTree* q = find_cluster3(m, centre);
    if (!q || q->getGP() == nullptr) return 0;
    std::shared_ptr<OnGPIS> gp = q->getGP();
    int n = gp->alpha.size();
    if (N) *N = gp->nSamples;
    if (n <= cap_n) {
        if (alpha) std::memcpy(alpha, gp->alpha.data(), sizeof(float) * n);
        if (L) std::memcpy(L, gp->L.data(), sizeof(float) * (size_t)n * n);
    }
    return n;
}
// Candidate clusters of a query exactly as test_kernel asks for them (GPisMap3.cpp:811-814):
// DFS order, with squared centre distances.
int ref3_candidates(void* m_, const float* x, float half, float* centres, float* sqdst, int cap) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<OcTree*> quads;
    std::vector<float> d;
    m->t->QueryNonEmptyLevelC(AABB3(x[0], x[1], x[2], half), quads, d);
    for (int i = 0; i < (int)quads.size() && i < cap; ++i) {
        Point3<float> c = quads[i]->getCenter();
        centres[3 * i] = c.x; centres[3 * i + 1] = c.y; centres[3 * i + 2] = c.z;
        sqdst[i] = d[i];
    }
    return (int)quads.size();
}
// Bulk-load samples straight into the tree (bypassing the sensor pipeline) using the
// reference's own Insert (octree.cpp:295-414) and mark the touched clusters active, as
// evalPoints does (GPisMap3.cpp:613-614, 687-691). Returns number inserted.
int ref3_insert_samples(void* m_, const float* samples, int N) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) m->t = new OcTree(Point3<float>(0.0, 0.0, 0.0));
    int cnt = 0;
    for (int i = 0; i < N; ++i) {
        std::shared_ptr<Node3> p = Samples3::make(samples + 9 * i);
        std::unordered_set<OcTree*> ins;
        bool ok = false;
        if (!m->t->IsNotNew(p)) {
            ok = m->t->Insert(p, ins);
            if (ok && !m->t->IsRoot()) m->t = m->t->getRoot();
        }
        if (!ok || ins.empty()) continue;
        for (auto q : ins) m->activeSet.insert(q);
        ++cnt;
    }
    return cnt;
}
// Restrict the active set to clusters whose centre lies inside the box [lo,hi] (bounded
// CPU samples of a large map), then train: GPisMap3::updateGPs (GPisMap3.cpp:720-792).
int ref3_update_gps(void* m_, const float* lo, const float* hi) {
    GPisMap3* m = (GPisMap3*)m_;
    if (lo && hi) {
        std::unordered_set<OcTree*> keep;
        for (auto q : m->activeSet) {
            Point3<float> c = q->getCenter();
            if (c.x >= lo[0] && c.x <= hi[0] && c.y >= lo[1] && c.y <= hi[1] && c.z >= lo[2] && c.z <= hi[2])
                keep.insert(q);
        }
        m->activeSet.swap(keep);
    }
    int n = (int)m->activeSet.size();
    m->updateGPs();
    return n;
}

// ------------------------------------------------------------------ GPisMap (2D)
void* ref2_create() { return new GPisMap(); }
void ref2_destroy(void* m) { delete (GPisMap*)m; }
void ref2_reset(void* m) { ((GPisMap*)m)->reset(); }
void ref2_update(void* m, float* theta, float* range, int N, const float* pose6) {
    std::vector<float> pose(pose6, pose6 + 6);
    ((GPisMap*)m)->update(theta, range, N, pose);
}
void ref2_update_timed(void* m_, float* theta, float* range, int N, const float* pose6, double* phases,
                       int* counts) {
    GPisMap* m = (GPisMap*)m_;
    std::vector<float> pose(pose6, pose6 + 6);
    for (int i = 0; i < 5; ++i) phases[i] = 0;
    counts[0] = counts[1] = 0;
    double t0 = now_s();
    bool ok = m->preproData(theta, range, N, pose);
    double t1 = now_s();
    phases[0] = t1 - t0;
    counts[0] = m->obs_numdata;
    if (!ok) return;
    ok = m->regressObs();
    double t2 = now_s();
    phases[1] = t2 - t1;
    if (!ok) return;
    m->updateMapPoints();
    double t3 = now_s();
    phases[2] = t3 - t2;
    m->addNewMeas();
    double t4 = now_s();
    phases[3] = t4 - t3;
    counts[1] = (int)m->activeSet.size();
    if (!m->activeSet.empty()) m->updateGPs();  // the 2D copy divides by zero on an empty set (GPisMap.cpp:625-632)
    phases[4] = now_s() - t4;
}
int ref2_test(void* m_, float* x, int n, float* res) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    return m->test(x, 2, n, res) ? 1 : 0;
}
int ref2_get_all_points(void* m_, float* out, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<std::shared_ptr<Node> > nodes;
    m->t->getAllChildrenNonEmptyNodes(nodes);
    int n = (int)nodes.size();
    if (out && cap >= n)
        for (int i = 0; i < n; ++i) { out[2 * i] = nodes[i]->getPosX(); out[2 * i + 1] = nodes[i]->getPosY(); }
    return n;
}
int ref2_all_samples(void* m_, float* out, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<std::shared_ptr<Node> > nodes;
    m->t->getAllChildrenNonEmptyNodes(nodes);
    int n = (int)nodes.size();
    if (out && cap >= n)
        for (int i = 0; i < n; ++i) Samples2::dump(nodes[i], out + 7 * i);
    return n;
}
int ref3_all_samples(void* m_, float* out, int cap) {
    GPisMap3* m = (GPisMap3*)m_;
    if (m->t == 0) return 0;
    std::vector<std::shared_ptr<Node3> > nodes;
    m->t->getAllChildrenNonEmptyNodes(nodes);
    int n = (int)nodes.size();
    if (out && cap >= n)
        for (int i = 0; i < n; ++i) Samples3::dump(nodes[i], out + 9 * i);
    return n;
}
int ref2_clusters(void* m_, float* centres, int* nsamples, int* trained, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<QuadTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB(0.f, 0.f, HUGE_HALF), oc);
    for (int i = 0; i < (int)oc.size() && i < cap; ++i) {
        Point<float> c = oc[i]->getCenter();
        if (centres) { centres[2 * i] = c.x; centres[2 * i + 1] = c.y; }
        if (nsamples) nsamples[i] = oc[i]->getNodeCount();
        if (trained) trained[i] = (oc[i]->getGP() != nullptr) ? oc[i]->getGP()->nSamples : -1;
    }
    return (int)oc.size();
}
void ref2_root(void* m_, float* c_half) {
    GPisMap* m = (GPisMap*)m_;
    c_half[0] = c_half[1] = c_half[2] = 0.f;
    if (m->t == 0) return;
    Point<float> c = m->t->getCenter();
    c_half[0] = c.x; c_half[1] = c.y; c_half[2] = m->t->getHalfLength();
}
// GPisMap.cpp:580-585: ball radius 4*l.
int ref2_train_set(void* m_, const float* centre, float half, float rtimes, float* samples, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    AABB bb(centre[0], centre[1], half * rtimes);
    std::vector<std::shared_ptr<Node> > res;
    m->t->QueryRange(bb, res);
    for (int i = 0; i < (int)res.size() && i < cap; ++i) Samples2::dump(res[i], samples + 7 * i);
    return (int)res.size();
}
int ref2_cluster_gp(void* m_, const float* centre, float* alpha, float* L, int* N, int cap_n) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<QuadTree*> oc;
    m->t->QueryNonEmptyLevelC(AABB(centre[0], centre[1], 1e-3f), oc);
    for (auto q : oc) {
        Point<float> c = q->getCenter();
        if (std::fabs(c.x - centre[0]) < 1e-3f && std::fabs(c.y - centre[1]) < 1e-3f) {
            std::shared_ptr<OnGPIS> gp = q->getGP();
            if (gp == nullptr) return 0;
            int n = gp->alpha.size();
            if (N) *N = gp->nSamples;
            if (n <= cap_n) {
                if (alpha) std::memcpy(alpha, gp->alpha.data(), sizeof(float) * n);
                if (L) std::memcpy(L, gp->L.data(), sizeof(float) * (size_t)n * n);
            }
            return n;
        }
    }
    return 0;
}
int ref2_candidates(void* m_, const float* x, float half, float* centres, float* sqdst, int cap) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) return 0;
    std::vector<QuadTree*> quads;
    std::vector<float> d;
    m->t->QueryNonEmptyLevelC(AABB(x[0], x[1], half), quads, d);
    for (int i = 0; i < (int)quads.size() && i < cap; ++i) {
        Point<float> c = quads[i]->getCenter();
        centres[2 * i] = c.x; centres[2 * i + 1] = c.y;
        sqdst[i] = d[i];
    }
    return (int)quads.size();
}
int ref2_insert_samples(void* m_, const float* samples, int N) {
    GPisMap* m = (GPisMap*)m_;
    if (m->t == 0) m->t = new QuadTree(Point<float>(0.0, 0.0));
    int cnt = 0;
    for (int i = 0; i < N; ++i) {
        std::shared_ptr<Node> p = Samples2::make(samples + 7 * i);
        std::unordered_set<QuadTree*> ins;
        bool ok = false;
        if (!m->t->IsNotNew(p)) {
            ok = m->t->Insert(p, ins);
            if (ok && !m->t->IsRoot()) m->t = m->t->getRoot();
        }
        if (!ok || ins.empty()) continue;
        for (auto q : ins) m->activeSet.insert(q);
        ++cnt;
    }
    return cnt;
}
int ref2_update_gps(void* m_) {
    GPisMap* m = (GPisMap*)m_;
    int n = (int)m->activeSet.size();
    if (n > 0) m->updateGPs();
    return n;
}

}  // extern "C"
