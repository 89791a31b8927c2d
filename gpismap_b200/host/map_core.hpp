// Dimension-independent part of the drop-in map classes: owns the host tree, the active set and
// the gpis_ctx; turns dirty leaves into the CSR the C ABI trains (replacing updateGPs,
// cpp/src/GPisMap3.cpp:720-792 / cpp/src/GPisMap.cpp:596-663) and keeps the device leaf table in
// step with the tree (which leaves are non-empty, root box for DFS tie-breaks).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <unordered_set>
#include <thread>
#include <vector>

#include "gpis_b200.h"
#include "prtree.hpp"

namespace gpismap_host {

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Coarse host-side profile (development aid, scripts/update_profile.py): seconds and call counts per slot.
//  0 gpis_obs_test  1 updateMapPoints: cull  2 updateMapPoints: collect+stage  3 updateMapPoints: serial apply
//  4 evalPoints: serial insert  5 train_active: dirty set  6 train_active: gather  7 gpis_leaves_update  8 sync_table
inline double g_prof_s[16] = {0};
inline long long g_prof_n[16] = {0};
struct ProfScope {
    int slot; double t0;
    explicit ProfScope(int s) : slot(s), t0(now_s()) {}
    ~ProfScope() { g_prof_s[slot] += now_s() - t0; g_prof_n[slot] += 1; }
};

struct LeafHandle {
    int cell; uint32_t gen;
    bool operator<(const LeafHandle& o) const { return cell < o.cell || (cell == o.cell && gen < o.gen); }
};

// activeSet (GPisMap3.h:91, std::unordered_set<OcTree*> in the reference): a list plus a per-cell mark instead of a
// node-based set — every accepted sample activates its leaf, so insert() runs ~10^5 times per frame. A handle is
// live while mark[cell] == gen + 1; handles of freed cells are dropped by erase() or fail cell_alive() later.
class ActiveSet {
public:
    void insert(const LeafHandle& h) {
        if ((size_t)h.cell >= mark_.size()) mark_.resize((size_t)h.cell + 1 + mark_.size() / 2, 0);
        if (mark_[h.cell] == h.gen + 1) return;
        if (mark_[h.cell] == 0) ++live_;
        mark_[h.cell] = h.gen + 1;
        items_.push_back(h);
    }
    void erase(const LeafHandle& h) {
        if ((size_t)h.cell < mark_.size() && mark_[h.cell] == h.gen + 1) { mark_[h.cell] = 0; --live_; }
    }
    void clear() {
        for (const LeafHandle& h : items_) mark_[h.cell] = 0;
        items_.clear();
        live_ = 0;
    }
    size_t size() const { return live_; }
    // live handles in (cell, generation) order — the iteration order of the std::set this replaces
    std::vector<LeafHandle> sorted() const {
        std::vector<LeafHandle> v;
        v.reserve(live_);
        for (const LeafHandle& h : items_) if (mark_[h.cell] == h.gen + 1) v.push_back(h);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end(), [](const LeafHandle& a, const LeafHandle& b) { return a.cell == b.cell && a.gen == b.gen; }), v.end());
        return v;
    }
private:
    std::vector<LeafHandle> items_;
    std::vector<uint32_t> mark_;
    size_t live_ = 0;
};

template <int D>
class MapCore {
public:
    using Tree = PRTree<D>;
    MapCore(const TreeParam& tp, float rtimes, int device) : tparam(tp), rtimes_(rtimes), device_(device) {}
    ~MapCore() { delete tree; if (ctx) gpis_destroy(ctx); }

    TreeParam tparam;
    float rtimes_;
    int device_;
    Tree* tree = nullptr;
    gpis_ctx* ctx = nullptr;
    ActiveSet active;                                 // activeSet (GPisMap3.h:91)
    std::map<uint64_t, std::array<int32_t, 3>> device_cells;   // leaves currently registered on the device
    std::map<uint64_t, std::array<float, 6>> device_boxes;     // effective boxes sent for them (only non-default ones)
    int last_trained = 0;
    float last_train_ms = 0.f;
    // device-side dirty set + training-set gather: the samples live on the device per leaf; the host
    // re-sends only the leaves whose samples changed (mutation log of the tree). GPIS_HOST_GATHER=1 keeps the first
    // path (host QueryRange per dirty leaf, gpis_leaves_update); results are identical.
    bool device_gather = std::getenv("GPIS_HOST_GATHER") == nullptr || std::atoi(std::getenv("GPIS_HOST_GATHER")) == 0;
    std::vector<std::pair<int32_t, uint32_t>> sample_log;

    bool ensure_ctx(const gpis_config& cfg) {
        if (ctx) return true;
        gpis_config c = cfg;
        c.device = device_;
        const int rc = gpis_create(&ctx, &c);
        if (rc != GPIS_OK) {
            std::fprintf(stderr, "gpismap_b200: gpis_create failed (%d): %s — there is no CPU fallback\n", rc,
                         ctx ? gpis_last_error(ctx) : "no context");
            if (ctx) { gpis_destroy(ctx); ctx = nullptr; }
            return false;
        }
        return true;
    }
    void ensure_tree() { if (!tree) { tree = new Tree(tparam); tree->mutation_log = &sample_log; } }   // addNewMeas, GPisMap3.cpp:573-575

    void reset() {   // GPisMap3::reset, GPisMap3.cpp:99-115
        delete tree; tree = nullptr;
        active.clear();
        device_cells.clear();
        device_boxes.clear();
        sample_log.clear();
        if (ctx) gpis_reset(ctx);
    }

    void cell_of(int cell_id, int32_t* out) const {
        const auto& n = tree->cell(cell_id);
        const double pitch = 2.0 * (double)tparam.cluster_half;
        for (int a = 0; a < 3; ++a) out[a] = (a < D) ? (int32_t)std::floor((double)n.c[a] / pitch) : 0;
    }
    static uint64_t pack(const int32_t* c) {
        return ((uint64_t)(uint32_t)(c[0] + (1 << 20)) << 42) | ((uint64_t)(uint32_t)(c[1] + (1 << 20)) << 21) |
               (uint64_t)(uint32_t)(c[2] + (1 << 20));
    }

    // IsNotNew + Insert + registration, the block shared by evalPoints and reEvalPoints
    // (GPisMap3.cpp:608-623, 541-556). Returns the sample id when the point went in AND registered
    // at least one leaf; -1 otherwise (the sample may still sit in the tree).
    int try_insert(const float* pos, std::vector<int>& touched, const float* full = nullptr) {
        touched.clear();
        const int at = tree->locate(pos);              // one look-up of the cluster-level cell for both steps
        if (tree->is_not_new_at(at, pos)) return -1;
        const int s = tree->new_sample(pos);
        if (full) {   // bulk load: the sample carries its data from the start, like ref_harness.cpp does
            Sample<D>& sm = tree->sample(s);
            for (int a = 0; a < D; ++a) sm.grad[a] = full[D + a];
            sm.val = full[2 * D]; sm.pose_sig = full[2 * D + 1]; sm.grad_sig = full[2 * D + 2];
        }
        const bool ok = tree->insert_at(at, s, touched);
        if (!ok || touched.empty()) return -1;
        return s;
    }
    void activate(const std::vector<int>& touched) {
        for (int c : touched) active.insert(LeafHandle{c, tree->cell(c).gen});
    }
    // every non-empty leaf becomes active: the next train_active() retrains the whole map on its current samples
    int activate_all() {
        if (!tree) return 0;
        std::vector<int> all;
        float zero[D];
        for (int a = 0; a < D; ++a) zero[a] = 0.f;
        tree->query_clusters(zero, 1.0e6f, all);
        for (int c : all) active.insert(LeafHandle{c, tree->cell(c).gen});
        return (int)all.size();
    }
    void drop_freed(const std::vector<int>& freed) {
        for (int c : freed) {
            // gen was bumped when the cell was freed; the stale handle carries gen-1
            active.erase(LeafHandle{c, tree->cell(c).gen - 1});
        }
    }

    // updateGPs: dirty set = active ∪ leaves whose box touches AABB(c, Rtimes*l), then one training
    // ball per dirty leaf, in QueryRange order, shipped as one CSR batch.
    // f-2 path: table first (new leaves, erases, effective boxes, root box), then the changed sample lists, then the
    // active cells; dirty-set expansion, ball gather and training happen on the device.
    // Returns 1 = done, 0 = not applicable (fall back to the host gather), -1 = error.
    int train_active_device() {
        if (!sync_table()) return -1;
        constexpr int W = 2 * D + 3;
        double tp0 = now_s();
        {
            std::sort(sample_log.begin(), sample_log.end());
            sample_log.erase(std::unique(sample_log.begin(), sample_log.end()), sample_log.end());
            // the changed leaves' sample lists are independent read-only walks of the tree: a few host threads, each
            // with its own part of the (sorted) log, concatenated in log order
            struct Part { std::vector<int32_t> cells, counts; std::vector<float> centres, samples; };
            const int nthr = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 8, (int)sample_log.size() / 64 + 1}));
            std::vector<Part> parts(nthr);
            auto gather = [&](int t) {
                Part& P = parts[t];
                const size_t b = sample_log.size() * t / nthr, e2 = sample_log.size() * (t + 1) / nthr;
                std::vector<int> ids;
                for (size_t i = b; i < e2; ++i) {
                    const auto& e = sample_log[i];
                    if (!tree->cell_alive(e.first, e.second)) continue;      // collapsed: sync_table erased it
                    const auto& n = tree->cell(e.first);
                    int32_t cc[3];
                    cell_of(e.first, cc);
                    ids.clear();
                    tree->collect_samples(e.first, ids);
                    if (ids.empty() && !device_cells.count(pack(cc))) continue;
                    for (int a = 0; a < D; ++a) { P.cells.push_back(cc[a]); P.centres.push_back(n.c[a]); }
                    for (int s : ids) {
                        const Sample<D>& sm = tree->sample(s);
                        for (int a = 0; a < D; ++a) P.samples.push_back(sm.pos[a]);
                        for (int a = 0; a < D; ++a) P.samples.push_back(sm.grad[a]);
                        P.samples.push_back(sm.val); P.samples.push_back(sm.pose_sig); P.samples.push_back(sm.grad_sig);
                    }
                    P.counts.push_back((int32_t)ids.size());
                }
            };
            if (nthr == 1) gather(0);
            else {
                std::vector<std::thread> th;
                for (int t = 1; t < nthr; ++t) th.emplace_back(gather, t);
                gather(0);
                for (auto& x : th) x.join();
            }
            std::vector<int32_t> cells, offsets(1, 0);
            std::vector<float> centres, samples;
            {
                size_t nc = 0, ns = 0;
                for (const Part& P : parts) { nc += P.cells.size(); ns += P.samples.size(); }
                cells.reserve(nc); centres.reserve(nc); samples.reserve(ns);
            }
            for (const Part& P : parts) {
                cells.insert(cells.end(), P.cells.begin(), P.cells.end());
                centres.insert(centres.end(), P.centres.begin(), P.centres.end());
                samples.insert(samples.end(), P.samples.begin(), P.samples.end());
                for (int32_t c : P.counts) offsets.push_back(offsets.back() + c);
            }
            sample_log.clear();
            const int nl = (int)offsets.size() - 1;
            if (nl > 0 && gpis_samples_set(ctx, nl, cells.data(), centres.data(), offsets.data(), samples.data()) != GPIS_OK) {
                std::fprintf(stderr, "gpismap_b200: gpis_samples_set failed: %s\n", gpis_last_error(ctx));
                return -1;
            }
        }
        g_prof_s[6] += now_s() - tp0; g_prof_n[6] += 1;
        std::vector<int32_t> acells;
        float radius = rtimes_ * tparam.cluster_half;
        for (const LeafHandle& h : active.sorted()) {
            if (!tree->cell_alive(h.cell, h.gen) || tree->is_empty_leaf(h.cell)) continue;
            int32_t cc[3];
            cell_of(h.cell, cc);
            for (int a = 0; a < D; ++a) acells.push_back(cc[a]);
            radius = rtimes_ * tree->cell(h.cell).half;                 // GPisMap3.cpp:733: Rtimes * l
        }
        active.clear();                                                   // GPisMap3.cpp:789
        if (acells.empty()) return 1;
        ProfScope ps(7);
        int32_t ntr = 0;
        const int rc = gpis_leaves_train_dirty(ctx, (int)acells.size() / D, acells.data(), radius, &ntr);
        if (rc != GPIS_OK) {
            std::fprintf(stderr, "gpismap_b200: gpis_leaves_train_dirty failed (%d): %s\n", rc, gpis_last_error(ctx));
            return -1;
        }
        gpis_stats st;
        gpis_get_stats(ctx, &st);
        last_trained = (int)st.last_train_leaves;
        last_train_ms = st.last_train_ms;
        return 1;
    }

    bool train_active() {
        last_trained = 0; last_train_ms = 0.f;
        if (!tree || !ctx) { active.clear(); return false; }
        if (device_gather) {
            const int r = train_active_device();
            if (r != 0) return r > 0;
        }
        sample_log.clear();
        std::set<int> update_set;
        std::vector<int> qs;
        double tp0 = now_s();
        for (const LeafHandle& h : active.sorted()) {
            if (!tree->cell_alive(h.cell, h.gen)) continue;
            update_set.insert(h.cell);
            const auto& n = tree->cell(h.cell);
            qs.clear();
            tree->query_clusters(n.c, rtimes_ * n.half, qs);     // GPisMap3.cpp:733-740
            for (int q : qs) update_set.insert(q);
        }
        active.clear();                                           // GPisMap3.cpp:789
        g_prof_s[5] += now_s() - tp0; g_prof_n[5] += 1; tp0 = now_s();
        // Training sets (QueryRange per dirty leaf, GPisMap3.cpp:705-709) are independent read-only tree
        // queries: gathered by a few host threads into per-chunk buffers, concatenated in leaf order.
        constexpr int W = 2 * D + 3;
        std::vector<int> cids;
        for (int cid : update_set)
            if (!tree->is_empty_leaf(cid)) cids.push_back(cid);   // an empty leaf is never a query candidate; its GP is unobservable
        struct Part { std::vector<int32_t> cells, counts; std::vector<float> centres, samples; };
        const int nthr = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, (int)cids.size() / 16 + 1}));
        std::vector<Part> parts(nthr);
        auto gather = [&](int t) {
            Part& P = parts[t];
            const size_t b = cids.size() * t / nthr, e = cids.size() * (t + 1) / nthr;
            std::vector<int> ids;
            for (size_t i = b; i < e; ++i) {
                const int cid = cids[i];
                const auto& n = tree->cell(cid);
                ids.clear();
                tree->query_range(n.c, n.half * rtimes_, ids);        // GPisMap3.cpp:705-709
                if (ids.empty()) continue;                             // GPisMap3.cpp:710
                int32_t cc[3];
                cell_of(cid, cc);
                for (int a = 0; a < D; ++a) { P.cells.push_back(cc[a]); P.centres.push_back(n.c[a]); }
                for (int s : ids) {
                    const Sample<D>& sm = tree->sample(s);
                    for (int a = 0; a < D; ++a) P.samples.push_back(sm.pos[a]);
                    for (int a = 0; a < D; ++a) P.samples.push_back(sm.grad[a]);
                    P.samples.push_back(sm.val); P.samples.push_back(sm.pose_sig); P.samples.push_back(sm.grad_sig);
                }
                P.counts.push_back((int32_t)ids.size());
            }
        };
        if (nthr == 1) gather(0);
        else {
            std::vector<std::thread> th;
            for (int t = 1; t < nthr; ++t) th.emplace_back(gather, t);
            gather(0);
            for (auto& x : th) x.join();
        }
        std::vector<int32_t> cells;
        std::vector<float> centres;
        std::vector<int32_t> offsets(1, 0);
        std::vector<float> samples;
        for (const Part& P : parts) {
            cells.insert(cells.end(), P.cells.begin(), P.cells.end());
            centres.insert(centres.end(), P.centres.begin(), P.centres.end());
            samples.insert(samples.end(), P.samples.begin(), P.samples.end());
            for (int32_t c : P.counts) offsets.push_back(offsets.back() + c);
        }
        for (size_t i = 0; i + D <= cells.size(); i += D) {
            int32_t cc[3] = {cells[i], cells[i + 1], D == 3 ? cells[i + 2] : 0};
            device_cells[pack(cc)] = {cc[0], cc[1], cc[2]};
        }
        const int nl = (int)offsets.size() - 1;
        g_prof_s[6] += now_s() - tp0; g_prof_n[6] += 1;
        if (nl > 0) {
            ProfScope ps(7);
            const int rc = gpis_leaves_update(ctx, nl, cells.data(), centres.data(), offsets.data(), samples.data(), nullptr);
            if (rc != GPIS_OK) {
                std::fprintf(stderr, "gpismap_b200: gpis_leaves_update failed (%d): %s\n", rc, gpis_last_error(ctx));
                return false;
            }
            gpis_stats st;
            gpis_get_stats(ctx, &st);
            last_trained = (int)st.last_train_leaves;
            last_train_ms = st.last_train_ms;
        }
        return sync_table();
    }

    // Make the device table list exactly the tree's non-empty leaves (QueryNonEmptyLevelC over
    // everything, octree.cpp:829-859) and tell it the root box.
    bool sync_table() {
        if (!tree || !ctx) return false;
        ProfScope ps(8);
        std::vector<int> all;
        float zero[D];
        for (int a = 0; a < D; ++a) zero[a] = 0.f;
        tree->query_clusters(zero, 1.0e6f, all);
        std::map<uint64_t, std::array<int32_t, 3>> current;
        std::vector<int32_t> mark_cells;
        std::vector<float> mark_centres;
        for (int cid : all) {
            int32_t cc[3];
            cell_of(cid, cc);
            const uint64_t k = pack(cc);
            current[k] = {cc[0], cc[1], cc[2]};
            if (!device_cells.count(k)) {
                for (int a = 0; a < D; ++a) { mark_cells.push_back(cc[a]); mark_centres.push_back(tree->cell(cid).c[a]); }
            }
        }
        std::vector<int32_t> erase_cells;
        for (auto& kv : device_cells)
            if (!current.count(kv.first))
                for (int a = 0; a < D; ++a) erase_cells.push_back(kv.second[a]);
        if (!erase_cells.empty() && gpis_leaves_erase(ctx, (int)erase_cells.size() / D, erase_cells.data()) != GPIS_OK) return false;
        if (!mark_cells.empty() && gpis_leaves_mark(ctx, (int)mark_cells.size() / D, mark_cells.data(), mark_centres.data()) != GPIS_OK) return false;
        device_cells.swap(current);
        // effective (ancestor-intersected) boxes: send the ones that differ from centre -/+ half or changed
        {
            std::vector<int32_t> bcells;
            std::vector<float> bvals;
            std::map<uint64_t, std::array<float, 6>> now;
            for (int cid : all) {
                float lo[D], hi[D];
                tree->effective_box(cid, lo, hi);
                const auto& n = tree->cell(cid);
                bool dflt = true;
                for (int a = 0; a < D; ++a) dflt = dflt && lo[a] == n.lo[a] && hi[a] == n.hi[a];
                int32_t cc[3];
                cell_of(cid, cc);
                const uint64_t k = pack(cc);
                std::array<float, 6> b{};
                for (int a = 0; a < D; ++a) { b[a] = lo[a]; b[3 + a] = hi[a]; }
                auto prev = device_boxes.find(k);
                const bool had = prev != device_boxes.end();
                if (!dflt) now[k] = b;
                if ((!dflt && (!had || prev->second != b)) || (dflt && had)) {
                    for (int a = 0; a < D; ++a) bcells.push_back(cc[a]);
                    for (int a = 0; a < D; ++a) bvals.push_back(lo[a]);
                    for (int a = 0; a < D; ++a) bvals.push_back(hi[a]);
                }
            }
            device_boxes.swap(now);
            if (!bcells.empty() && gpis_leaves_set_boxes(ctx, (int)bcells.size() / D, bcells.data(), bvals.data()) != GPIS_OK) return false;
        }
        // root box
        const auto& r = tree->cell(tree->root());
        const double pitch = 2.0 * (double)tparam.cluster_half;
        int32_t rmin[3] = {0, 0, 0};
        for (int a = 0; a < D; ++a) rmin[a] = (int32_t)std::llround(((double)r.c[a] - (double)r.half) / pitch);
        const int levels = (int)std::lround(std::log2((double)r.half / (double)tparam.cluster_half));
        return gpis_rebase(ctx, rmin, levels) == GPIS_OK;
    }

    bool query(const float* x, int leng, float* res) {
        if (!ctx || !tree) return false;   // the reference dereferences a null tree here (GPisMap3.cpp:814)
        return gpis_query(ctx, x, leng, res) == GPIS_OK;
    }

    void all_points(std::vector<float>& pos) const {
        pos.clear();
        if (!tree) return;
        std::vector<int> ids;
        tree->collect_samples(tree->root(), ids);
        pos.reserve(ids.size() * D);
        for (int s : ids) for (int a = 0; a < D; ++a) pos.push_back(tree->sample(s).pos[a]);
    }
    void all_samples(std::vector<float>& out) const {
        out.clear();
        if (!tree) return;
        std::vector<int> ids;
        tree->collect_samples(tree->root(), ids);
        for (int s : ids) {
            const Sample<D>& sm = tree->sample(s);
            for (int a = 0; a < D; ++a) out.push_back(sm.pos[a]);
            for (int a = 0; a < D; ++a) out.push_back(sm.grad[a]);
            out.push_back(sm.val); out.push_back(sm.pose_sig); out.push_back(sm.grad_sig);
        }
    }
    void leaves(std::vector<float>& centres, std::vector<int>& counts) const {
        centres.clear(); counts.clear();
        if (!tree) return;
        std::vector<int> all;
        float zero[D];
        for (int a = 0; a < D; ++a) zero[a] = 0.f;
        tree->query_clusters(zero, 1.0e6f, all);
        for (int cid : all) {
            for (int a = 0; a < D; ++a) centres.push_back(tree->cell(cid).c[a]);
            counts.push_back(tree->count(cid));
        }
    }
    // bulk load, mirroring oracle/ref_harness.cpp ref3_insert_samples
    int insert_samples(const float* smp, int n) {
        ensure_tree();
        constexpr int W = 2 * D + 3;
        std::vector<int> touched;
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const float* s = smp + (size_t)i * W;
            const int id = try_insert(s, touched, s);
            if (id < 0) continue;
            activate(touched);
            ++cnt;
        }
        return cnt;
    }
};

// occ_test (GPisMap3.cpp:38-41, GPisMap.cpp:39-42): float arguments, double inside, float result
inline float occ_test(float rinv, float rinv0, float a) {
    return (float)(2.0 * (1.0 / (1.0 + std::exp((double)(-a * (rinv - rinv0)))) - 0.5));
}
inline float saturate(float val, float min_val, float max_val) { return std::min(std::max(val, min_val), max_val); }

}  // namespace gpismap_host
