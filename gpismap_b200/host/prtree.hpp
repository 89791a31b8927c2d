// Host-side point-region tree (quadtree for D = 2, octree for D = 3) that decides leaf assignment.
//
// Behavioural contract = the reference's QuadTree / OcTree (cpp/src/quadtree.cpp, cpp/src/octree.cpp,
// cpp/include/octree.h, cpp/include/quadtree.h): same strict containsPoint / inclusive
// intersectsAABB float tests, same child visiting order (NW,NE,SW,SE[,then the back four]), same
// subdivision / minimum spacing / root growth rules, so that sample-to-leaf assignment, training
// set order (QueryRange DFS order = row order of K) and candidate lists stay bit-exact. The
// implementation is new: one template for both dimensions, nodes in an index-linked arena with
// children allocated as one block, samples in a pool, no per-node heap objects or shared_ptr.
//
// This is bookkeeping around the GPU hot path (SURVEY.md §8a "adjacent but not on the GPU path").
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace gpismap_host {

template <int D>
struct Sample {
    float pos[D];
    float grad[D];
    float val = 0.f, pose_sig = 0.f, grad_sig = 0.f;   // defaults = Node3(Point3) ctor (strct.cpp:72-79)
    Sample() { for (int c = 0; c < D; ++c) { pos[c] = 0.f; grad[c] = 0.f; } }
};

struct TreeParam {           // cpp/include/strct.h:175-199
    float initroot_half, min_half, min_half_sqr, max_half, cluster_half;
    float reg_tol_insert;    // cluster registration tolerance on insert: 1e-6 (octree.cpp:325) / 1e-3 (quadtree.cpp:262)
    bool reg_at_maxdepth;    // quadtree.cpp:240-241 registers at max depth, octree.cpp:305-313 does not
    TreeParam(float mi, float ma, float ini, float c, float tol, bool rmd)
        : initroot_half(ini), min_half(mi), min_half_sqr(mi * mi), max_half(ma), cluster_half(c),
          reg_tol_insert(tol), reg_at_maxdepth(rmd) {}
};

template <int D>
class PRTree {
public:
    static constexpr int NCH = 1 << D;
    static constexpr int kUnsafe = -1, kAbsent = -2;   // find_cluster: take the full descent / no cluster-level cell there
    struct Cell {
        float c[D];
        float half;
        float lo[D], hi[D];      // c - half, c + half in float (AABB ctor, octree.h:73-78)
        int32_t parent = -1;
        int32_t child0 = -1;     // first of NCH consecutive children, -1 = leaf
        int32_t sample = -1;
        uint32_t gen = 0;        // bumped when the cell is freed (stale-handle detection)
        bool max_depth = false, root_limit = false, alive = true;
    };

    explicit PRTree(const TreeParam& p) : P(p) {
        float c0[D];
        for (int a = 0; a < D; ++a) c0[a] = 0.f;     // GPisMap3.cpp:574, GPisMap.cpp:460: rooted at the origin
        root_ = new_cell(c0, P.initroot_half, -1, /*derive_flags=*/false);
    }

    int root() const { return root_; }
    const Cell& cell(int i) const { return cells_[i]; }
    int32_t count(int i) const { return count_[i]; }       // numNodes
    const Sample<D>& sample(int s) const { return samples_[s]; }
    Sample<D>& sample(int s) { return samples_[s]; }
    bool cell_alive(int i, uint32_t gen) const { return i >= 0 && i < (int)cells_.size() && cells_[i].alive && cells_[i].gen == gen; }
    bool is_cluster_level(int i) const { return std::fabs((double)cells_[i].half - (double)P.cluster_half) < 1e-3; }
    bool is_empty_leaf(int i) const { return cells_[i].child0 < 0 && cells_[i].sample < 0; }
    const TreeParam& param() const { return P; }

    int new_sample(const float* pos) {
        Sample<D> s;
        for (int a = 0; a < D; ++a) s.pos[a] = pos[a];
        samples_.push_back(s);
        return (int)samples_.size() - 1;
    }

    // ---- IsNotNew (octree.cpp:431-458)
    // locate(p) = the cluster-level cell to enter at (or kUnsafe / kAbsent); valid until the next structural change
    // above the cluster level, i.e. it may be shared by an is_not_new_at / insert_at pair on the same point
    int locate(const float* p) const { return find_cluster(p); }
    bool is_not_new(const float* p) const { return is_not_new_at(find_cluster(p), p); }
    bool is_not_new_at(int c, const float* p) const {
        if (c >= 0) return is_not_new(c, p);      // same recursion, entered at the leaf's cluster-level cell
        if (c == kAbsent) return false;            // no cluster-level cell there: the descent ends in an empty leaf
        return is_not_new(root_, p);
    }

    // ---- Insert(n, quads) from the root (octree.cpp:295-414 / quadtree.cpp:223-312); follows root
    // growth (t = t->getRoot(), GPisMap3.cpp:615-618). `touched` receives the cluster-level cells the
    // insertion registered (vecInserted).
    bool insert(int s, std::vector<int>& touched) { return insert_at(find_cluster(samples_[s].pos), s, touched); }
    bool insert_at(int c, int s, std::vector<int>& touched) {
        if (c >= 0) {
            // enter the recursion at the cluster-level cell; the ancestors would only have passed the call down
            // (their own boxes contain the point, one child can take it) and refreshed their counts on the way back
            const bool ok = insert_rec(c, s, &touched);
            if (ok) for (int a = cells_[c].parent; a >= 0; a = cells_[a].parent) update_count(a);
            return ok;
        }
        const bool ok = insert_rec(root_, s, &touched);
        while (cells_[root_].parent >= 0) root_ = cells_[root_].parent;
        return ok;
    }

    // ---- Remove(n, set) (octree.cpp:510-566); cells deleted by a collapse are reported in `freed`
    // (the reference erases them from the active set).
    bool remove_tracked(int s, std::vector<int>& freed) { return remove_from(samples_[s].pos, true, &freed); }
    // ---- Remove(n) (octree.cpp:460-508): visits every child, collapses, reports freed cells too
    bool remove_plain(int s, std::vector<int>& freed) { return remove_from(samples_[s].pos, false, &freed); }

    // ---- QueryRange (octree.cpp:777-804): samples with |p - c|^2 < (half)^2, DFS order
    void query_range(const float* c, float half, std::vector<int>& out) const {
        float lo[D], hi[D];
        for (int a = 0; a < D; ++a) { lo[a] = c[a] - half; hi[a] = c[a] + half; }
        query_range_rec(root_, c, half * half, lo, hi, out);
    }

    // ---- QueryNonEmptyLevelC (octree.cpp:829-893): non-empty cluster-level cells hitting the box
    void query_clusters(const float* c, float half, std::vector<int>& out) const {
        float lo[D], hi[D];
        for (int a = 0; a < D; ++a) { lo[a] = c[a] - half; hi[a] = c[a] + half; }
        query_clusters_rec(root_, lo, hi, out);
    }

    // ---- getAllChildrenNonEmptyNodes (octree.cpp:806-827)
    void collect_samples(int cell_id, std::vector<int>& out) const {
        const Cell& n = cells_[cell_id];
        if (n.child0 < 0) { if (n.sample >= 0) out.push_back(n.sample); return; }
        for (int k = 0; k < NCH; ++k) collect_samples(n.child0 + k, out);
    }

    // Intersection of the float boxes of a cell and all its ancestors: what the level-by-level pruning
    // of QueryNonEmptyLevelC (octree.cpp:864-867) lets through.
    void effective_box(int cell_id, float* lo, float* hi) const {
        for (int a = 0; a < D; ++a) { lo[a] = cells_[cell_id].lo[a]; hi[a] = cells_[cell_id].hi[a]; }
        for (int p = cells_[cell_id].parent; p >= 0; p = cells_[p].parent)
            for (int a = 0; a < D; ++a) { lo[a] = std::max(lo[a], cells_[p].lo[a]); hi[a] = std::min(hi[a], cells_[p].hi[a]); }
    }

    // Every placement or removal of a sample is reported as the (cluster-level cell, generation) it happened under,
    // so that the owner can keep a per-leaf copy of the samples elsewhere (the device sample store).
    std::vector<std::pair<int32_t, uint32_t>>* mutation_log = nullptr;

    // a sample's fields were changed in place (reEvalPoints doubling the noises, GPisMap3.cpp:451-454): log its leaf
    void touch(int s) {
        if (!mutation_log) return;
        const float* p = samples_[s].pos;
        int id = root_;
        while (id >= 0) {
            if (is_cluster_level(id)) { mutation_log->push_back({id, cells_[id].gen}); return; }
            if (cells_[id].child0 < 0) return;
            int nx = -1;
            for (int k = 0; k < NCH; ++k)
                if (contains(cells_[cells_[id].child0 + k], p)) { nx = cells_[id].child0 + k; break; }
            id = nx;
        }
    }

    size_t num_cells() const { return cells_.size(); }
    size_t num_samples() const { return samples_.size(); }

private:
    TreeParam P;
    std::vector<Cell> cells_;
    std::vector<int32_t> count_;         // numNodes per cell, kept apart from the 60-byte cells: the ancestor walks of
                                         // every insert / remove read the eight children's counts (one 32-byte run)
    std::vector<Sample<D>> samples_;
    std::vector<int32_t> free_blocks_;   // recycled child blocks
    int root_ = -1;

    static float sqdist(const float* a, const float* b) {   // octree.cpp:24-31
        float s = 0.f;
        for (int c = 0; c < D; ++c) { const float d = a[c] - b[c]; s = (c == 0) ? d * d : s + d * d; }
        return s;
    }
    // child k of the reference's visiting order: bit0 = east (+x); bit1 = south (-y); bit2 = back (-z)
    static float child_sign(int k, int axis) {
        if (axis == 0) return (k & 1) ? 1.f : -1.f;
        return ((k >> axis) & 1) ? -1.f : 1.f;
    }
    void init_cell(Cell& n, const float* c, float half, int parent, bool derive_flags) {
        for (int a = 0; a < D; ++a) { n.c[a] = c[a]; n.lo[a] = c[a] - half; n.hi[a] = c[a] + half; }
        n.half = half; n.parent = parent; n.child0 = -1; n.sample = -1; n.alive = true;
        n.max_depth = derive_flags && (half < P.min_half);     // octree.cpp:72-75
        n.root_limit = derive_flags && (half > P.max_half);
    }
    int new_cell(const float* c, float half, int parent, bool derive_flags) {
        cells_.emplace_back();
        count_.push_back(0);
        init_cell(cells_.back(), c, half, parent, derive_flags);
        return (int)cells_.size() - 1;
    }
    int alloc_children() {
        if (!free_blocks_.empty()) { const int b = free_blocks_.back(); free_blocks_.pop_back(); return b; }
        const int b = (int)cells_.size();
        cells_.resize(cells_.size() + NCH);
        count_.resize(cells_.size(), 0);
        return b;
    }
    bool contains(const Cell& n, const float* p) const {    // strict, octree.h:119-126
        for (int a = 0; a < D; ++a) if (!(p[a] > n.lo[a] && p[a] < n.hi[a])) return false;
        return true;
    }
    static bool intersects(const Cell& n, const float* lo, const float* hi) {   // inclusive, octree.h:128-135
        for (int a = 0; a < D; ++a) if (hi[a] < n.lo[a] || lo[a] > n.hi[a]) return false;
        return true;
    }
    // The one child whose box can contain p, when that is certain. The reference tries the children in visiting
    // order and each child tests its own float box (c +- l computed in float, so neighbouring boxes can overlap or
    // leave a gap of an ulp around the parent's centre planes). Away from those planes — by more than a few ulps
    // of |c| + half — every other child rejects p on its first comparison and has no side effect, so descending
    // into this child alone is the same computation; near a plane the callers fall back to the full scan.
    int certain_child(const Cell& n, const float* p) const {
        int k = 0;
        for (int a = 0; a < D; ++a) {
            const float d = p[a] - n.c[a];
            const float tol = 9.6e-7f * (std::fabs(n.c[a]) + n.half);   // 8 ulp
            if (!(std::fabs(d) > tol)) return -1;
            const bool up = d > 0.f;
            if (a == 0) { if (up) k |= 1; }
            else if (!up) k |= (1 << a);
        }
        return k;
    }
    // Subdivide (octree.cpp:672-713): children at c +- half/2, visiting order as above
    void subdivide(int id, int except_k = -1, int except_cell = -1) {
        const int b = alloc_children();
        const float l = (float)((double)cells_[id].half * 0.5);
        for (int k = 0; k < NCH; ++k) {
            float c[D];
            for (int a = 0; a < D; ++a) c[a] = cells_[id].c[a] + child_sign(k, a) * l;
            Cell& ch = cells_[b + k];
            const uint32_t g = ch.gen;
            init_cell(ch, c, l, id, true);
            ch.gen = g;
            count_[b + k] = 0;
            index_add(b + k);
        }
        cells_[id].child0 = b;
        (void)except_k; (void)except_cell;
    }
    void update_count(int id) {
        const int b = cells_[id].child0;
        if (b < 0) return;
        int s = 0;
        for (int k = 0; k < NCH; ++k) s += count_[b + k];
        count_[id] = s;
    }
    // ---- direct entry at the cluster level. Every descent from the root passes 7+ levels whose only effect, for a
    // point that is not within a few ulps of a lattice plane, is to hand the call to the one child that contains it.
    // The index maps a cluster-level lattice cell to its tree cell; points near a plane (or trees with an orphaned
    // subtree) take the full descent, so the visited nodes and every float comparison that can matter are the same.
    // open addressing, linear probing; value 0 = empty slot, -1 = tombstone, otherwise cell id + 1
    struct ClusterIndex {
        std::vector<uint64_t> keys;
        std::vector<int32_t> vals;
        size_t used = 0, live = 0;
        ClusterIndex() : keys(1024, 0), vals(1024, 0) {}
        static size_t mix(uint64_t k) { k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 32; return (size_t)k; }
        int32_t find(uint64_t k) const {
            const size_t m = keys.size() - 1;
            for (size_t i = mix(k) & m;; i = (i + 1) & m) {
                if (vals[i] == 0) return -1;
                if (vals[i] > 0 && keys[i] == k) return vals[i] - 1;
            }
        }
        void set(uint64_t k, int32_t id) {
            if ((used + 1) * 2 > keys.size()) rehash();
            const size_t m = keys.size() - 1;
            size_t tomb = (size_t)-1;
            for (size_t i = mix(k) & m;; i = (i + 1) & m) {
                if (vals[i] == 0) {
                    if (tomb != (size_t)-1) i = tomb; else ++used;
                    keys[i] = k; vals[i] = id + 1; ++live;
                    return;
                }
                if (vals[i] < 0) { if (tomb == (size_t)-1) tomb = i; }
                else if (keys[i] == k) { vals[i] = id + 1; return; }
            }
        }
        void erase_if(uint64_t k, int32_t id) {
            const size_t m = keys.size() - 1;
            for (size_t i = mix(k) & m;; i = (i + 1) & m) {
                if (vals[i] == 0) return;
                if (vals[i] > 0 && keys[i] == k) { if (vals[i] - 1 == id) { vals[i] = -1; --live; } return; }
            }
        }
        void rehash() {
            std::vector<uint64_t> ok; std::vector<int32_t> ov;
            ok.swap(keys); ov.swap(vals);
            const size_t cap = std::max<size_t>(1024, live * 4 <= ok.size() ? ok.size() : ok.size() * 2);
            keys.assign(cap, 0); vals.assign(cap, 0);
            used = 0; live = 0;
            for (size_t i = 0; i < ok.size(); ++i) if (ov[i] > 0) set(ok[i], ov[i] - 1);
        }
    };
    ClusterIndex cluster_index_;
    // the previous look-up: consecutive pixels of a depth image mostly fall into the same leaf
    mutable uint64_t last_key_ = 0;
    mutable int32_t last_id_ = kUnsafe;   // kUnsafe = nothing cached
    static uint64_t lattice_key(const long long* i) {
        uint64_t k = 0;
        for (int a = 0; a < D; ++a) k = k * 0x9E3779B97F4A7C15ull + (uint64_t)(i[a] + (1ll << 40));
        return k;
    }
    uint64_t key_of_cell(int id) const {
        long long i[D];
        const double pitch = 2.0 * (double)P.cluster_half;
        for (int a = 0; a < D; ++a) i[a] = (long long)std::floor((double)cells_[id].c[a] / pitch);
        return lattice_key(i);
    }
    void index_add(int id) { if (is_cluster_level(id)) { cluster_index_.set(key_of_cell(id), id); last_id_ = kUnsafe; } }
    void index_del(int id) {
        if (!is_cluster_level(id)) return;
        cluster_index_.erase_if(key_of_cell(id), id);
        last_id_ = kUnsafe;
    }
    int find_cluster(const float* p) const {
        if (orphaned_ || !contains(cells_[root_], p)) return kUnsafe;
        const double pitch = 2.0 * (double)P.cluster_half;
        const float rh = cells_[root_].half;
        long long i[D];
        for (int a = 0; a < D; ++a) {
            const double q = (double)p[a] / pitch;
            const double f = std::floor(q);
            // distance to the nearest lattice plane, against the largest tolerance any level of the descent would use
            // (certain_child: 8 ulp of |c| + half, here with the root's half and the root centre's magnitude)
            const double tol = 4.0 * 9.6e-7 * ((double)std::fabs(p[a]) + (double)std::fabs(cells_[root_].c[a]) + 2.0 * (double)rh);
            const double dist = std::min(q - f, f + 1.0 - q) * pitch;
            if (!(dist > tol)) return kUnsafe;
            i[a] = (long long)f;
        }
        const uint64_t key = lattice_key(i);
        if (last_id_ != kUnsafe && last_key_ == key) return last_id_;
        const int32_t id = cluster_index_.find(key);
        last_key_ = key;
        last_id_ = id < 0 ? kAbsent : id;
        return last_id_;
    }
    // Remove entered at the cluster-level cell, then the ancestors' part of remove_rec (collapse when all children are
    // empty leaves, refresh the count) walked upwards
    bool remove_from(const float* p, bool short_circuit, std::vector<int>* freed) {
        const int c = find_cluster(p);
        if (c == kAbsent) return false;
        if (c < 0) return remove_rec(root_, p, short_circuit, freed);
        const bool res = remove_rec(c, p, short_circuit, freed);
        for (int a = cells_[c].parent; a >= 0; a = cells_[a].parent) remove_finish(a, res, freed);
        return res;
    }
    void remove_finish(int id, bool res, std::vector<int>* freed) {
        if (res && cells_[id].child0 >= 0) {
            // a live empty leaf always carries count 0 (every site that clears `sample` of a leaf clears the count),
            // so the eight-cell scan can only succeed when the children's counts sum to 0
            bool all_empty = true;
            for (int k = 0; k < NCH; ++k) all_empty = all_empty && count_[cells_[id].child0 + k] == 0;
            for (int k = 0; all_empty && k < NCH; ++k) all_empty = is_empty_leaf(cells_[id].child0 + k);
            if (all_empty) {
                const int b = cells_[id].child0;
                for (int k = 0; k < NCH; ++k) {
                    if (freed) freed->push_back(b + k);
                    index_del(b + k);
                    cells_[b + k].alive = false;
                    cells_[b + k].gen++;
                }
                free_blocks_.push_back(b);
                cells_[id].child0 = -1;
                count_[id] = 0;
            }
        }
        update_count(id);
    }
    void log_mutation(int id) {
        if (!mutation_log) return;
        int c = id;
        while (c >= 0 && (double)cells_[c].half < (double)P.cluster_half - 1e-3) c = cells_[c].parent;
        if (c >= 0 && is_cluster_level(c)) mutation_log->push_back({c, cells_[c].gen});
    }
    bool reg_ok(const Cell& n) const { return std::fabs((double)n.half - (double)P.cluster_half) < (double)P.reg_tol_insert; }

    // InsertToParent (octree.cpp:151-212): the root grows toward the point; the old root becomes
    // the child diagonally opposite, and the insertion continues WITHOUT cluster registration
    // (it calls the overload that takes no set).
    bool insert_to_parent(int id, int s) {
        const float* np = samples_[s].pos;
        const float l = cells_[id].half;
        float pc[D];
        for (int a = 0; a < D; ++a) pc[a] = 0.f;     // Point3<float> par_c default-constructs to 0
        int child_k = -1;
        // the chain of ifs in the reference requires a strict inequality on every axis
        bool strict = true;
        for (int a = 0; a < D; ++a) if (!(np[a] < cells_[id].c[a] || np[a] > cells_[id].c[a])) strict = false;
        if (strict) {
            child_k = 0;
            for (int a = 0; a < D; ++a) {
                const bool up = np[a] > cells_[id].c[a];
                pc[a] = up ? cells_[id].c[a] + l : cells_[id].c[a] - l;
                // the old root sits on the side away from the point: sign = up ? -1 : +1
                const float sgn = up ? -1.f : 1.f;
                if (a == 0) { if (sgn > 0) child_k |= 1; }
                else { if (sgn < 0) child_k |= (1 << a); }
            }
        }
        const float ph = (float)(2.0 * (double)l);
        const int par = new_cell(pc, ph, -1, true);
        if (child_k >= 0) {
            // SubdivideExcept (octree.cpp:715-775): allocate the block, then let the old root take
            // the place of child_k. Arena children must be contiguous, so the old root's contents
            // are moved into the block slot and every reference to it is re-pointed.
            const int b = alloc_children();
            const float hl = (float)((double)ph * 0.5);
            for (int k = 0; k < NCH; ++k) {
                if (k == child_k) continue;
                float c[D];
                for (int a = 0; a < D; ++a) c[a] = pc[a] + child_sign(k, a) * hl;
                Cell& ch = cells_[b + k];
                const uint32_t g = ch.gen;
                init_cell(ch, c, hl, par, true);
                ch.gen = g;
                count_[b + k] = 0;
                index_add(b + k);
            }
            move_cell(id, b + child_k);
            cells_[b + child_k].parent = par;
            cells_[par].child0 = b;
        }
        // child_k < 0: a point exactly on one of the root's centre planes. The reference then builds
        // a fresh childless parent at the origin and the old subtree is orphaned (octree.cpp:96-99).
        if (child_k < 0) { cells_[id].parent = par; orphaned_ = true; }
        root_ = par;
        return insert_rec(par, s, nullptr);
    }
    // move cell `from` (with its subtree links) to slot `to`
    void move_cell(int from, int to) {
        const uint32_t g = cells_[to].gen;
        cells_[to] = cells_[from];
        cells_[to].gen = g;
        count_[to] = count_[from];
        if (cells_[to].child0 >= 0)
            for (int k = 0; k < NCH; ++k) cells_[cells_[to].child0 + k].parent = to;
        index_del(from);
        index_add(to);
        cells_[from].alive = false;
        cells_[from].gen++;
        cells_[from].child0 = -1; cells_[from].sample = -1;
        moved_from_ = from; moved_to_ = to;
    }

    bool insert_rec(int id, int s, std::vector<int>* quads) {
        const float* p = samples_[s].pos;
        if (!contains(cells_[id], p)) {
            if (cells_[id].parent < 0) {
                if (cells_[id].root_limit) return false;
                return insert_to_parent(id, s);
            }
            return false;
        }
        if (cells_[id].max_depth) {
            if (cells_[id].sample < 0) {
                cells_[id].sample = s; count_[id] = 1;
                log_mutation(id);
                if (quads && P.reg_at_maxdepth && reg_ok(cells_[id])) quads->push_back(id);
                return true;
            }
            return false;
        }
        if (cells_[id].child0 < 0) {
            if (cells_[id].half > P.cluster_half) {
                subdivide(id);
            } else {
                if (cells_[id].sample < 0) {
                    cells_[id].sample = s; count_[id] = 1;
                    log_mutation(id);
                    if (quads && reg_ok(cells_[id])) quads->push_back(id);
                    return true;
                }
                if (sqdist(samples_[cells_[id].sample].pos, p) < P.min_half_sqr) return false;
                const int old = cells_[id].sample;
                subdivide(id);
                {
                    const int ck = certain_child(cells_[id], samples_[old].pos);
                    if (ck >= 0) insert_rec(cells_[id].child0 + ck, old, quads);
                    else
                        for (int k = 0; k < NCH; ++k)
                            if (insert_rec(cells_[id].child0 + k, old, quads)) break;
                }
                cells_[id].sample = -1;   // dropped silently if no child took it (lattice-plane rejection)
                log_mutation(id);
            }
        }
        const int ck = certain_child(cells_[id], p);
        for (int k = (ck >= 0 ? ck : 0); k < (ck >= 0 ? ck + 1 : NCH); ++k) {
            if (insert_rec(cells_[id].child0 + k, s, quads)) {
                if (quads && reg_ok(cells_[id])) quads->push_back(id);
                update_count(id);
                return true;
            }
        }
        return false;
    }

    bool is_not_new(int id, const float* p) const {
        const Cell& n = cells_[id];
        if (!contains(n, p)) return false;
        if (n.child0 < 0 && n.sample < 0) return false;
        if (n.sample >= 0 && sqdist(samples_[n.sample].pos, p) < P.min_half_sqr) return true;
        if (n.child0 < 0) return false;
        const int ck = certain_child(n, p);
        if (ck >= 0) return is_not_new(n.child0 + ck, p);
        for (int k = 0; k < NCH; ++k) if (is_not_new(n.child0 + k, p)) return true;
        return false;
    }

    bool remove_rec(int id, const float* p, bool short_circuit, std::vector<int>* freed) {
        if (!contains(cells_[id], p)) return false;
        if (is_empty_leaf(id)) return false;
        if (cells_[id].sample >= 0 && (double)sqdist(samples_[cells_[id].sample].pos, p) < 1e-12) {   // EPS, octree.cpp:22
            cells_[id].sample = -1; count_[id] = 0;
            log_mutation(id);
            return true;
        }
        if (cells_[id].child0 < 0) return false;
        bool res = false;
        const int ck = certain_child(cells_[id], p);
        for (int k = (ck >= 0 ? ck : 0); k < (ck >= 0 ? ck + 1 : NCH); ++k) {
            if (short_circuit && res) break;
            res = remove_rec(cells_[id].child0 + k, p, short_circuit, freed) || res;
        }
        remove_finish(id, res, freed);
        return res;
    }

    void query_range_rec(int id, const float* c, float r2, const float* lo, const float* hi, std::vector<int>& out) const {
        const Cell& n = cells_[id];
        if (!intersects(n, lo, hi) || (n.child0 < 0 && n.sample < 0)) return;
        if (n.child0 < 0) {
            if (sqdist(samples_[n.sample].pos, c) < r2) out.push_back(n.sample);
            return;
        }
        for (int k = 0; k < NCH; ++k) query_range_rec(n.child0 + k, c, r2, lo, hi, out);
    }
    void query_clusters_rec(int id, const float* lo, const float* hi, std::vector<int>& out) const {
        const Cell& n = cells_[id];
        if (!intersects(n, lo, hi) || (n.child0 < 0 && n.sample < 0)) return;
        const bool above = (double)n.half > (double)P.cluster_half + 0.001;
        if (n.child0 < 0 && above) return;
        if (above) { for (int k = 0; k < NCH; ++k) query_clusters_rec(n.child0 + k, lo, hi, out); }
        else out.push_back(id);
    }

public:
    bool orphaned_ = false;
    int moved_from_ = -1, moved_to_ = -1;   // last root relocation (insert_to_parent), for handle fix-up
};

}  // namespace gpismap_host
