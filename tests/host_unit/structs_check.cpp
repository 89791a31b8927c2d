// TEST SCAFFOLDING (tests/test_host_structs.py): invariants of the host-side containers behind the serial passes of
// update() — the list-based active set, the per-cell count array and the cluster-level index of PRTree.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "map_core.hpp"

extern "C" void gpis_destroy(gpis_ctx*) {}

using namespace gpismap_host;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); if (++fails > 20) std::exit(1); } } while (0)

static void check_active_set() {
    std::mt19937 rng(3);
    ActiveSet a;
    std::set<LeafHandle> ref;
    for (int it = 0; it < 200000; ++it) {
        const int op = rng() % 100;
        LeafHandle h{(int)(rng() % 300), (uint32_t)(rng() % 3)};
        if (op < 70) {
            // a cell is only ever active with its current generation: activating a newer one supersedes the older
            for (uint32_t g = 0; g < 3; ++g) if (g != h.gen) ref.erase(LeafHandle{h.cell, g});
            a.insert(h); ref.insert(h);
        } else if (op < 95) { a.erase(h); ref.erase(h); }
        else if (op < 97) { a.clear(); ref.clear(); }
        if (it % 997 == 0) {
            std::vector<LeafHandle> s = a.sorted();
            CHECK(s.size() == ref.size() && a.size() == ref.size(), "active set size %zu / %zu vs %zu", s.size(), a.size(), ref.size());
            size_t i = 0;
            for (const LeafHandle& r : ref) { if (i < s.size()) CHECK(s[i].cell == r.cell && s[i].gen == r.gen, "order differs at %zu", i); ++i; }
        }
    }
}

template <int D>
static void check_tree(unsigned seed, int nops) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(-1.5f, 1.5f);
    const TreeParam tp = D == 3 ? TreeParam(0.0125f / 2.0f, 1.6f, 0.4f, 0.025f, 1e-6f, false)
                                : TreeParam(0.025f, 102.4f, 12.8f, 0.4f, 1e-3f, true);
    PRTree<D> t(tp);
    std::vector<int> live, touched, freed, ids;
    for (int it = 0; it < nops; ++it) {
        if (live.empty() || rng() % 100 < 65) {
            float p[D];
            // clustered points: many share a cluster-level cell, some fall next to lattice planes
            const float cx = 0.05f * (float)((int)(rng() % 40) - 20);
            for (int a = 0; a < D; ++a) p[a] = (D == 3 ? cx : 8.f * cx) + (D == 3 ? 0.03f : 0.5f) * U(rng) + (rng() % 50 == 0 ? 0.f : 1e-3f * U(rng));
            const int at = t.locate(p);
            const bool known = t.is_not_new_at(at, p);
            CHECK(known == t.is_not_new(p), "is_not_new_at disagrees with is_not_new");
            if (known) continue;
            const int s = t.new_sample(p);
            touched.clear();
            if (t.insert_at(at, s, touched)) live.push_back(s);
        } else {
            const size_t k = rng() % live.size();
            freed.clear();
            const bool ok = (rng() & 1) ? t.remove_tracked(live[k], freed) : t.remove_plain(live[k], freed);
            CHECK(ok, "a live sample could not be removed");
            live[k] = live.back(); live.pop_back();
        }
        if (it % 5000 == 4999 || it == nops - 1) {
            // counts: every inner cell = sum of its children; the root counts the samples in the tree
            ids.clear();
            t.collect_samples(t.root(), ids);
            CHECK((int)ids.size() == t.count(t.root()), "root count %d vs %zu samples", t.count(t.root()), ids.size());
            CHECK(ids.size() == live.size(), "tree holds %zu samples, expected %zu", ids.size(), live.size());
            for (int c = 0; c < (int)t.num_cells(); ++c) {
                const auto& n = t.cell(c);
                if (!n.alive) continue;
                if (n.child0 >= 0) {
                    int sum = 0;
                    for (int k = 0; k < PRTree<D>::NCH; ++k) sum += t.count(n.child0 + k);
                    // (a count goes stale only on the reference's own "insertion failed after a subdivision" path)
                    if (sum != t.count(c)) { ids.clear(); t.collect_samples(c, ids); CHECK((int)ids.size() == sum, "children's counts do not match their samples"); }
                } else {
                    CHECK(t.count(c) == (n.sample >= 0 ? 1 : 0), "leaf count %d with sample %d", t.count(c), n.sample);
                }
            }
            // the cluster index finds, for every sample, the cluster-level cell whose box holds it (or declines)
            for (int s : live) {
                const float* p = t.sample(s).pos;
                const int at = t.locate(p);
                if (at < 0) continue;
                const auto& n = t.cell(at);
                CHECK(n.alive && t.is_cluster_level(at), "index returned a dead or non-cluster cell");
                for (int a = 0; a < D; ++a) CHECK(p[a] > n.lo[a] && p[a] < n.hi[a], "index returned a cell that does not contain the point");
                CHECK(t.is_not_new_at(at, p), "a stored sample is not found from its cluster cell");
            }
        }
    }
}

int main() {
    check_active_set();
    check_tree<3>(11, 60000);
    check_tree<2>(12, 40000);
    std::printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
    return fails ? 1 : 0;
}
