// K3 (lookup half) + K4 — the SDF + gradient + variance query.
// Replaces GPisMap3::test_kernel / GPisMap::test_kernel (cpp/src/GPisMap3.cpp:794-902,
// cpp/src/GPisMap.cpp:665-763), QueryNonEmptyLevelC (cpp/src/octree.cpp:861-893,
// cpp/src/quadtree.cpp:643-671), OnGPIS::testSinglePoint / test2Dpoint
// (cpp/src/OnGPIS.cpp:177-239) and the test covariance (cpp/src/covFnc.cpp:258-314, 404-450).
//
// Pipeline per batch of queries:
//   k_candidates   one thread per query: probe the Morton-keyed leaf table around the query,
//                  keep the 4 nearest leaves in (sqdist, DFS rank) order, preset var_f, emit the
//                  pass-1 work item (query, nearest leaf)
//   k_eval_*       evaluate (query, leaf) pairs: k*, k*^T alpha, solve with L, variances
//   k_select2      queries whose nearest-leaf variance exceeds the threshold emit work items for
//                  the 2nd / 3rd nearest leaves (GPisMap3.cpp:837-854)
//   k_fuse         min-variance pick or weighted blend (GPisMap3.cpp:864-895), written in place
#pragma once
#include "common.cuh"

namespace gpis {

struct QueryWork {  // per-batch device scratch
    int4* cand;          // per query: nc, slot0, slot1, slot2
    int32_t* tie;        // per query tie flag (optional, may be null)
    float* evalout;      // per query 3 x 8 floats: [f, g0, g1, g2, vf, vg0, vg1, vg2] per candidate rank
    int2* pairs;         // work items: (query, rank<<28 | slot)   (rank 0..2)
    int32_t* counters;   // [0] = number of pairs
};

// DFS rank key of a cell under the current root (octree.cpp:844-851: rank = 4[z<c] + 2[y<c] + [x>c];
// quadtree.cpp:630-633: rank = 2[y<c] + [x>c]). Larger key = visited later.
__device__ inline uint64_t dfs_key(const QueryParams& P, int4 cell) {
    const int L = P.levels;
    const uint32_t mask = (L >= 31) ? 0x7fffffffu : ((1u << L) - 1u);
    const uint32_t lx = (uint32_t)(cell.x - P.root_min[0]) & mask;
    const uint32_t ly = (~(uint32_t)(cell.y - P.root_min[1])) & mask;
    const uint32_t lz = (P.dim == 3) ? ((~(uint32_t)(cell.z - P.root_min[2])) & mask) : 0u;
    if (P.dim == 3) return spread3(lx) | (spread3(ly) << 1) | (spread3(lz) << 2);
    // 2-D: interleave two coordinates (reuse spread3 on 21 bits: order is preserved)
    return spread3(lx) | (spread3(ly) << 1);
}

// (sq, dfs) lexicographic "a before b"
__device__ inline bool cand_before(const QueryParams& P, const LeafTable& T, float sa, int ia, float sb, int ib) {
    if (sa < sb) return true;
    if (sa > sb) return false;
    return dfs_key(P, T.cell[ia]) < dfs_key(P, T.cell[ib]);
}

__global__ void __launch_bounds__(256)
k_candidates(const float* __restrict__ x, int64_t nq, float* __restrict__ res, LeafTable T, QueryParams P, QueryWork W) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int dim = P.dim, w = 1 + dim;
    float xq[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < dim; ++c) xq[c] = x[q * dim + c];
    res[q * 2 * w + w] = P.var_preset;  // GPisMap3.cpp:816, GPisMap.cpp:682

    // query AABB in float, exactly as AABB3(x, half) builds it (octree.h:69-77)
    float qlo[3], qhi[3];
    int clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0};
    for (int c = 0; c < dim; ++c) {
        qlo[c] = xq[c] - P.search_half;
        qhi[c] = xq[c] + P.search_half;
        // conservative cell range; the exact float test below decides
        clo[c] = (int)floor((double)qlo[c] * P.inv_pitch - 1.0 - 1e-3);
        chi[c] = (int)floor((double)qhi[c] * P.inv_pitch + 1e-3);
    }
    int nc = 0;
    float bs[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
    int bi[4] = {-1, -1, -1, -1};
    for (int iz = clo[2]; iz <= chi[2]; ++iz)
        for (int iy = clo[1]; iy <= chi[1]; ++iy)
            for (int ix = clo[0]; ix <= chi[0]; ++ix) {
                const int slot = table_find(T, cell_key(ix, iy, iz));
                if (slot < 0) continue;
                const float4 ct = T.centre[slot], bl = T.lo[slot], bh = T.hi[slot];
                const float cc[3] = {ct.x, ct.y, ct.z};
                const float blo[3] = {bl.x, bl.y, bl.z}, bhi[3] = {bh.x, bh.y, bh.z};
                bool hit = true;
                float sq = 0.f;
                for (int c = 0; c < dim; ++c) {
                    const float lo = blo[c], hi = bhi[c];                                   // octree.h:73-78 (c -/+ l), all levels
                    if (qhi[c] < lo || qlo[c] > hi) hit = false;                            // octree.h:128-135
                    const float d = cc[c] - xq[c];                                          // octree.cpp:24-31
                    sq = (c == 0) ? d * d : sq + d * d;
                }
                if (!hit) continue;
                ++nc;
                // insert into the sorted top-4
                int pos = 4;
                for (int k = 3; k >= 0; --k)
                    if (bi[k] < 0 || cand_before(P, T, sq, slot, bs[k], bi[k])) pos = k;
                if (pos < 4) {
                    for (int k = 3; k > pos; --k) { bs[k] = bs[k - 1]; bi[k] = bi[k - 1]; }
                    bs[pos] = sq; bi[pos] = slot;
                }
            }
    const int numc = min(nc, 3);
    W.cand[q] = make_int4(nc, numc > 0 ? bi[0] : -1, numc > 1 ? bi[1] : -1, numc > 2 ? bi[2] : -1);
    int t = 0;
    for (int k = 0; k < numc && k + 1 < nc; ++k)
        if (bs[k] == bs[k + 1]) t = 1;
    W.tie[q] = t;
    // More than 16 candidates with an exact distance tie among the picks: the reference's std::sort
    // (introsort) does not keep DFS order there; k_candidates_exact replays it.
    if (t && nc > 16) return;
    if (nc > 0 && T.rec[bi[0]] != 0ull) {
        const int p = atomicAdd(&W.counters[0], 1);
        W.pairs[p] = make_int2((int)q, bi[0]);
    }
}

// ------------------------------------------------------------------ exact std::sort replay
// libstdc++ (GCC 13, bits/stl_algo.h) std::sort on the candidate index array with comparator
// sqdst[i1] < sqdst[i2], started from the tree's DFS order — what GPisMap3.cpp:826-829 executes.
// Only queries with > 16 candidates AND an exact tie among the picks come here (a symmetric query
// grid, e.g. the reference's own demo grids, produces them).
#define GPIS_MAXC 256
struct CandList { float key[GPIS_MAXC]; int slot[GPIS_MAXC]; };

__device__ inline void ssr_swap(CandList& L, int a, int b) {
    const float tk = L.key[a]; L.key[a] = L.key[b]; L.key[b] = tk;
    const int ts = L.slot[a]; L.slot[a] = L.slot[b]; L.slot[b] = ts;
}
__device__ inline void ssr_unguarded_linear_insert(CandList& L, int last) {
    const float vk = L.key[last]; const int vs = L.slot[last];
    int next = last - 1;
    while (vk < L.key[next]) { L.key[last] = L.key[next]; L.slot[last] = L.slot[next]; last = next; --next; }
    L.key[last] = vk; L.slot[last] = vs;
}
__device__ inline void ssr_insertion_sort(CandList& L, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (L.key[i] < L.key[first]) {
            const float vk = L.key[i]; const int vs = L.slot[i];
            for (int j = i; j > first; --j) { L.key[j] = L.key[j - 1]; L.slot[j] = L.slot[j - 1]; }
            L.key[first] = vk; L.slot[first] = vs;
        } else ssr_unguarded_linear_insert(L, i);
    }
}
// heapsort fallback of introsort (depth limit reached): __partial_sort(first, last, last) = __make_heap + __sort_heap
// (bits/stl_heap.h: __adjust_heap, __push_heap, __pop_heap), replayed move by move
__device__ inline void ssr_adjust_heap(CandList& L, int first, int hole, int len, float vk, int vs) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (L.key[first + child] < L.key[first + child - 1]) child--;
        L.key[first + hole] = L.key[first + child]; L.slot[first + hole] = L.slot[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        L.key[first + hole] = L.key[first + child - 1]; L.slot[first + hole] = L.slot[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && L.key[first + parent] < vk) {
        L.key[first + hole] = L.key[first + parent]; L.slot[first + hole] = L.slot[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    L.key[first + hole] = vk; L.slot[first + hole] = vs;
}
__device__ inline void ssr_heapsort(CandList& L, int first, int last) {
    const int len = last - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent) {
            ssr_adjust_heap(L, first, parent, len, L.key[first + parent], L.slot[first + parent]);
            if (parent == 0) break;
        }
    for (int end = last; end - first > 1;) {
        --end;
        const float vk = L.key[end]; const int vs = L.slot[end];
        L.key[end] = L.key[first]; L.slot[end] = L.slot[first];
        ssr_adjust_heap(L, first, 0, end - first, vk, vs);
    }
}
__device__ inline void std_sort_replay(CandList& L, int n) {
    if (n <= 1) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    // __introsort_loop, recursion on the right part turned into an explicit stack
    int stk_first[32], stk_last[32], stk_depth[32], sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        // the reference recurses into [cut, last) FIRST and then continues with [first, cut):
        // the two sub-ranges are disjoint, so the order of processing does not change the result.
        while (last - first > 16) {
            if (depth == 0) { ssr_heapsort(L, first, last); break; }
            --depth;
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            int med;
            if (L.key[a] < L.key[b]) { if (L.key[b] < L.key[c]) med = b; else if (L.key[a] < L.key[c]) med = c; else med = a; }
            else if (L.key[a] < L.key[c]) med = a;
            else if (L.key[b] < L.key[c]) med = c;
            else med = b;
            ssr_swap(L, first, med);
            int lo = first + 1, hi = last;
            for (;;) {
                while (L.key[lo] < L.key[first]) ++lo;
                --hi;
                while (L.key[first] < L.key[hi]) --hi;
                if (!(lo < hi)) break;
                ssr_swap(L, lo, hi);
                ++lo;
            }
            if (sp < 32) { stk_first[sp] = lo; stk_last[sp] = last; stk_depth[sp] = depth; ++sp; }
            last = lo;
        }
    }
    if (n > 16) {
        ssr_insertion_sort(L, 0, 16);
        for (int i = 16; i != n; ++i) ssr_unguarded_linear_insert(L, i);
    } else ssr_insertion_sort(L, 0, n);
}

__global__ void __launch_bounds__(128)
k_candidates_exact(const float* __restrict__ x, int64_t nq, LeafTable T, QueryParams P, QueryWork W) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (W.tie[q] == 0 || W.cand[q].x <= 16) return;
    const int dim = P.dim;
    float xq[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < dim; ++c) xq[c] = x[q * dim + c];
    float qlo[3], qhi[3];
    int clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0};
    for (int c = 0; c < dim; ++c) {
        qlo[c] = xq[c] - P.search_half;
        qhi[c] = xq[c] + P.search_half;
        clo[c] = (int)floor((double)qlo[c] * P.inv_pitch - 1.0 - 1e-3);
        chi[c] = (int)floor((double)qhi[c] * P.inv_pitch + 1e-3);
    }
    CandList L;
    uint64_t dk[GPIS_MAXC];
    int nc = 0;
    for (int iz = clo[2]; iz <= chi[2]; ++iz)
        for (int iy = clo[1]; iy <= chi[1]; ++iy)
            for (int ix = clo[0]; ix <= chi[0]; ++ix) {
                const int slot = table_find(T, cell_key(ix, iy, iz));
                if (slot < 0) continue;
                const float4 ct = T.centre[slot], bl = T.lo[slot], bh = T.hi[slot];
                const float cc[3] = {ct.x, ct.y, ct.z};
                const float blo[3] = {bl.x, bl.y, bl.z}, bhi[3] = {bh.x, bh.y, bh.z};
                bool hit = true;
                float sq = 0.f;
                for (int c = 0; c < dim; ++c) {
                    if (qhi[c] < blo[c] || qlo[c] > bhi[c]) hit = false;
                    const float d = cc[c] - xq[c];
                    sq = (c == 0) ? d * d : sq + d * d;
                }
                if (!hit || nc >= GPIS_MAXC) continue;
                // keep the list in DFS order (insertion by DFS key)
                const uint64_t k = dfs_key(P, T.cell[slot]);
                int pos = nc;
                while (pos > 0 && dk[pos - 1] > k) { dk[pos] = dk[pos - 1]; L.key[pos] = L.key[pos - 1]; L.slot[pos] = L.slot[pos - 1]; --pos; }
                dk[pos] = k; L.key[pos] = sq; L.slot[pos] = slot;
                ++nc;
            }
    std_sort_replay(L, nc);
    const int numc = min(nc, 3);
    W.cand[q] = make_int4(nc, numc > 0 ? L.slot[0] : -1, numc > 1 ? L.slot[1] : -1, numc > 2 ? L.slot[2] : -1);
    if (nc > 0 && T.rec[L.slot[0]] != 0ull) {
        const int p = atomicAdd(&W.counters[0], 1);
        W.pairs[p] = make_int2((int)q, L.slot[0]);
    }
}

// After pass 1: GPisMap3.cpp:837-854. rank-1/2 candidates of queries that are still too uncertain.
__global__ void __launch_bounds__(256)
k_select2(int64_t nq, const float* __restrict__ res, LeafTable T, QueryParams P, QueryWork W, int q_base_unused) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int4 c = W.cand[q];
    if (c.x < 2) return;
    const int w = 1 + P.dim;
    // variance of the nearest leaf, or the preset if that leaf is untrained
    float vf = P.var_preset;
    if (T.rec[c.y] != 0ull) vf = W.evalout[(q * 3 + 0) * 8 + w];
    if (!(vf > P.var_thre)) return;
    const int numc = min(c.x, 3);
    const int ids[3] = {c.y, c.z, c.w};
    for (int k = 1; k < numc; ++k) {
        if (T.rec[ids[k]] == 0ull) continue;
        const int p = atomicAdd(&W.counters[0], 1);
        W.pairs[p] = make_int2((int)q, (k << 28) | ids[k]);
    }
}

// ------------------------------------------------------------------ evaluation, version 1
// One CTA per (query, leaf) pair; warp c handles right-hand side c of k* (f, d/dx, d/dy, d/dz).
// Straightforward blocked forward substitution that streams the leaf's tiles from L2.
#define EVAL1_THREADS 128
__global__ void __launch_bounds__(EVAL1_THREADS)
k_eval_v1(const float* __restrict__ x, LeafTable T, QueryParams P, QueryWork W, const int2* __restrict__ pairs) {
    extern __shared__ __align__(16) float sm[];
    const int2 pr = pairs[blockIdx.x];
    const int q = pr.x, rank = (pr.y >> 28) & 3, slot = pr.y & 0x0fffffff;
    const int dim = P.dim, w = 1 + dim;
    const unsigned char* rec = reinterpret_cast<const unsigned char*>(T.rec[slot]);
    __builtin_assume(__isGlobal(rec));   // the record address travels as an integer: without the hint its loads are generic LD.E
    const int4 meta = T.meta[slot];
    const int N = meta.x, ng = meta.y, n = meta.z, nb = meta.w;
    const int npad = nb * 32;
    const float4* pts = reinterpret_cast<const float4*>(rec + rec_off_pts());
    const float* alpha = reinterpret_cast<const float*>(rec + rec_off_alpha(N));
    const float* dinv = reinterpret_cast<const float*>(rec + rec_off_dinv(N, nb));
    const float* tiles = reinterpret_cast<const float*>(rec + rec_off_tiles(N, nb));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float xq[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < dim; ++c) xq[c] = x[(int64_t)q * dim + c];
    float* Ks = sm;  // w columns of npad
    for (int i = tid; i < w * npad; i += EVAL1_THREADS) Ks[i] = 0.f;
    __syncthreads();
    // k*: covFnc.cpp:282-311 (3D) / 425-446 (2D), one test point
    for (int k = tid; k < N; k += EVAL1_THREADS) {
        const float4 p = pts[k];
        const int g = __float_as_int(p.w);
        const float xs[3] = {p.x, p.y, p.z};
        float d[3], s2 = 0.f;
        for (int c = 0; c < dim; ++c) { d[c] = xs[c] - xq[c]; s2 = (c == 0) ? d[c] * d[c] : s2 + d[c] * d[c]; }
        const float r = sqrtf(s2);
        const DF e = exp_df(-P.a * r);
        Ks[k] = kf_val(r, P.a, e);
        float k1[3];
        for (int c = 0; c < dim; ++c) { k1[c] = kf1_val(d[c], P.a, e); Ks[(1 + c) * npad + k] = k1[c]; }
        if (g >= 0) {
            for (int c = 0; c < dim; ++c) {
                const int row = N + c * ng + g;
                Ks[row] = -k1[c];
                for (int e2 = 0; e2 < dim; ++e2) {
                    const int c0 = min(c, e2), e0 = max(c, e2);
                    Ks[(1 + e2) * npad + row] = kf2_val(r, d[c0], d[e0], c == e2 ? 1.f : 0.f, P.a, e);
                }
            }
        }
    }
    __syncthreads();
    if (warp >= w) return;
    float* kc = Ks + warp * npad;
    // mean: k*^T alpha (OnGPIS.cpp:187)
    double mu = 0.0;   // double accumulation: see k_eval_v3
    for (int i = lane; i < n; i += 32) mu = fma((double)kc[i], (double)alpha[i], mu);
    for (int o = 16; o > 0; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
    // block elimination (OnGPIS.cpp:199): u_i = b_i - sum_j G_ij u_j, then v_i = inv(Lii) u_i
    float ss = 0.f;
    for (int bi = 0; bi < nb; ++bi) {
        float t = kc[bi * 32 + lane];
        for (int bj = 0; bj < bi; ++bj) {
            const float* Tl = tiles + (size_t)tile_index(bi, bj, nb) * GPIS_TILE_ELEMS;
            const float* u = kc + bj * 32;
            float s = 0.f;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) s = fmaf(__ldg(Tl + k * 32 + lane), u[k], s);
            t -= s;
        }
        __syncwarp();
        kc[bi * 32 + lane] = t;
        __syncwarp();
        const float* D = dinv + (size_t)bi * GPIS_TILE_ELEMS;
        float vr = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) vr = fmaf(__ldg(D + k * 32 + lane), __shfl_sync(0xffffffffu, t, k), vr);
        ss = fmaf(vr, vr, ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
        float* out = W.evalout + ((int64_t)q * 3 + rank) * 8;
        out[warp] = (float)mu;
        // priors: OnGPIS.cpp:203-212 / 235-237, evaluated in double like the reference
        const double prior = (warp == 0) ? (double)P.prior_f : P.prior_g;
        out[w + warp] = (float)(prior - (double)ss);
    }
}

// Fusion (GPisMap3.cpp:818-897, GPisMap.cpp:684-758), in place on res.
__global__ void __launch_bounds__(256)
k_fuse(int64_t nq, float* __restrict__ res, LeafTable T, QueryParams P, QueryWork W) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int4 c = W.cand[q];
    if (c.x == 0) return;
    const int w = 1 + P.dim, w2 = 2 * w;
    float* r = res + q * w2;
    const float* e0 = W.evalout + (q * 3) * 8;
    const bool t0 = T.rec[c.y] != 0ull;
    if (t0)
        for (int k = 0; k < w2; ++k) r[k] = e0[k];
    if (c.x == 1) return;
    if (!(r[w] > P.var_thre)) return;
    const int numc = min(c.x, 3);
    const int ids[3] = {c.y, c.z, c.w};
    float cand[3][8];
    for (int k = 0; k < w2; ++k) cand[0][k] = r[k];
    for (int m = 1; m < numc; ++m) {
        if (T.rec[ids[m]] != 0ull) {
            for (int k = 0; k < w2; ++k) cand[m][k] = W.evalout[(q * 3 + m) * 8 + k];
        } else {  // null GP at rank 1/2 is undefined behaviour in the reference (GPisMap3.cpp:852-853)
            for (int k = 0; k < w2; ++k) cand[m][k] = r[k];
            cand[m][w] = 1e30f;
        }
    }
    int ord[3] = {0, 1, 2};
    for (int i = 1; i < numc; ++i) {  // stable insertion sort by var_f, like std::sort on <= 3 items
        const int oi = ord[i];
        int j = i - 1;
        while (j >= 0 && cand[ord[j]][w] > cand[oi][w]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = oi;
    }
    const float* A = cand[ord[0]];
    if (A[w] < P.var_thre) {
        for (int k = 0; k < w2; ++k) r[k] = A[k];
    } else {
        const float* B = cand[ord[1]];
        const float w1 = A[w] - P.var_thre, w2f = B[w] - P.var_thre, w12 = w1 + w2f;
        for (int k = 0; k < w2; ++k) r[k] = (w2f * A[k] + w1 * B[k]) / w12;
    }
}

}  // namespace gpis
