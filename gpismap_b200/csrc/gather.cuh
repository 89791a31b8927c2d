// SURVEY.md 8 f-2 — device-side dirty set and training-set gather.
// Replaces, per frame, the host loops of updateGPs (cpp/src/GPisMap3.cpp:720-741, cpp/src/GPisMap.cpp:596-616: dirty set
// = active leaves + every non-empty leaf whose box touches AABB(c, Rtimes*l)) and updateGPs_kernel's QueryRange per
// dirty leaf (cpp/src/GPisMap3.cpp:698-709, cpp/src/octree.cpp:777-804: samples with |p - c|^2 < r^2, strict, in the
// tree's DFS order — that order is the row order of K).
//
// The samples live on the device, one block per leaf (rows of 2*dim+3 floats in the leaf's own DFS order; the host
// re-sends a leaf's block only when its samples changed: gpis_samples_set). A training ball is then the ordered
// concatenation, over the leaves whose box touches the ball's AABB taken in DFS order (dfs_key, the order the tree
// walk visits them), of the samples that pass the distance test in the reference's float arithmetic. The AABB
// pruning of the reference's descent never removes a sample that passes that test (a sample strictly inside a node's
// float box that lies outside the query box in one axis is at least one ulp beyond c +- r there), so the flat lookup
// returns the tree's set, in the tree's order.
//
//   k_dirty_mark   one thread per active leaf: flag every registered leaf whose effective box touches AABB(c, r)
//   k_ball_count   one warp per dirty leaf: neighbour leaves (sorted by DFS key), N and ng of the ball
//   k_ball_gather  one warp per dirty leaf: ordered compaction of the ball's rows into the CSR that K1 reads
// HBM/L2-latency bound integer and gather work, a few hundred microseconds per frame.
#pragma once
#include "common.cuh"
#include "query.cuh"

namespace gpis {

#define GATHER_MAXNBR 128   // leaves whose box can touch a training ball's AABB (5^3 = 125 for Rtimes <= 3)

struct SampleStore {      // per table slot
    uint64_t* ptr;        // device address of the leaf's sample rows, 0 = none
    int32_t* cnt;         // rows
};
struct StoreUpdate { int32_t slot, cnt; uint64_t ptr; };

__global__ void k_store_apply(SampleStore S, const StoreUpdate* __restrict__ ups, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S.ptr[ups[i].slot] = ups[i].ptr;
    S.cnt[ups[i].slot] = ups[i].cnt;
}

struct WordCopy { const uint32_t* src; uint32_t* dst; uint64_t words; };
__global__ void __launch_bounds__(256) k_copy_words(const WordCopy* __restrict__ jobs) {
    const WordCopy j = jobs[blockIdx.x];
    for (uint64_t i = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; i < j.words; i += (uint64_t)gridDim.y * blockDim.x) j.dst[i] = j.src[i];
}

// conservative lattice range of the float box [c - r, c + r]; the exact float test decides (as in k_candidates)
__device__ __forceinline__ void ball_cell_range(const QueryParams& P, const float* c, float r, float* qlo, float* qhi, int* clo, int* chi) {
    for (int a = 0; a < 3; ++a) { qlo[a] = 0.f; qhi[a] = 0.f; clo[a] = 0; chi[a] = 0; }
    for (int a = 0; a < P.dim; ++a) {
        qlo[a] = c[a] - r;                                   // AABB3(c, r): octree.h:73-78
        qhi[a] = c[a] + r;
        clo[a] = (int)floor((double)qlo[a] * P.inv_pitch - 1.0 - 1e-3);
        chi[a] = (int)floor((double)qhi[a] * P.inv_pitch + 1e-3);
    }
}
__device__ __forceinline__ bool box_touches(const LeafTable& T, int slot, int dim, const float* qlo, const float* qhi) {
    const float4 bl = T.lo[slot], bh = T.hi[slot];
    const float blo[3] = {bl.x, bl.y, bl.z}, bhi[3] = {bh.x, bh.y, bh.z};
    for (int a = 0; a < dim; ++a)
        if (qhi[a] < blo[a] || qlo[a] > bhi[a]) return false;   // octree.h:128-135 (inclusive), every level: effective box
    return true;
}

__global__ void __launch_bounds__(128)
k_dirty_mark(const int4* __restrict__ active, int n_active, LeafTable T, QueryParams P, float radius,
             int32_t* __restrict__ flag, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const int4 cell = active[i];
    const int self = table_find(T, cell_key(cell.x, cell.y, P.dim == 3 ? cell.z : 0));
    if (self < 0) return;
    const float4 ct = T.centre[self];
    const float c[3] = {ct.x, ct.y, ct.z};
    float qlo[3], qhi[3];
    int clo[3], chi[3];
    ball_cell_range(P, c, radius, qlo, qhi, clo, chi);
    if (atomicExch(&flag[self], 1) == 0) list[atomicAdd(count, 1)] = self;      // the active leaf itself (GPisMap3.cpp:729)
    for (int iz = clo[2]; iz <= chi[2]; ++iz)
        for (int iy = clo[1]; iy <= chi[1]; ++iy)
            for (int ix = clo[0]; ix <= chi[0]; ++ix) {
                const int slot = table_find(T, cell_key(ix, iy, iz));
                if (slot < 0 || !box_touches(T, slot, P.dim, qlo, qhi)) continue;
                if (atomicExch(&flag[slot], 1) == 0) list[atomicAdd(count, 1)] = slot;
            }
}

// gradflag rule (OnGPIS.cpp:63-66, 122-125), as in K1
__device__ __forceinline__ bool sample_grad_valid(const float* s, int dim) {
    bool allsmall = true;
    for (int c = 0; c < dim; ++c) allsmall = allsmall && (fabs((double)s[dim + c]) < 1e-6);
    return !((double)s[2 * dim + 2] > 0.1001 || allsmall);
}

// Neighbour leaves of one ball, sorted by DFS key, into nbr[0..m) (shared memory of the calling warp). Returns m.
__device__ __forceinline__ int ball_neighbours(const LeafTable& T, const QueryParams& P, const float* c, float radius,
                                               int* nbr, unsigned long long* nkey, int lane) {
    float qlo[3], qhi[3];
    int clo[3], chi[3];
    ball_cell_range(P, c, radius, qlo, qhi, clo, chi);
    const int nx = chi[0] - clo[0] + 1, ny = chi[1] - clo[1] + 1, nz = chi[2] - clo[2] + 1;
    const int total = nx * ny * nz;
    int m = 0;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        int slot = -1;
        if (t < total) {
            const int ix = clo[0] + t % nx, iy = clo[1] + (t / nx) % ny, iz = clo[2] + t / (nx * ny);
            slot = table_find(T, cell_key(ix, iy, iz));
            if (slot >= 0 && !box_touches(T, slot, P.dim, qlo, qhi)) slot = -1;
        }
        const unsigned msk = __ballot_sync(0xffffffffu, slot >= 0);
        if (slot >= 0) {
            const int pos = m + __popc(msk & ((1u << lane) - 1u));
            if (pos < GATHER_MAXNBR) { nbr[pos] = slot; nkey[pos] = dfs_key(P, T.cell[slot]); }
        }
        m += __popc(msk);
    }
    m = min(m, GATHER_MAXNBR);
    __syncwarp();
    // rank sort (m <= 128): every lane places its elements
    int myslot[4]; int myrank[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = lane + 32 * q;
        myslot[q] = -1; myrank[q] = 0;
        if (i < m) {
            const unsigned long long k = nkey[i];
            int r = 0;
            for (int j = 0; j < m; ++j) r += (nkey[j] < k || (nkey[j] == k && j < i)) ? 1 : 0;
            myslot[q] = nbr[i]; myrank[q] = r;
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) if (myslot[q] >= 0) nbr[myrank[q]] = myslot[q];
    __syncwarp();
    return m;
}

// One warp per dirty leaf. mode 0: count N and ng. mode 1: write the ball's rows, in order, at csr + off[d]*w9.
template <int MODE>
__global__ void __launch_bounds__(128)
k_ball(const int32_t* __restrict__ list, int ndirty, LeafTable T, QueryParams P, SampleStore S, float radius,
       int32_t* __restrict__ outN, int32_t* __restrict__ outNg, const int32_t* __restrict__ off, float* __restrict__ csr) {
    __shared__ int s_nbr[4][GATHER_MAXNBR];
    __shared__ unsigned long long s_key[4][GATHER_MAXNBR];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x * 4 + warp;
    if (d >= ndirty) return;
    const int slot = list[d];
    const float4 ct = T.centre[slot];
    const float c[3] = {ct.x, ct.y, ct.z};
    const int dim = P.dim, w9 = 2 * dim + 3;
    const float r2 = radius * radius;                         // octree.cpp:783 / prtree query_range: half * half
    const int m = ball_neighbours(T, P, c, radius, s_nbr[warp], s_key[warp], lane);
    int N = 0, ng = 0;
    float* dst = (MODE == 1) ? csr + (size_t)off[d] * w9 : nullptr;
    for (int q = 0; q < m; ++q) {
        const int ns = s_nbr[warp][q];
        const float* rows = reinterpret_cast<const float*>(S.ptr[ns]);
        __builtin_assume(__isGlobal(rows));
        const int cnt = S.cnt[ns];
        for (int base = 0; base < cnt; base += 32) {
            const int k = base + lane;
            bool in = false, gv = false;
            const float* s = rows + (size_t)k * w9;
            if (k < cnt) {
                float sq = 0.f;
                for (int a = 0; a < dim; ++a) { const float dd = s[a] - c[a]; sq = (a == 0) ? dd * dd : sq + dd * dd; }   // octree.cpp:24-31
                in = sq < r2;
                if (MODE == 0 && in) gv = sample_grad_valid(s, dim);
            }
            const unsigned msk = __ballot_sync(0xffffffffu, in);
            if (MODE == 0) ng += __popc(__ballot_sync(0xffffffffu, gv));
            if (MODE == 1 && in) {
                float* o = dst + (size_t)(N + __popc(msk & ((1u << lane) - 1u))) * w9;
                for (int a = 0; a < w9; ++a) o[a] = s[a];
            }
            N += __popc(msk);
        }
    }
    if (MODE == 0 && lane == 0) { outN[d] = N; outNg[d] = ng; }
}

}  // namespace gpis
