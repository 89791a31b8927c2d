// K4 bucketing: (query, leaf) pairs are grouped by leaf (histogram, scan, scatter on the device) and cut into
// work items of 8 / 6 / 4 queries by leaf size class, so that one CTA of k_eval_v3 streams a leaf's factor once
// for several queries. Also the per-pass algorithmic work counters that gpis_stats reports.
#pragma once
#include <string>

#include "common.cuh"
#include "query.cuh"

namespace gpis {

#define QB 8

struct SortBufs {
    int32_t* count;    // nslots
    int32_t* start;    // nslots + 1   (exclusive prefix of count)
    int32_t* istart;   // nslots + 1   (exclusive prefix of ceil(count/QB))
    int32_t* cursor;   // nslots
    int2* sorted;      // npairs
    int4* items;       // work items of the 8-query class: slot, first sorted pair, count, unused
    int4* itemsB;      // work items of the 4-query class (leaves too large for 32 right-hand sides in smem)
    int4* itemsM;      // work items of the 6-query class
    int2* pairsC;      // pairs of leaves too large for either (one CTA per pair, k_eval_v1)
    int32_t* totals;   // [0] = #items (legacy list), [1] = #itemsA, [2] = #itemsB, [3] = #pairsC, [4] = #itemsM
};

__global__ void k_pair_hist(const int2* __restrict__ pairs, int npairs, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) atomicAdd(&count[pairs[i].y & 0x0fffffff], 1);
}

// single-block exclusive scans over the slots
__global__ void __launch_bounds__(1024) k_slot_scan(SortBufs S, int nslots) {
    __shared__ int s_cnt[1024], s_itm[1024];
    const int t = threadIdx.x;
    const int per = (nslots + 1023) / 1024;
    const int lo = t * per, hi = min(nslots, lo + per);
    int c = 0, it = 0;
    for (int i = lo; i < hi; ++i) { c += S.count[i]; it += (S.count[i] + QB - 1) / QB; }
    s_cnt[t] = c; s_itm[t] = it;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int a = 0, b = 0;
        if (t >= off) { a = s_cnt[t - off]; b = s_itm[t - off]; }
        __syncthreads();
        s_cnt[t] += a; s_itm[t] += b;
        __syncthreads();
    }
    int pc = s_cnt[t] - c, pi = s_itm[t] - it;
    for (int i = lo; i < hi; ++i) {
        S.start[i] = pc; S.istart[i] = pi; S.cursor[i] = pc;
        pc += S.count[i]; pi += (S.count[i] + QB - 1) / QB;
    }
    if (t == 1023) { S.start[nslots] = s_cnt[1023]; S.istart[nslots] = s_itm[1023]; S.totals[0] = s_itm[1023]; }
}

__global__ void k_pair_scatter(const int2* __restrict__ pairs, int npairs, SortBufs S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int2 p = pairs[i];
    const int pos = atomicAdd(&S.cursor[p.y & 0x0fffffff], 1);
    S.sorted[pos] = p;
}

// Work lists by leaf size class (nb = 32-row blocks of the leaf system). A leaf's items are contiguous so that
// CTAs running at the same time share its tiles in L2.
__global__ void k_make_items_classed(SortBufs S, LeafTable T, int nslots, int nbA, int nbM, int nbB) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int c = S.count[s];
    if (c == 0) return;
    const int nb = T.meta[s].w;
    if (nb <= nbA) {
        int it = atomicAdd(&S.totals[1], (c + 7) / 8);
        for (int o = 0; o < c; o += 8) S.items[it++] = make_int4(s, S.start[s] + o, min(8, c - o), 0);
    } else if (nb <= nbM) {
        int it = atomicAdd(&S.totals[4], (c + 5) / 6);
        for (int o = 0; o < c; o += 6) S.itemsM[it++] = make_int4(s, S.start[s] + o, min(6, c - o), 0);
    } else if (nb <= nbB) {
        int it = atomicAdd(&S.totals[2], (c + 3) / 4);
        for (int o = 0; o < c; o += 4) S.itemsB[it++] = make_int4(s, S.start[s] + o, min(4, c - o), 0);
    } else {
        int it = atomicAdd(&S.totals[3], c);
        for (int o = 0; o < c; ++o) S.pairsC[it++] = S.sorted[S.start[s] + o];
    }
}

// Algorithmic work of one evaluation pass, accumulated over leaves with work:
//   acc[0] flops            sum over evaluations of 4n^2 + 16n + 80N
//   acc[1] gather bytes     sum over evaluations of 16N + 4n + 2n(n+1)
//   acc[2] compulsory bytes sum over DISTINCT leaves touched of the same record size
__global__ void k_query_stats(SortBufs S, LeafTable T, int nslots, double* acc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double f = 0, g = 0, c = 0;
    if (s < nslots && S.count[s] > 0) {
        const int4 m = T.meta[s];
        const double N = m.x, n = m.z, cnt = S.count[s];
        const double rec = 16.0 * N + 4.0 * n + 2.0 * n * (n + 1.0);
        f = cnt * (4.0 * n * n + 16.0 * n + 80.0 * N);
        g = cnt * rec;
        c = rec;
    }
    for (int o = 16; o > 0; o >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0 && (f != 0 || c != 0)) { atomicAdd(acc, f); atomicAdd(acc + 1, g); atomicAdd(acc + 2, c); }
}

}  // namespace gpis
