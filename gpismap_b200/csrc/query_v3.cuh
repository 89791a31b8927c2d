// K4, grouped evaluation — the production kernel.
//
// (query, leaf) pairs are bucketed by leaf; one CTA solves up to QBT queries (4*QBT right-hand
// sides) against one leaf. The record stores G(i,j) = L(i,j) inv(L(j,j)) for off-diagonal 32x32
// tiles and Dinv(j) = inv(L(j,j)); with U = right-hand sides after block elimination,
//      U_i = B_i - sum_{j<i} G(i,j) U_j,   V_j = Dinv(j) U_j,   var = prior - sum V^2
// which is L^{-1} k* (OnGPIS.cpp:199-213) evaluated block-wise.
//
// Mapping (B200): 8 warps per CTA, one CTA per SM. U (npad x 4*QBT floats) stays in shared memory.
// Block rows are dealt to the warps in snake order and processed in waves of 32 block rows (the partial wave
// first): inside a wave a warp keeps its (up to) 4 block rows as a 16x8-per-lane register accumulator, so one
// k-step costs 24 shared-memory words per lane for 128 FMAs (packed FFMA2, operands software-pipelined:
// mma_tile.cuh; a 4x8 lane tile was measured smem-bound, profiles/r01_history.md). G tiles stream L2 -> shared
// memory as 1 KB quarter tiles through per-warp double-buffered cp.async copies issued one step ahead (running
// across column and wave boundaries: G is read-only); 1 KB TMA bulk copies are kept behind -DE3_USE_TMA (12 %
// slower here). Synchronisation is dataflow: one mbarrier per block row, armed when the row's tile of U is final
// and written to shared memory (exactly once); the owner of row j+1 applies column j to that row first and
// publishes it before touching its other rows (lookahead). There is no block-wide barrier inside the elimination.
// The order of visits is not computed on the device: the host generates, per (nb, warp), a program of 16-byte
// records (e3_build_warp) and a list of the rows whose variance product the warp computes (e3_assign_variance);
// tests/test_eval_programs.py checks coverage, publish protocol and deadlock freedom of those programs on the CPU.
#pragma once
#include <algorithm>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "mma_tile.cuh"
#include "query.cuh"
#include "query_group.cuh"

namespace gpis {

// Tile staging: cp.async by default; -DE3_USE_TMA selects 1 KB TMA bulk copies (measured 12 % slower, profiles/r01)
#if !defined(E3_USE_TMA) && !defined(E3_USE_CPASYNC)
#define E3_USE_CPASYNC 1
#endif
struct Eval3Smem {
    static constexpr int off_bar = 0;                                // per warp 2 mbarriers
    static constexpr int off_stage = 256;                            // per warp 2 stages x 4 slots x 1 KB
    static constexpr int off_red = off_stage + E3_WARPS * 2 * E3_STAGE_FLOATS * 4;  // E3_WARPS x 32 floats
    static constexpr int off_ready = off_red + E3_WARPS * 32 * 4;     // one mbarrier per block row (<= 128)
    static constexpr int off_U = off_ready + 1024;
    static int total(int nbmax, int ncol) { return off_U + nbmax * 32 * ncol * 4; }
};

// ------------------------------------------------------------------ elimination programs
// The order in which a warp visits (block row, block column) pairs depends only on (nb, warp). It is generated
// once on the host (same rules the kernel used to evaluate inline: waves of E3_WAVE block rows, the partial
// wave first, snake dealing, split visits for the lookahead) and read by the kernel as a table with uniform
// loads, no per-visit index arithmetic on the device.
//   header  : {nwaves, nvisits, nvar, 0}, {0,0,0,0}
//   wave c  : {R, 0, 0, 0}, {row of slot 0, 1, 2, 3}                 (slot s = the warp's (R-1-s)-th row, ascending)
//   visit v : {flags, tiles of slots 0|1, tiles of slots 2|3, nact}     (one int4; a tile is a 16-bit index into the
//             leaf's tile array, 0xFFFF = idle slot)
//             flags: bits 0-2 s_lo, 3-5 cnt, 6 part, 7 split, 8 valid, 16-23 wave, 24-31 column j
//   4 terminator visits (valid clear), then the rows whose variance product this warp computes, 4 per record
struct EvalProg {
    const int4* recs;       // all programs
    const int32_t* off;     // [E3_PROG_MAXNB + 1][E3_WARPS] record offset of (nb, warp)
};
#define E3_PROG_MAXNB 80

struct E3Visit { int j, s_lo, cnt, part, split, c, nact, t[4]; };
struct E3WarpProg { int nwaves; int R[8]; int rows[8][4]; std::vector<E3Visit> vis; std::vector<int> var_rows; };

static inline void e3_build_warp(int nb, int warp, E3WarpProg& wp) {
    const int W_ = E3_WARPS;
    const int nwaves = (nb + E3_WAVE - 1) / E3_WAVE;
    const int first_rows = (nb % E3_WAVE) ? (nb % E3_WAVE) : E3_WAVE;
    auto wstart = [&](int c) { return c == 0 ? 0 : first_rows + (c - 1) * E3_WAVE; };
    auto wend = [&](int c) { return c == 0 ? std::min(nb, first_rows) : std::min(nb, first_rows + c * E3_WAVE); };
    auto off_of = [&](int r) { return r * W_ + ((r & 1) ? (W_ - 1 - warp) : warp); };
    auto wave_R = [&](int c) {
        const int m = wend(c) - wstart(c);
        int R = 0;
        for (int r = 0; r < E3_R; ++r) R += (off_of(r) < m) ? 1 : 0;
        return R;
    };
    auto row_of = [&](int c, int r) { return wstart(c) + off_of(r); };
    wp.nwaves = nwaves;
    wp.vis.clear(); wp.var_rows.clear();
    for (int c = 0; c < nwaves; ++c) {
        const int R = wave_R(c);
        wp.R[c] = R;
        for (int sl = 0; sl < 4; ++sl) wp.rows[c][sl] = (sl < R) ? row_of(c, R - 1 - sl) : -1;
    }
    for (int c = 0; c < nwaves; ++c) {
        const int R = wave_R(c);
        if (R == 0) continue;
        const int last = row_of(c, R - 1);
        for (int j = 0; j < last; ++j) {
            int below = 0;
            for (int r = 0; r < R; ++r) below += (row_of(c, r) <= j) ? 1 : 0;
            const int nact = R - below;
            if (nact <= 0) break;
            const bool split = row_of(c, R - nact) == j + 1;
            for (int part = 0; part < 2; ++part) {
                E3Visit v;
                if (part == 0) { v.s_lo = split ? nact - 1 : 0; v.cnt = split ? 1 : nact; }
                else { if (!(split && nact > 1)) break; v.s_lo = 0; v.cnt = nact - 1; }
                v.j = j; v.part = part; v.split = split ? 1 : 0; v.c = c; v.nact = nact;
                for (int sl = 0; sl < 4; ++sl)
                    v.t[sl] = (sl >= v.s_lo && sl < v.s_lo + v.cnt) ? tile_index(row_of(c, R - 1 - sl), j, nb) : -1;
                wp.vis.push_back(v);
            }
        }
    }
}

// Which warp computes the variance product V_j = Dinv(j) U_j of which row: a static list schedule on a coarse
// timing model of the elimination (visit costs measured with scripts/timing_probe.py), so that the warps which
// finish their elimination early take the rows of the warps that carry the end of the dependency chain.
static inline void e3_assign_variance(int nb, E3WarpProg (&wp)[E3_WARPS]) {
    const double cost[5] = {0., 3800., 5100., 7000., 7700.};   // cycles per visit by active slots (incl. control)
    const double pub = 300., cvar = 3600.;
    std::vector<double> ready(nb, -1.);
    ready[0] = 0.;
    double t[E3_WARPS];
    size_t pc[E3_WARPS];
    for (int w = 0; w < E3_WARPS; ++w) { t[w] = 0.; pc[w] = 0; }
    bool progress = true;
    while (progress) {
        progress = false;
        for (int w = 0; w < E3_WARPS; ++w) {
            while (pc[w] < wp[w].vis.size()) {
                const E3Visit& v = wp[w].vis[pc[w]];
                if (v.part == 0) {
                    if (ready[v.j] < 0.) break;
                    t[w] = std::max(t[w], ready[v.j]);
                }
                t[w] += cost[v.cnt];
                if (v.part == 0 && v.split) ready[v.j + 1] = t[w] + pub;
                ++pc[w];
                progress = true;
            }
        }
    }
    double avail[E3_WARPS];
    for (int w = 0; w < E3_WARPS; ++w) avail[w] = t[w];
    std::vector<int> order(nb);
    for (int j = 0; j < nb; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ready[a] < ready[b]; });
    for (int j : order) {
        int best = 0;
        double bt = 1e300;
        for (int w = 0; w < E3_WARPS; ++w) {
            const double s = std::max(avail[w], ready[j]);
            if (s < bt) { bt = s; best = w; }
        }
        avail[best] = bt + cvar;
        wp[best].var_rows.push_back(j);
    }
}

static inline void e3_emit_program(const E3WarpProg& wp, std::vector<int4>& out) {
    out.push_back(make_int4(wp.nwaves, (int)wp.vis.size(), (int)wp.var_rows.size(), 0));
    out.push_back(make_int4(0, 0, 0, 0));
    for (int c = 0; c < wp.nwaves; ++c) {
        out.push_back(make_int4(wp.R[c], 0, 0, 0));
        out.push_back(make_int4(wp.rows[c][0], wp.rows[c][1], wp.rows[c][2], wp.rows[c][3]));
    }
    auto pack2 = [](int a, int b) { return (int)(((uint32_t)a & 0xFFFFu) | (((uint32_t)b & 0xFFFFu) << 16)); };   // -1 -> 0xFFFF
    for (const E3Visit& v : wp.vis) {
        const int flags = v.s_lo | (v.cnt << 3) | (v.part << 6) | (v.split << 7) | (1 << 8) | (v.c << 16) | (v.j << 24);
        out.push_back(make_int4(flags, pack2(v.t[0], v.t[1]), pack2(v.t[2], v.t[3]), v.nact));
    }
    for (int k = 0; k < 4; ++k)   // terminators (valid bit clear); also the landing zone of the look-ahead loads
        out.push_back(make_int4(0, -1, -1, 0));
    for (size_t i = 0; i < wp.var_rows.size(); i += 4) {
        int r[4] = {-1, -1, -1, -1};
        for (size_t k = 0; k < 4 && i + k < wp.var_rows.size(); ++k) r[k] = wp.var_rows[i + k];
        out.push_back(make_int4(r[0], r[1], r[2], r[3]));
    }
}
// all (nb, warp) programs -> device
static inline int e3_upload_programs(int4** d_recs, int32_t** d_off, std::string& err) {
    std::vector<int4> recs;
    std::vector<int32_t> off((E3_PROG_MAXNB + 1) * E3_WARPS, 0);
    for (int nb = 1; nb <= E3_PROG_MAXNB; ++nb) {
        E3WarpProg wp[E3_WARPS];
        for (int w = 0; w < E3_WARPS; ++w) e3_build_warp(nb, w, wp[w]);
        e3_assign_variance(nb, wp);
        for (int w = 0; w < E3_WARPS; ++w) {
            off[nb * E3_WARPS + w] = (int32_t)recs.size();
            e3_emit_program(wp[w], recs);
        }
    }
    cudaError_t e = cudaMalloc(d_recs, recs.size() * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc(d_off, off.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(*d_recs, recs.data(), recs.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(*d_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("elimination programs: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

#ifdef E3_TIMING
__device__ long long g_e3_timing[64 * 32];   // [warp][phase] accumulated clock64 ticks of blockIdx.x == E3_TIMING
#define E3_T(var) const long long var = clock64();
#define E3_ACC(slot, t0, t1) if (blockIdx.x == E3_TIMING && lane == 0) g_e3_timing[warp * 32 + (slot)] += (t1) - (t0);
#else
#define E3_T(var)
#define E3_ACC(slot, t0, t1)
#endif

template <int QBT>
__global__ void __launch_bounds__(E3_THREADS, 1)
k_eval_v3(const float* __restrict__ x, LeafTable T, QueryParams P, QueryWork W, SortBufs S, const int4* __restrict__ items,
          EvalProg G) {
    constexpr int NCOL = 4 * QBT;      // right-hand sides per CTA
    constexpr int CPL = NCOL / 4;      // columns per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int4 item = items[blockIdx.x];
    const int slot = item.x, first = item.y, cnt = item.z;
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
    const int rg = lane >> 2, cg = lane & 3;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Eval3Smem::off_bar) + warp * 2;
    float* stg = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_stage) + warp * 2 * E3_STAGE_FLOATS;   // [2][E3_R][256]
    const uint32_t stg_s = smem_u32(stg) + lane * 16;   // this lane's 16-byte column of the warp's staging area (shared-space address)
    float* red = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_red);
    float* U = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_U);                       // [npad][NCOL]
    uint64_t* ready = reinterpret_cast<uint64_t*>(smem_raw + Eval3Smem::off_ready);  // mbarrier per block row: U_j final

    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    E3_T(t_begin)

    // ---- right-hand sides: k* of every query of the item (covFnc.cpp:282-311 / 425-446)
    {
        float4* U4 = reinterpret_cast<float4*>(U);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < npad * NCOL / 4; i += E3_THREADS) U4[i] = z4;
        for (int i = tid; i < nb; i += E3_THREADS) mbar_init(&ready[i], 1);
        fence_mbar_init();
    }
    // The tile staging area is idle until the elimination starts: park the leaf's points, alpha and the
    // query coordinates there (coalesced loads once, instead of a dependent global load per work item).
    float* stage_all = reinterpret_cast<float*>(smem_raw + Eval3Smem::off_stage);
    constexpr int kStageFloats = E3_WARPS * 2 * E3_STAGE_FLOATS;
    const bool park = (N * 4 + npad + 4 * QBT) <= kStageFloats;
    float4* s_pts = reinterpret_cast<float4*>(stage_all);
    float* s_alpha = stage_all + N * 4;
    float* s_xq = s_alpha + npad;
    if (park) {
        for (int i = tid; i < N; i += E3_THREADS) s_pts[i] = pts[i];
        for (int i = tid; i < npad; i += E3_THREADS) s_alpha[i] = alpha[i];
    }
    if (tid < cnt * 4) {
        const int qi = tid >> 2, c = tid & 3;
        const int q = S.sorted[first + qi].x;
        (park ? s_xq : red)[tid] = (c < dim) ? x[(int64_t)q * dim + c] : 0.f;
    }
    __syncthreads();
    const float* xq_s = park ? s_xq : red;
    // dimension-specialised (fully unrolled, no local arrays): this loop is 8-14 % of the kernel for small leaves
    auto kstar = [&](auto dimc) {
        constexpr int DIM = decltype(dimc)::value;
        for (int idx = tid; idx < N * QBT; idx += E3_THREADS) {
            const int qi = idx % QBT, k = idx / QBT;
            if (qi >= cnt) continue;
            const float4 p = park ? s_pts[k] : pts[k];
            const int g = __float_as_int(p.w);
            const float xs[3] = {p.x, p.y, p.z};
            float d[DIM], s2 = 0.f;
#pragma unroll
            for (int c = 0; c < DIM; ++c) { d[c] = xs[c] - xq_s[4 * qi + c]; s2 = (c == 0) ? d[c] * d[c] : s2 + d[c] * d[c]; }
            const float r = sqrtf(s2);
            const DF e = exp_df(-P.a * r);
            float* col = U + 4 * qi;
            float k1[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < DIM; ++c) k1[c] = kf1_val(d[c], P.a, e);
            // one 16-byte store per row: [k, dk/dx, dk/dy, dk/dz] (2-D: the fourth entry stays 0)
            *reinterpret_cast<float4*>(col + k * NCOL) = make_float4(kf_val(r, P.a, e), k1[0], k1[1], k1[2]);
            if (g >= 0) {
                float k2[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // symmetric (covFnc.cpp:300-309)
#pragma unroll
                for (int c = 0; c < DIM; ++c)
#pragma unroll
                    for (int e2 = c; e2 < DIM; ++e2) k2[c][e2] = k2[e2][c] = kf2_val(r, d[c], d[e2], c == e2 ? 1.f : 0.f, P.a, e);
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                    *reinterpret_cast<float4*>(col + (size_t)(N + c * ng + g) * NCOL) = make_float4(-k1[c], k2[c][0], k2[c][1], k2[c][2]);
            }
        }
    };
    if (dim == 3) kstar(std::integral_constant<int, 3>{});
    else kstar(std::integral_constant<int, 2>{});
    __syncthreads();

    // ---- mean: U^T alpha (OnGPIS.cpp:187); lane = column, warps split the rows. Accumulated in double: the sum
    // cancels heavily (sum |k_i alpha_i| >> |f|) and 4n DFMAs per query are free next to the 2n^2 FMAs of the solve;
    // the result is the correctly rounded fp32 mean of the fp32 inputs, so the distance to the reference is the
    // reference's own summation error.
    {
        // partial sums of the warps: doubles, parked behind the query coordinates in the (still idle) staging area
        double* redd = reinterpret_cast<double*>(stage_all + ((N * 4 + npad + 4 * QBT + 1) & ~1));
        double mu = 0.0;
        if (lane < NCOL)
            for (int i = warp; i < n; i += E3_WARPS) mu = fma((double)U[i * NCOL + lane], (double)(park ? s_alpha[i] : __ldg(alpha + i)), mu);
        redd[warp * 32 + lane] = mu;
        __syncthreads();
        if (warp == 0 && lane < NCOL) {
            double s = 0.0;
            for (int ww = 0; ww < E3_WARPS; ++ww) s += redd[ww * 32 + lane];
            const int qi = lane >> 2, c = lane & 3;
            if (qi < cnt && c < w) {
                const int2 pr = S.sorted[first + qi];
                W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + c] = (float)s;
            }
        }
        __syncthreads();   // `red` is reused by the final reduction: a warp without rows (tiny leaves) gets there at once
    }

    E3_T(t_elim0)
    E3_ACC(0, t_begin, t_elim0)   // k* build + mean
    // ---- block elimination in waves of E3_WAVE block rows, dataflow-synchronised
    // Warp w owns block rows base + E3_WARPS*r of each wave; accumulator slot s holds the (s+1)-th of them counted
    // from the END of the wave, so the rows still active at column j are always slots [0, nact). A (sub)visit is
    // one column j applied to a slot range: normally [0, nact); when this warp owns row j+1 the visit is split —
    // first the slot of row j+1 alone, then U_{j+1} is published (ready[j+1] = 1), then the remaining slots —
    // so the next column's operand is available long before the other warps ask for it (lookahead). Consumers
    // spin on ready[j]; there is no block-wide barrier inside the elimination.
    // a visit record as loaded (3 registers); the packed fields are decoded where they are used
    struct SV {
        int f; uint32_t t01, t23;
        __device__ __forceinline__ int j() const { return (int)((uint32_t)f >> 24); }
        __device__ __forceinline__ int s_lo() const { return f & 7; }
        __device__ __forceinline__ int cnt() const { return (f >> 3) & 7; }
        __device__ __forceinline__ int part() const { return (f >> 6) & 1; }
        __device__ __forceinline__ bool valid() const { return (f >> 8) & 1; }
        __device__ __forceinline__ int wave() const { return (f >> 16) & 255; }
        __device__ __forceinline__ bool solo() const { return (f & 0xC0) == 0x80; }   // split && part == 0: the lookahead part
        __device__ __forceinline__ int tile(int sl) const {                            // -1 = idle slot
            const uint32_t t = (sl == 0) ? (t01 & 0xFFFFu) : (sl == 1) ? (t01 >> 16) : (sl == 2) ? (t23 & 0xFFFFu) : (t23 >> 16);
            return t == 0xFFFFu ? -1 : (int)t;
        }
        __device__ __forceinline__ int solo_tile() const { return tile(s_lo()); }
    };
    const int4* prog = G.recs + G.off[nb * E3_WARPS + warp];
    const int nwaves = prog[0].x;
    const int4* pvis = prog + 2 + 2 * nwaves;
    auto load_sv = [&](int v) {
        const int4 a = __ldg(pvis + v);
        SV r;
        r.f = a.x; r.t01 = (uint32_t)a.y; r.t23 = (uint32_t)a.z;
        return r;
    };
#ifdef E3_USE_CPASYNC
    // cp.async (LDGSTS): every lane moves 2 x 16 B per quarter tile. Measured faster than 1 KB TMA bulk copies for
    // this access pattern (profiles/r01). The lookahead part of a split visit (one row, R = 1 work) fetches its
    // whole 4 KB tile at once: its four quarter steps are too short to hide a load each.
    auto issue = [&](const SV& v, int q, int stage) {
        if (v.solo()) {
            const float* src = tiles + (size_t)v.solo_tile() * GPIS_TILE_ELEMS + lane * 4;
            const uint32_t dst = stg_s + stage * (E3_STAGE_FLOATS * 4);
#pragma unroll
            for (int h = 0; h < 8; ++h)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + h * 512), "l"(src + h * 128) : "memory");
        } else {
#pragma unroll
            for (int sl = 0; sl < E3_R; ++sl) {
                const int ti = v.tile(sl);
                if (ti >= 0) {
                    const float* src = tiles + (size_t)ti * GPIS_TILE_ELEMS + q * 256 + lane * 4;
                    const uint32_t dst = stg_s + (stage * E3_R + sl) * 1024;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 128) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#else
    auto issue = [&](const SV& v, int q, int stage) {
        if (lane != 0) return;
        if (v.solo()) {
            mbar_expect_tx(&bars[stage], GPIS_TILE_BYTES);
            tma_load_1d(stg + stage * E3_STAGE_FLOATS, tiles + (size_t)v.solo_tile() * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[stage]);
            return;
        }
        mbar_expect_tx(&bars[stage], (uint32_t)v.cnt() * 1024u);
#pragma unroll
        for (int sl = 0; sl < E3_R; ++sl) {
            const int ti = v.tile(sl);
            if (ti >= 0) tma_load_1d(stg + (stage * E3_R + sl) * 256, tiles + (size_t)ti * GPIS_TILE_ELEMS + q * 256, 1024u, &bars[stage]);
        }
    };
#endif

    float acc[E3_R][4][CPL];
    uint32_t ph = 0;
    int st = 0;
    int vi = 0;
    SV cur = load_sv(0);
    SV nxt = load_sv(1);
    if (cur.valid()) issue(cur, 0, 0);
    if (tid == 0) mbar_arrive(&ready[0]);   // row 0 has nothing to eliminate: B_0 is U_0

    for (int c = 0; c < nwaves; ++c) {
        const int R = __ldg(prog + 2 + 2 * c).x;
        if (R == 0) continue;
        // accumulators <- right-hand sides of this warp's rows (slot s = its (R-1-s)-th row, ascending)
        {
            const int4 rows = __ldg(prog + 3 + 2 * c);
#pragma unroll
            for (int r = 0; r < E3_R; ++r) {
                if (r < R) {
                    const int row = r == 0 ? rows.x : r == 1 ? rows.y : r == 2 ? rows.z : rows.w;
                    const float* Ui = U + (size_t)row * 32 * NCOL;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        ld_cols<CPL>(Ui + (4 * rg + i) * NCOL, cg, acc[r][i]);
                }
            }
        }
        while (cur.valid() && cur.wave() == c) {
            const SV nxt2 = load_sv(vi + 2);   // two visits ahead: the records come from the L2 (no L1 left beside the
                                               // shared-memory carve-out) and must not be waited for
            // operand U_j must be final (published by the owner of row j)
            E3_T(t_r0)
            if (cur.part() == 0) mbar_wait(&ready[cur.j()], 0u);   // hardware-suspended wait, acquire semantics
            E3_T(t_r1)
            E3_ACC(1, t_r0, t_r1)   // waiting for the operand U_j
            const float* Uj = U + (size_t)cur.j() * 32 * NCOL;
            const bool solo = cur.solo();   // lookahead part: its whole tile was staged at once
            // One quarter step: prefetch (the next quarter of this visit, or the first step of the next one; a solo
            // visit owns its stage for all four quarters and prefetches only once), wait for this step's tiles, FMAs.
            // The slot count / slot index is a compile-time constant of the loop: the dispatch happens once per
            // visit, not once per step.
            auto step_pre = [&](int q) {
                E3_T(t_s0)
                const bool do_wait = !solo || q == 0;
                if (do_wait) {
                    bool issued = true;
                    if (!solo && q < 3) issue(cur, q + 1, st ^ 1);
                    else if (nxt.valid()) issue(nxt, 0, st ^ 1);
                    else issued = false;
#ifdef E3_USE_CPASYNC
                    if (issued) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
#else
                    (void)issued;
                    mbar_wait(&bars[st], (ph >> st) & 1u);
                    ph ^= (1u << st);
#endif
                }
                E3_T(t_s1)
                E3_ACC(2, t_s0, t_s1)   // issue + waiting for the staged tiles
            };
            auto step_post = [&](int q) {
                __syncwarp();
                if (!solo || q == 3) st ^= 1;
#ifdef E3_TIMING
                if (blockIdx.x == E3_TIMING && lane == 0) { g_e3_timing[warp * 32 + 8] += cur.cnt(); g_e3_timing[warp * 32 + 9] += 1; }
#endif
            };
            auto run_multi = [&](auto cc) {   // slots [0, C) against column j
                constexpr int C = decltype(cc)::value;
                for (int q = 0; q < 4; ++q) {
#ifdef E3_USE_CPASYNC
                    // the next quarter of this visit: slots [0, C) are known to be active, no predicates
                    if (q < 3) {
#pragma unroll
                        for (int sl = 0; sl < C; ++sl) {
                            const int ti = cur.tile(sl);
                            const float* src = tiles + (size_t)ti * GPIS_TILE_ELEMS + (q + 1) * 256 + lane * 4;
                            const uint32_t dst = stg_s + ((st ^ 1) * E3_R + sl) * 1024;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 128) : "memory");
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else if (nxt.valid()) {
                        issue(nxt, 0, st ^ 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    __syncwarp();
#else
                    step_pre(q);
#endif
                    qmma_sub<C, CPL, NCOL>(acc, stg + st * E3_STAGE_FLOATS, Uj + q * 8 * NCOL, rg, cg);
                    step_post(q);
                }
            };
            auto run_solo = [&](auto sc) {    // slot S alone; quarter q of its tile sits at stage + q*256
                constexpr int S = decltype(sc)::value;
#ifdef E3_USE_CPASYNC
                // the whole tile is already in flight: one prefetch (the next visit), one wait, four quarters
                if (nxt.valid()) { issue(nxt, 0, st ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                qmma_one<S, CPL, NCOL, 32>(acc, stg + st * E3_STAGE_FLOATS - S * 256, Uj, rg, cg);
                __syncwarp();
                st ^= 1;
#else
                for (int q = 0; q < 4; ++q) {
                    step_pre(q);
                    qmma_one<S, CPL, NCOL>(acc, stg + st * E3_STAGE_FLOATS + (q - S) * 256, Uj + q * 8 * NCOL, rg, cg);
                    step_post(q);
                }
#endif
            };
            if (solo) {
                switch (cur.s_lo()) {
#if E3_R >= 4
                    case 3: run_solo(std::integral_constant<int, 3>{}); break;
#endif
#if E3_R >= 3
                    case 2: run_solo(std::integral_constant<int, 2>{}); break;
#endif
#if E3_R >= 2
                    case 1: run_solo(std::integral_constant<int, 1>{}); break;
#endif
                    default: run_solo(std::integral_constant<int, 0>{}); break;
                }
            } else {
                switch (cur.cnt()) {
#if E3_R >= 4
                    case 4: run_multi(std::integral_constant<int, 4>{}); break;
#endif
#if E3_R >= 3
                    case 3: run_multi(std::integral_constant<int, 3>{}); break;
#endif
#if E3_R >= 2
                    case 2: run_multi(std::integral_constant<int, 2>{}); break;
#endif
                    default: run_multi(std::integral_constant<int, 1>{}); break;
                }
            }
            if (solo) {
                // row j+1 (slot s_lo) is final: publish it
                float* Un = U + (size_t)(cur.j() + 1) * 32 * NCOL;
#pragma unroll
                for (int r = 0; r < E3_R; ++r) {
                    if (r == cur.s_lo()) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            st_cols<CPL>(Un + (4 * rg + i) * NCOL, cg, acc[r][i]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[cur.j() + 1]);   // release: the stores above are visible to waiters
            }
            cur = nxt;
            nxt = nxt2;
            ++vi;
        }
    }
    E3_T(t_elim1)
    E3_T(t_elim2)
    E3_ACC(4, t_elim0, t_elim1)   // whole elimination of this warp

    // ---- V_j = Dinv(j) U_j and the column sums of V^2 (OnGPIS.cpp:200-201) for the rows the program assigns to
    // this warp; each row is waited for individually (ready[j]), so there is no block-wide barrier and a warp that
    // finishes its elimination early works here while the others are still eliminating.
    float ss[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) ss[c] = 0.f;
    {
        static_assert(E3_R == 4, "a stage must hold one full tile");
        // rows assigned to this warp by the program (not necessarily its own: see e3_assign_variance)
        const int nvar = prog[0].z;
        const int* vrows = reinterpret_cast<const int*>(pvis + (prog[0].y + 4));
        auto issue_d = [&](int j, int s2) {
            if (lane == 0) {
                mbar_expect_tx(&bars[s2], GPIS_TILE_BYTES);
                tma_load_1d(stg + s2 * E3_STAGE_FLOATS, dinv + (size_t)j * GPIS_TILE_ELEMS, GPIS_TILE_BYTES, &bars[s2]);
            }
        };
#ifdef E3_USE_CPASYNC
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#endif
        int vk = 0;
        int j = (vk < nvar) ? __ldg(vrows + vk) : -1;
        if (j >= 0) issue_d(j, st);
        while (j >= 0) {
            ++vk;
            const int jn = (vk < nvar) ? __ldg(vrows + vk) : -1;
            if (jn >= 0) issue_d(jn, st ^ 1);
            mbar_wait(&ready[j], 0u);   // U_j final (acquire); rows of other warps may still be in flight
            mbar_wait(&bars[st], (ph >> st) & 1u);
            ph ^= (1u << st);
            float v[E3_R][4][CPL];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < CPL; ++jj) v[0][i][jj] = 0.f;
            const float* Uj = U + (size_t)j * 32 * NCOL;
            qmma_one<0, CPL, NCOL, 32>(v, stg + st * E3_STAGE_FLOATS, Uj, rg, cg);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < CPL; ++jj) ss[jj] = fmaf(v[0][i][jj], v[0][i][jj], ss[jj]);   // (-V)^2 = V^2
            __syncwarp();
            st ^= 1;
            j = jn;
        }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        float t = ss[c];
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        ss[c] = t;
    }
    if (rg == 0) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) red[warp * 32 + lane_col<CPL>(cg, c)] = ss[c];
    }
    __syncthreads();
    if (warp == 0 && lane < NCOL) {
        float s = 0.f;
        for (int ww = 0; ww < E3_WARPS; ++ww) s += red[ww * 32 + lane];
        const int qi = lane >> 2, c = lane & 3;
        if (qi < cnt && c < w) {
            const int2 pr = S.sorted[first + qi];
            const double prior = (c == 0) ? (double)P.prior_f : P.prior_g;   // OnGPIS.cpp:203-212 / 235-237
            W.evalout[((int64_t)pr.x * 3 + ((pr.y >> 28) & 3)) * 8 + w + c] = (float)(prior - (double)s);
        }
    }
#ifdef E3_TIMING
    { const long long t_end = clock64(); if (blockIdx.x == E3_TIMING && lane == 0) { g_e3_timing[warp * 32 + 6] += t_end - t_elim2; g_e3_timing[warp * 32 + 7] = nb; } }
#endif
}

// ------------------------------------------------------------------ host side
#define E3_NB_A 40   // 8 queries per CTA: U = nb * 4 KB  (n <= 1280)
#define E3_NB_M 53   // 6 queries per CTA: U = nb * 3 KB  (n <= 1696)
#define E3_NB_B 80   // 4 queries per CTA: U = nb * 2 KB  (n <= 2560); larger leaves go to k_eval_v1

static inline int query_eval_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_eval_v3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval_v3<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval_v3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_eval_*): ") + cudaGetErrorString(e); return -2; }
    return 0;
}

// Buckets the npairs work items in W.pairs by leaf and evaluates them. *d_sort is a grow-only device
// buffer owned by the context. version: 1 = one CTA per pair (k_eval_v1), otherwise the production kernel.
static inline int query_eval(cudaStream_t st, const float* d_x, const LeafTable& T, const QueryParams& P,
                             const QueryWork& W, int npairs, int nslots, int max_nb, int32_t** d_sort,
                             int64_t* sort_cap, int64_t* launches, std::string& err, double* d_acc, int version,
                             const EvalProg& G, int64_t* items_out = nullptr) {
#define CK2(call)                                                                 \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return -2; } \
    } while (0)
    const int64_t max_items = npairs / 4 + nslots + 8;
    const int64_t need = (int64_t)nslots * 4 + 64 + (int64_t)npairs * 4 + max_items * 12 + 64;
    if (*sort_cap < need) {
        if (*d_sort) CK2(cudaFree(*d_sort));
        *d_sort = nullptr; *sort_cap = 0;
        CK2(cudaMalloc(d_sort, sizeof(int32_t) * (need + need / 4)));
        *sort_cap = need + need / 4;
    }
    SortBufs S;
    int32_t* p = *d_sort;
    S.totals = p; p += 16;
    S.count = p; p += nslots;
    S.start = p; p += nslots + 1;
    S.istart = p; p += nslots + 1;
    S.cursor = p; p += nslots;
    p += (4 - (((uintptr_t)p >> 2) & 3)) & 3;   // 16-byte align for int4
    S.items = reinterpret_cast<int4*>(p); p += max_items * 4;
    S.itemsB = reinterpret_cast<int4*>(p); p += max_items * 4;
    S.itemsM = reinterpret_cast<int4*>(p); p += max_items * 4;
    S.sorted = reinterpret_cast<int2*>(p); p += (int64_t)npairs * 2;
    S.pairsC = reinterpret_cast<int2*>(p);
    CK2(cudaMemsetAsync(S.totals, 0, sizeof(int32_t) * 16, st));
    CK2(cudaMemsetAsync(S.count, 0, sizeof(int32_t) * nslots, st));
    k_pair_hist<<<(npairs + 255) / 256, 256, 0, st>>>(W.pairs, npairs, S.count);
    k_slot_scan<<<1, 1024, 0, st>>>(S, nslots);
    k_pair_scatter<<<(npairs + 255) / 256, 256, 0, st>>>(W.pairs, npairs, S);
    *launches += 3;
    if (d_acc) { k_query_stats<<<(nslots + 255) / 256, 256, 0, st>>>(S, T, nslots, d_acc); *launches += 1; }
    const int v1_smem = (1 + P.dim) * max_nb * 32 * (int)sizeof(float);
    if (version == 1) {
        k_eval_v1<<<npairs, EVAL1_THREADS, v1_smem, st>>>(d_x, T, P, W, S.sorted);
        *launches += 1;
        CK2(cudaGetLastError());
        return 0;
    }
    k_make_items_classed<<<(nslots + 255) / 256, 256, 0, st>>>(S, T, nslots, E3_NB_A, E3_NB_M, E3_NB_B);
    *launches += 1;
    int32_t tot[5] = {0, 0, 0, 0, 0};
    CK2(cudaMemcpyAsync(tot, S.totals, sizeof(int32_t) * 5, cudaMemcpyDeviceToHost, st));
    CK2(cudaStreamSynchronize(st));
    if (items_out) { items_out[0] += tot[1]; items_out[1] += tot[4]; items_out[2] += tot[2]; items_out[3] += tot[3]; }
    if (tot[1] > 0) {
        const int nbm = max_nb < E3_NB_A ? max_nb : E3_NB_A;
        k_eval_v3<8><<<tot[1], E3_THREADS, Eval3Smem::total(nbm, 32), st>>>(d_x, T, P, W, S, S.items, G);
        *launches += 1;
    }
    if (tot[4] > 0) {
        const int nbm = max_nb < E3_NB_M ? max_nb : E3_NB_M;
        k_eval_v3<6><<<tot[4], E3_THREADS, Eval3Smem::total(nbm, 24), st>>>(d_x, T, P, W, S, S.itemsM, G);
        *launches += 1;
    }
    if (tot[2] > 0) {
        const int nbm = max_nb < E3_NB_B ? max_nb : E3_NB_B;
        k_eval_v3<4><<<tot[2], E3_THREADS, Eval3Smem::total(nbm, 16), st>>>(d_x, T, P, W, S, S.itemsB, G);
        *launches += 1;
    }
    if (tot[3] > 0) {
        k_eval_v1<<<tot[3], EVAL1_THREADS, v1_smem, st>>>(d_x, T, P, W, S.pairsC);
        *launches += 1;
    }
    CK2(cudaGetLastError());
#undef CK2
    return 0;
}

}  // namespace gpis
