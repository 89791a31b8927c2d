// libgpis_b200.so — C ABI (include/gpis_b200.h) over the hand-written sm_100a kernels.
// Host side of the boundary: context, device arena for leaf records, host mirror of the leaf
// table for memory management, parameter derivation with the reference's exact mixed
// float/double expressions. There is no CPU compute path in this library.
#include "../../include/gpis_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler is attached
#include <nccl.h>   // types only: the library is resolved at run time (gpis_comm_init), there is no link dependency

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "leaf_train.cuh"
#include "frame.cuh"
#include "gather.cuh"
#include "obs_gp.cuh"
#include "query.cuh"
#include "query_group.cuh"
#include "query_v3.cuh"

using namespace gpis;

// ------------------------------------------------------------------ K3: table maintenance kernels
namespace gpis {

struct SlotUpdate {
    uint64_t key;
    uint64_t rec;
    int32_t slot;
    int32_t live;     // 1 = upsert, 0 = erase
    int32_t cell[3];
    float centre[3];
    float lo[3], hi[3];
    int32_t meta[4];  // N, ng, n, nb
};

__global__ void k_table_apply(LeafTable T, const SlotUpdate* __restrict__ ups, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SlotUpdate u = ups[i];
    if (u.live) {
        // slot arrays first, then publish the key
        T.centre[u.slot] = make_float4(u.centre[0], u.centre[1], u.centre[2], 0.f);
        T.lo[u.slot] = make_float4(u.lo[0], u.lo[1], u.lo[2], 0.f);
        T.hi[u.slot] = make_float4(u.hi[0], u.hi[1], u.hi[2], 0.f);
        T.cell[u.slot] = make_int4(u.cell[0], u.cell[1], u.cell[2], 1);
        T.rec[u.slot] = u.rec;
        T.meta[u.slot] = make_int4(u.meta[0], u.meta[1], u.meta[2], u.meta[3]);
        uint32_t h = hash_key(u.key) & T.cap_mask;
        for (uint32_t probe = 0; probe <= T.cap_mask; ++probe) {
            const uint64_t prev = atomicCAS(reinterpret_cast<unsigned long long*>(T.keys + h),
                                            (unsigned long long)GPIS_KEY_EMPTY, (unsigned long long)u.key);
            if (prev == GPIS_KEY_EMPTY || prev == u.key) { T.vals[h] = u.slot; return; }
            h = (h + 1) & T.cap_mask;
        }
    } else {
        uint32_t h = hash_key(u.key) & T.cap_mask;
        for (uint32_t probe = 0; probe <= T.cap_mask; ++probe) {
            const uint64_t k = T.keys[h];
            if (k == u.key) { T.keys[h] = GPIS_KEY_TOMB; T.vals[h] = -1; break; }
            if (k == GPIS_KEY_EMPTY) break;
            h = (h + 1) & T.cap_mask;
        }
        T.rec[u.slot] = 0ull;
        T.cell[u.slot] = make_int4(0, 0, 0, 0);
    }
}

// Re-insert every live slot into a fresh key array (after growth or too many tombstones).
__global__ void k_table_rebuild(LeafTable T, int nslots) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int4 c = T.cell[s];
    if (c.w == 0) return;
    const uint64_t key = cell_key(c.x, c.y, c.z);
    uint32_t h = hash_key(key) & T.cap_mask;
    for (uint32_t probe = 0; probe <= T.cap_mask; ++probe) {
        const uint64_t prev = atomicCAS(reinterpret_cast<unsigned long long*>(T.keys + h),
                                        (unsigned long long)GPIS_KEY_EMPTY, (unsigned long long)key);
        if (prev == GPIS_KEY_EMPTY) { T.vals[h] = s; return; }
        h = (h + 1) & T.cap_mask;
    }
}

// Unpack a dense lower-triangular L (row-major n x n) from the tile array, for gpis_leaf_get.
__global__ void k_unpack_L(const float* __restrict__ tiles, int n, int nb, float* __restrict__ Ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const int r = i / n, c = i % n;
    float v = 0.f;
    if (c <= r) {
        const int bi = r >> 5, bj = c >> 5;
        if (bi == bj) v = tiles[(size_t)tile_index(bi, bj, nb) * GPIS_TILE_ELEMS + (c & 31) * 32 + (r & 31)];
        else {  // off-diagonal tiles hold G = L_ij inv(Ljj): L_ij = G Ljj
            const float* G = tiles + (size_t)tile_index(bi, bj, nb) * GPIS_TILE_ELEMS;
            const float* Ljj = tiles + (size_t)tile_index(bj, bj, nb) * GPIS_TILE_ELEMS;
            for (int k = (c & 31); k < 32; ++k) v = fmaf(G[k * 32 + (r & 31)], Ljj[(c & 31) * 32 + k], v);
        }
    }
    Ld[i] = v;
}

// K5: gather / scatter of whole leaf records between the arena and one contiguous staging buffer.
struct CopyJob { const unsigned char* src; unsigned char* dst; uint64_t bytes; };
__global__ void __launch_bounds__(256) k_copy_records(const CopyJob* __restrict__ jobs) {
    const CopyJob j = jobs[blockIdx.x];
    const uint4* s = reinterpret_cast<const uint4*>(j.src);
    uint4* d = reinterpret_cast<uint4*>(j.dst);
    const uint64_t n16 = j.bytes >> 4;   // records are multiples of 128 bytes
    for (uint64_t i = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.y * blockDim.x) d[i] = s[i];
}

}  // namespace gpis

// ------------------------------------------------------------------ context
struct HostLeaf {
    int slot;
    uint64_t rec;      // device address, 0 = untrained
    uint64_t rec_bytes;
    int N, ng, n, nb;
    int cell[3];
    float centre[3];
    float lo[3], hi[3];   // effective box (default: centre -/+ cluster_half)
    bool box_set;
    uint64_t smp;      // device sample block of this leaf (gpis_samples_set), 0 = none
    uint64_t smp_bytes;
    int smp_n;
};

struct ArenaChunk { unsigned char* base; uint64_t size; };

// One training batch between "records reserved, samples gathered, jobs uploaded" and "records installed in the table".
// With gpis_set_train_mode(ctx, 1 | 2) the batch stays in this state after gpis_leaves_train_dirty returns: K1 runs on
// its own stream while the host works on the next frame, and every entry point that reads or changes records, the
// table or the arena completes it first (train_flush).
struct TrainPlan { uint64_t key; int N, ng, n, nb; uint64_t rec, rb; };
struct PendingTrain {
    bool active = false, launched = false;
    std::vector<TrainPlan> plan;
    int njobs = 0, maxN = 1, maxnb = 1;
};

struct gpis_ctx {
    gpis_config cfg;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // asynchronous leaf training (gpis_set_train_mode)
    cudaStream_t copy_stream[2] = {nullptr, nullptr};        // gpis_query with pinned host buffers: H2D | D2H beside the evaluation
    cudaEvent_t ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t train_stream = nullptr;
    cudaEvent_t ev_train[3] = {nullptr, nullptr, nullptr};   // inputs ready (main stream) | K1 start | K1 done (training stream)
    int train_mode = 0;                                      // 0 synchronous, 1 launch at once, 2 launch at the next gpis_reeval
    PendingTrain pend;
    void* d_train_smp = nullptr; uint64_t train_smp_bytes = 0;     // gathered training balls (CSR) of the pending batch
    void* d_train_jobs = nullptr; uint64_t train_jobs_bytes = 0;   // its jobs + status
    int* d_k1_counter = nullptr;                                   // job counter of the persistent K1 CTAs
    bool train_ms_pending = false;                                 // ev_train[1..2] recorded, elapsed time not read yet
    // leaf table
    LeafTable T{};
    uint32_t table_cap = 0;
    int slot_cap = 0, slot_count = 0, tombs = 0;
    std::vector<int> free_slots;
    std::unordered_map<uint64_t, HostLeaf> leaves;
    // arena
    std::vector<ArenaChunk> chunks;
    std::map<uint64_t, uint64_t> free_blocks;  // address -> size
    std::set<std::pair<uint64_t, uint64_t>> free_by_size;   // (size, address) of the same blocks
    uint64_t arena_used = 0, arena_reserved = 0;
    // params
    QueryParams qp{};
    TrainParams tp{};
    int max_nb = 1;
    int max_N = 1;
    // scratch
    void* d_scratch = nullptr; uint64_t scratch_bytes = 0;       // generic upload buffer
    void* d_scratch2 = nullptr; uint64_t scratch2_bytes = 0;
    void* d_jobs = nullptr; uint64_t jobs_bytes = 0;             // training jobs + status
    int num_sms = 148;
    // device sample store (f-2)
    SampleStore store{nullptr, nullptr};
    std::vector<uint64_t> slot_key;                              // slot -> key of the leaf that owns it
    void* d_gather = nullptr; uint64_t gather_bytes = 0;         // flags, dirty list, counts, offsets
    void* d_frame = nullptr; uint64_t frame_bytes = 0;           // per-frame sensor pipeline (gpis_frame_eval)
    void* d_reeval = nullptr; uint64_t reeval_bytes = 0;         // gpis_reeval
    QueryWork W{}; int64_t work_cap = 0;
    void* d_x = nullptr; void* d_res = nullptr; int64_t q_cap = 0;
    int32_t* d_sort = nullptr; int64_t sort_cap = 0;
    EvalProg prog{nullptr, nullptr};   // elimination programs of k_eval_v3 (query_v3.cuh)
    double* d_acc = nullptr;
    // obs gp
    ObsTile* obs_tiles = nullptr; int obs_tile_cap = 0;
    ObsTileDesc* obs_desc = nullptr;
    float* obs_b0 = nullptr; float* obs_b1 = nullptr; int obs_b_cap = 0;
    ObsParams op{}; bool obs_trained = false;
    int obs_ni = -1, obs_nj = -1; bool obs_repartition = true;
    std::vector<float> obs_hb0, obs_hb1; std::vector<ObsTileDesc> obs_hdesc;
    // K5 replication (gpis_comm_init / gpis_replicate): NCCL resolved at run time
    void* nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    ncclResult_t (*p_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*p_ncclBroadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*p_ncclCommDestroy)(ncclComm_t) = nullptr;
    const char* (*p_ncclGetErrorString)(ncclResult_t) = nullptr;
    std::unordered_set<uint64_t> repl_touched;   // leaves whose table entry or record changed since the last gpis_replicate
    std::unordered_set<uint64_t> repl_trained;   // ... of which the record is new
    std::vector<uint64_t> repl_erased;           // leaves erased since then
    void* d_repl = nullptr; uint64_t repl_cap = 0;      // staging buffer (bounded: records travel in chunks)
    void* d_repl_idx = nullptr; uint64_t repl_idx_cap = 0;
    void* d_repl_jobs = nullptr; uint64_t repl_jobs_cap = 0;
    gpis_stats st{};
    int eval_version = 3;
};

// NVTX range per C-ABI call / phase (the reference has no tracing at all)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return GPIS_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

extern "C" {
static int train_launch(gpis_ctx* ctx);   // start the pending batch's K1 (no wait)
static int train_flush(gpis_ctx* ctx);    // ... and wait for it, install its records
static void train_discard(gpis_ctx* ctx); // wait and drop it (reset / destroy)
}

static int ensure(gpis_ctx* ctx, void** p, uint64_t* cap, uint64_t need) {
    if (*cap >= need) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    uint64_t sz = std::max<uint64_t>(need, 1 << 20);
    sz = sz + sz / 4;
    CK(cudaMalloc(p, sz));
    *cap = sz;
    return 0;
}

static uint64_t key_of(const gpis_ctx* ctx, const int32_t* cell) {
    return cell_key(cell[0], cell[1], ctx->cfg.dim == 3 ? cell[2] : 0);
}

// ---- arena: best-fit free list over cudaMalloc'd chunks
// Free blocks are indexed twice: by address (coalescing) and by size (best fit in O(log F): a frame reserves ~10^3
// records of megabytes between thousands of kilobyte-sized sample-list holes; a first-fit scan cost 5-15 ms per frame).
static void free_add(gpis_ctx* ctx, uint64_t addr, uint64_t size) {
    ctx->free_blocks[addr] = size;
    ctx->free_by_size.insert({size, addr});
}
static void free_del(gpis_ctx* ctx, std::map<uint64_t, uint64_t>::iterator it) {
    ctx->free_by_size.erase({it->second, it->first});
    ctx->free_blocks.erase(it);
}
static int arena_alloc(gpis_ctx* ctx, uint64_t bytes, uint64_t* out) {
    bytes = align_up(bytes, 256);
    auto bs = ctx->free_by_size.lower_bound({bytes, 0});
    if (bs != ctx->free_by_size.end()) {
        const uint64_t sz = bs->first, addr = bs->second;
        free_del(ctx, ctx->free_blocks.find(addr));
        if (sz > bytes) free_add(ctx, addr + bytes, sz - bytes);
        ctx->arena_used += bytes;
        *out = addr;
        return 0;
    }
    // Chunks grow geometrically (each new one as large as everything reserved so far, at most 32 GiB): cudaMalloc is
    // cheap on an idle device (0.6 ms for 1 GiB, 3 ms for 32 GiB, scripts/probe/malloc_probe.py) but waits for the
    // kernels in flight, so a map that grows by a 1 GiB chunk every other frame paid 5-100 ms each time.
    const uint64_t min_csz = std::max<uint64_t>(ctx->cfg.arena_chunk_bytes, align_up(bytes, 1 << 20));
    uint64_t csz = std::max<uint64_t>(min_csz, std::min<uint64_t>(ctx->arena_reserved, 32ull << 30));
    unsigned char* base = nullptr;
    static const bool prof_sync = std::getenv("GPIS_PROFILE") != nullptr;
    double sync_ms = 0.;
    if (prof_sync) {
        const auto ts = std::chrono::steady_clock::now();
        cudaDeviceSynchronize();
        sync_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts).count();
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (cudaMalloc(&base, csz) != cudaSuccess) {      // not that much left: take what is needed
        (void)cudaGetLastError();
        csz = min_csz;
        CK(cudaMalloc(&base, csz));
    }
    if (std::getenv("GPIS_PROFILE"))
        std::fprintf(stderr, "arena: new %.2f GiB chunk in %.2f ms (device sync before it: %.2f ms)\n", csz / 1073741824.0,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), sync_ms);
    ctx->chunks.push_back({base, csz});
    ctx->arena_reserved += csz;
    free_add(ctx, (uint64_t)base, csz);
    return arena_alloc(ctx, bytes, out);
}
static bool same_chunk(const gpis_ctx* ctx, uint64_t a, uint64_t b) {
    for (auto& c : ctx->chunks) if (a >= (uint64_t)c.base && b < (uint64_t)c.base + c.size) return true;
    return false;
}
static void arena_free(gpis_ctx* ctx, uint64_t addr, uint64_t bytes) {
    bytes = align_up(bytes, 256);
    ctx->arena_used -= bytes;
    // coalesce with the next / previous block when contiguous inside one chunk
    auto nx = ctx->free_blocks.lower_bound(addr);
    if (nx != ctx->free_blocks.end() && addr + bytes == nx->first && same_chunk(ctx, addr, nx->first)) {
        bytes += nx->second;
        auto dead = nx++;
        free_del(ctx, dead);
    }
    if (nx != ctx->free_blocks.begin()) {
        auto pv = std::prev(nx);
        if (pv->first + pv->second == addr && same_chunk(ctx, pv->first, addr)) {
            addr = pv->first;
            bytes += pv->second;
            free_del(ctx, pv);
        }
    }
    free_add(ctx, addr, bytes);
}

// ---- table storage
static int table_alloc(gpis_ctx* ctx, uint32_t cap, int slot_cap) {
    LeafTable N{};
    CK(cudaMalloc(&N.keys, sizeof(uint64_t) * cap));
    CK(cudaMalloc(&N.vals, sizeof(int32_t) * cap));
    CK(cudaMemsetAsync(N.keys, 0, sizeof(uint64_t) * cap, ctx->stream));
    CK(cudaMemsetAsync(N.vals, 0xff, sizeof(int32_t) * cap, ctx->stream));
    N.cap_mask = cap - 1;
    CK(cudaMalloc(&N.centre, sizeof(float4) * slot_cap));
    CK(cudaMalloc(&N.lo, sizeof(float4) * slot_cap));
    CK(cudaMalloc(&N.hi, sizeof(float4) * slot_cap));
    CK(cudaMalloc(&N.cell, sizeof(int4) * slot_cap));
    CK(cudaMalloc(&N.rec, sizeof(uint64_t) * slot_cap));
    CK(cudaMalloc(&N.meta, sizeof(int4) * slot_cap));
    CK(cudaMemsetAsync(N.cell, 0, sizeof(int4) * slot_cap, ctx->stream));
    CK(cudaMemsetAsync(N.rec, 0, sizeof(uint64_t) * slot_cap, ctx->stream));
    if (slot_cap != ctx->slot_cap || !ctx->store.ptr) {   // the sample store is indexed by slot too
        SampleStore S{nullptr, nullptr};
        CK(cudaMalloc(&S.ptr, sizeof(uint64_t) * slot_cap));
        CK(cudaMalloc(&S.cnt, sizeof(int32_t) * slot_cap));
        CK(cudaMemsetAsync(S.ptr, 0, sizeof(uint64_t) * slot_cap, ctx->stream));
        CK(cudaMemsetAsync(S.cnt, 0, sizeof(int32_t) * slot_cap, ctx->stream));
        if (ctx->store.ptr) {
            CK(cudaMemcpyAsync(S.ptr, ctx->store.ptr, sizeof(uint64_t) * ctx->slot_cap, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(S.cnt, ctx->store.cnt, sizeof(int32_t) * ctx->slot_cap, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->store.ptr); cudaFree(ctx->store.cnt);
        }
        ctx->store = S;
        ctx->slot_key.resize(slot_cap, 0);
    }
    if (ctx->T.keys) {
        const int old = ctx->slot_cap;
        CK(cudaMemcpyAsync(N.centre, ctx->T.centre, sizeof(float4) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(N.lo, ctx->T.lo, sizeof(float4) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(N.hi, ctx->T.hi, sizeof(float4) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(N.cell, ctx->T.cell, sizeof(int4) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(N.rec, ctx->T.rec, sizeof(uint64_t) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(N.meta, ctx->T.meta, sizeof(int4) * old, cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->slot_count > 0) {
            k_table_rebuild<<<(ctx->slot_count + 255) / 256, 256, 0, ctx->stream>>>(N, ctx->slot_count);
            ctx->st.kernel_launches++;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->T.keys); cudaFree(ctx->T.vals); cudaFree(ctx->T.centre); cudaFree(ctx->T.cell);
        cudaFree(ctx->T.rec); cudaFree(ctx->T.meta); cudaFree(ctx->T.lo); cudaFree(ctx->T.hi);
    }
    ctx->T = N;
    ctx->table_cap = cap;
    ctx->slot_cap = slot_cap;
    ctx->tombs = 0;
    return 0;
}
static int table_reserve(gpis_ctx* ctx, int extra) {
    const int need_slots = ctx->slot_count + extra;
    if (need_slots <= ctx->slot_cap && (uint64_t)(ctx->leaves.size() + ctx->tombs + extra) * 2 <= ctx->table_cap) return 0;
    int sc = std::max(ctx->slot_cap, 1024);
    while (sc < need_slots) sc *= 2;
    uint32_t cap = std::max<uint32_t>(ctx->table_cap, 4096);
    while ((uint64_t)(ctx->leaves.size() + extra) * 4 > cap) cap *= 2;
    return table_alloc(ctx, cap, sc);
}

static void derive_params(gpis_ctx* ctx) {
    const gpis_config& c = ctx->cfg;
    QueryParams& q = ctx->qp;
    q.dim = c.dim;
    q.cluster_half = c.cluster_half;
    q.search_half = c.search_half;
    q.var_thre = c.var_thre;
    q.var_preset = (float)(1.0 + (double)c.map_noise);                 // GPisMap3.cpp:816
    q.a = (float)(std::sqrt(3.0) / (double)c.map_scale);               // covFnc.cpp:263
    const float three_over_scale = (float)(3.0 / (double)(c.map_scale * c.map_scale));  // OnGPIS.h:58
    if (c.dim == 3) { q.prior_f = 1.001; q.prior_g = (double)three_over_scale + 0.001; }   // OnGPIS.cpp:203-212
    else            { q.prior_f = 1.01;  q.prior_g = (double)three_over_scale + 0.1; }     // OnGPIS.cpp:235-237
    q.inv_pitch = 1.0 / (2.0 * (double)c.cluster_half);
    q.root_min[0] = q.root_min[1] = q.root_min[2] = -(1 << 19);
    q.levels = 20;
    TrainParams& t = ctx->tp;
    t.dim = c.dim;
    t.scale = c.map_scale;
    t.a = (float)(std::sqrt(3.0) / (double)c.map_scale);               // covFnc.cpp:147
    t.a2 = t.a * t.a;
    t.refine = 1;
    if (const char* e = std::getenv("GPIS_REFINE")) t.refine = std::atoi(e) ? 1 : 0;   // development switch (profiles/r02_history.md)
    ObsParams& o = ctx->op;
    o.a = 1 / c.obs_scale;                                             // covFnc.cpp:51
    o.diag = (float)(1.0 + (double)c.obs_noise);                       // covFnc.cpp:57
    o.var_prior = 1 + c.obs_noise;                                     // ObsGP.cpp:61
}

extern "C" {

int gpis_config_default(gpis_config* cfg, int dim) {
    if (!cfg || (dim != 2 && dim != 3)) return GPIS_ERR_ARG;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->dim = dim;
    cfg->device = 0;
    if (dim == 3) {
        cfg->map_scale = 0.04f;                  // params.h:92
        cfg->map_noise = 5e-3f;                  // params.h:93
        cfg->cluster_half = (float)0.025;        // params.h:41
        cfg->search_half = (float)0.025 * 3.0;   // GPisMap3.cpp:811  C_leng*3.0
        cfg->var_thre = 0.5f;                    // GPisMap3.cpp:800
    } else {
        cfg->map_scale = 1.2f;                   // params.h:73
        cfg->map_noise = 1e-2f;                  // params.h:74
        cfg->cluster_half = (float)0.8;          // params.h:34
        cfg->search_half = 1.2f * 4.0;           // GPisMap.cpp:680   map_scale_param*4.0
        cfg->var_thre = 0.4f;                    // GPisMap.cpp:671
    }
    cfg->obs_scale = 0.5f;                       // params.h:97
    cfg->obs_noise = 0.01f;                      // params.h:98
    cfg->max_leaves = 1 << 16;
    cfg->arena_chunk_bytes = 1ull << 30;
    return GPIS_OK;
}

int gpis_create(gpis_ctx** out, const gpis_config* cfg) {
    if (!out || !cfg || (cfg->dim != 2 && cfg->dim != 3)) return GPIS_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) return GPIS_ERR_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return GPIS_ERR_NODEVICE;
    if (prop.major != 10) {
        std::fprintf(stderr, "gpis_b200: device %d is sm_%d%d; this library is built for sm_100a only and has no fallback\n",
                     cfg->device, prop.major, prop.minor);
        return GPIS_ERR_NODEVICE;
    }
    {   // the exact-tie replay keeps at most GPIS_MAXC candidates per query (query.cuh): refuse a search box that can exceed it
        const double per_axis = 2.0 * std::ceil((double)cfg->search_half / (2.0 * (double)cfg->cluster_half)) + 2.0;
        if (!(cfg->cluster_half > 0.f) || !(cfg->search_half > 0.f) || std::pow(per_axis, cfg->dim) > GPIS_MAXC) {
            std::fprintf(stderr, "gpis_b200: search_half / cluster_half = %g gives up to %g candidate leaves per query; the limit is %d\n",
                         (double)cfg->search_half / (double)cfg->cluster_half, std::pow(per_axis, cfg->dim), GPIS_MAXC);
            return GPIS_ERR_ARG;
        }
    }
    gpis_ctx* ctx = new gpis_ctx();
    ctx->cfg = *cfg;
    if (ctx->cfg.arena_chunk_bytes < (64ull << 20)) ctx->cfg.arena_chunk_bytes = 64ull << 20;
    *out = ctx;
    CK(cudaSetDevice(cfg->device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ctx->ev[i]));
    {   // the training stream yields to the main stream: a frame's small kernels should not queue behind K1's CTAs
        int lo_prio = 0, hi_prio = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CK(cudaStreamCreateWithPriority(&ctx->train_stream, cudaStreamNonBlocking, lo_prio));
        for (int i = 0; i < 3; ++i) CK(cudaEventCreate(&ctx->ev_train[i]));
        CK(cudaMalloc(&ctx->d_k1_counter, 256));
        for (int i = 0; i < 2; ++i) CK(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
    }
    derive_params(ctx);
    uint32_t cap = 4096;
    while (cap < (uint32_t)std::max(1, cfg->max_leaves) * 2u) cap *= 2;
    int rc = table_alloc(ctx, cap, std::max(1024, cfg->max_leaves));
    if (rc) return rc;
    CK(cudaFuncSetAttribute(k_leaf_train, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->num_sms = prop.multiProcessorCount;
    CK(cudaFuncSetAttribute(k_eval_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    rc = query_eval_init(ctx->err);
    if (rc == 0) rc = e3_upload_programs(const_cast<int4**>(&ctx->prog.recs), const_cast<int32_t**>(&ctx->prog.off), ctx->err);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return GPIS_OK;
}

void gpis_destroy(gpis_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    train_discard(ctx);
    cudaDeviceSynchronize();
    for (auto& c : ctx->chunks) cudaFree(c.base);
    cudaFree(ctx->T.keys); cudaFree(ctx->T.vals); cudaFree(ctx->T.centre); cudaFree(ctx->T.cell);
    cudaFree(ctx->T.rec); cudaFree(ctx->T.meta); cudaFree(ctx->T.lo); cudaFree(ctx->T.hi);
    cudaFree(ctx->d_scratch); cudaFree(ctx->d_scratch2); cudaFree(ctx->d_jobs);
    cudaFree(ctx->W.cand); cudaFree(ctx->W.tie); cudaFree(ctx->W.evalout); cudaFree(ctx->W.pairs); cudaFree(ctx->W.counters);
    cudaFree(ctx->d_x); cudaFree(ctx->d_res); cudaFree(ctx->d_sort);
    cudaFree(const_cast<int4*>(ctx->prog.recs)); cudaFree(const_cast<int32_t*>(ctx->prog.off));
    cudaFree(ctx->obs_tiles); cudaFree(ctx->obs_desc); cudaFree(ctx->obs_b0); cudaFree(ctx->obs_b1);
    cudaFree(ctx->d_acc); cudaFree(ctx->store.ptr); cudaFree(ctx->store.cnt); cudaFree(ctx->d_gather); cudaFree(ctx->d_frame); cudaFree(ctx->d_reeval);
    cudaFree(ctx->d_repl); cudaFree(ctx->d_repl_idx); cudaFree(ctx->d_repl_jobs);
    cudaFree(ctx->d_train_smp); cudaFree(ctx->d_train_jobs); cudaFree(ctx->d_k1_counter);
    for (int i = 0; i < 3; ++i) if (ctx->ev_train[i]) cudaEventDestroy(ctx->ev_train[i]);
    if (ctx->train_stream) cudaStreamDestroy(ctx->train_stream);
    for (int i = 0; i < 2; ++i) if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
    for (int i = 0; i < 4; ++i) if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    if (ctx->comm && ctx->p_ncclCommDestroy) ctx->p_ncclCommDestroy(ctx->comm);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* gpis_last_error(const gpis_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int gpis_device(const gpis_ctx* ctx) { return ctx ? ctx->cfg.device : -1; }

int gpis_reset(gpis_ctx* ctx) {
    if (!ctx) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    train_discard(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->leaves.clear();
    ctx->free_slots.clear();
    ctx->slot_count = 0;
    ctx->tombs = 0;
    ctx->free_blocks.clear();
    ctx->free_by_size.clear();
    for (auto& c : ctx->chunks) free_add(ctx, (uint64_t)c.base, c.size);
    ctx->arena_used = 0;
    CK(cudaMemsetAsync(ctx->T.keys, 0, sizeof(uint64_t) * ctx->table_cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->T.cell, 0, sizeof(int4) * ctx->slot_cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->T.rec, 0, sizeof(uint64_t) * ctx->slot_cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->store.ptr, 0, sizeof(uint64_t) * ctx->slot_cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->store.cnt, 0, sizeof(int32_t) * ctx->slot_cap, ctx->stream));
    ctx->obs_trained = false;
    ctx->obs_repartition = true;   // ObsGP2D::reset (ObsGP.cpp:198-203) runs when the map deletes gpo
    ctx->obs_ni = ctx->obs_nj = -1;
    ctx->repl_touched.clear(); ctx->repl_trained.clear(); ctx->repl_erased.clear();
    ctx->max_nb = 1; ctx->max_N = 1;
    derive_params(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return GPIS_OK;
}

int gpis_rebase(gpis_ctx* ctx, const int32_t* root_min_cell, int levels) {
    if (!ctx || !root_min_cell || levels < 0 || levels > 20) return GPIS_ERR_ARG;
    for (int c = 0; c < 3; ++c) ctx->qp.root_min[c] = (c < ctx->cfg.dim) ? root_min_cell[c] : 0;
    ctx->qp.levels = levels;
    return GPIS_OK;
}

int gpis_get_rebase(gpis_ctx* ctx, int32_t* root_min_cell, int* levels) {
    if (!ctx || !root_min_cell || !levels) return GPIS_ERR_ARG;
    for (int c = 0; c < 3; ++c) root_min_cell[c] = ctx->qp.root_min[c];
    *levels = ctx->qp.levels;
    return GPIS_OK;
}

// gradflag rule shared with the kernel (OnGPIS.cpp:63-66, 122-125)
static inline bool grad_valid(const float* s, int dim) {
    bool allsmall = true;
    for (int c = 0; c < dim; ++c) allsmall = allsmall && (std::fabs((double)s[dim + c]) < 1e-6);
    return !((double)s[2 * dim + 2] > 0.1001 || allsmall);
}

static int apply_updates(gpis_ctx* ctx, const std::vector<SlotUpdate>& ups) {
    if (ups.empty()) return 0;
    int rc = ensure(ctx, &ctx->d_scratch2, &ctx->scratch2_bytes, ups.size() * sizeof(SlotUpdate));
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_scratch2, ups.data(), ups.size() * sizeof(SlotUpdate), cudaMemcpyHostToDevice, ctx->stream));
    k_table_apply<<<((int)ups.size() + 127) / 128, 128, 0, ctx->stream>>>(ctx->T, (const SlotUpdate*)ctx->d_scratch2, (int)ups.size());
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int store_apply(gpis_ctx* ctx, const std::vector<StoreUpdate>& ups) {
    if (ups.empty()) return 0;
    int rc = ensure(ctx, &ctx->d_scratch2, &ctx->scratch2_bytes, ups.size() * sizeof(StoreUpdate));
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_scratch2, ups.data(), ups.size() * sizeof(StoreUpdate), cudaMemcpyHostToDevice, ctx->stream));
    k_store_apply<<<((int)ups.size() + 127) / 128, 128, 0, ctx->stream>>>(ctx->store, (const StoreUpdate*)ctx->d_scratch2, (int)ups.size());
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// centre / default box of a leaf; keeps an explicitly set box (gpis_leaves_set_boxes)
static void set_geometry(gpis_ctx* ctx, HostLeaf& hl, const int32_t* cell, const float* centre) {
    const int dim = ctx->cfg.dim;
    for (int c = 0; c < 3; ++c) { hl.cell[c] = 0; hl.centre[c] = 0.f; }
    for (int c = 0; c < dim; ++c) { hl.cell[c] = cell[c]; hl.centre[c] = centre[c]; }
    if (!hl.box_set)
        for (int c = 0; c < 3; ++c) {
            hl.lo[c] = (c < dim) ? hl.centre[c] - ctx->cfg.cluster_half : 0.f;   // octree.h:73-78
            hl.hi[c] = (c < dim) ? hl.centre[c] + ctx->cfg.cluster_half : 0.f;
        }
}
static SlotUpdate make_update(uint64_t key, const HostLeaf& hl) {
    SlotUpdate u{};
    u.key = key; u.rec = hl.rec; u.slot = hl.slot; u.live = 1;
    for (int c = 0; c < 3; ++c) { u.cell[c] = hl.cell[c]; u.centre[c] = hl.centre[c]; u.lo[c] = hl.lo[c]; u.hi[c] = hl.hi[c]; }
    u.meta[0] = hl.N; u.meta[1] = hl.ng; u.meta[2] = hl.n; u.meta[3] = hl.nb;
    return u;
}

static int take_slot(gpis_ctx* ctx, uint64_t key) {
    int s;
    if (!ctx->free_slots.empty()) { s = ctx->free_slots.back(); ctx->free_slots.pop_back(); }
    else s = ctx->slot_count++;
    if ((size_t)s >= ctx->slot_key.size()) ctx->slot_key.resize(s + 1, 0);
    ctx->slot_key[s] = key;
    return s;
}

int gpis_leaves_mark(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres) {
    if (!ctx || n_leaves < 0 || (n_leaves > 0 && (!cells || !centres))) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const int dim = ctx->cfg.dim;
    int rc = table_reserve(ctx, n_leaves);
    if (rc) return rc;
    std::vector<SlotUpdate> ups;
    for (int i = 0; i < n_leaves; ++i) {
        const uint64_t key = key_of(ctx, cells + (size_t)i * dim);
        if (ctx->leaves.count(key)) continue;
        HostLeaf hl{};
        hl.slot = take_slot(ctx, key);
        set_geometry(ctx, hl, cells + (size_t)i * dim, centres + (size_t)i * dim);
        ctx->leaves[key] = hl;
        ups.push_back(make_update(key, hl));
        ctx->repl_touched.insert(key);
    }
    return apply_updates(ctx, ups);
}

int gpis_leaves_set_boxes(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* boxes) {
    if (!ctx || n_leaves < 0 || (n_leaves > 0 && (!cells || !boxes))) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const int dim = ctx->cfg.dim;
    std::vector<SlotUpdate> ups;
    for (int i = 0; i < n_leaves; ++i) {
        const uint64_t key = key_of(ctx, cells + (size_t)i * dim);
        auto it = ctx->leaves.find(key);
        if (it == ctx->leaves.end()) continue;
        HostLeaf& hl = it->second;
        for (int c = 0; c < dim; ++c) { hl.lo[c] = boxes[(size_t)i * 2 * dim + c]; hl.hi[c] = boxes[(size_t)i * 2 * dim + dim + c]; }
        hl.box_set = true;
        ups.push_back(make_update(key, hl));
        ctx->repl_touched.insert(key);
    }
    return apply_updates(ctx, ups);
}

int gpis_leaves_erase(gpis_ctx* ctx, int n_leaves, const int32_t* cells) {
    if (!ctx || n_leaves < 0 || (n_leaves > 0 && !cells)) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const int dim = ctx->cfg.dim;
    std::vector<SlotUpdate> ups;
    std::vector<StoreUpdate> store_clear;
    for (int i = 0; i < n_leaves; ++i) {
        const uint64_t key = key_of(ctx, cells + (size_t)i * dim);
        auto it = ctx->leaves.find(key);
        if (it == ctx->leaves.end()) continue;
        SlotUpdate u{};
        u.key = key; u.slot = it->second.slot; u.live = 0;
        ups.push_back(u);
        if (it->second.rec) arena_free(ctx, it->second.rec, it->second.rec_bytes);
        if (it->second.smp) { arena_free(ctx, it->second.smp, it->second.smp_bytes); store_clear.push_back(StoreUpdate{it->second.slot, 0, 0ull}); }
        ctx->free_slots.push_back(it->second.slot);
        ctx->leaves.erase(it);
        ctx->tombs++;
        ctx->repl_touched.erase(key); ctx->repl_trained.erase(key);
        ctx->repl_erased.push_back(key);
    }
    int rc = apply_updates(ctx, ups);
    if (rc) return rc;
    rc = store_apply(ctx, store_clear);
    if (rc) return rc;
    if (ctx->tombs * 4 > (int)ctx->table_cap) return table_alloc(ctx, ctx->table_cap, ctx->slot_cap);
    return 0;
}

// Launch K1 on a batch of jobs whose records are allocated and whose samples sit in d_smp (device). Jobs are
// sorted biggest-first (one CTA per leaf, the hardware scheduler balances the tail). status_host (optional,
// job order) receives the non-positive-pivot counts.
static int train_jobs(gpis_ctx* ctx, std::vector<TrainJob>& jobs, const float* d_smp, int maxN, int maxnb,
                      int32_t* status_host, float* ms_out) {
    NvtxRange nvtx_("K1 leaf train");
    *ms_out = 0.f;
    if (jobs.empty()) return 0;
    std::vector<int> order(jobs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jobs[a].n > jobs[b].n; });
    std::vector<TrainJob> sorted(jobs.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = jobs[order[i]];
    const uint64_t b_jobs = align_up(sorted.size() * sizeof(TrainJob), 256);
    const uint64_t b_st = align_up(sorted.size() * sizeof(int32_t), 256);
    int rc = ensure(ctx, &ctx->d_jobs, &ctx->jobs_bytes, b_jobs + b_st);
    if (rc) return rc;
    TrainJob* d_jobs = (TrainJob*)ctx->d_jobs;
    int32_t* d_st = (int32_t*)((unsigned char*)ctx->d_jobs + b_jobs);
    CK(cudaMemcpyAsync(d_jobs, sorted.data(), sorted.size() * sizeof(TrainJob), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = launch_leaf_train(ctx->stream, d_jobs, (int)sorted.size(), d_smp, ctx->tp, d_st, maxN, maxnb, 2 * ctx->num_sms, ctx->d_k1_counter, ctx->err);
    if (rc) return rc;
    ctx->st.kernel_launches++;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    std::vector<int32_t> st(sorted.size());
    CK(cudaMemcpyAsync(st.data(), d_st, sorted.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(ms_out, ctx->ev[0], ctx->ev[1]));
    if (status_host)
        for (size_t i = 0; i < order.size(); ++i) status_host[order[i]] = st[i];
    return 0;
}

// ---- asynchronous training (gpis_set_train_mode). The batch's device-side inputs (gathered balls in d_train_smp, jobs
// in d_train_jobs) were written on the main stream; K1 runs on the training stream behind an event.
static int train_launch(gpis_ctx* ctx) {
    PendingTrain& pd = ctx->pend;
    if (!pd.active || pd.launched) return 0;
    NvtxRange nvtx_("K1 leaf train (launch)");
    const uint64_t b_jobs = align_up((uint64_t)pd.njobs * sizeof(TrainJob), 256);
    TrainJob* d_jobs = (TrainJob*)ctx->d_train_jobs;
    int32_t* d_st = (int32_t*)((unsigned char*)ctx->d_train_jobs + b_jobs);
    CK(cudaEventRecord(ctx->ev_train[0], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->train_stream, ctx->ev_train[0], 0));
    CK(cudaEventRecord(ctx->ev_train[1], ctx->train_stream));
    // overlapped launches keep a few CTA slots free: the on-demand observation tests of the frame in progress (a handful
    // of points each) must not wait for a leaf to finish
    static const int reserve = std::getenv("GPIS_TRAIN_RESERVE") ? std::atoi(std::getenv("GPIS_TRAIN_RESERVE")) : 32;   // CTA slots of 296
    const int slots = 2 * ctx->num_sms - (ctx->train_mode == 0 ? 0 : std::max(0, std::min(reserve, ctx->num_sms)));
    const int rc = launch_leaf_train(ctx->train_stream, d_jobs, pd.njobs, (const float*)ctx->d_train_smp, ctx->tp, d_st, pd.maxN, pd.maxnb, slots,
                                     ctx->d_k1_counter, ctx->err);
    if (rc) return rc;
    ctx->st.kernel_launches++;
    CK(cudaEventRecord(ctx->ev_train[2], ctx->train_stream));
    pd.launched = true;
    ctx->train_ms_pending = true;
    return 0;
}
static void train_read_ms(gpis_ctx* ctx, bool wait) {
    if (!ctx->train_ms_pending) return;
    if (!wait && cudaEventQuery(ctx->ev_train[2]) != cudaSuccess) return;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev_train[2]) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev_train[1], ctx->ev_train[2]) == cudaSuccess)
        ctx->st.last_train_ms = ms;
    ctx->train_ms_pending = false;
}
static int train_flush(gpis_ctx* ctx) {
    PendingTrain& pd = ctx->pend;
    if (!pd.active) return 0;
    NvtxRange nvtx_("K1 leaf train (wait + install)");
    int rc = train_launch(ctx);
    if (rc) {   // nothing ran: give the reserved records back, the leaves keep their previous GPs
        for (auto& q : pd.plan) arena_free(ctx, q.rec, q.rb);
        pd = PendingTrain{};
        return rc;
    }
    CK(cudaEventSynchronize(ctx->ev_train[2]));
    train_read_ms(ctx, true);
    std::vector<SlotUpdate> ups;
    std::vector<std::pair<uint64_t, uint64_t>> to_free;
    for (const TrainPlan& pl : pd.plan) {
        auto it = ctx->leaves.find(pl.key);
        if (it == ctx->leaves.end()) { arena_free(ctx, pl.rec, pl.rb); continue; }   // cannot happen: erase flushes first
        HostLeaf& hl = it->second;
        if (hl.rec) to_free.push_back({hl.rec, hl.rec_bytes});
        hl.rec = pl.rec; hl.rec_bytes = pl.rb; hl.N = pl.N; hl.ng = pl.ng; hl.n = pl.n; hl.nb = pl.nb;
        ups.push_back(make_update(pl.key, hl));
        ctx->repl_touched.insert(pl.key);
        ctx->repl_trained.insert(pl.key);
    }
    ctx->max_nb = std::max(ctx->max_nb, pd.maxnb);
    ctx->max_N = std::max(ctx->max_N, pd.maxN);
    pd = PendingTrain{};
    rc = apply_updates(ctx, ups);
    if (rc) return rc;
    for (auto& f : to_free) arena_free(ctx, f.first, f.second);
    return 0;
}
static void train_discard(gpis_ctx* ctx) {
    if (ctx->train_stream) cudaStreamSynchronize(ctx->train_stream);
    ctx->train_ms_pending = false;
    ctx->pend = PendingTrain{};
}

int gpis_set_train_mode(gpis_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 3) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    const int rc = train_flush(ctx);
    ctx->train_mode = mode;
    return rc;
}
int gpis_train_kick(gpis_ctx* ctx) {
    if (!ctx) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    return train_launch(ctx);
}
int gpis_train_wait(gpis_ctx* ctx) {
    if (!ctx) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    return train_flush(ctx);
}

int gpis_leaves_update(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres,
                       const int32_t* offsets, const float* samples, int32_t* status) {
    NvtxRange nvtx_("gpis_leaves_update");
    if (!ctx || n_leaves < 0) return GPIS_ERR_ARG;
    if (n_leaves == 0) return GPIS_OK;
    if (!cells || !centres || !offsets || !samples) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const int dim = ctx->cfg.dim, w9 = 2 * dim + 3;

    // ---- pass 1: validate and size every leaf; nothing is mutated yet. A leaf beyond the kernel's capacity
    // (the reference has no limit, OnGPIS.cpp:91-149) is skipped — it keeps whatever GP it had — and flagged in
    // status[]; the rest of the batch is trained.
    struct Plan { int N, ng, n, nb; bool train; uint64_t rec, rb; };
    std::vector<Plan> plan(n_leaves);
    int skipped = 0;
    for (int i = 0; i < n_leaves; ++i) {
        const int N = offsets[i + 1] - offsets[i];
        if (N < 0 || offsets[i] < 0) { ctx->err = "offsets must be non-negative and non-decreasing"; return GPIS_ERR_ARG; }
        Plan& pl = plan[i];
        pl = Plan{N, 0, 0, 0, false, 0, 0};
        if (N == 0) continue;   // GPisMap3.cpp:710: nothing in range -> the leaf keeps whatever GP it had
        if (N > GPIS_MAX_SAMPLES) { ++skipped; continue; }
        int ng = 0;
        for (int k = 0; k < N; ++k) ng += grad_valid(samples + (size_t)(offsets[i] + k) * w9, dim) ? 1 : 0;
        pl.ng = ng; pl.n = N + dim * ng; pl.nb = (pl.n + 31) / 32;
        if (pl.n > GPIS_MAX_N) { ++skipped; continue; }
        pl.train = true;
        pl.rb = rec_bytes(N, pl.nb);
    }
    // ---- pass 2: reserve table space and every record; roll back on failure
    int rc = table_reserve(ctx, n_leaves);
    if (rc) return rc;
    for (int i = 0; i < n_leaves; ++i) {
        if (!plan[i].train) continue;
        rc = arena_alloc(ctx, plan[i].rb, &plan[i].rec);
        if (rc) {
            for (int k = 0; k < i; ++k) if (plan[k].train) arena_free(ctx, plan[k].rec, plan[k].rb);
            return rc;
        }
    }
    // ---- pass 3: upload, train, then install (queries never see a half-built record)
    std::vector<TrainJob> jobs;
    std::vector<int> job_leaf;
    std::vector<SlotUpdate> ups;
    std::vector<std::pair<uint64_t, uint64_t>> to_free;
    std::vector<std::pair<uint64_t, int>> installs;   // key, plan index
    double flops = 0, bytes = 0;
    int64_t sumN = 0, sumn = 0;
    int maxN = 1, maxnb = 1;
    for (int i = 0; i < n_leaves; ++i) {
        const Plan& pl = plan[i];
        const uint64_t key = key_of(ctx, cells + (size_t)i * dim);
        auto it = ctx->leaves.find(key);
        if (it == ctx->leaves.end()) {
            HostLeaf hl{};
            hl.slot = take_slot(ctx, key);
            it = ctx->leaves.emplace(key, hl).first;
        }
        HostLeaf& hl = it->second;
        set_geometry(ctx, hl, cells + (size_t)i * dim, centres + (size_t)i * dim);
        if (status) status[i] = (pl.N > 0 && !pl.train) ? (int32_t)GPIS_ERR_CAPACITY : 0;
        ctx->repl_touched.insert(key);
        if (!pl.train) {   // registered (non-empty) but not retrained
            ups.push_back(make_update(key, hl));
            continue;
        }
        TrainJob j{};
        j.rec = pl.rec; j.sample_off = offsets[i]; j.N = pl.N; j.ng = pl.ng; j.n = pl.n; j.nb = pl.nb; j.slot = hl.slot;
        for (int c = 0; c < 3; ++c) { j.cell[c] = hl.cell[c]; j.centre[c] = hl.centre[c]; j.lo[c] = hl.lo[c]; j.hi[c] = hl.hi[c]; }
        jobs.push_back(j);
        job_leaf.push_back(i);
        installs.push_back({key, i});
        flops += (double)pl.n * pl.n * pl.n / 3.0 + 2.0 * pl.n * pl.n + 30.0 * pl.N * pl.N;
        bytes += 52.0 * pl.N + 4.0 * pl.n + 2.0 * pl.n * (pl.n + 1.0);
        sumN += pl.N; sumn += pl.n;
        maxN = std::max(maxN, pl.N); maxnb = std::max(maxnb, pl.nb);
    }
    float ms = 0.f;
    if (!jobs.empty()) {
        const uint64_t nsamp = (uint64_t)offsets[n_leaves];
        rc = ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, align_up(nsamp * w9 * sizeof(float), 256));
        if (rc == 0) {
            cudaError_t e = cudaMemcpyAsync(ctx->d_scratch, samples, nsamp * w9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { ctx->err = std::string("sample upload: ") + cudaGetErrorString(e); rc = GPIS_ERR_CUDA; }
        }
        std::vector<int32_t> st(jobs.size(), 0);
        if (rc == 0) rc = train_jobs(ctx, jobs, (const float*)ctx->d_scratch, maxN, maxnb, st.data(), &ms);
        if (rc) {   // nothing was installed: the leaves keep their previous records
            for (const Plan& pl : plan) if (pl.train) arena_free(ctx, pl.rec, pl.rb);
            apply_updates(ctx, ups);
            return rc;
        }
        if (status)
            for (size_t i = 0; i < jobs.size(); ++i) status[job_leaf[i]] = st[i];
    }
    for (auto& in : installs) {
        HostLeaf& hl = ctx->leaves[in.first];
        const Plan& pl = plan[in.second];
        if (hl.rec) to_free.push_back({hl.rec, hl.rec_bytes});
        hl.rec = pl.rec; hl.rec_bytes = pl.rb; hl.N = pl.N; hl.ng = pl.ng; hl.n = pl.n; hl.nb = pl.nb;
        ups.push_back(make_update(in.first, hl));
        ctx->repl_trained.insert(in.first);
    }
    ctx->max_nb = std::max(ctx->max_nb, maxnb);
    ctx->max_N = std::max(ctx->max_N, maxN);
    rc = apply_updates(ctx, ups);
    if (rc) return rc;
    for (auto& f : to_free) arena_free(ctx, f.first, f.second);
    ctx->st.last_train_leaves = (int64_t)jobs.size();
    ctx->st.last_train_sum_N = sumN; ctx->st.last_train_sum_n = sumn;
    ctx->st.last_train_flops = flops; ctx->st.last_train_bytes = bytes;
    ctx->st.last_train_ms = ms;
    ctx->st.last_train_skipped = skipped;
    if (skipped) ctx->err = "gpis_leaves_update: " + std::to_string(skipped) + " leaf/leaves exceed GPIS_MAX_SAMPLES / GPIS_MAX_N and were not retrained (see status[])";
    return GPIS_OK;
}

// ------------------------------------------------------------------ f-2: device sample store + device-side gather
int gpis_samples_set(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres, const int32_t* offsets,
                     const float* samples) {
    NvtxRange nvtx_("gpis_samples_set");
    if (!ctx || n_leaves < 0) return GPIS_ERR_ARG;
    if (n_leaves == 0) return GPIS_OK;
    if (!cells || !centres || !offsets || (offsets[n_leaves] > 0 && !samples)) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const int dim = ctx->cfg.dim, w9 = 2 * dim + 3;
    for (int i = 0; i < n_leaves; ++i)
        if (offsets[i] < 0 || offsets[i + 1] < offsets[i]) { ctx->err = "offsets must be non-negative and non-decreasing"; return GPIS_ERR_ARG; }
    int rc = table_reserve(ctx, n_leaves);
    if (rc) return rc;
    const uint64_t total = (uint64_t)offsets[n_leaves] * w9 * sizeof(float);
    rc = ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, std::max<uint64_t>(total, 256));
    if (rc) return rc;
    if (total) CK(cudaMemcpyAsync(ctx->d_scratch, samples, total, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<SlotUpdate> ups;
    std::vector<StoreUpdate> sup;
    std::vector<WordCopy> jobs;
    for (int i = 0; i < n_leaves; ++i) {
        const uint64_t key = key_of(ctx, cells + (size_t)i * dim);
        auto it = ctx->leaves.find(key);
        if (it == ctx->leaves.end()) {   // not registered yet: register it like gpis_leaves_mark
            HostLeaf hl{};
            hl.slot = take_slot(ctx, key);
            set_geometry(ctx, hl, cells + (size_t)i * dim, centres + (size_t)i * dim);
            it = ctx->leaves.emplace(key, hl).first;
            ups.push_back(make_update(key, it->second));
            ctx->repl_touched.insert(key);
        }
        HostLeaf& hl = it->second;
        const int N = offsets[i + 1] - offsets[i];
        if (hl.smp) arena_free(ctx, hl.smp, hl.smp_bytes);
        hl.smp = 0; hl.smp_bytes = 0; hl.smp_n = N;
        if (N > 0) {
            hl.smp_bytes = (uint64_t)N * w9 * sizeof(float);
            rc = arena_alloc(ctx, hl.smp_bytes, &hl.smp);
            if (rc) { hl.smp = 0; hl.smp_bytes = 0; hl.smp_n = 0; return rc; }
            jobs.push_back(WordCopy{(const uint32_t*)ctx->d_scratch + (size_t)offsets[i] * w9, (uint32_t*)hl.smp, (uint64_t)N * w9});
        }
        sup.push_back(StoreUpdate{hl.slot, N, hl.smp});
    }
    if (!jobs.empty()) {
        rc = ensure(ctx, &ctx->d_jobs, &ctx->jobs_bytes, jobs.size() * sizeof(WordCopy));
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->d_jobs, jobs.data(), jobs.size() * sizeof(WordCopy), cudaMemcpyHostToDevice, ctx->stream));
        k_copy_words<<<dim3((unsigned)jobs.size(), 1), 256, 0, ctx->stream>>>((const WordCopy*)ctx->d_jobs);
        ctx->st.kernel_launches++;
        CK(cudaGetLastError());
    }
    rc = store_apply(ctx, sup);     // synchronises: the host buffers above may go
    if (rc) return rc;
    return apply_updates(ctx, ups);
}

int gpis_leaves_train_dirty(gpis_ctx* ctx, int n_active, const int32_t* active_cells, float radius, int32_t* n_trained) {
    NvtxRange nvtx_("gpis_leaves_train_dirty");
    if (!ctx || n_active < 0 || !(radius > 0.f)) return GPIS_ERR_ARG;
    if (n_trained) *n_trained = 0;
    CK(cudaSetDevice(ctx->cfg.device));
    {   // the previous batch (asynchronous modes) completes first: its buffers and the arena are about to be reused
        const int rcf = train_flush(ctx);
        if (rcf) return rcf;
    }
    ctx->st.last_train_leaves = 0; ctx->st.last_train_skipped = 0;
    if (ctx->train_mode == 0) ctx->st.last_train_ms = 0.f;   // asynchronous modes: the most recently COMPLETED batch
    if (n_active == 0) return GPIS_OK;
    if (!active_cells) return GPIS_ERR_ARG;
    const int dim = ctx->cfg.dim, w9 = 2 * dim + 3;
    static const bool prof = std::getenv("GPIS_PROFILE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tq[8] = {0}; tq[0] = now();
    {   // the warp-local neighbour list holds GATHER_MAXNBR leaves
        const double per_axis = 2.0 * std::ceil((double)radius * ctx->qp.inv_pitch) + 1.0;
        if (std::pow(per_axis, dim) > GATHER_MAXNBR) { ctx->err = "training radius too large for the device-side gather"; return GPIS_ERR_CAPACITY; }
    }
    const int nslots = ctx->slot_count;
    // scratch: [active int4 x n][flag x nslots][list x nslots][count 16][N x nslots][ng x nslots][off x nslots]
    const uint64_t b_act = align_up(sizeof(int4) * (uint64_t)n_active, 256), b_sl = align_up(sizeof(int32_t) * (uint64_t)std::max(nslots, 1), 256);
    int rc = ensure(ctx, &ctx->d_gather, &ctx->gather_bytes, b_act + 5 * b_sl + 256);
    if (rc) return rc;
    unsigned char* base = (unsigned char*)ctx->d_gather;
    int4* d_act = (int4*)base;
    int32_t* d_flag = (int32_t*)(base + b_act);
    int32_t* d_list = (int32_t*)(base + b_act + b_sl);
    int32_t* d_cnt = (int32_t*)(base + b_act + 2 * b_sl);
    int32_t* d_N = (int32_t*)(base + b_act + 2 * b_sl + 256);
    int32_t* d_ng = (int32_t*)(base + b_act + 3 * b_sl + 256);
    int32_t* d_off = (int32_t*)(base + b_act + 4 * b_sl + 256);
    std::vector<int4> act(n_active);
    for (int i = 0; i < n_active; ++i) act[i] = make_int4(active_cells[(size_t)i * dim], active_cells[(size_t)i * dim + 1], dim == 3 ? active_cells[(size_t)i * dim + 2] : 0, 0);
    CK(cudaMemcpyAsync(d_act, act.data(), sizeof(int4) * (uint64_t)n_active, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_flag, 0, b_sl, ctx->stream));
    CK(cudaMemsetAsync(d_cnt, 0, 256, ctx->stream));
    k_dirty_mark<<<(n_active + 127) / 128, 128, 0, ctx->stream>>>(d_act, n_active, ctx->T, ctx->qp, radius, d_flag, d_list, d_cnt);
    ctx->st.kernel_launches++;
    int32_t ndirty = 0;
    CK(cudaMemcpyAsync(&ndirty, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tq[1] = now();
    if (ndirty <= 0) return GPIS_OK;
    k_ball<0><<<(ndirty + 3) / 4, 128, 0, ctx->stream>>>(d_list, ndirty, ctx->T, ctx->qp, ctx->store, radius, d_N, d_ng, nullptr, nullptr);
    ctx->st.kernel_launches++;
    std::vector<int32_t> list(ndirty), hN(ndirty), hng(ndirty), off(ndirty, 0);
    CK(cudaMemcpyAsync(list.data(), d_list, sizeof(int32_t) * ndirty, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hN.data(), d_N, sizeof(int32_t) * ndirty, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hng.data(), d_ng, sizeof(int32_t) * ndirty, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tq[2] = now();
    // plan: sizes, capacity, record reservation (rolled back on failure); nothing is installed before training succeeded
    using Plan = TrainPlan;
    std::vector<Plan> plan;
    std::vector<int> plan_of(ndirty, -1);
    int skipped = 0;
    int64_t nrows = 0;
    for (int d = 0; d < ndirty; ++d) {
        off[d] = (int32_t)nrows;
        const int N = hN[d];
        if (N <= 0) continue;                                   // GPisMap3.cpp:710: empty ball, the leaf keeps its GP
        const int n = N + dim * hng[d], nb = (n + 31) / 32;
        if (N > GPIS_MAX_SAMPLES || n > GPIS_MAX_N) { ++skipped; continue; }
        Plan pl{ctx->slot_key[list[d]], N, hng[d], n, nb, 0, rec_bytes(N, nb)};
        rc = arena_alloc(ctx, pl.rb, &pl.rec);
        if (rc) { for (auto& q : plan) arena_free(ctx, q.rec, q.rb); return rc; }
        plan_of[d] = (int)plan.size();
        plan.push_back(pl);
        nrows += N;
    }
    if (plan.empty()) { ctx->st.last_train_skipped = skipped; return GPIS_OK; }
    rc = ensure(ctx, &ctx->d_train_smp, &ctx->train_smp_bytes, (uint64_t)nrows * w9 * sizeof(float));
    if (rc) { for (auto& q : plan) arena_free(ctx, q.rec, q.rb); return rc; }
    // gather only the leaves that train: compact the dirty list to those
    std::vector<int32_t> tlist, toff;
    for (int d = 0; d < ndirty; ++d) if (plan_of[d] >= 0) { tlist.push_back(list[d]); toff.push_back(off[d]); }
    CK(cudaMemcpyAsync(d_list, tlist.data(), sizeof(int32_t) * tlist.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off, toff.data(), sizeof(int32_t) * toff.size(), cudaMemcpyHostToDevice, ctx->stream));
    k_ball<1><<<((int)tlist.size() + 3) / 4, 128, 0, ctx->stream>>>(d_list, (int)tlist.size(), ctx->T, ctx->qp, ctx->store, radius, nullptr, nullptr,
                                                                    d_off, (float*)ctx->d_train_smp);
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    std::vector<TrainJob> jobs(plan.size());
    double flops = 0, bytes = 0;
    int64_t sumN = 0, sumn = 0;
    int maxN = 1, maxnb = 1;
    for (size_t i = 0; i < plan.size(); ++i) {
        const Plan& pl = plan[i];
        const HostLeaf& hl = ctx->leaves[pl.key];
        TrainJob& j = jobs[i];
        j = TrainJob{};
        j.rec = pl.rec; j.sample_off = toff[i]; j.N = pl.N; j.ng = pl.ng; j.n = pl.n; j.nb = pl.nb; j.slot = hl.slot;
        for (int c = 0; c < 3; ++c) { j.cell[c] = hl.cell[c]; j.centre[c] = hl.centre[c]; j.lo[c] = hl.lo[c]; j.hi[c] = hl.hi[c]; }
        flops += (double)pl.n * pl.n * pl.n / 3.0 + 2.0 * pl.n * pl.n + 30.0 * pl.N * pl.N;
        bytes += 52.0 * pl.N + 4.0 * pl.n + 2.0 * pl.n * (pl.n + 1.0);
        sumN += pl.N; sumn += pl.n;
        maxN = std::max(maxN, pl.N); maxnb = std::max(maxnb, pl.nb);
    }
    tq[3] = now();
    {   // jobs biggest-first (one CTA per leaf, the hardware scheduler balances the tail), uploaded on the main stream
        std::vector<int> order(jobs.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jobs[a].n > jobs[b].n; });
        std::vector<TrainJob> sorted(jobs.size());
        for (size_t i = 0; i < order.size(); ++i) sorted[i] = jobs[order[i]];
        const uint64_t b_jobs = align_up(sorted.size() * sizeof(TrainJob), 256);
        const uint64_t b_st = align_up(sorted.size() * sizeof(int32_t), 256);
        rc = ensure(ctx, &ctx->d_train_jobs, &ctx->train_jobs_bytes, b_jobs + b_st);
        if (rc) { for (auto& q : plan) arena_free(ctx, q.rec, q.rb); return rc; }
        CK(cudaMemcpyAsync(ctx->d_train_jobs, sorted.data(), sorted.size() * sizeof(TrainJob), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // `sorted`, `tlist`, `toff` go out of scope; the gather has finished too
    }
    ctx->pend = PendingTrain{};
    ctx->pend.active = true;
    ctx->pend.plan = plan;
    ctx->pend.njobs = (int)jobs.size(); ctx->pend.maxN = maxN; ctx->pend.maxnb = maxnb;
    ctx->st.last_train_leaves = (int64_t)jobs.size();
    ctx->st.last_train_sum_N = sumN; ctx->st.last_train_sum_n = sumn;
    ctx->st.last_train_flops = flops; ctx->st.last_train_bytes = bytes;
    ctx->st.last_train_skipped = skipped;
    if (n_trained) *n_trained = (int32_t)jobs.size();
    // mode 0: train and install before returning; 1: K1 starts now on the training stream; 2: it starts at the next
    // gpis_reeval (the next frame's device work goes first, K1 then runs beside that frame's serial host passes).
    // In modes 1 and 2 the records are installed by the next entry point that needs them (train_flush).
    if (ctx->train_mode == 0) rc = train_flush(ctx);
    else if (ctx->train_mode == 1) rc = train_launch(ctx);
    if (rc) return rc;
    tq[4] = now();
    if (prof) std::fprintf(stderr, "train_dirty: mark %.2f count %.2f plan+gather %.2f train(wall) %.2f [kernel %.2f] ms, %d dirty, %zu trained\n",
                           tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[4] - tq[3], ctx->st.last_train_ms, ndirty, jobs.size());
    if (skipped) ctx->err = "gpis_leaves_train_dirty: " + std::to_string(skipped) + " leaf/leaves exceed GPIS_MAX_SAMPLES / GPIS_MAX_N and were not retrained";
    return GPIS_OK;
}

int gpis_leaf_index(gpis_ctx* ctx, const int32_t* cell) {
    if (!ctx || !cell) return -1;
    auto it = ctx->leaves.find(key_of(ctx, cell));
    return it == ctx->leaves.end() ? -1 : it->second.slot;
}

int gpis_leaf_get(gpis_ctx* ctx, const int32_t* cell, int32_t* N, int32_t* ng, float* alpha, float* L,
                  float* gradflag, int cap_n) {
    if (!ctx || !cell) return 0;
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return 0;
    if (train_flush(ctx)) return 0;
    auto it = ctx->leaves.find(key_of(ctx, cell));
    if (it == ctx->leaves.end() || it->second.rec == 0) return 0;
    const HostLeaf& hl = it->second;
    if (N) *N = hl.N;
    if (ng) *ng = hl.ng;
    if (hl.n > cap_n) return hl.n;
    const unsigned char* rec = (const unsigned char*)hl.rec;
    if (alpha) {
        if (cudaMemcpy(alpha, rec + rec_off_alpha(hl.N), sizeof(float) * hl.n, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    }
    if (gradflag) {
        std::vector<float4> pts(hl.N);
        if (cudaMemcpy(pts.data(), rec + rec_off_pts(), sizeof(float4) * hl.N, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
        for (int k = 0; k < hl.N; ++k) { int g; std::memcpy(&g, &pts[k].w, 4); gradflag[k] = g >= 0 ? 1.f : 0.f; }
    }
    if (L) {
        if (ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, sizeof(float) * (uint64_t)hl.n * hl.n)) return 0;
        const int tot = hl.n * hl.n;
        k_unpack_L<<<(tot + 255) / 256, 256, 0, ctx->stream>>>((const float*)(rec + rec_off_tiles(hl.N, hl.nb)), hl.n, hl.nb, (float*)ctx->d_scratch);
        ctx->st.kernel_launches++;
        if (cudaMemcpyAsync(L, ctx->d_scratch, sizeof(float) * (uint64_t)tot, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    }
    return hl.n;
}

// ------------------------------------------------------------------ queries
// h_x / h_res (optional, pinned host memory): the caller's buffers. The chunks of x and res then travel on two copy
// streams beside the evaluation of the neighbouring chunks (H2D of chunk c+1 and D2H of chunk c-1 overlap chunk c).
static int query_core(gpis_ctx* ctx, const float* d_x, int64_t n, float* d_res, int32_t* h_chosen, int32_t* h_tie,
                      const float* h_x = nullptr, float* h_res = nullptr) {
    NvtxRange nvtx_("gpis_query: candidates + eval + fuse");
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // queries see every record trained so far
    const int dim = ctx->cfg.dim;
    const int64_t CH = 1 << 22;  // queries per chunk (bounds scratch: ~160 B/query)
    const int64_t chunk_cap = std::min<int64_t>(n, CH);
    if (ctx->work_cap < chunk_cap) {
        cudaFree(ctx->W.cand); cudaFree(ctx->W.tie); cudaFree(ctx->W.evalout); cudaFree(ctx->W.pairs); cudaFree(ctx->W.counters);
        ctx->W = QueryWork{};
        ctx->work_cap = 0;
        const int64_t cap = chunk_cap + chunk_cap / 8 + 1024;
        CK(cudaMalloc(&ctx->W.cand, sizeof(int4) * cap));
        CK(cudaMalloc(&ctx->W.tie, sizeof(int32_t) * cap));
        CK(cudaMalloc(&ctx->W.evalout, sizeof(float) * 24 * cap));
        CK(cudaMalloc(&ctx->W.pairs, sizeof(int2) * 2 * cap));
        CK(cudaMalloc(&ctx->W.counters, sizeof(int32_t) * 16));
        ctx->work_cap = cap;
    }
    QueryWork W = ctx->W;
    int64_t evals = 0;
    int64_t items[4] = {0, 0, 0, 0};
    float ms_total = 0.f, ms_eval = 0.f;
    if (!ctx->d_acc) CK(cudaMalloc(&ctx->d_acc, sizeof(double) * 4));
    CK(cudaMemsetAsync(ctx->d_acc, 0, sizeof(double) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    const bool staged = h_x != nullptr && h_res != nullptr;
    const int w2 = 2 * (1 + dim);
    auto upload = [&](int64_t q0, int slot) -> int {
        const int64_t nq = std::min<int64_t>(CH, n - q0);
        CK(cudaMemcpyAsync(const_cast<float*>(d_x) + q0 * dim, h_x + q0 * dim, sizeof(float) * dim * nq, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
        CK(cudaMemcpyAsync(d_res + q0 * w2, h_res + q0 * w2, sizeof(float) * w2 * nq, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
        CK(cudaEventRecord(ctx->ev_copy[slot], ctx->copy_stream[0]));
        return 0;
    };
    if (staged) { const int rcu = upload(0, 0); if (rcu) return rcu; }
    int64_t chunk_index = 0;
    for (int64_t q0 = 0; q0 < n; q0 += CH, ++chunk_index) {
        const int64_t nq = std::min<int64_t>(CH, n - q0);
        const float* xq = d_x + q0 * dim;
        float* rq = d_res + q0 * 2 * (1 + dim);
        const int gq = (int)((nq + 255) / 256);
        if (staged) {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[chunk_index & 1], 0));
            if (q0 + CH < n) { const int rcu = upload(q0 + CH, (int)((chunk_index + 1) & 1)); if (rcu) return rcu; }
        }
        CK(cudaMemsetAsync(W.counters, 0, sizeof(int32_t) * 16, ctx->stream));
        k_candidates<<<gq, 256, 0, ctx->stream>>>(xq, nq, rq, ctx->T, ctx->qp, W);
        k_candidates_exact<<<(int)((nq + 127) / 128), 128, 0, ctx->stream>>>(xq, nq, ctx->T, ctx->qp, W);
        ctx->st.kernel_launches += 2;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                CK(cudaMemsetAsync(W.counters, 0, sizeof(int32_t) * 16, ctx->stream));
                k_select2<<<gq, 256, 0, ctx->stream>>>(nq, rq, ctx->T, ctx->qp, W, 0);
                ctx->st.kernel_launches++;
            }
            int32_t npairs = 0;
            CK(cudaMemcpyAsync(&npairs, W.counters, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (npairs > 0) {
                CK(cudaEventRecord(ctx->ev[2], ctx->stream));
                {
                    int rc = query_eval(ctx->stream, xq, ctx->T, ctx->qp, W, npairs, ctx->slot_count, ctx->max_nb,
                                        &ctx->d_sort, &ctx->sort_cap, &ctx->st.kernel_launches, ctx->err, ctx->d_acc,
                                        ctx->eval_version, ctx->prog, items);
                    if (rc) return rc;
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(ctx->ev[3], ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
                ms_eval += ms;
                evals += npairs;
            }
        }
        k_fuse<<<gq, 256, 0, ctx->stream>>>(nq, rq, ctx->T, ctx->qp, W);
        ctx->st.kernel_launches++;
        CK(cudaGetLastError());
        if (h_chosen) CK(cudaMemcpyAsync(h_chosen + q0 * 4, W.cand, sizeof(int4) * nq, cudaMemcpyDeviceToHost, ctx->stream));
        if (h_tie) CK(cudaMemcpyAsync(h_tie + q0, W.tie, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, ctx->stream));
        if (h_chosen || h_tie) CK(cudaStreamSynchronize(ctx->stream));
        if (staged) {   // this chunk's results go home while the next chunk is evaluated
            CK(cudaEventRecord(ctx->ev_copy[2 + (chunk_index & 1)], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream[1], ctx->ev_copy[2 + (chunk_index & 1)], 0));
            CK(cudaMemcpyAsync(h_res + q0 * w2, rq, sizeof(float) * w2 * nq, cudaMemcpyDeviceToHost, ctx->copy_stream[1]));
        }
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (staged) CK(cudaStreamSynchronize(ctx->copy_stream[1]));
    CK(cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[1]));
    ctx->st.last_query_n = n;
    ctx->st.last_query_evals = evals;
    ctx->st.last_query_ms = ms_total;
    ctx->st.last_query_eval_ms = ms_eval;
    for (int i = 0; i < 4; ++i) ctx->st.last_query_items[i] = items[i];
    double acc[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(acc, ctx->d_acc, sizeof(double) * 3, cudaMemcpyDeviceToHost));
    ctx->st.last_query_flops = acc[0];
    ctx->st.last_query_bytes_gather = 44.0 * (double)n + acc[1];
    // compulsory bytes: every leaf record touched counts once per pass over the batch (two passes and
    // query chunks may touch a leaf again; that is real re-reading and is charged)
    ctx->st.last_query_bytes_compulsory = 44.0 * (double)n + acc[2];
    return GPIS_OK;
}

int gpis_query_device(gpis_ctx* ctx, const float* x_device, int64_t n, float* res_inout_device) {
    if (!ctx || !x_device || !res_inout_device || n < 1) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    return query_core(ctx, x_device, n, res_inout_device, nullptr, nullptr);
}

static int query_host(gpis_ctx* ctx, const float* x, int64_t n, float* res, int32_t* chosen, int32_t* tie) {
    NvtxRange nvtx_("gpis_query (host buffers)");
    if (!ctx || !x || !res || n < 1) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    const int dim = ctx->cfg.dim, w2 = 2 * (1 + dim);
    if (ctx->q_cap < n) {
        cudaFree(ctx->d_x); cudaFree(ctx->d_res);
        ctx->d_x = ctx->d_res = nullptr; ctx->q_cap = 0;
        CK(cudaMalloc(&ctx->d_x, sizeof(float) * dim * n));
        CK(cudaMalloc(&ctx->d_res, sizeof(float) * w2 * n));
        ctx->q_cap = n;
    }
    // pinned (or registered) caller buffers: copies chunk by chunk beside the evaluation; pageable memory: cudaMemcpyAsync
    // would block the host thread inside the chunk loop, so everything goes up first and comes back at the end
    auto pinned = [](const void* p) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    if (pinned(x) && pinned(res)) return query_core(ctx, (const float*)ctx->d_x, n, (float*)ctx->d_res, chosen, tie, x, res);
    CK(cudaMemcpyAsync(ctx->d_x, x, sizeof(float) * dim * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_res, res, sizeof(float) * w2 * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = query_core(ctx, (const float*)ctx->d_x, n, (float*)ctx->d_res, chosen, tie);
    if (rc) return rc;
    CK(cudaMemcpyAsync(res, ctx->d_res, sizeof(float) * w2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GPIS_OK;
}
int gpis_query(gpis_ctx* ctx, const float* x, int64_t n, float* res_inout) { return query_host(ctx, x, n, res_inout, nullptr, nullptr); }
int gpis_query_debug(gpis_ctx* ctx, const float* x, int64_t n, float* res_inout, int32_t* chosen4, int32_t* tie) {
    return query_host(ctx, x, n, res_inout, chosen4, tie);
}

// ------------------------------------------------------------------ observation GPs
static int obs_upload_partition(gpis_ctx* ctx) {
    const int nt = (int)ctx->obs_hdesc.size();
    if (nt > ctx->obs_tile_cap) {
        cudaFree(ctx->obs_tiles); cudaFree(ctx->obs_desc);
        ctx->obs_tiles = nullptr; ctx->obs_desc = nullptr; ctx->obs_tile_cap = 0;
        CK(cudaMalloc(&ctx->obs_tiles, sizeof(ObsTile) * nt));
        CK(cudaMalloc(&ctx->obs_desc, sizeof(ObsTileDesc) * nt));
        ctx->obs_tile_cap = nt;
    }
    const int nbv = (int)std::max(ctx->obs_hb0.size(), ctx->obs_hb1.size());
    if (nbv > ctx->obs_b_cap) {
        cudaFree(ctx->obs_b0); cudaFree(ctx->obs_b1);
        ctx->obs_b0 = ctx->obs_b1 = nullptr; ctx->obs_b_cap = 0;
        CK(cudaMalloc(&ctx->obs_b0, sizeof(float) * nbv));
        CK(cudaMalloc(&ctx->obs_b1, sizeof(float) * nbv));
        ctx->obs_b_cap = nbv;
    }
    if (nt) CK(cudaMemcpyAsync(ctx->obs_desc, ctx->obs_hdesc.data(), sizeof(ObsTileDesc) * nt, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->obs_hb0.empty()) CK(cudaMemcpyAsync(ctx->obs_b0, ctx->obs_hb0.data(), sizeof(float) * ctx->obs_hb0.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->obs_hb1.empty()) CK(cudaMemcpyAsync(ctx->obs_b1, ctx->obs_hb1.data(), sizeof(float) * ctx->obs_hb1.size(), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// computePartition (ObsGP.cpp:204-265) from the host copy of the pixel grid; recomputed only when the grid size
// changes or after a reset, like the reference (ObsGP.cpp:335-337).
static int obs_partition_2d(gpis_ctx* ctx, const float* vu, int ni, int nj) {
    const int g = 5, ov = 3;                                   // params.h:109-110
    ctx->op.d = 2; ctx->op.ni = ni; ctx->op.margin = 0.005f;   // params.h:108
    if (ctx->obs_ni != ni || ctx->obs_nj != nj || ctx->obs_repartition) {
        const int ng0 = (ni - ov) / g + 1, ng1 = (nj - ov) / g + 1;
        if (ng0 < 1 || ng1 < 1) { ctx->err = "grid too small for the ObsGP partition"; return GPIS_ERR_ARG; }
        std::vector<int> i0(ng0), i1(ng0), j0(ng1), j1(ng1);
        ctx->obs_hb0.assign(1, vu[0]);
        for (int n = 0; n < ng0; ++n) {
            i0[n] = n * g; i1[n] = i0[n] + g + ov - 1;
            if (n < ng0 - 1) ctx->obs_hb0.push_back(vu[2 * (i1[n] - ov / 2)]);
            else { i1[n] = ni - 1; ctx->obs_hb0.push_back(vu[2 * i1[n]]); }
        }
        ctx->obs_hb1.assign(1, vu[1]);
        for (int m = 0; m < ng1; ++m) {
            j0[m] = m * g; j1[m] = j0[m] + g + ov - 1;
            if (m < ng1 - 1) ctx->obs_hb1.push_back(vu[2 * (size_t)(j1[m] - ov / 2) * ni + 1]);
            else { j1[m] = nj - 1; ctx->obs_hb1.push_back(vu[2 * (size_t)j1[m] * ni + 1]); }
        }
        ctx->obs_hdesc.clear();
        for (int m = 0; m < ng1; ++m)
            for (int n = 0; n < ng0; ++n) ctx->obs_hdesc.push_back(ObsTileDesc{i0[n], i1[n], j0[m], j1[m]});
        ctx->op.ng0 = ng0; ctx->op.ntiles = ng0 * ng1;
        ctx->op.nb0 = (int)ctx->obs_hb0.size(); ctx->op.nb1 = (int)ctx->obs_hb1.size();
        ctx->obs_ni = ni; ctx->obs_nj = nj; ctx->obs_repartition = false;
        return obs_upload_partition(ctx);
    }
    return 0;
}

int gpis_obs_train_2d(gpis_ctx* ctx, const float* vu, const float* zinv, int ni, int nj) {
    NvtxRange nvtx_("gpis_obs_train_2d");
    if (!ctx) return GPIS_ERR_ARG;
    if (!vu || !zinv || ni <= 0 || nj <= 0) return GPIS_OK;   // ObsGP.cpp:333: silently untrained
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = obs_partition_2d(ctx, vu, ni, nj);
    if (rc) return rc;
    const uint64_t bx = align_up(sizeof(float) * 2 * (uint64_t)ni * nj, 256), bf = align_up(sizeof(float) * (uint64_t)ni * nj, 256);
    rc = ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, bx + bf);
    if (rc) return rc;
    float* d_vu = (float*)ctx->d_scratch;
    float* d_f = (float*)((unsigned char*)ctx->d_scratch + bx);
    CK(cudaMemcpyAsync(d_vu, vu, sizeof(float) * 2 * (uint64_t)ni * nj, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_f, zinv, sizeof(float) * (uint64_t)ni * nj, cudaMemcpyHostToDevice, ctx->stream));
    k_obs_train<<<ctx->op.ntiles, OBS_MAXP, 0, ctx->stream>>>(d_vu, d_f, ctx->obs_desc, ctx->obs_tiles, ctx->op);
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->obs_trained = true;
    return GPIS_OK;
}

int gpis_obs_train_1d(gpis_ctx* ctx, const float* theta, const float* f, int N) {
    NvtxRange nvtx_("gpis_obs_train_1d");
    if (!ctx) return GPIS_ERR_ARG;
    if (!theta || !f || N <= 0) return GPIS_OK;                // ObsGP.cpp:89
    CK(cudaSetDevice(ctx->cfg.device));
    const int g = 20, ov = 6;                                  // params.h:101-102
    ctx->op.d = 1; ctx->op.ni = N; ctx->op.margin = 0.0175f;   // params.h:103
    // group ranges, ObsGP.cpp:91-137 (recomputed every frame: ObsGP1D::reset clears `range`)
    const int nGroup = N / g + 1;
    ctx->obs_hb0.assign(1, theta[0]);
    ctx->obs_hb1.clear();
    ctx->obs_hdesc.clear();
    for (int n = 0; n < nGroup - 1; ++n) {
        if (n < nGroup - 2) {
            const int a = n * g, b = a + g + ov;
            ctx->obs_hb0.push_back(theta[b - ov / 2]);
            ctx->obs_hdesc.push_back(ObsTileDesc{a, a + g + ov - 1, 0, 0});
        } else {
            int a = n * g;
            int b = a + (N - a) / 2 + ov;
            ctx->obs_hb0.push_back(theta[b - ov / 2]);
            ctx->obs_hdesc.push_back(ObsTileDesc{a, b, 0, 0});
            ++n;
            a = a + (N - a) / 2;
            b = N - 1;
            ctx->obs_hb0.push_back(theta[b]);
            ctx->obs_hdesc.push_back(ObsTileDesc{a, b, 0, 0});
        }
    }
    ctx->op.ng0 = (int)ctx->obs_hdesc.size(); ctx->op.ntiles = (int)ctx->obs_hdesc.size();
    ctx->op.nb0 = (int)ctx->obs_hb0.size(); ctx->op.nb1 = 0;
    ctx->obs_ni = -1; ctx->obs_nj = -1; ctx->obs_repartition = true;
    int rc = obs_upload_partition(ctx);
    if (rc) return rc;
    const uint64_t bx = align_up(sizeof(float) * (uint64_t)N, 256);
    rc = ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, 2 * bx);
    if (rc) return rc;
    float* d_t = (float*)ctx->d_scratch;
    float* d_f = (float*)((unsigned char*)ctx->d_scratch + bx);
    CK(cudaMemcpyAsync(d_t, theta, sizeof(float) * N, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_f, f, sizeof(float) * N, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->op.ntiles > 0) {
        k_obs_train<<<ctx->op.ntiles, OBS_MAXP, 0, ctx->stream>>>(d_t, d_f, ctx->obs_desc, ctx->obs_tiles, ctx->op);
        ctx->st.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->obs_trained = true;
    return GPIS_OK;
}

// Batched ObsGP*::test on device-resident points: bucket by tile, then one CTA per (tile, chunk).
static int obs_test_device(gpis_ctx* ctx, const float* d_x, int m, float* d_val, float* d_var, int32_t* d_tile,
                           int32_t* d_order, int32_t* d_count, int32_t* d_start, int32_t* d_cursor) {
    const int nt = ctx->op.ntiles;
    if (nt < 1 || m < 1) return 0;
    CK(cudaMemsetAsync(d_count, 0, sizeof(int32_t) * nt, ctx->stream));
    k_obs_locate<<<(m + 255) / 256, 256, 0, ctx->stream>>>(d_x, m, ctx->obs_b0, ctx->obs_b1, ctx->obs_tiles, ctx->op, d_tile, d_count, d_var);
    k_obs_scan<<<1, 1024, 0, ctx->stream>>>(d_count, nt, d_start, d_cursor);
    k_obs_scatter<<<(m + 255) / 256, 256, 0, ctx->stream>>>(d_tile, m, d_cursor, d_order);
    k_obs_test_grouped<<<dim3(nt, OBS_TEST_CHUNKS), OBS_TEST_THREADS, 0, ctx->stream>>>(d_x, d_start, d_order, ctx->obs_tiles, ctx->op, d_val, d_var);
    ctx->st.kernel_launches += 4;
    CK(cudaGetLastError());
    return 0;
}

int gpis_obs_test(gpis_ctx* ctx, const float* xt, int d, int m, float* val, float* var) {
    NvtxRange nvtx_("gpis_obs_test");
    if (!ctx || !xt || !val || !var || m < 0) return GPIS_ERR_ARG;
    if (m == 0) return GPIS_OK;
    if (!ctx->obs_trained) return GPIS_OK;                    // ObsGP.cpp:147-149, 412-414: outputs untouched
    if (d != ctx->op.d) return GPIS_OK;                       // ObsGP.cpp:412: wrong input dimension
    CK(cudaSetDevice(ctx->cfg.device));
    const uint64_t bx = align_up(sizeof(float) * (uint64_t)d * m, 256), bv = align_up(sizeof(float) * (uint64_t)m, 256);
    const int nt = ctx->op.ntiles;
    const uint64_t bt = align_up(sizeof(int32_t) * (uint64_t)(nt + 1), 256);
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->scratch_bytes, bx + 4 * bv + 3 * bt);
    if (rc) return rc;
    unsigned char* base = (unsigned char*)ctx->d_scratch;
    float* d_x = (float*)base;
    float* d_val = (float*)(base + bx);
    float* d_var = (float*)(base + bx + bv);
    int32_t* d_tile = (int32_t*)(base + bx + 2 * bv);
    int32_t* d_order = (int32_t*)(base + bx + 3 * bv);
    int32_t* d_count = (int32_t*)(base + bx + 4 * bv);
    int32_t* d_start = (int32_t*)(base + bx + 4 * bv + bt);
    int32_t* d_cursor = (int32_t*)(base + bx + 4 * bv + 2 * bt);
    CK(cudaMemcpyAsync(d_x, xt, sizeof(float) * (uint64_t)d * m, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_val, val, sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream));
    rc = obs_test_device(ctx, d_x, m, d_val, d_var, d_tile, d_order, d_count, d_start, d_cursor);
    if (rc) return rc;
    CK(cudaMemcpyAsync(val, d_val, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(var, d_var, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GPIS_OK;
}

// ------------------------------------------------------------------ f-3 + f-1: one depth frame on the device
int gpis_frame_eval(gpis_ctx* ctx, const float* depth, int N, const float* vu_grid, const gpis_frame_params* fp,
                    int32_t* n_valid, float* range_obs_max, int32_t cap, float* xyz_global, int32_t* status,
                    float* grad, float* noise, float* grad_noise) {
    NvtxRange nvtx_("gpis_frame_eval");
    if (!ctx || !depth || !vu_grid || !fp || !n_valid || !range_obs_max || N < 1) return GPIS_ERR_ARG;
    if (ctx->cfg.dim != 3 || fp->skip < 1) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    FrameParams P{};
    P.width = fp->width; P.height = fp->height; P.skip = fp->skip;
    P.n = fp->width / fp->skip; P.m = fp->height / fp->skip; P.N = N;
    for (int i = 0; i < 3; ++i) P.t[i] = fp->pose[i];
    for (int i = 0; i < 9; ++i) P.R[i] = fp->pose[3 + i];
    P.delx = fp->delx; P.obs_var_thre = fp->obs_var_thre;
    P.min_position_noise = fp->min_position_noise; P.min_grad_noise = fp->min_grad_noise;
    P.max_range = fp->max_range; P.min_range = fp->min_range;
    const int G = P.n * P.m;
    if (G < 1) return GPIS_ERR_ARG;
    *n_valid = 0; *range_obs_max = 0.f;
    int rc = obs_partition_2d(ctx, vu_grid, P.m, P.n);          // ni = rows (fast), nj = columns (GPisMap3.cpp:247-249)
    if (rc) return rc;
    // device layout (floats unless noted), sized for K = G
    const uint64_t a = 256;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t r = o; o += align_up(bytes, a); return r; };
    const uint64_t o_depth = take(4ull * N), o_vu = take(8ull * G), o_zinv = take(4ull * G), o_flag = take(4ull * G), o_start = take(4ull * (G + 1)),
                   o_cur = take(4ull * G), o_vuv = take(8ull * G), o_xl = take(12ull * G), o_xg = take(12ull * G), o_val1 = take(4ull * G),
                   o_var1 = take(4ull * G), o_vup = take(48ull * G), o_valp = take(24ull * G), o_varp = take(24ull * G), o_st = take(4ull * G),
                   o_grad = take(12ull * G), o_noise = take(4ull * G), o_gn = take(4ull * G), o_tile = take(24ull * G), o_order = take(24ull * G),
                   o_tc = take(4ull * (ctx->op.ntiles + 4096)), o_ts = take(4ull * (ctx->op.ntiles + 4096)), o_tcur = take(4ull * (ctx->op.ntiles + 4096)),
                   o_rmax = take(256);
    rc = ensure(ctx, &ctx->d_frame, &ctx->frame_bytes, o);
    if (rc) return rc;
    unsigned char* B = (unsigned char*)ctx->d_frame;
    float* d_depth = (float*)(B + o_depth); float* d_vu = (float*)(B + o_vu); float* d_zinv = (float*)(B + o_zinv);
    int32_t* d_flag = (int32_t*)(B + o_flag); int32_t* d_start = (int32_t*)(B + o_start); int32_t* d_cur = (int32_t*)(B + o_cur);
    float* d_vuv = (float*)(B + o_vuv); float* d_xl = (float*)(B + o_xl); float* d_xg = (float*)(B + o_xg);
    float* d_val1 = (float*)(B + o_val1); float* d_var1 = (float*)(B + o_var1); float* d_vup = (float*)(B + o_vup);
    float* d_valp = (float*)(B + o_valp); float* d_varp = (float*)(B + o_varp); int32_t* d_st = (int32_t*)(B + o_st);
    float* d_grad = (float*)(B + o_grad); float* d_noise = (float*)(B + o_noise); float* d_gn = (float*)(B + o_gn);
    int32_t* d_tile = (int32_t*)(B + o_tile); int32_t* d_order = (int32_t*)(B + o_order);
    int32_t* d_tc = (int32_t*)(B + o_tc); int32_t* d_ts = (int32_t*)(B + o_ts); int32_t* d_tcur = (int32_t*)(B + o_tcur);
    int32_t* d_rmax = (int32_t*)(B + o_rmax);
    CK(cudaMemcpyAsync(d_depth, depth, 4ull * N, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_vu, vu_grid, 8ull * G, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_rmax, 0, 256, ctx->stream));
    const int gb = (G + 255) / 256;
    k_frame_valid<<<gb, 256, 0, ctx->stream>>>(d_depth, P, d_zinv, d_flag);
    k_obs_scan<<<1, 1024, 0, ctx->stream>>>(d_flag, G, d_start, d_cur);
    k_frame_project<<<gb, 256, 0, ctx->stream>>>(d_depth, d_vu, P, d_flag, d_start, d_vuv, d_xl, d_xg, d_rmax);
    ctx->st.kernel_launches += 3;
    int32_t hk[2] = {0, 0};
    CK(cudaMemcpyAsync(&hk[0], d_start + G, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&hk[1], d_rmax, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int K = hk[0];
    *n_valid = K;
    std::memcpy(range_obs_max, &hk[1], 4);
    if (K <= 1) return GPIS_OK;                                  // GPisMap3.cpp:212-215: nothing to regress
    if (K > cap || !xyz_global || !status || !grad || !noise || !grad_noise) { ctx->err = "gpis_frame_eval: output capacity"; return GPIS_ERR_ARG; }
    // regressObs (GPisMap3.cpp:239-256): ObsGP2D::train on the device-resident grid
    k_obs_train<<<ctx->op.ntiles, OBS_MAXP, 0, ctx->stream>>>(d_vu, d_zinv, ctx->obs_desc, ctx->obs_tiles, ctx->op);
    ctx->st.kernel_launches++;
    ctx->obs_trained = true;
    // evalPoints: centre tests, probes, numerics
    rc = obs_test_device(ctx, d_vuv, K, d_val1, d_var1, d_tile, d_order, d_tc, d_ts, d_tcur);
    if (rc) return rc;
    k_frame_probes<<<(6 * K + 255) / 256, 256, 0, ctx->stream>>>(d_xl, K, P.delx, d_vup);
    ctx->st.kernel_launches++;
    rc = obs_test_device(ctx, d_vup, 6 * K, d_valp, d_varp, d_tile, d_order, d_tc, d_ts, d_tcur);
    if (rc) return rc;
    k_frame_numerics<<<(K + 255) / 256, 256, 0, ctx->stream>>>(d_xl, K, P, d_var1, d_valp, d_varp, d_st, d_grad, d_noise, d_gn);
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(xyz_global, d_xg, 12ull * K, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(status, d_st, 4ull * K, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(grad, d_grad, 12ull * K, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(noise, d_noise, 4ull * K, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(grad_noise, d_gn, 4ull * K, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GPIS_OK;
}

int gpis_reeval(gpis_ctx* ctx, int n, const float* samples8, const gpis_frame_params* fp, float map_noise_param,
                int32_t* action, float* pos_new, float* grad_new, float* noise, float* grad_noise) {
    NvtxRange nvtx_("gpis_reeval");
    if (!ctx || n < 0 || !fp) return GPIS_ERR_ARG;
    if (n == 0) return GPIS_OK;
    if (!samples8 || !action || !pos_new || !grad_new || !noise || !grad_noise || ctx->cfg.dim != 3) return GPIS_ERR_ARG;
    if (!ctx->obs_trained || ctx->op.d != 2) { for (int i = 0; i < n; ++i) action[i] = -1; return GPIS_OK; }
    CK(cudaSetDevice(ctx->cfg.device));
    ReevalParams P{};
    for (int i = 0; i < 3; ++i) P.t[i] = fp->pose[i];
    for (int i = 0; i < 9; ++i) P.R[i] = fp->pose[3 + i];
    P.delx = fp->delx; P.obs_var_thre = fp->obs_var_thre; P.min_position_noise = fp->min_position_noise;
    P.min_grad_noise = fp->min_grad_noise; P.map_noise_param = map_noise_param;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t r = o; o += align_up(bytes, 256); return r; };
    const uint64_t N = (uint64_t)n;
    const uint64_t o_smp = take(32 * N), o_loc = take(12 * N), o_vu = take(8 * N), o_val = take(4 * N), o_var = take(4 * N), o_alive = take(4 * N),
                   o_xn = take(12 * N), o_abs = take(4 * N), o_vup = take(48 * N), o_valp = take(24 * N), o_varp = take(24 * N), o_act = take(4 * N),
                   o_pos = take(12 * N), o_grad = take(12 * N), o_noise = take(4 * N), o_gn = take(4 * N), o_tile = take(24 * N), o_order = take(24 * N),
                   o_tc = take(4ull * (ctx->op.ntiles + 1)), o_ts = take(4ull * (ctx->op.ntiles + 1)), o_tcur = take(4ull * (ctx->op.ntiles + 1));
    int rc = ensure(ctx, &ctx->d_reeval, &ctx->reeval_bytes, o);
    if (rc) return rc;
    unsigned char* B = (unsigned char*)ctx->d_reeval;
    float* d_smp = (float*)(B + o_smp); float* d_loc = (float*)(B + o_loc); float* d_vu = (float*)(B + o_vu);
    float* d_val = (float*)(B + o_val); float* d_var = (float*)(B + o_var); int32_t* d_alive = (int32_t*)(B + o_alive);
    float* d_xn = (float*)(B + o_xn); float* d_abs = (float*)(B + o_abs); float* d_vup = (float*)(B + o_vup);
    float* d_valp = (float*)(B + o_valp); float* d_varp = (float*)(B + o_varp); int32_t* d_act = (int32_t*)(B + o_act);
    float* d_pos = (float*)(B + o_pos); float* d_grad = (float*)(B + o_grad); float* d_noise = (float*)(B + o_noise); float* d_gn = (float*)(B + o_gn);
    int32_t* d_tile = (int32_t*)(B + o_tile); int32_t* d_order = (int32_t*)(B + o_order);
    int32_t* d_tc = (int32_t*)(B + o_tc); int32_t* d_ts = (int32_t*)(B + o_ts); int32_t* d_tcur = (int32_t*)(B + o_tcur);
    CK(cudaMemcpyAsync(d_smp, samples8, 32 * N, cudaMemcpyHostToDevice, ctx->stream));
    const int gb = (n + 255) / 256;
    k_reeval_project<<<gb, 256, 0, ctx->stream>>>(d_smp, n, P, d_loc, d_vu);
    ctx->st.kernel_launches++;
    rc = obs_test_device(ctx, d_vu, n, d_val, d_var, d_tile, d_order, d_tc, d_ts, d_tcur);
    if (rc) return rc;
    k_reeval_walk<<<gb, 256, 0, ctx->stream>>>(d_smp, n, P, d_loc, d_val, d_var, d_alive, d_xn, d_abs, d_vup);
    ctx->st.kernel_launches++;
    rc = obs_test_device(ctx, d_vup, 6 * n, d_valp, d_varp, d_tile, d_order, d_tc, d_ts, d_tcur);
    if (rc) return rc;
    k_reeval_numerics<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_smp, n, P, d_loc, d_alive, d_xn, d_abs, d_valp, d_varp, d_act, d_pos, d_grad, d_noise, d_gn);
    ctx->st.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(action, d_act, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pos_new, d_pos, 12 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(grad_new, d_grad, 12 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(noise, d_noise, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(grad_noise, d_gn, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // train mode 2: the previous frame's K1 starts here, beside this frame's serial host passes
    return ctx->train_mode == 2 ? train_launch(ctx) : GPIS_OK;
}

// ------------------------------------------------------------------ K5 inside the library: NCCL replication
// One message per call: a 64-byte header, an index of table entries (every leaf whose entry or record changed on the
// root since the last call, and the erased keys), then the new records, packed by one kernel into a bounded staging
// buffer and broadcast chunk by chunk. Receivers allocate arena space from the index, scatter each chunk with one
// kernel and apply all table changes at the end; nothing synchronises per record.
struct ReplHeader { uint64_t magic, n_entries, payload_bytes; int32_t root_min[3]; int32_t levels; uint64_t pad[3]; };
static_assert(sizeof(ReplHeader) == 64, "header is 64 bytes");
struct ReplEntry {
    uint64_t key, bytes, offset;      // bytes = 0: table entry only (mark / box change); offset into the record stream
    int32_t cell[3]; int32_t erase;   // erase = 1: the leaf disappeared
    float centre[3]; int32_t box_set;
    float lo[3], hi[3];
    int32_t meta[4];                  // N, ng, n, nb of the record that travels (bytes > 0)
    int32_t pad[2];
};
#define GPIS_REPL_CHUNK (256ull << 20)

#define NCK(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + (ctx->p_ncclGetErrorString ? ctx->p_ncclGetErrorString(r_) : "nccl error"); \
            return GPIS_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

static void* nccl_open(std::string& err) {
    const char* env = std::getenv("GPIS_NCCL_LIB");
    // dlopen by soname returns the copy the process already holds (e.g. the one torch loaded), else the system one
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        if (void* h = dlopen(n, RTLD_NOW | RTLD_LOCAL)) return h;
    }
    err = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed");
    return nullptr;
}

int gpis_comm_unique_id(void* id128) {
    if (!id128) return GPIS_ERR_ARG;
    std::string err;
    void* h = nccl_open(err);
    if (!h) return GPIS_ERR_STATE;
    auto fn = (ncclResult_t(*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    if (!fn) return GPIS_ERR_STATE;
    ncclUniqueId id;
    if (fn(&id) != ncclSuccess) return GPIS_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return GPIS_OK;
}

int gpis_comm_init(gpis_ctx* ctx, int rank, int world, const void* id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    if (!ctx->nccl_lib) ctx->nccl_lib = nccl_open(ctx->err);
    if (!ctx->nccl_lib) return GPIS_ERR_STATE;
    ctx->p_ncclCommInitRank = (decltype(ctx->p_ncclCommInitRank))dlsym(ctx->nccl_lib, "ncclCommInitRank");
    ctx->p_ncclBroadcast = (decltype(ctx->p_ncclBroadcast))dlsym(ctx->nccl_lib, "ncclBroadcast");
    ctx->p_ncclCommDestroy = (decltype(ctx->p_ncclCommDestroy))dlsym(ctx->nccl_lib, "ncclCommDestroy");
    ctx->p_ncclGetErrorString = (decltype(ctx->p_ncclGetErrorString))dlsym(ctx->nccl_lib, "ncclGetErrorString");
    if (!ctx->p_ncclCommInitRank || !ctx->p_ncclBroadcast || !ctx->p_ncclCommDestroy) { ctx->err = "NCCL symbols missing"; return GPIS_ERR_STATE; }
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    if (ctx->comm) { ctx->p_ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
    NCK(ctx->p_ncclCommInitRank(&ctx->comm, world, id, rank));
    ctx->comm_rank = rank; ctx->comm_world = world;
    return GPIS_OK;
}

static int erase_one(gpis_ctx* ctx, uint64_t key, std::vector<SlotUpdate>& ups) {
    auto it = ctx->leaves.find(key);
    if (it == ctx->leaves.end()) return 0;
    SlotUpdate u{};
    u.key = key; u.slot = it->second.slot; u.live = 0;
    ups.push_back(u);
    if (it->second.rec) arena_free(ctx, it->second.rec, it->second.rec_bytes);
    if (it->second.smp) arena_free(ctx, it->second.smp, it->second.smp_bytes);   // replicas hold no samples; kept for symmetry
    ctx->free_slots.push_back(it->second.slot);
    ctx->leaves.erase(it);
    ctx->tombs++;
    return 1;
}

}  // extern "C" (templates below)

// The sender's half of a message: index of table entries (+ record sizes / stream offsets) and the header.
// full = every leaf of the table (snapshot); otherwise what changed since the last gpis_replicate.
static void build_message(gpis_ctx* ctx, bool full, ReplHeader& hd, std::vector<ReplEntry>& idx) {
    idx.clear();
    hd = ReplHeader{};
    if (!full)
        for (uint64_t k : ctx->repl_erased) { ReplEntry e{}; e.key = k; e.erase = 1; idx.push_back(e); }
    uint64_t off = 0;
    auto add = [&](uint64_t k, const HostLeaf& hl, bool with_record) {
        ReplEntry e{};
        e.key = k;
        for (int c = 0; c < 3; ++c) { e.cell[c] = hl.cell[c]; e.centre[c] = hl.centre[c]; e.lo[c] = hl.lo[c]; e.hi[c] = hl.hi[c]; }
        e.box_set = hl.box_set ? 1 : 0;
        if (with_record && hl.rec) {
            e.bytes = hl.rec_bytes; e.offset = off; off += align_up(hl.rec_bytes, 256);
            e.meta[0] = hl.N; e.meta[1] = hl.ng; e.meta[2] = hl.n; e.meta[3] = hl.nb;
        }
        idx.push_back(e);
    };
    if (full) {
        std::vector<uint64_t> keys;
        for (auto& kv : ctx->leaves) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());   // deterministic files
        for (uint64_t k : keys) add(k, ctx->leaves[k], true);
    } else {
        for (uint64_t k : ctx->repl_touched) {
            auto it = ctx->leaves.find(k);
            if (it != ctx->leaves.end()) add(k, it->second, ctx->repl_trained.count(k) != 0);
        }
    }
    hd.magic = 0x4750495352455031ull; hd.n_entries = idx.size(); hd.payload_bytes = off;
    for (int c = 0; c < 3; ++c) hd.root_min[c] = ctx->qp.root_min[c];
    hd.levels = ctx->qp.levels;
}

// Moves the record stream of a message through the bounded staging buffer, chunk by chunk. On the sending side the
// chunk is gathered from the arena by one kernel and handed to `xfer` (NCCL broadcast, or a copy to a file); on the
// receiving side `xfer` fills the staging buffer and one kernel scatters it to the freshly reserved records.
template <class Xfer>
static int stream_records(gpis_ctx* ctx, bool sender, const ReplHeader& hd, const std::vector<ReplEntry>& idx,
                          const std::vector<uint64_t>& new_rec, Xfer&& xfer) {
    if (!hd.payload_bytes) return 0;
    uint64_t biggest = 0;
    for (const ReplEntry& e : idx) biggest = std::max<uint64_t>(biggest, align_up(e.bytes, 256));
    const uint64_t chunk_cap = std::max<uint64_t>(GPIS_REPL_CHUNK, biggest);
    int rc = ensure(ctx, &ctx->d_repl, &ctx->repl_cap, std::min<uint64_t>(chunk_cap, hd.payload_bytes));
    if (rc) return rc;
    std::vector<CopyJob> jobs;
    size_t i = 0;
    while (i < idx.size()) {
        jobs.clear();
        uint64_t base = 0, used = 0;
        bool have = false;
        size_t j = i;
        for (; j < idx.size(); ++j) {   // maximal run of travelling records that fits the staging buffer
            const ReplEntry& e = idx[j];
            if (e.erase || !e.bytes) continue;
            const uint64_t sz = align_up(e.bytes, 256);
            if (!have) { base = e.offset; have = true; }
            if (used + sz > chunk_cap && used > 0) break;
            unsigned char* stage = (unsigned char*)ctx->d_repl + (e.offset - base);
            if (sender) jobs.push_back(CopyJob{(const unsigned char*)ctx->leaves[e.key].rec, stage, e.bytes});
            else jobs.push_back(CopyJob{stage, (unsigned char*)new_rec[j], e.bytes});
            used += sz;
        }
        i = j;
        if (!have) break;
        rc = ensure(ctx, &ctx->d_repl_jobs, &ctx->repl_jobs_cap, jobs.size() * sizeof(CopyJob));
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->d_repl_jobs, jobs.data(), jobs.size() * sizeof(CopyJob), cudaMemcpyHostToDevice, ctx->stream));
        if (sender) { k_copy_records<<<dim3((unsigned)jobs.size(), 16), 256, 0, ctx->stream>>>((const CopyJob*)ctx->d_repl_jobs); ctx->st.kernel_launches++; }
        rc = xfer(used);
        if (rc) return rc;
        if (!sender) { k_copy_records<<<dim3((unsigned)jobs.size(), 16), 256, 0, ctx->stream>>>((const CopyJob*)ctx->d_repl_jobs); ctx->st.kernel_launches++; }
        CK(cudaGetLastError());
    }
    return 0;
}

// Receiver, before the records arrive: erases, table capacity, arena space for every travelling record.
static int receive_prepare(gpis_ctx* ctx, const std::vector<ReplEntry>& idx, std::vector<uint64_t>& new_rec) {
    std::vector<SlotUpdate> ups;
    for (const ReplEntry& e : idx) if (e.erase) erase_one(ctx, e.key, ups);
    int rc = apply_updates(ctx, ups);
    if (rc) return rc;
    rc = table_reserve(ctx, (int)idx.size());
    if (rc) return rc;
    new_rec.assign(idx.size(), 0);
    for (size_t i = 0; i < idx.size(); ++i)
        if (!idx[i].erase && idx[i].bytes) {
            rc = arena_alloc(ctx, idx[i].bytes, &new_rec[i]);
            if (rc) { for (size_t k = 0; k < i; ++k) if (new_rec[k]) arena_free(ctx, new_rec[k], idx[k].bytes); return rc; }
        }
    return 0;
}
// Receiver, after the scatter kernels (stream order): install the table entries, release replaced records.
static int receive_install(gpis_ctx* ctx, const ReplHeader& hd, const std::vector<ReplEntry>& idx, const std::vector<uint64_t>& new_rec) {
    std::vector<SlotUpdate> ups;
    std::vector<std::pair<uint64_t, uint64_t>> to_free;
    for (size_t i = 0; i < idx.size(); ++i) {
        const ReplEntry& e = idx[i];
        if (e.erase) continue;
        auto it = ctx->leaves.find(e.key);
        if (it == ctx->leaves.end()) {
            HostLeaf hl{};
            hl.slot = take_slot(ctx, e.key);
            it = ctx->leaves.emplace(e.key, hl).first;
        }
        HostLeaf& hl = it->second;
        for (int c = 0; c < 3; ++c) { hl.cell[c] = e.cell[c]; hl.centre[c] = e.centre[c]; hl.lo[c] = e.lo[c]; hl.hi[c] = e.hi[c]; }
        hl.box_set = e.box_set != 0;
        if (e.bytes) {
            if (hl.rec) to_free.push_back({hl.rec, hl.rec_bytes});
            hl.rec = new_rec[i]; hl.rec_bytes = e.bytes;
            hl.N = e.meta[0]; hl.ng = e.meta[1]; hl.n = e.meta[2]; hl.nb = e.meta[3];
            ctx->max_nb = std::max(ctx->max_nb, hl.nb);
            ctx->max_N = std::max(ctx->max_N, hl.N);
        }
        ups.push_back(make_update(e.key, hl));
    }
    int rc = apply_updates(ctx, ups);
    if (rc) return rc;
    for (auto& f : to_free) arena_free(ctx, f.first, f.second);
    for (int c = 0; c < 3; ++c) ctx->qp.root_min[c] = hd.root_min[c];
    ctx->qp.levels = hd.levels;
    return 0;
}

extern "C" {

int gpis_replicate(gpis_ctx* ctx, int root) {
    NvtxRange nvtx_("gpis_replicate");
    if (!ctx) return GPIS_ERR_ARG;
    if (!ctx->comm) { ctx->err = "gpis_replicate before gpis_comm_init"; return GPIS_ERR_STATE; }
    if (root < 0 || root >= ctx->comm_world) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    const bool is_root = ctx->comm_rank == root;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    std::vector<ReplEntry> idx;
    ReplHeader hd{};
    if (is_root) build_message(ctx, false, hd, idx);
    int rc = ensure(ctx, &ctx->d_repl_idx, &ctx->repl_idx_cap, 256);
    if (rc) return rc;
    if (is_root) CK(cudaMemcpyAsync(ctx->d_repl_idx, &hd, sizeof(hd), cudaMemcpyHostToDevice, ctx->stream));
    NCK(ctx->p_ncclBroadcast(ctx->d_repl_idx, ctx->d_repl_idx, sizeof(hd), ncclUint8, root, ctx->comm, ctx->stream));
    if (!is_root) {
        CK(cudaMemcpyAsync(&hd, ctx->d_repl_idx, sizeof(hd), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (hd.magic != 0x4750495352455031ull) { ctx->err = "gpis_replicate: bad header"; return GPIS_ERR_STATE; }
        idx.resize(hd.n_entries);
    }
    const uint64_t idx_bytes = hd.n_entries * sizeof(ReplEntry);
    if (hd.n_entries) {
        rc = ensure(ctx, &ctx->d_repl_idx, &ctx->repl_idx_cap, idx_bytes);
        if (rc) return rc;
        if (is_root) CK(cudaMemcpyAsync(ctx->d_repl_idx, idx.data(), idx_bytes, cudaMemcpyHostToDevice, ctx->stream));
        NCK(ctx->p_ncclBroadcast(ctx->d_repl_idx, ctx->d_repl_idx, idx_bytes, ncclUint8, root, ctx->comm, ctx->stream));
        if (!is_root) {
            CK(cudaMemcpyAsync(idx.data(), ctx->d_repl_idx, idx_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    std::vector<uint64_t> new_rec;
    if (!is_root) { rc = receive_prepare(ctx, idx, new_rec); if (rc) return rc; }
    rc = stream_records(ctx, is_root, hd, idx, new_rec, [&](uint64_t used) -> int {
        NCK(ctx->p_ncclBroadcast(ctx->d_repl, ctx->d_repl, used, ncclUint8, root, ctx->comm, ctx->stream));
        return 0;
    });
    if (rc) return rc;
    if (!is_root) { rc = receive_install(ctx, hd, idx, new_rec); if (rc) return rc; }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->st.last_replicate_ms = ms;
    ctx->st.last_replicate_bytes = (int64_t)(hd.payload_bytes + idx_bytes + sizeof(hd));
    ctx->st.last_replicate_records = 0;
    for (const ReplEntry& e : idx) ctx->st.last_replicate_records += (e.bytes ? 1 : 0);
    if (is_root) { ctx->repl_touched.clear(); ctx->repl_trained.clear(); ctx->repl_erased.clear(); }
    return GPIS_OK;
}

// ------------------------------------------------------------------ snapshot
// The same message as gpis_replicate with every leaf in it, written to / read from a flat file:
//   [ReplHeader 64 B][gpis_config][n_entries x ReplEntry][record stream, 256-byte aligned records]
// The reference has no persistence (its map dies with the process, GPisMap3.cpp:951-972 only lists points); this
// gives the trained device map checkpoint/resume and a wire format a replica can be started from.
int gpis_snapshot_save(gpis_ctx* ctx, const char* path) {
    NvtxRange nvtx_("gpis_snapshot_save");
    if (!ctx || !path) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    std::FILE* f = std::fopen(path, "wb");
    if (!f) { ctx->err = std::string("cannot open ") + path; return GPIS_ERR_ARG; }
    ReplHeader hd{};
    std::vector<ReplEntry> idx;
    build_message(ctx, true, hd, idx);
    bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1 && std::fwrite(&ctx->cfg, sizeof(ctx->cfg), 1, f) == 1 &&
              (idx.empty() || std::fwrite(idx.data(), sizeof(ReplEntry), idx.size(), f) == idx.size());
    std::vector<unsigned char> host;
    int rc = 0;
    if (ok) rc = stream_records(ctx, true, hd, idx, std::vector<uint64_t>(), [&](uint64_t used) -> int {
        host.resize(used);
        CK(cudaMemcpyAsync(host.data(), ctx->d_repl, used, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (std::fwrite(host.data(), 1, used, f) != used) { ctx->err = "short write"; return GPIS_ERR_ARG; }
        return 0;
    });
    std::fclose(f);
    if (!ok) { ctx->err = "short write"; return GPIS_ERR_ARG; }
    return rc;
}

int gpis_snapshot_load(gpis_ctx* ctx, const char* path) {
    NvtxRange nvtx_("gpis_snapshot_load");
    if (!ctx || !path) return GPIS_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    { const int rcf_ = train_flush(ctx); if (rcf_) return rcf_; }   // an asynchronous training batch completes first
    std::FILE* f = std::fopen(path, "rb");
    if (!f) { ctx->err = std::string("cannot open ") + path; return GPIS_ERR_ARG; }
    ReplHeader hd{};
    gpis_config cfg{};
    std::vector<ReplEntry> idx;
    bool ok = std::fread(&hd, sizeof(hd), 1, f) == 1 && hd.magic == 0x4750495352455031ull && std::fread(&cfg, sizeof(cfg), 1, f) == 1;
    if (ok && (cfg.dim != ctx->cfg.dim || cfg.map_scale != ctx->cfg.map_scale || cfg.cluster_half != ctx->cfg.cluster_half)) {
        std::fclose(f);
        ctx->err = "snapshot was written with different map parameters (dim / map_scale / cluster_half)";
        return GPIS_ERR_ARG;
    }
    if (ok) { idx.resize(hd.n_entries); ok = idx.empty() || std::fread(idx.data(), sizeof(ReplEntry), idx.size(), f) == idx.size(); }
    if (!ok) { std::fclose(f); ctx->err = "not a gpis snapshot (or truncated)"; return GPIS_ERR_ARG; }
    int rc = gpis_reset(ctx);
    std::vector<uint64_t> new_rec;
    if (!rc) rc = receive_prepare(ctx, idx, new_rec);
    std::vector<unsigned char> host;
    if (!rc) rc = stream_records(ctx, false, hd, idx, new_rec, [&](uint64_t used) -> int {
        host.resize(used);
        if (std::fread(host.data(), 1, used, f) != used) { ctx->err = "truncated snapshot"; return GPIS_ERR_ARG; }
        CK(cudaMemcpyAsync(ctx->d_repl, host.data(), used, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // `host` is reused by the next chunk
        return 0;
    });
    std::fclose(f);
    if (!rc) rc = receive_install(ctx, hd, idx, new_rec);
    if (!rc) CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

int gpis_get_stats(gpis_ctx* ctx, gpis_stats* out) {
    if (!ctx || !out) return GPIS_ERR_ARG;
    train_read_ms(ctx, false);   // never blocks: last_train_ms is the most recently completed batch
    ctx->st.leaves = (int64_t)ctx->leaves.size();
    int64_t tr = 0;
    for (auto& kv : ctx->leaves) tr += kv.second.rec ? 1 : 0;
    ctx->st.leaves_trained = tr;
    ctx->st.arena_bytes_used = (int64_t)ctx->arena_used;
    ctx->st.arena_bytes_reserved = (int64_t)ctx->arena_reserved;
    *out = ctx->st;
    return GPIS_OK;
}

#ifdef E3_TIMING
int gpis_debug_timing(long long* out) {   // 64 warps x 32 counters of the instrumented CTA, then reset
    cudaMemcpyFromSymbol(out, g_e3_timing, sizeof(long long) * 64 * 32);
    static long long zeros[64 * 32];
    cudaMemcpyToSymbol(g_e3_timing, zeros, sizeof(zeros));
    return 0;
}
#endif

#ifdef K1_TIMING
int gpis_debug_train_timing(long long* out) {   // 8 counters of block 0 of k_leaf_train, then reset
    cudaMemcpyFromSymbol(out, g_k1_timing, sizeof(long long) * 8);
    static long long zeros[8];
    cudaMemcpyToSymbol(g_k1_timing, zeros, sizeof(zeros));
    return 0;
}
#endif

// Development/test aid (no device needed): the elimination program of k_eval_v3 for (nb, warp) as int32 words
// (query_v3.cuh: header, waves, visits, terminators, variance rows). Returns the number of words, or the
// required count when cap is too small.
int gpis_debug_program(int nb, int warp, int32_t* out, int cap) {
    if (nb < 1 || nb > E3_PROG_MAXNB || warp < 0 || warp >= E3_WARPS) return -1;
    E3WarpProg wp[E3_WARPS];
    for (int w = 0; w < E3_WARPS; ++w) e3_build_warp(nb, w, wp[w]);
    e3_assign_variance(nb, wp);
    std::vector<int4> recs;
    e3_emit_program(wp[warp], recs);
    const int words = (int)recs.size() * 4;
    if (out && cap >= words) std::memcpy(out, recs.data(), sizeof(int4) * recs.size());
    return words;
}

int gpis_set_eval_version(gpis_ctx* ctx, int v) {
    if (!ctx || (v != 1 && v != 3)) return GPIS_ERR_ARG;
    ctx->eval_version = v;
    return GPIS_OK;
}

}  // extern "C"
