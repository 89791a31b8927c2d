/*
 * gpis_b200.h — C ABI of the B200-native GPisMap hot path (libgpis_b200.so).
 *
 * The reference (leebhoram/GPisMap) has no plugin / FFI layer: its seam is the public C++
 * class API (cpp/include/GPisMap3.h:118-127, cpp/include/GPisMap.h:98-106) plus the MATLAB mex
 * command protocol. This ABI is the new boundary that sits directly underneath those class
 * methods: host C++ (sensor pre-processing, quad/octree bookkeeping — gpismap_b200/host/)
 * calls these entry points, and nothing above them sees CUDA. Every entry point names the
 * reference code it replaces. INTEGRATION.md shows how a maintainer of the reference would
 * bind it.
 *
 * Conventions: plain pointers and sizes only; all pointers are HOST memory unless the name ends
 * in `_device`; the caller owns every buffer; calls on one context must not overlap (the
 * reference classes are not thread-safe either); every function returns 0 on success and a
 * negative gpis_status on failure, with a message available from gpis_last_error(). Nothing
 * here falls back to the CPU: without a usable sm_100 device gpis_create fails.
 *
 * Geometry vocabulary: a *leaf* (reference: cluster, "level C" node) is a tree node
 * whose half-length equals cluster_half; it owns one local GP. Leaves are addressed by their
 * integer lattice cell: cell[c] = floor(centre[c] / (2*cluster_half)) (centres are odd multiples
 * of cluster_half because the reference roots its tree at the origin, GPisMap3.cpp:574,
 * GPisMap.cpp:460).
 */
#ifndef GPIS_B200_H
#define GPIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpis_ctx gpis_ctx;

typedef enum gpis_status {
    GPIS_OK = 0,
    GPIS_ERR_ARG = -1,      /* null / out-of-range argument */
    GPIS_ERR_CUDA = -2,     /* CUDA runtime error (message has the details) */
    GPIS_ERR_NODEVICE = -3, /* no sm_100-class device: there is no CPU path */
    GPIS_ERR_CAPACITY = -4, /* leaf too large (N > GPIS_MAX_SAMPLES) or table full */
    GPIS_ERR_STATE = -5     /* e.g. gpis_obs_test before gpis_obs_train_* */
} gpis_status;

#define GPIS_MAX_SAMPLES 1536 /* samples per training ball (reference max observed: 663) */
#define GPIS_MAX_N 6144       /* unknowns per leaf system, n = N + dim*ng */

/* Runtime parameters. Defaults reproduce the reference's compile-time macros
 * (cpp/include/params.h:27-110) and the constants hard-coded in its orchestrators. */
typedef struct gpis_config {
    int32_t dim;          /* 2 (GPisMap) or 3 (GPisMap3) */
    int32_t device;       /* CUDA device ordinal */
    float map_scale;      /* Matern length scale l: params.h:92 (0.04) / :73 (1.2) */
    float map_noise;      /* params.h:93 (5e-3) / :74 (1e-2); var_f preset = 1 + noise */
    float cluster_half;   /* params.h:41 (0.025) / :34 (0.8) */
    float search_half;    /* query AABB half-width: GPisMap3.cpp:811 (3*C_leng) / GPisMap.cpp:680 (4*l) */
    float var_thre;       /* fusion threshold: GPisMap3.cpp:800 (0.5) / GPisMap.cpp:671 (0.4) */
    float obs_scale;      /* params.h:97 (0.5) */
    float obs_noise;      /* params.h:98 (0.01) */
    int32_t max_leaves;   /* capacity of the device leaf table (grown on demand) */
    int32_t reserved0;
    uint64_t arena_chunk_bytes; /* granularity of the device arena that holds trained leaf records */
} gpis_config;

/* Fill *cfg with the reference defaults for dim = 2 or 3. */
int gpis_config_default(gpis_config* cfg, int dim);

int gpis_create(gpis_ctx** out, const gpis_config* cfg);
void gpis_destroy(gpis_ctx* ctx);
/* Drop every leaf and the observation GP: GPisMap3::reset (GPisMap3.cpp:99-115), GPisMap::reset
 * (GPisMap.cpp:87-103). */
int gpis_reset(gpis_ctx* ctx);
const char* gpis_last_error(const gpis_ctx* ctx);
int gpis_device(const gpis_ctx* ctx);

/* ---------------------------------------------------------------- leaf GPs (K1 + K3)
 * Train n_leaves local GPs and install them in the device leaf table.
 * Replaces updateGPs_kernel → OnGPIS::train → covFnc (GPisMap3.cpp:698-718, GPisMap.cpp:574-594,
 * OnGPIS.cpp:34-149, covFnc.cpp:142-256 / 317-402) and OcTree::Update(gp) (octree.cpp:569-572).
 *   cells    n_leaves x dim   integer lattice cell of each leaf
 *   centres  n_leaves x dim   the float centres held by the host tree (used verbatim in the
 *                             query's AABB / distance tests so neighbour sets stay bit-exact)
 *   offsets  n_leaves + 1     CSR into samples
 *   samples  offsets[n] x (2*dim+3) floats: pos[dim], grad[dim], val, pose_sig, grad_sig, in the
 *            tree's QueryRange DFS order (octree.cpp:777-804) — that order is the row order of K
 *   status   n_leaves         out, optional: number of non-positive Cholesky pivots (the
 *                             reference never checks LLT::info(), OnGPIS.cpp:139), or GPIS_ERR_CAPACITY
 *                             for a leaf beyond GPIS_MAX_SAMPLES / GPIS_MAX_N: such a leaf is registered
 *                             but not retrained (it keeps its previous GP), the rest of the batch is
 *                             trained and the call still returns GPIS_OK (gpis_last_error has a note).
 * Every leaf is validated and every record reserved before anything is mutated: on an error return
 * the table, the arena and the installed GPs are as before the call.
 * A leaf with an empty sample range is only registered as non-empty/untrained (the reference
 * skips training when QueryRange returns nothing, GPisMap3.cpp:710). */
int gpis_leaves_update(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres,
                       const int32_t* offsets, const float* samples, int32_t* status);

/* ---- device-side dirty set and training-set gather ---------------------------------------------
 * Instead of walking its tree once per dirty leaf and shipping every training ball (gpis_leaves_update), the host keeps
 * the samples on the device, one block per leaf, and only re-sends the leaves whose samples changed:
 *   gpis_samples_set        replace the sample lists of n_leaves leaves. CSR like gpis_leaves_update, but `samples` holds
 *                           each leaf's OWN samples, in the tree's DFS order (getAllChildrenNonEmptyNodes,
 *                           octree.cpp:806-827). An empty range drops the leaf's list. Unknown leaves are registered
 *                           (untrained). Leaves removed with gpis_leaves_erase lose their list.
 *   gpis_leaves_train_dirty replaces updateGPs for the leaves in `active_cells` (GPisMap3.cpp:720-792,
 *                           GPisMap.cpp:596-663): the device expands them to the dirty set (every registered leaf whose
 *                           effective box touches AABB(centre, radius), inclusive float compare), gathers each dirty
 *                           leaf's training ball (samples with |p - c|^2 < radius^2 in the reference's float arithmetic,
 *                           in QueryRange's DFS order) from the sample store, trains (K1) and installs the records.
 *                           radius = Rtimes * cluster_half (GPisMap3.cpp:707,733) or 4 * cluster_half (GPisMap.cpp).
 *                           The leaf table (boxes, root box: gpis_leaves_mark / _set_boxes / _rebase) must be current.
 * Results are identical to the host-gather path (tests/test_gpu_parity.py::test_device_gather_equals_host_gather). */
int gpis_samples_set(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres, const int32_t* offsets,
                     const float* samples);
int gpis_leaves_train_dirty(gpis_ctx* ctx, int n_active, const int32_t* active_cells, float radius, int32_t* n_trained);

/* Overlap of leaf training with the next frame's host work. The reference's update() trains inside updateGPs and returns
 * when every GP is ready (GPisMap3.cpp:218-237); nothing before the next test() reads those GPs (reEvalPoints and
 * evalPoints use the observation GP only). gpis_set_train_mode selects when K1 of gpis_leaves_train_dirty runs:
 *   0 (default)  before gpis_leaves_train_dirty returns;
 *   1            at once, on a low-priority stream; gpis_leaves_train_dirty returns after the launch;
 *   2            at the end of the next gpis_reeval (the next frame's device work goes first, K1 then runs beside that
 *                frame's serial host passes), or at the next entry point below, whichever comes first;
 *   3            when gpis_train_kick is called (the host picks the point of its frame from which it has no more device
 *                work of its own), or at the next entry point below.
 * In modes 1 and 2 every entry point that reads or changes records, the leaf table or the arena (gpis_query*,
 * gpis_leaf_get, gpis_leaves_*, gpis_samples_set, gpis_replicate, gpis_snapshot_*, gpis_reset) first waits for the
 * batch and installs its records, so results never depend on the mode. gpis_train_wait does only that.
 * gpis_stats.last_train_ms is then the most recently COMPLETED batch (gpis_get_stats never blocks). */
int gpis_set_train_mode(gpis_ctx* ctx, int mode);
int gpis_train_kick(gpis_ctx* ctx);
int gpis_train_wait(gpis_ctx* ctx);

/* Register leaves that hold samples but have no GP yet (a sample inserted through root growth
 * does not activate its leaf, octree.cpp:297-301, 209-211); they still count as query
 * candidates (octree.cpp:861-893). */
int gpis_leaves_mark(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* centres);

/* Override the float box used for candidate tests of already-registered leaves (default: centre -/+
 * cluster_half). boxes: n_leaves x 2*dim floats, lo[dim] then hi[dim]. The reference's
 * QueryNonEmptyLevelC prunes its descent at EVERY tree level with that level's float box
 * (octree.cpp:864-867), and a parent's c-l / c+l can round past its child's; a query box that merely
 * touches a lattice plane is then cut off above the leaf. The host tree therefore hands down, per
 * leaf, the intersection of the boxes of the leaf and all its ancestors, which makes the flat
 * device lookup return bit-identical candidate sets. */
int gpis_leaves_set_boxes(gpis_ctx* ctx, int n_leaves, const int32_t* cells, const float* boxes);

/* Remove leaves that became empty (OcTree::Remove collapsing a node, octree.cpp:548-561). */
int gpis_leaves_erase(gpis_ctx* ctx, int n_leaves, const int32_t* cells);

/* Tell the device the current root box so that candidate ties are broken in the tree's DFS child
 * order (octree.cpp:844-851: NWF,NEF,SWF,SEF,NWB,NEB,SWB,SEB; quadtree.cpp:630-633). root_min_cell
 * is the lattice cell of the root's minimum corner, levels = log2(root_half / cluster_half).
 * Called after root growth (octree.cpp:151-212). */
int gpis_rebase(gpis_ctx* ctx, const int32_t* root_min_cell, int levels);
/* Read the root box back (3 ints + levels): what a replica needs next to the imported records. */
int gpis_get_rebase(gpis_ctx* ctx, int32_t* root_min_cell, int* levels);

/* Read one trained leaf back (tests / debugging). Any out pointer may be NULL.
 * L is returned dense row-major n x n (lower), as OnGPIS holds it (OnGPIS.h:40-43). Returns n,
 * 0 if the leaf is unknown or untrained. */
int gpis_leaf_get(gpis_ctx* ctx, const int32_t* cell, int32_t* N, int32_t* ng, float* alpha, float* L,
                  float* gradflag, int cap_n);

/* ---------------------------------------------------------------- queries (K3 + K4)
 * SDF value, gradient and variances for n query points. Replaces GPisMap3::test → test_kernel →
 * QueryNonEmptyLevelC + OnGPIS::testSinglePoint (GPisMap3.cpp:794-949, octree.cpp:861-893,
 * OnGPIS.cpp:177-216) and the 2D copies (GPisMap.cpp:665-810, OnGPIS.cpp:218-239).
 *   x    n x dim, interleaved per point
 *   res  n x 2(1+dim), read-modify-write: [f, grad(dim), var_f, var_grad(dim)]; fields the
 *        reference's control flow does not reach keep the caller's contents; var_f is always
 *        preset to 1 + map_noise (GPisMap3.cpp:816). */
int gpis_query(gpis_ctx* ctx, const float* x, int64_t n, float* res_inout);
int gpis_query_device(gpis_ctx* ctx, const float* x_device, int64_t n, float* res_inout_device);
/* Same, also returning per query [n_candidates, id0, id1, id2] where id = index into the order
 * in which leaves were first registered (gpis_leaf_index) and -1 = none, and a flag set when an
 * exact centre-distance tie touches the picks. Host pointers. */
int gpis_query_debug(gpis_ctx* ctx, const float* x, int64_t n, float* res_inout, int32_t* chosen4,
                     int32_t* tie);
/* Index (slot) of a leaf in the device table, -1 if absent. */
int gpis_leaf_index(gpis_ctx* ctx, const int32_t* cell);

/* ---------------------------------------------------------------- observation GPs (K2)
 * Replaces ObsGP2D::train (ObsGP.cpp:204-350: partition + one GPou per tile with >= 1 valid
 * pixel) and ObsGP1D::train (ObsGP.cpp:85-143).
 *   vu    2*ni*nj floats, [v,u] interleaved, index j*ni+i;  zinv  ni*nj (<= 0 ⇒ invalid)
 * The partition boundaries are recomputed only when (ni,nj) changes or after gpis_reset, like the
 * reference (ObsGP.cpp:335-337). */
int gpis_obs_train_2d(gpis_ctx* ctx, const float* vu, const float* zinv, int ni, int nj);
int gpis_obs_train_1d(gpis_ctx* ctx, const float* theta, const float* f, int n);
/* Batched ObsGP*::test (ObsGP.cpp:145-187, 352-408) + GPou::test (ObsGP.cpp:50-62).
 *   xt   m x d (d = 2: [v,u]; d = 1: bearing); val / var are read-modify-write: where the
 *   reference would not evaluate, val is untouched and var = 1e6. */
int gpis_obs_test(gpis_ctx* ctx, const float* xt, int d, int m, float* val_inout, float* var_inout);

/* ---- one depth frame on the device (SURVEY.md 8 f-3 + f-1) --------------------------------------------------------
 * Replaces, for GPisMap3::update, preprocData (GPisMap3.cpp:125-216), regressObs (:239-256, = gpis_obs_train_2d) and
 * the numerics of evalPoints (:580-696): validity / inverse depth / back-projection / local->global of the
 * sub-sampled pixels, the observation GP of the frame, then for every valid measurement the GP test at its pixel and
 * at six finite-difference probes, occupancy values, surface normal and noise terms. What is left for the host is
 * the serial tree insertion. Scalar types follow the reference expression by expression (the map's samples stay
 * bit-identical to the reference's: tests/test_gpu_parity.py::test_bench_scale_map_matches_reference).
 *   depth      N floats, metres, column-major dataz[col*height + row] (GPisMap3.cpp:183)
 *   vu_grid    2*(width/skip)*(height/skip) floats [v,u], index (height/skip)*col_ + row_ (GPisMap3.cpp:155-170)
 *   outputs    n_valid = valid measurements K in the reference's order; range_obs_max; per measurement k < K:
 *              xyz_global[3], status (0 = centre test rejected: skipped; 1 = a probe rejected: inserted then removed,
 *              :652-655; 2 = ok), grad[3] (global normal), noise, grad_noise. cap = capacity of those arrays.
 * With K <= 1 nothing is regressed (GPisMap3.cpp:212-215). Afterwards gpis_obs_test serves this frame's GP. */
typedef struct gpis_frame_params {
    int32_t width, height, skip, reserved;
    float pose[12];                 /* [t(3) | R column-major(9)], local -> global */
    float delx, obs_var_thre, min_position_noise, min_grad_noise;
    double max_range, min_range;    /* params.h:77-78 (4.0, 0.4) */
} gpis_frame_params;
int gpis_frame_eval(gpis_ctx* ctx, const float* depth, int N, const float* vu_grid, const gpis_frame_params* fp,
                    int32_t* n_valid, float* range_obs_max, int32_t cap, float* xyz_global, int32_t* status,
                    float* grad, float* noise, float* grad_noise);

/* The numerics of reEvalPoints (GPisMap3.cpp:321-534) for n existing samples against the current frame's observation
 * GP: projection into the camera, first test + occupancy gate, the 10-step surface walk, six probes, fused position /
 * normal / noises. samples8: [pos(3), grad(3), pose_sig, grad_sig] per sample. Outputs per sample: action (-1 = not
 * re-evaluated: behind the camera, untrusted test or occupancy gate; 0 = nothing; 1 = double both noises, :451-454;
 * 2 = replace the sample by pos_new / grad_new / noise / grad_noise). The host then removes / re-inserts in the
 * reference's order (updateMapPoints, GPisMap3.cpp:258-319). fp->pose, delx, obs_var_thre and the noise floors are
 * read from fp; width / height / skip / ranges are ignored. */
int gpis_reeval(gpis_ctx* ctx, int n, const float* samples8, const gpis_frame_params* fp, float map_noise_param,
                int32_t* action, float* pos_new, float* grad_new, float* noise, float* grad_noise);

/* ---------------------------------------------------------------- replication (K5), snapshot and stats */
/* K5 inside the library. The trained leaf table is replicated over NCCL (NVLink / NVSwitch) so that query batches can
 * be sharded over GPUs; the query path itself has no collective.
 *   gpis_comm_unique_id  rank 0 calls it and hands the 128 bytes to the other ranks by any host channel
 *                        (ncclGetUniqueId); NCCL is resolved at run time, preferring the copy already in the process.
 *   gpis_comm_init       every rank, once per context (ncclCommInitRank).
 *   gpis_replicate       collective: every rank calls it after the root's update. The root ships everything that
 *                        changed since its previous call — new records (packed by one kernel, broadcast in bounded
 *                        chunks), table-only changes (leaves registered without a GP, effective boxes), erased
 *                        leaves, the root box — and the other ranks install it: afterwards they answer queries
 *                        bit-identically to the root. No per-record synchronisation. */
int gpis_comm_unique_id(void* id128);
int gpis_comm_init(gpis_ctx* ctx, int rank, int world, const void* id128);
int gpis_replicate(gpis_ctx* ctx, int root);

/* Snapshot of the trained device map: the message gpis_replicate ships, with every leaf in it, as a
 * flat file — [64-byte header][gpis_config][index of table entries][256-byte aligned leaf records]. The reference has
 * no persistence (GPisMap3.cpp:951-972 only lists sample positions). gpis_snapshot_load resets the context first and
 * refuses a file written with a different dim / map_scale / cluster_half. Queries after a load are bit-identical to
 * queries before the save. The host-side tree is not part of it (a loaded map answers test(), it cannot be updated). */
int gpis_snapshot_save(gpis_ctx* ctx, const char* path);
int gpis_snapshot_load(gpis_ctx* ctx, const char* path);

typedef struct gpis_stats {
    int64_t leaves;            /* leaves registered in the table */
    int64_t leaves_trained;    /* ... of which hold a GP */
    int64_t arena_bytes_used;  /* bytes of trained records */
    int64_t arena_bytes_reserved;
    /* last gpis_leaves_update */
    int64_t last_train_leaves, last_train_sum_N, last_train_sum_n;
    double last_train_flops;   /* sum n^3/3 + 2n^2 + 30N^2 */
    double last_train_bytes;   /* sum 52N + 4n + 2n(n+1) */
    float last_train_ms;       /* CUDA-event time of the training kernels */
    int32_t last_train_skipped; /* leaves beyond GPIS_MAX_SAMPLES / GPIS_MAX_N that were not retrained */
    /* last gpis_query* */
    int64_t last_query_n, last_query_evals; /* evaluated (query, leaf) pairs */
    double last_query_flops;   /* sum 4n^2 + 16n + 80N over evaluations */
    double last_query_bytes_gather;     /* 44/query + sum 16N + 4n + 2n(n+1) over evaluations */
    double last_query_bytes_compulsory; /* 44/query + sum over distinct leaves touched */
    float last_query_ms;       /* CUDA-event time of the query kernels (no H2D/D2H) */
    float last_query_eval_ms;  /* ... of which the leaf-evaluation (solve) kernels */
    int64_t kernel_launches;   /* kernels of this library launched since gpis_create */
    /* last gpis_query*: evaluation CTAs launched with 8, 6, 4 and 1 queries per CTA (batch fill = evals / capacity) */
    int64_t last_query_items[4];
    /* last gpis_replicate on this rank */
    int64_t last_replicate_bytes, last_replicate_records;
    float last_replicate_ms;   /* CUDA-event time on this rank's stream: pack, broadcasts, scatter, table apply */
    int32_t reserved1;
} gpis_stats;
int gpis_get_stats(gpis_ctx* ctx, gpis_stats* out);

/* ---- development and test aids (not part of the drop-in surface) ------------------------------
 * gpis_set_eval_version: 3 = production evaluation kernel (k_eval_v3, default), 1 = one CTA per
 *   (query, leaf) pair (k_eval_v1, also the fallback for leaves beyond k_eval_v3's shared memory);
 *   the parity tests run both.
 * gpis_debug_program: the host-generated visit program of k_eval_v3 for a leaf of nb block rows and
 *   one warp, as int32 words (layout: gpismap_b200/csrc/query_v3.cuh); needs no device. Returns the
 *   word count (also when cap is too small), -1 on bad arguments. tests/test_eval_programs.py checks
 *   coverage, the publish protocol and deadlock freedom on these programs. */
int gpis_set_eval_version(gpis_ctx* ctx, int version);
int gpis_debug_program(int nb, int warp, int32_t* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
