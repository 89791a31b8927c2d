/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
 *
 * Plain-C CPU restatement of GPisMap's data-parallel hot path (SURVEY.md §8a rows a1-a10,
 * a13, a14), each function citing the reference file:line it follows (paths are relative to
 * /root/reference). Built twice by oracle/Makefile: REAL=float (the parity oracle, with the
 * reference's selected double intermediates) and REAL=double (shadow truth used to
 * attribute error and calibrate tolerances).
 *
 * Parity status: PINNED against the reference itself. The reference ships no tests or
 * golden vectors, so the pin is the reference's own sources compiled
 * unmodified into oracle/_ref/libgpisref.so (dense LA from oracle/eigen_shim, because Eigen
 * is absent from the image) and the fixtures under tests/golden/ generated from it by
 * tests/golden/make_golden.py. tests/test_oracle_vs_ref.py and tests/test_oracle_golden.py
 * hold this restatement to those.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 */
#ifndef GPIS_ORACLE_H
#define GPIS_ORACLE_H

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#ifdef __cplusplus
extern "C" {
#endif

int gpo_real_bytes(void);

/* a1/a2: cpp/src/covFnc.cpp:29-33, 142-256 (3D), 317-402 (2D). x: N x dim row-major
 * (= dim x N column-major), K out: n x n row-major, n = N + dim*ng. Returns n. */
int gpo_matern_train(int dim, const float* x, const float* gradflag, int N, float scale, const float* sigx,
                     const float* siggrad, real* K);
/* a3: cpp/src/covFnc.cpp:258-314 (3D), 404-450 (2D); single test point. Ks out: n x (1+dim)
 * row-major. Returns n. */
int gpo_matern_test(int dim, const real* x, const int* gradidx, int N, int ng, const real* xt, real scale,
                    real* Ks);

/* a4: cpp/src/OnGPIS.cpp:34-89 (2D), 91-149 (3D). samples: N x (2*dim+3) floats
 * [pos, grad, val, pose_sig, grad_sig]. */
typedef struct gpo_gp gpo_gp;
gpo_gp* gpo_gp_train(int dim, const float* samples, int N, float scale, float noise);
void gpo_gp_free(gpo_gp* g);
int gpo_gp_n(const gpo_gp* g);
int gpo_gp_ng(const gpo_gp* g);
int gpo_gp_chol_fail(const gpo_gp* g);
/* alpha: n; L: n x n row-major dense lower; gradflag: N (0/1) */
void gpo_gp_get(const gpo_gp* g, real* alpha, real* L, float* gradflag);
/* a5: cpp/src/OnGPIS.cpp:177-216 (3D), 218-239 (2D). res: m rows of 2(1+dim), in/out. */
void gpo_gp_test(const gpo_gp* g, const real* x, int m, real* res);

/* a13+a14: cpp/src/octree.cpp:861-893, cpp/src/quadtree.cpp:643-671 (candidate rule) and
 * cpp/src/GPisMap3.cpp:794-902, cpp/src/GPisMap.cpp:665-763 (sort + fusion).
 * The map is a list of non-empty clusters in the tree's DFS order. gps[i] may be NULL
 * (non-empty but never trained). */
typedef struct gpo_map gpo_map;
/* boxes (optional, nclusters x 2*dim floats: lo[dim], hi[dim]): the effective box of each cluster =
 * intersection of its own float AABB with those of all its tree ancestors. The reference prunes
 * its DFS at every level with float boxes (octree.cpp:864-867), and c-l / c+l of a parent can round
 * past its child's, so a query box that only touches a lattice plane can be cut off above the
 * cluster level. NULL = own box only (centre -/+ cluster_half). */
gpo_map* gpo_map_create(int dim, int nclusters, const float* centres, float cluster_half, gpo_gp* const* gps,
                        float search_half, float var_thre, float noise, const float* boxes);
void gpo_map_free(gpo_map* m);
/* res: m rows of 2(1+dim), in/out (fields the logic does not reach stay untouched).
 * chosen (optional): m x 4 ints = [ncandidates, id0, id1, id2] (ids index the cluster list,
 * -1 = none; order = ascending centre distance, ties keep DFS order);
 * tie (optional): m ints, 1 if an exact distance tie touches the first min(nc,3) picks. */
void gpo_map_test(const gpo_map* m, const float* x, int n, real* res, int* chosen, int* tie);

/* a7-a10: cpp/src/covFnc.cpp:47-109, cpp/src/ObsGP.cpp:32-62 (GPou), 204-408 (ObsGP2D),
 * 85-187 (ObsGP1D). */
typedef struct gpo_obs gpo_obs;
gpo_obs* gpo_obs2d_train(const float* vu, const float* zinv, int ni, int nj);
gpo_obs* gpo_obs2d_retrain(const float* vu, const float* zinv, int ni, int nj, const gpo_obs* prev);
gpo_obs* gpo_obs1d_train(const float* theta, const float* f, int n);
void gpo_obs_free(gpo_obs* o);
int gpo_obs_ntiles(const gpo_obs* o);
void gpo_obs_tile_counts(const gpo_obs* o, int* counts);
int gpo_obs_bounds(const gpo_obs* o, float* bi, float* bj); /* returns nbi*65536+nbj */
/* xt: m x d (d=2: [v,u]; d=1: angle). val/var in/out: val untouched and var=1e6 where the
 * reference would not evaluate (cpp/src/ObsGP.cpp:363-377, 396-403, 152-186). */
void gpo_obs_test(const gpo_obs* o, const float* xt, int m, real* val, real* var);

void gpo_sort_replay(const float* keys, int n, int* idx_out);
int gpo_sort_replay_heapsorts(void);

#ifdef __cplusplus
}
#endif
#endif
