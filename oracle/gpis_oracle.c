/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
 * See gpis_oracle.h for scope, pinning status and who may call this.
 *
 * Arithmetic conventions (REAL=float build): storage and linear algebra are fp32; the
 * spots where the reference's unqualified exp()/sqrt()/double literals promote to double
 *  are reproduced with explicit casts. Linear-algebra operation order is
 * this file's own (row-oriented Cholesky-Banachiewicz, sequential sums): Eigen's order is
 * not reproducible without its sources and is absorbed by the stated tolerances.
 */
#include "gpis_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int gpo_real_bytes(void) { return (int)sizeof(real); }

/* Scalar type of the covariance ENTRIES. Normally the build's `real`. With -DCOV_FLOAT (libgpisoracle64r.so: REAL=double,
 * COV_FLOAT) every entry of K and k* is evaluated and rounded exactly as in the fp32 build and only the linear algebra
 * (Cholesky, solves, dot products, fusion) runs in double: "the reference's algorithm in exact arithmetic on the
 * reference's own float covariance entries". That is the arbiter for what fp32 pins: against it, the fp32 reference's
 * distance is its own linear-algebra rounding error, not the (common) rounding of the covariance entries. */
#ifdef COV_FLOAT
typedef float creal;
#else
typedef real creal;
#endif
/* ------------------------------------------------------------------ a1: covFnc.cpp:29-33 */
static real kf(creal r, creal a) { return (real)(creal)((1.0 + (double)(a * r)) * exp((double)(-a * r))); }
static real kf1(creal r, creal dx, creal a) { return (real)(creal)((double)(a * a * dx) * exp((double)(-a * r))); }
static real kf2(creal r, creal dx1, creal dx2, creal delta, creal a) {
    return (real)(creal)((double)(a * a * (delta - a * dx1 * dx2 / r)) * exp((double)(-a * r)));
}
static real sqrt_real(real v) { return sizeof(real) == 4 ? (real)sqrtf((float)v) : (real)sqrt((double)v); }
/* (x1.col(k)-x2.col(j)).norm(): sequential sum of squares, then sqrt in the scalar type */
static creal dist(const real* a, const real* b, int dim) {
    creal s = 0;
    for (int c = 0; c < dim; ++c) { creal d = (creal)a[c] - (creal)b[c]; s += d * d; }
    return sizeof(creal) == 4 ? (creal)sqrtf((float)s) : (creal)sqrt((double)s);
}
#define CDIFF(p, q) ((creal)(p) - (creal)(q))

/* ------------------------------------------------------------------ a2: train covariance
 * covFnc.cpp:142-256 (3D), 317-402 (2D). gradidx[k] = compacted index or -1
 * (covFnc.cpp:151-161). K is n x n row-major, fully (symmetrically) filled. */
static void matern_train_core(int dim, const real* x, const int* gradidx, int N, int ng, real scale,
                              const real* sigx, const real* siggrad, real* K) {
    const int n = N + dim * ng;
    const creal a = (creal)(sqrt(3.0) / (double)scale); /* covFnc.cpp:147 */
    const creal a2 = a * a;
    memset(K, 0, sizeof(real) * (size_t)n * n);
#define KK(i, j) K[(size_t)(i) * n + (j)]
    for (int k = 0; k < N; ++k) {
        const int gk = gradidx[k];
        for (int j = k; j < N; ++j) {
            if (k == j) {
                KK(k, k) = (real)(creal)(1.0 + (double)sigx[k]); /* :173 */
                if (gk >= 0) {
                    for (int c = 0; c < dim; ++c) {
                        const int kc = N + c * ng + gk;
                        if (dim == 2 && c == 0) /* 2D quirk, covFnc.cpp:352 */
                            KK(kc, kc) = (real)(creal)((double)a2 + sqrt((double)((creal)sigx[k] * (creal)siggrad[k])));
                        else
                            KK(kc, kc) = (real)(creal)(a2 + (creal)siggrad[k]); /* :182,186,190 / :355 */
                    }
                }
                continue;
            }
            const real* xk = x + (size_t)k * dim;
            const real* xj = x + (size_t)j * dim;
            const creal r = dist(xk, xj, dim);
            const int gj = gradidx[j];
            KK(k, j) = KK(j, k) = kf(r, a); /* :194-195 */
            if (gk >= 0) {
                for (int c = 0; c < dim; ++c) { /* :198-203 */
                    const int kc = N + c * ng + gk;
                    const real v = -kf1(r, CDIFF(xk[c], xj[c]), a);
                    KK(kc, j) = KK(j, kc) = v;
                }
                if (gj >= 0) {
                    for (int c = 0; c < dim; ++c) { /* :210-215: K(k,jind) = -K(j,kind) */
                        const int kc = N + c * ng + gk, jc = N + c * ng + gj;
                        const real v = -KK(j, kc);
                        KK(k, jc) = KK(jc, k) = v;
                    }
                    for (int c = 0; c < dim; ++c) /* :217-237, upper computed, lower mirrored */
                        for (int e = c; e < dim; ++e) {
                            const int kc = N + c * ng + gk, ke = N + e * ng + gk;
                            const int jc = N + c * ng + gj, je = N + e * ng + gj;
                            const real v = kf2(r, CDIFF(xk[c], xj[c]), CDIFF(xk[e], xj[e]), c == e ? (creal)1 : (creal)0, a);
                            KK(kc, je) = KK(je, kc) = v;
                            if (e != c) KK(ke, jc) = KK(jc, ke) = v;
                        }
                }
            } else if (gj >= 0) { /* :239-249 */
                for (int c = 0; c < dim; ++c) {
                    const int jc = N + c * ng + gj;
                    const real v = kf1(r, CDIFF(xk[c], xj[c]), a);
                    KK(k, jc) = KK(jc, k) = v;
                }
            }
        }
    }
#undef KK
}

int gpo_matern_train(int dim, const float* x, const float* gradflag, int N, float scale, const float* sigx,
                     const float* siggrad, real* K) {
    int* gi = (int*)malloc(sizeof(int) * (N > 0 ? N : 1));
    real* xr = (real*)malloc(sizeof(real) * (size_t)(N > 0 ? N : 1) * dim);
    real* sx = (real*)malloc(sizeof(real) * (N > 0 ? N : 1));
    real* sg = (real*)malloc(sizeof(real) * (N > 0 ? N : 1));
    int ng = 0;
    for (int k = 0; k < N; ++k) {
        gi[k] = gradflag[k] > 0.5f ? ng++ : -1;
        sx[k] = sigx[k]; sg[k] = siggrad[k];
        for (int c = 0; c < dim; ++c) xr[(size_t)k * dim + c] = x[(size_t)k * dim + c];
    }
    if (K) matern_train_core(dim, xr, gi, N, ng, (real)scale, sx, sg, K);
    free(gi); free(xr); free(sx); free(sg);
    return N + dim * ng;
}

/* ------------------------------------------------------------------ a3: test covariance
 * covFnc.cpp:258-314 (3D), 404-450 (2D), m = 1. Ks is n x (1+dim) row-major. */
int gpo_matern_test(int dim, const real* x, const int* gradidx, int N, int ng, const real* xt, real scale,
                    real* Ks) {
    const int n = N + dim * ng, w = 1 + dim;
    const creal a = (creal)(sqrt(3.0) / (double)scale);
    memset(Ks, 0, sizeof(real) * (size_t)n * w);
    for (int k = 0; k < N; ++k) {
        const real* xk = x + (size_t)k * dim;
        const creal r = dist(xk, xt, dim);
        real* row = Ks + (size_t)k * w;
        row[0] = kf(r, a);
        for (int c = 0; c < dim; ++c) row[1 + c] = kf1(r, CDIFF(xk[c], xt[c]), a);
        const int gk = gradidx[k];
        if (gk >= 0) {
            for (int c = 0; c < dim; ++c) {
                real* rc = Ks + (size_t)(N + c * ng + gk) * w;
                rc[0] = -row[1 + c];
                for (int e = 0; e < dim; ++e) {
                    /* the reference evaluates the upper triangle and copies it down */
                    const int c0 = c < e ? c : e, e0 = c < e ? e : c;
                    rc[1 + e] = kf2(r, CDIFF(xk[c0], xt[c0]), CDIFF(xk[e0], xt[e0]), c == e ? (creal)1 : (creal)0, a);
                }
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------ dense LA helpers */
/* Row-oriented Cholesky (lower, row-major, in place on the lower triangle). Returns the
 * number of non-positive pivots (the reference never checks, OnGPIS.cpp:139). */
static int chol_lower(real* A, int n) {
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        real* ai = A + (size_t)i * n;
        for (int j = 0; j <= i; ++j) {
            const real* aj = A + (size_t)j * n;
            real s = ai[j];
            for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
            if (j < i) ai[j] = s / aj[j];
            else {
                if (!(s > 0)) ++bad;
                ai[i] = sqrt_real(s);
            }
        }
        for (int j = i + 1; j < n; ++j) ai[j] = 0;
    }
    return bad;
}
/* L z = b, nrhs columns stored row-major b[n][nrhs] */
static void fwd_solve(const real* L, int n, real* b, int nrhs) {
    for (int i = 0; i < n; ++i) {
        const real* li = L + (size_t)i * n;
        for (int c = 0; c < nrhs; ++c) {
            real s = b[(size_t)i * nrhs + c];
            for (int k = 0; k < i; ++k) s -= li[k] * b[(size_t)k * nrhs + c];
            b[(size_t)i * nrhs + c] = s / li[i];
        }
    }
}
/* L^T a = z */
static void bwd_solve(const real* L, int n, real* b) {
    for (int i = n - 1; i >= 0; --i) {
        real s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
        b[i] = s / L[(size_t)i * n + i];
    }
}

/* ------------------------------------------------------------------ a4: OnGPIS::train */
struct gpo_gp {
    int dim, N, ng, n, bad;
    real scale;
    real three_over_scale;
    real* x;      /* N x dim */
    int* gradidx; /* N */
    real* alpha;  /* n */
    real* L;      /* n x n row-major */
};

gpo_gp* gpo_gp_train(int dim, const float* samples, int N, float scale, float noise) {
    (void)noise; /* "currently noise param is not effective" (OnGPIS.h:45-46) */
    if (N <= 0) return NULL; /* OnGPIS.cpp:40,97: stays untrained */
    const int w = 2 * dim + 3;
    gpo_gp* g = (gpo_gp*)calloc(1, sizeof(gpo_gp));
    g->dim = dim; g->N = N; g->scale = (real)scale;
    g->three_over_scale = (real)(3.0 / (double)((real)scale * (real)scale)); /* OnGPIS.h:58 */
    g->x = (real*)malloc(sizeof(real) * (size_t)N * dim);
    g->gradidx = (int*)malloc(sizeof(int) * N);
    real* sigx = (real*)malloc(sizeof(real) * N);
    real* sigg = (real*)malloc(sizeof(real) * N);
    real* f = (real*)malloc(sizeof(real) * N);
    real* gv = (real*)malloc(sizeof(real) * (size_t)N * dim);
    int ng = 0;
    for (int k = 0; k < N; ++k) {
        const float* s = samples + (size_t)k * w;
        for (int c = 0; c < dim; ++c) g->x[(size_t)k * dim + c] = s[c];
        f[k] = s[2 * dim];
        sigx[k] = s[2 * dim + 1];
        sigg[k] = s[2 * dim + 2];
        int allsmall = 1;
        for (int c = 0; c < dim; ++c)
            if (!(fabs((double)s[dim + c]) < 1e-6)) allsmall = 0;
        /* OnGPIS.cpp:63-66 / 122-125: float siggrad compared with the double literal */
        if ((double)s[2 * dim + 2] > 0.1001 || allsmall) {
            g->gradidx[k] = -1;
            sigx[k] = 2.0;
        } else {
            for (int c = 0; c < dim; ++c) gv[(size_t)ng * dim + c] = s[dim + c];
            g->gradidx[k] = ng++;
        }
    }
    g->ng = ng;
    const int n = g->n = N + dim * ng;
    g->alpha = (real*)malloc(sizeof(real) * n);
    g->L = (real*)malloc(sizeof(real) * (size_t)n * n);
    /* y = [f; gx(valid); gy(valid); gz(valid)]  (OnGPIS.cpp:75-76, 135-136) */
    for (int k = 0; k < N; ++k) g->alpha[k] = f[k];
    for (int c = 0; c < dim; ++c)
        for (int q = 0; q < ng; ++q) g->alpha[N + c * ng + q] = gv[(size_t)q * dim + c];
    matern_train_core(dim, g->x, g->gradidx, N, ng, g->scale, sigx, sigg, g->L);
    g->bad = chol_lower(g->L, n);              /* :139 */
    fwd_solve(g->L, n, g->alpha, 1);           /* :141-142 */
    bwd_solve(g->L, n, g->alpha);              /* :143 */
    free(sigx); free(sigg); free(f); free(gv);
    return g;
}
void gpo_gp_free(gpo_gp* g) {
    if (!g) return;
    free(g->x); free(g->gradidx); free(g->alpha); free(g->L); free(g);
}
int gpo_gp_n(const gpo_gp* g) { return g ? g->n : 0; }
int gpo_gp_ng(const gpo_gp* g) { return g ? g->ng : 0; }
int gpo_gp_chol_fail(const gpo_gp* g) { return g ? g->bad : 0; }
void gpo_gp_get(const gpo_gp* g, real* alpha, real* L, float* gradflag) {
    if (alpha) memcpy(alpha, g->alpha, sizeof(real) * g->n);
    if (L) memcpy(L, g->L, sizeof(real) * (size_t)g->n * g->n);
    if (gradflag)
        for (int k = 0; k < g->N; ++k) gradflag[k] = g->gradidx[k] >= 0 ? 1.f : 0.f;
}

/* ------------------------------------------------------------------ a5: testSinglePoint
 * OnGPIS.cpp:177-216 (3D) / 218-239 (2D). out = [f, grad(dim), var(1+dim)] written in place. */
static void gp_test_point(const gpo_gp* g, const real* xt, real* out) {
    const int n = g->n, dim = g->dim, w = 1 + dim;
    real* Ks = (real*)malloc(sizeof(real) * (size_t)n * w);
    gpo_matern_test(dim, g->x, g->gradidx, g->N, g->ng, xt, g->scale, Ks);
    for (int c = 0; c < w; ++c) { /* K^T alpha, :187 */
        real s = 0;
        for (int i = 0; i < n; ++i) s += Ks[(size_t)i * w + c] * g->alpha[i];
        out[c] = s;
    }
    fwd_solve(g->L, n, Ks, w); /* :199 */
    for (int c = 0; c < w; ++c) {
        real s = 0;
        for (int i = 0; i < n; ++i) { real v = Ks[(size_t)i * w + c]; s += v * v; }
        /* priors: :203-212 (3D: 1.001, 3/l^2+0.001), :235-237 (2D: 1.01, 3/l^2+0.1) */
        double prior;
        if (dim == 3) prior = c == 0 ? 1.001 : (double)g->three_over_scale + 0.001;
        else prior = c == 0 ? 1.01 : (double)g->three_over_scale + 0.1;
        out[w + c] = (real)(prior - (double)s);
    }
    free(Ks);
}
void gpo_gp_test(const gpo_gp* g, const real* x, int m, real* res) {
    if (!g) return; /* untrained: outputs untouched (OnGPIS.cpp:179-180) */
    const int w = 2 * (1 + g->dim);
    for (int i = 0; i < m; ++i) gp_test_point(g, x + (size_t)i * g->dim, res + (size_t)i * w);
}

/* ------------------------------------------------------------------ a13 + a14: map query */
struct gpo_map {
    int dim, nc;
    float* centres; /* nc x dim */
    float* boxes;   /* nc x 2*dim: effective lo, hi */
    float half, search_half;
    real var_thre, noise;
    gpo_gp** gps;
};
gpo_map* gpo_map_create(int dim, int nclusters, const float* centres, float cluster_half, gpo_gp* const* gps,
                        float search_half, float var_thre, float noise, const float* boxes) {
    gpo_map* m = (gpo_map*)calloc(1, sizeof(gpo_map));
    m->dim = dim; m->nc = nclusters; m->half = cluster_half; m->search_half = search_half;
    m->var_thre = (real)var_thre; m->noise = (real)noise;
    m->centres = (float*)malloc(sizeof(float) * (size_t)(nclusters > 0 ? nclusters : 1) * dim);
    memcpy(m->centres, centres, sizeof(float) * (size_t)nclusters * dim);
    m->gps = (gpo_gp**)malloc(sizeof(gpo_gp*) * (nclusters > 0 ? nclusters : 1));
    memcpy(m->gps, gps, sizeof(gpo_gp*) * nclusters);
    m->boxes = (float*)malloc(sizeof(float) * (size_t)(nclusters > 0 ? nclusters : 1) * 2 * dim);
    for (int i = 0; i < nclusters; ++i)
        for (int c = 0; c < dim; ++c) {
            m->boxes[(size_t)i * 2 * dim + c] = boxes ? boxes[(size_t)i * 2 * dim + c] : centres[(size_t)i * dim + c] - cluster_half;
            m->boxes[(size_t)i * 2 * dim + dim + c] = boxes ? boxes[(size_t)i * 2 * dim + dim + c] : centres[(size_t)i * dim + c] + cluster_half;
        }
    return m;
}
void gpo_map_free(gpo_map* m) {
    if (!m) return;
    free(m->centres); free(m->boxes); free(m->gps); free(m);
}

/* ------------------------------------------------------------------ libstdc++ std::sort replay
 * bits/stl_algo.h (GCC 13): __sort -> __introsort_loop (median-of-3 to first, unguarded partition,
 * threshold 16, depth limit 2*floor(log2 n), heapsort fallback) -> __final_insertion_sort.
 * Sorts the parallel arrays (idx, key) by key with comparator key[a] < key[b]; the arrays start in
 * the tree's DFS order. The heapsort fallback (__partial_sort = __heap_select + __sort_heap, bits/stl_heap.h) is
 * replayed too: it needs > 2 log2(n) bad partitions, which natural candidate lists do not produce, but an adversarial
 * one can (tests/test_oracle_vs_ref.py builds one with McIlroy's adversary against the real std::sort). */
static void ssr_swap(int* idx, float* key, int a, int b) {
    int ti = idx[a]; idx[a] = idx[b]; idx[b] = ti;
    float tk = key[a]; key[a] = key[b]; key[b] = tk;
}
static void ssr_unguarded_linear_insert(int* idx, float* key, int last) {
    const int vi = idx[last]; const float vk = key[last];
    int next = last - 1;
    while (vk < key[next]) { idx[last] = idx[next]; key[last] = key[next]; last = next; --next; }
    idx[last] = vi; key[last] = vk;
}
static void ssr_insertion_sort(int* idx, float* key, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (key[i] < key[first]) {
            const int vi = idx[i]; const float vk = key[i];
            for (int j = i; j > first; --j) { idx[j] = idx[j - 1]; key[j] = key[j - 1]; }
            idx[first] = vi; key[first] = vk;
        } else ssr_unguarded_linear_insert(idx, key, i);
    }
}
/* __adjust_heap + __push_heap (bits/stl_heap.h) on the range starting at `first` */
static void ssr_adjust_heap(int* idx, float* key, int first, int hole, int len, int vi, float vk) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (key[first + child] < key[first + child - 1]) child--;
        idx[first + hole] = idx[first + child]; key[first + hole] = key[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        idx[first + hole] = idx[first + child - 1]; key[first + hole] = key[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && key[first + parent] < vk) {
        idx[first + hole] = idx[first + parent]; key[first + hole] = key[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    idx[first + hole] = vi; key[first + hole] = vk;
}
/* __partial_sort(first, last, last): __make_heap, then __sort_heap by repeated __pop_heap */
static int g_heapsort_calls = 0;   /* test hook: how often the depth limit tripped */
int gpo_sort_replay_heapsorts(void) { return g_heapsort_calls; }
static void ssr_heapsort(int* idx, float* key, int first, int last) {
    const int len = last - first;
    ++g_heapsort_calls;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent) {
            ssr_adjust_heap(idx, key, first, parent, len, idx[first + parent], key[first + parent]);
            if (parent == 0) break;
        }
    for (int end = last; end - first > 1;) {
        --end;
        const int vi = idx[end]; const float vk = key[end];
        idx[end] = idx[first]; key[end] = key[first];
        ssr_adjust_heap(idx, key, first, 0, end - first, vi, vk);
    }
}
static void ssr_introsort_loop(int* idx, float* key, int first, int last, int depth) {
    while (last - first > 16) {
        if (depth == 0) { ssr_heapsort(idx, key, first, last); return; }
        --depth;
        /* __move_median_to_first(first, first+1, mid, last-1) */
        const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
        int med;
        if (key[a] < key[b]) { if (key[b] < key[c]) med = b; else if (key[a] < key[c]) med = c; else med = a; }
        else if (key[a] < key[c]) med = a;
        else if (key[b] < key[c]) med = c;
        else med = b;
        ssr_swap(idx, key, first, med);
        /* __unguarded_partition(first+1, last, pivot = first) */
        int lo = first + 1, hi = last;
        for (;;) {
            while (key[lo] < key[first]) ++lo;
            --hi;
            while (key[first] < key[hi]) --hi;
            if (!(lo < hi)) break;
            ssr_swap(idx, key, lo, hi);
            ++lo;
        }
        ssr_introsort_loop(idx, key, lo, last, depth);
        last = lo;
    }
}
static void std_sort_replay(int* idx, float* key, int n) {
    if (n <= 1) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    ssr_introsort_loop(idx, key, 0, n, 2 * lg);
    if (n > 16) {
        ssr_insertion_sort(idx, key, 0, 16);
        for (int i = 16; i != n; ++i) ssr_unguarded_linear_insert(idx, key, i);
    } else ssr_insertion_sort(idx, key, 0, n);
}

/* test hook: the replay on its own (idx starts as 0..n-1) */
void gpo_sort_replay(const float* keys, int n, int* idx_out) {
    float* k = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { idx_out[i] = i; k[i] = keys[i]; }
    std_sort_replay(idx_out, k, n);
    free(k);
}

/* One query: GPisMap3.cpp:803-900 / GPisMap.cpp:673-761. Geometry is always fp32 (it decides
 * the neighbour set and must be bit-exact): AABB c-l / c+l and x-h / x+h in float,
 * inclusive overlap (octree.h:128-135), sqdist as dx*dx+dy*dy+dz*dz (octree.cpp:24-31). */
static void map_test_point(const gpo_map* m, const float* x, real* res, int* chosen, int* tie, int* idx,
                           float* sq) {
    const int dim = m->dim, w = 1 + dim;
    int nc = 0;
    float qlo[3], qhi[3];
    for (int c = 0; c < dim; ++c) { qlo[c] = x[c] - m->search_half; qhi[c] = x[c] + m->search_half; }
    for (int i = 0; i < m->nc; ++i) {
        const float* ct = m->centres + (size_t)i * dim;
        int hit = 1;
        for (int c = 0; c < dim; ++c) {
            const float lo = m->boxes[(size_t)i * 2 * dim + c], hi = m->boxes[(size_t)i * 2 * dim + dim + c];
            if (qhi[c] < lo || qlo[c] > hi) { hit = 0; break; }
        }
        if (!hit) continue;
        float s = 0.f;
        for (int c = 0; c < dim; ++c) { float d = ct[c] - x[c]; s = c == 0 ? d * d : s + d * d; }
        idx[nc] = i; sq[nc] = s; ++nc;
    }
    /* std::sort over the index array with comparator sqdst[i1] < sqdst[i2] (GPisMap3.cpp:826-829):
     * replayed exactly (libstdc++ introsort), because for more than 16 candidates exact distance
     * ties are ordered by the partitioning, not by DFS order. */
    std_sort_replay(idx, sq, nc);
    const int numc = nc > 3 ? 3 : nc;
    if (chosen) {
        chosen[0] = nc;
        for (int k = 0; k < 3; ++k) chosen[1 + k] = k < numc ? idx[k] : -1;
    }
    if (tie) {
        *tie = 0;
        for (int k = 0; k < numc && k + 1 < nc; ++k)
            if (sq[k] == sq[k + 1]) *tie = 1;
    }
    res[w] = (real)(1.0 + (double)(float)m->noise); /* :816 */
    if (nc == 0) return;
    const real xt[3] = {x[0], x[1], dim == 3 ? x[2] : 0};
    if (nc == 1) { /* :818-823 */
        if (m->gps[idx[0]]) gp_test_point(m->gps[idx[0]], xt, res);
        return;
    }
    if (m->gps[idx[0]]) gp_test_point(m->gps[idx[0]], xt, res); /* :832-835 */
    if (!(res[w] > m->var_thre)) return;                          /* :837 */
    real cand[3][8];
    for (int c = 0; c < 2 * w; ++c) cand[0][c] = res[c];
    for (int k = 1; k < numc; ++k) { /* :847-854; a NULL gp there is UB in the reference */
        for (int c = 0; c < 2 * w; ++c) cand[k][c] = res[c];
        cand[k][w] = (real)1e30;
        if (m->gps[idx[k]]) gp_test_point(m->gps[idx[k]], xt, cand[k]);
    }
    int ord[3] = {0, 1, 2};
    for (int i = 1; i < numc; ++i) { /* :864-867: sort the <=3 by var_f (stable) */
        int oi = ord[i]; int j = i - 1;
        while (j >= 0 && cand[ord[j]][w] > cand[oi][w]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = oi;
    }
    const real* A = cand[ord[0]];
    if (A[w] < m->var_thre) { /* :869-880 */
        for (int c = 0; c < 2 * w; ++c) res[c] = A[c];
    } else { /* :881-895 */
        const real* B = cand[ord[1]];
        const real w1 = A[w] - m->var_thre, w2 = B[w] - m->var_thre, w12 = w1 + w2;
        for (int c = 0; c < 2 * w; ++c) res[c] = (w2 * A[c] + w1 * B[c]) / w12;
    }
}
void gpo_map_test(const gpo_map* m, const float* x, int n, real* res, int* chosen, int* tie) {
    int* idx = (int*)malloc(sizeof(int) * (m->nc > 0 ? m->nc : 1));
    float* sq = (float*)malloc(sizeof(float) * (m->nc > 0 ? m->nc : 1));
    const int w = 2 * (1 + m->dim);
    for (int i = 0; i < n; ++i)
        map_test_point(m, x + (size_t)i * m->dim, res + (size_t)i * w, chosen ? chosen + 4 * i : NULL,
                       tie ? tie + i : NULL, idx, sq);
    free(idx); free(sq);
}

/* ------------------------------------------------------------------ a7/a8: OU kernel, GPou */
typedef struct {
    int n, d;
    real* x;     /* n x d */
    real* alpha; /* n */
    real* L;     /* n x n */
} gpou;
#define OBS_SCALE 0.5f /* params.h:97 */
#define OBS_NOISE 0.01f /* params.h:98 */

/* ObsGP.cpp:32-48 + covFnc.cpp:47-68 */
static gpou* gpou_train(const real* x, const real* f, int n, int d) {
    gpou* g = (gpou*)calloc(1, sizeof(gpou));
    g->n = n; g->d = d;
    g->x = (real*)malloc(sizeof(real) * (size_t)n * d);
    memcpy(g->x, x, sizeof(real) * (size_t)n * d);
    g->alpha = (real*)malloc(sizeof(real) * n);
    memcpy(g->alpha, f, sizeof(real) * n);
    g->L = (real*)malloc(sizeof(real) * (size_t)n * n);
    const real a = (real)1 / (real)OBS_SCALE;
    for (int k = 0; k < n; ++k)
        for (int j = k; j < n; ++j) {
            real v;
            if (k == j) v = (real)(1.0 + (double)(real)OBS_NOISE);
            else v = (real)exp((double)(-a * dist(x + (size_t)k * d, x + (size_t)j * d, d)));
            g->L[(size_t)k * n + j] = g->L[(size_t)j * n + k] = v;
        }
    chol_lower(g->L, n);
    fwd_solve(g->L, n, g->alpha, 1);
    bwd_solve(g->L, n, g->alpha);
    return g;
}
static void gpou_free(gpou* g) {
    if (!g) return;
    free(g->x); free(g->alpha); free(g->L); free(g);
}
/* ObsGP.cpp:50-62 + covFnc.cpp:93-109, single test point */
static void gpou_test(const gpou* g, const real* xt, real* f, real* var) {
    const int n = g->n;
    real* k = (real*)malloc(sizeof(real) * n);
    const real a = (real)1 / (real)OBS_SCALE;
    for (int i = 0; i < n; ++i) k[i] = (real)exp((double)(-a * dist(g->x + (size_t)i * g->d, xt, g->d)));
    real s = 0;
    for (int i = 0; i < n; ++i) s += k[i] * g->alpha[i];
    *f = s;
    fwd_solve(g->L, n, k, 1);
    s = 0;
    for (int i = 0; i < n; ++i) s += k[i] * k[i];
    *var = ((real)1 + (real)OBS_NOISE) - s;
    free(k);
}

/* ------------------------------------------------------------------ a9/a10: partitioned GPs */
struct gpo_obs {
    int d;          /* 1 or 2 */
    int ng0, ng1;   /* tile grid (ng1 = 1 in 1-D) */
    int nb0, nb1;   /* boundary counts */
    int ni, nj;     /* grid the partition was computed for */
    float* b0;      /* Val_i / range */
    float* b1;      /* Val_j */
    gpou** gps;
    int ngps;
    float margin;
};
#define OBS_GROUP2 5    /* params.h:108-110 */
#define OBS_OVERLAP2 3
#define OBS_MARGIN2 0.005f
#define OBS_GROUP1 20   /* params.h:101-103 */
#define OBS_OVERLAP1 6
#define OBS_MARGIN1 0.0175f

/* ObsGP.cpp:204-265 (partition) + 280-329 (training of tiles with >= 1 valid pixel) */
gpo_obs* gpo_obs2d_retrain(const float* vu, const float* zinv, int ni, int nj, const gpo_obs* prev);
gpo_obs* gpo_obs2d_train(const float* vu, const float* zinv, int ni, int nj) {
    return gpo_obs2d_retrain(vu, zinv, ni, nj, NULL);
}
/* prev: the previous frame's regressor. ObsGP2D keeps its partition boundaries (Val_i / Val_j)
 * while the grid dimensions stay the same (ObsGP.cpp:335-337; regressObs calls the non-virtual base
 * reset(), GPisMap3.cpp:252), even if the pixel coordinates changed (resetCam). */
gpo_obs* gpo_obs2d_retrain(const float* vu, const float* zinv, int ni, int nj, const gpo_obs* prev) {
    if (ni <= 0 || nj <= 0 || !vu) return NULL;
    gpo_obs* o = (gpo_obs*)calloc(1, sizeof(gpo_obs));
    o->d = 2; o->margin = OBS_MARGIN2;
    const int g = OBS_GROUP2, ov = OBS_OVERLAP2;
    o->ng0 = (ni - ov) / g + 1;
    o->ng1 = (nj - ov) / g + 1;
    o->b0 = (float*)malloc(sizeof(float) * (o->ng0 + 1));
    o->b1 = (float*)malloc(sizeof(float) * (o->ng1 + 1));
    int* i0 = (int*)malloc(sizeof(int) * o->ng0), *i1 = (int*)malloc(sizeof(int) * o->ng0);
    int* j0 = (int*)malloc(sizeof(int) * o->ng1), *j1 = (int*)malloc(sizeof(int) * o->ng1);
    o->b0[0] = vu[0];
    for (int n = 0; n < o->ng0; ++n) {
        i0[n] = n * g;
        i1[n] = i0[n] + g + ov - 1;
        if (n < o->ng0 - 1) o->b0[n + 1] = vu[2 * (i1[n] - ov / 2)];
        else { i1[n] = ni - 1; o->b0[n + 1] = vu[2 * i1[n]]; }
    }
    o->b1[0] = vu[1];
    for (int m = 0; m < o->ng1; ++m) {
        j0[m] = m * g;
        j1[m] = j0[m] + g + ov - 1;
        if (m < o->ng1 - 1) o->b1[m + 1] = vu[2 * (j1[m] - ov / 2) * ni + 1];
        else { j1[m] = nj - 1; o->b1[m + 1] = vu[2 * j1[m] * ni + 1]; }
    }
    o->nb0 = o->ng0 + 1; o->nb1 = o->ng1 + 1;
    if (prev && prev->d == 2 && prev->ng0 == o->ng0 && prev->ng1 == o->ng1 && prev->ni == ni && prev->nj == nj) {
        memcpy(o->b0, prev->b0, sizeof(float) * o->nb0);
        memcpy(o->b1, prev->b1, sizeof(float) * o->nb1);
    }
    o->ni = ni; o->nj = nj;
    o->ngps = o->ng0 * o->ng1;
    o->gps = (gpou**)calloc(o->ngps > 0 ? o->ngps : 1, sizeof(gpou*));
    const int cap = (g + ov + 8) * (g + ov + 8) + ni + nj;
    real* xs = (real*)malloc(sizeof(real) * 2 * (size_t)cap * 4);
    real* fs = (real*)malloc(sizeof(real) * (size_t)cap * 4);
    for (int m = 0; m < o->ng1; ++m)
        for (int n = 0; n < o->ng0; ++n) {
            int cnt = 0;
            for (int j = j0[m]; j <= j1[m]; ++j)
                for (int i = i0[n]; i <= i1[n]; ++i) {
                    const int ind = j * ni + i;
                    if (zinv[ind] > 0) {
                        xs[2 * cnt] = vu[2 * ind]; xs[2 * cnt + 1] = vu[2 * ind + 1];
                        fs[cnt] = zinv[ind]; ++cnt;
                    }
                }
            if (cnt >= 1) o->gps[m * o->ng0 + n] = gpou_train(xs, fs, cnt, 2);
        }
    free(xs); free(fs); free(i0); free(i1); free(j0); free(j1);
    return o;
}

/* ObsGP.cpp:85-143 */
gpo_obs* gpo_obs1d_train(const float* theta, const float* f, int N) {
    if (N <= 0 || !theta) return NULL;
    gpo_obs* o = (gpo_obs*)calloc(1, sizeof(gpo_obs));
    o->d = 1; o->margin = OBS_MARGIN1;
    const int g = OBS_GROUP1, ov = OBS_OVERLAP1;
    const int nGroup = N / g + 1;
    o->ng0 = nGroup; o->ng1 = 1;
    o->b0 = (float*)malloc(sizeof(float) * (nGroup + 2));
    o->gps = (gpou**)calloc(nGroup + 1, sizeof(gpou*));
    real* xs = (real*)malloc(sizeof(real) * (size_t)(N + 1));
    real* fs = (real*)malloc(sizeof(real) * (size_t)(N + 1));
    for (int i = 0; i < N; ++i) { xs[i] = theta[i]; fs[i] = f[i]; }
    int nb = 0, ngp = 0;
    o->b0[nb++] = theta[0];
    for (int n = 0; n < nGroup - 1; ++n) {
        if (n < nGroup - 2) {
            const int a = n * g, b = a + g + ov;
            o->b0[nb++] = theta[b - ov / 2];
            o->gps[ngp++] = gpou_train(xs + a, fs + a, g + ov, 1);
        } else { /* the last two groups split the remainder in half (:114-136) */
            int a = n * g;
            int b = a + (N - a) / 2 + ov;
            o->b0[nb++] = theta[b - ov / 2];
            o->gps[ngp++] = gpou_train(xs + a, fs + a, b - a + 1, 1);
            ++n;
            a = a + (N - a) / 2;
            b = N - 1;
            o->b0[nb++] = theta[b];
            o->gps[ngp++] = gpou_train(xs + a, fs + a, b - a + 1, 1);
        }
    }
    o->nb0 = nb; o->nb1 = 0; o->ngps = ngp;
    free(xs); free(fs);
    return o;
}
void gpo_obs_free(gpo_obs* o) {
    if (!o) return;
    for (int i = 0; i < o->ngps; ++i) gpou_free(o->gps[i]);
    free(o->gps); free(o->b0); free(o->b1); free(o);
}
int gpo_obs_ntiles(const gpo_obs* o) { return o ? o->ngps : 0; }
void gpo_obs_tile_counts(const gpo_obs* o, int* counts) {
    for (int i = 0; i < o->ngps; ++i) counts[i] = o->gps[i] ? o->gps[i]->n : 0;
}
int gpo_obs_bounds(const gpo_obs* o, float* bi, float* bj) {
    if (bi) memcpy(bi, o->b0, sizeof(float) * o->nb0);
    if (bj && o->nb1) memcpy(bj, o->b1, sizeof(float) * o->nb1);
    return o->nb0 * 65536 + o->nb1;
}

void gpo_obs_test(const gpo_obs* o, const float* xt, int m, real* val, real* var) {
    if (!o) return; /* untrained: outputs untouched (ObsGP.cpp:147-149, 412-414) */
    for (int k = 0; k < m; ++k) {
        var[k] = (real)1e6;
        if (o->d == 2) { /* ObsGP.cpp:359-406 */
            const float x0 = xt[2 * k], x1 = xt[2 * k + 1];
            if (x0 < o->b0[0] + o->margin) continue;
            if (x0 > o->b0[o->nb0 - 1] - o->margin) continue;
            if (x1 < o->b1[0] + o->margin) continue;
            if (x1 > o->b1[o->nb1 - 1] - o->margin) continue;
            int n = 0, mm = 0;
            for (int i = 1; i < o->nb0; ++i, ++n)
                if (x0 < o->b0[i]) break;
            for (int i = 1; i < o->nb1; ++i, ++mm)
                if (x1 < o->b1[i]) break;
            const int ind = mm * o->ng0 + n;
            if (ind < o->ngps && o->gps[ind]) {
                const real p[2] = {x0, x1};
                gpou_test(o->gps[ind], p, &val[k], &var[k]);
            }
        } else { /* ObsGP.cpp:154-185 */
            const float x0 = xt[k];
            const float liml = o->b0[0] + o->margin, limr = o->b0[o->nb0 - 1] - o->margin;
            if (x0 < liml || x0 > limr) continue;
            for (int j = 0; j + 1 < o->nb0; ++j)
                if (x0 > o->b0[j] && x0 < o->b0[j + 1]) {
                    if (j < o->ngps && o->gps[j]) {
                        const real p[1] = {x0};
                        gpou_test(o->gps[j], p, &val[k], &var[k]);
                    }
                    break;
                }
        }
    }
}
