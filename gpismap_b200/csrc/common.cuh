// Shared device/host definitions for the B200 GPisMap hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GPIS_TILE 32            // leaf systems are stored as 32x32 tiles
#define GPIS_TILE_ELEMS 1024
#define GPIS_TILE_BYTES 4096

namespace gpis {

// ------------------------------------------------------------------ leaf record (device arena)
// One trained leaf GP = one contiguous, 128-byte aligned record:
//   [ header 128 B | pts: N x float4 (x,y,z|0, gradidx bits) | alpha: nb*32 floats |
//     dinv: nb tiles (inverse of the diagonal Cholesky blocks) | tiles: nb(nb+1)/2 tiles of L ]
// n = N + dim*ng unknowns are padded to nb*32 with an identity block (padded alpha = 0).
// Tiles are column-block-major: for block column bk, tiles bi = bk..nb-1 are consecutive, so the
// slab a factorization step streams is one contiguous range (TMA bulk-copy friendly). Inside a
// tile, element (r, k) lives at k*32 + r ("k-major"): a thread that needs 4 consecutive rows at
// one k issues a single 16-byte shared-memory load.
struct LeafHeader {
    int32_t N, ng, n, nb;
    int32_t dim, chol_fail, slot, pad0;
    uint64_t bytes;
    uint64_t key;
    int32_t cell[4];
    float centre[4];
    float lo[4], hi[4];   // effective candidate box (see gpis_leaves_set_boxes)
    uint8_t pad[128 - 32 - 16 - 16 - 16 - 32];
};
static_assert(sizeof(LeafHeader) == 128, "header must be 128 bytes");

__host__ __device__ inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline uint64_t rec_off_pts() { return 128; }
__host__ __device__ inline uint64_t rec_off_alpha(int N) { return 128 + align_up((uint64_t)N * 16, 128); }
__host__ __device__ inline uint64_t rec_off_dinv(int N, int nb) { return rec_off_alpha(N) + (uint64_t)nb * 128; }
__host__ __device__ inline uint64_t rec_off_tiles(int N, int nb) {
    return rec_off_dinv(N, nb) + (uint64_t)nb * GPIS_TILE_BYTES;
}
__host__ __device__ inline uint64_t rec_bytes(int N, int nb) {
    return rec_off_tiles(N, nb) + (uint64_t)nb * (nb + 1) / 2 * GPIS_TILE_BYTES;
}
// index of tile (bi, bk), bi >= bk, in the column-block-major tile array
__host__ __device__ inline int tile_index(int bi, int bk, int nb) { return bk * nb - bk * (bk - 1) / 2 + (bi - bk); }

// ------------------------------------------------------------------ leaf table (K3)
// Open-addressing hash keyed by the Morton code of the leaf's lattice cell (bias keeps the
// coordinates non-negative). Values are slot indices into the SoA slot arrays.
#define GPIS_CELL_BIAS (1 << 20)
#define GPIS_KEY_EMPTY 0ull
#define GPIS_KEY_TOMB 0xFFFFFFFFFFFFFFFFull

__host__ __device__ inline uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
// Morton key of a cell; +1 so that 0 can mean "empty". 2-D cells pass iz = 0.
__host__ __device__ inline uint64_t cell_key(int ix, int iy, int iz) {
    return (spread3((uint64_t)(ix + GPIS_CELL_BIAS)) | (spread3((uint64_t)(iy + GPIS_CELL_BIAS)) << 1) |
            (spread3((uint64_t)(iz + GPIS_CELL_BIAS)) << 2)) + 1ull;
}
__host__ __device__ inline uint32_t hash_key(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

struct LeafTable {       // device pointers, passed to kernels by value
    uint64_t* keys;      // cap
    int32_t* vals;       // cap
    uint32_t cap_mask;   // cap - 1 (cap is a power of two)
    // slot arrays
    float4* centre;      // xyz = centre used verbatim in the distance test, w = unused
    float4* lo;          // effective float box of the leaf (own box ∩ ancestors' boxes, see gpis_leaves_set_boxes)
    float4* hi;
    int4* cell;          // lattice cell (w: 1 = live, 0 = dead)
    uint64_t* rec;       // device address of the trained record, 0 = untrained
    int4* meta;          // N, ng, n, nb
};

__device__ inline int table_find(const LeafTable& t, uint64_t key) {
    uint32_t h = hash_key(key) & t.cap_mask;
    for (uint32_t probe = 0; probe <= t.cap_mask; ++probe) {
        const uint64_t k = __ldg(t.keys + h);
        if (k == key) return __ldg(t.vals + h);
        if (k == GPIS_KEY_EMPTY) return -1;
        h = (h + 1) & t.cap_mask;
    }
    return -1;
}

// ------------------------------------------------------------------ query parameters
struct QueryParams {
    int dim;
    float cluster_half, search_half, var_thre;
    float var_preset;   // (float)(1.0 + map_noise)            GPisMap3.cpp:816
    float a;            // (float)(sqrt(3)/scale)              covFnc.cpp:263
    double prior_f;     // 3D: 1.001, 2D: 1.01 (double literals) OnGPIS.cpp:203,235
    double prior_g;     // three_over_scale + 0.001 / + 0.1    OnGPIS.cpp:205-212,236-237
    double inv_pitch;   // 1 / (2*cluster_half)
    int root_min[3];    // root box for DFS tie-breaks (gpis_rebase)
    int levels;
};

// ------------------------------------------------------------------ exp in double-float arithmetic
// The reference evaluates exp() in double and rounds the final product to float (covFnc.cpp:29-33, SURVEY
// §3.3). B200's vector FP64 rate is a small fraction of its FP32 rate (profiles/r01: the double exp made the
// covariance build the longest phase of leaf training), so exp and the products are carried out in
// double-float (hi + lo, ~48 significant bits) on the FP32 pipe instead: error-free transforms with explicit
// FMAs, relative error < 1e-11 (scripts/dfexp_check.py), which rounds to the same float as the double path
// except when the exact value lies within ~2^-36 of a rounding boundary.
struct DF { float hi, lo; };
__device__ __forceinline__ DF df_two_sum(float a, float b) {
    const float s = a + b, bb = s - a;
    return DF{s, (a - (s - bb)) + (b - bb)};
}
__device__ __forceinline__ DF df_fast_two_sum(float a, float b) {   // |a| >= |b|
    const float s = a + b;
    return DF{s, b - (s - a)};
}
__device__ __forceinline__ DF df_mul(DF x, DF y) {
    const float p = x.hi * y.hi;
    float e = __fmaf_rn(x.hi, y.hi, -p);
    e = e + (x.hi * y.lo + x.lo * y.hi);
    return df_fast_two_sum(p, e);
}
__device__ __forceinline__ DF df_add(DF x, DF y) {
    DF s = df_two_sum(x.hi, y.hi);
    return df_fast_two_sum(s.hi, s.lo + (x.lo + y.lo));
}
// exp(x) for a float x in [-80, 0.4]: x = k ln2 + r, |r| <= 0.35; Taylor through r^5 in double-float, tail in float
__device__ __forceinline__ DF exp_df(float x) {
    const float k = rintf(x * 1.44269504f);
    const float LN2_HI = 0.693145751953125f;        // 0x1.62e4p-1: k * LN2_HI is exact for |k| < 2^9
    const float LN2_MID = 1.42860677e-06f;          // 0x1.7f7d1cp-20
    const float LN2_LO = 5.49560397e-14f;           // ln2 - LN2_HI - LN2_MID
    const float r1 = __fmaf_rn(-k, LN2_HI, x);
    const float p = k * LN2_MID, pe = __fmaf_rn(k, LN2_MID, -p);
    DF r = df_two_sum(r1, -p);
    r = df_fast_two_sum(r.hi, r.lo - (pe + k * LN2_LO));
    const float t = __fmaf_rn(__fmaf_rn(__fmaf_rn(2.75573192e-06f, r.hi, 2.48015873e-05f), r.hi, 1.98412698e-04f), r.hi, 1.38888889e-03f);
    DF u = df_two_sum(8.33333377e-03f, r.hi * t);   // 1/120 = 0x1.111112p-7 - 0x1.dddddep-32
    u.lo += -4.34617203e-10f;
    u = df_add(df_mul(u, r), DF{4.16666679e-02f, -1.24176347e-09f});   // 1/24
    u = df_add(df_mul(u, r), DF{1.66666672e-01f, -4.96705388e-09f});   // 1/6
    u = df_add(df_mul(u, r), DF{0.5f, 0.f});
    u = df_add(df_mul(u, r), DF{1.0f, 0.f});
    u = df_add(df_mul(u, r), DF{1.0f, 0.f});
    const float sc = __int_as_float(((int)k + 127) << 23);   // 2^k, exact scaling
    return DF{u.hi * sc, u.lo * sc};
}
__device__ __forceinline__ float df_round(DF v) { return v.hi + v.lo; }
// (float)((double)c * e)
__device__ __forceinline__ float df_mulf_round(float c, DF e) {
    const float p = e.hi * c;
    float err = __fmaf_rn(e.hi, c, -p);
    err = __fmaf_rn(e.lo, c, err);
    return p + err;
}

// ------------------------------------------------------------------ Matern-3/2 pieces
// Same mixed precision as the reference (covFnc.cpp:29-33 with SURVEY §3.3): products in float, exp and the
// final product beyond float precision, rounded once to float. -fmad=false keeps the float products unfused
// like the reference's SSE2 build. e = exp_df(-a*r).
__device__ __forceinline__ float kf_val(float r, float a, DF e) {
    const DF s = df_two_sum(1.0f, a * r);          // 1.0 + (double)(a*r), exact
    return df_round(df_mul(s, e));
}
__device__ __forceinline__ float kf1_val(float dx, float a, DF e) { return df_mulf_round(a * a * dx, e); }
__device__ __forceinline__ float kf2_val(float r, float dx1, float dx2, float delta, float a, DF e) {
    return df_mulf_round(a * a * (delta - a * dx1 * dx2 / r), e);
}

// ------------------------------------------------------------------ PTX helpers (TMA bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {   // required before the same location is initialised again
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Packed fp32x2 FMA (sm_100 FFMA2): two IEEE fma.rn results per issue slot, bit-identical to two fmaf().
// (c0, c1) -= (a0, a1) * b   and   (c0, c1) += (a0, a1) * b. ptxas folds the pack/unpack moves when the
// pairs sit in aligned adjacent registers (they do when they come from 8/16-byte loads) and the negation
// into the instruction's broadcast-operand modifier.
#ifndef GPIS_NO_FFMA2
__device__ __forceinline__ void fma2_sub(float& c0, float& c1, float a0, float a1, float b) {
    uint64_t aa, bb, cc;
    const float nb = -b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(nb));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(cc));
}
__device__ __forceinline__ void fma2_add(float& c0, float& c1, float a0, float a1, float b) {
    uint64_t aa, bb, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(cc));
}
#else
__device__ __forceinline__ void fma2_sub(float& c0, float& c1, float a0, float a1, float b) { c0 = fmaf(-a0, b, c0); c1 = fmaf(-a1, b, c1); }
__device__ __forceinline__ void fma2_add(float& c0, float& c1, float a0, float a1, float b) { c0 = fmaf(a0, b, c0); c1 = fmaf(a1, b, c1); }
#endif

}  // namespace gpis
