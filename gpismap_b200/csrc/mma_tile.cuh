// Register-tile micro-kernels of the evaluation kernel (query_v3.cuh): a warp updates up to E3_R stacked
// 32-row tiles against one shared right-hand operand,
// 16x8 accumulators per lane when all four slots are active. Operands come from shared memory (k-major
// 32-wide tiles, 8 k per call), FMAs are packed FFMA2, operand loads are software pipelined over k.
#pragma once
#include "common.cuh"

namespace gpis {

#ifndef E3_WARPS
#define E3_WARPS 8
#endif
#ifndef E3_R
#define E3_R 4                      // block rows per warp per wave
#endif
#define E3_THREADS (E3_WARPS * 32)
#define E3_WAVE (E3_WARPS * E3_R)   // block rows per wave
#define E3_STAGE_FLOATS (E3_R * 256)   // one stage: E3_R quarter tiles of 1 KB

// A lane's CPL columns of a row of U <-> registers. CPL = 8 / 4: columns CPL*cg .. CPL*cg + CPL-1, one or two 16-byte
// accesses. CPL = 6 (24-column rows): columns 4cg .. 4cg+3 and 16+2cg, 17+2cg — one 16-byte and one 8-byte access, both
// aligned, instead of three 8-byte ones at 24-byte strides (6 instead of 7 shared-memory instructions per k-step of the
// 16x6 lane tile). The column a register stands for only matters where results leave the tile: lane_col().
template <int CPL>
__device__ __forceinline__ int lane_col(int cg, int j) {
    if constexpr (CPL == 6) return j < 4 ? 4 * cg + j : 16 + 2 * cg + (j - 4);
    else return CPL * cg + j;
}
template <int CPL>
__device__ __forceinline__ void ld_cols(const float* __restrict__ row, int cg, float (&v)[CPL]) {
    if constexpr (CPL == 6) {
        const float4 a = *reinterpret_cast<const float4*>(row + 4 * cg);
        const float2 b = *reinterpret_cast<const float2*>(row + 16 + 2 * cg);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
    } else {
        const float* p = row + CPL * cg;
#pragma unroll
        for (int t = 0; t < CPL / 4; ++t) {
            const float4 b = *reinterpret_cast<const float4*>(p + 4 * t);
            v[4 * t] = b.x; v[4 * t + 1] = b.y; v[4 * t + 2] = b.z; v[4 * t + 3] = b.w;
        }
    }
}
template <int CPL>
__device__ __forceinline__ void st_cols(float* __restrict__ row, int cg, const float (&v)[CPL]) {
    if constexpr (CPL == 6) {
        *reinterpret_cast<float4*>(row + 4 * cg) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float2*>(row + 16 + 2 * cg) = make_float2(v[4], v[5]);
    } else {
        float* p = row + CPL * cg;
#pragma unroll
        for (int t = 0; t < CPL / 4; ++t) *reinterpret_cast<float4*>(p + 4 * t) = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
    }
}

// acc[s][i][j] -= sum_{k<8} A_s[k][4rg+i] * B[k][CPL*cg+j] for accumulator slots s < R;  A_s: quarter tile
// [8][32] at As + s*256, B: [8][NCOL] rows of U. Slots are ordered from the warp's LAST block row of the
// wave upwards, so the rows still active at a column are always a prefix and R is the only variant.
template <int R, int CPL, int NCOL>
__device__ __forceinline__ void qmma_sub(float (&acc)[E3_R][4][CPL], const float* __restrict__ As,
                                         const float* __restrict__ Bq, int rg, int cg) {
    // software pipelined over k: the operands of step k+1 are loaded before the FMAs of step k are issued, so
    // the shared-memory latency is covered by this warp's own FMAs (two register operand buffers)
    const float* Ap = As + 4 * rg;
    const float* Bp = Bq;
    float bv[2][CPL];
    float4 av[2][R];
    ld_cols<CPL>(Bp, cg, bv[0]);
#pragma unroll
    for (int r = 0; r < R; ++r) av[0][r] = *reinterpret_cast<const float4*>(Ap + r * 256);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int cu = k & 1, nx = cu ^ 1;
        if (k + 1 < 8) {
            ld_cols<CPL>(Bp + (k + 1) * NCOL, cg, bv[nx]);
#pragma unroll
            for (int r = 0; r < R; ++r) av[nx][r] = *reinterpret_cast<const float4*>(Ap + r * 256 + (k + 1) * 32);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float a4[4] = {av[cu][r].x, av[cu][r].y, av[cu][r].z, av[cu][r].w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < CPL; j += 2) fma2_sub(acc[r][i][j], acc[r][i][j + 1], bv[cu][j], bv[cu][j + 1], a4[i]);
        }
    }
}
// one accumulator slot S only (the lookahead part of a split visit), KS consecutive k (8 = a quarter tile,
// 32 = the whole tile in one software pipeline)
template <int S, int CPL, int NCOL, int KS = 8>
__device__ __forceinline__ void qmma_one(float (&acc)[E3_R][4][CPL], const float* __restrict__ As,
                                         const float* __restrict__ Bq, int rg, int cg) {
    const float* Ap = As + S * 256 + 4 * rg;
    const float* Bp = Bq;
    float bv[2][CPL];
    float4 av[2];
    ld_cols<CPL>(Bp, cg, bv[0]);
    av[0] = *reinterpret_cast<const float4*>(Ap);
#pragma unroll 8
    for (int k = 0; k < KS; ++k) {
        const int cu = k & 1, nx = cu ^ 1;
        if (k + 1 < KS) {
            ld_cols<CPL>(Bp + (k + 1) * NCOL, cg, bv[nx]);
            av[nx] = *reinterpret_cast<const float4*>(Ap + (k + 1) * 32);
        }
        const float a4[4] = {av[cu].x, av[cu].y, av[cu].z, av[cu].w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CPL; j += 2) fma2_sub(acc[S][i][j], acc[S][i][j + 1], bv[cu][j], bv[cu][j + 1], a4[i]);
    }
}
template <int CPL, int NCOL>
__device__ __forceinline__ void qmma_dispatch_one(int slot, float (&acc)[E3_R][4][CPL], const float* As, const float* Bq, int rg, int cg) {
    switch (slot) {
#if E3_R >= 4
        case 3: qmma_one<3, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
#if E3_R >= 3
        case 2: qmma_one<2, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
#if E3_R >= 2
        case 1: qmma_one<1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
        default: qmma_one<0, CPL, NCOL>(acc, As, Bq, rg, cg); break;
    }
}
template <int CPL, int NCOL>
__device__ __forceinline__ void qmma_dispatch(int nact, float (&acc)[E3_R][4][CPL], const float* As, const float* Bq, int rg, int cg) {
    switch (nact) {
#if E3_R >= 4
        case 4: qmma_sub<4, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
#if E3_R >= 3
        case 3: qmma_sub<3, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
#if E3_R >= 2
        case 2: qmma_sub<2, CPL, NCOL>(acc, As, Bq, rg, cg); break;
#endif
        case 1: qmma_sub<1, CPL, NCOL>(acc, As, Bq, rg, cg); break;
        default: break;
    }
}



}  // namespace gpis
