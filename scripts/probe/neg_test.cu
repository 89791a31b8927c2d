#include <cstdint>
__global__ void k(float* out, const float* in) {
    float a0 = in[threadIdx.x], a1 = in[threadIdx.x + 32], b = in[threadIdx.x + 64];
    float c0 = in[threadIdx.x + 96], c1 = in[threadIdx.x + 128];
    uint64_t aa, bb, cc;
    float nb = -b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(nb));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(cc));
    out[threadIdx.x] = c0; out[threadIdx.x + 32] = c1;
}
