// Micro-probe: FFMA issue rate per SM for a 16x8 register tile, 3-reg FFMA vs packed fma.rn.f32x2,
// as a function of warps per scheduler. Operands are register-resident (no memory traffic).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE, int TR, int MAXT>
__global__ void __launch_bounds__(MAXT) k_probe(float* out, int iters, float seed) {
    float acc[TR][8];
    float a[TR], b[8];
#pragma unroll
    for (int i = 0; i < TR; ++i) a[i] = seed * (i + 1) + threadIdx.x * 1e-9f;
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = seed * (j + 3);
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        } else if (MODE == 1) {
            // pairs along rows: (acc[i][j], acc[i+1][j]) += (a[i], a[i+1]) * (b[j], b[j])
            uint64_t bb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) asm("mov.b64 %0, {%1, %1};" : "=l"(bb[j]) : "f"(b[j]));
#pragma unroll
            for (int i = 0; i < TR; i += 2) {
                uint64_t aa;
                asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a[i]), "f"(a[i + 1]));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint64_t c;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[i][j]), "f"(acc[i + 1][j]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(aa), "l"(bb[j]));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][j]), "=f"(acc[i + 1][j]) : "l"(c));
                }
            }
        }
        // perturb operands a bit so nothing is hoisted
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] += 1e-7f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int TR, int MAXT>
void run(const char* name, int threads) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    const int iters = 20000;
    k_probe<MODE, TR, MAXT><<<sms, threads>>>(out, 100, 1e-3f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<MODE, TR, MAXT><<<sms, threads>>>(out, iters, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * threads * iters * (TR * 8.0);
    const double tf = 2.0 * fma / (ms * 1e-3) / 1e12;
    const double peak = 2.0 * sms * 128.0 * khz * 1e3 / 1e12;
    printf("%-11s threads=%3d  %.3f ms  %.1f TFLOP/s  (%.1f%% of %d SMs x128 x2 x %.0f MHz = %.1f; err=%s)\n", name, threads, ms, tf,
           100.0 * tf / peak, sms, khz / 1e3, peak, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0, 16, 256>("ffma 16x8", 128); run<0, 16, 256>("ffma 16x8", 256);
    run<0, 8, 512>("ffma 8x8", 128); run<0, 8, 512>("ffma 8x8", 256); run<0, 8, 512>("ffma 8x8", 512);
    run<1, 16, 256>("ffma2 16x8", 128); run<1, 16, 256>("ffma2 16x8", 256);
    run<1, 8, 512>("ffma2 8x8", 128); run<1, 8, 512>("ffma2 8x8", 256); run<1, 8, 512>("ffma2 8x8", 512);
    return 0;
}
