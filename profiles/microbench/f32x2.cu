// Issue cost of Blackwell's packed FP32 instructions (FFMA2 / FMUL2 / FADD2) against their scalar forms.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run: ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = a[i] + 0.5f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {                  // 2 scalar FFMA
                a[i] = __fmaf_rn(a[i], s, 1.0f); b[i] = __fmaf_rn(b[i], s, 1.0f);
            } else if (MODE == 1) {           // 1 FFMA2
                u64 x, y, z, r;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(b[i]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(s), "f"(s));
                asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(1.0f), "f"(1.0f));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(r));
            } else if (MODE == 2) {           // 2 scalar FMUL
                a[i] = __fmul_rn(a[i], s); b[i] = __fmul_rn(b[i], s);
            } else {                          // 1 FMUL2
                u64 x, y, r;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(b[i]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(s), "f"(s));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(r));
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 4, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 64, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops_pairs = (double)blocks * threads * iters * 8;        // pair-operations
    printf("%-14s %8.3f ms   %7.2f G pair-ops/s  (%.1f pair-ops / clk / SM at 1.9 GHz)\n", name, ms, flops_pairs / ms * 1e-6,
           flops_pairs / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 512 * sizeof(float));
    run<0>("2x FFMA", d); run<1>("FFMA2", d); run<2>("2x FMUL", d); run<3>("FMUL2", d);
    return 0;
}
