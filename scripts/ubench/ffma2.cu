// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE> __global__ void k(float *out, int iters) {
    float s = threadIdx.x * 1e-3f;
    if (MODE == 0) {
        float a[16]; for (int i = 0; i < 16; ++i) a[i] = s + i;
        const float b = 1.0000001f, c = 1e-7f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {
        unsigned long long a[8]; for (int i = 0; i < 8; ++i) { float2 v = make_float2(s + i, s - i); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
        float2 bb = make_float2(1.0000001f, 0.9999999f), cc = make_float2(1e-7f, 2e-7f);
        unsigned long long b = *reinterpret_cast<unsigned long long*>(&bb), c = *reinterpret_cast<unsigned long long*>(&cc);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], b, c);
        float r = 0; for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); r += v.x + v.y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}
int main() {
    int sms = 148, blocks = sms * 8, threads = 256, iters = 4096;
    float *d; cudaMalloc(&d, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(d, iters); else k<1><<<blocks, threads>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = 128.0 * iters * blocks * threads;          // scalar FMAs per launch (both modes do 128 per thread-iter)
            double winst = (mode == 0 ? 128.0 : 64.0) * iters * blocks * threads / 32;
            if (rep == 3) printf("%s: %.3f ms  %.2f TFLOP/s  %.3f warp-instr/clk/SM (at 1.965 GHz)\n", mode == 0 ? "FFMA " : "FFMA2", ms,
                                 2 * fmas / ms / 1e9, winst / (ms * 1e-3) / 1.965e9 / sms);
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
