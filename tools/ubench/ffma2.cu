// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate and dependent-chain latency on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
template <int ILP> __global__ void k_scalar(float* out, int iters, float b, float c) {
    float a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = ffma1(a[j], b, c);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_packed(float* out, int iters, float b, float c) {
    u64 a[ILP];
    float2 bb = make_float2(b, b), cc = make_float2(c, c);
    u64 B = *reinterpret_cast<u64*>(&bb), C = *reinterpret_cast<u64*>(&cc);
#pragma unroll
    for (int j = 0; j < ILP; ++j) { float2 t = make_float2(threadIdx.x + j, threadIdx.x - j); a[j] = *reinterpret_cast<u64*>(&t); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = ffma2(a[j], B, C);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { float2 t = *reinterpret_cast<float2*>(&a[j]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 64 * 1024 * sizeof(float));
    int sms = 148, iters = 4096;
    // warps per SM sweep: block of 128 threads (1 warp/SMSP) x blocks per SM
    for (int wps : {1, 2, 4, 8}) {           // warps per scheduler
        int threads = 128, blocks = sms * wps;
        auto rate = [&](float ms, double fma_per_thread_iter) { return blocks * (double)threads * iters * fma_per_thread_iter / (ms * 1e-3) / 1e12; };
        float m1 = time_ms([&] { k_scalar<1><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float m4 = time_ms([&] { k_scalar<4><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float m8 = time_ms([&] { k_scalar<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float p1 = time_ms([&] { k_packed<1><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float p4 = time_ms([&] { k_packed<4><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float p8 = time_ms([&] { k_packed<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        printf("warps/SMSP=%d  scalar TFMA/s ilp1 %.2f ilp4 %.2f ilp8 %.2f | packed TFMA/s ilp1 %.2f ilp4 %.2f ilp8 %.2f\n", wps,
               rate(m1, 1), rate(m4, 4), rate(m8, 8), rate(p1, 2), rate(p4, 8), rate(p8, 16));
        printf("              ms: %.4f %.4f %.4f | %.4f %.4f %.4f   (dependent-chain ns/instr: scalar %.2f packed %.2f)\n", m1, m4, m8, p1, p4, p8,
               m1 * 1e6 / iters, p1 * 1e6 / iters);
    }
    return 0;
}
