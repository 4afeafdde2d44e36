// Micro-benchmark: device -> page-locked host, 3.9 MB (one env step's obs/reward/done at 65 536 agents):
//   (a) copy engine (cudaMemcpyAsync), (b) zero-copy stores from SMs: 16 B per lane, 128 B-aligned rows,
//   (c) the same with a 16 B-misaligned base (what a warp-private 1664 B obs block looks like), (d) 4 B per lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pcie_d2h tools/ubench/pcie_d2h.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void copy16(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) dst[i] = src[i];
}
__global__ void copy4(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) dst[i] = src[i];
}
template <typename F> float time_us(F f, int reps = 20) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best * 1e3f;
}
int main() {
    const size_t bytes = 65536 * 60;   // 3 932 160
    float *d, *h;
    cudaMalloc(&d, bytes + 256); cudaHostAlloc(&h, bytes + 256, cudaHostAllocDefault);
    cudaMemset(d, 1, bytes + 256);
    float* hd; cudaHostGetDevicePointer((void**)&hd, h, 0);
    auto report = [&](const char* what, float us) { printf("%-58s %8.2f us  %6.2f GB/s\n", what, us, bytes / us * 1e-3); };
    report("copy engine, one cudaMemcpyAsync", time_us([&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, 0); }));
    report("copy engine, 4 chunks", time_us([&] { for (int c = 0; c < 4; ++c) cudaMemcpyAsync((char*)h + c * (bytes / 4), (char*)d + c * (bytes / 4), bytes / 4, cudaMemcpyDeviceToHost, 0); }));
    for (int grid : {148, 592, 1024}) {
        char buf[96];
        snprintf(buf, sizeof buf, "zero-copy float4 stores, aligned, grid %d x 64", grid);
        report(buf, time_us([&] { copy16<<<grid, 64>>>((const float4*)d, (float4*)hd, bytes / 16); }));
        snprintf(buf, sizeof buf, "zero-copy float4 stores, base +16 B, grid %d x 64", grid);
        report(buf, time_us([&] { copy16<<<grid, 64>>>((const float4*)d + 1, (float4*)hd + 1, bytes / 16); }));
    }
    report("zero-copy float stores (4 B per lane), grid 1024 x 64", time_us([&] { copy4<<<1024, 64>>>(d, hd, bytes / 4); }));
    return 0;
}
