// vf_sensors.cu — ingestion half of the renderer hand-off: what the reference does to the images
// a renderer hands back, per agent and per sensor, in numpy on the host (envs/base/droneEnv.py:296-331:
// stack -> expand_dims / transpose -> np.where(depth == 0, 20, depth)), as two HBM-bound byte/float kernels that read a
// batched image buffer — device memory, or the page-locked host buffer a renderer wrote (zero-copy over PCIe) — and
// write the observation tensors in the layout the policies' feature extractors take.
//   depth / semantic : (n, H, W)       -> (n, 1, H, W), depth background 0 -> `background`   (:303-305, :310-311)
//   color            : (n, H, W, 4) u8 -> (n, 3, H, W) u8, alpha dropped                     (:306-308)
// Pure streaming: 128-bit loads and stores, grid = a multiple of the SM count, no reuse.
#include <cuda_runtime.h>

#include <string>

#include "../../include/visfly_b200.h"

namespace {

thread_local std::string g_sensor_error;
int sensor_fail(const char* what, cudaError_t err = cudaSuccess) {
    g_sensor_error = what;
    if (err != cudaSuccess) {
        g_sensor_error += ": ";
        g_sensor_error += cudaGetErrorString(err);
    }
    return 1;
}

__global__ void __launch_bounds__(256)
vf_ingest_depth_kernel(size_t n4, size_t tail_start, size_t total, const float* __restrict__ src,
                       float* __restrict__ dst, float background) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        v.x = v.x == 0.f ? background : v.x;
        v.y = v.y == 0.f ? background : v.y;
        v.z = v.z == 0.f ? background : v.z;
        v.w = v.w == 0.f ? background : v.w;
        reinterpret_cast<float4*>(dst)[i] = v;
    }
    for (size_t i = tail_start + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float v = src[i];
        dst[i] = v == 0.f ? background : v;
    }
}

// one thread = 4 consecutive pixels of one image: 16 B in, 4 B to each of the three colour planes
__global__ void __launch_bounds__(256)
vf_ingest_color_kernel(size_t images, size_t hw, const unsigned char* __restrict__ src, unsigned char* __restrict__ dst) {
    const size_t quads_per_image = hw / 4, quads = images * quads_per_image;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < quads; q += stride) {
        const size_t img = q / quads_per_image, px = (q % quads_per_image) * 4;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (img * hw + px) * 4));
        const unsigned p[4] = {v.x, v.y, v.z, v.w};
        unsigned r = 0, g = 0, b = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r |= (p[k] & 0xFFu) << (8 * k);
            g |= ((p[k] >> 8) & 0xFFu) << (8 * k);
            b |= ((p[k] >> 16) & 0xFFu) << (8 * k);
        }
        unsigned char* out = dst + img * 3 * hw + px;
        *reinterpret_cast<unsigned*>(out) = r;
        *reinterpret_cast<unsigned*>(out + hw) = g;
        *reinterpret_cast<unsigned*>(out + 2 * hw) = b;
    }
}

// device pointer of a buffer that is device memory or page-locked host memory
int resolve(const void* p, const void** out, const char* what) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return sensor_fail(what);
    }
    if (attr.type == cudaMemoryTypeHost) *out = attr.devicePointer;
    else if (attr.type == cudaMemoryTypeUnregistered) return sensor_fail(what);
    else *out = p;
    return 0;
}

int grid_for(size_t work_items) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t want = (work_items + 255) / 256, cap = size_t(sms) * 8;       // 8 resident CTAs of 256 per SM
    return int(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

extern "C" {

const char* vf_sensor_last_error(void) { return g_sensor_error.c_str(); }

int vf_ingest_depth(long long n_images, int height, int width, const float* src, float* dst, float background,
                    void* stream) {
    if (n_images < 0 || height < 1 || width < 1) return sensor_fail("vf_ingest_depth: bad shape");
    const size_t total = size_t(n_images) * height * width;
    if (total == 0) return 0;
    if (!src || !dst) return sensor_fail("vf_ingest_depth: NULL buffer");
    const void* s = nullptr;
    if (resolve(src, &s, "vf_ingest_depth: src must be device memory or page-locked host memory")) return 1;
    if ((reinterpret_cast<size_t>(s) | reinterpret_cast<size_t>(dst)) & 15u)
        return sensor_fail("vf_ingest_depth: buffers must be 16-byte aligned");
    const size_t n4 = total / 4;
    vf_ingest_depth_kernel<<<grid_for(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n4, n4 * 4, total, static_cast<const float*>(s), dst, background);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : sensor_fail("vf_ingest_depth launch failed", err);
}

int vf_ingest_color(long long n_images, int height, int width, const unsigned char* src, unsigned char* dst,
                    void* stream) {
    if (n_images < 0 || height < 1 || width < 1) return sensor_fail("vf_ingest_color: bad shape");
    const size_t hw = size_t(height) * width;
    if (n_images == 0) return 0;
    if (hw % 4) return sensor_fail("vf_ingest_color: height * width must be a multiple of 4");
    if (!src || !dst) return sensor_fail("vf_ingest_color: NULL buffer");
    const void* s = nullptr;
    if (resolve(src, &s, "vf_ingest_color: src must be device memory or page-locked host memory")) return 1;
    if ((reinterpret_cast<size_t>(s) & 15u) || (reinterpret_cast<size_t>(dst) & 3u))
        return sensor_fail("vf_ingest_color: src must be 16-byte aligned, dst 4-byte aligned");
    vf_ingest_color_kernel<<<grid_for(size_t(n_images) * hw / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        size_t(n_images), hw, static_cast<const unsigned char*>(s), dst);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : sensor_fail("vf_ingest_color launch failed", err);
}

}  // extern "C"
