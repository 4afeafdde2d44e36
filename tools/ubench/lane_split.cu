// Experiment: does splitting ONE agent over FOUR lanes — one quaternion / body-rate component per
// lane, operands exchanged with __shfl_sync, warp-shuffle renormalisation, 4x the warps at the headline size — beat
// one agent per thread?  The RK4 attitude sub-step (utils/maths.py:353-386 with frozen torque: 4 stages of
// q' = 1/2 q (x) (0,w), w' = J^-1 tau - g (w x w terms), then renormalisation) is 60 % of the control step's
// instructions and is the part that parallelises best over components, so it is the most favourable case.
//   A: one agent per thread   (what vf_math.cuh::attitude_fwd does)
//   B: four lanes per agent   (lane c owns q[c], w[c]; each stage fetches its 6 operands with lane-indexed shuffles)
// Both run `substeps` sub-steps on n agents; results are compared, then timed at 65 536 and 4 194 304 agents.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/lane_split tools/ubench/lane_split.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <vector>

struct Coef { float jt[3], g[3], h; };

__device__ __forceinline__ float rsq(float x) { float y = rsqrtf(x); return y * (1.5f - 0.5f * x * y * y); }

__global__ void __launch_bounds__(64) k_thread(int n, int substeps, Coef c, const float4* __restrict__ q_in,
                                               const float4* __restrict__ w_in, float4* __restrict__ q_out,
                                               float4* __restrict__ w_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 Q = q_in[i], W = w_in[i];
    float q[4] = {Q.x, Q.y, Q.z, Q.w}, w[3] = {W.x, W.y, W.z};
    const float h = c.h, hh = 0.5f * h, qh = 0.5f * h, qhh = 0.25f * h, s6 = 1.f / 6.f, s3 = 2.f / 6.f;
    for (int s = 0; s < substeps; ++s) {
        float k[4][3], g[4][4], qs[4], ws[3];
#pragma unroll
        for (int j = 0; j < 4; ++j) qs[j] = q[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) ws[j] = w[j];
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            k[st][0] = c.jt[0] - c.g[0] * (ws[1] * ws[2]);
            k[st][1] = c.jt[1] - c.g[1] * (ws[2] * ws[0]);
            k[st][2] = c.jt[2] - c.g[2] * (ws[0] * ws[1]);
            g[st][0] = -(qs[1] * ws[0] + qs[2] * ws[1] + qs[3] * ws[2]);
            g[st][1] = qs[0] * ws[0] + qs[2] * ws[2] - qs[3] * ws[1];
            g[st][2] = qs[0] * ws[1] - qs[1] * ws[2] + qs[3] * ws[0];
            g[st][3] = qs[0] * ws[2] + qs[1] * ws[1] - qs[2] * ws[0];
            if (st < 3) {
                const float a = st == 2 ? h : hh, b = st == 2 ? qh : qhh;
#pragma unroll
                for (int j = 0; j < 3; ++j) ws[j] = w[j] + a * k[st][j];
#pragma unroll
                for (int j = 0; j < 4; ++j) qs[j] = q[j] + b * g[st][j];
            }
        }
        float qn[4];
#pragma unroll
        for (int j = 0; j < 3; ++j) w[j] += h * (s6 * k[0][j] + s3 * k[1][j] + s3 * k[2][j] + s6 * k[3][j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) qn[j] = q[j] + qh * (s6 * g[0][j] + s3 * g[1][j] + s3 * g[2][j] + s6 * g[3][j]);
        const float inv = rsq(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = qn[j] * inv;
    }
    q_out[i] = make_float4(q[0], q[1], q[2], q[3]);
    w_out[i] = make_float4(w[0], w[1], w[2], 0.f);
}

// lane c of an agent's quad owns q[c] and w[c] (c = 3: no body rate)
__global__ void __launch_bounds__(256) k_lanes(int n, int substeps, Coef c, const float* __restrict__ q_in,
                                               const float* __restrict__ w_in, float* __restrict__ q_out,
                                               float* __restrict__ w_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int agent = t >> 2, l = t & 3, base = threadIdx.x & 28 & 31;   // first lane of the quad inside the warp
    if (agent >= n) return;
    float q = q_in[t], w = w_in[t];
    // operands of component l of q (x) (0,w): g_l = s0 q[a0] w[b0] + s1 q[a1] w[b1] + s2 q[a2] w[b2]
    const int a0 = l == 0 ? 1 : 0, a1 = l == 0 ? 2 : (l == 1 ? 2 : (l == 2 ? 1 : 1)), a2 = l == 3 ? 2 : 3;
    const int b0 = l == 0 ? 0 : l - 1, b1 = l == 0 ? 1 : (l == 1 ? 2 : (l == 2 ? 2 : 1)), b2 = l == 0 ? 2 : (l == 1 ? 1 : 0);
    const float s0 = l == 0 ? -1.f : 1.f, s1 = (l == 0 || l == 2) ? -1.f : 1.f, s2 = (l == 0 || l == 1 || l == 3) ? -1.f : 1.f;
    const float jt = l < 3 ? c.jt[l] : 0.f, gc = l < 3 ? c.g[l] : 0.f;
    const int c1 = (l + 1) % 3, c2 = (l + 2) % 3;
    const float h = c.h, hh = 0.5f * h, qh = 0.5f * h, qhh = 0.25f * h, s6 = 1.f / 6.f, s3 = 2.f / 6.f;
    const unsigned full = 0xffffffffu;
    for (int s = 0; s < substeps; ++s) {
        float qs = q, ws = w, ksum = 0.f, gsum = 0.f;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            const float qa0 = __shfl_sync(full, qs, base + a0), qa1 = __shfl_sync(full, qs, base + a1),
                        qa2 = __shfl_sync(full, qs, base + a2);
            const float wb0 = __shfl_sync(full, ws, base + b0), wb1 = __shfl_sync(full, ws, base + b1),
                        wb2 = __shfl_sync(full, ws, base + b2);
            const float wc1 = __shfl_sync(full, ws, base + c1), wc2 = __shfl_sync(full, ws, base + c2);
            const float g = (s0 * qa0) * wb0 + (s1 * qa1) * wb1 + (s2 * qa2) * wb2;
            const float k = jt - gc * (wc1 * wc2);
            const float wt = (st == 0 || st == 3) ? s6 : s3;
            ksum += wt * k;
            gsum += wt * g;
            if (st < 3) {
                ws = w + (st == 2 ? h : hh) * k;
                qs = q + (st == 2 ? qh : qhh) * g;
            }
        }
        w += h * ksum;
        const float qn = q + qh * gsum;
        float n2 = qn * qn;
        n2 += __shfl_xor_sync(full, n2, 1);
        n2 += __shfl_xor_sync(full, n2, 2);
        q = qn * rsq(n2);
    }
    q_out[t] = q;
    w_out[t] = w;
}

template <class F> float time_us(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e3f / reps;
}

int main() {
    Coef c = {{0.8f, -0.5f, 0.3f}, {0.35f, -0.4f, 0.05f}, 0.0025f};
    for (int n : {65536, 1 << 22}) {
        std::vector<float> q(4 * n), w(4 * n);
        for (int i = 0; i < n; ++i) {
            float a = 0.3f * sinf(i * 0.37f), b = 0.2f * cosf(i * 0.11f), d = 0.1f * sinf(i * 0.05f);
            float nn = sqrtf(1 + a * a + b * b + d * d);
            q[4 * i] = 1 / nn; q[4 * i + 1] = a / nn; q[4 * i + 2] = b / nn; q[4 * i + 3] = d / nn;
            w[4 * i] = sinf(i * 0.7f); w[4 * i + 1] = cosf(i * 0.3f); w[4 * i + 2] = 0.5f * sinf(i * 0.9f); w[4 * i + 3] = 0.f;
        }
        float *dq, *dw, *q1, *w1, *q2, *w2;
        size_t bytes = sizeof(float) * 4 * n;
        cudaMalloc(&dq, bytes); cudaMalloc(&dw, bytes); cudaMalloc(&q1, bytes); cudaMalloc(&w1, bytes);
        cudaMalloc(&q2, bytes); cudaMalloc(&w2, bytes);
        cudaMemcpy(dq, q.data(), bytes, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), bytes, cudaMemcpyHostToDevice);
        const int S = 8;
        auto fa = [&] { k_thread<<<(n + 63) / 64, 64>>>(n, S, c, (float4*)dq, (float4*)dw, (float4*)q1, (float4*)w1); };
        auto fb = [&] { k_lanes<<<(4 * n + 255) / 256, 256>>>(n, S, c, dq, dw, q2, w2); };
        fa(); fb(); cudaDeviceSynchronize();
        std::vector<float> r1(4 * n), r2(4 * n);
        cudaMemcpy(r1.data(), q1, bytes, cudaMemcpyDeviceToHost); cudaMemcpy(r2.data(), q2, bytes, cudaMemcpyDeviceToHost);
        float err = 0; for (int i = 0; i < 4 * n; ++i) err = fmaxf(err, fabsf(r1[i] - r2[i]));
        const int reps = n > 100000 ? 20 : 200;
        const float ta = time_us(fa, reps), tb = time_us(fb, reps);
        printf("n=%8d  RK4 attitude x%d sub-steps:  one agent per thread %8.2f us (%5d warps)   four lanes per agent %8.2f us "
               "(%6d warps)   ratio %.2f   max |dq| = %.2e   [%s]\n", n, S, ta, n / 32, tb, n / 8, tb / ta, err, cudaGetErrorString(cudaGetLastError()));
        cudaFree(dq); cudaFree(dw); cudaFree(q1); cudaFree(w1); cudaFree(q2); cudaFree(w2);
    }
    return 0;
}
