// vf_policy.cu — the deterministic actor of the analytic-gradient trainers, forward and backward, one launch each.
//
// Why it is here: BASELINE configs[2] (NavigationEnv, 65 536 agents, requires_grad=True, APG) spends 8 % of its GPU
// time in the env-step kernels and the rest in the policy MLP as ~80 library launches per env step (SIMT sgemm with one
// output tile, elementwise tanh / bias / clip kernels, split-K weight gradients) — profiles/r01_s3_launches_apg.txt.
// The policy of the reference's trainers (utils/algorithms/BPTT.py:107-127 via the SB3-derived actor) is a small MLP
//     a = clip(tanh(W3 tanh(W2 tanh(W1 x + b1) + b2) + b3), -1, 1),     x (n, D <= 32),  hidden H in {32, 64},  a (n, 4)
// which is GEMM-shaped but tiny: 5.4 kFMA per agent, 21 KB of weights.  One CTA takes a tile of 128 agents, keeps the
// weights and all activations of the tile in shared memory (k-major, so both operands of every product are read with
// 128-bit loads) and runs register-tiled fp32 products on the CUDA cores (8 x H/8 accumulators per thread).  The
// backward kernel recomputes the forward tile instead of reading saved activations (nothing but the inputs is kept
// for the backward pass, as for the env-step adjoint), forms dZ3 -> dW3 -> dZ2 -> dW2 -> dZ1 -> dW1 -> dx in place,
// and writes its weight-gradient tile to a per-CTA partial; a second small kernel adds the partials in a fixed order
// (bit-reproducible, no atomics).  fp32 throughout (tensor cores would need TF32 / bf16 operands; the trainers'
// gradient parity is stated in fp32).
#include <cuda_runtime.h>

#include <string>

#include "../../include/visfly_b200.h"

namespace {

constexpr int TM = 128;      // agents per CTA
constexpr int TS = TM + 4;   // row stride of the k-major activation tiles: rows 16 bytes apart in bank space, so that
                             // 128-bit accesses to 8 or 16 DIFFERENT rows at one column are conflict-free
constexpr int NT = 128;      // threads per CTA
constexpr int DP = 32;       // padded input width
constexpr int NA = 4;        // action width

int policy_fail(const char* what, cudaError_t err = cudaSuccess);

// Thread (ty, tx) of the register-tiled products owns agents 8 ty .. 8 ty + 7 and the H/8 neurons tx, tx + 8, tx + 16 ...
// (interleaved, so that the 8 threads of a quarter warp write 8 CONSECUTIVE activation rows).  Its weights sit side by
// side in shared memory: column NJ tx + j of a weight tile holds neuron tx + 8 j.
template <int H> __host__ __device__ constexpr int neuron_of_column(int c) { return c / (H / 8) + 8 * (c % (H / 8)); }

template <int H> struct Smem {
    float x[DP][TS];         // inputs, k-major
    float h1[H][TS];         // layer-1 activations, row = neuron (backward: overwritten by dZ1)
    float h2[H][TS];         // layer-2 activations (backward: overwritten by dZ2)
    float w1t[DP][H];        // w1t[k][c] = W1[neuron(c)][k]
    float w2t[H][H];         // w2t[k][c] = W2[neuron(c)][k]
    float w3t[H][NA];        // w3t[k][o] = W3[o][k]
    float b1[H], b2[H], b3[NA];   // b1 / b2 by column
};
template <int H> struct SmemBwd : Smem<H> {
    float w1n[H][DP];        // w1n[j][k] = W1[j][k]
    float w2n[H][H];         // w2n[j][c] = W2[j][neuron(c)]   (k = j, columns permuted like the forward tiles)
    float w3n[NA][H];
    float dz3[NA][TS];
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// weights + input tile -> shared memory
template <int H, class S>
__device__ void load_tile(S& s, int n, int d, int first, const float* __restrict__ x, const float* __restrict__ w1,
                          const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                          const float* __restrict__ w3, const float* __restrict__ b3) {
    const int t = threadIdx.x;
    for (int e = t; e < H * DP; e += NT) {               // W1 (H, d) row-major -> w1t[k][c]
        const int c = e / DP, k = e % DP;
        s.w1t[k][c] = k < d ? __ldg(w1 + neuron_of_column<H>(c) * d + k) : 0.f;
    }
    for (int e = t; e < H * H; e += NT) {
        const int c = e / H, k = e % H;
        s.w2t[k][c] = __ldg(w2 + neuron_of_column<H>(c) * H + k);
    }
    for (int e = t; e < NA * H; e += NT) {
        const int o = e / H, k = e % H;
        s.w3t[k][o] = __ldg(w3 + e);
    }
    for (int e = t; e < H; e += NT) {
        s.b1[e] = __ldg(b1 + neuron_of_column<H>(e));
        s.b2[e] = __ldg(b2 + neuron_of_column<H>(e));
    }
    if (t < NA) s.b3[t] = __ldg(b3 + t);
    const int a = first + t;                             // one agent per thread: d contiguous floats
    for (int k = 0; k < DP; ++k) s.x[k][t] = (a < n && k < d) ? __ldg(x + size_t(a) * d + k) : 0.f;
}

// acc[8 agents][H/8 columns] = A[TM x K] B[K x H] for this thread's tile; A k-major [K][TS], B k-major [K][H].
template <int H, int K>
__device__ __forceinline__ void tile_gemm(const float (*A)[TS], const float (*B)[H], float acc[8][H / 8]) {
    constexpr int NJ = H / 8;
    const int ty = threadIdx.x / 8, tx = threadIdx.x % 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = lds4(&A[k][8 * ty]), a1 = lds4(&A[k][8 * ty + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j += 4) {
            const float4 b = lds4(&B[k][NJ * tx + j]);
            bv[j] = b.x; bv[j + 1] = b.y; bv[j + 2] = b.z; bv[j + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// forward of one tile: fills s.h1, s.h2 (tanh activations); returns this thread's agent's action in a[4]
template <int H, class S> __device__ void forward_tile(S& s, float a[NA]) {
    constexpr int NJ = H / 8;
    const int ty = threadIdx.x / 8, tx = threadIdx.x % 8;
    float acc[8][NJ];
    tile_gemm<H, DP>(s.x, s.w1t, acc);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float b = s.b1[NJ * tx + j];
        float4 lo = make_float4(tanhf(acc[0][j] + b), tanhf(acc[1][j] + b), tanhf(acc[2][j] + b), tanhf(acc[3][j] + b));
        float4 hi = make_float4(tanhf(acc[4][j] + b), tanhf(acc[5][j] + b), tanhf(acc[6][j] + b), tanhf(acc[7][j] + b));
        *reinterpret_cast<float4*>(&s.h1[tx + 8 * j][8 * ty]) = lo;
        *reinterpret_cast<float4*>(&s.h1[tx + 8 * j][8 * ty + 4]) = hi;
    }
    __syncthreads();
    tile_gemm<H, H>(s.h1, s.w2t, acc);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float b = s.b2[NJ * tx + j];
        float4 lo = make_float4(tanhf(acc[0][j] + b), tanhf(acc[1][j] + b), tanhf(acc[2][j] + b), tanhf(acc[3][j] + b));
        float4 hi = make_float4(tanhf(acc[4][j] + b), tanhf(acc[5][j] + b), tanhf(acc[6][j] + b), tanhf(acc[7][j] + b));
        *reinterpret_cast<float4*>(&s.h2[tx + 8 * j][8 * ty]) = lo;
        *reinterpret_cast<float4*>(&s.h2[tx + 8 * j][8 * ty + 4]) = hi;
    }
    __syncthreads();
    // output layer: one agent per thread
    float mu[NA] = {s.b3[0], s.b3[1], s.b3[2], s.b3[3]};
    const int t = threadIdx.x;
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
        const float h = s.h2[k][t];
        const float4 w = lds4(&s.w3t[k][0]);
        mu[0] = fmaf(h, w.x, mu[0]); mu[1] = fmaf(h, w.y, mu[1]); mu[2] = fmaf(h, w.z, mu[2]); mu[3] = fmaf(h, w.w, mu[3]);
    }
#pragma unroll
    for (int o = 0; o < NA; ++o) a[o] = tanhf(mu[o]);
}

template <int H>
__global__ void __launch_bounds__(NT)
vf_policy_fwd_kernel(int n, int d, const float* __restrict__ x, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     const float* __restrict__ w3, const float* __restrict__ b3, float lo, float hi,
                     float* __restrict__ action) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<H>& s = *reinterpret_cast<Smem<H>*>(smem_raw);
    const int first = blockIdx.x * TM;
    load_tile<H>(s, n, d, first, x, w1, b1, w2, b2, w3, b3);
    __syncthreads();
    float a[NA];
    forward_tile<H>(s, a);
    const int agent = first + threadIdx.x;
    if (agent < n) {
        // th.clip(actions, low, high) of the trainers (BPTT.py:113-115): NaN passes through like torch.clamp
#pragma unroll
        for (int o = 0; o < NA; ++o) a[o] = a[o] < lo ? lo : (a[o] > hi ? hi : a[o]);
        reinterpret_cast<float4*>(action)[agent] = make_float4(a[0], a[1], a[2], a[3]);
    }
}

// per-CTA partial weight gradients, in this order: dW1 (H, d) | db1 (H) | dW2 (H, H) | db2 (H) | dW3 (4, H) | db3 (4)
__host__ __device__ inline int partial_size(int h, int d) { return h * d + h + h * h + h + NA * h + NA; }

template <int H>
__global__ void __launch_bounds__(NT)
vf_policy_bwd_kernel(int n, int d, const float* __restrict__ x, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     const float* __restrict__ w3, const float* __restrict__ b3, float lo, float hi,
                     const float* __restrict__ g_action, float* __restrict__ g_x, float* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemBwd<H>& s = *reinterpret_cast<SmemBwd<H>*>(smem_raw);
    constexpr int NJ = H / 8;
    const int t = threadIdx.x;
    const int first = blockIdx.x * TM;
    load_tile<H>(s, n, d, first, x, w1, b1, w2, b2, w3, b3);
    for (int e = t; e < H * DP; e += NT) {
        const int j = e / DP, k = e % DP;
        s.w1n[j][k] = k < d ? __ldg(w1 + j * d + k) : 0.f;
    }
    for (int e = t; e < H * H; e += NT) s.w2n[e / H][e % H] = __ldg(w2 + (e / H) * H + neuron_of_column<H>(e % H));
    for (int e = t; e < NA * H; e += NT) s.w3n[e / H][e % H] = __ldg(w3 + e);
    __syncthreads();
    float a[NA];
    forward_tile<H>(s, a);
    // ---- output layer: dz3 = g_a * clip'(a) * (1 - a^2) --------------------------------------------------------
    {
        const int agent = first + t;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (agent < n) g = __ldg(reinterpret_cast<const float4*>(g_action) + agent);
        const float gv[NA] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int o = 0; o < NA; ++o) {
            const float gate = (a[o] >= lo && a[o] <= hi) ? 1.f : 0.f;      // torch.clamp: closed interval
            s.dz3[o][t] = gv[o] * gate * (1.f - a[o] * a[o]);
        }
    }
    __syncthreads();
    float* out = partial + size_t(blockIdx.x) * partial_size(H, d);
    float* p_w1 = out;
    float* p_b1 = p_w1 + H * d;
    float* p_w2 = p_b1 + H;
    float* p_b2 = p_w2 + H * H;
    float* p_w3 = p_b2 + H;
    float* p_b3 = p_w3 + NA * H;
    // ---- dW3[o][k] = sum_a dz3[o][a] h2[k][a],  db3 ---------------------------------------------------------------
    for (int e = t; e < NA * H; e += NT) {
        const int o = e / H, k = e % H;
        float acc = 0.f;
        for (int a4 = 0; a4 < TM; a4 += 4) {
            const float4 z = lds4(&s.dz3[o][a4]), h = lds4(&s.h2[k][a4]);
            acc += z.x * h.x + z.y * h.y + z.z * h.z + z.w * h.w;
        }
        p_w3[e] = acc;
    }
    if (t < NA) {
        float acc = 0.f;
        for (int a1 = 0; a1 < TM; ++a1) acc += s.dz3[t][a1];
        p_b3[t] = acc;
    }
    __syncthreads();
    // ---- dZ2 = (dz3 W3) * (1 - h2^2), in place over h2 ----------------------------------------------------------
    {
        const float z0 = s.dz3[0][t], z1 = s.dz3[1][t], z2 = s.dz3[2][t], z3 = s.dz3[3][t];
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float h = s.h2[k][t];
            const float dh = z0 * s.w3n[0][k] + z1 * s.w3n[1][k] + z2 * s.w3n[2][k] + z3 * s.w3n[3][k];
            s.h2[k][t] = dh * (1.f - h * h);
        }
    }
    __syncthreads();
    // ---- dW2[j][i] = sum_a dz2[j][a] h1[i][a]: thread block 8 (j) x H/16 (i),  db2 -------------------------------
    {
        constexpr int JB = H / 8, IB = H / 16;
        const int tj = t / 16, ti = t % 16;
        float acc[JB][IB];
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int i = 0; i < IB; ++i) acc[j][i] = 0.f;
        for (int a4 = 0; a4 < TM; a4 += 4) {              // rows interleaved over the threads: conflict-free 128-bit loads
            float4 zv[JB], hv[IB];
#pragma unroll
            for (int j = 0; j < JB; ++j) zv[j] = lds4(&s.h2[tj + 8 * j][a4]);
#pragma unroll
            for (int i = 0; i < IB; ++i) hv[i] = lds4(&s.h1[ti + 16 * i][a4]);
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
                for (int i = 0; i < IB; ++i)
                    acc[j][i] += zv[j].x * hv[i].x + zv[j].y * hv[i].y + zv[j].z * hv[i].z + zv[j].w * hv[i].w;
        }
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int i = 0; i < IB; ++i) p_w2[(tj + 8 * j) * H + ti + 16 * i] = acc[j][i];
        if (t < H) {
            float b = 0.f;
            for (int a4 = 0; a4 < TM; a4 += 4) {
                const float4 z = lds4(&s.h2[t][a4]);
                b += z.x + z.y + z.z + z.w;
            }
            p_b2[t] = b;
        }
    }
    // ---- dZ1 = (dZ2 W2) * (1 - h1^2), in place over h1 (each thread owns its 8 x NJ block) -----------------------
    {
        const int ty = t / 8, tx = t % 8;
        float acc[8][NJ];
        tile_gemm<H, H>(s.h2, s.w2n, acc);
        __syncthreads();                                   // every thread is done reading h1 (dW2) before it changes
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float* row = &s.h1[tx + 8 * j][8 * ty];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = row[i];
                row[i] = acc[i][j] * (1.f - h * h);
            }
        }
    }
    __syncthreads();
    // ---- dW1[j][k] = sum_a dz1[j][a] x[k][a]: thread block H/8 (j) x 2 (k),  db1 ---------------------------------
    {
        constexpr int JB = H / 8;
        const int tj = t / 16, ti = t % 16;
        float acc[JB][2];
#pragma unroll
        for (int j = 0; j < JB; ++j) acc[j][0] = acc[j][1] = 0.f;
        for (int a4 = 0; a4 < TM; a4 += 4) {
            const float4 x0 = lds4(&s.x[ti][a4]), x1 = lds4(&s.x[ti + 16][a4]);
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const float4 z = lds4(&s.h1[tj + 8 * j][a4]);
                acc[j][0] += z.x * x0.x + z.y * x0.y + z.z * x0.z + z.w * x0.w;
                acc[j][1] += z.x * x1.x + z.y * x1.y + z.z * x1.z + z.w * x1.w;
            }
        }
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (ti + 16 * c < d) p_w1[(tj + 8 * j) * d + ti + 16 * c] = acc[j][c];
        if (t < H) {
            float b = 0.f;
            for (int a4 = 0; a4 < TM; a4 += 4) {
                const float4 z = lds4(&s.h1[t][a4]);
                b += z.x + z.y + z.z + z.w;
            }
            p_b1[t] = b;
        }
    }
    // ---- dx[a][k] = sum_j dz1[j][a] W1[j][k]: one agent per thread -------------------------------------------------
    if (g_x) {
        float acc[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) acc[k] = 0.f;
#pragma unroll 2
        for (int j = 0; j < H; ++j) {
            const float z = s.h1[j][t];
#pragma unroll
            for (int k = 0; k < DP; k += 4) {
                const float4 w = lds4(&s.w1n[j][k]);
                acc[k] = fmaf(z, w.x, acc[k]); acc[k + 1] = fmaf(z, w.y, acc[k + 1]);
                acc[k + 2] = fmaf(z, w.z, acc[k + 2]); acc[k + 3] = fmaf(z, w.w, acc[k + 3]);
            }
        }
        const int agent = first + t;
        if (agent < n) {
            float* dst = g_x + size_t(agent) * d;
#pragma unroll
            for (int k = 0; k < DP; ++k)
                if (k < d) dst[k] = acc[k];
        }
    }
}

// out[e] = sum over CTAs of partial[c][e], fixed order
__global__ void vf_policy_reduce_kernel(int ctas, int size, const float* __restrict__ partial, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= size) return;
    float acc = 0.f;
    for (int c = 0; c < ctas; ++c) acc += __ldg(partial + size_t(c) * size + e);
    out[e] = acc;
}

thread_local std::string g_policy_error;
int policy_fail(const char* what, cudaError_t err) {
    g_policy_error = what;
    if (err != cudaSuccess) {
        g_policy_error += ": ";
        g_policy_error += cudaGetErrorString(err);
    }
    return 1;
}

template <class K> int allow_smem(K kernel, size_t bytes) {
    const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    return err == cudaSuccess ? 0 : policy_fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed", err);
}

int check_shapes(int n, int d, int h) {
    if (n < 0) return policy_fail("n must be >= 0");
    if (d < 1 || d > DP) return policy_fail("policy input width must be in 1..32");
    if (h != 32 && h != 64) return policy_fail("policy hidden width must be 32 or 64");
    return 0;
}

}  // namespace

extern "C" {

const char* vf_policy_last_error(void) { return g_policy_error.c_str(); }

int vf_policy_partial_floats(int n, int d, int h) { return ((n + TM - 1) / TM) * partial_size(h, d); }

int vf_policy_fwd(int n, int d, int h, const float* x, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* w3, const float* b3, float lo, float hi, float* action, void* stream) {
    if (check_shapes(n, d, h)) return 1;
    if (n == 0) return 0;
    if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !action) return policy_fail("vf_policy_fwd: NULL buffer");
    if (reinterpret_cast<size_t>(action) & 15u) return policy_fail("action must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (n + TM - 1) / TM;
    if (h == 64) {
        if (allow_smem(vf_policy_fwd_kernel<64>, sizeof(Smem<64>))) return 1;
        vf_policy_fwd_kernel<64><<<grid, NT, sizeof(Smem<64>), st>>>(n, d, x, w1, b1, w2, b2, w3, b3, lo, hi, action);
    } else {
        if (allow_smem(vf_policy_fwd_kernel<32>, sizeof(Smem<32>))) return 1;
        vf_policy_fwd_kernel<32><<<grid, NT, sizeof(Smem<32>), st>>>(n, d, x, w1, b1, w2, b2, w3, b3, lo, hi, action);
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : policy_fail("vf_policy_fwd launch failed", err);
}

int vf_policy_bwd(int n, int d, int h, const float* x, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* w3, const float* b3, float lo, float hi, const float* grad_action,
                  float* grad_x, float* partial, float* grad_params, void* stream) {
    if (check_shapes(n, d, h)) return 1;
    if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !grad_action || !partial || !grad_params)
        return policy_fail("vf_policy_bwd: NULL buffer");
    if (reinterpret_cast<size_t>(grad_action) & 15u) return policy_fail("grad_action must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (n + TM - 1) / TM;
    const int size = partial_size(h, d);
    if (n > 0) {
        if (h == 64) {
            if (allow_smem(vf_policy_bwd_kernel<64>, sizeof(SmemBwd<64>))) return 1;
            vf_policy_bwd_kernel<64><<<grid, NT, sizeof(SmemBwd<64>), st>>>(n, d, x, w1, b1, w2, b2, w3, b3, lo, hi,
                                                                          grad_action, grad_x, partial);
        } else {
            if (allow_smem(vf_policy_bwd_kernel<32>, sizeof(SmemBwd<32>))) return 1;
            vf_policy_bwd_kernel<32><<<grid, NT, sizeof(SmemBwd<32>), st>>>(n, d, x, w1, b1, w2, b2, w3, b3, lo, hi,
                                                                          grad_action, grad_x, partial);
        }
    }
    vf_policy_reduce_kernel<<<(size + 127) / 128, 128, 0, st>>>(grid, size, partial, grad_params);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : policy_fail("vf_policy_bwd launch failed", err);
}

}  // extern "C"
