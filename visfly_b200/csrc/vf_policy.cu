// vf_policy.cu — the deterministic actor of the analytic-gradient trainers, forward and backward, one launch each.
//
// Two kernel families behind vf_policy_fwd / vf_policy_bwd: the tensor-core kernels of vf_policy_tc.cuh (tcgen05,
// accumulators in tensor memory, 3xTF32 — the default wherever they apply) and the CUDA-core kernels of this file
// (VF_POLICY_NO_TC=1, and the backward for observation widths above 16).
//
// Why it is here: BASELINE configs[2] (NavigationEnv, 65 536 agents, requires_grad=True, APG) spends 8 % of its GPU
// time in the env-step kernels and the rest in the policy MLP as ~80 library launches per env step (SIMT sgemm with one
// output tile, elementwise tanh / bias / clip kernels, split-K weight gradients) — profiles/r01_s3_launches_apg.txt.
// The policy of the reference's trainers (utils/algorithms/BPTT.py:107-127 via the SB3-derived actor) is a small MLP
//     a = clip(tanh(W3 tanh(W2 tanh(W1 x + b1) + b2) + b3), lo, hi),    x (n, D <= 32),  hidden H in {32, 64},  a (n, 4)
// which is GEMM-shaped but tiny: 5.4 kFMA per agent, 21 KB of weights.  One CTA of 256 threads takes a tile of 128
// agents, keeps the weights and all activations of the tile in shared memory (k-major, so both operands of every
// product are read with 128-bit loads) and runs register-tiled fp32 products on the CUDA cores (8 agents x H/16
// neurons per thread).  The observation may arrive in two pieces (e.g. NavigationEnv: state (n,13) and target (n,3)),
// so no concatenated copy is ever made.  The backward kernel recomputes the forward tile instead of reading saved
// activations (nothing but the inputs is kept for the backward pass, as for the env-step adjoint), forms
// dZ3 -> dW3 -> dZ2 -> dW2 -> dZ1 -> dW1 -> dx in place, and writes its weight-gradient tile to a per-CTA partial; a
// second small kernel adds the partials in a fixed order (bit-reproducible, no atomics).  fp32 throughout (tensor
// cores would need TF32 / bf16 operands; the trainers' gradient parity is stated in fp32).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <set>
#include <string>
#include <utility>

#include "../../include/visfly_b200.h"

namespace {

constexpr int TM = 128;      // agents per CTA
constexpr int TS = TM + 4;   // row stride of the k-major activation tiles: rows 16 bytes apart in bank space, so that
                             // 128-bit accesses to 8 or 16 DIFFERENT rows at one column are conflict-free
constexpr int NT = 256;      // threads per CTA
constexpr int DP = 32;       // padded input width
constexpr int NA = 4;        // action width

int policy_fail(const char* what, cudaError_t err = cudaSuccess);

// Thread (ty, tx), ty = t / 16, tx = t % 16, of the register-tiled products owns agents 8 ty .. 8 ty + 7 and the N/16
// columns tx, tx + 16, tx + 32 ... of an N-column result (interleaved, so that the threads of a quarter warp write
// CONSECUTIVE activation rows).  Its weights sit side by side in shared memory: position NJ tx + j of a weight row
// holds column tx + 16 j.
template <int N> __host__ __device__ constexpr int column_at(int c) { return c / (N / 16) + 16 * (c % (N / 16)); }

// tanh from one ex2 and one reciprocal, (1 - e) / (1 + e) with e = exp(-2|x|), and its Taylor polynomial where that
// form cancels (|x| < 0.04): relative error ~2e-7 over the whole range.  The library tanhf is ~35 instructions, and
// the network evaluates 132 of them per agent.
__device__ __forceinline__ float fast_tanh(float x) {
    const float ax = fabsf(x);
    const float e = exp2f(-2.885390081777927f * ax);             // exp(-2|x|)
    const float big = __fdividef(1.f - e, 1.f + e);
    const float x2 = ax * ax;
    const float small = ax * fmaf(x2, fmaf(x2, 0.13333333f, -0.33333334f), 1.f);
    return copysignf(ax < 0.04f ? small : big, x);
}

// The weights in the layouts the tiles want, as ONE block: vf_policy_pack writes it to global memory once per weight
// update, and every CTA of the forward / backward kernels copies it into shared memory with straight 128-bit loads
// (transposing 21 KB per CTA on the fly cost as much as the products: 32-way bank conflicts on the strided stores).
template <int H> struct PackedFwd {
    float w1t[DP][H];        // w1t[k][c] = W1[column_at(c)][k]   (rows k >= d are zero)
    float w2t[H][H];         // w2t[k][c] = W2[column_at(c)][k]
    float w3t[H][NA];        // w3t[k][o] = W3[o][k]
    float b1[H], b2[H], b3[NA];   // b1 / b2 in weight-row order
};
template <int H> struct PackedBwd {
    float w1n[H][DP];        // w1n[j][c] = W1[j][column_at<DP>(c)]   (columns >= d are zero)
    float w2n[H][H];         // w2n[j][c] = W2[j][column_at<H>(c)]
    float w3n[NA][H];
};
}  // namespace
// per-CTA partial weight gradients, in this order: dW1 (H, d) | db1 (H) | dW2 (H, H) | db2 (H) | dW3 (4, H) | db3 (4)
__host__ __device__ inline int partial_size(int h, int d) { return h * d + h + h * h + h + 4 * h + 4; }
namespace {
#include "vf_policy_tc.cuh"

template <int H> struct Packed {
    PackedFwd<H> f;
    PackedBwd<H> b;
    tc::PackedTc<H> t;
};
template <int H> struct Smem : PackedFwd<H> {
    float x[DP][TS];         // inputs, k-major (backward: finally the dx tile)
    float h1[H][TS];         // layer-1 activations, row = neuron (backward: overwritten by dZ1)
    float h2[H][TS];         // layer-2 activations (backward: overwritten by dZ2)
};
// backward: the forward-layout weights are dead once the tile has been recomputed, so the natural-layout block is
// copied over them (137 KB -> 112 KB of shared memory: two CTAs per SM instead of one)
template <int H> struct SmemBwd : Smem<H> {
    float dz3[NA][TS];
};
static_assert(sizeof(PackedBwd<64>) <= sizeof(PackedFwd<64>) && sizeof(PackedBwd<32>) <= sizeof(PackedFwd<32>),
              "the backward weight block reuses the forward block's shared memory");
static_assert(sizeof(PackedFwd<64>) % 16 == 0 && sizeof(PackedBwd<64>) % 16 == 0 && sizeof(PackedFwd<32>) % 16 == 0 &&
              sizeof(PackedBwd<32>) % 16 == 0, "packed blocks are copied with 128-bit accesses");

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// one thread per element of the packed block (run once per weight update)
template <int H>
__global__ void vf_policy_pack_kernel(int d, const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                      const float* __restrict__ w3, const float* __restrict__ b3,
                                      Packed<H>* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    Packed<H>& p = *out;
    if (e < DP * H) {
        const int k = e / H, c = e % H;
        p.f.w1t[k][c] = k < d ? w1[column_at<H>(c) * d + k] : 0.f;
        const int j = e / DP, cc = e % DP, kk = column_at<DP>(cc);
        p.b.w1n[j][cc] = kk < d ? w1[j * d + kk] : 0.f;
    }
    if (e < H * H) {
        const int k = e / H, c = e % H;
        p.f.w2t[k][c] = w2[column_at<H>(c) * H + k];
        p.b.w2n[k][c] = w2[k * H + column_at<H>(c)];
    }
    if (e < NA * H) {
        p.f.w3t[e % H][e / H] = w3[e];
        p.b.w3n[e / H][e % H] = w3[e];
    }
    if (e < H) {
        p.f.b1[e] = b1[column_at<H>(e)];
        p.f.b2[e] = b2[column_at<H>(e)];
    }
    if (e < NA) p.f.b3[e] = b3[e];
    tc::pack_tc<H>(e, d, w1, b1, w2, b2, w3, b3, p.t);
}

template <class T> __device__ __forceinline__ void copy_block(T& dst, const T& src) {
    const float4* g = reinterpret_cast<const float4*>(&src);
    float4* s4 = reinterpret_cast<float4*>(&dst);
    for (int i = threadIdx.x; i < int(sizeof(T) / 16); i += NT) s4[i] = __ldg(g + i);
}

// input tile -> shared memory.  x = [xa (n, da) | xb (n, db)] row-major pieces, d = da + db.
template <class S>
__device__ void load_inputs(S& s, int n, int da, int db, int first, const float* __restrict__ xa,
                            const float* __restrict__ xb) {
    const int t = threadIdx.x, d = da + db;
    const int la = t % TM, a = first + la;               // two threads per agent, each takes every other feature
    for (int k = t / TM; k < DP; k += NT / TM) {
        float v = 0.f;
        if (a < n) {
            if (k < da) v = __ldg(xa + size_t(a) * da + k);
            else if (k < d) v = __ldg(xb + size_t(a) * db + (k - da));
        }
        s.x[k][la] = v;
    }
}

// acc[8 agents][N/16 columns] = A[TM x K] B[K x N] for this thread's tile; A k-major [K][TS], B k-major [K][N].
template <int N, int K>
__device__ __forceinline__ void tile_gemm(const float (*A)[TS], const float (*B)[N], float acc[8][N / 16]) {
    constexpr int NJ = N / 16;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a0 = lds4(&A[k][8 * ty]), a1 = lds4(&A[k][8 * ty + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        static_assert(NJ == 4 || NJ == 2, "tile widths: 64 or 32 columns");
        float bv[NJ];
        if constexpr (NJ == 4) {
            const float4 b = lds4(&B[k][NJ * tx]);
            bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
        } else {
            const float2 b = *reinterpret_cast<const float2*>(&B[k][NJ * tx]);
            bv[0] = b.x; bv[1] = b.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// rows[column tx + 16 j][agents 8 ty ..] = f(acc + bias)
template <int N, bool TANH>
__device__ __forceinline__ void store_tile(float (*dst)[TS], const float acc[8][N / 16], const float* bias) {
    constexpr int NJ = N / 16;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float b = bias ? bias[NJ * tx + j] : 0.f;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = TANH ? fast_tanh(acc[i][j] + b) : acc[i][j] + b;
        *reinterpret_cast<float4*>(&dst[tx + 16 * j][8 * ty]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(&dst[tx + 16 * j][8 * ty + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// forward of one tile: fills s.h1, s.h2 (tanh activations); threads 0..TM-1 return their agent's action in a[4]
template <int H, class S> __device__ void forward_tile(S& s, float a[NA]) {
    float acc[8][H / 16];
    tile_gemm<H, DP>(s.x, s.w1t, acc);
    store_tile<H, true>(s.h1, acc, s.b1);
    __syncthreads();
    tile_gemm<H, H>(s.h1, s.w2t, acc);
    store_tile<H, true>(s.h2, acc, s.b2);
    __syncthreads();
    const int t = threadIdx.x;
    if (t < TM) {                                        // output layer: one agent per thread
        float mu[NA] = {s.b3[0], s.b3[1], s.b3[2], s.b3[3]};
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float h = s.h2[k][t];
            const float4 w = lds4(&s.w3t[k][0]);
            mu[0] = fmaf(h, w.x, mu[0]); mu[1] = fmaf(h, w.y, mu[1]);
            mu[2] = fmaf(h, w.z, mu[2]); mu[3] = fmaf(h, w.w, mu[3]);
        }
#pragma unroll
        for (int o = 0; o < NA; ++o) a[o] = fast_tanh(mu[o]);
    }
}

template <int H>
__global__ void __launch_bounds__(NT)
vf_policy_fwd_kernel(int n, int da, int db, const float* __restrict__ xa, const float* __restrict__ xb,
                     const Packed<H>* __restrict__ packed, float lo, float hi, float* __restrict__ action) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<H>& s = *reinterpret_cast<Smem<H>*>(smem_raw);
    const int first = blockIdx.x * TM;
    copy_block<PackedFwd<H>>(s, packed->f);
    load_inputs(s, n, da, db, first, xa, xb);
    __syncthreads();
    float a[NA];
    forward_tile<H>(s, a);
    const int agent = first + threadIdx.x;
    if (threadIdx.x < TM && agent < n) {
        // th.clip(actions, low, high) of the trainers (BPTT.py:113-115): NaN passes through like torch.clamp
#pragma unroll
        for (int o = 0; o < NA; ++o) a[o] = a[o] < lo ? lo : (a[o] > hi ? hi : a[o]);
        reinterpret_cast<float4*>(action)[agent] = make_float4(a[0], a[1], a[2], a[3]);
    }
}

template <int H>
__global__ void __launch_bounds__(NT)
vf_policy_bwd_kernel(int n, int da, int db, const float* __restrict__ xa, const float* __restrict__ xb,
                     const Packed<H>* __restrict__ packed, float lo, float hi, const float* __restrict__ g_action,
                     float* __restrict__ g_xa, float* __restrict__ g_xb, float* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemBwd<H>& s = *reinterpret_cast<SmemBwd<H>*>(smem_raw);
    const int t = threadIdx.x, d = da + db;
    const int first = blockIdx.x * TM;
    copy_block<PackedFwd<H>>(s, packed->f);
    load_inputs(s, n, da, db, first, xa, xb);
    __syncthreads();
    float a[NA];
    forward_tile<H>(s, a);
    __syncthreads();                                       // every thread is done with the forward-layout weights
    PackedBwd<H>& nat = *reinterpret_cast<PackedBwd<H>*>(static_cast<PackedFwd<H>*>(&s));
    copy_block<PackedBwd<H>>(nat, packed->b);
    // ---- output layer: dz3 = g_a * clip'(a) * (1 - a^2) --------------------------------------------------------
    if (t < TM) {
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
    // ---- dZ2 = (dz3 W3) * (1 - h2^2), in place over h2: two threads per agent, every other neuron ---------------
    {
        const int la = t % TM;
        const float z0 = s.dz3[0][la], z1 = s.dz3[1][la], z2 = s.dz3[2][la], z3 = s.dz3[3][la];
#pragma unroll 8
        for (int k = t / TM; k < H; k += NT / TM) {
            const float h = s.h2[k][la];
            const float dh = z0 * nat.w3n[0][k] + z1 * nat.w3n[1][k] + z2 * nat.w3n[2][k] + z3 * nat.w3n[3][k];
            s.h2[k][la] = dh * (1.f - h * h);
        }
    }
    __syncthreads();
    // ---- dW2[j][i] = sum_a dz2[j][a] h1[i][a]: H/16 (j) x H/16 (i) outputs per thread,  db2 -----------------------
    {
        constexpr int JB = H / 16;
        const int tj = t / 16, ti = t % 16;
        float acc[JB][JB];
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int i = 0; i < JB; ++i) acc[j][i] = 0.f;
        for (int a4 = 0; a4 < TM; a4 += 4) {              // rows interleaved over the threads: conflict-free 128-bit loads
            float4 zv[JB], hv[JB];
#pragma unroll
            for (int j = 0; j < JB; ++j) zv[j] = lds4(&s.h2[tj + 16 * j][a4]);
#pragma unroll
            for (int i = 0; i < JB; ++i) hv[i] = lds4(&s.h1[ti + 16 * i][a4]);
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
                for (int i = 0; i < JB; ++i)
                    acc[j][i] += zv[j].x * hv[i].x + zv[j].y * hv[i].y + zv[j].z * hv[i].z + zv[j].w * hv[i].w;
        }
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int i = 0; i < JB; ++i) p_w2[(tj + 16 * j) * H + ti + 16 * i] = acc[j][i];
        if (t < H) {
            float b = 0.f;
            for (int a4 = 0; a4 < TM; a4 += 4) {
                const float4 z = lds4(&s.h2[t][a4]);
                b += z.x + z.y + z.z + z.w;
            }
            p_b2[t] = b;
        }
    }
    // ---- dZ1 = (dZ2 W2) * (1 - h1^2), in place over h1 (each thread owns its 8 x H/16 block) ---------------------
    {
        constexpr int NJ = H / 16;
        const int ty = t / 16, tx = t % 16;
        float acc[8][NJ];
        tile_gemm<H, H>(s.h2, nat.w2n, acc);
        __syncthreads();                                   // every thread is done reading h1 (dW2) before it changes
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float* row = &s.h1[tx + 16 * j][8 * ty];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = row[i];
                row[i] = acc[i][j] * (1.f - h * h);
            }
        }
    }
    __syncthreads();
    // ---- dW1[j][k] = sum_a dz1[j][a] x[k][a]: H/16 (j) x 2 (k) outputs per thread,  db1 ---------------------------
    {
        constexpr int JB = H / 16;
        const int tj = t / 16, ti = t % 16;
        float acc[JB][2];
#pragma unroll
        for (int j = 0; j < JB; ++j) acc[j][0] = acc[j][1] = 0.f;
        for (int a4 = 0; a4 < TM; a4 += 4) {
            const float4 x0 = lds4(&s.x[ti][a4]), x1 = lds4(&s.x[ti + 16][a4]);
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const float4 z = lds4(&s.h1[tj + 16 * j][a4]);
                acc[j][0] += z.x * x0.x + z.y * x0.y + z.z * x0.z + z.w * x0.w;
                acc[j][1] += z.x * x1.x + z.y * x1.y + z.z * x1.z + z.w * x1.w;
            }
        }
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (ti + 16 * c < d) p_w1[(tj + 16 * j) * d + ti + 16 * c] = acc[j][c];
        if (t < H) {
            float b = 0.f;
            for (int a4 = 0; a4 < TM; a4 += 4) {
                const float4 z = lds4(&s.h1[t][a4]);
                b += z.x + z.y + z.z + z.w;
            }
            p_b1[t] = b;
        }
    }
    // ---- dx[a][k] = sum_j dz1[j][a] W1[j][k]: tile product into the x tile, then one contiguous row per agent -----
    if (g_xa || g_xb) {
        float acc[8][DP / 16];
        tile_gemm<DP, H>(s.h1, nat.w1n, acc);
        __syncthreads();                                   // dW1 has read the x tile
        store_tile<DP, false>(s.x, acc, nullptr);
        __syncthreads();
        const int la = t % TM, agent = first + la;
        if (agent < n) {
            for (int k = t / TM; k < d; k += NT / TM) {
                const float v = s.x[k][la];
                if (k < da) { if (g_xa) g_xa[size_t(agent) * da + k] = v; }
                else if (g_xb) g_xb[size_t(agent) * db + (k - da)] = v;
            }
        }
    }
}

// out[e] = sum over CTAs of partial[c][e], fixed order: 32 elements x 16 slices of CTAs per block, then the slices
constexpr int RED_E = 32, RED_S = 16;
__global__ void __launch_bounds__(RED_E * RED_S)
vf_policy_reduce_kernel(int ctas, int size, const float* __restrict__ partial, float* __restrict__ out) {
    __shared__ float s[RED_S][RED_E];
    const int le = threadIdx.x % RED_E, sl = threadIdx.x / RED_E;
    const int e = blockIdx.x * RED_E + le;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (e < size) {
        int c = sl;
        for (; c + 3 * RED_S < ctas; c += 4 * RED_S) {
            a0 += __ldg(partial + size_t(c) * size + e);
            a1 += __ldg(partial + size_t(c + RED_S) * size + e);
            a2 += __ldg(partial + size_t(c + 2 * RED_S) * size + e);
            a3 += __ldg(partial + size_t(c + 3 * RED_S) * size + e);
        }
        for (; c < ctas; c += RED_S) a0 += __ldg(partial + size_t(c) * size + e);
    }
    s[sl][le] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (sl == 0 && e < size) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < RED_S; ++k) acc += s[k][le];
        out[e] = acc;
    }
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

// opt-in to > 48 KB of dynamic shared memory: once per kernel and device (the attribute is per device), not per launch
template <class K> int allow_smem(K kernel, size_t bytes) {
    static std::mutex lock;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kernel));
    {
        std::lock_guard<std::mutex> g(lock);
        if (done.count(key)) return 0;
    }
    const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (err != cudaSuccess) return policy_fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed", err);
    std::lock_guard<std::mutex> g(lock);
    done.insert(key);
    return 0;
}

// VF_POLICY_NO_TC=1 keeps the CUDA-core kernels (A/B measurements, profiles/)
bool use_tensor_cores() {
    static const bool on = [] {
        const char* e = getenv("VF_POLICY_NO_TC");
        return !(e && e[0] && e[0] != '0');
    }();
    return on;
}

// persistent CTAs: `per_sm` per SM of the current device, never more than there are tiles
int tc_grid(int tiles, int per_sm) {
    static const int sms = [] {
        int dev = 0, count = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        return count;
    }();
    const int cap = sms * per_sm;
    return tiles < cap ? tiles : cap;
}

int check_shapes(int n, int da, int db, int h, const void* xb) {
    if (n < 0) return policy_fail("n must be >= 0");
    if (da < 1 || db < 0 || da + db > DP) return policy_fail("policy input width must be in 1..32");
    if (db > 0 && !xb && n > 0) return policy_fail("second input piece is NULL");
    if (h != 32 && h != 64) return policy_fail("policy hidden width must be 32 or 64");
    return 0;
}

}  // namespace

extern "C" {

const char* vf_policy_last_error(void) { return g_policy_error.c_str(); }

int vf_policy_partial_floats(int n, int d, int h) { return ((n + TM - 1) / TM) * partial_size(h, d); }

int vf_policy_packed_floats(int h) { return int((h == 64 ? sizeof(Packed<64>) : sizeof(Packed<32>)) / sizeof(float)); }

int vf_policy_pack(int d, int h, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                   const float* b3, float* packed, void* stream) {
    if (check_shapes(0, d, 0, h, nullptr)) return 1;
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !packed) return policy_fail("vf_policy_pack: NULL buffer");
    if (reinterpret_cast<size_t>(packed) & 15u) return policy_fail("packed must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = h * (h > DP ? h : DP);
    if (h == 64)
        vf_policy_pack_kernel<64><<<(threads + 255) / 256, 256, 0, st>>>(d, w1, b1, w2, b2, w3, b3,
                                                                         reinterpret_cast<Packed<64>*>(packed));
    else
        vf_policy_pack_kernel<32><<<(threads + 255) / 256, 256, 0, st>>>(d, w1, b1, w2, b2, w3, b3,
                                                                         reinterpret_cast<Packed<32>*>(packed));
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : policy_fail("vf_policy_pack launch failed", err);
}

int vf_policy_fwd(int n, int da, int db, int h, const float* xa, const float* xb, const float* packed, float lo,
                  float hi, float* action, void* stream) {
    if (check_shapes(n, da, db, h, xb)) return 1;
    if (n == 0) return 0;
    if (!xa || !packed || !action) return policy_fail("vf_policy_fwd: NULL buffer");
    if ((reinterpret_cast<size_t>(action) | reinterpret_cast<size_t>(packed)) & 15u)
        return policy_fail("action and packed must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (n + TM - 1) / TM;
    if (use_tensor_cores()) {
        const int d = da + db;
#define VF_TC_FWD(H_, DK_)                                                                                            \
        do {                                                                                                          \
            if (allow_smem(tc::vf_policy_fwd_tc_kernel<H_, DK_>, sizeof(tc::FwdSmem<H_, DK_>))) return 1;             \
            tc::vf_policy_fwd_tc_kernel<H_, DK_><<<tc_grid(grid, 2), tc::FWD_THREADS, sizeof(tc::FwdSmem<H_, DK_>), st>>>( \
                n, da, db, xa, xb, &reinterpret_cast<const Packed<H_>*>(packed)->t, lo, hi, action);                  \
        } while (0)
        if (h == 64) { if (d <= 16) VF_TC_FWD(64, 16); else VF_TC_FWD(64, 32); }
        else { if (d <= 16) VF_TC_FWD(32, 16); else VF_TC_FWD(32, 32); }
#undef VF_TC_FWD
        const cudaError_t err = cudaGetLastError();
        return err == cudaSuccess ? 0 : policy_fail("vf_policy_fwd (tensor-core kernel) launch failed", err);
    }
    if (h == 64) {
        if (allow_smem(vf_policy_fwd_kernel<64>, sizeof(Smem<64>))) return 1;
        vf_policy_fwd_kernel<64><<<grid, NT, sizeof(Smem<64>), st>>>(n, da, db, xa, xb,
                                                                    reinterpret_cast<const Packed<64>*>(packed), lo, hi, action);
    } else {
        if (allow_smem(vf_policy_fwd_kernel<32>, sizeof(Smem<32>))) return 1;
        vf_policy_fwd_kernel<32><<<grid, NT, sizeof(Smem<32>), st>>>(n, da, db, xa, xb,
                                                                    reinterpret_cast<const Packed<32>*>(packed), lo, hi, action);
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : policy_fail("vf_policy_fwd launch failed", err);
}

int vf_policy_bwd(int n, int da, int db, int h, const float* xa, const float* xb, const float* packed, float lo,
                  float hi, const float* grad_action, float* grad_xa, float* grad_xb, float* partial,
                  float* grad_params, void* stream) {
    if (check_shapes(n, da, db, h, xb)) return 1;
    if (!xa || !packed || !grad_action || !partial || !grad_params) return policy_fail("vf_policy_bwd: NULL buffer");
    if ((reinterpret_cast<size_t>(grad_action) | reinterpret_cast<size_t>(packed)) & 15u)
        return policy_fail("grad_action and packed must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = (n + TM - 1) / TM;
    const int size = partial_size(h, da + db);
    if (n > 0 && use_tensor_cores() && da + db <= 16) {
        grid = tc_grid(grid, 1);                       // persistent CTAs: one partial per CTA
        if (h == 64) {
            if (allow_smem(tc::vf_policy_bwd_tc_kernel<64>, sizeof(tc::BwdSmem<64>))) return 1;
            tc::vf_policy_bwd_tc_kernel<64><<<grid, tc::BWD_THREADS, sizeof(tc::BwdSmem<64>), st>>>(
                n, da, db, xa, xb, &reinterpret_cast<const Packed<64>*>(packed)->t, lo, hi, grad_action, grad_xa, grad_xb,
                partial);
        } else {
            if (allow_smem(tc::vf_policy_bwd_tc_kernel<32>, sizeof(tc::BwdSmem<32>))) return 1;
            tc::vf_policy_bwd_tc_kernel<32><<<grid, tc::BWD_THREADS, sizeof(tc::BwdSmem<32>), st>>>(
                n, da, db, xa, xb, &reinterpret_cast<const Packed<32>*>(packed)->t, lo, hi, grad_action, grad_xa, grad_xb,
                partial);
        }
    } else if (n > 0) {
        if (h == 64) {
            if (allow_smem(vf_policy_bwd_kernel<64>, sizeof(SmemBwd<64>))) return 1;
            vf_policy_bwd_kernel<64><<<grid, NT, sizeof(SmemBwd<64>), st>>>(
                n, da, db, xa, xb, reinterpret_cast<const Packed<64>*>(packed), lo, hi, grad_action, grad_xa, grad_xb, partial);
        } else {
            if (allow_smem(vf_policy_bwd_kernel<32>, sizeof(SmemBwd<32>))) return 1;
            vf_policy_bwd_kernel<32><<<grid, NT, sizeof(SmemBwd<32>), st>>>(
                n, da, db, xa, xb, reinterpret_cast<const Packed<32>*>(packed), lo, hi, grad_action, grad_xa, grad_xb, partial);
        }
    }
    vf_policy_reduce_kernel<<<(size + RED_E - 1) / RED_E, RED_E * RED_S, 0, st>>>(grid, size, partial, grad_params);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : policy_fail("vf_policy_bwd launch failed", err);
}

}  // extern "C"
