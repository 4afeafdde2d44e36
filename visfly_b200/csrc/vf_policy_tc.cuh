// vf_policy_tc.cuh — the actor MLP on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in tensor memory).
// Included by vf_policy.cu (inside its anonymous namespace); fast_tanh, NA, partial_size come from there.
//
// One CTA of 128 threads takes a tile of 128 agents; thread t owns agent t = tensor-memory lane t.
//   * every product whose rows are agents (the three layers, dZ1 = dZ2 W2, dx = dZ1 W1) reads its A operand from
//     TENSOR MEMORY: the epilogue thread that has just produced its agent's activation row in registers stores it there
//     (tcgen05.st), so activations never travel through shared memory in this role; B = the weights, K-major in
//     shared memory in the un-swizzled 8 x 16-byte core-matrix layout, packed once per weight update;
//   * every product that contracts over agents (dW2, dW1 and the bias gradients) reads both operands from shared memory, K-major with
//     K = agents, in 128-byte-swizzled rows: a warp storing one feature of its 32 agents writes one 128-byte row
//     (conflict-free scalar stores);
//   * fp32 accuracy on kind::tf32: every operand is split into hi = tf32(x) and lo = tf32(x - hi) and every product
//     issued as hi*hi + lo*hi + hi*lo into the same fp32 accumulator (3xTF32; tools/ubench/umma_probe.cu measures
//     6e-7 of the result scale against 3e-4 for a single pass).
// tools/ubench/umma_probe.cu is the known-answer probe of each building block used here (descriptor encodings, A from
// tensor memory, the swizzled K = agents operands) — MN-major tf32 operands read as zeros in the un-swizzled layout
// (profiles/r02_umma_probe.txt), hence the second, transposed copy for the weight gradients.
#pragma once



namespace tc {

constexpr int TILE = 128;

__host__ __device__ constexpr int blk(int r, int c, int C) {          // float index in the blocked K-major layout
    return ((r / 8) * (C / 4) + c / 4) * 32 + (r % 8) * 4 + (c % 4);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t((lbo >> 4) & 0x3FFF) << 16) | (uint64_t((sbo >> 4) & 0x3FFF) << 32) |
           (uint64_t(1) << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {     // fp32 accumulate, tf32 x tf32, both K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                    "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                    "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
                    "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// tanh(x) = 1 - 2 / (1 + exp(2 x)) from one ex2 and one reciprocal (5 instructions with the bias folded into the scale):
// absolute error ~1e-7 over the whole range (the cancellation near 0 costs relative, not absolute accuracy), exact limits
// +-1 for large |x|, NaN passes through.  `y` = 2 log2(e) * x.
__device__ __forceinline__ float tanh_scaled(float y) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + exp2f(y)));
    return fmaf(-2.f, r, 1.f);
}
constexpr float TANH_SCALE = 2.885390081777927f;

__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = tf32_round(x);
    lo = tf32_round(x - hi);
}
// operands written by the threads (tensor-memory stores, generic-proxy shared-memory stores) -> visible to the MMAs the
// elected thread issues after the barrier
__device__ __forceinline__ void publish_operands() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
template <int COLS> __device__ __forceinline__ void tmem_free(uint32_t tmem, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(COLS));
}

// ---- the weights as the tensor cores want them (written by vf_policy_pack, copied verbatim into shared memory) --------
// [0] = hi, [1] = lo parts.  "blocked (R x C)" = blk(r, c, C): B operand with N = R, K = C.
template <int H> struct PackedTc {
    float w1_16[2][H * 16];      // blocked (H x 16):  W1[j][k], k < d <= 16 (zero beyond d)
    float w1_32[2][H * 32];      // blocked (H x 32):  the same for 16 < d <= 32
    float w2[2][H * H];          // blocked (H x H):   W2[j][i]
    float w1t[2][32 * H];        // blocked (32 x H):  W1[j][k] at (k, j)      (dx = dZ1 W1; the first 16 rows serve d <= 16)
    float w2t[2][H * H];         // blocked (H x H):   W2[j][i] at (i, j)      (dZ1 = dZ2 W2)
    float b1[H], b2[H], b3[4], w3n[4][H];
};

template <int H>
__device__ void pack_tc(int e, int d, const float* __restrict__ w1, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                        const float* __restrict__ b3, PackedTc<H>& p) {
    float hi, lo;
    if (e < H * 32) {
        const int j = e / 32, k = e % 32;
        split(k < d ? w1[j * d + k] : 0.f, hi, lo);
        if (k < 16) { p.w1_16[0][blk(j, k, 16)] = hi; p.w1_16[1][blk(j, k, 16)] = lo; }
        p.w1_32[0][blk(j, k, 32)] = hi; p.w1_32[1][blk(j, k, 32)] = lo;
        p.w1t[0][blk(k, j, H)] = hi; p.w1t[1][blk(k, j, H)] = lo;
    }
    if (e < H * H) {
        const int j = e / H, i = e % H;
        split(w2[j * H + i], hi, lo);
        p.w2[0][blk(j, i, H)] = hi; p.w2[1][blk(j, i, H)] = lo;
        p.w2t[0][blk(i, j, H)] = hi; p.w2t[1][blk(i, j, H)] = lo;
    }
    if (e < NA * H) p.w3n[e / H][e % H] = w3[e];          // the output layer runs on the CUDA cores
    if (e < H) { p.b1[e] = b1[e]; p.b2[e] = b2[e]; }
    if (e < NA) p.b3[e] = b3[e];
}

template <int THREADS> __device__ __forceinline__ void copy_floats(float* dst, const float* src, int count) {
    const float4* g = reinterpret_cast<const float4*>(src);
    float4* s4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < count / 4; i += THREADS) s4[i] = __ldg(g + i);
}

// issue D[128 x N] (+)= A[128 x K] B[N x K]^T as 3xTF32: A hi / lo in tensor memory at a_hi / a_lo (K columns each), B hi /
// lo blocked (N x K) in shared memory.  One thread.
template <int N, int K>
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, const float* b_hi, const float* b_lo,
                                         bool accumulate) {
    constexpr uint32_t idesc = idesc_tf32(TILE, N);
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t a = p == 1 ? a_lo : a_hi;
        const uint32_t b = smem_u32(p == 2 ? b_lo : b_hi);
#pragma unroll
        for (int k = 0; k < K / 8; ++k) {
            umma_ts(d, a + k * 8, desc_kmajor(b + k * 256, 128, (K / 4) * 128), idesc, acc);
            acc = 1u;
        }
    }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// 512 threads per CTA: thread t owns agent (row) r = t % 128 of the tile — tensor-memory lane r, reachable from warp
// (t / 32) because (t / 32) % 4 == r / 32 — and the 16 columns [16 q, 16 q + 16), q = t / 128, of every activation
// row: four threads share an agent, so the 128-lane tile keeps 16 warps busy with the tanh / split work between the
// products.  CTAs are persistent (grid = 2 per SM): weights and tensor memory are set up once, tiles are strided.
// tensor-memory columns: [0, H) accumulator (Z1, Z2, then Z3 in its first 16), [H, 3H) the A operand hi | lo
constexpr int FWD_THREADS = 512;
template <int H, int DK> struct FwdSmem {
    float w1[2][H * DK];
    float w2[2][H * H];
    float w3n[NA][H];
    float b1s[H], b2s[H], b3[4];              // b1, b2 pre-multiplied by TANH_SCALE
    float4 mu[4][TILE];                       // output-layer partial sums of the four column groups
    uint64_t bar;
    uint32_t tmem;
};

// epilogue of a hidden layer, this thread's 16 columns: h = tanh(acc + bias) (bias pre-multiplied by TANH_SCALE)
template <int H>
__device__ __forceinline__ void hidden_values(uint32_t lane_base, uint32_t acc_col, const float* bias_scaled, int q, float h[16]) {
    float v[16];
    tmem_ld16(lane_base + acc_col + 16 * q, v);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_scaled + 16 * q + i);
        h[i] = tanh_scaled(fmaf(v[i], TANH_SCALE, b.x));
        h[i + 1] = tanh_scaled(fmaf(v[i + 1], TANH_SCALE, b.y));
        h[i + 2] = tanh_scaled(fmaf(v[i + 2], TANH_SCALE, b.z));
        h[i + 3] = tanh_scaled(fmaf(v[i + 3], TANH_SCALE, b.w));
    }
}
// ... and its hi / lo parts to the A-operand columns of the next product
template <int H> __device__ __forceinline__ void store_operand(uint32_t lane_base, uint32_t op_col, int q, const float h[16]) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) split(h[i], hi[i], lo[i]);
    tmem_st16(lane_base + op_col + 16 * q, hi);
    tmem_st16(lane_base + op_col + H + 16 * q, lo);
}
// output layer on the CUDA cores (4 x H: too thin for an MMA phase of its own): this thread's 16 columns of the
// agent's h2 row against W3, the four column groups meet in shared memory
template <int H>
__device__ __forceinline__ float4 output_partial(const float (*w3n)[H], int q, const float h[16]) {
    float mu[NA] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < NA; ++o)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 w = *reinterpret_cast<const float4*>(&w3n[o][16 * q + i]);
            mu[o] = fmaf(h[i], w.x, fmaf(h[i + 1], w.y, fmaf(h[i + 2], w.z, fmaf(h[i + 3], w.w, mu[o]))));
        }
    return make_float4(mu[0], mu[1], mu[2], mu[3]);
}

// this thread's 16 observation columns [16 q, 16 q + 16) of row `agent` of [xa | xb | 0...]
__device__ __forceinline__ void load_cols(float x[16], int q, int n, int da, int db, int agent, const float* __restrict__ xa,
                                          const float* __restrict__ xb) {
    const int d = da + db;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = 16 * q + i;
        float v = 0.f;
        if (agent < n) {
            if (k < da) v = __ldg(xa + size_t(agent) * da + k);
            else if (k < d) v = __ldg(xb + size_t(agent) * db + (k - da));
        }
        x[i] = v;
    }
}

template <int H, int DK>
__global__ void __launch_bounds__(FWD_THREADS, 2)
vf_policy_fwd_tc_kernel(int n, int da, int db, const float* __restrict__ xa, const float* __restrict__ xb,
                        const PackedTc<H>* __restrict__ packed, float lo, float hi, float* __restrict__ action) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdSmem<H, DK>& s = *reinterpret_cast<FwdSmem<H, DK>*>(smem_raw);
    constexpr int COLS = 4 * H >= 256 ? 256 : 128;                 // 3H columns used
    const int t = threadIdx.x, warp = t >> 5, q = t >> 7, r = t & 127;
    const int tiles = (n + TILE - 1) / TILE;
    copy_floats<FWD_THREADS>(&s.w1[0][0], DK == 16 ? &packed->w1_16[0][0] : &packed->w1_32[0][0], 2 * H * DK);
    copy_floats<FWD_THREADS>(&s.w2[0][0], &packed->w2[0][0], 2 * H * H);
    copy_floats<FWD_THREADS>(&s.w3n[0][0], &packed->w3n[0][0], NA * H);
    if (t < H) { s.b1s[t] = packed->b1[t] * TANH_SCALE; s.b2s[t] = packed->b2[t] * TANH_SCALE; }
    if (t < NA) s.b3[t] = packed->b3[t];
    if (t == 0) mbar_init(&s.bar);
    const uint32_t tmem = tmem_alloc<COLS>(&s.tmem, warp);
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    constexpr uint32_t ACC = 0, OP = H;
    uint32_t phase = 0;
    float x[16];
    if (16 * q < DK) load_cols(x, q, n, da, db, blockIdx.x * TILE + r, xa, xb);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int agent = tile * TILE + r;
        if (16 * q < DK) {
            float xh[16], xl[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) split(x[k], xh[k], xl[k]);
            tmem_st16(lane_base + OP + 16 * q, xh);
            tmem_st16(lane_base + OP + DK + 16 * q, xl);
            if (tile + int(gridDim.x) < tiles)                    // the next tile's rows travel while this one computes
                load_cols(x, q, n, da, db, (tile + gridDim.x) * TILE + r, xa, xb);
        }
        publish_operands();
        if (t == 0) {
            issue_ts<H, DK>(tmem + ACC, tmem + OP, tmem + OP + DK, s.w1[0], s.w1[1], false);
            umma_commit(&s.bar);
        }
        mbar_wait(&s.bar, phase); phase ^= 1;
        if (16 * q < H) {
            float h[16];
            hidden_values<H>(lane_base, ACC, s.b1s, q, h);
            store_operand<H>(lane_base, OP, q, h);
        }
        publish_operands();
        if (t == 0) {
            issue_ts<H, H>(tmem + ACC, tmem + OP, tmem + OP + H, s.w2[0], s.w2[1], false);
            umma_commit(&s.bar);
        }
        mbar_wait(&s.bar, phase); phase ^= 1;
        if (16 * q < H) {
            float h[16];
            hidden_values<H>(lane_base, ACC, s.b2s, q, h);
            s.mu[q][r] = output_partial<H>(s.w3n, q, h);
        } else {
            s.mu[q][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (q == 0 && agent < n) {
            const float4 m0 = s.mu[0][r], m1 = s.mu[1][r], m2 = s.mu[2][r], m3 = s.mu[3][r];
            float a[NA] = {(m0.x + m1.x) + (m2.x + m3.x), (m0.y + m1.y) + (m2.y + m3.y), (m0.z + m1.z) + (m2.z + m3.z),
                           (m0.w + m1.w) + (m2.w + m3.w)};
#pragma unroll
            for (int o = 0; o < NA; ++o) {
                a[o] = tanh_scaled((a[o] + s.b3[o]) * TANH_SCALE);
                a[o] = a[o] < lo ? lo : (a[o] > hi ? hi : a[o]);       // th.clip of the trainers: NaN passes through
            }
            reinterpret_cast<float4*>(action)[agent] = make_float4(a[0], a[1], a[2], a[3]);
        }
        // the next tile's x overwrites operand columns layer 2 has finished reading (its commit was awaited); s.mu is
        // rewritten only after the next tile's two barriers
    }
    tmem_free<COLS>(tmem, warp);
}

// ---- backward --------------------------------------------------------------------------------------------------------
// One persistent CTA of 512 threads per SM, all of its tensor memory, d <= 16.  Per tile of 128 agents:
//   P0  x -> tensor memory (A of layer 1) and shared memory (K = agents, B of dW1)              MMA: Z1 = x W1^T
//   P1  h1 = tanh(Z1 + b1) -> tensor memory (A of layer 2), shared memory (B of dW2)            MMA: Z2 = h1 W2^T
//   P2  h2 = tanh(Z2 + b2); output layer, dz3 and dW3 += h2^T dz3 (warp butterfly) on the CUDA cores
//       dZ2 = (dz3 W3)(1 - h2^2) -> tensor memory (A of dZ1), shared memory (A of dW2)           MMA: dZ1' = dZ2 W2,
//                                                                                                    [dW2 | db2] += dZ2^T [h1 | 1]
//   P3  dZ1 = dZ1' (1 - h1^2) -> tensor memory (A of dx), shared memory (A of dW1, over dZ2)    MMA: dx = dZ1 W1,
//                                                                                                    [dW1 | db1] += dZ1^T [x | 1]
//   P4  dx -> global memory
// Two completion barriers: A for the products the next phase consumes (agents as rows), B for the weight-gradient
// products (contracted over agents, M = 64: the accumulator row j sits in tensor-memory lane 32 (j / 16) + j % 16),
// which only have to be finished before their shared-memory operands are overwritten — they run under the next phase's
// CUDA-core work.  Their accumulators stay in tensor memory across the CTA's tiles and are written out once, as this
// CTA's partial; vf_policy_reduce_kernel adds the partials in a fixed order.  The bias gradients ride along as one more
// operand row of ones.  W2 and W2^T take turns in one shared-memory buffer (cp.async from the packed block,
// L2-resident): both plus the K = agents operands do not fit in 227 KB.
constexpr int BWD_THREADS = 512;
template <int H> struct BwdSmem {
    // K = agents operands, [hi | lo][4 groups of 32 agents][F rows][32 floats], 16-byte chunks XOR row % 8
    float kx[2][4 * 24 * 32];          // x (16 rows) + a row group whose hi part is ones
    float kh1[2][4 * (H + 8) * 32];    // h1 (H rows) + a row group whose hi part is ones
    float kh2[2][4 * H * 32];          // dZ2, then dZ1
    float w1[2][H * 16];               // blocked (H x 16)
    float w1t[2][16 * H];              // blocked (16 x H)
    float w2[2][H * H];                // blocked W2 (layer 2) / W2^T (dZ1), swapped per tile
    float w3n[NA][H];
    float b1s[H], b2s[H], b3[4];
    float4 mu[3][TILE];                // output-layer partial sums of column groups 1..3
    float4 dz3[TILE];
    float db3[4][NA];
    float dw3[16][64];                 // end of the kernel: every warp's share of dW3
    uint64_t bar[2];
    uint32_t tmem;
};

// D[64 x N] (+)= A B^T contracted over the tile's 128 agents: A (FA rows per agent group, the first 64 used) and B (FB
// rows per group, the first N used) are K = agents operands
template <int N, int FA, int FB>
__device__ __forceinline__ void issue_kk(uint32_t d, const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                                         bool accumulate) {
    constexpr uint32_t idesc = idesc_tf32(64, N);
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t a = smem_u32(p == 1 ? a_lo : a_hi);
        const uint32_t b = smem_u32(p == 2 ? b_lo : b_hi);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            umma_ss(d, desc_sw128(a + (k / 4) * FA * 128 + (k % 4) * 32), desc_sw128(b + (k / 4) * FB * 128 + (k % 4) * 32),
                    idesc, acc);
            acc = 1u;
        }
    }
}

// this thread's 16 features [16 q, 16 q + 16) of agent r into a K = agents operand with F rows per agent group
template <int F>
__device__ __forceinline__ void store_krole(float* hi_buf, float* lo_buf, int q, int r, const float v[16]) {
    const int l = r & 31, base = (r >> 5) * F * 32 + 16 * q * 32;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = base + i * 32 + ((((l >> 2) ^ (i & 7)) << 2) | (l & 3));
        float hi, lo;
        split(v[i], hi, lo);
        hi_buf[idx] = hi;
        lo_buf[idx] = lo;
    }
}

// ... both at once (one split per value): the A operand of the next agents-as-rows product in tensor memory and the
// K = agents operand in shared memory
template <int H, int F>
__device__ __forceinline__ void store_both(uint32_t lane_base, uint32_t op_col, float* hi_buf, float* lo_buf, int q, int r,
                                           const float v[16]) {
    const int l = r & 31, base = (r >> 5) * F * 32 + 16 * q * 32;
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = base + i * 32 + ((((l >> 2) ^ (i & 7)) << 2) | (l & 3));
        split(v[i], hi[i], lo[i]);
        hi_buf[idx] = hi[i];
        lo_buf[idx] = lo[i];
    }
    tmem_st16(lane_base + op_col + 16 * q, hi);
    tmem_st16(lane_base + op_col + H + 16 * q, lo);
}

// sum over the 32 lanes of a warp of 64 per-lane values: lane L ends up with the totals of values 2 L and 2 L + 1 in
// val[0], val[1] (butterfly that halves the number of live values per step: 62 shuffles instead of 320)
__device__ __forceinline__ void warp_sum64(float val[64], int lane) {
#pragma unroll
    for (int s = 16, count = 64; s > 0; s >>= 1, count >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < count / 2; ++k) {
            const float send = up ? val[k] : val[k + count / 2];
            const float keep = up ? val[k + count / 2] : val[k];
            val[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
}

template <int THREADS> __device__ __forceinline__ void copy_async(float* dst, const float* src, int count) {
    const uint32_t d = smem_u32(dst);
    for (int i = threadIdx.x; i < count / 4; i += THREADS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + 16 * i), "l"(src + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void copy_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int H>
__global__ void __launch_bounds__(BWD_THREADS, 1)
vf_policy_bwd_tc_kernel(int n, int da, int db, const float* __restrict__ xa, const float* __restrict__ xb,
                        const PackedTc<H>* __restrict__ packed, float lo, float hi, const float* __restrict__ g_action,
                        float* __restrict__ g_xa, float* __restrict__ g_xb, float* __restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    BwdSmem<H>& s = *reinterpret_cast<BwdSmem<H>*>(smem_raw);
    constexpr int DK = 16;
    constexpr uint32_t ACC0 = 0, ACC1 = H, OPA = 2 * H, OPB = 4 * H, DW2 = 6 * H, DW1 = 7 * H + 8;
    static_assert(7 * H + 32 <= 512, "tensor-memory columns");
    const int t = threadIdx.x, warp = t >> 5, q = t >> 7, r = t & 127, d = da + db;
    const int tiles = (n + TILE - 1) / TILE;
    copy_floats<BWD_THREADS>(&s.w1[0][0], &packed->w1_16[0][0], 2 * H * 16);
    // the first 16 rows of the blocked (32 x H) W1^T are its first two row groups: contiguous, per hi / lo part
    copy_floats<BWD_THREADS>(&s.w1t[0][0], &packed->w1t[0][0], 16 * H);
    copy_floats<BWD_THREADS>(&s.w1t[1][0], &packed->w1t[1][0], 16 * H);
    copy_floats<BWD_THREADS>(&s.w2[0][0], &packed->w2[0][0], 2 * H * H);
    copy_floats<BWD_THREADS>(&s.w3n[0][0], &packed->w3n[0][0], NA * H);
    for (int i = t; i < 4 * 8 * 32; i += BWD_THREADS) {       // the row groups of ones (hi) / zeros (lo), per agent group
        const int grp = i / (8 * 32), e = i % (8 * 32);
        s.kx[0][grp * 24 * 32 + 16 * 32 + e] = 1.f;
        s.kx[1][grp * 24 * 32 + 16 * 32 + e] = 0.f;
        s.kh1[0][grp * (H + 8) * 32 + H * 32 + e] = 1.f;
        s.kh1[1][grp * (H + 8) * 32 + H * 32 + e] = 0.f;
    }
    if (t < H) { s.b1s[t] = packed->b1[t] * TANH_SCALE; s.b2s[t] = packed->b2[t] * TANH_SCALE; }
    if (t < NA) s.b3[t] = packed->b3[t];
    if (t == 0) { mbar_init(&s.bar[0]); mbar_init(&s.bar[1]); }
    const uint32_t tmem = tmem_alloc<512>(&s.tmem, warp);
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    uint32_t phase_a = 0, phase_b = 0;
    float db3_acc[NA] = {0.f, 0.f, 0.f, 0.f};
    float dw3_acc[2] = {0.f, 0.f};           // this lane's two entries of the warp's dW3 share (warp_sum64)
    float x[16];
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q == 0) {
        const int agent = blockIdx.x * TILE + r;
        load_cols(x, 0, n, da, db, agent, xa, xb);
        if (agent < n) g = __ldg(reinterpret_cast<const float4*>(g_action) + agent);
    }
    bool first = true;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, first = false) {
        const int agent = tile * TILE + r;
        const float4 g_now = g;
        // ---- P0 ------------------------------------------------------------------------------------------------
        if (!first) { mbar_wait(&s.bar[1], phase_b); phase_b ^= 1; }       // dW1 of the previous tile has read kx / kh2
        if (q == 0) {
            float xh[16], xl[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) split(x[k], xh[k], xl[k]);
            tmem_st16(lane_base + OPA, xh);
            tmem_st16(lane_base + OPA + DK, xl);
            store_krole<24>(s.kx[0], s.kx[1], 0, r, x);
            const int next = (tile + int(gridDim.x)) * TILE + r;
            if (tile + int(gridDim.x) < tiles) {
                load_cols(x, 0, n, da, db, next, xa, xb);
                g = next < n ? __ldg(reinterpret_cast<const float4*>(g_action) + next) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        copy_async_wait();                                     // W2 is back in its buffer (second and later tiles)
        publish_operands();
        if (t == 0) {
            issue_ts<H, DK>(tmem + ACC0, tmem + OPA, tmem + OPA + DK, s.w1[0], s.w1[1], false);
            umma_commit(&s.bar[0]);
        }
        mbar_wait(&s.bar[0], phase_a); phase_a ^= 1;
        // ---- P1 ------------------------------------------------------------------------------------------------
        if (16 * q < H) {
            float h[16];
            hidden_values<H>(lane_base, ACC0, s.b1s, q, h);
            store_both<H, H + 8>(lane_base, OPA, s.kh1[0], s.kh1[1], q, r, h);
        }
        publish_operands();
        if (t == 0) {
            issue_ts<H, H>(tmem + ACC1, tmem + OPA, tmem + OPA + H, s.w2[0], s.w2[1], false);
            umma_commit(&s.bar[0]);
        }
        mbar_wait(&s.bar[0], phase_a); phase_a ^= 1;
        copy_async<BWD_THREADS>(&s.w2[0][0], &packed->w2t[0][0], 2 * H * H);      // layer 2 is done with W2: W2^T moves in
        // ---- P2 ------------------------------------------------------------------------------------------------
        float h2[16];
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (16 * q < H) {
            hidden_values<H>(lane_base, ACC1, s.b2s, q, h2);
            mine = output_partial<H>(s.w3n, q, h2);
        }
        if (q > 0) s.mu[q - 1][r] = mine;
        __syncthreads();
        if (q == 0) {
            const float4 m1 = s.mu[0][r], m2 = s.mu[1][r], m3 = s.mu[2][r];
            const float mu[NA] = {(mine.x + m1.x) + (m2.x + m3.x), (mine.y + m1.y) + (m2.y + m3.y),
                                  (mine.z + m1.z) + (m2.z + m3.z), (mine.w + m1.w) + (m2.w + m3.w)};
            const float gv[NA] = {g_now.x, g_now.y, g_now.z, g_now.w};
            float z[NA];
#pragma unroll
            for (int o = 0; o < NA; ++o) {
                const float a = tanh_scaled((mu[o] + s.b3[o]) * TANH_SCALE);
                const float gate = (a >= lo && a <= hi) ? 1.f : 0.f;      // torch.clamp: closed interval
                z[o] = agent < n ? gv[o] * gate * (1.f - a * a) : 0.f;
                float v = z[o];                                            // db3: fixed-order sum over the warp's agents
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                db3_acc[o] += v;
            }
            s.dz3[r] = make_float4(z[0], z[1], z[2], z[3]);
        }
        __syncthreads();
        if (16 * q < H) {
            const float4 z = s.dz3[r];
            {   // dW3[o][16 q + i] += sum over the warp's agents of dz3[o] h2[i] (4 x 16 values per lane, CUDA cores: an
                // MMA over h2 would need a third K = agents buffer)
                float val[64];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    val[i] = z.x * h2[i]; val[16 + i] = z.y * h2[i]; val[32 + i] = z.z * h2[i]; val[48 + i] = z.w * h2[i];
                }
                warp_sum64(val, t & 31);
                dw3_acc[0] += val[0];
                dw3_acc[1] += val[1];
            }
            float dz2[16];                                     // dZ2 = (dz3 W3) * (1 - h2^2), this thread's 16 columns
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 wa = *reinterpret_cast<const float4*>(&s.w3n[0][16 * q + i]);
                const float4 wb = *reinterpret_cast<const float4*>(&s.w3n[1][16 * q + i]);
                const float4 wc = *reinterpret_cast<const float4*>(&s.w3n[2][16 * q + i]);
                const float4 wd = *reinterpret_cast<const float4*>(&s.w3n[3][16 * q + i]);
                dz2[i] = (z.x * wa.x + z.y * wb.x + z.z * wc.x + z.w * wd.x) * (1.f - h2[i] * h2[i]);
                dz2[i + 1] = (z.x * wa.y + z.y * wb.y + z.z * wc.y + z.w * wd.y) * (1.f - h2[i + 1] * h2[i + 1]);
                dz2[i + 2] = (z.x * wa.z + z.y * wb.z + z.z * wc.z + z.w * wd.z) * (1.f - h2[i + 2] * h2[i + 2]);
                dz2[i + 3] = (z.x * wa.w + z.y * wb.w + z.z * wc.w + z.w * wd.w) * (1.f - h2[i + 3] * h2[i + 3]);
            }
            store_both<H, H>(lane_base, OPB, s.kh2[0], s.kh2[1], q, r, dz2);     // kh2: free since barrier B at P0
        }
        copy_async_wait();                                     // W2^T has landed
        publish_operands();
        if (t == 0) {
            issue_ts<H, H>(tmem + ACC0, tmem + OPB, tmem + OPB + H, s.w2[0], s.w2[1], false);
            umma_commit(&s.bar[0]);
            issue_kk<H + 8, H, H + 8>(tmem + DW2, s.kh2[0], s.kh2[1], s.kh1[0], s.kh1[1], !first);
            umma_commit(&s.bar[1]);
        }
        mbar_wait(&s.bar[0], phase_a); phase_a ^= 1;
        copy_async<BWD_THREADS>(&s.w2[0][0], &packed->w2[0][0], 2 * H * H);       // W2 back for the next tile's layer 2
        // ---- P3 ------------------------------------------------------------------------------------------------
        float dz1[16];
        if (16 * q < H) {
            float v[16], hh[16], hl[16];
            tmem_ld16(lane_base + ACC0 + 16 * q, v);
            tmem_ld16(lane_base + OPA + 16 * q, hh);
            tmem_ld16(lane_base + OPA + H + 16 * q, hl);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h1 = hh[i] + hl[i];
                dz1[i] = v[i] * (1.f - h1 * h1);
            }
        }
        mbar_wait(&s.bar[1], phase_b); phase_b ^= 1;           // dW2 has read dZ2 from kh2: dZ1 may move in
        if (16 * q < H) store_both<H, H>(lane_base, OPA, s.kh2[0], s.kh2[1], q, r, dz1);
        publish_operands();
        if (t == 0) {
            issue_ts<16, H>(tmem + ACC1, tmem + OPA, tmem + OPA + H, s.w1t[0], s.w1t[1], false);
            umma_commit(&s.bar[0]);
            issue_kk<24, H, 24>(tmem + DW1, s.kh2[0], s.kh2[1], s.kx[0], s.kx[1], !first);
            umma_commit(&s.bar[1]);
        }
        mbar_wait(&s.bar[0], phase_a); phase_a ^= 1;
        // ---- P4 ------------------------------------------------------------------------------------------------
        if (q == 0 && (g_xa || g_xb)) {
            float v[16];
            tmem_ld16(lane_base + ACC1, v);
            if (agent < n) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k < da) { if (g_xa) g_xa[size_t(agent) * da + k] = v[k]; }
                    else if (k < d) { if (g_xb) g_xb[size_t(agent) * db + (k - da)] = v[k]; }
                }
            }
        }
        // next tile: its x overwrites operand columns whose products were awaited above, kx / kh2 after barrier B;
        // ACC1 is rewritten by its layer 2, two barriers from here
    }
    mbar_wait(&s.bar[1], phase_b);                             // the last dW1
    copy_async_wait();
    // ---- this CTA's partial weight gradients: accumulator row j = tensor-memory lane 32 (j / 16) + j % 16 ---------------
    float* out = partial + size_t(blockIdx.x) * partial_size(H, d);
    float* p_w1 = out;
    float* p_b1 = p_w1 + H * d;
    float* p_w2 = p_b1 + H;
    float* p_b2 = p_w2 + H * H;
    float* p_w3 = p_b2 + H;
    float* p_b3 = p_w3 + NA * H;
    {
        const int j = 16 * (r >> 5) + (r & 31);
        const bool row = (r & 31) < 16 && j < H;
        float v[16];
        if (16 * q < H) {
            tmem_ld16(lane_base + DW2 + 16 * q, v);
            if (row) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(p_w2 + j * H + 16 * q + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        if (q == 0) {
            tmem_ld16(lane_base + DW2 + H, v);
            if (row) p_b2[j] = v[0];
        } else if (q == 2) {
            tmem_ld16(lane_base + DW1, v);
            if (row) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < d) p_w1[j * d + k] = v[k];
            }
        } else {
            tmem_ld16(lane_base + DW1 + 16, v);
            if (row) p_b1[j] = v[0];
        }
    }
    if (q == 0 && (r & 31) == 0) {
#pragma unroll
        for (int o = 0; o < NA; ++o) s.db3[r >> 5][o] = db3_acc[o];
    }
    s.dw3[warp][2 * (t & 31)] = dw3_acc[0];                   // value index o * 16 + i of column group q = warp / 4
    s.dw3[warp][2 * (t & 31) + 1] = dw3_acc[1];
    __syncthreads();
    if (t < NA) p_b3[t] = (s.db3[0][t] + s.db3[1][t]) + (s.db3[2][t] + s.db3[3][t]);
    if (t < NA * H) {                                          // the four warps (agent groups) of each column group, fixed order
        const int o = t / H, k = t % H, qq = k / 16, e = o * 16 + k % 16;
        p_w3[t] = (s.dw3[4 * qq][e] + s.dw3[4 * qq + 1][e]) + (s.dw3[4 * qq + 2][e] + s.dw3[4 * qq + 3][e]);
    }
    tmem_free<512>(tmem, warp);
}

}  // namespace tc
