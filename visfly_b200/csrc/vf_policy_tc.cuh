// vf_policy_tc.cuh — the actor MLP on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in tensor memory).
// Included by vf_policy.cu (inside its anonymous namespace); fast_tanh, NA, partial_size come from there.
//
// One CTA of 128 threads takes a tile of 128 agents; thread t owns agent t = tensor-memory lane t.
//   * every product whose rows are agents (the three layers, dZ1 = dZ2 W2, dx = dZ1 W1) reads its A operand from
//     TENSOR MEMORY: the epilogue thread that has just produced its agent's activation row in registers stores it there
//     (tcgen05.st), so activations never travel through shared memory in this role; B = the weights, K-major in
//     shared memory in the un-swizzled 8 x 16-byte core-matrix layout, packed once per weight update;
//   * every product that contracts over agents (dW3, dW2, dW1) reads both operands from shared memory, K-major with
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
// float index of (feature f, agent a) in an (F x 128) K = agents operand: [a / 32][f][32 floats], 16-byte chunks ^ f % 8
__device__ __forceinline__ int sw128(int f, int a, int F) {
    const int l = a & 31;
    return (a >> 5) * F * 32 + f * 32 + ((((l >> 2) ^ (f & 7)) << 2) | (l & 3));
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
    float w3[2][16 * H];         // blocked (16 x H):  W3[o][k], rows >= 4 zero
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
    if (e < 16 * H) {
        const int o = e / H, k = e % H;
        split(o < NA ? w3[o * H + k] : 0.f, hi, lo);
        p.w3[0][blk(o, k, H)] = hi; p.w3[1][blk(o, k, H)] = lo;
        if (o < NA) p.w3n[o][k] = w3[o * H + k];
    }
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

}  // namespace tc
