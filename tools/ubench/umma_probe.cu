// umma_probe.cu — known-answer probe of the tcgen05 (UMMA) plumbing the actor kernels use: TMEM allocation, shared-
// memory matrix descriptors for the un-swizzled core-matrix layout in both majors, kind::tf32 with fp32 accumulation,
// the 3-pass hi/lo split (3xTF32) that brings the products back to fp32 accuracy, mbarrier completion, tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/ubench/umma_probe.cu && ./umma_probe
// Test 1 (forward-shaped):  D[128 x 64] = A[128 x 32] * B[64 x 32]^T      A, B K-major
// Test 2 (weight-gradient): D[f][g] = sum_a P[a][f] R[a][g]               P (128 x 64), R (128 x 32) read MN-major
//                           M = 128 issued, rows >= 64 are don't-care
// The blocked layout of an (R x C) fp32 matrix: [r / 8][c / 4][r % 8][c % 4] — 128-byte core matrices of 8 rows x 16 B.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__host__ __device__ inline int blk(int r, int c, int C) { return ((r / 8) * (C / 4) + c / 4) * 32 + (r % 8) * 4 + (c % 4); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// un-swizzled descriptor: start, leading byte offset, stride byte offset (all multiples of 16 B), sm_100 version bit
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t((lbo >> 4) & 0x3FFF) << 16) | (uint64_t((sbo >> 4) & 0x3FFF) << 32) |
           (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
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
// A from TMEM (lanes = rows, 32-bit columns = K), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand in 128-byte-swizzled rows: SBO = 1024 B between 8-row groups, layout type 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
    return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// byte offset of element (row f, column a) of an (F x 128) K-major operand: [a / 32][f][128 B], 16-byte chunks XOR f % 8
__host__ __device__ inline int sw128_off(int f, int a, int F) {
    const int l = a % 32;
    return (a / 32) * F * 128 + f * 128 + (((l / 4) ^ (f % 8)) * 16) + (l % 4) * 4;
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct Smem {
    float a_hi[128 * 64], a_lo[128 * 64];      // test 1: 128 x 32 (first half); test 2: P 128 x 64
    float b_hi[128 * 32], b_lo[128 * 32];      // test 1: 64 x 32; test 2: R 128 x 32
    float slack[1024];                         // test 2 reads up to 2 KB past a_lo / b_lo for the don't-care rows
    uint64_t bar;
    uint32_t tmem;
};

// passes = 1: plain TF32; passes = 3: hi*hi + lo*hi + hi*lo
__global__ void __launch_bounds__(128) probe(int test, int variant, int passes, const float* A, int na, const float* B, int nb, float* D,
                                             const float* A2) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int t = threadIdx.x, warp = t / 32;
    for (int i = t; i < na; i += 128) {
        const float x = A[i], h = tf32_hi(x);
        s.a_hi[i] = h;
        s.a_lo[i] = x - h;
    }
    for (int i = t; i < nb; i += 128) {
        const float x = B[i], h = tf32_hi(x);
        s.b_hi[i] = h;
        s.b_lo[i] = x - h;
    }
    for (int i = t; i < 1024; i += 128) s.slack[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        mbar_init(&s.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem;
    if (test == 3) {          // A (128 x 32, row-major in global memory at A2) -> hi / lo in TMEM columns [64,96) / [96,128)
        float hi[32], lo[32];
        for (int k = 0; k < 32; ++k) {
            const float x = A2[t * 32 + k];
            hi[k] = tf32_hi(x);
            lo[k] = x - hi[k];
        }
        const uint32_t lane = uint32_t(warp * 32) << 16;
        tmem_st16(tmem + lane + 64, hi); tmem_st16(tmem + lane + 80, hi + 16);
        tmem_st16(tmem + lane + 96, lo); tmem_st16(tmem + lane + 112, lo + 16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (t == 0 && test >= 3) {
        uint32_t acc = 0;
        for (int p = 0; p < passes; ++p) {
            const float* a = p == 1 ? s.a_lo : s.a_hi;
            const float* b = p == 2 ? s.b_lo : s.b_hi;
            if (test == 3) {
                constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
                for (int k = 0; k < 32 / 8; ++k) {
                    umma_tf32_ts(tmem, tmem + (p == 1 ? 96 : 64) + k * 8, make_desc(smem_u32(b) + k * 256, 128, 8 * 128), idesc, acc);
                    acc = 1;
                }
            } else {          // test 4: both operands K-major (K = 128 agents) in 128-byte-swizzled rows; A has 64 rows, M = 128
                const uint32_t idesc = make_idesc(test == 5 ? 64 : 128, 32, 0, 0);        // test 5: M = 64
                for (int k = 0; k < 128 / 8; ++k) {
                    umma_tf32(tmem, make_desc_sw128(smem_u32(a) + (k / 4) * 64 * 128 + (k % 4) * 32),
                              make_desc_sw128(smem_u32(b) + (k / 4) * 32 * 128 + (k % 4) * 32), idesc, acc);
                    acc = 1;
                }
            }
        }
        umma_commit(&s.bar);
    }
    if (t == 0 && test < 3) {
        uint32_t acc = 0;
        for (int p = 0; p < passes; ++p) {
            const float* a = p == 1 ? s.a_lo : s.a_hi;
            const float* b = p == 2 ? s.b_lo : s.b_hi;
            if (test == 1) {
                constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
                for (int k = 0; k < 32 / 8; ++k) {          // K-major: LBO = 128 B between K chunks, SBO = (32/4)*128 B
                    umma_tf32(tmem, make_desc(smem_u32(a) + k * 256, 128, 8 * 128), make_desc(smem_u32(b) + k * 256, 128, 8 * 128),
                              idesc, acc);
                    acc = 1;
                }
            } else {
                const int a_mn = variant & 1, b_mn = (variant >> 1) & 1, swap = (variant >> 2) & 1;
                const uint32_t idesc = make_idesc(128, 32, a_mn, b_mn);
                for (int k = 0; k < 128 / 8; ++k) {         // MN-major: SBO = 128 B between MN units, one 8-deep K group per MMA
                    const uint64_t da = a_mn ? (swap ? make_desc(smem_u32(a) + k * 16 * 128, 128, 16 * 128)
                                                     : make_desc(smem_u32(a) + k * 16 * 128, 16 * 128, 128))
                                             : make_desc(smem_u32(a) + k * 256, 128, 32 * 128);     // (64 x 128) K-major copy
                    const uint64_t db = b_mn ? (swap ? make_desc(smem_u32(b) + k * 8 * 128, 128, 8 * 128)
                                                     : make_desc(smem_u32(b) + k * 8 * 128, 8 * 128, 128))
                                             : make_desc(smem_u32(b) + k * 256, 128, 32 * 128);     // (32 x 128) K-major copy
                    umma_tf32(tmem, da, db, idesc, acc);
                    acc = 1;
                }
            }
        }
        umma_commit(&s.bar);
    }
    mbar_wait(&s.bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int ncol = (test == 1 || test == 3) ? 64 : 32;
    for (int c0 = 0; c0 < ncol; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[t * ncol + c0 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main() {
    srand(7);
    auto rnd = [] { return float(rand()) / RAND_MAX * 2.f - 1.f; };
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(Smem)));
    for (int test = 1; test <= 5; ++test) {
        const bool fwd = test == 1 || test == 3;
        const int ar = 128, ac = fwd ? 32 : 64, br = fwd ? 64 : 128, bc = 32;
        std::vector<float> A(ar * ac), B(br * bc), Ab(ar * ac), Bb(br * bc);
        for (auto& x : A) x = rnd();
        for (auto& x : B) x = rnd();
    for (int variant = 0; variant < (test == 2 ? 2 : 1); ++variant) {
        if (test == 2 && (variant & 4) && !(variant & 3)) continue;
        for (int r = 0; r < ar; ++r) for (int c = 0; c < ac; ++c) {
            if (test >= 4) Ab[sw128_off(c, r, ac) / 4] = A[r * ac + c];
            else if (test == 2 && !(variant & 1)) Ab[blk(c, r, ar)] = A[r * ac + c];       // transposed copy, K-major operand
            else Ab[blk(r, c, ac)] = A[r * ac + c];
        }
        for (int r = 0; r < br; ++r) for (int c = 0; c < bc; ++c) {
            if (test >= 4) Bb[sw128_off(c, r, bc) / 4] = B[r * bc + c];
            else if (test == 2 && !(variant & 2)) Bb[blk(c, r, br)] = B[r * bc + c];
            else Bb[blk(r, c, bc)] = B[r * bc + c];
        }
        const int dm = fwd ? 128 : 64, dn = fwd ? 64 : 32;
        std::vector<double> ref(dm * dn, 0.0);
        for (int i = 0; i < dm; ++i)
            for (int j = 0; j < dn; ++j) {
                double acc = 0;
                if (fwd) for (int k = 0; k < 32; ++k) acc += double(A[i * 32 + k]) * B[j * 32 + k];
                else for (int a = 0; a < 128; ++a) acc += double(A[a * 64 + i]) * B[a * 32 + j];
                ref[i * dn + j] = acc;
            }
        float *dA, *dB, *dD, *dA2;
        cudaMalloc(&dA2, A.size() * 4); cudaMemcpy(dA2, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 64 * 4);
        cudaMemcpy(dA, Ab.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bb.data(), B.size() * 4, cudaMemcpyHostToDevice);
        for (int passes : {1, 3}) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            probe<<<1, 128, sizeof(Smem)>>>(test, variant, passes, dA, int(A.size()), dB, int(B.size()), dD, dA2);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("test %d passes %d: CUDA error %s\n", test, passes, cudaGetErrorString(e)); return 1; }
            std::vector<float> D(128 * dn);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0, scale = 0;
            for (int i = 0; i < dm; ++i)
                for (int j = 0; j < dn; ++j) {
                    worst = fmax(worst, fabs(D[i * dn + j] - ref[i * dn + j]));
                    scale = fmax(scale, fabs(ref[i * dn + j]));
                }
            if (test == 5) {          // which tensor-memory lane holds which row of the M = 64 accumulator
                printf("test 5 passes %d lane->row:", passes);
                for (int lane = 0; lane < 128; ++lane) {
                    int hit = -1;
                    for (int i = 0; i < 64; ++i) {
                        bool same = true;
                        for (int j = 0; j < dn; ++j) same = same && fabs(D[lane * dn + j] - ref[i * dn + j]) < 5e-3 * (1 + fabs(ref[i * dn + j]));
                        if (same) hit = i;
                    }
                    printf(" %d", hit);
                }
                printf("\n");
                continue;
            }
            printf("test %d variant %d passes %d: max |err| %.3e (max |ref| %.3f)  D[0][0..3] = %.5f %.5f %.5f %.5f  ref %.5f %.5f %.5f %.5f\n",
                   test, variant, passes, worst, scale, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    }
    return 0;
}
