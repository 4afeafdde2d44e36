// vf_kernels.cu — sm_100a kernels and the C-ABI of include/visfly_b200.h.
//
// Design (B200): the control step is ~1-4 kFLOP of dependent fp32 arithmetic per agent on 24 input floats,
// no reuse across agents and no contraction, so there is nothing for tensor cores / TMA tiles to do; the
// levers that matter are (i) one HBM round trip per control step — all sub-steps run in registers,
// (ii) 128-bit coalesced accesses on an SoA-of-float4 state (a warp touches 512 contiguous bytes per plane),
// (iii) enough independent warps per SM sub-partition to cover the fp32 dependency chains (one agent per
// thread, small CTAs so that 65 536 agents spread evenly over 148 SMs), and (iv) no forward intermediates in
// HBM for the backward pass: the adjoint kernel re-runs the sub-steps and keeps its tape in thread-local
// memory (L1-resident).  The AoS observation (n,13) the reference returns is transposed through shared
// memory so its global stores/loads are 128-bit and coalesced as well.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "vf_env.cuh"
#include "vf_math.cuh"

namespace {

thread_local std::string g_last_error;

int fail(const char* what, cudaError_t err = cudaSuccess) {
    g_last_error = what;
    if (err != cudaSuccess) {
        g_last_error += ": ";
        g_last_error += cudaGetErrorString(err);
    }
    return 1;
}

constexpr int kObs = VF_OBS_FLOATS;       // 13
constexpr int kWarpObs = 32 * kObs;       // floats one warp's observations occupy (416 = 104 float4)

// Programmatic dependent launch (sm_90+): the step kernels are launched with the programmatic-stream-serialization
// attribute, so the CTAs of step t+1 become resident while step t is still running (65 536 agents fill less than
// half of the chip's warp slots) and park at `pdl_wait` — launch latency, CTA rasterisation and the first
// instruction-cache misses of step t+1 overlap with the arithmetic of step t.  `griddepcontrol.wait` returns once the
// preceding grid has completed and its writes are visible, so every global access below it is ordered as before.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Ahead of griddepcontrol.wait: pull this thread's input lines towards L2.  A prefetch has no architectural effect and L2
// is the point of coherence, so it is safe even where the preceding grid is still writing those lines; where the
// inputs come from HBM (a loop over more env replicas than fit in L2) the round trip overlaps the previous launch's tail.
constexpr unsigned kEnvFlagPrefetch = 0x40000000u;      // internal bit of env_flags, set by vf_env_step_fwd unless VF_NO_PREFETCH=1
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ float4 ldg4(const float* base, size_t idx4) {
    return __ldg(reinterpret_cast<const float4*>(base) + idx4);
}
__device__ __forceinline__ void stg4(float* base, size_t idx4, float4 v) {
    reinterpret_cast<float4*>(base)[idx4] = v;
}

__device__ __forceinline__ void load_state(const float* __restrict__ st, int n, int i, vf::State<float>& s) {
    const float4 a = ldg4(st, size_t(i));
    const float4 b = ldg4(st, size_t(n) + i);
    const float4 c = ldg4(st, size_t(2) * n + i);
    const float4 d = ldg4(st, size_t(3) * n + i);
    const float4 e = ldg4(st, size_t(4) * n + i);
    s.p[0] = a.x; s.p[1] = a.y; s.p[2] = a.z; s.al[0] = a.w;
    s.q[0] = b.x; s.q[1] = b.y; s.q[2] = b.z; s.q[3] = b.w;
    s.v[0] = c.x; s.v[1] = c.y; s.v[2] = c.z; s.al[1] = c.w;
    s.w[0] = d.x; s.w[1] = d.y; s.w[2] = d.z; s.al[2] = d.w;
    s.mot[0] = e.x; s.mot[1] = e.y; s.mot[2] = e.z; s.mot[3] = e.w;
}

__device__ __forceinline__ void store_state(float* __restrict__ st, int n, int i, const vf::State<float>& s) {
    stg4(st, size_t(i), make_float4(s.p[0], s.p[1], s.p[2], s.al[0]));
    stg4(st, size_t(n) + i, make_float4(s.q[0], s.q[1], s.q[2], s.q[3]));
    stg4(st, size_t(2) * n + i, make_float4(s.v[0], s.v[1], s.v[2], s.al[1]));
    stg4(st, size_t(3) * n + i, make_float4(s.w[0], s.w[1], s.w[2], s.al[2]));
    stg4(st, size_t(4) * n + i, make_float4(s.mot[0], s.mot[1], s.mot[2], s.mot[3]));
}

// Warp-cooperative transpose of 32 agents x 13 floats between registers and the row-major (n,13) array.
// `smem` is this warp's 416-float slice.  lane*13+j is conflict-free (13 is odd).
// `obs2` (optional) is a second destination with the same layout (the page-locked host mirror).
__device__ __forceinline__ void warp_store_obs(float* __restrict__ obs, float* smem, int n, int warp_first,
                                               int lane, const float o[kObs], float* __restrict__ obs2 = nullptr) {
    if (warp_first + 32 <= n) {
#pragma unroll
        for (int j = 0; j < kObs; ++j) smem[lane * kObs + j] = o[j];
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(obs + size_t(warp_first) * kObs);
        float4* dst2 = reinterpret_cast<float4*>(obs2 + size_t(warp_first) * kObs);
        const float4* src = reinterpret_cast<const float4*>(smem);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = lane + 32 * k;
            if (idx < kWarpObs / 4) {
                const float4 v = src[idx];
                dst[idx] = v;
                if (obs2) dst2[idx] = v;
            }
        }
        __syncwarp();
    } else if (warp_first + lane < n) {
        float* dst = obs + size_t(warp_first + lane) * kObs;
#pragma unroll
        for (int j = 0; j < kObs; ++j) dst[j] = o[j];
        if (obs2) {
            float* dst2 = obs2 + size_t(warp_first + lane) * kObs;
#pragma unroll
            for (int j = 0; j < kObs; ++j) dst2[j] = o[j];
        }
    }
}

__device__ __forceinline__ void warp_load_obs(const float* __restrict__ obs, float* smem, int n, int warp_first,
                                              int lane, float o[kObs]) {
    if (warp_first + 32 <= n) {
        const float4* src = reinterpret_cast<const float4*>(obs + size_t(warp_first) * kObs);
        float4* dst = reinterpret_cast<float4*>(smem);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = lane + 32 * k;
            if (idx < kWarpObs / 4) dst[idx] = __ldg(src + idx);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kObs; ++j) o[j] = smem[lane * kObs + j];
        __syncwarp();
    } else if (warp_first + lane < n) {
        const float* src = obs + size_t(warp_first + lane) * kObs;
#pragma unroll
        for (int j = 0; j < kObs; ++j) o[j] = __ldg(src + j);
    } else {
#pragma unroll
        for (int j = 0; j < kObs; ++j) o[j] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward control step
// ---------------------------------------------------------------------------------------------
template <int INTEG, int ACT, bool LAG, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
vf_step_fwd_kernel(const __grid_constant__ VfParams params, int n, int substeps,
                   const float* __restrict__ state_in, const float* __restrict__ action,
                   float* __restrict__ state_out, float* __restrict__ obs_out, float* __restrict__ ext_out,
                   const float* __restrict__ wind, const float* __restrict__ fifo_push,
                   float* __restrict__ fifo_copy, const __grid_constant__ VfFifoRows ring) {
    __shared__ __align__(16) float s_obs[(BLOCK / 32) * kWarpObs];
    const vf::Params<float>& P = reinterpret_cast<const vf::Params<float>&>(params);
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    const bool live = i < n;
    pdl_trigger();
    pdl_wait();

    vf::State<float> s;
    vf::Wrench<float> k;
    float wd[3] = {P.wind[0], P.wind[1], P.wind[2]};
    if (live) {
        load_state(state_in, n, i, s);
        float4 a4;
        if (ring.depth > 0) {
            // device-resident FIFO ring (vf_step_fwd_ring): consume the oldest row, shift, append — this agent's rows only
            a4 = *(reinterpret_cast<const float4*>(ring.row[0]) + i);
            for (int r = 0; r + 1 < ring.depth; ++r)
                stg4(ring.row[r], size_t(i), *(reinterpret_cast<const float4*>(ring.row[r + 1]) + i));
            stg4(ring.row[ring.depth - 1], size_t(i), ldg4(fifo_push, size_t(i)));
        } else {
            a4 = ldg4(action, size_t(i));
            // comm-delay FIFO: the engine-owned copy of the action that arrived this step (dynamics.py:324 `clone()`)
            if (fifo_copy) stg4(fifo_copy, size_t(i), ldg4(fifo_push, size_t(i)));
        }
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        if (wind) {                                  // per-agent wind of this control step (wind functions)
            const float4 w4 = ldg4(wind, size_t(i));
            wd[0] = w4.x; wd[1] = w4.y; wd[2] = w4.z;
        }
        vf::step_fwd<float>(P, substeps, INTEG, ACT, LAG, a, s, k, wd);
        store_state(state_out, n, i, s);
        if (ext_out) {
            stg4(ext_out, size_t(2) * i, make_float4(k.acc[0], k.acc[1], k.acc[2], 0.f));
            stg4(ext_out, size_t(2) * i + 1, make_float4(k.thr[0], k.thr[1], k.thr[2], k.thr[3]));
        }
    }
    if (obs_out) {
        float o[kObs];
        if (live) {
            o[0] = s.p[0]; o[1] = s.p[1]; o[2] = s.p[2];
            o[3] = s.q[0]; o[4] = s.q[1]; o[5] = s.q[2]; o[6] = s.q[3];
            o[7] = s.v[0] + wd[0]; o[8] = s.v[1] + wd[1]; o[9] = s.v[2] + wd[2];
            o[10] = s.w[0]; o[11] = s.w[1]; o[12] = s.w[2];
        }
        warp_store_obs(obs_out, s_obs + warp * kWarpObs, n, warp_first, lane, o);
    }
}

// ---------------------------------------------------------------------------------------------
// reverse control step
// ---------------------------------------------------------------------------------------------
template <int INTEG, int ACT, bool LAG, int SMAX, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
vf_step_bwd_kernel(const __grid_constant__ VfParams params, int n, int substeps,
                   const float* __restrict__ state_in, const float* __restrict__ action,
                   const float* __restrict__ g_state_out, const float* __restrict__ g_obs,
                   float* __restrict__ g_state_in, float* __restrict__ g_action, const float* __restrict__ wind) {
    __shared__ __align__(16) float s_obs[(BLOCK / 32) * kWarpObs];
    const vf::Params<float>& P = reinterpret_cast<const vf::Params<float>&>(params);
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    const bool live = i < n;
    pdl_trigger();
    pdl_wait();

    vf::State<float> g;
    if (g_state_out && live) {
        load_state(g_state_out, n, i, g);
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) g.p[j] = g.v[j] = g.w[j] = g.al[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) g.q[j] = g.mot[j] = 0.f;
    }
    if (g_obs) {
        float o[kObs];
        warp_load_obs(g_obs, s_obs + warp * kWarpObs, n, warp_first, lane, o);
        g.p[0] += o[0]; g.p[1] += o[1]; g.p[2] += o[2];
        g.q[0] += o[3]; g.q[1] += o[4]; g.q[2] += o[5]; g.q[3] += o[6];
        g.v[0] += o[7]; g.v[1] += o[8]; g.v[2] += o[9];
        g.w[0] += o[10]; g.w[1] += o[11]; g.w[2] += o[12];
    }
    if (!live) return;

    vf::State<float> s0;
    load_state(state_in, n, i, s0);
    const float4 a4 = ldg4(action, size_t(i));
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    vf::Tape<float> tape[SMAX];
    float ga[4];
    float wd[3] = {P.wind[0], P.wind[1], P.wind[2]};
    if (wind) {                                      // only the position clamp's gradient gate depends on it
        const float4 w4 = ldg4(wind, size_t(i));
        wd[0] = w4.x; wd[1] = w4.y; wd[2] = w4.z;
    }
    vf::step_bwd<float>(P, substeps, INTEG, ACT, LAG, a, s0, g, ga, tape, wd);
    store_state(g_state_in, n, i, g);
    stg4(g_action, size_t(i), make_float4(ga[0], ga[1], ga[2], ga[3]));
}

// ---------------------------------------------------------------------------------------------
// fused env step: control step + collision + task reward + termination + episode record + auto-reset
// ---------------------------------------------------------------------------------------------
template <int INTEG, int ACT, bool LAG, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
vf_env_step_fwd_kernel(const __grid_constant__ VfParams params, const __grid_constant__ VfEnvSpec E, int n,
                       int substeps, unsigned env_flags, unsigned long long step_index,
                       const unsigned long long* __restrict__ step_base,
                       const float* __restrict__ state_in, const float* __restrict__ action,
                       const float* __restrict__ wind, const float* __restrict__ fifo_push,
                       const float* __restrict__ reset_table, const int* __restrict__ status_in,
                       float* __restrict__ state_out, int* __restrict__ status_out, float* __restrict__ fifo_copy,
                       float* __restrict__ obs_out, float* __restrict__ reward_out,
                       unsigned char* __restrict__ done_out, float* __restrict__ record_out,
                       float* __restrict__ term_obs_out, long long* __restrict__ gate_out,
                       const VfEnvMirror mirror, const __grid_constant__ VfPeerScatter peers) {
    __shared__ __align__(16) float s_obs[(BLOCK / 32) * kWarpObs];
    const vf::Params<float>& P = reinterpret_cast<const vf::Params<float>&>(params);
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    const bool live = i < n;
    pdl_trigger();
    if (live && (env_flags & kEnvFlagPrefetch)) {
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) prefetch_l2(reinterpret_cast<const float4*>(state_in) + size_t(pl) * n + i);
        prefetch_l2(reinterpret_cast<const int4*>(status_in) + i);
        prefetch_l2(reinterpret_cast<const float4*>(action) + i);
    }
    pdl_wait();

    vf::State<float> s;
    float wd[3] = {P.wind[0], P.wind[1], P.wind[2]};
    int g = 0;
    if (live) {
        // every input of the step is requested up front: the env status record is only needed after the sub-step
        // loop, but fetching it there would put a second (cold) HBM round trip on each warp's critical path
        load_state(state_in, n, i, s);
        const int4 st4 = __ldg(reinterpret_cast<const int4*>(status_in) + i);
        float4 a4 = ldg4(action, size_t(i));
        if (wind) {
            const float4 w4 = ldg4(wind, size_t(i));
            wd[0] = w4.x; wd[1] = w4.y; wd[2] = w4.z;
        }
        if (fifo_copy) stg4(fifo_copy, size_t(i), ldg4(fifo_push, size_t(i)));
        const int age = st4.x;
        const float ret_in = __int_as_float(st4.y);
        const unsigned eb = unsigned(st4.z) & 0xFFu;
        const int g_in = (st4.z >> 8) & 0xFF;
        int passed = st4.w;
        if (age < E.fifo_depth) a4 = make_float4(0.f, 0.f, 0.f, 0.f);   // FIFO rows of a reset agent read as zero
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        vf::Wrench<float> k;
        vf::step_fwd<float>(P, substeps, INTEG, ACT, LAG, a, s, k, wd);

        int sc = age + 1;
        vf::EnvEval<float> ev;
        vf::env_eval<float>(P, E, s, wd, sc, g_in, (eb & VF_EBIT_EPISODE_DONE) != 0, ev);
        float vel[3] = {ev.vel[0], ev.vel[1], ev.vel[2]};
        g = ev.gate;
        passed += ev.pass ? 1 : 0;
        bool once = (eb & VF_EBIT_ONCE_COLLIDED) || ev.is_col;
        const bool success = ev.success;
        const float reward = ev.reward;
        float ret = ret_in + reward;
        bool ep_done = ev.ep_done;
        const bool done = ev.done;

        unsigned rbits = (done ? VF_RBIT_DONE : 0u) | (ep_done ? VF_RBIT_EPISODE_DONE : 0u) |
                         (success ? VF_RBIT_SUCCESS : 0u) | (sc >= E.max_episode_steps ? VF_RBIT_TRUNCATED : 0u) |
                         (once ? VF_RBIT_COLLIDED : 0u);
        stg4(record_out, size_t(i), make_float4(ret, float(sc), float(rbits), float(passed)));
        reward_out[i] = reward;
        done_out[i] = done ? 1 : 0;
        if (mirror.reward) mirror.reward[i] = reward;
        if (mirror.done) mirror.done[i] = done ? 1 : 0;

        if (done && term_obs_out) {                 // rare: scalar stores are fine
            if (E.obs_kind == VF_OBS_STATE13) {
                float* t = term_obs_out + size_t(i) * 13;
                t[0] = s.p[0]; t[1] = s.p[1]; t[2] = s.p[2];
                t[3] = s.q[0]; t[4] = s.q[1]; t[5] = s.q[2]; t[6] = s.q[3];
                t[7] = vel[0]; t[8] = vel[1]; t[9] = vel[2];
                t[10] = s.w[0]; t[11] = s.w[1]; t[12] = s.w[2];
            } else {
                float* t = term_obs_out + size_t(i) * 16;
                // the reference builds the terminal observation before the gate index advances (RacingEnv.py:254)
                const int g0 = g_in, g1 = g_in + 1 >= E.n_gates ? g_in + 1 - E.n_gates : g_in + 1;
                for (int j = 0; j < 3; ++j) {
                    t[j] = (E.gates[g0][j] - s.p[j]) / 10.f;
                    t[3 + j] = (E.gates[g1][j] - s.p[j]) / 10.f;
                }
                t[6] = s.q[0]; t[7] = s.q[1]; t[8] = s.q[2]; t[9] = s.q[3];
                for (int j = 0; j < 3; ++j) { t[10 + j] = vel[j] / 10.f; t[13 + j] = s.w[j] / 10.f; }
            }
        }
        if (done && !(env_flags & VF_ENV_FLAG_NO_RESET)) {
            if (E.task == VF_TASK_RACING) {         // chosen from the terminal position (RacingEnv.py:150-163)
                g = vf::racing_first_gate<float>(s.p);
                passed = 0;
            }
            const unsigned long long step = step_index + (step_base ? *step_base : 0ull);
            // out-of-line call (see sample_reset): results come back through temporaries so that the state itself
            // never has its address taken and stays in registers on the per-step path
            float rp[3], rq[4], rv[3], rw[3];
            vf::sample_reset(E, E.agent_offset + unsigned(i), step, reset_table ? reset_table + size_t(i) * 13 : nullptr,
                             rp, rq, rv, rw);
            for (int j = 0; j < 3; ++j) { s.p[j] = rp[j]; s.v[j] = rv[j]; s.w[j] = rw[j]; }
            for (int j = 0; j < 4; ++j) s.q[j] = rq[j];
            for (int j = 0; j < 4; ++j) s.mot[j] = E.init_motor_omega;
            s.al[0] = s.al[1] = s.al[2] = 0.f;
            sc = 0; ret = 0.f; ep_done = false; once = false;
            vel[0] = s.v[0] + wd[0]; vel[1] = s.v[1] + wd[1]; vel[2] = s.v[2] + wd[2];
        }
        store_state(state_out, n, i, s);
        const int ebo = int((ep_done ? VF_EBIT_EPISODE_DONE : 0u) | (once ? VF_EBIT_ONCE_COLLIDED : 0u));
        reinterpret_cast<int4*>(status_out)[i] = make_int4(sc, __float_as_int(ret), ebo | (g << 8), passed);
        if (gate_out) gate_out[i] = g;
        // the rollout's all-gather, fused: this agent's episode return goes to every rank's gather buffer through
        // peer-mapped memory (NVLink); a warp stores 128 contiguous bytes per peer
        for (int r = 0; r < peers.world; ++r) peers.dst[r][peers.offset + i] = ret;

        if (E.obs_kind == VF_OBS_RACING16) {
            float o[16];
            const int g1 = g + 1 >= E.n_gates ? g + 1 - E.n_gates : g + 1;      // (g + 1) % n_gates
            for (int j = 0; j < 3; ++j) {
                o[j] = (E.gates[g][j] - s.p[j]) / 10.f;
                o[3 + j] = (E.gates[g1][j] - s.p[j]) / 10.f;
            }
            o[6] = s.q[0]; o[7] = s.q[1]; o[8] = s.q[2]; o[9] = s.q[3];
            for (int j = 0; j < 3; ++j) { o[10 + j] = vel[j] / 10.f; o[13 + j] = s.w[j] / 10.f; }
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v = make_float4(o[4 * kk], o[4 * kk + 1], o[4 * kk + 2], o[4 * kk + 3]);
                stg4(obs_out, size_t(4) * i + kk, v);
                if (mirror.obs) stg4(mirror.obs, size_t(4) * i + kk, v);
            }
        }
    }
    if (E.obs_kind == VF_OBS_STATE13) {
        float o[kObs];
        if (live) {
            o[0] = s.p[0]; o[1] = s.p[1]; o[2] = s.p[2];
            o[3] = s.q[0]; o[4] = s.q[1]; o[5] = s.q[2]; o[6] = s.q[3];
            o[7] = s.v[0] + wd[0]; o[8] = s.v[1] + wd[1]; o[9] = s.v[2] + wd[2];
            o[10] = s.w[0]; o[11] = s.w[1]; o[12] = s.w[2];
        }
        warp_store_obs(obs_out, s_obs + warp * kWarpObs, n, warp_first, lane, o, mirror.obs);
    }
    if (mirror.flag) {
        // completion word for a host that spins instead of synchronising the stream: every thread makes its host
        // stores visible system-wide, the block counts itself in, and the last block of the grid raises the flag
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned arrived = atomicAdd(mirror.counter, 1u);
            if (arrived == gridDim.x - 1) {
                *mirror.counter = 0u;               // ready for the next launch (stream-ordered behind this one)
                __threadfence_system();
                *reinterpret_cast<volatile unsigned*>(mirror.flag) = mirror.flag_value;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// wrapper tail for caller-defined tasks: collision flags, accumulation, termination, record, auto-reset
// ---------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
vf_env_finish_kernel(const __grid_constant__ VfParams params, const __grid_constant__ VfEnvSpec E, int n,
                     unsigned env_flags, unsigned long long step_index,
                     const unsigned long long* __restrict__ step_base, const float* __restrict__ state_in,
                     const float* __restrict__ wind, const float* __restrict__ reset_table,
                     const int* __restrict__ status_in, const float* __restrict__ reward_in,
                     const unsigned char* __restrict__ success_in, const unsigned char* __restrict__ failure_in,
                     float* __restrict__ state_out, int* __restrict__ status_out, float* __restrict__ obs_out,
                     unsigned char* __restrict__ done_out, float* __restrict__ record_out,
                     const __grid_constant__ VfFifoRows ring) {
    __shared__ __align__(16) float s_obs[(BLOCK / 32) * kWarpObs];
    const vf::Params<float>& P = reinterpret_cast<const vf::Params<float>&>(params);
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    const bool live = i < n;
    pdl_trigger();
    pdl_wait();
    vf::State<float> s;
    float wd[3] = {P.wind[0], P.wind[1], P.wind[2]};
    if (live) {
        load_state(state_in, n, i, s);
        const int4 st4 = __ldg(reinterpret_cast<const int4*>(status_in) + i);
        const float reward = __ldg(reward_in + i);
        const bool success = success_in && __ldg(success_in + i) != 0;
        const bool failure = failure_in && __ldg(failure_in + i) != 0;
        if (wind) {
            const float4 w4 = ldg4(wind, size_t(i));
            wd[0] = w4.x; wd[1] = w4.y; wd[2] = w4.z;
        }
        int sc = st4.x + 1;
        const unsigned eb = unsigned(st4.z) & 0xFFu;
        const vf::BoxHit<float> hit = vf::box_hit<float>(s.p, E.bbox_lo, E.bbox_hi);
        const bool is_col = hit.dis < E.uav_radius;
        bool once = (eb & VF_EBIT_ONCE_COLLIDED) || is_col;
        float ret = __int_as_float(st4.y) + reward;
        // droneGymEnv.py:188-193
        bool ep_done = (eb & VF_EBIT_EPISODE_DONE) || success || failure || hit.out || (E.collision_reset && is_col);
        const bool done = ep_done || sc >= E.max_episode_steps;
        const unsigned rbits = (done ? VF_RBIT_DONE : 0u) | (ep_done ? VF_RBIT_EPISODE_DONE : 0u) |
                               (success ? VF_RBIT_SUCCESS : 0u) | (sc >= E.max_episode_steps ? VF_RBIT_TRUNCATED : 0u) |
                               (once ? VF_RBIT_COLLIDED : 0u);
        stg4(record_out, size_t(i), make_float4(ret, float(sc), float(rbits), 0.f));
        done_out[i] = done ? 1 : 0;
        if (done && !(env_flags & VF_ENV_FLAG_NO_RESET)) {
            const unsigned long long step = step_index + (step_base ? *step_base : 0ull);
            float rp[3], rq[4], rv[3], rw[3];
            vf::sample_reset(E, E.agent_offset + unsigned(i), step, reset_table ? reset_table + size_t(i) * 13 : nullptr,
                             rp, rq, rv, rw);
            for (int j = 0; j < 3; ++j) { s.p[j] = rp[j]; s.v[j] = rv[j]; s.w[j] = rw[j]; }
            for (int j = 0; j < 4; ++j) { s.q[j] = rq[j]; s.mot[j] = E.init_motor_omega; }
            s.al[0] = s.al[1] = s.al[2] = 0.f;
            sc = 0; ret = 0.f; ep_done = false; once = false;
            for (int r = 0; r < ring.depth; ++r)         // actions still in flight to a re-initialised agent: dropped
                stg4(ring.row[r], size_t(i), make_float4(0.f, 0.f, 0.f, 0.f));
        }
        store_state(state_out, n, i, s);
        const int ebo = int((ep_done ? VF_EBIT_EPISODE_DONE : 0u) | (once ? VF_EBIT_ONCE_COLLIDED : 0u));
        reinterpret_cast<int4*>(status_out)[i] = make_int4(sc, __float_as_int(ret), ebo | (st4.z & 0xFF00), st4.w);
    }
    if (obs_out) {
        float o[kObs];
        if (live) {
            o[0] = s.p[0]; o[1] = s.p[1]; o[2] = s.p[2];
            o[3] = s.q[0]; o[4] = s.q[1]; o[5] = s.q[2]; o[6] = s.q[3];
            o[7] = s.v[0] + wd[0]; o[8] = s.v[1] + wd[1]; o[9] = s.v[2] + wd[2];
            o[10] = s.w[0]; o[11] = s.w[1]; o[12] = s.w[2];
        }
        warp_store_obs(obs_out, s_obs + warp * kWarpObs, n, warp_first, lane, o);
    }
}

// ---------------------------------------------------------------------------------------------
// reverse mode of the fused env step
// ---------------------------------------------------------------------------------------------
template <int INTEG, int ACT, bool LAG, int SMAX, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
vf_env_step_bwd_kernel(const __grid_constant__ VfParams params, const __grid_constant__ VfEnvSpec E, int n,
                       int substeps, unsigned env_flags, const float* __restrict__ state_in,
                       const float* __restrict__ action, const float* __restrict__ wind,
                       const int* __restrict__ status_in,
                       const float* __restrict__ g_state_out, const float* __restrict__ g_obs,
                       const float* __restrict__ g_reward, float* __restrict__ g_state_in,
                       float* __restrict__ g_action) {
    __shared__ __align__(16) float s_obs[(BLOCK / 32) * kWarpObs];
    const vf::Params<float>& P = reinterpret_cast<const vf::Params<float>&>(params);
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    const bool live = i < n;
    pdl_trigger();
    pdl_wait();

    float o[16];
    const bool have_obs = g_obs != nullptr;
    if (have_obs) {
        if (E.obs_kind == VF_OBS_STATE13) {
            warp_load_obs(g_obs, s_obs + warp * kWarpObs, n, warp_first, lane, o);
        } else if (live) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v = ldg4(g_obs, size_t(4) * i + kk);
                o[4 * kk] = v.x; o[4 * kk + 1] = v.y; o[4 * kk + 2] = v.z; o[4 * kk + 3] = v.w;
            }
        }
    }
    if (!live) return;

    vf::State<float> g;
    if (g_state_out) {
        load_state(g_state_out, n, i, g);
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) g.p[j] = g.v[j] = g.w[j] = g.al[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) g.q[j] = g.mot[j] = 0.f;
    }
    vf::State<float> s0;
    load_state(state_in, n, i, s0);
    const int4 st4 = __ldg(reinterpret_cast<const int4*>(status_in) + i);
    const int age_in = st4.x, gate_in = (st4.z >> 8) & 0xFF;
    float4 a4 = ldg4(action, size_t(i));
    const bool masked = age_in < E.fifo_depth;        // the forward replaced the delayed action by zero
    if (masked) a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float gr = g_reward ? __ldg(g_reward + i) : 0.f;
    float wd[3] = {P.wind[0], P.wind[1], P.wind[2]};
    if (wind) {
        const float4 w4 = ldg4(wind, size_t(i));
        wd[0] = w4.x; wd[1] = w4.y; wd[2] = w4.z;
    }
    vf::Tape<float> tape[SMAX];
    float ga[4];
    vf::env_step_bwd_agent<float>(P, E, substeps, INTEG, ACT, LAG, (env_flags & VF_ENV_FLAG_NO_RESET) != 0, a, s0,
                                  age_in, gate_in, have_obs ? o : nullptr, gr, g, ga, tape, wd);
    store_state(g_state_in, n, i, g);
    stg4(g_action, size_t(i), masked ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(ga[0], ga[1], ga[2], ga[3]));
}

// ---------------------------------------------------------------------------------------------
// layout conversion (reset / property views)
// ---------------------------------------------------------------------------------------------
__global__ void vf_pack_kernel(int n, int m, const long long* __restrict__ index,
                               const float* __restrict__ pos, const float* __restrict__ quat,
                               const float* __restrict__ vel, const float* __restrict__ rate,
                               const float* __restrict__ motor, const float* __restrict__ alpha,
                               float* __restrict__ st) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const long long i = index ? index[r] : r;
    if (i < 0 || i >= n) return;
    vf::State<float> s;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        s.p[j] = pos ? pos[size_t(r) * 3 + j] : 0.f;
        s.v[j] = vel ? vel[size_t(r) * 3 + j] : 0.f;
        s.w[j] = rate ? rate[size_t(r) * 3 + j] : 0.f;
        s.al[j] = alpha ? alpha[size_t(r) * 3 + j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s.q[j] = quat ? quat[size_t(r) * 4 + j] : (j == 0 ? 1.f : 0.f);
        s.mot[j] = motor ? motor[size_t(r) * 4 + j] : 0.f;
    }
    store_state(st, n, int(i), s);
}

__global__ void vf_unpack_kernel(int n, const float* __restrict__ st, float* __restrict__ pos,
                                 float* __restrict__ quat, float* __restrict__ vel, float* __restrict__ rate,
                                 float* __restrict__ motor, float* __restrict__ alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vf::State<float> s;
    load_state(st, n, i, s);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (pos) pos[size_t(i) * 3 + j] = s.p[j];
        if (vel) vel[size_t(i) * 3 + j] = s.v[j];
        if (rate) rate[size_t(i) * 3 + j] = s.w[j];
        if (alpha) alpha[size_t(i) * 3 + j] = s.al[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (quat) quat[size_t(i) * 4 + j] = s.q[j];
        if (motor) motor[size_t(i) * 4 + j] = s.mot[j];
    }
}

// ---------------------------------------------------------------------------------------------
// renderer hand-off: poses in Habitat-sim's frame (pure sign / axis permutation of the packed state)
// ---------------------------------------------------------------------------------------------
// Warp-cooperative store of 32 rows of W floats (row-major, W odd => lane*W+j is conflict-free in shared memory).
template <int W>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst_base, float* smem, int n, int warp_first,
                                                int lane, const float* row) {
    if (warp_first + 32 <= n) {
#pragma unroll
        for (int j = 0; j < W; ++j) smem[lane * W + j] = row[j];
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(dst_base + size_t(warp_first) * W);
        const float4* src = reinterpret_cast<const float4*>(smem);
#pragma unroll
        for (int k = 0; k < (8 * W + 31) / 32; ++k) {
            const int idx = lane + 32 * k;
            if (idx < 8 * W) dst[idx] = src[idx];
        }
        __syncwarp();
    } else if (warp_first + lane < n) {
        float* dst = dst_base + size_t(warp_first + lane) * W;
#pragma unroll
        for (int j = 0; j < W; ++j) dst[j] = row[j];
    }
}

constexpr int kPoseBlock = 128;

__global__ void __launch_bounds__(kPoseBlock)
vf_export_pose_kernel(const __grid_constant__ VfParams params, int n, const float* __restrict__ st,
                      float* __restrict__ pose_out, float* __restrict__ vel_out) {
    __shared__ __align__(16) float s_rows[(kPoseBlock / 32) * 32 * 7];
    const int i = blockIdx.x * kPoseBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warp_first = i - lane;
    float pose[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vel[3] = {0.f, 0.f, 0.f};
    if (i < n) {
        const float4 p = ldg4(st, size_t(i));
        const float4 q = ldg4(st, size_t(n) + i);
        // utils/common.py:159-176 : pos @ [[0,0,-1],[-1,0,0],[0,1,0]],  ori @ [[1,0,0,0],[0,0,0,-1],[0,-1,0,0],[0,0,1,0]]
        pose[0] = -p.y; pose[1] = p.z; pose[2] = -p.x;
        pose[3] = q.x; pose[4] = -q.z; pose[5] = q.w; pose[6] = -q.y;
        if (vel_out) {
            const float4 v = ldg4(st, size_t(2) * n + i);
            const float vx = v.x + params.wind[0], vy = v.y + params.wind[1], vz = v.z + params.wind[2];
            vel[0] = -vy; vel[1] = vz; vel[2] = -vx;
        }
    }
    warp_store_rows<7>(pose_out, s_rows + warp * 32 * 7, n, warp_first, lane, pose);
    if (vel_out) warp_store_rows<3>(vel_out, s_rows + warp * 32 * 7, n, warp_first, lane, vel);
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
constexpr int kBlock = 64;   // 65 536 agents -> 1024 CTAs -> 6.9 per SM: <= 14 warps on the fullest SM

// Tuning knob for experiments only (tools/run_kernels.py): VF_BLOCK=32|64|128|256 overrides the CTA size of the
// forward kernels.  Unset in production.
int block_override() {
    static const int v = [] {
        const char* e = std::getenv("VF_BLOCK");
        const int b = e ? std::atoi(e) : 0;
        return (b == 32 || b == 64 || b == 128 || b == 256) ? b : 0;
    }();
    return v;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_trigger / pdl_wait above).  VF_NO_PDL=1 turns
// the attribute off (plain stream order) for A/B measurements.
bool pdl_enabled() {
    static const bool v = std::getenv("VF_NO_PDL") == nullptr;
    return v;
}

template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(block));
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);   // errors surface through cudaGetLastError in the caller
}

bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15u) == 0; }

int check_common(const VfParams* params, int n, int substeps, int integrator, int action_type, bool backward = false) {
    if (!params) return fail("params is NULL");
    if (n < 0) return fail("n must be >= 0");
    if (substeps < 1) return fail("substeps must be >= 1");
    if (integrator != VF_INTEGRATOR_EULER && integrator != VF_INTEGRATOR_RK4)
        return fail("integrator must be VF_INTEGRATOR_EULER or VF_INTEGRATOR_RK4");
    if (action_type < VF_ACTION_THRUST || action_type > VF_ACTION_POSITION)
        return fail("action_type must be a VF_ACTION_* value");
    if (backward && action_type != VF_ACTION_THRUST && action_type != VF_ACTION_BODYRATE)
        return fail("no gradient for the velocity / position action types: the reference's autograd graph is broken "
                    "there (in-place writes in a per-agent loop, envs/base/dynamics.py:446-450)");
    return 0;
}

int check_spec(const VfEnvSpec* spec, bool allow_custom = false) {
    if (!spec) return fail("spec is NULL");
    if (spec->task < VF_TASK_HOVER || spec->task > (allow_custom ? VF_TASK_CUSTOM : VF_TASK_RACING))
        return fail("spec.task must be a VF_TASK_* value (VF_TASK_CUSTOM only for vf_env_finish)");
    if (spec->obs_kind != VF_OBS_STATE13 && spec->obs_kind != VF_OBS_RACING16)
        return fail("spec.obs_kind must be a VF_OBS_* value");
    if (spec->gen_kind < VF_GEN_UNIFORM || spec->gen_kind > VF_GEN_TABLE) return fail("spec.gen_kind must be a VF_GEN_* value");
    if (spec->gen_boxes < 1 || spec->gen_boxes > VF_GEN_MAX_BOXES) return fail("spec.gen_boxes out of range");
    if (spec->task == VF_TASK_RACING && (spec->n_gates < 1 || spec->n_gates > 4)) return fail("spec.n_gates out of range");
    return 0;
}

template <int INTEG, int ACT, bool LAG>
void launch_fwd(const VfParams& p, int n, int substeps, const float* si, const float* a, float* so, float* obs,
                float* ext, const float* wind, const float* push, float* copy, const VfFifoRows& ring, cudaStream_t st) {
    switch (block_override()) {
        case 32: launch_pdl(vf_step_fwd_kernel<INTEG, ACT, LAG, 32>, (n + 31) / 32, 32, st, p, n, substeps, si, a, so, obs, ext, wind, push, copy, ring); return;
        case 128: launch_pdl(vf_step_fwd_kernel<INTEG, ACT, LAG, 128>, (n + 127) / 128, 128, st, p, n, substeps, si, a, so, obs, ext, wind, push, copy, ring); return;
        case 256: launch_pdl(vf_step_fwd_kernel<INTEG, ACT, LAG, 256>, (n + 255) / 256, 256, st, p, n, substeps, si, a, so, obs, ext, wind, push, copy, ring); return;
        default: break;
    }
    const int grid = (n + kBlock - 1) / kBlock;
    launch_pdl(vf_step_fwd_kernel<INTEG, ACT, LAG, kBlock>, grid, kBlock, st, p, n, substeps, si, a, so, obs, ext, wind, push, copy, ring);
}

template <int INTEG, int ACT, bool LAG>
void launch_bwd(const VfParams& p, int n, int substeps, const float* si, const float* a, const float* gso,
                const float* gobs, float* gsi, float* ga, const float* wind, cudaStream_t st) {
    const int grid = (n + kBlock - 1) / kBlock;
    if (substeps <= 8)
        launch_pdl(vf_step_bwd_kernel<INTEG, ACT, LAG, 8, kBlock>, grid, kBlock, st, p, n, substeps, si, a, gso, gobs, gsi, ga, wind);
    else if (substeps <= 16)
        launch_pdl(vf_step_bwd_kernel<INTEG, ACT, LAG, 16, kBlock>, grid, kBlock, st, p, n, substeps, si, a, gso, gobs, gsi, ga, wind);
    else
        launch_pdl(vf_step_bwd_kernel<INTEG, ACT, LAG, VF_MAX_SUBSTEPS_BWD, kBlock>, grid, kBlock, st, p, n, substeps, si, a, gso, gobs, gsi, ga, wind);
}

template <int INTEG, int ACT, bool LAG>
void launch_env_fwd(const VfParams& p, const VfEnvSpec& e, int n, int substeps, unsigned env_flags,
                    unsigned long long step_index, const unsigned long long* step_base, const float* si,
                    const float* a, const float* wind, const float* push, const float* table, const int* status_in,
                    float* so, int* status_out, float* copy, float* obs, float* rew, unsigned char* done, float* rec,
                    float* tobs, long long* gate_out, const VfEnvMirror& mirror, const VfPeerScatter& peers,
                    cudaStream_t st) {
    const int grid = (n + kBlock - 1) / kBlock;
    launch_pdl(vf_env_step_fwd_kernel<INTEG, ACT, LAG, kBlock>, grid, kBlock, st, p, e, n, substeps, env_flags,
               step_index, step_base, si, a, wind, push, table, status_in, so, status_out, copy, obs, rew, done, rec,
               tobs, gate_out, mirror, peers);
}

template <int INTEG, int ACT, bool LAG>
void launch_env_bwd(const VfParams& p, const VfEnvSpec& e, int n, int substeps, unsigned env_flags, const float* si,
                    const float* a, const float* wind, const int* status_in, const float* gso, const float* gobs,
                    const float* gr, float* gsi, float* ga, cudaStream_t st) {
    const int grid = (n + kBlock - 1) / kBlock;
    if (substeps <= 8)
        launch_pdl(vf_env_step_bwd_kernel<INTEG, ACT, LAG, 8, kBlock>, grid, kBlock, st, p, e, n, substeps, env_flags, si, a, wind, status_in, gso, gobs, gr, gsi, ga);
    else if (substeps <= 16)
        launch_pdl(vf_env_step_bwd_kernel<INTEG, ACT, LAG, 16, kBlock>, grid, kBlock, st, p, e, n, substeps, env_flags, si, a, wind, status_in, gso, gobs, gr, gsi, ga);
    else
        launch_pdl(vf_env_step_bwd_kernel<INTEG, ACT, LAG, VF_MAX_SUBSTEPS_BWD, kBlock>, grid, kBlock, st, p, e, n, substeps, env_flags, si, a, wind, status_in, gso, gobs, gr, gsi, ga);
}

// forward-only dispatch: all four action types
#define VF_DISPATCH_FWD(FN, ...)                                                              \
    do {                                                                                      \
        if (action_type == VF_ACTION_VELOCITY || action_type == VF_ACTION_POSITION) {         \
            const bool lag = (flags & VF_FLAG_CTRL_DELAY) != 0;                               \
            if (integrator == VF_INTEGRATOR_RK4) {                                            \
                if (action_type == VF_ACTION_VELOCITY) {                                      \
                    if (lag) FN<VF_INTEGRATOR_RK4, VF_ACTION_VELOCITY, true>(__VA_ARGS__);    \
                    else     FN<VF_INTEGRATOR_RK4, VF_ACTION_VELOCITY, false>(__VA_ARGS__);   \
                } else {                                                                      \
                    if (lag) FN<VF_INTEGRATOR_RK4, VF_ACTION_POSITION, true>(__VA_ARGS__);    \
                    else     FN<VF_INTEGRATOR_RK4, VF_ACTION_POSITION, false>(__VA_ARGS__);   \
                }                                                                             \
            } else {                                                                          \
                if (action_type == VF_ACTION_VELOCITY) {                                      \
                    if (lag) FN<VF_INTEGRATOR_EULER, VF_ACTION_VELOCITY, true>(__VA_ARGS__);  \
                    else     FN<VF_INTEGRATOR_EULER, VF_ACTION_VELOCITY, false>(__VA_ARGS__); \
                } else {                                                                      \
                    if (lag) FN<VF_INTEGRATOR_EULER, VF_ACTION_POSITION, true>(__VA_ARGS__);  \
                    else     FN<VF_INTEGRATOR_EULER, VF_ACTION_POSITION, false>(__VA_ARGS__); \
                }                                                                             \
            }                                                                                 \
        } else {                                                                              \
            VF_DISPATCH(FN, __VA_ARGS__);                                                     \
        }                                                                                     \
    } while (0)

#define VF_DISPATCH(FN, ...)                                                                  \
    do {                                                                                      \
        const bool lag = (flags & VF_FLAG_CTRL_DELAY) != 0;                                   \
        if (integrator == VF_INTEGRATOR_RK4) {                                                \
            if (action_type == VF_ACTION_BODYRATE) {                                          \
                if (lag) FN<VF_INTEGRATOR_RK4, VF_ACTION_BODYRATE, true>(__VA_ARGS__);        \
                else     FN<VF_INTEGRATOR_RK4, VF_ACTION_BODYRATE, false>(__VA_ARGS__);       \
            } else {                                                                          \
                if (lag) FN<VF_INTEGRATOR_RK4, VF_ACTION_THRUST, true>(__VA_ARGS__);          \
                else     FN<VF_INTEGRATOR_RK4, VF_ACTION_THRUST, false>(__VA_ARGS__);         \
            }                                                                                 \
        } else {                                                                              \
            if (action_type == VF_ACTION_BODYRATE) {                                          \
                if (lag) FN<VF_INTEGRATOR_EULER, VF_ACTION_BODYRATE, true>(__VA_ARGS__);      \
                else     FN<VF_INTEGRATOR_EULER, VF_ACTION_BODYRATE, false>(__VA_ARGS__);     \
            } else {                                                                          \
                if (lag) FN<VF_INTEGRATOR_EULER, VF_ACTION_THRUST, true>(__VA_ARGS__);        \
                else     FN<VF_INTEGRATOR_EULER, VF_ACTION_THRUST, false>(__VA_ARGS__);       \
            }                                                                                 \
        }                                                                                     \
    } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int vf_abi_version(void) { return VF_ABI_VERSION; }

const char* vf_last_error(void) { return g_last_error.c_str(); }

int vf_params_size(void) { return int(sizeof(VfParams)); }

int vf_device_sm_count(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

int vf_step_fwd(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
                const float* state_in, const float* action, float* state_out, float* obs_out, float* ext_out,
                const float* wind, const float* fifo_push, float* fifo_copy, void* stream) {
    if (check_common(params, n, substeps, integrator, action_type)) return 1;
    if (n == 0) return 0;
    if (!state_in || !action || !state_out) return fail("state_in, action and state_out must not be NULL");
    if (state_in == state_out) return fail("state_out must not alias state_in");
    if ((fifo_push == nullptr) != (fifo_copy == nullptr)) return fail("fifo_push and fifo_copy must be given together");
    if (fifo_copy && fifo_copy == action) return fail("fifo_copy must not alias the consumed action");
    if (!aligned16(state_in) || !aligned16(action) || !aligned16(state_out) || !aligned16(obs_out) ||
        !aligned16(ext_out) || !aligned16(wind) || !aligned16(fifo_push) || !aligned16(fifo_copy))
        return fail("all buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const VfFifoRows ring = {};
    VF_DISPATCH_FWD(launch_fwd, *params, n, substeps, state_in, action, state_out, obs_out, ext_out, wind, fifo_push,
                    fifo_copy, ring, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_step_fwd launch failed", err);
    return 0;
}

static int check_ring(const VfFifoRows* r) {
    if (r->depth < 1 || r->depth > VF_FIFO_MAX_ROWS) return fail("fifo_rows: depth must be in 1..VF_FIFO_MAX_ROWS");
    for (int j = 0; j < r->depth; ++j) {
        if (!r->row[j] || !aligned16(r->row[j])) return fail("fifo_rows: every row must be a 16-byte aligned device pointer");
        for (int k = 0; k < j; ++k)
            if (r->row[k] == r->row[j]) return fail("fifo_rows: rows must be distinct buffers");
    }
    return 0;
}

int vf_step_fwd_ring(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
                     const float* state_in, const VfFifoRows* fifo_rows, const float* fifo_push, float* state_out,
                     float* obs_out, float* ext_out, const float* wind, void* stream) {
    if (check_common(params, n, substeps, integrator, action_type)) return 1;
    if (n == 0) return 0;
    if (!state_in || !state_out || !fifo_rows || !fifo_push)
        return fail("state_in, state_out, fifo_rows and fifo_push must not be NULL");
    if (state_in == state_out) return fail("state_out must not alias state_in");
    if (check_ring(fifo_rows)) return 1;
    for (int j = 0; j < fifo_rows->depth; ++j)
        if (fifo_rows->row[j] == fifo_push) return fail("fifo_push must not alias a FIFO row");
    if (!aligned16(state_in) || !aligned16(state_out) || !aligned16(obs_out) || !aligned16(ext_out) || !aligned16(wind) ||
        !aligned16(fifo_push))
        return fail("all buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* action = nullptr;
    float* fifo_copy = nullptr;
    VF_DISPATCH_FWD(launch_fwd, *params, n, substeps, state_in, action, state_out, obs_out, ext_out, wind, fifo_push,
                    fifo_copy, *fifo_rows, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_step_fwd_ring launch failed", err);
    return 0;
}

int vf_step_bwd(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
                const float* state_in, const float* action, const float* grad_state_out, const float* grad_obs,
                float* grad_state_in, float* grad_action, const float* wind, void* stream) {
    if (check_common(params, n, substeps, integrator, action_type, true)) return 1;
    if (substeps > VF_MAX_SUBSTEPS_BWD) return fail("substeps exceeds VF_MAX_SUBSTEPS_BWD for the reverse sweep");
    if (n == 0) return 0;
    if (!state_in || !action || !grad_state_in || !grad_action)
        return fail("state_in, action, grad_state_in and grad_action must not be NULL");
    if (!aligned16(state_in) || !aligned16(action) || !aligned16(grad_state_out) || !aligned16(grad_obs) ||
        !aligned16(grad_state_in) || !aligned16(grad_action) || !aligned16(wind))
        return fail("all buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VF_DISPATCH(launch_bwd, *params, n, substeps, state_in, action, grad_state_out, grad_obs, grad_state_in,
                grad_action, wind, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_step_bwd launch failed", err);
    return 0;
}

int vf_step_fwd_host(const VfParams* params, int n, int substeps, int integrator, int action_type,
                     unsigned flags, const float* state_in, const float* action_host, float* action_dev,
                     float* state_out, float* obs_dev, float* obs_host, void* stream) {
    if (!action_host || !action_dev) return fail("action_host and action_dev must not be NULL");
    if ((obs_host == nullptr) != (obs_dev == nullptr)) return fail("obs_host and obs_dev must be given together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t err = cudaMemcpyAsync(action_dev, action_host, sizeof(float) * 4 * size_t(n),
                                      cudaMemcpyHostToDevice, st);
    if (err != cudaSuccess) return fail("vf_step_fwd_host: H2D copy of the actions failed", err);
    if (vf_step_fwd(params, n, substeps, integrator, action_type, flags, state_in, action_dev, state_out, obs_dev,
                    nullptr, nullptr, nullptr, nullptr, stream))
        return 1;
    if (obs_host) {
        err = cudaMemcpyAsync(obs_host, obs_dev, sizeof(float) * VF_OBS_FLOATS * size_t(n),
                              cudaMemcpyDeviceToHost, st);
        if (err != cudaSuccess) return fail("vf_step_fwd_host: D2H copy of the observation failed", err);
    }
    err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) return fail("vf_step_fwd_host: stream synchronise failed", err);
    return 0;
}

int vf_env_spec_size(void) { return int(sizeof(VfEnvSpec)); }

int vf_wait_flag(const volatile unsigned* flag, unsigned value, long long timeout_us) {
    if (!flag) return fail("vf_wait_flag: flag is NULL");
    if (*flag == value) return 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1;; ++spins) {
        if (*flag == value) return 0;
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if (timeout_us > 0 && (spins & 1023u) == 0) {
            const auto dt = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
            if (dt.count() > timeout_us) return fail("vf_wait_flag: timed out waiting for the kernel's completion word");
        }
    }
}

int vf_env_step_fwd(const VfParams* params, const VfEnvSpec* spec, int n, int substeps, int integrator,
                    int action_type, unsigned flags, unsigned env_flags, unsigned long long step_index,
                    const unsigned long long* step_base,
                    const float* state_in, const float* action, const float* wind, const float* fifo_push,
                    const float* reset_table, const int* status_in,
                    float* state_out, int* status_out, float* fifo_copy, float* obs_out, float* reward_out,
                    unsigned char* done_out, float* record_out, float* term_obs_out, long long* gate_out,
                    const VfEnvMirror* host_mirror, const VfPeerScatter* peer_returns, void* stream) {
    if (check_common(params, n, substeps, integrator, action_type)) return 1;
    if (check_spec(spec)) return 1;
    VfPeerScatter peers = {};
    if (peer_returns && peer_returns->world > 0) {
        if (peer_returns->world > VF_MAX_PEERS) return fail("peer_returns.world exceeds VF_MAX_PEERS");
        if (peer_returns->offset < 0) return fail("peer_returns.offset must be >= 0");
        for (int r = 0; r < peer_returns->world; ++r)
            if (!peer_returns->dst[r]) return fail("peer_returns.dst has a NULL entry");
        peers = *peer_returns;
    }
    if (n == 0) return 0;
    if (!state_in || !action || !state_out || !status_in || !status_out || !obs_out || !reward_out || !done_out ||
        !record_out)
        return fail("vf_env_step_fwd: a required buffer is NULL");
    if (spec->gen_kind == VF_GEN_TABLE && !reset_table) return fail("VF_GEN_TABLE needs reset_table");
    if (state_in == state_out) return fail("state_out must not alias state_in");
    if ((fifo_push == nullptr) != (fifo_copy == nullptr)) return fail("fifo_push and fifo_copy must be given together");
    if (fifo_copy && fifo_copy == action) return fail("fifo_copy must not alias the consumed action");
    if (!aligned16(state_in) || !aligned16(action) || !aligned16(state_out) || !aligned16(obs_out) ||
        !aligned16(record_out) || !aligned16(status_in) || !aligned16(status_out) || !aligned16(wind) ||
        !aligned16(fifo_push) || !aligned16(fifo_copy))
        return fail("all float4 / int4 buffers must be 16-byte aligned");
    // page-locked host destinations -> their device aliases (identity under unified addressing; an error here means
    // the caller passed pageable memory)
    VfEnvMirror mirror = {nullptr, nullptr, nullptr, nullptr, nullptr, 0u};
    if (host_mirror) {
        void* d = nullptr;
        if (host_mirror->obs) {
            if (cudaHostGetDevicePointer(&d, host_mirror->obs, 0) != cudaSuccess || !aligned16(d)) {
                (void)cudaGetLastError();
                return fail("host_mirror.obs must be page-locked (cudaHostAlloc / cudaHostRegister) and 16-byte aligned");
            }
            mirror.obs = static_cast<float*>(d);
        }
        if (host_mirror->reward) {
            if (cudaHostGetDevicePointer(&d, host_mirror->reward, 0) != cudaSuccess) {
                (void)cudaGetLastError();
                return fail("host_mirror.reward must be page-locked host memory");
            }
            mirror.reward = static_cast<float*>(d);
        }
        if (host_mirror->done) {
            if (cudaHostGetDevicePointer(&d, host_mirror->done, 0) != cudaSuccess) {
                (void)cudaGetLastError();
                return fail("host_mirror.done must be page-locked host memory");
            }
            mirror.done = static_cast<int*>(d);
        }
        if (host_mirror->flag) {
            if (!host_mirror->counter) return fail("host_mirror.flag needs host_mirror.counter (a zeroed device word)");
            if (cudaHostGetDevicePointer(&d, host_mirror->flag, 0) != cudaSuccess) {
                (void)cudaGetLastError();
                return fail("host_mirror.flag must be page-locked host memory");
            }
            mirror.flag = static_cast<unsigned*>(d);
            mirror.counter = host_mirror->counter;
            mirror.flag_value = host_mirror->flag_value;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool prefetch = [] {
        const char* e = getenv("VF_NO_PREFETCH");
        return !(e && e[0] && e[0] != '0');
    }();
    env_flags = (env_flags & ~kEnvFlagPrefetch) | (prefetch ? kEnvFlagPrefetch : 0u);
    VF_DISPATCH_FWD(launch_env_fwd, *params, *spec, n, substeps, env_flags, step_index, step_base, state_in, action,
                    wind, fifo_push, reset_table, status_in, state_out, status_out, fifo_copy, obs_out, reward_out,
                    done_out, record_out, term_obs_out, gate_out, mirror, peers, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_env_step_fwd launch failed", err);
    return 0;
}

int vf_env_step_bwd(const VfParams* params, const VfEnvSpec* spec, int n, int substeps, int integrator,
                    int action_type, unsigned flags, unsigned env_flags, const float* state_in, const float* action,
                    const float* wind, const int* status_in, const float* grad_state_out, const float* grad_obs,
                    const float* grad_reward, float* grad_state_in, float* grad_action, void* stream) {
    if (check_common(params, n, substeps, integrator, action_type, true)) return 1;
    if (check_spec(spec)) return 1;
    if (substeps > VF_MAX_SUBSTEPS_BWD) return fail("substeps exceeds VF_MAX_SUBSTEPS_BWD for the reverse sweep");
    if (n == 0) return 0;
    if (!state_in || !action || !status_in || !grad_state_in || !grad_action)
        return fail("vf_env_step_bwd: a required buffer is NULL");
    if (!aligned16(state_in) || !aligned16(action) || !aligned16(grad_state_out) || !aligned16(grad_obs) ||
        !aligned16(grad_state_in) || !aligned16(grad_action) || !aligned16(status_in) || !aligned16(wind))
        return fail("all buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VF_DISPATCH(launch_env_bwd, *params, *spec, n, substeps, env_flags, state_in, action, wind, status_in,
                grad_state_out, grad_obs, grad_reward, grad_state_in, grad_action, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_env_step_bwd launch failed", err);
    return 0;
}

int vf_env_finish(const VfParams* params, const VfEnvSpec* spec, int n, unsigned env_flags,
                  unsigned long long step_index, const unsigned long long* step_base, const float* state_in,
                  const float* wind, const float* reset_table, const int* status_in, const float* reward,
                  const unsigned char* success, const unsigned char* failure, float* state_out, int* status_out,
                  float* obs_out, unsigned char* done_out, float* record_out, const VfFifoRows* fifo_rows,
                  void* stream) {
    if (!params) return fail("params is NULL");
    if (check_spec(spec, true)) return 1;
    if (fifo_rows && check_ring(fifo_rows)) return 1;
    const VfFifoRows ring = fifo_rows ? *fifo_rows : VfFifoRows{};
    if (n < 0) return fail("n must be >= 0");
    if (n == 0) return 0;
    if (!state_in || !status_in || !reward || !state_out || !status_out || !done_out || !record_out)
        return fail("vf_env_finish: a required buffer is NULL");
    if (spec->gen_kind == VF_GEN_TABLE && !reset_table) return fail("VF_GEN_TABLE needs reset_table");
    if (state_in == state_out) return fail("state_out must not alias state_in");
    if (!aligned16(state_in) || !aligned16(state_out) || !aligned16(status_in) || !aligned16(status_out) ||
        !aligned16(obs_out) || !aligned16(record_out) || !aligned16(wind))
        return fail("all float4 / int4 buffers must be 16-byte aligned");
    const int grid = (n + kBlock - 1) / kBlock;
    launch_pdl(vf_env_finish_kernel<kBlock>, grid, kBlock, static_cast<cudaStream_t>(stream), *params, *spec, n,
               env_flags, step_index, step_base, state_in, wind, reset_table, status_in, reward, success, failure,
               state_out, status_out, obs_out, done_out, record_out, ring);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_env_finish launch failed", err);
    return 0;
}

int vf_export_pose_habitat(const VfParams* params, int n, const float* state, float* pose_out, float* vel_out,
                           void* stream) {
    if (!params) return fail("params is NULL");
    if (n < 0) return fail("n must be >= 0");
    if (n == 0) return 0;
    if (!state || !pose_out) return fail("state and pose_out must not be NULL");
    // page-locked host destinations resolve to their device alias; device pointers pass through unchanged
    float* outs[2] = {pose_out, vel_out};
    for (float*& o : outs) {
        if (!o) continue;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, o) != cudaSuccess) {
            (void)cudaGetLastError();
            return fail("vf_export_pose_habitat: output is neither device nor page-locked host memory");
        }
        if (attr.type == cudaMemoryTypeHost) o = static_cast<float*>(attr.devicePointer);
        else if (attr.type == cudaMemoryTypeUnregistered)
            return fail("vf_export_pose_habitat: output is pageable host memory (use cudaHostAlloc / pin_memory)");
        if (!aligned16(o)) return fail("vf_export_pose_habitat: outputs must be 16-byte aligned");
    }
    if (!aligned16(state)) return fail("state must be 16-byte aligned");
    vf_export_pose_kernel<<<(n + kPoseBlock - 1) / kPoseBlock, kPoseBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        *params, n, state, outs[0], outs[1]);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_export_pose_habitat launch failed", err);
    return 0;
}

int vf_pack_state(int n, int m, const long long* index, const float* pos, const float* quat, const float* vel,
                  const float* rate, const float* motor, const float* alpha, float* state, void* stream) {
    if (n < 0 || m < 0) return fail("n and m must be >= 0");
    if (m == 0) return 0;
    if (!state || !aligned16(state)) return fail("state must be a 16-byte aligned device pointer");
    vf_pack_kernel<<<(m + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(n, m, index, pos, quat, vel,
                                                                                   rate, motor, alpha, state);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_pack_state launch failed", err);
    return 0;
}

int vf_unpack_state(int n, const float* state, float* pos, float* quat, float* vel, float* rate, float* motor,
                    float* alpha, void* stream) {
    if (n < 0) return fail("n must be >= 0");
    if (n == 0) return 0;
    if (!state || !aligned16(state)) return fail("state must be a 16-byte aligned device pointer");
    vf_unpack_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(n, state, pos, quat, vel,
                                                                                     rate, motor, alpha);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("vf_unpack_state launch failed", err);
    return 0;
}

}  // extern "C"
