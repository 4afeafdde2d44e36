// TEST INFRASTRUCTURE — host instantiation of visfly_b200/csrc/vf_math.cuh.
//
// Compiles the exact per-agent arithmetic the sm_100a kernels run (same header, same templates) with g++
// for float and double, behind the same packed-state layout as the C-ABI in include/visfly_b200.h.  It exists
// so that the hand-derived adjoint can be checked against torch.autograd (oracle/torch_oracle.py) and the
// forward against the golden vectors on a machine with no GPU, before GPU minutes are spent.
//
// It is NOT part of the product: visfly_b200 never loads this library (the product has no CPU path and
// fails loudly without the CUDA extension).  Only tests/ use it.
#include <cstddef>
#include <vector>

#include "../visfly_b200/csrc/vf_env.cuh"
#include "../visfly_b200/csrc/vf_math.cuh"

namespace {

template <class T> void load_state(const T* packed, int n, int i, vf::State<T>& s) {
    const T* p0 = packed + (size_t(0) * n + i) * 4;
    const T* p1 = packed + (size_t(1) * n + i) * 4;
    const T* p2 = packed + (size_t(2) * n + i) * 4;
    const T* p3 = packed + (size_t(3) * n + i) * 4;
    const T* p4 = packed + (size_t(4) * n + i) * 4;
    for (int k = 0; k < 3; ++k) { s.p[k] = p0[k]; s.v[k] = p2[k]; s.w[k] = p3[k]; }
    for (int k = 0; k < 4; ++k) { s.q[k] = p1[k]; s.mot[k] = p4[k]; }
    s.al[0] = p0[3]; s.al[1] = p2[3]; s.al[2] = p3[3];
}

template <class T> void store_state(T* packed, int n, int i, const vf::State<T>& s) {
    T* p0 = packed + (size_t(0) * n + i) * 4;
    T* p1 = packed + (size_t(1) * n + i) * 4;
    T* p2 = packed + (size_t(2) * n + i) * 4;
    T* p3 = packed + (size_t(3) * n + i) * 4;
    T* p4 = packed + (size_t(4) * n + i) * 4;
    for (int k = 0; k < 3; ++k) { p0[k] = s.p[k]; p2[k] = s.v[k]; p3[k] = s.w[k]; }
    for (int k = 0; k < 4; ++k) { p1[k] = s.q[k]; p4[k] = s.mot[k]; }
    p0[3] = s.al[0]; p2[3] = s.al[1]; p3[3] = s.al[2];
}

template <class T>
void fwd(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
         const T* state_in, const T* action, T* state_out, T* obs_out, T* ext_out) {
    const vf::Params<T> P(*params);
    const bool lag = flags & VF_FLAG_CTRL_DELAY;
    for (int i = 0; i < n; ++i) {
        vf::State<T> s;
        load_state(state_in, n, i, s);
        vf::Wrench<T> k;
        vf::step_fwd<T>(P, substeps, integrator, action_type, lag, action + size_t(i) * 4, s, k);
        store_state(state_out, n, i, s);
        if (obs_out) {
            T* o = obs_out + size_t(i) * VF_OBS_FLOATS;
            for (int j = 0; j < 3; ++j) { o[j] = s.p[j]; o[7 + j] = s.v[j] + P.wind[j]; o[10 + j] = s.w[j]; }
            for (int j = 0; j < 4; ++j) o[3 + j] = s.q[j];
        }
        if (ext_out) {
            T* e = ext_out + size_t(i) * VF_EXT_FLOATS;
            for (int j = 0; j < 3; ++j) e[j] = k.acc[j];
            e[3] = T(0);
            for (int j = 0; j < 4; ++j) e[4 + j] = k.thr[j];
        }
    }
}

template <class T>
int bwd(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
        const T* state_in, const T* action, const T* g_state_out, const T* g_obs, T* g_state_in,
        T* g_action) {
    if (substeps > VF_MAX_SUBSTEPS_BWD) return 1;
    const vf::Params<T> P(*params);
    const bool lag = flags & VF_FLAG_CTRL_DELAY;
    std::vector<vf::Tape<T>> tape(substeps);
    for (int i = 0; i < n; ++i) {
        vf::State<T> s0, g;
        load_state(state_in, n, i, s0);
        if (g_state_out) {
            load_state(g_state_out, n, i, g);
        } else {
            g = vf::State<T>();
            for (int j = 0; j < 3; ++j) g.p[j] = g.v[j] = g.w[j] = g.al[j] = T(0);
            for (int j = 0; j < 4; ++j) g.q[j] = g.mot[j] = T(0);
        }
        if (g_obs) {
            const T* o = g_obs + size_t(i) * VF_OBS_FLOATS;
            for (int j = 0; j < 3; ++j) { g.p[j] += o[j]; g.v[j] += o[7 + j]; g.w[j] += o[10 + j]; }
            for (int j = 0; j < 4; ++j) g.q[j] += o[3 + j];
        }
        T ga[4];
        vf::step_bwd<T>(P, substeps, integrator, action_type, lag, action + size_t(i) * 4, s0, g, ga,
                        tape.data());
        store_state(g_state_in, n, i, g);
        for (int j = 0; j < 4; ++j) g_action[size_t(i) * 4 + j] = ga[j];
    }
    return 0;
}

// One fused env step WITHOUT the reset (reward / flags of the step + the pre-reset state), and its adjoint.
template <class T>
void env_fwd(const VfParams* params, const VfEnvSpec* E, int n, int substeps, int integrator, int action_type,
             unsigned flags, const T* state_in, const T* action, const int* saved, T* state_out, T* reward,
             int* done, int* gate_out) {
    const vf::Params<T> P(*params);
    const bool lag = flags & VF_FLAG_CTRL_DELAY;
    for (int i = 0; i < n; ++i) {
        vf::State<T> s;
        load_state(state_in, n, i, s);
        T a[4];
        for (int j = 0; j < 4; ++j) a[j] = saved[2 * i] < E->fifo_depth ? T(0) : action[size_t(i) * 4 + j];
        vf::Wrench<T> k;
        vf::step_fwd<T>(P, substeps, integrator, action_type, lag, a, s, k);
        vf::EnvEval<T> ev;
        vf::env_eval<T>(P, *E, s, P.wind, saved[2 * i] + 1, saved[2 * i + 1], false, ev);
        store_state(state_out, n, i, s);
        reward[i] = ev.reward;
        done[i] = ev.done ? 1 : 0;
        gate_out[i] = ev.gate;
    }
}

template <class T>
void env_bwd(const VfParams* params, const VfEnvSpec* E, int n, int substeps, int integrator, int action_type,
             unsigned flags, unsigned env_flags, const T* state_in, const T* action, const int* saved,
             const T* g_state_out, const T* g_obs, const T* g_reward, T* g_state_in, T* g_action) {
    const vf::Params<T> P(*params);
    const bool lag = flags & VF_FLAG_CTRL_DELAY;
    const int width = E->obs_kind == VF_OBS_STATE13 ? 13 : 16;
    std::vector<vf::Tape<T>> tape(substeps);
    for (int i = 0; i < n; ++i) {
        vf::State<T> s0, g;
        load_state(state_in, n, i, s0);
        if (g_state_out) {
            load_state(g_state_out, n, i, g);
        } else {
            for (int j = 0; j < 3; ++j) g.p[j] = g.v[j] = g.w[j] = g.al[j] = T(0);
            for (int j = 0; j < 4; ++j) g.q[j] = g.mot[j] = T(0);
        }
        const bool masked = saved[2 * i] < E->fifo_depth;
        T a[4], ga[4];
        for (int j = 0; j < 4; ++j) a[j] = masked ? T(0) : action[size_t(i) * 4 + j];
        vf::env_step_bwd_agent<T>(P, *E, substeps, integrator, action_type, lag,
                                  (env_flags & VF_ENV_FLAG_NO_RESET) != 0, a, s0, saved[2 * i], saved[2 * i + 1],
                                  g_obs ? g_obs + size_t(i) * width : nullptr, g_reward ? g_reward[i] : T(0), g, ga,
                                  tape.data());
        store_state(g_state_in, n, i, g);
        for (int j = 0; j < 4; ++j) g_action[size_t(i) * 4 + j] = masked ? T(0) : ga[j];
    }
}

}  // namespace

extern "C" {

void vfm_env_fwd_f64(const VfParams* p, const VfEnvSpec* e, int n, int s, int integ, int at, unsigned fl,
                     const double* si, const double* a, const int* saved, double* so, double* rew, int* done,
                     int* gate) {
    env_fwd<double>(p, e, n, s, integ, at, fl, si, a, saved, so, rew, done, gate);
}
void vfm_env_bwd_f64(const VfParams* p, const VfEnvSpec* e, int n, int s, int integ, int at, unsigned fl, unsigned efl,
                     const double* si, const double* a, const int* saved, const double* gso, const double* gobs,
                     const double* gr, double* gsi, double* ga) {
    env_bwd<double>(p, e, n, s, integ, at, fl, efl, si, a, saved, gso, gobs, gr, gsi, ga);
}
void vfm_env_bwd_f32(const VfParams* p, const VfEnvSpec* e, int n, int s, int integ, int at, unsigned fl, unsigned efl,
                     const float* si, const float* a, const int* saved, const float* gso, const float* gobs,
                     const float* gr, float* gsi, float* ga) {
    env_bwd<float>(p, e, n, s, integ, at, fl, efl, si, a, saved, gso, gobs, gr, gsi, ga);
}


void vfm_step_fwd_f32(const VfParams* p, int n, int s, int integ, int at, unsigned fl, const float* si,
                      const float* a, float* so, float* obs, float* ext) {
    fwd<float>(p, n, s, integ, at, fl, si, a, so, obs, ext);
}
void vfm_step_fwd_f64(const VfParams* p, int n, int s, int integ, int at, unsigned fl, const double* si,
                      const double* a, double* so, double* obs, double* ext) {
    fwd<double>(p, n, s, integ, at, fl, si, a, so, obs, ext);
}
int vfm_step_bwd_f32(const VfParams* p, int n, int s, int integ, int at, unsigned fl, const float* si,
                     const float* a, const float* gso, const float* gobs, float* gsi, float* ga) {
    return bwd<float>(p, n, s, integ, at, fl, si, a, gso, gobs, gsi, ga);
}
int vfm_step_bwd_f64(const VfParams* p, int n, int s, int integ, int at, unsigned fl, const double* si,
                     const double* a, const double* gso, const double* gobs, double* gsi, double* ga) {
    return bwd<double>(p, n, s, integ, at, fl, si, a, gso, gobs, gsi, ga);
}

}  // extern "C"
