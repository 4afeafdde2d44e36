// vf_env.cuh — per-agent arithmetic of the env wrapper tail fused behind the control step (see the "Fused env
// step" block of include/visfly_b200.h for the reference lines each piece follows).
#pragma once

#include "vf_math.cuh"

namespace vf {

// ---------------------------------------------------------------------------------------------
// counter-based RNG for the reset sampler: Philox4x32-10 (Salmon et al., SC'11).  Key = seed, counter =
// (agent, draw, step_lo, step_hi): every (agent, step) pair owns an independent stream, results do not
// depend on launch geometry or on how agents are sharded over GPUs.
// ---------------------------------------------------------------------------------------------
struct Philox {
    unsigned c[4];
    unsigned k[2];
};
VF_HD unsigned mulhi32(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
VF_HD void philox_round(Philox& s) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const unsigned hi0 = mulhi32(M0, s.c[0]), lo0 = M0 * s.c[0];
    const unsigned hi1 = mulhi32(M1, s.c[2]), lo1 = M1 * s.c[2];
    const unsigned n0 = hi1 ^ s.c[1] ^ s.k[0], n2 = hi0 ^ s.c[3] ^ s.k[1];
    s.c[0] = n0; s.c[1] = lo1; s.c[2] = n2; s.c[3] = lo0;
}
VF_HD void philox4x32(unsigned long long seed, unsigned agent, unsigned draw, unsigned long long step, unsigned out[4]) {
    Philox s;
    s.c[0] = agent; s.c[1] = draw; s.c[2] = (unsigned)step; s.c[3] = (unsigned)(step >> 32);
    s.k[0] = (unsigned)seed; s.k[1] = (unsigned)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        philox_round(s);
        s.k[0] += 0x9E3779B9u; s.k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = s.c[i];
}
// uniform in [0,1) with 24 random bits (what torch.rand produces for float32)
VF_HD float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------
// analytic bounding box (reference droneEnv.py:345-369): nearest face, its distance, out-of-bounds test
// ---------------------------------------------------------------------------------------------
template <class T> struct BoxHit {
    int axis;       // coordinate in which the nearest face lies
    T delta;        // collision_vector[axis] = face - p[axis]  (all other components are 0)
    T dis;          // |delta|
    bool out;       // is_out_bounds
};
// component `axis` of a 3-vector held in registers: selects, not an indexed access (a register array indexed at run
// time is demoted to local memory together with the struct around it — measured: it cost the forward kernel a
// 152-byte stack frame and 45 local stores per agent)
template <class T> VF_HD T pick3(const T v[3], int axis) { return axis == 0 ? v[0] : (axis == 1 ? v[1] : v[2]); }

template <class T> VF_HD BoxHit<T> box_hit(const T p[3], const float lo[3], const float hi[3]) {
    BoxHit<T> h;
    // gaps to the six faces in the reference's order [p-lo (x,y,z), hi-p (x,y,z)]; first minimum wins (torch.min)
    const T gap[6] = {p[0] - T(lo[0]), p[1] - T(lo[1]), p[2] - T(lo[2]),
                      T(hi[0]) - p[0], T(hi[1]) - p[1], T(hi[2]) - p[2]};
    T best = gap[0];
    int face = 0;
    for (int j = 1; j < 6; ++j)
        if (gap[j] < best) { best = gap[j]; face = j; }
    h.axis = face < 3 ? face : face - 3;
    // collision_vector[axis] = wall - p[axis]: lo - p = -(p - lo) and hi - p are the gaps themselves (exact)
    h.delta = face < 3 ? -best : best;
    // the reference takes the 2-norm of a vector with one non-zero component: sqrt(delta^2) = |delta| (exact in
    // binary floating point short of underflow of the square)
    h.dis = vabs(best);
    h.out = (p[0] < T(lo[0])) | (p[1] < T(lo[1])) | (p[2] < T(lo[2])) | (p[0] > T(hi[0])) | (p[1] > T(hi[1])) |
            (p[2] > T(hi[2]));
    return h;
}

template <class T> VF_HD T norm3(T x, T y, T z) { return vsqrt(x * x + y * y + z * z); }

// hover-style shaping shared by HoverEnv (HoverEnv.py:83-94) and RacingEnv (RacingEnv.py:203-215)
template <class T>
VF_HD T reward_hover(const T p[3], const T q[4], const T v[3], const T w[3], const T tgt[3]) {
    const T d = norm3(p[0] - tgt[0], p[1] - tgt[1], p[2] - tgt[2]);
    const T qe = vsqrt((q[0] - T(1)) * (q[0] - T(1)) + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    return T(0.1) + d * T(-0.1 * 1 / 9) + qe * T(-0.00001) + norm3(v[0], v[1], v[2]) * T(-0.002) +
           norm3(w[0], w[1], w[2]) * T(-0.002);
}

VF_HD float vacos(float x) { return acosf(x); }
VF_HD double vacos(double x) { return acos(x); }

// NavigationEnv.py:85-99, term by term
template <class T>
VF_HD T reward_navigation(const T p[3], const T q[4], const T v[3], const T w[3], const T tgt[3],
                          const BoxHit<T>& hit, bool success, int max_steps, int step_count) {
    const T tx = tgt[0] - p[0], ty = tgt[1] - p[1], tz = tgt[2] - p[2];
    const T vn = norm3(v[0], v[1], v[2]);
    T approach = (v[0] * tx + v[1] * ty + v[2] * tz) / (T(1e-6) + norm3(tx, ty, tz));
    approach = approach > T(10) ? T(10) : approach;
    // heading direction = x axis of the body frame (maths.py:129-131)
    const T dx = T(1) - T(2) * (q[2] * q[2] + q[3] * q[3]);
    const T dy = T(2) * (q[1] * q[2] + q[3] * q[0]);
    const T dz = T(2) * (q[1] * q[3] - q[2] * q[0]);
    const T thrd = T(3.14159265358979323846 / 18);
    T c = (dx * v[0] + dy * v[1] + dz * v[2]) / (T(1e-6) + vn) / T(1);
    c = vclamp(c, T(-1), T(1));
    T ang = vacos(c);
    ang = (ang < thrd ? thrd : ang) - thrd;
    const T qe = vsqrt((q[0] - T(1)) * (q[0] - T(1)) + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const T stable = qe * T(-0.00001) + vn * T(-0.002) + norm3(w[0], w[1], w[2]) * T(-0.002);
    const T near = T(1) / (hit.dis + T(0.2)) * T(-0.01);
    T prox = T(1) - hit.dis;
    prox = prox > T(0) ? prox : T(0);
    T closing = (hit.delta * pick3(v, hit.axis)) / (T(1e-6) + hit.dis);
    closing = closing > T(0) ? closing : T(0);
    const T bonus = success ? T(max_steps - step_count) * T(0.1) * (T(0.2) + T(0.8) / (T(1) + T(1) * vn)) : T(0);
    return T(0.1) * T(0) + approach * T(0.01) + ang * T(-0.01) + stable + near + prox * closing * T(-0.005) + bonus;
}

// ---- adjoints of the rewards (torch.autograd conventions: ||x|| has gradient 0 at x = 0, clamp / clamp_max /
// clamp_min pass the gradient on the closed side, relu'(0) = 0).  One deliberate deviation: where
// the reference's heading term would produce acos'(+-1) = -inf times a zero mask (NaN), the gradient is 0 here. ----
template <class T>
VF_HD void reward_hover_adj(const T p[3], const T q[4], const T v[3], const T w[3], const T tgt[3], T gr,
                            T gp[3], T gq[4], T gv[3], T gw[3]) {
    const T dx = p[0] - tgt[0], dy = p[1] - tgt[1], dz = p[2] - tgt[2];
    const T d = norm3(dx, dy, dz);
    if (d > T(0)) {
        const T k = gr * T(-0.1 * 1 / 9) / d;
        gp[0] += k * dx; gp[1] += k * dy; gp[2] += k * dz;
    }
    const T qe = vsqrt((q[0] - T(1)) * (q[0] - T(1)) + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (qe > T(0)) {
        const T k = gr * T(-0.00001) / qe;
        gq[0] += k * (q[0] - T(1)); gq[1] += k * q[1]; gq[2] += k * q[2]; gq[3] += k * q[3];
    }
    const T vn = norm3(v[0], v[1], v[2]);
    if (vn > T(0)) {
        const T k = gr * T(-0.002) / vn;
        gv[0] += k * v[0]; gv[1] += k * v[1]; gv[2] += k * v[2];
    }
    const T wn = norm3(w[0], w[1], w[2]);
    if (wn > T(0)) {
        const T k = gr * T(-0.002) / wn;
        gw[0] += k * w[0]; gw[1] += k * w[1]; gw[2] += k * w[2];
    }
}

template <class T>
VF_HD void reward_navigation_adj(const T p[3], const T q[4], const T v[3], const T w[3], const T tgt[3],
                                 const BoxHit<T>& hit, bool success, int max_steps, int step_count, T gr,
                                 T gp[3], T gq[4], T gv[3], T gw[3]) {
    const T t[3] = {tgt[0] - p[0], tgt[1] - p[1], tgt[2] - p[2]};
    const T tn = norm3(t[0], t[1], t[2]);
    const T vn = norm3(v[0], v[1], v[2]);
    // approach speed
    {
        const T den = T(1e-6) + tn, dot = v[0] * t[0] + v[1] * t[1] + v[2] * t[2];
        if (dot / den <= T(10)) {
            const T g = gr * T(0.01);
            for (int i = 0; i < 3; ++i) {
                gv[i] += g * t[i] / den;
                T gt = g * v[i] / den;
                if (tn > T(0)) gt -= g * dot / (den * den) * t[i] / tn;
                gp[i] -= gt;
            }
        }
    }
    // heading alignment
    {
        const T d[3] = {T(1) - T(2) * (q[2] * q[2] + q[3] * q[3]), T(2) * (q[1] * q[2] + q[3] * q[0]),
                        T(2) * (q[1] * q[3] - q[2] * q[0])};
        const T thrd = T(3.14159265358979323846 / 18);
        const T denv = T(1e-6) + vn, dv = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
        const T c0 = dv / denv;
        const T c = vclamp(c0, T(-1), T(1));
        const T ang0 = vacos(c);
        if (ang0 >= thrd && c0 >= T(-1) && c0 <= T(1) && c > T(-1) && c < T(1)) {
            const T gc0 = gr * T(-0.01) * (T(-1) / vsqrt(T(1) - c * c));
            T gd[3];
            for (int i = 0; i < 3; ++i) {
                gd[i] = gc0 * v[i] / denv;
                T g = gc0 * d[i] / denv;
                if (vn > T(0)) g -= gc0 * dv / (denv * denv) * v[i] / vn;
                gv[i] += g;
            }
            // d = (1-2(y^2+z^2), 2(xy+zw), 2(xz-yw)) with q = (w,x,y,z)
            gq[0] += gd[1] * T(2) * q[3] - gd[2] * T(2) * q[2];
            gq[1] += gd[1] * T(2) * q[2] + gd[2] * T(2) * q[3];
            gq[2] += -gd[0] * T(4) * q[2] + gd[1] * T(2) * q[1] - gd[2] * T(2) * q[0];
            gq[3] += -gd[0] * T(4) * q[3] + gd[1] * T(2) * q[0] + gd[2] * T(2) * q[1];
        }
    }
    // stability terms
    {
        const T qe = vsqrt((q[0] - T(1)) * (q[0] - T(1)) + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (qe > T(0)) {
            const T k = gr * T(-0.00001) / qe;
            gq[0] += k * (q[0] - T(1)); gq[1] += k * q[1]; gq[2] += k * q[2]; gq[3] += k * q[3];
        }
        if (vn > T(0)) { const T k = gr * T(-0.002) / vn; for (int i = 0; i < 3; ++i) gv[i] += k * v[i]; }
        const T wn = norm3(w[0], w[1], w[2]);
        if (wn > T(0)) { const T k = gr * T(-0.002) / wn; for (int i = 0; i < 3; ++i) gw[i] += k * w[i]; }
    }
    // obstacle proximity and closing speed.  collision_vector = point.detach() - p: value (0,..,delta,..,0) but
    // d(vector_j)/dp_j = -1 for every j (droneEnv.py:351-365)
    {
        const int ax = hit.axis;
        const T dis = hit.dis, delta = hit.delta;
        T g_dis = gr * T(0.01) / ((dis + T(0.2)) * (dis + T(0.2)));
        const T den = T(1e-6) + dis;
        const T num = delta * pick3(v, ax);
        const T cl0 = num / den;
        const T A = T(1) - dis > T(0) ? T(1) - dis : T(0);
        const T Bc = cl0 > T(0) ? cl0 : T(0);
        if (T(1) - dis > T(0)) g_dis -= gr * T(-0.005) * Bc;
        if (cl0 > T(0)) {
            const T g_cl0 = gr * T(-0.005) * A;
            for (int j = 0; j < 3; ++j) gp[j] -= g_cl0 * v[j] / den;      // d num / d p_j = -v_j
            for (int j = 0; j < 3; ++j)
                if (j == ax) gv[j] += g_cl0 * delta / den;               // d num / d v_ax = delta
            g_dis -= g_cl0 * num / (den * den);
        }
        if (dis > T(0))
            for (int j = 0; j < 3; ++j)
                if (j == ax) gp[j] -= g_dis * delta / dis;               // dis = |delta|, d delta / d p_ax = -1
    }
    // success bonus
    if (success && vn > T(0)) {
        const T k = gr * T(max_steps - step_count) * T(0.1) * (T(-0.8) / ((T(1) + vn) * (T(1) + vn))) / vn;
        for (int i = 0; i < 3; ++i) gv[i] += k * v[i];
    }
}

// ---- everything the wrapper decides from the state after the control step (shared by forward, backward, mirror) ----
template <class T> struct EnvEval {
    T vel[3];            // reported velocity = v + wind
    BoxHit<T> hit;
    bool is_col, success, pass, ep_done, done;
    int gate;            // next gate index after this step's gate test (racing)
    T reward;
};

// `wd`: this agent's wind of the control step (the constant P.wind, or the per-agent vector of the wind functions)
template <class T>
VF_HD void env_eval(const Params<T>& P, const VfEnvSpec& E, const State<T>& s, const T wd[3], int sc, int gate_in,
                    bool ep_done_in, EnvEval<T>& ev) {
    for (int j = 0; j < 3; ++j) ev.vel[j] = s.v[j] + wd[j];
    ev.hit = box_hit<T>(s.p, E.bbox_lo, E.bbox_hi);
    ev.is_col = ev.hit.dis < T(E.uav_radius);
    ev.success = false;
    ev.pass = false;
    ev.gate = gate_in;
    T tgt[3];
    if (E.task == VF_TASK_HOVER) {
        for (int j = 0; j < 3; ++j) tgt[j] = T(E.target[j]);
        ev.reward = reward_hover<T>(s.p, s.q, ev.vel, s.w, tgt);
    } else if (E.task == VF_TASK_NAVIGATION) {
        for (int j = 0; j < 3; ++j) tgt[j] = T(E.target[j]);
        ev.success = norm3(s.p[0] - tgt[0], s.p[1] - tgt[1], s.p[2] - tgt[2]) <= T(E.success_radius);
        ev.reward = reward_navigation<T>(s.p, s.q, ev.vel, s.w, tgt, ev.hit, ev.success, E.max_episode_steps, sc);
    } else {
        for (int j = 0; j < 3; ++j) tgt[j] = T(E.gates[gate_in][j]);
        ev.pass = norm3(s.p[0] - tgt[0], s.p[1] - tgt[1], s.p[2] - tgt[2]) <= T(E.success_radius);
        ev.gate = gate_in + (ev.pass ? 1 : 0);
        if (ev.gate >= E.n_gates) ev.gate -= E.n_gates;          // (gate_in + pass) % n_gates with gate_in < n_gates
        for (int j = 0; j < 3; ++j) tgt[j] = T(E.gates[ev.gate][j]);
        ev.reward = reward_hover<T>(s.p, s.q, ev.vel, s.w, tgt) + (ev.pass ? T(20) : T(0));
    }
    ev.ep_done = ep_done_in || ev.success || ev.hit.out || (E.collision_reset && ev.is_col);
    ev.done = ev.ep_done || sc >= E.max_episode_steps;
}

// Reverse mode of one fused env step for one agent.
//   a        : delayed action as the forward saw it (already zeroed when the agent was younger than the FIFO)
//   age_in, gate_in : step count / gate index at the start of the step
//   g        : in = dL/d(packed state returned by the step), out = dL/d(s0)
//   gobs     : dL/d(observation returned by the step) (13 or 16 wide);  gr : dL/d(reward)
// A finished agent was re-initialised inside the step: its returned state / observation are constants, only the
// reward (computed before the reset) carries gradient (reference: in-place overwrite, dynamics.py:249-263).
template <class T>
VF_HD void env_step_bwd_agent(const Params<T>& P, const VfEnvSpec& E, int substeps, int integrator, int action_type,
                              bool ctrl_delay, bool no_reset, const T a[4], const State<T>& s0, int age_in,
                              int gate_in, const T* gobs, T gr, State<T>& g, T ga[4], Tape<T>* tape,
                              const T* wind = nullptr) {
    Command<T> c;
    State<T> s_raw;
    step_fwd_taped(P, substeps, integrator, action_type, ctrl_delay, a, s0, c, s_raw, tape, wind);
    State<T> s = s_raw;
    clamp_state(P, s);
    EnvEval<T> ev;
    env_eval(P, E, s, wind ? wind : P.wind, age_in + 1, gate_in, false, ev);
    if (ev.done && !no_reset) {
        for (int j = 0; j < 3; ++j) g.p[j] = g.v[j] = g.w[j] = g.al[j] = T(0);
        for (int j = 0; j < 4; ++j) g.q[j] = g.mot[j] = T(0);
    } else if (gobs) {
        if (E.obs_kind == VF_OBS_STATE13) {
            for (int j = 0; j < 3; ++j) { g.p[j] += gobs[j]; g.v[j] += gobs[7 + j]; g.w[j] += gobs[10 + j]; }
            for (int j = 0; j < 4; ++j) g.q[j] += gobs[3 + j];
        } else {
            for (int j = 0; j < 3; ++j) {
                g.p[j] -= (gobs[j] + gobs[3 + j]) / T(10);
                g.v[j] += gobs[10 + j] / T(10);
                g.w[j] += gobs[13 + j] / T(10);
            }
            for (int j = 0; j < 4; ++j) g.q[j] += gobs[6 + j];
        }
    }
    T tgt[3];
    if (E.task == VF_TASK_NAVIGATION) {
        for (int j = 0; j < 3; ++j) tgt[j] = T(E.target[j]);
        reward_navigation_adj<T>(s.p, s.q, ev.vel, s.w, tgt, ev.hit, ev.success, E.max_episode_steps, age_in + 1, gr,
                                 g.p, g.q, g.v, g.w);
    } else {
        for (int j = 0; j < 3; ++j) tgt[j] = E.task == VF_TASK_HOVER ? T(E.target[j]) : T(E.gates[ev.gate][j]);
        reward_hover_adj<T>(s.p, s.q, ev.vel, s.w, tgt, gr, g.p, g.q, g.v, g.w);
    }
    step_bwd_taped(P, substeps, integrator, action_type, ctrl_delay, s0, c, s_raw, tape, g, ga);
}

// first gate from where the agent stands (RacingEnv.py:173-185)
template <class T> VF_HD int racing_first_gate(const T p[3]) {
    const T rx = p[0] - T(4), ry = p[1] - T(0);
    if (rx < T(0)) return ry > T(0) ? 0 : 3;
    return rx > T(0) ? 1 : 2;
}

// euler (roll, pitch, yaw) -> quaternion, order zyx (maths.py:257-269)
VF_HD void euler_to_quat(const float e[3], float q[4]) {
    const float cr = cosf(e[0] * 0.5f), sr = sinf(e[0] * 0.5f);
    const float cp = cosf(e[1] * 0.5f), sp = sinf(e[1] * 0.5f);
    const float cy = cosf(e[2] * 0.5f), sy = sinf(e[2] * 0.5f);
    q[0] = cr * cp * cy + sr * sp * sy;
    q[1] = sr * cp * cy - cr * sp * sy;
    q[2] = cr * sp * cy + sr * cp * sy;
    q[3] = cr * cp * sy - sr * sp * cy;
}

// Fresh initial state of one agent (position, quaternion, velocity, body rates).
// (An out-of-line version of this function was measured and dropped, see VF_HD_COLD in vf_math.cuh.)
VF_HD_COLD void sample_reset(const VfEnvSpec& E, unsigned agent, unsigned long long step, const float* table_row,
                        float p[3], float q[4], float v[3], float w[3]) {
    if (E.gen_kind == VF_GEN_TABLE) {
        for (int j = 0; j < 3; ++j) { p[j] = table_row[j]; v[j] = table_row[7 + j]; w[j] = table_row[10 + j]; }
        for (int j = 0; j < 4; ++j) q[j] = table_row[3 + j];
        return;
    }
    unsigned r[16];
    for (int d = 0; d < 4; ++d) philox4x32(E.seed, agent, (unsigned)d, step, r + 4 * d);
    const int box = E.gen_boxes > 1 ? (int)(r[12] % (unsigned)E.gen_boxes) : 0;
    float f[12];
    if (E.gen_kind == VF_GEN_NORMAL) {
        // Box-Muller on pairs; the reference draws (2*randn - 1) * std + mean (randomization.py:201-204)
        unsigned r2[12];
        for (int d = 0; d < 3; ++d) philox4x32(E.seed, agent, (unsigned)(4 + d), step, r2 + 4 * d);
        for (int j = 0; j < 12; ++j) {
            const float u1 = 1.0f - u01(r[j]), u2 = u01(r2[j]);
            const float z = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
            f[j] = 2.0f * z - 1.0f;
        }
    } else {
        for (int j = 0; j < 12; ++j) f[j] = 2.0f * u01(r[j]) - 1.0f;
    }
    float e[3], off[3];
    for (int j = 0; j < 3; ++j) {
        off[j] = f[j] * E.gen_half[box][0][j];
        p[j] = off[j] + E.gen_mean[box][0][j];
        e[j] = f[3 + j] * E.gen_half[box][1][j] + E.gen_mean[box][1][j];
        v[j] = f[6 + j] * E.gen_half[box][2][j] + E.gen_mean[box][2][j];
        w[j] = f[9 + j] * E.gen_half[box][3][j] + E.gen_mean[box][3][j];
    }
    if (E.gen_heading[box]) {
        // randomization.py:162-165 with calculate_yaw_pitch (:27-28): the yaw points from the drawn position back to
        // the centre of the position box, roll = pitch = 0, plus the orientation noise (the orientation mean is unused)
        const float dx = -off[0], dy = -off[1];
        const float yaw = acosf(dx / sqrtf(dx * dx + dy * dy)) * (dy >= 0.f ? 1.f : -1.f);
        e[0] = f[3] * E.gen_half[box][1][0];
        e[1] = f[4] * E.gen_half[box][1][1];
        e[2] = yaw + f[5] * E.gen_half[box][1][2];
    }
    euler_to_quat(e, q);
}

}  // namespace vf
