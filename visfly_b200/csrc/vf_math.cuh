// vf_math.cuh — per-agent arithmetic of one quadrotor control step and of its hand-derived adjoint.
//
// One agent = one thread.  Everything here is straight-line scalar code on values that live in registers
// for the whole control step (all `substeps` integration sub-steps), so a step costs one HBM round trip.
// The file is written against a scalar type T so that the very same source can be instantiated
//   * as float inside the sm_100a kernels (vf_kernels.cu), and
//   * as float/double on the host by the test-only mirror under oracle/ (used to check the adjoint on a
//     machine without a GPU; never linked into the product library).
//
// Reference semantics followed (all paths relative to the VisFly tree):
//   command de-normalisation      envs/base/dynamics.py:704-713
//   body-rate PID -> rotor thrust envs/base/dynamics.py:400-413, clamp :501
//   rotor speed model             envs/base/dynamics.py:505-554
//   allocation, drag, acceleration envs/base/dynamics.py:339-349
//   Euler / RK4 update            utils/maths.py:300-389   (RK4 with repairs R1-R3)
//   renormalisation               envs/base/dynamics.py:367
//   post-step clamps              envs/base/dynamics.py:374-382
// What is deliberately different from the reference's op sequence (same real-number function, fewer
// flops, rounding-level differences only):
//   * q (x) (0,u) (x) q*  is evaluated as (w^2-r.r) u + 2 (r.u) r + 2 w (r x u), which is the same
//     polynomial for non-unit q (the reference does not assume |q| = 1 inside the sandwich);
//   * the desired rotor speed (one sqrt per rotor) is hoisted out of the sub-step loop: the reference
//     recomputes the identical value every sub-step (dynamics.py:511);
//   * w x (J w) uses the diagonal-J closed form;
//   * RK4 of v and p: acceleration is frozen over the four stages (dynamics.py:347,358), so the stage
//     sums collapse to v += a dt and p += (v + wind) dt + a dt^2 / 2.
#pragma once

#include "../../include/visfly_b200.h"

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
// "cold" helpers (the reset sampler) stay INLINE: an out-of-line call made the per-step text contiguous, but a kernel
// with a call stack no longer overlaps its launch with the previous grid under programmatic dependent launch — the
// step loop lost 1.1 us per step (hot 10.5 -> 11.7 us, cold 12.4 -> 13.8; tools/ab_loop.py, same box)
#define VF_HD_COLD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#define VF_HD_COLD inline
#endif

#if !defined(__CUDACC__)
#include <cmath>
#endif

namespace vf {

// ---------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------
VF_HD float  vsqrt(float x) { return sqrtf(x); }
VF_HD double vsqrt(double x) { return sqrt(x); }
VF_HD float  vabs(float x) { return fabsf(x); }
VF_HD double vabs(double x) { return fabs(x); }
VF_HD float  vsin(float x) { return sinf(x); }
VF_HD double vsin(double x) { return sin(x); }
VF_HD float  vcos(float x) { return cosf(x); }
VF_HD double vcos(double x) { return cos(x); }
VF_HD float  vatan2(float y, float x) { return atan2f(y, x); }
VF_HD double vatan2(double y, double x) { return atan2(y, x); }
// 1/sqrt(x): one MUFU.RSQ and one Newton step on the device (<= 1 ulp, no slow-path branches, instead of an IEEE
// sqrt followed by an IEEE division); plain 1/sqrt on the host mirror.
VF_HD float vrsqrt(float x) {
#if defined(__CUDA_ARCH__)
    const float y = rsqrtf(x);
    return y * (1.5f - (0.5f * x) * y * y);
#else
    return 1.0f / sqrtf(x);
#endif
}
VF_HD double vrsqrt(double x) { return 1.0 / sqrt(x); }
template <class T> VF_HD T vclamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
// torch.clamp backward: gradient passes on the closed interval [lo, hi]
template <class T> VF_HD T vgate(T x, T lo, T hi, T g) { return (x >= lo && x <= hi) ? g : T(0); }

// Parameters converted to the working scalar type (float: bit copy of VfParams).
template <class T> struct Params {
    T dt, mass, inv_mass;
    T J[3], J_inv[3];
    T B[16], B_inv[16];
    T thrust_map[3];
    T motor_c, thrust_min, thrust_max;
    T k_lin[3], k_quad[3];
    T JKp[9], Kd[9];
    T act_half[4], act_mean[4];
    T gravity[3], wind[3];
    T pos_lo[3], pos_hi[3];
    T vel_lim, rate_lim;
    T Kp[9];
    T vel_kp, vel_kd, pos_kd;

    Params() = default;
    explicit Params(const VfParams& s) {
        const float* src = reinterpret_cast<const float*>(&s);
        T* dst = reinterpret_cast<T*>(this);
        for (unsigned i = 0; i < sizeof(VfParams) / sizeof(float); ++i) dst[i] = T(src[i]);
    }
};
static_assert(sizeof(Params<float>) == sizeof(VfParams), "Params<float> must mirror VfParams");

// The 20 floats that carry across control steps.
template <class T> struct State {
    T p[3];    // position (world)
    T q[4];    // orientation w,x,y,z
    T v[3];    // ground velocity (world)
    T w[3];    // body rates
    T mot[4];  // rotor speeds
    T al[3];   // angular acceleration applied in the last sub-step
};

// ---------------------------------------------------------------------------------------------
// quaternion sandwich  y = Im( q (x) (0,u) (x) q* )       (sgn = +1)   reference maths.py:38
//                      y = Im( q* (x) (0,u) (x) q )       (sgn = -1)   reference maths.py:49
// valid for any q (scales by |q|^2), exactly like the reference's two Hamilton products.
// ---------------------------------------------------------------------------------------------
template <class T> VF_HD void sandwich(const T q[4], const T u[3], T sgn, T y[3]) {
    const T w = q[0], x = q[1], yy = q[2], z = q[3];
    const T A = w * w - (x * x + yy * yy + z * z);
    const T ru2 = T(2) * (x * u[0] + yy * u[1] + z * u[2]);
    const T sw2 = T(2) * sgn * w;
    const T cx = yy * u[2] - z * u[1];
    const T cy = z * u[0] - x * u[2];
    const T cz = x * u[1] - yy * u[0];
    y[0] = A * u[0] + ru2 * x + sw2 * cx;
    y[1] = A * u[1] + ru2 * yy + sw2 * cy;
    y[2] = A * u[2] + ru2 * z + sw2 * cz;
}

// The matrix of the same map, M(q) = (w^2 - r.r) I + 2 r r^T + 2 w [r]x  (= |q|^2 R(q), exact for non-unit q):
//   sandwich(q, u, +1) = M u        sandwich(q, u, -1) = M^T u
// One sub-step rotates twice with the same q (world->body for the drag, body->world for the force), so the ten
// quaternion products are formed once: 22 + 9 + 9 operations instead of 2 x 25.
template <class T> struct RotMat {
    T m[3][3];
};
template <class T> VF_HD void rotmat(const T q[4], RotMat<T>& R) {
    const T w = q[0], x = q[1], y = q[2], z = q[3];
    const T A = w * w - (x * x + y * y + z * z);
    const T x2 = x + x, y2 = y + y, z2 = z + z;
    const T wx = w * x2, wy = w * y2, wz = w * z2;
    R.m[0][0] = A + x2 * x;  R.m[1][1] = A + y2 * y;  R.m[2][2] = A + z2 * z;
    R.m[0][1] = x2 * y - wz; R.m[1][0] = x2 * y + wz;
    R.m[0][2] = x2 * z + wy; R.m[2][0] = x2 * z - wy;
    R.m[1][2] = y2 * z - wx; R.m[2][1] = y2 * z + wx;
}
template <class T> VF_HD void rot_fwd(const RotMat<T>& R, const T u[3], T y[3]) {      // body -> world
    for (int i = 0; i < 3; ++i) y[i] = R.m[i][0] * u[0] + R.m[i][1] * u[1] + R.m[i][2] * u[2];
}
template <class T> VF_HD void rot_inv(const RotMat<T>& R, const T u[3], T y[3]) {      // world -> body
    for (int i = 0; i < 3; ++i) y[i] = R.m[0][i] * u[0] + R.m[1][i] * u[1] + R.m[2][i] * u[2];
}

// adjoint of sandwich: accumulates into gq (w,x,y,z) and writes gu
template <class T>
VF_HD void sandwich_adj(const T q[4], const T u[3], T sgn, const T gy[3], T gq[4], T gu[3]) {
    const T w = q[0];
    const T r[3] = {q[1], q[2], q[3]};
    // gu = sandwich with the opposite sign (the inverse rotation, same |q|^2 scale)
    sandwich(q, gy, -sgn, gu);
    const T ugy = u[0] * gy[0] + u[1] * gy[1] + u[2] * gy[2];
    const T ru = r[0] * u[0] + r[1] * u[1] + r[2] * u[2];
    const T rgy = r[0] * gy[0] + r[1] * gy[1] + r[2] * gy[2];
    // (r x u) . gy
    const T rxu_gy = (r[1] * u[2] - r[2] * u[1]) * gy[0] + (r[2] * u[0] - r[0] * u[2]) * gy[1] +
                     (r[0] * u[1] - r[1] * u[0]) * gy[2];
    gq[0] += T(2) * (w * ugy + sgn * rxu_gy);
    // u x gy
    const T c[3] = {u[1] * gy[2] - u[2] * gy[1], u[2] * gy[0] - u[0] * gy[2], u[0] * gy[1] - u[1] * gy[0]};
    const T sw = sgn * w;
    for (int i = 0; i < 3; ++i) gq[1 + i] += T(2) * (ru * gy[i] + rgy * u[i] - ugy * r[i] + sw * c[i]);
}

// quaternion kinematics  g = 1/2 q (x) (0,w)            reference maths.py:311
template <class T> VF_HD void qdot(const T q[4], const T w[3], T g[4]) {
    g[0] = T(-0.5) * (q[1] * w[0] + q[2] * w[1] + q[3] * w[2]);
    g[1] = T(0.5) * (q[0] * w[0] + q[2] * w[2] - q[3] * w[1]);
    g[2] = T(0.5) * (q[0] * w[1] - q[1] * w[2] + q[3] * w[0]);
    g[3] = T(0.5) * (q[0] * w[2] + q[1] * w[1] - q[2] * w[0]);
}
// q (x) (0,w) = 2 * qdot: the forward pass folds the 1/2 into its step sizes (exact, a power of two)
template <class T> VF_HD void qdot2(const T q[4], const T w[3], T g[4]) {
    g[0] = -(q[1] * w[0] + q[2] * w[1] + q[3] * w[2]);
    g[1] = q[0] * w[0] + q[2] * w[2] - q[3] * w[1];
    g[2] = q[0] * w[1] - q[1] * w[2] + q[3] * w[0];
    g[3] = q[0] * w[2] + q[1] * w[1] - q[2] * w[0];
}
template <class T> VF_HD void qdot_adj(const T q[4], const T w[3], const T gg[4], T gq[4], T gw[3]) {
    const T h = T(0.5);
    // g0 = -h (q1 w0 + q2 w1 + q3 w2)
    // g1 =  h (q0 w0 + q2 w2 - q3 w1)
    // g2 =  h (q0 w1 - q1 w2 + q3 w0)
    // g3 =  h (q0 w2 + q1 w1 - q2 w0)
    gq[0] += h * (gg[1] * w[0] + gg[2] * w[1] + gg[3] * w[2]);
    gq[1] += h * (-gg[0] * w[0] - gg[2] * w[2] + gg[3] * w[1]);
    gq[2] += h * (-gg[0] * w[1] + gg[1] * w[2] - gg[3] * w[0]);
    gq[3] += h * (-gg[0] * w[2] - gg[1] * w[1] + gg[2] * w[0]);
    gw[0] += h * (-gg[0] * q[1] + gg[1] * q[0] + gg[2] * q[3] - gg[3] * q[2]);
    gw[1] += h * (-gg[0] * q[2] - gg[1] * q[3] + gg[2] * q[0] + gg[3] * q[1]);
    gw[2] += h * (-gg[0] * q[3] + gg[1] * q[2] - gg[2] * q[1] + gg[3] * q[0]);
}

// gyroscopic term  c = w x (J w), J diagonal              reference maths.py:314, dynamics.py:406
template <class T> VF_HD void gyro(const Params<T>& P, const T w[3], T c[3]) {
    c[0] = (P.J[2] - P.J[1]) * w[1] * w[2];
    c[1] = (P.J[0] - P.J[2]) * w[2] * w[0];
    c[2] = (P.J[1] - P.J[0]) * w[0] * w[1];
}
template <class T> VF_HD void gyro_adj(const Params<T>& P, const T w[3], const T gc[3], T gw[3]) {
    const T a = (P.J[2] - P.J[1]) * gc[0], b = (P.J[0] - P.J[2]) * gc[1], c = (P.J[1] - P.J[0]) * gc[2];
    gw[0] += b * w[2] + c * w[1];
    gw[1] += a * w[2] + c * w[0];
    gw[2] += a * w[1] + b * w[0];
}

// body-rate dynamics  f = J^-1 (tau - w x J w)            reference maths.py:314
template <class T> VF_HD void wdot(const Params<T>& P, const T w[3], const T tau[3], T f[3]) {
    T c[3];
    gyro(P, w, c);
    for (int i = 0; i < 3; ++i) f[i] = P.J_inv[i] * (tau[i] - c[i]);
}
// the same with J^-1 tau formed once per sub-step (tau is frozen over the RK4 stages) and the inertia ratios as
// constants: 2 operations per component
template <class T> struct WdotCoef {
    T jt[3];   // J^-1 tau
    T g[3];    // J^-1 (J_k - J_j)
};
template <class T> VF_HD void wdot_coef(const Params<T>& P, const T tau[3], WdotCoef<T>& c) {
    c.g[0] = P.J_inv[0] * (P.J[2] - P.J[1]);
    c.g[1] = P.J_inv[1] * (P.J[0] - P.J[2]);
    c.g[2] = P.J_inv[2] * (P.J[1] - P.J[0]);
    for (int i = 0; i < 3; ++i) c.jt[i] = P.J_inv[i] * tau[i];
}
template <class T> VF_HD void wdot_lean(const WdotCoef<T>& c, const T w[3], T f[3]) {
    f[0] = c.jt[0] - c.g[0] * (w[1] * w[2]);
    f[1] = c.jt[1] - c.g[1] * (w[2] * w[0]);
    f[2] = c.jt[2] - c.g[2] * (w[0] * w[1]);
}
// adjoint: accumulates gw and gtau
template <class T>
VF_HD void wdot_adj(const Params<T>& P, const T w[3], const T gf[3], T gw[3], T gtau[3]) {
    T h[3], gc[3];
    for (int i = 0; i < 3; ++i) {
        h[i] = P.J_inv[i] * gf[i];
        gtau[i] += h[i];
        gc[i] = -h[i];
    }
    gyro_adj(P, w, gc, gw);
}

// ---------------------------------------------------------------------------------------------
// once per control step: action -> clamped desired rotor thrusts -> desired rotor speeds
// ---------------------------------------------------------------------------------------------
template <class T> struct Command {
    T t_pre[4];   // desired rotor thrusts before the clamp (kept for the clamp's gradient gate)
    T t_des[4];   // clamped                                             dynamics.py:501
    T w_des[4];   // desired rotor speeds                                dynamics.py:545-554
    T disc[4];    // sqrt(b^2 - 4 a (c - T_des))  (= 1 / dW_des/dT_des)
    T w_in[4];    // (1 - c_m) * w_des: what the first-order rotor lag adds every sub-step
};

// Outer loops of the velocity / position action types: set-point -> desired force -> geometric attitude
// controller -> [collective thrust, body torque]                     reference dynamics.py:414-496.
// Forward only (the reference's autograd graph is broken on this branch: per-agent in-place writes, :446-450).
template <class T>
VF_HD void geometric_cmd(const Params<T>& P, int action_type, const T a[4], const T p[3], const T q[4],
                         const T v[3], const T w[3], const T al[3], T ft[4]) {
    T cmd[4];                                                     // [yaw, x, y, z]               :714-729
    for (int i = 0; i < 4; ++i) cmd[i] = a[i] * P.act_half[i] + P.act_mean[i];
    const T qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    const T yaw = vatan2(T(2) * (qw * qz + qx * qy), T(1) - T(2) * (qy * qy + qz * qz));   // maths.py:248
    T a_des[3], yaw_des, yaw_gain;
    if (action_type == VF_ACTION_VELOCITY) {
        for (int i = 0; i < 3; ++i) a_des[i] = P.vel_kp * (cmd[1 + i] - v[i]);              // :416
        const T vn = vsqrt(v[0] * v[0] + v[1] * v[1]);                                      // :421
        yaw_des = vn > T(0.1) ? vatan2(v[1], v[0]) : yaw;                                   // :423-427
        yaw_gain = P.vel_kd * T(2);                                                         // :433
    } else {
        for (int i = 0; i < 3; ++i) {
            const T v_des = P.pos_kd * (cmd[1 + i] - p[i]);                                 // :457
            a_des[i] = P.vel_kd * (v_des - v[i]);                                           // :458
        }
        yaw_des = cmd[0];                                                                   // :462
        yaw_gain = P.pos_kd * T(2);                                                         // :469
    }
    T F[3];
    for (int i = 0; i < 3; ++i) F[i] = P.mass * (a_des[i] - P.gravity[i]);                  // :417 / :459
    T e = yaw_des - yaw;                                                                    // :430-432
    e = vatan2(vsin(e), vcos(e));
    const T yaw_spd = e * yaw_gain;
    T Fb[3];
    sandwich(q, F, T(-1), Fb);                                                              // :435 transform()
    ft[0] = Fb[2];
    // R(q), unit-norm form                                                                 maths.py:113-117
    const T R[3][3] = {
        {T(1) - T(2) * (qy * qy + qz * qz), T(2) * (qx * qy - qz * qw), T(2) * (qx * qz + qy * qw)},
        {T(2) * (qx * qy + qz * qw), T(1) - T(2) * (qx * qx + qz * qz), T(2) * (qy * qz - qx * qw)},
        {T(2) * (qx * qz - qy * qw), T(2) * (qy * qz + qx * qw), T(1) - T(2) * (qx * qx + qy * qy)}};
    // desired body frame                                                                   :437-442
    const T Fn = vsqrt(F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
    const T b3[3] = {F[0] / Fn, F[1] / Fn, F[2] / Fn};
    const T c1[3] = {vcos(yaw_des), vsin(yaw_des), T(0)};
    T b2[3] = {b3[1] * c1[2] - b3[2] * c1[1], b3[2] * c1[0] - b3[0] * c1[2], b3[0] * c1[1] - b3[1] * c1[0]};
    const T b2n = vsqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
    for (int i = 0; i < 3; ++i) b2[i] = b2[i] / b2n;
    const T b1[3] = {b2[1] * b3[2] - b2[2] * b3[1], b2[2] * b3[0] - b2[0] * b3[2], b2[0] * b3[1] - b2[1] * b3[0]};
    // M = R_des^T R  (R_des has columns b1 b2 b3)
    const T* bb[3] = {b1, b2, b3};
    T M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = bb[i][0] * R[0][j] + bb[i][1] * R[1][j] + bb[i][2] * R[2][j];
    // attitude error  -vee(0.5 (M - M^T))  and rate error                                  :446-450
    const T pe[3] = {T(0.5) * (M[1][2] - M[2][1]), -T(0.5) * (M[0][2] - M[2][0]), T(0.5) * (M[0][1] - M[1][0])};
    T ave[3];
    for (int i = 0; i < 3; ++i) ave[i] = M[i][2] * yaw_spd - w[i];
    T gy[3] = {T(0), T(0), T(0)};
    T rate_gain = T(1);
    if (action_type == VF_ACTION_POSITION) {                                                // :486-491
        gyro(P, w, gy);
        rate_gain = T(1.2);
    }
    for (int i = 0; i < 3; ++i) {
        T t = P.Kp[3 * i] * pe[0] + P.Kp[3 * i + 1] * pe[1] + P.Kp[3 * i + 2] * pe[2] +
              rate_gain * (P.Kp[3 * i] * ave[0] + P.Kp[3 * i + 1] * ave[1] + P.Kp[3 * i + 2] * ave[2]);
        if (action_type == VF_ACTION_POSITION)
            t -= (P.Kd[3 * i] * al[0] + P.Kd[3 * i + 1] * al[1] + P.Kd[3 * i + 2] * al[2]) + gy[i];
        ft[1 + i] = P.J[i] * t;                                                             // :452 (J diagonal)
    }
}

template <class T>
VF_HD void command_fwd(const Params<T>& P, int action_type, const T a[4], const State<T>& s, Command<T>& c) {
    const T* w = s.w;
    const T* al = s.al;
    if (action_type == VF_ACTION_VELOCITY || action_type == VF_ACTION_POSITION) {
        T ft[4];
        geometric_cmd(P, action_type, a, s.p, s.q, s.v, s.w, s.al, ft);
        for (int i = 0; i < 4; ++i)
            c.t_pre[i] = P.B_inv[4 * i] * ft[0] + P.B_inv[4 * i + 1] * ft[1] + P.B_inv[4 * i + 2] * ft[2] +
                         P.B_inv[4 * i + 3] * ft[3];
    } else if (action_type == VF_ACTION_BODYRATE) {
        // dynamics.py:705-707 : collective thrust [N] and body-rate set-points [rad/s]
        T ft[4];
        ft[0] = (a[0] * P.act_half[0] + P.act_mean[0]) * P.mass;
        T e[3];
        for (int i = 0; i < 3; ++i) e[i] = (a[1 + i] * P.act_half[1 + i] + P.act_mean[1 + i]) - w[i];
        T gy[3];
        gyro(P, w, gy);
        // dynamics.py:404-407 : tau_des = J Kp e + w x J w - Kd alpha
        for (int i = 0; i < 3; ++i) {
            ft[1 + i] = P.JKp[3 * i] * e[0] + P.JKp[3 * i + 1] * e[1] + P.JKp[3 * i + 2] * e[2] + gy[i] -
                        (P.Kd[3 * i] * al[0] + P.Kd[3 * i + 1] * al[1] + P.Kd[3 * i + 2] * al[2]);
        }
        // dynamics.py:412-413
        for (int i = 0; i < 4; ++i)
            c.t_pre[i] = P.B_inv[4 * i] * ft[0] + P.B_inv[4 * i + 1] * ft[1] + P.B_inv[4 * i + 2] * ft[2] +
                         P.B_inv[4 * i + 3] * ft[3];
    } else {
        // dynamics.py:713 : per-rotor thrust command
        for (int i = 0; i < 4; ++i) c.t_pre[i] = P.mass * (a[i] * P.act_half[i] + P.act_mean[i]);
    }
    const T ta = P.thrust_map[0], tb = P.thrust_map[1], tc = P.thrust_map[2];
    const T scale = T(1) / (T(2) * ta);
    for (int i = 0; i < 4; ++i) {
        c.t_des[i] = vclamp(c.t_pre[i], P.thrust_min, P.thrust_max);
        c.disc[i] = vsqrt(tb * tb - T(4) * ta * (tc - c.t_des[i]));
        c.w_des[i] = scale * (c.disc[i] - tb);
        c.w_in[i] = (T(1) - P.motor_c) * c.w_des[i];
    }
}

// adjoint of command_fwd.  g_tdes: dL/dT_des (already includes the rotor-speed branch).
// Accumulates into gw / gal (start-of-step body rates and angular acceleration), writes ga.
template <class T>
VF_HD void command_adj(const Params<T>& P, int action_type, const T w[3], const Command<T>& c,
                       const T g_tdes[4], T ga[4], T gw[3], T gal[3]) {
    T g_pre[4];
    for (int i = 0; i < 4; ++i) g_pre[i] = vgate(c.t_pre[i], P.thrust_min, P.thrust_max, g_tdes[i]);
    if (action_type == VF_ACTION_BODYRATE) {
        T gft[4];
        for (int j = 0; j < 4; ++j)
            gft[j] = P.B_inv[j] * g_pre[0] + P.B_inv[4 + j] * g_pre[1] + P.B_inv[8 + j] * g_pre[2] +
                     P.B_inv[12 + j] * g_pre[3];
        const T* gt = gft + 1;  // dL/d tau_des
        T ge[3];
        for (int j = 0; j < 3; ++j) {
            ge[j] = P.JKp[j] * gt[0] + P.JKp[3 + j] * gt[1] + P.JKp[6 + j] * gt[2];
            gal[j] -= P.Kd[j] * gt[0] + P.Kd[3 + j] * gt[1] + P.Kd[6 + j] * gt[2];
            gw[j] -= ge[j];
        }
        gyro_adj(P, w, gt, gw);
        ga[0] = gft[0] * P.mass * P.act_half[0];
        for (int j = 0; j < 3; ++j) ga[1 + j] = ge[j] * P.act_half[1 + j];
    } else {
        for (int i = 0; i < 4; ++i) ga[i] = g_pre[i] * P.mass * P.act_half[i];
    }
}

// ---------------------------------------------------------------------------------------------
// one integration sub-step
// ---------------------------------------------------------------------------------------------
// Rotor model + allocation + drag + acceleration (everything that is frozen over the RK4 stages).
template <class T> struct Wrench {
    T thr[4];   // rotor thrusts actually produced this sub-step
    T tau[3];   // body torque
    T acc[3];   // world acceleration
    T vb[3];    // body-frame velocity (kept for the drag adjoint)
    T fb[3];    // body-frame force = (0,0,F) - drag
};

template <class T>
VF_HD void wrench_fwd(const Params<T>& P, bool ctrl_delay, const Command<T>& c, State<T>& s, Wrench<T>& k) {
    if (ctrl_delay) {
        for (int i = 0; i < 4; ++i) {
            s.mot[i] = P.motor_c * s.mot[i] + c.w_in[i];                                   // dynamics.py:514
            k.thr[i] = (P.thrust_map[0] * s.mot[i] + P.thrust_map[1]) * s.mot[i] + P.thrust_map[2];   // :530-534
        }
    } else {
        for (int i = 0; i < 4; ++i) k.thr[i] = c.t_des[i];                                 // dynamics.py:518
    }
    T ft[4];
    for (int i = 0; i < 4; ++i)                                                           // dynamics.py:339
        ft[i] = P.B[4 * i] * k.thr[0] + P.B[4 * i + 1] * k.thr[1] + P.B[4 * i + 2] * k.thr[2] +
                P.B[4 * i + 3] * k.thr[3];
    k.tau[0] = ft[1]; k.tau[1] = ft[2]; k.tau[2] = ft[3];
    RotMat<T> R;
    rotmat(s.q, R);
    rot_inv(R, s.v, k.vb);                                                                // dynamics.py:342
    for (int i = 0; i < 3; ++i)                                                           // dynamics.py:343-345
        k.fb[i] = -k.vb[i] * (P.k_lin[i] + P.k_quad[i] * vabs(k.vb[i]));
    k.fb[2] += ft[0];
    T aw[3];
    rot_fwd(R, k.fb, aw);                                                                 // dynamics.py:347
    for (int i = 0; i < 3; ++i) k.acc[i] = aw[i] * P.inv_mass + P.gravity[i];
}

// adjoint of wrench_fwd.
//   s      : the state at the START of the sub-step (rotor speeds before the lag update)
//   mot1   : rotor speeds after the lag update (what the thrust map saw)
//   gacc, gtau : incoming adjoints of acc and tau
// accumulates: gq, gv (start-of-sub-step q, v), gmot (in/out: on entry adjoint of the post-update rotor
// speed, on exit adjoint of the pre-update one), g_wdes / g_tdes (desired rotor speed / thrust).
template <class T>
VF_HD void wrench_adj(const Params<T>& P, bool ctrl_delay, const State<T>& s, const T mot1[4],
                      const Wrench<T>& k, const T gacc[3], const T gtau[3], T gq[4], T gv[3], T gmot[4],
                      T g_wdes[4], T g_tdes[4]) {
    T gaw[3];
    for (int i = 0; i < 3; ++i) gaw[i] = gacc[i] * P.inv_mass;
    T gfb[3];
    sandwich_adj(s.q, k.fb, T(1), gaw, gq, gfb);
    const T gF = gfb[2];
    T gvb[3];
    for (int i = 0; i < 3; ++i) gvb[i] = -gfb[i] * (P.k_lin[i] + T(2) * P.k_quad[i] * vabs(k.vb[i]));
    T gvv[3];
    sandwich_adj(s.q, s.v, T(-1), gvb, gq, gvv);
    for (int i = 0; i < 3; ++i) gv[i] += gvv[i];
    const T gft[4] = {gF, gtau[0], gtau[1], gtau[2]};
    for (int j = 0; j < 4; ++j) {
        const T gthr = P.B[j] * gft[0] + P.B[4 + j] * gft[1] + P.B[8 + j] * gft[2] + P.B[12 + j] * gft[3];
        if (ctrl_delay) {
            const T gm1 = gmot[j] + gthr * (T(2) * P.thrust_map[0] * mot1[j] + P.thrust_map[1]);
            g_wdes[j] += (T(1) - P.motor_c) * gm1;
            gmot[j] = P.motor_c * gm1;
        } else {
            g_tdes[j] += gthr;
        }
    }
}

// RK4 stage points of (q, w) with torque frozen (maths.py:363-379 with repair R1).
template <class T> struct Rk4Stages {
    T q2[4], q3[4], q4[4];
    T w2[3], w3[3], w4[3];
};

// Integrate q and w over one sub-step; returns the applied angular acceleration in `al`
// (Euler: maths.py:351; RK4: weighted stage mean, repair R3).  `st` (optional) receives the stage points.
template <class T>
VF_HD void attitude_fwd(const Params<T>& P, int integrator, const T tau[3], T q[4], T w[3], T al[3],
                        T* qn_norm, Rk4Stages<T>* st) {
    const T h = P.dt;
    T qn[4];
    WdotCoef<T> wc;
    wdot_coef(P, tau, wc);
    // g* below are 2 * qdot (q (x) (0,w)); the 1/2 lives in the quaternion step sizes qh = h/2, qhh = h/4
    const T qh = T(0.5) * h;
    if (integrator == VF_INTEGRATOR_RK4) {
        T k1[3], k2[3], k3[3], k4[3], g1[4], g2[4], g3[4], g4[4];
        T q2[4], q3[4], q4[4], w2[3], w3[3], w4[3];
        const T hh = T(0.5) * h, qhh = T(0.25) * h;
        wdot_lean(wc, w, k1);
        qdot2(q, w, g1);
        for (int i = 0; i < 3; ++i) w2[i] = w[i] + hh * k1[i];
        for (int i = 0; i < 4; ++i) q2[i] = q[i] + qhh * g1[i];
        wdot_lean(wc, w2, k2);
        qdot2(q2, w2, g2);
        for (int i = 0; i < 3; ++i) w3[i] = w[i] + hh * k2[i];
        for (int i = 0; i < 4; ++i) q3[i] = q[i] + qhh * g2[i];
        wdot_lean(wc, w3, k3);
        qdot2(q3, w3, g3);
        for (int i = 0; i < 3; ++i) w4[i] = w[i] + h * k3[i];
        for (int i = 0; i < 4; ++i) q4[i] = q[i] + qh * g3[i];
        wdot_lean(wc, w4, k4);
        qdot2(q4, w4, g4);
        const T s6 = T(1) / T(6), s3 = T(2) / T(6);
        for (int i = 0; i < 3; ++i) {
            al[i] = s6 * k1[i] + s3 * k2[i] + s3 * k3[i] + s6 * k4[i];      // maths.py:384 / R3
            w[i] = w[i] + al[i] * h;
        }
        for (int i = 0; i < 4; ++i)
            qn[i] = q[i] + (s6 * g1[i] + s3 * g2[i] + s3 * g3[i] + s6 * g4[i]) * qh;   // maths.py:382
        if (st) {
            for (int i = 0; i < 4; ++i) { st->q2[i] = q2[i]; st->q3[i] = q3[i]; st->q4[i] = q4[i]; }
            for (int i = 0; i < 3; ++i) { st->w2[i] = w2[i]; st->w3[i] = w3[i]; st->w4[i] = w4[i]; }
        }
    } else {
        T g[4];
        qdot2(q, w, g);
        wdot_lean(wc, w, al);
        for (int i = 0; i < 4; ++i) qn[i] = q[i] + g[i] * qh;               // maths.py:345
        for (int i = 0; i < 3; ++i) w[i] = w[i] + al[i] * h;                // maths.py:347
    }
    // dynamics.py:367, maths.py:226-230
    const T n2 = qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3];
    const T inv = vrsqrt(n2);
    for (int i = 0; i < 4; ++i) q[i] = qn[i] * inv;
    if (qn_norm) *qn_norm = n2 * inv;
}

// adjoint of attitude_fwd.
//   q0, w0 : start-of-sub-step values; tau frozen
//   gq, gw : in = adjoints of the outputs (normalised q', w'), out = adjoints of q0, w0
//   gal    : adjoint of the returned angular acceleration (non-zero only for the last sub-step)
//   gtau   : accumulated
template <class T>
VF_HD void attitude_adj(const Params<T>& P, int integrator, const T tau[3], const T q0[4], const T w0[3],
                        T gq[4], T gw[3], const T gal[3], T gtau[3]) {
    const T h = P.dt;
    // recompute the forward pass of this sub-step (stage points + normalisation)
    T q1[4] = {q0[0], q0[1], q0[2], q0[3]};
    T w1[3] = {w0[0], w0[1], w0[2]};
    T al[3], nrm;
    Rk4Stages<T> st;
    attitude_fwd(P, integrator, tau, q1, w1, al, &nrm, &st);   // q1 = normalised output
    // normalisation: q' = qn / |qn|
    const T dot = q1[0] * gq[0] + q1[1] * gq[1] + q1[2] * gq[2] + q1[3] * gq[3];
    const T inv = T(1) / nrm;
    T gqn[4];
    for (int i = 0; i < 4; ++i) gqn[i] = (gq[i] - q1[i] * dot) * inv;
    // adjoint of the mean angular acceleration: w' = w + h al ; al also returned
    T gk[3];
    for (int i = 0; i < 3; ++i) gk[i] = h * gw[i] + gal[i];
    // running adjoints of the base point
    T aq[4] = {gqn[0], gqn[1], gqn[2], gqn[3]};
    T aw[3] = {gw[0], gw[1], gw[2]};
    if (integrator == VF_INTEGRATOR_RK4) {
        const T s6 = T(1) / T(6), s3 = T(2) / T(6), hh = T(0.5) * h;
        T gg[4], gkk[3], sq[4], sw[3];
        // stage 4  (weight 1/6; reached from stage 3 with step h)
        for (int i = 0; i < 4; ++i) { gg[i] = s6 * h * gqn[i]; sq[i] = T(0); }
        for (int i = 0; i < 3; ++i) { gkk[i] = s6 * gk[i]; sw[i] = T(0); }
        qdot_adj(st.q4, st.w4, gg, sq, sw);
        wdot_adj(P, st.w4, gkk, sw, gtau);
        for (int i = 0; i < 4; ++i) aq[i] += sq[i];
        for (int i = 0; i < 3; ++i) aw[i] += sw[i];
        // stage 3  (weight 2/6 + h * adjoint of stage-4 point)
        for (int i = 0; i < 4; ++i) { gg[i] = s3 * h * gqn[i] + h * sq[i]; }
        for (int i = 0; i < 3; ++i) { gkk[i] = s3 * gk[i] + h * sw[i]; }
        for (int i = 0; i < 4; ++i) sq[i] = T(0);
        for (int i = 0; i < 3; ++i) sw[i] = T(0);
        qdot_adj(st.q3, st.w3, gg, sq, sw);
        wdot_adj(P, st.w3, gkk, sw, gtau);
        for (int i = 0; i < 4; ++i) aq[i] += sq[i];
        for (int i = 0; i < 3; ++i) aw[i] += sw[i];
        // stage 2  (weight 2/6 + h/2 * adjoint of stage-3 point)
        for (int i = 0; i < 4; ++i) { gg[i] = s3 * h * gqn[i] + hh * sq[i]; }
        for (int i = 0; i < 3; ++i) { gkk[i] = s3 * gk[i] + hh * sw[i]; }
        for (int i = 0; i < 4; ++i) sq[i] = T(0);
        for (int i = 0; i < 3; ++i) sw[i] = T(0);
        qdot_adj(st.q2, st.w2, gg, sq, sw);
        wdot_adj(P, st.w2, gkk, sw, gtau);
        for (int i = 0; i < 4; ++i) aq[i] += sq[i];
        for (int i = 0; i < 3; ++i) aw[i] += sw[i];
        // stage 1  (weight 1/6 + h/2 * adjoint of stage-2 point), evaluated at the base point itself
        for (int i = 0; i < 4; ++i) { gg[i] = s6 * h * gqn[i] + hh * sq[i]; }
        for (int i = 0; i < 3; ++i) { gkk[i] = s6 * gk[i] + hh * sw[i]; }
        qdot_adj(q0, w0, gg, aq, aw);
        wdot_adj(P, w0, gkk, aw, gtau);
    } else {
        T gg[4];
        for (int i = 0; i < 4; ++i) gg[i] = h * gqn[i];
        qdot_adj(q0, w0, gg, aq, aw);
        wdot_adj(P, w0, gk, aw, gtau);
    }
    for (int i = 0; i < 4; ++i) gq[i] = aq[i];
    for (int i = 0; i < 3; ++i) gw[i] = aw[i];
}

// One full sub-step on the state (rotor lag -> wrench -> translation -> attitude -> renormalise).
// `wind`: this agent's wind vector for the control step (time-varying wind functions, dynamics.py:136-165,384-388);
// NULL = the constant P.wind.
template <class T>
VF_HD void substep_fwd(const Params<T>& P, int integrator, bool ctrl_delay, const Command<T>& c, State<T>& s,
                       Wrench<T>& k, const T* wind = nullptr) {
    wrench_fwd(P, ctrl_delay, c, s, k);
    const T h = P.dt;
    const T* wd = wind ? wind : P.wind;
    if (integrator == VF_INTEGRATOR_RK4) {
        const T h2 = T(0.5) * h * h;
        for (int i = 0; i < 3; ++i) {
            s.p[i] = s.p[i] + (s.v[i] + wd[i]) * h + k.acc[i] * h2;          // maths.py:381 (stage velocities)
            s.v[i] = s.v[i] + k.acc[i] * h;                                   // maths.py:383
        }
    } else {
        for (int i = 0; i < 3; ++i) {
            s.p[i] = s.p[i] + (s.v[i] + wd[i]) * h;                           // maths.py:344
            s.v[i] = s.v[i] + k.acc[i] * h;                                   // maths.py:346
        }
    }
    attitude_fwd<T>(P, integrator, k.tau, s.q, s.w, s.al, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// the control step
// ---------------------------------------------------------------------------------------------
// Post-step clamps (dynamics.py:374-382).  `raw` keeps the unclamped values for the gradient gates.
template <class T> VF_HD void clamp_state(const Params<T>& P, State<T>& s) {
    for (int i = 0; i < 3; ++i) {
        s.p[i] = vclamp(s.p[i], P.pos_lo[i], P.pos_hi[i]);
        s.v[i] = vclamp(s.v[i], -P.vel_lim, P.vel_lim);
        s.w[i] = vclamp(s.w[i], -P.rate_lim, P.rate_lim);
    }
}

// Forward control step.  `a` is the (already comm-delayed) normalised action.
// On return s is the state after substeps*dt; k holds the last sub-step's wrench (acc, thrusts).
template <class T>
VF_HD void step_fwd(const Params<T>& P, int substeps, int integrator, int action_type, bool ctrl_delay,
                    const T a[4], State<T>& s, Wrench<T>& k, const T* wind = nullptr) {
    Command<T> c;
    command_fwd(P, action_type, a, s, c);
    for (int it = 0; it < substeps; ++it) substep_fwd(P, integrator, ctrl_delay, c, s, k, wind);
    clamp_state(P, s);
}

// Per-sub-step inputs the reverse sweep needs (14 scalars).
template <class T> struct Tape {
    T q[4], v[3], w[3], mot[4];
};

// Forward re-run for the reverse sweep: records each sub-step's inputs; `s` ends as the UNCLAMPED end state.
template <class T>
VF_HD void step_fwd_taped(const Params<T>& P, int substeps, int integrator, int action_type, bool ctrl_delay,
                          const T a[4], const State<T>& s0, Command<T>& c, State<T>& s, Tape<T>* tape,
                          const T* wind = nullptr) {
    command_fwd(P, action_type, a, s0, c);
    s = s0;
    Wrench<T> k;
    for (int it = 0; it < substeps; ++it) {
        Tape<T>& t = tape[it];
        for (int i = 0; i < 4; ++i) { t.q[i] = s.q[i]; t.mot[i] = s.mot[i]; }
        for (int i = 0; i < 3; ++i) { t.v[i] = s.v[i]; t.w[i] = s.w[i]; }
        substep_fwd(P, integrator, ctrl_delay, c, s, k, wind);
    }
}

// Reverse sweep over a recorded step.
//   s_raw: unclamped end state from step_fwd_taped (gradient gates of the post-step clamps)
//   g    : in = dL/d(clamped state after the step), out = dL/d(s0);   ga: out = dL/d(action)
template <class T>
VF_HD void step_bwd_taped(const Params<T>& P, int substeps, int integrator, int action_type, bool ctrl_delay,
                          const State<T>& s0, const Command<T>& c, const State<T>& s_raw, const Tape<T>* tape,
                          State<T>& g, T ga[4]) {
    // ---- gates of the post-step clamps (gradient w.r.t. the unclamped values) ----
    for (int i = 0; i < 3; ++i) {
        g.p[i] = vgate(s_raw.p[i], P.pos_lo[i], P.pos_hi[i], g.p[i]);
        g.v[i] = vgate(s_raw.v[i], -P.vel_lim, P.vel_lim, g.v[i]);
        g.w[i] = vgate(s_raw.w[i], -P.rate_lim, P.rate_lim, g.w[i]);
    }
    // ---- reverse sweep ----
    T g_wdes[4] = {T(0), T(0), T(0), T(0)};
    T g_tdes[4] = {T(0), T(0), T(0), T(0)};
    T gal[3] = {g.al[0], g.al[1], g.al[2]};     // adjoint of the returned angular acceleration
    const T h = P.dt;
    for (int it = substeps - 1; it >= 0; --it) {
        const Tape<T>& t = tape[it];
        State<T> si;                          // start-of-sub-step state (p is not needed by any adjoint)
        for (int i = 0; i < 4; ++i) { si.q[i] = t.q[i]; si.mot[i] = t.mot[i]; }
        for (int i = 0; i < 3; ++i) { si.v[i] = t.v[i]; si.w[i] = t.w[i]; si.p[i] = T(0); si.al[i] = T(0); }
        // recompute this sub-step's wrench (it also advances the rotor speeds in a copy)
        State<T> sc = si;
        Wrench<T> kk;
        wrench_fwd(P, ctrl_delay, c, sc, kk);
        // attitude part: (q,w) <- (q', w', alpha)
        T gtau[3] = {T(0), T(0), T(0)};
        attitude_adj<T>(P, integrator, kk.tau, si.q, si.w, g.q, g.w, gal, gtau);
        gal[0] = gal[1] = gal[2] = T(0);      // earlier sub-steps' alpha is overwritten, carries no gradient
        // translation part
        T gacc[3];
        if (integrator == VF_INTEGRATOR_RK4) {
            const T h2 = T(0.5) * h * h;
            for (int i = 0; i < 3; ++i) {
                gacc[i] = h * g.v[i] + h2 * g.p[i];
                g.v[i] = g.v[i] + h * g.p[i];
            }
        } else {
            for (int i = 0; i < 3; ++i) {
                gacc[i] = h * g.v[i];
                g.v[i] = g.v[i] + h * g.p[i];
            }
        }
        // wrench part
        wrench_adj(P, ctrl_delay, si, sc.mot, kk, gacc, gtau, g.q, g.v, g.mot, g_wdes, g_tdes);
    }
    // ---- once-per-step part ----
    if (ctrl_delay)
        for (int i = 0; i < 4; ++i) g_tdes[i] += g_wdes[i] / c.disc[i];
    g.al[0] = g.al[1] = g.al[2] = T(0);         // start-of-step alpha only enters through the PID D term
    command_adj(P, action_type, s0.w, c, g_tdes, ga, g.w, g.al);
}

// Reverse-mode control step.
//   s0  : state at the start of the step
//   g   : in  = dL/d(state after the step) [gradient w.r.t. the observation already folded in by the caller],
//         out = dL/d(s0)
//   ga  : out = dL/d(action)
//   tape: scratch of at least `substeps` entries
template <class T>
VF_HD void step_bwd(const Params<T>& P, int substeps, int integrator, int action_type, bool ctrl_delay,
                    const T a[4], const State<T>& s0, State<T>& g, T ga[4], Tape<T>* tape,
                    const T* wind = nullptr) {
    Command<T> c;
    State<T> s;
    step_fwd_taped(P, substeps, integrator, action_type, ctrl_delay, a, s0, c, s, tape, wind);
    step_bwd_taped(P, substeps, integrator, action_type, ctrl_delay, s0, c, s, tape, g, ga);
}

}  // namespace vf
