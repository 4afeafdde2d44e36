"""TEST INFRASTRUCTURE — CPU/PyTorch restatement of the reference's quadrotor control step.

This module is the *oracle* for the CUDA engine in ``visfly_b200``: it restates, operation by operation on
``(3,N)`` / ``(4,N)`` component-major tensors, what the reference executes for one call of
``Dynamics.step`` (VisFly envs/base/dynamics.py:319-372) and ``Integrator.integrate``
(utils/maths.py:317-389), so that

  * ``torch.autograd`` through it is the gradient oracle for the hand-derived adjoint kernel, and
  * timing it is the "reference arm" / ``cpu_baseline`` of ``bench.py`` (kind = "port": it issues the same
    sequence of tiny aten ops the reference issues, ~2k per Euler step, ~8k per RK4 step).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import it.  The product
package never does (it has no CPU fallback).

Parity status: **pinned against the executed reference** — ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference`` (with the RK4 repairs R1-R3 of SURVEY.md §8c applied as monkeypatches)
and records state trajectories and autograd gradients into ``tests/golden/*.npz``; ``tests/test_oracle.py``
checks this restatement against those fixtures (bit-exact on the state, see the test for the tolerance
on gradients) and, when ``/root/reference`` is present, against the live reference.  The reference itself
ships no tests or golden vectors.

RK4 note: the reference's ``integrator="rk4"`` crashes as shipped; the oracle implements the repaired
semantics frozen in SURVEY.md §8c: R1 wind is passed to every stage, R2 stage buffers live on the state's
device/dtype, R3 the returned angular acceleration is the weighted stage mean ``d_ori_vel @ ks``.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import torch as th

_DEFAULT_CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "visfly_b200", "configs", "drone")


# ---------------------------------------------------------------------------------------------------
# quaternion algebra on 4-tuples of (N,) tensors                          reference utils/maths.py:4-293
# ---------------------------------------------------------------------------------------------------
def q_mul(a, b):
    """Hamilton product, term order as reference maths.py:170-173."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return (aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw)


def q_conj(q):
    return (q[0], -q[1], -q[2], -q[3])


def q_pure(v, like):
    """(0, v) with a 0-dim zero real part (reference builds it with ``th.tensor(0)``, maths.py:38)."""
    return (th.tensor(0, device=like.device), v[0], v[1], v[2])


def q_rotate(q, v):
    """Im(q (x) (0,v) (x) q*) — body to world (maths.py:38)."""
    r = q_mul(q_mul(q, q_pure(v, q[0])), q_conj(q))
    return th.stack([r[1], r[2], r[3]])


def q_inv_rotate(q, v):
    """Im(q* (x) (0,v) (x) q) — world to body (maths.py:49)."""
    r = q_mul(q_mul(q_conj(q), q_pure(v, q[0])), q)
    return th.stack([r[1], r[2], r[3]])


def q_normalize(q):
    n = th.sqrt(q[0].pow(2) + q[1].pow(2) + q[2].pow(2) + q[3].pow(2))     # maths.py:227
    return (q[0] / n, q[1] / n, q[2] / n, q[3] / n)


def cross3(a, b):
    """maths.py:392-394"""
    return th.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]) + 0


# ---------------------------------------------------------------------------------------------------
# constants                                                        reference envs/base/dynamics.py:562-689
# ---------------------------------------------------------------------------------------------------
class OracleModel:
    def __init__(self, cfg: str, dt: float, dtype, device):
        path = cfg if os.path.isfile(cfg) else os.path.join(_DEFAULT_CFG_DIR, cfg + ".json")
        with open(path) as f:
            d = json.load(f)
        d.setdefault("max_acc", 3.0)
        t = lambda x: th.tensor(x, dtype=dtype)
        self.g = t([[0, 0, -9.81]]).T
        self.z = t([[0, 0, 1]]).T
        self.m = t(d["mass"])
        self.k_quad = t([d["quad_drag_coeffs"]]).T * 0.5 * 1.225 * t([d["cross_sections"]]).T
        self.k_lin = t([d["linear_drag_coeffs"]]).T
        self.J = th.diag(t(d["inertia"]))
        self.J_inv = th.inverse(self.J)
        self.Kp = t(d["BODYRAYE_PID"]["p"])
        self.Kd = t(d["BODYRAYE_PID"]["d"])
        self.thrust_map = t(d["thrust_map"])
        self.c = th.exp(-t(1 / d["motor_tau"]) * dt)
        w_max = d["motor_omega_max"]
        self.thrust_max = self.thrust_map[0] * w_max ** 2 + self.thrust_map[1] * w_max + self.thrust_map[2]
        self.thrust_min = 0
        arm = t(d["arm_length"])
        dirs = t([[1, -1, -1, 1.0], [-1, -1, 1, 1], [0, 0, 0, 0.0]])
        dirs = dirs / dirs.norm(dim=0)
        self.B = th.vstack([th.ones(1, 4, dtype=dtype), (arm * dirs)[:2], t(d["kappa"]) * t([1, -1, 1, -1])])
        self.B_inv = th.inverse(self.B)
        # "max_min" normalisation, dynamics.py:625-642
        acc_max, acc_min = (d["max_acc"] * -self.g[2]).clone(), t(0)                # dynamics.py:597-599
        rate_max, rate_min = t(d["max_rate"]), t(-d["max_rate"])
        self.acc_half = th.atleast_1d((acc_max - acc_min) / 2)
        self.acc_mean = th.atleast_1d(acc_max - self.acc_half * 1)
        self.rate_half = th.atleast_1d((rate_max - rate_min) / 2)
        self.rate_mean = th.atleast_1d(rate_max - self.rate_half * 1)
        # velocity / position action types: outer-loop gains and set-point scaling        dynamics.py:574-575, :655-687
        self.vel_p, self.vel_d = t(d["VELOCITY_PID"]["p"]), t(d["VELOCITY_PID"]["d"])
        self.pos_d = t(d["POSITION_PID"]["d"])
        spd_max, pos_max = t(d["max_spd"]), t(d["max_pos"])
        self.spd_half = th.atleast_1d((spd_max - (-spd_max)) / 2)
        self.spd_mean = th.atleast_1d(spd_max - self.spd_half * 1)
        self.pos_half = th.atleast_1d((pos_max - (-pos_max)) / 2)
        self.pos_mean = th.atleast_1d(pos_max - self.pos_half * 1)
        yaw_scale = th.as_tensor(th.pi - (-th.pi), dtype=dtype) / 2
        self.yaw_mean = th.atleast_1d(th.pi - yaw_scale * 1)
        self.yaw_half_velocity = self.yaw_mean.clone()       # the reference passes half=yaw_bias here (:671)
        self.yaw_half_position = th.atleast_1d(yaw_scale)
        for k, v in list(vars(self).items()):
            if isinstance(v, th.Tensor):
                setattr(self, k, v.to(device))

    def rotor_omega(self, thrust):                                                # dynamics.py:545-554
        a, b, c = self.thrust_map[0], self.thrust_map[1], self.thrust_map[2]
        return (1 / (2 * a)) * (-b + th.sqrt(b.pow(2) - 4 * a * (c - thrust)))

    def rotor_thrust(self, omega):                                                # dynamics.py:530-534
        a, b, c = self.thrust_map[0], self.thrust_map[1], self.thrust_map[2]
        return (a * (omega + 0).pow(2)) + b * omega + c


# ---------------------------------------------------------------------------------------------------
# the control step
# ---------------------------------------------------------------------------------------------------
class OracleDynamics:
    """Restatement of reference ``Dynamics`` for the four action types (``velocity`` / ``position``: forward
    only — the reference's own backward raises there, and its per-agent Python loop is kept as a loop).

    State is held like the reference holds it: ``pos/vel/ang_vel/ang_acc (3,N)``, ``motor/thrusts (4,N)``,
    quaternion as four ``(N,)`` tensors, ``t (N,)``, FIFO of ``(4,N)`` delayed actions.
    """

    def __init__(self, num: int = 1, action_type: str = "bodyrate", dt: float = 0.005, ctrl_dt: float = 0.03,
                 ctrl_delay: bool = True, comm_delay: float = 0.06, integrator: str = "euler",
                 cfg: str = "drone_state", wind: Sequence = (0, 0, 0), device="cpu",
                 dtype=th.float32, random_reset_time: bool = False, drag_random: float = 0):
        assert action_type in ("bodyrate", "thrust", "velocity", "position")
        assert integrator in ("euler", "rk4")
        self.num, self.action_type, self.integrator = num, action_type, integrator
        self.random_reset_time = random_reset_time     # reference quirk C9: partial reset draws t ~ U(0, 6.28)
        self.dt, self.ctrl_dt, self.ctrl_delay = dt, ctrl_dt, ctrl_delay
        self.device, self.dtype = th.device(device), dtype
        if not th.as_tensor(ctrl_dt) % th.as_tensor(dt) == 0:
            raise ValueError("ctrl_dt should be a multiple of dt")
        self.substeps = int(ctrl_dt / dt)                                          # dynamics.py:74
        self.fifo_depth = int(comm_delay / ctrl_dt)                                # dynamics.py:75
        self.M = OracleModel(cfg, dt, dtype, self.device)
        self.drag_random = drag_random
        self.k_lin_mean, self.k_quad_mean = self.M.k_lin, self.M.k_quad            # dynamics.py:129-130
        self.init_thrust = -(self.M.m * self.M.g / 4)[-1]
        self.init_omega = self.M.rotor_omega(self.init_thrust)
        self._constructing = True      # the reference does not reset (hence does not draw drag) in its constructor
        self.wind = th.zeros((3, 1), dtype=dtype, device=self.device)
        self.reset()
        self._constructing = False
        self._create_wind(wind)

    def _create_wind(self, wind):
        """dynamics.py:132-174.  Numbers: constant wind.  Six strings: ``wind = f1(t, wind_1) + f2(t, wind_2)``,
        each component an expression in x (= t, (N,)) and y (= its previous value), re-evaluated at the start of
        every control step (``update_wind``, :384-388).  (The reference's 3-string form builds 3-argument lambdas that
        ``update_wind`` calls with two arguments: it raises at construction and is not restated.)"""
        self.wind_fn = None
        if isinstance(wind[0], str):
            if len(wind) != 6:
                raise ValueError("wind functions: a list of six expression strings")
            fx = [eval("lambda x,y:" + w) for w in wind]
            self.wind_fn = (lambda x, y: th.stack([fx[0](x, y[0]), fx[1](x, y[1]), fx[2](x, y[2])]),
                            lambda x, y: th.stack([fx[3](x, y[0]), fx[4](x, y[1]), fx[5](x, y[2])]))
            self.wind_1, self.wind_2 = self._zeros(3), self._zeros(3)              # dynamics.py:172-173
            self.update_wind()                                                     # dynamics.py:174
        else:
            self.wind = th.tensor(list(wind), dtype=self.dtype, device=self.device).reshape(3, 1)

    def update_wind(self):                                                         # dynamics.py:384-388
        if self.wind_fn is None:
            return
        self.wind_1 = self.wind_fn[0](self.t, self.wind_1)
        self.wind_2 = self.wind_fn[1](self.t, self.wind_2)
        self.wind = self.wind_1 + self.wind_2

    # -- state ---------------------------------------------------------------------------------
    def _zeros(self, k, n=None):
        return th.zeros((k, self.num if n is None else n), dtype=self.dtype, device=self.device)

    def reset(self, pos=None, ori=None, vel=None, ori_vel=None, motor_omega=None, thrusts=None, t=None,
              ang_acc=None, indices=None):
        """Inputs are (n,k) row-major like the reference (dynamics.py:229-236); they are copied (C5)."""
        cv = lambda x: None if x is None else th.as_tensor(x, dtype=self.dtype, device=self.device).clone()
        pos, ori, vel, ori_vel, motor_omega, thrusts = map(cv, (pos, ori, vel, ori_vel, motor_omega, thrusts))
        if indices is None:
            n = self.num
            self.pos = self._zeros(3) if pos is None else pos.T
            q = th.tensor([[1.0, 0, 0, 0]], dtype=self.dtype, device=self.device).repeat(n, 1) if ori is None else ori
            self.q = tuple(q.T)
            self.vel = self._zeros(3) if vel is None else vel.T
            self.ang_vel = self._zeros(3) if ori_vel is None else ori_vel.T
            self.thrusts = th.ones((4, n), dtype=self.dtype, device=self.device) * self.init_thrust \
                if thrusts is None else thrusts.T
            self.motor = th.ones((4, n), dtype=self.dtype, device=self.device) * self.init_omega \
                if motor_omega is None else motor_omega.T
            self.t = th.zeros((n,), dtype=self.dtype, device=self.device) if t is None else cv(t)
            self.ang_acc = self._zeros(3) if ang_acc is None else cv(ang_acc).T
            self.acc = self._zeros(3)
            self.fifo: List[th.Tensor] = [self._zeros(4) for _ in range(self.fifo_depth)]
            if self.drag_random and not self._constructing:                        # dynamics.py:244-246
                dr = self.drag_random
                self.M.k_lin = self.k_lin_mean * (((th.rand_like(self.k_lin_mean) - 0.5) * 2 * dr).clamp(-0.5, .5) + 1)
                self.M.k_quad = self.k_quad_mean * (((th.rand_like(self.k_quad_mean) - 0.5) * 2 * dr).clamp(-0.5, .5) + 1)
        else:
            idx = th.as_tensor(indices, device=self.device)
            m = len(idx)
            # functional masked writes (same values as the reference's in-place index_put, dynamics.py:249-263;
            # upstream gradient of the overwritten agents is cut exactly like there)
            def put(dst, src, k):
                dst = dst.clone()
                dst[:, idx] = (th.zeros((k, m), dtype=self.dtype, device=self.device) if src is None else src.T)
                return dst
            self.pos = put(self.pos, pos, 3)
            qt = th.stack(self.q)
            qn = th.tensor([[1.0, 0, 0, 0]], dtype=self.dtype, device=self.device).repeat(m, 1) if ori is None else ori
            qt = qt.clone()
            qt[:, idx] = qn.T
            self.q = tuple(qt)
            self.vel = put(self.vel, vel, 3)
            self.ang_vel = put(self.ang_vel, ori_vel, 3)
            mo = self.motor.clone()
            mo[:, idx] = (th.ones((4, m), dtype=self.dtype, device=self.device) * self.init_omega
                          if motor_omega is None else motor_omega.T)
            self.motor = mo
            tr = self.thrusts.clone()
            tr[:, idx] = (th.ones((4, m), dtype=self.dtype, device=self.device) * self.init_thrust
                          if thrusts is None else thrusts.T)
            self.thrusts = tr
            tt = self.t.clone()
            if t is None and self.random_reset_time:                              # dynamics.py:256
                tt[idx] = th.zeros((m,), dtype=self.dtype, device=self.device) + th.rand((m,)) * 3.14 * 2
            else:
                tt[idx] = th.zeros((m,), dtype=self.dtype, device=self.device) if t is None else cv(t)
            self.t = tt
            self.ang_acc = put(self.ang_acc, None, 3)
            self.acc = put(self.acc, None, 3)
            fifo = []
            for a in self.fifo:
                a = a.clone()
                a[:, idx] = a[:, idx] * 0
                fifo.append(a)
            self.fifo = fifo
        return self.state

    def detach(self):                                                              # dynamics.py:176-190
        for k in ("pos", "vel", "ang_vel", "motor", "thrusts", "ang_acc", "acc", "t"):
            setattr(self, k, getattr(self, k).clone().detach())
        self.q = tuple(c.clone().detach() for c in self.q)
        self.fifo = [a.clone().detach() for a in self.fifo]

    # -- pieces of the step ----------------------------------------------------------------------
    def _denormalize(self, action):                                               # dynamics.py:704-713
        M = self.M
        if self.action_type == "bodyrate":
            cmd = th.hstack([(action[:, :1] * M.acc_half + M.acc_mean) * M.m,
                             action[:, 1:] * M.rate_half + M.rate_mean])
            return cmd.T
        if self.action_type == "thrust":
            return M.m * (action * M.acc_half + M.acc_mean).T
        if self.action_type == "velocity":                                        # dynamics.py:714-720 (N,4)
            return th.hstack([action[:, :1] * M.yaw_half_velocity + M.yaw_mean,
                              action[:, 1:] * M.spd_half + M.spd_mean])
        return th.hstack([action[:, :1] * M.yaw_half_position + M.yaw_mean,       # dynamics.py:722-728
                          action[:, 1:] * M.pos_half + M.pos_mean])

    def _yaw(self):                                                               # maths.py:248
        w, x, y, z = self.q
        return th.atan2(2 * (w * z + x * y), 1 - 2 * (y.pow(2) + z.pow(2)))

    def _R(self):                                                                 # maths.py:113-117
        w, x, y, z = self.q
        return th.stack([
            th.stack([1 - 2 * (y.pow(2) + z.pow(2)), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
            th.stack([2 * (x * y + z * w), 1 - 2 * (x.pow(2) + z.pow(2)), 2 * (y * z - x * w)]),
            th.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x.pow(2) + y.pow(2))])])

    def _geometric(self, F_des, yaw_des, yaw_spd_des):
        """Desired frame, attitude / rate errors (dynamics.py:435-450, :471-485), incl. the per-agent loop."""
        gross = q_inv_rotate(self.q, F_des)[2]
        R = self._R()
        b3 = F_des / F_des.norm(dim=0)
        c1 = th.stack([yaw_des.cos(), yaw_des.sin(), th.zeros_like(yaw_des)], dim=0)
        b2 = cross3(b3, c1)
        b2 = b2 / b2.norm(dim=0)
        b1 = cross3(b2, b3)
        R_des = th.stack([b1, b2, b3]).transpose(0, 1)
        pose_err, rate_err = th.zeros_like(self.pos), th.zeros_like(self.pos)
        for i in range(self.num):
            m = 0.5 * (R_des[..., i].T @ R[..., i] - R[..., i].T @ R_des[..., i])
            pose_err[:, i] = -th.as_tensor([-m[1, 2], m[0, 2], -m[0, 1]], device=self.device)
            rate_err[:, i] = (R_des[..., i].T @ R[..., i]
                              @ th.tensor([[0], [0], [yaw_spd_des[i]]], device=self.device, dtype=self.dtype).squeeze()
                              - self.ang_vel[:, i])
        return gross, pose_err, rate_err

    def _thrust_des(self, cmd):                                                   # dynamics.py:398-501
        M = self.M
        if self.action_type == "bodyrate":
            err = cmd[1:] - self.ang_vel
            tau_des = M.J @ M.Kp @ err + cross3(self.ang_vel + 0, M.J @ (self.ang_vel + 0)) - M.Kd @ self.ang_acc
            t_des = M.B_inv @ th.cat([cmd[0:1, :], tau_des])
        elif self.action_type == "thrust":
            t_des = cmd
        elif self.action_type == "velocity":                                      # dynamics.py:414-454
            cmd = cmd.T
            a_des = M.vel_p * (cmd[1:] - self.vel)
            F_des = M.m * (a_des - M.g)
            vh = self.vel[:2, :]
            yaw = self._yaw()
            yaw_des = th.where(vh.norm(dim=0) > 0.1, th.atan2(vh[1], vh[0]), yaw)
            err = yaw_des - yaw
            err = th.atan2(th.sin(err), th.cos(err))
            gross, pose_err, rate_err = self._geometric(F_des, yaw_des, err * M.vel_d * 2.0)
            tau_des = M.J @ (M.Kp @ pose_err + M.Kp @ rate_err - cross3(self.ang_vel, self.ang_vel))
            t_des = M.B_inv @ th.vstack([gross, tau_des])
        else:                                                                     # dynamics.py:455-496
            cmd = cmd.T
            v_des = M.pos_d * (cmd[1:] - self.pos)
            a_des = M.vel_d * (v_des - self.vel)
            F_des = M.m * (a_des - M.g)
            yaw_des = cmd[0]
            err = yaw_des - self._yaw()
            err = th.atan2(th.sin(err), th.cos(err))
            gross, pose_err, rate_err = self._geometric(F_des, yaw_des, err * M.pos_d * 2.0)
            tau_des = M.J @ (M.Kp @ pose_err + 1.2 * M.Kp @ rate_err - M.Kd @ self.ang_acc
                             - cross3(self.ang_vel, M.J @ self.ang_vel))
            t_des = M.B_inv @ th.vstack([gross, tau_des])
        return th.clamp(t_des, M.thrust_min, M.thrust_max)

    def _derivs(self, vel, q, acc, w, tau):                                       # maths.py:300-315
        M = self.M
        d_pos = vel + self.wind
        qd = q_mul(q, q_pure(w, q[0]))
        d_q = th.stack([qd[0] * 0.5, qd[1] * 0.5, qd[2] * 0.5, qd[3] * 0.5])
        d_w = M.J_inv @ (tau - th.linalg.cross(w.T, (M.J @ w).T).T)
        return d_pos, d_q, acc, d_w

    def _integrate(self, acc, tau):                                               # maths.py:317-389
        dt = self.dt
        pos, vel, w = self.pos, self.vel, self.ang_vel
        q = th.stack(self.q)
        if self.integrator == "euler":
            d_pos, d_q, d_vel, d_w = self._derivs(vel, self.q, acc, w, tau)
            pos = pos + d_pos * dt
            q = q + d_q * dt
            vel = vel + d_vel * dt
            w = w + d_w * dt
            alpha = d_w
        else:
            ks = th.tensor([1., 2., 2., 1.], dtype=self.dtype, device=self.device) / 6
            frac = (0.5, 0.5, 1.0)
            dps, dqs, dvs, dws = [], [], [], []
            qc, vc, wc = self.q, vel, w
            for i in range(4):
                if i:
                    qc = tuple(q + dqs[i - 1] * frac[i - 1] * dt)
                    vc = vel + dvs[i - 1] * frac[i - 1] * dt
                    wc = w + dws[i - 1] * frac[i - 1] * dt
                d_pos, d_q, d_vel, d_w = self._derivs(vc, qc, acc, wc, tau)
                dps.append(d_pos); dqs.append(d_q); dvs.append(d_vel); dws.append(d_w)
            # the reference stacks the four stage derivatives on a trailing axis and contracts with ks
            # (maths.py:358-361, :381-384)
            stk = lambda xs: th.stack([x.expand_as(xs[0]) if x.shape != xs[0].shape else x for x in xs], dim=-1)
            pos = pos + stk([d.expand_as(pos) for d in dps]) @ ks * dt
            q = q + stk(dqs) @ ks * dt
            vel = vel + stk(dvs) @ ks * dt
            dw = stk(dws) @ ks
            w = w + stk(dws) @ ks * dt
            alpha = dw                                                            # repair R3
        self.pos, self.vel, self.ang_vel, self.ang_acc = pos, vel, w, alpha
        self.q = q_normalize(tuple(q))                                            # dynamics.py:367

    def step(self, action):
        """``action`` (N,4) in [-1,1]  ->  ``state`` (N,13)                        dynamics.py:319-372"""
        M = self.M
        self.update_wind()                                                        # dynamics.py:320
        action = th.as_tensor(action, dtype=self.dtype, device=self.device)
        if self.fifo_depth:                                                       # dynamics.py:323-326
            self.fifo.append(action.T.clone())
            action = self.fifo.pop(0).T
        cmd = self._denormalize(action)
        t_des = self._thrust_des(cmd)
        for _ in range(self.substeps):
            if self.ctrl_delay:                                                   # dynamics.py:510-516
                w_des = M.rotor_omega(t_des)
                self.motor = M.c * self.motor + (1 - M.c) * w_des
                self.thrusts = M.rotor_thrust(self.motor)
            else:
                self.thrusts = t_des
            ft = M.B @ self.thrusts                                               # dynamics.py:339
            v_body = q_inv_rotate(self.q, self.vel + 0)                           # dynamics.py:342
            drag = M.k_lin * v_body + M.k_quad * v_body * v_body.abs()            # dynamics.py:343-345
            self.acc = q_rotate(self.q, M.z * ft[0] - drag) / M.m + M.g           # dynamics.py:347
            self._integrate(self.acc, ft[1:])
        self.t = self.t + self.ctrl_dt                                            # dynamics.py:368
        self.pos = th.vstack([self.pos[0:2].clamp(-100, 100), self.pos[2].clamp(0, 20)])   # :378-380
        self.vel = self.vel.clamp(-20, 20)
        self.ang_vel = self.ang_vel.clamp(-10, 10)
        return self.state

    # -- views (dynamics.py:735-819) --------------------------------------------------------------
    @property
    def position(self): return self.pos.T
    @property
    def orientation(self): return th.stack(self.q).T
    @property
    def velocity(self): return (self.vel + self.wind).T
    @property
    def angular_velocity(self): return self.ang_vel.T
    @property
    def angular_acceleration(self): return self.ang_acc.T
    @property
    def acceleration(self): return self.acc.T
    @property
    def motor_omega(self): return self.motor.T
    @property
    def direction(self):                                                          # maths.py:122-133
        w, x, y, z = self.q
        return th.stack([1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w)]).T
    @property
    def state(self):
        return th.hstack([self.position, self.orientation, self.velocity, self.angular_velocity])
    @property
    def full_state(self):
        return th.hstack([self.position, self.orientation, self.velocity, self.angular_velocity,
                          self.motor.T, self.thrusts.T, self.t.unsqueeze(1)])

    # -- packed-state bridge used by the parity tests (layout of include/visfly_b200.h) --------------
    def packed(self) -> th.Tensor:
        n = self.num
        out = th.zeros((5, n, 4), dtype=self.dtype, device=self.device)
        out[0, :, :3], out[0, :, 3] = self.pos.T, self.ang_acc[0]
        out[1] = th.stack(self.q).T
        out[2, :, :3], out[2, :, 3] = self.vel.T, self.ang_acc[1]
        out[3, :, :3], out[3, :, 3] = self.ang_vel.T, self.ang_acc[2]
        out[4] = self.motor.T
        return out

    def load_packed(self, packed: th.Tensor):
        p = packed.to(self.dtype).to(self.device)
        self.pos, self.vel, self.ang_vel = p[0, :, :3].T, p[2, :, :3].T, p[3, :, :3].T
        self.q = tuple(p[1].T)
        self.motor = p[4].T
        self.ang_acc = th.stack([p[0, :, 3], p[2, :, 3], p[3, :, 3]])
        self.thrusts = self.M.rotor_thrust(self.motor)
