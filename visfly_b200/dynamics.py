"""Drop-in ``Dynamics`` (reference envs/base/dynamics.py:19-826) on top of the fused sm_100a control-step kernel.

Same constructor keywords, methods and properties as the reference class, so ``DroneEnvsBase`` and task code
keep working; what is different is where the work happens:

* the whole control step (command de-normalisation, body-rate PID, thrust clamp, ``ctrl_dt/dt`` sub-steps of
  rotor lag / allocation / drag / Euler-or-RK4 / renormalisation, post-step clamps, the ``state`` view) is ONE
  kernel launch (``vf_step_fwd``) instead of ~2 000 (Euler x4) to ~8 000 (RK4 x8) aten ops;
* its gradient is ONE launch of the hand-derived adjoint kernel (``vf_step_bwd``) behind a
  ``torch.autograd.Function``; nothing but the step's inputs is kept for the backward pass (the reference's
  autograd graph keeps ~7.9 KB per agent per control step);
* the agent state lives in HBM as five planes of float4 (see include/visfly_b200.h); the (N,k) tensors the
  reference's properties return are views of the kernel's outputs;
* there is no CPU implementation: without the CUDA extension every call raises.

Deliberate deviations from the reference (documented in DESIGN.md): inputs of ``reset`` are copied (the
reference aliases and later mutates them); the two per-step device synchronising asserts
(dynamics.py:333, droneGymEnv.py:144) are opt-in via ``debug_checks``; ``t`` after a partial reset is 0 unless
``random_reset_time=True`` (reference draws U(0, 2*pi), C9).  The reference's wind functions (six expression
strings, dynamics.py:136-151) are evaluated on the device once per control step and reach the kernel as a per-agent
wind vector; ``drag_random`` re-draws the (shared) drag coefficients on every full reset like dynamics.py:244-246.
Action types ``velocity`` / ``position`` (geometric attitude
controller, reference dynamics.py:414-496, per-agent Python loop there) run forward in the same kernel; their
backward raises, as the reference's own autograd does on that branch.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch as th

from . import _lib
from .maths import Quaternion
from .params import action_scaling, build_vf_params, load_drone_model
from .type import ACTION_TYPE

_ALIAS = {"thrust": ACTION_TYPE.THRUST, "bodyrate": ACTION_TYPE.BODYRATE,
          "velocity": ACTION_TYPE.VELOCITY, "position": ACTION_TYPE.POSITION}


class StepConfig:
    """Everything ``vf_step_fwd`` / ``vf_step_bwd`` need besides tensors (immutable per Dynamics object)."""

    __slots__ = ("params", "substeps", "integrator", "action_type", "flags", "params_addr")

    def __init__(self, params, substeps, integrator, action_type, flags):
        import ctypes
        self.params, self.substeps, self.integrator = params, substeps, integrator
        self.action_type, self.flags = action_type, flags
        #: address of the ``VfParams`` struct (owned by this object) for the C++ plumbing
        self.params_addr = ctypes.addressof(params)

    def __deepcopy__(self, memo):
        import copy
        return StepConfig(copy.deepcopy(self.params, memo), self.substeps, self.integrator, self.action_type,
                          self.flags)


class ControlStep(th.autograd.Function):
    """``(state[5,N,4], action[N,4], push[N,4] | None) -> (state'[5,N,4], obs[N,13], copy[N,4] | None)`` — one kernel
    each way.  ``action`` is the (comm-delayed) action the step consumes; ``push`` is the action that arrived this
    step: the kernel leaves an engine-owned copy of it next to its other outputs — the reference's
    ``action.T.clone()`` (dynamics.py:324) — which is what waits in the FIFO.  Its gradient passes straight through to
    ``push``, exactly like the clone's."""

    @staticmethod
    def forward(ctx, state: th.Tensor, action: th.Tensor, push: Optional[th.Tensor], cfg: StepConfig,
                wind: Optional[th.Tensor] = None):
        state_out, obs, copy = _lib.fast().step_fwd(cfg.params_addr, cfg.substeps, cfg.integrator, cfg.action_type,
                                                    cfg.flags, state, action, wind, push)
        ctx.cfg, ctx.wind = cfg, wind
        ctx.save_for_backward(state, action)
        ctx.set_materialize_grads(False)
        return state_out, obs, copy

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, g_state_out, g_obs, g_copy):
        state, action = ctx.saved_tensors
        cfg = ctx.cfg
        g_state = th.empty_like(state)
        g_action = th.empty_like(action)
        if g_state_out is not None:
            g_state_out = g_state_out.contiguous()
        if g_obs is not None:
            g_obs = g_obs.contiguous()
        _lib.step_bwd(cfg.params, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags,
                      state, action, g_state_out, g_obs, g_state, g_action, ctx.wind)
        return g_state, g_action, g_copy, None, None


class Dynamics:
    action_type_alias: Dict = _ALIAS

    def __init__(
            self,
            num: int = 1,
            action_type: str = "bodyrate",
            ori_output_type: str = "quaternion",
            seed: int = 42,
            dt: float = 0.005,
            ctrl_dt: float = 0.03,
            ctrl_delay: bool = True,
            comm_delay: float = 0.06,
            action_space: Tuple[float, float] = (-1, 1),
            device: Union[str, th.device] = "cuda",
            integrator: str = "euler",
            drag_random: float = 0,
            cfg: str = "drone_state",
            wind_settings: Optional[List] = (0, 0, 0),
            rotor_sim: bool = True,
            debug_checks: bool = False,
            random_reset_time: bool = False,
    ):
        assert action_type in ["bodyrate", "thrust", "velocity", "position"]
        assert ori_output_type in ["quaternion", "euler"]
        if integrator not in ("euler", "rk4"):
            raise ValueError("type should be one of ['euler', 'rk4']")
        self.device = th.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(
                "visfly_b200.Dynamics computes on a CUDA device only (sm_100a kernels, no CPU fallback); "
                f"got device={device!r}")
        if self.device.index is None:
            self.device = th.device("cuda", th.cuda.current_device() if th.cuda.is_available() else 0)
        _lib.load(require_cuda=True)

        self.num = num
        self.action_type = _ALIAS[action_type]
        self.angular_output_type = ori_output_type
        self._is_quat_output = ori_output_type == "quaternion"
        self.dt, self.ctrl_dt = dt, ctrl_dt
        if not th.as_tensor(ctrl_dt) % th.as_tensor(dt) == 0:
            raise ValueError("ctrl_dt should be a multiple of dt")
        self._interval_steps = int(ctrl_dt / dt)
        self._comm_delay_steps = int(comm_delay / ctrl_dt)
        self._integrator = integrator
        self._ctrl_delay = ctrl_delay
        self._rotor_sim = rotor_sim
        self._debug_checks = debug_checks
        self._random_reset_time = random_reset_time
        self._drag_random = drag_random

        self.set_seed(seed)
        self._model = load_drone_model(cfg, dt)
        self.m = self._model.m.to(self.device)
        self.name = self._model.name
        self._normal_params = action_scaling(self._model, self.action_type, action_space)
        self._wind, self._wind_fn = self._parse_wind(wind_settings)
        self._params = build_vf_params(self._model, self.action_type, self._normal_params, self._wind)
        for v in self._normal_params.values():
            v.to(self.device)
        self._bd_thrust = self._model.bd_thrust
        self._init_thrust = self._model.init_thrust.to(self.device)
        self._init_motor_omega = self._model.init_motor_omega.to(self.device)
        self.wind_velocity = th.tensor(self._wind, dtype=th.float32, device=self.device).reshape(3, 1)
        self._cfg = StepConfig(self._params, self._interval_steps, _lib.INTEGRATOR_ID[integrator],
                               self.action_type.value, _lib.FLAG_CTRL_DELAY if ctrl_delay else 0)
        self._wind_rows = None                    # (N,4) per-agent wind handed to the kernel (wind functions only)
        self._constructing = True
        self.reset()
        self._constructing = False
        if self._wind_fn is not None:             # reference dynamics.py:172-174
            self._wind_1 = th.zeros((3, num), device=self.device)
            self._wind_2 = th.zeros((3, num), device=self.device)
            self.update_wind()

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _parse_wind(wind_settings):
        """``(constant wind, None)`` or ``((0,0,0), (f1, f2))`` for the reference's wind functions
        (dynamics.py:132-165): six expression strings in ``x`` (= per-agent time ``t``, shape (N,)) and ``y`` (= the
        previous value of that wind component, shape (N,)); ``wind = f1(t, wind_1) + f2(t, wind_2)`` is re-evaluated
        once per control step (:384-388).  The expressions are arbitrary Python, evaluated on the engine's device with
        ordinary tensor ops; the kernel receives the resulting per-agent wind vector."""
        if wind_settings is None:
            return (0.0, 0.0, 0.0), None
        if isinstance(wind_settings, (list, tuple)) and len(wind_settings) == 3 and \
                all(isinstance(w, (int, float)) for w in wind_settings):
            return tuple(float(w) for w in wind_settings), None
        if isinstance(wind_settings, (list, tuple)) and len(wind_settings) == 6 and \
                all(isinstance(w, str) for w in wind_settings):
            fx = [eval("lambda x,y:" + w) for w in wind_settings]       # noqa: S307 — same contract as the reference
            f1 = lambda x, y: th.stack([fx[0](x, y[0]), fx[1](x, y[1]), fx[2](x, y[2])])
            f2 = lambda x, y: th.stack([fx[3](x, y[0]), fx[4](x, y[1]), fx[5](x, y[2])])
            return (0.0, 0.0, 0.0), (f1, f2)
        raise ValueError("wind_settings should be [wx, wy, wz] or a list of six expression strings in x (time) and "
                         "y (previous value); the reference's 3-string form raises at construction "
                         "(dynamics.py:158-161 builds 3-argument lambdas that update_wind calls with two)")

    def update_wind(self):
        """Reference dynamics.py:384-388: called at the start of every control step."""
        if self._wind_fn is None:
            return
        t = self.t
        self._wind_1 = self._wind_fn[0](t, self._wind_1)
        self._wind_2 = self._wind_fn[1](t, self._wind_2)
        self.wind_velocity = (self._wind_1 + self._wind_2).to(th.float32)
        rows = th.zeros((self.num, 4), dtype=th.float32, device=self.device)
        rows[:, :3] = self.wind_velocity.detach().T
        self._wind_rows = rows
        self._obs_t = None            # the reported velocity includes the wind (dynamics.py:750-752)

    def _randomize_drag(self):
        """Reference dynamics.py:244-246: the drag means are (3,1), so ONE random scale per axis is drawn and shared by
        all agents; it becomes the k_lin / k_quad of the parameter block the kernels read.  (The reference's partial
        reset, :265-267, indexes the (3,1) means with agent indices and raises for any index but 0; here a partial
        reset keeps the coefficients.)"""
        dr = self._drag_random
        lin, quad = self._model.linear_drag, self._model.quad_drag
        lin = lin * (((th.rand_like(lin) - 0.5) * 2 * dr).clamp(-0.5, .5) + 1)
        quad = quad * (((th.rand_like(quad) - 0.5) * 2 * dr).clamp(-0.5, .5) + 1)
        for i in range(3):
            self._params.k_lin[i] = float(lin[i, 0])
            self._params.k_quad[i] = float(quad[i, 0])

    def set_seed(self, seed=42):
        th.manual_seed(seed)

    def close(self):
        pass

    def _f(self, x, cols):
        """(n,cols) float32 contiguous copy on the engine's device."""
        t = x.detach() if isinstance(x, th.Tensor) else th.as_tensor(x)
        return t.to(device=self.device, dtype=th.float32, copy=True).reshape(-1, cols).contiguous()

    def _assemble(self, n, pos, ori, vel, ori_vel, motor_omega, wind_t=None):
        """Rows of packed state (5,n,4) and observation (n,13) for freshly (re)initialised agents.
        ``wind_t``: (n,3) wind of exactly these agents when the wind is per-agent (wind functions)."""
        dev = self.device
        pos = th.zeros((n, 3), device=dev) if pos is None else self._f(pos, 3)
        vel = th.zeros((n, 3), device=dev) if vel is None else self._f(vel, 3)
        rate = th.zeros((n, 3), device=dev) if ori_vel is None else self._f(ori_vel, 3)
        if ori is None:
            quat = th.zeros((n, 4), device=dev)
            quat[:, 0] = 1
        else:
            quat = self._f(ori, 4)
        motor = self._init_motor_omega.expand(n, 4) if motor_omega is None else self._f(motor_omega, 4)
        zero = th.zeros((n, 1), device=dev)       # angular acceleration restarts at 0 (dynamics.py:239,258)
        packed = th.stack([th.cat([pos, zero], 1), quat, th.cat([vel, zero], 1), th.cat([rate, zero], 1),
                           motor.to(th.float32)])
        obs = th.cat([pos, quat, vel + (self.wind_velocity.T if wind_t is None else wind_t), rate], 1)
        return packed.contiguous(), obs.contiguous()

    def reset(self, pos=None, ori=None, vel=None, ori_vel=None, motor_omega=None, thrusts=None, t=None,
              indices=None):
        """Reference ``Dynamics.reset`` (dynamics.py:218-269); inputs are (n,k) row-major and are copied."""
        dev = self.device
        if indices is None:
            n = self.num
            self._state, self._obs = self._assemble(n, pos, ori, vel, ori_vel, motor_omega)
            self._t_base = th.zeros((n,), device=dev) if t is None else \
                th.as_tensor(t, dtype=th.float32, device=dev).reshape(n).clone()
            self._n_steps = 0
            self._pre_action = [th.zeros((n, 4), device=dev) for _ in range(self._comm_delay_steps)]
            self._prev, self._ext, self._fresh = None, None, None
            self._t_provider = None
            self._t_custom = t is not None
            self._thrusts_given = None if thrusts is None else self._f(thrusts, 4)
            if self._drag_random and not self._constructing:
                self._randomize_drag()
        else:
            idx = th.as_tensor(indices, device=dev, dtype=th.int64).reshape(-1)
            m = idx.numel()
            rows, obs_rows = self._assemble(m, pos, ori, vel, ori_vel, motor_omega,
                                            self.wind_velocity.T[idx] if self.wind_velocity.shape[1] > 1 else None)
            # functional masked overwrite: the reset agents' upstream gradient is cut, the fresh state is a
            # constant — the reference's in-place index_put gives exactly this (dynamics.py:249-263, App. F)
            self._state = self._state.index_copy(1, idx, rows)
            self._obs = self._obs.index_copy(0, idx, obs_rows)
            if t is None:
                t_new = th.rand((m,), device=dev) * 3.14 * 2 if self._random_reset_time else th.zeros((m,), device=dev)
            else:
                t_new = th.as_tensor(t, dtype=th.float32, device=dev).reshape(m)
            self._t_base = self._t_base.index_copy(0, idx, t_new - self._n_steps * self.ctrl_dt)
            self._pre_action = [a.index_fill(0, idx, 0.0) for a in self._pre_action]
            self._t_custom = self._t_custom or t is not None or self._random_reset_time
            fresh = th.zeros((self.num,), dtype=th.bool, device=dev).index_fill(0, idx, True)
            self._mark_fresh(fresh)
            if thrusts is not None:
                base = self.thrusts.detach().clone()
                base[idx] = self._f(thrusts, 4)
                self._thrusts_given = base
        return self.state

    def reset_where(self, mask: th.Tensor, pos=None, ori=None, vel=None, ori_vel=None, motor_omega=None, t=None):
        """Mask form of the partial reset: ``pos`` etc. are full (N,k) batches, rows where ``mask`` is True
        replace the current state.  Same semantics as ``reset(indices=where(mask))`` (gradient of the replaced
        agents is cut, FIFO rows zeroed) but with no device->host round trip, so it can sit in the step loop."""
        n = self.num
        mask = mask.to(self.device).reshape(n)
        rows, obs_rows = self._assemble(n, pos, ori, vel, ori_vel, motor_omega)
        self._state = th.where(mask.view(1, n, 1), rows, self._state)
        m1 = mask.view(n, 1)
        self._obs = th.where(m1, obs_rows, self._obs)
        if t is None:
            t_new = th.rand((n,), device=self.device) * 3.14 * 2 if self._random_reset_time else 0.0
        else:
            t_new = th.as_tensor(t, dtype=th.float32, device=self.device).reshape(n)
        self._t_base = th.where(mask, t_new - self._n_steps * self.ctrl_dt, self._t_base)
        self._pre_action = [th.where(m1, 0.0, a) for a in self._pre_action]
        self._t_custom = self._t_custom or t is not None or self._random_reset_time
        self._mark_fresh(mask)
        return self.state

    def _mark_fresh(self, mask: th.Tensor):
        """Agents re-initialised since the last step: their diagnostics (acceleration, thrusts) read as the
        reset values (reference dynamics.py:254,259) instead of the last step's."""
        self._fresh = mask if self._fresh is None else (self._fresh | mask)
        self._ext = None

    def detach(self):
        """Cut the autograd graph at the current state (reference dynamics.py:176-190)."""
        self._state = self._state.detach()
        self._obs = self._obs.detach()
        self._pre_action = [a.detach() for a in self._pre_action]

    # -- comm-delay FIFO (reference dynamics.py:323-328) ----------------------------------------------------
    # The reference stores a CLONE of every action (`action.T.clone()`).  Here the clone is made by the step kernel
    # itself: the launch that consumes the delayed action also copies the newly arrived one into its own output slab
    # (`fifo_push` -> `fifo_copy`, +32 B per agent, no extra launch), and that engine-owned copy is what waits in
    # `_pre_action`.  The caller may therefore overwrite its action buffer as soon as step() returns, like with the
    # reference.  Actions that reach the device through a conversion (numpy / CPU tensors, other dtypes, strided
    # views) are already private copies and enter the FIFO directly.
    def _as_device_action(self, action):
        """``(action on the device, owned)``: owned = the tensor was created here and nobody else can write to it."""
        owned = False
        if not isinstance(action, th.Tensor):
            action = th.from_numpy(np.asarray(action))
        if action.dtype is not th.float32 or action.device != self.device:
            action, owned = action.to(device=self.device, dtype=th.float32), True
        if action.shape != (self.num, 4):
            raise ValueError(f"action must have shape ({self.num}, 4), got {tuple(action.shape)}")
        if not action.is_contiguous():
            action, owned = action.contiguous(), True
        return action, owned

    # ------------------------------------------------------------------------------------------
    _fifo_ring = False          # True while a step is being recorded as a CUDA graph (see _step_ring)

    def step(self, action) -> th.Tensor:
        """One control step; ``action`` is (N,4) in ``action_space``; returns ``state`` (reference :319-372)."""
        action, owned = self._as_device_action(action)
        if self._fifo_ring and self._comm_delay_steps and not th.is_grad_enabled():
            return self._step_ring(action)
        push = None
        if self._comm_delay_steps:                                   # dynamics.py:323-326
            if owned:
                self._pre_action.append(action)
            else:
                push = action                                        # cloned by the launch below
            action = self._pre_action.pop(0)
        state, cfg = self._state, self._cfg
        if self._wind_fn is not None:
            self.update_wind()                                       # dynamics.py:320
        wind = self._wind_rows
        if th.is_grad_enabled() and (state.requires_grad or action.requires_grad or
                                     (push is not None and push.requires_grad)):
            self._prev = (state.detach(), action.detach(), None, None)
            self._state, self._obs, copy = ControlStep.apply(state, action, push, cfg, wind)
        else:
            # nothing to differentiate: straight to the launch (an autograd.Function costs ~8 us of host time per
            # call even when no input requires grad; the kernel takes ~11 us at 65 536 agents)
            # without a FIFO the consumed action is the caller's own tensor: remember its version so that the lazy
            # diagnostics (_extras) can tell if it was overwritten in the meantime
            self._prev = (state, action, None if (owned or self._comm_delay_steps) else action._version, None)
            self._state, self._obs, copy = _lib.fast().step_fwd(cfg.params_addr, cfg.substeps, cfg.integrator,
                                                                cfg.action_type, cfg.flags, state, action, wind, push)
        if push is not None:
            self._pre_action.append(copy)
        self._n_steps += 1
        self._ext, self._fresh, self._thrusts_given = None, None, None
        if self._debug_checks:                                       # dynamics.py:333 (device sync!)
            assert bool(th.isfinite(self._obs).all()), "non-finite state after step"
        return self.state

    def _step_ring(self, action: th.Tensor) -> th.Tensor:
        """``step`` with the comm-delay FIFO as a device-resident ring (a step being recorded as a CUDA graph,
        envs/base/task_graph.py): the launch consumes the oldest entry, shifts the others and appends this step's
        action, all in place — the entries keep their addresses.  The diagnostics come out of the same launch (the
        consumed action is gone afterwards)."""
        state, cfg = self._state, self._cfg
        if self._wind_fn is not None:
            self.update_wind()
        state_out = th.empty_like(state)
        obs = th.empty((self.num, 13), dtype=th.float32, device=self.device)
        ext = th.empty((self.num, 8), dtype=th.float32, device=self.device)
        _lib.step_fwd_ring(cfg.params, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, state, self._pre_action,
                           action if action.is_contiguous() else action.contiguous(), state_out, obs, ext,
                           self._wind_rows)
        self._prev = ("ext", ext)
        self._state, self._obs = state_out, obs
        self._n_steps += 1
        self._ext, self._fresh, self._thrusts_given = None, None, None
        return self.state

    def _extras(self) -> th.Tensor:
        """(N,8) [acc, 0, thrusts] of the last sub-step; produced on demand by re-running the step kernel on
        the saved inputs (the hot loop never pays for diagnostics nobody reads)."""
        if self._ext is None:
            ext = th.zeros((self.num, 8), dtype=th.float32, device=self.device)
            rest = self._model.thrust_from_rotor_omega(self._state[4].detach())
            if self._prev is None:
                ext[:, 4:] = rest
            elif len(self._prev) == 2:                       # ("ext", tensor): written by the step's own launch
                ext = self._prev[1]
                if self._fresh is not None:
                    m1 = self._fresh.view(-1, 1)
                    ext = th.cat([th.where(m1, 0.0, ext[:, :4]), th.where(m1, rest, ext[:, 4:])], 1)
            else:
                state_in, action, version, status = self._prev
                if status is not None and self._comm_delay_steps:
                    # fused env step: agents younger than the FIFO flew a zero action (the kernel masks by age)
                    action = th.where((status[:, 0] < self._comm_delay_steps).view(-1, 1), 0.0, action)
                if version is not None and action._version != version:
                    raise RuntimeError(
                        "acceleration / thrusts of the last step are produced on demand from the step's inputs, but the "
                        "action tensor passed to step() was modified in place since (comm_delay=0 keeps no copy): read "
                        "these diagnostics before reusing the action buffer, or pass action.clone()")
                scratch = th.empty_like(state_in)
                _lib.step_fwd(self._cfg.params, self._cfg.substeps, self._cfg.integrator, self._cfg.action_type,
                              self._cfg.flags, state_in, action, scratch, None, ext, self._wind_rows)
                if self._fresh is not None:
                    m1 = self._fresh.view(-1, 1)
                    ext = th.cat([th.where(m1, 0.0, ext[:, :4]), th.where(m1, rest, ext[:, 4:])], 1)
            self._ext = ext
        return self._ext

    # ------------------------------------------------------------------------------------------
    def _normalize(self, action):
        """real units -> [-1,1] (reference dynamics.py:271-317; used by deployment code only)."""
        if not isinstance(action, th.Tensor):
            action = th.from_numpy(np.asarray(action))
        action = action.to(self.device)
        p = self._normal_params
        if self.action_type == ACTION_TYPE.BODYRATE:
            return th.hstack([(action[:, :1] - p["acc"].mean) / p["acc"].half,
                              (action[:, 1:] - p["bodyrate"].mean) / p["bodyrate"].half])
        if self.action_type == ACTION_TYPE.THRUST:
            return (action - p["acc"].mean) / p["acc"].half
        return th.hstack([(action[:, :1] - p["yaw"].mean) / p["yaw"].half,
                          (action[:, 1:] - p["velocity"].mean) / p["velocity"].half])

    def _de_normalize(self, command):
        """[-1,1] -> real units, same return layout as the reference (dynamics.py:692-733)."""
        if not isinstance(command, th.Tensor):
            command = th.from_numpy(np.asarray(command))
        command = command.to(self.device)
        p = self._normal_params
        if self.action_type == ACTION_TYPE.BODYRATE:
            return th.hstack([(command[:, :1] * p["acc"].half + p["acc"].mean) * self.m,
                              command[:, 1:] * p["bodyrate"].half + p["bodyrate"].mean]).T
        if self.action_type == ACTION_TYPE.THRUST:
            return self.m * (command * p["acc"].half + p["acc"].mean).T
        return th.hstack([command[:, :1] * p["yaw"].half + p["yaw"].mean,
                          command[:, 1:] * p["velocity"].half + p["velocity"].mean])

    # -- views (reference dynamics.py:735-826) -----------------------------------------------------
    @property
    def _orientation(self) -> Quaternion:
        return Quaternion.from_tensor(self._obs[:, 3:7].T)

    @property
    def position(self):
        return self._obs[:, 0:3]

    @property
    def orientation(self):
        if self._is_quat_output:
            return self._obs[:, 3:7]
        return self._orientation.toEuler().T

    @property
    def direction(self):
        return self._orientation.x_axis.T

    @property
    def velocity(self):
        return self._obs[:, 7:10]

    @property
    def angular_velocity(self):
        return self._obs[:, 10:13]

    @property
    def acceleration(self):
        return self._extras()[:, 0:3]

    @property
    def angular_acceleration(self):
        s = self._state
        return th.stack([s[0, :, 3], s[2, :, 3], s[3, :, 3]], 1)

    @property
    def t(self):
        if self._t_provider is not None:     # fused env step: per-agent step count (+ the offsets given at reset)
            return self._t_provider()
        return self._t_base + self._n_steps * self.ctrl_dt

    # the (N,13) observation is a kernel output; it is rebuilt from the packed state only if a path that does
    # not produce it (fused env step with a task-specific observation) ran last
    @property
    def _obs(self):
        if self._obs_t is None:
            s = self._state
            self._obs_t = th.cat([s[0, :, :3], s[1], s[2, :, :3] + self.wind_velocity.T, s[3, :, :3]], 1)
        return self._obs_t

    @_obs.setter
    def _obs(self, value):
        self._obs_t = value

    @property
    def motor_omega(self):
        return self._state[4]

    @property
    def thrusts(self):
        if self._thrusts_given is not None:
            return self._thrusts_given
        if self._ctrl_delay:
            return self._model.thrust_from_rotor_omega(self._state[4])   # dynamics.py:516, differentiable
        return self._extras()[:, 4:8]

    @property
    def state(self):
        if self._is_quat_output:
            return self._obs
        return th.hstack([self.position, self.orientation, self.velocity, self.angular_velocity])

    @property
    def is_quat_output(self):
        return self._is_quat_output

    @property
    def full_state(self):
        return th.hstack([self.position, self.orientation, self.velocity, self.angular_velocity,
                          self.motor_omega, self.thrusts, self.t.unsqueeze(1)])

    @property
    def extend_state(self):
        return th.hstack([self.position, self.orientation, self.velocity, self.angular_velocity,
                          self.acceleration, self.angular_acceleration, self.motor_omega, self.thrusts,
                          self.t.unsqueeze(1)])

    @property
    def R(self):
        return self._orientation.R

    @property
    def xz_axis(self):
        return self._orientation.xz_axis

    @property
    def packed_state(self) -> th.Tensor:
        """The HBM-resident (5,N,4) state (see include/visfly_b200.h)."""
        return self._state
