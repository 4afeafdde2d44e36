"""Host side of the fused env step (``vf_env_step_fwd``): one kernel launch per ``env.step`` for the built-in tasks.

``DroneGymEnvsBase.step`` takes this path when no Python-defined task code is involved: ``is_test=False``, the
task's ``get_reward / get_success / get_failure / get_observation`` are the built-in ones, the IMU noise model is
zero, the state generator is one the kernel can sample (Uniform / Normal / Union of those, or a reset table), the
orientation output is the quaternion and ``env.use_fused_step`` is True.  With ``requires_grad=True`` the step goes
through ``EnvControlStep`` (backward = one launch of ``vf_env_step_bwd``, the hand-derived adjoint of step + reward).
Everything else (custom tasks) keeps using the generic tensor-op path, which has identical semantics;
``tests/test_gpu_env.py`` replays the reference's golden runs and gradients through both.
"""
from __future__ import annotations

import ctypes
import os
from collections.abc import Sequence
from typing import Dict, Optional

import numpy as np
import torch as th

from ... import _lib
from ... import params as P
from ...randomization import NormalStateRandomizer, UniformStateRandomizer, UnionRandomizer


_raw_stream = getattr(th._C, "_cuda_getCurrentRawStream", None) or \
    (lambda index: th.cuda.current_stream(index).cuda_stream)


def generator_spec(gen, spec: P.VfEnvSpec) -> bool:
    """Fill the reset-sampler part of ``spec`` from a state generator; False if the kernel cannot express it."""
    boxes = gen.randomizers if isinstance(gen, UnionRandomizer) else [gen]
    if not 1 <= len(boxes) <= P.GEN_MAX_BOXES:
        return False
    kinds = set()
    for b, g in enumerate(boxes):
        if isinstance(g, UniformStateRandomizer):
            if g.heading:                      # yaw towards the box centre: sampled by the generic path
                return False
            kinds.add(P.GEN_UNIFORM)
            mean, half = g._mean, g._half
        elif isinstance(g, NormalStateRandomizer):
            kinds.add(P.GEN_NORMAL)
            mean, half = g._mean, g._std
        else:
            return False
        mean, half = mean.cpu(), half.cpu()
        for f in range(4):
            for j in range(3):
                spec.gen_mean[b][f][j] = float(mean[f, j])
                spec.gen_half[b][f][j] = float(half[f, j])
    if len(kinds) != 1:
        return False
    spec.gen_kind, spec.gen_boxes = kinds.pop(), len(boxes)
    return True


class RecordInfo(Sequence):
    """Lazy ``info`` list backed by the kernel's per-step episode record ``[return, length, bits, gates]``."""

    _IDLE = {"TimeLimit.truncated": False, "episode_done": False}

    def __init__(self, n, record: th.Tensor, term_obs, ctrl_dt: float, racing: bool, wrap_obs=None):
        """``term_obs``: dict of terminal-observation fields, or (with ``wrap_obs``) the kernel's raw terminal-
        observation tensor — it is turned into the dict only if somebody reads a finished agent's info."""
        self._n, self._record, self._term_raw, self._ctrl_dt, self._racing = n, record, term_obs, ctrl_dt, racing
        self._wrap = wrap_obs
        self._host = None
        self._cache: Dict[int, dict] = {}

    @property
    def _term(self):
        if self._wrap is not None:
            self._term_raw = {} if self._term_raw is None else self._wrap(self._term_raw)
            self._wrap = None
        return self._term_raw

    def _fetch(self):
        if self._host is None:
            self._host = self._record.cpu().numpy()
        return self._host

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        if i in self._cache:
            return self._cache[i]
        ret, length, bits, gates = self._fetch()[i]
        bits = int(bits)
        if not bits & P.RBIT_DONE:
            info = dict(self._IDLE)
        else:
            length = np.asarray(int(length), dtype=np.int32)
            info = {
                "episode_done": bool(bits & P.RBIT_EPISODE_DONE),
                "is_success": bool(bits & P.RBIT_SUCCESS),
                "episode": {"r": np.asarray(ret, dtype=np.float32), "l": length,
                            "t": np.asarray(length * self._ctrl_dt, dtype=np.float32),
                            "extra": {"collision": np.asarray(bool(bits & P.RBIT_COLLIDED))}},
                "terminal_observation": {k: v[i] for k, v in self._term.items()},
                "TimeLimit.truncated": bool(bits & P.RBIT_TRUNCATED),
            }
            if self._racing:
                info["episode"]["extra"]["past_gate"] = int(gates)
        self._cache[i] = info
        return info

    def copy(self):
        return self

    def done_indices(self):
        return np.nonzero(self._fetch()[:, 2].astype(np.int64) & P.RBIT_DONE)[0]

    def episode_done_tensor(self) -> th.Tensor:
        """``[info[i]["episode_done"] for i in range(n)]`` as one bool device tensor (no host round trip)."""
        return (self._record[:, 2].to(th.int32) & P.RBIT_EPISODE_DONE) != 0


class FusedEnvStep:
    """Owns the spec and the in-place per-agent env state of one env object while the fused path is active."""

    def __init__(self, env, task: int, obs_kind: int, target=None, gates=None, success_radius: float = 0.5):
        dyn = env.envs.dynamics
        self.env, self.n, self.device = env, env.num_agent, env.device
        self.task, self.obs_kind = task, obs_kind
        self.obs_width = 13 if obs_kind == P.OBS_STATE13 else 16
        s = P.VfEnvSpec()
        s.task, s.obs_kind = task, obs_kind
        s.uav_radius = env.envs.uav_radius
        lo, hi = env.envs._bboxes[0][0].tolist(), env.envs._bboxes[0][1].tolist()
        for j in range(3):
            s.bbox_lo[j], s.bbox_hi[j] = lo[j], hi[j]
            s.target[j] = 0.0 if target is None else float(target[j])
        s.success_radius = success_radius
        s.n_gates = 0
        if gates is not None:
            g = th.as_tensor(gates).cpu()
            s.n_gates = g.shape[0]
            for a in range(g.shape[0]):
                for j in range(3):
                    s.gates[a][j] = float(g[a, j])
        s.fifo_depth = dyn._comm_delay_steps
        s.init_motor_omega = float(dyn._init_motor_omega)
        s.seed = (int(env.envs.seed) * 0x9E3779B97F4A7C15 + 0x1234567) & 0xFFFFFFFFFFFFFFFF
        self.spec = s
        self.table: Optional[th.Tensor] = None
        self.active = False
        self.global_step = 0
        self.sc = self.ret = self.eb = self.gate = self.passed = None
        self._key, self._ok = None, False
        self._fn = None
        self._bind()

    _CTYPES_REFS = ("_fn", "_stepper", "_params_addr", "_spec_addr", "_host_ring", "_host_turn")

    def __deepcopy__(self, memo):
        """ctypes references are per-object handles: the copy re-creates them against its own env / spec."""
        import copy
        twin = self.__class__.__new__(self.__class__)
        memo[id(self)] = twin
        for k, v in self.__dict__.items():
            if k not in self._CTYPES_REFS:
                setattr(twin, k, copy.deepcopy(v, memo))
        twin._fn = twin._stepper = None   # re-bound on first use (the twin env may not be fully copied yet)
        return twin

    def _bind(self):
        self._fn = _lib.fast().EnvStepper
        self._stepper = None
        self._params_addr = ctypes.addressof(self.env.envs.dynamics._cfg.params)
        self._spec_addr = ctypes.addressof(self.spec)

    def _make_stepper(self):
        """Bind what does not change from step to step (parameter blocks, kernel variant, in-place status buffers,
        reset table) into the C++ stepper; rebuilt whenever one of those is replaced (``enter``, ``_refresh``)."""
        cfg = self.env.envs.dynamics._cfg
        self._stepper = self._fn(self._params_addr, self._spec_addr, cfg.substeps, cfg.integrator, cfg.action_type,
                                 cfg.flags, self.n, self.sc, self.ret, self.eb, self.gate, self.passed, self.table,
                                 self.obs_width)
        return self._stepper

    # -- eligibility ------------------------------------------------------------------------------------
    def refresh(self) -> bool:
        """Re-read the settings that may change between steps; False => the generic path must be used.
        The answer is cached on the identity of everything it depends on (this runs once per step)."""
        env = self.env
        envs = env.envs
        key = (id(envs.stateGenerator), id(envs._reset_table), env.max_episode_steps, env.is_collision_reset,
               "_generate_state" in envs.__dict__, env.use_fused_step)
        if key == self._key:
            return self._ok
        self._key = key
        self._ok = self._refresh()
        return self._ok

    def _refresh(self) -> bool:
        env, s = self.env, self.spec
        self._stepper = None
        if os.environ.get("VISFLY_B200_NO_FUSED_ENV") or not env.use_fused_step or not env.envs._imu_noise_free:
            return False
        if not env.envs.dynamics.is_quat_output or "_generate_state" in vars(env.envs):
            return False
        if env.envs.dynamics._wind_fn is not None:      # per-agent wind functions: generic path (vf_step_fwd + wind)
            return False
        s.max_episode_steps = int(env.max_episode_steps)
        s.collision_reset = int(bool(env.is_collision_reset))
        table = env.envs._reset_table
        if table is not None:
            s.gen_kind, s.gen_boxes, self.table = P.GEN_TABLE, 1, table
            return True
        self.table = None
        return generator_spec(env.envs.stateGenerator, s)

    # -- switching between the two paths ----------------------------------------------------------------
    def enter(self):
        env, dyn = self.env, self.env.envs.dynamics
        self.sc = env._step_count.to(th.int32).clone()
        self.ret = env._rewards.detach().to(th.float32).clone()
        self.eb = env.envs._once_collided.to(th.uint8) * P.EBIT_ONCE_COLLIDED \
            + env._episode_done.to(th.uint8) * P.EBIT_EPISODE_DONE
        if self.task == P.TASK_RACING:
            self.gate = env._next_target_i.to(th.int32).clone()
            self.passed = env._past_targets_num.to(th.int32).clone()
        env.envs._fused = self
        for t, dt in ((self.sc, th.int32), (self.ret, th.float32), (self.eb, th.uint8)):
            assert t.is_cuda and t.is_contiguous() and t.dtype == dt
        self._stepper = None
        self.active = True

    def leave(self):
        """Hand the per-agent state back to the attributes the generic path works on."""
        if not self.active:
            return
        env, dyn = self.env, self.env.envs.dynamics
        env._step_count = self.sc.clone()
        env._rewards = self.ret.clone()
        env._episode_done = (self.eb & P.EBIT_EPISODE_DONE).bool()
        env.envs._once_collided = (self.eb & P.EBIT_ONCE_COLLIDED).bool()
        if self.task == P.TASK_RACING:
            env._next_target_i = self.gate.to(th.int64)
            env._past_targets_num = self.passed.to(th.int64)
        # FIFO rows the kernel treated as zero (agent younger than the entry) become real zeros again
        d = len(dyn._pre_action)
        dyn._pre_action = [th.where((self.sc < d - j).view(-1, 1), 0.0, a) for j, a in enumerate(dyn._pre_action)]
        dyn._fifo_versions = [None] * len(dyn._pre_action)
        dyn._t_base = self.sc * dyn.ctrl_dt - dyn._n_steps * dyn.ctrl_dt
        dyn._t_steps = None
        env.envs.update_observation()
        env.envs.update_collision()
        env.envs._once_collided = (self.eb & P.EBIT_ONCE_COLLIDED).bool()
        env.envs._fused = None
        self.active = False

    # -- the step -------------------------------------------------------------------------------------------
    def _launch(self, state_in: th.Tensor, action: th.Tensor, want_saved: bool, mirror=None):
        """Allocate the step's outputs and launch ``vf_env_step_fwd`` (status buffers are updated in place).
        ``mirror``: address of a ``VfEnvMirror`` (page-locked host destinations for obs / reward / done), or None."""
        # hot call: one Python->C++ transition allocates the outputs (torch caching allocator) and launches through
        # the C-ABI on the current stream of the state's device (csrc/vf_torch.cpp, EnvStepper)
        out = (self._stepper or self._make_stepper()).step(state_in, action, self.global_step, 0,
                                                           self.env.keep_terminal_observation, want_saved, mirror or 0)
        self.global_step += 1
        return out

    def host_slot(self):
        """Page-locked host destinations of obs / reward / done for the numpy output mode: two alternating sets (the
        arrays handed out by the previous step stay valid for one more step), each with its ``VfEnvMirror``."""
        ring = getattr(self, "_host_ring", None)
        if ring is None:
            ring = []
            for _ in range(2):
                obs = th.empty((self.n, self.obs_width), dtype=th.float32, pin_memory=True)
                reward = th.empty((self.n,), dtype=th.float32, pin_memory=True)
                done = th.empty((self.n,), dtype=th.int32, pin_memory=True)
                m = P.VfEnvMirror(obs.data_ptr(), reward.data_ptr(), done.data_ptr())
                ring.append({"obs": obs, "reward": reward, "done": done, "mirror": m, "ref": ctypes.addressof(m),
                             "np": (obs.numpy(), reward.numpy(), done.numpy())})
            self._host_ring, self._host_turn = ring, 0
        self._host_turn ^= 1
        return ring[self._host_turn]

    def step(self, action, grad: bool = False, mirror=None, late_action=None):
        """One env step = one launch.  ``grad=True`` routes through ``EnvControlStep`` so that the returned state,
        observation and reward carry autograd history (backward = one launch of ``vf_env_step_bwd``).
        ``late_action`` (numpy mode with a comm-delay FIFO): a callable that stages this step's host action; it is
        called AFTER the launch — the kernel consumes an older FIFO entry, so the staging overlaps with it."""
        env, dyn = self.env, self.env.envs.dynamics
        if self._fn is None:
            self._bind()
        if not self.active:
            self.enter()
        if late_action is not None:
            action = dyn._fifo_pop()
        elif dyn._comm_delay_steps:
            dyn._fifo_push(action)
            action = dyn._fifo_pop()
        if not action.is_contiguous():
            action = action.contiguous()
        state_in = dyn._state
        if grad:
            state_out, obs, reward, done, record, term = EnvControlStep.apply(state_in, action, self)
        else:
            state_out, obs, reward, done, record, term, _ = self._launch(state_in, action, False, mirror)
        if late_action is not None:
            dyn._fifo_push(late_action())
        # keep the Dynamics object coherent (lazy views, diagnostics)
        dyn._prev = (state_in.detach(), action.detach()) if grad else (state_in, action)
        dyn._state = state_out
        dyn._obs_t = obs if self.obs_kind == P.OBS_STATE13 else None
        dyn._n_steps += 1
        dyn._ext, dyn._thrusts_given = None, None
        dyn._fresh = done
        dyn._t_steps = self.sc
        env.envs._collision_stale = True
        env._step_count, env._rewards, env._reward, env._done = self.sc, self.ret, reward, done
        if self.gate is not None:
            env._next_target_i, env._past_targets_num = self.gate, self.passed
        return obs, reward, done, record, term


class EnvControlStep(th.autograd.Function):
    """``(state, action) -> (state', obs, reward | done, record, terminal obs)`` — the fused env step with its
    hand-derived adjoint; saves nothing but the step's inputs and 8 bytes per agent of start-of-step status."""

    @staticmethod
    def forward(ctx, state: th.Tensor, action: th.Tensor, fz: FusedEnvStep):
        state_out, obs, reward, done, record, term, saved = fz._launch(state, action, True)
        ctx.fz = fz
        ctx.save_for_backward(state, action, saved)
        ctx.set_materialize_grads(False)
        outs = (state_out, obs, reward, done, record) + (() if term is None else (term,))
        ctx.mark_non_differentiable(done, record, *(() if term is None else (term,)))
        return outs if term is not None else outs + (None,)

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, g_state_out, g_obs, g_reward, *_):
        state, action, saved = ctx.saved_tensors
        fz = ctx.fz
        cfg = fz.env.envs.dynamics._cfg
        g_state, g_action = th.empty_like(state), th.empty_like(action)
        c = lambda t: None if t is None else t.contiguous()
        _lib.env_step_bwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0,
                          state, action, saved, c(g_state_out), c(g_obs), c(g_reward), g_state, g_action)
        return g_state, g_action, None
