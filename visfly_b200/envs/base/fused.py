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

import collections
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


def generator_spec(gen, spec: P.VfEnvSpec, watch: Optional[list] = None) -> bool:
    """Fill the reset-sampler part of ``spec`` from a state generator; False if the kernel cannot express it.
    The generator's tables are appended to ``watch`` (their version counters are checked every step)."""
    boxes = gen.randomizers if isinstance(gen, UnionRandomizer) else [gen]
    if not 1 <= len(boxes) <= P.GEN_MAX_BOXES:
        return False
    kinds = set()
    for b, g in enumerate(boxes):
        spec.gen_heading[b] = 0
        if isinstance(g, UniformStateRandomizer):
            if g.test:                         # deterministic evaluation grid: sampled by the generic path
                return False
            spec.gen_heading[b] = int(bool(g.heading))     # yaw towards the box centre (randomization.py:162-165)
            kinds.add(P.GEN_UNIFORM)
            mean, half = g._mean, g._half
        elif isinstance(g, NormalStateRandomizer):
            kinds.add(P.GEN_NORMAL)
            mean, half = g._mean, g._std
        else:
            return False
        if watch is not None:
            watch.extend((mean, half))
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


_TASK_METHODS = ("get_reward", "get_success", "get_failure", "get_observation")


def _ver(t):
    return -1 if t is None else t._version


def watched(name: str) -> property:
    """Attribute whose assignment bumps the owner's ``_gen`` counter (see ``FusedEnvStep.refresh``)."""
    key = "_w_" + name

    def get(self):
        try:
            return self.__dict__[key]
        except KeyError:
            raise AttributeError(name) from None

    def put(self, value):
        d = self.__dict__
        d[key] = value
        d["_gen"] = d.get("_gen", 0) + 1

    return property(get, put)


class FusedEnvStep:
    """Owns the spec and the per-agent env status records of one env object while the fused path is active.

    Status lives in ONE ``int32 (N,4)`` tensor (``VfEnvStatus``: step count, return, flag bits | gate, gates passed);
    every step reads the current one and writes a fresh one into its output slab, so the record a step started from
    stays intact (the adjoint kernel reads it; ``rewind`` goes back to it).  ``sc / ret / eb / gate / passed`` are
    views of the current record built on demand."""

    def __init__(self, env, task: int, obs_kind: int):
        dyn = env.envs.dynamics
        self.env, self.n, self.device = env, env.num_agent, env.device
        self.task, self.obs_kind = task, obs_kind
        self.obs_width = 13 if obs_kind == P.OBS_STATE13 else 16
        s = P.VfEnvSpec()
        s.task, s.obs_kind = task, obs_kind
        s.fifo_depth = dyn._comm_delay_steps
        s.init_motor_omega = float(dyn._init_motor_omega)
        s.seed = (int(env.envs.seed) * 0x9E3779B97F4A7C15 + 0x1234567) & 0xFFFFFFFFFFFFFFFF
        self.spec = s
        self.table: Optional[th.Tensor] = None
        self.active = False
        self.global_step = 0
        self.status: Optional[th.Tensor] = None
        self.gate_obs: Optional[th.Tensor] = None    # racing: next gate index as the observation dict carries it
        self.record: Optional[th.Tensor] = None      # episode record of the last step (success / failure views)
        self.t_off: Optional[th.Tensor] = None       # per-agent time offsets given at reset (None: all zero)
        self.step_base: Optional[th.Tensor] = None   # device word added to the Philox step index (graph replays)
        self.sc_mid: Optional[th.Tensor] = None      # split path: the step count task code sees while a step is open
        self.sc_open = False                         # ... built on first read (most task code never looks at it)
        self.finish_out = None                       # (state, status) tensors vf_env_finish writes into (recorded steps)
        self.peer_next = 0                           # address of a VfPeerScatter for the NEXT launch (fused all-gather)
        self.peer_done = False                       # ... and whether that launch has happened
        self.spec_gen = 0                            # bumped whenever the spec is rebuilt (recorded steps go stale)
        self._views = (None, None)
        self._gen_seen, self._ndict, self._watch, self._watch_sum, self._ok = -1, None, (), 0, False
        # host-driven (numpy) mode runs one step ahead of its caller: steps launched but not yet handed out, and the
        # (state, status) the next launch starts from.  See step_host / rewind.
        self._ahead = collections.deque()
        self._front = None
        self._fn = None
        self._bind()

    _CTYPES_REFS = ("_fn", "_stepper", "_params_addr", "_spec_addr", "_host_ring", "_host_turn", "_views", "_ahead",
                    "_front", "_counter", "_flag_seq", "peer_next", "peer_done")

    def __deepcopy__(self, memo):
        """ctypes references are per-object handles: the copy re-creates them against its own env / spec."""
        import copy
        self.rewind()                     # steps launched ahead of the caller are not part of the copied state
        twin = self.__class__.__new__(self.__class__)
        memo[id(self)] = twin
        for k, v in self.__dict__.items():
            if k not in self._CTYPES_REFS:
                setattr(twin, k, copy.deepcopy(v, memo))
        twin._fn = twin._stepper = None   # re-bound on first use (the twin env may not be fully copied yet)
        twin._views = (None, None)
        twin._ahead, twin._front = collections.deque(), None
        twin.peer_next, twin.peer_done = 0, False
        return twin

    def _bind(self):
        self._fn = _lib.fast().EnvStepper
        self._stepper = None
        self._params_addr = ctypes.addressof(self.env.envs.dynamics._cfg.params)
        self._spec_addr = ctypes.addressof(self.spec)

    def _make_stepper(self):
        """Bind what does not change from step to step (parameter blocks, kernel variant, reset table, the device
        word of the Philox step base) into the C++ stepper; rebuilt whenever one of those is replaced."""
        if self._fn is None:                  # a deep copy re-binds its ctypes handles on first use
            self._bind()
        cfg = self.env.envs.dynamics._cfg
        self._stepper = self._fn(self._params_addr, self._spec_addr, cfg.substeps, cfg.integrator, cfg.action_type,
                                 cfg.flags, self.n, self.table, self.obs_width, self.step_base)
        return self._stepper

    # -- views of the current status record ---------------------------------------------------------------
    # built on demand, one field at a time (sc / ret / passed are strided views — no kernel; eb / gate unpack bits)
    def _field(self, k):
        if self._views[0] is not self.status:
            self._views = (self.status, {})
        cache = self._views[1]
        v = cache.get(k)
        if v is None:
            st = self.status
            if k == 0:
                v = st[:, 0]
            elif k == 1:
                v = st[:, 1].view(th.float32)
            elif k == 2:
                v = st[:, 2] & 0xFF
            elif k == 3:
                v = (st[:, 2] >> 8) & 0xFF
            else:
                v = st[:, 3]
            cache[k] = v
        return v

    @property
    def sc(self):
        if not self.sc_open:
            return self._field(0)
        if self.sc_mid is None:
            self.sc_mid = self._field(0) + 1
        return self.sc_mid

    ret = property(lambda self: self._field(1))
    eb = property(lambda self: self._field(2))
    gate = property(lambda self: self._field(3) if self.task == P.TASK_RACING else None)
    passed = property(lambda self: self._field(4) if self.task == P.TASK_RACING else None)

    def t_now(self) -> th.Tensor:
        t = self.sc * self.env.envs.dynamics.ctrl_dt
        return t if self.t_off is None else t + self.t_off

    # -- eligibility ------------------------------------------------------------------------------------
    def refresh(self) -> bool:
        """Re-read the settings that may change between steps; False => the generic path must be used.
        The settings the spec is built from are *watched* attributes of the env objects (``watched`` below): assigning
        one bumps a generation counter, and in-place edits of the tensors among them (``env.target[:] = ...``, generator
        tables) show in their version counters.  Runs once per step: two integer compares and a short sum."""
        d, de = self.env.__dict__, self.env.envs.__dict__
        # instance-level overrides (a monkeypatched task method, a replaced state generator function) are plain
        # attributes: their presence is part of the key
        inst = ("_generate_state" in de, "get_reward" in d or "get_success" in d or "get_failure" in d
                or "get_observation" in d)
        if d.get("_gen", 0) + de.get("_gen", 0) == self._gen_seen and inst == self._ndict:
            vsum = 0
            for t in self._watch:
                vsum += t._version
            if vsum == self._watch_sum:
                return self._ok
        self._ok = self._refresh()
        self._gen_seen, self._ndict = d.get("_gen", 0) + de.get("_gen", 0), inst
        self._watch_sum = sum(t._version for t in self._watch)
        return self._ok

    def _refresh(self) -> bool:
        """Rebuild the spec from the env's live attributes (host reads: only when something changed)."""
        env, envs, s = self.env, self.env.envs, self.spec
        self.rewind()                     # a step launched ahead saw the old settings
        self._stepper = None
        self.spec_gen += 1
        watch = []
        ok = True
        if os.environ.get("VISFLY_B200_NO_FUSED_ENV") or not env.use_fused_step or not envs._imu_noise_free:
            ok = False
        if not envs.dynamics.is_quat_output or "_generate_state" in vars(envs):
            ok = False
        if self.task != P.TASK_CUSTOM and any(m in vars(env) for m in _TASK_METHODS):
            ok = False                                      # instance-level override of a built-in task method
        s.max_episode_steps = int(env.max_episode_steps)
        s.collision_reset = int(bool(env.is_collision_reset))
        s.uav_radius = float(envs.uav_radius)
        s.agent_offset = 0 if envs.shard is None else envs.shard[0]
        box = envs._bboxes[0]
        watch.append(box)
        lo, hi = box[0].tolist(), box[1].tolist()
        for j in range(3):
            s.bbox_lo[j], s.bbox_hi[j] = lo[j], hi[j]
        s.success_radius = float(getattr(env, "success_radius", 0.5))
        if self.task == P.TASK_RACING:
            gates = env.targets
            watch.append(gates)
            g = gates.detach().cpu()
            if not 1 <= g.shape[0] <= 4:
                ok = False
            else:
                s.n_gates = g.shape[0]
                for a in range(g.shape[0]):
                    for j in range(3):
                        s.gates[a][j] = float(g[a, j])
        elif self.task != P.TASK_CUSTOM:
            tgt = env.target
            watch.append(tgt)
            if bool((tgt != tgt[0]).any()):                 # per-agent targets: tensor-op path
                ok = False
            row = tgt[0].tolist()
            for j in range(3):
                s.target[j] = float(row[j])
        table = envs._reset_table
        if table is not None:
            s.gen_kind, s.gen_boxes, self.table = P.GEN_TABLE, 1, table
        else:
            self.table = None
            ok = generator_spec(envs.stateGenerator, s, watch) and ok
        self._watch = tuple(watch)
        return ok

    # -- switching between the two paths ----------------------------------------------------------------
    def enter(self):
        env, dyn = self.env, self.env.envs.dynamics
        racing = self.task == P.TASK_RACING
        sc = env._step_count.to(th.int32)
        if dyn._t_custom:                                   # times given at reset (t=..., random_reset_time)
            self.t_off = (dyn.t - sc * dyn.ctrl_dt).to(th.float32)
        else:
            self.t_off = None
        self.status = _lib.pack_status(
            sc, env._rewards, env.envs._once_collided.to(th.int32) * P.EBIT_ONCE_COLLIDED
            + env._episode_done.to(th.int32) * P.EBIT_EPISODE_DONE,
            env._next_target_i if racing else None, env._past_targets_num if racing else None)
        self.record = None
        env.envs._fused = self
        dyn._t_provider = self.t_now
        self._stepper = None
        self.active = True

    def leave(self):
        """Hand the per-agent state back to the attributes the generic path works on."""
        if not self.active:
            return
        self.rewind()
        env, dyn = self.env, self.env.envs.dynamics
        sc, ret, eb = self.sc.clone(), self.ret.clone(), self.eb
        t_now = self.t_now()
        self.active = False                     # from here on env._step_count / _rewards are plain attributes again
        env._step_count = sc
        env._rewards = ret
        env._episode_done = (eb & P.EBIT_EPISODE_DONE).bool()
        if self.record is not None:
            env._success = (self.record[:, 2].to(th.int32) & P.RBIT_SUCCESS) != 0
        if self.task == P.TASK_RACING:
            env._next_target_i = self.gate.to(th.int64)
            env._past_targets_num = self.passed.to(th.int64)
        # FIFO rows the kernel treated as zero (agent younger than the entry) become real zeros again
        d = len(dyn._pre_action)
        dyn._pre_action = [th.where((sc < d - j).view(-1, 1), 0.0, a) for j, a in enumerate(dyn._pre_action)]
        dyn._t_base = t_now - dyn._n_steps * dyn.ctrl_dt
        dyn._t_provider = None
        env.envs.update_observation()
        env.envs.update_collision()
        env.envs._once_collided = (eb & P.EBIT_ONCE_COLLIDED).bool()
        env.envs._fused = None

    # -- the step -------------------------------------------------------------------------------------------
    def _launch(self, state_in: th.Tensor, action: th.Tensor, status_in: th.Tensor, mirror=None, wind=None, push=None):
        """Allocate the step's outputs and launch ``vf_env_step_fwd``.
        ``mirror``: address of a ``VfEnvMirror`` (page-locked host destinations for obs / reward / done), or None."""
        # hot call: one Python->C++ transition allocates the outputs (torch caching allocator) and launches through
        # the C-ABI on the current stream of the state's device (csrc/vf_torch.cpp, EnvStepper)
        peer = self.peer_next
        out = (self._stepper or self._make_stepper()).step(state_in, action, status_in, self.global_step, 0,
                                                           self.env.keep_terminal_observation, mirror or 0, wind, push,
                                                           peer)
        if peer:
            self.peer_next, self.peer_done = 0, True
        self.global_step += 1
        return out

    def host_slot(self):
        """Page-locked host destinations of obs / reward / done for the numpy output mode: a ring of
        ``env.host_ring_depth`` sets (arrays handed out by a step stay valid for ``depth - 1`` further steps), each
        with its ``VfEnvMirror`` and completion word."""
        ring = getattr(self, "_host_ring", None)
        if ring is None:
            depth = max(2, int(getattr(self.env, "host_ring_depth", 4)))
            ring = []
            self._counter = th.zeros((1,), dtype=th.int32, device=self.device)
            for _ in range(depth):
                obs = th.empty((self.n, self.obs_width), dtype=th.float32, pin_memory=True)
                reward = th.empty((self.n,), dtype=th.float32, pin_memory=True)
                done = th.empty((self.n,), dtype=th.int32, pin_memory=True)
                flag = th.zeros((16,), dtype=th.int32, pin_memory=True)       # one cache line of its own
                m = P.VfEnvMirror(obs.data_ptr(), reward.data_ptr(), done.data_ptr(), flag.data_ptr(),
                                  self._counter.data_ptr(), 0)
                ring.append({"obs": obs, "reward": reward, "done": done, "flag": flag, "mirror": m,
                             "ref": ctypes.addressof(m), "np": (obs.numpy(), reward.numpy(), done.numpy())})
            self._host_ring, self._host_turn, self._flag_seq = ring, 0, 0
        self._host_turn = (self._host_turn + 1) % len(ring)
        slot = ring[self._host_turn]
        self._flag_seq = (self._flag_seq % 0x7FFFFFF0) + 1
        slot["mirror"].flag_value = self._flag_seq
        slot["expect"] = self._flag_seq
        return slot

    def step(self, action, owned: bool = True, grad: bool = False, mirror=None):
        """One env step = one launch.  ``grad=True`` routes through ``EnvControlStep`` so that the returned state,
        observation and reward carry autograd history (backward = one launch of ``vf_env_step_bwd``).
        ``owned``: the action tensor is a private device copy (it came through a host->device conversion) and may
        wait in the comm-delay FIFO as it is; otherwise the launch clones it (``fifo_push`` -> ``fifo_copy``)."""
        env, dyn = self.env, self.env.envs.dynamics
        if self._fn is None:
            self._bind()
        if not self.active:
            self.enter()
        if self._ahead:
            self.rewind()
        push = None
        if dyn._comm_delay_steps:
            if not action.is_contiguous():
                action, owned = action.contiguous(), True
            if owned:
                dyn._pre_action.append(action)
            else:
                push = action
            action = dyn._pre_action.pop(0)
        elif not action.is_contiguous():
            action = action.contiguous()
        if dyn._wind_fn is not None:
            dyn.update_wind()                         # reference dynamics.py:320, from the per-agent time t
        wind = dyn._wind_rows
        state_in, status_in = dyn._state, self.status
        if grad:
            state_out, status, obs, reward, done, record, term, copy, gate = EnvControlStep.apply(
                state_in, action, push, status_in, wind, self)
            dyn._prev = (state_in.detach(), action.detach(), None, status_in)
        else:
            state_out, status, obs, reward, done, record, term, copy, gate = self._launch(
                state_in, action, status_in, mirror, wind, push)
            dyn._prev = (state_in, action, None, status_in)
        if push is not None:
            dyn._pre_action.append(copy)
        self.status, self.record, self.gate_obs = status, record, gate
        if self.t_off is not None:
            self.t_off = th.where(done, 0.0, self.t_off)
        # keep the Dynamics object coherent (lazy views, diagnostics).  The diagnostics re-run the step on its inputs
        # (dyn._prev, set above): agents younger than the FIFO flew a zero action — the kernel masks by age — so the
        # start-of-step status record rides along and _extras applies the same mask
        dyn._state = state_out
        dyn._obs_t = obs if self.obs_kind == P.OBS_STATE13 else None
        dyn._n_steps += 1
        dyn._ext, dyn._thrusts_given = None, None
        dyn._fresh = done
        env.envs._collision_stale = True
        env._reward, env._done = reward, done
        return obs, reward, done, record, term


    # -- split path: caller-defined task between two launches --------------------------------------------------------
    def finish(self, state: th.Tensor, reward: th.Tensor, success: th.Tensor, failure: th.Tensor, grad: bool):
        """Second launch of the split env step (``vf_env_finish``): step count, collision flags, accumulation,
        termination, episode record and auto-reset, given the task's ``reward / success / failure`` on the state the
        control step reached.  Returns ``(state', obs13 | None, done, record)``; with ``grad`` the state passes its
        gradient through for agents that were not re-initialised (the reference's in-place overwrite cuts it for the
        others, dynamics.py:249-263)."""
        dyn = self.env.envs.dynamics
        cfg = dyn._cfg
        # comm-delay FIFO rows of re-initialised agents (dynamics.py:262-263): zeroed by the launch itself where the
        # entries are plain engine-owned tensors; with autograd history in the FIFO the caller blends them (where)
        fifo = dyn._pre_action if dyn._pre_action and not grad else None
        if fifo is not None and any(a.requires_grad or not a.is_contiguous() for a in fifo):
            fifo = None
        out_state, out_status = self.finish_out if (self.finish_out is not None and not grad) else (None, None)
        state_out, status, obs, done, record = _lib.env_finish(
            cfg.params, self.spec, 0, self.global_step, state.detach(), self.status,
            reward.detach().to(th.float32).contiguous(), success.contiguous(), failure.contiguous(),
            want_obs=not grad, wind=dyn._wind_rows, reset_table=self.table, step_base=self.step_base, fifo=fifo,
            state_out=out_state, status_out=out_status)
        self.fifo_zeroed = fifo is not None
        self.global_step += 1
        if grad and state.requires_grad:
            state_out = _ResetBlend.apply(state, state_out, done)
        self.status, self.record = status, record
        if self.t_off is not None:
            self.t_off = th.where(done, 0.0, self.t_off)
        return state_out, obs, done, record

    # -- host-driven mode: one step ahead of the caller ---------------------------------------------------------
    # With a comm-delay FIFO of depth d >= 1 the action handed to call t is consumed by step t+d, so when call t
    # arrives everything step t+1 needs is already known.  step_host therefore launches step t+1 BEFORE it waits for
    # step t: the GPU always has the next kernel queued behind the running one and the zero-copy PCIe stores of one
    # step overlap with the caller's work on the previous one.  The results handed out are bit-identical to stepping
    # synchronously (steps are launched in order, each from its predecessor's outputs); what the env / Dynamics
    # objects show (state, status, info) is always the step that was handed out, and anything that touches the env
    # other than step() first drops the step in flight (rewind) — it is recomputed by the next call.
    def _launch_ahead(self):
        dyn = self.env.envs.dynamics
        action = dyn._pre_action.pop(0)
        slot = self.host_slot()
        state_in, status_in = self._front if self._front is not None else (dyn._state, self.status)
        out = self._launch(state_in, action, status_in, slot["ref"], None, None)
        self._front = (out[0], out[1])
        self._ahead.append((out, slot, state_in, status_in, action))

    def rewind(self):
        """Drop the steps launched ahead of the caller: their actions go back to the head of the FIFO."""
        if not self._ahead:
            self._front = None
            return
        dyn = self.env.envs.dynamics
        while self._ahead:
            _, _, _, _, action = self._ahead.pop()
            dyn._pre_action.insert(0, action)
            self.global_step -= 1
        self._front = None

    def step_host(self, stage_action):
        """One env step for a caller that lives in host memory.  ``stage_action()`` starts the host->device copy of
        this call's action and returns the device tensor.  Returns ``(obs, reward, done, record, term, slot)`` of the
        step this call hands out; ``slot`` holds the page-locked numpy views the kernel wrote."""
        env, dyn = self.env, self.env.envs.dynamics
        if self._fn is None:
            self._bind()
        if not self.active:
            self.enter()
        dyn._pre_action.append(stage_action())
        while len(self._ahead) < 2 and dyn._pre_action:
            self._launch_ahead()
        out, slot, state_in, status_in, action = self._ahead.popleft()
        state_out, status, obs, reward, done, record, term, _, gate = out
        _lib.fast().wait_flag(slot["flag"].data_ptr(), slot["expect"])
        self.status, self.record, self.gate_obs = status, record, gate
        if self.t_off is not None:
            self.t_off = th.where(done, 0.0, self.t_off)
        dyn._prev = (state_in, action, None, status_in)
        dyn._state = state_out
        dyn._obs_t = obs if self.obs_kind == P.OBS_STATE13 else None
        dyn._n_steps += 1
        dyn._ext, dyn._thrusts_given = None, None
        dyn._fresh = done
        env.envs._collision_stale = True
        env._reward, env._done = reward, done
        return obs, reward, done, record, term, slot


class _ResetBlend(th.autograd.Function):
    """``where(done, fresh, state)`` as computed by ``vf_env_finish``: value = the kernel's output, gradient = the
    incoming one for agents that kept their state, zero for re-initialised ones."""

    @staticmethod
    def forward(ctx, state, state_out, done):
        ctx.save_for_backward(done)
        return state_out.view_as(state_out)

    @staticmethod
    def backward(ctx, g):
        done, = ctx.saved_tensors
        return th.where(done.view(1, -1, 1), 0.0, g), None, None


class EnvControlStep(th.autograd.Function):
    """``(state, action, push) -> (state', status', obs, reward, done, record, terminal obs, fifo copy)`` — the fused
    env step with its hand-derived adjoint; saves nothing but the step's inputs (the start-of-step status record is
    one of them: the forward writes the new record elsewhere)."""

    @staticmethod
    def forward(ctx, state: th.Tensor, action: th.Tensor, push, status_in: th.Tensor, wind, fz: FusedEnvStep):
        state_out, status, obs, reward, done, record, term, copy, gate = fz._launch(state, action, status_in, None,
                                                                                    wind, push)
        ctx.fz, ctx.wind = fz, wind
        ctx.save_for_backward(state, action, status_in)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(status, done, record, *(t for t in (term, gate) if t is not None))
        return state_out, status, obs, reward, done, record, term, copy, gate

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, g_state_out, _g_status, g_obs, g_reward, _g_done, _g_record, _g_term, g_copy, _g_gate):
        state, action, status_in = ctx.saved_tensors
        fz = ctx.fz
        cfg = fz.env.envs.dynamics._cfg
        g_state, g_action = th.empty_like(state), th.empty_like(action)
        c = lambda t: None if t is None else t.contiguous()
        _lib.env_step_bwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0,
                          state, action, status_in, c(g_state_out), c(g_obs), c(g_reward), g_state, g_action, ctx.wind)
        return g_state, g_action, g_copy, None, None, None
