"""``DroneGymEnvsBase`` — the Gym/SB3-style vectorised wrapper (surface of reference
envs/base/droneGymEnv.py:19-633): ``reset() -> obs``, ``step(a) -> (obs, reward, done, info)``, the three output
modes (tensors with grad / detached tensors / numpy), reward accumulation, success / failure / out-of-bounds /
collision / time-limit termination, per-episode ``info`` records and auto-reset of finished agents.

Differences that make it run at 65 536+ agents (the reference's wrapper caps at ~2.5e5 agent-steps/s because of
per-agent Python loops):
  * bookkeeping is functional tensor code on the device, no per-agent loops and no host synchronisation in
    ``step``; finished agents are re-initialised with a mask blend (``reset_agents_where``);
  * ``info`` is a lazy sequence: dictionaries with the reference's keys (droneGymEnv.py:238-275) are built only
    for the entries somebody reads;
  * the range assert on the action (a device sync per step, droneGymEnv.py:144) is opt-in (``debug_checks``);
  * task methods receive ``predicted_obs=None`` uniformly (the reference's Hover/Racing envs crash on it, C4).
World-model latent hooks (``world``, ``deter``/``stoch``) are not part of this path.
"""
from __future__ import annotations

import contextlib
import os
from collections.abc import Sequence
from typing import Dict, List, Optional

import numpy as np
import torch as th

from ...type import ACTION_TYPE, TensorDict
from ._compat import VecEnv, spaces
from .droneEnv import DroneEnvsBase


_RecordInfo = None          # envs/base/fused.py:RecordInfo, bound on first use
_wait_flag = None


class LazyInfo(Sequence):
    """Per-agent info dicts materialised on demand from one snapshot of device tensors."""

    _IDLE = {"TimeLimit.truncated": False, "episode_done": False}

    def __init__(self, n, done, episode_done, success, step_count, rewards, once_collided, terminal_obs,
                 max_episode_steps, ctrl_dt, indiv_rewards=None, extra_fn=None, detach_obs=True):
        self._n = n
        self._dev = dict(done=done, episode_done=episode_done, success=success, step_count=step_count,
                         rewards=rewards, once_collided=once_collided)
        self._terminal_obs, self._indiv, self._extra_fn = terminal_obs, indiv_rewards, extra_fn
        self._max_steps, self._ctrl_dt, self._host = max_episode_steps, ctrl_dt, None
        self._cache: Dict[int, dict] = {}

    def _fetch(self):
        if self._host is None:        # one batched device->host transfer, only if somebody looks
            self._host = {k: v.detach().cpu().numpy() for k, v in self._dev.items()}
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
        h = self._fetch()
        if not h["done"][i]:
            info = dict(self._IDLE)
        else:
            length = h["step_count"][i]
            info = {
                "episode_done": bool(h["episode_done"][i]),
                "is_success": bool(h["success"][i]),
                "episode": {"r": h["rewards"][i], "l": length, "t": np.asarray(length * self._ctrl_dt),
                            "extra": {"collision": h["once_collided"][i]}},
                "terminal_observation": {
                    k: (v[i].detach() if isinstance(v, th.Tensor) else v[i]) for k, v in self._terminal_obs.items()},
                "TimeLimit.truncated": bool(length >= self._max_steps),
            }
            if self._indiv is not None:
                for k, v in self._indiv.items():
                    info["episode"]["extra"][k] = v[i].detach().clone()
            if self._extra_fn is not None:
                self._extra_fn(i, info)
        self._cache[i] = info
        return info

    def copy(self):
        return self

    def done_indices(self):
        return np.nonzero(self._fetch()["done"])[0]

    def episode_done_tensor(self) -> th.Tensor:
        """``[info[i]["episode_done"] for i in range(n)]`` as one bool device tensor (no host round trip)."""
        return self._dev["episode_done"] & self._dev["done"]


def _watched(name):
    from .fused import watched
    return watched(name)


class DroneGymEnvsBase(VecEnv):
    # settings the one-kernel env step bakes into its spec: assigning one is noticed on the next step (fused.watched)
    _gen = 0
    target = _watched("target")
    targets = _watched("targets")
    success_radius = _watched("success_radius")
    max_episode_steps = _watched("max_episode_steps")
    is_collision_reset = _watched("is_collision_reset")
    use_fused_step = _watched("use_fused_step")
    # task envs with their own reward / success / failure code: True records the whole step (control step kernel, the
    # task's tensor ops, wrapper-tail kernel) as a CUDA graph and replays it per step (task_graph.py); the returned
    # tensors are then the graph's buffers, overwritten by the next step — "copy" hands out fresh copies instead
    capture_task_step = bool(os.environ.get("VISFLY_B200_CAPTURE_TASK_STEP"))
    _task_graph, _split_warm = None, 0

    def __init__(
            self,
            num_agent_per_scene: int = 1,
            num_scene: int = 1,
            seed: int = 42,
            visual: bool = False,
            max_episode_steps: int = 1000,
            device="cuda",
            dynamics_kwargs=None,
            random_kwargs=None,
            requires_grad: bool = False,
            scene_kwargs: Optional[Dict] = None,
            sensor_kwargs: Optional[List] = None,
            tensor_output: bool = True,
            is_train: bool = False,
            is_collision_reset: bool = True,
            debug_checks: bool = False,
            shard=None,
    ):
        device = th.device(device)
        self.envs = DroneEnvsBase(
            num_agent_per_scene=num_agent_per_scene, num_scene=num_scene, seed=seed, visual=visual, device=device,
            shard=shard,
            dynamics_kwargs=dict(dynamics_kwargs or {}), random_kwargs=dict(random_kwargs or {}),
            scene_kwargs=dict(scene_kwargs or {}), sensor_kwargs=sensor_kwargs or [])
        self.device = self.envs.device
        self.num_agent = self.num_envs = num_agent_per_scene * num_scene
        self.num_scene, self.num_agent_per_scene = num_scene, num_agent_per_scene
        self.requires_grad = requires_grad
        self.max_sense_radius = 10
        self.tensor_output = tensor_output
        self.is_train = is_train
        self.is_collision_reset = is_collision_reset
        self.debug_checks = debug_checks
        self.max_episode_steps = max_episode_steps

        ori_dim = 3 if self.envs.dynamics.angular_output_type == "euler" else 4
        self.observation_space = spaces.Dict(
            {"state": spaces.Box(low=-np.inf, high=np.inf, shape=(9 + ori_dim,), dtype=np.float32)})
        if self.envs.dynamics.action_type not in (ACTION_TYPE.BODYRATE, ACTION_TYPE.THRUST, ACTION_TYPE.VELOCITY,
                                                  ACTION_TYPE.POSITION):
            raise ValueError("action_type should be one of ['bodyrate', 'thrust', 'velocity', 'position']")
        self.action_space = spaces.Box(low=-1, high=1, shape=(4,), dtype=np.float32)
        self.deter = self.stoch = None

        n, dev = self.num_agent, self.device
        self._step_count = th.zeros((n,), dtype=th.int32, device=dev)
        self._reward = th.zeros((n,), device=dev)
        self._rewards = th.zeros((n,), device=dev)
        self._action = th.zeros((n, 4), device=dev)
        self._obs_tensors = TensorDict({})
        self._observations = TensorDict({})
        self._success = th.zeros(n, dtype=th.bool, device=dev)
        self._failure = th.zeros(n, dtype=th.bool, device=dev)
        self._episode_done = th.zeros(n, dtype=th.bool, device=dev)
        self._done = th.zeros(n, dtype=th.bool, device=dev)
        self._info = None
        self._indiv_rewards = self._indiv_reward = None
        self.keep_terminal_observation = True     # False: skip writing info["terminal_observation"] rows
        self.host_ring_depth = 4                  # numpy mode: arrays returned by a step stay valid for depth-1 more steps
        self.use_fused_step = True                # False: force the generic tensor-op path (debugging / comparison)
        self._fused = None                        # FusedEnvStep, created by built-in tasks (_make_fused)
        self._split = None                        # FusedEnvStep(TASK_CUSTOM) for user-defined tasks (_make_split)
        self.render_mode = ["None"] * n
        self._is_initial = False

    # -- the step ------------------------------------------------------------------------------------
    def _grad_ctx(self):
        return contextlib.nullcontext() if self.requires_grad else th.no_grad()

    def step(self, _action, is_test=False, predict=False, world=None):
        fz = self._fused
        if fz is not None and type(_action) is th.Tensor and self.tensor_output and not (
                is_test or predict or self.requires_grad or self.debug_checks or world is not None) \
                and _action.is_cuda and _action.dtype is th.float32 and self._is_initial and fz.refresh():
            return self._step_fused_tensor(_action)          # the common case, kept in one frame
        assert self._is_initial, "You should call reset() before step()"
        if world is not None or predict:
            raise NotImplementedError("world-model rollouts are not part of the dynamics path")
        fused = self._fused is not None and not is_test and self._fused.refresh()
        # numpy mode + comm-delay FIFO: a step consumes an OLDER action, so the engine runs one step ahead of its
        # caller (FusedEnvStep.step_host) and the host->device copy of the new action rides on a side stream
        ahead = fused and not self.tensor_output and not self.debug_checks and not self.requires_grad \
            and self._fused.spec.fifo_depth >= 1 and self.envs.dynamics._wind_fn is None \
            and not (isinstance(_action, th.Tensor) and _action.is_cuda)
        if ahead:
            return self._step_fused_host(_action)
        self._action = self._stage_action(_action)
        # a device tensor handed in by the caller is still the caller's: the step launch clones it for the FIFO
        self._action_owned = self._action is not _action
        if self.debug_checks:                                   # reference droneGymEnv.py:144 (host sync)
            assert self._action.max() <= 1 and self._action.min() >= -1
        if self._fused is not None:
            if fused:
                return self._step_fused()
            self._fused.leave()
        elif self._split is not None:
            if not is_test and self._split.refresh():
                if self.capture_task_step and self.tensor_output and not self.requires_grad \
                        and not self.debug_checks and self._action.shape == (self.num_agent, 4):
                    return self._step_split_graph()
                return self._step_split()
            self._split.leave()
        with self._grad_ctx():
            self.envs.step(self._action)
            self.get_full_observation()
            self._step_count = self._step_count + 1
            self._success = self.get_success()
            self._failure = self.get_failure()
            assert self._success.dtype == th.bool and self._failure.dtype == th.bool
            if self._indiv_reward is None:
                self._reward = self.get_reward(predicted_obs={})
            else:
                self._indiv_reward = self.get_reward(predicted_obs={})
                assert isinstance(self._indiv_reward, dict) and "reward" in self._indiv_reward
                self._reward = self._indiv_reward["reward"]
                self._indiv_rewards = {k: self._indiv_rewards[k] + v.detach() for k, v in self._indiv_reward.items()}
            self._rewards = self._rewards + self._reward

            episode_done = self._episode_done | self._success | self._failure | self.is_out_bounds
            if self.is_collision_reset:
                episode_done = episode_done | self.is_collision
            self._episode_done = episode_done
            self._done = episode_done | (self._step_count >= self.max_episode_steps)

            done, reward = self._done, self._reward
            info = self._snapshot_info()
            self._info = info
            if not is_test:
                self._auto_reset(done)
        return self._format_step_output(reward, done, info)

    # -- two-launch path for user-defined tasks ---------------------------------------------------------------
    def _make_split(self):
        """Task envs that define their own ``get_success / get_failure / get_reward`` as tensor code get the split env
        step: launch 1 = the control step (``vf_step_fwd``), then the task's tensor ops on the state reached, launch 2 =
        the rest of the wrapper (``vf_env_finish``: collision flags, accumulation, termination, episode record,
        auto-reset with the in-kernel sampler) — instead of the ~100 small kernels of the generic path.  Rewards
        returned as dicts (per-term logging) keep the generic path."""
        from ... import params as P
        from .fused import FusedEnvStep
        if self._indiv_rewards is not None or type(self)._extra_info is not DroneGymEnvsBase._extra_info:
            return None
        return FusedEnvStep(self, P.TASK_CUSTOM, P.OBS_STATE13)

    def _step_split(self):
        global _RecordInfo
        if _RecordInfo is None:
            from .fused import RecordInfo as _RecordInfo
        fz, envs = self._split, self.envs
        dyn = envs.dynamics
        if not fz.active:
            fz.enter()
        with self._grad_ctx():
            dyn.step(self._action)                             # launch 1 (FIFO, wind, autograd history)
            envs._collision_stale = True                       # collision views: recomputed only if the task reads them
            envs.update_observation()
            self.get_full_observation()
            pre_obs = self._obs_tensors
            fz.sc_open, fz.sc_mid = True, None                 # `self._step_count` as the task code sees it: +1
            self._success = self.get_success()
            self._failure = self.get_failure()
            reward = self.get_reward(predicted_obs={})
            if not isinstance(reward, th.Tensor):
                raise ValueError("get_reward changed its return type after reset()")
            fz.sc_open, fz.sc_mid = False, None
            state_out, obs13, done, record = fz.finish(dyn._state, reward, self._success, self._failure,
                                                       self.requires_grad)             # launch 2
            dyn._state, dyn._obs_t = state_out, obs13
            dyn._ext, dyn._fresh = None, done
            if dyn._pre_action and not fz.fifo_zeroed:         # FIFO rows of re-initialised agents (dynamics.py:262-263)
                m1 = done.view(-1, 1)
                dyn._pre_action = [th.where(m1, 0.0, a) for a in dyn._pre_action]
            envs._collision_stale = True
            self._reward, self._done = reward, done
            self._on_reset_where(done)                         # task hook for its own per-agent state
            self.get_full_observation()                        # observation after the auto-reset
        info = _RecordInfo(self.num_agent, record, pre_obs.detach(), dyn.ctrl_dt, False)
        self._info = info
        return self._format_step_output(reward, done, info)

    def _step_split_graph(self):
        """``_step_split`` as one CUDA-graph replay (task_graph.py); the first steps after a reset, a spec change or an
        env the recording does not fit run the eager two-launch path."""
        from . import task_graph as tg
        g, fz = self._task_graph, self._split
        if g is not None and not fz.active:
            fz.enter()                        # back from the generic path (reset_agent_by_id, examine, ...)
        if g is not None and g.matches(self._action):
            return g.replay(self._action, self.capture_task_step == "copy")
        self._task_graph = None
        if self._split_warm < 3 or not fz.active or fz.t_off is not None or self.envs.dynamics._wind_fn is not None:
            self._split_warm += 1
            return self._step_split()
        action = self._action
        g = tg.capture_or_none(self, action)
        if g is None:
            self.capture_task_step = False
            self._action, self._action_owned = action, False
            return self._step_split()
        self._task_graph, self._split_warm = g, 0
        return g.replay(action, self.capture_task_step == "copy")

    # -- one-kernel path (built-in tasks, no autograd) -------------------------------------------------------
    def _make_fused(self):
        """Built-in task envs return a ``FusedEnvStep`` here; ``None`` keeps the generic tensor-op path."""
        return None

    def _builtin_task(self, owner) -> bool:
        """True if this object's task methods are exactly ``owner``'s (not overridden by a user subclass)."""
        return all(getattr(type(self), m) is getattr(owner, m)
                   for m in ("get_reward", "get_success", "get_failure", "get_observation"))

    def _fused_obs(self, obs: th.Tensor) -> TensorDict:
        return TensorDict({"state": obs})

    def _fused_np_obs(self, obs: np.ndarray) -> TensorDict:
        """numpy-mode twin of ``_fused_obs``: the kernel's host copy of the observation, plus host copies of whatever
        else the task puts next to it (constant tensors such as the navigation target are converted once)."""
        out = TensorDict({"state": obs})
        cache = self.__dict__.setdefault("_np_const", {})
        for k, v in self._obs_tensors.items():
            if k == "state":
                continue
            hit = cache.get(k)
            if hit is None or hit[0] is not v:
                hit = cache[k] = (v, v.detach().cpu().numpy())
            out[k] = hit[1]
        return out

    def _step_fused_tensor(self, action):
        """``_stage_action`` + ``_step_fused`` + ``FusedEnvStep.step`` + ``_launch`` for the common case — device
        actions in, tensors out, no autograd — written out in ONE Python frame: the kernel takes ~12 us at 65 536
        agents, every call layer costs a fraction of a microsecond of the ~10 us host budget."""
        global _RecordInfo
        if _RecordInfo is None:
            from .fused import RecordInfo as _RecordInfo
        fz, envs = self._fused, self.envs
        dyn = envs.dynamics
        if not fz.active:
            fz.enter()
        if fz._ahead:
            fz.rewind()
        self._action = action
        if action.device != self.device:
            action = action.to(self.device)
        push, fifo = None, dyn._pre_action
        if fifo:                                              # comm-delay FIFO (dynamics.py:323-328)
            if action.is_contiguous():
                push = action                                 # the caller's tensor: cloned by the launch
            else:
                fifo.append(action.contiguous())
            delayed = fifo.pop(0)
        else:
            delayed = action if action.is_contiguous() else action.contiguous()
        if dyn._wind_fn is not None:
            dyn.update_wind()
        state_in, status_in, peer = dyn._state, fz.status, fz.peer_next
        state_out, status, obs, reward, done, record, term, copy, gate = (fz._stepper or fz._make_stepper()).step(
            state_in, delayed, status_in, fz.global_step, 0, self.keep_terminal_observation, 0, dyn._wind_rows, push, peer)
        if peer:
            fz.peer_next, fz.peer_done = 0, True
        fz.global_step += 1
        if push is not None:
            fifo.append(copy)
        fz.status, fz.record, fz.gate_obs = status, record, gate
        if fz.t_off is not None:
            fz.t_off = th.where(done, 0.0, fz.t_off)
        dyn._prev = (state_in, delayed, None, status_in)
        dyn._state = state_out
        dyn._obs_t = obs if fz.obs_kind == 0 else None       # params.OBS_STATE13
        dyn._n_steps += 1
        dyn._ext = dyn._thrusts_given = None
        dyn._fresh = done
        envs._collision_stale = True
        self._reward, self._done = reward, done
        self._obs_tensors = self._observations = out = self._fused_obs(obs)
        if fz.task != 2:                                      # params.TASK_RACING
            info = _RecordInfo(self.num_agent, record, term, dyn.ctrl_dt, False, self._fused_obs)
        else:
            info = _RecordInfo(self.num_agent, record, term, dyn.ctrl_dt, True, lambda t: self._fused_obs(t, gate))
        self._info = info
        return out, reward, done, info

    def _step_fused_host(self, host_action):
        """numpy in / numpy out through the one-kernel path, one step ahead of the caller (see step_host)."""
        global _RecordInfo, _wait_flag
        if _RecordInfo is None:
            from .fused import RecordInfo as _RecordInfo
            from ... import _lib
            _wait_flag = _lib.fast().wait_flag
        fz = self._fused

        def stage():
            # with a FIFO of depth 1 the step launched in this call consumes this very action: stage it on the main
            # stream; deeper FIFOs take the side stream (the copy is joined below, long before it is consumed)
            self._action = self._stage_action(host_action, side_stream=fz.spec.fifo_depth > 1)
            return self._action
        obs, reward, done, record, term, slot = fz.step_host(stage)
        self._obs_tensors = self._fused_obs(obs)
        if fz.task != 2:                         # params.TASK_RACING
            info = _RecordInfo(self.num_agent, record, term, self.envs.dynamics.ctrl_dt, False, self._fused_obs)
        else:
            gate = fz.gate_obs
            info = _RecordInfo(self.num_agent, record, term, self.envs.dynamics.ctrl_dt, True,
                               lambda t: self._fused_obs(t, gate))
        self._info = info
        evt = self.__dict__.pop("_h2d_evt", None)
        if evt is not None:
            evt.synchronize()                        # join the side-stream copy of this call's action
        np_obs, np_reward, np_done = slot["np"]
        self._observations = self._fused_np_obs(np_obs)
        return self._observations, np_reward, np_done, info

    def _step_fused(self):
        global _RecordInfo, _wait_flag
        if _RecordInfo is None:                  # resolved once (fused.py imports this module's siblings)
            from .fused import RecordInfo as _RecordInfo
            from ... import _lib
            _wait_flag = _lib.fast().wait_flag
        if self.requires_grad and not self.tensor_output:
            raise ValueError("requires_grad should be False if tensor_output is False")
        fz = self._fused
        slot = None if self.tensor_output else fz.host_slot()
        obs, reward, done, record, term = fz.step(self._action, self._action_owned, self.requires_grad,
                                                  None if slot is None else slot["ref"])
        self._obs_tensors = self._fused_obs(obs)
        # the terminal-observation dict is built only if somebody reads a finished agent's info
        if fz.task != 2:                         # params.TASK_RACING
            info = _RecordInfo(self.num_agent, record, term, self.envs.dynamics.ctrl_dt, False, self._fused_obs)
        else:                                    # racing: the dict also carries this step's gate tensor
            gate = fz.gate_obs
            info = _RecordInfo(self.num_agent, record, term, self.envs.dynamics.ctrl_dt, True,
                               lambda t: self._fused_obs(t, gate))
        self._info = info
        if self.tensor_output:                   # kernel outputs never carry autograd history: nothing to detach
            self._observations = self._obs_tensors
            return self._obs_tensors, reward, done, info
        # numpy mode without a FIFO to run ahead in (reference droneGymEnv.py:218): the kernel has already written
        # obs / reward / done into the page-locked host slot (zero-copy stores over PCIe); its last thread block
        # raises the slot's completion word, which the host spins on (no driver call)
        _wait_flag(slot["flag"].data_ptr(), slot["expect"])
        np_obs, np_reward, np_done = slot["np"]
        self._observations = self._fused_np_obs(np_obs)
        return self._observations, np_reward, np_done, info

    def _snapshot_info(self) -> LazyInfo:
        return LazyInfo(self.num_agent, self._done, self._episode_done, self._success, self._step_count,
                        self._rewards, self.envs.once_collided, self._obs_tensors,
                        self.max_episode_steps, self.envs.dynamics.ctrl_dt, self._indiv_rewards,
                        extra_fn=self._extra_info)

    def _extra_info(self, indice: int, info: dict):
        """Hook for task envs that add per-episode extras (reference RacingEnv.collect_info)."""

    def collect_info(self, indice, observations=None):
        """Info record of one agent for the step just taken (reference droneGymEnv.py:238-275)."""
        return self._snapshot_info()[indice]

    def _auto_reset(self, done: th.Tensor):
        """Re-initialise finished agents in place of the reference's ``examine()`` (droneGymEnv.py:207-208,
        :420-423) — same effect, expressed with a mask so that nothing leaves the device."""
        self._on_reset_where(done)
        self.envs.reset_agents_where(done)
        self.get_full_observation()
        self._reset_attr_where(done)

    def _on_reset_where(self, mask: th.Tensor):
        """Hook: task-specific per-agent state to re-initialise (called before the dynamics reset)."""

    def _format_step_output(self, reward, done, info):
        if self.requires_grad:
            if not self.tensor_output:
                raise ValueError("requires_grad should be False if tensor_output is False")
            self._observations = self._obs_tensors
            return self._observations, reward, done, info
        if self.tensor_output:
            self._observations = self._obs_tensors
            return self._observations.detach(), reward.detach(), done, info
        # numpy mode (reference droneGymEnv.py:218): everything goes to page-locked host buffers with
        # asynchronous copies and ONE stream synchronisation, instead of one blocking .cpu() per tensor
        host = self._host_buffers(self._obs_tensors, reward, done)
        for k, v in self._obs_tensors.items():
            host["obs"][k].copy_(v.detach(), non_blocking=True)
        host["reward"].copy_(reward.detach(), non_blocking=True)
        host["done"].copy_(done, non_blocking=True)
        th.cuda.current_stream(self.device).synchronize()
        self._observations = TensorDict({k: v.numpy() for k, v in host["obs"].items()})
        return self._observations, host["reward"].numpy(), host["done"].numpy().astype(np.int32), info

    def _host_buffers(self, obs, reward, done):
        """Two alternating sets of pinned host buffers (the arrays handed out last step stay valid for one step)."""
        ring = getattr(self, "_host_ring", None)
        if ring is None or any(k not in ring[0]["obs"] or ring[0]["obs"][k].shape != v.shape for k, v in obs.items()):
            pin = lambda t: th.empty(t.shape, dtype=t.dtype, pin_memory=True)
            ring = [{"obs": {k: pin(v) for k, v in obs.items()}, "reward": pin(reward), "done": pin(done)}
                    for _ in range(2)]
            self._host_ring, self._host_turn = ring, 0
        self._host_turn ^= 1
        return ring[self._host_turn]

    def _stage_action(self, action, side_stream: bool = False) -> th.Tensor:
        """Host actions (numpy / CPU tensors) reach the device through a pinned staging buffer."""
        if isinstance(action, th.Tensor) and action.is_cuda:
            if action.dtype is th.float32 and action.device == self.device:
                return action                                   # the common case: nothing to convert
            return action.to(self.device, dtype=th.float32)
        src = action if isinstance(action, th.Tensor) else th.as_tensor(np.asarray(action))
        if not self.tensor_output and src.dtype == th.float32 and src.is_contiguous() and src.is_pinned():
            # numpy mode ends every step with a stream synchronisation, so the DMA engine may read the caller's
            # page-locked array directly (no host-side staging copy)
            if not side_stream:
                return src.to(self.device, non_blocking=True)
            main = th.cuda.current_stream(self.device)
            cs = getattr(self, "_copy_stream", None)
            if cs is None:
                cs = self._copy_stream = th.cuda.Stream(self.device)
            with th.cuda.stream(cs):                  # destination comes from the side stream's own pool
                dev = src.to(self.device, non_blocking=True)
                self._h2d_evt = cs.record_event()
            dev.record_stream(main)                   # consumed by main-stream kernels a few steps later
            return dev
        pin = getattr(self, "_act_pin", None)
        if pin is None or pin.shape != src.shape:
            pin = self._act_pin = th.empty(src.shape, dtype=th.float32, pin_memory=True)
        evt = getattr(self, "_act_evt", None)
        if evt is not None:
            evt.synchronize()                 # the previous step's H2D copy must have left the staging buffer
        pin.copy_(src)
        dev = pin.to(self.device, non_blocking=True)
        self._act_evt = th.cuda.Event()
        self._act_evt.record(th.cuda.current_stream(self.device))
        return dev

    _TRANSIENT = ("_host_ring", "_host_turn", "_act_pin", "_act_evt", "_np_const", "_copy_stream", "_h2d_evt",
                  "_task_graph", "_split_warm")

    def __deepcopy__(self, memo):
        """Deep-copyable like the reference env (utils/algorithms/shac.py:121); host staging buffers and CUDA
        events are per-object scratch and are re-created lazily by the copy."""
        import copy
        if self.__dict__.get("_fused") is not None:
            self._fused.rewind()          # a step launched ahead of the caller (host mode) is not part of the state
        twin = self.__class__.__new__(self.__class__)
        memo[id(self)] = twin
        for k, v in self.__dict__.items():
            if k not in self._TRANSIENT:
                setattr(twin, k, copy.deepcopy(v, memo))
        return twin

    def _format_obs(self, obs):
        if not self.tensor_output:
            return TensorDict({k: v.detach().cpu().numpy() for k, v in obs.items()})
        return obs

    # -- reset ---------------------------------------------------------------------------------------------
    def reset(self, state=None, predicted_obs=None, is_test=False, stoch=None, deter=None):
        self._is_initial = True
        for f in (self._fused, self._split):
            if f is not None:
                f.active = False                  # a full reset re-initialises everything the fused path owns
        self.envs._fused = None
        with self._grad_ctx():
            self.envs.reset(state=state)
            self._on_reset_where(th.ones(self.num_agent, dtype=th.bool, device=self.device))
            self._reset_attr()
            self.get_full_observation()
            probe = self.get_reward()
            if isinstance(probe, dict):
                self._indiv_reward = {k: th.zeros((self.num_agent,), device=self.device) for k in probe}
                self._indiv_rewards = {k: th.zeros((self.num_agent,), device=self.device) for k in probe}
            elif isinstance(probe, th.Tensor):
                self._indiv_reward = self._indiv_rewards = None
            else:
                raise ValueError(f"get_reward should return a dict or a tensor, but got {type(probe)}")
        self._fused = self._make_fused()
        self._split = None if self._fused is not None else self._make_split()
        self._task_graph, self._split_warm = None, 0
        self._observations = self._format_obs(self._obs_tensors)
        return self._observations

    def reset_agent_by_id(self, agent_indices=None, state=None, reset_obs=None):
        """Index-based reset of selected agents (reference droneGymEnv.py:339-349)."""
        assert not isinstance(agent_indices, bool)
        self.get_state_for_generic_path()
        with self._grad_ctx():
            if agent_indices is None:
                mask = th.ones(self.num_agent, dtype=th.bool, device=self.device)
            else:
                idx = th.as_tensor(agent_indices, device=self.device, dtype=th.int64).reshape(-1)
                mask = th.zeros(self.num_agent, dtype=th.bool, device=self.device).index_fill(0, idx, True)
            self._on_reset_where(mask)
            self.envs.reset_agents(agent_indices, state=state, pos_reset_by_state=True)
            self.get_full_observation()
            self._reset_attr_where(mask)
        self._observations = self._format_obs(self._obs_tensors)
        return self._observations

    def reset_env_by_id(self, scene_indices=None):
        scene_indices = th.arange(self.num_scene) if scene_indices is None else th.atleast_1d(th.as_tensor(scene_indices))
        agents = (scene_indices.unsqueeze(1) * self.num_agent_per_scene + th.arange(self.num_agent_per_scene)).flatten()
        return self.reset_agent_by_id(agents)

    def examine(self):
        self.get_state_for_generic_path()
        self._auto_reset(self._done)
        self._observations = self._format_obs(self._obs_tensors)
        return self._observations

    @th.no_grad()
    def _reset_attr(self, indices=None, reset_latent=True):
        if indices is not None:
            idx = th.as_tensor(indices, device=self.device, dtype=th.int64).reshape(-1)
            return self._reset_attr_where(
                th.zeros(self.num_agent, dtype=th.bool, device=self.device).index_fill(0, idx, True))
        n, dev = self.num_agent, self.device
        self._reward = th.zeros((n,), device=dev)
        self._rewards = th.zeros((n,), device=dev)
        self._done = th.zeros(n, dtype=th.bool, device=dev)
        self._episode_done = th.zeros(n, dtype=th.bool, device=dev)
        self._step_count = th.zeros((n,), dtype=th.int32, device=dev)
        if self._indiv_rewards is not None:
            self._indiv_rewards = {k: th.zeros((n,), device=dev) for k in self._indiv_rewards}
            self._indiv_reward = {k: th.zeros((n,), device=dev) for k in self._indiv_reward}

    @th.no_grad()
    def _reset_attr_where(self, mask: th.Tensor):
        keep = ~mask
        self._reward = th.where(mask, 0.0, self._reward.detach())
        self._rewards = th.where(mask, 0.0, self._rewards)
        self._done = self._done & keep
        self._episode_done = self._episode_done & keep
        self._step_count = self._step_count * keep
        if self._indiv_rewards is not None:
            self._indiv_rewards = {k: th.where(mask, 0.0, v) for k, v in self._indiv_rewards.items()}
            self._indiv_reward = {k: th.where(mask, 0.0, v) for k, v in self._indiv_reward.items()}

    def detach(self):
        self.envs.detach()
        self.simple_detach()

    def get_state_for_generic_path(self):
        """Make sure the per-agent attributes (``_step_count``, ``_rewards``, ...) are the live ones."""
        for f in (self._fused, self._split):
            if f is not None:
                f.leave()

    def simple_detach(self):
        self._rewards = self._rewards.detach()
        self._reward = self._reward.detach()
        self._action = self._action.detach()
        self._obs_tensors = self._obs_tensors.detach()
        if isinstance(self._observations, TensorDict) and self.tensor_output:
            self._observations = self._obs_tensors      # the dict handed out last must not keep the old graph alive

    # -- task interface ---------------------------------------------------------------------------------------
    def get_done(self):
        return th.zeros(self.num_agent, dtype=th.bool, device=self.device)

    def get_success(self) -> th.Tensor:
        return th.zeros(self.num_agent, dtype=th.bool, device=self.device)

    def get_failure(self) -> th.Tensor:
        return th.zeros(self.num_agent, dtype=th.bool, device=self.device)

    def get_reward(self, predicted_obs=None):
        raise NotImplementedError

    def get_observation(self, indices=None, predicted_obs=None) -> TensorDict:
        raise NotImplementedError

    def get_full_observation(self, indice=None, predicted_obs=None):
        obs = self.get_observation(predicted_obs=predicted_obs)
        assert isinstance(obs, TensorDict)
        self._obs_tensors = obs.as_tensor(device=self.device)
        return self._obs_tensors

    # -- misc VecEnv surface -------------------------------------------------------------------------------------
    def close(self):
        self.envs.close()

    def render(self, **kwargs):
        return self.envs.render(**kwargs)

    def env_is_wrapped(self, wrapper_class=None, indices=None):
        return False

    def step_async(self, actions=None):
        raise NotImplementedError("This method is not implemented")

    def step_wait(self):
        raise NotImplementedError("This method is not implemented")

    def get_attr(self, attr_name, indices=None):
        if indices is None:
            return getattr(self, attr_name)

    def set_attr(self, attr_name, value, indices=None):
        raise NotImplementedError("This method is not implemented")

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        raise NotImplementedError("This method is not implemented")

    def to(self, device):
        self.device = device if not isinstance(device, str) else th.device(device)

    def eval(self):
        self.envs.eval()

    def set_requires_grad(self, requires_grad: bool):
        self.requires_grad = requires_grad

    def __len__(self):
        return self.num_envs

    def __repr__(self):
        return (f"{self.__class__.__name__}(Env={self.envs.__class__}, NumAgentPerScene={self.num_agent_per_scene}, "
                f"NumScene={self.num_scene}, tensorOut={self.tensor_output}, RequiresGrad={self.requires_grad})")

    # -- per-agent bookkeeping that the one-kernel path keeps in its status / episode records ------------------------
    # While the fused path is active these read straight from the kernel's records (no per-step copies); otherwise
    # they are the plain tensors the generic path maintains.
    def _fused_live(self):
        f = self.__dict__.get("_fused") or self.__dict__.get("_split")
        return f if (f is not None and f.active) else None

    def _record_bit(self, bit, fallback):
        f = self._fused_live()
        if f is None or f.record is None:
            return fallback
        return (f.record[:, 2].to(th.int32) & bit) != 0

    @property
    def _step_count(self):
        f = self._fused_live()
        return f.sc if f is not None else self.__dict__["_step_count_v"]

    @_step_count.setter
    def _step_count(self, v):
        self.__dict__["_step_count_v"] = v

    @property
    def _rewards(self):
        f = self._fused_live()
        return f.ret if f is not None else self.__dict__["_rewards_v"]

    @_rewards.setter
    def _rewards(self, v):
        self.__dict__["_rewards_v"] = v

    @property
    def episode_done(self):
        f = self._fused_live()
        return (f.eb & 1) != 0 if f is not None else self._episode_done

    @property
    def success(self):
        return self._record_bit(4, self._success)               # params.RBIT_SUCCESS

    @property
    def failure(self):
        f = self._fused_live()
        # built-in tasks have no failure condition of their own (get_failure is all-False, e.g. HoverEnv.py:79-80)
        return th.zeros_like(self._failure) if f is not None else self._failure

    reward = property(lambda s: s._reward)
    sensor_obs = property(lambda s: s.envs.sensor_obs)
    state = property(lambda s: s.envs.state)
    info = property(lambda s: s._info)
    is_collision = property(lambda s: s.envs.is_collision)
    is_out_bounds = property(lambda s: s.envs.is_out_bounds)
    done = property(lambda s: s._done)
    direction = property(lambda s: s.envs.direction)
    position = property(lambda s: s.envs.position)
    orientation = property(lambda s: s.envs.orientation)
    velocity = property(lambda s: s.envs.velocity)
    angular_velocity = property(lambda s: s.envs.angular_velocity)
    t = property(lambda s: s.envs.t)
    visual = property(lambda s: s.envs.visual)
    collision_vector = property(lambda s: s.envs.collision_vector)
    collision_dis = property(lambda s: s.envs.collision_dis)
    collision_point = property(lambda s: s.envs.collision_point)
    full_state = property(lambda s: s.envs.full_state)
    extend_state = property(lambda s: s.envs.extend_state)
    dynamic_object_position = property(lambda s: s.envs.dynamic_object_position)
