"""``RacingEnv`` / ``RacingEnv2`` — fly through a loop of four gates (reference envs/RacingEnv.py:16-267).

The reference's classes crash at ``reset()`` (signature drift, undefined ``self.latent``); the gate
logic, rewards, observations and initial-state distribution below follow the reference's code, with the per-agent
Python loop of ``_choose_target`` (:175-185) written as tensor arithmetic.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch as th

from ..type import TensorDict
from .base._compat import spaces
from .base.droneGymEnv import DroneGymEnvsBase

is_pos_reward = True


def _start_boxes():
    box = lambda c: {"class": "Uniform", "kwargs": {"position": {"mean": c, "half": [.2, .2, 0.2]}}}
    return {"state_generator": {"class": "Union", "kwargs": [{"randomizers_kwargs": [
        box([2., 2., 1]), box([6., 2., 1.5]), box([6., -2., 1.5]), box([2., 0., 1])]}]}}


class RacingEnv(DroneGymEnvsBase):
    def __init__(
            self,
            num_agent_per_scene: int = 1,
            num_scene: int = 1,
            seed: int = 42,
            visual: bool = False,
            requires_grad: bool = False,
            random_kwargs: dict = None,
            dynamics_kwargs: dict = None,
            scene_kwargs: dict = None,
            sensor_kwargs: list = None,
            device: str = "cuda",
            target: Optional[th.Tensor] = None,
            max_episode_steps: int = 256,
            latent_dim=None,
            **kwargs,
    ):
        super().__init__(num_agent_per_scene=num_agent_per_scene, num_scene=num_scene, seed=seed, visual=visual,
                         requires_grad=requires_grad, random_kwargs=_start_boxes(),       # reference :33-70
                         dynamics_kwargs=dynamics_kwargs, sensor_kwargs=sensor_kwargs, scene_kwargs=scene_kwargs,
                         device=device, max_episode_steps=max_episode_steps, **kwargs)
        dev = self.device
        self.targets = th.as_tensor([[4, 4, 1.], [8, 0, 2.], [5, -4, 1.], [1, -1, 1.]], device=dev)
        self._next_target_num = 2
        self._next_target_i = th.zeros((self.num_envs,), dtype=th.int64, device=dev)
        self._past_targets_num = th.zeros((self.num_envs,), dtype=th.int64, device=dev)
        self._is_pass_next = th.zeros((self.num_envs,), dtype=th.bool, device=dev)
        self.success_radius = 0.3
        self.observation_space["gate"] = spaces.Box(low=0, high=len(self.targets), shape=(1,), dtype=np.int32)
        self.observation_space["state"] = spaces.Box(
            low=-np.inf, high=np.inf,
            shape=(3 * (self._next_target_num - 1) + self.observation_space["state"].shape[0],), dtype=np.float32)
        self.success_r = 5
        self.latent = None
        self._unit_quat = th.tensor([1., 0, 0, 0], device=dev)
        self._gate_ref = th.as_tensor([4., 0, 1], device=dev)

    is_pass_next = property(lambda s: s._is_pass_next)

    # gate bookkeeping: while the one-kernel path is active it lives in the kernel's status records
    @property
    def _next_target_i(self):
        f = self._fused_live()
        return f.gate.to(th.int64) if f is not None else self.__dict__["_next_target_i_v"]

    @_next_target_i.setter
    def _next_target_i(self, v):
        self.__dict__["_next_target_i_v"] = v

    @property
    def _past_targets_num(self):
        f = self._fused_live()
        return f.passed.to(th.int64) if f is not None else self.__dict__["_past_targets_num_v"]

    @_past_targets_num.setter
    def _past_targets_num(self, v):
        self.__dict__["_past_targets_num_v"] = v

    def _extra_info(self, indice, info):
        info["episode"]["extra"]["past_gate"] = int(self._past_targets_num_at_done[indice])

    def _snapshot_info(self):
        self._past_targets_num_at_done = self._past_targets_num
        return super()._snapshot_info()

    def get_observation(self, indices=None, predicted_obs=None) -> Dict:
        return TensorDict({"state": self.state, "gate": self._next_target_i})

    _FUSED_OBS = 0          # params.OBS_STATE13

    def _make_fused(self):
        from .. import params as P
        from .base.fused import FusedEnvStep
        owner = RacingEnv2 if isinstance(self, RacingEnv2) else RacingEnv
        if not is_pos_reward or not self._builtin_task(owner) or type(self)._extra_info is not RacingEnv._extra_info:
            return None
        return FusedEnvStep(self, P.TASK_RACING, self._FUSED_OBS)     # gates / radius: read from the live attributes

    def _fused_obs(self, obs, gate=None):
        # the kernel writes the gate index in the dtype / shape the observation dict carries ((N,) here, (N,1) for v2)
        return TensorDict({"state": obs, "gate": self._fused.gate_obs if gate is None else gate})

    def get_success(self) -> th.Tensor:
        """Gate passing (reference :142-148); never ends the episode."""
        gate = self.targets[self._next_target_i]
        self._is_pass_next = (self.position.detach() - gate).norm(dim=1) <= self.success_radius
        self._next_target_i = (self._next_target_i + self._is_pass_next) % len(self.targets)
        self._past_targets_num = self._past_targets_num + self._is_pass_next
        return th.zeros((self.num_envs,), dtype=th.bool, device=self.device)

    def _choose_target(self, mask: Optional[th.Tensor] = None):
        """First gate from where the agent stands relative to (4,0,1) (reference :173-185, vectorised)."""
        r = self.position.detach() - self._gate_ref
        left, right = r[:, 0] < 0, r[:, 0] > 0
        choice = th.where(left, th.where(r[:, 1] > 0, 0, 3), th.where(right, 1, 2))
        self._next_target_i = choice if mask is None else th.where(mask, choice, self._next_target_i)

    def _on_reset_where(self, mask: th.Tensor):
        self._choose_target(mask)
        self._past_targets_num = self._past_targets_num * ~mask
        self._is_pass_next = self._is_pass_next & ~mask

    def get_reward(self, predicted_obs=None) -> th.Tensor:
        gate = self.targets[self._next_target_i]
        if not is_pos_reward:                                                 # reference :188-201
            dis_vector = gate - self.position
            dis = (dis_vector - 0).norm(dim=1, keepdim=True)
            approaching_v = (((self.velocity - 0) * dis_vector).sum(dim=1, keepdim=True) / (dis + 1e-6)).clamp_max(15.)
            away_v_vector = self.velocity - dis_vector / (dis + 1e-6) * approaching_v
            away_v = (away_v_vector - 0).norm(dim=1) * (1 / (dis.squeeze() + 1))
            reward = approaching_v.squeeze() * 0.02 - away_v * 0.02 + self.is_pass_next * self.success_r
            return reward + (self.angular_velocity - 0).norm(dim=1) * -0.001
        base_r, pos_factor = 0.1, -0.1 * 1 / 9                                 # reference :203-215
        self.success_r = 20
        return (base_r
                + (self.position - gate).norm(dim=1) * pos_factor
                + (self.orientation - self._unit_quat).norm(dim=1) * -0.00001
                + (self.velocity - 0).norm(dim=1) * -0.002
                + (self.angular_velocity - 0).norm(dim=1) * -0.002
                + self.is_pass_next * self.success_r)


class RacingEnv2(RacingEnv):
    _FUSED_OBS = 1          # params.OBS_RACING16

    def get_observation(self, indices=None, predicted_obs=None) -> Dict:
        """Relative positions of the next two gates, attitude, scaled velocities (reference :250-267)."""
        nxt = th.stack([self._next_target_i + i for i in range(self._next_target_num)]).T % len(self.targets)
        relative_pos = (self.targets[nxt] - self.position.unsqueeze(1)).reshape(self.num_envs, -1)
        state = th.hstack([relative_pos / self.max_sense_radius, self.orientation, self.velocity / 10,
                           self.angular_velocity / 10])
        return TensorDict({"state": state, "gate": self._next_target_i.unsqueeze(1).clone().detach()})
