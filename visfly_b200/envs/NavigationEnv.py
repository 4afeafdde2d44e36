"""``NavigationEnv`` — fly to a target inside the bounding box (reference envs/NavigationEnv.py:27-99)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch as th

from ..type import TensorDict
from .base._compat import spaces
from .base.droneGymEnv import DroneGymEnvsBase


class NavigationEnv(DroneGymEnvsBase):
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
            **kwargs,
    ):
        super().__init__(num_agent_per_scene=num_agent_per_scene, num_scene=num_scene, seed=seed, visual=visual,
                         requires_grad=requires_grad, random_kwargs=random_kwargs, dynamics_kwargs=dynamics_kwargs,
                         scene_kwargs=scene_kwargs, sensor_kwargs=sensor_kwargs, device=device,
                         max_episode_steps=max_episode_steps, **kwargs)
        tgt = th.as_tensor([9, 0., 1] if target is None else target, dtype=th.float32).reshape(1, -1)
        self.target = (th.ones((self.num_envs, 1)) @ tgt).to(self.device)
        self.observation_space["target"] = spaces.Box(low=-np.inf, high=np.inf, shape=(3,), dtype=np.float32)
        self.success_radius = 0.5
        self._unit_quat = th.tensor([1., 0, 0, 0], device=self.device)

    def get_observation(self, indices=None, predicted_obs=None) -> Dict:
        return TensorDict({"state": self.state, "target": self.target})

    def get_success(self) -> th.Tensor:
        return (self.position - self.target).norm(dim=1) <= self.success_radius

    def _make_fused(self):
        from .. import params as P
        from .base.fused import FusedEnvStep
        if not self._builtin_task(NavigationEnv):
            return None
        return FusedEnvStep(self, P.TASK_NAVIGATION, P.OBS_STATE13)   # target / radius: read from the live attributes

    def _fused_obs(self, obs):
        return TensorDict({"state": obs, "target": self.target})

    def get_reward(self, predicted_obs=None) -> th.Tensor:
        """reference NavigationEnv.py:85-99, term by term."""
        base_r = 0.1
        thrd_perce = th.pi / 18
        to_target = self.target - self.position
        v = self.velocity
        approach = ((v * to_target).sum(dim=1) / (1e-6 + to_target.norm(dim=1))).clamp_max(10) * 0.01
        heading = (((self.direction * v).sum(dim=1) / (1e-6 + v.norm(dim=1)) / 1).clamp(-1., 1.).acos()
                   .clamp_min(thrd_perce) - thrd_perce) * -0.01
        stable = ((self.orientation - self._unit_quat).norm(dim=1) * -0.00001
                  + (v - 0).norm(dim=1) * -0.002 + (self.angular_velocity - 0).norm(dim=1) * -0.002)
        near = 1 / (self.collision_dis + 0.2) * -0.01
        closing = ((1 - self.collision_dis).relu()
                   * ((self.collision_vector * (v - 0)).sum(dim=1) / (1e-6 + self.collision_dis)).relu() * -0.005)
        bonus = (self._success * (self.max_episode_steps - self._step_count) * base_r
                 * (0.2 + 0.8 / (1 + 1 * v.norm(dim=1))))
        return base_r * 0 + approach + heading + stable + near + closing + bonus
