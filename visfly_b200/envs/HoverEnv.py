"""``HoverEnv`` — hold position at a target (reference envs/HoverEnv.py:14-94)."""
from __future__ import annotations

from typing import Dict, Optional

import torch as th

from ..type import TensorDict
from .base.droneGymEnv import DroneGymEnvsBase


class HoverEnv(DroneGymEnvsBase):
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
            tensor_output: bool = False,
            **kwargs,
    ):
        if random_kwargs is None:      # reference HoverEnv.py:32-41
            random_kwargs = {"state_generator": {"class": "Uniform", "kwargs": [
                {"position": {"mean": [1., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}}
        super().__init__(num_agent_per_scene=num_agent_per_scene, num_scene=num_scene, seed=seed, visual=visual,
                         requires_grad=requires_grad, random_kwargs=random_kwargs, dynamics_kwargs=dynamics_kwargs,
                         sensor_kwargs=sensor_kwargs, scene_kwargs=scene_kwargs, device=device,
                         max_episode_steps=max_episode_steps, tensor_output=tensor_output, **kwargs)
        tgt = th.as_tensor([1, 0., 1.5] if target is None else target, dtype=th.float32).reshape(1, -1)
        self.target = (th.ones((self.num_envs, 1)) @ tgt).to(self.device)
        self.success_radius = 0.5
        self._unit_quat = th.tensor([1., 0, 0, 0], device=self.device)

    def get_observation(self, indices=None, predicted_obs=None) -> Dict:
        return TensorDict({"state": self.state})

    def _make_fused(self):
        from .. import params as P
        from .base.fused import FusedEnvStep
        if not self._builtin_task(HoverEnv):
            return None
        return FusedEnvStep(self, P.TASK_HOVER, P.OBS_STATE13)     # target etc. are read from the live attributes

    def get_success(self) -> th.Tensor:
        return th.zeros(self.num_agent, dtype=th.bool, device=self.device)      # reference HoverEnv.py:79-80

    def get_reward(self, predicted_obs=None) -> th.Tensor:
        """reference HoverEnv.py:83-94"""
        base_r = 0.1
        pos_factor = -0.1 * 1 / 9
        return (base_r
                + (self.position - self.target).norm(dim=1) * pos_factor
                + (self.orientation - self._unit_quat).norm(dim=1) * -0.00001
                + (self.velocity - 0).norm(dim=1) * -0.002
                + (self.angular_velocity - 0).norm(dim=1) * -0.002)
