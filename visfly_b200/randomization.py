"""Initial-state generators (reference utils/randomization.py), vectorised.

The reference keeps one generator object per agent and calls ``safe_generate(num=1)`` in a Python loop
(envs/base/droneEnv.py:243-249; 17 s for 65 536 agents).  With ``visual=False`` every agent shares the same
generator (droneEnv.py:221-230), so drawing all requested agents in one call samples the same distribution.
Everything is generated on the engine's device; nothing here synchronises with the host.

Same classes / kwargs as the reference: ``Uniform`` (:106-170), ``Normal`` (:173-206), ``Union`` (:249-296);
``state_generator = {"class": ..., "kwargs": [...]}`` is resolved by ``load_generator`` (:299-310).
Rejection sampling against scene geometry (``safe_generate`` with ``is_collision_func``, :64-96) only exists for
Habitat scenes and is out of scope with rendering.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch as th

_ZERO3 = {"mean": [0.0, 0.0, 0.0], "half": [0.0, 0.0, 0.0]}
_ZERO3N = {"mean": [0.0, 0.0, 0.0], "std": [0.0, 0.0, 0.0]}


def euler_zyx_to_quat(e: th.Tensor) -> th.Tensor:
    """(n,3) roll,pitch,yaw -> (n,4) w,x,y,z  (reference maths.py:257-269, order zyx)."""
    h = e * 0.5
    cr, sr, cp, sp, cy, sy = h[:, 0].cos(), h[:, 0].sin(), h[:, 1].cos(), h[:, 1].sin(), h[:, 2].cos(), h[:, 2].sin()
    return th.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)


class StateRandomizer:
    def __init__(self, device="cuda", is_collision_func=None, scene_id=None, seed: int = 42, **unknown):
        if unknown:             # a typo in a generator kwarg must not silently change the start distribution
            raise TypeError(f"{type(self).__name__}: unknown keyword argument(s) {sorted(unknown)}")
        self.device = th.device(device)
        self.is_collision_func, self.scene_id = is_collision_func, scene_id

    def to(self, device):
        self.device = th.device(device)
        return self

    def set_seed(self, seed=42):
        th.manual_seed(seed)

    def _generate(self, num: int, **kw) -> Tuple[th.Tensor, th.Tensor, th.Tensor, th.Tensor]:
        raise NotImplementedError

    def generate(self, num: int, **kw):
        """(position, euler orientation, velocity, angular velocity), each (num,3)."""
        return self._generate(num, **kw)

    def safe_generate(self, num: int = 1, **kw):
        """(position (num,3), quaternion (num,4), velocity, angular velocity) on ``self.device``."""
        if self.is_collision_func is not None:
            raise NotImplementedError("rejection sampling against scene geometry needs the renderer (out of scope)")
        pos, eul, vel, rate = self.generate(num, **kw)
        return pos, euler_zyx_to_quat(eul), vel, rate


class UniformStateRandomizer(StateRandomizer):
    def __init__(self, position=_ZERO3, orientation=_ZERO3, velocity=_ZERO3, angular_velocity=_ZERO3,
                 heading=False, test: bool = False, xyz_num=(1, 1, 1), xyz_half=(0, 2, 0.), **kw):
        super().__init__(**kw)
        #: reference :141-151 — evaluation mode: start positions walk over a regular grid of the position box
        #: (``xyz_num`` points per axis, `indexing="ij"`), one grid point per generated agent, plus U(-1,1)*xyz_half
        self.test = bool(test)
        if self.test:
            axes = [th.linspace(-1, 1, int(k)) for k in xyz_num]
            gx, gy, gz = th.meshgrid(*axes, indexing="ij")
            self._grid = th.stack([gx.flatten(), gy.flatten(), gz.flatten()], dim=1)
            self._grid_index = 0
            self._xyz_half = th.atleast_2d(th.tensor(list(xyz_half), dtype=th.float32))
        # one (4,3) mean and half-width table: a single rand call covers all four fields
        fields = (position, orientation, velocity, angular_velocity)
        self._mean = th.tensor([list(f["mean"]) for f in fields], dtype=th.float32)
        self._half = th.tensor([list(f["half"]) for f in fields], dtype=th.float32)
        self._deterministic = bool((self._half == 0).all())
        #: reference :162-165 — the initial yaw points from the drawn position back to the centre of the position
        #: box (roll = pitch = 0), plus the orientation noise; the orientation mean is not used
        self.heading = bool(heading)
        if self.heading and bool((self._half[0, :2] == 0).all()):
            raise ValueError("heading=True needs a position box with a non-zero x or y half-width (the reference "
                             "divides by the horizontal offset from the box centre, randomization.py:29)")

    def to(self, device):
        super().to(device)
        self._mean, self._half = self._mean.to(device), self._half.to(device)
        if self.test:
            self._grid, self._xyz_half = self._grid.to(device), self._xyz_half.to(device)
        return self

    def _generate(self, num, **kw):
        if self._mean.device != self.device:
            self.to(self.device)
        if self._deterministic and not self.test:
            s = self._mean.unsqueeze(0).expand(num, 4, 3)
        else:
            s = (2 * th.rand((num, 4, 3), device=self.device) - 1) * self._half + self._mean     # :153-169
        if self.test:
            # the reference generates agent by agent and advances the grid index once per call (:156-160): agent k of
            # this batch gets grid point (index + k) mod len(grid)
            k = (self._grid_index + th.arange(num, device=self.device)) % self._grid.shape[0]
            self._grid_index += num
            pos = self._grid[k] * self._half[0] + self._mean[0] + \
                (2 * th.rand((num, 3), device=self.device) - 1) * self._xyz_half
            s = th.cat([pos.unsqueeze(1), s[:, 1:]], dim=1)
        if self.heading:
            d = self._mean[0] - s[:, 0]                                   # direction = -half  (:163)
            yaw = th.arccos(d[:, 0] / d[:, :2].norm(dim=1)) * th.where(d[:, 1].sign() >= 0, 1, -1)   # :27-28
            ori = s[:, 1] - self._mean[1]                                 # the noise term of :165
            ori = th.stack([ori[:, 0], ori[:, 1], ori[:, 2] + yaw], dim=1)
            return s[:, 0], ori, s[:, 2], s[:, 3]
        return s[:, 0], s[:, 1], s[:, 2], s[:, 3]


class NormalStateRandomizer(StateRandomizer):
    def __init__(self, position=_ZERO3N, orientation=_ZERO3N, velocity=_ZERO3N, angular_velocity=_ZERO3N, **kw):
        super().__init__(**kw)
        fields = (position, orientation, velocity, angular_velocity)
        self._mean = th.tensor([list(f["mean"]) for f in fields], dtype=th.float32)
        self._std = th.tensor([list(f["std"]) for f in fields], dtype=th.float32)

    def to(self, device):
        super().to(device)
        self._mean, self._std = self._mean.to(device), self._std.to(device)
        return self

    def _generate(self, num, **kw):
        if self._mean.device != self.device:
            self.to(self.device)
        # the reference draws (2*randn - 1) * std + mean (:201-204); kept, it is the behaviour
        s = (2 * th.randn((num, 4, 3), device=self.device) - 1) * self._std + self._mean
        return s[:, 0], s[:, 1], s[:, 2], s[:, 3]


class UnionRandomizer(StateRandomizer):
    """Each agent picks one of several generators uniformly at random (:249-296)."""

    Randomizer_alias = {"Uniform": UniformStateRandomizer, "Normal": NormalStateRandomizer}

    def __init__(self, randomizers_kwargs: List[Dict], **kw):
        super().__init__(**kw)
        self.randomizers = [self.Randomizer_alias[r["class"]](**{**r["kwargs"], **kw}) for r in randomizers_kwargs]

    def to(self, device):
        super().to(device)
        for r in self.randomizers:
            r.to(device)
        return self

    def __len__(self):
        return len(self.randomizers)

    def _generate(self, num, **kw):
        draws = [r.generate(num) for r in self.randomizers]
        pick = th.randint(0, len(self.randomizers), (num,), device=self.device)
        row = th.arange(num, device=self.device)
        return tuple(th.stack([d[f] for d in draws])[pick, row] for f in range(4))


_CLS = {"Uniform": UniformStateRandomizer, "Normal": NormalStateRandomizer, "Union": UnionRandomizer}


def load_generator(cls, kwargs, is_collision_func=None, scene_id=None, device="cuda"):
    if isinstance(cls, str):
        if cls not in _CLS:
            raise NotImplementedError(f"state generator '{cls}' is not available without the renderer")
        cls = _CLS[cls]
    return cls(is_collision_func=is_collision_func, scene_id=scene_id, device=device, **kwargs)
