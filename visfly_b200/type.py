"""Small value types that appear at the boundary of the dynamics engine.

Same names and behaviour as the reference's ``utils/type.py`` (``bound`` :8-11, ``ACTION_TYPE`` :14-18,
``Uniform`` :21-38, ``Normal`` :41-58, ``PID`` :61-85, ``TensorDict`` :101-193) so that task code written
against VisFly keeps working; the implementations are new.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Any, Iterable, Union

import numpy as np
import torch as th


@dataclass
class bound:
    min: Any
    max: Any


class ACTION_TYPE(Enum):
    THRUST = 0
    BODYRATE = 1
    VELOCITY = 2
    POSITION = 3


class Uniform:
    """``mean`` / ``half`` pair.  ``generate(n)`` draws ``(rand-0.5)*half+mean`` like reference type.py:37-38."""

    def __init__(self, mean=0, half=0):
        self.mean = th.atleast_1d(th.as_tensor(mean))
        self.half = th.atleast_1d(th.as_tensor(half))

    def to(self, device):
        self.mean, self.half = self.mean.to(device), self.half.to(device)
        return self

    def generate(self, size, generator=None):
        r = th.rand(size, len(self.mean), device=self.mean.device, generator=generator)
        return (r - 0.5) * self.half + self.mean

    @property
    def is_zero(self) -> bool:
        return bool((self.half == 0).all()) and bool((self.mean == 0).all())


class Normal:
    def __init__(self, mean=0, std=0):
        self.mean = th.as_tensor(mean)
        self.std = th.as_tensor(std)

    def to(self, device):
        self.mean, self.std = self.mean.to(device), self.std.to(device)
        return self

    def generate(self, size, generator=None):
        return th.normal(self.mean, self.std, size, generator=generator)


@dataclass
class PID:
    p: th.Tensor = None
    i: th.Tensor = None
    d: th.Tensor = None

    def __post_init__(self):
        eye = th.eye(3)
        self.p = eye.clone() if self.p is None else self.p
        self.i = eye.clone() if self.i is None else self.i
        self.d = eye.clone() if self.d is None else self.d

    def to(self, device):
        self.p, self.i, self.d = self.p.to(device), self.i.to(device), self.d.to(device)
        return self

    def clone(self):
        self.p, self.i, self.d = self.p.clone(), self.i.clone(), self.d.clone()
        return self

    def detach(self):
        self.p, self.i, self.d = self.p.detach(), self.i.detach(), self.d.detach()
        return self


class TensorDict(dict):
    """Observation container: a dict of equally long tensors that can be indexed by agent.

    String keys address fields; int / slice / index-tensor keys address agents and return a new
    ``TensorDict`` of (at least 2-d) rows (reference type.py:115-126).
    """

    def __init__(self, data=()):
        super().__init__(data)

    def detach(self):
        return TensorDict({k: v.detach() for k, v in self.items()})

    def clone(self):
        for k in list(self.keys()):
            super().__setitem__(k, self[k].clone())
        return self

    def __getitem__(self, key: Any) -> Any:
        if isinstance(key, str):
            return super().__getitem__(key)
        if isinstance(key, (int, slice)):
            return TensorDict({k: th.atleast_2d(v[key]) for k, v in self.items()})
        if hasattr(key, "__iter__"):
            return TensorDict({k: th.atleast_2d(v[_index_on(key, v)]) for k, v in self.items()})
        raise TypeError("TensorDict keys are field names (str) or agent indices")

    def __setitem__(self, key: Any, value: Any) -> None:
        if isinstance(key, str):
            super().__setitem__(key, value)
        elif isinstance(key, (int, th.Tensor, np.ndarray, list, slice)):
            for k in self.keys():
                super().__getitem__(k)[key] = value[k]
        else:
            raise TypeError("TensorDict keys are field names (str) or agent indices")

    def append(self, data):
        if isinstance(data, TensorDict):
            for k in data.keys():
                super().__setitem__(k, th.cat([self[k], data[k]]))

    def cpu(self):
        for k in list(self.keys()):
            super().__setitem__(k, self[k].cpu())
        return self

    def as_tensor(self, device=th.device("cpu")):
        return TensorDict({k: th.as_tensor(v, device=device) for k, v in self.items()})

    def to(self, device):
        for k in list(self.keys()):
            super().__setitem__(k, self[k].to(device))
        return self

    def reshape(self, shape):
        for k in list(self.keys()):
            super().__setitem__(k, self[k].reshape(shape))
        return self

    @staticmethod
    def stack(items: Iterable["TensorDict"]):
        items = list(items)
        return TensorDict({k: th.stack([it[k] for it in items]) for k in items[0].keys()})

    def numpy(self):
        for k in list(self.keys()):
            super().__setitem__(k, self[k].detach().cpu().numpy())
        return self

    def __len__(self):
        sizes = {len(v) for v in self.values()}
        assert len(sizes) == 1, "fields of a TensorDict must have the same length"
        return sizes.pop()

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def _index_on(key, like):
    """Index tensors must live where the indexed tensor lives (or on the CPU)."""
    if isinstance(key, th.Tensor) and isinstance(like, th.Tensor) and key.device != like.device:
        return key.to(like.device)
    return key


Number = Union[int, float]
