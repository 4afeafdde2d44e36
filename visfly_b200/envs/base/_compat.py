"""Optional third-party surface: gymnasium ``spaces`` and stable-baselines3 ``VecEnv``.

The reference's wrapper subclasses SB3's ``VecEnv`` and describes itself with gymnasium spaces
(envs/base/droneGymEnv.py:4,7,78-115).  Both are used when importable; otherwise minimal stand-ins with the
attributes the reference's algorithms read (``shape``, ``low``, ``high``, ``dtype``, ``spaces``, item access) keep
the env usable — the dynamics path itself depends on neither.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium import spaces  # type: ignore
except Exception:  # noqa: BLE001
    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class _Dict:
        def __init__(self, spaces=None):
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def __setitem__(self, k, v):
            self.spaces[k] = v

        def __contains__(self, k):
            return k in self.spaces

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return f"Dict({self.spaces})"

    class spaces:  # noqa: N801 - mirrors the module name
        Box = _Box
        Dict = _Dict

try:  # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv  # type: ignore
except Exception:  # noqa: BLE001
    class VecEnv:  # noqa: D401 - stand-in base class
        """Stand-in for ``stable_baselines3.common.vec_env.VecEnv`` (only used as a base class)."""
