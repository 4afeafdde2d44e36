"""Import the real VisFly reference (``/root/reference``) for oracle validation and golden generation.

Only usable where the reference tree is mounted (the build container).  Nothing that runs on the GPU box
may depend on it: callers must check ``reference_available()`` and skip otherwise.

The reference is imported unmodified; the three RK4 repairs frozen in SURVEY.md §8c are applied as
runtime monkeypatches (the reference's ``integrator="rk4"`` raises ``TypeError`` as shipped):
  R1  pass ``wind`` to every RK4 stage                    (utils/maths.py:370-379 vs :300-309)
  R2  stage buffers on the state's device/dtype          (utils/maths.py:354-361)
  R3  return the weighted stage mean ``d_ori_vel @ ks``  (utils/maths.py:386)
"""
from __future__ import annotations

import os
import sys
import tempfile

import torch as th

REFERENCE_ROOT = os.environ.get("VISFLY_REFERENCE", "/root/reference")
_state = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "base", "dynamics.py"))


def load_reference():
    """Returns the reference's ``envs.base.dynamics`` module (package name ``VisFly``), repaired."""
    if "mod" in _state:
        return _state["mod"]
    root = tempfile.mkdtemp(prefix="visfly_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(root, "VisFly"))
    sys.path.insert(0, root)
    from VisFly.envs.base import dynamics as dynmod          # noqa
    from VisFly.utils import maths                           # noqa

    orig_integrate = maths.Integrator.integrate
    orig_derivs = maths.Integrator._get_derivatives
    cur = {}

    def derivs(vel, ori, acc, ori_vel, tau, J, J_inv, wind=None):
        return orig_derivs(vel, ori, acc, ori_vel, tau, J, J_inv, cur["wind"] if wind is None else wind)   # R1

    def integrate(**kw):
        cur["wind"] = kw.get("wind")
        if kw.get("type") != "rk4":
            return orig_integrate(**kw)
        # R2: the reference allocates its stage buffers with the default dtype on the CPU; run it under the
        # state's dtype so float64 oracles work.
        prev = th.get_default_dtype()
        th.set_default_dtype(kw["pos"].dtype)
        try:
            out = orig_integrate(**kw)
        finally:
            th.set_default_dtype(prev)
        ks = th.tensor([1., 2., 2., 1.], dtype=kw["pos"].dtype) / 6
        return (*out[:4], out[4] @ ks)                                                                  # R3

    maths.Integrator._get_derivatives = staticmethod(derivs)
    maths.Integrator.integrate = staticmethod(integrate)
    _state["mod"] = dynmod
    return dynmod


class default_dtype:
    """The reference allocates with torch's default dtype everywhere (reset, FIFO, RK4 buffers): run float64
    reference sessions entirely inside this context."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        self.prev = th.get_default_dtype()
        th.set_default_dtype(self.dtype)

    def __exit__(self, *exc):
        th.set_default_dtype(self.prev)


def make_reference_dynamics(num, dtype=th.float32, **kw):
    """Construct a reference ``Dynamics``; for float64 the module constants are rebuilt in that dtype."""
    dynmod = load_reference()
    prev = th.get_default_dtype()
    th.set_default_dtype(dtype)
    try:
        dynmod.g = th.tensor([[0, 0, -9.81]]).T
        dynmod.z = th.tensor([[0, 0, 1.0]]).T
        d = dynmod.Dynamics(num=num, **kw)
    finally:
        th.set_default_dtype(prev)
    return d


# ---------------------------------------------------------------------------------------------------
# env-level reference: the real wrapper + task envs with their absent third-party imports stubbed
# ---------------------------------------------------------------------------------------------------
def load_reference_envs():
    """Import the reference's ``HoverEnv`` / ``NavigationEnv`` / ``RacingEnv2`` (``visual=False``).

    habitat_sim, stable_baselines3, gymnasium, the Habitat ``SceneManager`` / ``ObjectManager`` and the
    depth auto-encoder module are not installed here and not on the dynamics path; they are replaced by
    inert ``sys.modules`` stubs *before* the reference modules are imported.
    The reference's signature drift is bridged by thin subclasses, nothing else is touched.
    """
    if "envs" in _state:
        return _state["envs"]
    import types

    load_reference()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class SensorType:
        DEPTH, COLOR, SEMANTIC = 1, 2, 3

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(dict):
        def __init__(self, spaces=None):
            super().__init__(spaces or {})
            self.spaces = self

    class VecEnv:
        pass

    class SceneManager:
        def __init__(self, num_agent_per_scene=1, num_scene=1, sensor_settings=None, **kw):
            self.num_scene, self.num_agent_per_scene = num_scene, num_agent_per_scene
            self.num_agent = num_scene * num_agent_per_scene
            self.col_refine_steps = 0
            self.scenes = [None]
            self.sensor_settings = sensor_settings or []
            self.dynamic_object_position = [[None] for _ in range(self.num_agent)]
            self.dynamic_object_velocity = [[None] for _ in range(self.num_agent)]
            self.dynamic_object_acceleration = [[None] for _ in range(self.num_agent)]

        def close(self):
            pass

    hs = mod("habitat_sim", SensorType=SensorType)
    hs.sensor = mod("habitat_sim.sensor", SensorType=SensorType)
    mod("stable_baselines3")
    mod("stable_baselines3.common")
    mod("stable_baselines3.common.vec_env", VecEnv=VecEnv)
    gym = mod("gymnasium")
    gym.spaces = mod("gymnasium.spaces", Box=Box, Dict=Dict)
    mod("VisFly.utils.SceneManager", SceneManager=SceneManager)
    mod("VisFly.utils.ObjectManger", ObjectManager=object)
    mod("VisFly.utils.tools")
    mod("VisFly.utils.tools.train_encoder", model=None)

    from VisFly.envs.HoverEnv import HoverEnv as _Hover          # noqa
    from VisFly.envs.NavigationEnv import NavigationEnv          # noqa
    from VisFly.envs.RacingEnv import RacingEnv2 as _Racing2     # noqa

    class HoverEnv(_Hover):
        def get_reward(self, predicted_obs=None):
            return super().get_reward()

    class RacingEnv2(_Racing2):
        latent = None

        def get_observation(self, indices=None, predicted_obs=None):
            return super().get_observation(indices)

        def get_reward(self, predicted_obs=None):
            return super().get_reward()

    _state["envs"] = {"HoverEnv": HoverEnv, "NavigationEnv": NavigationEnv, "RacingEnv2": RacingEnv2}
    return _state["envs"]
