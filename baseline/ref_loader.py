"""Loader of the UNMODIFIED VisFly reference for the reference arm of bench.py and for oracle validation.

The reference is pure Python/PyTorch.  Where its tree is present it is imported as package ``VisFly`` from
  1. ``$VISFLY_REFERENCE`` or ``/root/reference`` (the build container), else
  2. ``baseline/_ref/VisFly`` — a verbatim copy of the few files of the path made by ``baseline/install_ref.py``
     during ``__graft_entry__.build()`` (git-ignored: reference sources never enter this repository's history; it
     travels to the GPU box with the gpurun snapshot like the built .so files do).
Nothing of the product imports this module: only ``bench.py``'s baseline legs and ``tests/``.

The source files are never edited.  What the reference needs to run at all is applied as runtime monkeypatches,
listed in ``baseline/REF_PATCHES.md``:
  R1  pass ``wind`` to every RK4 stage                    (utils/maths.py:370-379 vs :300-309)
  R2  stage buffers on the state's device/dtype          (utils/maths.py:354-361)
  R3  return the weighted stage mean ``d_ori_vel @ ks``  (utils/maths.py:386)
  D1-D3 (CUDA runs only): the reference allocates its wind vector, RK4 stage buffers, FIFO and restart times with
      device-less factory calls; ``reference_on_device`` runs it under ``torch.set_default_device`` so that those
      land on the GPU too.
  C4  signature shims for HoverEnv / RacingEnv2 (``predicted_obs`` kwarg), as thin subclasses.
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile

import torch as th

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("VISFLY_REFERENCE", "/root/reference"), os.path.join(_HERE, "_ref", "VisFly")]
_state = {}


def find_reference_root():
    """Directory holding the reference tree, or None."""
    for root in _CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "envs", "base", "dynamics.py")):
            return root
    return None


REFERENCE_ROOT = find_reference_root()


def reference_available() -> bool:
    return REFERENCE_ROOT is not None


def reference_origin() -> str:
    """"live" (the mounted reference tree) or "copy" (baseline/_ref) — reported next to every reference number."""
    if REFERENCE_ROOT is None:
        return "absent"
    return "copy" if os.path.abspath(REFERENCE_ROOT).startswith(_HERE) else "live"


@contextlib.contextmanager
def reference_on_device(device):
    """D1-D3: run reference code with ``device`` as torch's default device (and restore it afterwards)."""
    prev = th.get_default_device()
    th.set_default_device(device)
    try:
        yield
    finally:
        th.set_default_device(prev)

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
        ks = th.tensor([1., 2., 2., 1.], dtype=kw["pos"].dtype, device=kw["pos"].device) / 6
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


# ---------------------------------------------------------------------------------------------------
# the reference's own analytic-gradient trainers (utils/algorithms/BPTT.py, shac.py), stable-baselines3 stubbed
# ---------------------------------------------------------------------------------------------------
def load_reference_algorithms():
    """Import the reference's ``BPTT`` / ``shac`` trainer classes unmodified.

    They sit on stable-baselines3 for logging, learning-rate schedules and the policy classes.  None of that is on
    the path this repository replaces, and the package is not installed here, so the handful of names the two modules
    import are provided as inert ``sys.modules`` stubs (a logger that records into a dict, ``get_schedule_fn`` for
    constant rates, ``update_learning_rate``, ``polyak_update`` ...).  The trainer loops themselves —
    ``BPTT.learn`` (BPTT.py:77-180), ``TemporalDifferBase.__init__/_build`` (shac.py:53-136) — run as written, on
    whatever env object they are given; the policy is passed in as a class (``_create_policy``, shac.py:156-175)."""
    if "algos" in _state:
        return _state["algos"]
    import types

    load_reference_envs()                       # gymnasium / vec_env stubs, VisFly package on sys.path

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Logger:
        def __init__(self):
            self.values, self.dumps = {}, []

        def record(self, key, value, **_):
            self.values[key] = value

        def dump(self, step=0):
            self.dumps.append((step, dict(self.values)))

    def configure(folder=None, format_strings=None):
        return Logger()

    def get_schedule_fn(value):
        return value if callable(value) else (lambda _progress: float(value))

    def update_learning_rate(optimizer, lr):
        for group in optimizer.param_groups:
            group["lr"] = lr

    def safe_mean(xs):
        xs = [float(x) for x in xs]
        return sum(xs) / len(xs) if xs else float("nan")

    def polyak_update(params, target_params, tau):
        with th.no_grad():
            for p, t in zip(params, target_params):
                t.mul_(1 - tau).add_(p, alpha=tau)

    def get_parameters_by_name(model, included_names):
        return [p for n, p in model.state_dict().items() if any(k in n for k in included_names)]

    class _Anything:                            # names that are only used in annotations / never on this path
        def __init__(self, *a, **k):
            pass

    for name in ("Space", "Discrete", "MultiDiscrete", "MultiBinary"):      # annotations in utils/algorithms/common.py
        sys.modules["gymnasium.spaces"].__dict__.setdefault(name, _Anything)
    sb3 = mod("stable_baselines3")
    common = mod("stable_baselines3.common")
    sb3.common = common
    common.logger = mod("stable_baselines3.common.logger", Logger=Logger, configure=configure)
    mod("stable_baselines3.common.type_aliases", Schedule=object, RolloutBufferSamples=_Anything,
        DictRolloutBufferSamples=_Anything, ReplayBufferSamples=_Anything, DictReplayBufferSamples=_Anything)
    mod("stable_baselines3.common.utils", get_schedule_fn=get_schedule_fn, safe_mean=safe_mean,
        update_learning_rate=update_learning_rate, polyak_update=polyak_update,
        get_parameters_by_name=get_parameters_by_name)
    mod("stable_baselines3.common.buffers", BaseBuffer=_Anything)
    sys.modules["stable_baselines3.common.vec_env"].__dict__.setdefault("VecNormalize", _Anything)
    mod("stable_baselines3.sac")
    mod("stable_baselines3.sac.policies", MultiInputPolicy=_Anything)
    gym = mod("gym")
    gym.vector = mod("gym.vector")
    gym.vector.utils = mod("gym.vector.utils", spaces=sys.modules["gymnasium.spaces"])
    mod("VisFly.utils.policies")
    mod("VisFly.utils.policies.td_policies", CnnPolicy=_Anything, BasePolicy=_Anything, MultiInputPolicy=_Anything)
    mod("VisFly.utils.test")
    mod("VisFly.utils.test.debug", get_network_statistics=lambda *a, **k: None,
        check_none_parameters=lambda *a, **k: None)
    for absent in ("cv2", "matplotlib", "matplotlib.pyplot"):        # imported by VisFly/utils/common.py (set_seed)
        if absent not in sys.modules:
            try:
                __import__(absent)
            except Exception:
                mod(absent)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    from VisFly.utils.algorithms.BPTT import BPTT           # noqa
    from VisFly.utils.algorithms.shac import shac           # noqa
    _state["algos"] = {"BPTT": BPTT, "shac": shac}
    return _state["algos"]
