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
