"""Shared helpers of the test-suite: seeded inputs, packed-state conversion, the host mirror, the oracle."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from visfly_b200.params import VfParams, action_scaling, build_vf_params, load_drone_model  # noqa: E402
from visfly_b200.type import ACTION_TYPE  # noqa: E402
from oracle.torch_oracle import OracleDynamics  # noqa: E402

INTEGRATOR = {"euler": 0, "rk4": 1}
ACTION = {"thrust": 0, "bodyrate": 1, "velocity": 2, "position": 3}
FLAG_CTRL_DELAY = 1


def vf_params(action_type="bodyrate", dt=0.005, cfg="drone_state", wind=(0.0, 0.0, 0.0)) -> VfParams:
    at = {"bodyrate": ACTION_TYPE.BODYRATE, "thrust": ACTION_TYPE.THRUST, "velocity": ACTION_TYPE.VELOCITY,
          "position": ACTION_TYPE.POSITION}[action_type]
    model = load_drone_model(cfg, dt)
    return build_vf_params(model, at, action_scaling(model, at), wind)


def random_flight_state(n, seed=0, spread=1.0, dtype=th.float32):
    """A batch of plausible mid-flight states: (pos, quat, vel, rate, motor, alpha), all (n,k) row-major."""
    g = th.Generator().manual_seed(seed)
    r = lambda *s: th.rand(*s, generator=g, dtype=th.float64)
    rn = lambda *s: th.randn(*s, generator=g, dtype=th.float64)
    pos = th.stack([r(n) * 4 - 2, r(n) * 4 - 2, r(n) * 3 + 0.5], 1)
    axis = rn(n, 3)
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = rn(n, 1) * 0.4 * spread
    quat = th.cat([th.cos(ang / 2), axis * th.sin(ang / 2)], 1)
    vel = rn(n, 3) * 1.5 * spread
    rate = rn(n, 3) * 1.0 * spread
    motor = 1658.0 + rn(n, 4) * 250 * spread
    alpha = rn(n, 3) * 8 * spread
    return tuple(x.to(dtype) for x in (pos, quat, vel, rate, motor, alpha))


def pack(pos, quat, vel, rate, motor, alpha):
    n = pos.shape[0]
    out = th.zeros((5, n, 4), dtype=pos.dtype)
    out[0, :, :3], out[0, :, 3] = pos, alpha[:, 0]
    out[1] = quat
    out[2, :, :3], out[2, :, 3] = vel, alpha[:, 1]
    out[3, :, :3], out[3, :, 3] = rate, alpha[:, 2]
    out[4] = motor
    return out


def unpack(packed):
    pos, quat, vel, rate, motor = packed[0, :, :3], packed[1], packed[2, :, :3], packed[3, :, :3], packed[4]
    alpha = th.stack([packed[0, :, 3], packed[2, :, 3], packed[3, :, 3]], 1)
    return pos, quat, vel, rate, motor, alpha


def make_oracle(n, action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, ctrl_delay=True,
                comm_delay=0.0, wind=(0, 0, 0), dtype=th.float32, cfg="drone_state", device="cpu"):
    return OracleDynamics(n, action_type, dt=dt, ctrl_dt=ctrl_dt, ctrl_delay=ctrl_delay, comm_delay=comm_delay,
                          integrator=integrator, wind=wind, dtype=dtype, cfg=cfg, device=device)


def oracle_step_packed(orc: OracleDynamics, packed, action):
    """One control step of the oracle from a packed state; returns (packed_out, obs13)."""
    orc.load_packed(packed)
    obs = orc.step(action)
    return orc.packed(), obs


def rel_l2(a, b):
    a, b = th.as_tensor(a, dtype=th.float64), th.as_tensor(b, dtype=th.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


# ---------------------------------------------------------------------------------------------------
# host mirror (oracle/host_mirror.cpp): the kernels' arithmetic compiled for the CPU, test-only
# ---------------------------------------------------------------------------------------------------
_MIRROR = {}


def host_mirror():
    if "lib" in _MIRROR:
        return _MIRROR["lib"]
    so = os.path.join(ROOT, "oracle", "libvf_host_mirror.so")
    src = os.path.join(ROOT, "oracle", "host_mirror.cpp")
    hdr = os.path.join(ROOT, "visfly_b200", "csrc", "vf_math.cuh")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    _MIRROR["lib"] = lib
    return lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def mirror_fwd(params, packed, action, substeps, integrator, action_type="bodyrate", ctrl_delay=True,
               want_ext=False):
    lib = host_mirror()
    dt = packed.dtype
    fn = lib.vfm_step_fwd_f32 if dt == th.float32 else lib.vfm_step_fwd_f64
    n = packed.shape[1]
    packed, action = packed.contiguous(), action.to(dt).contiguous()
    out = th.empty_like(packed)
    obs = th.empty((n, 13), dtype=dt)
    ext = th.empty((n, 8), dtype=dt) if want_ext else None
    fn(ctypes.byref(params), n, substeps, INTEGRATOR[integrator], ACTION[action_type],
       FLAG_CTRL_DELAY if ctrl_delay else 0, _ptr(packed), _ptr(action), _ptr(out), _ptr(obs), _ptr(ext))
    return (out, obs, ext) if want_ext else (out, obs)


def mirror_bwd(params, packed, action, g_out, g_obs, substeps, integrator, action_type="bodyrate",
               ctrl_delay=True):
    lib = host_mirror()
    dt = packed.dtype
    fn = lib.vfm_step_bwd_f32 if dt == th.float32 else lib.vfm_step_bwd_f64
    n = packed.shape[1]
    packed, action = packed.contiguous(), action.to(dt).contiguous()
    g_out = None if g_out is None else g_out.to(dt).contiguous()
    g_obs = None if g_obs is None else g_obs.to(dt).contiguous()
    g_in = th.empty_like(packed)
    g_act = th.empty((n, 4), dtype=dt)
    rc = fn(ctypes.byref(params), n, substeps, INTEGRATOR[integrator], ACTION[action_type],
            FLAG_CTRL_DELAY if ctrl_delay else 0, _ptr(packed), _ptr(action), _ptr(g_out), _ptr(g_obs),
            _ptr(g_in), _ptr(g_act))
    assert rc == 0
    return g_in, g_act


def oracle_grads(orc: OracleDynamics, packed, action, g_out, g_obs):
    """torch.autograd through the oracle: d<g_out,packed_out> + <g_obs,obs> / d(packed, action)."""
    packed = packed.clone().requires_grad_(True)
    action = action.clone().to(packed.dtype).requires_grad_(True)
    out, obs = oracle_step_packed(orc, packed, action)
    loss = 0
    if g_out is not None:
        loss = loss + (out * g_out.to(out.dtype)).sum()
    if g_obs is not None:
        loss = loss + (obs * g_obs.to(obs.dtype)).sum()
    gp, ga = th.autograd.grad(loss, (packed, action), allow_unused=True)
    gp = th.zeros_like(packed) if gp is None else gp
    ga = th.zeros_like(action) if ga is None else ga
    return gp.detach(), ga.detach()
