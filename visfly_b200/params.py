"""Drone parameter loading and the ``VfParams`` POD handed to the CUDA kernels.

The physical constants are derived from the drone JSON with float32 torch arithmetic in the same order as
the reference (``Dynamics.load`` envs/base/dynamics.py:562-608, ``_init`` :94-114, ``_get_scale_factor``
:610-689) so that every constant the kernel sees is the float32 number the reference computes with.
Keys the reference requires but some shipped JSONs lack (``THRUST_PID``, ``max_acc``)
fall back to the ``drone_state`` values instead of raising ``KeyError``.
"""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass, field
from typing import Dict, Sequence, Tuple

import torch as th

from .type import ACTION_TYPE, PID, Uniform, bound

_CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "drone")

GRAVITY = 9.81

# post-step clamps of the reference ("_ugly_fix", dynamics.py:374-382; the code clamps z to 20, not 10)
POS_LO = (-100.0, -100.0, 0.0)
POS_HI = (100.0, 100.0, 20.0)
VEL_LIM = 20.0
RATE_LIM = 10.0

_FALLBACK = {
    "max_acc": 3.0,
    "THRUST_PID": {"p": 1.0, "i": 0.0, "d": 0.0},
}


class VfParams(ctypes.Structure):
    """ctypes mirror of ``struct VfParams`` in include/visfly_b200.h (field order is the ABI)."""

    _fields_ = [
        ("dt", ctypes.c_float),
        ("mass", ctypes.c_float),
        ("inv_mass", ctypes.c_float),
        ("J", ctypes.c_float * 3),
        ("J_inv", ctypes.c_float * 3),
        ("B", ctypes.c_float * 16),
        ("B_inv", ctypes.c_float * 16),
        ("thrust_map", ctypes.c_float * 3),
        ("motor_c", ctypes.c_float),
        ("thrust_min", ctypes.c_float),
        ("thrust_max", ctypes.c_float),
        ("k_lin", ctypes.c_float * 3),
        ("k_quad", ctypes.c_float * 3),
        ("JKp", ctypes.c_float * 9),
        ("Kd", ctypes.c_float * 9),
        ("act_half", ctypes.c_float * 4),
        ("act_mean", ctypes.c_float * 4),
        ("gravity", ctypes.c_float * 3),
        ("wind", ctypes.c_float * 3),
        ("pos_lo", ctypes.c_float * 3),
        ("pos_hi", ctypes.c_float * 3),
        ("vel_lim", ctypes.c_float),
        ("rate_lim", ctypes.c_float),
        ("Kp", ctypes.c_float * 9),
        ("vel_kp", ctypes.c_float),
        ("vel_kd", ctypes.c_float),
        ("pos_kd", ctypes.c_float),
    ]

    def as_dict(self) -> Dict[str, object]:
        out = {}
        for name, ctype in self._fields_:
            v = getattr(self, name)
            out[name] = float(v) if ctype is ctypes.c_float else [float(x) for x in v]
        return out


def resolve_cfg_path(cfg: str) -> str:
    """``cfg`` is a name under configs/drone (reference dynamics.py:96) or a path to a JSON file."""
    for cand in (cfg, cfg + ".json", os.path.join(_CFG_DIR, cfg + ".json"), os.path.join(_CFG_DIR, cfg)):
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError(f"drone config '{cfg}' not found (looked in {_CFG_DIR})")


@dataclass
class DroneModel:
    """Float32 constants of one drone model, as torch CPU tensors (the host-side source of truth)."""

    name: str
    dt: float
    m: th.Tensor
    inertia: th.Tensor            # (3,3) diagonal
    inertia_inv: th.Tensor
    B_allocation: th.Tensor       # (4,4)
    B_allocation_inv: th.Tensor
    thrust_map: th.Tensor         # (3,)
    motor_c: th.Tensor
    bd_rotor_omega: bound
    bd_thrust: bound
    bd_rate: bound
    bd_acc: bound
    bd_spd: bound
    bd_pos: bound
    linear_drag: th.Tensor        # (3,1)
    quad_drag: th.Tensor          # (3,1)
    BODYRATE_PID: PID
    THRUST_PID: PID
    VELOCITY_PID: PID
    POSITION_PID: PID
    init_thrust: th.Tensor = field(default=None)
    init_motor_omega: th.Tensor = field(default=None)

    def rotor_omega_from_thrust(self, thrust: th.Tensor) -> th.Tensor:
        a, b, c = self.thrust_map[0], self.thrust_map[1], self.thrust_map[2]
        return (1 / (2 * a)) * (-b + th.sqrt(b.pow(2) - 4 * a * (c - thrust)))

    def thrust_from_rotor_omega(self, omega: th.Tensor) -> th.Tensor:
        a, b, c = self.thrust_map[0], self.thrust_map[1], self.thrust_map[2]
        return a * omega.pow(2) + b * omega + c


def load_drone_model(cfg: str, dt: float) -> DroneModel:
    with open(resolve_cfg_path(cfg), "r") as f:
        data = json.load(f)
    for k, v in _FALLBACK.items():
        data.setdefault(k, v)

    g = th.tensor([[0, 0, -GRAVITY]]).T
    m = th.tensor(data["mass"])
    cross_sections = th.tensor([data["cross_sections"]]).T
    quad_drag = th.tensor([data["quad_drag_coeffs"]]).T * 0.5 * 1.225 * cross_sections
    linear_drag = th.tensor([data["linear_drag_coeffs"]]).T
    inertia = th.diag(th.tensor(data["inertia"]))

    def _pid(key):
        d = data[key]
        return PID(p=th.tensor(d["p"]), i=th.tensor(d["i"]), d=th.tensor(d["d"]))

    kappa = th.tensor(data["kappa"])
    arm = th.tensor(data["arm_length"])
    thrust_map = th.tensor(data["thrust_map"])
    motor_c = th.exp(-th.tensor(1 / data["motor_tau"]) * dt)
    omega_bd = bound(max=data["motor_omega_max"], min=data["motor_omega_min"])
    thrust_bd = bound(
        max=thrust_map[0] * omega_bd.max ** 2 + thrust_map[1] * omega_bd.max + thrust_map[2], min=0)

    # rotor arms: X configuration, unit directions scaled by the arm length (dynamics.py:100-113)
    direction = th.tensor([[1, -1, -1, 1.0], [-1, -1, 1, 1], [0, 0, 0, 0.0]])
    direction = direction / direction.norm(dim=0)
    t_bm = arm * direction
    B = th.vstack([th.ones(1, 4), t_bm[:2], kappa * th.tensor([1, -1, 1, -1])])

    model = DroneModel(
        name=data["name"], dt=dt, m=m, inertia=inertia, inertia_inv=th.inverse(inertia),
        B_allocation=B, B_allocation_inv=th.inverse(B), thrust_map=thrust_map, motor_c=motor_c,
        bd_rotor_omega=omega_bd, bd_thrust=thrust_bd,
        bd_rate=bound(max=th.tensor(data["max_rate"]), min=th.tensor(-data["max_rate"])),
        bd_acc=bound(max=(data["max_acc"] * -g[2]).clone(), min=th.tensor(0)),
        bd_spd=bound(max=th.tensor(data["max_spd"]), min=th.tensor(-data["max_spd"])),
        bd_pos=bound(max=th.tensor(data["max_pos"]), min=th.tensor(-data["max_pos"])),
        linear_drag=linear_drag, quad_drag=quad_drag,
        BODYRATE_PID=_pid("BODYRAYE_PID"), THRUST_PID=_pid("THRUST_PID"),
        VELOCITY_PID=_pid("VELOCITY_PID"), POSITION_PID=_pid("POSITION_PID"),
    )
    model.init_thrust = -(m * g / 4)[-1]                           # dynamics.py:85
    model.init_motor_omega = model.rotor_omega_from_thrust(model.init_thrust)   # dynamics.py:86
    return model


def action_scaling(model: DroneModel, action_type: ACTION_TYPE,
                   normal_range: Tuple[float, float] = (-1, 1)) -> Dict[str, Uniform]:
    """``_normal_params`` of the reference ("max_min" thrust normalisation, dynamics.py:610-689)."""
    lo, hi = normal_range

    def _affine(bd):
        scale = (bd.max - bd.min) / (hi - lo)
        return Uniform(mean=bd.max - scale * hi, half=scale)

    if action_type == ACTION_TYPE.BODYRATE:
        return {"acc": _affine(model.bd_acc), "bodyrate": _affine(model.bd_rate)}
    if action_type == ACTION_TYPE.THRUST:
        return {"acc": _affine(model.bd_acc)}
    yaw_scale = th.as_tensor(th.pi - (-th.pi)) / (hi - lo)
    yaw_bias = th.pi - yaw_scale * hi
    if action_type == ACTION_TYPE.VELOCITY:
        # the reference stores half=yaw_bias here (dynamics.py:671); kept, it is part of the behaviour
        return {"velocity": _affine(model.bd_spd), "yaw": Uniform(mean=yaw_bias, half=yaw_bias)}
    return {"velocity": _affine(model.bd_pos), "yaw": Uniform(mean=yaw_bias, half=yaw_scale)}


def build_vf_params(model: DroneModel, action_type: ACTION_TYPE, scaling: Dict[str, Uniform],
                    wind: Sequence[float] = (0.0, 0.0, 0.0)) -> VfParams:
    p = VfParams()
    p.dt = float(th.tensor(model.dt, dtype=th.float32))
    p.mass = float(model.m)
    p.inv_mass = float(1.0 / model.m)
    for i in range(3):
        p.J[i] = float(model.inertia[i, i])
        p.J_inv[i] = float(model.inertia_inv[i, i])
        p.k_lin[i] = float(model.linear_drag[i, 0])
        p.k_quad[i] = float(model.quad_drag[i, 0])
        p.gravity[i] = (0.0, 0.0, -GRAVITY)[i]
        p.wind[i] = float(wind[i])
        p.pos_lo[i] = POS_LO[i]
        p.pos_hi[i] = POS_HI[i]
        p.thrust_map[i] = float(model.thrust_map[i])
    for i in range(16):
        p.B[i] = float(model.B_allocation.flatten()[i])
        p.B_inv[i] = float(model.B_allocation_inv.flatten()[i])
    jkp = model.inertia @ model.BODYRATE_PID.p.to(th.float32)
    kd = model.BODYRATE_PID.d.to(th.float32)
    for i in range(9):
        p.JKp[i] = float(jkp.flatten()[i])
        p.Kd[i] = float(kd.flatten()[i])
    p.motor_c = float(model.motor_c)
    p.thrust_min = float(model.bd_thrust.min)
    p.thrust_max = float(model.bd_thrust.max)
    p.vel_lim = VEL_LIM
    p.rate_lim = RATE_LIM
    if action_type == ACTION_TYPE.BODYRATE:
        halves = [scaling["acc"].half] + [scaling["bodyrate"].half] * 3
        means = [scaling["acc"].mean] + [scaling["bodyrate"].mean] * 3
    elif action_type == ACTION_TYPE.THRUST:
        halves = [scaling["acc"].half] * 4
        means = [scaling["acc"].mean] * 4
    else:                                            # velocity / position: [yaw, x, y, z]   dynamics.py:714-729
        halves = [scaling["yaw"].half] + [scaling["velocity"].half] * 3
        means = [scaling["yaw"].mean] + [scaling["velocity"].mean] * 3
    kp = model.BODYRATE_PID.p.to(th.float32)
    for i in range(9):
        p.Kp[i] = float(kp.flatten()[i])
    p.vel_kp = float(model.VELOCITY_PID.p)
    p.vel_kd = float(model.VELOCITY_PID.d)
    p.pos_kd = float(model.POSITION_PID.d)
    for i in range(4):
        p.act_half[i] = float(halves[i])
        p.act_mean[i] = float(means[i])
    return p


# ---------------------------------------------------------------------------------------------------
# fused env step: ctypes mirror of ``struct VfEnvSpec`` (include/visfly_b200.h)
# ---------------------------------------------------------------------------------------------------
TASK_HOVER, TASK_NAVIGATION, TASK_RACING, TASK_CUSTOM = 0, 1, 2, 3
OBS_STATE13, OBS_RACING16 = 0, 1
GEN_UNIFORM, GEN_NORMAL, GEN_TABLE = 0, 1, 2
GEN_MAX_BOXES = 4
ENV_FLAG_NO_RESET = 1
RBIT_DONE, RBIT_EPISODE_DONE, RBIT_SUCCESS, RBIT_TRUNCATED, RBIT_COLLIDED = 1, 2, 4, 8, 16
EBIT_EPISODE_DONE, EBIT_ONCE_COLLIDED = 1, 2


class VfEnvSpec(ctypes.Structure):
    _fields_ = [
        ("task", ctypes.c_int),
        ("obs_kind", ctypes.c_int),
        ("max_episode_steps", ctypes.c_int),
        ("collision_reset", ctypes.c_int),
        ("fifo_depth", ctypes.c_int),
        ("uav_radius", ctypes.c_float),
        ("bbox_lo", ctypes.c_float * 3),
        ("bbox_hi", ctypes.c_float * 3),
        ("target", ctypes.c_float * 3),
        ("success_radius", ctypes.c_float),
        ("n_gates", ctypes.c_int),
        ("gates", (ctypes.c_float * 3) * 4),
        ("gen_kind", ctypes.c_int),
        ("gen_boxes", ctypes.c_int),
        ("gen_mean", ((ctypes.c_float * 3) * 4) * GEN_MAX_BOXES),
        ("gen_half", ((ctypes.c_float * 3) * 4) * GEN_MAX_BOXES),
        ("gen_heading", ctypes.c_int * GEN_MAX_BOXES),
        ("init_motor_omega", ctypes.c_float),
        ("seed", ctypes.c_ulonglong),
        ("agent_offset", ctypes.c_uint),
    ]


class VfEnvMirror(ctypes.Structure):
    """ctypes mirror of ``struct VfEnvMirror``: page-locked host destinations of obs / reward / done (numpy mode) and
    the completion word the last thread block raises (``flag`` page-locked host, ``counter`` device, zeroed)."""
    _fields_ = [("obs", ctypes.c_void_p), ("reward", ctypes.c_void_p), ("done", ctypes.c_void_p),
                ("flag", ctypes.c_void_p), ("counter", ctypes.c_void_p), ("flag_value", ctypes.c_uint)]


MAX_PEERS = 8


class VfPeerScatter(ctypes.Structure):
    """ctypes mirror of ``struct VfPeerScatter``: peer-mapped gather buffers the rollout's last env step scatters the
    per-agent episode returns into (the fused all-gather of SURVEY.md §8e)."""
    _fields_ = [("dst", ctypes.c_void_p * MAX_PEERS), ("offset", ctypes.c_longlong), ("world", ctypes.c_int)]


FIFO_MAX_ROWS = 8


class VfFifoRows(ctypes.Structure):
    """ctypes mirror of ``struct VfFifoRows``: the comm-delay FIFO as a device-resident ring of (n,4) rows."""
    _fields_ = [("row", ctypes.c_void_p * FIFO_MAX_ROWS), ("depth", ctypes.c_int)]
