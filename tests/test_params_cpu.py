"""visfly_b200.params (host logic of the product, no GPU needed): the parameter block the kernels read against the
tensors the live reference derives from the same drone file (envs/base/dynamics.py:100-114 `_init`, :562-608 `load`,
:610-689 `_get_scale_factor`), for every file the reference can load and all four action types — bit for bit."""
import numpy as np
import pytest
import torch as th

from _reference import make_reference_dynamics, reference_available
from visfly_b200.params import action_scaling, build_vf_params, load_drone_model
from visfly_b200.type import ACTION_TYPE

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
REF_LOADS = ["drone_state", "drone_d435i_jetson_orin_nx", "drone_d435i_jetson_orin_nx_fast"]
AT = {"bodyrate": ACTION_TYPE.BODYRATE, "thrust": ACTION_TYPE.THRUST, "velocity": ACTION_TYPE.VELOCITY,
      "position": ACTION_TYPE.POSITION}


def f32(x):
    return np.asarray(th.as_tensor(x, dtype=th.float32).reshape(-1).numpy())


@pytest.mark.parametrize("cfg", REF_LOADS)
@pytest.mark.parametrize("at", list(AT))
@pytest.mark.parametrize("dt", [0.005, 0.0025])
def test_parameter_block_equals_reference_tensors(cfg, at, dt):
    ref = make_reference_dynamics(2, action_type=at, dt=dt, ctrl_dt=0.02, cfg=cfg, comm_delay=0.0)
    model = load_drone_model(cfg, dt)
    p = build_vf_params(model, AT[at], action_scaling(model, AT[at]), (0.0, 0.0, 0.0))
    eq = lambda got, want: np.testing.assert_array_equal(np.asarray(got, dtype=np.float32), f32(want))
    eq([p.mass], ref.m)
    eq(p.J[:], th.diagonal(ref._inertia))
    eq(p.J_inv[:], th.diagonal(ref._inertia_inv))
    eq(p.B[:], ref._B_allocation)
    eq(p.B_inv[:], ref._B_allocation_inv)
    eq(p.thrust_map[:], ref._thrust_map)
    eq([p.motor_c], ref._c)
    eq([p.thrust_min, p.thrust_max], [ref._bd_thrust.min, ref._bd_thrust.max])
    eq(p.k_lin[:], ref._linear_drag_coeffs_mean)
    eq(p.k_quad[:], ref._quad_drag_coeffs_mean)
    eq(p.JKp[:], ref._inertia @ ref._BODYRATE_PID.p)
    eq(p.Kd[:], ref._BODYRATE_PID.d)
    eq(p.Kp[:], ref._BODYRATE_PID.p)
    eq([p.vel_kp, p.vel_kd, p.pos_kd], [ref._VELOCITY_PID.p, ref._VELOCITY_PID.d, ref._POSITION_PID.d])
    n = ref._normal_params
    if at == "bodyrate":
        half = [n["acc"].half] + [n["bodyrate"].half] * 3
        mean = [n["acc"].mean] + [n["bodyrate"].mean] * 3
    elif at == "thrust":
        half, mean = [n["acc"].half] * 4, [n["acc"].mean] * 4
    else:
        half = [n["yaw"].half] + [n["velocity"].half] * 3
        mean = [n["yaw"].mean] + [n["velocity"].mean] * 3
    eq(p.act_half[:], th.stack([th.as_tensor(h, dtype=th.float32).reshape(()) for h in half]))
    eq(p.act_mean[:], th.stack([th.as_tensor(m, dtype=th.float32).reshape(()) for m in mean]))
    # hover rotor speed / thrust after reset (dynamics.py:85-86)
    eq([float(model.init_thrust)], ref._init_thrust)
    eq([float(model.init_motor_omega)], ref._init_motor_omega)
