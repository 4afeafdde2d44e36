"""Golden fixtures for the two Dynamics options outside the default path, from the REAL reference:

    python tests/golden/make_wind_drag_golden.py         (build container only: needs /root/reference)

  wind_drag.npz
    windfn_<integ>_*  time-varying wind functions (six expression strings, reference dynamics.py:136-151, :384-388):
                      24 control steps of 16 agents, a partial reset (explicit t) after step 10
    drag_*            drag_random=0.3 (dynamics.py:244-246): coefficients drawn by the full reset under a fixed
                      torch seed, 16 control steps of fast-flying agents
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from _reference import default_dtype, make_reference_dynamics  # noqa: E402
from _util import random_flight_state  # noqa: E402

WIND_FN = ["0.5*th.sin(2*x)", "0.3*th.cos(x)", "0*x", "0.9*y+0.05", "0.8*y-0.02", "0*y+0.01*x"]
WIND_KW = {
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
}
DRAG_KW = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True,
               drag_random=0.3)
DRAG_SEED = 123
RESET_AT, RESET_IDX = 10, [3, 7]


def actions(T, n, seed):
    g = th.Generator().manual_seed(seed)
    a = (th.rand(T, n, 4, generator=g) * 2 - 1) * 0.3
    a[..., 0] += -1.0 / 3.0
    return a


def main():
    out = {"wind_fn": np.array(WIND_FN)}
    n, T = 16, 24
    init = random_flight_state(n, seed=31, spread=0.5, dtype=th.float64)
    acts = actions(T, n, 33)
    out["windfn_actions"] = acts.numpy()
    for k, x in zip(("pos", "quat", "vel", "rate"), init[:4]):
        out["windfn_init_" + k] = x.numpy()
    out["windfn_reset_pos"] = np.array([[0.5, -0.5, 2.0], [-1.0, 1.0, 1.0]])
    out["windfn_reset_t"] = np.array([1.25, 3.5])
    for integ, kw in WIND_KW.items():
        for dtype, tag in ((th.float32, "f32"), (th.float64, "f64")):
            with default_dtype(dtype):
                d = make_reference_dynamics(n, dtype=dtype, wind_settings=list(WIND_FN), **kw)
                pos, quat, vel, rate = (x.to(dtype) for x in init[:4])
                d.reset(pos=pos.clone(), ori=quat.clone(), vel=vel.clone(), ori_vel=rate.clone())
                states, winds = [], []
                for t in range(T):
                    if t == RESET_AT:
                        d.reset(pos=th.tensor(out["windfn_reset_pos"], dtype=dtype), indices=RESET_IDX,
                                t=th.tensor(out["windfn_reset_t"], dtype=dtype))
                    states.append(d.step(acts[t].to(dtype).clone()).clone())
                    winds.append(d.wind_velocity.clone())
                out[f"windfn_{integ}_states_{tag}"] = th.stack(states).numpy()
                out[f"windfn_{integ}_wind_{tag}"] = th.stack(winds).numpy()
                out[f"windfn_{integ}_final_full_state_{tag}"] = d.full_state.numpy()

    n, T = 16, 16
    init = random_flight_state(n, seed=41, spread=1.0, dtype=th.float64)
    vel = init[2] * 6.0                       # fast enough for the drag to matter
    acts = actions(T, n, 43)
    out["drag_actions"] = acts.numpy()
    for k, x in zip(("pos", "quat", "vel", "rate"), (init[0], init[1], vel, init[3])):
        out["drag_init_" + k] = x.numpy()
    for dtype, tag in ((th.float32, "f32"), (th.float64, "f64")):
        with default_dtype(dtype):
            d = make_reference_dynamics(n, dtype=dtype, **DRAG_KW)
            th.manual_seed(DRAG_SEED)
            d.reset(pos=init[0].to(dtype).clone(), ori=init[1].to(dtype).clone(), vel=vel.to(dtype).clone(),
                    ori_vel=init[3].to(dtype).clone())
            out[f"drag_k_lin_{tag}"] = d._linear_drag_coeffs.numpy()
            out[f"drag_k_quad_{tag}"] = d._quad_drag_coeffs.numpy()
            out[f"drag_states_{tag}"] = th.stack([d.step(acts[t].to(dtype).clone()).clone() for t in range(T)]).numpy()
    out["drag_seed"] = np.array(DRAG_SEED)
    np.savez_compressed(os.path.join(HERE, "wind_drag.npz"), **out)
    print("wrote wind_drag.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
