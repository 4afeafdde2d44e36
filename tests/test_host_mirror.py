"""The kernels' per-agent arithmetic (visfly_b200/csrc/vf_math.cuh) instantiated on the host, against the
oracle and the reference's golden vectors.  Lets the forward step and — above all — the hand-derived adjoint be
validated without a GPU.  (The library under test here is test infrastructure: oracle/host_mirror.cpp.)"""
import os

import numpy as np
import pytest
import torch as th

from _util import (make_oracle, mirror_bwd, mirror_fwd, oracle_grads, oracle_step_packed, pack,
                   random_flight_state, rel_l2, unpack, vf_params)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_CFG = {
    "euler": ("bodyrate", "euler", 0.005, 0.02, True),
    "rk4": ("bodyrate", "rk4", 0.0025, 0.02, True),
    "rk4_nolag": ("bodyrate", "rk4", 0.0025, 0.02, False),
    "euler_thrust": ("thrust", "euler", 0.005, 0.02, True),
    "rk4_s12": ("bodyrate", "rk4", 0.0025, 0.03, True),
    "euler_velocity": ("velocity", "euler", 0.005, 0.02, True),
    "rk4_velocity": ("velocity", "rk4", 0.0025, 0.02, True),
    "euler_position": ("position", "euler", 0.005, 0.02, False),
    "rk4_position": ("position", "rk4", 0.0025, 0.02, True),
}


@pytest.mark.parametrize("cfg", list(STEP_CFG))
def test_forward_step_matches_reference_golden(cfg):
    at, integ, dt, ctrl_dt, lag = STEP_CFG[cfg]
    z = np.load(os.path.join(GOLD, f"step_{cfg}.npz"))
    P = vf_params(at, dt)
    S = int(ctrl_dt / dt)
    before = [th.from_numpy(z[f"before_{k}_f32"]) for k in ("pos", "quat", "vel", "rate", "motor", "alpha")]
    out, obs, ext = mirror_fwd(P, pack(*before), th.from_numpy(z["action"]), S, integ, at, lag, want_ext=True)
    # judged against the float64 reference when there is one: the float32 reference itself is ~1e-7 away from it
    tag = "f64" if "obs_f64" in z.files else "f32"
    ref32_err = rel_l2(z["obs_f32"], z[f"obs_{tag}"])
    assert rel_l2(obs, z[f"obs_{tag}"]) < max(2e-7, 2 * ref32_err)
    for k, got in zip(("pos", "quat", "vel", "rate", "motor", "alpha"), unpack(out)):
        ref = z[f"after_{k}_{tag}"]
        assert rel_l2(got, ref) < max(5e-7, 3 * rel_l2(z[f"after_{k}_f32"], ref)), k
    assert rel_l2(ext[:, :3], z[f"acc_{tag}"]) < 1e-5
    assert rel_l2(ext[:, 4:], z[f"thrusts_{tag}"]) < 1e-6


@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
@pytest.mark.parametrize("at", ["bodyrate", "thrust"])
@pytest.mark.parametrize("lag", [True, False])
def test_adjoint_matches_autograd_float64(integ, dt, at, lag):
    n, S, wind = 96, int(0.02 / dt), (0.3, -0.2, 0.1)
    P = vf_params(at, dt, wind=wind)
    packed = pack(*random_flight_state(n, seed=3)).double()
    g = th.Generator().manual_seed(5)
    action = (th.rand(n, 4, generator=g) * 2 - 1).double()
    g_out = th.randn(5, n, 4, generator=g, dtype=th.float64)
    g_obs = th.randn(n, 13, generator=g, dtype=th.float64)
    orc = make_oracle(n, at, integ, dt, ctrl_delay=lag, wind=wind, dtype=th.float64)
    ref_gs, ref_ga = oracle_grads(orc, packed, action, g_out, g_obs)
    got_gs, got_ga = mirror_bwd(P, packed, action, g_out, g_obs, S, integ, at, lag)
    # constants differ at float32 rounding level (the oracle derives them in float64), hence 5e-6, not 1e-12
    assert rel_l2(got_gs, ref_gs) < 5e-6
    assert rel_l2(got_ga, ref_ga) < 5e-6
    # every block of the Jacobian on its own (catches a wrong small term hiding behind a large one)
    for name, sl in (("p", (0, slice(0, 3))), ("q", (1, slice(0, 4))), ("v", (2, slice(0, 3))),
                     ("w", (3, slice(0, 3))), ("mot", (4, slice(0, 4))), ("al", None)):
        go = th.zeros_like(g_out)
        if sl is None:
            go[0, :, 3], go[2, :, 3], go[3, :, 3] = g_out[0, :, 3], g_out[2, :, 3], g_out[3, :, 3]
        else:
            go[sl[0], :, sl[1]] = g_out[sl[0], :, sl[1]]
        r_gs, r_ga = oracle_grads(orc, packed, action, go, None)
        m_gs, m_ga = mirror_bwd(P, packed, action, go, None, S, integ, at, lag)
        for blk_ref, blk_got in zip(unpack(r_gs) + (r_ga,), unpack(m_gs) + (m_ga,)):
            if float(blk_ref.norm()) > 0:
                assert rel_l2(blk_got, blk_ref) < 2e-5, name
            else:
                assert float(blk_got.norm()) == 0.0, name


def test_adjoint_float32_within_tolerance_of_float64_autograd():
    n, dt, S = 256, 0.0025, 8
    P = vf_params("bodyrate", dt)
    packed = pack(*random_flight_state(n, seed=9))
    g = th.Generator().manual_seed(6)
    action = th.rand(n, 4, generator=g) * 2 - 1
    g_out, g_obs = th.randn(5, n, 4, generator=g), th.randn(n, 13, generator=g)
    orc = make_oracle(n, "bodyrate", "rk4", dt, dtype=th.float64)
    ref_gs, ref_ga = oracle_grads(orc, packed.double(), action.double(), g_out.double(), g_obs.double())
    got_gs, got_ga = mirror_bwd(P, packed, action, g_out, g_obs, S, "rk4")
    assert rel_l2(got_gs, ref_gs) < 1e-4 and rel_l2(got_ga, ref_ga) < 1e-4   # north-star gradient tolerance


def test_gradient_gates_follow_torch_conventions():
    """Saturated thrust clamp and saturated post-step clamps stop the gradient."""
    n, dt, S = 8, 0.005, 4
    P = vf_params("bodyrate", dt)
    pos, quat, vel, rate, motor, alpha = random_flight_state(n, seed=4)
    pos[:, 2] = 25.0              # above the z clamp (20): dL/dz must be gated to 0
    vel[:, 0] = 30.0              # beyond the velocity clamp
    packed = pack(pos, quat, vel, rate, motor, alpha).double()
    action = th.ones(n, 4, dtype=th.float64)            # full collective + max rates: rotor clamp saturates
    g_out = th.zeros(5, n, 4, dtype=th.float64)
    g_out[0, :, 2] = 1.0
    g_out[2, :, 0] = 1.0
    orc = make_oracle(n, "bodyrate", "euler", dt, dtype=th.float64)
    ref_gs, ref_ga = oracle_grads(orc, packed, action, g_out, None)
    got_gs, got_ga = mirror_bwd(P, packed, action, g_out, None, S, "euler")
    assert float(ref_gs.abs().max()) == 0.0 and float(got_gs.abs().max()) == 0.0
    assert float(got_ga.abs().max()) == 0.0 and float(ref_ga.abs().max()) == 0.0


def test_orientation_stays_unit_and_state_finite_long_run():
    n, dt, S = 64, 0.0025, 8
    P = vf_params("bodyrate", dt)
    packed = pack(*random_flight_state(n, seed=12))
    g = th.Generator().manual_seed(1)
    for _ in range(300):
        packed, obs = mirror_fwd(P, packed, th.rand(n, 4, generator=g) * 2 - 1, S, "rk4")
    assert bool(th.isfinite(packed).all())
    assert float((packed[1].norm(dim=1) - 1).abs().max()) < 1e-6


@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
@pytest.mark.parametrize("at,lag", [("bodyrate", True), ("thrust", False)])
def test_adjoint_matches_central_differences_float64(integ, dt, at, lag):
    """An autograd-free check of the hand-derived adjoint: J^T g from the reverse sweep against central differences
    of the forward step, coordinate by coordinate (all 24 inputs of every agent), in float64.  States are kept away
    from the clamp boundaries and from v_body = 0, where the step is not differentiable."""
    n, S, eps = 5, int(0.02 / dt), 1e-6
    P = vf_params(at, dt, wind=(0.3, -0.2, 0.1))
    packed = pack(*random_flight_state(n, seed=13, spread=0.5)).double()
    g = th.Generator().manual_seed(15)
    action = ((th.rand(n, 4, generator=g) * 2 - 1) * 0.5).double()
    if at == "bodyrate":
        action[:, 0] -= 0.3                                   # collective thrust well inside the rotor limits
    g_out = th.randn(5, n, 4, generator=g, dtype=th.float64)
    g_obs = th.randn(n, 13, generator=g, dtype=th.float64)

    def loss(pk, ac):
        out, obs = mirror_fwd(P, pk, ac, S, integ, at, lag)
        return (out * g_out).sum(dim=(0, 2)) + (obs * g_obs).sum(dim=1)      # per agent (agents are independent)

    gs, ga = mirror_bwd(P, packed, action, g_out, g_obs, S, integ, at, lag)
    fd_s, fd_a = th.zeros_like(packed), th.zeros_like(action)
    for plane in range(5):
        for lane in range(4):
            d = th.zeros_like(packed)
            d[plane, :, lane] = eps
            fd_s[plane, :, lane] = (loss(packed + d, action) - loss(packed - d, action)) / (2 * eps)
    for j in range(4):
        d = th.zeros_like(action)
        d[:, j] = eps
        fd_a[:, j] = (loss(packed, action + d) - loss(packed, action - d)) / (2 * eps)
    assert rel_l2(gs, fd_s) < 2e-7, rel_l2(gs, fd_s)
    assert rel_l2(ga, fd_a) < 2e-7, rel_l2(ga, fd_a)
    for got, ref in zip(unpack(gs), unpack(fd_s)):            # block by block
        assert float((got - ref).abs().max()) < 1e-6 * max(1.0, float(ref.abs().max()))
