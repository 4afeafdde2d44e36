"""The oracle (oracle/torch_oracle.py) against the golden vectors recorded from the real reference, the
known-answer numbers of SURVEY.md App. B, and — where the reference tree is mounted — the live reference."""
import os

import numpy as np
import pytest
import torch as th

from _reference import default_dtype, make_reference_dynamics, reference_available
from _util import make_oracle, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = {
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4_nolag": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=False),
    "euler_thrust": dict(action_type="thrust", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
    "rk4_s12": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.03, comm_delay=0.06, ctrl_delay=True),
    "euler_velocity": dict(action_type="velocity", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4_velocity": dict(action_type="velocity", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
    "euler_position": dict(action_type="position", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=False),
    "rk4_position": dict(action_type="position", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
}
FWD_ONLY = {"euler_velocity", "rk4_velocity", "euler_position", "rk4_position"}   # the reference's backward raises there
DT = {"f32": th.float32, "f64": th.float64}


def load(name):
    return np.load(os.path.join(GOLD, name))


def tags(z, prefix):
    return [t for t in ("f32", "f64") if f"{prefix}_{t}" in z.files]


# -- known answers --------------------------------------------------------------------------------------
def test_kat_matches_survey_appendix_b():
    """The fixture regenerated from the reference reproduces the numbers printed in SURVEY.md App. B."""
    z = load("kat.npz")
    b1 = z["B1_euler_full_state"]
    np.testing.assert_allclose(b1[0, :3], [1.0, -2.6705285e-08, 1.5000829], rtol=0, atol=1e-7)
    np.testing.assert_allclose(b1[0, 13:17], [2311.0203, 1135.9850, 2191.8691, 1184.1414], rtol=2e-7)
    np.testing.assert_allclose(z["B1_euler_alpha"][0], [13.989368, -6.0003133, 24.186993], rtol=2e-7)
    b2 = z["B2_euler_fifo3_full_state"]
    np.testing.assert_allclose(b2[:3, 2], [1.5001488, 1.5011138, 1.5033994], rtol=2e-7)
    np.testing.assert_allclose(b2[:3, 13], [1827.4973, 1919.6969, 1969.9915], rtol=2e-7)
    b3 = z["B3_rk4_full_state"]
    np.testing.assert_allclose(b3[2, 10:13], [0.62325293, -0.34816104, 1.0154399], rtol=2e-7)
    np.testing.assert_allclose(z["B3_rk4_alpha"][0], [13.992664, -5.9894352, 24.187946], rtol=2e-7)


@pytest.mark.parametrize("name,kw,steps", [
    ("B1_euler", dict(dt=0.005, integrator="euler", comm_delay=0.0), 3),
    ("B2_euler_fifo3", dict(dt=0.005, integrator="euler", comm_delay=0.06), 4),
    ("B3_rk4", dict(dt=0.0025, integrator="rk4", comm_delay=0.0), 3),
])
def test_oracle_kat(name, kw, steps):
    z = load("kat.npz")
    orc = make_oracle(1, "bodyrate", ctrl_dt=0.02, **kw)
    orc.reset(pos=[[1.0, 0.0, 1.5]])
    a = th.from_numpy(z["action"])
    for k in range(steps):
        orc.step(a)
        np.testing.assert_allclose(orc.full_state.numpy()[0], z[name + "_full_state"][k], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(orc.angular_acceleration.numpy()[0], z[name + "_alpha"][k], rtol=1e-6, atol=1e-9)


# -- trajectories ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("law", ["uniform", "hover"])
def test_oracle_trajectory_matches_reference_golden(cfg, law):
    z = load(f"traj_{cfg}.npz")
    for tag in tags(z, f"states_{law}"):
        dtype = DT[tag]
        kw = CONFIGS[cfg]
        n = z["init_pos"].shape[0]
        orc = make_oracle(n, dtype=dtype, **kw)
        orc.reset(pos=z["init_pos"], ori=z["init_quat"], vel=z["init_vel"], ori_vel=z["init_rate"])
        acts = th.from_numpy(z[f"actions_{law}"]).to(dtype)
        states = th.stack([orc.step(acts[t]).clone() for t in range(acts.shape[0])])
        gold = th.from_numpy(z[f"states_{law}_{tag}"])
        # same op sequence as the reference: only libm / BLAS-kernel differences between hosts may show up
        tol = 2e-6 if tag == "f32" else 1e-12
        assert rel_l2(states, gold) < tol
        assert rel_l2(orc.full_state, z[f"final_full_state_{law}_{tag}"]) < tol
        assert rel_l2(orc.angular_acceleration, z[f"final_alpha_{law}_{tag}"]) < 50 * tol


# -- gradients -----------------------------------------------------------------------------------------------
def hover_loss(dyn, acts, gamma=0.99):
    target = th.tensor([[1.0, 0.0, 1.5]], dtype=acts.dtype, device=acts.device)
    one = th.tensor([1.0, 0, 0, 0], dtype=acts.dtype, device=acts.device)
    total = 0.0
    for t in range(acts.shape[0]):
        dyn.step(acts[t])
        r = 0.1 - (dyn.position - target).norm(dim=1) / 90 - (dyn.orientation - one).norm(dim=1) * 1e-5 \
            - (dyn.velocity - 0).norm(dim=1) * 0.002 - (dyn.angular_velocity - 0).norm(dim=1) * 0.002
        total = total + (gamma ** t) * r
    return -total.mean()


@pytest.mark.parametrize("cfg", [c for c in CONFIGS if c not in FWD_ONLY])
def test_oracle_gradients_match_reference_autograd_golden(cfg):
    z = load(f"grad_{cfg}.npz")
    for tag in tags(z, "grad_actions"):
        dtype = DT[tag]
        n = z["init_pos"].shape[0]
        orc = make_oracle(n, dtype=dtype, **CONFIGS[cfg])
        leaves = [th.from_numpy(z["init_" + k]).to(dtype).requires_grad_(True) for k in ("pos", "quat", "vel", "rate")]
        acts = th.from_numpy(z["actions"]).to(dtype).requires_grad_(True)
        orc.reset()
        orc.pos, orc.vel, orc.ang_vel = leaves[0].T, leaves[2].T, leaves[3].T
        orc.q = tuple(leaves[1].T)
        loss = hover_loss(orc, acts)
        grads = th.autograd.grad(loss, [acts] + leaves)
        tol = 1e-5 if tag == "f32" else 1e-11
        assert abs(loss.item() - float(z[f"loss_{tag}"])) < tol * 10
        for k, g in zip(("actions", "pos", "quat", "vel", "rate"), grads):
            assert rel_l2(g, z[f"grad_{k}_{tag}"]) < tol, (k, tag)


# -- live reference (build container only) ---------------------------------------------------------------------
@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_oracle_is_bit_exact_with_live_reference(cfg):
    kw = CONFIGS[cfg]
    n = 48
    g = th.Generator().manual_seed(3)
    pos, vel, rate = th.rand(n, 3, generator=g) * 2, th.randn(n, 3, generator=g), th.randn(n, 3, generator=g)
    quat = th.randn(n, 4, generator=g)
    quat = quat / quat.norm(dim=1, keepdim=True)
    with default_dtype(th.float32):
        ref = make_reference_dynamics(n, **kw)
        ref.reset(pos=pos.clone(), ori=quat.clone(), vel=vel.clone(), ori_vel=rate.clone())
    orc = make_oracle(n, **kw)
    orc.reset(pos=pos, ori=quat, vel=vel, ori_vel=rate)
    for _ in range(40):
        a = th.rand(n, 4, generator=g) * 2 - 1
        s_ref, s_orc = ref.step(a.clone()), orc.step(a)
        assert th.equal(s_ref, s_orc)
    assert th.equal(ref.full_state, orc.full_state)
    assert th.equal(ref.angular_acceleration, orc.angular_acceleration)
    assert th.equal(ref.acceleration, orc.acceleration)
    assert th.equal(ref.direction, orc.direction)
