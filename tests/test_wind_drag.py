"""The two Dynamics options beside the default path — time-varying wind functions (reference dynamics.py:136-151,
:384-388) and ``drag_random`` (:244-246) — against fixtures recorded from the real reference
(``tests/golden/make_wind_drag_golden.py``): the oracle restatement on the CPU, the CUDA engine on the GPU."""
import os

import numpy as np
import pytest
import torch as th

from _reference import default_dtype, make_reference_dynamics, reference_available
from _util import OracleDynamics, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wind_drag.npz")
WIND_KW = {
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
}
DRAG_KW = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True)
RESET_AT, RESET_IDX = 10, [3, 7]
DT = {"f32": th.float32, "f64": th.float64}


def wind_rollout(make, z, dtype, dev="cpu"):
    """Replays the fixture's wind-function run through `make(n, wind_fn)`; returns (states, winds)."""
    acts = th.from_numpy(z["windfn_actions"]).to(dtype).to(dev)
    n = acts.shape[1]
    d = make(n, [str(s) for s in z["wind_fn"]])
    cv = lambda k: th.from_numpy(z[k]).to(dtype).to(dev)
    d.reset(pos=cv("windfn_init_pos"), ori=cv("windfn_init_quat"), vel=cv("windfn_init_vel"),
            ori_vel=cv("windfn_init_rate"))
    states, winds = [], []
    for t in range(acts.shape[0]):
        if t == RESET_AT:
            d.reset(pos=cv("windfn_reset_pos"), indices=RESET_IDX, t=cv("windfn_reset_t"))
        states.append(d.step(acts[t].clone()).clone())
        w = d.wind_velocity if hasattr(d, "wind_velocity") else d.wind
        winds.append(w.clone())
    return th.stack(states).cpu(), th.stack(winds).cpu()


# -- CPU: the oracle ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("integ", list(WIND_KW))
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_wind_functions_match_reference_golden(integ, tag):
    z = np.load(GOLD)
    kw = dict(WIND_KW[integ])
    make = lambda n, fn: OracleDynamics(n, kw.pop("action_type"), wind=fn, dtype=DT[tag], **kw)
    states, winds = wind_rollout(make, z, DT[tag])
    tol = 2e-6 if tag == "f32" else 1e-12
    assert rel_l2(winds, z[f"windfn_{integ}_wind_{tag}"]) < tol
    assert rel_l2(states, z[f"windfn_{integ}_states_{tag}"]) < tol


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_drag_random_matches_reference_golden(tag):
    z = np.load(GOLD)
    dtype = DT[tag]
    acts = th.from_numpy(z["drag_actions"]).to(dtype)
    orc = OracleDynamics(acts.shape[1], drag_random=0.3, dtype=dtype, **DRAG_KW)
    cv = lambda k: th.from_numpy(z[k]).to(dtype)
    with default_dtype(dtype):
        th.manual_seed(int(z["drag_seed"]))
        orc.reset(pos=cv("drag_init_pos"), ori=cv("drag_init_quat"), vel=cv("drag_init_vel"), ori_vel=cv("drag_init_rate"))
    np.testing.assert_array_equal(orc.M.k_lin.numpy(), z[f"drag_k_lin_{tag}"])
    np.testing.assert_array_equal(orc.M.k_quad.numpy(), z[f"drag_k_quad_{tag}"])
    states = th.stack([orc.step(acts[t]).clone() for t in range(acts.shape[0])])
    assert rel_l2(states, z[f"drag_states_{tag}"]) < (2e-6 if tag == "f32" else 1e-12)
    # the drawn coefficients differ from the means, and they matter for this flight
    plain = OracleDynamics(acts.shape[1], dtype=dtype, **DRAG_KW)
    plain.reset(pos=cv("drag_init_pos"), ori=cv("drag_init_quat"), vel=cv("drag_init_vel"), ori_vel=cv("drag_init_rate"))
    other = th.stack([plain.step(acts[t]).clone() for t in range(acts.shape[0])])
    assert rel_l2(other, z[f"drag_states_{tag}"]) > 1e-4


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("integ", list(WIND_KW))
def test_oracle_wind_functions_bit_exact_with_live_reference(integ):
    z = np.load(GOLD)
    kw = dict(WIND_KW[integ])
    ref = wind_rollout(lambda n, fn: make_reference_dynamics(n, wind_settings=fn, **kw), z, th.float32)
    kw2 = dict(kw)
    orc = wind_rollout(lambda n, fn: OracleDynamics(n, kw2.pop("action_type"), wind=fn, **kw2), z, th.float32)
    assert th.equal(ref[1], orc[1]) and th.equal(ref[0], orc[0])


def test_three_string_wind_form_is_rejected_like_the_reference():
    with pytest.raises((ValueError, TypeError)):
        OracleDynamics(2, wind=["0*x", "0*x", "0*x"])


# -- GPU: the engine ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("integ", list(WIND_KW))
def test_engine_wind_functions_match_reference_golden(integ):
    from visfly_b200.dynamics import Dynamics
    z = np.load(GOLD)
    make = lambda n, fn: Dynamics(num=n, wind_settings=fn, device="cuda", **WIND_KW[integ])
    states, winds = wind_rollout(make, z, th.float32, dev="cuda")
    assert rel_l2(winds, z[f"windfn_{integ}_wind_f32"]) < 2e-6
    err, floor = rel_l2(states, z[f"windfn_{integ}_states_f64"]), rel_l2(z[f"windfn_{integ}_states_f32"],
                                                                          z[f"windfn_{integ}_states_f64"])
    assert err < max(1e-5, 2 * floor), (err, floor)
    # a constant-wind engine fed the same actions ends somewhere else: the per-agent wind really reached the kernel
    still = wind_rollout(lambda n, fn: Dynamics(num=n, device="cuda", **WIND_KW[integ]), z, th.float32, dev="cuda")[0]
    assert rel_l2(still, z[f"windfn_{integ}_states_f64"]) > 1e-3


@pytest.mark.gpu
def test_engine_wind_function_gradient_matches_autograd_through_the_oracle():
    from visfly_b200.dynamics import ControlStep, Dynamics
    from _util import oracle_grads, pack, random_flight_state
    n, dt = 128, 0.0025
    dyn = Dynamics(num=n, action_type="bodyrate", dt=dt, ctrl_dt=0.02, integrator="rk4", comm_delay=0.0, device="cuda")
    g = th.Generator().manual_seed(5)
    fields = list(random_flight_state(n, seed=3))
    fields[0][:, 2] = 19.99                                   # next to the z clamp: the gate depends on the wind
    packed = pack(*fields)
    action = th.rand(n, 4, generator=g) * 2 - 1
    wind = th.randn(n, 3, generator=g) * 2
    g_out, g_obs = th.randn(5, n, 4, generator=g), th.randn(n, 13, generator=g)
    st, ac = packed.cuda().requires_grad_(True), action.cuda().requires_grad_(True)
    wind4 = th.cat([wind, th.zeros(n, 1)], 1).cuda().contiguous()
    out, obs, _ = ControlStep.apply(st, ac, None, dyn._cfg, wind4)
    ((out * g_out.cuda()).sum() + (obs * g_obs.cuda()).sum()).backward()
    orc = OracleDynamics(n, "bodyrate", dt=dt, ctrl_dt=0.02, integrator="rk4", comm_delay=0.0, dtype=th.float64)
    orc.wind = wind.double().T.contiguous()
    ref_gs, ref_ga = oracle_grads(orc, packed.double(), action.double(), g_out.double(), g_obs.double())
    assert rel_l2(st.grad.cpu(), ref_gs) < 1e-4 and rel_l2(ac.grad.cpu(), ref_ga) < 1e-4
    orc.load_packed(packed.double())
    ref_obs = orc.step(action.double())
    assert rel_l2(obs.detach().cpu(), ref_obs) < 1e-5 and rel_l2(out.detach().cpu(), orc.packed()) < 1e-5


@pytest.mark.gpu
def test_engine_drag_random_matches_reference_golden():
    from visfly_b200.dynamics import Dynamics
    z = np.load(GOLD)
    acts = th.from_numpy(z["drag_actions"]).cuda()
    d = Dynamics(num=acts.shape[1], drag_random=0.3, device="cuda", **DRAG_KW)
    cv = lambda k: th.from_numpy(z[k]).float()
    th.manual_seed(int(z["drag_seed"]))
    d.reset(pos=cv("drag_init_pos"), ori=cv("drag_init_quat"), vel=cv("drag_init_vel"), ori_vel=cv("drag_init_rate"))
    np.testing.assert_array_equal(np.array(d._params.k_lin[:], dtype=np.float32), z["drag_k_lin_f32"][:, 0])
    np.testing.assert_array_equal(np.array(d._params.k_quad[:], dtype=np.float32), z["drag_k_quad_f32"][:, 0])
    states = th.stack([d.step(acts[t]).clone() for t in range(acts.shape[0])]).cpu()
    err, floor = rel_l2(states, z["drag_states_f64"]), rel_l2(z["drag_states_f32"], z["drag_states_f64"])
    assert err < max(1e-5, 2 * floor), (err, floor)


@pytest.mark.gpu
def test_env_with_wind_functions_runs_the_one_kernel_path_and_matches_the_generic_one():
    """Per-agent wind (the reference's wind functions, dynamics.py:136-165,384-388) inside the fused env step: the
    wind is evaluated with tensor ops from the per-agent time, the kernel takes it as a (N,4) vector.  Same env through
    the generic tensor-op path must give the same observations / rewards / dones, across an auto-reset; the gradient of
    a short rollout agrees as well."""
    from visfly_b200.envs import HoverEnv, NavigationEnv
    fn = ["0.5*th.sin(2*x)", "0.3*th.cos(x)", "0*x", "0.9*y+0.05", "0.8*y-0.02", "0*y"]
    n = 64
    g = th.Generator().manual_seed(11)
    pos = th.tensor([1.0, 0.0, 1.5]) + (th.rand(n, 3, generator=g) * 2 - 1) * 0.5
    quat = th.tensor([[1.0, 0, 0, 0]]).repeat(n, 1)
    acts = ((th.rand(12, n, 4, generator=g) * 2 - 1) * 0.2).cuda()
    acts[..., 0] -= 1.0 / 3.0
    runs = {}
    for fused in (True, False):
        env = HoverEnv(num_agent_per_scene=n, visual=False, device="cuda", max_episode_steps=8, tensor_output=True,
                       dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02,
                                            wind_settings=fn))
        env.envs.set_reset_table(pos.cuda(), quat.cuda())
        env.use_fused_step = fused
        env.reset()
        out = []
        for t in range(12):                                   # crosses an auto-reset at step 8
            obs, reward, done, info = env.step(acts[t])
            out.append((obs["state"].clone(), reward.clone(), done.clone()))
        assert (env._fused is not None and env._fused.active) == fused
        assert env.envs.dynamics.wind_velocity.shape == (3, n)
        assert float(env.envs.dynamics.wind_velocity.abs().max()) > 0.1
        runs[fused] = out
    for (o1, r1, d1), (o2, r2, d2) in zip(runs[True], runs[False]):
        assert rel_l2(o1.cpu(), o2.cpu()) < 1e-5 and th.allclose(r1, r2, atol=1e-5) and th.equal(d1, d2)
    grads = {}
    for fused in (True, False):
        env = NavigationEnv(num_agent_per_scene=n, visual=False, device="cuda", max_episode_steps=5, requires_grad=True,
                            dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02,
                                                 wind_settings=fn))
        env.envs.set_reset_table(pos.cuda(), quat.cuda())
        env.use_fused_step = fused
        env.reset()
        a = acts[:8].clone().requires_grad_(True)
        loss = 0.0
        for t in range(8):
            obs, reward, done, info = env.step(a[t])
            loss = loss - reward.mean() + 1e-3 * obs["state"].pow(2).mean()
        grads[fused], = th.autograd.grad(loss, a)
    assert rel_l2(grads[True].cpu(), grads[False].cpu()) < 1e-4
