"""Parity of the sm_100a kernels (called through the C-ABI) with the oracle and with the golden vectors recorded
from the real reference.  Tolerances are the north-star's: state trajectories within 1e-5 rel-L2 of the
reference, analytic gradients within 1e-4 of torch.autograd (both are met with >10x margin)."""
import os

import numpy as np
import pytest
import torch as th

from _util import (ACTION, FLAG_CTRL_DELAY, INTEGRATOR, make_oracle, oracle_grads, oracle_step_packed, pack,
                   random_flight_state, rel_l2, unpack, vf_params)

pytestmark = pytest.mark.gpu

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
TRAJ_TOL = 1e-5      # north-star: state trajectories within 1e-5 rel-L2 of the reference
GRAD_TOL = 1e-4      # north-star: analytic gradients within 1e-4 of PyTorch autograd


def cuda_fwd(P, packed, action, S, integ, at="bodyrate", lag=True, obs=True, ext=False):
    from visfly_b200 import _lib
    st = packed.cuda().contiguous()
    ac = action.cuda().float().contiguous()
    out = th.empty_like(st)
    n = st.shape[1]
    o = th.empty((n, 13), device="cuda") if obs else None
    e = th.empty((n, 8), device="cuda") if ext else None
    _lib.step_fwd(P, S, INTEGRATOR[integ], ACTION[at], FLAG_CTRL_DELAY if lag else 0, st, ac, out, o, e)
    th.cuda.synchronize()
    return out.cpu(), (None if o is None else o.cpu()), (None if e is None else e.cpu())


def cuda_bwd(P, packed, action, g_out, g_obs, S, integ, at="bodyrate", lag=True):
    from visfly_b200 import _lib
    st, ac = packed.cuda().contiguous(), action.cuda().float().contiguous()
    go = None if g_out is None else g_out.cuda().float().contiguous()
    gb = None if g_obs is None else g_obs.cuda().float().contiguous()
    gs, ga = th.empty_like(st), th.empty_like(ac)
    _lib.step_bwd(P, S, INTEGRATOR[integ], ACTION[at], FLAG_CTRL_DELAY if lag else 0, st, ac, go, gb, gs, ga)
    th.cuda.synchronize()
    return gs.cpu(), ga.cpu()


def make_dynamics(n, **kw):
    from visfly_b200.dynamics import Dynamics
    return Dynamics(num=n, device="cuda", **kw)


# -- one step, seeded inputs, all kernel variants ---------------------------------------------------------
@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
@pytest.mark.parametrize("at", ["bodyrate", "thrust", "velocity", "position"])
@pytest.mark.parametrize("lag", [True, False])
def test_forward_one_step_vs_oracle(integ, dt, at, lag):
    # 1000: ragged last warp and last CTA (velocity / position: the oracle keeps the reference's per-agent loop)
    n, S, wind = 1000 if at in ("bodyrate", "thrust") else 200, int(0.02 / dt), (0.3, -0.2, 0.1)
    P = vf_params(at, dt, wind=wind)
    packed = pack(*random_flight_state(n, seed=3))
    g = th.Generator().manual_seed(5)
    action = th.rand(n, 4, generator=g) * 2 - 1
    o32 = make_oracle(n, at, integ, dt, ctrl_delay=lag, wind=wind)
    o64 = make_oracle(n, at, integ, dt, ctrl_delay=lag, wind=wind, dtype=th.float64)
    r32, robs32 = oracle_step_packed(o32, packed, action)
    r64, robs64 = oracle_step_packed(o64, packed.double(), action.double())
    out, obs, ext = cuda_fwd(P, packed, action, S, integ, at, lag, ext=True)
    floor = max(rel_l2(r32, r64), 1e-7)
    assert rel_l2(out, r64) < 3 * floor
    assert rel_l2(obs, robs64) < 3 * max(rel_l2(robs32, robs64), 1e-7)
    for got, ref in zip(unpack(out), unpack(r64)):
        assert rel_l2(got, ref) < 2e-6
    assert rel_l2(ext[:, :3], o64.acceleration) < 1e-5
    assert rel_l2(ext[:, 4:], o64.thrusts.T) < 1e-6
    # the observation is the packed state re-laid out (+ wind on the velocity)
    assert th.equal(obs[:, 0:3], out[0, :, :3]) and th.equal(obs[:, 3:7], out[1])
    assert th.equal(obs[:, 10:13], out[3, :, :3])
    assert th.allclose(obs[:, 7:10], out[2, :, :3] + th.tensor(wind), atol=1e-6)


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_forward_one_step_vs_reference_golden(cfg):
    kw = CONFIGS[cfg]
    z = np.load(os.path.join(GOLD, f"step_{cfg}.npz"))
    P = vf_params(kw["action_type"], kw["dt"])
    S = int(kw["ctrl_dt"] / kw["dt"])
    before = [th.from_numpy(z[f"before_{k}_f32"]) for k in ("pos", "quat", "vel", "rate", "motor", "alpha")]
    out, obs, ext = cuda_fwd(P, pack(*before), th.from_numpy(z["action"]), S, kw["integrator"], kw["action_type"],
                             kw["ctrl_delay"], ext=True)
    assert rel_l2(obs, z["obs_f32"]) < 1e-6
    if "obs_f64" in z.files:      # float64 reference as arbiter: as close to it as the float32 reference is
        assert rel_l2(obs, z["obs_f64"]) < max(2e-7, 3 * rel_l2(z["obs_f32"], z["obs_f64"]))
    for k, got in zip(("pos", "quat", "vel", "rate", "motor", "alpha"), unpack(out)):
        assert rel_l2(got, z[f"after_{k}_f32"]) < 2e-6, k
    assert rel_l2(ext[:, :3], z["acc_f32"]) < 1e-5 and rel_l2(ext[:, 4:], z["thrusts_f32"]) < 1e-6


# -- trajectories through the drop-in Dynamics class ---------------------------------------------------------
def test_known_answer_vectors():
    z = np.load(os.path.join(GOLD, "kat.npz"))
    a = th.from_numpy(z["action"]).cuda()
    for name, kw, steps in [("B1_euler", dict(dt=0.005, integrator="euler", comm_delay=0.0), 3),
                            ("B2_euler_fifo3", dict(dt=0.005, integrator="euler", comm_delay=0.06), 4),
                            ("B3_rk4", dict(dt=0.0025, integrator="rk4", comm_delay=0.0), 3)]:
        d = make_dynamics(1, action_type="bodyrate", ctrl_dt=0.02, **kw)
        d.reset(pos=th.tensor([[1.0, 0.0, 1.5]]))
        for k in range(steps):
            s = d.step(a)
            fs = z[name + "_full_state"][k]
            assert s.shape == (1, 13)
            np.testing.assert_allclose(s.cpu().numpy()[0], fs[:13], rtol=2e-5, atol=2e-7)
            np.testing.assert_allclose(d.full_state.cpu().numpy()[0], fs, rtol=2e-5, atol=2e-7)
            np.testing.assert_allclose(d.angular_acceleration.cpu().numpy()[0], z[name + "_alpha"][k], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("law", ["uniform", "hover"])
def test_trajectory_vs_reference_golden(cfg, law):
    kw = CONFIGS[cfg]
    z = np.load(os.path.join(GOLD, f"traj_{cfg}.npz"))
    n = z["init_pos"].shape[0]
    d = make_dynamics(n, **kw)
    d.reset(pos=th.from_numpy(z["init_pos"]), ori=th.from_numpy(z["init_quat"]), vel=th.from_numpy(z["init_vel"]),
            ori_vel=th.from_numpy(z["init_rate"]))
    acts = th.from_numpy(z[f"actions_{law}"]).cuda()
    states = th.stack([d.step(acts[t]).clone() for t in range(acts.shape[0])]).cpu()
    ref32 = z[f"states_{law}_f32"]
    assert rel_l2(states, ref32) < TRAJ_TOL
    if f"states_{law}_f64" in z.files:          # float64 reference arbitrates: we are as close to it as float32 gets
        ref64 = z[f"states_{law}_f64"]
        assert rel_l2(states, ref64) < max(TRAJ_TOL, 2 * rel_l2(ref32, ref64))
    fs = d.full_state.cpu()
    assert rel_l2(fs[:, :21], z[f"final_full_state_{law}_f32"][:, :21]) < TRAJ_TOL
    assert th.allclose(fs[:, 21], th.from_numpy(z[f"final_full_state_{law}_f32"][:, 21]), atol=1e-5)


@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
def test_256_step_trajectory_vs_oracle(integ, dt):
    """256 control steps, N=256, both action laws; 1e-5 rel-L2 with the float64 oracle as arbiter."""
    n, T = 256, 256
    init = random_flight_state(n, seed=40, spread=0.3)
    for law_seed, hover in ((0, False), (1, True)):
        g = th.Generator().manual_seed(law_seed)
        acts = th.rand(T, n, 4, generator=g) * 2 - 1
        if hover:
            acts = acts * 0.1
            acts[..., 0] += -1.0 / 3.0
        d = make_dynamics(n, action_type="bodyrate", integrator=integ, dt=dt, ctrl_dt=0.02)
        o32 = make_oracle(n, "bodyrate", integ, dt, comm_delay=0.06)
        o64 = make_oracle(n, "bodyrate", integ, dt, comm_delay=0.06, dtype=th.float64)
        for obj in (d, o32, o64):
            obj.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
        acts_gpu = acts.cuda()
        got = th.stack([d.step(acts_gpu[t]).clone() for t in range(T)]).cpu()
        r32 = th.stack([o32.step(acts[t]).clone() for t in range(T)])
        r64 = th.stack([o64.step(acts[t].double()).clone() for t in range(T)])
        assert rel_l2(got, r32) < TRAJ_TOL
        assert rel_l2(got, r64) < max(TRAJ_TOL, 2 * rel_l2(r32, r64))


# -- adjoint ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("integ,dt,ctrl_dt", [("euler", 0.005, 0.02), ("rk4", 0.0025, 0.02), ("rk4", 0.0025, 0.03),
                                              ("euler", 0.001, 0.04)])
@pytest.mark.parametrize("at", ["bodyrate", "thrust"])
@pytest.mark.parametrize("lag", [True, False])
def test_backward_one_step_vs_autograd(integ, dt, ctrl_dt, at, lag):
    n, S, wind = 777, int(round(ctrl_dt / dt)), (0.3, -0.2, 0.1)
    P = vf_params(at, dt, wind=wind)
    packed = pack(*random_flight_state(n, seed=8))
    g = th.Generator().manual_seed(9)
    action = th.rand(n, 4, generator=g) * 2 - 1
    g_out, g_obs = th.randn(5, n, 4, generator=g), th.randn(n, 13, generator=g)
    # (the oracle repeats the reference's float32 `ctrl_dt % dt` check, which rejects e.g. 0.04/0.001:
    #  build it with one sub-step and set the count directly)
    orc = make_oracle(n, at, integ, dt, ctrl_dt=dt, ctrl_delay=lag, wind=wind, dtype=th.float64)
    orc.substeps = S
    ref_gs, ref_ga = oracle_grads(orc, packed.double(), action.double(), g_out.double(), g_obs.double())
    got_gs, got_ga = cuda_bwd(P, packed, action, g_out, g_obs, S, integ, at, lag)
    assert rel_l2(got_gs, ref_gs) < GRAD_TOL and rel_l2(got_ga, ref_ga) < GRAD_TOL
    for got, ref in zip(unpack(got_gs), unpack(ref_gs)):
        assert rel_l2(got, ref) < GRAD_TOL
    # NULL gradient inputs mean zeros
    z_gs, z_ga = cuda_bwd(P, packed, action, None, None, S, integ, at, lag)
    assert float(z_gs.abs().max()) == 0.0 and float(z_ga.abs().max()) == 0.0
    o_gs, o_ga = cuda_bwd(P, packed, action, None, g_obs, S, integ, at, lag)
    r_gs, r_ga = oracle_grads(orc, packed.double(), action.double(), None, g_obs.double())
    assert rel_l2(o_gs, r_gs) < GRAD_TOL and rel_l2(o_ga, r_ga) < GRAD_TOL


def _hover_loss(dyn, acts, gamma=0.99):
    dev, dt = acts.device, acts.dtype
    target = th.tensor([[1.0, 0.0, 1.5]], dtype=dt, device=dev)
    one = th.tensor([1.0, 0, 0, 0], dtype=dt, device=dev)
    total = 0.0
    for t in range(acts.shape[0]):
        dyn.step(acts[t])
        r = 0.1 - (dyn.position - target).norm(dim=1) / 90 - (dyn.orientation - one).norm(dim=1) * 1e-5 \
            - (dyn.velocity - 0).norm(dim=1) * 0.002 - (dyn.angular_velocity - 0).norm(dim=1) * 0.002
        total = total + (gamma ** t) * r
    return -total.mean()


@pytest.mark.parametrize("at", ["velocity", "position"])
def test_velocity_and_position_have_no_gradient_like_the_reference(at):
    """The reference's backward raises for these action types (in-place writes in its per-agent loop,
    dynamics.py:446-450); the engine runs them forward and refuses the adjoint with a clear error."""
    d = make_dynamics(8, action_type=at, integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0)
    a = th.zeros(8, 4, device="cuda", requires_grad=True)
    s = d.step(a)
    assert s.shape == (8, 13) and bool(th.isfinite(s).all())
    with pytest.raises(RuntimeError, match="no gradient for the velocity / position"):
        s.sum().backward()


@pytest.mark.parametrize("cfg", [c for c in CONFIGS if c not in FWD_ONLY])
def test_rollout_gradients_vs_reference_autograd_golden(cfg):
    """8-step BPTT through the drop-in Dynamics + autograd.Function against gradients recorded from
    torch.autograd through the real reference (tests/golden/grad_*.npz)."""
    from visfly_b200.dynamics import ControlStep
    kw = CONFIGS[cfg]
    z = np.load(os.path.join(GOLD, f"grad_{cfg}.npz"))
    n = z["init_pos"].shape[0]
    d = make_dynamics(n, **kw)
    leaves = [th.from_numpy(z["init_" + k]).float().cuda().requires_grad_(True) for k in ("pos", "quat", "vel", "rate")]
    acts = th.from_numpy(z["actions"]).cuda().requires_grad_(True)
    d.reset()
    zero = th.zeros((n, 1), device="cuda")
    motor = d.motor_omega.detach()
    d._state = th.stack([th.cat([leaves[0], zero], 1), leaves[1], th.cat([leaves[2], zero], 1),
                         th.cat([leaves[3], zero], 1), motor])
    loss = _hover_loss(d, acts)
    grads = th.autograd.grad(loss, [acts] + leaves)
    tag = "f64" if "grad_actions_f64" in z.files else "f32"
    assert abs(loss.item() - float(z[f"loss_{tag}"])) < 1e-5
    for k, g in zip(("actions", "pos", "quat", "vel", "rate"), grads):
        assert rel_l2(g.cpu(), z[f"grad_{k}_{tag}"]) < GRAD_TOL, k
        assert rel_l2(g.cpu(), z[f"grad_{k}_f32"]) < GRAD_TOL, k


def test_apg_horizon_gradients_vs_oracle_autograd():
    """Config 3 at a parity-sized batch: H=32 RK4 steps with the comm-delay FIFO, gradients of a discounted
    return w.r.t. all actions and the initial state, vs torch.autograd through the oracle (float32 and float64)."""
    n, H, dt = 512, 32, 0.0025
    init = random_flight_state(n, seed=50, spread=0.3)
    g = th.Generator().manual_seed(51)
    acts = (th.rand(H, n, 4, generator=g) * 2 - 1) * 0.5
    d = make_dynamics(n, action_type="bodyrate", integrator="rk4", dt=dt, ctrl_dt=0.02)
    d.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
    a_gpu = acts.cuda().requires_grad_(True)
    s0 = d._state.detach().clone().requires_grad_(True)
    d._state = s0
    loss = _hover_loss(d, a_gpu)
    g_a, g_s = th.autograd.grad(loss, [a_gpu, s0])
    res = {}
    for dtype in (th.float32, th.float64):
        o = make_oracle(n, "bodyrate", "rk4", dt, comm_delay=0.06, dtype=dtype)
        o.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
        a = acts.to(dtype).requires_grad_(True)
        p0 = o.packed().clone().requires_grad_(True)
        o.load_packed(p0)
        lo = _hover_loss(o, a)
        res[dtype] = th.autograd.grad(lo, [a, p0]) + (lo,)
    assert abs(loss.item() - res[th.float64][2].item()) < 1e-5
    # actions still in the FIFO at the end of the horizon get exactly zero gradient
    assert float(g_a[-3:].abs().max()) == 0.0 and float(res[th.float64][0][-3:].abs().max()) == 0.0
    assert rel_l2(g_a.cpu(), res[th.float64][0]) < GRAD_TOL
    assert rel_l2(g_s.cpu(), res[th.float64][1]) < GRAD_TOL
    assert rel_l2(g_a.cpu(), res[th.float32][0]) < GRAD_TOL


# -- edge cases and size-independent properties --------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 33, 64, 65])
def test_ragged_and_tiny_batches(n):
    from visfly_b200 import _lib
    P = vf_params("bodyrate", 0.0025)
    if n == 0:
        e = th.empty((5, 0, 4), device="cuda")
        _lib.step_fwd(P, 8, 1, 1, 1, e, th.empty((0, 4), device="cuda"), th.empty((5, 0, 4), device="cuda"), None, None)
        return
    packed = pack(*random_flight_state(n, seed=n))
    g = th.Generator().manual_seed(n)
    action = th.rand(n, 4, generator=g) * 2 - 1
    out, obs, _ = cuda_fwd(P, packed, action, 8, "rk4")
    ref, robs = oracle_step_packed(make_oracle(n, "bodyrate", "rk4", 0.0025, dtype=th.float64), packed.double(), action.double())
    assert rel_l2(out, ref) < 1e-6 and rel_l2(obs, robs) < 1e-6
    g_obs = th.randn(n, 13, generator=g)
    gs, ga = cuda_bwd(P, packed, action, None, g_obs, 8, "rk4")
    rgs, rga = oracle_grads(make_oracle(n, "bodyrate", "rk4", 0.0025, dtype=th.float64), packed.double(),
                            action.double(), None, g_obs.double())
    assert rel_l2(gs, rgs) < GRAD_TOL and rel_l2(ga, rga) < GRAD_TOL


def test_full_size_properties_and_shard_invariance():
    """BASELINE config 2 size (65 536 agents, RK4 x8): unit quaternions, determinism, and bitwise identical
    results whether the batch is stepped whole or as two shards (per-agent arithmetic only)."""
    n, dt, S = 65536, 0.0025, 8
    P = vf_params("bodyrate", dt)
    packed = pack(*random_flight_state(n, seed=77))
    g = th.Generator().manual_seed(78)
    for _ in range(4):
        action = th.rand(n, 4, generator=g) * 2 - 1
        out, obs, _ = cuda_fwd(P, packed, action, S, "rk4")
        out2, obs2, _ = cuda_fwd(P, packed, action, S, "rk4")
        assert th.equal(out, out2) and th.equal(obs, obs2)
        h = n // 2
        lo, lo_obs, _ = cuda_fwd(P, packed[:, :h].contiguous(), action[:h], S, "rk4")
        hi, hi_obs, _ = cuda_fwd(P, packed[:, h:].contiguous(), action[h:], S, "rk4")
        assert th.equal(th.cat([lo, hi], 1), out) and th.equal(th.cat([lo_obs, hi_obs]), obs)
        assert float((out[1].norm(dim=1) - 1).abs().max()) < 1e-6
        assert bool(th.isfinite(out).all())
        packed = out
    # spot-check a slice of the big batch against the oracle
    sl = slice(1000, 1256)
    ref, _ = oracle_step_packed(make_oracle(256, "bodyrate", "rk4", dt, dtype=th.float64),
                                packed[:, sl].double(), action[sl].double())
    got, _, _ = cuda_fwd(P, packed, action, S, "rk4")
    assert rel_l2(got[:, sl], ref) < 1e-6


def test_post_step_clamps_and_gradient_gates():
    n, dt, S = 64, 0.005, 4
    P = vf_params("bodyrate", dt)
    pos, quat, vel, rate, motor, alpha = random_flight_state(n, seed=4)
    pos[:, 2], vel[:, 0], rate[:, 1] = 25.0, 30.0, -40.0
    packed = pack(pos, quat, vel, rate, motor, alpha)
    action = th.ones(n, 4)
    out, obs, _ = cuda_fwd(P, packed, action, S, "euler")
    assert float(out[0, :, 2].max()) <= 20.0 and float(out[2, :, 0].max()) <= 20.0 and float(out[3, :, 1].min()) >= -10.0
    g_out = th.zeros(5, n, 4)
    g_out[0, :, 2] = 1.0
    g_out[2, :, 0] = 1.0
    gs, ga = cuda_bwd(P, packed, action, g_out, None, S, "euler")
    rgs, rga = oracle_grads(make_oracle(n, "bodyrate", "euler", dt, dtype=th.float64), packed.double(),
                            action.double(), g_out.double(), None)
    assert float(rgs.abs().max()) == 0.0 and float(gs.abs().max()) == 0.0 and float(ga.abs().max()) == 0.0


def test_pack_unpack_roundtrip_and_scatter():
    from visfly_b200 import _lib
    n = 300
    fields = [f.cuda().contiguous() for f in random_flight_state(n, seed=5)]
    st = th.zeros((5, n, 4), device="cuda")
    _lib.pack_state(n, st, None, *fields)
    assert th.equal(st.cpu(), pack(*[f.cpu() for f in fields]))
    outs = [th.empty_like(f) for f in fields]
    _lib.unpack_state(n, st, *outs)
    for a, b in zip(outs, fields):
        assert th.equal(a, b)
    idx = th.tensor([5, 17, 299, 0], device="cuda")
    rows = [f[:4].contiguous() * 2 for f in fields]
    _lib.pack_state(n, st, idx, *rows)
    ref = pack(*[f.cpu() for f in fields])
    ref[:, idx.cpu()] = pack(*[r.cpu() for r in rows])
    assert th.equal(st.cpu(), ref)


def test_cuda_graph_capture_replays_the_step():
    from visfly_b200 import _lib
    n, S = 4096, 8
    P = vf_params("bodyrate", 0.0025)
    st = pack(*random_flight_state(n, seed=6)).cuda()
    ac = (th.rand(n, 4) * 2 - 1).cuda()
    out, obs = th.empty_like(st), th.empty((n, 13), device="cuda")
    _lib.step_fwd(P, S, 1, 1, 1, st, ac, out, obs, None)
    eager = out.clone()
    th.cuda.synchronize()
    graph = th.cuda.CUDAGraph()
    out.zero_()
    with th.cuda.graph(graph):
        _lib.step_fwd(P, S, 1, 1, 1, st, ac, out, obs, None)
    graph.replay()
    th.cuda.synchronize()
    assert th.equal(out, eager)


def test_host_buffer_entry_point():
    from visfly_b200 import _lib
    n, S = 5000, 8
    P = vf_params("bodyrate", 0.0025)
    st = pack(*random_flight_state(n, seed=7)).cuda()
    a_host = (th.rand(n, 4) * 2 - 1).pin_memory()
    o_host = th.empty((n, 13)).pin_memory()
    a_dev, o_dev, out = th.empty((n, 4), device="cuda"), th.empty((n, 13), device="cuda"), th.empty_like(st)
    _lib.step_fwd_host(P, S, 1, 1, 1, st, a_host, a_dev, out, o_dev, o_host)
    ref_out, ref_obs, _ = cuda_fwd(P, st.cpu(), a_host, S, "rk4")
    assert th.equal(out.cpu(), ref_out) and th.equal(o_host, ref_obs)


# -- Dynamics surface -------------------------------------------------------------------------------------------
def test_dynamics_surface_shapes_and_reset_semantics():
    from visfly_b200.type import ACTION_TYPE
    n = 128
    d = make_dynamics(n, action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02)
    assert d.action_type == ACTION_TYPE.BODYRATE and d.num == n and d.is_quat_output
    assert d.state.shape == (n, 13) and d.full_state.shape == (n, 22) and d.extend_state.shape == (n, 28)
    assert d.R.shape == (3, 3, n) and d.xz_axis.shape == (2, 3, n) and d.direction.shape == (n, 3)
    assert d._orientation.toTensor().shape == (4, n)
    o = make_oracle(n, "bodyrate", "rk4", 0.0025, comm_delay=0.06)
    init = random_flight_state(n, seed=60, spread=0.3)
    d.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
    o.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
    assert rel_l2(d.full_state.cpu(), o.full_state) < 1e-6
    g = th.Generator().manual_seed(61)
    for t in range(10):
        a = th.rand(n, 4, generator=g) * 2 - 1
        d.step(a.cuda()), o.step(a)
        if t == 4:                      # partial reset in the middle (auto-reset path of the env wrapper)
            idx = [3, 50, 127]
            newp = th.tensor([[0.0, 0, 1], [1, 1, 1], [2, 2, 2.0]])
            d.reset(pos=newp, indices=idx)
            o.reset(pos=newp, indices=idx)
            assert float(d.t[idx].abs().max()) == 0.0
    assert rel_l2(d.full_state.cpu()[:, :21], o.full_state[:, :21]) < 1e-5
    assert rel_l2(d.acceleration.cpu(), o.acceleration) < 1e-4
    assert rel_l2(d.direction.cpu(), o.direction) < 1e-5
    assert rel_l2(d.angular_acceleration.cpu(), o.angular_acceleration) < 1e-4
    assert th.allclose(d.t.cpu(), o.t, atol=1e-6)
    with pytest.raises(ValueError):
        make_dynamics(4, dt=0.003, ctrl_dt=0.02)


def test_partial_reset_cuts_the_gradient_like_the_reference():
    n, H = 64, 6
    d = make_dynamics(n, action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0)
    o = make_oracle(n, "bodyrate", "euler", 0.005, comm_delay=0.0, dtype=th.float64)
    init = random_flight_state(n, seed=70, spread=0.3)
    d.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
    o.reset(pos=init[0], ori=init[1], vel=init[2], ori_vel=init[3])
    g = th.Generator().manual_seed(71)
    acts = th.rand(H, n, 4, generator=g) * 2 - 1
    a_gpu, a_cpu = acts.cuda().requires_grad_(True), acts.double().requires_grad_(True)
    idx = list(range(0, n, 2))
    lg = lc = 0.0
    for t in range(H):
        d.step(a_gpu[t]), o.step(a_cpu[t])
        lg = lg + d.position.norm(dim=1).sum() + d.angular_velocity.pow(2).sum()
        lc = lc + o.position.norm(dim=1).sum() + o.angular_velocity.pow(2).sum()
        if t == 2:
            d.reset(indices=idx), o.reset(indices=idx)
    gg, = th.autograd.grad(lg, a_gpu)
    gc, = th.autograd.grad(lc, a_cpu)
    assert rel_l2(gg.cpu(), gc) < GRAD_TOL
    d.detach()
    assert not d._state.requires_grad and all(not a.requires_grad for a in d._pre_action)


def test_habitat_pose_export_matches_reference_transform():
    """Renderer hand-off: poses in Habitat's frame written by one kernel into page-locked host
    memory / device memory == the reference's std_to_habitat (utils/common.py:131-179) on Dynamics.position /
    orientation / velocity — exact, it is a signed axis permutation."""
    from visfly_b200.render_handoff import HabitatPoseExporter, habitat_to_std, std_to_habitat
    for n in (1, 31, 257, 4096):
        d = make_dynamics(n, action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0,
                          wind_settings=[0.3, -0.2, 0.1])
        pos, quat, vel, rate, _, _ = random_flight_state(n, seed=n)
        d.reset(pos=pos, ori=quat, vel=vel, ori_vel=rate)
        d.step(th.zeros(n, 4, device="cuda"))
        # the reference's matrices, applied with numpy exactly as it does
        ref_pos = d.position.cpu().numpy() @ np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]])
        ref_ori = d.orientation.cpu().numpy() @ np.array([[1, 0, 0, 0], [0, 0, 0, -1], [0, -1, 0, 0], [0, 0, 1, 0]])
        ref_vel = d.velocity.cpu().numpy() @ np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]])
        for host in (True, False):
            pose, hv = HabitatPoseExporter(d, host=host).export()
            pose, hv = (pose, hv) if host else (pose.cpu().numpy(), hv.cpu().numpy())
            assert np.array_equal(pose[:, :3], ref_pos.astype(np.float32))
            assert np.array_equal(pose[:, 3:], ref_ori.astype(np.float32))
            assert np.array_equal(hv, ref_vel.astype(np.float32))
        hp, ho = std_to_habitat(d.position, d.orientation)
        assert np.array_equal(hp, ref_pos.astype(np.float32)) and np.array_equal(ho, ref_ori.astype(np.float32))
        sp, so = habitat_to_std(hp, ho)                      # round trip
        assert th.equal(sp, d.position.cpu()) and th.equal(so, d.orientation.cpu())


def test_comm_delay_fifo_owns_a_copy_of_every_action_like_the_reference():
    """The reference clones every action into its FIFO (`action.T.clone()`, dynamics.py:323-328), so a caller may
    refill ONE preallocated action buffer in place every step.  Here the step launch itself makes that copy
    (`fifo_push` -> `fifo_copy`): the in-place-reuse loop must fly exactly the trajectory of fresh tensors and of the
    reference golden, through `Dynamics.step`, the fused env step and the generic env path; gradients reach the
    action through the copy, and actions still waiting in the FIFO at the end of a rollout get exactly zero."""
    from _util import make_oracle
    n = 64
    kw = dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06)
    g = th.Generator().manual_seed(3)
    acts = (th.rand(8, n, 4, generator=g) * 2 - 1).cuda()
    ok = make_dynamics(n, **kw)
    ref = th.stack([ok.step(acts[t]).clone() for t in range(8)])          # fresh tensors
    orc = make_oracle(n, **kw)                                             # the reference's own FIFO (clone per step)
    ref_cpu = th.stack([orc.step(acts[t].cpu()).clone() for t in range(8)])
    assert rel_l2(ref.cpu(), ref_cpu) < TRAJ_TOL
    reuse = make_dynamics(n, **kw)
    buf = th.empty(n, 4, device="cuda")
    out = []
    for t in range(8):
        buf.copy_(acts[t])                                                  # legal with the reference: same buffer
        out.append(reuse.step(buf).clone())
        assert all(a.data_ptr() != buf.data_ptr() for a in reuse._pre_action)
    assert th.equal(th.stack(out), ref)
    # host actions (numpy) are converted on the way in: already private, no second copy
    host = make_dynamics(n, **kw)
    out = [host.step(acts[t].cpu().numpy()).clone() for t in range(8)]
    assert th.equal(th.stack(out), ref)
    from visfly_b200.envs import HoverEnv
    runs = {}
    for fused in (True, False):
        for mode in ("fresh", "reuse"):
            env = HoverEnv(num_agent_per_scene=n, visual=False, device="cuda", tensor_output=True, seed=5,
                           dynamics_kwargs=dict(kw))
            env.use_fused_step = fused
            env.reset()
            states = []
            for t in range(8):
                if mode == "reuse":
                    buf.copy_(acts[t])
                states.append(env.step(buf if mode == "reuse" else acts[t])[0]["state"].clone())
            assert env._fused.active == fused
            runs[fused, mode] = th.stack(states)
        assert th.equal(runs[fused, "fresh"], runs[fused, "reuse"])
    assert rel_l2(runs[True, "fresh"].cpu(), runs[False, "fresh"].cpu()) < 1e-6
    # gradients: through the copy to the action that was pushed; nothing for actions that never left the FIFO
    d = make_dynamics(n, **kw)
    a = acts.clone().requires_grad_(True)
    loss = 0.0
    for t in range(8):
        loss = loss + d.step(a[t]).pow(2).sum()
    ga, = th.autograd.grad(loss, a)
    assert float(ga[:5].abs().sum()) > 0 and float(ga[5:].abs().sum()) == 0.0       # depth 3: a[5:] were never flown
    o = make_oracle(n, dtype=th.float64, **kw)
    a64 = acts.cpu().double().requires_grad_(True)
    lo = 0.0
    for t in range(8):
        lo = lo + o.step(a64[t]).pow(2).sum()
    go, = th.autograd.grad(lo, a64)
    assert rel_l2(ga.cpu(), go) < GRAD_TOL


def test_fifo_ring_entry_point_flies_the_list_fifo_trajectory_and_finish_zeroes_rows():
    """``vf_step_fwd_ring`` (the comm-delay FIFO as a device-resident ring, shifted in place by the step launch) against
    ``Dynamics.step`` with its list-of-tensors FIFO: same states, same diagnostics, entries keep their addresses; and
    ``vf_env_finish`` zeroes the ring rows of exactly the agents it re-initialises (reference dynamics.py:262-263)."""
    from visfly_b200 import _lib
    from visfly_b200 import params as P
    n = 200
    kw = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06)
    g = th.Generator().manual_seed(11)
    acts = (th.rand(9, n, 4, generator=g) * 2 - 1).cuda()
    a, b = make_dynamics(n, **kw), make_dynamics(n, **kw)
    b._fifo_ring = True
    ptrs = [t.data_ptr() for t in b._pre_action]
    with th.no_grad():
        for t in range(9):
            sa, sb = a.step(acts[t]), b.step(acts[t])
            assert th.equal(sa, sb) and th.equal(a.acceleration, b.acceleration) and th.equal(a.thrusts, b.thrusts)
            assert [x.data_ptr() for x in b._pre_action] == ptrs
            for x, y in zip(a._pre_action, b._pre_action):
                assert th.equal(x, y)
    # wrapper tail: rows of the agents that end here (time limit for odd agents) are zeroed in every FIFO entry
    spec = P.VfEnvSpec()
    spec.task, spec.obs_kind, spec.max_episode_steps, spec.fifo_depth = P.TASK_CUSTOM, P.OBS_STATE13, 5, 3
    spec.uav_radius, spec.success_radius, spec.init_motor_omega, spec.seed = 0.1, 0.5, float(b._init_motor_omega), 9
    for j in range(3):
        spec.bbox_lo[j], spec.bbox_hi[j] = -1e3, 1e3
        spec.gen_mean[0][0][j], spec.gen_half[0][0][j] = 0.0, 1.0
    spec.gen_kind, spec.gen_boxes = P.GEN_UNIFORM, 1
    sc = th.where(th.arange(n, device="cuda") % 2 == 1, 4, 0).to(th.int32)
    status = _lib.pack_status(sc, th.zeros(n, device="cuda"), th.zeros(n, dtype=th.int32, device="cuda"))
    before = [x.clone() for x in b._pre_action]
    state_out, status_out, obs, done, record = _lib.env_finish(
        b._cfg.params, spec, 0, 0, b._state, status, th.zeros(n, device="cuda"), None, None, fifo=b._pre_action,
        status_out=status)
    assert status_out is status and th.equal(done, sc == 4) and bool(done.any()) and not bool(done.all())
    for x, y in zip(before, b._pre_action):
        assert th.equal(y, th.where(done.view(-1, 1), 0.0, x))
    assert th.equal(status[:, 0], th.where(done, 0, sc + 1).to(th.int32))


def test_diagnostics_notice_an_overwritten_action_without_a_fifo():
    """comm_delay=0 keeps no copy of the action; the on-demand diagnostics re-run the step on its inputs and must
    refuse if the caller has overwritten the action buffer in the meantime (instead of reporting a different step)."""
    n = 32
    d = make_dynamics(n, action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0)
    buf = th.zeros(n, 4, device="cuda")
    d.step(buf)
    acc = d.acceleration.clone()
    d.step(buf)
    buf.add_(0.5)
    with pytest.raises(RuntimeError, match="modified in place"):
        d.acceleration
    d.step(buf)
    assert d.acceleration.shape == acc.shape


@pytest.mark.parametrize("host", [True, False])
def test_sensor_ingest_matches_the_reference_post_processing(host):
    """Renderer hand-off, ingestion side: per-agent frames -> observation tensors exactly as the
    reference's `update_observation` builds them on the host (envs/base/droneEnv.py:296-312: np.stack, expand_dims /
    transpose, np.where(depth == 0, 20, depth)), from page-locked host buffers (zero-copy) and from device buffers."""
    from visfly_b200.render_handoff import SensorIngest
    rng = np.random.default_rng(4)
    for n, h, w in ((37, 64, 64), (5, 9, 7), (4096, 64, 64) if not host else (300, 48, 36)):
        per_agent = []
        for _ in range(n if n < 400 else 3):           # per-agent dicts as a renderer returns them
            d = rng.random((h, w), dtype=np.float32) * 8
            d[rng.random((h, w)) < 0.2] = 0.0           # no-return pixels
            per_agent.append({"depth": d, "color": rng.integers(0, 256, (h, w, 4), dtype=np.uint8),
                              "semantic": rng.integers(0, 40, (h, w)).astype(np.float32)})
        per_agent = (per_agent * (n // len(per_agent) + 1))[:n]
        # the reference's lines, verbatim in numpy
        ref_depth = np.expand_dims(np.stack([o["depth"] for o in per_agent]), 1)
        ref_depth = np.where(ref_depth == 0, 20, ref_depth)
        ref_color = np.transpose(np.stack([o["color"] for o in per_agent])[..., :3], (0, 3, 1, 2))
        ref_sem = np.expand_dims(np.stack([o["semantic"] for o in per_agent]), 1)
        sensors = {"depth": (h, w), "semantic": (h, w)}
        if (h * w) % 4 == 0:
            sensors["color"] = (h, w)
        ing = SensorIngest(n, sensors, device="cuda", host=host)
        for uuid in sensors:
            ing.buffer(uuid).copy_(th.from_numpy(np.stack([o[uuid] for o in per_agent])))
        out = ing.ingest()
        th.cuda.synchronize()
        assert out["depth"].shape == (n, 1, h, w) and np.array_equal(out["depth"].cpu().numpy(), ref_depth)
        assert np.array_equal(out["semantic"].cpu().numpy(), ref_sem)
        if "color" in sensors:
            assert out["color"].dtype == th.uint8 and np.array_equal(out["color"].cpu().numpy(), ref_color)
    with pytest.raises(KeyError):
        SensorIngest(4, {"lidar": (8, 8)})
