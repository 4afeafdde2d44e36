"""All seven drone parameter files (reference configs/drone/*.json): the oracle against the live
reference for the files the reference itself can load (four of them lack THRUST_PID and raise KeyError there,
dynamics.py:574), the CUDA engine against the oracle for all seven — forward trajectory and one-step gradient."""
import pytest
import torch as th

from _reference import make_reference_dynamics, reference_available
from _util import make_oracle, oracle_grads, oracle_step_packed, pack, random_flight_state, rel_l2, vf_params

ALL = ["drone_state", "drone_state_fast", "drone_d435i", "drone_d435i_n100", "drone_d435i_jetson_orin_nx",
       "drone_d435i_jetson_orin_nx_fast", "example"]
REF_LOADS = ["drone_state", "drone_d435i_jetson_orin_nx", "drone_d435i_jetson_orin_nx_fast"]


def hover_actions(T, n, seed):
    g = th.Generator().manual_seed(seed)
    a = (th.rand(T, n, 4, generator=g) * 2 - 1) * 0.3
    a[..., 0] -= 1.0 / 3.0
    return a


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("cfg", REF_LOADS)
@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
def test_oracle_bit_exact_with_live_reference_per_config(cfg, integ, dt):
    n, T = 16, 12
    init = random_flight_state(n, seed=51, spread=0.5)
    acts = hover_actions(T, n, 53)
    ref = make_reference_dynamics(n, action_type="bodyrate", dt=dt, ctrl_dt=0.02, integrator=integ, cfg=cfg,
                                  comm_delay=0.0)
    orc = make_oracle(n, "bodyrate", integ, dt, cfg=cfg)
    for d in (ref, orc):
        d.reset(pos=init[0].clone(), ori=init[1].clone(), vel=init[2].clone(), ori_vel=init[3].clone())
    for t in range(T):
        assert th.equal(ref.step(acts[t].clone()), orc.step(acts[t].clone())), (cfg, t)
    assert th.equal(ref.full_state, orc.full_state)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ALL)
@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
def test_engine_matches_oracle_per_config(cfg, integ, dt):
    from visfly_b200.dynamics import ControlStep, Dynamics
    n, T = 257, 32
    init = random_flight_state(n, seed=61, spread=0.5)
    acts = hover_actions(T, n, 63)
    eng = Dynamics(num=n, action_type="bodyrate", dt=dt, ctrl_dt=0.02, integrator=integ, cfg=cfg, comm_delay=0.0,
                   device="cuda")
    o32 = make_oracle(n, "bodyrate", integ, dt, cfg=cfg)
    o64 = make_oracle(n, "bodyrate", integ, dt, cfg=cfg, dtype=th.float64)
    for d, dt_ in ((eng, th.float32), (o32, th.float32), (o64, th.float64)):
        d.reset(pos=init[0].to(dt_), ori=init[1].to(dt_), vel=init[2].to(dt_), ori_vel=init[3].to(dt_))
    got = th.stack([eng.step(acts[t].cuda()).cpu() for t in range(T)])
    r32 = th.stack([o32.step(acts[t]).clone() for t in range(T)])
    r64 = th.stack([o64.step(acts[t].double()).clone() for t in range(T)])
    assert rel_l2(got, r64) < max(1e-5, 2 * rel_l2(r32, r64)), (cfg, rel_l2(got, r64), rel_l2(r32, r64))
    assert rel_l2(eng.full_state.cpu(), o64.full_state) < 1e-5

    # one-step gradient from a mid-flight state
    packed = pack(*random_flight_state(n, seed=65))
    g = th.Generator().manual_seed(67)
    action = th.rand(n, 4, generator=g) * 2 - 1
    g_out, g_obs = th.randn(5, n, 4, generator=g), th.randn(n, 13, generator=g)
    st, ac = packed.cuda().requires_grad_(True), action.cuda().requires_grad_(True)
    out, obs, _ = ControlStep.apply(st, ac, None, eng._cfg)
    ((out * g_out.cuda()).sum() + (obs * g_obs.cuda()).sum()).backward()
    ref_gs, ref_ga = oracle_grads(make_oracle(n, "bodyrate", integ, dt, cfg=cfg, dtype=th.float64), packed.double(),
                                  action.double(), g_out.double(), g_obs.double())
    assert rel_l2(st.grad.cpu(), ref_gs) < 1e-4 and rel_l2(ac.grad.cpu(), ref_ga) < 1e-4
