"""Adjoint of the fused env step (csrc/vf_env.cuh: reward adjoints + reset masking + control-step adjoint),
instantiated on the host in float64, against torch.autograd through the env oracle — one step, all three tasks,
with agents that finish in the step (collision, out of bounds, success, time limit), agents younger than the
comm-delay FIFO, and agents passing a gate."""
import ctypes

import pytest
import torch as th

from _util import INTEGRATOR, host_mirror, pack, random_flight_state, rel_l2, unpack, vf_params
from oracle.env_oracle import OracleEnv
from visfly_b200 import params as P

TASK_ID = {"hover": P.TASK_HOVER, "navigation": P.TASK_NAVIGATION, "racing2": P.TASK_RACING}
GATES = [[4, 4, 1.], [8, 0, 2.], [5, -4, 1.], [1, -1, 1.]]


def make_spec(task, max_steps=20, fifo=3):
    s = P.VfEnvSpec()
    s.task, s.obs_kind = TASK_ID[task], (P.OBS_RACING16 if task == "racing2" else P.OBS_STATE13)
    s.max_episode_steps, s.collision_reset, s.fifo_depth, s.uav_radius = max_steps, 1, fifo, 0.1
    for j, (lo, hi) in enumerate(zip((-30., -30., 0.), (30., 30., 8.))):
        s.bbox_lo[j], s.bbox_hi[j] = lo, hi
    tgt = {"hover": (1., 0., 1.5), "navigation": (9., 0., 1.)}.get(task, (0., 0., 0.))
    for j in range(3):
        s.target[j] = tgt[j]
    s.success_radius = 0.5 if task != "racing2" else 0.3
    s.n_gates = 4
    for a in range(4):
        for j in range(3):
            s.gates[a][j] = GATES[a][j]
    s.gen_kind, s.gen_boxes, s.init_motor_omega, s.seed = P.GEN_TABLE, 1, 1658.4774, 1
    return s


def scenario(task, n, seed):
    """States that exercise every branch of the tail."""
    pos, quat, vel, rate, motor, alpha = random_flight_state(n, seed=seed, spread=0.5, dtype=th.float64)
    g = th.Generator().manual_seed(seed + 100)
    age = th.randint(0, 18, (n,), generator=g, dtype=th.int32)
    age[:4] = th.tensor([0, 1, 2, 19])                 # younger than the FIFO (3) ... and one hitting the time limit
    gate = th.randint(0, 4, (n,), generator=g, dtype=th.int32)
    pos[4] = th.tensor([1.0, 0.5, 0.09]); vel[4] = th.tensor([0.2, 0.0, -0.5])      # collides with the floor
    pos[5] = th.tensor([29.99, 0.0, 2.0]); vel[5] = th.tensor([3.0, 0.0, 0.0])      # leaves the box
    pos[6] = th.tensor([2.0, 1.0, 0.6]); vel[6] = th.tensor([0.1, 0.2, -0.8])       # inside the proximity band
    pos[7] = th.tensor([3.0, -29.5, 3.0]); vel[7] = th.tensor([0.0, -1.0, 0.1])     # near the -y wall, closing in
    if task == "navigation":
        pos[8] = th.tensor([8.9, 0.05, 1.0]); vel[8] = th.tensor([0.3, 0.0, 0.0])   # success
        pos[9] = th.tensor([8.0, 0.0, 1.0]); vel[9] = th.tensor([12.0, 0.0, 0.0])   # approach speed above the clamp
        vel[10] = th.zeros(3)                                                        # exactly at rest
    if task == "racing2":
        pos[8] = th.tensor([3.95, 3.9, 1.0]); gate[8] = 0                           # passes gate 0
        pos[9] = th.tensor([7.9, 0.1, 1.95]); gate[9] = 1
    return pack(pos, quat, vel, rate, motor, alpha), age, gate


def oracle_env_grads(task, integ, dt, packed, action, age, gate, g_out, g_obs, g_rew, spec):
    n = packed.shape[1]
    table = (th.full((n, 3), 1.0, dtype=th.float64), th.tensor([[1.0, 0, 0, 0]], dtype=th.float64).repeat(n, 1),
             th.zeros(n, 3, dtype=th.float64), th.zeros(n, 3, dtype=th.float64))
    env = OracleEnv(task, n, dict(action_type="bodyrate", integrator=integ, dt=dt, ctrl_dt=0.02, comm_delay=0.0),
                    max_episode_steps=spec.max_episode_steps, requires_grad=True, dtype=th.float64, faithful_rng=False,
                    generate_state=lambda idx=None: tuple(x if idx is None else x[idx] for x in table))
    env._is_initial = True
    packed = packed.clone().requires_grad_(True)
    action = action.clone().requires_grad_(True)
    env.dyn.load_packed(packed)
    env.step_count = age.clone()
    if task == "racing2":
        env.next_target_i = gate.clone()
    a_eff = th.where((age < spec.fifo_depth).unsqueeze(1), th.zeros_like(action), action)
    obs, r, d, info = env.step(a_eff)
    loss = (env.dyn.packed() * g_out).sum() + (obs["state"] * g_obs).sum() + (r * g_rew).sum()
    gp, ga = th.autograd.grad(loss, (packed, action))
    return gp, ga, r.detach(), d


def mirror_env(task, integ, dt, packed, action, age, gate, g_out, g_obs, g_rew, spec, dtype=th.float64):
    lib = host_mirror()
    n = packed.shape[1]
    params = vf_params("bodyrate", dt)
    saved = th.stack([age, gate], 1).to(th.int32).contiguous()
    c = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    cast = lambda t: None if t is None else t.to(dtype).contiguous()
    packed, action, g_out, g_obs, g_rew = map(cast, (packed, action, g_out, g_obs, g_rew))
    S = int(round(0.02 / dt))
    so, rew = th.empty_like(packed), th.empty(n, dtype=th.float64)
    done, gate_out = th.empty(n, dtype=th.int32), th.empty(n, dtype=th.int32)
    if dtype == th.float64:
        lib.vfm_env_fwd_f64(ctypes.byref(params), ctypes.byref(spec), n, S, INTEGRATOR[integ], 1, 1, c(packed), c(action),
                            c(saved), c(so), c(rew), c(done), c(gate_out))
    gs, ga = th.empty_like(packed), th.empty_like(action)
    fn = lib.vfm_env_bwd_f64 if dtype == th.float64 else lib.vfm_env_bwd_f32
    fn(ctypes.byref(params), ctypes.byref(spec), n, S, INTEGRATOR[integ], 1, 1, 0, c(packed), c(action), c(saved),
       c(g_out), c(g_obs), c(g_rew), c(gs), c(ga))
    return gs, ga, rew, done.bool()


@pytest.mark.parametrize("task", ["hover", "navigation", "racing2"])
@pytest.mark.parametrize("integ,dt", [("euler", 0.005), ("rk4", 0.0025)])
def test_env_step_adjoint_matches_autograd(task, integ, dt):
    n = 64
    spec = make_spec(task)
    packed, age, gate = scenario(task, n, seed=5)
    g = th.Generator().manual_seed(9)
    action = (th.rand(n, 4, generator=g, dtype=th.float64) * 2 - 1) * 0.6
    width = 16 if task == "racing2" else 13
    g_out = th.randn(5, n, 4, generator=g, dtype=th.float64)
    g_obs = th.randn(n, width, generator=g, dtype=th.float64)
    g_rew = th.randn(n, generator=g, dtype=th.float64)
    ref_gs, ref_ga, ref_r, ref_d = oracle_env_grads(task, integ, dt, packed, action, age, gate.long(), g_out, g_obs, g_rew, spec)
    got_gs, got_ga, got_r, got_d = mirror_env(task, integ, dt, packed, action, age, gate, g_out, g_obs, g_rew, spec)
    assert th.equal(got_d, ref_d) and int(ref_d.sum()) >= 3
    assert rel_l2(got_r, ref_r) < 1e-6
    assert rel_l2(got_gs, ref_gs) < 5e-6 and rel_l2(got_ga, ref_ga) < 5e-6
    assert float(got_ga[:3].abs().max()) == 0.0 and float(ref_ga[:3].abs().max()) == 0.0     # FIFO-masked actions
    # each source of gradient separately, per agent (a wrong small term cannot hide behind a big one)
    for only in ("state", "obs", "reward"):
        go = g_out if only == "state" else th.zeros_like(g_out)
        gb = g_obs if only == "obs" else th.zeros_like(g_obs)
        gr = g_rew if only == "reward" else th.zeros_like(g_rew)
        r_gs, r_ga, _, _ = oracle_env_grads(task, integ, dt, packed, action, age, gate.long(), go, gb, gr, spec)
        m_gs, m_ga, _, _ = mirror_env(task, integ, dt, packed, action, age, gate, go, gb, gr, spec)
        for i in range(n):
            ref_i = th.cat([r_gs[:, i].flatten(), r_ga[i]])
            got_i = th.cat([m_gs[:, i].flatten(), m_ga[i]])
            if float(ref_i.norm()) > 0:
                assert rel_l2(got_i, ref_i) < 2e-5, (only, i)
            else:
                assert float(got_i.norm()) == 0.0, (only, i)


def test_env_step_adjoint_float32_within_tolerance():
    n, task, integ, dt = 64, "navigation", "rk4", 0.0025
    spec = make_spec(task)
    packed, age, gate = scenario(task, n, seed=6)
    g = th.Generator().manual_seed(10)
    action = (th.rand(n, 4, generator=g, dtype=th.float64) * 2 - 1) * 0.6
    g_out, g_obs = th.randn(5, n, 4, generator=g, dtype=th.float64), th.randn(n, 13, generator=g, dtype=th.float64)
    g_rew = th.randn(n, generator=g, dtype=th.float64)
    ref_gs, ref_ga, _, _ = oracle_env_grads(task, integ, dt, packed, action, age, gate.long(), g_out, g_obs, g_rew, spec)
    got_gs, got_ga, _, _ = mirror_env(task, integ, dt, packed.float(), action, age, gate, g_out, g_obs, g_rew, spec, dtype=th.float32)
    assert rel_l2(got_gs, ref_gs) < 1e-4 and rel_l2(got_ga, ref_ga) < 1e-4


@pytest.mark.parametrize("task", ["hover", "navigation", "racing2"])
def test_env_reward_adjoint_matches_central_differences(task):
    """Autograd-free check of the reward adjoints (csrc/vf_env.cuh): d(sum_i g_i * reward_i) / d(state, action) from
    the reverse sweep against central differences of the forward env step, float64, for agents in smooth flight (not
    finishing, older than the FIFO, away from the walls' proximity band and from the clamps)."""
    n, integ, dt, eps = 6, "rk4", 0.0025, 1e-6
    spec = make_spec(task, max_steps=1000)
    pos, quat, vel, rate, motor, alpha = random_flight_state(n, seed=23, spread=0.5, dtype=th.float64)
    pos = pos + th.tensor([2.0, 0.5, 2.0], dtype=th.float64)          # mid-air, metres away from every wall / target
    vel = vel + th.tensor([0.8, -0.4, 0.3], dtype=th.float64)         # speed well above zero (norms are smooth)
    packed = pack(pos, quat, vel, rate, motor, alpha)
    age = th.full((n,), 7, dtype=th.int32)
    gate = th.arange(n, dtype=th.int32) % 4
    g = th.Generator().manual_seed(29)
    action = (th.rand(n, 4, generator=g, dtype=th.float64) * 2 - 1) * 0.4
    action[:, 0] -= 0.3
    g_rew = th.randn(n, generator=g, dtype=th.float64)
    zs, zo = th.zeros(5, n, 4, dtype=th.float64), th.zeros(n, 16 if task == "racing2" else 13, dtype=th.float64)

    def weighted_reward(pk, ac):
        _, _, rew, done = mirror_env(task, integ, dt, pk, ac, age, gate, zs, zo, g_rew, spec)
        assert not bool(done.any())
        return rew * g_rew                                            # per agent (agents are independent)

    gs, ga, _, _ = mirror_env(task, integ, dt, packed, action, age, gate, zs, zo, g_rew, spec)
    fd_s, fd_a = th.zeros_like(packed), th.zeros_like(action)
    for plane in range(5):
        for lane in range(4):
            d = th.zeros_like(packed)
            d[plane, :, lane] = eps
            fd_s[plane, :, lane] = (weighted_reward(packed + d, action) - weighted_reward(packed - d, action)) / (2 * eps)
    for j in range(4):
        d = th.zeros_like(action)
        d[:, j] = eps
        fd_a[:, j] = (weighted_reward(packed, action + d) - weighted_reward(packed, action - d)) / (2 * eps)
    assert float(fd_s.abs().max()) > 1e-4                             # the reward really depends on the state
    assert rel_l2(gs, fd_s) < 1e-6, rel_l2(gs, fd_s)
    assert rel_l2(ga, fd_a) < 1e-6, rel_l2(ga, fd_a)
