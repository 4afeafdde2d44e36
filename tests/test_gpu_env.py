"""The product's Gym-style envs (visfly_b200.envs, CUDA dynamics underneath) against golden runs of the real
reference envs: observations, rewards, termination flags, episode records, auto-reset, and autograd gradients
through a requires_grad rollout.  Inputs are the recorded ones (start table, action sequence)."""
import copy
import os

import numpy as np
import pytest
import torch as th

from _env_util import DYN, GOLD, episode_records, load_env_golden, table_of
from _util import rel_l2

pytestmark = pytest.mark.gpu
TASKS = ["hover", "navigation", "racing2"]


def make_env(task, n, integ, max_episode_steps, table, path="generic", **kw):
    from visfly_b200.envs import HoverEnv, NavigationEnv, RacingEnv2
    cls = {"hover": HoverEnv, "navigation": NavigationEnv, "racing2": RacingEnv2}[task]
    if task == "hover":
        kw.setdefault("tensor_output", True)
    env = cls(num_agent_per_scene=n, visual=False, device="cuda", dynamics_kwargs=dict(DYN[integ], comm_delay=0.06),
              max_episode_steps=max_episode_steps, **kw)
    tbl = tuple(x.cuda() for x in table)
    if path == "fused":          # the official way to pin restarts: both paths honour it, the kernel reads the table
        env.envs.set_reset_table(*tbl)
        return env

    def generate(indices=None, num=None):
        if indices is None:
            return tbl
        idx = th.as_tensor(indices, device="cuda")
        return tuple(x[idx] for x in tbl)

    env.envs._generate_state = generate
    return env


@pytest.mark.parametrize("task", TASKS)
@pytest.mark.parametrize("integ", ["euler", "rk4"])
@pytest.mark.parametrize("path", ["generic", "fused"])
def test_env_replays_reference_golden(task, integ, path):
    """generic = tensor-op wrapper around the fused dynamics kernel; fused = the one-kernel env step."""
    z = load_env_golden(task, integ)
    acts = th.from_numpy(z["actions"]).cuda()
    T, n = acts.shape[:2]
    env = make_env(task, n, integ, int(z["max_episode_steps"]), table_of(z), path=path)
    obs = env.reset()
    assert (env._fused is not None) and env._fused.refresh() == (path == "fused")
    ref0 = z["reset_obs_state"]
    cols = slice(6, None) if task == "racing2" else slice(None)    # reference reset() returns stale gate columns
    assert rel_l2(obs["state"].cpu()[:, cols], ref0[:, cols]) < 1e-6
    for t in range(T):
        obs, r, d, info = env.step(acts[t])
        assert isinstance(r, th.Tensor) and r.shape == (n,) and d.dtype == th.bool
        assert np.array_equal(d.cpu().numpy(), z["done"][t]), t
        assert rel_l2(obs["state"].cpu(), z["obs_state"][t]) < 1e-5, t
        if "gate" in obs:
            assert np.array_equal(obs["gate"].cpu().numpy(), z["obs_gate"][t])
        if "target" in obs:
            assert np.array_equal(obs["target"].cpu().numpy(), z["obs_target"][t])
        np.testing.assert_allclose(r.cpu().numpy(), z["reward"][t], rtol=1e-4, atol=1e-5)
        er, el, tr, sc, co, pg = episode_records(d.cpu().numpy(), info, n)
        np.testing.assert_allclose(er, z["episode_r"][t], rtol=1e-4, atol=1e-5)
        assert np.array_equal(el, z["episode_l"][t]) and np.array_equal(tr, z["truncated"][t])
        assert np.array_equal(sc, z["is_success"][t]) and np.array_equal(co, z["collision"][t])
        assert np.array_equal(pg, z["past_gate"][t])
        idle = [i for i in range(n) if not z["done"][t][i]][:2]
        for i in idle:
            assert info[i] == {"TimeLimit.truncated": False, "episode_done": False}
        assert env._fused.active == (path == "fused")


@pytest.mark.parametrize("integ", ["euler", "rk4"])
@pytest.mark.parametrize("path", ["generic", "fused"])
def test_env_rollout_gradients_match_reference_autograd(integ, path):
    """generic: autograd through tensor-op rewards + ControlStep; fused: EnvControlStep (vf_env_step_bwd)."""
    z = np.load(os.path.join(GOLD, "envgrad_navigation.npz"))
    acts = th.from_numpy(z["actions"]).cuda().requires_grad_(True)
    H, n = acts.shape[:2]
    env = make_env("navigation", n, integ, int(z["max_episode_steps"]), table_of(z), path=path, requires_grad=True)
    env.reset()
    loss = 0.0
    for t in range(H):
        obs, r, d, info = env.step(acts[t])
        loss = loss - (0.99 ** t) * r
    loss = loss.mean()
    g, = th.autograd.grad(loss, acts)
    assert abs(loss.item() - float(z[f"loss_{integ}"])) < 1e-5
    assert rel_l2(g.cpu(), z[f"grad_actions_{integ}"]) < 1e-4
    assert env._fused.active == (path == "fused")
    env.detach()
    assert not env.envs.dynamics.packed_state.requires_grad


def test_output_modes_spaces_and_deepcopy():
    from visfly_b200.envs import HoverEnv, NavigationEnv
    n = 64
    dyn = dict(DYN["euler"])
    env = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dyn, max_episode_steps=4)   # numpy mode
    assert env.observation_space["state"].shape == (13,) and env.action_space.shape == (4,)
    assert env.num_envs == n and len(env) == n
    obs = env.reset()
    assert isinstance(obs["state"], np.ndarray) and obs["state"].shape == (n, 13)
    with pytest.raises(AssertionError):
        HoverEnv(num_agent_per_scene=2, visual=False, dynamics_kwargs=dyn).step(np.zeros((2, 4)))    # step before reset
    for t in range(5):
        obs, r, d, info = env.step(np.random.uniform(-1, 1, (n, 4)).astype(np.float32))
        assert isinstance(obs["state"], np.ndarray) and isinstance(r, np.ndarray) and d.dtype == np.int32
    assert d.sum() == 0 and len(info) == n        # step 4 ended every episode, step 5 is the first of the next ones
    nav = NavigationEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dyn, requires_grad=True,
                        random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                            {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
    assert "target" in nav.observation_space.spaces
    nav.reset()
    twin = copy.deepcopy(nav)                       # reference utils/algorithms/shac.py:121 deep-copies the env
    a = th.zeros(n, 4, device="cuda")
    o1 = nav.step(a)[0]["state"]
    o2 = twin.step(a)[0]["state"]
    assert th.equal(o1, o2)
    hov = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dyn, tensor_output=True, max_episode_steps=6)
    hov.reset()
    for _ in range(3):
        hov.step(a)
    assert hov._fused.active
    twin2 = copy.deepcopy(hov)                      # mid-run copy while the one-kernel path owns the env status
    for _ in range(5):                              # crosses an auto-reset: same seed + same step index => same draws
        r1, r2 = hov.step(a), twin2.step(a)
        assert th.equal(r1[0]["state"], r2[0]["state"]) and th.equal(r1[1], r2[1]) and th.equal(r1[2], r2[2])
    nav.requires_grad = False                        # settable (reference PPO.py:80-82)
    nav.tensor_output = True
    assert not nav.step(a)[1].requires_grad
    with pytest.raises(NotImplementedError):
        HoverEnv(num_agent_per_scene=2, visual=True)


def test_random_state_generators_cover_the_configured_box():
    from visfly_b200.envs import HoverEnv, RacingEnv2
    n = 4096
    env = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["euler"]), tensor_output=True)
    p = env.reset()["state"][:, :3]
    lo, hi = th.tensor([0., -1, 1], device="cuda"), th.tensor([2., 1, 2], device="cuda")
    assert bool((p >= lo).all() and (p <= hi).all())
    assert float((p.mean(0) - th.tensor([1., 0, 1.5], device="cuda")).abs().max()) < 0.05
    race = RacingEnv2(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["euler"]))
    race.reset()
    centres = th.tensor([[2., 2., 1], [6., 2., 1.5], [6., -2., 1.5], [2., 0., 1]], device="cuda")
    dist = (race.position.unsqueeze(1) - centres).abs().amax(dim=2)
    which = dist.argmin(dim=1)
    assert bool((dist.amin(dim=1) <= 0.2 + 1e-6).all())
    assert all(abs(float((which == k).float().mean()) - 0.25) < 0.05 for k in range(4))
    # first gate from the start box (reference RacingEnv._choose_target)
    assert bool((race._next_target_i[which == 0] == 0).all() and (race._next_target_i[which == 1] == 1).all())


def test_fused_and_generic_paths_agree_and_hand_over_mid_run():
    """Same env stepped (a) fused all the way, (b) generic all the way, (c) fused then generic then fused."""
    from visfly_b200.envs import NavigationEnv
    z = load_env_golden("navigation", "rk4")
    acts = th.from_numpy(z["actions"]).cuda()
    T, n = acts.shape[:2]

    def build():
        env = NavigationEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["rk4"]), max_episode_steps=20)
        env.envs.set_reset_table(*[x.cuda() for x in table_of(z)])
        env.reset()
        return env

    def run(env, fused_at):
        outs = []
        for t in range(T):
            env.use_fused_step = fused_at(t)
            obs, r, d, info = env.step(acts[t])
            assert env._fused.active == fused_at(t)
            outs.append((obs["state"].detach().clone(), r.detach().clone(), d.clone(),
                         [float(info[i]["episode"]["r"]) for i in d.nonzero().flatten().tolist()]))
        return outs

    a = run(build(), lambda t: True)
    b = run(build(), lambda t: False)
    c = run(build(), lambda t: (t // 7) % 2 == 0)
    for x, y, w in zip(a, b, c):
        for other in (y, w):
            assert th.allclose(x[0], other[0], atol=1e-5, rtol=1e-5) and th.allclose(x[1], other[1], atol=1e-5)
            assert th.equal(x[2], other[2]) and np.allclose(x[3], other[3], atol=1e-5)


def test_fused_path_samples_resets_on_device():
    """Random (Philox) restarts of the fused step: right box, right bookkeeping, deterministic given the seed."""
    from visfly_b200.envs import HoverEnv, RacingEnv2
    n = 8192
    outs = []
    for _ in range(2):
        env = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["euler"]), max_episode_steps=5,
                       tensor_output=True, seed=7)
        env.envs.set_reset_table(th.tensor([[1.0, 0, 1.5]]).repeat(n, 1), th.tensor([[1.0, 0, 0, 0]]).repeat(n, 1))
        env.reset()
        env.envs.set_reset_table(None, None)          # from here on the generator (uniform box) is used
        a = th.zeros(n, 4, device="cuda")
        for t in range(5):
            obs, r, d, info = env.step(a)
        assert env._fused.active and bool(d.all())     # step 5 truncates everybody -> everybody restarts
        p = obs["state"][:, :3]
        lo, hi = th.tensor([0., -1, 1], device="cuda"), th.tensor([2., 1, 2], device="cuda")
        assert bool((p >= lo).all() and (p <= hi).all())
        assert float((p.mean(0) - th.tensor([1., 0, 1.5], device="cuda")).abs().max()) < 0.05
        assert float((p.std(0) - th.tensor([2., 2, 1], device="cuda") / 12 ** 0.5).abs().max()) < 0.03
        assert float(obs["state"][:, 3:7].sub(th.tensor([1., 0, 0, 0], device="cuda")).abs().max()) == 0
        assert int(env._step_count.sum()) == 0 and float(env._rewards.abs().sum()) == 0
        assert info[0]["TimeLimit.truncated"] and int(info[0]["episode"]["l"]) == 5
        assert info[3]["terminal_observation"]["state"].shape == (13,)
        outs.append(p.clone())
    assert th.equal(outs[0], outs[1])
    race = RacingEnv2(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["euler"]), max_episode_steps=3)
    race.reset()
    for t in range(3):
        obs, r, d, info = race.step(th.zeros(n, 4, device="cuda"))
    assert race._fused.active and bool(d.all()) and obs["state"].shape == (n, 16) and obs["gate"].shape == (n, 1)
    centres = th.tensor([[2., 2., 1], [6., 2., 1.5], [6., -2., 1.5], [2., 0., 1]], device="cuda")
    dist = (race.position.unsqueeze(1) - centres).abs().amax(dim=2)
    assert bool((dist.amin(dim=1) <= 0.2 + 1e-6).all())
    which = dist.argmin(dim=1)
    assert all(abs(float((which == k).float().mean()) - 0.25) < 0.03 for k in range(4))


@pytest.mark.parametrize("task", TASKS)
def test_fused_env_gradients_match_generic_autograd(task):
    """APG-style rollout (H=12, auto-resets inside the horizon) through all three tasks: gradients of the discounted
    return w.r.t. every action and the initial state, fused adjoint kernel vs autograd through the generic path."""
    z = load_env_golden(task, "rk4")
    acts0 = th.from_numpy(z["actions"])[:12].cuda()
    n = acts0.shape[1]
    res = {}
    for fused in (True, False):
        env = make_env(task, n, "rk4", 7, table_of(z), path="fused", requires_grad=True)
        env.use_fused_step = fused
        env.reset()
        s0 = env.envs.dynamics._state.detach().clone().requires_grad_(True)
        env.envs.dynamics._state = s0
        acts = acts0.clone().requires_grad_(True)
        loss = 0.0
        for t in range(acts.shape[0]):
            obs, r, d, info = env.step(acts[t])
            loss = loss - (0.99 ** t) * r.mean() + 1e-3 * obs["state"].pow(2).mean()
        assert env._fused.active == fused
        res[fused] = th.autograd.grad(loss, [acts, s0]) + (loss.detach(),)
    assert abs(float(res[True][2] - res[False][2])) < 1e-5
    assert rel_l2(res[True][0].cpu(), res[False][0].cpu()) < 1e-4
    assert rel_l2(res[True][1].cpu(), res[False][1].cpu()) < 1e-4


@pytest.mark.parametrize("task", TASKS)
def test_numpy_mode_host_mirror_equals_tensor_mode(task):
    """numpy output mode of the fused step (kernel writes obs / reward / done straight into page-locked host memory)
    returns exactly what the tensor mode returns; host actions may be page-locked, pageable numpy or CPU tensors."""
    z = load_env_golden(task, "rk4")
    acts = z["actions"]
    T, n = acts.shape[:2]
    steps = int(z["max_episode_steps"]) if "max_episode_steps" in z.files else 20

    def build(tensor_output):
        env = make_env(task, n, "rk4", steps, table_of(z), path="fused", tensor_output=tensor_output)
        env.reset()
        return env

    dev_env, np_env = build(True), build(False)
    pinned = th.empty((n, 4), pin_memory=True)
    held = []
    for t in range(T):
        o1, r1, d1, i1 = dev_env.step(th.from_numpy(acts[t]).cuda())
        if t % 3 == 0:                                   # page-locked host array
            pinned.copy_(th.from_numpy(acts[t]))
            a = pinned.numpy()
        elif t % 3 == 1:                                 # pageable numpy
            a = acts[t].copy()
        else:                                            # CPU tensor
            a = th.from_numpy(acts[t].copy())
        o2, r2, d2, i2 = np_env.step(a)
        assert np_env._fused.active and dev_env._fused.active
        assert isinstance(o2["state"], np.ndarray) and isinstance(r2, np.ndarray) and d2.dtype == np.int32
        assert set(o1.keys()) == set(o2.keys())
        for k in o1.keys():
            assert np.array_equal(o1[k].cpu().numpy(), np.asarray(o2[k])), (t, k)
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(d1.cpu().numpy().astype(np.int32), d2)
        for i in np.nonzero(d2)[0]:
            assert float(i1[int(i)]["episode"]["r"]) == float(i2[int(i)]["episode"]["r"])
        held.append((o2["state"], o2["state"].copy()))
        if len(held) >= 3:                                # arrays handed out two steps ago are still intact
            view, snap = held[-3]
            assert np.array_equal(view, snap)
        # the env objects show the step that was handed out, although the numpy env runs one step ahead inside
        assert th.equal(dev_env.state, np_env.state) and th.equal(dev_env._step_count, np_env._step_count)
        assert th.equal(dev_env._rewards, np_env._rewards)


@pytest.mark.parametrize("depth", [1, 3])
def test_host_mode_runs_ahead_and_rewinds_when_the_env_is_touched(depth):
    """numpy mode with a comm-delay FIFO launches step t+1 before it hands out step t (FusedEnvStep.step_host).
    Anything else done to the env in between (index-based reset, hand-over to the generic path, deep copy, switching
    to tensor output) must see the env exactly as after the step that was handed out: compared with a tensor-mode twin
    that steps synchronously and gets the same treatment."""
    z = load_env_golden("navigation", "rk4")
    acts = np.concatenate([z["actions"], z["actions"][::-1]])
    T, n = acts.shape[:2]
    dyn = dict(DYN["rk4"], comm_delay=0.02 * depth)

    def build(tensor_output):
        from visfly_b200.envs import NavigationEnv
        env = NavigationEnv(num_agent_per_scene=n, visual=False, device="cuda", dynamics_kwargs=dict(dyn),
                            max_episode_steps=9, tensor_output=tensor_output)
        env.envs.set_reset_table(*[x.cuda() for x in table_of(z)])
        env.reset()
        return env

    a_env, b_env = build(True), build(False)
    twins = None
    for t in range(T):
        o1, r1, d1, _ = a_env.step(th.from_numpy(acts[t]).cuda())
        o2, r2, d2, _ = b_env.step(acts[t].copy())
        if b_env.tensor_output:
            o2, r2, d2 = {k: v.cpu().numpy() for k, v in o2.items()}, r2.cpu().numpy(), d2.cpu().numpy()
        assert np.array_equal(o1["state"].cpu().numpy(), o2["state"]), t
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(d1.cpu().numpy().astype(np.int32), d2.astype(np.int32))
        if not b_env.tensor_output and b_env.use_fused_step:
            assert len(b_env._fused._ahead) == 1              # one step in flight between calls
        if t == 5:                                            # index-based reset of a few agents
            for e in (a_env, b_env):
                e.reset_agent_by_id([1, 4, 7])
            assert th.equal(a_env.state, b_env.state)
        if t == 11:                                           # a few steps on the generic tensor-op path
            a_env.use_fused_step = b_env.use_fused_step = False
        if t == 15:
            a_env.use_fused_step = b_env.use_fused_step = True
        if t == 20:                                           # deep copies continue identically
            twins = (copy.deepcopy(a_env), copy.deepcopy(b_env))
        if t == 30:                                           # switch the host env to tensor output mid-run
            b_env.tensor_output = True
    assert twins is not None
    for t in range(21, 26):
        o1 = twins[0].step(th.from_numpy(acts[t]).cuda())[0]["state"]
        o2 = twins[1].step(acts[t].copy())[0]["state"]
        assert np.array_equal(o1.cpu().numpy(), o2)


def test_numpy_mode_host_mirror_ragged_batch():
    """Agent counts that are not a multiple of the warp size take the scalar store path of the mirror as well."""
    from visfly_b200.envs import HoverEnv
    n = 83
    envs = [HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN["rk4"]), seed=5, max_episode_steps=6,
                     tensor_output=mode) for mode in (True, False)]
    th.manual_seed(5)                                     # the initial placement draws from torch's global generator
    o1 = envs[0].reset()
    th.manual_seed(5)
    o2 = envs[1].reset()
    assert np.array_equal(o1["state"].cpu().numpy(), o2["state"])
    g = th.Generator().manual_seed(3)
    for t in range(14):                                   # crosses two auto-resets (Philox keyed by seed/agent/step)
        a = th.rand(n, 4, generator=g) * 2 - 1
        o1, r1, d1, _ = envs[0].step(a.cuda())
        o2, r2, d2, _ = envs[1].step(a.numpy())
        assert np.array_equal(o1["state"].cpu().numpy(), o2["state"]) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy().astype(np.int32), d2)


@pytest.mark.parametrize("action_type", ["velocity", "position"])
def test_fused_env_step_runs_velocity_and_position_action_types(action_type):
    """The one-kernel env step is instantiated for the geometric-controller action types too: same results as the
    generic path (control-step kernel + tensor ops), which is itself checked against the reference golden runs."""
    from visfly_b200.envs import NavigationEnv
    n, T = 96, 40
    dyn = dict(DYN["rk4"], action_type=action_type)
    z = load_env_golden("navigation", "rk4")
    g = th.Generator().manual_seed(8)
    acts = (th.rand(T, n, 4, generator=g) * 2 - 1) * 0.3
    table = tuple(x.repeat(3, 1)[:n] for x in table_of(z))

    def run(fused):
        env = NavigationEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(dyn), max_episode_steps=15)
        env.envs.set_reset_table(*[x.cuda() for x in table])
        env.use_fused_step = fused
        env.reset()
        out = []
        for t in range(T):
            obs, r, d, info = env.step(acts[t].cuda())
            assert env._fused.active == fused
            out.append((obs["state"].clone(), r.clone(), d.clone()))
        return out

    for (o1, r1, d1), (o2, r2, d2) in zip(run(True), run(False)):
        assert th.allclose(o1, o2, atol=2e-5, rtol=2e-5) and th.allclose(r1, r2, atol=2e-5) and th.equal(d1, d2)
    assert bool(th.stack([d for _, _, d in run(True)]).any())          # the runs cross auto-resets


def test_config5_size_racing_shards_equal_the_whole_batch():
    """BASELINE config 5 size: RacingEnv semantics, 524 288 agents = 8 shards of 65 536.  The fused env step of the
    whole batch and of the eight contiguous shards (what eight ranks would run) give bitwise identical
    observations, rewards and termination flags across auto-resets; unit quaternions, finite rewards, gate indices in
    range throughout."""
    from visfly_b200.envs import RacingEnv2
    n, shards, T = 524288, 8, 7
    m = n // shards
    g = th.Generator().manual_seed(91)
    pos = th.rand(n, 3, generator=g) * th.tensor([6.0, 6.0, 1.5]) + th.tensor([0.5, -3.0, 0.5])
    quat = th.nn.functional.normalize(th.randn(n, 4, generator=g) * 0.1 + th.tensor([1.0, 0, 0, 0]), dim=1)
    vel, rate = th.randn(n, 3, generator=g), th.randn(n, 3, generator=g) * 0.3
    acts = (th.rand(T, n, 4, generator=g) * 2 - 1).cuda()

    def run(lo, hi):
        env = RacingEnv2(num_agent_per_scene=hi - lo, visual=False, device="cuda", tensor_output=True,
                         dynamics_kwargs=dict(DYN["rk4"], comm_delay=0.06), max_episode_steps=4)
        env.envs.set_reset_table(*(x[lo:hi].cuda() for x in (pos, quat, vel, rate)))
        env.reset()
        out = []
        for t in range(T):
            obs, r, d, info = env.step(acts[t, lo:hi])
            out.append((obs["state"].clone(), obs["gate"].clone(), r.clone(), d.clone()))
        assert env._fused.active
        return out

    whole = run(0, n)
    for t in range(T):
        st, gate, r, d = whole[t]
        assert bool(th.isfinite(st).all() and th.isfinite(r).all())
        assert float((st[:, 6:10].norm(dim=1) - 1).abs().max()) < 1e-5
        assert int(gate.min()) >= 0 and int(gate.max()) <= 3
    assert bool(whole[3][3].all())                      # step 4 truncates everybody: the resets are crossed
    for k in range(shards):
        part = run(k * m, (k + 1) * m)
        for t in range(T):
            for a, b in zip(whole[t], part[t]):
                assert th.equal(a[k * m:(k + 1) * m], b), (k, t)


def test_gradients_at_the_headline_batch_size_match_the_oracle_on_a_slice():
    """BASELINE config 3 size: NavigationEnv, 65 536 agents, requires_grad=True, RK4 x8.  Agents are independent, so the
    gradient of a short discounted-return rollout w.r.t. the actions of any subset of agents must equal what the
    float64 env oracle computes for that subset alone (same starts, same actions): checked on three slices of 128
    agents (first, middle, last rows of the batch) while the kernel runs the whole 65 536-agent launch."""
    from oracle.env_oracle import OracleEnv
    from visfly_b200.envs import NavigationEnv
    n, H, m = 65536, 6, 128
    g = th.Generator().manual_seed(65536)
    pos = th.stack([th.rand(n, generator=g) * 8, th.rand(n, generator=g) * 4 - 2, th.rand(n, generator=g) * 2 + 0.5], 1)
    quat = th.nn.functional.normalize(th.tensor([[1.0, 0, 0, 0]]) + 0.1 * th.randn(n, 4, generator=g), dim=1)
    vel, rate = 0.5 * th.randn(n, 3, generator=g), 0.3 * th.randn(n, 3, generator=g)
    acts = (th.rand(H, n, 4, generator=g) * 2 - 1) * 0.3
    acts[..., 0] -= 1.0 / 3.0
    dyn = dict(DYN["rk4"], comm_delay=0.04)
    env = NavigationEnv(num_agent_per_scene=n, visual=False, device="cuda", requires_grad=True, max_episode_steps=4,
                        dynamics_kwargs=dict(dyn))
    env.envs.set_reset_table(pos.cuda(), quat.cuda(), vel.cuda(), rate.cuda())
    env.reset()
    a_dev = acts.cuda().requires_grad_(True)
    loss = 0.0
    for t in range(H):                                   # every agent auto-resets after step 4
        obs, r, d, info = env.step(a_dev[t])
        loss = loss - (0.99 ** t) * r.sum()
    assert env._fused.active
    ga, = th.autograd.grad(loss, a_dev)
    for lo in (0, n // 2 - m // 2, n - m):
        sl = slice(lo, lo + m)
        table = tuple(x[sl].double() for x in (pos, quat, vel, rate))
        orc = OracleEnv("navigation", m, dict(dyn), max_episode_steps=4, requires_grad=True, dtype=th.float64,
                        faithful_rng=False,
                        generate_state=lambda idx=None: table if idx is None else tuple(x[th.as_tensor(idx)] for x in table))
        orc.reset()
        a_ref = acts[:, sl].double().requires_grad_(True)
        ref = 0.0
        for t in range(H):
            obs, r, d, info = orc.step(a_ref[t])
            ref = ref - (0.99 ** t) * r.sum()
        gr, = th.autograd.grad(ref, a_ref)
        assert rel_l2(ga[:, sl].cpu(), gr) < 1e-4, lo
