"""Trainer utilities that need no GPU: TD(lambda) returns against vectors recorded from the
reference's own function (and against the live function where the reference tree is mounted), the policy networks,
and the data-parallel gradient exchange on world_size-2 gloo."""
import importlib.util
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from visfly_b200.algorithms import ActorCritic, all_reduce_gradients, compute_td_returns, polyak_update
from visfly_b200.algorithms.common import RolloutBuffer, broadcast_parameters
from visfly_b200.algorithms.policies import flatten_obs, obs_dim
from visfly_b200.envs.base._compat import spaces
from visfly_b200.type import TensorDict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_COMMON = "/root/reference/utils/algorithms/common.py"


def _cases(z):
    i = 0
    while f"case{i}_r" in z.files:
        yield i, [th.from_numpy(z[f"case{i}_{k}"]) for k in ("r", "done", "ep", "v")]
        i += 1


def test_td_returns_match_reference_golden_bit_exactly():
    z = np.load(os.path.join(GOLD, "td_returns.npz"))
    seen = 0
    for i, (r, done, ep, v) in _cases(z):
        for tag, kw in (("a", dict(gamma=0.99, lamda=0.95)), ("b", dict(gamma=0.9, lamda=0.5))):
            got = th.stack(compute_td_returns(list(r), list(done), list(v), episode_done=list(ep), **kw))
            assert np.array_equal(got.numpy(), z[f"case{i}{tag}_returns"]), (i, tag)
        got = th.stack(compute_td_returns(list(r), list(done), list(v)))
        assert np.array_equal(got.numpy(), z[f"case{i}c_returns"])
        seen += 1
    assert seen == 4


@pytest.mark.skipif(not os.path.isfile(REF_COMMON), reason="reference tree not mounted")
def test_td_returns_match_live_reference_function():
    spec = importlib.util.spec_from_file_location("make_algo_golden", os.path.join(GOLD, "make_algo_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.reference_function(REF_COMMON, "compute_td_returns")
    g = th.Generator().manual_seed(9)
    r, v = th.randn(20, 33, generator=g), th.randn(20, 33, generator=g)
    done = th.rand(20, 33, generator=g) < 0.1
    ep = done & (th.rand(20, 33, generator=g) < 0.5)
    a = th.stack(ref(list(r), list(done), list(v), episode_done=list(ep), gamma=0.97, lamda=0.9))
    b = th.stack(compute_td_returns(list(r), list(done), list(v), episode_done=list(ep), gamma=0.97, lamda=0.9))
    assert th.equal(a, b)


def test_td_returns_reduce_to_discounted_sum_without_bootstrap():
    """lambda = 1-eps, no terminations, zero values: the target is the plain discounted reward-to-go."""
    h, n, gamma = 6, 3, 0.9
    r = [th.full((n,), float(t + 1)) for t in range(h)]
    done = [th.zeros(n, dtype=th.bool) for _ in range(h)]
    v = [th.zeros(n) for _ in range(h)]
    out = compute_td_returns(r, done, v, gamma=gamma, lamda=1 - 1e-6)
    expect = [sum(gamma ** (k - t) * (k + 1) for k in range(t, h)) for t in range(h)]
    assert th.allclose(th.stack(out)[:, 0], th.tensor(expect), rtol=1e-4)


def _spaces():
    return (spaces.Dict({"state": spaces.Box(-np.inf, np.inf, (13,), np.float32),
                         "target": spaces.Box(-np.inf, np.inf, (3,), np.float32)}),
            spaces.Box(-1, 1, (4,), np.float32))


def test_policy_shapes_determinism_and_reparameterisation():
    obs_space, act_space = _spaces()
    assert obs_dim(obs_space) == 16
    th.manual_seed(0)
    pol = ActorCritic(obs_space, act_space, net_arch=[32, 32])
    obs = TensorDict({"state": th.randn(10, 13), "target": th.randn(10, 3)})
    assert flatten_obs(obs).shape == (10, 16)
    a1, _ = pol.actor(obs, deterministic=True)
    a2, _ = pol.actor(obs, deterministic=True)
    assert th.equal(a1, a2) and a1.shape == (10, 4) and bool((a1.abs() <= 1).all())
    act, logp, _ = pol.actor.action_log_prob(obs)
    assert act.shape == (10, 4) and logp.shape == (10,) and act.requires_grad
    act.sum().backward()                                    # the sample is differentiable w.r.t. the weights
    assert pol.actor.mu.weight.grad is not None and float(pol.actor.mu.weight.grad.abs().sum()) > 0
    q1, q2 = pol.critic(obs, a1.detach())
    assert q1.shape == (10, 1) and q2.shape == (10, 1)
    # stacked (H, N, k) observations of the horizon buffer pass through the critic as well
    obs3 = TensorDict.stack([obs, obs, obs])
    assert pol.critic(obs3, th.zeros(3, 10, 4))[0].shape == (3, 10, 1)
    before = [p.clone() for p in pol.critic_target.parameters()]
    with th.no_grad():
        for p in pol.critic.parameters():
            p.add_(1.0)
    polyak_update(pol.critic.parameters(), pol.critic_target.parameters(), tau=0.25)
    for b, s, t in zip(before, pol.critic.parameters(), pol.critic_target.parameters()):
        assert th.allclose(t, 0.75 * b + 0.25 * s, atol=1e-6)


def test_rollout_buffer_flattens_like_the_reference():
    buf = RolloutBuffer(gamma=0.99)
    n, h = 5, 4
    for t in range(h):
        obs = TensorDict({"state": th.full((n, 13), float(t))})
        buf.add(obs=obs, reward=th.ones(n), action=th.zeros(n, 4), next_obs=obs, done=th.zeros(n, dtype=th.bool),
                episode_done=th.zeros(n, dtype=th.bool), value=th.zeros(n))
    buf.compute_returns()
    assert buf.returns.shape == (h * n,) and buf.obs["state"].shape == (h * n, 13) and buf.action.shape == (h * n, 4)
    assert buf.returns[0] > buf.returns[-1] > 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        obs_space, act_space = _spaces()
        th.manual_seed(100 + rank)                                   # different initial weights per rank ...
        pol = ActorCritic(obs_space, act_space, net_arch=[16])
        broadcast_parameters(pol)                                    # ... until rank 0's are broadcast
        th.manual_seed(7)
        full = th.randn(12, 16)
        shard = full[rank * 6:(rank + 1) * 6]
        loss = pol.actor(shard, deterministic=True)[0].pow(2).sum(dim=1).mean()      # mean over THIS rank's agents
        loss.backward()
        assert all_reduce_gradients(pol.actor.parameters()) == world
        got = th.cat([p.grad.reshape(-1) for p in pol.actor.parameters() if p.grad is not None])
        # single-process answer: mean over all 12 agents
        ref = ActorCritic(obs_space, act_space, net_arch=[16])
        ref.load_state_dict(pol.state_dict())
        ref.actor(full, deterministic=True)[0].pow(2).sum(dim=1).mean().backward()
        want = th.cat([(p.grad if p.grad is not None else th.zeros_like(p)).reshape(-1) for p in ref.actor.parameters()])
        out[rank] = bool(th.allclose(got, want, atol=1e-6)) and \
            float(sum(p.sum() for p in pol.parameters())) == float(sum(p.sum() for p in ref.parameters()))
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_world2_gloo_equals_single_process_gradient():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_all_reduce_is_a_no_op_without_a_process_group():
    lin = th.nn.Linear(3, 2)
    lin(th.ones(1, 3)).sum().backward()
    g = lin.weight.grad.clone()
    assert all_reduce_gradients(lin.parameters()) == 1 and th.equal(g, lin.weight.grad)


@pytest.mark.parametrize("n", [7, 512, 4096, 65536 // 8 + 512])
def test_wide_batch_linear_equals_nn_linear(n):
    """Same outputs, same gradients (weight, bias, input) as nn.Linear; sliced weight gradient for large batches."""
    from visfly_b200.algorithms.policies import WideBatchLinear, _slices
    th.manual_seed(n)
    ref = th.nn.Linear(16, 64).double()
    new = WideBatchLinear(16, 64).double()
    new.load_state_dict(ref.state_dict())
    x = th.randn(n, 16, dtype=th.float64)
    g = th.randn(n, 64, dtype=th.float64)
    xs = [x.clone().requires_grad_(True) for _ in range(2)]
    outs = [m(xi) for m, xi in ((ref, xs[0]), (new, xs[1]))]
    assert th.equal(outs[0], outs[1])
    for o in outs:
        (o * g).sum().backward()
    assert th.allclose(xs[0].grad, xs[1].grad, rtol=1e-12, atol=1e-12)
    assert th.allclose(ref.weight.grad, new.weight.grad, rtol=1e-11, atol=1e-11)
    assert th.allclose(ref.bias.grad, new.bias.grad, rtol=1e-11, atol=1e-11)
    assert _slices(65536) == 128 and _slices(7) == 1 and _slices(512) == 2 and 65536 % _slices(65536) == 0
    with th.no_grad():
        assert th.equal(new(x), ref(x))
