"""Analytic-gradient trainers on the GPU env: BPTT improves a hover policy, its horizon
gradient is the same through the one-kernel adjoint path and through the generic autograd path, SHAC runs with
TD(lambda) critic targets."""
import copy

import pytest
import torch as th

from _env_util import DYN
from _util import rel_l2

pytestmark = pytest.mark.gpu


def hover_env(n, seed=3, **kw):
    from visfly_b200.envs import HoverEnv
    return HoverEnv(num_agent_per_scene=n, visual=False, device="cuda", seed=seed, max_episode_steps=64,
                    dynamics_kwargs=dict(DYN["rk4"]), requires_grad=True, tensor_output=True, **kw)


def test_bptt_horizon_gradient_fused_adjoint_equals_generic_autograd():
    from visfly_b200.algorithms import BPTT
    grads = {}
    for fused in (True, False):
        th.manual_seed(11)
        env = hover_env(256)
        env.use_fused_step = fused
        algo = BPTT(env, horizon=24, policy_kwargs=dict(net_arch=[32, 32]), seed=5, make_eval_env=False)
        assert env.requires_grad
        loss = algo.rollout_loss()
        assert (env._fused is not None and env._fused.active) == fused
        algo.actor.optimizer.zero_grad()
        loss.backward()
        grads[fused] = (float(loss), th.cat([p.grad.reshape(-1) for p in algo.actor.parameters() if p.grad is not None]))
    assert abs(grads[True][0] - grads[False][0]) < 1e-5
    assert rel_l2(grads[True][1].cpu(), grads[False][1].cpu()) < 1e-4


def test_bptt_learns_to_hover_better():
    from visfly_b200.algorithms import BPTT
    env = hover_env(2048)
    algo = BPTT(env, horizon=32, learning_rate=3e-3, policy_kwargs=dict(net_arch=[64, 64]), seed=1,
                dump_step=2048 * 32 * 5)
    before = algo.evaluate()
    algo.learn(total_timesteps=2048 * 32 * 40)
    after = algo.evaluate()
    assert after["ep_rew_mean"] > before["ep_rew_mean"] + 0.05 * abs(before["ep_rew_mean"]), (before, after)
    assert len(algo.history) >= 7 and all(th.isfinite(th.tensor(h["actor_loss"])) for h in algo.history)
    assert algo.history[-1]["actor_loss"] < algo.history[0]["actor_loss"]


def test_shac_updates_actor_critic_and_target():
    from visfly_b200.algorithms import SHAC
    from visfly_b200.envs import NavigationEnv
    env = NavigationEnv(num_agent_per_scene=512, visual=False, device="cuda", seed=2, max_episode_steps=48,
                        dynamics_kwargs=dict(DYN["rk4"]), requires_grad=True,
                        random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                            {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
    algo = SHAC(env, horizon=16, gradient_steps=3, policy_kwargs=dict(net_arch=[32, 32]), seed=4,
                dump_step=512 * 16, make_eval_env=True)
    actor0 = copy.deepcopy(algo.actor.state_dict())
    target0 = copy.deepcopy(algo.critic_target.state_dict())
    algo.learn(total_timesteps=512 * 16 * 4)
    assert len(algo.history) == 4
    for h in algo.history:
        assert all(th.isfinite(th.tensor(float(h[k]))) for k in ("actor_loss", "critic_loss", "ep_rew_mean"))
    assert any(not th.equal(v, algo.actor.state_dict()[k]) for k, v in actor0.items())
    assert any(not th.equal(v, algo.critic_target.state_dict()[k]) for k, v in target0.items())
    assert env._fused is not None and env._fused.active          # the horizon ran on the one-kernel path


def test_bptt_update_replayed_as_a_cuda_graph_equals_the_eager_update():
    """BPTT(cuda_graph=True): horizon forward, backward through the fused adjoint, clipping and the Adam step captured
    once and replayed.  With deterministic restarts (reset table) and a noise-free policy the replayed updates must
    leave the same weights and the same env state as eager updates."""
    from visfly_b200.algorithms import BPTT
    from visfly_b200.envs import NavigationEnv
    n, H = 256, 8
    g = th.Generator().manual_seed(2)
    pos = th.stack([th.rand(n, generator=g) * 6, th.rand(n, generator=g) * 4 - 2, th.rand(n, generator=g) + 1], 1)
    quat = th.tensor([[1.0, 0, 0, 0]]).repeat(n, 1)
    out = {}
    for graph in (False, True):
        env = NavigationEnv(num_agent_per_scene=n, visual=False, device="cuda", requires_grad=True,
                            dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02),
                            max_episode_steps=12)
        env.envs.set_reset_table(pos.cuda(), quat.cuda())
        algo = BPTT(env, horizon=H, learning_rate=1e-3, policy_kwargs=dict(net_arch=[32, 32]), seed=3,
                    make_eval_env=False, dump_step=1 << 62, cuda_graph=graph)
        algo.learn(total_timesteps=9 * n * H)
        assert (algo._graph is not None) == graph and algo.num_timesteps == 9 * n * H
        th.cuda.synchronize()
        out[graph] = (th.cat([p.detach().reshape(-1) for p in algo.actor.parameters()]).clone(),
                      env.envs.dynamics.packed_state.detach().clone(), env._step_count.clone())
    assert th.allclose(out[True][0], out[False][0], atol=2e-5, rtol=1e-4)
    assert th.allclose(out[True][1], out[False][1], atol=1e-4, rtol=1e-4) and th.equal(out[True][2], out[False][2])


@pytest.mark.parametrize("task", ["hover", "navigation"])
def test_bptt_rollout_loss_and_policy_gradient_match_autograd_through_the_env_oracle(task):
    """n3 parity against the ORACLE, not against ourselves: `BPTT.rollout_loss` on the GPU env (one fused launch per
    step forward, one adjoint launch per step backward) versus the same horizon written out on the float64 CPU
    restatement of the reference env (`oracle/env_oracle.py`) with a float64 copy of the same policy: the loss and the
    gradient of every policy parameter.  Auto-resets happen inside the horizon (max_episode_steps < H)."""
    import copy as _copy
    from oracle.env_oracle import OracleEnv
    from visfly_b200.algorithms import BPTT
    from visfly_b200.envs import HoverEnv, NavigationEnv
    n, H, gamma = 192, 12, 0.99
    g = th.Generator().manual_seed(21)
    pos = th.stack([th.rand(n, generator=g) * 6 + 1, th.rand(n, generator=g) * 4 - 2, th.rand(n, generator=g) + 1], 1)
    quat = th.nn.functional.normalize(th.tensor([[1.0, 0, 0, 0]]) + 0.05 * th.randn(n, 4, generator=g), dim=1)
    vel, rate = 0.3 * th.randn(n, 3, generator=g), 0.2 * th.randn(n, 3, generator=g)
    dyn = dict(DYN["rk4"], comm_delay=0.04)
    cls = HoverEnv if task == "hover" else NavigationEnv
    kw = dict(tensor_output=True) if task == "hover" else {}
    env = cls(num_agent_per_scene=n, visual=False, device="cuda", requires_grad=True, max_episode_steps=7,
              dynamics_kwargs=dict(dyn), **kw)
    env.envs.set_reset_table(pos.cuda(), quat.cuda(), vel.cuda(), rate.cuda())
    algo = BPTT(env, horizon=H, gamma=gamma, policy_kwargs=dict(net_arch=[32, 32]), seed=9, make_eval_env=False)
    loss = algo.rollout_loss()
    assert env._fused is not None and env._fused.active
    algo.actor.optimizer.zero_grad()
    loss.backward()
    grads = [p.grad.detach().cpu().double() for p in algo.actor.parameters() if p.grad is not None]

    actor64 = _copy.deepcopy(algo.actor).cpu().double()
    table = tuple(x.double() for x in (pos, quat, vel, rate))
    orc = OracleEnv(task, n, dict(dyn), max_episode_steps=7, requires_grad=True, dtype=th.float64, faithful_rng=False,
                    generate_state=lambda idx=None: table if idx is None else tuple(x[th.as_tensor(idx)] for x in table))
    obs = orc.reset()
    ref_loss, discount = 0.0, th.ones(n, dtype=th.float64)
    for _ in range(H):
        flat = th.cat([obs[k].double() for k in sorted(obs.keys())], dim=-1)       # flatten_obs, kept in float64
        act, _, _ = actor64.action_log_prob(flat, noise_scale=0.0)
        obs, r, d, info = orc.step(act.clip(-1, 1))
        ref_loss = ref_loss - r * discount
        discount = discount * gamma * ~d + d
    ref_loss = ref_loss.mean()
    ref_grads = th.autograd.grad(ref_loss, [p for p in actor64.parameters()], allow_unused=True)
    ref_grads = [gr for gr in ref_grads if gr is not None]
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * max(1.0, abs(float(ref_loss)))
    assert len(grads) == len(ref_grads)
    assert rel_l2(th.cat([x.reshape(-1) for x in grads]), th.cat([x.reshape(-1) for x in ref_grads])) < 1e-4
