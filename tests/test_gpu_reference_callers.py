"""The reference's OWN analytic-gradient trainer on the new env.

INTEGRATION.md claims that `utils/algorithms/BPTT.py` works on `visfly_b200.envs` unchanged.  This test imports the
reference's `BPTT` class unmodified (`baseline/ref_loader.load_reference_algorithms`: stable-baselines3's logger /
schedule helpers stubbed, nothing of the trainer loop touched), hands it a `visfly_b200` NavigationEnv on the GPU and a
policy class, and lets `BPTT.learn` (reference BPTT.py:77-180) run: `env.get_observation()`, `env.step(actions)` with
autograd history, `actor_loss.backward()` through the fused adjoint kernels, `env.detach()`, the deep-copied evaluation
env stepped with `is_test=True`, per-agent `info[i]["episode"]` records.  Needs the reference tree (the build container's
`/root/reference` or the `baseline/_ref` copy that travels to the GPU box)."""
import tempfile

import pytest
import torch as th

from _reference import reference_available

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_reference_bptt_learn_runs_unchanged_on_the_new_env():
    from baseline.ref_loader import load_reference_algorithms
    from visfly_b200.algorithms.policies import ActorCritic
    from visfly_b200.envs import NavigationEnv

    class Policy(ActorCritic):                     # what shac._create_policy instantiates (shac.py:156-175)
        def __init__(self, observation_space, action_space, lr_schedule, **kw):
            super().__init__(observation_space, action_space, learning_rate=lr_schedule(1.0), **kw)

    RefBPTT = load_reference_algorithms()["BPTT"]
    n, H = 128, 8
    env = NavigationEnv(num_agent_per_scene=n, visual=False, device="cuda", requires_grad=True, max_episode_steps=24,
                        dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02),
                        random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                            {"position": {"mean": [7., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
    algo = RefBPTT(env=env, policy=Policy, policy_kwargs=dict(net_arch=[32, 32]), horizon=H, learning_rate=1e-3,
                   dump_step=4 * n * H, device="cuda", seed=1, save_path=tempfile.mkdtemp())
    assert algo.eval_env is not env and algo.eval_env.num_envs == n          # deepcopy(env) worked (shac.py:121)
    before = th.cat([p.detach().reshape(-1).clone() for p in algo.policy.actor.parameters()])
    algo.learn(total_timesteps=10 * n * H)
    after = th.cat([p.detach().reshape(-1) for p in algo.policy.actor.parameters()])
    assert bool(th.isfinite(after).all()) and float((after - before).abs().max()) > 1e-5
    assert env._fused is not None and env._fused.active                      # the horizon ran on the one-kernel path
    assert not env.envs.dynamics.packed_state.requires_grad                  # env.detach() at the end of the update
    # the evaluation pass ran (is_test=True steps, info records of finished agents) and was logged
    logged = algo._logger.dumps
    assert logged and "rollout/ep_rew_mean" in logged[-1][1] and "rollout/success_rate" in logged[-1][1]
    assert 1 <= float(logged[-1][1]["rollout/ep_len_mean"]) <= 24
