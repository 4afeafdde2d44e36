"""User-defined task envs (own get_success / get_reward as tensor code) on the two-launch path: control step kernel,
the task's tensor ops, vf_env_finish — against the generic tensor-op path, which replays the reference's golden runs."""
import numpy as np
import pytest
import torch as th

from _env_util import DYN, load_env_golden, table_of
from _util import rel_l2

pytestmark = pytest.mark.gpu


def make_env(n, table=None, **kw):
    from visfly_b200.envs import NavigationEnv

    class CustomNav(NavigationEnv):
        """A task a user would write: different success radius, own shaping, a bonus that reads the step counter."""

        def get_success(self):
            return (self.position - self.target).norm(dim=1) <= 0.8

        def get_reward(self, predicted_obs=None):
            dist = (self.position - self.target).norm(dim=1)
            shaping = -0.05 * dist - 0.01 * self.angular_velocity.norm(dim=1) - 0.002 * self.velocity.pow(2).sum(dim=1)
            wall = -0.02 / (self.collision_dis + 0.3)
            bonus = self._success * (self.max_episode_steps - self._step_count) * 0.05
            return shaping + wall + bonus

    env = CustomNav(num_agent_per_scene=n, visual=False, device="cuda", dynamics_kwargs=dict(DYN["rk4"], comm_delay=0.04),
                    max_episode_steps=11, **kw)
    if table is not None:
        env.envs.set_reset_table(*[x.cuda() for x in table])
    return env


def test_custom_task_env_takes_the_split_path_and_matches_the_generic_one():
    z = load_env_golden("navigation", "rk4")
    acts = th.from_numpy(z["actions"]).cuda()
    T, n = acts.shape[:2]
    runs = {}
    for split in (True, False):
        env = make_env(n, table_of(z))
        env.use_fused_step = split
        env.reset()
        assert env._fused is None and env._split is not None
        out = []
        for t in range(T):
            obs, r, d, info = env.step(acts[t])
            assert env._split.active == split
            recs = [(i, float(info[i]["episode"]["r"]), int(info[i]["episode"]["l"]), bool(info[i]["is_success"]),
                     bool(info[i]["TimeLimit.truncated"]), bool(info[i]["episode"]["extra"]["collision"]))
                    for i in d.nonzero().flatten().tolist()]
            term = [info[i]["terminal_observation"]["state"].clone() for i in d.nonzero().flatten().tolist()[:2]]
            out.append((obs["state"].clone(), r.clone(), d.clone(), recs, term, env._step_count.clone(),
                        env._rewards.clone(), env.success.clone(), env.episode_done.clone()))
        runs[split] = out
    assert any(len(o[3]) for o in runs[True]) and any(r[3] for o in runs[True] for r in o[3])    # ends, successes
    for a, b in zip(runs[True], runs[False]):
        assert rel_l2(a[0].cpu(), b[0].cpu()) < 1e-6 and th.allclose(a[1], b[1], atol=1e-5) and th.equal(a[2], b[2])
        assert len(a[3]) == len(b[3])
        for ra, rb in zip(a[3], b[3]):
            assert ra[0] == rb[0] and abs(ra[1] - rb[1]) < 1e-4 and ra[2:] == rb[2:]
        for ta, tb in zip(a[4], b[4]):
            assert th.allclose(ta, tb, atol=1e-5)
        assert th.equal(a[5].to(th.int32), b[5].to(th.int32)) and th.allclose(a[6], b[6], atol=1e-4)
        assert th.equal(a[7], b[7]) and th.equal(a[8], b[8])


def test_custom_task_env_gradients_and_modes():
    z = load_env_golden("navigation", "rk4")
    acts = th.from_numpy(z["actions"])[:14].cuda()
    n = acts.shape[1]
    grads = {}
    for split in (True, False):
        env = make_env(n, table_of(z), requires_grad=True)
        env.use_fused_step = split
        env.reset()
        a = acts.clone().requires_grad_(True)
        loss = 0.0
        for t in range(a.shape[0]):                          # crosses the time limit (11): resets inside the horizon
            obs, r, d, info = env.step(a[t])
            loss = loss - (0.98 ** t) * r.mean() + 1e-3 * obs["state"].pow(2).mean()
        assert env._split.active == split
        grads[split] = th.autograd.grad(loss, a)[0]
        env.detach()
    assert rel_l2(grads[True].cpu(), grads[False].cpu()) < 1e-4
    # numpy mode, random (Philox) restarts from the generator box, hand-over to the generic path and back
    env = make_env(4096, tensor_output=False,
                   random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                       {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
    env.reset()
    for t in range(25):
        if t == 14:
            env.use_fused_step = False
        if t == 18:
            env.use_fused_step = True
        obs, r, d, info = env.step(np.zeros((4096, 4), dtype=np.float32))
        assert isinstance(obs["state"], np.ndarray) and r.shape == (4096,) and d.dtype == np.int32
    assert env._split.active
    p = env.position
    assert bool((p[:, 0] > -5).all()) and bool(th.isfinite(p).all())
