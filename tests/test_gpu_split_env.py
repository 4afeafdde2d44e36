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


def _graph_env(n, capture, **kw):
    from visfly_b200.envs import NavigationEnv

    class CarryNav(NavigationEnv):
        """Task code with per-agent state of its own, re-bound every step (progress shaping against the previous
        distance) and re-initialised for finished agents in the reset hook."""

        def get_success(self):
            return (self.position - self.target).norm(dim=1) <= 0.8

        def get_reward(self, predicted_obs=None):
            dist = (self.position - self.target).norm(dim=1)
            prev = getattr(self, "_prev_dist", None)
            progress = 0.0 if prev is None else prev - dist
            self._prev_dist = dist.detach()
            return progress - 0.01 * self.angular_velocity.norm(dim=1) + self._success * 1.0 \
                - 0.001 * self._step_count - 0.05 * self.is_collision

        def _on_reset_where(self, mask):
            if getattr(self, "_prev_dist", None) is not None:
                self._prev_dist = th.where(mask, 3.0, self._prev_dist)

    env = CarryNav(num_agent_per_scene=n, visual=False, device="cuda", seed=5, max_episode_steps=17,
                   dynamics_kwargs=dict(DYN["rk4"], comm_delay=0.04),
                   random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                       {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}}, **kw)
    env.capture_task_step = capture
    env.reset()
    return env


def test_recorded_task_step_replays_bitwise_what_the_two_launch_path_computes():
    """capture_task_step: control step kernel + the task's tensor ops + vf_env_finish as one CUDA-graph replay.  Same
    seeds, same actions => the same numbers as the eager two-launch path, through time-limit restarts (in-kernel
    Philox sampler keyed by the step number), the task's own carried tensor, eager steps in between, an agent reset
    from outside and a changed setting (re-recorded)."""
    n, T = 2048, 90
    g = th.Generator(device="cuda").manual_seed(3)
    acts = (th.rand((T, n, 4), device="cuda", generator=g) * 2 - 1) * 0.6
    logs = {}
    for capture in (False, True, "copy"):
        env = _graph_env(n, capture)
        out, kept = [], []
        for t in range(T):
            if t == 40:
                env.capture_task_step = False                 # two eager steps in between: the step numbering goes on
            if t == 42:
                env.capture_task_step = capture
            if t == 55:
                env.reset_agent_by_id(th.arange(0, n, 7, device="cuda"))
                env._prev_dist = th.where(th.arange(n, device="cuda") % 7 == 0, 3.0, env._prev_dist)
            if t == 70:
                env.max_episode_steps = 9                     # the spec changes: the recording is dropped and re-made
            obs, r, d, info = env.step(acts[t])
            kept.append((obs["state"], r, d))
            done_ids = d.nonzero().flatten().tolist()[:3]
            recs = [(i, float(info[i]["episode"]["r"]), int(info[i]["episode"]["l"]),
                     info[i]["terminal_observation"]["state"].clone()) for i in done_ids]
            out.append((obs["state"].clone(), r.clone(), d.clone(), recs, env._step_count.clone(), env._rewards.clone(),
                        env.position.clone(), env._prev_dist.clone()))
            if t % 30 == 29:        # on-demand diagnostics of the step just taken (a replay writes them in its own launch)
                out[-1] = out[-1] + (env.envs.dynamics.acceleration.clone(), env.envs.dynamics.thrusts.clone())
        logs[capture] = (out, kept)
        if capture:
            assert env._task_graph is not None and env._task_graph.replays > 10
        else:
            assert env._task_graph is None
    assert sum(int(o[2].sum()) for o in logs[False][0]) > n          # every agent restarted at least once
    for mode in (True, "copy"):
        for t, (a, b) in enumerate(zip(logs[False][0], logs[mode][0])):
            for x, y in zip(a[:3] + a[4:], b[:3] + b[4:]):
                assert th.equal(x, y), (mode, t)
            assert len(a[3]) == len(b[3])
            for ra, rb in zip(a[3], b[3]):
                assert ra[:3] == rb[:3] and th.equal(ra[3], rb[3])
    # "copy" hands out fresh tensors (they still hold their step's values); True hands out the graph's own buffers
    out, kept = logs["copy"]
    assert th.equal(kept[60][0], out[60][0]) and th.equal(kept[60][1], out[60][1])
    out, kept = logs[True]
    assert kept[60][0].data_ptr() == kept[61][0].data_ptr()


_FALLBACK = """
import warnings
import torch as th
from visfly_b200.envs import HoverEnv

class Syncing(HoverEnv):
    def get_reward(self, predicted_obs=None):
        scale = float(self.position.abs().max())          # host read of device data: not capturable
        return -(self.position - self.target).norm(dim=1) / max(scale, 1.0)

env = Syncing(num_agent_per_scene=256, visual=False, device="cuda", seed=1, tensor_output=True,
              dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02))
env.capture_task_step = True
env.reset()
a = th.zeros((256, 4), device="cuda")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    for _ in range(6):
        env.step(a)
assert any("cannot be recorded" in str(x.message) for x in w), [str(x.message) for x in w]
assert env.capture_task_step is False and env._task_graph is None and env._split.active
obs, r, d, info = env.step(a)
th.cuda.synchronize()
assert bool(th.isfinite(r).all())
print("fallback ok")
"""


def test_task_code_that_cannot_be_recorded_falls_back_with_a_warning():
    """Runs in its own process: a failed stream capture is not something the other tests' CUDA context should see."""
    import subprocess
    import sys
    from _util import ROOT
    out = subprocess.run([sys.executable, "-c", _FALLBACK], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fallback ok" in out.stdout, out.stderr[-2000:]
