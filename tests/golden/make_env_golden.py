"""Golden fixtures for the env wrapper tail, generated from the REAL reference envs (``/root/reference``).

    python tests/golden/make_env_golden.py        (build container only)

The reference's ``HoverEnv`` / ``NavigationEnv`` / ``RacingEnv2`` are imported with their absent third-party
imports stubbed (``tests/_reference.py``).  To make runs replayable by any implementation, the per-agent random
initial-state generator is replaced by a fixed table (agent i always restarts from row i) — everything else is
the reference's own code: bounding-box collision, rewards, success / out-of-bounds / collision / time-limit
termination, episode bookkeeping, auto-reset.

Fixtures: ``env_<task>_<integrator>.npz`` with the table, the actions, and per step the observation, reward,
done flags and the episode records of finished agents; ``envgrad_navigation.npz`` with autograd gradients of a
discounted return through a ``requires_grad=True`` rollout that auto-resets mid-horizon.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from _reference import load_reference_envs, reference_available  # noqa: E402

DYN = {
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02),
}
TASKS = {"hover": "HoverEnv", "navigation": "NavigationEnv", "racing2": "RacingEnv2"}


def start_table(task, n, seed=0):
    """Row i = where agent i (re)starts.  A few rows sit close to walls / gates / the target on purpose so that
    collisions, out-of-bounds, gate passes and successes all occur within the recorded horizon."""
    g = th.Generator().manual_seed(seed)
    r = lambda *s: th.rand(*s, generator=g)
    if task == "racing2":
        centres = th.tensor([[2., 2., 1], [6., 2., 1.5], [6., -2., 1.5], [2., 0., 1]])
        pos = centres[th.arange(n) % 4] + (r(n, 3) * 2 - 1) * 0.2
        pos[0] = th.tensor([3.9, 3.8, 1.0])           # next to gate 0
        pos[1] = th.tensor([7.9, 0.1, 1.9])           # next to gate 1
    elif task == "navigation":
        pos = th.stack([r(n) * 8, r(n) * 4 - 2, r(n) * 2 + 0.5], 1)
        pos[0] = th.tensor([8.7, 0.1, 1.0])           # inside the success radius of (9,0,1) after a step or two
        pos[1] = th.tensor([8.4, 0.0, 1.1])
    else:
        pos = th.stack([r(n) * 2, r(n) * 2 - 1, r(n) + 1.0], 1)
    pos[2] = th.tensor([1.0, 0.5, 0.13])              # just above the floor: collision (< 0.1) within a few steps
    pos[3] = th.tensor([29.95, 0.0, 2.0])             # at the x wall, flying outwards: out of bounds
    vel = (r(n, 3) * 2 - 1) * 0.5
    vel[2] = th.tensor([0.0, 0.0, -1.0])
    vel[3] = th.tensor([3.0, 0.0, 0.0])
    if task == "navigation":
        vel[0] = th.tensor([0.5, 0.0, 0.0])
        vel[1] = th.tensor([1.5, 0.0, 0.0])
    if task == "racing2":
        vel[0] = th.tensor([0.5, 1.0, 0.0])
        vel[1] = th.tensor([0.5, -0.5, 0.5])
    ang = (r(n, 1) * 2 - 1) * 0.3
    axis = th.randn(n, 3, generator=g)
    axis = axis / axis.norm(dim=1, keepdim=True)
    quat = th.cat([th.cos(ang / 2), axis * th.sin(ang / 2)], 1)
    rate = (r(n, 3) * 2 - 1) * 0.5
    return pos, quat, vel, rate


def action_seq(T, n, seed=1):
    g = th.Generator().manual_seed(seed)
    a = (th.rand(T, n, 4, generator=g) * 2 - 1) * 0.3
    a[..., 0] += -1.0 / 3.0                            # around 1 g collective
    return a


def make_env(cls, n, dyn, table, max_episode_steps, requires_grad=False):
    kw = dict(num_agent_per_scene=n, visual=False, device="cpu", dynamics_kwargs=dict(dyn, comm_delay=0.06),
              max_episode_steps=max_episode_steps, requires_grad=requires_grad)
    if cls.__name__ == "HoverEnv":
        kw["tensor_output"] = True
    env = cls(**kw)

    def generate(indices=None):
        idx = th.arange(n) if indices is None else th.as_tensor(indices)
        return tuple(x[idx].clone() for x in table)

    env.envs._generate_state = generate
    return env


def record(task, integ, n=32, T=48, max_episode_steps=20):
    cls = load_reference_envs()[TASKS[task]]
    table = start_table(task, n)
    env = make_env(cls, n, DYN[integ], table, max_episode_steps)
    acts = action_seq(T, n)
    obs0 = env.reset()
    out = {"actions": acts.numpy(), "max_episode_steps": np.array(max_episode_steps)}
    for k, x in zip(("pos", "quat", "vel", "rate"), table):
        out["table_" + k] = x.numpy()
    for k, v in obs0.items():
        out["reset_obs_" + k] = v.numpy()
    per = {k: [] for k in list(obs0.keys())}
    rew, done, ep_r, ep_l, trunc, succ, coll, gates = [], [], [], [], [], [], [], []
    for t in range(T):
        obs, r, d, info = env.step(acts[t].clone())
        for k in per:
            per[k].append(obs[k].clone().numpy())
        rew.append(r.numpy().copy())
        done.append(d.numpy().copy())
        er, el, tr, sc, co, pg = (np.full(n, np.nan, np.float32), np.full(n, -1, np.int32), np.zeros(n, bool),
                                  np.zeros(n, bool), np.zeros(n, bool), np.full(n, -1, np.int32))
        for i in range(n):
            if d[i]:
                er[i], el[i] = info[i]["episode"]["r"], info[i]["episode"]["l"]
                tr[i], sc[i] = info[i]["TimeLimit.truncated"], info[i]["is_success"]
                co[i] = bool(info[i]["episode"]["extra"]["collision"])
                pg[i] = info[i]["episode"]["extra"].get("past_gate", -1)
        ep_r.append(er); ep_l.append(el); trunc.append(tr); succ.append(sc); coll.append(co); gates.append(pg)
    for k in per:
        out["obs_" + k] = np.stack(per[k])
    out.update(reward=np.stack(rew), done=np.stack(done), episode_r=np.stack(ep_r), episode_l=np.stack(ep_l),
               truncated=np.stack(trunc), is_success=np.stack(succ), collision=np.stack(coll), past_gate=np.stack(gates))
    np.savez_compressed(os.path.join(HERE, f"env_{task}_{integ}.npz"), **out)
    return out


def record_grad(n=16, H=10, max_episode_steps=6):
    """NavigationEnv with requires_grad: -sum_t gamma^t r_t, every agent auto-resets once inside the horizon."""
    cls = load_reference_envs()["NavigationEnv"]
    table = start_table("navigation", n, seed=3)
    out = {}
    for integ in ("euler", "rk4"):
        env = make_env(cls, n, DYN[integ], table, max_episode_steps, requires_grad=True)
        acts = action_seq(H, n, seed=4).requires_grad_(True)
        env.reset()
        loss = 0.0
        for t in range(H):
            obs, r, d, info = env.step(acts[t])
            loss = loss - (0.99 ** t) * r
        loss = loss.mean()
        g, = th.autograd.grad(loss, acts)
        out[f"loss_{integ}"] = np.array(loss.item())
        out[f"grad_actions_{integ}"] = g.numpy()
        out["actions"] = acts.detach().numpy()
    for k, x in zip(("pos", "quat", "vel", "rate"), table):
        out["table_" + k] = x.numpy()
    out["max_episode_steps"] = np.array(max_episode_steps)
    np.savez_compressed(os.path.join(HERE, "envgrad_navigation.npz"), **out)


def main():
    if not reference_available():
        raise SystemExit("reference tree not found")
    for task in TASKS:
        for integ in DYN:
            o = record(task, integ)
            print(task, integ, "dones:", int(o["done"].sum()), "success:", int(o["is_success"].sum()),
                  "collision:", int(o["collision"].sum()), "truncated:", int(o["truncated"].sum()),
                  "gates:", int((o["past_gate"] > 0).sum()))
    record_grad()
    for f in sorted(os.listdir(HERE)):
        if f.startswith("env") and f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
