"""The env-tail oracle (oracle/env_oracle.py) against golden runs of the real reference envs and, where the
reference tree is mounted, bit-exactly against the live reference (same seeds, same RNG consumption)."""
import numpy as np
import pytest
import torch as th

from _env_util import DYN, episode_records, load_env_golden, table_of
from _reference import load_reference_envs, reference_available
from _util import rel_l2
from oracle.env_oracle import OracleEnv

TASKS = {"hover": "HoverEnv", "navigation": "NavigationEnv", "racing2": "RacingEnv2"}


def table_generator(table):
    def generate(indices=None):
        idx = th.arange(table[0].shape[0]) if indices is None else th.as_tensor(indices)
        return tuple(x[idx].clone() for x in table)
    return generate


@pytest.mark.parametrize("task", list(TASKS))
@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_env_oracle_replays_reference_golden(task, integ):
    z = load_env_golden(task, integ)
    acts = th.from_numpy(z["actions"])
    T, n = acts.shape[:2]
    env = OracleEnv(task, n, dict(DYN[integ], comm_delay=0.06), max_episode_steps=int(z["max_episode_steps"]),
                    generate_state=table_generator(table_of(z)))
    obs = env.reset()
    for k in obs:
        assert rel_l2(obs[k].double(), z["reset_obs_" + k]) < 1e-7 or float(np.abs(z["reset_obs_" + k]).max()) == 0
    for t in range(T):
        obs, r, d, info = env.step(acts[t])
        assert np.array_equal(d.numpy(), z["done"][t]), t
        assert rel_l2(obs["state"], z["obs_state"][t]) < 2e-6, t
        if "gate" in obs:
            assert np.array_equal(obs["gate"].numpy(), z["obs_gate"][t])
        np.testing.assert_allclose(r.numpy(), z["reward"][t], rtol=2e-5, atol=2e-6)
        er, el, tr, sc, co, pg = episode_records(d.numpy(), info, n)
        np.testing.assert_allclose(er, z["episode_r"][t], rtol=2e-5, atol=2e-6)
        assert np.array_equal(el, z["episode_l"][t]) and np.array_equal(tr, z["truncated"][t])
        assert np.array_equal(sc, z["is_success"][t]) and np.array_equal(co, z["collision"][t])
        assert np.array_equal(pg, z["past_gate"][t])


@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_env_oracle_gradients_match_reference_golden(integ):
    z = np.load(__import__("os").path.join(__import__("_env_util").GOLD, "envgrad_navigation.npz"))
    acts = th.from_numpy(z["actions"]).requires_grad_(True)
    H, n = acts.shape[:2]
    env = OracleEnv("navigation", n, dict(DYN[integ], comm_delay=0.06), max_episode_steps=int(z["max_episode_steps"]),
                    requires_grad=True, generate_state=table_generator(table_of(z)))
    env.reset()
    loss = 0.0
    for t in range(H):
        obs, r, d, info = env.step(acts[t])
        loss = loss - (0.99 ** t) * r
    loss = loss.mean()
    g, = th.autograd.grad(loss, acts)
    assert abs(loss.item() - float(z[f"loss_{integ}"])) < 1e-6
    assert rel_l2(g, z[f"grad_actions_{integ}"]) < 1e-5


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("task", list(TASKS))
def test_env_oracle_is_bit_exact_with_live_reference(task):
    """Random initial states from the reference's own generators: identical seeds => identical runs."""
    n, T, dyn = 12, 40, dict(DYN["euler"])
    rk = {"state_generator": {"class": "Uniform", "kwargs": [{
        "position": {"mean": [1., 0., 1.5], "half": [1.0, 1.0, 0.5]},
        "orientation": {"mean": [0., 0, 0], "half": [0.2, 0.2, 3.0]},
        "velocity": {"mean": [0., 0, 0], "half": [1., 1, 1]}}]}}
    cls = load_reference_envs()[TASKS[task]]
    kw = dict(num_agent_per_scene=n, visual=False, device="cpu", dynamics_kwargs=dict(dyn), max_episode_steps=9)
    if task == "hover":
        kw["tensor_output"] = True
    if task != "racing2":
        kw["random_kwargs"] = rk
    g = th.Generator().manual_seed(1)
    acts = th.rand(T, n, 4, generator=g) * 2 - 1

    def run(env):
        th.manual_seed(123)
        out = [env.reset()]
        for t in range(T):
            out.append(env.step(acts[t].clone()))
        return out

    ref = run(cls(**kw))
    orc = run(OracleEnv(task, n, dict(dyn), max_episode_steps=9, random_kwargs=None if task == "racing2" else rk))
    for k in orc[0]:
        assert th.equal(ref[0][k], orc[0][k])
    n_done = 0
    for a, b in zip(ref[1:], orc[1:]):
        for k in b[0]:
            assert th.equal(a[0][k], b[0][k])
        assert th.equal(a[1], b[1]) and th.equal(a[2], b[2])
        for i in range(n):
            assert a[3][i].keys() == b[3][i].keys()
            if "episode" in a[3][i]:
                n_done += 1
                assert float(a[3][i]["episode"]["r"]) == float(b[3][i]["episode"]["r"])
                assert int(a[3][i]["episode"]["l"]) == int(b[3][i]["episode"]["l"])
                assert a[3][i]["TimeLimit.truncated"] == b[3][i]["TimeLimit.truncated"]
                assert a[3][i]["episode"]["extra"].keys() == b[3][i]["episode"]["extra"].keys()
                for k in b[3][i]["terminal_observation"]:
                    assert th.equal(a[3][i]["terminal_observation"][k], b[3][i]["terminal_observation"][k])
    assert n_done >= n * 3
