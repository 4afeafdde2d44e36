"""Host logic of the one-kernel env step that needs no GPU: the watched-attribute generation counters that tell
`FusedEnvStep.refresh` when the spec must be rebuilt (a changed target / radius / generator table must not
be ignored), and the status-record packing."""
import types

import torch as th

from visfly_b200 import _lib
from visfly_b200.envs.base.fused import FusedEnvStep, watched


class _Envs:
    _gen = 0
    stateGenerator = watched("stateGenerator")
    uav_radius = watched("uav_radius")


class _Env:
    _gen = 0
    target = watched("target")
    max_episode_steps = watched("max_episode_steps")


def _refresher(env):
    """A FusedEnvStep shell with a counting `_refresh` (the real one reads CUDA tensors)."""
    fz = FusedEnvStep.__new__(FusedEnvStep)
    fz.env = env
    fz._gen_seen, fz._ndict, fz._watch, fz._watch_sum, fz._ok = -1, None, (), 0, False
    fz.calls = 0

    def rebuild():
        fz.calls += 1
        fz._watch = (env.target,)
        return True
    fz._refresh = rebuild
    return fz


def test_watched_attributes_bump_the_generation_and_behave_like_attributes():
    e = _Env()
    assert not hasattr(e, "target") and getattr(e, "target", None) is None
    e.target = th.zeros(3)
    g1 = e._gen
    e.max_episode_steps = 256
    assert e._gen == g1 + 1 and e.max_episode_steps == 256 and th.equal(e.target, th.zeros(3))
    assert "_w_target" in e.__dict__ and _Env._gen == 0             # per-instance counter


def test_refresh_rebuilds_exactly_when_something_it_depends_on_changed():
    env = _Env()
    env.envs = _Envs()
    env.envs.stateGenerator, env.envs.uav_radius = object(), 0.1
    env.target, env.max_episode_steps = th.zeros(4, 3), 100
    fz = _refresher(env)
    assert fz.refresh() and fz.calls == 1
    for _ in range(5):
        assert fz.refresh()
    assert fz.calls == 1                                             # steady state: nothing is re-read
    env.max_episode_steps = 50                                       # plain re-assignment
    assert fz.refresh() and fz.calls == 2
    env.target[:] = 1.0                                              # in-place edit of a watched tensor
    assert fz.refresh() and fz.calls == 3
    env.target = th.ones(4, 3)                                       # replaced object
    assert fz.refresh() and fz.calls == 4
    env.envs.uav_radius = 0.2                                        # the other owner's counter
    assert fz.refresh() and fz.calls == 5
    env.get_reward = lambda predicted_obs=None: None                 # instance-level task override
    assert fz.refresh() and fz.calls == 6
    env.envs._generate_state = lambda indices=None: None             # replaced generator function
    assert fz.refresh() and fz.calls == 7
    env._scratch = 1                                                 # unrelated attributes do not trigger
    assert fz.refresh() and fz.calls == 7


def test_status_record_roundtrip():
    n = 9
    sc = th.arange(n, dtype=th.int32)
    ret = th.linspace(-3, 5, n)
    eb = th.tensor([0, 1, 2, 3, 0, 1, 2, 3, 0], dtype=th.int32)
    gate, passed = th.arange(n, dtype=th.int64) % 4, th.arange(n, dtype=th.int64) * 3
    st = _lib.pack_status(sc, ret, eb, gate, passed)
    assert st.shape == (n, 4) and st.dtype == th.int32 and st.is_contiguous()
    s2, r2, e2, g2, p2 = _lib.unpack_status(st)
    assert th.equal(s2, sc) and th.equal(r2, ret) and th.equal(e2, eb)
    assert th.equal(g2.long(), gate) and th.equal(p2.long(), passed)
