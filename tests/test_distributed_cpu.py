"""Host-side logic of the multi-GPU path on world_size-2 gloo (CPU tensors): sharding covers the agent range,
the one collective of the path (all-gather of episode returns) works for equal and ragged shards."""
import os
import socket

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from visfly_b200.distributed import gather_episode_returns, rollout_stats, shard_range, shard_seed


def test_shard_range_partitions_exactly():
    for n, world in [(65536, 8), (524288, 8), (10, 3), (7, 8), (1, 2), (0, 4)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)
    assert len({shard_seed(42, r) for r in range(8)}) == 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        local = th.arange(lo, hi, dtype=th.float32) * 0.5          # "episode return" of agent i is i/2
        full = gather_episode_returns(local, n_total=n_total).clone()
        again = gather_episode_returns(local + 1.0, n_total=n_total)      # cached gather object, buffers reused
        ref = th.arange(n_total, dtype=th.float32) * 0.5
        ok = th.equal(full, ref) and th.equal(again, ref + 1.0)
        mean_r, mean_l, cnt = rollout_stats(local.sum(), th.tensor(float(hi - lo) * 10), th.tensor(float(hi - lo)))
        ok = ok and cnt == n_total and abs(mean_l - 10.0) < 1e-9 and abs(mean_r - 0.5 * (n_total - 1) / 2) < 1e-6
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 9])
def test_gather_episode_returns_world2_gloo(n_total):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_single_process_is_identity():
    x = th.arange(5.0)
    assert gather_episode_returns(x) is x


def test_bind_host_to_gpu_degrades_to_a_no_op_without_a_gpu():
    from visfly_b200.distributed import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    got = bind_host_to_gpu(0)
    assert got is None or set(got) <= before
    if got is None:
        assert os.sched_getaffinity(0) == before


def _fused_fallback_worker(rank, world, port, n_total, out):
    import types
    from visfly_b200.distributed import FusedReturnsGather
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        g = FusedReturnsGather(hi - lo, n_total, rank, world, "cpu")     # no peer-mapped memory on CPU: NCCL-style path
        env = types.SimpleNamespace(_fused=None, use_fused_step=True, num_agent=hi - lo,
                                    _rewards=th.arange(lo, hi, dtype=th.float32) * 2.0)
        g.arm(env)
        full = g.finish(env)
        out[rank] = (g.fused, bool(th.equal(full, th.arange(n_total, dtype=th.float32) * 2.0)))
    finally:
        dist.destroy_process_group()


def test_fused_returns_gather_falls_back_to_the_collective_without_peer_memory():
    """`FusedReturnsGather` (episode returns scattered by the rollout's last env step over NVLink peer mappings) must
    degrade to the one-shot all-gather on every rank where peer mappings do not exist — here: CPU tensors over gloo,
    ragged shards."""
    world, n_total = 2, 11
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_fused_fallback_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
        assert dict(out) == {0: (False, True), 1: (False, True)}
