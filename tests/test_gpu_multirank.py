"""BASELINE config 5 across real ranks: RacingEnv2 shards, one process per shard, one all-gather of episode returns.

Two processes are spawned with ``torch.multiprocessing`` (rendezvous on 127.0.0.1).  With two or more GPUs visible each
rank owns its own GPU and the collective runs over NCCL; on a one-GPU box both ranks share cuda:0 and the collective
runs over gloo on host copies (NCCL cannot place two ranks on one device) — the env kernels are the same either way.
Every rank builds its shard with ``shard=(offset, total)`` and the SAME seed: initial placements, in-kernel Philox
restarts and therefore every observation / reward / done flag must equal, bit for bit, the corresponding rows of the
whole batch stepped by one process; the gathered episode returns must equal the whole batch's returns.
"""
import os
import socket

import pytest
import torch as th

pytestmark = pytest.mark.gpu

TOTAL, T, MAX_STEPS, SEED = 8192 + 6, 11, 4, 17          # ragged on purpose: shards of 4099 + 4099
DYN = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06)


def _actions():
    g = th.Generator().manual_seed(3)
    a = (th.rand(T, TOTAL, 4, generator=g) * 2 - 1) * 0.5
    a[..., 0] -= 1.0 / 3.0
    return a


def _run(lo, hi, device, gather=None):
    from visfly_b200.envs import RacingEnv2
    env = RacingEnv2(num_agent_per_scene=hi - lo, visual=False, device=device, tensor_output=True, seed=SEED,
                     dynamics_kwargs=dict(DYN), max_episode_steps=MAX_STEPS,
                     shard=None if (lo, hi) == (0, TOTAL) else (lo, TOTAL))
    th.manual_seed(SEED)
    env.reset()
    acts = _actions()[:, lo:hi].to(device)
    out = []
    returns = th.zeros(hi - lo, device=device)
    fused_returns = []
    for t in range(T):
        last_of_rollout = gather is not None and t % 4 == 3          # a "rollout" of 4 steps, then the all-gather
        if last_of_rollout:
            gather.arm(env)
        obs, r, d, info = env.step(acts[t])
        if last_of_rollout:
            fused_returns.append(gather.finish(env).clone())
        else:
            fused_returns.append(env._rewards.clone())
        returns = returns + r
        out.append(th.cat([obs["state"], obs["gate"].float(), r.unsqueeze(1), d.float().unsqueeze(1)], 1).cpu())
    assert env._fused.active
    return th.stack(out), returns, fused_returns


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nccl, out):
    import torch.distributed as dist
    from visfly_b200.distributed import gather_episode_returns, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    device = th.device("cuda", rank if nccl else 0)
    th.cuda.set_device(device)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world,
                            **({"device_id": device} if nccl else {}))
    try:
        lo, hi = shard_range(TOTAL, rank, world)
        fg = None
        if nccl:
            from visfly_b200.distributed import FusedReturnsGather
            fg = FusedReturnsGather(hi - lo, TOTAL, rank, world, device)
        rows, returns, fused_returns = _run(lo, hi, device, fg)
        full = gather_episode_returns(returns if nccl else returns.cpu(), n_total=TOTAL)
        out[rank] = (lo, hi, rows, full.cpu().clone(), [x.cpu() for x in fused_returns],
                     None if fg is None else (fg.fused, fg.why_not))
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_one_gpu_racing_batch_bitwise():
    import torch.multiprocessing as mp
    whole, whole_returns, whole_acc = _run(0, TOTAL, th.device("cuda", 0))
    assert bool(whole[MAX_STEPS - 1][:, -1].all())          # everybody truncated once: Philox restarts are crossed
    nccl = th.cuda.device_count() >= 2
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), nccl, out), nprocs=world, join=True)
        res = dict(out)
    assert sorted(res) == [0, 1]
    for rank, (lo, hi, rows, full, fused_returns, how) in res.items():
        assert th.equal(rows, whole[:, lo:hi]), f"rank {rank}: shard rows differ from the whole batch"
        assert th.equal(full, whole_returns.cpu()), f"rank {rank}: gathered returns differ"
        if how is not None:
            # the all-gather fused into the rollout's last env step (peer-memory stores + barrier), or its NCCL
            # fallback: either way every rank ends up with the whole batch's per-agent episode accumulators
            print(f"rank {rank}: fused gather = {how}")
            for t in range(3, T, 4):
                assert th.equal(fused_returns[t], whole_acc[t].cpu()), (rank, t, how)
