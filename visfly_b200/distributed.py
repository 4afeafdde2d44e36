"""Multi-GPU plumbing: agents are independent, so the path shards trivially.

One process per GPU (torchrun), rank r owns the contiguous agent range ``shard_range(N, r, world)`` with its own
packed state, FIFO, env bookkeeping and RNG stream; the kernels are per-agent arithmetic, so results do not depend
on how agents are split (tests: one batch vs. two shards are bit-identical).  There is no per-step communication.
The only collective of the path is one all-gather of per-agent episode returns per rollout (NCCL over
NVLink/NVSwitch on GPUs; the same code runs on gloo/CPU tensors for the host-side tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch as th
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, th.device]:
    """Initialise ``torch.distributed`` from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = th.cuda.is_available() and (backend or "nccl") == "nccl"
    device = th.device("cuda", local) if use_cuda else th.device("cpu")
    if use_cuda:
        th.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if use_cuda else {}
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, device


def bind_host_to_gpu(device_index: int) -> Optional[list]:
    """Pin this process to the CPU cores next to its GPU (NVML's ideal CPU affinity = the NUMA node the GPU's PCIe
    root hangs off).  In numpy mode every rank streams ~4 MB per step into page-locked host memory; with one rank per
    GPU on a two-socket host, pages first-touched on the wrong socket turn each PCIe write into cross-socket traffic.
    Call before the first page-locked allocation.  Returns the CPU list, or None where NVML / the affinity call is
    unavailable (the process then simply stays unpinned)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(th.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:           # no NVML, no permission, exotic topology: run unpinned
        return None


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous agent range of ``rank``; the first ``n_total % world`` ranks take one extra agent."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_seed(seed: int, rank: int) -> int:
    """Per-rank seed: every shard draws its initial states / resets from its own stream."""
    return int(seed) + 1_000_003 * int(rank)


def gather_episode_returns(returns: th.Tensor, group=None) -> th.Tensor:
    """All-gather the per-agent episode returns of every shard: ``(n_local,) -> (sum of n_local over ranks,)`` in
    rank order.  Shards may differ in size by one (see ``shard_range``); single-process runs return the input."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return returns
    world = dist.get_world_size(group)
    n_local = th.tensor([returns.numel()], device=returns.device, dtype=th.int64)
    sizes = [th.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s) for s in sizes]
    width = max(sizes)
    padded = returns.new_zeros(width)
    padded[:returns.numel()] = returns.reshape(-1)
    parts = [returns.new_empty(width) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return th.cat([p[:k] for p, k in zip(parts, sizes)])


def rollout_stats(ep_return_sum: th.Tensor, ep_len_sum: th.Tensor, ep_count: th.Tensor, group=None):
    """Global mean episode return / length for logging (one 3-scalar all-reduce)."""
    v = th.stack([ep_return_sum.double(), ep_len_sum.double(), ep_count.double()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, group=group)
    cnt = v[2].clamp_min(1)
    return float(v[0] / cnt), float(v[1] / cnt), int(v[2])
