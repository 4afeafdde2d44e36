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


class EpisodeReturnsGather:
    """The one collective of the path: all-gather of the per-agent episode returns of every shard,
    once per rollout, ``(n_local,) -> (n_total,)`` in rank order.

    Everything that can be decided ahead of the rollout is decided here: shard sizes follow from ``shard_range``
    (no size exchange, no ``.item()``), the receive buffer is allocated once, and ``__call__`` is ONE
    ``all_gather_into_tensor`` enqueued on the current stream — it never synchronises with the host, so the launches
    queued behind it keep flowing.  The returned tensor is valid in stream order and is overwritten by the next call.
    Shards that differ in size by one (``n_total % world != 0``) go through a ``base+1``-wide staging row and one
    precomputed ``index_select``; equal shards take the direct path with no extra kernel."""

    def __init__(self, n_total: int, rank: int, world: int, device, dtype=th.float32, group=None):
        self.n_total, self.rank, self.world, self.group = int(n_total), int(rank), int(world), group
        lo, hi = shard_range(n_total, rank, world)
        self.n_local = hi - lo
        base, extra = divmod(self.n_total, self.world)
        self.ragged = extra != 0
        if not self.ragged:
            self.out = th.empty((self.n_total,), device=device, dtype=dtype)
        else:
            self.width = base + 1
            self.stage = th.zeros((self.width,), device=device, dtype=dtype)
            self.wide = th.empty((self.world * self.width,), device=device, dtype=dtype)
            idx = [r * self.width + k for r in range(world) for k in range(base + (1 if r < extra else 0))]
            self.index = th.tensor(idx, device=device, dtype=th.int64)
            self.out = th.empty((self.n_total,), device=device, dtype=dtype)

    def __call__(self, returns: th.Tensor) -> th.Tensor:
        if returns.numel() != self.n_local:
            raise ValueError(f"rank {self.rank} owns {self.n_local} agents, got {returns.numel()} returns")
        src = returns.detach().reshape(-1)
        if self.world == 1:
            return src
        if not self.ragged:
            dist.all_gather_into_tensor(self.out, src.contiguous(), group=self.group)
            return self.out
        self.stage[:self.n_local].copy_(src)
        dist.all_gather_into_tensor(self.wide, self.stage, group=self.group)
        th.index_select(self.wide, 0, self.index, out=self.out)
        return self.out


_GATHERS = {}


def gather_episode_returns(returns: th.Tensor, n_total: Optional[int] = None, group=None) -> th.Tensor:
    """All-gather the per-agent episode returns of every shard: ``(n_local,) -> (n_total,)`` in rank order, one
    asynchronous ``all_gather_into_tensor`` (see ``EpisodeReturnsGather``; the gather object is cached per shape).
    ``n_total`` = number of agents over all ranks; ``None`` means equal shards (``n_local * world``).  Shards must
    follow ``shard_range``.  Single-process runs return the input."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return returns
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_total = returns.numel() * world if n_total is None else int(n_total)
    key = (n_total, rank, world, returns.device, returns.dtype, id(group))
    g = _GATHERS.get(key)
    if g is None:
        g = _GATHERS[key] = EpisodeReturnsGather(n_total, rank, world, returns.device, returns.dtype, group)
    return g(returns)


def rollout_stats(ep_return_sum: th.Tensor, ep_len_sum: th.Tensor, ep_count: th.Tensor, group=None):
    """Global mean episode return / length for logging (one 3-scalar all-reduce)."""
    v = th.stack([ep_return_sum.double(), ep_len_sum.double(), ep_count.double()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, group=group)
    cnt = v[2].clamp_min(1)
    return float(v[0] / cnt), float(v[1] / cnt), int(v[2])
