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
        if self.world == 1:
            return returns
        src = returns.detach().reshape(-1)
        if not self.ragged:
            dist.all_gather_into_tensor(self.out, src.contiguous(), group=self.group)
            return self.out
        self.stage[:self.n_local].copy_(src)
        dist.all_gather_into_tensor(self.wide, self.stage, group=self.group)
        th.index_select(self.wide, 0, self.index, out=self.out)
        return self.out


class FusedReturnsGather:
    """The rollout's all-gather of episode returns, fused into the rollout's last env step.

    ``EpisodeReturnsGather`` launches a collective behind the last step (copy to a contiguous send buffer, NCCL
    all-gather, two stream hand-overs: ~40 us on 2 GPUs, as long as three env steps).  Here the last step's own launch
    stores each agent's return directly into the gather buffer of every rank through peer-mapped device memory
    (``torch.distributed._symmetric_memory``: NVLink / NVSwitch P2P mappings exchanged once at construction), so the
    transfer rides inside the step kernel and what is left of the collective is one cross-rank barrier on the stream.

        gather = FusedReturnsGather(n_local, n_total, rank, world, device)    # once (collective: all ranks)
        ... K-1 x env.step(a) ...
        gather.arm(env)                                               # the NEXT env.step also scatters the returns
        env.step(a)
        returns = gather.finish(env)                                  # (n_total,) in rank order, valid in stream order

    Two buffers alternate between rollouts, so a rank that is already scattering rollout k+1 never writes into the
    buffer a slower rank is still reading for rollout k (``finish`` keeps ranks within one rollout of each other).
    Falls back to ``EpisodeReturnsGather`` (NCCL) where peer mappings are unavailable, or for envs on the generic
    tensor-op path; ``self.fused`` says which one is in use."""

    def __init__(self, n_local: int, n_total: int, rank: int, world: int, device, group=None):
        from .params import MAX_PEERS, VfPeerScatter
        self.n_local, self.n_total, self.rank, self.world = int(n_local), int(n_total), int(rank), int(world)
        self.device = th.device(device)
        lo, hi = shard_range(n_total, rank, world)
        if hi - lo != self.n_local:
            raise ValueError(f"rank {rank} owns {hi - lo} of {n_total} agents, not {n_local}")
        self.offset = lo
        self.fallback = EpisodeReturnsGather(n_total, rank, world, self.device, group=group)
        self.fused, self.why_not = False, None
        self._turn, self._armed = 0, False
        if world == 1:
            self.why_not = "single process"
            return
        if world > MAX_PEERS:
            self.why_not = f"more than {MAX_PEERS} ranks"
            return
        try:
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else dist.group.WORLD
            self._bufs, self._hdls, self._peers = [], [], []
            for _ in range(2):
                buf = symm.empty(self.n_total, dtype=th.float32, device=self.device)
                hdl = symm.rendezvous(buf, grp)
                ptrs = list(hdl.buffer_ptrs)
                if len(ptrs) != world or not all(ptrs):
                    raise RuntimeError("incomplete peer mapping")
                ps = VfPeerScatter()
                for r, ptr in enumerate(ptrs):
                    ps.dst[r] = ptr
                ps.offset, ps.world = lo, world
                buf.zero_()
                self._bufs.append(buf); self._hdls.append(hdl); self._peers.append(ps)
            th.cuda.synchronize(self.device)
            self._hdls[0].barrier(channel=0)
            self.fused = True
        except Exception as e:          # no P2P / fabric handles on this box: the NCCL collective does the job
            self.why_not = repr(e)[:200]
            self.fused = False
        # every rank must take the same path (the barrier is collective)
        flag = th.tensor([1 if self.fused else 0], device=self.device, dtype=th.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0 and self.fused:
            self.fused, self.why_not = False, "a peer could not map the buffers"

    def arm(self, env):
        """The next ``env.step`` scatters the accumulated episode returns to every rank (fused path only)."""
        fz = getattr(env, "_fused", None)
        self._armed = bool(self.fused and fz is not None and env.use_fused_step and env.num_agent == self.n_local)
        if self._armed:
            import ctypes
            fz.peer_next = ctypes.addressof(self._peers[self._turn])

    def finish(self, env) -> th.Tensor:
        """All ranks' returns ``(n_total,)`` in rank order; valid in stream order, overwritten two rollouts later."""
        fz = getattr(env, "_fused", None)
        if self._armed and fz is not None and fz.peer_next == 0 and fz.peer_done:
            fz.peer_done = False
            buf, hdl = self._bufs[self._turn], self._hdls[self._turn]
            hdl.barrier(channel=0)                   # every rank's step kernel has finished storing
            self._turn ^= 1
            self._armed = False
            return buf
        if fz is not None:
            fz.peer_next = 0
        self._armed = False
        return self.fallback(env._rewards)


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
