"""Shared pieces of the analytic-gradient trainers: TD(lambda) returns, the horizon buffer, target-network update and
the data-parallel gradient exchange."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch as th

from ..type import TensorDict


def compute_td_returns(r: Sequence[th.Tensor], done: Sequence[th.Tensor], next_value: Sequence[th.Tensor],
                       episode_done: Optional[Sequence[th.Tensor]] = None, gamma: float = 0.99,
                       lamda: float = 0.95) -> List[th.Tensor]:
    """TD(lambda) targets of one horizon, backward recursion of reference utils/algorithms/common.py:893-923.

    ``r[t]``, ``done[t]`` (bool), ``next_value[t]`` are ``(N,)`` tensors for t = 0..H-1.  ``done`` ends the bootstrap
    chain (time-limit truncation bootstraps from ``next_value``), ``episode_done`` marks real terminations (no
    bootstrap).  Returns the list of H ``(N,)`` targets.
    """
    h = len(r)
    episode_done = done if episode_done is None else episode_done
    dev = r[0].device
    n = r[0].shape[0]
    returns: List[th.Tensor] = [None] * h
    a_i = th.zeros(n, dtype=th.float32, device=dev)
    lam = th.ones(n, dtype=th.float32, device=dev)
    b_i = next_value[-1] * (~done[-1])
    for t in reversed(range(h)):
        active, ended = ~done[t], done[t]
        lam = lam * lamda * active + ended
        a_i = active * (lamda * gamma * a_i + gamma * next_value[t] + ((1.0 - lam) / (1.0 - lamda)) * r[t])
        b_i = gamma * (next_value[t] * ended * (~episode_done[t]) + b_i * active) + r[t]
        returns[t] = (1.0 - lamda) * a_i + lam * b_i
    return returns


class RolloutBuffer:
    """One horizon of detached transitions for the critic update (reference ``SimpleRolloutBuffer``,
    common.py:1198-1249)."""

    def __init__(self, gamma: float):
        self.gamma = gamma
        self.clear()

    def clear(self):
        self.obs, self.action, self.reward, self.next_obs = [], [], [], []
        self.done, self.episode_done, self.value, self.returns = [], [], [], []

    def add(self, obs, reward, action, next_obs, done, episode_done, value):
        self.obs.append(obs)
        self.reward.append(reward)
        self.action.append(action)
        self.next_obs.append(next_obs)
        self.done.append(done)
        self.episode_done.append(episode_done)
        self.value.append(value)

    def compute_returns(self):
        # the reference hard-codes gamma=0.99 here (common.py:1238); kept
        self.returns = compute_td_returns(r=self.reward, done=self.done, next_value=self.value,
                                          episode_done=self.episode_done, gamma=0.99)
        self.flatten()

    def flatten(self):
        self.reward = th.vstack(self.reward).flatten()
        # (H, N, ...) -> (H*N, ...): rows line up with the flattened action / return vectors
        self.obs = TensorDict({k: v.flatten(0, 1) for k, v in TensorDict.stack(self.obs).items()})
        self.action = th.vstack(self.action)
        self.next_obs = TensorDict({k: v.flatten(0, 1) for k, v in TensorDict.stack(self.next_obs).items()})
        self.done = th.vstack(self.done).flatten()
        self.episode_done = th.vstack(self.episode_done).flatten()
        self.returns = th.vstack(self.returns).flatten()


@th.no_grad()
def polyak_update(params: Iterable[th.Tensor], target_params: Iterable[th.Tensor], tau: float):
    """target <- (1 - tau) * target + tau * source."""
    for p, tp in zip(params, target_params):
        tp.mul_(1.0 - tau).add_(p, alpha=tau)


def all_reduce_gradients(params: Iterable[th.Tensor], group=None) -> int:
    """Data-parallel exchange of one update: gradients of all parameters are averaged over the ranks of ``group``
    with ONE all-reduce of one flat bucket (NCCL over NVLink on GPUs, gloo in the CPU tests).  Agents are sharded
    across ranks and every rank's loss is the mean over ITS agents, so the average of the rank gradients is the
    gradient of the mean over all agents.  Returns the number of ranks (1 = nothing to do)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = th.zeros_like(p)
    flat = th.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for p in params:
        k = p.numel()
        p.grad.copy_(flat[off:off + k].view_as(p.grad))
        off += k
    return world


def broadcast_parameters(module: th.nn.Module, src: int = 0, group=None):
    """Every rank starts from rank ``src``'s weights."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
