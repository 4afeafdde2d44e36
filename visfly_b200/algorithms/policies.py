"""Minimal actor / critic networks for the analytic-gradient trainers (stand-ins for the reference's SB3-derived
``MultiInputPolicy``, utils/policies/td_policies.py): observation dicts are flattened and concatenated, the actor is a
tanh-squashed Gaussian sampled with the reparameterisation trick so the action stays differentiable, the critic is
a pair of Q networks."""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, Optional, Sequence, Tuple, Type

import torch as th
from torch import nn

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0


def flatten_obs(obs) -> th.Tensor:
    """``TensorDict`` / dict of ``(..., k_i)`` tensors -> ``(..., sum k_i)`` (keys in sorted order); tensors pass."""
    if isinstance(obs, th.Tensor):
        return obs
    parts = []
    for k in sorted(obs.keys()):
        v = obs[k]
        v = v.to(th.float32)
        parts.append(v.unsqueeze(-1) if v.dim() == 1 else v)
    return parts[0] if len(parts) == 1 else th.cat(parts, dim=-1)


def obs_dim(observation_space) -> int:
    spaces = getattr(observation_space, "spaces", None)
    if spaces is None:
        return int(observation_space.shape[-1])
    return int(sum(int(th.tensor(s.shape).prod()) if len(s.shape) else 1 for _, s in sorted(spaces.items())))


def mlp(sizes: Sequence[int], activation: Type[nn.Module], out_activation: Optional[Type[nn.Module]] = None):
    layers = []
    for i in range(len(sizes) - 1):
        layers.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2:
            layers.append(activation())
        elif out_activation is not None:
            layers.append(out_activation())
    return nn.Sequential(*layers)


class Actor(nn.Module):
    def __init__(self, in_dim: int, act_dim: int = 4, net_arch: Sequence[int] = (64, 64),
                 activation_fn: Type[nn.Module] = nn.Tanh, log_std_init: float = -2.0):
        super().__init__()
        self.body = mlp([in_dim, *net_arch], activation_fn, activation_fn)
        self.mu = nn.Linear(net_arch[-1], act_dim)
        self.log_std = nn.Parameter(th.full((act_dim,), float(log_std_init)))
        self.optimizer: Optional[th.optim.Optimizer] = None

    def _dist(self, obs) -> Tuple[th.Tensor, th.Tensor]:
        h = self.body(flatten_obs(obs))
        return self.mu(h), self.log_std.clamp(LOG_STD_MIN, LOG_STD_MAX)

    def forward(self, obs, deterministic: bool = True):
        """-> (action in [-1,1], hidden)"""
        mean, log_std = self._dist(obs)
        if deterministic:
            return th.tanh(mean), None
        return th.tanh(mean + th.randn_like(mean) * log_std.exp()), None

    def action_log_prob(self, obs, noise_scale: float = 1.0):
        """Reparameterised sample, its log-probability and the hidden features (reference actor surface)."""
        mean, log_std = self._dist(obs)
        std = log_std.exp() * noise_scale
        pre = mean + th.randn_like(mean) * std if noise_scale > 0 else mean
        act = th.tanh(pre)
        if noise_scale > 0:
            logp = (-0.5 * ((pre - mean) / std).pow(2) - std.log() - 0.9189385332046727).sum(-1)
            logp = logp - th.log(1 - act.pow(2) + 1e-6).sum(-1)
        else:
            logp = th.zeros(act.shape[:-1], device=act.device)
        return act, logp, None


class TwinCritic(nn.Module):
    def __init__(self, in_dim: int, act_dim: int = 4, net_arch: Sequence[int] = (64, 64),
                 activation_fn: Type[nn.Module] = nn.Tanh, n_critics: int = 2):
        super().__init__()
        self.q = nn.ModuleList([mlp([in_dim + act_dim, *net_arch, 1], activation_fn) for _ in range(n_critics)])
        self.optimizer: Optional[th.optim.Optimizer] = None

    def forward(self, obs, action):
        x = th.cat([flatten_obs(obs), action], dim=-1)
        return tuple(q(x) for q in self.q)


class ActorCritic(nn.Module):
    """``policy.actor / policy.critic / policy.critic_target`` with their optimisers, as the trainers expect."""

    def __init__(self, observation_space, action_space, learning_rate: float = 1e-3, net_arch: Sequence[int] = (64, 64),
                 activation_fn: Type[nn.Module] = nn.Tanh, log_std_init: float = -2.0,
                 optimizer_class: Type[th.optim.Optimizer] = th.optim.Adam, optimizer_kwargs: Optional[Dict] = None):
        super().__init__()
        in_dim, act_dim = obs_dim(observation_space), int(action_space.shape[-1])
        self.actor = Actor(in_dim, act_dim, net_arch, activation_fn, log_std_init)
        self.critic = TwinCritic(in_dim, act_dim, net_arch, activation_fn)
        self.critic_target = deepcopy(self.critic)
        for p in self.critic_target.parameters():
            p.requires_grad_(False)
        kw = dict(optimizer_kwargs or {})
        self.actor.optimizer = optimizer_class(self.actor.parameters(), lr=learning_rate, **kw)
        self.critic.optimizer = optimizer_class(self.critic.parameters(), lr=learning_rate, **kw)

    def predict(self, obs, deterministic: bool = True):
        with th.no_grad():
            return self.actor(obs, deterministic=deterministic)[0]
