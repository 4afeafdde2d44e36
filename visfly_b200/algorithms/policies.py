"""Minimal actor / critic networks for the analytic-gradient trainers (stand-ins for the reference's SB3-derived
``MultiInputPolicy``, utils/policies/td_policies.py): observation dicts are flattened and concatenated, the actor is a
tanh-squashed Gaussian sampled with the reparameterisation trick so the action stays differentiable, the critic is
a pair of Q networks."""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, Optional, Sequence, Tuple, Type

import weakref

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


def _slices(rows: int, max_slices: int = 128, min_rows: int = 256) -> int:
    """Number of equal row slices for the weight-gradient contraction: the largest power of two <= max_slices that
    divides ``rows`` and leaves at least ``min_rows`` rows per slice (1 = do not slice)."""
    b = 1
    while b * 2 <= max_slices and rows % (b * 2) == 0 and rows // (b * 2) >= min_rows:
        b *= 2
    return b


class _WideBatchLinearFn(th.autograd.Function):
    """``y = x W^T + b`` for a batch of tens of thousands of agents and a few dozen features.

    The forward and the input gradient are ordinary library GEMMs.  The weight gradient ``dW = dy^T x`` contracts over
    the batch: a (64 x 65 536) . (65 536 x 64) product has ONE output tile, and the library's split-K choice for it
    runs at ~280 GB/s on a B200 (121 us per layer per step — 58 % of the GPU time of a BPTT update, see
    profiles/r01_final_launches_apg.txt).  Here the batch is cut into up to 128 equal slices, one batched GEMM
    produces the per-slice products (one CTA tile each, all SMs busy) and a small reduction adds them."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gy @ weight
        g2, x2 = gy.reshape(-1, gy.shape[-1]), x.reshape(-1, x.shape[-1])
        if ctx.needs_input_grad[1]:
            b = _slices(x2.shape[0])
            if b > 1:
                gw = th.bmm(g2.view(b, -1, g2.shape[1]).transpose(1, 2), x2.view(b, -1, x2.shape[1])).sum(0)
            else:
                gw = g2.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


class WideBatchLinear(nn.Linear):
    """``nn.Linear`` (same parameters, same initialisation, same state dict) with the sliced weight gradient."""

    def forward(self, x):
        if x.requires_grad or self.weight.requires_grad:
            return _WideBatchLinearFn.apply(x, self.weight, self.bias)
        return nn.functional.linear(x, self.weight, self.bias)


def mlp(sizes: Sequence[int], activation: Type[nn.Module], out_activation: Optional[Type[nn.Module]] = None):
    layers = []
    for i in range(len(sizes) - 1):
        layers.append(WideBatchLinear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2:
            layers.append(activation())
        elif out_activation is not None:
            layers.append(out_activation())
    return nn.Sequential(*layers)


_FLAT_CACHE = weakref.WeakKeyDictionary()


class _FusedActorFn(th.autograd.Function):
    """``clip(tanh(W3 tanh(W2 tanh(W1 x + b1) + b2) + b3), lo, hi)`` as ONE launch forward (``vf_policy_fwd``) and one
    backward (``vf_policy_bwd`` + a fixed-order reduction of the per-tile weight gradients) instead of ~40 library
    launches per env step.  ``x = [xa | xb]`` may arrive in two pieces (no concatenated copy); the backward recomputes
    the activations from the inputs, nothing else is saved."""

    @staticmethod
    def forward(ctx, xa, xb, lo, hi, packed, flat, h):
        from .. import _lib
        xa = xa.contiguous()
        xb = None if xb is None else xb.contiguous()
        ctx.save_for_backward(xa, packed, *(() if xb is None else (xb,)))
        ctx.lo, ctx.hi, ctx.two, ctx.h = lo, hi, xb is not None, h
        return _lib.policy_fwd(xa, xb, packed, h, lo, hi)

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, g_action):
        from .. import _lib
        saved = ctx.saved_tensors
        xa, packed, xb = saved[0], saved[1], (saved[2] if ctx.two else None)
        g_a, g_b, flat = _lib.policy_bwd(xa, xb, packed, ctx.h, ctx.lo, ctx.hi, g_action.contiguous(),
                                         ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        # `flat` = [dW1 | db1 | dW2 | db2 | dW3 | db3]: the gradient of the concatenated parameter vector the forward
        # was given (Actor._flat_params) — ONE accumulation per env step instead of six
        return g_a, g_b, None, None, None, flat, None


class Actor(nn.Module):
    def __init__(self, in_dim: int, act_dim: int = 4, net_arch: Sequence[int] = (64, 64),
                 activation_fn: Type[nn.Module] = nn.Tanh, log_std_init: float = -2.0):
        super().__init__()
        self.body = mlp([in_dim, *net_arch], activation_fn, activation_fn)
        self.mu = WideBatchLinear(net_arch[-1], act_dim)
        self.log_std = nn.Parameter(th.full((act_dim,), float(log_std_init)))
        self.optimizer: Optional[th.optim.Optimizer] = None

    def fused_ok(self, x: th.Tensor) -> bool:
        """True if the one-launch actor kernels cover this network: two tanh hidden layers of equal width 32 or 64, at
        most 32 inputs, four outputs, float32 on a CUDA device."""
        ok = self.__dict__.get("_fused_ok")
        if ok is None:
            lin = [m for m in self.body if isinstance(m, nn.Linear)]
            act = [m for m in self.body if not isinstance(m, nn.Linear)]
            ok = (len(lin) == 2 and len(act) == 2 and all(isinstance(m, nn.Tanh) for m in act)
                  and lin[0].out_features == lin[1].out_features == lin[1].in_features and lin[0].out_features in (32, 64)
                  and lin[0].in_features <= 32 and self.mu.out_features == 4 and self.mu.in_features == lin[1].out_features
                  and all(m.bias is not None for m in lin + [self.mu]))
            self.__dict__["_fused_ok"] = ok
        return bool(ok) and x.is_cuda and x.dtype is th.float32 and x.dim() == 2 and self.mu.weight.is_cuda

    def deterministic_action(self, obs, lo: float = -1.0, hi: float = 1.0) -> th.Tensor:
        """``clip(tanh(mean(obs)), lo, hi)`` — the noise-free action of the trainers' rollouts; one kernel each way where
        ``fused_ok``, the library ops otherwise (same function, same gradients).  An observation dict of two float32
        matrices (e.g. ``{"state": (N,13), "target": (N,3)}``) is handed to the kernel piecewise, in ``flatten_obs``'s
        key order, without being concatenated."""
        pieces = None
        if not isinstance(obs, th.Tensor) and len(obs.keys()) == 2:      # (len() of a TensorDict is its batch size)
            pieces = [obs[k] for k in sorted(obs.keys())]
            if not all(p.dim() == 2 and p.dtype is th.float32 and p.is_cuda for p in pieces) or \
                    pieces[0].shape[1] + pieces[1].shape[1] != self.body[0].in_features:
                pieces = None
        if pieces is not None and self.fused_ok(pieces[0]):
            xa, xb = pieces
        else:
            xa, xb = flatten_obs(obs), None
        if self.fused_ok(xa):
            l1, l2 = self.body[0], self.body[2]
            params = (l1.weight, l1.bias, l2.weight, l2.bias, self.mu.weight, self.mu.bias)
            return _FusedActorFn.apply(xa, xb, float(lo), float(hi), self._packed_weights(params), self._flat_params(params),
                                       l1.out_features)
        return th.clip(th.tanh(self.mu(self.body(xa))), lo, hi)

    def _packed_weights(self, params) -> th.Tensor:
        """The weights in the kernels' tile layouts (``vf_policy_pack``), rebuilt when a parameter has changed (its
        version counter moves with every optimiser step / ``load_state_dict`` / in-place edit): once per update, not
        once per env step.  Under CUDA-graph capture the cache is bypassed, so that the pack launch is part of the
        captured update and every replay repacks the weights its optimiser step has just written."""
        from .. import _lib
        key = tuple((p.data_ptr(), p._version) for p in params)
        cache = self.__dict__.get("_packed")
        capturing = th.cuda.is_current_stream_capturing()
        if cache is None or cache[0] != key or capturing != cache[2]:
            cache = (key, _lib.policy_pack(params), capturing)
            self.__dict__["_packed"] = cache
        return cache[1]

    def _flat_params(self, params) -> th.Tensor:
        """``cat`` of the six parameter tensors, in the order of the kernels' gradient vector — the one autograd input
        through which ``_FusedActorFn`` hands the parameter gradients back (the horizon's H contributions are summed on
        this one tensor, the split into per-parameter gradients happens once per update).  Rebuilt with the packed
        weights, and whenever the autograd mode differs from the one it was built under."""
        key = (tuple((p.data_ptr(), p._version) for p in params), th.is_grad_enabled(),
               th.cuda.is_current_stream_capturing())
        cache = _FLAT_CACHE.get(self)          # not an attribute: a tensor with autograd history cannot be deep-copied
        if cache is None or cache[0] != key:
            cache = (key, th.cat([p.reshape(-1) for p in params]))
            _FLAT_CACHE[self] = cache
        return cache[1]

    def release_graph(self):
        """Drop the cached parameter vector (and with it the autograd nodes it keeps alive); the trainers call this
        after every backward pass."""
        _FLAT_CACHE.pop(self, None)

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
