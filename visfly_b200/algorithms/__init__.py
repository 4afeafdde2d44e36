"""Analytic-gradient trainers on top of the differentiable env step: BPTT and SHAC restated
without stable-baselines3 (reference utils/algorithms/BPTT.py, shac.py, common.py:893-923)."""
from .bptt import BPTT
from .common import RolloutBuffer, all_reduce_gradients, compute_td_returns, polyak_update
from .policies import Actor, ActorCritic, TwinCritic
from .shac import SHAC

__all__ = ["BPTT", "SHAC", "Actor", "TwinCritic", "ActorCritic", "RolloutBuffer", "compute_td_returns",
           "polyak_update", "all_reduce_gradients"]
