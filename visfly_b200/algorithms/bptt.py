"""BPTT — back-propagation through time of the discounted return over a horizon of differentiable env steps
(restatement of reference utils/algorithms/BPTT.py:77-180 without stable-baselines3): no critic, the actor loss is
``-sum_t discount_t * r_t`` with the discount restarting at 1 for agents whose episode ended inside the horizon."""
from __future__ import annotations

import time

import torch as th

from .shac import AnalyticGradientBase


class BPTT(AnalyticGradientBase):
    name = "BPTT"

    def rollout_loss(self) -> th.Tensor:
        """One horizon with autograd history; returns the mean actor loss (reference BPTT.py:107-127)."""
        n = self.num_envs
        actor_loss = 0.0
        discount = th.ones((n,), dtype=th.float32, device=self.device)
        for _ in range(self.H):
            obs = self.env.get_observation()
            action = self._act(obs)
            obs, reward, done, info = self.env.step(action)
            self.num_timesteps += n
            actor_loss = actor_loss - reward * discount
            discount = discount * self.gamma * ~done + done
        return actor_loss.mean()

    def learn(self, total_timesteps: int):
        assert self.H >= 1, "horizon must be at least 1"
        self.policy.train()
        start, last_dump, t0 = self.num_timesteps, self.num_timesteps, time.time()
        while self.num_timesteps - start < total_timesteps:
            actor_loss = self.rollout_loss()
            self._actor_update(actor_loss)
            self.env.detach()
            if self.num_timesteps - last_dump >= self._dump_step:
                rec = {"timesteps": self.num_timesteps, "actor_loss": float(actor_loss),
                       "fps": (self.num_timesteps - last_dump) / max(time.time() - t0, 1e-9)}
                if self.eval_env is not None:
                    rec.update(self.evaluate())
                self._log(**rec)
                last_dump, t0 = self.num_timesteps, time.time()
        return self.policy
