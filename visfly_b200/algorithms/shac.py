"""SHAC — short-horizon actor-critic with analytic policy gradients through the differentiable env step
(restatement of reference utils/algorithms/shac.py:184-326 without stable-baselines3).

One update = one horizon of H env steps with autograd history (forward: one fused launch per step, backward: one
adjoint launch per step, see envs/base/fused.py), actor loss ``-sum_t gamma^t r_t - gamma^H V(s_H)``, one backward
through the whole horizon, then ``gradient_steps`` critic regressions onto TD(lambda) targets.  With several ranks
(torchrun, one process per GPU, agents sharded) the actor / critic gradients are averaged with one all-reduce per
update (``all_reduce_gradients``); there is no other communication.
"""
from __future__ import annotations

import time
from collections import deque
from copy import deepcopy
from typing import Any, Dict, List, Optional

import torch as th

from .common import RolloutBuffer, all_reduce_gradients, broadcast_parameters, polyak_update
from .policies import ActorCritic


def episode_done_of(info, env) -> th.Tensor:
    """Per-agent ``info[i]["episode_done"]`` of the step just taken as one device tensor (the reference loops over
    the info list in Python, shac.py:229-230)."""
    fn = getattr(info, "episode_done_tensor", None)
    if fn is not None:
        return fn()
    return th.tensor([bool(info[i]["episode_done"]) for i in range(len(info))], device=env.device)


class AnalyticGradientBase:
    name = "SHAC"

    def __init__(self, env, policy="MultiInputPolicy", policy_kwargs: Optional[Dict[str, Any]] = None,
                 learning_rate: float = 1e-3, logger_kwargs=None, comment: Optional[str] = None,
                 save_path: Optional[str] = None, dump_step: int = int(1e4), horizon: int = 32, tau: float = 0.005,
                 gamma: float = 0.99, gradient_steps: int = 5, buffer_size: int = int(1e6),
                 batch_size: int = int(2e5), clip_range_vf: float = 0.1, pre_stop: float = 0.1,
                 policy_noise: float = 0., device=None, seed: int = 42, max_grad_norm: float = 0.5,
                 make_eval_env: bool = True, verbose: int = 0, cuda_graph: bool = False):
        self.env = env
        self.device = th.device(device) if device is not None else env.device
        self.num_envs = env.num_envs
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.H, self.tau, self.gamma = int(horizon), tau, gamma
        self.gradient_steps, self.max_grad_norm = gradient_steps, max_grad_norm
        self.policy_noise, self._dump_step, self.verbose = policy_noise, dump_step, verbose
        self.learning_rate, self.comment, self.save_path = learning_rate, comment, save_path
        #: replay a whole update (horizon forward, backward, gradient exchange, optimiser step) as ONE CUDA graph: an
        #: update is ~80 small launches per env step and is otherwise bound by the host side of those launches
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        th.manual_seed(seed)
        if isinstance(policy, th.nn.Module):
            self.policy = policy.to(self.device)
        else:
            policy_kwargs = dict(policy_kwargs or {})
            if self.cuda_graph:           # optimiser state (step counters) must live on the device to be replayable
                policy_kwargs["optimizer_kwargs"] = dict(policy_kwargs.get("optimizer_kwargs") or {}, capturable=True)
            self.policy = ActorCritic(self.observation_space, self.action_space, learning_rate=learning_rate,
                                      **policy_kwargs).to(self.device)
        broadcast_parameters(self.policy)
        self.actor, self.critic, self.critic_target = self.policy.actor, self.policy.critic, self.policy.critic_target
        self.eval_env = None
        if make_eval_env:
            self.eval_env = deepcopy(env)                      # reference shac.py:121
            self.eval_env.requires_grad = False
            self.eval_env.reset()
        self.env.reset()
        self.env.requires_grad = True                           # reference shac.py:124
        self.rollout_buffer = RolloutBuffer(gamma=gamma)
        self._lo = th.as_tensor(self.action_space.low, device=self.device)
        self._hi = th.as_tensor(self.action_space.high, device=self.device)
        lo, hi = self._lo.flatten().tolist(), self._hi.flatten().tolist()
        #: (low, high) as plain floats when the action box is the same in every dimension (it is: Box(-1, 1, (4,)))
        self._scalar_bounds = (lo[0], hi[0]) if len(set(lo)) == 1 and len(set(hi)) == 1 else None
        self.history: List[Dict[str, float]] = []
        self.num_timesteps = 0

    # -- pieces shared by SHAC and BPTT ---------------------------------------------------------------------
    def _act(self, obs):
        if self.policy_noise == 0 and self._scalar_bounds is not None and hasattr(self.actor, "deterministic_action"):
            return self.actor.deterministic_action(obs, *self._scalar_bounds)      # one launch where supported
        actions, _, _ = self.actor.action_log_prob(obs, noise_scale=self.policy_noise)
        return th.clip(actions, self._lo, self._hi)

    def _actor_update(self, actor_loss: th.Tensor):
        self.actor.optimizer.zero_grad()
        actor_loss.backward()
        if hasattr(self.actor, "release_graph"):
            self.actor.release_graph()
        all_reduce_gradients(self.actor.parameters())
        th.nn.utils.clip_grad_norm_(self.actor.parameters(), self.max_grad_norm)
        self.actor.optimizer.step()

    def _critic_update(self) -> float:
        buf, loss = self.rollout_buffer, 0.0
        for _ in range(self.gradient_steps):
            values, _ = th.cat(self.critic(buf.obs, buf.action), dim=-1).min(dim=-1)
            loss = th.nn.functional.mse_loss(buf.returns.view_as(values), values)
            self.critic.optimizer.zero_grad()
            loss.backward()
            all_reduce_gradients(self.critic.parameters())
            th.nn.utils.clip_grad_norm_(self.critic.parameters(), self.max_grad_norm)
            self.critic.optimizer.step()
            polyak_update(self.critic.parameters(), self.critic_target.parameters(), self.tau)
        return float(loss)

    @th.no_grad()
    def evaluate(self, max_steps: Optional[int] = None) -> Dict[str, float]:
        """Every agent of the evaluation env flies one episode with the deterministic policy
        (reference shac.py:286-303)."""
        env = self.eval_env
        env.reset_agent_by_id()
        obs = env.get_observation()
        pending = th.ones(env.num_envs, dtype=th.bool, device=env.device)
        ret = th.zeros(env.num_envs, device=env.device)
        length = th.zeros(env.num_envs, device=env.device)
        success = th.zeros(env.num_envs, dtype=th.bool, device=env.device)
        ep_ret = th.zeros(env.num_envs, device=env.device)
        for _ in range(max_steps or env.max_episode_steps + 1):
            a = th.clip(self.actor(obs, deterministic=True)[0], self._lo, self._hi)
            obs, reward, done, info = env.step(a, is_test=True)
            ep_ret = ep_ret + reward * pending
            first = done & pending
            ret = th.where(first, ep_ret, ret)
            length = th.where(first, env._step_count.to(th.float32), length)
            success = success | (first & env.success)
            pending = pending & ~done
            if not bool(pending.any()):
                break
        fin = ~pending
        k = fin.sum().clamp_min(1)
        return {"ep_rew_mean": float((ret * fin).sum() / k), "ep_len_mean": float((length * fin).sum() / k),
                "success_rate": float((success & fin).sum() / k), "episodes": int(fin.sum())}

    def _log(self, **kw):
        self.history.append(kw)
        if self.verbose:
            print(f"[{self.name}] " + "  ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}"
                                                for k, v in kw.items()), flush=True)

    def save(self, path: str):
        th.save(self.policy.state_dict(), path)

    def load(self, path: str):
        self.policy.load_state_dict(th.load(path, map_location=self.device))


class SHAC(AnalyticGradientBase):
    name = "SHAC"

    def learn(self, total_timesteps: int):
        self.policy.train()
        start, last_dump, t0 = self.num_timesteps, self.num_timesteps, time.time()
        n = self.num_envs
        while self.num_timesteps - start < total_timesteps:
            actor_loss = 0.0
            discount = th.ones((n,), dtype=th.float32, device=self.device)
            for inner_step in range(self.H):
                pre_obs = self.env.get_observation()
                action = self._act(pre_obs)
                obs, reward, done, info = self.env.step(action)
                episode_done = episode_done_of(info, self.env)
                self.num_timesteps += n
                # bootstrap value of the state reached (target critic, detached)                   shac.py:236-241
                next_action = th.clip(self.actor(obs, deterministic=True)[0], self._lo, self._hi)
                next_value, _ = th.cat(self.critic_target(obs.detach(), next_action.detach()), dim=-1).min(dim=-1)
                actor_loss = actor_loss - reward * discount
                boot = (done | (inner_step == self.H - 1)) & ~episode_done                         # shac.py:246
                actor_loss = actor_loss - next_value * discount * self.gamma * boot
                discount = discount * self.gamma * ~done + done
                self.rollout_buffer.add(obs=pre_obs.detach(), reward=reward.detach(), action=action.detach(),
                                        next_obs=obs.detach(), done=done.clone(), episode_done=episode_done,
                                        value=next_value.detach())
            actor_loss = actor_loss.mean()
            self._actor_update(actor_loss)
            self.rollout_buffer.compute_returns()
            self.env.detach()
            critic_loss = self._critic_update()
            self.rollout_buffer.clear()
            if self.num_timesteps - last_dump >= self._dump_step:
                rec = {"timesteps": self.num_timesteps, "actor_loss": float(actor_loss), "critic_loss": critic_loss,
                       "fps": (self.num_timesteps - last_dump) / max(time.time() - t0, 1e-9)}
                if self.eval_env is not None:
                    rec.update(self.evaluate())
                self._log(**rec)
                last_dump, t0 = self.num_timesteps, time.time()
        return self.policy
