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
        n, H = self.num_envs, self.H
        rewards, dones = [], []
        for _ in range(H):
            obs = self.env.get_observation()
            action = self._act(obs)
            obs, reward, done, info = self.env.step(action)
            self.num_timesteps += n
            rewards.append(reward)
            dones.append(done)
        # actor_loss = -sum_t reward_t * discount_t with discount_0 = 1, discount_{t+1} = 1 if done_t else gamma * discount_t
        # (reference BPTT.py:116-127), evaluated once per horizon from the stacked rewards and dones instead of eight small
        # launches (and their autograd nodes) per step: discount_t = gamma ** (steps since the last episode end before t)
        done_t = th.stack(dones)                                                    # (H, n)
        step_no = th.arange(1, H + 1, device=self.device, dtype=th.int32).view(H, 1)
        last_end = th.cummax(th.where(done_t, step_no, 0), dim=0).values            # 1 + index of the last done_j, j <= t
        since = step_no - 1 - th.cat([last_end.new_zeros((1, n)), last_end[:-1]])   # t - (1 + last j < t with done_j)
        discounts = th.pow(th.full((), self.gamma, device=self.device, dtype=th.float32), since.to(th.float32))
        return -(th.stack(rewards) * discounts).sum(0).mean()

    # -- the whole update as one CUDA graph -------------------------------------------------------------------------
    # What makes an update replayable: (i) everything that carries over from one update to the next — packed state,
    # env status records, the comm-delay FIFO, the last observation — lives in fixed holder tensors that the graph
    # reads first and overwrites last; (ii) the in-kernel restart sampler takes its step number from a device word
    # (VfEnvSpec step_base) that the graph itself advances by H, so every replay draws fresh restarts; (iii) the
    # optimiser is constructed with capturable=True; (iv) nothing in the horizon synchronises with the host (lazy info
    # records, no .item()).  torch's generator (policy noise) is graph-safe by itself.
    def _holders_from_env(self):
        dyn, fz = self.env.envs.dynamics, self.env._fused
        self._g_state = dyn._state.detach().clone()
        self._g_status = fz.status.clone()
        self._g_fifo = [a.detach().clone() for a in dyn._pre_action]
        self._g_obs = None if dyn._obs_t is None else dyn._obs_t.detach().clone()

    def _env_to_holders(self):
        """Point the env at the holder tensors (start of an update)."""
        dyn, fz = self.env.envs.dynamics, self.env._fused
        dyn._state, fz.status = self._g_state, self._g_status
        dyn._pre_action = list(self._g_fifo)
        dyn._obs_t = self._g_obs

    def _holders_from_outputs(self):
        """Copy what the update left in the env back into the holders (end of an update, inside the graph)."""
        dyn, fz = self.env.envs.dynamics, self.env._fused
        self._g_state.copy_(dyn._state)
        self._g_status.copy_(fz.status)
        for h, a in zip(self._g_fifo, dyn._pre_action):
            h.copy_(a)
        if self._g_obs is not None:
            self._g_obs.copy_(dyn._obs_t)
        fz.step_base.add_(self.H)

    def _one_update(self):
        loss = self.rollout_loss()
        self._actor_update(loss)
        self.env.detach()
        return loss.detach()

    def _capture(self):
        env = self.env
        fz, dyn = env._fused, env.envs.dynamics
        if fz is None or not fz.active or not fz.refresh():
            raise RuntimeError("cuda_graph=True needs the one-kernel env step (built-in task, fused path active)")
        if fz.t_off is not None or dyn._wind_fn is not None:
            raise NotImplementedError("cuda_graph=True with per-agent time offsets / wind functions")
        # from here on the restart sampler's step number is (host counter) + (device word); the device word takes over
        # the count so far, so that numbers stay unique across eager steps, warm-up updates and replays
        import gc
        self.actor.optimizer.zero_grad(set_to_none=True)
        env.detach()
        if hasattr(self.actor, "release_graph"):
            self.actor.release_graph()
        gc.collect()            # no autograd graph of an earlier (default-stream) update may survive into the capture
        fz.step_base = th.full((1,), int(fz.global_step), dtype=th.int64, device=self.device)
        fz.global_step = 0
        fz._stepper = None                                    # re-bound with the step-base word
        self._holders_from_env()
        main = th.cuda.current_stream(self.device)
        side = th.cuda.Stream(self.device)
        side.wait_stream(main)
        with th.cuda.stream(side):                            # warm-up on a side stream, as graph capture wants it
            for _ in range(2):
                self._env_to_holders()
                self._one_update()
                self._holders_from_outputs()
        main.wait_stream(side)
        th.cuda.synchronize(self.device)
        steps0 = self.num_timesteps
        graph = th.cuda.CUDAGraph()
        self.actor.optimizer.zero_grad(set_to_none=True)
        self._env_to_holders()
        with th.cuda.graph(graph):
            self._g_loss = self._one_update()
            self._holders_from_outputs()
        self._g_steps = self.num_timesteps - steps0           # time steps one replay stands for
        self.num_timesteps = steps0                           # the capture pass itself computed nothing
        self._graph = graph

    def _replay(self):
        self._graph.replay()
        self.num_timesteps += self._g_steps
        self.env.envs.dynamics._n_steps += self.H
        return self._g_loss

    def learn(self, total_timesteps: int):
        assert self.H >= 1, "horizon must be at least 1"
        self.policy.train()
        start, last_dump, t0 = self.num_timesteps, self.num_timesteps, time.time()
        while self.num_timesteps - start < total_timesteps:
            if self.cuda_graph:
                if self._graph is None:
                    if not (self.env._fused is not None and self.env._fused.active):
                        actor_loss = self._one_update()       # the first eager update switches the env to the fused path
                        continue
                    self._capture()
                actor_loss = self._replay()
            else:
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
