"""TEST INFRASTRUCTURE — CPU/PyTorch restatement of the reference's env wrapper tail (``visual=False``).

Restates, on top of ``OracleDynamics``, what the reference executes around the dynamics step:
  * ``DroneEnvsBase``     envs/base/droneEnv.py — per-agent state generation loop (:237-251), analytic bounding
                          box collision (:345-369), ``reset_agents`` (:260-288), ``step`` (:373-379);
  * ``DroneGymEnvsBase``  envs/base/droneGymEnv.py — ``step`` (:141-218) with its per-agent info loop (:197-201,
                          :238-275), ``reset`` (:302-327), ``reset_agent_by_id`` (:339-349), ``_reset_attr`` (:357-418);
  * task envs             HoverEnv.py:59-94, NavigationEnv.py:58-99, RacingEnv.py:87-215, :250-267
                          (with the ``predicted_obs`` / ``latent`` signature repairs of SURVEY.md C4).
The per-agent Python loops are kept on purpose: this is also the "reference arm" of ``bench.py`` (kind = "port"),
and those loops are what the reference spends most of its env-step time in at large N.

Parity status: pinned against the executed reference — ``tests/golden/make_env_golden.py`` runs the real
reference envs (third-party imports stubbed, ``tests/_reference.py``) and ``tests/test_env_oracle.py`` checks this
port against those fixtures and, where the reference tree is mounted, bit-exactly against the live reference.
Only tests/, smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch as th

from .torch_oracle import OracleDynamics


def _uniform_state(cfg: Dict, num: int):
    """reference utils/randomization.py:153-169 (one Uniform generator) + :98 euler->quaternion."""
    z3 = {"mean": [0., 0., 0.], "half": [0., 0., 0.]}
    out = []
    for key in ("position", "orientation", "velocity", "angular_velocity"):
        f = cfg.get(key, z3)
        out.append((2 * th.rand(num, 3) - 1) * th.tensor(f["half"]).unsqueeze(0) + th.tensor(f["mean"]).unsqueeze(0))
    pos, eul, vel, rate = out
    h = eul * 0.5
    cr, sr, cp, sp, cy, sy = h[:, 0].cos(), h[:, 0].sin(), h[:, 1].cos(), h[:, 1].sin(), h[:, 2].cos(), h[:, 2].sin()
    quat = th.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)
    return pos, quat, vel, rate


class OracleEnv:
    """Wrapper + task in one class; ``task`` in {"hover", "navigation", "racing2"}."""

    RACING_BOXES = [[2., 2., 1], [6., 2., 1.5], [6., -2., 1.5], [2., 0., 1]]

    def __init__(self, task: str, num_agents: int, dynamics_kwargs: Dict, max_episode_steps: int = 256,
                 requires_grad: bool = False, tensor_output: bool = True, random_kwargs: Optional[Dict] = None,
                 target=None, is_collision_reset: bool = True, dtype=th.float32,
                 generate_state: Optional[Callable] = None, faithful_rng: bool = True):
        assert task in ("hover", "navigation", "racing2")
        self.task, self.n, self.dtype = task, num_agents, dtype
        kw = dict(dynamics_kwargs)
        kw["wind"] = kw.pop("wind_settings", (0, 0, 0))
        # faithful_rng: consume the global RNG exactly where the reference does (per-step IMU noise draw even
        # when the noise is zero, C11; random t on partial reset, C9) so that seeded runs match it draw for draw
        self.faithful_rng = faithful_rng
        self.dyn = OracleDynamics(num=num_agents, dtype=dtype, random_reset_time=faithful_rng, **kw)
        self.max_episode_steps = max_episode_steps
        self.requires_grad, self.tensor_output = requires_grad, tensor_output
        self.is_collision_reset = is_collision_reset
        self.uav_radius = 0.1
        self.bbox = th.tensor([[-30., -30., 0.], [30., 30., 8.]], dtype=dtype)
        if task == "hover":
            default = {"state_generator": {"class": "Uniform", "kwargs": [
                {"position": {"mean": [1., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}}
            self.random_kwargs = default if random_kwargs is None else random_kwargs
            self.target = th.ones((self.n, 1), dtype=dtype) @ th.as_tensor(
                [1, 0., 1.5] if target is None else target, dtype=dtype).reshape(1, -1)
        elif task == "navigation":
            self.random_kwargs = random_kwargs or {}
            self.target = th.ones((self.n, 1), dtype=dtype) @ th.as_tensor(
                [9, 0., 1] if target is None else target, dtype=dtype).reshape(1, -1)
            self.success_radius = 0.5
        else:
            self.random_kwargs = None
            self.targets = th.as_tensor([[4, 4, 1.], [8, 0, 2.], [5, -4, 1.], [1, -1, 1.]], dtype=dtype)
            self.next_target_i = th.zeros((self.n,), dtype=th.int)
            self.past_targets_num = th.zeros((self.n,), dtype=th.int)
            self.is_pass_next = th.zeros((self.n,), dtype=th.bool)
            self.success_radius = 0.3
        self._generate_override = generate_state
        self.once_collided = th.zeros(self.n, dtype=th.bool)
        self.step_count = th.zeros((self.n,), dtype=th.int32)
        self.reward = th.zeros((self.n,), dtype=dtype)
        self.rewards = th.zeros((self.n,), dtype=dtype)
        self.success = th.zeros(self.n, dtype=th.bool)
        self.failure = th.zeros(self.n, dtype=th.bool)
        self.episode_done = th.zeros(self.n, dtype=th.bool)
        self.done = th.zeros(self.n, dtype=th.bool)
        self.info: List[dict] = [{"TimeLimit.truncated": False} for _ in range(self.n)]
        self.observations: Dict[str, th.Tensor] = {}
        self._is_initial = False

    # -- DroneEnvsBase ---------------------------------------------------------------------------------------
    def _generate_one(self):
        if self.task == "racing2":        # Union of four Uniform boxes: all four are drawn, one is picked (:278-296)
            draws = [_uniform_state({"position": {"mean": c, "half": [.2, .2, 0.2]}}, 1) for c in self.RACING_BOXES]
            pick = int(th.randint(0, 4, (1,)))
            return draws[pick]
        cfg = self.random_kwargs.get("state_generator", {}).get("kwargs", [{}])[0]
        return _uniform_state(cfg, 1)

    def _generate_state(self, indices=None):
        """One generator call per agent, like the reference (droneEnv.py:243-249)."""
        if self._generate_override is not None:
            return self._generate_override(indices)
        indices = np.arange(self.n) if indices is None else indices
        m = len(indices)
        pos, quat, vel, rate = th.empty((m, 3)), th.empty((m, 4)), th.empty((m, 3)), th.empty((m, 3))
        for k in range(m):
            p, q, v, w = self._generate_one()
            pos[k], quat[k], vel[k], rate[k] = p[0], q[0], v[0], w[0]
        return pos, quat, vel, rate

    def _imu_draw(self):
        """droneEnv.py:114-116, :333 — the (zero) IMU noise is drawn for every agent at every observation update."""
        if self.faithful_rng:
            th.rand(self.n, 13)

    def _update_collision(self):                                                   # droneEnv.py:345-369
        p = self.dyn.position.clone().detach()
        value, index = th.hstack([p - self.bbox[0], self.bbox[1] - p]).min(dim=1)
        cp = p.clone()
        cp[th.arange(self.n), index % 3] = self.bbox.flatten()[index]
        self.collision_point = cp
        self.is_out_bounds = (self.dyn.position < self.bbox[0]).any(dim=1) | (self.dyn.position > self.bbox[1]).any(dim=1)
        self.collision_vector = cp - self.dyn.position
        self.collision_dis = (self.collision_vector - 0).norm(dim=1)
        self.is_collision = self.collision_dis < self.uav_radius
        self.once_collided = self.once_collided | self.is_collision

    def _reset_agents(self, indices=None):                                         # droneEnv.py:260-288
        pos, quat, vel, rate = self._generate_state(indices)
        cv = lambda x: x.to(self.dtype)
        self.dyn.reset(pos=cv(pos), ori=cv(quat), vel=cv(vel), ori_vel=cv(rate), indices=indices)
        self._imu_draw()
        self._update_collision()
        if indices is None:
            self.once_collided = th.zeros(self.n, dtype=th.bool)
        else:
            self.once_collided = self.once_collided.clone()
            self.once_collided[indices] = False

    # -- task ------------------------------------------------------------------------------------------------
    def _observation(self):
        d = self.dyn
        if self.task == "hover":
            return {"state": d.state}
        if self.task == "navigation":
            return {"state": d.state, "target": self.target}
        nxt = th.stack([self.next_target_i + i for i in range(2)]).T % len(self.targets)       # RacingEnv.py:254-267
        rel = (self.targets[nxt] - d.position.unsqueeze(1)).reshape(self.n, -1)
        state = th.hstack([rel / 10, d.orientation, d.velocity / 10, d.angular_velocity / 10])
        return {"state": state, "gate": self.next_target_i.unsqueeze(1).clone().detach()}

    def _success(self):
        if self.task == "hover":
            return th.full((self.n,), False)
        if self.task == "navigation":
            return (self.dyn.position - self.target).norm(dim=1) <= self.success_radius
        self.is_pass_next = (self.dyn.position - self.targets[self.next_target_i]).norm(dim=1) <= self.success_radius
        self.next_target_i = (self.next_target_i + self.is_pass_next) % len(self.targets)       # RacingEnv.py:142-148
        self.past_targets_num = self.past_targets_num + self.is_pass_next
        return th.zeros((self.n,), dtype=th.bool)

    def _reward(self):
        d = self.dyn
        one = th.tensor([1, 0, 0, 0])
        if self.task in ("hover", "racing2"):                                       # HoverEnv.py:83-94, RacingEnv.py:203-215
            tgt = self.target if self.task == "hover" else self.targets[self.next_target_i]
            r = (0.1 + (d.position - tgt).norm(dim=1) * (-0.1 * 1 / 9)
                 + (d.orientation - one).norm(dim=1) * -0.00001
                 + (d.velocity - 0).norm(dim=1) * -0.002
                 + (d.angular_velocity - 0).norm(dim=1) * -0.002)
            return r if self.task == "hover" else r + self.is_pass_next * 20
        base_r, thrd = 0.1, th.pi / 18                                              # NavigationEnv.py:85-99
        return base_r * 0 + \
            ((d.velocity * (self.target - d.position)).sum(dim=1) / (1e-6 + (self.target - d.position).norm(dim=1))).clamp_max(10) * 0.01 + \
            (((d.direction * d.velocity).sum(dim=1) / (1e-6 + d.velocity.norm(dim=1)) / 1).clamp(-1., 1.).acos().clamp_min(thrd) - thrd) * -0.01 + \
            (d.orientation - one).norm(dim=1) * -0.00001 + \
            (d.velocity - 0).norm(dim=1) * -0.002 + \
            (d.angular_velocity - 0).norm(dim=1) * -0.002 + \
            1 / (self.collision_dis + 0.2) * -0.01 + \
            (1 - self.collision_dis).relu() * ((self.collision_vector * (d.velocity - 0)).sum(dim=1) / (1e-6 + self.collision_dis)).relu() * -0.005 + \
            self.success * (self.max_episode_steps - self.step_count) * base_r * (0.2 + 0.8 / (1 + 1 * d.velocity.norm(dim=1)))

    def _choose_target(self, indices=None):                                          # RacingEnv.py:173-185
        indices = th.arange(self.n) if indices is None else indices
        rela = self.dyn.position - th.as_tensor([4, 0, 1])
        for index in indices:
            if rela[index][0] < 0:
                self.next_target_i[index] = 0 if rela[index][1] > 0 else 3
            else:
                self.next_target_i[index] = 1 if rela[index][0] > 0 else 2

    # -- DroneGymEnvsBase ------------------------------------------------------------------------------------------
    def reset(self):                                                                 # droneGymEnv.py:302-327
        self._is_initial = True
        self._reset_agents(None)
        self.observations = self._observation()
        self._reset_attr(None)
        self.observations = self._observation()
        if self.task == "racing2":
            # RacingEnv.py:165-170: the first gate is chosen AFTER the base class built the observation, so the
            # observation returned by reset() still refers to the previous gate indices (reference quirk, kept)
            self.next_target_i = th.zeros((self.n,), dtype=th.int)
            self._choose_target()
        return self.observations

    def _reset_attr(self, indices):                                                  # droneGymEnv.py:357-418
        with th.no_grad():
            if indices is None:
                self.reward = th.zeros((self.n,), dtype=self.dtype)
                self.rewards = th.zeros((self.n,), dtype=self.dtype)
                self.done = th.zeros(self.n, dtype=th.bool)
                self.episode_done = th.zeros(self.n, dtype=th.bool)
                self.step_count = th.zeros((self.n,), dtype=th.int32)
            else:
                self.reward[indices] = 0
                self.rewards[indices] = 0
                self.done[indices] = False
                self.episode_done[indices] = False
                self.step_count[indices] = 0
        for i in (range(self.n) if indices is None else indices):
            self.info[int(i)] = {"TimeLimit.truncated": False, "episode_done": False}

    def _collect_info(self, i):                                                       # droneGymEnv.py:238-275
        length = self.step_count[i].cpu().clone().detach().numpy()
        info = {
            "episode_done": self.episode_done[i].item(),
            "is_success": bool(self.success[i]),
            "episode": {"r": self.rewards[i].cpu().clone().detach().numpy(), "l": length,
                        "t": (self.step_count[i] * self.dyn.ctrl_dt).cpu().clone().detach().numpy()},
            "terminal_observation": {k: (v[i].detach() if self.requires_grad else v[i])
                                     for k, v in self.observations.items()},
            "TimeLimit.truncated": bool(self.step_count[i] >= self.max_episode_steps),
        }
        info["episode"]["extra"] = {"collision": self.once_collided[i].clone().detach().cpu().numpy()}
        if self.task == "racing2":
            info["episode"]["extra"]["past_gate"] = self.past_targets_num[i].item()
        return info

    def step(self, action, is_test=False):                                            # droneGymEnv.py:141-218
        assert self._is_initial
        action = action if isinstance(action, th.Tensor) else th.as_tensor(action)
        assert action.max() <= 1 and action.min() >= -1
        self.dyn.step(action)
        self._imu_draw()
        self._update_collision()
        self.observations = self._observation()
        self.step_count += 1
        self.success = self._success()
        self.failure = th.full((self.n,), False)
        self.reward = self._reward()
        self.rewards = self.rewards + self.reward
        self.episode_done = self.episode_done | self.success | self.failure | self.is_out_bounds
        if self.is_collision_reset:
            self.episode_done = self.episode_done | self.is_collision
        self.done = self.episode_done | (self.step_count >= self.max_episode_steps)
        for i in range(self.n):                       # the reference's per-agent loop, kept (droneGymEnv.py:197-201)
            if self.done[i]:
                self.info[i] = self._collect_info(i)
        done, reward, info = self.done.clone(), self.reward.clone(), self.info.copy()
        if self.done.any() and not is_test:                                            # examine(), :420-423
            idx = th.where(self.done)[0]
            if self.task == "racing2":                                                 # RacingEnv.py:150-163
                self._choose_target(idx)
                self.past_targets_num[idx] = 0
                self.is_pass_next[idx] = False
            self._reset_agents(idx)
            self.observations = self._observation()
            self._reset_attr(idx)
        if self.requires_grad:
            return self.observations, reward, done, info
        if self.tensor_output:
            return {k: v.detach() for k, v in self.observations.items()}, reward.detach(), done, info
        return ({k: v.detach().cpu().numpy() for k, v in self.observations.items()}, reward.cpu().numpy(),
                done.cpu().numpy().astype(np.int32), info)

    def detach(self):                                                                  # droneGymEnv.py:286-300
        self.dyn.detach()
        self.rewards = self.rewards.clone().detach()
        self.reward = self.reward.clone().detach()
        self.step_count = self.step_count.clone().detach()
        self.done = self.done.clone().detach()
