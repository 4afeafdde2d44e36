"""``DroneEnvsBase`` — owner of the dynamics engine, the initial-state generators and (``visual=False``) the
analytic bounding-box collision model.  Surface of reference envs/base/droneEnv.py:18-525.

What changed underneath: ``Dynamics`` is the fused CUDA engine; state generation is one vectorised draw for all
requested agents instead of a per-agent Python loop (reference :243-249); collision bookkeeping always covers the
whole batch with a handful of tensor ops and never synchronises with the host; a mask-based reset
(``reset_agents_where``) lets the wrapper auto-reset finished agents without ``where(done)`` round trips.
Rendering (Habitat-Sim ``SceneManager``) is out of scope: ``visual=True`` raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch as th

from ...dynamics import Dynamics
from ...randomization import load_generator
from ...type import Normal, Uniform

IS_BBOX_COLLISION = True


def _watched(name):
    from .fused import watched
    return watched(name)


class DroneEnvsBase:
    # settings the one-kernel env step bakes into its spec: assigning one is noticed on the next step (fused.watched)
    _gen = 0
    stateGenerator = _watched("stateGenerator")
    _reset_table = _watched("_reset_table")
    uav_radius = _watched("uav_radius")
    _bboxes = _watched("_bboxes")

    def __init__(
            self,
            num_agent_per_scene: int = 1,
            num_scene: int = 1,
            seed: int = 42,
            visual: bool = False,
            random_kwargs: Optional[Dict] = None,
            dynamics_kwargs: Optional[Dict] = None,
            scene_kwargs: Optional[Dict] = None,
            sensor_kwargs: Optional[Dict] = None,
            uav_radius: float = 0.1,
            sensitive_radius: float = 10.,
            multi_drone: bool = False,
            device="cuda",
            shard=None,
    ):
        if visual:
            raise NotImplementedError(
                "visual=True needs Habitat-Sim (reference utils/SceneManager.py); rendering is left untouched and "
                "out of scope of this engine — construct the env with visual=False")
        random_kwargs = dict(random_kwargs or {})
        self.device = th.device(device)
        self.seed = seed
        #: (agent_offset, total_agents) when this env is one contiguous shard of a larger batch spread over several
        #: GPUs: initial placements and in-kernel restarts are then the rows the whole batch would draw
        self.shard = None if shard is None else (int(shard[0]), int(shard[1]))
        self.visual = visual
        self.uav_radius = uav_radius
        self.is_multi_drone = multi_drone
        if multi_drone and num_agent_per_scene == 1:
            raise ValueError("Num of agents should not be 1 in multi drone env.")
        self.num_scene, self.num_agent_per_scene = num_scene, num_agent_per_scene

        self.noise_settings = dict(random_kwargs.get("noise_kwargs", {}))
        self.dynamics = Dynamics(num=num_agent_per_scene * num_scene, seed=seed, device=device,
                                 **(dynamics_kwargs or {}))
        self.device = self.dynamics.device
        self._create_noise_model()
        self.sceneManager = None
        self.stateGenerator = self._create_randomizer(random_kwargs)
        self._scene_iter = random_kwargs.get("scene_iter", False)
        self._create_bbox()
        self._sensor_list = []
        self._sensor_obs = {}
        n = self.dynamics.num
        self._once_collided = th.zeros(n, dtype=th.bool, device=self.device)
        self._is_collision = th.zeros(n, dtype=th.bool, device=self.device)
        self._is_out_bounds = th.zeros(n, dtype=th.bool, device=self.device)
        self._collision_point = self._collision_vector = self._collision_dis = None
        self._collision_stale = False      # set by the fused env step: collision views are recomputed on demand
        self._reset_table = None           # optional (N,13) [p q v w] rows agents restart from (set_reset_table)
        self._fused = None                 # FusedEnvStep while the one-kernel env step owns the per-agent status
        self._eval = False

    # -- construction helpers ---------------------------------------------------------------------
    def _create_noise_model(self):
        """IMU noise on the 13-d state (reference :99-125).  All-zero noise (the default) draws nothing."""
        cfg = self.noise_settings.get("IMU", None)
        dim = self.dynamics.state.shape[1]
        if cfg is None:
            model = Uniform(mean=th.zeros(dim), half=th.zeros(dim))
        elif cfg["model"] == "UniformNoiseModel":
            model = Uniform(**cfg.get("kwargs", {}))
        elif cfg["model"] == "GaussianNoiseModel":
            model = Normal(**cfg.get("kwargs", {}))
        else:
            raise ValueError("IMU Noise model does not exist.")
        self.noise_settings["IMU"] = model.to(self.device)
        self._imu_noise_free = isinstance(model, Uniform) and model.is_zero

    def _create_bbox(self):
        self._bboxes = [th.tensor([[-30., -30., 0.], [30., 30., 8.]], device=self.device)]     # reference :129
        self._flatten_bboxes = [b.flatten() for b in self._bboxes]

    def _create_randomizer(self, random_kwargs: Dict):
        cfg = random_kwargs.get("state_generator", {})
        kwargs_list = cfg.get("kwargs", [{}])
        gen = load_generator(cls=cfg.get("class", "Uniform"), device=self.device, is_collision_func=None,
                             scene_id=0, kwargs=kwargs_list[0])
        gen.to(self.device)
        return gen

    # -- state generation / reset ---------------------------------------------------------------------
    def set_reset_table(self, pos, quat, vel=None, ang_vel=None):
        """Deterministic (re)starts: agent i always restarts from row i.  Replaces the random state generator
        for both the fused and the generic path (``None`` as ``pos`` switches back to the generator)."""
        if pos is None:
            self._reset_table = None
            return
        n, dev = self.dynamics.num, self.device
        f = lambda x, k: th.zeros((n, k), device=dev) if x is None else \
            th.as_tensor(x, dtype=th.float32, device=dev).reshape(n, k)
        self._reset_table = th.cat([f(pos, 3), f(quat, 4), f(vel, 3), f(ang_vel, 3)], 1).contiguous()

    def _generate_state(self, indices=None, num: Optional[int] = None):
        if self._reset_table is not None:
            t = self._reset_table if indices is None else self._reset_table[th.as_tensor(indices, device=self.device)]
            return t[:, 0:3], t[:, 3:7], t[:, 7:10], t[:, 10:13]
        n = self.dynamics.num if indices is None else len(indices)
        if self.shard is not None and indices is None and num is None:
            # draw for the whole batch and keep this shard's rows: same seed => same placement as one big env
            lo, total = self.shard
            return tuple(x[lo:lo + n] for x in self.stateGenerator.safe_generate(num=total))
        return self.stateGenerator.safe_generate(num=n if num is None else num)

    def reset(self, state=None):
        self.reset_agents(indices=None, state=state)
        return self.state, self.sensor_obs

    @staticmethod
    def _split_state(state, device):
        if isinstance(state, th.Tensor):
            s = state.to(device).detach()
            return s[:, :3], s[:, 3:7], s[:, 7:10], s[:, 10:13], s[:, 13:17], s[:, 17:21], s[:, 21]
        if len(state) == 4:
            return (*state, None, None, None)
        if len(state) == 6:
            return (*state, None)
        raise ValueError("State should be a tuple of 4 or 6 elements.")

    def reset_agents(self, indices=None, state=None, pos_reset_by_state=False):
        """Index-based (re)initialisation, reference :260-288."""
        if indices is not None and not hasattr(indices, "__iter__"):
            indices = [indices]
        if indices is not None:
            indices = th.as_tensor(indices, device=self.device, dtype=th.int64).reshape(-1)
        motor, thrust, t = None, None, None
        if state is not None:
            pos, ori, vel, ori_vel, motor, thrust, t = self._split_state(state, self.device)
            if not pos_reset_by_state:
                pos, _, _, _ = self._generate_state(indices)
        else:
            pos, ori, vel, ori_vel = self._generate_state(indices)
        self.dynamics.reset(pos=pos, ori=ori, vel=vel, ori_vel=ori_vel, motor_omega=motor, thrusts=thrust, t=t,
                            indices=indices)
        self.update_observation()
        self.update_collision()
        if indices is None:
            self._once_collided = th.zeros_like(self._once_collided)
        else:
            self._once_collided = self._once_collided.index_fill(0, indices, False)

    def reset_agents_where(self, mask: th.Tensor):
        """Re-initialise the agents selected by a boolean mask without leaving the device: fresh states are
        drawn for the whole batch and blended in (same distribution as reset_agents on ``where(mask)``)."""
        pos, ori, vel, ori_vel = self._generate_state(None)
        self.dynamics.reset_where(mask, pos=pos, ori=ori, vel=vel, ori_vel=ori_vel)
        self.update_observation()
        self.update_collision()
        self._once_collided = self._once_collided & ~mask

    # -- per-step bookkeeping -------------------------------------------------------------------------------
    def _generate_noise_obs(self, sensor):
        if sensor != "IMU":
            return None
        if self._imu_noise_free:
            return self.state
        noisy = self.state + self.noise_settings["IMU"].generate(self.dynamics.num).to(self.device)
        if self.dynamics.is_quat_output:      # renormalise the perturbed quaternion (reference :118-124)
            noisy = th.cat([noisy[:, :3], th.nn.functional.normalize(noisy[:, 3:7], p=2, dim=1), noisy[:, 7:]], dim=1)
        return noisy

    def update_observation(self, indices=None):
        self._sensor_obs["IMU"] = self._generate_noise_obs("IMU")

    def update_collision(self, indices=None):
        """Closest point on the scene bounding box (reference :345-367); whole batch, no host sync."""
        pos = self.dynamics.position
        p = pos.detach()
        lo, hi = self._bboxes[0][0], self._bboxes[0][1]
        gap = th.cat([p - lo, hi - p], dim=1)                      # (N,6) distance to each face
        face = gap.argmin(dim=1)
        cp = p.scatter(1, (face % 3).unsqueeze(1), self._flatten_bboxes[0][face].unsqueeze(1))
        self._collision_point = cp
        self._is_out_bounds = (p < lo).any(dim=1) | (p > hi).any(dim=1)
        self._collision_vector = cp - pos
        self._collision_dis = (self._collision_vector - 0).norm(dim=1)
        self._is_collision = self._collision_dis < self.uav_radius
        self._once_collided = self._once_collided | self._is_collision
        self._collision_stale = False

    def _collision_view(self, name):
        if self._collision_stale:           # after a fused step: recompute from the current position, keep the flags
            keep = self._once_collided
            self.update_collision()
            self._once_collided = keep
        return getattr(self, name)

    def step(self, action):
        self.dynamics.step(action)
        self.update_observation()
        self.update_collision()

    # -- misc ----------------------------------------------------------------------------------------------------
    def set_seed(self, seed=42):
        self.dynamics.set_seed(self.seed if seed is None else seed)

    def stack(self):
        self._stack_cache = tuple(x.clone().detach() for x in
                                  (self.position, self.orientation, self.velocity, self.angular_velocity))

    def recover(self):
        self.reset_agents(state=self._stack_cache, pos_reset_by_state=True)

    def detach(self):
        self.dynamics.detach()
        if self._collision_vector is not None:
            self._collision_vector = self._collision_vector.detach()
            self._collision_dis = self._collision_dis.detach()

    def close(self):
        self.dynamics.close()

    def eval(self):
        self._eval = True

    def render(self, **kwargs):
        return None

    # -- views -----------------------------------------------------------------------------------------------------
    state = property(lambda s: s.dynamics.state)
    sensor_obs = property(lambda s: s._sensor_obs)
    is_collision = property(lambda s: s._collision_view("_is_collision"))
    is_out_bounds = property(lambda s: s._collision_view("_is_out_bounds"))
    direction = property(lambda s: s.dynamics.direction)
    position = property(lambda s: s.dynamics.position)
    orientation = property(lambda s: s.dynamics.orientation)
    velocity = property(lambda s: s.dynamics.velocity)
    angular_velocity = property(lambda s: s.dynamics.angular_velocity)
    t = property(lambda s: s.dynamics.t)
    thrusts = property(lambda s: s.dynamics.thrusts)
    full_state = property(lambda s: s.dynamics.full_state)
    extend_state = property(lambda s: s.dynamics.extend_state)
    acceleration = property(lambda s: s.dynamics.acceleration)
    angular_acceleration = property(lambda s: s.dynamics.angular_acceleration)
    collision_point = property(lambda s: s._collision_view("_collision_point"))
    collision_vector = property(lambda s: s._collision_view("_collision_vector"))
    collision_dis = property(lambda s: s._collision_view("_collision_dis"))

    @property
    def once_collided(self):
        if self._fused is not None and self._fused.active:
            return (self._fused.eb & 2).bool()
        return self._once_collided

    @property
    def dynamic_object_position(self):
        return [[None] for _ in range(self.dynamics.num)]

    dynamic_object_velocity = dynamic_object_position
    dynamic_object_acceleration = dynamic_object_position
