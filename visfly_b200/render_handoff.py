"""Renderer hand-off: agent poses in Habitat-sim's frame without a host round trip per field.

The reference converts ``Dynamics.position / orientation / velocity`` with numpy matrix products on host copies
every step (``std_to_habitat``, utils/common.py:131-179, called from ``SceneManager.set_pose``,
utils/SceneManager.py:347-348) and then loops over the agents in Python.  Here one kernel reads the packed state and
writes ``[hab_pos, hab_ori]`` rows (+ velocities) either to device memory (for a GPU-resident renderer) or straight
into page-locked host memory (for Habitat-sim's scene-node API).  Habitat-sim itself is not part of this repository
(``visual=True`` raises); this module is the boundary a renderer binds to: ``HabitatPoseExporter`` hands the poses
over, ``SensorIngest`` takes the rendered frames back.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch as th

from . import _lib


class HabitatPoseExporter:
    """Reusable page-locked (or device) destination for the poses of one ``Dynamics`` object."""

    def __init__(self, dynamics, host: bool = True, with_velocity: bool = True):
        self.dynamics, self.host = dynamics, host
        n = dynamics.num
        kw = dict(pin_memory=True) if host else dict(device=dynamics.device)
        self.pose = th.empty((n, 7), dtype=th.float32, **kw)
        self.velocity = th.empty((n, 3), dtype=th.float32, **kw) if with_velocity else None

    def export(self, synchronize: bool = True):
        """-> ``(pose (n,7) = [hab_pos, hab_ori(w first)], hab_vel (n,3) | None)``; numpy views of the page-locked
        buffers when ``host=True`` (valid until the next ``export``), device tensors otherwise."""
        dyn = self.dynamics
        _lib.export_pose_habitat(dyn._cfg.params, dyn.packed_state.detach(), self.pose, self.velocity)
        if not self.host:
            return self.pose, self.velocity
        if synchronize:
            th.cuda.current_stream(dyn.device).synchronize()
        return self.pose.numpy(), None if self.velocity is None else self.velocity.numpy()


class SensorIngest:
    """Ingestion side of the hand-off (reference ``DroneEnvsBase.update_observation``, envs/base/droneEnv.py:296-331):
    the images of all agents, as one batched buffer per sensor, become the observation tensors on the device.

    A renderer binding writes its frames into ``buffer(uuid)`` — page-locked host memory (``host=True``, read by the
    kernel zero-copy over PCIe) or device memory (a GPU-resident renderer) — instead of returning one numpy array per
    agent to be stacked; ``ingest()`` then does the reference's post-processing in one streaming launch per sensor:
    ``depth``: ``(n,H,W) -> (n,1,H,W)`` with the no-return value 0 replaced by 20 (``:303-305``);
    ``color``: ``(n,H,W,4)`` RGBA uint8 ``-> (n,3,H,W)`` (``:306-308``); ``semantic``: ``(n,H,W) -> (n,1,H,W)``.
    Sensor kinds are told from the uuid exactly like the reference does (substring match, ``:301-312``)."""

    def __init__(self, num_agents: int, sensors: dict, device="cuda", host: bool = True, depth_background: float = 20.0):
        """``sensors``: ``{uuid: (H, W)}``."""
        self.n, self.device, self.background = int(num_agents), th.device(device), float(depth_background)
        self._src, self._dst = {}, {}
        kw = dict(pin_memory=True) if host else dict(device=self.device)
        for uuid, (h, w) in sensors.items():
            if "depth" in uuid or "semantic" in uuid:
                self._src[uuid] = th.zeros((self.n, h, w), dtype=th.float32, **kw)
                self._dst[uuid] = th.empty((self.n, 1, h, w), dtype=th.float32, device=self.device)
            elif "color" in uuid:
                self._src[uuid] = th.zeros((self.n, h, w, 4), dtype=th.uint8, **kw)
                self._dst[uuid] = th.empty((self.n, 3, h, w), dtype=th.uint8, device=self.device)
            else:
                raise KeyError("Can not find uuid of sensors")

    def buffer(self, uuid: str) -> th.Tensor:
        """Where the renderer writes the frames of sensor ``uuid`` (all agents, agent-major)."""
        return self._src[uuid]

    def ingest(self) -> dict:
        """-> ``{uuid: device tensor}`` in the reference's observation layout (valid until the next ``ingest``)."""
        for uuid, src in self._src.items():
            if "color" in uuid:
                _lib.ingest_color(src, self._dst[uuid])
            else:       # semantic ids are copied as they are: only depth has a background value
                _lib.ingest_depth(src, self._dst[uuid], self.background if "depth" in uuid else 0.0)
        return dict(self._dst)


def std_to_habitat(std_pos: Optional[th.Tensor] = None, std_ori: Optional[th.Tensor] = None, format="enu") \
        -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """Same contract as the reference function (utils/common.py:131-179) for tensors already on the host or the
    device: pure axis permutation / sign flips, so it is exact in any precision."""
    assert format in ["enu"]
    hab_ori = None
    if std_ori is not None:
        o = std_ori.detach()
        hab_ori = th.stack([o[..., 0], -o[..., 2], o[..., 3], -o[..., 1]], dim=-1).cpu().numpy()
    hab_pos = None
    if std_pos is not None:
        p = std_pos.detach()
        if p.dim() != 1 and p.shape[1] != 3:
            raise ValueError("std_pos shape error")
        hab_pos = th.stack([-p[..., 1], p[..., 2], -p[..., 0]], dim=-1).cpu().numpy()
    return hab_pos, hab_ori


def habitat_to_std(habitat_pos=None, habitat_ori=None, format="enu"):
    """Inverse transformation (reference utils/common.py:89-128); returns float32 / input-precision tensors."""
    assert format in ["enu"]
    std_pos = std_ori = None
    if habitat_pos is not None:
        p = np.atleast_2d(np.asarray(habitat_pos))
        std_pos = th.as_tensor(np.stack([-p[:, 2], -p[:, 0], p[:, 1]], axis=1), dtype=th.float32)
    if habitat_ori is not None:
        o = np.atleast_2d(np.asarray(habitat_ori))
        std_ori = th.from_numpy(np.stack([o[:, 0], -o[:, 3], -o[:, 1], o[:, 2]], axis=1))
    return std_pos, std_ori
