"""Quaternion helper kept for surface compatibility with the reference's ``utils/maths.py``.

The engine itself never uses this class — orientation lives in plane 1 of the packed state and all
quaternion arithmetic of the control step happens inside the CUDA kernel.  Task code written against VisFly
does touch ``dynamics._orientation`` (reference envs/base/droneEnv.py:376 calls ``.toTensor()``) and uses
``Quaternion`` for a few geometric helpers, so a batched implementation with the same method names is
provided on top of one ``(4,N)`` tensor (component-major like the reference's four ``(N,)`` tensors).
"""
from __future__ import annotations

import torch as th


class Quaternion:
    def __init__(self, w=None, x=None, y=None, z=None, num=1, device=th.device("cpu")):
        if w is None:
            q = th.zeros((4, num), device=device)
            q[0] = 1
        elif isinstance(w, (int, float)):
            q = th.tensor([[float(w)], [float(x)], [float(y)], [float(z)]], device=device).repeat(1, num)
        elif isinstance(w, th.Tensor):
            q = th.stack([th.atleast_1d(c) for c in (w, x, y, z)])
        else:
            raise ValueError("unsupported type")
        self._q = q

    @classmethod
    def from_tensor(cls, q: th.Tensor) -> "Quaternion":
        """``q`` is (4,N), rows w,x,y,z (no copy)."""
        obj = cls.__new__(cls)
        obj._q = q
        return obj

    # components -------------------------------------------------------------------------------
    w = property(lambda s: s._q[0])
    x = property(lambda s: s._q[1])
    y = property(lambda s: s._q[2])
    z = property(lambda s: s._q[3])
    real = property(lambda s: s._q[0])
    imag = property(lambda s: s._q[1:])
    shape = property(lambda s: (4, s._q.shape[1]))

    def toTensor(self):
        return self._q

    def to(self, device):
        self._q = self._q.to(device)
        return self

    def clone(self):
        return Quaternion.from_tensor(self._q.clone())

    def detach(self):
        return Quaternion.from_tensor(self._q.detach())

    def __len__(self):
        return self._q.shape[1]

    def __getitem__(self, idx):
        return Quaternion.from_tensor(self._q[:, idx].reshape(4, -1))

    def __repr__(self):
        return f"Quaternion(wxyz={self._q.T})"

    # algebra ------------------------------------------------------------------------------------
    def conjugate(self):
        return Quaternion.from_tensor(self._q * th.tensor([[1.0], [-1.0], [-1.0], [-1.0]], device=self._q.device))

    def norm(self):
        return self._q.norm(dim=0)

    def normalize(self):
        return Quaternion.from_tensor(self._q / self.norm())

    def inverse(self):
        return Quaternion.from_tensor(self.conjugate()._q / self.norm())

    def __neg__(self):
        return Quaternion.from_tensor(-self._q)

    def __add__(self, other):
        return Quaternion.from_tensor(self._q + (other._q if isinstance(other, Quaternion) else other))

    def __sub__(self, other):
        return Quaternion.from_tensor(self._q - other._q)

    def __truediv__(self, other):
        return Quaternion.from_tensor(self._q / other)

    def __mul__(self, other):
        if isinstance(other, Quaternion):
            a, b = self._q, other._q
            return Quaternion.from_tensor(th.stack([
                a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]))
        return Quaternion.from_tensor(self._q * other)

    def _sandwich(self, v: th.Tensor, sign: float):
        """(w^2 - r.r) v + 2 (r.v) r + 2 sign w (r x v); equals the double Hamilton product for any q."""
        w, r = self._q[0], self._q[1:]
        rv = (r * v).sum(0)
        return (w * w - (r * r).sum(0)) * v + 2 * rv * r + 2 * sign * w * th.linalg.cross(r, v.expand_as(r), dim=0)

    def rotate(self, other):
        """body -> world for a (3,N) vector; quaternion product for a Quaternion."""
        return self * other if isinstance(other, Quaternion) else self._sandwich(other, 1.0)

    def inv_rotate(self, other):
        """world -> body."""
        return self.conjugate() * other if isinstance(other, Quaternion) else self._sandwich(other, -1.0)

    transform = inv_rotate
    inv_transform = rotate

    # frames ---------------------------------------------------------------------------------------
    @property
    def R(self):
        w, x, y, z = self._q
        return th.stack([
            th.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
            th.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]),
            th.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)])])

    @property
    def x_axis(self):
        w, x, y, z = self._q
        return th.stack([1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w)])

    @property
    def xz_axis(self):
        w, x, y, z = self._q
        return th.stack([
            th.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
            th.stack([2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)])])

    def toEuler(self, order="zyx"):
        w, x, y, z = self._q
        if order == "zyx":
            roll = th.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
            pitch = th.asin(2 * (w * y - z * x))
            yaw = th.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        elif order == "xyz":
            roll = th.atan2(2 * (w * y - x * z), 1 - 2 * (x * x + y * y))
            pitch = th.asin(2 * (w * z - y * x))
            yaw = th.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + z * z))
        else:
            raise ValueError("order should be 'zyx' or 'xyz'")
        return th.stack([roll, pitch, yaw])

    @staticmethod
    def from_euler(roll, pitch, yaw, order="zyx"):
        roll, pitch, yaw = (th.as_tensor(a, dtype=th.float32) * 0.5 for a in (roll, pitch, yaw))
        cr, sr, cp, sp, cy, sy = roll.cos(), roll.sin(), pitch.cos(), pitch.sin(), yaw.cos(), yaw.sin()
        s = 1.0 if order == "zyx" else -1.0
        return Quaternion(cr * cp * cy + s * sr * sp * sy,
                          sr * cp * cy - s * cr * sp * sy,
                          cr * sp * cy + s * sr * cp * sy,
                          cr * cp * sy - s * sr * sp * cy)


def cross(a: th.Tensor, b: th.Tensor):
    """Cross product of two (3,N) tensors (reference utils/maths.py:392-394)."""
    return th.linalg.cross(a, b, dim=0)
