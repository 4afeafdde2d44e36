"""Host logic of the comm-delay FIFO (visfly_b200.dynamics.Dynamics._as_device_action) on CPU tensors: which actions
are already private copies (conversions) and which are still the caller's tensor and must be cloned by the step launch
(`fifo_push` -> `fifo_copy`, the reference's `action.T.clone()`, envs/base/dynamics.py:323-328).  The launch itself is
covered on the GPU by tests/test_gpu_parity.py::test_comm_delay_fifo_owns_a_copy_of_every_action_like_the_reference."""
import types

import numpy as np
import pytest
import torch as th

from visfly_b200.dynamics import Dynamics


def _stage(action, n=2, device="cpu"):
    d = types.SimpleNamespace(device=th.device(device), num=n)
    return Dynamics._as_device_action(d, action)


def test_a_tensor_that_needs_no_conversion_is_still_the_callers():
    a = th.zeros(2, 4)
    out, owned = _stage(a)
    assert out is a and not owned                       # -> the kernel clones it


def test_conversions_produce_private_copies():
    out, owned = _stage(th.zeros(2, 4, dtype=th.float64))
    assert owned and out.dtype is th.float32
    out, owned = _stage(np.zeros((2, 4)))               # float64 numpy
    assert owned and out.dtype is th.float32
    base = th.zeros(2, 8)
    out, owned = _stage(base[:, ::2])                   # strided view -> contiguous copy
    assert owned and out.is_contiguous() and out.data_ptr() != base.data_ptr()


def test_float32_numpy_shares_memory_until_it_moves_to_the_device():
    """th.from_numpy aliases the array: on the engine's (CUDA) device the .to() makes the private copy; with no device
    change the alias is reported as not owned, so the launch clones it."""
    arr = np.zeros((2, 4), dtype=np.float32)
    out, owned = _stage(arr)
    assert not owned and out.data_ptr() == arr.ctypes.data


def test_shape_is_checked():
    with pytest.raises(ValueError, match="shape"):
        _stage(th.zeros(3, 4))
