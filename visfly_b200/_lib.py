"""ctypes binding of the C-ABI in include/visfly_b200.h (``libvisfly_b200.so``, built in-tree).

There is exactly one compute path: the sm_100a kernels.  If the shared library is missing, or CUDA is not
available, every entry point raises — there is no CPU fallback and no PyTorch re-implementation behind it.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch as th

from .params import FIFO_MAX_ROWS, VfEnvMirror, VfEnvSpec, VfFifoRows, VfParams, VfPeerScatter

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvisfly_b200.so")
ABI_VERSION = 8

INTEGRATOR_ID = {"euler": 0, "rk4": 1}
FLAG_CTRL_DELAY = 1
MAX_SUBSTEPS_BWD = 64

_P = ctypes.POINTER
_vp = ctypes.c_void_p
_i = ctypes.c_int
_u = ctypes.c_uint

# name -> (restype, argtypes); mirrors include/visfly_b200.h one to one
SIGNATURES = {
    "vf_abi_version": (_i, []),
    "vf_last_error": (ctypes.c_char_p, []),
    "vf_params_size": (_i, []),
    "vf_device_sm_count": (_i, []),
    "vf_step_fwd": (_i, [_P(VfParams), _i, _i, _i, _i, _u, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_step_fwd_ring": (_i, [_P(VfParams), _i, _i, _i, _i, _u, _vp, _P(VfFifoRows), _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_step_bwd": (_i, [_P(VfParams), _i, _i, _i, _i, _u, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_step_fwd_host": (_i, [_P(VfParams), _i, _i, _i, _i, _u, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_pack_state": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_unpack_state": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_export_pose_habitat": (_i, [_P(VfParams), _i, _vp, _vp, _vp, _vp]),
    "vf_env_spec_size": (_i, []),
    "vf_wait_flag": (_i, [_vp, _u, ctypes.c_longlong]),
    "vf_ingest_depth": (_i, [ctypes.c_longlong, _i, _i, _vp, _vp, ctypes.c_float, _vp]),
    "vf_ingest_color": (_i, [ctypes.c_longlong, _i, _i, _vp, _vp, _vp]),
    "vf_sensor_last_error": (ctypes.c_char_p, []),
    "vf_env_finish": (_i, [_P(VfParams), _P(VfEnvSpec), _i, _u, ctypes.c_ulonglong, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp, _vp, _vp, _vp, _vp, _vp, _P(VfFifoRows), _vp]),
    "vf_policy_packed_floats": (_i, [_i]),
    "vf_policy_pack": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vf_policy_fwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp]),
    "vf_policy_bwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp,
                           _vp]),
    "vf_policy_partial_floats": (_i, [_i, _i, _i]),
    "vf_policy_last_error": (ctypes.c_char_p, []),
    "vf_env_step_fwd": (_i, [_P(VfParams), _P(VfEnvSpec), _i, _i, _i, _i, _u, _u, ctypes.c_ulonglong, _vp,
                             _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _P(VfEnvMirror), _P(VfPeerScatter), _vp]),
    "vf_env_step_bwd": (_i, [_P(VfParams), _P(VfEnvSpec), _i, _i, _i, _i, _u, _u,
                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib: Optional[ctypes.CDLL] = None


class ExtensionMissing(RuntimeError):
    pass


def load(require_cuda: bool = False) -> ctypes.CDLL:
    """Load (once) and type the shared library.  ``require_cuda`` additionally insists on a usable GPU."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). visfly_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype, fn.argtypes = res, args
        got = lib.vf_abi_version()
        if got != ABI_VERSION:
            raise ExtensionMissing(f"libvisfly_b200.so has ABI {got}, python side expects {ABI_VERSION}: rebuild")
        if lib.vf_params_size() != ctypes.sizeof(VfParams):
            raise ExtensionMissing("struct VfParams differs between libvisfly_b200.so and visfly_b200/params.py")
        if lib.vf_env_spec_size() != ctypes.sizeof(VfEnvSpec):
            raise ExtensionMissing("struct VfEnvSpec differs between libvisfly_b200.so and visfly_b200/params.py")
        _lib = lib
    if require_cuda and not th.cuda.is_available():
        raise RuntimeError("visfly_b200 needs a CUDA device (sm_100a): torch.cuda.is_available() is False "
                           "and there is no CPU fallback")
    return _lib


_fast = None


def fast():
    """The torch-extension plumbing module (``visfly_b200/_vf_torch*.so``, csrc/vf_torch.cpp): output allocation +
    C-ABI launch in one Python->C++ transition for the two per-step hot calls.  Same kernels, same entry points as
    the ctypes binding; required (no silent slow path) — build it with ``__graft_entry__.build()``."""
    global _fast
    if _fast is None:
        load()                                   # the ABI / struct-size checks, and the library itself
        try:
            from . import _vf_torch
        except ImportError as e:
            raise ExtensionMissing(
                "visfly_b200/_vf_torch*.so not found or not loadable: build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` ({e})") from e
        if _vf_torch.abi_version() != ABI_VERSION:
            raise ExtensionMissing("_vf_torch was built against a different libvisfly_b200.so: rebuild")
        _fast = _vf_torch
    return _fast


def _check(rc: int):
    if rc != 0:
        raise RuntimeError("visfly_b200: " + load().vf_last_error().decode())


def _dev_ptr(t: Optional[th.Tensor], what: str) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda or t.dtype != th.float32 or not t.is_contiguous():
        raise ValueError(f"{what} must be a contiguous float32 CUDA tensor (got {t.dtype}, {t.device}, "
                         f"contiguous={t.is_contiguous()})")
    return t.data_ptr()


def _stream(device) -> int:
    return th.cuda.current_stream(device).cuda_stream


def step_fwd(params: VfParams, substeps: int, integrator: int, action_type: int, flags: int,
             state_in: th.Tensor, action: th.Tensor, state_out: th.Tensor, obs_out: Optional[th.Tensor],
             ext_out: Optional[th.Tensor], wind: Optional[th.Tensor] = None, fifo_push: Optional[th.Tensor] = None,
             fifo_copy: Optional[th.Tensor] = None) -> None:
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    with th.cuda.device(state_in.device):
        _check(lib.vf_step_fwd(ctypes.byref(params), n, substeps, integrator, action_type, flags,
                               _dev_ptr(state_in, "state_in"), _dev_ptr(action, "action"),
                               _dev_ptr(state_out, "state_out"), _dev_ptr(obs_out, "obs_out"),
                               _dev_ptr(ext_out, "ext_out"), _dev_ptr(wind, "wind"),
                               _dev_ptr(fifo_push, "fifo_push"), _dev_ptr(fifo_copy, "fifo_copy"),
                               _stream(state_in.device)))


def fifo_rows(rows) -> Optional[VfFifoRows]:
    """``struct VfFifoRows`` over a list of engine-owned (n,4) FIFO entries (oldest first); None for an empty list."""
    if not rows:
        return None
    if len(rows) > FIFO_MAX_ROWS:
        raise ValueError(f"comm-delay FIFO deeper than VF_FIFO_MAX_ROWS ({FIFO_MAX_ROWS})")
    r = VfFifoRows()
    for j, t in enumerate(rows):
        r.row[j] = _dev_ptr(t, "fifo row")
    r.depth = len(rows)
    return r


def step_fwd_ring(params: VfParams, substeps: int, integrator: int, action_type: int, flags: int,
                  state_in: th.Tensor, rows, fifo_push: th.Tensor, state_out: th.Tensor,
                  obs_out: Optional[th.Tensor], ext_out: Optional[th.Tensor] = None,
                  wind: Optional[th.Tensor] = None) -> None:
    """Binding of ``vf_step_fwd_ring``: the control step with the comm-delay FIFO (``rows``, oldest first) shifted in
    place by the launch — consumes ``rows[0]``, appends ``fifo_push``."""
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    ring = fifo_rows(rows)
    with th.cuda.device(state_in.device):
        _check(lib.vf_step_fwd_ring(ctypes.byref(params), n, substeps, integrator, action_type, flags,
                                    _dev_ptr(state_in, "state_in"), None if ring is None else ctypes.byref(ring),
                                    _dev_ptr(fifo_push, "fifo_push"), _dev_ptr(state_out, "state_out"),
                                    _dev_ptr(obs_out, "obs_out"), _dev_ptr(ext_out, "ext_out"), _dev_ptr(wind, "wind"),
                                    _stream(state_in.device)))


def step_bwd(params: VfParams, substeps: int, integrator: int, action_type: int, flags: int,
             state_in: th.Tensor, action: th.Tensor, g_state_out: Optional[th.Tensor],
             g_obs: Optional[th.Tensor], g_state_in: th.Tensor, g_action: th.Tensor,
             wind: Optional[th.Tensor] = None) -> None:
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    with th.cuda.device(state_in.device):
        _check(lib.vf_step_bwd(ctypes.byref(params), n, substeps, integrator, action_type, flags,
                               _dev_ptr(state_in, "state_in"), _dev_ptr(action, "action"),
                               _dev_ptr(g_state_out, "grad_state_out"), _dev_ptr(g_obs, "grad_obs"),
                               _dev_ptr(g_state_in, "grad_state_in"), _dev_ptr(g_action, "grad_action"),
                               _dev_ptr(wind, "wind"), _stream(state_in.device)))


def step_fwd_host(params: VfParams, substeps: int, integrator: int, action_type: int, flags: int,
                  state_in: th.Tensor, action_host: th.Tensor, action_dev: th.Tensor, state_out: th.Tensor,
                  obs_dev: Optional[th.Tensor], obs_host: Optional[th.Tensor]) -> None:
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    for t, what in ((action_host, "action_host"), (obs_host, "obs_host")):
        if t is not None and (t.is_cuda or t.dtype != th.float32 or not t.is_contiguous()):
            raise ValueError(f"{what} must be a contiguous float32 host tensor")
    with th.cuda.device(state_in.device):
        _check(lib.vf_step_fwd_host(ctypes.byref(params), n, substeps, integrator, action_type, flags,
                                    _dev_ptr(state_in, "state_in"), action_host.data_ptr(),
                                    _dev_ptr(action_dev, "action_dev"), _dev_ptr(state_out, "state_out"),
                                    _dev_ptr(obs_dev, "obs_dev"),
                                    None if obs_host is None else obs_host.data_ptr(),
                                    _stream(state_in.device)))


def pack_state(n: int, state: th.Tensor, index: Optional[th.Tensor] = None, pos=None, quat=None, vel=None,
               rate=None, motor=None, alpha=None) -> None:
    lib = load(require_cuda=True)
    fields = [pos, quat, vel, rate, motor, alpha]
    m = n
    for f in fields:
        if f is not None:
            m = f.shape[0]
    if index is not None:
        if index.dtype != th.int64 or not index.is_cuda or not index.is_contiguous():
            raise ValueError("index must be a contiguous int64 CUDA tensor")
        m = index.numel()
    with th.cuda.device(state.device):
        _check(lib.vf_pack_state(n, m, None if index is None else index.data_ptr(),
                                 *[_dev_ptr(f, "field") for f in fields], _dev_ptr(state, "state"),
                                 _stream(state.device)))


def unpack_state(n: int, state: th.Tensor, pos=None, quat=None, vel=None, rate=None, motor=None,
                 alpha=None) -> None:
    lib = load(require_cuda=True)
    with th.cuda.device(state.device):
        _check(lib.vf_unpack_state(n, _dev_ptr(state, "state"),
                                   *[_dev_ptr(f, "field") for f in (pos, quat, vel, rate, motor, alpha)],
                                   _stream(state.device)))


def _any_ptr(t: Optional[th.Tensor], what: str, dtype) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{what} must be a contiguous {dtype} CUDA tensor (got {t.dtype}, {t.device})")
    return t.data_ptr()


def pack_status(step_count: th.Tensor, returns: th.Tensor, ebits: th.Tensor, gate: Optional[th.Tensor] = None,
                gates_passed: Optional[th.Tensor] = None) -> th.Tensor:
    """Per-agent env status records ``int32[n][4]`` (``VfEnvStatus`` in include/visfly_b200.h) from its fields."""
    n = step_count.numel()
    st = th.zeros((n, 4), dtype=th.int32, device=step_count.device)
    st[:, 0] = step_count.to(th.int32)
    st[:, 1] = returns.detach().to(th.float32).contiguous().view(th.int32)
    st[:, 2] = ebits.to(th.int32) & 0xFF
    if gate is not None:
        st[:, 2] |= gate.to(th.int32) << 8
    if gates_passed is not None:
        st[:, 3] = gates_passed.to(th.int32)
    return st


def unpack_status(status: th.Tensor):
    """``(step_count int32, returns float32, ebits int32, gate int32, gates_passed int32)`` views / copies."""
    return (status[:, 0], status[:, 1].view(th.float32), status[:, 2] & 0xFF, (status[:, 2] >> 8) & 0xFF,
            status[:, 3])


def env_step_fwd(params: VfParams, spec: VfEnvSpec, substeps: int, integrator: int, action_type: int, flags: int,
                 env_flags: int, step_index: int, state_in: th.Tensor, action: th.Tensor,
                 reset_table: Optional[th.Tensor], status_in: th.Tensor, state_out: th.Tensor, status_out: th.Tensor,
                 obs_out: th.Tensor, reward_out: th.Tensor, done_out: th.Tensor, record_out: th.Tensor,
                 term_obs_out: Optional[th.Tensor], host_mirror: Optional[VfEnvMirror] = None,
                 wind: Optional[th.Tensor] = None, fifo_push: Optional[th.Tensor] = None,
                 fifo_copy: Optional[th.Tensor] = None, step_base: Optional[th.Tensor] = None,
                 gate_out: Optional[th.Tensor] = None, peer_returns: Optional[VfPeerScatter] = None) -> None:
    """Binding of ``vf_env_step_fwd`` (fused control step + env wrapper tail, one launch)."""
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    with th.cuda.device(state_in.device):
        _check(lib.vf_env_step_fwd(
            ctypes.byref(params), ctypes.byref(spec), n, substeps, integrator, action_type, flags, env_flags,
            step_index, _any_ptr(step_base, "step_base", th.int64),
            _dev_ptr(state_in, "state_in"), _dev_ptr(action, "action"), _dev_ptr(wind, "wind"),
            _dev_ptr(fifo_push, "fifo_push"), _dev_ptr(reset_table, "reset_table"),
            _any_ptr(status_in, "status_in", th.int32),
            _dev_ptr(state_out, "state_out"), _any_ptr(status_out, "status_out", th.int32),
            _dev_ptr(fifo_copy, "fifo_copy"), _dev_ptr(obs_out, "obs_out"), _dev_ptr(reward_out, "reward_out"),
            _any_ptr(done_out, "done_out", th.bool), _dev_ptr(record_out, "record_out"),
            _dev_ptr(term_obs_out, "term_obs_out"), _any_ptr(gate_out, "gate_out", th.int64),
            None if host_mirror is None else ctypes.byref(host_mirror),
            None if peer_returns is None else ctypes.byref(peer_returns), _stream(state_in.device)))


def env_step_bwd(params: VfParams, spec: VfEnvSpec, substeps: int, integrator: int, action_type: int, flags: int,
                 env_flags: int, state_in: th.Tensor, action: th.Tensor, status_in: th.Tensor,
                 g_state_out: Optional[th.Tensor], g_obs: Optional[th.Tensor], g_reward: Optional[th.Tensor],
                 g_state_in: th.Tensor, g_action: th.Tensor, wind: Optional[th.Tensor] = None) -> None:
    """Binding of ``vf_env_step_bwd`` (adjoint of the fused env step, one launch)."""
    lib = load(require_cuda=True)
    n = state_in.shape[1]
    with th.cuda.device(state_in.device):
        _check(lib.vf_env_step_bwd(
            ctypes.byref(params), ctypes.byref(spec), n, substeps, integrator, action_type, flags, env_flags,
            _dev_ptr(state_in, "state_in"), _dev_ptr(action, "action"), _dev_ptr(wind, "wind"),
            _any_ptr(status_in, "status_in", th.int32),
            _dev_ptr(g_state_out, "grad_state_out"), _dev_ptr(g_obs, "grad_obs"), _dev_ptr(g_reward, "grad_reward"),
            _dev_ptr(g_state_in, "grad_state_in"), _dev_ptr(g_action, "grad_action"), _stream(state_in.device)))


def env_finish(params: VfParams, spec: VfEnvSpec, env_flags: int, step_index: int, state_in: th.Tensor,
               status_in: th.Tensor, reward: th.Tensor, success: Optional[th.Tensor], failure: Optional[th.Tensor],
               want_obs: bool = True, wind: Optional[th.Tensor] = None, reset_table: Optional[th.Tensor] = None,
               step_base: Optional[th.Tensor] = None, fifo=None, state_out: Optional[th.Tensor] = None,
               status_out: Optional[th.Tensor] = None):
    """Binding of ``vf_env_finish`` (wrapper tail for caller-defined tasks).  Allocates and returns
    ``(state_out, status_out, obs | None, done, record)``.  ``fifo``: engine-owned comm-delay FIFO entries whose rows
    are zeroed in place for re-initialised agents; ``state_out / status_out``: write there instead of allocating
    (``status_out`` may be ``status_in`` itself)."""
    lib = load(require_cuda=True)
    n, dev = state_in.shape[1], state_in.device
    state_out = th.empty_like(state_in) if state_out is None else state_out
    status_out = th.empty_like(status_in) if status_out is None else status_out
    ring = fifo_rows(fifo)
    obs = th.empty((n, 13), dtype=th.float32, device=dev) if want_obs else None
    done = th.empty((n,), dtype=th.bool, device=dev)
    record = th.empty((n, 4), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _check(lib.vf_env_finish(
            ctypes.byref(params), ctypes.byref(spec), n, env_flags, step_index,
            _any_ptr(step_base, "step_base", th.int64), _dev_ptr(state_in, "state_in"), _dev_ptr(wind, "wind"),
            _dev_ptr(reset_table, "reset_table"), _any_ptr(status_in, "status_in", th.int32),
            _dev_ptr(reward, "reward"), _any_ptr(success, "success", th.bool), _any_ptr(failure, "failure", th.bool),
            _dev_ptr(state_out, "state_out"), _any_ptr(status_out, "status_out", th.int32),
            None if obs is None else obs.data_ptr(), done.data_ptr(), record.data_ptr(),
            None if ring is None else ctypes.byref(ring), _stream(dev)))
    return state_out, status_out, obs, done, record


def wait_flag(flag_addr: int, value: int, timeout_us: int = 10_000_000) -> None:
    """Spin (in C, no GIL games needed: microseconds) until the page-locked completion word equals ``value``."""
    _check(load().vf_wait_flag(flag_addr, value, timeout_us))


def export_pose_habitat(params: VfParams, state: th.Tensor, pose_out: th.Tensor, vel_out: Optional[th.Tensor]) -> None:
    """Binding of ``vf_export_pose_habitat``; outputs are float32 contiguous, CUDA or page-locked host tensors."""
    lib = load(require_cuda=True)
    n = state.shape[1]
    for t, what, w in ((pose_out, "pose_out", 7), (vel_out, "vel_out", 3)):
        if t is None:
            continue
        if t.dtype != th.float32 or not t.is_contiguous() or tuple(t.shape) != (n, w):
            raise ValueError(f"{what} must be a contiguous float32 ({n},{w}) tensor")
        if not t.is_cuda and not t.is_pinned():
            raise ValueError(f"{what} must live on the device or in page-locked host memory")
    with th.cuda.device(state.device):
        _check(lib.vf_export_pose_habitat(ctypes.byref(params), n, _dev_ptr(state, "state"), pose_out.data_ptr(),
                                          None if vel_out is None else vel_out.data_ptr(), _stream(state.device)))


def _policy_check(rc: int):
    if rc != 0:
        raise RuntimeError("visfly_b200: " + load().vf_policy_last_error().decode())


def policy_pack(params) -> th.Tensor:
    """Binding of ``vf_policy_pack``: (W1, b1, W2, b2, W3, b3) -> the packed weight block the actor kernels read."""
    lib = load(require_cuda=True)
    h, d = params[0].shape
    packed = th.empty((lib.vf_policy_packed_floats(h),), dtype=th.float32, device=params[0].device)
    with th.cuda.device(packed.device):
        _policy_check(lib.vf_policy_pack(d, h, *[_dev_ptr(p.detach(), "param") for p in params], packed.data_ptr(),
                                         _stream(packed.device)))
    return packed


def policy_fwd(xa: th.Tensor, xb: Optional[th.Tensor], packed: th.Tensor, h: int, lo: float, hi: float) -> th.Tensor:
    """Binding of ``vf_policy_fwd``: observation pieces ``xa (n, da)`` [, ``xb (n, db)``] (contiguous float32 CUDA
    tensors) and the packed weight block of ``policy_pack``."""
    lib = load(require_cuda=True)
    n, da = xa.shape
    db = 0 if xb is None else xb.shape[1]
    action = th.empty((n, 4), dtype=th.float32, device=xa.device)
    with th.cuda.device(xa.device):
        _policy_check(lib.vf_policy_fwd(n, da, db, h, _dev_ptr(xa, "xa"), _dev_ptr(xb, "xb"), packed.data_ptr(),
                                        float(lo), float(hi), action.data_ptr(), _stream(xa.device)))
    return action


def policy_bwd(xa: th.Tensor, xb: Optional[th.Tensor], packed: th.Tensor, h: int, lo: float, hi: float,
               g_action: th.Tensor, want_ga: bool, want_gb: bool):
    """Binding of ``vf_policy_bwd``: returns ``(grad_xa | None, grad_xb | None, flat parameter gradients)``."""
    lib = load(require_cuda=True)
    n, da = xa.shape
    db = 0 if xb is None else xb.shape[1]
    d = da + db
    g_a = th.empty_like(xa) if want_ga else None
    g_b = th.empty_like(xb) if (want_gb and xb is not None) else None
    partial = th.empty((lib.vf_policy_partial_floats(n, d, h),), dtype=th.float32, device=xa.device)
    flat = th.empty((h * d + h + h * h + h + 4 * h + 4,), dtype=th.float32, device=xa.device)
    with th.cuda.device(xa.device):
        _policy_check(lib.vf_policy_bwd(n, da, db, h, _dev_ptr(xa, "xa"), _dev_ptr(xb, "xb"), packed.data_ptr(),
                                        float(lo), float(hi), _dev_ptr(g_action, "grad_action"),
                                        _dev_ptr(g_a, "grad_xa"), _dev_ptr(g_b, "grad_xb"), partial.data_ptr(),
                                        flat.data_ptr(), _stream(xa.device)))
    return g_a, g_b, flat


def ingest_depth(src: th.Tensor, dst: th.Tensor, background: float = 20.0) -> None:
    """Binding of ``vf_ingest_depth``: ``src`` float32 (n,H,W) on the device or in page-locked host memory, ``dst``
    float32 (n,1,H,W) on the device."""
    lib = load(require_cuda=True)
    n, h, w = src.shape
    with th.cuda.device(dst.device):
        if lib.vf_ingest_depth(n, h, w, src.data_ptr(), dst.data_ptr(), float(background), _stream(dst.device)):
            raise RuntimeError("visfly_b200: " + lib.vf_sensor_last_error().decode())


def ingest_color(src: th.Tensor, dst: th.Tensor) -> None:
    """Binding of ``vf_ingest_color``: ``src`` uint8 (n,H,W,4), ``dst`` uint8 (n,3,H,W) on the device."""
    lib = load(require_cuda=True)
    n, h, w, _ = src.shape
    with th.cuda.device(dst.device):
        if lib.vf_ingest_color(n, h, w, src.data_ptr(), dst.data_ptr(), _stream(dst.device)):
            raise RuntimeError("visfly_b200: " + lib.vf_sensor_last_error().decode())
