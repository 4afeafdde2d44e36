"""The C-ABI shared library: it loads, exports every symbol include/visfly_b200.h declares, agrees with the
Python mirror of ``struct VfParams``, validates arguments without a GPU, and the product fails loudly (no CPU
fallback) when CUDA or the extension is missing."""
import ctypes
import os
import re

import pytest
import torch as th

from _util import ROOT, vf_params
from visfly_b200 import _lib
from visfly_b200.params import VfParams

HEADER = os.path.join(ROOT, "include", "visfly_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert {"vf_step_fwd", "vf_step_bwd", "vf_step_fwd_host", "vf_pack_state", "vf_unpack_state",
            "vf_abi_version", "vf_last_error", "vf_params_size", "vf_device_sm_count"} <= set(names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in visfly_b200/_lib.py"


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    hdr = open(HEADER).read()
    assert lib.vf_abi_version() == int(re.search(r"#define VF_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION
    assert lib.vf_params_size() == ctypes.sizeof(VfParams)
    # field order in the header == field order of the ctypes mirror
    body = re.search(r"typedef struct VfParams \{(.*?)\} VfParams;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"float\s+([A-Za-z_]+)", body)
    assert fields == [f[0] for f in VfParams._fields_]
    from visfly_b200.params import VfEnvMirror, VfEnvSpec, VfFifoRows, VfPeerScatter
    assert lib.vf_env_spec_size() == ctypes.sizeof(VfEnvSpec)
    for name, cls in (("VfEnvSpec", VfEnvSpec), ("VfEnvMirror", VfEnvMirror), ("VfFifoRows", VfFifoRows),
                      ("VfPeerScatter", VfPeerScatter)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = re.findall(r"(?:float|int|unsigned long long|long long|unsigned)\s*\*?\s+([A-Za-z_]+)", body)
        assert fields == [f[0] for f in cls._fields_], name


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    p = vf_params()
    assert lib.vf_step_fwd(None, 4, 4, 0, 1, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"params" in lib.vf_last_error()
    assert lib.vf_step_fwd(ctypes.byref(p), 4, 0, 0, 1, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"substeps" in lib.vf_last_error()
    assert lib.vf_step_fwd(ctypes.byref(p), 4, 4, 7, 1, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"integrator" in lib.vf_last_error()
    assert lib.vf_step_fwd(ctypes.byref(p), 4, 4, 0, 4, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"action_type" in lib.vf_last_error()
    # velocity (2) / position (3) run forward only: the adjoint entry points refuse them
    assert lib.vf_step_bwd(ctypes.byref(p), 4, 4, 0, 3, 1, None, None, None, None, None, None, None, None) != 0
    assert b"no gradient for the velocity / position" in lib.vf_last_error()
    assert lib.vf_step_fwd(ctypes.byref(p), 4, 4, 0, 1, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"NULL" in lib.vf_last_error()
    assert lib.vf_step_bwd(ctypes.byref(p), 4, 65, 0, 1, 1, None, None, None, None, None, None, None, None) != 0
    assert b"VF_MAX_SUBSTEPS_BWD" in lib.vf_last_error()
    x = ctypes.c_void_p(16)
    assert lib.vf_step_fwd(ctypes.byref(p), 4, 4, 0, 1, 1, x, x, ctypes.c_void_p(32), None, None, None, x, None, None) != 0
    assert b"fifo_push and fifo_copy" in lib.vf_last_error()
    # the device-resident FIFO ring
    from visfly_b200.params import VfFifoRows
    ring = VfFifoRows()
    assert lib.vf_step_fwd_ring(ctypes.byref(p), 4, 4, 0, 1, 1, x, None, x, ctypes.c_void_p(32), None, None, None, None) != 0
    assert b"fifo_rows" in lib.vf_last_error()
    ring.depth = 9
    assert lib.vf_step_fwd_ring(ctypes.byref(p), 4, 4, 0, 1, 1, x, ctypes.byref(ring), x, ctypes.c_void_p(32), None, None,
                                None, None) != 0
    assert b"VF_FIFO_MAX_ROWS" in lib.vf_last_error()
    ring.depth, ring.row[0], ring.row[1] = 2, 64, 64
    assert lib.vf_step_fwd_ring(ctypes.byref(p), 4, 4, 0, 1, 1, x, ctypes.byref(ring), x, ctypes.c_void_p(32), None, None,
                                None, None) != 0
    assert b"distinct" in lib.vf_last_error()
    ring.row[1] = 16
    assert lib.vf_step_fwd_ring(ctypes.byref(p), 4, 4, 0, 1, 1, x, ctypes.byref(ring), x, ctypes.c_void_p(32), None, None,
                                None, None) != 0
    assert b"alias a FIFO row" in lib.vf_last_error()
    assert lib.vf_wait_flag(None, 1, 10) != 0
    flag = ctypes.c_uint(7)
    assert lib.vf_wait_flag(ctypes.byref(flag), 7, 10) == 0                 # already raised
    assert lib.vf_wait_flag(ctypes.byref(flag), 8, 2000) != 0 and b"timed out" in lib.vf_last_error()
    # empty batch is a no-op, not an error
    assert lib.vf_step_fwd(ctypes.byref(p), 0, 4, 0, 1, 1, None, None, None, None, None, None, None, None, None) == 0


def test_params_match_reference_constants():
    """SURVEY.md App. A values (printed from the live reference object)."""
    p = vf_params("bodyrate", 0.005)
    assert abs(p.mass - 0.46) < 1e-7 and abs(p.thrust_max - 5.0155) < 1e-4
    assert abs(p.motor_c - 0.8594) < 1e-4
    assert abs(p.act_half[0] - 14.715) < 1e-5 and abs(p.act_mean[0] - 14.715) < 1e-5
    assert p.act_half[1] == 2.0 and p.act_mean[1] == 0.0
    assert abs(p.B[4] - 0.0530330) < 1e-6 and abs(p.B_inv[1] - 4.7140) < 1e-3 and abs(p.B_inv[3] - 15.625) < 1e-3
    assert abs(p.k_quad[2] - 9.1875e-3) < 1e-8
    assert abs(p.JKp[0] - 0.00101 * 60) < 1e-8 and abs(p.Kd[8] - 0.002) < 1e-9
    assert abs(vf_params("bodyrate", 0.0025).motor_c - 0.9270) < 1e-4


@pytest.mark.skipif(th.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    from visfly_b200.dynamics import Dynamics
    with pytest.raises(RuntimeError, match="CUDA"):
        Dynamics(num=4, device="cuda")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Dynamics(num=4, device="cpu")
    with pytest.raises(RuntimeError):
        _lib.step_fwd(vf_params(), 4, 0, 1, 1, th.zeros(5, 4, 4), th.zeros(4, 4), th.zeros(5, 4, 4), None, None)


def test_missing_extension_raises(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvisfly_b200.so")
    with pytest.raises(_lib.ExtensionMissing):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; nothing under visfly_b200/ may reference it."""
    pkg = os.path.join(ROOT, "visfly_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "torch_oracle" not in text and "host_mirror.so" not in text, f
