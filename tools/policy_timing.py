"""Device time of the fused actor kernels (vf_policy_fwd / vf_policy_bwd) at the APG size; run once per kernel family:
    python tools/policy_timing.py            (tensor-core kernels)
    VF_POLICY_NO_TC=1 python tools/policy_timing.py   (CUDA-core kernels)"""
import os
import sys

import torch as th

sys.path.insert(0, ".")
from visfly_b200 import _lib  # noqa: E402

n, d, h = 65536, 16, 64
th.manual_seed(0)
dev = "cuda"
params = [th.randn(h, d, device=dev) * 0.3, th.randn(h, device=dev) * 0.1, th.randn(h, h, device=dev) * 0.2,
          th.randn(h, device=dev) * 0.1, th.randn(4, h, device=dev) * 0.2, th.randn(4, device=dev) * 0.1]
packed = _lib.policy_pack(params)
xa, xb = th.randn(n, 13, device=dev), th.randn(n, 3, device=dev)
g = th.randn(n, 4, device=dev)


def timed(fn, per_graph=20, reps=10):
    """Device time per call: the calls are recorded into a CUDA graph (the ctypes binding costs more host time per call
    than these kernels take on the device) and the replay is timed with events."""
    for _ in range(5):
        fn()
    th.cuda.synchronize()
    graph = th.cuda.CUDAGraph()
    with th.cuda.graph(graph):
        for _ in range(per_graph):
            fn()
    graph.replay()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * per_graph)


a = _lib.policy_fwd(xa, xb, packed, h, -1.0, 1.0)
x = th.cat([xa, xb], 1)
ref = th.clip(th.tanh(th.tanh(th.tanh(x @ params[0].T + params[1]) @ params[2].T + params[3]) @ params[4].T + params[5]), -1, 1)
x64 = x.double()
p64 = [p.double() for p in params]
ref64 = th.clip(th.tanh(th.tanh(th.tanh(x64 @ p64[0].T + p64[1]) @ p64[2].T + p64[3]) @ p64[4].T + p64[5]), -1, 1)
rl2 = lambda u, v: float((u.double() - v).norm() / v.norm())
print("kernels:", "CUDA cores" if os.environ.get("VF_POLICY_NO_TC") else "tensor cores (tcgen05)")
print(f"fwd rel-L2 vs fp64: {rl2(a, ref64):.2e}   (torch fp32 ops: {rl2(ref, ref64):.2e})")
print(f"fwd  {timed(lambda: _lib.policy_fwd(xa, xb, packed, h, -1.0, 1.0)):7.2f} us")
print(f"bwd  {timed(lambda: _lib.policy_bwd(xa, xb, packed, h, -1.0, 1.0, g, True, True)):7.2f} us  (incl. the partial-sum reduction)")
