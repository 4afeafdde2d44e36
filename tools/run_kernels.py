"""Launch the fused control-step kernels a few times (driver script for ncu captures and quick timings).

    python tools/run_kernels.py [--agents N] [--integrator rk4|euler] [--iters K] [--time]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import torch as th  # noqa: E402

from _util import pack, random_flight_state, vf_params  # noqa: E402
from visfly_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=65536)
    ap.add_argument("--integrator", default="rk4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    n = a.agents
    dt = 0.0025 if a.integrator == "rk4" else 0.005
    S = int(0.02 / dt)
    integ = 1 if a.integrator == "rk4" else 0
    P = vf_params("bodyrate", dt)
    st = pack(*random_flight_state(n, seed=1, spread=0.3)).cuda()
    ac = ((th.rand(n, 4) * 2 - 1) * 0.1).cuda()
    ac[:, 0] -= 1 / 3
    out, obs = th.empty_like(st), th.empty((n, 13), device="cuda")
    g_out, g_obs = th.randn_like(st), th.randn((n, 13), device="cuda")
    gs, ga = th.empty_like(st), th.empty_like(ac)
    flush = th.empty(256 << 20, dtype=th.uint8, device="cuda")

    def fwd():
        _lib.step_fwd(P, S, integ, 1, 1, st, ac, out, obs, None)

    def bwd():
        _lib.step_bwd(P, S, integ, 1, 1, st, ac, g_out, g_obs, gs, ga)

    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        th.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            th.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        if a.time:
            ts.sort()
            byt = (176 if name == "fwd" else 272) * n
            print(f"{name} {a.integrator} n={n}: median {ts[len(ts)//2]:.2f} us  min {ts[0]:.2f} us  "
                  f"-> {n / (ts[len(ts)//2] * 1e-6):.3e} agent-steps/s, {byt / (ts[len(ts)//2] * 1e-6) / 1e9:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
