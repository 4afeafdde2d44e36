"""env.step rate of one 65 536-agent HoverEnv, back to back: host enqueue time, wall time and device time per step.
Run twice to see what programmatic dependent launch buys:  VF_NO_PDL=1 python tools/step_rate_probe.py ; python tools/step_rate_probe.py
(diagnostic, not part of the product)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from bench import DYN, hover_actions  # noqa: E402
from visfly_b200.envs import HoverEnv  # noqa: E402

dev = th.device("cuda", 0)


def run(n, steps, replicas):
    envs = []
    for r in range(replicas):
        env = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=42 + r,
                       max_episode_steps=256, tensor_output=True)
        env.reset()
        envs.append(env)
    acts = list(hover_actions(n, 16, dev).unbind(0))
    for i in range(600):
        envs[i % replicas].step(acts[i % 16])
    th.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            envs[i % replicas].step(acts[i % 16])
        e1.record()
        t_launch = time.perf_counter() - t0
        th.cuda.synchronize()
        wall = time.perf_counter() - t0
        row = (wall / steps * 1e6, t_launch / steps * 1e6, e0.elapsed_time(e1) / steps * 1e3)
        best = row if best is None or row[0] < best[0] else best
    print(f"block={os.environ.get('VF_BLOCK', '64'):>3s} pdl={'off' if os.environ.get('VF_NO_PDL') else 'on '} n={n:8d} replicas={replicas:2d} steps={steps}: "
          f"wall {best[0]:6.2f} us/step  host-enqueue {best[1]:6.2f}  device {best[2]:6.2f}", flush=True)


if __name__ == "__main__":
    run(65536, 3000, 1)
    run(65536, 3000, 16)
    if not os.environ.get("VF_PROBE_SHORT"):
        run(16384, 3000, 1)      # host-bound: the enqueue cost alone
        run(1 << 20, 1000, 1)
