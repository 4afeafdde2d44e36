"""cProfile of the Python side of env.step (fused path) — where do the microseconds go?"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from visfly_b200.envs import HoverEnv  # noqa: E402

n = 65536
env = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(action_type="bodyrate", integrator="rk4",
               dt=0.0025, ctrl_dt=0.02), tensor_output=True)
env.reset()
a = th.zeros(n, 4, device="cuda")
a[:, 0] = -1 / 3
for _ in range(50):
    env.step(a)
th.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    env.step(a)
th.cuda.synchronize()
dt = (time.perf_counter() - t0) / 2000
print(f"env.step wall: {dt * 1e6:.1f} us/step -> {n / dt:.3e} agent-steps/s")
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    env.step(a)
pr.disable()
th.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
