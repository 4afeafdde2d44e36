"""Where do the host microseconds of env.step go?  Host-bound batch (4096 agents), cProfile over the step loop and a
direct loop over the C++ stepper alone.  (diagnostic, not part of the product)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from bench import DYN, hover_actions  # noqa: E402
from visfly_b200.envs import HoverEnv  # noqa: E402

dev = th.device("cuda", 0)
n, steps = 4096, 20000
env = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=42,
               max_episode_steps=256, tensor_output=True)
env.reset()
acts = list(hover_actions(n, 16, dev).unbind(0))
for i in range(2000):
    env.step(acts[i % 16])
th.cuda.synchronize()

t0 = time.perf_counter()
for i in range(steps):
    env.step(acts[i % 16])
t1 = time.perf_counter()
th.cuda.synchronize()
print(f"env.step host loop: {(t1 - t0) / steps * 1e6:.2f} us/step")

fz = env._fused
stepper = fz._stepper or fz._make_stepper()
st, ac = env.envs.dynamics._state, acts[0]
t0 = time.perf_counter()
for i in range(steps):
    out = stepper.step(st, ac, fz.status, i, 0, True, 0)
t1 = time.perf_counter()
th.cuda.synchronize()
print(f"EnvStepper.step alone (alloc + carve + launch + 7-tuple): {(t1 - t0) / steps * 1e6:.2f} us/call")
t0 = time.perf_counter()
for i in range(steps):
    out = stepper.step(st, ac, fz.status, i, 0, False, 0)
t1 = time.perf_counter()
th.cuda.synchronize()
print(f"EnvStepper.step alone, no terminal obs: {(t1 - t0) / steps * 1e6:.2f} us/call")

pr = cProfile.Profile()
pr.enable()
for i in range(steps):
    env.step(acts[i % 16])
pr.disable()
th.cuda.synchronize()
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(18)
