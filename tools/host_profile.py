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

alloc_us, launch_us = stepper.profile(st, ac, fz.status, steps)
th.cuda.synchronize()
print(f"inside EnvStepper.step: slab allocation + 8 carved tensors {alloc_us:.2f} us, vf_env_step_fwd (checks + "
      f"cudaLaunchKernelEx) {launch_us:.2f} us  -> argument parsing + result tuple = the rest")

pr = cProfile.Profile()
pr.enable()
for i in range(steps):
    env.step(acts[i % 16])
pr.disable()
th.cuda.synchronize()
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(18)

# short brackets: what does a K-step loop cost right after a device synchronisation (the driver's K = 20 bracket)?
for K in (20, 200):
    rows = []
    for rep in range(8):
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        marks = []
        for i in range(K):
            env.step(acts[i % 16])
            if i in (0, 1, 4, K - 1):
                marks.append(time.perf_counter() - t0)
        t_enq = time.perf_counter() - t0
        e1.record()
        e1.synchronize()
        t_all = time.perf_counter() - t0
        rows.append((t_enq * 1e6 / K, t_all * 1e6 / K, e0.elapsed_time(e1) * 1e3 / K, [round(m * 1e6, 1) for m in marks]))
    print(f"K={K}: per-step us (enqueue, wall incl. sync, device) and cumulative enqueue marks at steps 1,2,5,K:")
    for r in rows:
        print("   %.2f  %.2f  %.2f  %s" % r)
