"""Where does env.step's wall time go under the bench's exact conditions?  (diagnostic, not part of the product)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from bench import DYN, hover_actions  # noqa: E402
from visfly_b200.envs import HoverEnv  # noqa: E402

n = 65536
dev = th.device("cuda", 0)


def run(label, steps, pool, max_ep, const_action=False, keep_term=True):
    env = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=42,
                   max_episode_steps=max_ep, tensor_output=True)
    env.keep_terminal_observation = keep_term
    env.reset()
    acts = list(hover_actions(n, pool, dev).unbind(0))
    for i in range(50):
        env.step(acts[i % pool])
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        env.step(acts[0] if const_action else acts[i % pool])
    e1.record()
    t_launch = time.perf_counter() - t0
    th.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"{label:44s} steps={steps:5d} host-enqueue {t_launch / steps * 1e6:6.2f} us/step  wall {wall / steps * 1e6:6.2f} "
          f"us/step  device {e0.elapsed_time(e1) / steps * 1e3:6.2f} us/step", flush=True)


run("const action, 2000 steps, max_ep 1000", 2000, 16, 1000, const_action=True)
run("rotating actions, 2000 steps, max_ep 1000", 2000, 16, 1000)
run("rotating actions, 200 steps, max_ep 1000", 200, 16, 1000)
run("rotating actions, 200 steps, max_ep 256", 200, 16, 256)
run("rotating actions, 2000 steps, max_ep 256", 2000, 16, 256)
run("rotating, 2000 steps, no terminal obs", 2000, 16, 1000, keep_term=False)

# --- replicate bench.py's sequence: cold pass (flush + event pair per step), barrier, hot pass of 200 steps -----
import gc  # noqa: E402

from bench import timed_steps  # noqa: E402


def bench_like(label, K=200, gc_off=False, per_step_times=False):
    env = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=42,
                   max_episode_steps=256, tensor_output=True)
    env.reset()
    acts = list(hover_actions(n, 16, dev).unbind(0))
    flush = th.empty(256 << 20, dtype=th.uint8, device=dev)
    stream = th.cuda.current_stream(dev)
    step = lambda i: env.step(acts[i % 16])
    for i in range(20):
        step(i)
    th.cuda.synchronize()
    timed_steps(step, K, flush, stream)
    th.cuda.synchronize()
    if gc_off:
        gc.collect()
        gc.disable()
    ts = []
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(K):
        if per_step_times:
            a = time.perf_counter()
        step(i)
        if per_step_times:
            ts.append(time.perf_counter() - a)
    e1.record(stream)
    th.cuda.synchronize()
    wall = time.perf_counter() - t0
    gc.enable()
    msg = f"{label:44s} K={K} wall {wall / K * 1e6:6.2f} us/step device {e0.elapsed_time(e1) / K * 1e3:6.2f} us/step"
    if per_step_times:
        big = sorted(((t, i) for i, t in enumerate(ts)), reverse=True)[:5]
        msg += "  slowest host steps: " + ", ".join(f"#{i}:{t * 1e6:.0f}us" for t, i in big)
    print(msg, flush=True)


bench_like("bench-like")
bench_like("bench-like, gc disabled", gc_off=True)
bench_like("bench-like, per-step host times", per_step_times=True)
