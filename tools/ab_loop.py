"""A/B probe of the env.step loop (diagnostic): hot single-env loop and cold 16-replica loop, device us per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from bench import DYN, hover_actions
from visfly_b200.envs import HoverEnv
dev = th.device("cuda", 0)
n = 65536
delay = float(os.environ.get("VF_AB_DELAY", DYN["comm_delay"]))
dyn = dict(DYN, comm_delay=delay)
envs = [HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(dyn), seed=42 + j,
                 max_episode_steps=256, tensor_output=True) for j in range(16)]
for e in envs:
    e.reset()
    e.keep_terminal_observation = os.environ.get("VF_AB_NOTERM") is None
acts = list(hover_actions(n, 16, dev).unbind(0))
def loop(es, k):
    for i in range(600):
        es[i % len(es)].step(acts[i % 16])
    out = []
    for rep in range(5):
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            es[i % len(es)].step(acts[i % 16])
        e1.record(); e1.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / k)
    return sorted(out)[2]
print(f"delay={delay} prefetch_off={os.environ.get('VF_NO_PREFETCH')} noterm={os.environ.get('VF_AB_NOTERM')}: "
      f"hot {loop(envs[:1], 400):.2f} us/step   cold(16 replicas) {loop(envs, 400):.2f} us/step")
