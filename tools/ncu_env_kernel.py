"""A few env.step calls of the headline workload (driver script for ncu captures of the fused env-step kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from bench import AGENTS, DYN, hover_actions  # noqa: E402
from visfly_b200.envs import HoverEnv  # noqa: E402

n = int(os.environ.get("VF_AGENTS", AGENTS))
env = HoverEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN), tensor_output=True)
env.reset()
acts = hover_actions(n, 4, "cuda")
flush = th.empty(256 << 20, dtype=th.uint8, device="cuda")
for i in range(int(os.environ.get("VF_STEPS", 12))):
    flush.zero_()
    env.step(acts[i % 4])
th.cuda.synchronize()
