"""Kernel list of one recorded task-env step (capture_task_step) of bench.py's custom-task env: torch.profiler over
a few graph replays, device time per kernel."""
import sys

import torch as th
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench  # noqa: E402
from visfly_b200.envs import HoverEnv  # noqa: E402


class UserHover(HoverEnv):
    def get_reward(self, predicted_obs=None):
        return (0.1 - (self.position - self.target).norm(dim=1) * (0.1 / 9)
                - (self.orientation - self._unit_quat).norm(dim=1) * 1e-5
                - self.velocity.norm(dim=1) * 0.002 - self.angular_velocity.norm(dim=1) * 0.002)


dev = th.device("cuda", 0)
n = bench.AGENTS
env = UserHover(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(bench.DYN), seed=77,
                max_episode_steps=256, tensor_output=True)
env.capture_task_step = True
env.reset()
acts = list(bench.hover_actions(n, 4, dev, seed=3).unbind(0))
for i in range(50):
    env.step(acts[i % 4])
th.cuda.synchronize()
R = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(R):
        env.step(acts[i % 4])
    th.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"replays {R}, device time per step {tot / R:.1f} us, kernels per step {sum(e.count for e in rows) / R:.1f}")
for e in rows:
    print(f"{e.device_time_total / R:8.2f} us  x{e.count / R:4.1f}  {e.key[:110]}")
