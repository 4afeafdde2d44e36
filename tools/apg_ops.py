"""Which library ops (and how many launches of which kernels) one eager BPTT update is made of: torch.profiler over
one update of H steps (NavigationEnv, 65 536 agents), per-step averages."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import AGENTS, DYN  # noqa: E402
from visfly_b200.algorithms import BPTT  # noqa: E402
from visfly_b200.envs import NavigationEnv  # noqa: E402

n, H = AGENTS, 8
env = NavigationEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN), requires_grad=True,
                    max_episode_steps=256, random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                        {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
algo = BPTT(env, horizon=H, policy_kwargs=dict(net_arch=[64, 64]), make_eval_env=False, dump_step=1 << 62)
algo.learn(total_timesteps=3 * n * H)
th.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=bool(os.environ.get('VF_STACKS'))) as prof:
    algo.learn(total_timesteps=n * H)
    th.cuda.synchronize()
ev = prof.key_averages()
print("---- device kernels, per env step ----")
rows = sorted((e for e in ev if e.device_time_total > 0 and e.cpu_time_total == 0), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step {tot / H:.1f} us, launches per step {sum(e.count for e in rows) / H:.1f}")
for e in rows[:28]:
    print(f"{e.device_time_total / H:8.2f} us  x{e.count / H:5.2f}  {e.key[:100]}")
print("---- aten ops, per env step (count) ----")
ops = sorted((e for e in ev if e.key.startswith("aten::") or "Backward" in e.key or "Function" in e.key),
             key=lambda e: -e.count)
for e in ops[:45]:
    print(f"x{e.count / H:6.2f}  {e.key[:80]}")

if os.environ.get("VF_STACKS"):
    print("---- where the small ops come from ----")
    for e in sorted(prof.key_averages(group_by_stack_n=8), key=lambda e: -e.count):
        if e.key in ("aten::cat", "aten::add_", "aten::mul", "aten::add", "aten::sub", "aten::neg", "aten::clone", "aten::copy_",
                     "aten::fill_", "aten::bitwise_not") and e.count >= H - 1:
            frames = [f for f in e.stack if "visfly_b200" in f or "autograd" in f.lower()][:4]
            print(f"x{e.count / H:5.2f} {e.key:18s} {' <- '.join(f.split('/')[-1][:60] for f in frames)}")
