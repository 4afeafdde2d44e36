"""A short BPTT horizon through NavigationEnv (driver script for ncu captures of the adjoint kernel vf_env_step_bwd)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from bench import AGENTS, DYN  # noqa: E402
from visfly_b200.algorithms import BPTT  # noqa: E402
from visfly_b200.envs import NavigationEnv  # noqa: E402

n = int(os.environ.get("VF_AGENTS", AGENTS))
env = NavigationEnv(num_agent_per_scene=n, visual=False, dynamics_kwargs=dict(DYN), requires_grad=True,
                    max_episode_steps=256, random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                        {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
algo = BPTT(env, horizon=int(os.environ.get("VF_H", 4)), policy_kwargs=dict(net_arch=[64, 64]), make_eval_env=False,
            dump_step=1 << 62)
algo.learn(total_timesteps=int(os.environ.get("VF_UPDATES", 3)) * n * algo.H)
th.cuda.synchronize()
