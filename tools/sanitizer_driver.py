"""Small driver for compute-sanitizer: every kernel of the library a few times, ragged batch (not a multiple of the
warp / block size), programmatic dependent launch on, comm-delay FIFO clone, host mirror + completion word, reset
branch taken.  Run as

    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitizer_driver.py

and keep the log under profiles/ (the four step kernels rely on PDL ordering and on out-of-place status records)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from visfly_b200.dynamics import Dynamics  # noqa: E402
from visfly_b200.envs import HoverEnv, NavigationEnv, RacingEnv2  # noqa: E402
from visfly_b200.render_handoff import HabitatPoseExporter  # noqa: E402

n = 1000 + 37
dyn_kw = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06)
g = th.Generator().manual_seed(0)

# plain control step, forward + adjoint, FIFO clone (caller's buffer reused in place)
d = Dynamics(num=n, device="cuda", **dyn_kw)
buf = th.zeros(n, 4, device="cuda")
for t in range(5):
    buf.copy_((th.rand(n, 4, generator=g) * 2 - 1).cuda())
    d.step(buf)
a = ((th.rand(6, n, 4, generator=g) * 2 - 1) * 0.3).cuda().requires_grad_(True)
loss = 0.0
for t in range(6):
    loss = loss + d.step(a[t]).pow(2).sum()
loss.backward()
d.detach()
_ = d.acceleration, d.thrusts, d.full_state, d.extend_state
HabitatPoseExporter(d, host=True).export()
HabitatPoseExporter(d, host=False).export()
for act in ("velocity", "position", "thrust"):
    dd = Dynamics(num=n, device="cuda", **dict(dyn_kw, action_type=act, integrator="euler", dt=0.005))
    for t in range(3):
        dd.step((th.rand(n, 4, generator=g) * 2 - 1).cuda())

# fused env step: tensor mode, autograd mode, numpy mode (host mirror + completion word, one step ahead), resets
for cls in (HoverEnv, NavigationEnv, RacingEnv2):
    kw = dict(tensor_output=True) if cls is HoverEnv else {}
    env = cls(num_agent_per_scene=n, visual=False, device="cuda", dynamics_kwargs=dict(dyn_kw), max_episode_steps=4, **kw)
    env.reset()
    for t in range(10):                                   # crosses two rounds of Philox restarts
        env.step(((th.rand(n, 4, generator=g) * 2 - 1) * 0.3).cuda())
    assert env._fused.active
    env.requires_grad = True
    acts = ((th.rand(6, n, 4, generator=g) * 2 - 1) * 0.3).cuda().requires_grad_(True)
    loss = 0.0
    for t in range(6):
        obs, r, dn, info = env.step(acts[t])
        loss = loss - r.mean() + 1e-3 * obs["state"].pow(2).mean()
    loss.backward()
    env.detach()
    env.requires_grad = False
    env.tensor_output = False
    for t in range(9):
        obs, r, dn, info = env.step(((np.random.rand(n, 4) * 2 - 1) * 0.3).astype(np.float32))
    _ = [info[i] for i in np.nonzero(dn)[0][:3]]
    env.reset_agent_by_id([0, 5, n - 1])                  # drops the step in flight (rewind) and hands over
    env.tensor_output = True
    env.step(th.zeros(n, 4, device="cuda"))

# actor kernels: tensor-core path (tcgen05, default) for widths <= 16, CUDA-core backward above; two-piece observation
from visfly_b200.algorithms.policies import Actor  # noqa: E402
for d_obs, h in ((16, 64), (13, 32), (17, 64)):
    actor = Actor(d_obs, 4, (h, h)).cuda()
    xa = th.randn(n, d_obs - 3, device="cuda").requires_grad_(True)
    xb = th.randn(n, 3, device="cuda").requires_grad_(True)
    act = actor.deterministic_action({"a": xa, "b": xb}, -0.9, 0.9)
    (act * th.randn(n, 4, device="cuda")).sum().backward()
    actor.release_graph()

# task env with its own reward code: control step + vf_env_finish (FIFO rows zeroed in the launch), and the FIFO ring
class UserHover(HoverEnv):
    def get_reward(self, predicted_obs=None):
        return 0.1 - (self.position - self.target).norm(dim=1) * 0.01 - 0.001 * self._step_count


env = UserHover(num_agent_per_scene=n, visual=False, device="cuda", dynamics_kwargs=dict(dyn_kw), max_episode_steps=4,
                tensor_output=True)
env.reset()
for t in range(9):
    env.step(((th.rand(n, 4, generator=g) * 2 - 1) * 0.3).cuda())
assert env._split.active
ring = Dynamics(num=n, device="cuda", **dyn_kw)
ring._fifo_ring = True
with th.no_grad():
    for t in range(4):
        ring.step(((th.rand(n, 4, generator=g) * 2 - 1) * 0.3).cuda())
_ = ring.acceleration
th.cuda.synchronize()
print("sanitizer driver finished")
