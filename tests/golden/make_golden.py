"""Generate the golden fixtures in this directory from the REAL reference (``/root/reference``).

Run in the build container (the reference tree is not available on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified through ``tests/_reference.py`` (RK4 repairs R1-R3 of SURVEY.md §8c as
monkeypatches).  Every fixture records its inputs, so tests replay the same seeded inputs through the oracle
(``oracle/torch_oracle.py``) and through the CUDA engine and compare with what the reference produced.

Fixtures
  kat.npz         single-agent known-answer vectors (the setups of SURVEY.md App. B)
  traj_<cfg>.npz  64-step closed-loop-free trajectories of 32 agents, float32 and float64 reference
  step_<cfg>.npz  one control step from a mid-flight batch incl. internal state (motor speeds, angular acc.)
  grad_<cfg>.npz  autograd gradients of a discounted hover-style return through an 8-step rollout
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from _reference import default_dtype, make_reference_dynamics, reference_available  # noqa: E402
from _util import random_flight_state  # noqa: E402

CONFIGS = {
    # name: (dynamics kwargs)
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4_nolag": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=False),
    "euler_thrust": dict(action_type="thrust", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
    # 12 sub-steps (the reference's default ctrl_dt).  float32 only: the reference's own `ctrl_dt % dt == 0`
    # check (dynamics.py:71) is evaluated in the default dtype and rejects 0.03 / 0.0025 in float64.
    "rk4_s12": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.03, comm_delay=0.06, ctrl_delay=True),
    # velocity / position set-points (geometric attitude controller, reference dynamics.py:414-496).  Forward only:
    # the reference's backward raises on this branch (in-place writes in its per-agent loop, :446-450).
    "euler_velocity": dict(action_type="velocity", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
    "rk4_velocity": dict(action_type="velocity", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=True),
    "euler_position": dict(action_type="position", integrator="euler", dt=0.005, ctrl_dt=0.02, comm_delay=0.0, ctrl_delay=False),
    "rk4_position": dict(action_type="position", integrator="rk4", dt=0.0025, ctrl_dt=0.02, comm_delay=0.06, ctrl_delay=True),
}
F32_ONLY = {"rk4_s12"}
FWD_ONLY = {"euler_velocity", "rk4_velocity", "euler_position", "rk4_position"}


def dtypes_for(name):
    return ((th.float32, "f32"),) if name in F32_ONLY else ((th.float32, "f32"), (th.float64, "f64"))


def actions(T, n, seed, law):
    g = th.Generator().manual_seed(seed)
    if law == "uniform":
        return th.rand(T, n, 4, generator=g) * 2 - 1
    # smooth-hover: a0 = -1/3 is 1 g for the default drone
    a = (th.rand(T, n, 4, generator=g) * 2 - 1) * 0.1
    a[..., 0] += -1.0 / 3.0
    return a


def ref_internal(d):
    """(pos, quat, vel, rate, motor, alpha) as (n,k) float arrays from a reference Dynamics object."""
    return [x.detach().clone() for x in (d._position.T, d._orientation.toTensor().T, d._velocity.T,
                                         d._angular_velocity.T, d._motor_omega.T, d._angular_acc.T)]


def make_kat():
    out = {}
    a = th.tensor([[0.1, 0.3, -0.2, 0.5]])
    for name, kw, steps in [
        ("B1_euler", dict(dt=0.005, integrator="euler", comm_delay=0.0), 3),
        ("B2_euler_fifo3", dict(dt=0.005, integrator="euler", comm_delay=0.06), 4),
        ("B3_rk4", dict(dt=0.0025, integrator="rk4", comm_delay=0.0), 3),
    ]:
        d = make_reference_dynamics(1, action_type="bodyrate", ctrl_dt=0.02, **kw)
        d.reset(pos=th.tensor([[1.0, 0.0, 1.5]]), ori=th.tensor([[1.0, 0, 0, 0]]), vel=th.zeros(1, 3),
                ori_vel=th.zeros(1, 3))
        fs, al = [], []
        for _ in range(steps):
            d.step(a.clone())
            fs.append(d.full_state.clone())
            al.append(d.angular_acceleration.clone())
        out[name + "_full_state"] = th.cat(fs).numpy()
        out[name + "_alpha"] = th.cat(al).numpy()
    out["action"] = a.numpy()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)


def make_traj(name, kw, n=32, T=64):
    out = {}
    init = random_flight_state(n, seed=11, spread=0.5, dtype=th.float64)
    for law in ("uniform", "hover"):
        acts = actions(T, n, 7, law)
        out[f"actions_{law}"] = acts.numpy()
        for dtype, tag in dtypes_for(name):
            with default_dtype(dtype):
                d = make_reference_dynamics(n, dtype=dtype, **kw)
                pos, quat, vel, rate = (x.to(dtype) for x in init[:4])
                d.reset(pos=pos.clone(), ori=quat.clone(), vel=vel.clone(), ori_vel=rate.clone())
                states = []
                for t in range(T):
                    states.append(d.step(acts[t].to(dtype).clone()).clone())
                out[f"states_{law}_{tag}"] = th.stack(states).numpy()
                out[f"final_full_state_{law}_{tag}"] = d.full_state.numpy()
                out[f"final_alpha_{law}_{tag}"] = d.angular_acceleration.numpy()
    for k, x in zip(("pos", "quat", "vel", "rate"), init[:4]):
        out["init_" + k] = x.numpy()
    np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), **out)


def make_step(name, kw, n=64, warm=6):
    """One step from mid-flight: internal state before, delayed action, internal state + outputs after."""
    kw = dict(kw, comm_delay=0.0)       # the FIFO is host-side plumbing; the kernel sees the delayed action
    out = {}
    init = random_flight_state(n, seed=21, spread=1.0, dtype=th.float64)
    acts = actions(warm + 1, n, 23, "uniform")
    for dtype, tag in dtypes_for(name):
        with default_dtype(dtype):
            d = make_reference_dynamics(n, dtype=dtype, **kw)
            pos, quat, vel, rate = (x.to(dtype) for x in init[:4])
            d.reset(pos=pos.clone(), ori=quat.clone(), vel=vel.clone(), ori_vel=rate.clone())
            for t in range(warm):
                d.step(acts[t].to(dtype).clone())
            before = ref_internal(d)
            obs = d.step(acts[warm].to(dtype).clone()).clone()
            after = ref_internal(d)
            for k, b, a in zip(("pos", "quat", "vel", "rate", "motor", "alpha"), before, after):
                out[f"before_{k}_{tag}"] = b.numpy()
                out[f"after_{k}_{tag}"] = a.numpy()
            out[f"obs_{tag}"] = obs.numpy()
            out[f"acc_{tag}"] = d.acceleration.clone().numpy()
            out[f"thrusts_{tag}"] = d.thrusts.clone().numpy()
    out["action"] = acts[warm].numpy()
    np.savez_compressed(os.path.join(HERE, f"step_{name}.npz"), **out)


def hover_return(d, acts, gamma=0.99):
    """Discounted hover-style return (reference envs/HoverEnv.py:83-94 reward terms) through H steps."""
    target = th.tensor([[1.0, 0.0, 1.5]], dtype=acts.dtype)
    total = 0.0
    for t in range(acts.shape[0]):
        d.step(acts[t])
        r = 0.1 - (d.position - target).norm(dim=1) / 90 \
            - (d.orientation - th.tensor([1.0, 0, 0, 0], dtype=acts.dtype)).norm(dim=1) * 1e-5 \
            - (d.velocity - 0).norm(dim=1) * 0.002 - (d.angular_velocity - 0).norm(dim=1) * 0.002
        total = total + (gamma ** t) * r
    return total.mean()


def make_grad(name, kw, n=16, H=8):
    out = {}
    init = random_flight_state(n, seed=31, spread=0.5, dtype=th.float64)
    acts0 = actions(H, n, 33, "uniform") * 0.7
    out["actions"] = acts0.numpy()
    for k, x in zip(("pos", "quat", "vel", "rate"), init[:4]):
        out["init_" + k] = x.numpy()
    for dtype, tag in dtypes_for(name):
        with default_dtype(dtype):
            d = make_reference_dynamics(n, dtype=dtype, **kw)
            leaves = [x.to(dtype).clone().requires_grad_(True) for x in init[:4]]
            acts = acts0.to(dtype).clone().requires_grad_(True)
            # reference reset aliases its inputs and the integrator then updates them in place (SURVEY.md
            # C5): hand it non-leaf copies so autograd stays legal
            d.reset(pos=leaves[0] * 1, ori=leaves[1] * 1, vel=leaves[2] * 1, ori_vel=leaves[3] * 1)
            loss = -hover_return(d, acts)
            grads = th.autograd.grad(loss, [acts] + leaves)
            out[f"loss_{tag}"] = np.array(loss.item())
            for k, g in zip(("actions", "pos", "quat", "vel", "rate"), grads):
                out[f"grad_{k}_{tag}"] = g.numpy()
    np.savez_compressed(os.path.join(HERE, f"grad_{name}.npz"), **out)


def main():
    if not reference_available():
        raise SystemExit("reference tree not found; golden vectors can only be regenerated where it is mounted")
    th.manual_seed(0)
    only = sys.argv[1:]                       # optional: names of the configs to (re)generate
    if not only:
        make_kat()
    for name, kw in CONFIGS.items():
        if only and name not in only:
            continue
        make_traj(name, kw)
        make_step(name, kw)
        if name not in FWD_ONLY:
            make_grad(name, kw)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
