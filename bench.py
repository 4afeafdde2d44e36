#!/usr/bin/env python
"""bench.py — agent control-steps/sec of the north-star path (BASELINE.json configs[1]):

    HoverEnv, 65 536 agents per GPU, visual=False, RK4, dt=0.0025, ctrl_dt=0.02 (8 sub-steps), bodyrate actions,
    motor lag on, comm-delay FIFO of 3 steps, float32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  What is measured
  value     env.step throughput with actions and state resident in HBM (tensor mode): exactly K steps inside
            barrier + synchronize, 16 independent env replicas taking turns so that every step's inputs are evicted
            from L2 (inputs larger than L2, no flush kernel in the timed region); device time between two CUDA events
            on the launching stream, max over ranks; the median of five consecutive such brackets (all listed in
            bracket_ms, wall times beside them in bracket_detail_rank0).  Two diagnostics
            ride along: one env with a 256 MiB flush + CUDA-event pair per step (cold_l2_device_value) and one env
            back to back (hot_l2_bracketed_value).
  e2e       the same env driven like an SB3/numpy training loop: actions arrive in (pinned) host memory every step,
            observation / reward / done come back as numpy arrays — host<->device copies inside the timed region.
  roofline  the dominant kernel (the fused control step) timed alone, cold L2, against the measured HBM peak.
  cpu_baseline  the real reference (baseline/_ref: unmodified VisFly sources + the runtime patches of
            baseline/REF_PATCHES.md) on this box's host cores, on a bounded sample of the same workload; the oracle port
            (oracle/env_oracle.py) is timed beside it.
  apg / racing / custom_task  sub-objects: a whole BPTT update (BASELINE configs[2]) as one CUDA-graph replay, RacingEnv
            weak + strong scaling (configs[4]), a task env with its own reward code (recorded step / two-launch / generic).
`--impl reference` runs only that CPU arm for K steps and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402

METRIC = "agent control-steps/sec (visual=False)"
UNIT = "agent-steps/s"
AGENTS = 65536
DYN = dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02, ctrl_delay=True, comm_delay=0.06)
WORKLOAD = "HoverEnv 65536 agents/GPU visual=False RK4 dt=0.0025 ctrl_dt=0.02 bodyrate (BASELINE configs[1])"
ALGO_BYTES_FWD = 176          # 96 B read (20 state + 4 action floats) + 80 B written per agent-step
MOVED_BYTES_FWD = 176 + 52 + 39   # + the (n,13) observation + reward/done/episode record/env status of the tail
FLOP_PER_AGENT_STEP = 4100    # SURVEY.md §8(d) lean count, RK4 x 8 sub-steps
FP32_PEAK_TFLOPS = 74.0       # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal)
L2_FLUSH_BYTES = 256 << 20
REPLICAS = 16                 # env copies rotated in the timed loop: 16 x ~17 MB per step > 126 MB L2
BRACKETS = 5                  # consecutive K-step brackets; value = the median one
HOT_PREROLL = 600             # untimed back-to-back steps before the bracketed loop (host clock ramp, see run_ours)
E2E_PREROLL = 400             # untimed numpy-mode steps before the e2e brackets: the zero-copy PCIe write rate of the
                              # first ~20 ms after an idle period is 20-25 % below its steady state (measured: brackets
                              # of 200 steps took 24.5 / 20.8 / 19.9 ms in a row)


# ---------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML; nvidia-smi as fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._ready = threading.Event()          # set after the first sample: NVML import / init is over by then
        self._thread = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            t0 = time.perf_counter()
            self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            self._ready.set()
            mask = get_reasons(h)
            for bit, name in names.items():
                if mask & bit:
                    self.reasons.add(name)
            # NVML answers under a driver-wide lock: keep the sampler's duty cycle below ~4 % whatever a query costs
            # on this box (50 samples/s where a query takes < 0.8 ms, never fewer than 4 per second)
            self.period = min(max(0.02, 25.0 * (time.perf_counter() - t0)), 0.25)
            self._stop.wait(self.period)

    def _smi_loop(self):
        import subprocess
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout.strip().split(",")
            if len(out) >= 6:
                self.samples.append(int(out[0]))
                self._ready.set()
                self.max_mhz = int(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(name)
            time.sleep(0.1)

    def _loop(self):
        try:
            self._nvml_loop()
        except Exception:  # noqa: BLE001
            try:
                self._smi_loop()
            except Exception:  # noqa: BLE001
                pass

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        self._ready.wait(timeout=5.0)            # keep the sampler's start-up (imports, nvmlInit) out of timed regions
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": sorted(self.reasons)}


def hover_actions(n, count, device, seed=0):
    """smooth-hover law of SURVEY.md §8d config 2: a = [-1/3, 0, 0, 0] + U(-0.1, 0.1)^4  (-1/3 <-> 1 g)."""
    g = th.Generator(device="cpu").manual_seed(seed)
    a = (th.rand(count, n, 4, generator=g) * 2 - 1) * 0.1
    a[..., 0] += -1.0 / 3.0
    return a.to(device) if device != "cpu" else a


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (baseline/_ref copy or the live tree, runtime monkeypatches
# R1-R3 + signature shims only, baseline/REF_PATCHES.md) on the host cores; the oracle port rides along
# ---------------------------------------------------------------------------------------------------------
def base_config(world: int) -> dict:
    """`config` of the JSON line — identical in both arms."""
    return {"workload": WORKLOAD, "agents_per_gpu": AGENTS, "substeps": 8, "actions": "smooth-hover law",
            "l2": "inputs larger than L2 (16 rotating replicas of the 65536-agent env, no flush kernel in the timed region)",
            "parallelism": f"agents sharded over {world} GPU(s), one all_gather of episode returns per rollout"}


def _time_steps(step, acts, steps, warmup):
    for i in range(warmup):
        step(acts[i % len(acts)])
    t0 = time.perf_counter()
    for i in range(steps):
        step(acts[i % len(acts)])
    return time.perf_counter() - t0


def reference_env_run(steps: int, warmup: int, budget_s: float, agents: int = AGENTS, with_port: bool = True):
    """Times the real reference's HoverEnv.step (same dynamics kwargs; C4 shim + RK4 repairs as monkeypatches) on
    the CPU.  The agent count of the sample is halved until `warmup + steps` steps fit the time budget.  Falls back
    to the oracle port (kind "port") only where no reference tree is present."""
    import warnings
    warnings.filterwarnings("ignore")
    from baseline.ref_loader import reference_available, reference_origin
    th.set_num_threads(os.cpu_count() or 1)
    cores = th.get_num_threads()
    kind = "reference" if reference_available() else "port"

    def build(n):
        if kind == "reference":
            from baseline.ref_loader import load_reference_envs
            env = load_reference_envs()["HoverEnv"](num_agent_per_scene=n, visual=False, device="cpu",
                                                    dynamics_kwargs=dict(DYN), max_episode_steps=256,
                                                    tensor_output=True)
        else:
            env = port_env(n)
        env.reset()
        return env

    n = agents
    while True:
        t0 = time.perf_counter()
        env = build(n)
        reset_s = time.perf_counter() - t0
        acts = hover_actions(n, 4, "cpu")
        t0 = time.perf_counter()
        env.step(acts[0])
        probe = time.perf_counter() - t0
        if probe * (steps + warmup) <= budget_s or n <= 1024:
            break
        n //= 2
    dt = _time_steps(env.step, acts, steps, warmup)
    out = {"value": n * steps / dt, "ms_per_step": 1e3 * dt / steps, "agents": n, "cores": cores, "steps": steps,
           "kind": kind, "origin": reference_origin(), "reset_s": reset_s,
           "what": ("VisFly HoverEnv.step, unmodified source + runtime patches R1-R3/C4" if kind == "reference"
                    else "oracle/env_oracle.py OracleEnv(hover).step (op-by-op port)")}
    k = max(1, min(5, steps))
    if kind == "reference":
        from baseline.ref_loader import load_reference_envs, make_reference_dynamics
        dyn = make_reference_dynamics(n, **DYN)
        dyn.reset()
        out["dynamics_only_value"] = n * k / _time_steps(dyn.step, acts, k, 1)
        # the one task env whose signatures match the base class as shipped: no shim at all
        nav = load_reference_envs()["NavigationEnv"](num_agent_per_scene=min(n, 8192), visual=False, device="cpu",
                                                     dynamics_kwargs=dict(DYN), max_episode_steps=256)
        nav.reset()
        nav_acts = hover_actions(min(n, 8192), 4, "cpu")
        out["navigation_env_value"] = min(n, 8192) * k / _time_steps(nav.step, nav_acts, k, 1)
        out["navigation_env_agents"] = min(n, 8192)
    if with_port or kind == "port":
        from oracle.torch_oracle import OracleDynamics
        m = min(n, 16384)
        pacts = hover_actions(m, 4, "cpu")
        penv = port_env(m)
        penv.reset()
        out["port_value"] = m * k / _time_steps(penv.step, pacts, k, 1)
        pdyn = OracleDynamics(m, **{kk: v for kk, v in DYN.items()})
        out["port_dynamics_only_value"] = m * k / _time_steps(pdyn.step, pacts, k, 1)
        out["port_agents"] = m
        if kind == "port":
            out["dynamics_only_value"] = out["port_dynamics_only_value"]
    return out


def port_env(n):
    """oracle/env_oracle.py hover env with a vectorised initial placement (reset is not part of the timed step)."""
    from oracle.env_oracle import OracleEnv

    def generate(indices=None):
        m = n if indices is None else len(indices)
        pos = th.tensor([1., 0., 1.5]) + (th.rand(m, 3) * 2 - 1) * th.tensor([1.0, 1.0, 0.5])
        quat = th.zeros(m, 4)
        quat[:, 0] = 1
        return pos, quat, th.zeros(m, 3), th.zeros(m, 3)
    return OracleEnv("hover", n, dict(DYN), max_episode_steps=256, generate_state=generate)


def reference_dynamics_on_gpu(n, dev, steps=5):
    """north_star's second baseline: the reference's PyTorch dynamics executed on the B200.  The real reference
    `Dynamics.step` (RK4 repairs R1-R3; device hygiene D1-D3 = run under torch.set_default_device, see
    baseline/ref_loader.py) — ~8 200 aten launches per RK4x8 control step; the oracle port with its tensors on the
    device is the fallback where the reference tree is absent or the reference raises on CUDA."""
    acts = hover_actions(n, 4, dev)

    def timed(step):
        with th.no_grad():
            for i in range(2):
                step(acts[i])
            th.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                step(acts[i % 4])
            th.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    out = {}
    try:
        from baseline.ref_loader import make_reference_dynamics, reference_available, reference_on_device
        if reference_available():
            with reference_on_device(dev):
                dyn = make_reference_dynamics(n, device=dev, **DYN)
                dyn.reset()
                dt = timed(dyn.step)
            out = {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "agents": n, "kind": "reference",
                   "what": "VisFly Dynamics.step (no env wrapper), unmodified source + runtime patches R1-R3, "
                           "D1-D3 via torch.set_default_device(cuda)"}
    except Exception as e:  # noqa: BLE001 - a baseline must never take the bench down
        out = {"reference_error": repr(e)[:300]}
    if "value" not in out:
        try:
            from oracle.torch_oracle import OracleDynamics
            dyn = OracleDynamics(n, device=dev, **{k: v for k, v in DYN.items()})
            dt = timed(dyn.step)
            out.update({"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "agents": n, "kind": "port",
                        "what": "OracleDynamics.step (Dynamics.step only, no env wrapper) with all tensors on cuda"})
        except Exception as e:  # noqa: BLE001
            out["error"] = repr(e)[:200]
    return out


def cpu_baseline_record(r):
    sample = (f"{r['what']} on {r['agents']} agents x {r['steps']} steps ({r['ms_per_step']:.0f} ms/step), "
              f"{r['cores']} host threads, reference tree: {r['origin']}")
    rec = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample}
    for k in ("dynamics_only_value", "navigation_env_value", "navigation_env_agents", "port_value",
              "port_dynamics_only_value", "port_agents", "reset_s"):
        if k in r:
            rec[k] = r[k]
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = reference_env_run(args.steps, args.warmup, budget_s=150.0, with_port=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.gpus),
        "cpu_baseline": cpu_baseline_record(r),
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def timed_steps(step_fn, steps, flush, stream):
    """One CUDA-event pair per step on the launching stream, L2 flushed before each timed step."""
    starts = [th.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [th.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        starts[i].record(stream)
        step_fn(i)
        ends[i].record(stream)
    th.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in zip(starts, ends)]       # ms


def rotating_brackets(envs, act_list, K, preroll, brackets, stream, barrier, collective=None):
    """The contract's bracket over a set of independent env replicas that take turns (so that every step finds its
    inputs evicted from L2 while launches stay back to back): after `preroll` untimed steps, `brackets` consecutive
    brackets of exactly K steps (+ the rollout's one collective, if given), each opened behind barrier + synchronize
    and closed by synchronising on the bracket's last CUDA event — no barrier, no other collective inside the
    timed region.  `collective`: None, or a FusedReturnsGather — armed before the bracket's last step (whose launch then
    scatters the per-agent returns to every rank) and finished (cross-rank barrier) right after it.
    Returns per bracket the device time (CUDA events on the launching stream) and the wall time."""
    R, pool = len(envs), len(act_list)
    for i in range(preroll):
        envs[i % R].step(act_list[i % pool])
    dev_ms, wall_ms = [], []
    for _ in range(brackets):
        barrier()
        r0, r1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        r0.record(stream)
        for i in range(K - 1):
            envs[i % R].step(act_list[i % pool])
        last = envs[(K - 1) % R]
        if collective is not None:
            collective.arm(last)
        last.step(act_list[(K - 1) % pool])
        if collective is not None:
            collective.finish(last)
        r1.record(stream)
        r1.synchronize()
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(r0.elapsed_time(r1))
    return dev_ms, wall_ms


def median(xs):
    return sorted(xs)[len(xs) // 2]


def max_over_ranks(values, dev, world):
    import torch.distributed as dist
    t = th.tensor(list(values), device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def racing_leg(n_weak, dev, rank, world, K, W, stream, barrier):
    """BASELINE configs[4]: RacingEnv semantics (gates, radius 0.3, hover-style reward + 20 per gate, 16-wide
    observation; reference envs/RacingEnv.py:87-98,142-148,203-215,250-267), RK4 x8, agents sharded over the GPUs:
    weak scaling (65 536 agents per GPU) and strong scaling (524 288 agents in total, rank r owns shard_range(r))."""
    from visfly_b200.distributed import FusedReturnsGather, shard_range
    from visfly_b200.envs import RacingEnv2
    out = {"workload": "RacingEnv2 visual=False RK4 dt=0.0025 ctrl_dt=0.02 bodyrate (BASELINE configs[4])"}
    total_strong = 524288
    lo, hi = shard_range(total_strong, rank, world)
    for name, n, n_total, first in (("weak", n_weak, n_weak * world, rank * n_weak),
                                    ("strong", hi - lo, total_strong, lo)):
        replicas = max(2, -(-16 * 65536 // n))          # keep >= ~280 MB of per-step traffic in rotation (> L2)
        # every rank uses the SAME seed and declares its shard: placements and in-kernel restarts are the rows the
        # whole n_total-agent batch would draw (tests/test_gpu_multirank.py), whatever the number of GPUs
        envs = [RacingEnv2(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN),
                           seed=42 + 1000 * j, max_episode_steps=256, tensor_output=True,
                           shard=None if world == 1 else (first, n_total))
                for j in range(replicas)]
        for e in envs:
            e.reset()
        pool = 4
        act_list = list(hover_actions(n, pool, dev, seed=rank).unbind(0))
        gather = FusedReturnsGather(n, n_total, rank, world, dev)
        dev_ms, wall_ms = rotating_brackets(envs, act_list, K, max(W * replicas, HOT_PREROLL), BRACKETS, stream, barrier,
                                            collective=gather)
        ms = max_over_ranks(dev_ms, dev, world)
        t = median(ms)
        out[name] = {"value": n_total * K / (t * 1e-3), "unit": UNIT, "agents_total": n_total, "agents_this_gpu": n,
                     "ms_per_step": t / K, "bracket_ms": ms, "replicas": replicas,
                     "fused": bool(envs[0]._fused is not None and envs[0]._fused.active),
                     "collective": "fused into the last step (peer-memory stores + barrier)" if gather.fused else
                                   f"NCCL all_gather_into_tensor ({gather.why_not})",
                     "scaling": name}
        del envs, gather
        th.cuda.empty_cache()
    return out


def custom_task_leg(n, dev, K, W, stream, barrier):
    """A task env as a user would write it — HoverEnv's reward re-stated in a subclass, so that the env no longer
    qualifies for the one-kernel step — as one CUDA-graph replay per step (env.capture_task_step), on the eager
    two-launch path (control step kernel + the task's tensor ops + vf_env_finish) and on the generic tensor-op path
    (what such envs ran on in round 1)."""
    from visfly_b200.envs import HoverEnv

    class UserHover(HoverEnv):
        def get_reward(self, predicted_obs=None):
            return (0.1 - (self.position - self.target).norm(dim=1) * (0.1 / 9)
                    - (self.orientation - self._unit_quat).norm(dim=1) * 1e-5
                    - self.velocity.norm(dim=1) * 0.002 - self.angular_velocity.norm(dim=1) * 0.002)

    out = {"workload": "HoverEnv subclass with its own get_reward (tensor code), 65536 agents, RK4 x8"}
    acts = list(hover_actions(n, 4, dev, seed=3).unbind(0))
    for name, split, capture in (("recorded_step", True, True), ("recorded_step_copied_outputs", True, "copy"),
                                 ("two_launch_path", True, False), ("generic_path", False, False)):
        env = UserHover(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=77,
                        max_episode_steps=256, tensor_output=True)
        env.use_fused_step = split
        env.capture_task_step = capture
        env.reset()
        for i in range(max(W, 30)):
            env.step(acts[i % 4])
        steps = max(K, 50)
        barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            env.step(acts[i % 4])
        e1.record(stream)
        e1.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        out[name] = {"value": n * steps / (ms * 1e-3), "unit": UNIT, "us_per_step": ms * 1e3 / steps,
                     "active": bool(env._split is not None and env._split.active),
                     "graph_replays": 0 if env._task_graph is None else env._task_graph.replays}
    return out


def run_ours(args):
    import torch.distributed as dist
    from visfly_b200 import _lib
    from visfly_b200.distributed import FusedReturnsGather
    from visfly_b200.envs import HoverEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    # keep the rank's host thread and its page-locked buffers on the CPU set next to its GPU (NVML's affinity mask):
    # env.step costs ~10 us of host time against ~12 us of device time, so where the launching thread runs shows
    # directly (measured on one box: 16.7 us per step unpinned at N=1, 13.8 us pinned at N=4)
    from visfly_b200.distributed import bind_host_to_gpu
    numa_cpus = bind_host_to_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K, W = args.agents, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    flush = th.empty(L2_FLUSH_BYTES, dtype=th.uint8, device=dev)
    stream = th.cuda.current_stream(dev)
    pool = 16
    acts = hover_actions(n, pool, dev, seed=rank)

    # ---- value: tensor mode, everything resident in HBM ----------------------------------------------
    env = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=42 + rank,
                   max_episode_steps=256, tensor_output=True)
    env.reset()
    act_list = list(acts.unbind(0))

    def env_step(i):
        env.step(act_list[i % pool])

    for i in range(W):
        env_step(i)
    # the one collective of the path: episode returns of all shards, once per rollout (a no-op at world 1); buffers
    # and shard sizes are fixed ahead of the rollout, the call is one asynchronous all_gather_into_tensor
    # fused: the rollout's last step stores the returns into every rank's buffer through peer-mapped memory and a
    # cross-rank barrier follows (FusedReturnsGather); NCCL all_gather_into_tensor where P2P mappings are unavailable
    gather = FusedReturnsGather(n, n * world, rank, world, dev)
    for _ in range(3):                    # communicator / barrier set-up happens here, not in a timed bracket
        gather.arm(env)
        env_step(0)
        gather.finish(env)
        gather.fallback(env._rewards)
    # long-lived objects (modules, the envs) leave the garbage collector's working set: a full collection walking
    # them costs ~1 ms, which is 50 env steps on this path
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    clk = ClockSampler(local)
    clk.__enter__()                         # sampled across every timed region below (value, roofline, e2e, apg)
    try:
        # (A) cold L2: one CUDA-event pair per step, 256 MiB flush before every timed step
        per_step = timed_steps(env_step, K, flush, stream)
        barrier()
        # (B) hot L2: one env back to back (its 12 MB working set stays in L2), same bracket as (C).  The host thread
        #     slept in the barrier above while the GPU drained the flush-heavy cold pass and its core needs a few ms of
        #     work to clock back up (measured: 19.8 -> 17.1 -> 14.5 us/step over three consecutive 200-step loops),
        #     hence the untimed pre-roll
        hot_dev, hot_wall = rotating_brackets([env], act_list, K, max(W, HOT_PREROLL), 1, stream, barrier,
                                              collective=gather)
        # (C) THE reported value: the contract's bracket with inputs larger than L2 and no flush kernel inside it —
        #     REPLICAS independent copies of the 65 536-agent env take turns, so every step finds its state, actions
        #     and env status evicted (REPLICAS x ~17 MB per step >> 126 MB L2) while launches stay back to back.
        #     BRACKETS consecutive brackets of exactly K steps + the rollout's collective; per bracket the max over
        #     ranks of the device time between the bracket's two CUDA events (the contract's clock; the wall time, which
        #     adds the wake-up of the closing synchronize, is listed beside it); the value is the MEDIAN bracket.
        envs = [env] + [HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN),
                                 seed=42 + rank + 1000 * j, max_episode_steps=256, tensor_output=True)
                        for j in range(1, REPLICAS)]
        for e in envs[1:]:
            e.reset()
        preroll = max(W * REPLICAS, HOT_PREROLL)
        bracket_dev_ms, bracket_wall_ms = rotating_brackets(envs, act_list, K, preroll, BRACKETS, stream, barrier,
                                                            collective=gather)
        # (D) the same loop WITHOUT the collective: the step kernel's own launch-to-launch time (roofline.kernel_us)
        nocoll_dev_ms, _ = rotating_brackets(envs, act_list, K, 50, BRACKETS, stream, barrier)
        # (E) the collective alone, back to back
        c0, c1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        barrier()
        c0.record(stream)
        for _ in range(10):
            gather.fallback(env._rewards)
        c1.record(stream)
        c1.synchronize()
        nccl_collective_us = c0.elapsed_time(c1) * 1e3 / 10 if world > 1 else 0.0
        # the fused form: what the bracket with the collective costs beyond the collective-free one
        collective_us = max(0.0, (median(bracket_dev_ms) - median(nocoll_dev_ms)) * 1e3) if world > 1 else 0.0
        del envs
    except BaseException:
        clk.__exit__(None, None, None)
        raise
    bracket_ms = max_over_ranks(bracket_dev_ms, dev, world)       # device time (CUDA events on the launching stream)
    per_rank = None
    if world > 1:                       # which rank the others wait for: every rank's median collective-free bracket
        import torch.distributed as dist
        mine = th.tensor([median(nocoll_dev_ms), median(bracket_dev_ms)], device=dev, dtype=th.float64)
        allr = [th.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"collective_free_device_ms": [float(x[0]) for x in allr], "device_ms": [float(x[1]) for x in allr]}
    cold_ms, hot_ms = max_over_ranks([sum(per_step), max(hot_dev[0], hot_wall[0])], dev, world)
    total_ms = median(bracket_ms)
    cold_value = world * n * K / (cold_ms * 1e-3)
    hot_value = world * n * K / (hot_ms * 1e-3)
    value = world * n * K / (total_ms * 1e-3)

    # ---- Dynamics.step alone through the drop-in class (SURVEY.md §8d reports both) -------------------------
    from visfly_b200.dynamics import Dynamics
    dyn_only = Dynamics(num=n, device=dev, **DYN)
    with th.no_grad():
        for i in range(W):
            dyn_only.step(act_list[i % pool])
        barrier()
        d0, d1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        d0.record(stream)
        for i in range(K):
            dyn_only.step(act_list[i % pool])
        d1.record(stream)
        th.cuda.synchronize()
        dyn_ms = max(d0.elapsed_time(d1), (time.perf_counter() - t0) * 1e3)
    dynamics_step_value = n * K / (dyn_ms * 1e-3)

    # ---- roofline: the fused control-step kernel alone, cold L2 -----------------------------------------
    dynm = env.envs.dynamics
    cfg = dynm._cfg
    st_in = dynm.packed_state.detach().clone()
    st_out, obs_out = th.empty_like(st_in), th.empty((n, 13), device=dev)

    # the kernel env.step launches: the fused env step (control step + wrapper tail) on a private copy of the
    # per-agent env status, so that timing it does not disturb the env
    fz = env._fused
    status, status_o = fz.status.clone(), th.empty_like(fz.status)
    rew_o, done_o = th.empty(n, device=dev), th.empty(n, dtype=th.bool, device=dev)
    rec_o = th.empty((n, 4), device=dev)

    def kernel_only(i):
        _lib.env_step_fwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0, 10_000 + i,
                          st_in, acts[i % pool], None, status, st_out, status_o, obs_out, rew_o, done_o, rec_o, None)

    for i in range(5):
        kernel_only(i)
    k_ms = timed_steps(kernel_only, max(K, 50), flush, stream)
    k_iso = float(np.mean(k_ms)) * 1e-3
    # the dominant kernel's average launch duration over the timed region: the median bracket holds exactly K launches
    # of it (one per env.step, nothing else on the stream), CUDA events on the launching stream at both ends — an
    # upper bound of the kernel time (inter-launch gaps included), inputs cold (16 rotating replicas).  k_iso is the
    # same kernel launched alone behind a 256 MiB L2 flush with one event pair per launch (event overhead included).
    k_avg = median(nocoll_dev_ms) * 1e-3 / K
    peak, peak_src = measured_peaks()
    achieved = ALGO_BYTES_FWD * n / k_avg / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("vf_env_step_fwd_kernel_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "vf_env_step_fwd_kernel<RK4,BODYRATE,LAG>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture "
                                  "named in profiles/roofline_traffic.json (not re-measured in this run)",
                "peak_source": peak_src,
                "kernel_us": k_avg * 1e6, "kernel_us_isolated_cold_event_pair": k_iso * 1e6,
                "timing": "device time of the median collective-free K-step bracket / K launches (CUDA events on the "
                          "launching stream, cold inputs, one launch of this kernel per step, nothing else on the stream)",
                "algorithmic_bytes_per_launch": ALGO_BYTES_FWD * n,
                "bytes_moved_per_launch": MOVED_BYTES_FWD * n,
                "fp32": {"achieved_tflops": FLOP_PER_AGENT_STEP * n / k_avg / 1e12, "peak_tflops": FP32_PEAK_TFLOPS,
                         "frac": FLOP_PER_AGENT_STEP * n / k_avg / 1e12 / FP32_PEAK_TFLOPS,
                         "note": "RK4 x8 is fp32-pipe/latency bound (23 FLOP/B, ridge ~11.5), SURVEY.md §8d"}}

    # the same kernel where the chip is full: 4 194 304 agents (1.1 GB of traffic per launch, far beyond L2) — the
    # asymptote the 65 536-agent workload cannot reach because it is shorter than a launch ramp
    try:
        big = 1 << 22
        b_in = st_in.repeat(1, big // n, 1).contiguous()
        b_out, b_obs = th.empty_like(b_in), th.empty((big, 13), device=dev)
        b_act = acts[0].repeat(big // n, 1).contiguous()
        b_status = status.repeat(big // n, 1).contiguous()
        b_status_o = th.empty_like(b_status)
        b_rew, b_done = th.empty(big, device=dev), th.empty(big, dtype=th.bool, device=dev)
        b_rec = th.empty((big, 4), device=dev)

        def kernel_big(i):
            _lib.env_step_fwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0,
                              30_000 + i, b_in, b_act, None, b_status, b_out, b_status_o, b_obs, b_rew, b_done, b_rec,
                              None)

        for i in range(3):
            kernel_big(i)
        kb = float(np.mean(timed_steps(kernel_big, 10, None, stream))) * 1e-3
        roofline["asymptote_4194304_agents"] = {
            "kernel_us": kb * 1e6, "achieved": ALGO_BYTES_FWD * big / kb / 1e9, "frac": ALGO_BYTES_FWD * big / kb / 1e9 / peak,
            "moved_gbs": MOVED_BYTES_FWD * big / kb / 1e9, "fp32_tflops": FLOP_PER_AGENT_STEP * big / kb / 1e12,
            "fp32_frac": FLOP_PER_AGENT_STEP * big / kb / 1e12 / FP32_PEAK_TFLOPS, "agent_steps_per_s": big / kb}
        del b_in, b_out, b_obs, b_act, b_status, b_status_o, b_rew, b_done, b_rec
    except Exception as e:  # noqa: BLE001 - a diagnostic must never take the bench down
        roofline["asymptote_4194304_agents"] = {"error": repr(e)[:200]}

    # the same kernel with the page-locked host mirror as a second destination (what the e2e step launches)
    from visfly_b200.params import VfEnvMirror
    m_obs, m_rew = th.empty((n, 13), pin_memory=True), th.empty(n, pin_memory=True)
    m_done = th.empty(n, dtype=th.int32, pin_memory=True)
    mirror = VfEnvMirror(m_obs.data_ptr(), m_rew.data_ptr(), m_done.data_ptr(), None, None, 0)

    def kernel_mirror(i):
        _lib.env_step_fwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0, 20_000 + i,
                          st_in, acts[i % pool], None, status, st_out, status_o, obs_out, rew_o, done_o, rec_o, None,
                          mirror)

    for i in range(5):
        kernel_mirror(i)
    km_avg = float(np.mean(timed_steps(kernel_mirror, 50, flush, stream))) * 1e-3
    # ... and with the completion word the host spins on (fence.sys per thread + one atomic per block at the end)
    m_flag = th.zeros(16, dtype=th.int32, pin_memory=True)
    m_counter = th.zeros(1, dtype=th.int32, device=dev)
    mirror_f = VfEnvMirror(m_obs.data_ptr(), m_rew.data_ptr(), m_done.data_ptr(), m_flag.data_ptr(),
                           m_counter.data_ptr(), 1)

    def kernel_mirror_flag(i):
        _lib.env_step_fwd(cfg.params, fz.spec, cfg.substeps, cfg.integrator, cfg.action_type, cfg.flags, 0, 20_000 + i,
                          st_in, acts[i % pool], None, status, st_out, status_o, obs_out, rew_o, done_o, rec_o, None,
                          mirror_f)

    for i in range(5):
        kernel_mirror_flag(i)
    kmf_avg = float(np.mean(timed_steps(kernel_mirror_flag, 50, flush, stream))) * 1e-3
    # back to back (what the run-ahead host loop keeps queued): K launches between one event pair
    b0, b1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    b0.record(stream)
    for i in range(50):
        kernel_mirror_flag(i)
    b1.record(stream)
    b1.synchronize()
    kmf_b2b = b0.elapsed_time(b1) * 1e-3 / 50

    # ---- e2e: numpy in / numpy out through the public env API -------------------------------------------
    env_np = HoverEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=142 + rank,
                      max_episode_steps=256, tensor_output=False)
    env_np.reset()
    host_acts = [th.empty((n, 4), pin_memory=True).copy_(acts[i].cpu()).numpy() for i in range(pool)]
    sink = [0.0, 0.0, 0]

    def e2e_step(i):
        obs, reward, done, info = env_np.step(host_acts[i % pool])
        sink[0] += float(reward[0]) + float(reward[-1])        # the host really reads what came back
        sink[1] += float(obs["state"][0, 2]) + float(obs["state"][-1, 2])
        sink[2] += int(done[0]) + int(done[-1])

    for i in range(max(W, E2E_PREROLL)):
        e2e_step(i)
    # three consecutive K-step brackets (barrier + synchronize on both sides, max over ranks), the median is reported:
    # this loop is one host thread ping-ponging with the GPU, and a single bracket has been seen 20 % off
    e2e_brackets, e2e_dev = [], []
    for _ in range(3):
        barrier()
        x0, x1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        x0.record(stream)
        for i in range(K):
            e2e_step(i)
        x1.record(stream)
        th.cuda.synchronize()
        e2e_brackets.append(time.perf_counter() - t0)
        e2e_dev.append(x0.elapsed_time(x1))
    e2e_all = th.tensor(e2e_brackets, device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(e2e_all, op=dist.ReduceOp.MAX)
    e2e_s = e2e_all.sort().values[1]
    e2e = {"value": world * n * K / float(e2e_s), "unit": UNIT, "h2d_bytes_per_step": n * 16,
           "d2h_bytes_per_step": n * (13 * 4 + 4 + 4), "ms_per_step": 1e3 * float(e2e_s) / K,
           "bracket_ms": [1e3 * float(x) for x in e2e_all],
           "api": "HoverEnv(tensor_output=False).step(numpy actions) -> numpy obs/reward/done",
           "kernel_with_host_mirror_us": km_avg * 1e6, "kernel_with_mirror_and_flag_us": kmf_avg * 1e6,
           "kernel_with_mirror_and_flag_back_to_back_us": kmf_b2b * 1e6, "bracket_device_ms_rank0": e2e_dev,
           "launches_in_bracket": "K handed out + 1 running ahead (the engine launches step t+1 before it hands out "
                                  "step t; the bracket's closing synchronize waits for it)",
           "pcie_write_gbs_of_that_kernel": n * (13 * 4 + 4 + 4) / km_avg / 1e9,
           "transfers": "H2D: async DMA of the page-locked action array; D2H: the kernel stores obs/reward/done "
                        "straight into page-locked host memory (zero-copy over PCIe); its last thread block raises a "
                        "page-locked completion word the host spins on (no stream synchronisation per step)"}

    # ---- config[2]: APG-style analytic policy gradient through NavigationEnv (requires_grad=True) -----------------
    apg = None
    if not args.no_apg:
        try:
            apg = apg_benchmark(n, dev, rank, world, barrier, graph=True)      # whole update replayed as one CUDA graph
        except Exception as e:  # noqa: BLE001 - fall back to the eager update rather than lose the leg
            apg = {"graph_error": repr(e)[:300]}
        eager = apg_benchmark(n, dev, rank, world, barrier, graph=False)
        if "value" in apg:
            apg["eager"] = {k: eager[k] for k in ("value", "ms_per_update")}
        else:
            eager.update(apg)
            apg = eager

    racing = None if args.no_racing else racing_leg(n, dev, rank, world, K, W, stream, barrier)
    custom = None if args.no_racing else custom_task_leg(n, dev, K, W, stream, barrier)
    clk.__exit__(None, None, None)

    ref_gpu = None
    if rank == 0 and not args.no_cpu_baseline:
        ref_gpu = reference_dynamics_on_gpu(n, dev)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_record(reference_env_run(steps=10, warmup=1, budget_s=20.0, agents=16384))
        cfg_line = base_config(world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            # untimed env.step calls that really precede the first timed bracket (the requested W is kept beside it)
            "warmup": W + max(W, HOT_PREROLL) + preroll, "warmup_requested": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg_line,
            "measurement": {
                "value": f"median of {BRACKETS} consecutive brackets of exactly K env.step calls + the rollout's one "
                         "all_gather; each bracket opens behind barrier + synchronize and closes by synchronising on its "
                         "last CUDA event (no barrier / extra collective inside); per bracket max over ranks of "
                         "the device time between the two CUDA events on the launching stream; bracket_ms lists them all, bracket_detail_rank0 has the wall times beside them (they add the wake-up of the closing synchronize, ~15 us)",
                "cold_l2_device_value": "one env, 256 MiB flush + one CUDA-event pair per step",
                "hot_l2_bracketed_value": "one env back to back (12 MB working set resident in L2)",
                "host_affinity": None if numa_cpus is None else f"rank 0 pinned to {len(numa_cpus)} CPUs next to its GPU"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": K * env_launches_per_step(env), "apg": apg,
            "racing": racing, "custom_task": custom, "roofline": roofline, "cpu_baseline": cpu, "reference_dynamics_on_gpu": ref_gpu,
            "rollout_collective_us": collective_us,
            "rollout_collective": {"fused_into_last_step": gather.fused, "why_not": gather.why_not,
                                   "us_added_to_the_bracket": collective_us,
                                   "nccl_all_gather_into_tensor_us": nccl_collective_us},
            "dynamics_step_value_per_gpu": dynamics_step_value,
            "cold_l2_device_value": cold_value, "hot_l2_bracketed_value": hot_value, "kernel_only_value": n / k_avg,
            "bracket_ms": bracket_ms,
            "bracket_detail_rank0": {"device_ms": bracket_dev_ms, "wall_ms": bracket_wall_ms,
                                     "collective_free_device_ms": nocoll_dev_ms},
            "bracket_median_per_rank": per_rank,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def apg_benchmark(n, dev, rank, world, barrier, H=32, rollouts=5, gamma=0.99, graph=True):
    """BPTT updates (reference utils/algorithms/BPTT.py:107-134) through visfly_b200.algorithms.BPTT: H steps with grad
    through NavigationEnv, loss = -sum discount_t r_t, backward, gradient all-reduce over the ranks, Adam step,
    env.detach().  Forward = one fused env-step launch per step, backward = one adjoint launch per step; the policy
    MLP (16-64-64-4, tanh) is plain torch."""
    import torch.distributed as dist
    from visfly_b200.algorithms import BPTT
    from visfly_b200.envs import NavigationEnv
    env = NavigationEnv(num_agent_per_scene=n, visual=False, device=dev, dynamics_kwargs=dict(DYN), seed=7 + rank,
                        requires_grad=True, max_episode_steps=256,
                        random_kwargs={"state_generator": {"class": "Uniform", "kwargs": [
                            {"position": {"mean": [2., 0., 1.5], "half": [1.0, 1.0, 0.5]}}]}})
    algo = BPTT(env, horizon=H, gamma=gamma, learning_rate=1e-3, policy_kwargs=dict(net_arch=[64, 64]), seed=0,
                make_eval_env=False, dump_step=1 << 62, cuda_graph=graph)
    algo.learn(total_timesteps=(4 if graph else 2) * n * H)   # warm-up (graph: eager update, side-stream warm-up, capture)
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    algo.learn(total_timesteps=rollouts * n * H)
    e1.record()
    barrier()
    ms = th.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    finite = all(bool(th.isfinite(p_).all()) for p_ in algo.actor.parameters())
    return {"workload": "NavigationEnv 65536 agents/GPU requires_grad=True RK4 (BASELINE configs[2]), BPTT updates of "
                        "H=32 (visfly_b200.algorithms.BPTT, MLP 16-64-64-4 policy, grad all-reduce over ranks)",
            "value": world * n * H * rollouts / (float(ms) * 1e-3), "unit": UNIT + " (fwd+bwd+update)",
            "ms_per_update": float(ms) / rollouts, "horizon": H, "weights_finite": finite,
            "fused": bool(env._fused is not None and env._fused.active),
            "cuda_graph": bool(graph and algo._graph is not None)}


def env_launches_per_step(env) -> int:
    """Kernels of OUR library launched by one env.step (the torch elementwise glue is not counted)."""
    return int(getattr(env, "native_launches_per_step", 1))


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (e.g. "NCCL version ..." on the first
    collective), so fd 1 is pointed at stderr for the duration of the run and the JSON line goes to the saved fd."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=AGENTS, help="agents per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-apg", action="store_true", help="skip the BASELINE configs[2] leg")
    ap.add_argument("--no-racing", action="store_true", help="skip the BASELINE configs[4] leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        if not th.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
