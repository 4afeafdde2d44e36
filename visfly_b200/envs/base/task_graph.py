"""One env step of a task env with caller-written reward / success / failure code as ONE CUDA-graph replay.

The two-launch path (``DroneGymEnvsBase._step_split``: ``vf_step_fwd`` -> the task's tensor ops -> ``vf_env_finish``)
is host-bound: the task's own code is a dozen or two small tensor kernels, each costing ~10 us of Python + launch
time against ~2 us on the device.  ``TaskStepGraph`` records that whole sequence once and replays it per step.

What makes one step replayable:
* everything a step carries over to the next one — packed state, env status records, the comm-delay FIFO, and any
  tensor attribute the task code itself re-binds during a step (``self.foo = where(done, 0, self.foo)``) — lives in a
  fixed tensor the graph reads first and overwrites last (found by comparing the attribute tables of the env objects
  before and after the recorded step);
* the in-kernel restart sampler takes its step number from a device word the graph advances (``VfEnvSpec`` seed /
  ``step_base``), so replays draw the same restarts an eager run would (bitwise: tests/test_gpu_split_env.py);
* the action is copied into the graph's input buffer (one launch).

Limits (the env falls back to the eager two-launch path where it can tell, the rest is the caller's contract and is
stated in INTEGRATION.md): task code must be capturable (no ``.item()``, no host branches on device data), Python
scalars / lists it mutates per step are not seen by replays, and the tensors a step returns are the graph's own
buffers — overwritten by the next step — unless ``env.capture_task_step == "copy"``."""
from __future__ import annotations

import warnings

import torch as th

from ...type import TensorDict

# attributes a step (re)produces from scratch: never carried from one step into the next
# (fz.finish_out / dyn._fifo_ring are switches of the recording itself)
_OUTPUTS = {
    "env": ("_reward", "_done", "_success", "_failure", "_obs_tensors", "_observations", "_action", "_info"),
    "envs": ("_collision_point", "_is_out_bounds", "_collision_vector", "_collision_dis", "_is_collision"),
    "dyn": ("_obs", "_obs_t", "_ext", "_fresh", "_prev", "_thrusts_given"),
    "fz": ("record", "gate_obs", "sc_mid"),
}

class CaptureUnsupported(RuntimeError):
    pass


class TaskStepGraph:
    def __init__(self, env, action: th.Tensor):
        from .fused import RecordInfo
        self._RecordInfo = RecordInfo
        fz, envs = env._split, env.envs
        dyn = envs.dynamics
        self.env, self.fz, self.envs, self.dyn = env, fz, envs, dyn
        dev = env.device
        if fz.t_off is not None or dyn._wind_fn is not None:
            raise CaptureUnsupported("per-agent time offsets / wind functions")
        if fz.step_base is None:            # restart sampler: step number = host counter (frozen) + this device word
            fz.step_base = th.zeros((1,), dtype=th.int64, device=dev)
            fz._stepper = None
        self.g_step, self.spec_gen = fz.global_step, fz.spec_gen
        self.a_in = th.empty((env.num_agent, 4), dtype=th.float32, device=dev)
        self.a_in.copy_(action)
        owners = {"env": env, "envs": envs, "dyn": dyn, "fz": fz}
        saved = {k: dict(o.__dict__) for k, o in owners.items()}
        fifo = list(dyn._pre_action)
        graph = th.cuda.CUDAGraph()
        try:
            env._action, env._action_owned = self.a_in, False
            # recorded steps keep every carried tensor at its address: the FIFO is shifted in place by the control
            # step's launch, vf_env_finish writes the state / status records it read from
            state_holder, status_holder = dyn._state, fz.status
            dyn._fifo_ring, fz.finish_out = True, (state_holder, status_holder)
            with th.cuda.graph(graph):
                obs, reward, done, info = env._step_split()
                if dyn._state is not state_holder or fz.status is not status_holder:
                    raise CaptureUnsupported("the step did not write its state / status records in place")
                # written in place by the recorded launches: no copy, but replays must find them where they were
                carried = [(dyn, "_state", state_holder), (fz, "status", status_holder)]
                for name, o in owners.items():
                    for k, old in saved[name].items():
                        new = o.__dict__.get(k)
                        if new is old or k in _OUTPUTS[name] or not isinstance(old, th.Tensor) or not old.is_cuda:
                            continue
                        if not isinstance(new, th.Tensor) or new.shape != old.shape or new.dtype != old.dtype:
                            raise CaptureUnsupported(f"attribute {k!r} changed type or shape during the step")
                        old.copy_(new)
                        carried.append((o, k, old))
                if len(dyn._pre_action) != len(fifo):
                    raise CaptureUnsupported("comm-delay FIFO changed its depth during the step")
                for old, new in zip(fifo, dyn._pre_action):
                    if new is not old:
                        old.copy_(new)
                fz.step_base.add_(1)
        except BaseException:
            for k, o in owners.items():     # nothing ran: the Python side goes back to where it was
                o.__dict__.clear()
                o.__dict__.update(saved[k])
            raise
        finally:
            dyn._fifo_ring, fz.finish_out = False, None
        # the recording computed nothing; what a replay leaves behind is re-installed after every replay
        self.produced = [(o, k, o.__dict__[k]) for name, o in owners.items() for k in _OUTPUTS[name]
                         if k in o.__dict__ and k not in ("_info", "sc_mid")]
        for o, k, old in carried:
            o.__dict__[k] = old
        dyn._pre_action = list(fifo)
        dyn._n_steps, fz.global_step = saved["dyn"]["_n_steps"], self.g_step
        self.carried, self.fifo, self.graph = carried, fifo, graph
        self.obs, self.reward, self.done = obs, reward, done
        self.record, self.pre_obs = info._record, info._term_raw
        self.replays = 0

    def matches(self, action: th.Tensor) -> bool:
        """False => the recording no longer fits the env (spec rebuilt, FIFO depth changed, a carried attribute was
        deleted or re-bound to something of another shape): the caller records a new one."""
        if not (self.fz.spec_gen == self.spec_gen and action.shape == self.a_in.shape
                and len(self.dyn._pre_action) == len(self.fifo) and self.fz.t_off is None):
            return False
        for o, k, holder in self.carried:
            cur = o.__dict__.get(k)
            if cur is not holder and not (isinstance(cur, th.Tensor) and cur.shape == holder.shape
                                          and cur.dtype == holder.dtype and cur.device == holder.device):
                return False
        return all(a.shape == h.shape for a, h in zip(self.dyn._pre_action, self.fifo))

    def replay(self, action: th.Tensor, copy: bool):
        env, fz, dyn = self.env, self.fz, self.dyn
        if fz.global_step != self.g_step:              # eager steps since the last replay advanced the host counter
            fz.step_base.add_(fz.global_step - self.g_step)
            fz.global_step = self.g_step
        for o, k, holder in self.carried:              # re-bound from outside (reset(), a copied env state, ...)
            cur = o.__dict__.get(k)
            if cur is not holder:
                holder.copy_(cur)
                o.__dict__[k] = holder
        for j, holder in enumerate(self.fifo):
            if dyn._pre_action[j] is not holder:
                holder.copy_(dyn._pre_action[j])
                dyn._pre_action[j] = holder
        self.a_in.copy_(action)
        self.graph.replay()
        self.replays += 1
        for o, k, v in self.produced:
            o.__dict__[k] = v
        dyn._n_steps += 1
        fz._views = (None, None)            # the status holder keeps its identity: drop the unpacked fields by hand
        self.envs._collision_stale = True
        fz.sc_mid = None
        obs, reward, done, record, pre_obs = self.obs, self.reward, self.done, self.record, self.pre_obs
        if copy:
            obs = TensorDict({k: v.clone() for k, v in obs.items()})
            reward, done, record = reward.clone(), done.clone(), record.clone()
            pre_obs = {k: v.clone() for k, v in pre_obs.items()}
        info = env._info = self._RecordInfo(env.num_agent, record, pre_obs, dyn.ctrl_dt, False)
        return obs, reward, done, info


def capture_or_none(env, action):
    """A recorded step, or None (with a warning, once) if the task code cannot be recorded."""
    try:
        return TaskStepGraph(env, action)
    except Exception as e:                                                   # noqa: BLE001 - any failure => eager path
        th.cuda.synchronize(env.device)
        warnings.warn(f"capture_task_step: this env's step cannot be recorded as a CUDA graph ({type(e).__name__}: {e}); "
                      "staying on the two-launch path")
        return None
