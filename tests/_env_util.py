"""Replay helpers shared by the env-level tests (oracle port on the CPU, product envs on the GPU)."""
from __future__ import annotations

import os

import numpy as np
import torch as th

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DYN = {
    "euler": dict(action_type="bodyrate", integrator="euler", dt=0.005, ctrl_dt=0.02),
    "rk4": dict(action_type="bodyrate", integrator="rk4", dt=0.0025, ctrl_dt=0.02),
}


def load_env_golden(task, integ):
    return np.load(os.path.join(GOLD, f"env_{task}_{integ}.npz"))


def table_of(z, device="cpu"):
    return tuple(th.from_numpy(z["table_" + k]).to(device) for k in ("pos", "quat", "vel", "rate"))


def episode_records(done, info, n):
    """Arrays in the layout of the golden files from one step's (done, info)."""
    er, el = np.full(n, np.nan, np.float32), np.full(n, -1, np.int32)
    tr, sc, co, pg = np.zeros(n, bool), np.zeros(n, bool), np.zeros(n, bool), np.full(n, -1, np.int32)
    for i in np.nonzero(np.asarray(done))[0]:
        rec = info[int(i)]
        er[i], el[i] = rec["episode"]["r"], rec["episode"]["l"]
        tr[i], sc[i] = rec["TimeLimit.truncated"], rec["is_success"]
        co[i] = bool(rec["episode"]["extra"]["collision"])
        pg[i] = rec["episode"]["extra"].get("past_gate", -1)
    return er, el, tr, sc, co, pg
