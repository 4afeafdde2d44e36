"""Golden vectors for the trainer utilities from the REAL reference.

    python tests/golden/make_algo_golden.py

``compute_td_returns`` lives in reference utils/algorithms/common.py, a module that imports stable-baselines3 at the
top (absent here), so the function is lifted out of the file by its AST node and executed unmodified.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VISFLY_REFERENCE", "/root/reference")


def reference_function(path: str, name: str):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"th": th}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def td_cases(seed=0):
    g = th.Generator().manual_seed(seed)
    cases = []
    for h, n, p_done, p_ep in ((8, 16, 0.15, 0.5), (32, 64, 0.05, 0.7), (5, 7, 0.5, 0.0), (12, 9, 0.0, 0.0)):
        r = th.randn(h, n, generator=g)
        done = th.rand(h, n, generator=g) < p_done
        episode_done = done & (th.rand(h, n, generator=g) < p_ep)
        v = th.randn(h, n, generator=g)
        cases.append((r, done, episode_done, v))
    return cases


def main():
    fn = reference_function(os.path.join(REF, "utils", "algorithms", "common.py"), "compute_td_returns")
    out = {}
    for i, (r, done, ep, v) in enumerate(td_cases()):
        for tag, kw in (("a", dict(gamma=0.99, lamda=0.95)), ("b", dict(gamma=0.9, lamda=0.5))):
            ret = fn(list(r), list(done), list(v), episode_done=list(ep), **kw)
            out[f"case{i}{tag}_returns"] = th.stack(ret).numpy()
        ret = fn(list(r), list(done), list(v))                      # episode_done defaults to done
        out[f"case{i}c_returns"] = th.stack(ret).numpy()
        out[f"case{i}_r"], out[f"case{i}_done"], out[f"case{i}_ep"], out[f"case{i}_v"] = \
            r.numpy(), done.numpy(), ep.numpy(), v.numpy()
    np.savez_compressed(os.path.join(HERE, "td_returns.npz"), **out)
    print("td_returns.npz", os.path.getsize(os.path.join(HERE, "td_returns.npz")), "bytes")


if __name__ == "__main__":
    main()
