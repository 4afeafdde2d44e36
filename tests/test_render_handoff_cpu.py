"""Frame conversion of the renderer hand-off against the reference's own functions, lifted out of
utils/common.py by AST (the module imports packages that are absent here) — exact equality, both directions."""
import ast
import os
from typing import Optional, Tuple

import numpy as np
import pytest
import torch as th

from visfly_b200.render_handoff import habitat_to_std, std_to_habitat

REF = "/root/reference/utils/common.py"


def lift(name):
    node = next(n for n in ast.parse(open(REF).read()).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"th": th, "np": np, "Optional": Optional, "Tuple": Tuple, "Tensor": th.Tensor}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns[name]


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not mounted")
def test_frame_conversion_equals_live_reference():
    ref_s2h, ref_h2s = lift("std_to_habitat"), lift("habitat_to_std")
    g = th.Generator().manual_seed(0)
    pos, ori = th.randn(37, 3, generator=g), th.randn(37, 4, generator=g)
    hp, ho = std_to_habitat(pos, ori)
    rp, ro = ref_s2h(pos, ori)
    assert np.array_equal(hp, rp) and np.array_equal(ho, ro)
    assert np.array_equal(std_to_habitat(pos[0], None)[0], ref_s2h(pos[0], None)[0])     # single vector form
    assert std_to_habitat(None, None) == (None, None)
    sp, so = habitat_to_std(hp, ho)
    qp, qo = ref_h2s(rp, ro)
    assert th.equal(sp, qp) and th.equal(so.float(), qo.float())
    assert th.equal(sp, pos) and th.equal(so.float(), ori)


def test_frame_conversion_known_axes():
    """x forward, y left, z up  ->  Habitat: -z forward, -x left... : (x,y,z) -> (-y, z, -x)."""
    hp, ho = std_to_habitat(th.tensor([[1., 2., 3.]]), th.tensor([[0.5, 0.1, 0.2, 0.3]]))
    assert hp.tolist() == [[-2., 3., -1.]]
    assert np.allclose(ho, [[0.5, -0.2, 0.3, -0.1]])
