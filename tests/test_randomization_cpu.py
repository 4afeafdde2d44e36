"""State generators (reference utils/randomization.py) — the `heading=True` option of the uniform generator."""
import math

import pytest
import torch as th

from _reference import load_reference, reference_available
from visfly_b200.randomization import UniformStateRandomizer, euler_zyx_to_quat

BOX = dict(position={"mean": [1.0, -2.0, 1.5], "half": [3.0, 2.0, 0.5]},
           orientation={"mean": [0.3, 0.2, 0.1], "half": [0.0, 0.0, 0.2]},
           velocity={"mean": [0.0, 0.0, 0.0], "half": [1.0, 1.0, 1.0]},
           angular_velocity={"mean": [0.0, 0.0, 0.0], "half": [0.5, 0.5, 0.5]})


def test_heading_points_back_to_the_box_centre():
    th.manual_seed(0)
    gen = UniformStateRandomizer(heading=True, device="cpu", **BOX)
    pos, eul, vel, rate = gen.generate(4096)
    mean = th.tensor(BOX["position"]["mean"])
    d = mean - pos
    yaw = th.atan2(d[:, 1], d[:, 0])
    # roll = pitch = 0 (their half-widths are 0 and the orientation mean is ignored), yaw = heading + noise in +-0.2
    assert float(eul[:, :2].abs().max()) == 0.0
    err = th.atan2(th.sin(eul[:, 2] - yaw), th.cos(eul[:, 2] - yaw))
    assert float(err.abs().max()) <= 0.2 + 1e-5 and float(err.abs().max()) > 0.15
    assert float((pos - mean).abs().max(0).values.sub(th.tensor(BOX["position"]["half"])).max()) <= 1e-6
    q = euler_zyx_to_quat(eul)
    assert th.allclose(q.norm(dim=1), th.ones(4096), atol=1e-6)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_heading_formula_equals_the_reference_function():
    load_reference()                                          # puts the VisFly package on sys.path
    from VisFly.utils.randomization import calculate_yaw_pitch
    th.manual_seed(1)
    box = dict(BOX, orientation={"mean": [0.0, 0.0, 0.0], "half": [0.0, 0.0, 0.0]})
    gen = UniformStateRandomizer(heading=True, device="cpu", **box)
    pos, eul, _, _ = gen.generate(1000)
    half = pos - th.tensor(box["position"]["mean"])
    yaw_ref, _ = calculate_yaw_pitch(-half)
    assert th.equal(eul[:, 2], yaw_ref) and float(eul[:, :2].abs().max()) == 0.0


def test_heading_needs_a_horizontal_extent():
    with pytest.raises(ValueError):
        UniformStateRandomizer(heading=True, device="cpu",
                               position={"mean": [0.0, 0.0, 1.0], "half": [0.0, 0.0, 1.0]})
    assert math.isfinite(1.0)
