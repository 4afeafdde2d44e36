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


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_euler_to_quaternion_equals_the_reference_conversion():
    load_reference()
    from VisFly.utils.maths import Quaternion as RefQuaternion
    g = th.Generator().manual_seed(5)
    eul = (th.rand(500, 3, generator=g) * 2 - 1) * th.tensor([3.1, 1.5, 3.1])
    ref = RefQuaternion.from_euler(*eul.T).toTensor().T                       # randomization.py:95
    assert th.allclose(euler_zyx_to_quat(eul), ref, rtol=0, atol=1e-7)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_generators_draw_from_the_reference_distributions():
    """Same families and moments as the reference classes (the random streams differ: one (num,4,3) draw here, four
    (num,3) draws there), incl. the reference's `(2*randn - 1) * std + mean` for the normal generator (:201-204)."""
    load_reference()
    from VisFly.utils import randomization as R
    from visfly_b200.randomization import NormalStateRandomizer, UnionRandomizer
    n = 200_000
    box = dict(position={"mean": [1.0, -2.0, 1.5], "half": [3.0, 2.0, 0.5]},
               orientation={"mean": [0.0, 0.1, 0.2], "half": [0.1, 0.2, 0.3]},
               velocity={"mean": [0.5, 0.0, -0.5], "half": [1.0, 2.0, 0.5]},
               angular_velocity={"mean": [0.0, 0.0, 0.0], "half": [0.5, 0.5, 0.5]})
    th.manual_seed(0)
    ours = UniformStateRandomizer(device="cpu", **box).generate(n)
    ref = R.UniformStateRandomizer(**box)._generate(n)
    for a, b in zip(ours, ref):
        assert th.allclose(a.mean(0), b.mean(0), atol=0.02) and th.allclose(a.std(0), b.std(0), atol=0.02)
        assert th.allclose(a.amin(0), b.amin(0), atol=0.01) and th.allclose(a.amax(0), b.amax(0), atol=0.01)
    nbox = {k: {"mean": v["mean"], "std": v["half"]} for k, v in box.items()}
    # the reference's NormalStateRandomizer converts only `position` to a Normal (:196) and raises AttributeError on
    # the orientation line of _generate (:200); the formula it writes for all four fields is what is implemented here
    with pytest.raises(AttributeError):
        R.NormalStateRandomizer(**nbox)._generate(4)
    ours = NormalStateRandomizer(device="cpu", **nbox).generate(n)
    for a, k in zip(ours, nbox):
        mean, std = th.tensor(nbox[k]["mean"]), th.tensor(nbox[k]["std"])
        assert th.allclose(a.mean(0), mean - std, atol=0.03) and th.allclose(a.std(0), 2 * std, atol=0.03)
    two = [{"class": "Uniform", "kwargs": {"position": {"mean": [10.0, 0, 1], "half": [0.1, 0.1, 0.1]}}},
           {"class": "Uniform", "kwargs": {"position": {"mean": [-10.0, 0, 1], "half": [0.1, 0.1, 0.1]}}}]
    pos = UnionRandomizer(two, device="cpu").generate(20_000)[0]
    assert abs(float((pos[:, 0] > 0).float().mean()) - 0.5) < 0.02                # each agent picks a box uniformly
    assert float((pos[:, 0].abs() - 10).abs().max()) <= 0.1 + 1e-6


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_evaluation_grid_walks_the_box_like_the_reference():
    """`test=True` (reference randomization.py:141-160): start positions walk over a regular grid of the position box,
    one grid point per generated agent.  The reference is called agent by agent (droneEnv.py:243-249); the vectorised
    generator must hand agent k the same grid point (the U(-1,1)*xyz_half jitter is switched off to compare)."""
    load_reference()
    from VisFly.utils import randomization as R
    box = dict(position={"mean": [1.0, -2.0, 1.5], "half": [3.0, 2.0, 0.5]})
    kw = dict(test=True, xyz_num=[3, 2, 2], xyz_half=[0.0, 0.0, 0.0])
    ref = R.UniformStateRandomizer(**box, **kw)
    ref_pos = th.cat([ref._generate(1)[0] for _ in range(30)])
    ours = UniformStateRandomizer(device="cpu", **box, **kw)
    pos = th.cat([ours.generate(17)[0], ours.generate(13)[0]])            # two batches continue the walk
    assert th.allclose(pos, ref_pos, atol=1e-6)
    jit = UniformStateRandomizer(device="cpu", **box, test=True, xyz_num=[3, 2, 2], xyz_half=[0.0, 2.0, 0.0])
    p2 = jit.generate(30)[0]
    assert float((p2 - ref_pos)[:, [0, 2]].abs().max()) < 1e-6 and 0.5 < float((p2 - ref_pos)[:, 1].abs().max()) <= 2.0


def test_unknown_generator_kwargs_are_rejected():
    with pytest.raises(TypeError, match="unknown keyword"):
        UniformStateRandomizer(device="cpu", positon={"mean": [0, 0, 0], "half": [1, 1, 1]})     # typo
