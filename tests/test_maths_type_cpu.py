"""The compatibility helpers task code touches — ``visfly_b200.maths.Quaternion`` and ``visfly_b200.type`` — against the
live reference classes (utils/maths.py:4-293, utils/type.py) on the CPU, method by method."""
import pytest
import torch as th

from _reference import load_reference, reference_available
from visfly_b200.maths import Quaternion
from visfly_b200.type import ACTION_TYPE, TensorDict

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


def ref_modules():
    load_reference()
    from VisFly.utils import maths, type as rtype
    return maths, rtype


def rand_quats(n, seed, unit=True):
    g = th.Generator().manual_seed(seed)
    q = th.randn(4, n, generator=g)
    return q / q.norm(dim=0) if unit else q


def test_quaternion_methods_match_reference():
    maths, _ = ref_modules()
    n = 257
    qa, qb = rand_quats(n, 1), rand_quats(n, 2, unit=False)
    v = th.randn(3, n, generator=th.Generator().manual_seed(3))
    ours_a, ours_b = Quaternion.from_tensor(qa.clone()), Quaternion.from_tensor(qb.clone())
    ref_a, ref_b = maths.Quaternion(*qa.clone()), maths.Quaternion(*qb.clone())
    close = lambda x, y, tol=2e-6: th.testing.assert_close(th.as_tensor(x), th.as_tensor(y), rtol=tol, atol=tol)
    close(ours_a.toTensor(), ref_a.toTensor())
    close((ours_a * ours_b).toTensor(), (ref_a * ref_b).toTensor())          # Hamilton product, maths.py:168-178
    close(ours_b.rotate(v), ref_b.rotate(v))                                  # non-unit q: scales by |q|^2, :32-38
    close(ours_b.inv_rotate(v), ref_b.inv_rotate(v))                          # :40-49
    close(ours_b.normalize().toTensor(), ref_b.normalize().toTensor())        # :229-230
    close(ours_b.norm(), ref_b.norm())
    close(ours_a.conjugate().toTensor(), ref_a.conjugate().toTensor())
    close((ours_a + ours_b).toTensor(), (ref_a + ref_b).toTensor())
    close(ours_a.x_axis, ref_a.x_axis)                                        # :122-133
    close(ours_a.R, ref_a.R)
    close(ours_a.toEuler(), ref_a.toEuler(), 1e-5)                            # :244-249
    eul = ref_a.toEuler()
    close(Quaternion.from_euler(*eul).toTensor(), maths.Quaternion.from_euler(*eul).toTensor(), 1e-6)
    assert len(ours_a) == n and ours_a[5].toTensor().shape[0] == 4


def test_type_helpers_match_reference():
    _, rtype = ref_modules()
    assert [a.name for a in ACTION_TYPE] == [a.name for a in rtype.ACTION_TYPE]
    assert [a.value for a in ACTION_TYPE] == [a.value for a in rtype.ACTION_TYPE]
    data = {"state": th.arange(26.).reshape(2, 13), "target": th.ones(2, 3)}
    ours, ref = TensorDict(data), rtype.TensorDict(data)
    for idx in (0, slice(0, 1), th.tensor([1])):
        a, b = ours[idx], ref[idx]
        assert set(a.keys()) == set(b.keys())
        for k in a.keys():
            assert th.equal(th.as_tensor(a[k]), th.as_tensor(b[k])), (idx, k)
    assert th.equal(ours.detach()["state"], ref.detach()["state"]) and len(ours) == len(ref)
    assert th.equal(ours.clone()["target"], ref.clone()["target"])
