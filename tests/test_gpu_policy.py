"""The one-launch actor kernels (csrc/vf_policy.cu) against the library ops they replace: same action, same gradients
w.r.t. the observation and every parameter (float32 reference on the GPU, float64 reference as arbiter)."""
import copy

import pytest
import torch as th

from _util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,h", [(65536, 16, 64), (1000, 13, 64), (4099, 17, 32), (37, 16, 32), (128, 32, 64)])
def test_fused_actor_matches_the_library_ops(n, d, h):
    from visfly_b200.algorithms.policies import Actor
    th.manual_seed(n + d + h)
    actor = Actor(d, 4, (h, h)).cuda()
    for p in actor.parameters():                       # larger weights than the default init: tanh in its curved range
        p.data.mul_(2.0)
    x = (th.randn(n, d, device="cuda") * 1.5).requires_grad_(True)
    assert actor.fused_ok(x)
    g = th.randn(n, 4, device="cuda")
    a = actor.deterministic_action(x, -1.0, 1.0)
    params = [p for p in actor.parameters() if p is not actor.log_std]
    grads = th.autograd.grad((a * g).sum(), [x] + params)
    ref32 = th.clip(th.tanh(actor.mu(actor.body(x))), -1.0, 1.0)
    grads32 = th.autograd.grad((ref32 * g).sum(), [x] + params)
    actor64 = copy.deepcopy(actor).double()
    x64 = x.detach().double().requires_grad_(True)
    ref64 = th.clip(th.tanh(actor64.mu(actor64.body(x64))), -1.0, 1.0)
    params64 = [p for p in actor64.parameters() if p is not actor64.log_std]
    grads64 = th.autograd.grad((ref64 * g.double()).sum(), [x64] + params64)
    assert rel_l2(a.detach().cpu(), ref64.detach().cpu()) < 2e-6
    if d > 3:               # the same observation handed over as a dict of two pieces: identical results, no cat
        xa = x.detach()[:, :d - 3].contiguous().requires_grad_(True)
        xb = x.detach()[:, d - 3:].contiguous().requires_grad_(True)
        a2 = actor.deterministic_action({"a_first": xa, "b_second": xb}, -1.0, 1.0)
        g2 = th.autograd.grad((a2 * g).sum(), [xa, xb] + params)
        assert th.equal(a2, a) and th.equal(th.cat([g2[0], g2[1]], 1), grads[0])
        assert all(th.equal(u, v) for u, v in zip(g2[2:], grads[1:]))
    for got, r32, r64 in zip(grads, grads32, grads64):
        err, floor = rel_l2(got.cpu(), r64.cpu()), rel_l2(r32.cpu(), r64.cpu())
        assert err < max(5e-6, 3 * floor), (tuple(got.shape), err, floor)
    # bit-reproducible: the per-tile weight gradients are added in a fixed order
    again = th.autograd.grad((actor.deterministic_action(x, -1.0, 1.0) * g).sum(), [x] + params)
    assert all(th.equal(u, v) for u, v in zip(grads, again))


def test_unsupported_networks_take_the_library_path():
    from visfly_b200.algorithms.policies import Actor
    x = th.randn(64, 16, device="cuda")
    for arch in ((16, 16), (64, 32), (64, 64, 64)):
        actor = Actor(16, 4, arch).cuda()
        assert not actor.fused_ok(x)
        a = actor.deterministic_action(x)
        assert a.shape == (64, 4) and bool((a.abs() <= 1).all())
