"""Host logic of the comm-delay FIFO guard (visfly_b200.dynamics.Dynamics._fifo_push / _fifo_pop) on CPU tensors: the
reference clones every action into its FIFO (dynamics.py:324); the engine keeps the caller's tensor and raises if it
was modified in place before being consumed."""
import copy
import types

import pytest
import torch as th

from visfly_b200.dynamics import Dynamics


def fifo(depth=3):
    d = types.SimpleNamespace(_pre_action=[th.zeros(2, 4) for _ in range(depth)], _fifo_versions=[None] * depth,
                              _comm_delay_steps=depth)
    return d, (lambda a: Dynamics._fifo_push(d, a)), (lambda: Dynamics._fifo_pop(d))


def test_fresh_tensors_come_out_delayed_and_unchanged():
    d, push, pop = fifo()
    for t in range(10):
        push(th.full((2, 4), float(t)))
        out = pop()
        assert float(out[0, 0]) == max(t - 3, 0) and len(d._pre_action) == len(d._fifo_versions) == 3


def test_one_unmodified_tensor_every_step_is_fine():
    d, push, pop = fifo()
    a = th.ones(2, 4)
    for _ in range(10):
        push(a)
        assert pop() is not None


def test_in_place_reuse_raises_when_the_stale_entry_is_consumed():
    d, push, pop = fifo()
    buf = th.zeros(2, 4)
    with pytest.raises(RuntimeError, match="modified in place"):
        for t in range(10):
            buf.copy_(th.full((2, 4), float(t)))
            push(buf)
            pop()
    d, push, pop = fifo()
    for t in range(10):                                   # the remedy named in the message
        buf.copy_(th.full((2, 4), float(t)))
        push(buf.clone())
        assert float(pop()[0, 0]) == max(t - 3, 0)


def test_engine_made_copies_are_not_checked():
    d, push, pop = fifo()
    a = th.ones(2, 4)
    for _ in range(4):
        push(a)
        pop()
    twin = copy.deepcopy(d)                               # deepcopy of the env: entries are private copies now
    a.add_(1)                                             # ... so touching the original afterwards is harmless there
    for _ in range(4):
        Dynamics._fifo_push(twin, th.zeros(2, 4))
        Dynamics._fifo_pop(twin)
    d._pre_action = [x.detach() for x in d._pre_action]   # what Dynamics.detach() does: new objects, check skipped
    for _ in range(3):
        push(th.zeros(2, 4))
        pop()
