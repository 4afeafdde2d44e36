"""Index arithmetic of the tensor-core actor kernels (csrc/vf_policy_tc.cuh), restated in Python: the operand layouts the
MMA descriptors assume and the properties the kernels rely on (every slot written once, warp-wide stores free of bank
conflicts, the lane -> value map of the warp butterfly, the lane map of an M = 64 accumulator).  No GPU needed."""
import itertools

import numpy as np


def blk(r, c, C):                       # blocked K-major layout: [r / 8][c / 4][r % 8][c % 4]
    return ((r // 8) * (C // 4) + c // 4) * 32 + (r % 8) * 4 + (c % 4)


def sw128(f, a, F):                     # (feature f, agent a) of an (F x 128) K = agents operand, 128-byte-swizzled rows
    l = a & 31
    return (a >> 5) * F * 32 + f * 32 + ((((l >> 2) ^ (f & 7)) << 2) | (l & 3))


def test_blocked_layout_is_a_bijection_with_the_descriptor_strides():
    for R, C in ((64, 16), (64, 64), (16, 64), (32, 32)):
        idx = {blk(r, c, C) for r in range(R) for c in range(C)}
        assert idx == set(range(R * C))
        # descriptor: 16-byte chunks of one row 128 B apart (LBO), 8-row groups (C / 4) * 128 B apart (SBO)
        for r, c in itertools.product(range(R), range(0, C, 4)):
            assert blk(r, c, C) * 4 == (r % 8) * 16 + (c // 4) * 128 + (r // 8) * (C // 4) * 128
        # hi and lo parts stored back to back: the lo rows continue the row-group stride (one B operand of twice the N)
        assert blk(R, 0, C) == R * C


def test_swizzled_rows_every_slot_once_and_warp_stores_hit_32_banks():
    for F in (24, 64, 72):
        idx = {sw128(f, a, F) for f in range(F) for a in range(128)}
        assert idx == set(range(F * 128))
        for f in range(F):
            for w in range(4):                                   # one warp = 32 consecutive agents, one feature
                banks = {sw128(f, 32 * w + l, F) % 32 for l in range(32)}
                assert len(banks) == 32
            # the hardware pattern: 16-byte chunk index XOR (row % 8) inside 1024-byte atoms of 8 rows x 128 B
            for a in range(0, 128, 4):
                byte = sw128(f, a, F) * 4
                row_base = (a // 32) * F * 128 + f * 128
                assert byte - row_base == (((a % 32) // 4) ^ (f % 8)) * 16


def test_warp_butterfly_leaves_values_2l_and_2l_plus_1_in_lane_l():
    rng = np.random.default_rng(0)
    vals = rng.standard_normal((32, 64))                        # [lane][value]
    val = vals.copy()
    count, s = 64, 16
    while s > 0:
        new = val.copy()
        for lane in range(32):
            up = (lane & s) != 0
            for k in range(count // 2):
                send_from_partner = val[lane ^ s][k] if (((lane ^ s) & s) != 0) else val[lane ^ s][k + count // 2]
                keep = val[lane][k + count // 2] if up else val[lane][k]
                new[lane][k] = keep + send_from_partner
        val, count, s = new, count // 2, s >> 1
    total = vals.sum(0)
    for lane in range(32):
        assert np.allclose(val[lane][0], total[2 * lane]) and np.allclose(val[lane][1], total[2 * lane + 1])


def test_m64_accumulator_lane_map_covers_the_rows_once():
    lanes = [32 * (j // 16) + j % 16 for j in range(64)]         # measured: profiles/r02_umma_probe.txt, test 5
    assert len(set(lanes)) == 64 and max(lanes) < 128
    for lane in range(128):                                      # the read-out's inverse map
        j = 16 * (lane >> 5) + (lane & 31)
        assert ((lane & 31) < 16) == (lane in lanes) and (lane not in lanes or lanes[j] == lane)
