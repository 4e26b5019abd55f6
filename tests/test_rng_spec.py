"""The draw stream both sides of every parity check share: Philox4x32-10 known answers, the
three independent implementations agree (pure Python in ref_harness, C in the oracle; the CUDA
one is checked in test_cuda_parity.py), and the integer thresholds the kernel uses are exactly
the reference's float comparisons (env_wrappers.py:82-89,96,105,123)."""
from fractions import Fraction

import numpy as np

from oracle import oracle as orc
from ref_harness import philox4x32_10 as py_philox

# Random123 kat_vectors for philox4x32-10
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers():
    for ctr, key, out in KAT:
        assert py_philox(ctr, key) == out
        assert orc.philox4x32_10(ctr, key) == out


def test_oracle_draw_indexing_matches_python_shim():
    rs = np.random.RandomState(0)
    for _ in range(200):
        seed = int(rs.randint(0, 2**62))
        env, ep, stream, k = (int(v) for v in rs.randint(0, 2**31, size=4))
        k %= 64
        want = py_philox((env, ep, stream, k >> 2), (seed & 0xFFFFFFFF, seed >> 32))[k & 3]
        assert orc.draw_u32(seed, env, ep, stream, k) == want


def _int_threshold_lt(p):
    """smallest T with: (r / 2**32 < p)  <=>  (r < T) for u32 r, p a binary64 constant."""
    f = Fraction(p) * 2**32
    return int(f) if f == int(f) else int(f) + 1


def test_integer_thresholds_equal_the_float_comparisons():
    from importlib import import_module
    th = import_module("inversus_b200.constants")
    assert th.THRESH_SHOOT_HARD == _int_threshold_lt(0.2)
    assert th.THRESH_RANDMOVE_HARD == _int_threshold_lt(0.05)
    # easy: `random() > 0.001` -> NONE; i.e. moves iff r/2^32 <= 0.001 iff r < T_le
    f = Fraction(0.001) * 2**32
    assert th.THRESH_MOVE_EASY == int(f) + 1
    # brute-force the neighbourhood of each threshold against the float expression
    for p, T in ((0.2, th.THRESH_SHOOT_HARD), (0.05, th.THRESH_RANDMOVE_HARD)):
        for r in range(T - 3, T + 3):
            assert ((r / 4294967296.0) < p) == (r < T)
    T = th.THRESH_MOVE_EASY
    for r in range(T - 3, T + 3):
        assert ((r / 4294967296.0) > 0.001) == (not r < T)


def test_kernel_lookup_tables_are_the_reference_arithmetic():
    """The fp64 proximity table and the fp32 ammo table in the CUDA source are exactly what the
    reference's Python expressions evaluate to (env_wrappers.py:378-382 and :238-240)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = open(os.path.join(root, "inversus-reinforcement-learning_b200", "csrc", "inversus_kernels.cuh")).read()
    prox = re.search(r"kProximity\[24\] = \{(.*?)\};", cu, re.S).group(1)
    vals = [float.fromhex(v.strip()) for v in prox.replace("\n", " ").split(",")]
    assert len(vals) == 24
    for dist, v in enumerate(vals):
        assert v == 0.002 * (1.0 - dist / 25), dist  # max_dist = width + height = 25
    ammo = re.search(r"kAmmoNorm\[8\] = \{(.*?)\};", cu, re.S).group(1)
    avals = [float.fromhex(v.strip().rstrip("f")) for v in ammo.replace("\n", " ").split(",")]
    for k in range(7):  # MAX_AMMO = 6
        assert np.float32(avals[k]) == np.float32(k / 6), k
        assert float(np.float32(avals[k])) == avals[k]
