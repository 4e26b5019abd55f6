"""CPU checks of the drop-in boundary: the C-ABI library builds/loads without a GPU, exports every
symbol include/inversus_b200.h declares, the ctypes mirror matches the header's layout constants,
and -- with no device present -- every compute entry point fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "inversus_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(inv_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from inversus_b200 import _capi, build_library
    path = build_library()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "ctypes binding and header drifted apart"
    assert _capi.load().inv_abi_version() == 1


def test_header_constants_match_the_python_mirror():
    from inversus_b200 import constants as K
    src = open(HEADER).read()

    def define(name):
        m = re.search(rf"#define {name}\s+\(?([0-9xA-Fa-fu]+)", src)
        assert m, name
        return int(m.group(1).rstrip("u"), 0)
    assert define("INV_BOARD_W") == K.BOARD_W and define("INV_BOARD_H") == K.BOARD_H
    assert define("INV_MAX_BULLETS") == K.MAX_BULLETS
    assert define("INV_PACKED_STATE_BYTES") == K.PACKED_STATE_BYTES
    assert define("INV_TABLE_STRIDE") == K.TABLE_STRIDE and define("INV_TABLE_RESET_OFF") == K.TABLE_RESET_OFF
    assert define("INV_STREAM_RESET") == K.STREAM_RESET
    assert K.OBS_ELEMS == 1800
    assert K.algorithmic_bytes_per_env_step(4, False) == 7200 + 16 + 18 + 160 + 1


def test_state_struct_layout_matches_oracle_and_harness():
    from inversus_b200 import _capi
    from oracle import oracle as orc
    import ref_harness
    assert _capi.STATE_DTYPE == orc.STATE_DTYPE == ref_harness.STATE_DTYPE
    assert _capi.STATE_DTYPE.itemsize == 144


def test_kernel_thresholds_in_the_cuda_source_match_constants():
    from inversus_b200 import constants as K
    cu = open(os.path.join(ROOT, "inversus-reinforcement-learning_b200", "csrc", "inversus_kernels.cuh")).read()
    for name, val in (("kThreshShootHard", K.THRESH_SHOOT_HARD), ("kThreshRandMoveHard", K.THRESH_RANDMOVE_HARD),
                      ("kThreshMoveEasy", K.THRESH_MOVE_EASY)):
        m = re.search(rf"{name}\s*=\s*(\d+)u", cu)
        assert m and int(m.group(1)) == val, name


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from inversus_b200 import BatchedInversus, InversusError, MultiEnvRunner
    with pytest.raises(InversusError, match="no CPU fallback"):
        BatchedInversus(4, seed=0)
    with pytest.raises(InversusError):
        MultiEnvRunner(4, seed=0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "inversus-reinforcement-learning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle and the live-reference harness", ""), f
                # nor the staged copy of the reference (oracle/_ref, baseline/_ref) or the reference itself
                assert "_ref/" not in txt and "_ref\"" not in txt and "/root/reference" not in txt, f
                assert "import inversus_rl" not in txt and "from inversus_rl" not in txt, f


@pytest.mark.parametrize("no_avx512", [False, True])
def test_host_side_observation_expansion(no_avx512):
    """inv_host_expand_f32 (the CPU half of inv_step_host's packed path) against numpy.unpackbits:
    random rows, odd/even ranges, 1 and many threads, 64-byte-aligned and unaligned destinations.
    Runs in a subprocess so that INV_NO_AVX512 can select the AVX2 path as well."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from inversus_b200 import _capi
lib = _capi.load()
rs = np.random.RandomState(0)
n = 1031
bits = rs.randint(0, 2**32, size=(n, 64), dtype=np.uint64).astype(np.uint32)
want = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :1800].astype(np.float32)
for off in (0, 8):                      # aligned / unaligned destination
    raw = np.full(n * 1800 + 64, -7.0, np.float32)
    base = (-raw.ctypes.data // 4) %% 16 + off
    dst = raw[base: base + n * 1800].reshape(n, 1800)
    for first, count, nt in ((0, n, 1), (0, n, 7), (1, n - 1, 3), (5, 600, 2), (6, 1, 1), (7, 0, 4)):
        dst[:] = -7.0
        rc = lib.inv_host_expand_f32(bits.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), first, count, nt)
        assert rc == 0
        assert np.array_equal(dst[first:first + count], want[first:first + count]), (off, first, count, nt)
        assert (dst[:first] == -7.0).all() and (dst[first + count:] == -7.0).all()
        assert raw[base - 1] == -7.0 and raw[base + n * 1800] == -7.0
print("expand ok")
''' % ROOT
    env = dict(os.environ)
    if no_avx512:
        env["INV_NO_AVX512"] = "1"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "expand ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("no_avx2", [False, True])
def test_host_side_action_id_check(no_avx2):
    """inv_host_stage_action_ids (what every *_host call does before anything is enqueued): ids copied
    verbatim, -3 exactly when an id is outside 0..12 -- at every position of the vector body and the
    scalar tail, for negative ids too. Subprocess so that INV_NO_AVX2 selects the scalar path."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from inversus_b200 import _capi
lib = _capi.load()
rs = np.random.RandomState(1)
for n in (0, 1, 31, 64, 65, 127, 128, 1000, 4097):
    ids = rs.randint(0, 13, size=n).astype(np.int8)
    out = np.full(n + 2, 99, np.int8)
    dst = out[1:n + 1]
    assert lib.inv_host_stage_action_ids(ids.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n) == 0
    assert np.array_equal(dst, ids) and out[0] == 99 and out[n + 1] == 99
    for pos in sorted(set([0, n // 2, n - 1, max(n - 33, 0), min(63, max(n - 1, 0))])) if n else []:
        for val in (13, 127, -1, -128):
            bad = ids.copy()
            bad[pos] = val
            rc = lib.inv_host_stage_action_ids(bad.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n)
            assert rc == -3, (n, pos, val, rc)
    twelve = np.full(n, 12, np.int8)            # the largest legal id everywhere
    assert lib.inv_host_stage_action_ids(twelve.ctypes.data_as(C.c_void_p), twelve.ctypes.data_as(C.c_void_p), n) == 0
print("ids ok")
''' % ROOT
    env = dict(os.environ)
    if no_avx2:
        env["INV_NO_AVX2"] = "1"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ids ok" in r.stdout, r.stdout + r.stderr
