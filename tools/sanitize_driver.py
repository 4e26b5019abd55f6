"""Small torch-free workload for compute-sanitizer: drives the C ABI through ctypes with host
buffers only, covering both tile sizes (E=32 and E=128), both modes, host-expand staging, the
packed-snapshot rebuild and the debug entry.

    compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inversus_b200 import _capi  # noqa: E402

lib = _capi.load()


def run(n, mode, steps, flags):
    cfg = _capi.Config(n, 0, 7, _capi.MODE[mode], 1, 25, 0, 0, flags)
    h = C.c_void_p()
    _capi.check(lib.inv_create(C.byref(cfg), C.byref(h)))
    selfplay = mode == "selfplay"
    obs = np.empty((n, 12, 10, 15), np.float32)
    obs2 = np.empty((n, 12, 10, 15), np.float32) if selfplay else None
    extra, extra2 = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
    rew, done, info = np.empty(n, np.float32), np.empty(n, np.uint8), np.empty(n, np.uint8)
    steps_, ret = np.empty(n, np.int32), np.empty(n, np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _capi.check(lib.inv_reset_host(h, p(obs), p(extra), p(obs2), p(extra2) if selfplay else None))
    rs = np.random.RandomState(n)
    for t in range(steps):
        a1 = rs.randint(0, 13, n).astype(np.int8)
        a2 = rs.randint(0, 13, n).astype(np.int8) if selfplay else None
        _capi.check(lib.inv_step_host(h, p(a1), p(a2), p(obs), p(extra), p(obs2), p(extra2) if selfplay else None,
                                      p(rew), p(done), p(info), p(steps_), p(ret)))
        assert ((obs[:, 0] + obs[:, 1]) == 1).all()
    # K3 on the handle's own packed state into its own observation buffer
    ps, ob, ex = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nb = C.c_int64()
    _capi.check(lib.inv_get_buffer(h, _capi.BUF_PACKED_STATE, C.byref(ps), C.byref(nb)))
    _capi.check(lib.inv_get_buffer(h, _capi.BUF_OBS_P1, C.byref(ob), C.byref(nb)))
    _capi.check(lib.inv_get_buffer(h, _capi.BUF_EXTRA_P1, C.byref(ex), C.byref(nb)))
    for view in (0, 1):
        _capi.check(lib.inv_obs_from_packed(h, ps, n, n, view, 0, ob, ex, None))
    for phase in range(8):
        _capi.check(lib.inv_debug_phase(h, phase, 1, 2, 5, None))
    st = np.zeros(min(n, 64), _capi.STATE_DTYPE)
    _capi.check(lib.inv_export_state(h, p(st), 0, len(st)))
    _capi.check(lib.inv_import_state(h, p(st), 0, len(st)))
    bits = C.c_uint32()
    _capi.check(lib.inv_poll_status(h, None, C.byref(bits)))
    assert bits.value == 0
    _capi.check(lib.inv_destroy(h))
    print(f"ok n={n} mode={mode} steps={steps} episodes_done={int(done.sum())}", flush=True)


if __name__ == "__main__":
    run(300, "dummy", 30, _capi.FLAG_AUTO_RESET)
    run(333, "selfplay", 30, _capi.FLAG_AUTO_RESET)
    run(40000, "dummy", 6, _capi.FLAG_AUTO_RESET)       # E = 128 tiles, host-expand path (n >= 4096)
    run(38000, "selfplay", 4, 0)                        # two views, ragged last tile
    print("sanitize driver finished")
