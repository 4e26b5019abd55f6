"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. It wraps ``oracle/inversus_oracle.c``,
the plain-C restatement of the reference's ``inversus/core.py`` +
``inversus_rl/env_wrappers.py`` rollout path (see that file's header for citations).

Parity status: PINNED (tests/golden/, tests/test_oracle_golden.py,
tests/test_oracle_vs_reference_live.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libinversus_oracle.so")

MAX_DIM = 16
MAX_BULLETS = 64
TABLE_STRIDE = 64
TABLE_RESET_OFF = 16
STREAM_RESET = 0xFFFFFFFF
BLACK, WHITE = 0, 1
UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3


class Bullet(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("dir", C.c_int32), ("owner", C.c_int32)]


class Player(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("ammo", C.c_int32),
                ("reload", C.c_int32), ("alive", C.c_int32)]


class Env(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("grid", C.c_uint8 * (MAX_DIM * MAX_DIM)),
        ("p", Player * 2),
        ("n_bullets", C.c_int32),
        ("bullets", Bullet * MAX_BULLETS),
        ("step_count", C.c_int32),
        ("prev_alive", C.c_int32 * 2),
        ("episode_return", C.c_double),
        ("max_episode_steps", C.c_int32),
        ("difficulty", C.c_int32),
        ("mode", C.c_int32),
        ("episode", C.c_uint32),
        ("env_gid", C.c_uint32),
        ("seed", C.c_uint64),
        ("table", C.POINTER(C.c_uint32)),
        ("draws_used", C.c_int32),
        ("bullet_overflow", C.c_int32),
    ]


class StepOut(C.Structure):
    _fields_ = [("reward", C.c_double), ("done", C.c_int32), ("landed_hit", C.c_int32),
                ("got_hit", C.c_int32), ("win", C.c_int32), ("lose", C.c_int32),
                ("episode_steps", C.c_int32), ("episode_return", C.c_double), ("a2", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile). Returns the .so path."""
    src = os.path.join(_HERE, "inversus_oracle.c")
    hdr = os.path.join(_HERE, "inversus_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libinversus_oracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER
        L.orc_philox4x32_10.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
        L.orc_draw_u32.restype = C.c_uint32
        L.orc_draw_u32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_init.argtypes = [P(Env), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32]
        for name in ("orc_engine_reset", "orc_reload_ammo", "orc_update_bullets", "orc_rl_reset"):
            getattr(L, name).argtypes = [P(Env)]
            getattr(L, name).restype = None
        for name in ("orc_try_move", "orc_spawn_bullet", "orc_spawn_wide_shot"):
            getattr(L, name).argtypes = [P(Env), C.c_int, C.c_int]
            getattr(L, name).restype = C.c_int
        L.orc_apply_action.argtypes = [P(Env), C.c_int, C.c_int]
        L.orc_apply_action.restype = None
        L.orc_step_players.argtypes = [P(Env), C.c_int, C.c_int]
        L.orc_step_players.restype = None
        for name in ("orc_is_round_over", "orc_get_winner", "orc_dummy_policy"):
            getattr(L, name).argtypes = [P(Env)]
            getattr(L, name).restype = C.c_int
        L.orc_build_obs.argtypes = [P(Env), C.c_int, C.c_void_p, C.c_void_p]
        L.orc_build_obs.restype = None
        L.orc_rl_step.argtypes = [P(Env), C.c_int, C.c_int, P(StepOut)]
        L.orc_rl_step.restype = C.c_int
        L.orc_batch_reset.argtypes = [C.c_void_p, C.c_int64, C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        L.orc_batch_reset.restype = C.c_int
        L.orc_batch_step.argtypes = ([C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
                                     + [C.c_void_p] * 9 + [C.c_int])
        L.orc_batch_step.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_sizeof_env.restype = C.c_int64
        assert L.orc_sizeof_env() == C.sizeof(Env), (L.orc_sizeof_env(), C.sizeof(Env))
        _lib = L
    return _lib


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def draw_u32(seed, env_gid, episode, stream, k):
    return int(lib().orc_draw_u32(seed, env_gid, episode, stream, k))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N oracle envs stepped together (MultiEnvRunner semantics, env_wrappers.py:447-528),
    optionally with the trainer's auto-reset (training.py:140-151)."""

    def __init__(self, n, mode="dummy", difficulty="hard", max_episode_steps=500, seed=0,
                 env_id_base=0, width=15, height=10, nthreads=1):
        self.n = int(n)
        self.width, self.height = width, height
        self.nthreads = nthreads
        self.envs = (Env * self.n)()
        L = lib()
        m = {"dummy": 0, "selfplay": 1}[mode]
        d = {"easy": 0, "hard": 1}[difficulty]
        for i in range(self.n):
            L.orc_init(C.byref(self.envs[i]), width, height, m, d, max_episode_steps, seed, env_id_base + i)
        self.mode = mode
        hw = width * height
        self.obs1 = np.zeros((self.n, 12, height, width), np.float32)
        self.extra1 = np.zeros((self.n, 4), np.float32)
        self.obs2 = np.zeros((self.n, 12, height, width), np.float32) if mode == "selfplay" else None
        self.extra2 = np.zeros((self.n, 4), np.float32) if mode == "selfplay" else None
        self.reward = np.zeros(self.n, np.float32)
        self.done = np.zeros(self.n, np.uint8)
        self.flags = np.zeros(self.n, np.uint8)
        self.episode_steps = np.zeros(self.n, np.int32)
        self.episode_return = np.zeros(self.n, np.float64)
        del hw

    def reset(self, table=None):
        if table is not None:
            table = np.ascontiguousarray(table, np.uint32)
            assert table.shape == (self.n, TABLE_STRIDE)
        lib().orc_batch_reset(C.addressof(self.envs), self.n, _ptr(table), _ptr(self.obs1),
                              _ptr(self.extra1), _ptr(self.obs2), _ptr(self.extra2), self.nthreads)
        return self.obs1, self.extra1

    def step(self, a1, a2=None, table=None, auto_reset=False, want_obs=True):
        a1 = np.ascontiguousarray(a1, np.int8)
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, np.int8)
        if table is not None:
            table = np.ascontiguousarray(table, np.uint32)
            assert table.shape == (self.n, TABLE_STRIDE)
        o1 = self.obs1 if want_obs else None
        o2 = self.obs2 if want_obs else None
        rc = lib().orc_batch_step(C.addressof(self.envs), self.n, _ptr(a1), _ptr(a2), _ptr(table),
                                  int(auto_reset), _ptr(o1), _ptr(self.extra1), _ptr(o2),
                                  _ptr(self.extra2), _ptr(self.reward), _ptr(self.done),
                                  _ptr(self.flags), _ptr(self.episode_steps),
                                  _ptr(self.episode_return), self.nthreads)
        if rc != 0:
            raise ValueError("Invalid action_id: must be 0-12")
        return (self.obs1, self.extra1), self.reward, self.done.astype(bool), self.flags

    # ---- canonical state, same field layout the CUDA export uses (include/inversus_b200.h) ----
    def export_state(self):
        return export_state(self.envs, self.n)


STATE_DTYPE = np.dtype([
    ("tiles", np.uint32, (5,)),      # bit y*15+x set = WHITE
    ("p1", np.int32, (5,)),          # x, y, ammo, reload, alive
    ("p2", np.int32, (5,)),
    ("n_bullets", np.int32),
    ("bullets", np.int8, (16, 4)),   # x, y, dir, owner  (unused slots = 0)
    ("step_count", np.int32),
    ("episode", np.uint32),
    ("episode_return", np.float64),
], align=True)


def export_state(envs, n):
    """Canonical unpacked state of 15x10 oracle envs as a structured array (STATE_DTYPE)."""
    out = np.zeros(n, STATE_DTYPE)
    for i in range(n):
        e = envs[i]
        assert e.width == 15 and e.height == 10
        g = np.frombuffer(e.grid, np.uint8, 150)
        bits = np.zeros(160, np.uint8)
        bits[:150] = g
        out["tiles"][i] = np.packbits(bits, bitorder="little").view(np.uint32)
        for k, name in enumerate(("p1", "p2")):
            p = e.p[k]
            out[name][i] = (p.x, p.y, p.ammo, p.reload, p.alive)
        nb = e.n_bullets
        assert nb <= 16, "oracle state exceeds the device's 16 bullet slots"
        out["n_bullets"][i] = nb
        for b in range(nb):
            bb = e.bullets[b]
            out["bullets"][i, b] = (bb.x, bb.y, bb.dir, bb.owner)
        out["step_count"][i] = e.step_count
        out["episode"][i] = e.episode
        out["episode_return"][i] = e.episode_return
    return out


def import_state(envs, state):
    """Inverse of export_state (15x10 only)."""
    for i in range(len(state)):
        e = envs[i]
        s = state[i]
        bits = np.unpackbits(np.ascontiguousarray(s["tiles"]).view(np.uint8), bitorder="little")[:150]
        for t in range(150):
            e.grid[t] = int(bits[t])
        for k, name in enumerate(("p1", "p2")):
            p = e.p[k]
            p.x, p.y, p.ammo, p.reload, p.alive = (int(v) for v in s[name])
            e.prev_alive[k] = p.alive
        e.n_bullets = int(s["n_bullets"])
        for b in range(e.n_bullets):
            bb = e.bullets[b]
            bb.x, bb.y, bb.dir, bb.owner = (int(v) for v in s["bullets"][b])
        e.step_count = int(s["step_count"])
        e.episode = int(s["episode"])
        e.episode_return = float(s["episode_return"])
