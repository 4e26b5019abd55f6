"""InversusEnv-shaped facades so the reference's own unit tests can be restated verbatim.

`OracleEngine` drives oracle/inversus_oracle.c one engine call at a time (core.py method names);
`CudaEngine` (tests/test_cuda_reference_kat.py) drives the CUDA product through its debug-phase
entry point with the same surface. Directions/ids use the product's integer encoding:
UP, RIGHT, DOWN, LEFT = 0..3; P1, P2 = 0, 1; BLACK, WHITE = 0, 1.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

Bullet = namedtuple("Bullet", "x y dir owner")
UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3
P1, P2 = 0, 1
BLACK, WHITE = 0, 1
MAX_AMMO, RELOAD_TICKS_PER_AMMO, WIDE_SHOT_AMMO_COST = 6, 30, 3
NONE = 0


def MOVE(d):
    return 1 + d


def SHOOT(d):
    return 5 + d


def CHARGE(d):
    return 9 + d


class _PlayerView:
    def __init__(self, get, set_):
        object.__setattr__(self, "_get", get)
        object.__setattr__(self, "_set", set_)

    def __getattr__(self, k):
        v = self._get(k)
        return bool(v) if k == "alive" else v

    def __setattr__(self, k, v):
        self._set(k, int(v))


class OracleEngine:
    """oracle/inversus_oracle.c behind core.py's InversusEnv method names."""

    color = {P1: BLACK, P2: WHITE}

    def __init__(self, width=15, height=10, seed=1234):
        from oracle import oracle as orc
        self.orc, self.L = orc, orc.lib()
        self.e = orc.Env()
        self.L.orc_init(C.byref(self.e), width, height, 1, 1, 500, seed, 0)
        self.width, self.height = width, height
        self.player_color = BLACK
        self.reset()

    # --- players
    def _pv(self, i):
        names = {"x": "x", "y": "y", "ammo": "ammo", "reload_counter": "reload", "alive": "alive"}
        return _PlayerView(lambda k: getattr(self.e.p[i], names[k]),
                           lambda k, v: setattr(self.e.p[i], names[k], v))

    @property
    def player1(self):
        return self._pv(0)

    @property
    def player2(self):
        return self._pv(1)

    # legacy single-player accessors (core.py:156-181)
    @property
    def player_x(self):
        return self.e.p[0].x

    @player_x.setter
    def player_x(self, v):
        self.e.p[0].x = v

    @property
    def player_y(self):
        return self.e.p[0].y

    @player_y.setter
    def player_y(self, v):
        self.e.p[0].y = v
        self.e.n_bullets = 0  # core.py:180-181: the legacy y setter clears the bullet list

    # --- bullets
    @property
    def bullets(self):
        return [Bullet(b.x, b.y, b.dir, b.owner) for b in self.e.bullets[: self.e.n_bullets]]

    @bullets.setter
    def bullets(self, lst):
        self.e.n_bullets = len(lst)
        for i, b in enumerate(lst):
            s = self.e.bullets[i]
            s.x, s.y, s.dir, s.owner = b.x, b.y, b.dir, b.owner

    def get_bullets(self):
        return self.bullets

    # --- tiles
    def _get_tile(self, x, y):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        return self.e.grid[y * self.width + x]

    def _set_tile(self, x, y, c):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        self.e.grid[y * self.width + x] = c

    # --- engine calls
    def reset(self):
        self.L.orc_engine_reset(C.byref(self.e))

    def try_move_player(self, d, pid=P1):
        return bool(self.L.orc_try_move(C.byref(self.e), pid, d))

    def spawn_bullet(self, d, pid=P1):
        return bool(self.L.orc_spawn_bullet(C.byref(self.e), pid, d))

    def spawn_wide_shot(self, pid, d):
        return bool(self.L.orc_spawn_wide_shot(C.byref(self.e), pid, d))

    def _reload_ammo(self):
        self.L.orc_reload_ammo(C.byref(self.e))

    def update_bullets(self):
        self.L.orc_update_bullets(C.byref(self.e))

    def step_players(self, a1, a2):
        self.L.orc_step_players(C.byref(self.e), a1, a2)

    def step(self, a1):
        self.step_players(a1, NONE)

    def is_round_over(self):
        return bool(self.L.orc_is_round_over(C.byref(self.e)))

    def get_winner(self):
        return {0: None, 1: P1, 2: P2}[self.L.orc_get_winner(C.byref(self.e))]

    def observation(self, viewer=P1):
        import numpy as np
        g = np.zeros((12, self.height, self.width), np.float32)
        x = np.zeros(4, np.float32)
        self.L.orc_build_obs(C.byref(self.e), viewer, g.ctypes.data, x.ctypes.data)
        return g, x


class CudaEngine:
    """The CUDA product behind core.py's method names: each call runs ONE engine method on the
    device through inv_debug_phase (the same __device__ functions the fused step kernel inlines).
    Fixed 15x10 board; field pokes go through inv_export_state / inv_import_state."""

    def __init__(self, width=15, height=10, seed=1234):
        from inversus_b200 import BatchedInversus, _capi
        assert width <= 15 and height <= 10
        self.capi = _capi
        self.sim = BatchedInversus(1, "selfplay", "hard", 500, seed=seed, auto_reset=False)
        self.width, self.height = 15, 10
        self.player_color = BLACK
        self.sim.reset()
        self._pull()

    def _pull(self):
        self._st = self.sim.export_state()
        self._dirty = False

    def _push(self):
        if self._dirty:
            self.sim.import_state(self._st)
            self._dirty = False

    def _call(self, phase, pid=0, arg=0, arg2=0):
        self._push()
        res = int(self.sim.debug_phase(phase, pid, arg, arg2).cpu()[0])
        assert self.sim.poll_status() == 0
        self._pull()
        return res

    def _pv(self, name):
        idx = {"x": 0, "y": 1, "ammo": 2, "reload_counter": 3, "alive": 4}

        def get(k):
            return int(self._st[name][0][idx[k]])

        def set_(k, v):
            self._st[name][0][idx[k]] = v
            self._dirty = True
        return _PlayerView(get, set_)

    @property
    def player1(self):
        return self._pv("p1")

    @property
    def player2(self):
        return self._pv("p2")

    @property
    def player_x(self):
        return int(self._st["p1"][0][0])

    @player_x.setter
    def player_x(self, v):
        self._st["p1"][0][0] = v
        self._dirty = True

    @property
    def player_y(self):
        return int(self._st["p1"][0][1])

    @player_y.setter
    def player_y(self, v):
        self._st["p1"][0][1] = v
        self._st["n_bullets"][0] = 0  # core.py:180-181
        self._dirty = True

    @property
    def bullets(self):
        n = int(self._st["n_bullets"][0])
        return [Bullet(*(int(v) for v in self._st["bullets"][0][i])) for i in range(n)]

    @bullets.setter
    def bullets(self, lst):
        self._st["n_bullets"][0] = len(lst)
        self._st["bullets"][0][:] = 0
        for i, b in enumerate(lst):
            self._st["bullets"][0][i] = (b.x, b.y, b.dir, b.owner)
        self._dirty = True

    def get_bullets(self):
        return self.bullets

    def _get_tile(self, x, y):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        i = y * 15 + x
        return int(self._st["tiles"][0][i >> 5] >> (i & 31)) & 1

    def _set_tile(self, x, y, c):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        i = y * 15 + x
        w = int(self._st["tiles"][0][i >> 5])
        w = (w | (1 << (i & 31))) if c else (w & ~(1 << (i & 31)))
        self._st["tiles"][0][i >> 5] = w
        self._dirty = True

    def reset(self):
        self._call(self.capi.PHASE_ENGINE_RESET)

    def try_move_player(self, d, pid=P1):
        return bool(self._call(self.capi.PHASE_TRY_MOVE, pid, d))

    def spawn_bullet(self, d, pid=P1):
        return bool(self._call(self.capi.PHASE_SPAWN_BULLET, pid, d))

    def spawn_wide_shot(self, pid, d):
        return bool(self._call(self.capi.PHASE_WIDE_SHOT, pid, d))

    def _reload_ammo(self):
        self._call(self.capi.PHASE_RELOAD)

    def update_bullets(self):
        self._call(self.capi.PHASE_UPDATE_BULLETS)

    def step_players(self, a1, a2):
        self._call(self.capi.PHASE_STEP_PLAYERS, 0, a1, a2)

    def step(self, a1):
        self.step_players(a1, NONE)

    def is_round_over(self):
        return not (self.player1.alive and self.player2.alive)

    def get_winner(self):
        a1, a2 = self.player1.alive, self.player2.alive
        if a1 and a2:
            return None
        return P2 if (a2 and not a1) else (P1 if (a1 and not a2) else None)

    def observation(self, viewer=P1):
        self._push()
        g, e = self.sim.obs_from_packed(self.sim.snapshot(), view=viewer, obs_dtype="f32")
        return g[0].cpu().numpy(), e[0].cpu().numpy()


class HostKernelEngine:
    """The product's engine functions compiled for the host (tests/host_kernel) behind core.py's
    method names: the CPU-side twin of CudaEngine, same device functions, no GPU needed."""

    def __init__(self, width=15, height=10, seed=1234):
        import ctypes as C_
        import numpy as np
        from backends import HostKernelBackend
        assert width <= 15 and height <= 10
        self.C, self.np = C_, np
        self.lib = HostKernelBackend.lib()
        self.b = HostKernelBackend(dict(n=1, mode="selfplay", difficulty="hard", max_steps=500, seed=seed))
        self.width, self.height = 15, 10
        self.player_color = BLACK
        self.b.reset(None)

    # packed-plane accessors
    def _pw(self, i):
        return int(self.b.planes[1, 0, 1 + i])

    def _set_pw(self, i, w):
        self.b.planes[1, 0, 1 + i] = w

    def _pv(self, i):
        fields = {"x": (0, 15), "y": (4, 15), "ammo": (8, 7), "reload_counter": (11, 31), "alive": (16, 1)}

        def get(k):
            sh, m = fields[k]
            return (self._pw(i) >> sh) & m

        def set_(k, v):
            sh, m = fields[k]
            self._set_pw(i, (self._pw(i) & ~(m << sh)) | ((int(v) & m) << sh))
        return _PlayerView(get, set_)

    @property
    def player1(self):
        return self._pv(0)

    @property
    def player2(self):
        return self._pv(1)

    player_x = property(lambda s: s.player1.x, lambda s, v: setattr(s.player1, "x", v))

    @property
    def player_y(self):
        return self.player1.y

    @player_y.setter
    def player_y(self, v):
        self.player1.y = v
        self.bullets = []  # core.py:180-181

    @property
    def bullets(self):
        n = (self._pw(0) >> 20) & 31
        words = [int(w) for w in list(self.b.planes[3, 0]) + list(self.b.planes[4, 0])]
        out = []
        for s in range(n):
            b = (words[s >> 1] >> ((s & 1) * 16)) & 0xFFFF
            out.append(Bullet(b & 15, (b >> 4) & 15, (b >> 8) & 3, (b >> 10) & 1))
        return out

    @bullets.setter
    def bullets(self, lst):
        words = [0] * 8
        for s, b in enumerate(lst):
            words[s >> 1] |= (b.x | (b.y << 4) | (b.dir << 8) | (b.owner << 10)) << ((s & 1) * 16)
        self.b.planes[3, 0] = words[:4]
        self.b.planes[4, 0] = words[4:]
        self._set_pw(0, (self._pw(0) & ~(31 << 20)) | (len(lst) << 20))

    def get_bullets(self):
        return self.bullets

    def _tile_word(self, i):
        return (0, i >> 5) if (i >> 5) < 4 else (1, 0)

    def _get_tile(self, x, y):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        i = y * 15 + x
        pl, w = self._tile_word(i)
        return (int(self.b.planes[pl, 0, w]) >> (i & 31)) & 1

    def _set_tile(self, x, y, c):
        if not (0 <= x < self.width and 0 <= y < self.height):
            raise IndexError
        i = y * 15 + x
        pl, w = self._tile_word(i)
        v = int(self.b.planes[pl, 0, w])
        self.b.planes[pl, 0, w] = (v | (1 << (i & 31))) if c else (v & ~(1 << (i & 31)))

    def _call(self, phase, pid=0, arg=0, arg2=0):
        C_, np = self.C, self.np
        res = np.zeros(1, np.uint8)
        rc = self.lib.hk_debug(self.b.planes.ctypes.data_as(C_.c_void_p), C_.c_int64(1), phase, pid, arg, arg2, 1,
                               C_.c_uint64(self.b.seed), C_.c_uint32(0), res.ctypes.data_as(C_.c_void_p),
                               self.b.status.ctypes.data_as(C_.c_void_p))
        assert rc == 0 and self.b.status[0] == 0
        return int(res[0])

    def reset(self):
        self._call(6)

    def try_move_player(self, d, pid=P1):
        return bool(self._call(0, pid, d))

    def spawn_bullet(self, d, pid=P1):
        return bool(self._call(1, pid, d))

    def spawn_wide_shot(self, pid, d):
        return bool(self._call(2, pid, d))

    def _reload_ammo(self):
        self._call(3)

    def update_bullets(self):
        self._call(4)

    def step_players(self, a1, a2):
        self._call(5, 0, a1, a2)

    def step(self, a1):
        self.step_players(a1, NONE)

    def is_round_over(self):
        return not (self.player1.alive and self.player2.alive)

    def get_winner(self):
        a1, a2 = self.player1.alive, self.player2.alive
        if a1 and a2:
            return None
        return P2 if (a2 and not a1) else (P1 if (a1 and not a2) else None)

    def observation(self, viewer=P1):
        C_, np = self.C, self.np
        g, e = np.zeros((12, 10, 15), np.float32), np.zeros(4, np.float32)
        self.lib.hk_observation(self.b.planes.ctypes.data_as(C_.c_void_p), C_.c_int64(1), C_.c_int64(0), viewer,
                                g.ctypes.data_as(C_.c_void_p), e.ctypes.data_as(C_.c_void_p))
        return g, e
