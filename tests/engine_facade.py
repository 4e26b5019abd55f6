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
