"""Pure-Python restatement of the reference's step loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Why a second oracle: the north star asks for "the reference's Python step loop timed on the box's
own host cores" beside every GPU number, and the reference itself cannot travel to the GPU box.
This module restates that loop with the reference's own data-structure choices -- Enum tiles in a
list of lists, dataclass players and bullets, `copy.deepcopy` of the grid inside the observation
builder (inversus/core.py:183-185 called from inversus_rl/env_wrappers.py:199), a 150-iteration
Python loop filling a numpy array -- so that its cost profile is the reference's (the survey
measured ~75 % of a reference step inside that deepcopy). It is also a third, independent
implementation of the rules: tests/test_py_loop_golden.py replays the committed golden fixtures
(outputs of the live reference) through it bit for bit, so it is PINNED like the C oracle.

Only tests/ and bench.py's cpu_baseline leg may import it. Small cases only: it runs at a few
thousand env-steps per second per core.
"""
from __future__ import annotations

import copy
import enum
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

W, H = 15, 10                       # config.py:7-8
MAX_AMMO, RELOAD_TICKS, WIDE_COST = 6, 30, 3   # config.py:14-16
STREAM_RESET = 0xFFFFFFFF
TABLE_RESET_OFF = 16
M32 = 0xFFFFFFFF


class Tile(enum.Enum):              # game_types.py:8-11
    BLACK = 0
    WHITE = 1


class Dir(enum.Enum):               # order of the action ids / obs channels (env_wrappers.py:24-37)
    UP = 0
    RIGHT = 1
    DOWN = 2
    LEFT = 3


DX = {Dir.UP: 0, Dir.RIGHT: 1, Dir.DOWN: 0, Dir.LEFT: -1}
DY = {Dir.UP: -1, Dir.RIGHT: 0, Dir.DOWN: 1, Dir.LEFT: 0}
DIRS = [Dir.UP, Dir.RIGHT, Dir.DOWN, Dir.LEFT]


@dataclass
class Player:                       # game_types.py:53-63
    pid: int
    x: int
    y: int
    color: Tile                     # the colour it can NOT stand on
    ammo: int = MAX_AMMO
    reload_counter: int = 0
    alive: bool = True


@dataclass
class Bullet:                       # game_types.py:66-71
    x: int
    y: int
    dir: Dir
    owner: int


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Draws:
    """The injected draw stream: Philox block per 4 draws, or a table row."""

    def __init__(self, seed, gid, episode, stream, table=None):
        self.key = (seed & M32, (seed >> 32) & M32)
        self.ctr = (gid, episode & M32, stream)
        self.table, self.k, self.block = table, 0, None

    def u32(self):
        k = self.k
        self.k += 1
        if self.table is not None:
            return int(self.table[k])
        if k & 3 == 0:
            self.block = philox4x32_10(self.ctr + (k >> 2,), self.key)
        return self.block[k & 3]

    def random(self):
        return self.u32() / 4294967296.0

    def below(self, n):
        return (self.u32() * n) >> 32

    def shuffle(self, x):
        for i in reversed(range(1, len(x))):
            j = self.below(i + 1)
            x[i], x[j] = x[j], x[i]


class PyEnv:
    """One environment: InversusEnv (core.py) + SingleInversusRLEnv (env_wrappers.py:248-444)."""

    def __init__(self, mode="dummy", difficulty="hard", max_episode_steps=500, seed=0, gid=0):
        self.mode, self.difficulty, self.max_steps = mode, difficulty, max_episode_steps
        self.seed, self.gid = seed, gid
        self.episode = -1
        self.grid: List[List[Tile]] = []
        self.p1: Optional[Player] = None
        self.p2: Optional[Player] = None
        self.bullets: List[Bullet] = []
        self.step_count = 0
        self.episode_return = 0.0
        self.prev_alive = [True, True]

    # ------------------------------------------------------------------ core.py
    def _paint_plus(self, cx, cy, color):
        for x, y in ((cx, cy), (cx + 1, cy), (cx - 1, cy), (cx, cy + 1), (cx, cy - 1)):
            if 0 <= x < W and 0 <= y < H:
                self.grid[y][x] = color

    def reset(self, table=None):                      # core.py:55-154 + env_wrappers.py:272-284
        self.episode += 1
        dr = Draws(self.seed, self.gid, self.episode, STREAM_RESET, None if table is None else table[TABLE_RESET_OFF:])
        self.grid = [[Tile.BLACK for _ in range(W)] for _ in range(H)]      # config.py:31
        self._paint_plus(1, 1, Tile.WHITE)                                   # config.py:34-54
        p1x, p1y = 1 + dr.below(W - 2), 1 + dr.below(H - 2)                  # core.py:69-70
        p2x = p2y = 0
        for _ in range(20):                                                  # core.py:85-90
            p2x, p2y = 1 + dr.below(W - 2), 1 + dr.below(H - 2)
            if abs(p2x - p1x) + abs(p2y - p1y) > 4:
                break
        self._paint_plus(p2x, p2y, Tile.WHITE)                               # core.py:96-108
        self._paint_plus(p1x, p1y, Tile.WHITE)                               # core.py:112-121
        self._paint_plus(p2x, p2y, Tile.BLACK)                               # core.py:136-146
        self.p1 = Player(0, p1x, p1y, Tile.BLACK)
        self.p2 = Player(1, p2x, p2y, Tile.WHITE)
        self.bullets = []
        self.step_count = 0
        self.episode_return = 0.0
        self.prev_alive = [True, True]

    def _try_move(self, pl, d):                       # core.py:249-296
        if not pl.alive:
            return
        nx, ny = pl.x + DX[d], pl.y + DY[d]
        if 0 <= nx < W and 0 <= ny < H and self.grid[ny][nx] != pl.color:
            pl.x, pl.y = nx, ny

    def _spawn_bullet(self, pl, d):                   # core.py:298-326
        if pl.alive and pl.ammo > 0:
            pl.ammo -= 1
            self.bullets.append(Bullet(pl.x, pl.y, d, pl.pid))

    def _spawn_wide(self, pl, d):                     # core.py:328-381
        if not pl.alive or pl.ammo < WIDE_COST:
            return
        pl.ammo -= WIDE_COST
        if d in (Dir.UP, Dir.DOWN):
            lanes = [(pl.x, pl.y), (pl.x - 1, pl.y), (pl.x + 1, pl.y)]
        else:
            lanes = [(pl.x, pl.y), (pl.x, pl.y - 1), (pl.x, pl.y + 1)]
        for x, y in lanes:
            if 0 <= x < W and 0 <= y < H:
                self.bullets.append(Bullet(x, y, d, pl.pid))

    def _apply(self, pl, a):                          # env_wrappers.py:20-66 + core.py:510-525
        if not pl.alive or a == 0:
            return
        d = DIRS[(a - 1) % 4]
        if a <= 4:
            self._try_move(pl, d)
        elif a <= 8:
            self._spawn_bullet(pl, d)
        else:
            self._spawn_wide(pl, d)

    def _reload(self):                                # core.py:383-397
        for pl in (self.p1, self.p2):
            if pl.alive and pl.ammo < MAX_AMMO:
                pl.reload_counter += 1
                if pl.reload_counter >= RELOAD_TICKS:
                    pl.ammo += 1
                    pl.reload_counter = 0

    def _update_bullets(self):                        # core.py:399-475
        targets = {}
        for b in self.bullets:
            nx, ny = b.x + DX[b.dir], b.y + DY[b.dir]
            if 0 <= nx < W and 0 <= ny < H:
                targets.setdefault((nx, ny), []).append(Bullet(nx, ny, b.dir, b.owner))
        kept = []
        for (x, y), here in targets.items():
            if len({b.owner for b in here}) > 1:
                continue
            b = here[0]
            color = self.p1.color if b.owner == 0 else self.p2.color
            if self.grid[y][x] == color:
                self.grid[y][x] = Tile.WHITE if color == Tile.BLACK else Tile.BLACK
            if self.p1.alive and b.owner != 0 and (x, y) == (self.p1.x, self.p1.y):
                self.p1.alive = False
            if self.p2.alive and b.owner != 1 and (x, y) == (self.p2.x, self.p2.y):
                self.p2.alive = False
            kept.append(b)
        self.bullets = kept

    # ------------------------------------------------------------------ env_wrappers.py
    def _p2_can_step(self, d):
        nx, ny = self.p2.x + DX[d], self.p2.y + DY[d]
        return 0 <= nx < W and 0 <= ny < H and self.grid[ny][nx] != self.p2.color

    def _dummy(self, dr):                             # env_wrappers.py:69-170
        p1, p2 = self.p1, self.p2
        if not p2.alive:
            return 0
        if self.difficulty == "easy":
            move_prob, shoot_prob, random_move_prob = 0.001, 0.0, 0.0
        else:
            move_prob, shoot_prob, random_move_prob = 0.9, 0.2, 0.05
        x_al, y_al = p2.x == p1.x, p2.y == p1.y
        if dr.random() < shoot_prob and p2.ammo > 0 and (x_al or y_al):
            if x_al:
                return 5 + (Dir.UP if p1.y < p2.y else Dir.DOWN).value
            return 5 + (Dir.LEFT if p1.x < p2.x else Dir.RIGHT).value
        dirs = [Dir.UP, Dir.DOWN, Dir.LEFT, Dir.RIGHT]
        if dr.random() < random_move_prob:
            dr.shuffle(dirs)
            if self._p2_can_step(dirs[0]):
                return 1 + dirs[0].value
        if self.difficulty == "easy" and dr.random() > move_prob:
            return 0
        dx, dy = p1.x - p2.x, p1.y - p2.y
        cands = []
        if dx != 0:
            cands.append(Dir.RIGHT if dx > 0 else Dir.LEFT)
        if dy != 0:
            cands.append(Dir.DOWN if dy > 0 else Dir.UP)
        dr.shuffle(cands)
        for d in cands:
            if self._p2_can_step(d):
                return 1 + d.value
        dr.shuffle(dirs)
        for d in dirs:
            if self._p2_can_step(d):
                return 1 + d.value
        return 0

    def observation(self, viewer=0):                  # env_wrappers.py:173-245
        g = np.zeros((12, H, W), dtype=np.float32)
        me, en = (self.p1, self.p2) if viewer == 0 else (self.p2, self.p1)
        grid = copy.deepcopy(self.grid)               # core.py:183-185 via env_wrappers.py:199
        for y in range(H):
            for x in range(W):
                if grid[y][x] == Tile.BLACK:
                    g[0, y, x] = 1.0
                else:
                    g[1, y, x] = 1.0
        if me.alive:
            g[2, me.y, me.x] = 1.0
        if en.alive:
            g[3, en.y, en.x] = 1.0
        for b in self.bullets:
            g[(4 if b.owner == viewer else 8) + b.dir.value, b.y, b.x] = 1.0
        extra = np.array([me.ammo / MAX_AMMO if me.alive else 0.0, en.ammo / MAX_AMMO if en.alive else 0.0,
                          1.0 if me.alive else 0.0, 1.0 if en.alive else 0.0], dtype=np.float32)
        return g, extra

    def step(self, a1, a2=None, table=None):          # env_wrappers.py:286-444
        if not 0 <= a1 <= 12:
            raise ValueError(f"Invalid action_id: {a1}, must be 0-12")
        if self.mode == "dummy":
            a2 = self._dummy(Draws(self.seed, self.gid, self.episode, self.step_count, table))
        elif a2 is None:
            raise ValueError("opponent_policy required for selfplay mode")
        elif not 0 <= a2 <= 12:
            raise ValueError(f"Invalid action_id: {a2}, must be 0-12")
        prev1, prev2 = self.prev_alive
        prev_white = sum(row.count(Tile.WHITE) for row in self.grid)
        self._apply(self.p1, a1)                      # core.py:497-531
        self._apply(self.p2, a2)
        self._reload()
        self._update_bullets()
        self.step_count += 1
        p1, p2 = self.p1, self.p2
        reward, done, flags = 0.0, False, 0
        diff = sum(row.count(Tile.WHITE) for row in self.grid) - prev_white
        if diff > 0:
            reward += diff * 0.01
        if prev2 and not p2.alive:
            reward += 1.0
            flags |= 1
        if prev1 and not p1.alive:
            reward -= 0.01
            flags |= 2
        if p1.alive and p1.ammo == 0:
            reward -= 0.001
        if p1.alive and p2.alive:
            dist = abs(p1.x - p2.x) + abs(p1.y - p2.y)
            reward += 0.002 * (1.0 - dist / (W + H))
            aligned = p1.x == p2.x or p1.y == p2.y
            if aligned:
                reward += 0.002
            if 5 <= a1 <= 12 and aligned and p1.ammo > 0:
                sd = DIRS[(a1 - 1) % 4]
                aim = False
                if p1.x == p2.x:
                    aim = (p1.y < p2.y and sd == Dir.DOWN) or (p1.y > p2.y and sd == Dir.UP)
                elif p1.y == p2.y:
                    aim = (p1.x < p2.x and sd == Dir.RIGHT) or (p1.x > p2.x and sd == Dir.LEFT)
                if aim:
                    reward += 0.05
        over = not (p1.alive and p2.alive)
        if over:
            done = True
            if p1.alive:
                reward += 10.0
                flags |= 4
            elif p2.alive:
                reward -= 0.1
                flags |= 8
        else:
            reward -= 0.001
        self.prev_alive = [p1.alive, p2.alive]
        if self.step_count >= self.max_steps:
            done = True
            if not over:
                reward -= 2.0
        self.episode_return += reward
        return reward, done, flags


class PyRunner:
    """MultiEnvRunner (env_wrappers.py:447-528): a sequential Python loop over the envs."""

    def __init__(self, n, mode="dummy", difficulty="hard", max_episode_steps=500, seed=0, env_id_base=0):
        self.envs = [PyEnv(mode, difficulty, max_episode_steps, seed, env_id_base + i) for i in range(n)]
        self.n, self.mode = n, mode

    def reset(self, table=None):
        for i, e in enumerate(self.envs):
            e.reset(None if table is None else table[i])
        return self.observations(0)

    def observations(self, viewer=0):
        obs = [e.observation(viewer) for e in self.envs]
        return np.stack([o[0] for o in obs]), np.stack([o[1] for o in obs])

    def step(self, a1, a2=None, table=None, auto_reset=False, both_views=False):
        n = self.n
        out = dict(reward=np.zeros(n, np.float64), done=np.zeros(n, np.uint8), flags=np.zeros(n, np.uint8),
                   episode_steps=np.zeros(n, np.int32), episode_return=np.zeros(n, np.float64))
        grids, extras, grids2, extras2 = [], [], [], []
        for i, e in enumerate(self.envs):
            row = None if table is None else table[i]
            r, d, f = e.step(int(a1[i]), None if a2 is None else int(a2[i]), row)
            out["reward"][i], out["done"][i], out["flags"][i] = r, d, f
            out["episode_steps"][i], out["episode_return"][i] = e.step_count, e.episode_return
            if d and auto_reset:                      # training.py:148-151
                e.reset(row)
            g, x = e.observation(0)
            grids.append(g)
            extras.append(x)
            if both_views:
                g2, x2 = e.observation(1)
                grids2.append(g2)
                extras2.append(x2)
        out["obs1"], out["extra1"] = np.stack(grids), np.stack(extras)
        if both_views:
            out["obs2"], out["extra2"] = np.stack(grids2), np.stack(extras2)
        return out
