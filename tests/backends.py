"""Scenario backends (see tests/golden/scenarios.py) for the CPU oracle and the CUDA product."""
from __future__ import annotations

import ctypes as C

import numpy as np


class OracleBackend:
    """The C oracle (oracle/inversus_oracle.c) as a scenario backend. Checker only."""

    def __init__(self, sc, env_id_base=0, n=None, nthreads=1):
        from oracle import oracle as orc
        self.orc = orc
        self.b = orc.OracleBatch(n or sc["n"], sc["mode"], sc["difficulty"], sc["max_steps"],
                                 seed=sc["seed"], env_id_base=env_id_base, nthreads=nthreads)
        self.selfplay = sc["mode"] == "selfplay"

    def reset(self, table):
        self.b.reset(table)

    def step(self, a1, a2, table, auto_reset):
        b = self.b
        b.step(a1, a2, table, auto_reset)
        out = dict(reward=b.reward.copy(), done=b.done.copy(), flags=b.flags.copy(),
                   episode_steps=b.episode_steps.copy(), episode_return=b.episode_return.copy(),
                   obs1=b.obs1, extra1=b.extra1)
        if self.selfplay:
            out["obs2"], out["extra2"] = b.obs2, b.extra2
        return out

    def reset_envs(self, idx, table):
        L = self.orc.lib()
        for i in idx:
            e = self.b.envs[int(i)]
            if table is not None:
                row = np.ascontiguousarray(table[int(i)], np.uint32)
                e.table = row.ctypes.data_as(C.POINTER(C.c_uint32))
            L.orc_rl_reset(C.byref(e))
            e.table = None

    def state(self):
        return self.b.export_state()

    def obs(self):
        b = self.b
        L = self.orc.lib()
        o2 = np.zeros_like(b.obs1)
        e2 = np.zeros_like(b.extra1)
        for i in range(b.n):
            L.orc_build_obs(C.byref(b.envs[i]), 0, b.obs1[i].ctypes.data, b.extra1[i].ctypes.data)
            L.orc_build_obs(C.byref(b.envs[i]), 1, o2[i].ctypes.data, e2[i].ctypes.data)
        return b.obs1, b.extra1, o2, e2


class CudaBackend:
    """The CUDA product (through BatchedInversus -> C ABI) as a scenario backend."""

    def __init__(self, sc, env_id_base=0, n=None, obs_dtype="f32"):
        import torch
        from inversus_b200 import BatchedInversus
        self.torch = torch
        self.selfplay = sc["mode"] == "selfplay"
        self.sim = BatchedInversus(n or sc["n"], sc["mode"], sc["difficulty"], sc["max_steps"], seed=sc["seed"],
                                   auto_reset=(sc["resets"] == "auto"), env_id_base=env_id_base,
                                   obs_dtype=obs_dtype, p2_view=True)

    def _table(self, table):
        self.sim.set_draw_table(table)

    def reset(self, table):
        self._table(table)
        self.sim.reset()

    def step(self, a1, a2, table, auto_reset):
        assert auto_reset == self.sim.auto_reset
        self._table(table)
        s = self.sim
        s.step(self.torch.from_numpy(np.asarray(a1, np.int8)).cuda(),
               None if a2 is None else self.torch.from_numpy(np.asarray(a2, np.int8)).cuda())
        out = dict(reward=s.reward.cpu().numpy(), done=s.done.cpu().numpy(), flags=s.info.cpu().numpy(),
                   episode_steps=s.episode_steps.cpu().numpy(), episode_return=s.episode_return.cpu().numpy(),
                   obs1=s.obs.float().cpu().numpy(), extra1=s.extra.cpu().numpy())
        if self.selfplay:
            out["obs2"], out["extra2"] = s.obs_p2.float().cpu().numpy(), s.extra_p2.cpu().numpy()
        assert s.poll_status() == 0
        return out

    def reset_envs(self, idx, table):
        self._table(table)
        self.sim.reset_envs(np.asarray(idx, np.int64))

    def state(self):
        return self.sim.export_state()

    def obs(self):
        s = self.sim
        return (s.obs.float().cpu().numpy(), s.extra.cpu().numpy(),
                s.obs_p2.float().cpu().numpy(), s.extra_p2.cpu().numpy())


class PyLoopBackend:
    """The pure-Python restatement (oracle/py_loop.py) as a scenario backend. Checker only."""

    def __init__(self, sc, env_id_base=0):
        from oracle import py_loop
        self.pl = py_loop
        self.r = py_loop.PyRunner(sc["n"], sc["mode"], sc["difficulty"], sc["max_steps"], seed=sc["seed"],
                                  env_id_base=env_id_base)
        self.selfplay = sc["mode"] == "selfplay"

    def reset(self, table):
        self.r.reset(table)

    def step(self, a1, a2, table, auto_reset):
        return self.r.step(a1, a2, table, auto_reset, both_views=self.selfplay)

    def reset_envs(self, idx, table):
        for i in idx:
            self.r.envs[int(i)].reset(None if table is None else table[int(i)])

    def state(self):
        from oracle.oracle import STATE_DTYPE
        T = self.pl.Tile
        out = np.zeros(self.r.n, STATE_DTYPE)
        for i, e in enumerate(self.r.envs):
            bits = np.zeros(160, np.uint8)
            for y in range(10):
                for x in range(15):
                    bits[y * 15 + x] = 1 if e.grid[y][x] == T.WHITE else 0
            out["tiles"][i] = np.packbits(bits, bitorder="little").view(np.uint32)
            for name, p in (("p1", e.p1), ("p2", e.p2)):
                out[name][i] = (p.x, p.y, p.ammo, p.reload_counter, int(p.alive))
            out["n_bullets"][i] = len(e.bullets)
            for k, b in enumerate(e.bullets):
                out["bullets"][i, k] = (b.x, b.y, b.dir.value, b.owner)
            out["step_count"][i] = e.step_count
            out["episode"][i] = e.episode
            out["episode_return"][i] = e.episode_return
        return out

    def obs(self):
        o1, o2 = self.r.observations(0), self.r.observations(1)
        return o1[0], o1[1], o2[0], o2[1]
