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
