"""Drop-in for the reference's `inversus_rl.env_wrappers` rollout interface.

Same names, argument meaning and error behaviour as the reference
(inversus_rl/env_wrappers.py:20-66, :248-284, :447-528), numpy in / numpy out, so the rollout loops
of inversus_rl/training.py:119-157 and :287-325 run unchanged:

    from inversus_b200.env_wrappers import MultiEnvRunner
    env_runner = MultiEnvRunner(num_envs, opponent_type="dummy", max_episode_steps=500, difficulty="hard")
    obs_grid, obs_extra = env_runner.reset()
    next_obs, rewards, dones, infos = env_runner.step(actions)
    reset_obs = env_runner.envs[i].reset()

Every call runs the fused CUDA kernel through the C ABI and copies the results to fresh numpy
arrays (the reference's ownership contract: the trainer mutates `next_obs_grid[i]` in place,
training.py:150). For throughput use `BatchedInversus` directly (device tensors, no copies).
"""
from __future__ import annotations

import enum
from collections import namedtuple
from collections.abc import Sequence
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .constants import (BOARD_H, BOARD_W, INFO_GOT_HIT, INFO_LANDED_HIT, INFO_LOSE, INFO_WIN, STREAM_RESET,
                        TABLE_RESET_OFF, TABLE_STRIDE)
from .simulator import BatchedInversus

# (type, direction) names of the reference's Action dataclass (game_types.py:21-49)
Action = namedtuple("Action", "type direction")
_DIRS = ("up", "right", "down", "left")  # env_wrappers.py:24-37


def discrete_to_action(action_id: int) -> Action:
    """env_wrappers.py:20-66: 0 NONE; 1-4 MOVE; 5-8 SHOOT; 9-12 CHARGE_SHOOT; UP, RIGHT, DOWN, LEFT."""
    a = int(action_id)
    if a == 0:
        return Action("none", None)
    if 1 <= a <= 4:
        return Action("move", _DIRS[a - 1])
    if 5 <= a <= 8:
        return Action("shoot", _DIRS[a - 5])
    if 9 <= a <= 12:
        return Action("charge_shoot", _DIRS[a - 9])
    raise ValueError(f"Invalid action_id: {action_id}, must be 0-12")


def _check_actions(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    if a.size and (a.min() < 0 or a.max() > 12):
        bad = a[(a < 0) | (a > 12)][0]
        raise ValueError(f"Invalid action_id: {int(bad)}, must be 0-12")  # env_wrappers.py:66
    return np.ascontiguousarray(a, dtype=np.int8)


def _info_dict(bits: int, steps: int, ret: float) -> Dict[str, Any]:
    """Keys and types of the reference's info dict (env_wrappers.py:360-442)."""
    return {"landed_hit": bool(bits & INFO_LANDED_HIT), "got_hit": bool(bits & INFO_GOT_HIT),
            "win": bool(bits & INFO_WIN), "lose": bool(bits & INFO_LOSE),
            "episode_steps": int(steps), "episode_return": float(ret)}


class InfoList(Sequence):
    """list[dict]-compatible view over the info arrays; dicts are built on access so that a
    million-env step does not allocate a million dicts."""

    def __init__(self, bits, steps, rets):
        self._b, self._s, self._r = bits, steps, rets

    def __len__(self):
        return len(self._b)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return _info_dict(self._b[i], self._s[i], self._r[i])


class PlayerId(enum.Enum):
    """game_types.py:29-32."""
    P1 = 1
    P2 = 2


class _EngineView:
    """Stands in for `SingleInversusRLEnv.env` (the reference's InversusEnv): callers read the board
    size (policies.py:111-128 make_policy_from_env) and pass it to `build_observation`."""
    width, height = BOARD_W, BOARD_H

    def __init__(self, sim, index):
        self._sim, self._index = sim, index


def build_observation(env: "_EngineView", player_id=PlayerId.P1) -> Tuple[np.ndarray, np.ndarray]:
    """env_wrappers.py:173-245 for one env of a runner: (grid f32[12,10,15], extra f32[4]) from the
    given player's perspective, rebuilt from the env's current device state (K3 kernel)."""
    pid = getattr(player_id, "value", player_id)
    if pid not in (1, 2):
        raise ValueError(f"Invalid player ID: {player_id}")  # core.py:198
    sim, i = env._sim, env._index
    packed = sim.packed_state[:, i:i + 1].contiguous()
    grid, extra = sim.obs_from_packed(packed, view=pid - 1, obs_dtype="f32")
    return grid[0].cpu().numpy(), extra[0].cpu().numpy()


def _seeded_reset_draws(seed: int) -> np.ndarray:
    """The 42 spawn draws of one reset as a pure function of `seed` alone: Philox4x32-10 with key =
    the 64-bit seed and counter (0, 0, STREAM_RESET, block). Random numbers only -- the spawn
    logic that consumes them stays in the reset kernel (injected-draw mode of the C ABI)."""
    m32 = 0xFFFFFFFF
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    out = []
    for blk in range(11):  # 44 >= 2 + 2*20 draws (core.py:69-90)
        c0, c1, c2, c3 = 0, 0, STREAM_RESET, blk
        k0, k1 = s & m32, s >> 32
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
            c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & m32, p1 & m32, ((p0 >> 32) ^ c3 ^ k1) & m32, p0 & m32
            k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
        out += [c0, c1, c2, c3]
    return np.array(out, dtype=np.uint32)


def _reset_envs_seeded(sim: BatchedInversus, indices, seed: int) -> None:
    """`reset(seed=s)` of the reference (env_wrappers.py:272-276 -> core.py:64-67 reseeds the env's
    spawn generator): the new episode's spawn positions are a function of `s` only -- two envs reset
    with the same seed start identically, whatever the runner's own seed. Implemented with the
    C ABI's injected-draw table, so the game logic still runs in the kernel."""
    table = np.zeros((sim.num_envs, TABLE_STRIDE), np.uint32)
    table[:, TABLE_RESET_OFF:TABLE_RESET_OFF + 44] = _seeded_reset_draws(seed)
    sim.set_draw_table(table)
    try:
        sim.reset_envs(indices)
    finally:
        sim.set_draw_table(None)


class _EnvSlot:
    """`MultiEnvRunner.envs[i]`: supports what the trainer calls on it (training.py:149)."""

    def __init__(self, runner: "MultiEnvRunner", index: int):
        self._runner, self._index = runner, index
        self.env = _EngineView(runner.sim, index)
        self.opponent_type = runner.opponent_type
        self.difficulty = runner.difficulty
        self.max_episode_steps = runner.max_episode_steps

    def reset(self, seed: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        sim, i = self._runner.sim, self._index
        if seed is not None:  # env_wrappers.py:272-276
            _reset_envs_seeded(sim, [i], seed)
        else:
            sim.reset_envs([i])
        return sim.obs[i].float().cpu().numpy(), sim.extra[i].cpu().numpy()

    @property
    def step_count(self) -> int:
        return int(self._runner.sim.export_state(self._index, 1)["step_count"][0])


class MultiEnvRunner:
    """env_wrappers.py:447-528, backed by one fused CUDA kernel per step. No auto-reset: like the
    reference, finished envs are reset by the caller through `envs[i].reset()`."""

    def __init__(self, num_envs: int, opponent_type: str = "dummy", difficulty: str = "easy",
                 max_episode_steps: int = 500, seed: Optional[int] = None, *, device=0, env_id_base: int = 0,
                 reward_f64: bool = False):
        self.num_envs = num_envs
        self.opponent_type = opponent_type
        self.difficulty = difficulty
        self.max_episode_steps = max_episode_steps
        self.sim = BatchedInversus(num_envs, opponent_type, difficulty, max_episode_steps, seed, device=device,
                                   obs_dtype="f32", auto_reset=False, env_id_base=env_id_base, reward_f64=reward_f64)
        self.envs = [_EnvSlot(self, i) for i in range(num_envs)]
        self.episode_returns = [0.0] * num_envs
        self.episode_lengths = [0] * num_envs
        self.episode_wins = [0] * num_envs
        self.episode_losses = [0] * num_envs

    def reset(self) -> Tuple[np.ndarray, np.ndarray]:
        out = {"obs": np.empty((self.num_envs, 12, BOARD_H, BOARD_W), np.float32),
               "extra": np.empty((self.num_envs, 4), np.float32)}
        self.sim.reset_host(out)
        return out["obs"], out["extra"]

    def _opponent_actions(self, opponent_policy, opponent_actions):
        if opponent_actions is not None:
            return _check_actions(opponent_actions)
        if opponent_policy is None:
            raise ValueError("opponent_policy required for selfplay mode")  # env_wrappers.py:309
        # P2's view of the PRE-step state (env_wrappers.py:311) = what the last reset/step emitted
        g2 = self.sim.obs_p2.float().cpu().numpy()
        e2 = self.sim.extra_p2.cpu().numpy()
        if getattr(opponent_policy, "batched", False):
            return _check_actions(opponent_policy((g2, e2)))
        return _check_actions([opponent_policy((g2[i], e2[i])) for i in range(self.num_envs)])

    def step(self, action_ids, opponent_policy=None, *, opponent_actions=None):
        """Returns ((grid f32[N,12,10,15], extra f32[N,4]), rewards f32[N], dones bool[N], infos).

        `opponent_policy(obs_p2) -> action id` is called per env exactly like the reference
        (env_wrappers.py:311-314) unless it has a truthy `.batched` attribute, in which case it is
        called once with the stacked P2 observations; `opponent_actions` passes the ids directly.
        """
        a1 = _check_actions(action_ids)
        if a1.shape != (self.num_envs,):
            raise ValueError(f"action_ids must have shape ({self.num_envs},)")
        a2 = None
        if self.opponent_type == "selfplay":
            a2 = self._opponent_actions(opponent_policy, opponent_actions)
        n = self.num_envs
        out = {"obs": np.empty((n, 12, BOARD_H, BOARD_W), np.float32), "extra": np.empty((n, 4), np.float32),
               "reward": np.empty(n, np.float32), "done": np.empty(n, np.uint8), "info": np.empty(n, np.uint8),
               "episode_steps": np.empty(n, np.int32), "episode_return": np.empty(n, np.float64)}
        self.sim.step_host(a1, a2, out)
        dones = out["done"].astype(bool)
        for i in np.nonzero(dones)[0]:  # env_wrappers.py:513-519
            self.episode_returns[i] = float(out["episode_return"][i])
            self.episode_lengths[i] = int(out["episode_steps"][i])
            if out["info"][i] & INFO_WIN:
                self.episode_wins[i] += 1
            if out["info"][i] & INFO_LOSE:
                self.episode_losses[i] += 1
        infos = InfoList(out["info"], out["episode_steps"], out["episode_return"])
        return (out["obs"], out["extra"]), out["reward"], dones, infos


class SingleInversusRLEnv:
    """env_wrappers.py:248-444 for one env (used for shape probing, training.py:79-80, and play)."""

    def __init__(self, opponent_type: str = "dummy", difficulty: str = "easy", max_episode_steps: int = 500,
                 seed: Optional[int] = None, *, device=0):
        self._runner = MultiEnvRunner(1, opponent_type, difficulty, max_episode_steps, seed, device=device,
                                      reward_f64=True)
        self.opponent_type, self.difficulty, self.max_episode_steps = opponent_type, difficulty, max_episode_steps
        self.env = self._runner.envs[0].env
        self._runner.reset()

    def reset(self, seed: Optional[int] = None):
        if seed is not None:  # env_wrappers.py:272-276
            return self._runner.envs[0].reset(seed=seed)
        g, e = self._runner.reset()
        return g[0], e[0]

    def step(self, action_id: int, opponent_policy=None):
        (g, e), _, d, infos = self._runner.step(np.array([action_id]), opponent_policy)
        # the reference returns the binary64 reward here (env_wrappers.py:444); only MultiEnvRunner
        # casts to float32 (env_wrappers.py:525)
        return (g[0], e[0]), float(self._runner.sim.reward_f64[0]), bool(d[0]), infos[0]

    @property
    def step_count(self) -> int:
        return self._runner.envs[0].step_count
