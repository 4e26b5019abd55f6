"""Tensor-native host interface of the batched simulator: `BatchedInversus`.

Thin Python over the C ABI (include/inversus_b200.h). PyTorch is used for three things only: the
current CUDA stream, zero-copy views of the device buffers the extension owns, and device
placement of caller-provided action tensors. All game logic runs in the sm_100a kernels of
csrc/; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _capi
from .constants import BOARD_H, BOARD_W, OBS_CHANNELS, TABLE_STRIDE

_TORCH_OBS = {0: torch.float32, 1: torch.bfloat16, 2: torch.uint8}


class _DeviceArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.owner = owner  # keeps the handle (and thus the allocation) alive
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class _Handle:
    """Owns the inv_sim* and destroys it exactly once."""

    def __init__(self, cfg: _capi.Config):
        self.lib = _capi.load()
        h = C.c_void_p()
        _capi.check(self.lib.inv_create(C.byref(cfg), C.byref(h)))
        self.ptr = h

    def close(self):
        if self.ptr:
            self.lib.inv_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedInversus:
    """N INVERSUS environments advanced by one fused CUDA kernel per step.

    Mirrors `MultiEnvRunner(num_envs, opponent_type, difficulty, max_episode_steps, seed)`
    (inversus_rl/env_wrappers.py:450) with device tensors in and out. `auto_reset=True` fuses the
    trainer's reset-on-done (inversus_rl/training.py:140-151) into the step: `reward/done/info`
    describe the finished step, the returned observation is already the new episode's first.

    Returned tensors are views of buffers owned by the extension; they are overwritten by the next
    `reset`/`step` call (clone them to keep a value).
    """

    def __init__(self, num_envs: int, opponent_type: str = "dummy", difficulty: str = "easy",
                 max_episode_steps: int = 500, seed: Optional[int] = None, *, device=0,
                 obs_dtype: str = "f32", auto_reset: bool = True, env_id_base: int = 0,
                 p2_view: Optional[bool] = None, reward_f64: bool = False):
        if opponent_type not in _capi.MODE:
            raise ValueError(f"Unknown opponent_type: {opponent_type}")  # env_wrappers.py:316
        if difficulty not in _capi.DIFFICULTY:
            raise ValueError(f"Unknown difficulty: {difficulty}")
        if obs_dtype not in _capi.OBS_DTYPE:
            raise ValueError(f"Unknown obs_dtype: {obs_dtype}")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if seed is None:  # the reference leaves its envs unseeded (training.py:76)
            seed = int.from_bytes(os.urandom(8), "little")
        self.num_envs = int(num_envs)
        self.opponent_type = opponent_type
        self.difficulty = difficulty
        self.max_episode_steps = int(max_episode_steps)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.auto_reset = bool(auto_reset)
        self.env_id_base = int(env_id_base)
        flags = (_capi.FLAG_AUTO_RESET if auto_reset else 0)
        if p2_view or (p2_view is None and opponent_type == "selfplay"):
            flags |= _capi.FLAG_P2_VIEW
        if reward_f64:  # also keep the unrounded binary64 reward (what SingleInversusRLEnv.step returns)
            flags |= _capi.FLAG_REWARD_F64
        self._dt = _capi.OBS_DTYPE[obs_dtype]
        cfg = _capi.Config(self.num_envs, self.env_id_base, self.seed, _capi.MODE[opponent_type],
                           _capi.DIFFICULTY[difficulty], self.max_episode_steps, self.device.index,
                           self._dt, flags)
        self._h = _Handle(cfg)
        self._lib = self._h.lib
        self.has_p2_view = bool(flags & _capi.FLAG_P2_VIEW)
        n = self.num_envs
        grid = (n, OBS_CHANNELS, BOARD_H, BOARD_W)
        self.has_obs = self._dt != _capi.OBS_DTYPE["none"]  # "none": consumers read packed_state (policy encoder)
        self.obs = self._view(_capi.BUF_OBS_P1, grid, obs=True) if self.has_obs else None
        self.extra = self._view(_capi.BUF_EXTRA_P1, (n, 4), "<f4")
        self.obs_p2 = self._view(_capi.BUF_OBS_P2, grid, obs=True) if self.has_p2_view and self.has_obs else None
        self.extra_p2 = self._view(_capi.BUF_EXTRA_P2, (n, 4), "<f4") if self.has_p2_view else None
        self.reward = self._view(_capi.BUF_REWARD, (n,), "<f4")
        self.done = self._view(_capi.BUF_DONE, (n,), "|u1")
        self.info = self._view(_capi.BUF_INFO, (n,), "|u1")
        self.episode_steps = self._view(_capi.BUF_EPISODE_STEPS, (n,), "<i4")
        self.episode_return = self._view(_capi.BUF_EPISODE_RETURN, (n,), "<f8")
        self.reward_f64 = self._view(_capi.BUF_REWARD_F64, (n,), "<f8") if reward_f64 else None
        self.packed_state = self._view(_capi.BUF_PACKED_STATE, (5, n, 4), "<i4")
        self.debug_result = self._view(_capi.BUF_DEBUG_RESULT, (n,), "|u1")
        self._table = None

    # ------------------------------------------------------------------ plumbing
    def _view(self, which, shape, typestr=None, obs=False):
        p, nb = C.c_void_p(), C.c_int64()
        _capi.check(self._lib.inv_get_buffer(self._h.ptr, which, C.byref(p), C.byref(nb)))
        if obs:
            typestr = {0: "<f4", 1: "<i2", 2: "|u1"}[self._dt]
        t = torch.as_tensor(_DeviceArray(p.value, shape, typestr, self._h), device=self.device)
        if obs and self._dt == 1:
            t = t.view(torch.bfloat16)
        assert t.data_ptr() == p.value
        return t

    def _actions(self, a, name):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            t = a
            if t.device != self.device or t.dtype != torch.int8 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.int8).contiguous()
        else:
            t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.int8)).to(self.device, non_blocking=True)
        if t.numel() != self.num_envs:
            raise ValueError(f"{name} must have num_envs={self.num_envs} entries, got {t.numel()}")
        return t

    @property
    def launch_count(self) -> int:
        return int(self._lib.inv_launch_count(self._h.ptr))

    def close(self):
        self._h.close()

    # ------------------------------------------------------------------ MultiEnvRunner surface
    def reset(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """MultiEnvRunner.reset (env_wrappers.py:471-483)."""
        _capi.check(self._lib.inv_reset(self._h.ptr, _stream_ptr(self.device)))
        return self.obs, self.extra

    def step(self, action_ids, opponent_actions=None):
        """MultiEnvRunner.step (env_wrappers.py:485-528) on device tensors.

        action_ids: [N] ids in 0..12 (int8 CUDA tensor = zero-copy). opponent_actions: [N] ids of
        P2 in selfplay mode. Returns ((obs, extra), reward f32[N], done u8[N], info u8[N]) --
        info bits: 1 landed_hit, 2 got_hit, 4 win, 8 lose; `episode_steps` / `episode_return`
        are attributes. Out-of-range ids are reported by `poll_status()` (sticky), not raised here,
        because raising would need a device sync per step; the numpy front-ends check on the host.
        """
        a1 = self._actions(action_ids, "action_ids")
        a2 = self._actions(opponent_actions, "opponent_actions")
        if self.opponent_type == "selfplay" and a2 is None:
            raise ValueError("opponent_policy required for selfplay mode")  # env_wrappers.py:309
        self._keep = (a1, a2)  # keep inputs alive until the kernel has consumed them
        _capi.check(self._lib.inv_step(self._h.ptr, a1.data_ptr(), a2.data_ptr() if a2 is not None else None,
                                       _stream_ptr(self.device)))
        return (self.obs, self.extra), self.reward, self.done, self.info

    def reset_envs(self, indices) -> None:
        """MultiEnvRunner.envs[i].reset() for every i in `indices` (training.py:149)."""
        idx = torch.as_tensor(indices, dtype=torch.int64).reshape(-1).to(self.device).contiguous()
        if idx.numel() == 0:
            return
        self._keep_idx = idx
        _capi.check(self._lib.inv_reset_envs(self._h.ptr, idx.data_ptr(), idx.numel(), _stream_ptr(self.device)))

    # ------------------------------------------------------------------ host-buffer (numpy) calls
    def step_host(self, a1: np.ndarray, a2: Optional[np.ndarray], out: dict) -> None:
        """inv_step_host: numpy in, numpy out, synchronous. `out` maps any of obs/extra/obs_p2/
        extra_p2/reward/done/info/episode_steps/episode_return to preallocated arrays."""
        def p(k):
            v = out.get(k)
            return None if v is None else v.ctypes.data_as(C.c_void_p)
        a1 = np.ascontiguousarray(a1, np.int8)
        a2p = None
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, np.int8)
            a2p = a2.ctypes.data_as(C.c_void_p)
        if a1.size != self.num_envs:
            raise ValueError("action_ids has the wrong length")
        _capi.check(self._lib.inv_step_host(self._h.ptr, a1.ctypes.data_as(C.c_void_p), a2p, p("obs"), p("extra"),
                                            p("obs_p2"), p("extra_p2"), p("reward"), p("done"), p("info"),
                                            p("episode_steps"), p("episode_return")))

    def step_host_events(self, a1: np.ndarray, a2: Optional[np.ndarray], out: dict) -> np.ndarray:
        """inv_step_host_events: the step as a trainer with a GPU-resident policy consumes it
        (training.py:140-151). Observations -- grid and extra -- stay on the device; `out` maps
        reward/done/info to preallocated arrays (any may be missing) and "events" to an array of
        _capi.EVENT_DTYPE records. Returns the view of `out["events"]` holding the episodes that
        ended in this step, in env order."""
        def p(k):
            v = out.get(k)
            return None if v is None else v.ctypes.data_as(C.c_void_p)
        a1 = np.ascontiguousarray(a1, np.int8)
        a2p = None
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, np.int8)
            a2p = a2.ctypes.data_as(C.c_void_p)
        if a1.size != self.num_envs:
            raise ValueError("action_ids has the wrong length")
        ev = out["events"]
        if ev.dtype != _capi.EVENT_DTYPE or not ev.flags.c_contiguous:
            raise ValueError("out['events'] must be a contiguous array of _capi.EVENT_DTYPE")
        count = C.c_int64(0)
        _capi.check(self._lib.inv_step_host_events(self._h.ptr, a1.ctypes.data_as(C.c_void_p), a2p, p("reward"),
                                                   p("done"), p("info"), ev.ctypes.data_as(C.c_void_p), ev.size,
                                                   C.byref(count)))
        return ev[:count.value]

    def host_event_buffers(self, pinned: bool = True, capacity: Optional[int] = None) -> dict:
        """Output buffers for step_host_events (page-locked by default); capacity defaults to num_envs,
        which can never overflow."""
        n = self.num_envs
        cap = n if capacity is None else int(capacity)
        spec = {"reward": (n, np.float32), "done": (n, np.uint8), "info": (n, np.uint8)}
        out = {}
        for k, (count, dt) in spec.items():
            if pinned:
                t = torch.empty(count, dtype=torch.from_numpy(np.zeros(0, dt)).dtype, pin_memory=True)
                out[k] = t.numpy()
                out.setdefault("_pins", []).append(t)
            else:
                out[k] = np.empty(count, dt)
        if pinned:
            raw = torch.empty(max(cap, 1) * _capi.EVENT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            out["_pins"].append(raw)
            out["events"] = raw.numpy().view(_capi.EVENT_DTYPE)[:cap]
        else:
            out["events"] = np.empty(cap, _capi.EVENT_DTYPE)
        return out

    def reset_host(self, out: dict) -> None:
        def p(k):
            v = out.get(k)
            return None if v is None else v.ctypes.data_as(C.c_void_p)
        _capi.check(self._lib.inv_reset_host(self._h.ptr, p("obs"), p("extra"), p("obs_p2"), p("extra_p2")))

    def set_host_path(self, nthreads: Optional[int] = None, dma_fraction: float = -1.0) -> None:
        """How step_host delivers f32 observations (inv_set_host_path): nthreads=0 -> one plain
        PCIe copy; nthreads>0 -> packed rows over PCIe + host-side expansion on that many threads,
        with `dma_fraction` of the envs still copied directly (<0 = auto-balance)."""
        if nthreads is None:
            nthreads = min(os.cpu_count() or 1, 32)
        _capi.check(self._lib.inv_set_host_path(self._h.ptr, int(nthreads), float(dma_fraction)))

    def host_path(self) -> dict:
        nt, fr, td, te = C.c_int(), C.c_double(), C.c_double(), C.c_double()
        _capi.check(self._lib.inv_get_host_path(self._h.ptr, C.byref(nt), C.byref(fr), C.byref(td), C.byref(te)))
        return {"threads": nt.value, "dma_fraction": fr.value, "last_dma_s": td.value, "last_expand_s": te.value}

    def host_buffers(self, pinned: bool = True) -> dict:
        """Allocate one set of numpy output buffers for step_host (page-locked by default)."""
        n = self.num_envs
        np_obs = {0: np.float32, 1: np.uint16, 2: np.uint8}[self._dt]
        spec = {"obs": ((n, OBS_CHANNELS, BOARD_H, BOARD_W), np_obs), "extra": ((n, 4), np.float32),
                "reward": ((n,), np.float32), "done": ((n,), np.uint8), "info": ((n,), np.uint8),
                "episode_steps": ((n,), np.int32), "episode_return": ((n,), np.float64)}
        if self.has_p2_view:
            spec["obs_p2"] = spec["obs"]
            spec["extra_p2"] = spec["extra"]
        out = {}
        for k, (shape, dt) in spec.items():
            if pinned:
                t = torch.empty(shape, dtype=torch.from_numpy(np.zeros(0, dt)).dtype, pin_memory=True)
                out[k] = t.numpy()
                out.setdefault("_pins", []).append(t)
            else:
                out[k] = np.empty(shape, dt)
        return out

    # ------------------------------------------------------------------ parity surface
    def set_draw_table(self, table) -> None:
        """Injected-draw mode: [N, 64] uint32 draws for the NEXT reset/step call (None = Philox)."""
        if table is None:
            self._table = None
            _capi.check(self._lib.inv_set_draw_table(self._h.ptr, None))
            return
        t = torch.as_tensor(np.ascontiguousarray(table, np.uint32).view(np.int32)).to(self.device).contiguous()
        if tuple(t.shape) != (self.num_envs, TABLE_STRIDE):
            raise ValueError("draw table must be [num_envs, 64] uint32")
        self._table = t
        _capi.check(self._lib.inv_set_draw_table(self._h.ptr, t.data_ptr()))

    def export_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.num_envs - first if count is None else count
        out = np.zeros(count, _capi.STATE_DTYPE)
        torch.cuda.synchronize(self.device)
        _capi.check(self._lib.inv_export_state(self._h.ptr, out.ctypes.data_as(C.c_void_p), first, count))
        return out

    def import_state(self, state: np.ndarray, first: int = 0) -> None:
        st = np.ascontiguousarray(state, _capi.STATE_DTYPE)
        torch.cuda.synchronize(self.device)
        _capi.check(self._lib.inv_import_state(self._h.ptr, st.ctypes.data_as(C.c_void_p), first, len(st)))

    def snapshot(self) -> torch.Tensor:
        """Copy of the packed state planes ([5, N, 4] int32 = 80 B/env): what a rollout stores."""
        return self.packed_state.clone()

    def restore(self, packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Resume from `snapshot()`: loads the packed planes and rewrites the observation buffers.
        With the same seed the continuation is bit-identical to the run the snapshot came from
        (every draw is a function of seed, env id, episode, step and draw index)."""
        assert packed.is_cuda and packed.is_contiguous() and tuple(packed.shape) == (5, self.num_envs, 4)
        _capi.check(self._lib.inv_load_packed_state(self._h.ptr, packed.data_ptr(), _stream_ptr(self.device)))
        for view, (o, e) in enumerate(((self.obs, self.extra), (self.obs_p2, self.extra_p2))):
            if o is not None:
                _capi.check(self._lib.inv_obs_from_packed(self._h.ptr, self.packed_state.data_ptr(), self.num_envs,
                                                          self.num_envs, view, self._dt, o.data_ptr(), e.data_ptr(),
                                                          _stream_ptr(self.device)))
        return self.obs, self.extra

    def obs_from_packed(self, packed: torch.Tensor, view: int = 0, obs_dtype: Optional[str] = None,
                        count: Optional[int] = None):
        """Rebuild (obs, extra) from packed-state snapshots ([5, M, 4] int32 planes)."""
        assert packed.is_cuda and packed.is_contiguous() and packed.dim() == 3 and packed.shape[0] == 5
        stride = packed.shape[1]
        count = stride if count is None else count
        dt = self._dt if obs_dtype is None else _capi.OBS_DTYPE[obs_dtype]
        obs = torch.empty((count, OBS_CHANNELS, BOARD_H, BOARD_W), dtype=_TORCH_OBS[dt], device=self.device)
        extra = torch.empty((count, 4), dtype=torch.float32, device=self.device)
        _capi.check(self._lib.inv_obs_from_packed(self._h.ptr, packed.data_ptr(), stride, count, view, dt,
                                                  obs.data_ptr(), extra.data_ptr(), _stream_ptr(self.device)))
        return obs, extra

    def debug_phase(self, phase: int, pid: int = 0, arg: int = 0, arg2: int = 0) -> torch.Tensor:
        _capi.check(self._lib.inv_debug_phase(self._h.ptr, phase, pid, arg, arg2, _stream_ptr(self.device)))
        return self.debug_result

    def poll_status(self) -> int:
        """Synchronise and return+clear the sticky device status bits (constants.STATUS_*)."""
        bits = C.c_uint32()
        _capi.check(self._lib.inv_poll_status(self._h.ptr, _stream_ptr(self.device), C.byref(bits)))
        return int(bits.value)
