"""PPO agent with a device-resident rollout: the "next" row of the hot path.

Same constructor, hyper-parameters, loss and update schedule as the reference's `PPOAgent`
(inversus_rl/ppo_agent.py:13-247): clipped surrogate + value_coef * MSE - entropy_coef * entropy,
`epochs` passes over shuffled minibatches of `batch_size`, grad-norm clip 0.5, Adam. What changes
is where the data lives:

* `act` takes the simulator's device observations and returns device tensors (numpy in -> numpy
  out is kept for the reference's calling convention, ppo_agent.py:68-106);
* the rollout is stored time-major `[T, N]` on the device (`DeviceRollout`), observations as
  80-byte packed-state snapshots that are decoded per minibatch by the simulator's K3 kernel;
* GAE runs as one CUDA kernel (`inv_gae`), per env. The reference runs GAE over the flat
  `[t0e0, t0e1, ..., t1e0, ...]` list so that for num_envs > 1 `values[t+1]` belongs to a
  different env (ppo_agent.py:145-152 vs training.py:128-137); `gae_mode="reference"` reproduces
  that bit for bit, the default `"per_env"` is the textbook estimator;
* under `torch.distributed` the gradients are averaged with one flat NCCL all-reduce per
  minibatch (10.25 M parameters = 41 MB fp32).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi


def gae_numpy(rewards: np.ndarray, values: np.ndarray, dones: np.ndarray, last_value, gamma: float, lam: float):
    """The reference's loop (ppo_agent.py:139-154) over [T, N] float32 arrays, per column.
    Used for CPU tensors only (the reference's default device); CUDA tensors go through inv_gae."""
    T, N = rewards.shape
    adv = np.zeros_like(rewards, dtype=np.float32)
    g, gl = np.float32(gamma), np.float32(gamma * lam)
    nxt = np.zeros(N, np.float32) if last_value is None else np.asarray(last_value, np.float32).reshape(N).copy()
    last = np.zeros(N, np.float32)
    for t in range(T - 1, -1, -1):
        r, v, d = rewards[t].astype(np.float32), values[t].astype(np.float32), dones[t].astype(bool)
        delta_run = (r + g * nxt) - v
        last = np.where(d, r - v, delta_run + gl * last).astype(np.float32)
        adv[t] = last
        nxt = v
    return adv, adv + values.astype(np.float32)


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, last_value: Optional[torch.Tensor],
                gamma: float, lam: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """GAE over a time-major [T, N] rollout. CUDA tensors: one `inv_gae` kernel launch."""
    T, N = rewards.shape
    if not rewards.is_cuda:
        a, r = gae_numpy(rewards.numpy(), values.numpy(), dones.numpy(),
                         None if last_value is None else last_value.numpy(), gamma, lam)
        return torch.from_numpy(a), torch.from_numpy(r)
    rewards = rewards.contiguous().float()
    values = values.contiguous().float()
    dones = dones.contiguous().to(torch.uint8)
    lv = None if last_value is None else last_value.contiguous().float()
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        _capi.check(_capi.load().inv_gae(rewards.data_ptr(), values.data_ptr(), dones.data_ptr(),
                                         None if lv is None else lv.data_ptr(), gamma, lam, T, N,
                                         adv.data_ptr(), ret.data_ptr(),
                                         int(torch.cuda.current_stream(rewards.device).cuda_stream)))
    return adv, ret


class DeviceRollout:
    """Time-major rollout storage on the device. Observations are kept either as packed-state
    snapshots (80 B per env-step; decoded per minibatch through `sim.obs_from_packed`) or as the
    observation tensors themselves (`store="obs"`, small runs and parity tests)."""

    def __init__(self, T: int, N: int, device, store: str = "packed", obs_dtype=torch.float32):
        self.T, self.N, self.device, self.store = T, N, torch.device(device), store
        self.t = 0
        if store == "packed":
            self.packed = torch.empty((5, T, N, 4), dtype=torch.int32, device=device)
        else:
            self.obs = torch.empty((T, N, 12, 10, 15), dtype=obs_dtype, device=device)
            self.extra = torch.empty((T, N, 4), dtype=torch.float32, device=device)
        self.actions = torch.empty((T, N), dtype=torch.int64, device=device)
        self.log_probs = torch.empty((T, N), dtype=torch.float32, device=device)
        self.values = torch.empty((T, N), dtype=torch.float32, device=device)
        self.rewards = torch.empty((T, N), dtype=torch.float32, device=device)
        self.dones = torch.empty((T, N), dtype=torch.uint8, device=device)

    def reset(self):
        self.t = 0

    @property
    def full(self) -> bool:
        return self.t >= self.T

    def store_pre(self, sim_or_obs, actions, log_probs, values):
        """Before the env step: what the policy saw (state snapshot or obs) and what it did."""
        t = self.t
        if self.store == "packed":
            self.packed[:, t].copy_(sim_or_obs.packed_state)
        else:
            obs, extra = sim_or_obs
            self.obs[t].copy_(obs)
            self.extra[t].copy_(extra)
        self.actions[t].copy_(actions)
        self.log_probs[t].copy_(log_probs)
        self.values[t].copy_(values)

    def store_post(self, rewards, dones):
        """After the env step: its reward and done flag; advances the time index."""
        self.rewards[self.t].copy_(rewards)
        self.dones[self.t].copy_(dones)
        self.t += 1

    # CUDA-graph friendly variants: the time index lives on the device (a 1-element int64 tensor
    # that the captured graph increments), so one captured step can be replayed T times.
    def store_pre_at(self, t_dev: torch.Tensor, sim, actions, log_probs, values):
        assert self.store == "packed"
        self.packed.index_copy_(1, t_dev, sim.packed_state.unsqueeze(1))
        self.actions.index_copy_(0, t_dev, actions.unsqueeze(0))
        self.log_probs.index_copy_(0, t_dev, log_probs.unsqueeze(0))
        self.values.index_copy_(0, t_dev, values.unsqueeze(0))

    def store_post_at(self, t_dev: torch.Tensor, rewards, dones):
        self.rewards.index_copy_(0, t_dev, rewards.unsqueeze(0))
        self.dones.index_copy_(0, t_dev, dones.unsqueeze(0))

    def minibatch_obs(self, idx: torch.Tensor, sim, obs_dtype: Optional[str] = None):
        """Observations of flat sample indices `idx` (m = t*N + n)."""
        if self.store == "packed":
            T = self.t
            flat = self.packed[:, :T].reshape(5, T * self.N, 4)
            sel = flat.index_select(1, idx).contiguous()
            if obs_dtype == "packed":  # the policy's encoder reads the 80-byte states directly
                from .fused_ops import PackedStates
                return PackedStates(sel, 0), None
            return sim.obs_from_packed(sel, view=0, obs_dtype=obs_dtype)
        T = self.t
        return (self.obs[:T].reshape(T * self.N, 12, 10, 15).index_select(0, idx),
                self.extra[:T].reshape(T * self.N, 4).index_select(0, idx))


class PPOAgent:
    """ppo_agent.py:13 with device-resident data. Constructor arguments as in the reference;
    keyword-only extras: `gae_mode` ("per_env" | "reference"), `precision` ("fp32" | "bf16":
    bf16 autocast for inference and the training forward, fp32 master weights and optimiser),
    `shuffle` ("torch" | "numpy": the reference shuffles with np.random, ppo_agent.py:195)."""

    def __init__(self, policy: nn.Module, lr: float = 1e-4, gamma: float = 0.99, lam: float = 0.95,
                 clip_ratio: float = 0.2, epochs: int = 4, batch_size: int = 512, entropy_coef: float = 0.02,
                 value_coef: float = 0.1, device: str = "cpu", *, gae_mode: str = "per_env",
                 precision: str = "fp32", shuffle: str = "torch", generator: Optional[torch.Generator] = None,
                 graph_update: bool = False, packed_encoder: bool = False):
        assert gae_mode in ("per_env", "reference") and precision in ("fp32", "bf16") and shuffle in ("torch", "numpy")
        self.device = torch.device(device)
        self.policy = policy.to(self.device)
        # same Adam as the reference (ppo_agent.py:48); on CUDA the single-kernel fused implementation
        cuda = self.device.type == "cuda"
        self.graph_update = bool(graph_update) and cuda  # capture one minibatch step in a CUDA graph
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=lr, fused=cuda, capturable=self.graph_update)
        self._ug = None  # captured update graph + its static buffers
        self.gamma, self.lam, self.clip_ratio = gamma, lam, clip_ratio
        self.epochs, self.batch_size = epochs, batch_size
        self.entropy_coef, self.value_coef = entropy_coef, value_coef
        self.gae_mode, self.precision, self.shuffle, self.generator = gae_mode, precision, shuffle, generator
        self.max_grad_norm = 0.5  # ppo_agent.py:228
        self.act_chunk = 65536    # samples per inference micro-batch in act()
        # packed_encoder: the policy's first block reads packed env states (csrc/encoder_kernels.cu);
        # the update then never materialises observations
        self.packed_encoder = bool(packed_encoder) and cuda and precision == "bf16"
        self._flat_grad = None    # all gradients as views of one buffer (built on first use, CUDA only)
        self.grad_buckets = 4     # NCCL all-reduces per minibatch, launched while backward still runs
        self.comm_dtype = torch.float32  # torch.bfloat16 halves the all-reduce bytes (fp32 accumulation in Adam)
        self.allreduce_ms = None  # filled by time_allreduce()
        self.reset_buffers()

    # ------------------------------------------------------------------ reference-shaped list buffers
    def reset_buffers(self) -> None:
        self.obs_grid_buffer: List = []
        self.obs_extra_buffer: List = []
        self.action_buffer: List[int] = []
        self.log_prob_buffer: List[float] = []
        self.reward_buffer: List[float] = []
        self.value_buffer: List[float] = []
        self.done_buffer: List[bool] = []

    def store_step(self, grid_tensor, extra_vector, action, log_prob, value, reward, done) -> None:
        """ppo_agent.py:108-125 (one sample at a time, kept for the reference's rollout loop)."""
        self.obs_grid_buffer.append(grid_tensor)
        self.obs_extra_buffer.append(extra_vector)
        self.action_buffer.append(action)
        self.log_prob_buffer.append(log_prob)
        self.reward_buffer.append(reward)
        self.value_buffer.append(value)
        self.done_buffer.append(done)

    def compute_advantages(self, last_value: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
        """ppo_agent.py:127-157 on the list buffers: GAE over the flat list (N = 1, T = len)."""
        r = torch.tensor(self.reward_buffer, dtype=torch.float32, device=self.device).view(-1, 1)
        v = torch.tensor(self.value_buffer, dtype=torch.float32, device=self.device).view(-1, 1)
        d = torch.tensor(self.done_buffer, dtype=torch.uint8, device=self.device).view(-1, 1)
        lv = torch.tensor([last_value], dtype=torch.float32, device=self.device)
        adv, ret = compute_gae(r, v, d, lv, self.gamma, self.lam)
        return adv.view(-1).cpu().numpy(), ret.view(-1).cpu().numpy()

    # ------------------------------------------------------------------ acting
    def _forward(self, grid, extra, train: bool):
        if not isinstance(grid, torch.Tensor):  # fused_ops.PackedStates: always the bf16 tensor-core path
            return self.policy.forward_bf16(grid, None) if train else self.policy.infer(grid, None)
        if self.precision == "bf16" and grid.is_cuda:
            return self.policy.forward_bf16(grid, extra) if train else self.policy.infer(grid, extra)
        return self.policy(grid.float() if grid.dtype != torch.float32 else grid, extra)

    def act(self, grid_tensors, extra_vectors):
        """ppo_agent.py:68-106. numpy in -> numpy (actions, log_probs, values) like the reference;
        tensors in -> device tensors (no host round trip, no sync)."""
        self.policy.eval()
        packed = hasattr(grid_tensors, "planes")  # fused_ops.PackedStates (extra_vectors is ignored)
        as_numpy = not packed and not isinstance(grid_tensors, torch.Tensor)
        with torch.no_grad():
            if packed:
                grid, extra = grid_tensors, None
            else:
                grid = torch.as_tensor(grid_tensors).to(self.device)
                extra = torch.as_tensor(extra_vectors).to(self.device)
            n = grid.shape[0]
            if n > self.act_chunk:  # bound activation memory: 1M envs x 19200 features would be 38 GB per layer
                def part(i):
                    j = min(i + self.act_chunk, n)
                    return (grid.chunk(i, j), None) if packed else (grid[i:j], extra[i:j])
                parts = [self._forward(*part(i), train=False)
                         for i in range(0, n, self.act_chunk)]
                logits = torch.cat([p[0] for p in parts])
                values = torch.cat([p[1] for p in parts])
            else:
                logits, values = self._forward(grid, extra, train=False)
            if as_numpy:  # the reference's exact sampling ops (ppo_agent.py:94-97)
                dist = torch.distributions.Categorical(logits=logits)
                actions = dist.sample()
                log_probs = dist.log_prob(actions)
            else:  # same distribution without Categorical's argument validation (a host sync):
                logp = F.log_softmax(logits, dim=-1)  # capturable in a CUDA graph
                actions = torch.multinomial(logp.exp(), 1).squeeze(1)
                log_probs = logp.gather(1, actions.unsqueeze(1)).squeeze(1)
            values = values.squeeze(-1)
        if as_numpy:
            return actions.cpu().numpy(), log_probs.cpu().numpy(), values.cpu().numpy()
        return actions, log_probs, values

    # ------------------------------------------------------------------ update
    @staticmethod
    def _world() -> int:
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def _setup_flat_grads(self) -> None:
        """Every `p.grad` becomes a view of ONE persistent buffer, laid out in the order gradients
        become ready in backward (heads first: their two 19204x256 matrices are 96 % of the bytes),
        cut into `grad_buckets` contiguous buckets. Zeroing the gradients is one memset; under
        torch.distributed each bucket is all-reduced (averaged) by NCCL as soon as its last
        gradient has been accumulated, while backward is still running on the trunk."""
        params = [p for p in reversed(list(self.policy.parameters())) if p.requires_grad]
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        target = -(-total // max(int(self.grad_buckets), 1))
        buckets, off, start, members = [], 0, 0, []
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            members.append(p)
            if off - start >= target or off == total:
                buckets.append({"lo": start, "hi": off, "n": len(members), "pending": len(members), "work": None})
                for q in members:
                    q._inv_bucket = len(buckets) - 1
                start, members = off, []
        self._flat_grad, self._buckets, self._sync_active = flat, buckets, False
        for p in params:
            p.register_post_accumulate_grad_hook(self._grad_ready)

    def _launch_bucket(self, b: dict) -> None:
        d = torch.distributed
        view = self._flat_grad[b["lo"]:b["hi"]]
        if self.comm_dtype != view.dtype:  # narrower wire format: convert, reduce, convert back
            wire = view.to(self.comm_dtype)
            d.all_reduce(wire, op=d.ReduceOp.SUM)
            view.copy_(wire)
            b["work"] = None
            return
        avg = d.get_backend() == "nccl"
        b["avg"] = avg
        b["work"] = d.all_reduce(view, op=d.ReduceOp.AVG if avg else d.ReduceOp.SUM, async_op=True)

    def _grad_ready(self, p) -> None:
        if not self._sync_active:
            return
        b = self._buckets[p._inv_bucket]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch_bucket(b)

    def _begin_grad_sync(self) -> None:
        self._sync_active = self._world() > 1
        for b in self._buckets:
            b["pending"], b["work"], b["avg"] = b["n"], None, True

    def _finish_grad_sync(self) -> None:
        """Wait for the bucket all-reduces launched from the backward hooks (and launch any bucket
        whose gradients did not all arrive through a hook); leaves averaged gradients in place."""
        if not self._sync_active:
            return
        self._sync_active = False
        world = self._world()
        for b in self._buckets:
            if b["pending"] != 0:
                self._launch_bucket(b)
        for b in self._buckets:
            if b["work"] is not None:
                b["work"].wait()
            if not b["avg"] or self.comm_dtype != self._flat_grad.dtype:
                self._flat_grad[b["lo"]:b["hi"]].div_(world)

    def _sync_grads(self) -> None:
        """Blocking gradient average without the flat buffer (kept for callers that own .grad)."""
        if self._world() == 1:
            return
        world = self._world()
        for p in self.policy.parameters():
            if p.grad is not None:
                torch.distributed.all_reduce(p.grad, op=torch.distributed.ReduceOp.SUM)
                p.grad.div_(world)

    def time_allreduce(self, iters: int = 20) -> Optional[dict]:
        """Time the gradient all-reduce alone (all buckets, back to back): ms per minibatch and bytes."""
        if self._world() == 1 or self._flat_grad is None:
            return None
        d = torch.distributed
        dev = self._flat_grad.device
        scratch = torch.zeros_like(self._flat_grad)
        views = [scratch[b["lo"]:b["hi"]] for b in self._buckets]
        for _ in range(3):
            for v in views:
                d.all_reduce(v)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                for v in views:
                    d.all_reduce(v)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / iters
        else:
            import time as _t
            t0 = _t.perf_counter()
            for _ in range(iters):
                for v in views:
                    d.all_reduce(v)
            ms = (_t.perf_counter() - t0) * 1e3 / iters
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        d.all_reduce(t, op=d.ReduceOp.MAX)
        self.allreduce_ms = float(t.item())
        return {"allreduce_ms_per_minibatch": self.allreduce_ms, "bytes": scratch.numel() * scratch.element_size(),
                "buckets": len(views), "dtype": str(scratch.dtype).replace("torch.", "")}

    def _permutation(self, n: int, state: dict) -> torch.Tensor:
        if self.shuffle == "numpy":
            # the reference shuffles ONE index array in place every epoch (ppo_agent.py:192-195),
            # so epoch k's order is a shuffle of epoch k-1's
            if "idx" not in state:
                state["idx"] = np.arange(n)
            np.random.shuffle(state["idx"])
            return torch.from_numpy(state["idx"].copy()).to(self.device)
        return torch.randperm(n, device=self.device, generator=self.generator)

    def _minibatch_step(self, idx, fetch, actions, old_log_probs, advantages, returns, sums, validate=True):
        """One optimisation step on the samples `idx` (ppo_agent.py:198-231); adds the three loss
        statistics to `sums`. With validate=False it contains no host sync (CUDA-graph capturable)."""
        grid, extra = fetch(idx)
        logits, values = self._forward(grid, extra, train=True)
        dist = torch.distributions.Categorical(logits=logits, validate_args=validate)  # ppo_agent.py:211-213
        new_log_probs = dist.log_prob(actions[idx])
        entropy = dist.entropy().mean()
        ratio = torch.exp(new_log_probs - old_log_probs[idx])
        adv = advantages[idx]
        policy_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1.0 - self.clip_ratio, 1.0 + self.clip_ratio) * adv).mean()
        value_loss = F.mse_loss(values.squeeze(-1), returns[idx])
        loss = policy_loss + self.value_coef * value_loss - self.entropy_coef * entropy
        if self._flat_grad is None and (self.device.type == "cuda" or self._world() > 1):
            self._setup_flat_grads()
        if self._flat_grad is not None:
            self._flat_grad.zero_()
            self._begin_grad_sync()
            loss.backward()
            self._finish_grad_sync()
        else:
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
        self.optimizer.step()
        sums += torch.stack([policy_loss.detach(), value_loss.detach(), entropy.detach()])

    def _update_graph(self, n, fetch, fetch_key, actions, old_log_probs, advantages, returns):
        """Static buffers + a CUDA graph of one full-size minibatch step (forward, loss, backward,
        the bucketed NCCL gradient all-reduce when distributed, gradient clip, fused Adam). Warm-up
        steps needed for capture are undone afterwards. The graph is reused only for the same
        sample count, batch size and observation source (`fetch_key` names the tensors `fetch`
        reads); anything else re-captures."""
        bs = self.batch_size
        key = (n, bs, fetch_key)
        if self._ug is not None and self._ug["key"] == key:
            ug = self._ug
        else:
            dev = self.device
            ug = self._ug = {"key": key, "graph": None,
                             "idx": torch.zeros(bs, dtype=torch.int64, device=dev),
                             "sums": torch.zeros(3, device=dev),
                             "actions": torch.empty(n, dtype=torch.int64, device=dev),
                             "old_lp": torch.empty(n, dtype=torch.float32, device=dev),
                             "adv": torch.empty(n, dtype=torch.float32, device=dev),
                             "ret": torch.empty(n, dtype=torch.float32, device=dev)}
        ug["actions"].copy_(actions)
        ug["old_lp"].copy_(old_log_probs)
        ug["adv"].copy_(advantages)
        ug["ret"].copy_(returns)
        if ug["graph"] is None:
            ug["fetch"] = fetch
            params = list(self.policy.parameters())
            p_backup = [p.detach().clone() for p in params]
            o_backup = {p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
                        for p, st in self.optimizer.state.items()}

            def body():
                self._minibatch_step(ug["idx"], ug["fetch"], ug["actions"], ug["old_lp"], ug["adv"],
                                     ug["ret"], ug["sums"], validate=False)
            ug["idx"].copy_(torch.arange(bs, device=self.device) % n)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            ug["graph"] = graph
            with torch.no_grad():  # undo the warm-up/capture steps in place (the graph holds these addresses)
                for p, b in zip(params, p_backup):
                    p.copy_(b)
                for p, st in self.optimizer.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if p in o_backup and k in o_backup[p]:
                                v.copy_(o_backup[p][k])
                            else:
                                v.zero_()
            ug["sums"].zero_()
        return ug

    def _run_epochs(self, n: int, fetch: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                    actions, old_log_probs, advantages, returns, fetch_key=None) -> Dict[str, float]:
        """`epochs` passes over shuffled minibatches (ppo_agent.py:192-231). Under torch.distributed
        every rank runs the SAME number of optimisation steps (each contains collectives): the count
        follows the largest shard, and a rank whose shard is one env smaller wraps its permutation
        around so that its minibatches stay full."""
        self.policy.train()
        world = self._world()
        n_steps_from = n
        if world > 1:
            t = torch.tensor([n], dtype=torch.int64, device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            n_steps_from = int(t.item())
        ug = None
        # the list path hands fresh tensors to every update (fetch_key None): nothing stable to capture
        if self.graph_update and fetch_key is not None and self.shuffle == "torch" and n >= self.batch_size:
            ug = self._update_graph(n, fetch, fetch_key, actions, old_log_probs, advantages, returns)
            ug["sums"].zero_()
        sums = ug["sums"] if ug is not None else torch.zeros(3, device=self.device)
        num_updates = 0
        perm_state: dict = {}
        for _ in range(self.epochs):
            perm = self._permutation(n, perm_state)
            if world > 1 and n_steps_from > n:  # smaller shard: wrap around (see docstring)
                perm = torch.cat([perm, perm[:n_steps_from - n]])
            for start in range(0, n_steps_from, self.batch_size):
                idx = perm[start:start + self.batch_size]
                if world > 1 and idx.numel() < self.batch_size and n >= self.batch_size:
                    idx = torch.cat([idx, perm[:self.batch_size - idx.numel()]])  # same minibatch size on every rank
                if ug is not None and idx.numel() == self.batch_size:
                    ug["idx"].copy_(idx)
                    ug["graph"].replay()
                else:
                    self._minibatch_step(idx, fetch, actions, old_log_probs, advantages, returns, sums)
                num_updates += 1
        if hasattr(self.policy, "mark_updated"):
            self.policy.mark_updated()  # fused optimizers do not bump tensor versions
        p, v, e = (sums / max(num_updates, 1)).tolist()  # the only host sync of the update
        return {"policy_loss": p, "value_loss": v, "entropy": e}

    def _normalize_advantages(self, adv: torch.Tensor) -> torch.Tensor:
        """ppo_agent.py:173 (numpy mean / std with ddof 0). Under torch.distributed the two moments
        are taken over ALL ranks' samples, so the update does not depend on how envs are sharded."""
        if self._world() == 1:
            return (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
        m = torch.stack([adv.sum(dtype=torch.float64), (adv.double() ** 2).sum(),
                         torch.tensor(float(adv.numel()), dtype=torch.float64, device=adv.device)])
        torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.SUM)
        mean = m[0] / m[2]
        std = torch.sqrt(torch.clamp(m[1] / m[2] - mean * mean, min=0.0))
        return ((adv - mean.float()) / (std.float() + 1e-8))

    def update(self, rollout: Optional[DeviceRollout] = None, sim=None, last_value: Optional[torch.Tensor] = None
               ) -> Dict[str, float]:
        """ppo_agent.py:159-247. With no arguments it consumes the list buffers filled by
        `store_step` (reference calling convention); with a `DeviceRollout` it trains from the
        device-resident rollout, decoding packed observations through `sim` per minibatch."""
        if rollout is None:
            return self._update_from_lists()
        T, N = rollout.t, rollout.N
        if T == 0:
            return {}
        r, v, d = rollout.rewards[:T], rollout.values[:T], rollout.dones[:T]
        if self.gae_mode == "reference":  # flat-list GAE, bootstrap 0 (ppo_agent.py:127,170)
            adv, ret = compute_gae(r.reshape(-1, 1), v.reshape(-1, 1), d.reshape(-1, 1), None, self.gamma, self.lam)
        else:
            adv, ret = compute_gae(r, v, d, last_value, self.gamma, self.lam)
        adv, ret = adv.reshape(-1), ret.reshape(-1)
        adv = self._normalize_advantages(adv)
        if self.packed_encoder and rollout.store == "packed":
            obs_dtype = "packed"
        else:
            obs_dtype = "bf16" if self.precision == "bf16" else "f32"
        src = rollout.packed if rollout.store == "packed" else rollout.obs
        fetch_key = (obs_dtype, src.data_ptr(), T, N, None if sim is None else id(sim))
        stats = self._run_epochs(T * N, lambda idx: rollout.minibatch_obs(idx, sim, obs_dtype),
                                 rollout.actions[:T].reshape(-1), rollout.log_probs[:T].reshape(-1), adv, ret,
                                 fetch_key=fetch_key)
        rollout.reset()
        return stats

    def _update_from_lists(self) -> Dict[str, float]:
        if len(self.obs_grid_buffer) == 0:
            return {}
        adv_np, ret_np = self.compute_advantages()
        adv_np = (adv_np - adv_np.mean()) / (adv_np.std() + 1e-8)
        dev = self.device
        obs_grid = torch.as_tensor(np.stack([np.asarray(o) for o in self.obs_grid_buffer]), dtype=torch.float32).to(dev)
        obs_extra = torch.as_tensor(np.stack([np.asarray(o) for o in self.obs_extra_buffer]), dtype=torch.float32).to(dev)
        actions = torch.as_tensor(self.action_buffer, dtype=torch.int64).to(dev)
        old_lp = torch.as_tensor(self.log_prob_buffer, dtype=torch.float32).to(dev)
        adv = torch.as_tensor(adv_np, dtype=torch.float32).to(dev)
        ret = torch.as_tensor(ret_np, dtype=torch.float32).to(dev)
        stats = self._run_epochs(len(self.action_buffer), lambda idx: (obs_grid[idx], obs_extra[idx]),
                                 actions, old_lp, adv, ret)
        self.reset_buffers()
        return stats
