"""Policy/value network of the reference, checkpoint-compatible, with a bf16 tensor-core path.

Architecture and parameter names follow inversus_rl/policies.py:11-108 so that the reference's
`.pt` state dicts load unchanged: conv1..conv4 (3x3, 12->32->64->128->128), norm1..norm4
(LayerNorm over [C,H,W]), one residual around conv4, and two MLP heads `fc_actor` / `fc_critic`
(19204 -> 256 -> 128 -> 13 | 1) on the flattened features concatenated with the 4 extra
features. This is the only dense contraction in the system and stays in PyTorch (north star):
`forward` is the plain module-dtype path (fp32 = the parity path); `forward_bf16` / `infer` run the
same weights in bf16, channels-last end to end, with the two 19204-wide head GEMMs fused into one.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .constants import BOARD_H, BOARD_W, EXTRA_ELEMS, NUM_ACTIONS, OBS_CHANNELS

_CONV_WIDTHS = (32, 64, 128, 128)


def _mlp_head(in_dim: int, hidden: int, out_dim: int) -> nn.Sequential:
    # indices 0, 2, 4 hold the Linear layers -- the reference's state_dict keys (fc_actor.0.weight ...)
    return nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden // 2), nn.ReLU(),
                         nn.Linear(hidden // 2, out_dim))


class InversusCNNPolicy(nn.Module):
    def __init__(self, channels: int = OBS_CHANNELS, height: int = BOARD_H, width: int = BOARD_W,
                 extra_dim: int = EXTRA_ELEMS, hidden_dim: int = 256):
        super().__init__()
        self.channels, self.height, self.width, self.extra_dim = channels, height, width, extra_dim
        c_in = channels
        for i, c_out in enumerate(_CONV_WIDTHS, start=1):
            setattr(self, f"conv{i}", nn.Conv2d(c_in, c_out, kernel_size=3, padding=1))
            setattr(self, f"norm{i}", nn.LayerNorm([c_out, height, width]))
            c_in = c_out
        self.relu = nn.ReLU()
        self.flatten = nn.Flatten()
        self.feature_dim = _CONV_WIDTHS[-1] * height * width
        self.use_fused_kernels = True  # CUDA only: fused LayerNorm(+residual)+ReLU kernels in forward_bf16
        self.use_custom_wgrad = True   # CUDA only: tcgen05 weight gradient for conv2 / conv3 / conv4 when training
        self._weights_epoch = 0        # bumped by mark_updated(); part of the inference-cache key
        self.fc_actor = _mlp_head(self.feature_dim + extra_dim, hidden_dim, NUM_ACTIONS)
        self.fc_critic = _mlp_head(self.feature_dim + extra_dim, hidden_dim, 1)

    # ------------------------------------------------------------------ shared trunk
    def _trunk(self, grid: torch.Tensor) -> torch.Tensor:
        x = grid
        for i in (1, 2, 3):
            x = F.relu(getattr(self, f"norm{i}")(getattr(self, f"conv{i}")(x)))
        x = F.relu(self.norm4(self.conv4(x) + x))  # residual around conv4 (policies.py:98-100)
        return x

    def forward(self, grid_tensor: torch.Tensor, extra_vector: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B,12,H,W), (B,4) -> logits (B,13), value (B,1). Same math as policies.py:77-108."""
        p = self.conv1.weight
        grid = grid_tensor.to(p.dtype) if grid_tensor.dtype != p.dtype else grid_tensor
        x = torch.cat([self._trunk(grid).flatten(1), extra_vector.to(p.dtype)], dim=1)
        return self.fc_actor(x), self.fc_critic(x)

    # ------------------------------------------------------------------ bf16 tensor-core path
    def _prepare_group(self, group: str) -> dict:
        """bf16 working copies of one group of parameters in the layouts the fast path wants
        (differentiable w.r.t. the fp32 masters). Groups: "l1" (block 1 evaluated from packed states:
        fp32 masters + HWC LayerNorm affine), "c1".."c4" (channels-last conv filter, conv bias,
        HWC-ordered LayerNorm affine), "head" (the fused head matrix with its K dimension split into
        HWC-ordered trunk columns + padded extras, and the small head layers)."""
        bf = torch.bfloat16
        H, W, nf = self.height, self.width, self.feature_dim
        prep = {}
        if group == "l1":  # csrc/encoder_kernels.cu reads the fp32 masters directly
            prep["w1"], prep["b1"] = self.conv1.weight, self.conv1.bias
            prep["g1_hwc"] = self.norm1.weight.permute(1, 2, 0).reshape(-1)
            prep["be1_hwc"] = self.norm1.bias.permute(1, 2, 0).reshape(-1)
            return prep
        if group in ("c1", "c2", "c3", "c4"):
            i = int(group[1])
            conv, norm = getattr(self, f"conv{i}"), getattr(self, f"norm{i}")
            prep[f"cw{i}"] = conv.weight.to(bf).contiguous(memory_format=torch.channels_last)
            prep[f"cb{i}"] = conv.bias.to(bf)
            prep[f"nw{i}"] = norm.weight.permute(1, 2, 0).to(bf).contiguous()
            prep[f"nb{i}"] = norm.bias.permute(1, 2, 0).to(bf).contiguous()
            return prep
        assert group == "head"
        a0, c0 = self.fc_actor[0], self.fc_critic[0]
        w0 = torch.cat([a0.weight, c0.weight], 0)                # [2*hidden, 19200 + extra]
        c = nf // (H * W)
        if w0.is_cuda and self.use_fused_kernels:  # one tiled transpose+cast kernel each way (csrc/policy_kernels.cu)
            from .fused_ops import head_weight_to_hwc
            prep["w_feat"] = head_weight_to_hwc(w0, c, H * W)
        else:
            prep["w_feat"] = w0[:, :nf].to(bf).reshape(-1, c, H * W).permute(0, 2, 1).reshape(-1, nf)  # CHW -> HWC columns
        pad = (8 - self.extra_dim % 8) % 8
        prep["w_extra"] = F.pad(w0[:, nf:], (0, pad)).to(bf)
        prep["b0"] = torch.cat([a0.bias, c0.bias], 0).to(bf)
        for name, seq in (("actor", self.fc_actor), ("critic", self.fc_critic)):
            for k in (2, 4):
                prep[f"{name}_w{k}"] = seq[k].weight.to(bf)
                prep[f"{name}_b{k}"] = seq[k].bias.to(bf)
        return prep

    _PREP_GROUPS = ("l1", "c1", "c2", "c3", "c4", "head")

    def _prepare_bf16(self) -> dict:
        """All groups at once (what the inference cache holds)."""
        prep = {}
        for g in self._PREP_GROUPS:
            prep.update(self._prepare_group(g))
        return prep

    def _forward_prepared(self, prep: dict, grid_tensor, extra_vector: Optional[torch.Tensor]):
        bf = torch.bfloat16
        H, W = self.height, self.width
        packed = not isinstance(grid_tensor, torch.Tensor)       # fused_ops.PackedStates
        if packed:
            # block 1 (conv1 + bias + LayerNorm + ReLU) straight from the 80-byte env states: no
            # observation tensor, no layout conversion, no K=108 convolution
            from .fused_ops import encode_layer1
            flat, extra_vector = encode_layer1(grid_tensor, prep["w1"], prep["b1"], prep["g1_hwc"], prep["be1_hwc"],
                                               self.norm1.eps)
            nb = flat.shape[0]
            x = flat.view(nb, H, W, _CONV_WIDTHS[0]).permute(0, 3, 1, 2)
            fused = True
        else:
            x = grid_tensor.to(bf).contiguous(memory_format=torch.channels_last)
            nb = x.shape[0]
            fused = x.is_cuda and self.use_fused_kernels
        res = None
        for i in ((2, 3, 4) if packed else (1, 2, 3, 4)):
            # fused path: cuDNN runs bias-free, the per-channel bias is added inside the LayerNorm kernel
            w_i = prep[f"cw{i}"]
            if fused and self.use_custom_wgrad and w_i.requires_grad and i >= 2:
                # training: library forward / input gradient, tcgen05 weight gradient (csrc/wgrad_kernels.cu)
                from .fused_ops import conv3x3, conv3x3_supported
                y = conv3x3(x, w_i) if conv3x3_supported(x, w_i) else F.conv2d(x, w_i, None, padding=1)
            else:
                y = F.conv2d(x, w_i, None if fused else prep[f"cb{i}"], padding=1)
            eps = getattr(self, f"norm{i}").eps
            yp = y.permute(0, 2, 3, 1)                           # [B,H,W,C] view of channels-last memory
            c = yp.shape[-1]
            if fused:  # one hand-written kernel: +bias (+residual) -> LayerNorm -> affine -> ReLU (csrc/policy_kernels.cu)
                from .fused_ops import layer_norm_relu
                flat = layer_norm_relu(yp.reshape(nb, -1), prep[f"nw{i}"].reshape(-1), prep[f"nb{i}"].reshape(-1), eps,
                                       residual=res.permute(0, 2, 3, 1).reshape(nb, -1) if i == 4 else None,
                                       channel_bias=prep[f"cb{i}"], channels=c)
                x = flat.view(nb, H, W, c).permute(0, 3, 1, 2)
            else:
                if i == 4:
                    yp = yp + res.permute(0, 2, 3, 1)            # residual around conv4 (policies.py:98-100)
                yp = F.layer_norm(yp, (H, W, c), prep[f"nw{i}"], prep[f"nb{i}"], eps)
                x = F.relu(yp).permute(0, 3, 1, 2)               # back to an NCHW view, still channels-last
            if i == 3:
                res = x
        feat = x.permute(0, 2, 3, 1).reshape(nb, -1)             # [B, H*W*C] without a copy
        ex = F.pad(extra_vector.to(bf), (0, prep["w_extra"].shape[1] - self.extra_dim))
        h = F.relu(F.linear(feat, prep["w_feat"], prep["b0"]) + F.linear(ex, prep["w_extra"]))
        n_a = self.fc_actor[0].out_features
        ha, hc = h[:, :n_a], h[:, n_a:]
        logits = F.linear(F.relu(F.linear(ha, prep["actor_w2"], prep["actor_b2"])), prep["actor_w4"], prep["actor_b4"])
        value = F.linear(F.relu(F.linear(hc, prep["critic_w2"], prep["critic_b2"])), prep["critic_w4"], prep["critic_b4"])
        return logits.float(), value.float()

    def forward_bf16(self, grid_tensor: torch.Tensor, extra_vector: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same function in bf16 with fp32 master weights (differentiable; used for the PPO update
        when precision="bf16"). Everything stays channels-last so cuDNN never converts layouts:
        LayerNorm([C,H,W]) runs on the contiguous [B,H,W,C] view with an HWC-permuted copy of its
        affine parameters and bf16 I/O (fp32 statistics inside the kernel); the two 19204-wide head
        GEMMs are fused into one whose K dimension is split into the 19200 trunk features (16-byte
        aligned rows, weight columns permuted to HWC order so activations are never transposed)
        and the 4 extra features. Accepts f32/bf16/u8 observation planes (they hold only 0/1).
        Returns fp32 logits and value."""
        # the working copies are made at their point of use: autograd then runs each conversion's
        # backward right after the layer's own, so the head gradients (96 % of all bytes) are final
        # -- and their all-reduce can start -- while backward is still in the trunk
        return self._forward_prepared(_LazyPrep(self), grid_tensor, extra_vector)

    @torch.no_grad()
    def infer(self, grid_tensor: torch.Tensor, extra_vector: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rollout inference (no grad) through the bf16 path. The prepared bf16 weights are cached
        and rebuilt only when a parameter changed (optimizer steps bump tensor versions)."""
        return self._forward_prepared(self.inference_weights(), grid_tensor, extra_vector)

    def mark_updated(self) -> None:
        """Tell the inference cache that the parameters changed (call after optimizer steps)."""
        self._weights_epoch += 1

    @torch.no_grad()
    def inference_weights(self) -> dict:
        """The cached bf16 working copies used by `infer`. When a parameter changed they are
        refreshed IN PLACE (same tensors, same addresses), so a CUDA graph that captured `infer`
        keeps seeing current weights after an optimizer step."""
        # tensor versions catch load_state_dict and ordinary in-place updates; fused optimizers
        # (torch._fused_adam_) do NOT bump them, so trainers also call mark_updated()
        key = tuple(p._version for p in self.parameters()) + (self._weights_epoch, str(self.conv1.weight.device))
        cache = getattr(self, "_infer_cache", None)
        if cache is None or cache[0][-1] != key[-1]:
            cache = [key, self._prepare_bf16()]
            object.__setattr__(self, "_infer_cache", cache)
        elif cache[0] != key:
            fresh = self._prepare_bf16()
            for k, v in cache[1].items():
                if fresh[k] is not v and fresh[k].data_ptr() != v.data_ptr():  # fp32 masters are cached as themselves
                    v.copy_(fresh[k])
            cache[0] = key
        return cache[1]


class _LazyPrep(dict):
    """`prep[key]` prepares the parameter group that `key` belongs to on first use."""

    _GROUP = {"w1": "l1", "b1": "l1", "g1_hwc": "l1", "be1_hwc": "l1"}
    _GROUP.update({f"{k}{i}": f"c{i}" for i in (1, 2, 3, 4) for k in ("cw", "cb", "nw", "nb")})

    def __init__(self, policy: InversusCNNPolicy):
        super().__init__()
        self._policy = policy

    def __missing__(self, key):
        self.update(self._policy._prepare_group(self._GROUP.get(key, "head")))
        return dict.__getitem__(self, key)


def make_policy_from_env(env=None) -> InversusCNNPolicy:
    """policies.py:111-128: shapes from one reset observation (or the fixed 15x10 board)."""
    if env is None:
        return InversusCNNPolicy()
    grid, extra = env.reset()
    c, h, w = grid.shape
    return InversusCNNPolicy(c, h, w, extra.shape[0])


# parameter names and shapes of the reference checkpoint (10 249 582 parameters)
REFERENCE_STATE_DICT_SHAPES = {
    "conv1.weight": (32, 12, 3, 3), "conv1.bias": (32,), "norm1.weight": (32, 10, 15), "norm1.bias": (32, 10, 15),
    "conv2.weight": (64, 32, 3, 3), "conv2.bias": (64,), "norm2.weight": (64, 10, 15), "norm2.bias": (64, 10, 15),
    "conv3.weight": (128, 64, 3, 3), "conv3.bias": (128,), "norm3.weight": (128, 10, 15), "norm3.bias": (128, 10, 15),
    "conv4.weight": (128, 128, 3, 3), "conv4.bias": (128,), "norm4.weight": (128, 10, 15), "norm4.bias": (128, 10, 15),
    "fc_actor.0.weight": (256, 19204), "fc_actor.0.bias": (256,), "fc_actor.2.weight": (128, 256),
    "fc_actor.2.bias": (128,), "fc_actor.4.weight": (13, 128), "fc_actor.4.bias": (13,),
    "fc_critic.0.weight": (256, 19204), "fc_critic.0.bias": (256,), "fc_critic.2.weight": (128, 256),
    "fc_critic.2.bias": (128,), "fc_critic.4.weight": (1, 128), "fc_critic.4.bias": (1,),
}
