"""Policy/value network of the reference, checkpoint-compatible, with a bf16 tensor-core path.

Architecture and parameter names follow inversus_rl/policies.py:11-108 so that the reference's
`.pt` state dicts load unchanged: conv1..conv4 (3x3, 12->32->64->128->128), norm1..norm4
(LayerNorm over [C,H,W]), one residual around conv4, and two MLP heads `fc_actor` / `fc_critic`
(19204 -> 256 -> 128 -> 13 | 1) on the flattened features concatenated with the 4 extra
features. This is the only dense contraction in the system and stays in PyTorch (north star):
`forward` is the plain module-dtype path (fp32 = the parity path); `infer` runs the same weights
under bf16 autocast in channels-last layout with the two 19204-wide head GEMMs fused into one.
"""
from __future__ import annotations

from typing import Tuple

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

    # ------------------------------------------------------------------ bf16 tensor-core inference
    def infer(self, grid_tensor: torch.Tensor, extra_vector: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Inference under bf16 autocast, channels-last, actor/critic first layers as one GEMM.
        Accepts f32/bf16/u8 observation planes (all hold only 0/1). Returns fp32 logits and value."""
        with torch.autocast(device_type=grid_tensor.device.type, dtype=torch.bfloat16):
            grid = grid_tensor.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            feat = self._trunk(grid).flatten(1)
            x = torch.cat([feat.to(torch.bfloat16), extra_vector.to(torch.bfloat16)], dim=1)
            a0, c0 = self.fc_actor[0], self.fc_critic[0]
            h = F.relu(F.linear(x, torch.cat([a0.weight, c0.weight], 0), torch.cat([a0.bias, c0.bias], 0)))
            ha, hc = h[:, : a0.out_features], h[:, a0.out_features:]
            logits = self.fc_actor[4](F.relu(self.fc_actor[2](ha)))
            value = self.fc_critic[4](F.relu(self.fc_critic[2](hc)))
        return logits.float(), value.float()


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
