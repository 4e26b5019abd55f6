"""torch.autograd binding of the fused LayerNorm(+residual)+ReLU kernels (csrc/policy_kernels.cu).

Used by `InversusCNNPolicy.forward_bf16` on CUDA tensors: the convolutions stay in cuDNN, this
replaces the LayerNorm / residual add / ReLU / cast kernels between them with one kernel forward
and one (plus a tiny reduction) backward. Operates on the flat `[B, H*W*C]` view of channels-last
bf16 activations; gamma/beta are bf16 `[H*W*C]` in the same HWC order.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi


def _stream(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


class _LNReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
        B, D = x.shape
        assert gamma.dtype == torch.bfloat16 and gamma.numel() == D and beta.numel() == D
        gamma, beta = gamma.contiguous(), beta.contiguous()
        if res is not None:
            assert res.shape == x.shape and res.dtype == torch.bfloat16 and res.is_contiguous()
        y = torch.empty_like(x)
        mean = torch.empty(B, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().inv_ln_relu_fwd(
                x.data_ptr(), None if res is None else res.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, D,
                float(eps), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _stream(x)))
        ctx.save_for_backward(x, res if res is not None else x.new_empty(0), gamma, beta, mean, rstd)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, gamma, beta, mean, rstd = ctx.saved_tensors
        B, D = x.shape
        dy = dy.contiguous()
        lib = _capi.load()
        dx = torch.empty_like(x)
        dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            partials = torch.empty((lib.inv_ln_relu_partials(D), 2 * D), dtype=torch.float32, device=x.device)
            _capi.check(lib.inv_ln_relu_bwd(
                dy.data_ptr(), x.data_ptr(), res.data_ptr() if ctx.has_res else None, gamma.data_ptr(),
                beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), B, D, dx.data_ptr(), dgamma.data_ptr(),
                dbeta.data_ptr(), partials.data_ptr(), _stream(x)))
        return dx, (dx if ctx.has_res else None), dgamma.to(gamma.dtype), dbeta.to(beta.dtype), None


def layer_norm_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                    residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu(LayerNorm_over_dim1(x [+ residual]) * gamma + beta) for bf16 CUDA tensors [B, D]."""
    return _LNReLU.apply(x, residual, gamma, beta, eps)
