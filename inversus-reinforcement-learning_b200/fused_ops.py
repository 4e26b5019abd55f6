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
    def forward(ctx, x, res, cbias, gamma, beta, eps, channels):
        assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
        B, D = x.shape
        assert gamma.dtype == torch.bfloat16 and gamma.numel() == D and beta.numel() == D
        gamma, beta = gamma.contiguous(), beta.contiguous()
        if res is not None:
            assert res.shape == x.shape and res.dtype == torch.bfloat16 and res.is_contiguous()
        if cbias is not None:
            assert cbias.dtype == torch.bfloat16 and cbias.numel() == channels
            cbias = cbias.contiguous()
        y = torch.empty_like(x)
        mean = torch.empty(B, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().inv_ln_relu_fwd(
                x.data_ptr(), None if res is None else res.data_ptr(), None if cbias is None else cbias.data_ptr(),
                gamma.data_ptr(), beta.data_ptr(), B, D, channels, float(eps), y.data_ptr(), mean.data_ptr(),
                rstd.data_ptr(), _stream(x)))
        empty = x.new_empty(0)
        ctx.save_for_backward(x, res if res is not None else empty, cbias if cbias is not None else empty,
                              gamma, beta, mean, rstd)
        ctx.has_res, ctx.has_cb, ctx.channels = res is not None, cbias is not None, channels
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, cbias, gamma, beta, mean, rstd = ctx.saved_tensors
        B, D = x.shape
        C_ = ctx.channels
        dy = dy.contiguous()
        lib = _capi.load()
        dx = torch.empty_like(x)
        dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
        dcb = torch.empty(C_, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            partials = torch.empty((lib.inv_ln_relu_partials(D), 2 * D + C_), dtype=torch.float32, device=x.device)
            _capi.check(lib.inv_ln_relu_bwd(
                dy.data_ptr(), x.data_ptr(), res.data_ptr() if ctx.has_res else None,
                cbias.data_ptr() if ctx.has_cb else None, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                rstd.data_ptr(), B, D, C_, dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), dcb.data_ptr(),
                partials.data_ptr(), _stream(x)))
        return (dx, (dx if ctx.has_res else None), (dcb.to(cbias.dtype) if ctx.has_cb else None),
                dgamma.to(gamma.dtype), dbeta.to(beta.dtype), None, None)


def layer_norm_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                    residual: Optional[torch.Tensor] = None, channel_bias: Optional[torch.Tensor] = None,
                    channels: Optional[int] = None) -> torch.Tensor:
    """relu(LayerNorm_over_dim1(x + channel_bias [+ residual]) * gamma + beta) for bf16 CUDA tensors
    [B, D] whose rows are HWC-ordered feature maps with `channels` channels (channel = index % channels)."""
    if channels is None:
        channels = channel_bias.numel() if channel_bias is not None else 8
    return _LNReLU.apply(x, residual, channel_bias, gamma, beta, eps, channels)


class _HeadWeightToHWC(torch.autograd.Function):
    """fp32 [R, >= C*P] (columns in CHW order) -> bf16 [R, P*C] (HWC order); gradient back."""

    @staticmethod
    def forward(ctx, w, C_, P_):
        assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.stride(1) == 1
        R = w.shape[0]
        out = torch.empty((R, C_ * P_), dtype=torch.bfloat16, device=w.device)
        with torch.cuda.device(w.device):
            _capi.check(_capi.load().inv_transpose_cast(w.data_ptr(), 1, w.stride(0), out.data_ptr(), 0, C_ * P_,
                                                        R, C_, P_, _stream(w)))
        ctx.shape, ctx.cp = w.shape, (C_, P_)
        return out

    @staticmethod
    def backward(ctx, g):
        C_, P_ = ctx.cp
        g = g.contiguous()
        gw = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)  # extra columns get no gradient here
        with torch.cuda.device(g.device):
            _capi.check(_capi.load().inv_transpose_cast(g.data_ptr(), 0, C_ * P_, gw.data_ptr(), 1, gw.stride(0),
                                                        g.shape[0], P_, C_, _stream(g)))
        return gw, None, None


def head_weight_to_hwc(w: torch.Tensor, channels: int, positions: int) -> torch.Tensor:
    """bf16 HWC-ordered copy of the first channels*positions columns of the fp32 head weight."""
    return _HeadWeightToHWC.apply(w, channels, positions)


class PackedStates:
    """Packed env states as the policy's input: `planes` is [5, M, 4] int32 (a view or copy of
    `BatchedInversus.packed_state`, 80 bytes per env), `view` 0 = P1's perspective, 1 = P2's. The
    policy's first block reads this directly (`encode_layer1`); no observation tensor exists."""

    __slots__ = ("planes", "view")

    def __init__(self, planes: torch.Tensor, view: int = 0):
        assert planes.is_cuda and planes.dtype == torch.int32 and planes.dim() == 3 and planes.shape[0] == 5 \
            and planes.shape[2] == 4 and planes.is_contiguous(), "packed planes must be a contiguous [5, M, 4] int32 CUDA tensor"
        self.planes, self.view = planes, int(view)

    @property
    def shape(self):
        return (self.planes.shape[1],)

    def __len__(self):
        return self.planes.shape[1]

    def chunk(self, lo: int, hi: int) -> "PackedStates":
        return _PackedSlice(self.planes, self.view, lo, hi)


class _PackedSlice(PackedStates):
    """Entries [lo, hi) of a PackedStates without copying (the kernels take a plane stride)."""

    __slots__ = ("lo", "hi")

    def __init__(self, planes, view, lo, hi):
        self.planes, self.view, self.lo, self.hi = planes, view, int(lo), int(hi)

    @property
    def shape(self):
        return (self.hi - self.lo,)

    def __len__(self):
        return self.hi - self.lo


def _packed_args(ps: PackedStates):
    stride = ps.planes.shape[1]
    lo = getattr(ps, "lo", 0)
    count = len(ps)
    # entry lo of plane k sits at planes[k, lo]: offset the base pointer, keep the plane stride
    return ps.planes.data_ptr() + lo * 16, stride, count


class _EncodeLayer1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w1, b1, gamma_hwc, beta_hwc, ps, eps):
        assert w1.is_cuda and w1.dtype == torch.float32 and tuple(w1.shape) == (32, 12, 3, 3)
        w1, b1 = w1.contiguous(), b1.contiguous()
        gamma_hwc, beta_hwc = gamma_hwc.contiguous(), beta_hwc.contiguous()
        assert gamma_hwc.dtype == torch.float32 and gamma_hwc.numel() == 4800 and beta_hwc.numel() == 4800
        ptr, stride, count = _packed_args(ps)
        dev = w1.device
        y = torch.empty((count, 4800), dtype=torch.bfloat16, device=dev)
        extra = torch.empty((count, 4), dtype=torch.float32, device=dev)
        mean = torch.empty(count, dtype=torch.float32, device=dev)
        rstd = torch.empty(count, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.load().inv_encode_fwd(ptr, stride, count, ps.view, w1.data_ptr(), b1.data_ptr(),
                                                    gamma_hwc.data_ptr(), beta_hwc.data_ptr(), float(eps), y.data_ptr(),
                                                    extra.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _stream(w1)))
        ctx.save_for_backward(w1, b1, gamma_hwc, beta_hwc, mean, rstd)
        ctx.ps = ps
        ctx.mark_non_differentiable(extra)
        return y, extra

    @staticmethod
    def backward(ctx, dy, _dextra):
        w1, b1, gamma_hwc, beta_hwc, mean, rstd = ctx.saved_tensors
        ps = ctx.ps
        ptr, stride, count = _packed_args(ps)
        dev = w1.device
        dy = dy.contiguous()
        lib = _capi.load()
        dw1, db1 = torch.empty_like(w1), torch.empty_like(b1)
        dg, db = torch.empty_like(gamma_hwc), torch.empty_like(beta_hwc)
        with torch.cuda.device(dev):
            partials = torch.empty(lib.inv_encode_partials_floats(), dtype=torch.float32, device=dev)
            _capi.check(lib.inv_encode_bwd(ptr, stride, count, ps.view, w1.data_ptr(), b1.data_ptr(), gamma_hwc.data_ptr(),
                                           beta_hwc.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dy.data_ptr(),
                                           dw1.data_ptr(), db1.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                           partials.data_ptr(), _stream(w1)))
        return dw1, db1, dg, db, None, None


def encode_layer1(ps: PackedStates, w1: torch.Tensor, b1: torch.Tensor, gamma_hwc: torch.Tensor,
                  beta_hwc: torch.Tensor, eps: float = 1e-5):
    """relu(LayerNorm(conv1(observation(state)) + b1) * gamma + beta) straight from packed states:
    returns (y [M, 4800] bf16 in HWC order, extra [M, 4] f32). fp32 master weights in, fp32 grads out."""
    return _EncodeLayer1.apply(w1, b1, gamma_hwc, beta_hwc, ps, eps)


class _Conv3x3(torch.autograd.Function):
    """3x3 / padding-1 convolution of channels-last bf16 activations on the 15 x 10 board. Forward and
    input gradient are the library's (cuDNN runs both at the tensor-core peak); the WEIGHT gradient,
    where the library reaches 17 % of that peak, is the hand-written tcgen05 kernel of
    csrc/wgrad_kernels.cu."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.nn.functional.conv2d(x, w, None, padding=1)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        dy = dy.contiguous(memory_format=torch.channels_last)
        if ctx.needs_input_grad[0]:
            dx = torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                                                     (True, False, False))[0]
        if ctx.needs_input_grad[1]:
            B, cin = x.shape[0], x.shape[1]
            cout = w.shape[0]
            lib = _capi.load()
            out = torch.empty((cout, 3, 3, cin), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                scratch = torch.empty(lib.inv_conv3x3_wgrad_scratch_floats(cin, cout), dtype=torch.float32, device=x.device)
                _capi.check(lib.inv_conv3x3_wgrad(dy.data_ptr(), x.data_ptr(), B, cin, cout, out.data_ptr(),
                                                  scratch.data_ptr(), _stream(x)))
            dw = out.permute(0, 3, 1, 2).to(w.dtype)  # logical [co, ci, ky, kx], channels-last strides like w
        return dx, dw


def conv3x3_supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 4
            and tuple(x.shape[2:]) == (10, 15) and (w.shape[0], w.shape[1]) in ((128, 128), (128, 64), (64, 32))
            and tuple(w.shape[2:]) == (3, 3) and x.is_contiguous(memory_format=torch.channels_last))


def conv3x3(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """F.conv2d(x, w, None, padding=1) with the hand-written weight gradient (conv3 / conv4 shapes)."""
    return _Conv3x3.apply(x, w)
