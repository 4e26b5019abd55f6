"""GPU check of the tcgen05 weight-gradient kernel (csrc/wgrad_kernels.cu) against the library's
convolution backward on the same bf16 inputs (fp32 accumulation on both sides)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cin,cout", [(128, 128), (64, 128), (32, 64)])
@pytest.mark.parametrize("B", [1, 7, 300])
def test_conv3x3_weight_gradient_matches_the_library(cin, cout, B):
    import torch
    from inversus_b200.fused_ops import conv3x3, conv3x3_supported
    torch.manual_seed(B * 1000 + cin)
    x = torch.randn(B, cin, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, cout, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert conv3x3_supported(x, w)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    torch.nn.functional.conv2d(xr, wr, None, padding=1).backward(dy)
    xg, wg = x.clone().requires_grad_(), w.clone().requires_grad_()
    conv3x3(xg, wg).backward(dy)
    assert torch.equal(xg.grad, xr.grad)  # the input gradient is the library's in both
    # fp32 reference of the weight gradient from the same bf16 values
    want = torch.ops.aten.convolution_backward(dy.float(), x.float(), w.float(), None, (1, 1), (1, 1), (1, 1), False,
                                               (0, 0), 1, (False, True, False))[1]
    got = wg.grad.float()
    scale = want.abs().max().item()
    # both are fp32 sums of exact bf16 products; the result is then rounded to bf16 (2^-9 relative)
    assert (got - want).abs().max() <= 1.2e-2 * scale, ((got - want).abs().max().item(), scale)
    assert (wr.grad.float() - want).abs().max() <= 1.2e-2 * scale
    # every tap separately (a shifted or mirrored tap would pass a norm test but not this)
    for ky in range(3):
        for kx in range(3):
            a, b = got[:, :, ky, kx].flatten().double(), want[:, :, ky, kx].flatten().double()
            assert torch.dot(a, b) / (a.norm() * b.norm()) > 0.9999, (ky, kx)
