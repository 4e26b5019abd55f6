"""Round-2 exploration of the policy path on one B200 (not part of the product): what cuDNN does
with autotuning, with a 16-channel channels-last input, how a padded-layout shifted-GEMM weight
gradient compares with cuDNN's wgrad, and the write-only HBM ceiling (fill) next to the copy peak."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from inversus_b200 import InversusCNNPolicy, PPOAgent  # noqa: E402

dev = "cuda"
bf = torch.bfloat16


def bench(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def prof(fn, n=3, rows=16):
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=rows, max_name_column_width=90))


# ---------------------------------------------------------------- 1. write-only ceiling
big = torch.empty(7_700_000_000 // 4, dtype=torch.float32, device=dev)
ms = bench(lambda: big.zero_(), 10)
print(f"fill (cudaMemset via zero_) 7.7 GB: {ms:.3f} ms  {big.numel() * 4 / ms / 1e6:.1f} GB/s")
ms = bench(lambda: big.fill_(1.0), 10)
print(f"fill_ kernel 7.7 GB: {ms:.3f} ms  {big.numel() * 4 / ms / 1e6:.1f} GB/s")
half = big[: big.numel() // 2]
other = big[big.numel() // 2:]
ms = bench(lambda: other.copy_(half), 10)
print(f"copy 3.85 GB -> 3.85 GB: {ms:.3f} ms  {big.numel() * 4 / ms / 1e6:.1f} GB/s (read+write)")
del big, half, other
torch.cuda.empty_cache()

# ---------------------------------------------------------------- 2. conv1 input channels 12 vs 16
for B in (8192, 65536):
    x12 = (torch.rand(B, 12, 10, 15, device=dev) > 0.7).to(bf).contiguous(memory_format=torch.channels_last)
    x16 = F.pad(x12, (0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last)
    w12 = torch.randn(32, 12, 3, 3, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
    w16 = F.pad(w12, (0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last)
    for bm in (False, True):
        torch.backends.cudnn.benchmark = bm
        with torch.no_grad():
            t12 = bench(lambda: F.conv2d(x12, w12, None, padding=1))
            t16 = bench(lambda: F.conv2d(x16, w16, None, padding=1))
        print(f"conv1 fwd B={B} cudnn.benchmark={bm}: C=12 {t12 * 1e3:.1f} us   C=16 {t16 * 1e3:.1f} us")
    torch.backends.cudnn.benchmark = False
    if B == 8192:
        with torch.no_grad():
            prof(lambda: F.conv2d(x16, w16, None, padding=1), rows=4)
        w16g = w16.clone().requires_grad_(True)
        x16g = x16.clone()

        def c1_train():
            y = F.conv2d(x16g, w16g, None, padding=1)
            y.backward(y)
            w16g.grad = None
        prof(c1_train, rows=6)

# ---------------------------------------------------------------- 3. whole policy: benchmark flag, batch sizes
torch.manual_seed(0)
m = InversusCNNPolicy().to(dev)
for bm in (False, True):
    torch.backends.cudnn.benchmark = bm
    for B in (8192, 65536):
        g = (torch.rand(B, 12, 10, 15, device=dev) > 0.7).to(bf)
        e = torch.rand(B, 4, device=dev)
        with torch.no_grad():
            ms = bench(lambda: m.infer(g, e))
        print(f"infer B={B} benchmark={bm}: {ms:.3f} ms  {B / ms / 1e3:.2f} M samples/s  {B * 93e6 / ms / 1e9:.0f} TFLOP/s")
    for B in (8192, 32768):
        agent = PPOAgent(m, device=dev, precision="bf16", batch_size=B, epochs=1)
        g = (torch.rand(B, 12, 10, 15, device=dev) > 0.7).to(bf)
        e = torch.rand(B, 4, device=dev)
        act = torch.randint(0, 13, (B,), device=dev)
        adv, ret, olp = torch.randn(B, device=dev), torch.randn(B, device=dev), -torch.rand(B, device=dev)

        def train_step():
            agent._run_epochs(B, lambda idx: (g[idx], e[idx]), act, olp, adv, ret)
        ms = bench(train_step, 5)
        print(f"train minibatch B={B} benchmark={bm}: {ms:.3f} ms  {B / ms / 1e3:.3f} M samples/s  "
              f"{B * 3 * 93e6 / ms / 1e9:.0f} TFLOP/s")
        if B == 8192 and bm:
            prof(train_step, rows=26)
        del agent
torch.backends.cudnn.benchmark = False

# ---------------------------------------------------------------- 4. conv4 wgrad: cuDNN vs shifted GEMMs on a padded layout
B = 8192
C = 128
x = torch.randn(B, C, 10, 15, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
dy = torch.randn(B, C, 10, 15, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
w = torch.randn(C, C, 3, 3, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
for bm in (False, True):
    torch.backends.cudnn.benchmark = bm
    ms = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                                                           (False, True, False)))
    print(f"conv4 wgrad cuDNN benchmark={bm}: {ms * 1e3:.1f} us  {B * 150 * C * C * 9 * 2 / ms / 1e9:.0f} TFLOP/s")
    ms = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                                                           (True, False, False)))
    print(f"conv4 dgrad cuDNN benchmark={bm}: {ms * 1e3:.1f} us  {B * 150 * C * C * 9 * 2 / ms / 1e9:.0f} TFLOP/s")
    with torch.no_grad():
        ms = bench(lambda: F.conv2d(x, w, None, padding=1))
    print(f"conv4 fprop cuDNN benchmark={bm}: {ms * 1e3:.1f} us  {B * 150 * C * C * 9 * 2 / ms / 1e9:.0f} TFLOP/s")
torch.backends.cudnn.benchmark = False

# padded flat layout: [B, 11, 16, C] with column 15 and row 10 zero; a (dy, dx) tap is a row offset
# of dy*16 + dx in the flat [B*176, C] matrix
P = 176
xp = torch.zeros(B * P + 64, C, device=dev, dtype=bf)
xv = xp[32:32 + B * P].view(B, 11, 16, C)
xv[:, :10, :15] = x.permute(0, 2, 3, 1)
dyp = torch.zeros(B * P, C, device=dev, dtype=bf)
dyp.view(B, 11, 16, C)[:, :10, :15] = dy.permute(0, 2, 3, 1)
M = B * P


def wgrad_shift():
    outs = []
    for oy in (-1, 0, 1):
        a = torch.as_strided(xp, (3, M, C), (C, C, 1), storage_offset=(32 + oy * 16 - 1) * C)
        outs.append(torch.matmul(a.transpose(1, 2), dyp))   # [3, Cin, Cout]
    return torch.stack(outs)                                # [3(dy), 3(dx), Cin, Cout]


ms = bench(wgrad_shift)
print(f"conv4 wgrad as 3 strided-batched GEMMs on the padded layout: {ms * 1e3:.1f} us  "
      f"{B * 150 * C * C * 9 * 2 / ms / 1e9:.0f} useful TFLOP/s")
ref = torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (False, True, False))[1]
got = wgrad_shift().permute(3, 2, 0, 1)  # [Cout, Cin, ky, kx]
print("wgrad shift-vs-cudnn max rel err:", ((got.float() - ref.float()).abs().max() / ref.float().abs().max()).item())
prof(wgrad_shift, rows=5)


wt = w.permute(2, 3, 1, 0).contiguous()  # [ky, kx, Cin, Cout]


def fprop_shift_bf16():
    out = None
    for ky in range(3):
        a = torch.as_strided(xp, (3, M, C), (C, C, 1), storage_offset=(32 + (ky - 1) * 16 - 1) * C)
        y = torch.bmm(a, wt[ky])  # [3, M, Cout] -- separate outputs, summed afterwards (cost shown for reference)
        out = y if out is None else out + y
    return out.sum(0)


ms = bench(fprop_shift_bf16)
print(f"conv4 fprop as 3 strided-batched GEMMs (+ sums): {ms * 1e3:.1f} us")
print("done", time.strftime("%H:%M:%S"))
