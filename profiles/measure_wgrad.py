"""conv3 / conv4 weight gradient: the tcgen05 kernel (csrc/wgrad_kernels.cu) against the library's
wgrad on the PPO minibatch shapes, CUDA-event timed, and one PPO minibatch step with and without it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200 import _capi, InversusCNNPolicy, PPOAgent  # noqa: E402


def bench(fn, n=20, warm=5):
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


lib = _capi.load()
st = int(torch.cuda.current_stream().cuda_stream)
for cin, cout in ((128, 128), (64, 128), (32, 64)):
    for B in (8192, 32768):
        x = torch.randn(B, cin, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dy = torch.randn(B, cout, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        out = torch.empty((cout, 3, 3, cin), dtype=torch.float32, device="cuda")
        scratch = torch.empty(lib.inv_conv3x3_wgrad_scratch_floats(cin, cout), dtype=torch.float32, device="cuda")
        ms_lib = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                                                                   (False, True, False)))
        ms_own = bench(lambda: lib.inv_conv3x3_wgrad(dy.data_ptr(), x.data_ptr(), B, cin, cout, out.data_ptr(),
                                                     scratch.data_ptr(), st))
        flop = B * 150 * cout * cin * 9 * 2
        print(f"wgrad cin={cin} cout={cout} B={B}: library {ms_lib * 1e3:.1f} us ({flop / ms_lib / 1e9:.0f} TFLOP/s)   "
              f"tcgen05 kernel {ms_own * 1e3:.1f} us ({flop / ms_own / 1e9:.0f} TFLOP/s, "
              f"{(x.numel() + dy.numel()) * 2 / ms_own / 1e6:.0f} GB/s of unique operand bytes)", flush=True)

torch.manual_seed(0)
m = InversusCNNPolicy().cuda()
for B in (8192, 32768):
    g = (torch.rand(B, 12, 10, 15, device="cuda") > 0.7).to(torch.bfloat16)
    e = torch.rand(B, 4, device="cuda")
    act = torch.randint(0, 13, (B,), device="cuda")
    adv, ret, olp = torch.randn(B, device="cuda"), torch.randn(B, device="cuda"), -torch.rand(B, device="cuda")
    for custom in (False, True):
        m.use_custom_wgrad = custom
        agent = PPOAgent(m, device="cuda", precision="bf16", batch_size=B, epochs=1)
        ms = bench(lambda: agent._run_epochs(B, lambda idx: (g[idx], e[idx]), act, olp, adv, ret), 5)
        print(f"train minibatch B={B} custom_wgrad={custom}: {ms:.3f} ms  {B / ms / 1e3:.3f} M samples/s", flush=True)
