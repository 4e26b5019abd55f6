"""torch.profiler breakdown of the policy (the dense-contraction part that stays in PyTorch):
bf16 inference at rollout batch sizes and one PPO minibatch step. Run on a B200."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from inversus_b200 import InversusCNNPolicy, PPOAgent  # noqa: E402

torch.manual_seed(0)
m = InversusCNNPolicy().cuda()
agent = PPOAgent(m, device="cuda", precision="bf16", batch_size=8192)


def bench(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B in (4096, 65536):
    g = (torch.rand(B, 12, 10, 15, device="cuda") > 0.7).to(torch.bfloat16)
    e = torch.rand(B, 4, device="cuda")
    with torch.no_grad():
        ms = bench(lambda: m.infer(g, e))
    print(f"infer B={B}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.2f} M samples/s  {B * 93e6 / (ms * 1e-3) / 1e12:.1f} TFLOP/s")
    if B == 4096:
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                m.infer(g, e)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

B = 8192
g = (torch.rand(B, 12, 10, 15, device="cuda") > 0.7).to(torch.bfloat16)
e = torch.rand(B, 4, device="cuda")
act = torch.randint(0, 13, (B,), device="cuda")
adv, ret, olp = torch.randn(B, device="cuda"), torch.randn(B, device="cuda"), -torch.rand(B, device="cuda")
agent.epochs = 1


def train_step():
    agent._run_epochs(B, lambda idx: (g[idx], e[idx]), act, olp, adv, ret)


ms = bench(train_step, 5)
print(f"train minibatch B={B}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.3f} M samples/s")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
