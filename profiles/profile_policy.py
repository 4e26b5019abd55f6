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

# ---- the packed-state path (round 2): block 1 straight from the 80-byte env states, no observation
from inversus_b200 import BatchedInversus  # noqa: E402
from inversus_b200.fused_ops import PackedStates  # noqa: E402

for Bp in (8192, 65536):
    sim = BatchedInversus(Bp, "dummy", "hard", 500, seed=0, obs_dtype="none")
    sim.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    for _ in range(40):
        sim.step(torch.randint(0, 13, (Bp,), device="cuda", generator=gen).to(torch.int8))
    ps = PackedStates(sim.snapshot(), 0)
    with torch.no_grad():
        ms_p = bench(lambda: m.infer(ps, None))
    print(f"infer from packed states B={Bp}: {ms_p:.3f} ms  {Bp / ms_p * 1e3 / 1e6:.2f} M samples/s  "
          f"{Bp * 93e6 / (ms_p * 1e-3) / 1e12:.1f} TFLOP/s")
    if Bp == 65536:
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                m.infer(ps, None)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
    for Bt in ((8192, 32768) if Bp == 65536 else ()):
        agent_p = PPOAgent(m, device="cuda", precision="bf16", batch_size=Bt, epochs=1, packed_encoder=True)
        sel = ps.planes[:, :Bt].contiguous()
        actp = torch.randint(0, 13, (Bt,), device="cuda")
        advp, retp, olpp = torch.randn(Bt, device="cuda"), torch.randn(Bt, device="cuda"), -torch.rand(Bt, device="cuda")

        def train_step_packed():
            agent_p._run_epochs(Bt, lambda idx: (PackedStates(sel.index_select(1, idx).contiguous(), 0), None),
                                actp, olpp, advp, retp)
        ms_t = bench(train_step_packed, 5)
        print(f"train minibatch from packed states B={Bt}: {ms_t:.3f} ms  {Bt / ms_t * 1e3 / 1e6:.3f} M samples/s  "
              f"{Bt * 3 * 93e6 / (ms_t * 1e-3) / 1e12:.0f} TFLOP/s")
        if Bt == 8192:
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    train_step_packed()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
    sim.close()

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
