"""Round-2 driver for `ncu --set full`: one or two launches of every hand-written kernel at the
shapes the bench / PPO update use, so one capture covers them all.

    ncu --set full --clock-control none --import-source on \
        -k regex:'inv_kernel|encode_|ln_relu|conv3x3_wgrad' -o gpurun_out/prof_r2_all python profiles/ncu_targets_r2.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200 import BatchedInversus, InversusCNNPolicy  # noqa: E402
from inversus_b200.fused_ops import PackedStates, conv3x3, encode_layer1, layer_norm_relu  # noqa: E402

n = 1 << 20
g = torch.Generator(device="cuda")
g.manual_seed(0)
a = torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g)
b = torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g)
snap = None
for mode, dt in (("dummy", "f32"), ("selfplay", "f32"), ("dummy", "none")):
    sim = BatchedInversus(n, mode, "hard", 500, seed=0, obs_dtype=dt)
    sim.reset()
    for _ in range(2):
        sim.step(a, b if mode == "selfplay" else None)
    if (mode, dt) == ("dummy", "f32"):
        snap = sim.snapshot()
        sim.obs_from_packed(snap, 0)
    torch.cuda.synchronize()
    sim.close()
    del sim
    torch.cuda.empty_cache()

B = 8192
m = InversusCNNPolicy().cuda()
ps = PackedStates(snap[:, :B].contiguous(), 0)
w1, b1 = m.conv1.weight, m.conv1.bias
g1 = m.norm1.weight.permute(1, 2, 0).reshape(-1)
be1 = m.norm1.bias.permute(1, 2, 0).reshape(-1)
y, _ = encode_layer1(ps, w1, b1, g1, be1)
y.backward(torch.randn_like(y))

for D, C, with_res in ((9600, 64, False), (19200, 128, False), (19200, 128, True)):
    x = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
    res = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_() if with_res else None
    ga = torch.ones(D, device="cuda", dtype=torch.bfloat16).requires_grad_()
    be = torch.zeros(D, device="cuda", dtype=torch.bfloat16).requires_grad_()
    cb = torch.zeros(C, device="cuda", dtype=torch.bfloat16).requires_grad_()
    y = layer_norm_relu(x, ga, be, 1e-5, residual=res, channel_bias=cb, channels=C)
    y.backward(torch.randn_like(y))

for cin in (128, 64):
    x = torch.randn(B, cin, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(128, cin, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
    conv3x3(x, w).backward(torch.randn(B, 128, 10, 15, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
torch.cuda.synchronize()
print("ncu targets done")
