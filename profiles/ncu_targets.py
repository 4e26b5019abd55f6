"""Small driver for `ncu --set full`: launches every kernel of the path a few times at 1 048 576 envs
(step f32/bf16/u8, selfplay f32, reset, obs-from-packed) plus the PPO-side kernels (GAE, fused
LayerNorm+ReLU forward/backward), so one capture covers them all.

    ncu --set full --clock-control none --import-source on -k regex:'inv_kernel|gae_kernel|ln_relu' \
        -o gpurun_out/prof_all python profiles/ncu_targets.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200 import BatchedInversus, compute_gae  # noqa: E402
from inversus_b200.fused_ops import layer_norm_relu  # noqa: E402

n = 1 << 20
g = torch.Generator(device="cuda")
g.manual_seed(0)
a = torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g)
b = torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g)
for mode, dt in (("dummy", "f32"), ("dummy", "bf16"), ("dummy", "u8"), ("selfplay", "f32")):
    sim = BatchedInversus(n, mode, "hard", 500, seed=0, obs_dtype=dt)
    sim.reset()
    for _ in range(3):
        sim.step(a, b if mode == "selfplay" else None)
    if (mode, dt) == ("dummy", "f32"):
        snap = sim.snapshot()
        sim.obs_from_packed(snap, 0)
    torch.cuda.synchronize()
    sim.close()
    del sim
    torch.cuda.empty_cache()

T, N = 128, 1 << 18
r, v = torch.randn(T, N, device="cuda"), torch.randn(T, N, device="cuda")
d = (torch.rand(T, N, device="cuda") < 0.05).to(torch.uint8)
compute_gae(r, v, d, None, 0.99, 0.95)

B = 8192
for D, with_res in ((4800, False), (19200, False), (19200, True)):
    x = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
    res = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_() if with_res else None
    ga = torch.ones(D, device="cuda", dtype=torch.bfloat16).requires_grad_()
    be = torch.zeros(D, device="cuda", dtype=torch.bfloat16).requires_grad_()
    y = layer_norm_relu(x, ga, be, 1e-5, residual=res)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ncu targets done")
