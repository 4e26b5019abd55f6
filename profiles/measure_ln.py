"""Fused (conv bias +) (residual +) LayerNorm + ReLU kernels alone (csrc/policy_kernels.cu) and the
packed-state encoder (csrc/encoder_kernels.cu): CUDA-event time at the PPO minibatch shapes and
the fraction of the measured HBM copy peak their algorithmic traffic amounts to.
    python profiles/measure_ln.py [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200.fused_ops import layer_norm_relu  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0


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


torch.manual_seed(0)
for D, C in ((9600, 64), (19200, 128)):
    for with_res in (False, True):
        x = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
        r = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_() if with_res else None
        g = (torch.rand(D, device="cuda") + 0.5).to(torch.bfloat16).requires_grad_()
        b = (torch.rand(D, device="cuda") - 0.5).to(torch.bfloat16).requires_grad_()
        cb = torch.randn(C, device="cuda").to(torch.bfloat16).requires_grad_()
        dy = torch.randn(B, D, device="cuda").to(torch.bfloat16)
        with torch.no_grad():
            ms_f = bench(lambda: layer_norm_relu(x, g, b, 1e-5, residual=r, channel_bias=cb, channels=C))
        y = layer_norm_relu(x, g, b, 1e-5, residual=r, channel_bias=cb, channels=C)

        def bwd():
            torch.autograd.grad(y, [x, g, b, cb] + ([r] if with_res else []), dy, retain_graph=True)
        ms_b = bench(bwd)
        if D == 19200:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    bwd()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))
        fb = B * D * 2 * (3 if with_res else 2)   # read x (+res), write y
        bb = B * D * 2 * (4 if with_res else 3)   # read x, dy (+res), write dx
        print(json.dumps({"kernel": "ln_relu", "B": B, "D": D, "residual": with_res,
                          "fwd_ms": round(ms_f, 4), "fwd_gbs": round(fb / ms_f / 1e6, 1), "fwd_frac": round(fb / ms_f / 1e6 / peak, 3),
                          "bwd_ms(incl. reduce + grad casts)": round(ms_b, 4), "bwd_gbs": round(bb / ms_b / 1e6, 1),
                          "bwd_frac": round(bb / ms_b / 1e6 / peak, 3)}), flush=True)
