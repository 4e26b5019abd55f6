"""Time every kernel variant of the path at a fixed batch (default 1 048 576 envs) with CUDA events:
fused step (f32/bf16/u8 obs x dummy/selfplay), reset (K2), observation rebuild from packed
snapshots (K3). Prints one JSON line per variant with algorithmic GB/s and the fraction of the
measured HBM copy peak. Run on a B200:  python profiles/measure_variants.py [n_envs]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200 import BatchedInversus, constants  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
ELEM = {"f32": 4, "bf16": 2, "u8": 1}


def timeit(fn, iters=100, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return statistics.median(ms), min(ms)


def report(name, nbytes, ms, ms_min):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"variant": name, "envs": n, "median_ms": round(ms, 4), "min_ms": round(ms_min, 4),
                      "env_per_s": round(n / (ms * 1e-3)), "alg_bytes_per_env": nbytes // n,
                      "alg_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4)}), flush=True)


g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = [torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g) for _ in range(8)]
for mode in ("dummy", "selfplay"):
    for dt in ("f32", "bf16", "u8"):
        sim = BatchedInversus(n, mode, "hard", 500, seed=0, obs_dtype=dt, auto_reset=True)
        sim.reset()
        k = [0]

        def step():
            k[0] += 1
            sim.step(acts[k[0] % 8], acts[(k[0] + 3) % 8] if mode == "selfplay" else None)
        ms, mn = timeit(step)
        report(f"step_{mode}_{dt}", constants.algorithmic_bytes_per_env_step(ELEM[dt], mode == "selfplay") * n, ms, mn)
        if mode == "dummy":
            views = 1
            ms, mn = timeit(sim.reset, iters=30, warm=5)
            report(f"reset_{dt}", (views * (1800 * ELEM[dt] + 16) + 160) * n, ms, mn)
            snap = sim.snapshot()
            ms, mn = timeit(lambda: sim.obs_from_packed(snap, 0), iters=30, warm=5)
            report(f"obs_from_packed_{dt}(incl. torch.empty)", (1800 * ELEM[dt] + 16 + 80) * n, ms, mn)
        sim.close()
        del sim
        torch.cuda.empty_cache()
