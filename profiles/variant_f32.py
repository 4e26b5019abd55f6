"""Time the f32 step / reset / obs-from-packed kernels of whichever library INVERSUS_B200_LIB names
(kernel-shape experiments). One line per kernel: median ms, env-steps/s, algorithmic GB/s."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200 import BatchedInversus, constants  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dts = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32"]
tag = os.path.basename(os.environ.get("INVERSUS_B200_LIB", "default"))
ELEM = {"f32": 4, "bf16": 2, "u8": 1}


def timeit(fn, iters=200, warm=30):
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
    return statistics.median(ms), ev[0].elapsed_time(ev[iters]) / iters


g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = [torch.randint(0, 13, (n,), device="cuda", dtype=torch.int8, generator=g) for _ in range(8)]
for dt in dts:
    for mode in ("dummy", "selfplay"):
        sim = BatchedInversus(n, mode, "hard", 500, seed=0, obs_dtype=dt, auto_reset=True)
        sim.reset()
        k = [0]

        def step():
            k[0] += 1
            sim.step(acts[k[0] % 8], acts[(k[0] + 3) % 8] if mode == "selfplay" else None)
        ms, avg = timeit(step)
        b = constants.algorithmic_bytes_per_env_step(ELEM[dt], mode == "selfplay") * n
        print(json.dumps({"lib": tag, "kernel": f"step_{mode}_{dt}", "median_ms": round(ms, 4), "avg_ms": round(avg, 4),
                          "env_steps_per_s": round(n / (avg * 1e-3)), "alg_gbs": round(b / avg / 1e6, 1)}), flush=True)
        if mode == "dummy":
            snap = sim.snapshot()
            obs = torch.empty_like(sim.obs)
            ext = torch.empty_like(sim.extra)
            import ctypes as C
            from inversus_b200 import _capi
            lib = _capi.load()
            st = int(torch.cuda.current_stream().cuda_stream)

            def k3():
                lib.inv_obs_from_packed(sim._h.ptr, snap.data_ptr(), n, n, 0, sim._dt, obs.data_ptr(), ext.data_ptr(), st)
            ms, avg = timeit(k3, iters=50, warm=5)
            b = (1800 * ELEM[dt] + 16 + 80) * n
            print(json.dumps({"lib": tag, "kernel": f"obs_from_packed_{dt}", "median_ms": round(ms, 4),
                              "avg_ms": round(avg, 4), "alg_gbs": round(b / avg / 1e6, 1)}), flush=True)
        sim.close()
        del sim
        torch.cuda.empty_cache()
