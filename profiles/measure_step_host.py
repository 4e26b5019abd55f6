"""inv_step_host timing (1 048 576 envs, fp32 handle): small outputs only ("observations stay on the
GPU") and the full fp32-observation delivery, for 1/2/4/8 env chunks (INV_HOST_CHUNKS)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from inversus_b200 import BatchedInversus  # noqa: E402

n = 1 << 20
sim = BatchedInversus(n, "dummy", "hard", 500, seed=0, obs_dtype="f32", auto_reset=True)
sim.reset()
out = sim.host_buffers(pinned=True)
small = {k: v for k, v in out.items() if not k.startswith("obs")}
rs = np.random.RandomState(0)
acts = [rs.randint(0, 13, size=n).astype(np.int8) for _ in range(4)]
for what, bufs, steps in (("small outputs only", small, 30), ("full fp32 observations", out, 6)):
    for chunks in (1, 2, 4, 8):
        os.environ["INV_HOST_CHUNKS"] = str(chunks)
        sim.set_host_path(0 if bufs is small else min(os.cpu_count() or 1, 32), -1.0)
        for k in range(4):
            sim.step_host(acts[k % 4], None, bufs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            sim.step_host(acts[k % 4], None, bufs)
        dt = (time.perf_counter() - t0) / steps
        print(f"{what}: chunks={chunks}  {dt * 1e3:.3f} ms/step  {n / dt / 1e6:.1f} M env-steps/s  {sim.host_path()}", flush=True)
