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
_pins = [torch.empty(n, dtype=torch.int8, pin_memory=True) for _ in range(4)]
acts = [t.numpy() for t in _pins]
for a in acts:
    a[:] = rs.randint(0, 13, size=n)
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

# the trainer's view: reward / done / info dense + the finished episodes as a compact list
os.environ.pop("INV_HOST_CHUNKS", None)
ev = sim.host_event_buffers(pinned=True)
dense3 = {k: small[k] for k in ("reward", "done", "info")}
for what, fn in (("reward/done/info only (no extra, no episode stats)", lambda a: sim.step_host(a, None, dense3)),
                 ("reward/done/info + finished-episode list", lambda a: sim.step_host_events(a, None, ev))):
    for chunks in (0, 1, 2, 4):
        if chunks:
            os.environ["INV_HOST_CHUNKS"] = str(chunks)
        else:
            os.environ.pop("INV_HOST_CHUNKS", None)
        for k in range(20):
            fn(acts[k % 4])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cnt = 0
        for k in range(100):
            r = fn(acts[k % 4])
            cnt += 0 if r is None else len(r)
        dt = (time.perf_counter() - t0) / 100
        print(f"{what}: chunks={chunks or 'default'}  {dt * 1e3:.3f} ms/step  {n / dt / 1e6:.1f} M env-steps/s  "
              f"finished episodes per step {cnt / 100:.0f}", flush=True)
