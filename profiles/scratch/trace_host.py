"""Scratch: INV_HOST_TRACE=1 python profiles/scratch/trace_host.py -> per-call laps of inv_step_host_events on stderr."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from inversus_b200 import BatchedInversus  # noqa: E402

n = 1 << 20
sim = BatchedInversus(n, "dummy", "hard", 500, seed=0, obs_dtype="f32", auto_reset=True)
sim.reset()
ev = sim.host_event_buffers(pinned=True)
rs = np.random.RandomState(0)
acts = [rs.randint(0, 13, size=n).astype(np.int8) for _ in range(4)]
for k in range(30):
    sim.step_host_events(acts[k % 4], None, ev)
