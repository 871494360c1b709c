"""Host-path latency of SMALL batches (R replicas x 47 atoms x 3 grids, R x 47 <= 4096 particles): the host-mapped
one-launch path against the chunked copy pipeline (GFB_SMALL_BATCH=0 in a second process), ctypes loop.
    python tools/small_batch_perf.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = gf.Device(0)
w = W.c2_single_ligand()
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
kern = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
rng = np.random.default_rng(0)
print("GFB_SMALL_BATCH =", os.environ.get("GFB_SMALL_BATCH", "1"))
for r in (1, 2, 8, 21, 64, 87):
    pos = np.stack([w.pos.reshape(-1, 3) + rng.uniform(-0.2, 0.2, size=3) for _ in range(r)])
    f = np.zeros_like(pos)
    e = np.zeros(r)
    for _ in range(50):
        kern.execute_host(pos, forces=f, energies_out=e)
    t0 = time.perf_counter()
    for _ in range(steps):
        kern.execute_host(pos, forces=f, energies_out=e)
    dt = (time.perf_counter() - t0) / steps
    print(f"R = {r:3d}: {dt * 1e6:7.2f} us per call   E[0] = {e[0]:.9f}  sum = {e.sum():.9f}", flush=True)
