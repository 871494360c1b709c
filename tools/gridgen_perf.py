"""Grid generation throughput: 208x278x231 points x 9133 receptor atoms (the reference's test shape,
python/tests/test_simple_grid_energy.py:29-31) on the GPU vs the CPU oracle on a bounded sub-grid."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W
from oracle import bindings
dev = gf.Device(0)
rng = np.random.default_rng(0)
n = 9133
counts, sp, og = W.TEST_GRID_COUNTS, (W.TEST_GRID_SPACING,) * 3, W.TEST_GRID_ORIGIN
L = np.array(sp) * (np.array(counts) - 1)
pos = np.array(og) + rng.uniform(-0.3, 1.3, size=(n, 3)) * L
q, sg, ep = rng.normal(size=n) * 0.4, rng.uniform(0.1, 0.2, n), rng.uniform(0.1, 1.0, n)
pairs = np.prod(counts) * n
for t in ("charge", "ljr", "lja"):
    gf.Grid.generate(dev, (32, 32, 32), sp, og, t, pos, q, sg, ep, want_values=False, want_grid=False)   # warm-up
    t0 = time.perf_counter()
    grid, vals = gf.Grid.generate(dev, counts, sp, og, t, pos, q, sg, ep, want_values=False, want_grid=True)
    dt = time.perf_counter() - t0
    print(f"GPU {t:6s}: {dt*1e3:8.1f} ms incl. repack  {pairs/dt/1e9:8.1f} G pairs/s")
    grid.close()
sub = (24, 24, 24)
t0 = time.perf_counter()
bindings.port_generate_grid(sub, sp, og, "ljr", pos, q, sg, ep, n_threads=1)
dt = time.perf_counter() - t0
print(f"CPU oracle 1 thread ljr: {np.prod(sub)*n/dt/1e9:.4f} G pairs/s  -> full grid would take {pairs/(np.prod(sub)*n/dt):.0f} s")
