import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W
import bench
n = 74 * (1 << 20) // 8
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 8 / 1e9
print(f"H2D {gb/timeit(h2d):.1f} GB/s   D2H {gb/timeit(d2h):.1f} GB/s   both: {gb/timeit(both):.1f} GB/s each direction")
# execute_host timing on C5
dev = gf.Device(0)
w = W.c5_sharded_replicas()
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, 0) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling)
pos_h, _a = bench.pinned_array(w.pos.shape); pos_h[...] = w.pos
f_h, _b = bench.pinned_array(w.pos.shape); e_h, _c = bench.pinned_array((w.n_replicas,))
for env in (None,):
    secs = bench.time_e2e_steps(gf, k, pos_h, f_h, e_h, 20, 3)
    print(f"execute_host pinned: {secs/20*1e3:.3f} ms/step  {w.evals*20/secs/1e9:.2f} G evals/s")
secs = bench.time_e2e_steps(gf, k, w.pos, np.zeros_like(w.pos), np.zeros(w.n_replicas), 5, 2)
print(f"execute_host pageable: {secs/5*1e3:.3f} ms/step")
for ch in (4, 8, 16, 32, 64):
    os.environ["GFB_HOST_CHUNKS"] = str(ch)
    secs = bench.time_e2e_steps(gf, k, pos_h, f_h, e_h, 20, 3)
    print(f"chunks={ch:3d}: {secs/20*1e3:.3f} ms/step  {w.evals*20/secs/1e9:.2f} G evals/s")
