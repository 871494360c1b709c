"""Scratch timing probe (not the bench): device-resident C3 / C5-shard kernel time with CUDA events, and the
random-sector-gather microbenchmark at several footprints."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

dev = gf.Device(0)
print(dev.props())
for mb in (16, 32, 64, 96, 256, 1024):
    print(f"sector gather {mb:5d} MB: {dev.bench_sector_gather(mb << 20, 1 << 24, 10):8.1f} GB/s")

tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
stream = side.cuda_stream
assert stream != 0

def time_kernel(k, R, P, d_pos, fmode, iters=20, order=None):
    n = R * P
    stride = ((n + 31) // 32) * 32
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = torch.zeros(R, dtype=torch.float64, device=tdev)
    for _ in range(3):
        k.execute_device(R, P, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), fmode, stride, order, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        k.execute_device(R, P, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), fmode, stride, order, stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

t = time.time()
w = W.c3_million_atoms()
print("gen c3", time.time() - t)
for prec in (0, 1):
    g = gf.Grid(dev, w.counts, w.spacing, w.origin, w.grids[0], prec)
    k = gf.Kernel(dev, [g], w.scaling)
    d_pos = torch.from_numpy(w.pos).to(tdev)
    for fm, name in ((gf.FORCE_FIXED_ADD, "fixed"), (gf.FORCE_F64_STORE, "f64store")):
        us = time_kernel(k, 1, w.n_atoms, d_pos, fm)
        print(f"C3 prec={prec} {name}: {us:8.2f} us  {w.evals / us / 1e3:8.2f} G evals/s")
    d_order = torch.empty(w.n_atoms, dtype=torch.int32, device=tdev)
    k.sort_atoms(1, w.n_atoms, d_pos.data_ptr(), d_order.data_ptr(), stream)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        k.sort_atoms(1, w.n_atoms, d_pos.data_ptr(), d_order.data_ptr(), stream)
    torch.cuda.synchronize()
    print(f"   sort: {(time.time() - t0) / 5 * 1e6:.1f} us")
    us = time_kernel(k, 1, w.n_atoms, d_pos, gf.FORCE_FIXED_ADD, order=d_order.data_ptr())
    print(f"C3 prec={prec} fixed sorted: {us:8.2f} us  {w.evals / us / 1e3:8.2f} G evals/s")
    k.close(); g.close()

t = time.time()
w = W.c5_sharded_replicas()
print("gen c5", time.time() - t)
for prec in (0,):
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, prec) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling)
    d_pos = torch.from_numpy(w.pos).to(tdev)
    for R in (65536, 8192):
        for fm, name in ((gf.FORCE_FIXED_ADD, "fixed"), (gf.FORCE_F64_STORE, "f64store")):
            us = time_kernel(k, R, w.n_atoms, d_pos, fm)
            print(f"C5 R={R} prec={prec} {name}: {us:8.2f} us  {R * 47 * 3 / us / 1e3:8.2f} G evals/s")
