"""Scratch probe: device-resident step time of the cubic B-spline path on the C5 shape (R replicas x 47 atoms x 3 grids
of 192^3), CUDA events. Usage: python tools/bspline_perf.py [replicas] [bspline|bspline_points|points|hermite]   (points / hermite =
tricubic Hermite, interpolation method 2, on raw points / on records)"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
LNAME = sys.argv[2] if len(sys.argv) > 2 else "bspline"
LAYOUT = {v: k for k, v in gf.LAYOUT_NAMES.items()}[LNAME]      # bspline | bspline_points | hermite | points
dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
w = W.c5_sharded_replicas(n_local=R)
for prec, pname in ((0, "mixed"), (1, "double")):
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, prec, layout=LAYOUT) for v in w.grids]
    print(pname, "grid bytes", [g.device_bytes for g in grids])
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    n = R * w.n_atoms
    stride = ((n + 31) // 32) * 32
    d_pos = torch.from_numpy(w.pos).to(tdev)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = torch.zeros(R, dtype=torch.float64, device=tdev)
    for fm, fname in ((gf.FORCE_FIXED_ADD, "fixed_add"), (gf.FORCE_F64_STORE, "f64_store")):
        for i in range(3):
            k.execute_device(R, w.n_atoms, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), fm, stride, None, side.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for i in range(iters):
            k.execute_device(R, w.n_atoms, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), fm, stride, None, side.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        print(f"{LNAME} {pname} C5x{R} path {k.eval_path()} {fname}: {us:9.1f} us  {w.evals / us / 1e3:8.2f} G evals/s", flush=True)
    k.close()
    for g in grids:
        g.close()
