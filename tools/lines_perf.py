"""Scratch timing probe: device-resident C3 / C4 / C5 step time (CUDA events, rotating input sets) under the current
GFB_LINES / GFB_FORCE_PATH environment. Usage: python tools/lines_perf.py [C3] [C4] [C5]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
stream = side.cuda_stream
tag = f"LINES={os.environ.get('GFB_LINES', '1')} FPATH={os.environ.get('GFB_FORCE_PATH', '0')} PDL={os.environ.get('GFB_PDL', '0')}"


def time_kernel(k, R, P, pos_sets, fmode, iters=60):
    n = R * P
    stride = ((n + 31) // 32) * 32
    d_f = [torch.zeros(3 * stride, dtype=torch.int64, device=tdev) for _ in pos_sets]
    d_e = torch.zeros(R, dtype=torch.float64, device=tdev)
    def step(i):
        s = i % len(pos_sets)
        k.execute_device(R, P, pos_sets[s].data_ptr(), d_e.data_ptr(), None, d_f[s].data_ptr(), fmode, stride, None, stream)
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


names = sys.argv[1:] or ["C3", "C4", "C5"]
for name in names:
    if name == "C3":
        w = W.c3_million_atoms()
        rng = np.random.default_rng(99)
        length = w.spacing[0] * (w.counts[0] - 1)
        sets = [w.pos] + [rng.uniform(0.0, 0.999 * length, size=w.pos.shape) for _ in range(7)]
    elif name == "C4":
        w = W.c4_batched_replicas()
        sets = [w.pos] + [W.ligand_replicas(w.n_replicas, W.ligand47()[0].mean(axis=0), seed=W.SEED + 10 + i,
                                            escape_shift=(1.0, 0.0, 0.0)) for i in range(15)]
    else:
        w = W.c5_sharded_replicas()
        sets = [w.pos]
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    k.set_launch_overlap(os.environ.get("GFB_PDL", "0") == "1")
    pos_sets = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
    for fm, fname in ((gf.FORCE_FIXED_ADD, "fixed_add"), (gf.FORCE_F64_STORE, "f64_store"), (gf.FORCE_F64_ADD, "f64_add")):
        us = time_kernel(k, w.n_replicas, w.n_atoms, pos_sets, fm)
        print(f"{tag} {name} {fname}: {us:8.2f} us  {w.evals / us / 1e3:8.2f} G evals/s", flush=True)
    k.close()
    for g in grids:
        g.close()
