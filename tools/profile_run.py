"""Small fixed launch sequence for ncu: a few device-resident launches of one configuration.
usage: profile_run.py {c3|c5|c4} [precision] [force_mode] [iters]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fmode = int(sys.argv[3]) if len(sys.argv) > 3 else gf.FORCE_FIXED_ADD
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
w = {"c3": W.c3_million_atoms, "c5": lambda: W.c5_sharded_replicas(n_local=8192), "c5full": W.c5_sharded_replicas, "c4": W.c4_batched_replicas}[cfg]()
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, prec) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling)
R, P = w.n_replicas, w.n_atoms
n = R * P
stride = ((n + 31) // 32) * 32
d_pos = torch.from_numpy(w.pos).to(tdev)
d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
d_e = torch.zeros(R, dtype=torch.float64, device=tdev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters // 2:
        e0.record()
    k.execute_device(R, P, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), fmode, stride, None, side.cuda_stream)
e1.record()
torch.cuda.synchronize()
print(cfg, "prec", prec, "fmode", fmode, "us/launch", e0.elapsed_time(e1) / (iters - iters // 2) * 1e3, "E0", d_e[0].item())
