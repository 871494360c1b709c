"""Small fixed launch sequence for ncu: a few device-resident launches of one configuration, rotating pose sets the way
bench.py does (so that the captured launch sees the cache state of the timed loop, not a re-run of the same inputs).
usage: profile_run.py {c3|c4|c5full|c5shard2|c5shard4|c5shard8} [precision] [force_mode (-1 = energy only)] [iters] [pdl]
       [layout: auto|cells|rows|pairs|bspline|points|hermite]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fmode = int(sys.argv[3]) if len(sys.argv) > 3 else gf.FORCE_FIXED_ADD
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
pdl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
layout = {v: k for k, v in gf.LAYOUT_NAMES.items()}[sys.argv[6]] if len(sys.argv) > 6 else None
dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
if cfg == "c3":
    w = W.c3_million_atoms()
    rng = np.random.default_rng(99)
    length = w.spacing[0] * (w.counts[0] - 1)
    sets = [w.pos] + [rng.uniform(0.0, 0.999 * length, size=w.pos.shape) for _ in range(3)]
elif cfg == "c4":
    w = W.c4_batched_replicas()
    sets = [w.pos] + [W.ligand_replicas(w.n_replicas, W.ligand47()[0].mean(axis=0), seed=W.SEED + 10 + i, escape_shift=(1.0, 0.0, 0.0))
                      for i in range(7)]
else:
    n_gpu = {"c5full": 1, "c5shard2": 2, "c5shard4": 4, "c5shard8": 8}[cfg]
    r = 65536 // n_gpu
    w = W.c5_sharded_replicas(n_local=r)
    half = 0.5 * w.spacing[0] * 191
    sets = [w.pos] + [W.ligand_replicas(r, (half, half, half), seed=W.SEED + 7919 * j, escape_shift=(0.9, 0.0, 0.0)) for j in range(1, 2 * n_gpu)]
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, prec, layout=layout) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
k.set_launch_overlap(bool(pdl))
R, P = w.n_replicas, w.n_atoms
n = R * P
stride = ((n + 31) // 32) * 32
d_pos = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
if fmode == gf.FORCE_FIXED_ADD:
    d_f = [torch.zeros(3 * stride, dtype=torch.int64, device=tdev) for _ in sets]
elif fmode == gf.FORCE_F32_STORE:
    d_f = [torch.zeros(3 * n, dtype=torch.float32, device=tdev) for _ in sets]
elif fmode < 0:
    d_f = [None for _ in sets]
else:
    d_f = [torch.zeros(3 * n, dtype=torch.float64, device=tdev) for _ in sets]
d_e = [torch.zeros(R, dtype=torch.float64, device=tdev) for _ in range(2)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters // 2:
        e0.record()
    s = i % len(sets)
    k.execute_device(R, P, d_pos[s].data_ptr(), d_e[i % 2].data_ptr(), None, d_f[s].data_ptr() if d_f[s] is not None else None,
                     max(fmode, 0), stride, None, side.cuda_stream, d_energies_clear=d_e[(i + 1) % 2].data_ptr())
e1.record()
torch.cuda.synchronize()
print(cfg, "prec", prec, "fmode", fmode, "path", k.eval_path(), "us/launch", e0.elapsed_time(e1) / (iters - iters // 2) * 1e3, "E0", d_e[(iters - 1) % 2][0].item())
