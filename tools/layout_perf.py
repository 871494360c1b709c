"""Device-resident kernel time per layout for C3 / C4 / C5 (CUDA events on the launching stream, rotating sets)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W
import bench

dev = gf.Device(0)
tdev = torch.device("cuda:0")
stream = torch.cuda.Stream()
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c3", "c4", "c5"]
layouts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3]
for name in which:
    if name == "c3":
        w = W.c3_million_atoms()
        rng = np.random.default_rng(99)
        L = w.spacing[0] * (w.counts[0] - 1)
        sets = [w.pos] + [rng.uniform(0, 0.999 * L, size=w.pos.shape) for _ in range(7)]
    elif name == "c4":
        w = W.c4_batched_replicas()
        sets = [w.pos] + [W.ligand_replicas(w.n_replicas, W.ligand47()[0].mean(axis=0), seed=W.SEED + 10 + i, escape_shift=(1.0, 0, 0)) for i in range(15)]
    else:
        w = W.c5_sharded_replicas()
        sets = [w.pos]
    pos_sets = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
    for layout in layouts:
        grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, 0, layout=layout) for v in w.grids]
        k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
        for fm, fname in ((gf.FORCE_FIXED_ADD, "fixed"), (gf.FORCE_F64_STORE, "store")):
            secs, launches, _ = bench.time_device_steps(torch, gf, k, pos_sets, w.n_replicas, w.n_atoms, 100, 5, stream, force_mode=fm)
            us = secs / 100 * 1e6
            print(f"{name} layout={gf.LAYOUT_NAMES[layout]:5s} {fname:5s} grid_bytes={sum(g.device_bytes for g in grids)/2**20:7.1f}MB "
                  f"{us:8.2f} us  {w.evals/us/1e3:7.2f} G evals/s  hbm_frac={w.evals*bench.b_alg(w.n_grids)/us/1e3/6541.8:.3f}", flush=True)
        k.close()
        for g in grids: g.close()
