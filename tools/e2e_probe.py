"""Scratch probe: wall-clock time of gfb_kernel_execute_host on C5 / C3 / C4 with pinned host buffers, for the current
GFB_ZEROCOPY_FORCES / GFB_HOST_CHUNKS environment. Usage: python tools/e2e_probe.py [C5] [C3] [C4]"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

dev = gf.Device(0)
tag = f"ZC={os.environ.get('GFB_ZEROCOPY_FORCES', '1')}"
for name in (sys.argv[1:] or ["C5"]):
    w = {"C5": W.c5_sharded_replicas, "C3": W.c3_million_atoms, "C4": W.c4_batched_replicas}[name]()
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, 0) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    pos = torch.from_numpy(np.ascontiguousarray(w.pos)).pin_memory()
    f = torch.zeros_like(pos).pin_memory()
    e = torch.zeros(w.n_replicas, dtype=torch.float64).pin_memory()
    pn, fn, en = pos.numpy(), f.numpy(), e.numpy()
    for chunks in ([None] + [int(c) for c in os.environ.get("SWEEP", "").split(",") if c]):
        if chunks is None:
            os.environ.pop("GFB_HOST_CHUNKS", None)
        else:
            os.environ["GFB_HOST_CHUNKS"] = str(chunks)
        res = []
        for rep in range(3):
            for _ in range(3):
                k.execute_host(pn, forces=fn, energies_out=en)
            t0 = time.perf_counter()
            n = 20
            for _ in range(n):
                k.execute_host(pn, forces=fn, energies_out=en)
            res.append((time.perf_counter() - t0) / n * 1e3)
        print(f"{tag} {name} chunks={chunks}: " + " ".join(f"{r:7.3f}" for r in res) + f" ms  best {w.evals / min(res) / 1e6:7.2f} G evals/s", flush=True)
    k.close()
    for g in grids:
        g.close()
