"""Scratch probe: per-call latency of the host path for one ligand (configs[1]) — zero-copy small path vs the copy
pipeline (GFB_SMALL_PATH=0) — and for small batches. Usage: python tools/latency_probe.py"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

dev = gf.Device(0)
w = W.c2_single_ligand()
tag = "small_path=" + os.environ.get("GFB_SMALL_PATH", "1")
for prec, pname in ((0, "mixed"), (1, "double")):
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, prec) for v in w.grids]
    for ng in (3, 1):
        k = gf.Kernel(dev, grids[:ng], w.scaling[:ng], oob_k=w.oob_k[:ng])
        for mode, mname in ((gf.FORCE_F64_STORE, "store"), (gf.FORCE_F64_ADD, "add")):
            f = np.zeros_like(w.pos)
            e = np.zeros(1)
            for _ in range(200):
                k.execute_host(w.pos, forces=f, energies_out=e, force_mode=mode)
            n = 3000
            t0 = time.perf_counter()
            for _ in range(n):
                k.execute_host(w.pos, forces=f, energies_out=e, force_mode=mode)
            us = (time.perf_counter() - t0) / n * 1e6
            print(f"{tag} {pname} grids={ng} {mname}: {us:6.2f} us/call  E={e[0]:.9f}", flush=True)
        k.close()
    for g in grids:
        g.close()
