"""Scratch probe: does an event record / a wait on an already-complete event between two launches defeat PDL overlap?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W
dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
other = torch.cuda.Stream()
torch.cuda.set_stream(side)
w = W.c5_sharded_replicas()
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, 0) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
R, A = w.n_replicas, w.n_atoms
n = R * A
stride = ((n + 31) // 32) * 32
d_pos = torch.from_numpy(w.pos).to(tdev)
d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
d_e = [torch.zeros(R, dtype=torch.float64, device=tdev) for _ in range(3)]
def run(mode, pdl, iters=100):
    k.set_launch_overlap(pdl)
    evs = []
    def step(i):
        if mode == "wait" and i >= 2:
            side.wait_event(evs[i - 2])
        k.execute_device(R, A, d_pos.data_ptr(), d_e[i % 3].data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None,
                         side.cuda_stream, d_energies_clear=d_e[(i + 1) % 3].data_ptr())
        if mode in ("record", "wait"):
            ev = torch.cuda.Event()
            ev.record(side)
            evs.append(ev)
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    evs.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    print(f"mode={mode:7s} pdl={int(pdl)}: {e0.elapsed_time(e1) / iters * 1e3:7.2f} us", flush=True)
for mode in ("plain", "record", "wait"):
    for pdl in (False, True):
        run(mode, pdl)
