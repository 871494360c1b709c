"""Scratch probe: C5 step with the kernel reading positions from / storing forces to PINNED HOST memory directly
(zero-copy over PCIe) versus gfb_kernel_execute_host's chunk pipeline. CUDA events + wall clock."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = gf.Device(0)
tdev = torch.device("cuda:0")
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
w = W.c5_sharded_replicas(n_local=R)
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, 0) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
A = w.n_atoms
h_pos = torch.from_numpy(w.pos).pin_memory()
h_f = torch.zeros(R, A, 3, dtype=torch.float64).pin_memory()
d_e = torch.zeros(R, dtype=torch.float64, device=tdev)
h_e = torch.zeros(R, dtype=torch.float64).pin_memory()
d_pos = h_pos.to(tdev)
d_f = torch.zeros(R, A, 3, dtype=torch.float64, device=tdev)


def run(pos_ptr, f_ptr, iters=10, label=""):
    for _ in range(2):
        k.execute_device(R, A, pos_ptr, d_e.data_ptr(), None, f_ptr, gf.FORCE_F64_STORE, 0, None, side.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        d_e.zero_()
        k.execute_device(R, A, pos_ptr, d_e.data_ptr(), None, f_ptr, gf.FORCE_F64_STORE, 0, None, side.cuda_stream)
        h_e.copy_(d_e, non_blocking=True)
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / iters * 1e3
    print(f"{label:40s} {ms:8.3f} ms/step  {w.evals / ms / 1e6:8.2f} G evals/s", flush=True)


run(d_pos.data_ptr(), d_f.data_ptr(), label="device pos, device forces")
run(h_pos.data_ptr(), d_f.data_ptr(), label="HOST pos (zero-copy), device forces")
run(d_pos.data_ptr(), h_f.data_ptr(), label="device pos, HOST forces (zero-copy)")
run(h_pos.data_ptr(), h_f.data_ptr(), label="HOST pos, HOST forces (zero-copy both)")
f_ref = d_f.cpu()
print("forces equal:", bool(torch.equal(f_ref, h_f)))
pos_np, f_np, e_np = h_pos.numpy(), h_f.numpy(), h_e.numpy()
for _ in range(2):
    k.execute_host(pos_np, forces=f_np, energies_out=e_np)
t0 = time.perf_counter()
for _ in range(10):
    k.execute_host(pos_np, forces=f_np, energies_out=e_np)
ms = (time.perf_counter() - t0) / 10 * 1e3
print(f"{'execute_host chunk pipeline':40s} {ms:8.3f} ms/step  {w.evals / ms / 1e6:8.2f} G evals/s")

# ---- hybrids: one direction zero-copy inside the kernel, the other by the copy engine in chunks on a second stream
copy_stream = torch.cuda.Stream()
NCH = int(os.environ.get("NCH", "8"))
bounds = [R * c // NCH for c in range(NCH + 1)]
esz = A * 3 * 8


def hybrid(mode, iters=10):
    def step():
        d_e.zero_()
        evs = []
        for c in range(NCH):
            r0, r1 = bounds[c], bounds[c + 1]
            if mode == "zc_read_dma_write":
                k.execute_device(r1 - r0, A, h_pos.data_ptr() + r0 * esz, d_e.data_ptr() + 8 * r0, None, d_f.data_ptr() + r0 * esz,
                                 gf.FORCE_F64_STORE, 0, None, side.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(side)
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    h_f[r0:r1].copy_(d_f[r0:r1], non_blocking=True)
            else:  # dma_read_zc_write
                with torch.cuda.stream(copy_stream):
                    d_pos[r0:r1].copy_(h_pos[r0:r1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                side.wait_event(ev)
                k.execute_device(r1 - r0, A, d_pos.data_ptr() + r0 * esz, d_e.data_ptr() + 8 * r0, None, h_f.data_ptr() + r0 * esz,
                                 gf.FORCE_F64_STORE, 0, None, side.cuda_stream)
        h_e.copy_(d_e, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    ms = (time.perf_counter() - t0) / iters * 1e3
    print(f"{mode + ' x' + str(NCH):40s} {ms:8.3f} ms/step  {w.evals / ms / 1e6:8.2f} G evals/s", flush=True)


hybrid("zc_read_dma_write")
hybrid("dma_read_zc_write")
print("forces equal:", bool(torch.equal(f_ref, h_f)))
