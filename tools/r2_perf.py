"""Round-2 timing probe (CUDA events on the launching stream, rotating pose sets so that the footprint between
re-touches exceeds L2). Usage: python tools/r2_perf.py [modes] [strong] [double] [e2e] [c3]

  modes   C5 on one GPU (65,536 replicas x 47 x 3 grids of 192^3): FIXED_ADD / F64_STORE / F32_STORE / energy-only, PDL on
  strong  one rank's shard of C5 at N = 2/4/8 (32,768 / 16,384 / 8,192 replicas), N rotating pose sets:
          PDL off / PDL on / PDL + CUDA graph of the K-step loop
  double  C5 in DOUBLE precision: 256-byte record kernel (GFB_LINES_F64=0 in the environment: general kernel)
  e2e     gfb_kernel_execute_host on C5 with pinned buffers: F64 forces / F32 forces / energy only
  c3      C3 (1M atoms x 256^3): FIXED_ADD / F64_STORE / F32_STORE / energy-only, sorted order on and off
"""
import os
import sys
import time

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
NONE = -1


def time_steps(k, R, P, pos_sets, fmode, iters=60, graph=False, order=None, windows=5):
    n = R * P
    stride = ((n + 31) // 32) * 32
    if fmode == gf.FORCE_FIXED_ADD:
        d_f = [torch.zeros(3 * stride, dtype=torch.int64, device=tdev) for _ in pos_sets]
    elif fmode == gf.FORCE_F32_STORE:
        d_f = [torch.zeros(n * 3, dtype=torch.float32, device=tdev) for _ in pos_sets]
    elif fmode == NONE:
        d_f = [None for _ in pos_sets]
    else:
        d_f = [torch.zeros(n * 3, dtype=torch.float64, device=tdev) for _ in pos_sets]
    d_e = [torch.zeros(R, dtype=torch.float64, device=tdev) for _ in range(2)]

    def step(i):
        s = i % len(pos_sets)
        k.execute_device(R, P, pos_sets[s].data_ptr(), d_e[i % 2].data_ptr(), None, d_f[s].data_ptr() if d_f[s] is not None else None,
                         max(fmode, 0), stride, order, stream, d_energies_clear=d_e[(i + 1) % 2].data_ptr())
    for i in range(2 * len(pos_sets)):
        step(i)
    torch.cuda.synchronize()
    g = None
    if graph:
        gf.Graph.begin(dev, stream)
        for i in range(iters):
            step(i)
        g = gf.Graph.end(dev, stream)
        g.launch(stream)
        torch.cuda.synchronize()
    out = []
    for _ in range(windows):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if g is not None:
            g.launch(stream)
        else:
            for i in range(iters):
                step(i)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / iters * 1e3)
    if g is not None:
        g.close()
    out.sort()
    return out[len(out) // 2], out[0], out[-1]


def report(tag, evals, t):
    med, lo, hi = t
    print(f"{tag:58s} {med:8.2f} us (min {lo:7.2f} max {hi:7.2f})  {evals / med / 1e3:8.2f} G evals/s", flush=True)


names = sys.argv[1:] or ["modes", "strong", "double", "e2e", "c4", "c3", "c3sort"]
MODES = ((gf.FORCE_FIXED_ADD, "fixed_add"), (gf.FORCE_F64_STORE, "f64_store"), (gf.FORCE_F32_STORE, "f32_store"), (NONE, "energy_only"))

if any(n in names for n in ("modes", "strong", "double", "e2e", "sweep")):
    w = W.c5_sharded_replicas()
    w2 = W.c5_sharded_replicas(pose_seed=W.SEED + 101)
    for precision, pname in ((gf.PRECISION_MIXED, "mixed"), (gf.PRECISION_DOUBLE, "double")):
        if precision == gf.PRECISION_DOUBLE and "double" not in names:
            continue
        if precision == gf.PRECISION_MIXED and not any(n in names for n in ("modes", "strong", "e2e", "sweep")):
            continue
        grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, precision) for v in w.grids]
        k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
        path = int(gf.load_library().gfb_kernel_eval_path(k._h))
        full = [torch.from_numpy(w.pos).to(tdev), torch.from_numpy(w2.pos).to(tdev)]
        if "modes" in names or precision == gf.PRECISION_DOUBLE:
            for pdl in (False, True):
                k.set_launch_overlap(pdl)
                for fm, fname in MODES:
                    report(f"C5 {pname} path={path} pdl={int(pdl)} {fname}", w.evals, time_steps(k, w.n_replicas, w.n_atoms, full, fm, iters=40))
        if "strong" in names and precision == gf.PRECISION_MIXED:
            for n_gpu in (2, 4, 8):
                r = w.n_replicas // n_gpu
                # N pose sets of the shard size: the same bytes between re-touches as the single-GPU run
                sets = [full[j % 2][(j // 2) * r:(j // 2 + 1) * r].contiguous() for j in range(n_gpu)]
                for pdl, graph in ((False, False), (True, False), (True, True), (False, True)):
                    k.set_launch_overlap(pdl)
                    t = time_steps(k, r, w.n_atoms, sets, gf.FORCE_FIXED_ADD, iters=20 * n_gpu, graph=graph)
                    report(f"C5 shard 1/{n_gpu} ({r} replicas) pdl={int(pdl)} graph={int(graph)}", r * w.n_atoms * w.n_grids, t)
        if "sweep" in names and precision == gf.PRECISION_MIXED:
            # batch-size sweep on the C5 grids, launch overlap + graph: where the small-launch machinery matters
            for r in (512, 1024, 2048, 3072, 4096, 6144, 8192, 12288, 16384):
                n_sets = max(2, min(32, 65536 // r))
                sets = [full[j % 2][(j // 2) * r:(j // 2 + 1) * r].contiguous() for j in range(n_sets)]
                k.set_launch_overlap(True)
                t = time_steps(k, r, w.n_atoms, sets, gf.FORCE_FIXED_ADD, iters=4 * n_sets, graph=True)
                report(f"C5 grids, {r} replicas per launch, pdl=1 graph=1", r * w.n_atoms * w.n_grids, t)
        if "e2e" in names and precision == gf.PRECISION_MIXED:
            k.set_launch_overlap(False)
            pos_h = torch.from_numpy(w.pos.copy()).pin_memory()
            e_h = torch.zeros(w.n_replicas, dtype=torch.float64).pin_memory()
            f64_h = torch.zeros(w.pos.shape, dtype=torch.float64).pin_memory()
            f32_h = torch.zeros(w.pos.shape, dtype=torch.float32).pin_memory()
            for label, f, fm in (("f64 forces", f64_h, gf.FORCE_F64_STORE), ("f32 forces", f32_h, gf.FORCE_F32_STORE), ("energy only", None, 0)):
                for _ in range(3):
                    k.execute_host(pos_h.numpy(), forces=f.numpy() if f is not None else None, force_mode=fm, want_forces=f is not None,
                                   energies_out=e_h.numpy())
                t0 = time.perf_counter()
                for _ in range(10):
                    k.execute_host(pos_h.numpy(), forces=f.numpy() if f is not None else None, force_mode=fm, want_forces=f is not None,
                                   energies_out=e_h.numpy())
                ms = (time.perf_counter() - t0) / 10 * 1e3
                print(f"C5 e2e execute_host {label:12s} {ms:7.3f} ms  {w.evals / ms / 1e6:7.2f} G evals/s", flush=True)
            print("host copy GB/s (h2d, d2h, both):", dev.bench_host_copy(64 << 20, 5), flush=True)
        k.close()
        for g in grids:
            g.close()
        del full

if "c4" in names:
    w = W.c4_batched_replicas()
    sets = [w.pos] + [W.ligand_replicas(w.n_replicas, W.ligand47()[0].mean(axis=0), seed=W.SEED + 10 + i,
                                        escape_shift=(1.0, 0.0, 0.0)) for i in range(15)]
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    pos_sets = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
    for pdl, graph in ((False, False), (True, False), (True, True)):
        k.set_launch_overlap(pdl)
        for fm, fname in MODES:
            report(f"C4 pdl={int(pdl)} graph={int(graph)} {fname}", w.evals, time_steps(k, w.n_replicas, w.n_atoms, pos_sets, fm, iters=64, graph=graph))
    k.close()
    for g in grids:
        g.close()
    del pos_sets

if "c3" in names or "c3sort" in names:
    w = W.c3_million_atoms()
    rng = np.random.default_rng(99)
    length = w.spacing[0] * (w.counts[0] - 1)
    sets = [w.pos] + [rng.uniform(0.0, 0.999 * length, size=w.pos.shape) for _ in range(7)]
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    pos_sets = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
    for pdl in (False, True):
        k.set_launch_overlap(pdl)
        for fm, fname in MODES:
            report(f"C3 pdl={int(pdl)} {fname}", w.evals, time_steps(k, 1, w.n_atoms, pos_sets, fm, iters=40))
    if "c3sort" not in names:
        k.close()
        for g in grids:
            g.close()
        sys.exit(0) if "c4" not in names else None
    # sorted evaluation order (indirection) and physically sorted positions (what a platform that owns the atom order does)
    k.set_launch_overlap(False)
    d_order = torch.empty(w.n_atoms, dtype=torch.int32, device=tdev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k.sort_atoms(1, w.n_atoms, pos_sets[0].data_ptr(), d_order.data_ptr(), stream)
    e0.record()
    for _ in range(10):
        k.sort_atoms(1, w.n_atoms, pos_sets[0].data_ptr(), d_order.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"C3 sort_atoms (1M atoms): {e0.elapsed_time(e1) / 10 * 1e3:8.2f} us", flush=True)
    report("C3 fixed_add, order = sorted (indirect), 1 set", w.evals,
           time_steps(k, 1, w.n_atoms, pos_sets[:1], gf.FORCE_FIXED_ADD, iters=40, order=d_order.data_ptr()))
    sorted_sets = []
    for p in pos_sets:
        k.sort_atoms(1, w.n_atoms, p.data_ptr(), d_order.data_ptr(), stream)
        torch.cuda.synchronize()
        sorted_sets.append(p.view(-1, 3)[d_order.long()].contiguous().view(1, -1, 3))
    # scaling factors are per atom ordinal: permuting atoms needs a matching kernel state; C3's are U(0.5,1.5) and the
    # timing does not depend on them, so the probe reuses the state
    for fm, fname in MODES:
        report(f"C3 physically sorted positions {fname}", w.evals, time_steps(k, 1, w.n_atoms, sorted_sets, fm, iters=40))
    k.close()
    for g in grids:
        g.close()
