"""Where a strong-scaled window's time goes (run under torchrun, N >= 2): per-rank K-step window times with and without
the energy gather, with and without the start rendezvous, for each gather method. Uses bench.py's own DeviceLoop.
usage: python -m torch.distributed.run --nproc-per-node N tools/gather_diag.py [steps]"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

import torch
import torch.distributed as dist
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tdev = torch.device("cuda", local)
dev = gf.Device(local)
stream = torch.cuda.Stream(device=tdev)
dist.init_process_group("nccl", device_id=tdev)
uid = [gf.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = gf.Comm(dev, world, rank, uid[0])
handles = [None] * world
dist.all_gather_object(handles, comm.gather_alloc(bench.REPLICAS_TOTAL))
comm.gather_attach(handles)


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


w, sets, (lo, hi) = bench.pose_sets(W, rank, world, True, 2 * world)
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
kern = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
loop = bench.DeviceLoop(torch, gf, dev, kern, sets, bench.N_ATOMS, stream, comm=comm, gather_mode="none", lo=lo, total=bench.REPLICAS_TOTAL)


def per_rank(ms_list):
    t = torch.tensor(ms_list, dtype=torch.float64, device=tdev)
    allt = torch.empty(world * len(ms_list), dtype=torch.float64, device=tdev)
    dist.all_gather_into_tensor(allt, t)
    return allt.view(world, len(ms_list)).cpu().numpy()


for mode in ("none", "ll", "push", "nccl"):
    for rdv in (True, False):
        loop.gather_mode = mode
        bench.RENDEZVOUS = rdv
        # "none" never rendezvouses inside time_windows (gather_mode == none): do it by hand through a wrapper barrier
        ms, _ = loop.time_windows(steps, 3, 7, True, True, barrier, None)
        allms = per_rank(ms) * 1e3 / steps
        if rank == 0:
            med = np.median(allms, axis=1)
            print(f"gather={mode:5s} rendezvous={int(rdv)}  us/step per rank (median of 7 windows): " + " ".join(f"{x:6.2f}" for x in med) +
                  f"   max-over-ranks median {np.median(allms.max(axis=0)):6.2f}", flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
