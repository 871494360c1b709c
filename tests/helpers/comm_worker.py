"""One rank of a multi-process gfb_comm run (tests/test_gpu_multi.py launches one of these per GPU):
    python comm_worker.py <rank> <world> <exchange_dir> <n_total>
The NCCL id and the IPC handles travel through files in exchange_dir (any launcher-side channel would do: torchrun's
store, MPI, a socket). Each of three steps evaluates this rank's shard with the fused in-kernel gather, waits for all
peers' slices, and also runs gfb_comm_all_gather (NCCL) on the same energies; rank r writes what it saw to
<exchange_dir>/out<r>.npz."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def put(path, blob):
    with open(path + ".tmp", "wb") as f:
        f.write(blob)
    os.rename(path + ".tmp", path)


def get(path, timeout=180.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit(f"timed out waiting for {path}")
        time.sleep(0.01)
    return open(path, "rb").read()


def main():
    rank, world, xdir, n_total = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    from openmmgridforce_b200.sharding import shard_bounds
    torch.cuda.set_device(rank)
    tdev = torch.device("cuda", rank)
    dev = gf.Device(rank)
    if rank == 0:
        put(os.path.join(xdir, "uid"), gf.Comm.unique_id())
    comm = gf.Comm(dev, world, rank, get(os.path.join(xdir, "uid")))
    w = W.c5_sharded_replicas(n_replicas=n_total, n=48)
    lo, hi = shard_bounds(n_total, world, rank)
    put(os.path.join(xdir, f"ipc{rank}"), comm.gather_alloc(n_total))
    comm.gather_attach([get(os.path.join(xdir, f"ipc{r}")) for r in range(world)])
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    k.set_launch_overlap(True)
    r, a = hi - lo, w.n_atoms
    width = -(-n_total // world)        # padded shard width for the equal-count NCCL all-gather
    stride = ((r * a + 31) // 32) * 32
    rng = np.random.default_rng(100)
    shifts = rng.uniform(-0.02, 0.02, size=(3, 3))          # every step evaluates different poses (same on all ranks)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = [torch.zeros(width, dtype=torch.float64, device=tdev) for _ in range(2)]
    stream = torch.cuda.Stream(device=tdev)
    fused, nccl, ll = [], [], []
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for step in range(3):           # three gathers: both parities, and reuse of a parity
            d_pos = torch.from_numpy(np.ascontiguousarray(w.pos[lo:hi] + shifts[step])).to(tdev)
            cur, nxt = d_e[step % 2], d_e[(step + 1) % 2]
            # device-side rendezvous of all ranks before the step: plain, then held until the step has been enqueued
            comm.rendezvous(stream.cuda_stream, hold=(step == 2))
            if step == 1:               # the stand-alone producer (same protocol) for one of the three gathers
                k.execute_device(r, a, d_pos.data_ptr(), cur.data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None,
                                 stream.cuda_stream, d_energies_clear=nxt.data_ptr())
                comm.gather_push(cur.data_ptr(), r, lo, stream.cuda_stream)
            else:
                k.execute_device_gather(comm, lo, r, a, d_pos.data_ptr(), cur.data_ptr(), d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride,
                                        stream.cuda_stream, d_energies_clear=nxt.data_ptr())
            got = torch.full((n_total,), float("nan"), dtype=torch.float64, device=tdev)
            comm.gather_wait(got.data_ptr(), stream.cuda_stream)
            fused.append(got)
            got_ll = torch.full((n_total,), float("nan"), dtype=torch.float64, device=tdev)
            comm.gather(cur.data_ptr(), r, lo, got_ll.data_ptr(), stream.cuda_stream)      # the one-kernel flag-in-data gather
            ll.append(got_ll)
            padded = torch.empty(world * width, dtype=torch.float64, device=tdev)
            comm.all_gather(cur.data_ptr(), padded.data_ptr(), width, stream.cuda_stream)
            nccl.append(padded)
            if step == 2:
                comm.rendezvous_release()
    stream.synchronize()
    comm.gather_status()
    np.savez(os.path.join(xdir, f"out{rank}.npz"), fused=torch.stack(fused).cpu().numpy(), nccl=torch.stack(nccl).cpu().numpy(),
             ll=torch.stack(ll).cpu().numpy(),
             lo=lo, hi=hi, width=width, shifts=shifts)
    # leave together: a rank must not unmap its gather memory while a peer may still store into it
    put(os.path.join(xdir, f"done{rank}"), b"1")
    for q in range(world):
        get(os.path.join(xdir, f"done{q}"))
    k.close()
    for g in grids:
        g.close()
    comm.close()
    dev.close()


if __name__ == "__main__":
    main()
