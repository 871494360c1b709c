"""Scratch probe (torchrun, N >= 2): per-replica energy gather through torch symmetric memory — every rank copies its
512 KB into its slot of every peer's buffer with the copy engine over NVLink (no SMs), then a signal-pad barrier — versus
NCCL all_gather_into_tensor. Checks values and times 200 rounds of each, with a dummy compute kernel running alongside."""
import os
import sys
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
R = 65536
buf = symm.empty(2 * world * R, dtype=torch.float64, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok; multicast", hdl.has_multicast_support, flush=True)
peers = [hdl.get_buffer(p, (2, world, R), torch.float64) for p in range(world)]
mine = torch.full((R,), float(rank + 1), dtype=torch.float64, device=dev)
copy_stream = torch.cuda.Stream()


def symm_gather(step, src):
    b = step % 2
    ev = torch.cuda.Event()
    ev.record()
    copy_stream.wait_event(ev)
    with torch.cuda.stream(copy_stream):
        for p in range(world):
            peers[(rank + p) % world][b, rank].copy_(src, non_blocking=True)
        hdl.barrier(channel=b)


for s in range(4):
    symm_gather(s, mine * (s + 1))
torch.cuda.synchronize()
dist.barrier()
got = peers[rank][1, :, 0].cpu().tolist()          # step 3 -> buffer 1, value (r+1)*4
assert got == [4.0 * (r + 1) for r in range(world)], got
print(rank, "values ok", got, flush=True)

work = torch.empty(64 << 20, dtype=torch.float32, device=dev)
out = torch.empty(world * R, dtype=torch.float64, device=dev)
for name in ("none", "nccl", "symm"):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    pend = None
    for s in range(200):
        work.mul_(1.0001)                            # ~85 us HBM-bound stand-in for the evaluation kernel
        if name == "nccl":
            if pend is not None:
                pend.wait()
            pend = dist.all_gather_into_tensor(out, mine, async_op=True)
        elif name == "symm":
            symm_gather(s, mine)
    if pend is not None:
        pend.wait()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200 * 1e6
    if rank == 0:
        print(f"{name:5s}: {dt:7.1f} us per step", flush=True)
dist.barrier()
dist.destroy_process_group()
