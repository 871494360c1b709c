import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import openmmgridforce_b200 as gf
from openmmgridforce_b200 import workloads as W
dev = gf.Device(0)
w = W.c5_sharded_replicas(n_local=6000, n=96)
grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
k = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
tdev = torch.device("cuda:0")
r, a = w.n_replicas, w.n_atoms
n = r * a
stride = ((n + 31) // 32) * 32
pos = torch.from_numpy(w.pos).to(tdev)
stream = torch.cuda.Stream()
res = {}
for pdl in (False, True):
    k.set_launch_overlap(pdl)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = [torch.zeros(r, dtype=torch.float64, device=tdev) for _ in range(2)]
    torch.cuda.synchronize()
    nl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    with torch.cuda.stream(stream):
        for i in range(nl):
            k.execute_device(r, a, pos.data_ptr(), d_e[i % 2].data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None, stream.cuda_stream,
                             d_energies_clear=d_e[(i + 1) % 2].data_ptr())
    stream.synchronize()
    torch.cuda.synchronize()
    res[pdl] = (d_f.cpu().numpy().reshape(3, stride), d_e[(nl - 1) % 2].cpu().numpy())
f0, e0 = res[False]
f1, e1 = res[True]
bad = np.nonzero((f0 != f1).any(axis=0))[0]
print("atoms", n, "tiles", (n + 63) // 64, "differing atoms", bad.size)
if bad.size:
    print("first", bad[:20], "tile", bad[:20] // 64, "lane", bad[:20] % 64)
    t = np.unique(bad // 64)
    print("tiles affected", t.size, t[:20], "min", t.min(), "max", t.max())
    i = bad[0]
    print("f0", f0[:, i], "f1", f1[:, i], "ratio", f1[:, i] / np.where(f0[:, i] == 0, 1, f0[:, i]))
print("energy max rel diff", np.abs(e0 - e1).max() / np.abs(e0).max())
