"""Generates tests/golden/*.npz by running the reference's OWN kernel (oracle/_ref, built from /root/reference by
oracle/Makefile) on the cases in cases.py. Run in the build container:  python tests/golden/make_golden.py
The .npz files hold inputs and the reference's outputs; tests never need /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from oracle import bindings  # noqa: E402
import cases  # noqa: E402

bindings.build()
assert bindings.ref_available(), "needs /root/reference to build oracle/_ref"
only = sys.argv[1:]       # optional: names to (re)generate; default all
for name, fn in cases.CASES.items():
    if only and name not in only:
        continue
    c = fn()
    pos = np.ascontiguousarray(c["pos"], dtype=np.float64)
    n, g = pos.shape[0], len(c["grids"])
    ref = bindings.RefOracle(n, c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], oob_k=c["oob_k"],
                             inv_power=c["inv_power"], interpolation_method=c.get("interp", 0))
    e_all, f_all = ref.execute(pos)
    ge, gf = [], []
    for k in range(g):      # one force group per GridForce -> per-grid energy/forces from the reference itself
        e, f = ref.execute(pos, groups=1 << k)
        ge.append(e)
        gf.append(f)
    grids = {}
    for k, v in enumerate(c["grids"]):
        v = np.asarray(v, dtype=np.float64)
        grids[f"grid{k}"] = v.astype(np.float32) if np.array_equal(v.astype(np.float32).astype(np.float64), v) else v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n_grids=g, counts=np.array(c["counts"]), spacing=np.array(c["spacing"]),
                        origin=np.array(c["origin"]), scaling=np.asarray(c["scaling"], dtype=np.float64), pos=pos,
                        oob_k=np.array(c["oob_k"]), inv_power=np.array(c["inv_power"]), interp=np.array(c.get("interp", 0)),
                        ref_energy=e_all, ref_forces=f_all,
                        ref_grid_energies=np.array(ge), ref_grid_forces=np.array(gf), **grids)
    print(f"{name:22s} atoms={n:5d} grids={g} E={e_all:.12g}")
    ref.close()

if not only or "inv_power_transform" in only:
    rng = np.random.default_rng(17)
    vals = rng.normal(size=(9, 8, 11)) * 10 ** rng.uniform(-3, 4, size=(9, 8, 11))
    vals[::4, ::3, ::5] = 0.0
    out, mode = bindings.ref_inv_power_transform(vals, 4.0)
    np.savez_compressed(os.path.join(HERE, "inv_power_transform.npz"), values=vals, inv_power=4.0, ref_transformed=out,
                        mode_after=mode)
    print(f"inv_power_transform    {vals.size} values, mode after = {mode}")
