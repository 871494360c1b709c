"""Definitions of the golden-vector cases. Inputs small enough are stored in the .npz next to the outputs;
make_golden.py (run in the build container, where the reference compiles) evaluates them with the reference's
own kernel (oracle/_ref) and writes tests/golden/<name>.npz."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _ligand():
    doc = json.load(open(os.path.join(HERE, "..", "..", "openmmgridforce_b200", "data", "ligand47.json")))
    return np.array(doc["positions_nm"]), np.array(doc["charges_e"])


def case_ones_grid():
    """C1: python/tests/test_auto_scaling.py:22-25 toy grid (10^3 ones, 0.1 nm): E = sum(s), F = 0."""
    lig, q = _ligand()
    pos = (lig - lig.min(axis=0)) * 0.3 + 0.05
    return dict(counts=(10, 10, 10), spacing=(0.1, 0.1, 0.1), origin=(0.0, 0.0, 0.0), grids=[np.ones((10, 10, 10))],
                scaling=q[None, :], pos=pos, oob_k=[10000.0], inv_power=[0.0])


def case_ramp_grid():
    """python/tests/test_auto_grid.py:58-67 style 5^3 ramp (value 0.5*i over the flat index)."""
    counts = (5, 5, 5)
    vals = (0.5 * np.arange(125, dtype=np.float64)).reshape(counts)
    rng = np.random.default_rng(11)
    pos = rng.uniform(-0.05, 0.45, size=(64, 3))
    return dict(counts=counts, spacing=(0.1, 0.1, 0.1), origin=(0.0, 0.0, 0.0), grids=[vals],
                scaling=rng.uniform(0.5, 1.5, size=(1, 64)), pos=pos, oob_k=[10000.0], inv_power=[0.0])


def case_linear_field():
    """V = 2x + 4y + 6z + 1: interpolation is exact, F = -s*(2,4,6); one atom outside (restraint)."""
    counts, sp = (11, 11, 11), (0.1, 0.1, 0.1)
    i, j, k = np.meshgrid(*(np.arange(n) * 0.1 for n in counts), indexing="ij")
    pos = np.array([[0.33, 0.41, 0.27], [1.5, -0.01, 0.5], [0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    return dict(counts=counts, spacing=sp, origin=(0.0, 0.0, 0.0), grids=[2 * i + 4 * j + 6 * k + 1],
                scaling=np.array([[10.0, 1.0, 2.0, 0.0]]), pos=pos, oob_k=[10000.0], inv_power=[0.0])


def case_random_aniso():
    """Anisotropic spacing, shifted origin, 2 grids with different restraint constants, ~45 % of atoms outside,
    zero scaling factors sprinkled in, atoms on grid nodes and on the lower faces."""
    rng = np.random.default_rng(0)
    counts, sp, og = (17, 13, 19), (0.11, 0.07, 0.13), (0.3, -0.2, 1.0)
    grids = [rng.normal(size=counts) * 5 for _ in range(2)]
    n = 2000
    length = np.array(sp) * (np.array(counts) - 1)
    pos = np.array(og) + rng.uniform(-0.1, 1.1, size=(n, 3)) * length
    # exact grid nodes (fraction 0) and lower-face atoms
    nodes = rng.integers(0, np.array(counts) - 1, size=(50, 3))
    pos[:50] = np.array(og) + nodes * np.array(sp)
    pos[50:60, 0] = og[0]
    pos[60:70, 1] = og[1]
    pos[70:80, 2] = og[2]
    sc = rng.normal(size=(2, n))
    sc[:, ::17] = 0.0
    return dict(counts=counts, spacing=sp, origin=og, grids=grids, scaling=sc, pos=pos, oob_k=[1234.0, 10000.0],
                inv_power=[0.0, 0.0])


def case_ligand_three_grids():
    """The 47-atom ligand in three smooth+noise grids (40^3 @ 0.05 nm) around it — C2's shape, shrunk."""
    lig, q = _ligand()
    counts, sp = (40, 40, 40), (0.05, 0.05, 0.05)
    og = tuple(lig.mean(axis=0) - 0.5 * 0.05 * 39)
    rng = np.random.default_rng(5)
    x = np.arange(40) * 0.05
    grids = []
    for g in range(3):
        smooth = 10.0 * np.sin(7 * x + g)[:, None, None] * np.cos(5 * x)[None, :, None] * np.sin(6 * x - g)[None, None, :]
        grids.append((smooth + rng.uniform(-1, 1, size=counts)).astype(np.float32).astype(np.float64))
    sc = np.stack([q, rng.uniform(0.5, 1.5, 47), rng.uniform(0.5, 1.5, 47)])
    return dict(counts=counts, spacing=sp, origin=og, grids=grids, scaling=sc, pos=lig, oob_k=[10000.0] * 3,
                inv_power=[0.0] * 3)


def case_inv_power():
    """inv_power = 4 on a strictly positive grid (stored as G^(1/4)): pow + chain rule (:1057-1080)."""
    rng = np.random.default_rng(9)
    counts, sp = (12, 12, 12), (0.08, 0.08, 0.08)
    grids = [rng.uniform(0.5, 3.0, size=counts)]
    pos = rng.uniform(0.0, 0.88, size=(300, 3))
    return dict(counts=counts, spacing=sp, origin=(0.0, 0.0, 0.0), grids=grids, scaling=rng.uniform(0.5, 1.5, size=(1, 300)),
                pos=pos, oob_k=[10000.0], inv_power=[4.0])


def case_bspline_random_aniso():
    """Cubic B-spline (interpolation method 1, ReferenceGridForceKernels.cpp:727-795) on the random_aniso inputs plus
    atoms in the first/last cells of every axis (clamped stencil) and on the upper faces (index n-1, fraction 0)."""
    c = case_random_aniso()
    counts, sp, og = np.array(c["counts"]), np.array(c["spacing"]), np.array(c["origin"])
    rng = np.random.default_rng(21)
    pos = c["pos"]
    length = sp * (counts - 1)
    pos[100:130] = og + rng.uniform(0.0, 1.0, size=(30, 3)) * sp                      # first cells
    pos[130:160] = og + length - rng.uniform(0.0, 1.0, size=(30, 3)) * sp             # last cells
    pos[160:165, 0] = (og + length)[0]                                                 # upper faces
    pos[165:170, 1] = (og + length)[1]
    pos[170:175, 2] = (og + length)[2]
    pos[175] = og + length
    c["interp"] = 1
    return c


def case_bspline_ligand_three_grids():
    c = case_ligand_three_grids()
    c["interp"] = 1
    return c


def case_bspline_inv_power():
    c = case_inv_power()
    c["interp"] = 1
    return c


def case_bspline_thin_grid():
    """Smallest legal grids (2 and 3 points on an axis): every stencil index is clamped."""
    rng = np.random.default_rng(33)
    counts, sp = (2, 3, 7), (0.2, 0.15, 0.1)
    length = np.array(sp) * (np.array(counts) - 1)
    pos = rng.uniform(-0.05, 1.05, size=(200, 3)) * length
    return dict(counts=counts, spacing=sp, origin=(0.0, 0.0, 0.0), grids=[rng.normal(size=counts)],
                scaling=rng.uniform(0.5, 1.5, size=(1, 200)), pos=pos, oob_k=[10000.0], inv_power=[0.0], interp=1)


def _tricubic_safe(c, rng):
    """Interpolation method 2 is undefined in the reference in the last x layer (ix == nx-2: the ix+2 neighbour is past the
    end of the value vector, ReferenceGridForceKernels.cpp:825-832) and, like every method, on the upper faces (quirk Q2):
    atoms there are moved to a random interior x. Everything else stays, the last y/z cells included."""
    counts, sp, og = np.array(c["counts"]), np.array(c["spacing"]), np.array(c["origin"])
    pos = c["pos"]
    length = sp * (counts - 1)
    px = pos[:, 0] - og[0]
    bad = (px >= sp[0] * (counts[0] - 2) * (1 - 1e-9)) & (px <= length[0])
    pos[bad, 0] = og[0] + rng.uniform(0.0, 1.0, size=int(bad.sum())) * sp[0] * (counts[0] - 2) * (1 - 1e-9)
    for k in (1, 2):
        on_face = pos[:, k] - og[k] == length[k]
        pos[on_face, k] -= 0.25 * sp[k]
    return c


def case_tricubic_random_aniso():
    """Tricubic Hermite (interpolation method 2, ReferenceGridForceKernels.cpp:796-893) on the random_aniso inputs plus
    atoms in the first cells of every axis (derivative estimates switched off) and in the last y/z cells (neighbour reads
    by flat index land in the next row / x-slab)."""
    c = case_random_aniso()
    counts, sp, og = np.array(c["counts"]), np.array(c["spacing"]), np.array(c["origin"])
    rng = np.random.default_rng(41)
    pos = c["pos"]
    length = sp * (counts - 1)
    pos[100:130] = og + rng.uniform(0.0, 1.0, size=(30, 3)) * sp                                   # first cells
    pos[130:160, 1:] = (og + length - rng.uniform(1e-6, 1.0, size=(30, 3)) * sp)[:, 1:]            # last y and z cells
    pos[160:175, 1] = (og + length - rng.uniform(1e-6, 1.0, size=(15, 3)) * sp)[:, 1]              # last y cells only
    pos[175:190, 2] = (og + length - rng.uniform(1e-6, 1.0, size=(15, 3)) * sp)[:, 2]              # last z cells only
    c["interp"] = 2
    return _tricubic_safe(c, rng)


def case_tricubic_ligand_three_grids():
    c = case_ligand_three_grids()
    c["interp"] = 2
    return _tricubic_safe(c, np.random.default_rng(42))


def case_tricubic_inv_power():
    c = case_inv_power()
    c["interp"] = 2
    return _tricubic_safe(c, np.random.default_rng(43))


CASES = {
    "ones_grid": case_ones_grid,
    "ramp_grid": case_ramp_grid,
    "linear_field": case_linear_field,
    "random_aniso": case_random_aniso,
    "ligand_three_grids": case_ligand_three_grids,
    "inv_power": case_inv_power,
    "bspline_random_aniso": case_bspline_random_aniso,
    "bspline_ligand_three_grids": case_bspline_ligand_three_grids,
    "bspline_inv_power": case_bspline_inv_power,
    "bspline_thin_grid": case_bspline_thin_grid,
    "tricubic_random_aniso": case_tricubic_random_aniso,
    "tricubic_ligand_three_grids": case_tricubic_ligand_three_grids,
    "tricubic_inv_power": case_tricubic_inv_power,
}


def load_golden(name):
    """Returns (inputs dict, outputs dict) from the committed .npz (inputs are stored, not regenerated)."""
    z = np.load(os.path.join(HERE, name + ".npz"))
    n_grids = int(z["n_grids"])
    inp = dict(counts=tuple(int(c) for c in z["counts"]), spacing=tuple(z["spacing"]), origin=tuple(z["origin"]),
               grids=[z[f"grid{g}"].astype(np.float64) for g in range(n_grids)], scaling=z["scaling"], pos=z["pos"],
               oob_k=list(z["oob_k"]), inv_power=list(z["inv_power"]), interp=int(z["interp"]) if "interp" in z else 0)
    out = dict(grid_energies=z["ref_grid_energies"], energy=float(z["ref_energy"]), forces=z["ref_forces"],
               grid_forces=z["ref_grid_forces"])
    return inp, out
