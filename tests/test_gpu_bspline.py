"""GPU parity of the cubic B-spline path (GridForce::setInterpolationMethod(1); reference
platforms/reference/src/ReferenceGridForceKernels.cpp:727-795) and of the RUNTIME inv-power transformation
(openmmapi/src/GridForce.cpp:221-272), through the C ABI, against the golden vectors of the reference kernel and the C oracle.

Tolerances as for the trilinear path: MIXED energy 1e-6 / forces 1e-5 (max-norm), DOUBLE 1e-12.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {0: (1e-6, 1e-5), 1: (1e-12, 1e-12)}
BSPLINE = sorted(n for n in cases.CASES if n.startswith("bspline_"))


def _rel_f(f, ref):
    return np.abs(f - ref).max() / max(np.abs(ref).max(), 1e-300)


def _make(gf, dev, c, precision, particles=None, layout=4):
    grids = [gf.Grid(dev, c["counts"], c["spacing"], c["origin"], g, precision, layout=layout) for g in c["grids"]]
    k = gf.Kernel(dev, grids, c["scaling"], particles=particles, inv_power=c["inv_power"], oob_k=c["oob_k"])
    return grids, k


def _close(grids, k):
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
@pytest.mark.parametrize("name", BSPLINE)
def test_bspline_golden_vectors(gpu_device, name, precision):
    import openmmgridforce_b200 as gf
    c, ref = cases.load_golden(name)
    assert c["interp"] == 1
    grids, k = _make(gf, gpu_device, c, precision)
    assert all(g.layout == gf.LAYOUT_BSPLINE for g in grids) and not k.uses_lines_kernel()
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    tol_e, tol_f = TOL[precision]
    for g in range(len(grids)):
        assert abs(ge[0, g] - ref["grid_energies"][g]) <= tol_e * abs(ref["grid_energies"][g]), (name, g)
    assert abs(en[0] - ref["energy"]) <= tol_e * abs(ref["energy"])
    assert _rel_f(forces[0], ref["forces"]) <= tol_f
    _close(grids, k)


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
@pytest.mark.parametrize("name", BSPLINE)
def test_bspline_on_raw_points_golden_vectors(gpu_device, name, precision):
    """GFB_LAYOUT_BSPLINE_POINTS: the same method on the raw points with clamped indexing (general kernel) — the layout the
    platform falls back to when the 32x record copy does not fit. Same tolerances, same golden vectors (incl. the thin
    grid where every index clamps, and the upper faces)."""
    import openmmgridforce_b200 as gf
    c, ref = cases.load_golden(name)
    grids, k = _make(gf, gpu_device, c, precision, layout=gf.LAYOUT_BSPLINE_POINTS)
    points = int(np.prod(c["counts"]))
    assert all(g.device_bytes == points * (4 if precision == 0 else 8) for g in grids) and k.eval_path() == 0
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    tol_e, tol_f = TOL[precision]
    for g in range(len(grids)):
        assert abs(ge[0, g] - ref["grid_energies"][g]) <= tol_e * abs(ref["grid_energies"][g]), (name, g)
    assert abs(en[0] - ref["energy"]) <= tol_e * abs(ref["energy"])
    assert _rel_f(forces[0], ref["forces"]) <= tol_f
    _close(grids, k)


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
def test_bspline_batched_replicas_mixed_geometries(gpu_device, oracle_built, precision):
    """64 replicas x 47 atoms x 3 B-spline grids; the oracle evaluates replica by replica, grid by grid. Odd point counts
    exercise odd and even brick counts on every axis."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    lig, q = W.ligand47()
    rng = np.random.default_rng(4)
    for counts in ((37, 41, 43), (36, 40, 44), (35, 39, 45), (34, 38, 46), (33, 37, 47)):
        sp = (0.06, 0.055, 0.05)
        og = tuple(lig.mean(axis=0) - 0.5 * np.array(sp) * (np.array(counts) - 1))
        grids_v = [(rng.normal(size=counts) * 3).astype(np.float32).astype(np.float64) for _ in range(3)]
        sc = np.stack([q, rng.uniform(0.5, 1.5, 47), rng.uniform(0.5, 1.5, 47)])
        pos = np.stack([lig + rng.uniform(-0.45, 0.45, size=3) for _ in range(64)])
        c = dict(counts=counts, spacing=sp, origin=og, grids=grids_v, scaling=sc, oob_k=[10000.0, 5000.0, 100.0], inv_power=[0.0] * 3)
        port = oracle_built.PortOracle(counts, sp, og, grids_v, sc, oob_k=c["oob_k"], interpolation_method=1)
        ge_ref, f_ref = port.execute_batched(pos, n_threads=4)
        grids, k = _make(gf, gpu_device, c, precision)
        assert k.eval_path() == (3 if precision == 0 else 4)      # the record kernels: gf_eval_bspline_kernel / _f64_kernel
        en, forces, ge = k.execute_host(pos, want_grid_energies=True)
        tol_e, tol_f = TOL[precision]
        assert np.abs(ge - ge_ref).max() <= tol_e * np.abs(ge_ref).max(), counts
        assert np.abs(en - ge_ref.sum(axis=1)).max() <= tol_e * np.abs(ge_ref.sum(axis=1)).max()
        assert _rel_f(forces, f_ref) <= tol_f, counts
        _close(grids, k)


def test_bspline_device_path_fixed_point_and_subset(gpu_device, oracle_built):
    """execute_device with OpenMM's fixed-point force buffer and a particle subset, DOUBLE precision."""
    import torch
    import openmmgridforce_b200 as gf
    c, _ = cases.load_golden("bspline_random_aniso")
    rng = np.random.default_rng(2)
    n_particles = 300
    particles = rng.permutation(n_particles)[:120].astype(np.int32)
    pos = c["pos"][:n_particles]
    sc = c["scaling"][:, :120]
    cc = dict(c, scaling=sc)
    grids, k = _make(gf, gpu_device, cc, 1, particles=particles)
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], sc, oob_k=c["oob_k"], interpolation_method=1)
    f_ref = np.zeros((n_particles, 3))
    e_ref = 0.0
    for g in range(2):
        e, f, _ = port.execute(pos, g, ligand_atoms=particles)      # force written at the ordinal (Q1) -> scatter to particles
        f_ref[particles] += f
        e_ref += e
    tdev = torch.device("cuda:0")
    d_pos = torch.from_numpy(pos.copy()).to(tdev)
    stride = 320
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = torch.zeros(1, dtype=torch.float64, device=tdev)
    stream = torch.cuda.current_stream().cuda_stream
    k.execute_device(1, n_particles, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None, stream)
    torch.cuda.synchronize()
    f = (d_f.view(3, stride)[:, :n_particles].T.double() / 2.0 ** 32).cpu().numpy()
    assert abs(d_e.item() - e_ref) <= 1e-12 * abs(e_ref)
    assert np.abs(f - f_ref).max() <= 2.0 ** -31 + 1e-12 * np.abs(f_ref).max()     # fixed-point quantum
    _close(grids, k)


def test_bspline_memory_footprint(gpu_device):
    import openmmgridforce_b200 as gf
    counts = (50, 60, 72)
    g = gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), 0, layout=gf.LAYOUT_BSPLINE)
    records = (counts[0] + 1) * (counts[1] - 1) * (counts[2] - 1)     # two 4x4 (y,z) windows per padded plane and cell
    assert g.device_bytes == records * 128
    g.close()
    with pytest.raises(gf.GridForceB200Error):
        gf.Kernel(gpu_device, [gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), 0, layout=gf.LAYOUT_BSPLINE),
                               gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), 0, layout=gf.LAYOUT_CELLS)],
                  np.ones((2, 4)))


def test_inv_power_transform_on_device(gpu_device, oracle_built):
    """gfb_inv_power_transform (host and device-pointer forms) against the committed output of the reference's own
    GridForce::applyInvPowerTransformation and against the C restatement."""
    import torch
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inv_power_transform.npz"))
    got = gpu_device.inv_power_transform(z["values"], float(z["inv_power"]))
    ref = z["ref_transformed"]
    assert np.array_equal(got == 0.0, ref == 0.0)
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    assert np.all(np.abs(got - ref) <= 4e-16 * np.abs(ref))                  # <= 2 ulp per value
    rng = np.random.default_rng(8)
    big = rng.normal(size=200_001) * 10 ** rng.uniform(-5, 5, size=200_001)
    big[::1000] = 0.0
    want = oracle_built.port_inv_power_transform(big, 3.5)
    d = torch.from_numpy(big.copy()).to("cuda:0")
    gpu_device.inv_power_transform(None, 3.5, device_ptr=d.data_ptr(), n=d.numel())
    assert np.all(np.abs(d.cpu().numpy() - want) <= 4e-16 * np.abs(want))
    import openmmgridforce_b200 as gf
    with pytest.raises(gf.GridForceB200Error, match="non-zero"):
        gpu_device.inv_power_transform(big, 0.0)
