"""GPU parity of the tricubic Hermite path (GridForce::setInterpolationMethod(2); reference
platforms/reference/src/ReferenceGridForceKernels.cpp:796-893), through the C ABI and the plugin, against the golden
vectors of the reference kernel and the C oracle.

Tolerances: DOUBLE 1e-12. MIXED stores the points in FP32 and computes in FP64, so against the reference's FP64 grids the
north-star bounds apply (energy 1e-6 relative with the FP32-storage floor, forces 1e-5 max-norm), and on grids that are
FP32-representable only the order of the FP64 operations differs (1e-10 asserted).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {0: (1e-6, 1e-5), 1: (1e-12, 1e-12)}
TRICUBIC = sorted(n for n in cases.CASES if n.startswith("tricubic_"))


def _rel_f(f, ref):
    return np.abs(f - ref).max() / max(np.abs(ref).max(), 1e-300)


# (precision, layout): raw points and HERMITE records, both precisions
MODES = [(0, 5), (1, 5), (0, 6), (1, 6)]
MODE_IDS = ["mixed-points", "double-points", "mixed-records", "double-records"]


def _path(gf, precision, layout):
    """gfb_kernel_eval_path of a one-geometry state: the record kernels for HERMITE, the general kernel for POINTS."""
    return 0 if layout != gf.LAYOUT_HERMITE else (5 if precision == 0 else 6)


def _make(gf, dev, c, precision, particles=None, layout=5):
    grids = [gf.Grid(dev, c["counts"], c["spacing"], c["origin"], g, precision, layout=layout) for g in c["grids"]]
    k = gf.Kernel(dev, grids, c["scaling"], particles=particles, inv_power=c["inv_power"], oob_k=c["oob_k"])
    return grids, k


def _close(grids, k):
    k.close()
    for g in grids:
        g.close()


def _storage_floor(c, g):
    """FP32 storage of the points: each stored value is off by <= 6e-8 |v|; the Hermite weights of a stencil sum to at
    most ~4 in absolute value."""
    return 6e-8 * 4.0 * np.abs(c["scaling"][g]).sum() * np.abs(c["grids"][g]).max()


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("name", TRICUBIC)
def test_tricubic_golden_vectors(gpu_device, name, mode):
    import openmmgridforce_b200 as gf
    precision, layout = mode
    c, ref = cases.load_golden(name)
    assert c["interp"] == 2
    grids, k = _make(gf, gpu_device, c, precision, layout=layout)
    assert all(g.layout == layout for g in grids) and k.eval_path() == _path(gf, precision, layout)
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    tol_e, tol_f = TOL[precision]
    floor = 0.0
    for g in range(len(grids)):
        fl = _storage_floor(c, g) if precision == 0 else 0.0
        floor += fl
        assert abs(ge[0, g] - ref["grid_energies"][g]) <= max(tol_e * abs(ref["grid_energies"][g]), fl), (name, g)
    assert abs(en[0] - ref["energy"]) <= max(tol_e * abs(ref["energy"]), floor)
    assert _rel_f(forces[0], ref["forces"]) <= tol_f
    _close(grids, k)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_tricubic_batched_replicas(gpu_device, oracle_built, mode):
    """64 replicas x 47 atoms x 3 tricubic grids on three grid shapes; the oracle evaluates replica by replica, grid by
    grid. Grids FP32-representable: MIXED differs from the oracle by the order of FP64 operations only. HERMITE records
    run the record kernel (gf_eval_bspline_kernel<.., 2>)."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    precision, layout = mode
    lig, q = W.ligand47()
    rng = np.random.default_rng(4)
    for counts in ((37, 41, 43), (36, 40, 44), (20, 21, 22)):
        sp = (0.06, 0.055, 0.05)
        og = tuple(lig.mean(axis=0) - 0.5 * np.array(sp) * (np.array(counts) - 1))
        grids_v = [(rng.normal(size=counts) * 3).astype(np.float32).astype(np.float64) for _ in range(3)]
        sc = np.stack([q, rng.uniform(0.5, 1.5, 47), rng.uniform(0.5, 1.5, 47)])
        pos = np.stack([lig + rng.uniform(-0.45, 0.45, size=3) for _ in range(64)])
        c = dict(counts=counts, spacing=sp, origin=og, grids=grids_v, scaling=sc, oob_k=[10000.0, 5000.0, 100.0], inv_power=[0.0] * 3)
        port = oracle_built.PortOracle(counts, sp, og, grids_v, sc, oob_k=c["oob_k"], interpolation_method=2)
        ge_ref, f_ref = port.execute_batched(pos, n_threads=4)
        grids, k = _make(gf, gpu_device, c, precision, layout=layout)
        assert k.eval_path() == _path(gf, precision, layout)
        en, forces, ge = k.execute_host(pos, want_grid_energies=True)
        tol = 1e-10 if precision == 0 else 1e-12
        assert np.abs(ge - ge_ref).max() <= tol * np.abs(ge_ref).max(), counts
        assert np.abs(en - ge_ref.sum(axis=1)).max() <= tol * np.abs(ge_ref.sum(axis=1)).max()
        assert _rel_f(forces, f_ref) <= tol, counts
        _close(grids, k)


@pytest.mark.parametrize("mode", [(1, 5), (0, 6), (1, 6)], ids=["double-points", "mixed-records", "double-records"])
@pytest.mark.parametrize("counts", [(2, 2, 2), (2, 3, 7), (3, 2, 4), (4, 4, 4), (9, 5, 6)])
def test_tricubic_edges_and_thin_grids(gpu_device, oracle_built, counts, mode):
    """Every cell of small grids — first layers (derivative estimates off), last y/z cells (neighbours by flat index in
    the next row / slab), the last x layer (the reference reads past its vector there; the oracle, the zero guard slab
    of the POINTS layout and the flat-index fill of the HERMITE records all supply 0), atoms on the faces and outside —
    against the oracle: DOUBLE at 1e-12, MIXED records on an FP32-representable grid at 1e-10."""
    import openmmgridforce_b200 as gf
    precision, layout = mode
    rng = np.random.default_rng(sum(counts))
    sp = (0.2, 0.15, 0.1)
    length = np.array(sp) * (np.array(counts) - 1)
    n = 600
    pos = rng.uniform(-0.05, 1.05, size=(n, 3)) * length
    pos[0] = length                                       # the upper corner (quirk Q2: last cell at fraction 1)
    pos[1] = 0.0
    pos[2:12, 0] = length[0]
    pos[12:22, 1] = length[1]
    pos[22:32, 2] = length[2]
    grid = (rng.normal(size=counts) * 4).astype(np.float32).astype(np.float64)
    sc = rng.uniform(0.5, 1.5, size=(1, n))
    c = dict(counts=counts, spacing=sp, origin=(0.0, 0.0, 0.0), grids=[grid], scaling=sc, oob_k=[10000.0], inv_power=[0.0])
    port = oracle_built.PortOracle(counts, sp, (0.0, 0.0, 0.0), [grid], sc, interpolation_method=2)
    e_ref, f_ref, _ = port.execute(pos, 0)
    grids, k = _make(gf, gpu_device, c, precision, layout=layout)
    en, forces, _ = k.execute_host(pos)
    tol = 1e-12 if precision == 1 else 1e-10
    assert abs(en[0] - e_ref) <= tol * abs(e_ref)
    assert _rel_f(forces[0], f_ref) <= tol
    _close(grids, k)


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
def test_tricubic_records_general_kernel_fallback(gpu_device, oracle_built, precision):
    """HERMITE records of two DIFFERENT geometries in one state: the record kernel does not qualify, the general kernel
    reads the records (tricubic_interpolate<float, HERMITE>); and an evaluation order on a one-geometry state."""
    import torch
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(12)
    n = 500
    geoms = [((9, 11, 13), (0.11, 0.09, 0.08)), ((12, 10, 9), (0.08, 0.1, 0.12))]
    grids_v = [(rng.normal(size=cnt) * 3).astype(np.float32).astype(np.float64) for cnt, _ in geoms]
    sc = rng.uniform(0.5, 1.5, size=(2, n))
    pos = rng.uniform(-0.05, 0.95, size=(n, 3))
    grids = [gf.Grid(gpu_device, cnt, sp, (0.0, 0.0, 0.0), v, precision, layout=gf.LAYOUT_HERMITE) for (cnt, sp), v in zip(geoms, grids_v)]
    k = gf.Kernel(gpu_device, grids, sc)
    assert k.eval_path() == 0
    en, forces, ge = k.execute_host(pos, want_grid_energies=True)
    f_ref = np.zeros((n, 3))
    for g, ((cnt, sp), v) in enumerate(zip(geoms, grids_v)):
        port = oracle_built.PortOracle(cnt, sp, (0.0, 0.0, 0.0), [v], sc[g:g + 1], interpolation_method=2)
        e, f, _ = port.execute(pos, 0)
        f_ref += f
        assert abs(ge[0, g] - e) <= 1e-10 * abs(e)
    assert _rel_f(forces[0], f_ref) <= 1e-10
    _close(grids, k)
    # one geometry + an evaluation order -> general kernel as well
    cnt, sp = geoms[0]
    g1 = [gf.Grid(gpu_device, cnt, sp, (0.0, 0.0, 0.0), grids_v[0], precision, layout=gf.LAYOUT_HERMITE)]
    k1 = gf.Kernel(gpu_device, g1, sc[:1])
    port = oracle_built.PortOracle(cnt, sp, (0.0, 0.0, 0.0), [grids_v[0]], sc[:1], interpolation_method=2)
    e_ref, f_ref, _ = port.execute(pos, 0)
    tdev = torch.device("cuda:0")
    d_pos = torch.from_numpy(pos.copy()).to(tdev)
    d_order = torch.from_numpy(rng.permutation(n).astype(np.int32)).to(tdev)
    d_f = torch.zeros(n, 3, dtype=torch.float64, device=tdev)
    d_e = torch.zeros(1, dtype=torch.float64, device=tdev)
    k1.execute_device(1, n, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), gf.FORCE_F64_STORE, 0, d_order.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert abs(d_e.item() - e_ref) <= 1e-10 * abs(e_ref)
    assert _rel_f(d_f.cpu().numpy(), f_ref) <= 1e-10
    _close(g1, k1)


def test_tricubic_device_path_fixed_point_and_subset(gpu_device, oracle_built):
    """execute_device with OpenMM's fixed-point force buffer and a particle subset, DOUBLE precision."""
    import torch
    import openmmgridforce_b200 as gf
    c, _ = cases.load_golden("tricubic_random_aniso")
    rng = np.random.default_rng(2)
    n_particles = 300
    particles = rng.permutation(n_particles)[:120].astype(np.int32)
    pos = c["pos"][:n_particles]
    sc = c["scaling"][:, :120]
    cc = dict(c, scaling=sc)
    grids, k = _make(gf, gpu_device, cc, 1, particles=particles)      # DOUBLE points
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], sc, oob_k=c["oob_k"], interpolation_method=2)
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


def test_tricubic_memory_footprint_and_layout_mix(gpu_device):
    import openmmgridforce_b200 as gf
    counts = (50, 60, 72)
    points = counts[0] * counts[1] * counts[2] + counts[1] * counts[2]        # + the zero guard slab
    for precision, size in ((0, 4), (1, 8)):
        g = gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), precision, layout=gf.LAYOUT_POINTS)
        assert g.device_bytes == points * size
        g.close()
    for precision, size in ((0, 128), (1, 256)):                                                  # the BSPLINE record format
        g = gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), precision, layout=gf.LAYOUT_HERMITE)
        assert g.device_bytes == (counts[0] + 1) * (counts[1] - 1) * (counts[2] - 1) * size
        g.close()
    with pytest.raises(gf.GridForceB200Error):
        gf.Kernel(gpu_device, [gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), 0, layout=gf.LAYOUT_POINTS),
                               gf.Grid(gpu_device, counts, (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(counts), 0, layout=gf.LAYOUT_CELLS)],
                  np.ones((2, 4)))
