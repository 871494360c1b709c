"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerances (BASELINE.json north_star):
  classification (inside flag, cell index)   bit-exact
  MIXED   energy 1e-6 relative, forces 1e-5 relative in max-norm
  DOUBLE  energy and forces 1e-12 relative
"Relative" is to |E_ref| for energies and to max|F_ref| for forces.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {0: (1e-6, 1e-5), 1: (1e-12, 1e-12)}     # precision -> (energy, force)
GOLDEN = sorted(n for n in cases.CASES if not n.startswith(("bspline_", "tricubic_")))      # trilinear cases; B-spline: test_gpu_bspline.py, tricubic: test_gpu_tricubic.py
# (precision, layout): every device layout of include/gridforce_b200.h's gfb_layout, in both arithmetic modes
MODES = [(0, 1), (0, 2), (0, 3), (1, 1), (1, 2)]
MODE_IDS = ["mixed-cells", "mixed-rows", "mixed-pairs", "double-cells", "double-rows"]


def _rel_e(e, ref):
    return abs(e - ref) / max(abs(ref), 1e-300)


def _rel_f(f, ref):
    return np.abs(f - ref).max() / max(np.abs(ref).max(), 1e-300)


def _make(gf, dev, c, precision, particles=None, layout=None):
    grids = [gf.Grid(dev, c["counts"], c["spacing"], c["origin"], g, precision, layout=layout) for g in c["grids"]]
    k = gf.Kernel(dev, grids, c["scaling"], particles=particles, inv_power=c["inv_power"], oob_k=c["oob_k"])
    return grids, k


def _close(grids, k):
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("name", GOLDEN)
def test_golden_vectors(gpu_device, name, mode):
    import openmmgridforce_b200 as gf
    precision, layout = mode
    c, ref = cases.load_golden(name)
    grids, k = _make(gf, gpu_device, c, precision, layout=layout)
    assert all(g.layout == layout for g in grids)
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    tol_e, tol_f = TOL[precision]
    for g in range(len(grids)):
        assert _rel_e(ge[0, g], ref["grid_energies"][g]) <= tol_e, (name, g, ge[0, g], ref["grid_energies"][g])
    assert _rel_e(en[0], ref["energy"]) <= tol_e
    assert abs(en[0] - ge[0].sum()) <= 1e-12 * max(1.0, np.abs(ge[0]).sum())
    assert _rel_f(forces[0], ref["forces"]) <= tol_f
    _close(grids, k)


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", GOLDEN)
def test_classification_bit_exact(gpu_device, oracle_built, name, precision):
    import openmmgridforce_b200 as gf
    c, _ = cases.load_golden(name)
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], oob_k=c["oob_k"],
                                   inv_power=c["inv_power"])
    grids, k = _make(gf, gpu_device, c, precision)
    for g in range(len(grids)):
        _, _, want = port.execute(c["pos"], g, classify=True)
        got = k.classify_host(c["pos"], g)
        assert np.array_equal(got["inside"], want["inside"])
        assert np.array_equal(got["cell"], want["cell"])
    _close(grids, k)


def test_classification_adversarial_quotients(gpu_device, oracle_built):
    """Positions whose quotient pi/spacing sits within a few ulps of an integer: the reciprocal-multiply fast
    path must hand these to the exact division so that (int)(pi/spacing) is the reference's."""
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(7)
    counts = (300, 41, 57)
    sp = (0.0125, 0.1 / 3, 0.07)
    og = (1.00175115, -0.3, 0.0)
    grid = rng.normal(size=counts)
    n = 60000
    node = np.stack([rng.integers(1, c - 1, size=n) for c in counts], 1).astype(np.float64)
    pos = np.array(og) + node * np.array(sp)
    for _ in range(3):          # walk a few ulps either side of the node
        step = rng.integers(-3, 4, size=pos.shape)
        pos = np.where(step > 0, np.nextafter(pos, np.inf), np.where(step < 0, np.nextafter(pos, -np.inf), pos))
    sc = np.ones((1, n))
    port = oracle_built.PortOracle(counts, sp, og, [grid], sc)
    _, _, want = port.execute(pos, 0, classify=True)
    for precision in (0, 1):
        g = gf.Grid(gpu_device, counts, sp, og, grid, precision)
        k = gf.Kernel(gpu_device, [g], sc)
        got = k.classify_host(pos, 0)
        assert np.array_equal(got["inside"], want["inside"])
        assert np.array_equal(got["cell"], want["cell"]), int((got["cell"] != want["cell"]).any(axis=1).sum())
        _close([g], k)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_batched_replicas_vs_oracle(gpu_device, oracle_built, mode):
    """C4's shape, shrunk: 96 replicas x 47 atoms x 3 grids, some replicas partly outside the grid."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    precision, layout = mode
    w = W.c4_batched_replicas(n_replicas=96, counts=(96, 120, 104))
    # the ligand sits near the middle of the full test grid; move the small grid under it
    lig, _ = W.ligand47()
    og = tuple(lig.mean(axis=0) - 0.5 * np.array(w.spacing) * (np.array(w.counts) - 1))
    c = dict(counts=w.counts, spacing=w.spacing, origin=og, grids=w.grids, scaling=w.scaling, oob_k=w.oob_k, inv_power=w.inv_power)
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], oob_k=c["oob_k"])
    ge_ref, f_ref = port.execute_batched(w.pos, n_threads=4)
    grids, k = _make(gf, gpu_device, c, precision, layout=layout)
    en, forces, ge = k.execute_host(w.pos, want_grid_energies=True)
    tol_e, tol_f = TOL[precision]
    outside = int((np.abs(f_ref).max(axis=(1, 2)) > 1e3).sum())
    assert outside >= 1, "workload should exercise the restraint branch"
    # per replica (north_star: 1e-6 relative); MIXED adds the floor the FP32-stored grid values impose
    e_ref = ge_ref.sum(axis=1)
    wl = W.Workload("c4 small", c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], w.pos, c["oob_k"], c["inv_power"])
    floor = 6e-8 * W.mixed_energy_bound(wl, w.pos) if precision == gf.PRECISION_MIXED else 0.0
    term = np.abs(ge_ref).max(axis=1)
    assert (np.abs(ge - ge_ref) <= np.maximum(tol_e * np.maximum(np.abs(ge_ref), term[:, None] * (precision == gf.PRECISION_DOUBLE)),
                                              np.asarray(floor).reshape(-1, 1) if precision == gf.PRECISION_MIXED else 0.0)).all()
    assert (np.abs(en - e_ref) <= np.maximum(tol_e * np.maximum(np.abs(e_ref), term * (precision == gf.PRECISION_DOUBLE)), floor)).all()
    assert _rel_f(forces, f_ref) <= tol_f
    _close(grids, k)


def test_particle_subset_and_force_modes(gpu_device, oracle_built):
    """setLigandAtoms-style subset: 20 of 60 particles evaluated, forces written at the PARTICLE index; STORE leaves
    other entries untouched, ADD accumulates onto what is there."""
    import openmmgridforce_b200 as gf
    c = cases.case_random_aniso()
    rng = np.random.default_rng(1)
    particles = rng.permutation(60)[:20].astype(np.int32)
    pos = np.ascontiguousarray(np.stack([c["pos"][:60], c["pos"][60:120], c["pos"][120:180]]))
    sc = c["scaling"][:, :20]
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], sc, oob_k=c["oob_k"])
    grids = [gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], g, 1) for g in c["grids"]]
    k = gf.Kernel(gpu_device, grids, sc, particles=particles, oob_k=c["oob_k"])
    base = rng.normal(size=pos.shape)
    f_store = base.copy()
    en, _, _ = k.execute_host(pos, forces=f_store, force_mode=gf.FORCE_F64_STORE)
    f_add = base.copy()
    k.execute_host(pos, forces=f_add, force_mode=gf.FORCE_F64_ADD)
    for r in range(3):
        want_f = np.zeros((20, 3))
        want_e = 0.0
        for g in range(2):
            e, f, _ = port.execute(pos[r][particles], g)
            want_e += e
            want_f += f
        assert _rel_e(en[r], want_e) <= 1e-12
        touched = np.zeros(60, dtype=bool)
        touched[particles] = True
        assert np.array_equal(f_store[r][~touched], base[r][~touched])
        assert _rel_f(f_store[r][particles], want_f) <= 1e-12
        assert np.allclose(f_add[r][particles], base[r][particles] + want_f, rtol=1e-12, atol=1e-9)
        assert np.array_equal(f_add[r][~touched], base[r][~touched])
    _close(grids, k)


def test_device_path_fixed_point_and_sort(gpu_device, oracle_built):
    """CUDA-platform style call: device-resident positions, OpenMM 64-bit fixed-point planar force buffer
    (accumulated with RED.ADD.64), energy accumulated into a device double; then the same with a Morton order."""
    import torch
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(2)
    counts, sp = (64, 64, 64), (0.0125,) * 3
    grid = rng.normal(size=counts) * 3
    n = 50_000
    length = sp[0] * 63
    pos = rng.uniform(-0.02 * length, 1.02 * length, size=(n, 3))
    sc = rng.uniform(0.5, 1.5, size=(1, n))
    port = oracle_built.PortOracle(counts, sp, (0, 0, 0), [grid], sc)
    e_ref, f_ref, _ = port.execute(pos, 0)
    g = gf.Grid(gpu_device, counts, sp, (0, 0, 0), grid, 0)
    k = gf.Kernel(gpu_device, [g], sc)
    dev = torch.device("cuda:0")
    d_pos = torch.from_numpy(pos).to(dev)
    stride = ((n + 31) // 32) * 32
    side = torch.cuda.Stream()            # a real (non-NULL) stream handle: NULL means the library's own stream
    stream = side.cuda_stream
    torch.cuda.synchronize()
    for use_order in (False, True):
        d_f = torch.zeros(3 * stride, dtype=torch.int64, device=dev)
        d_e = torch.zeros(1, dtype=torch.float64, device=dev)
        d_out = torch.empty(n, 3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()            # allocations/zero-fills ran on torch's stream; ours is `side`
        d_order = None
        if use_order:
            d_order = torch.empty(n, dtype=torch.int32, device=dev)
            k.sort_atoms(1, n, d_pos.data_ptr(), d_order.data_ptr(), stream)
            torch.cuda.synchronize()
            order = d_order.cpu().numpy()
            assert np.array_equal(np.sort(order), np.arange(n)), "order must be a permutation"
        for _ in range(2):      # two launches: the buffers accumulate
            k.execute_device(1, n, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride,
                             d_order.data_ptr() if use_order else None, stream)
        gpu_device.fixed_to_f64(d_f.data_ptr(), stride, n, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
        assert _rel_e(d_e.item(), 2 * e_ref) <= 1e-6
        assert _rel_f(d_out.cpu().numpy(), 2 * f_ref) <= 1e-5
    _close([g], k)


def test_upper_face_and_degenerate_inputs(gpu_device, oracle_built):
    """Atom exactly on the upper face: the reference reads past the grid (UB); the CUDA path and the restatement
    both evaluate the last cell at fraction 1. Also: NaN position (falls to the restraint branch, adds 0 like the
    reference's comparisons do), zero replicas, zero atoms."""
    import openmmgridforce_b200 as gf
    counts, sp = (6, 5, 4), (0.1, 0.2, 0.3)
    rng = np.random.default_rng(4)
    grid = rng.normal(size=counts)
    h = np.array(sp) * (np.array(counts) - 1)
    pos = np.array([h, [h[0], 0.1, 0.1], [0.05, h[1], 0.2], [0.0, 0.0, h[2]], [np.nan, 0.1, 0.1]])
    sc = np.ones((1, 5))
    port = oracle_built.PortOracle(counts, sp, (0, 0, 0), [grid], sc)
    e_ref, f_ref, cls_ref = port.execute(pos, 0, classify=True)
    for precision, layout in MODES:
        g = gf.Grid(gpu_device, counts, sp, (0, 0, 0), grid, precision, layout=layout)
        k = gf.Kernel(gpu_device, [g], sc)
        cls = k.classify_host(pos, 0)
        assert np.array_equal(cls["cell"], cls_ref["cell"]) and np.array_equal(cls["inside"], cls_ref["inside"])
        assert tuple(cls["cell"][0]) == (4, 3, 2)
        en, f, _ = k.execute_host(pos)
        assert _rel_e(en[0], e_ref) <= TOL[precision][0] * 10
        assert np.abs(f[0] - f_ref).max() <= TOL[precision][1] * 10 * np.abs(f_ref).max()
        en0, f0, _ = k.execute_host(np.zeros((0, 5, 3)))
        assert en0.shape == (0,)
        _close([g], k)
    g = gf.Grid(gpu_device, counts, sp, (0, 0, 0), grid, 0)
    k0 = gf.Kernel(gpu_device, [g], np.zeros((1, 0)))
    en, f, _ = k0.execute_host(np.zeros((2, 3, 3)))
    assert np.array_equal(en, [0.0, 0.0]) and not f.any()
    _close([g], k0)


def test_mixed_geometry_grids(gpu_device, oracle_built):
    """Grids of different counts/spacing/origin in one kernel (the non-SAME code path): each classifies separately."""
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(8)
    specs = [((9, 11, 13), (0.1, 0.09, 0.08), (0.0, 0.0, 0.0)), ((14, 8, 10), (0.07, 0.12, 0.1), (-0.1, 0.05, 0.0))]
    vals = [rng.normal(size=s[0]) for s in specs]
    n = 500
    pos = rng.uniform(-0.15, 1.1, size=(n, 3))
    sc = rng.normal(size=(2, n))
    want_e, want_f = 0.0, np.zeros((n, 3))
    for g, (cn, sp, og) in enumerate(specs):
        port = oracle_built.PortOracle(cn, sp, og, [vals[g]], sc[g:g + 1])
        e, f, _ = port.execute(pos, 0)
        want_e += e
        want_f += f
    for layout in (1, 2):
        grids = [gf.Grid(gpu_device, cn, sp, og, vals[g], 1, layout=layout) for g, (cn, sp, og) in enumerate(specs)]
        k = gf.Kernel(gpu_device, grids, sc)
        en, f, _ = k.execute_host(pos)
        assert _rel_e(en[0], want_e) <= 1e-12 and _rel_f(f[0], want_f) <= 1e-12
        _close(grids, k)


def test_argument_errors_raise(gpu_device):
    import openmmgridforce_b200 as gf
    with pytest.raises(gf.GridForceB200Error):      # PAIRS is a MIXED-only layout
        gf.Grid(gpu_device, (4, 4, 4), (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(64), gf.PRECISION_DOUBLE, layout=gf.LAYOUT_PAIRS)
    with pytest.raises(gf.GridForceB200Error):
        gf.Grid(gpu_device, (1, 4, 4), (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(16))
    with pytest.raises(gf.GridForceB200Error):
        gf.Grid(gpu_device, (4, 4, 4), (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(10))
    g = gf.Grid(gpu_device, (4, 4, 4), (0.1, 0.1, 0.1), (0, 0, 0), np.zeros(64))
    with pytest.raises(gf.GridForceB200Error):
        gf.Kernel(gpu_device, [g] * 9, np.zeros((9, 3)))
    k = gf.Kernel(gpu_device, [g], np.ones((1, 3)), particles=[0, 5, 2])
    with pytest.raises(gf.GridForceB200Error):
        k.execute_host(np.zeros((1, 4, 3)))       # particle index 5 needs n_particles >= 6
    _close([g], k)


def test_update_parameters(gpu_device, oracle_built):
    """copyParametersToContext: new scaling factors take effect (linearity: 2x scaling -> 2x E, F)."""
    import openmmgridforce_b200 as gf
    c, ref = cases.load_golden("ramp_grid")
    grids, k = _make(gf, gpu_device, c, 1)
    k.update_parameters(scaling=2.0 * c["scaling"])
    en, f, _ = k.execute_host(c["pos"])
    # the restraint part is unscaled; compare against the port with doubled scaling
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], 2.0 * c["scaling"], oob_k=c["oob_k"])
    e2, f2, _ = port.execute(c["pos"], 0)
    assert _rel_e(en[0], e2) <= 1e-12 and _rel_f(f[0], f2) <= 1e-12
    _close(grids, k)


def test_energy_slots_particle_groups(gpu_device, oracle_built):
    """gfb_kernel_set_energy_slots: three particle groups flattened into one atom list, two replicas; energies come
    back per (replica, group) and equal the oracle evaluated group by group; forces accumulate per particle, including
    a particle that belongs to two groups."""
    import openmmgridforce_b200 as gf
    c = cases.case_random_aniso()
    rng = np.random.default_rng(12)
    groups = [list(range(0, 40)), list(range(40, 75)), [5, 80, 81, 82]]          # particle 5 is in groups 0 and 2
    particles = np.array(sum(groups, []), dtype=np.int32)
    slots = np.concatenate([np.full(len(g), i) for i, g in enumerate(groups)]).astype(np.int32)
    scaling = rng.normal(size=(1, particles.size))
    pos = np.ascontiguousarray(np.stack([c["pos"][:100], c["pos"][100:200]]))
    g = gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], c["grids"][0], gf.PRECISION_DOUBLE)
    k = gf.Kernel(gpu_device, [g], scaling, particles=particles, oob_k=c["oob_k"][:1])
    k.set_energy_slots(slots, 3)
    forces = np.zeros_like(pos)
    en, _, ge = k.execute_host(pos, forces=forces, force_mode=gf.FORCE_F64_ADD, want_grid_energies=True)
    en = en.reshape(2, 3)
    for r in range(2):
        want_f = np.zeros((100, 3))
        lo = 0
        for s, grp in enumerate(groups):
            port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"][:1], scaling[:, lo:lo + len(grp)],
                                           oob_k=c["oob_k"][:1])
            e, f, _ = port.execute(pos[r][grp], 0)
            np.add.at(want_f, grp, f)
            assert _rel_e(en[r, s], e) <= 1e-12, (r, s)
            lo += len(grp)
        assert _rel_f(forces[r], want_f) <= 1e-12
    assert np.allclose(ge.reshape(2, 3), en, rtol=1e-13)
    _close([g], k)
