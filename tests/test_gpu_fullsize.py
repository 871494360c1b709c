"""Full-size configurations of BASELINE.json on the GPU (-m gpu): C3 (1M atoms x 256^3) and C4 (4096 replicas x 3
grids) compared with the oracle directly (the C restatement finishes these in seconds), C5's shard through
size-independent properties (ones-grid sum, replica permutation invariance, shard-union = whole)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_c3_million_atoms_mixed(gpu_device, oracle_built):
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c3_million_atoms()
    port = oracle_built.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling)
    e_ref, f_ref, cls_ref = port.execute(w.pos[0], 0, classify=True)
    g = gf.Grid(gpu_device, w.counts, w.spacing, w.origin, w.grids[0], gf.PRECISION_MIXED)
    assert g.layout == gf.LAYOUT_CELLS and g.device_bytes == 255 ** 3 * 32            # AUTO: packed cells, 506 MiB
    k = gf.Kernel(gpu_device, [g], w.scaling)
    cls = k.classify_host(w.pos, 0)
    assert np.array_equal(cls["cell"], cls_ref["cell"]) and np.array_equal(cls["inside"], cls_ref["inside"])
    en, f, _ = k.execute_host(w.pos)
    assert abs(en[0] - e_ref) <= 1e-6 * abs(e_ref), (en[0], e_ref)
    assert _rel(f[0], f_ref) <= 1e-5
    k.close()
    g.close()


def test_c3_all_layouts_agree(gpu_device, oracle_built):
    """A 200k-atom slice of C3 through every MIXED layout: same classification, energies/forces within tolerance of the
    oracle, and identical forces between layouts (the layouts differ in where the 8 corners are read from, not in
    arithmetic)."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c3_million_atoms(n_atoms=200_000)
    port = oracle_built.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling)
    e_ref, f_ref, _ = port.execute(w.pos[0], 0)
    forces = {}
    for layout in (gf.LAYOUT_CELLS, gf.LAYOUT_ROWS, gf.LAYOUT_PAIRS):
        g = gf.Grid(gpu_device, w.counts, w.spacing, w.origin, w.grids[0], gf.PRECISION_MIXED, layout=layout)
        k = gf.Kernel(gpu_device, [g], w.scaling)
        en, f, _ = k.execute_host(w.pos)
        assert abs(en[0] - e_ref) <= 1e-6 * abs(e_ref) and _rel(f[0], f_ref) <= 1e-5
        forces[layout] = f
        k.close()
        g.close()
    # same corners, same formulas; only FMA contraction may differ between the template instantiations
    fmax = np.abs(f_ref).max()
    assert np.abs(forces[gf.LAYOUT_CELLS] - forces[gf.LAYOUT_ROWS]).max() <= 2e-6 * fmax
    assert np.abs(forces[gf.LAYOUT_CELLS] - forces[gf.LAYOUT_PAIRS]).max() <= 2e-6 * fmax


def test_c4_batched_mixed_and_double(gpu_device, oracle_built):
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c4_batched_replicas()
    port = oracle_built.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling, oob_k=w.oob_k)
    ge_ref, f_ref = port.execute_batched(w.pos, n_threads=8)
    e_ref = ge_ref.sum(axis=1)
    for precision, te, tf in ((gf.PRECISION_MIXED, 1e-6, 1e-5), (gf.PRECISION_DOUBLE, 1e-12, 1e-12)):
        grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, precision) for v in w.grids]
        k = gf.Kernel(gpu_device, grids, w.scaling, oob_k=w.oob_k)
        en, f, ge = k.execute_host(w.pos, want_grid_energies=True)
        assert np.abs(ge - ge_ref).max() <= te * np.abs(ge_ref).max()
        if precision == gf.PRECISION_MIXED:      # per replica, with the FP32-storage floor
            ok, worst = _per_replica_energy_ok(w, w.pos, en, e_ref)
            assert ok, worst
        else:
            assert (np.abs(en - e_ref) <= te * np.maximum(np.abs(e_ref), np.abs(ge_ref).max(axis=1))).all()
        assert _rel(f, f_ref) <= tf
        k.close()
        for g in grids:
            g.close()


def test_c5_shard_properties(gpu_device):
    """One rank's shard of C5 at 8 GPUs (8192 replicas x 47 x 3 grids of 192^3) on all-ones grids: every replica
    fully inside has E = sum over grids of sum(s) exactly representable to 1e-6; permuting replicas permutes the
    energies; evaluating two half-shards equals evaluating the shard."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    n = 192
    sp = (W.TEST_GRID_SPACING,) * 3
    half = 0.5 * sp[0] * (n - 1)
    pos = W.ligand_replicas(8192, (half, half, half), escape_shift=(0.9, 0.0, 0.0))
    scaling = W._ligand_scaling(3)
    ones = np.ones((n, n, n))
    grid = gf.Grid(gpu_device, (n, n, n), sp, (0, 0, 0), ones, gf.PRECISION_MIXED)
    k = gf.Kernel(gpu_device, [grid, grid, grid], scaling)
    en, f, _ = k.execute_host(pos)
    length = sp[0] * (n - 1)
    inside = ((pos >= 0) & (pos <= length)).all(axis=(1, 2))
    assert 0.9 < inside.mean() < 1.0
    assert np.abs(en[inside] - scaling.sum()).max() <= 1e-6 * abs(scaling.sum())
    assert np.abs(f[inside]).max() == 0.0
    # closed form for every replica: inside atoms contribute s (ones grid), outside atoms the harmonic wall, per grid
    atom_in = ((pos >= 0) & (pos <= length)).all(axis=2)                      # [R, A]
    dev = np.where(pos < 0, pos, np.where(pos > length, pos - length, 0.0))   # [R, A, 3]
    wall = 0.5 * 10000.0 * (dev ** 2).sum(axis=2)                             # per grid
    expect = (atom_in[:, None, :] * scaling[None, :, :]).sum(axis=(1, 2)) + 3 * np.where(atom_in, 0.0, wall).sum(axis=1)
    assert np.abs(en - expect).max() <= 1e-6 * np.abs(expect).max()
    f_expect = -3 * 10000.0 * np.where(atom_in[..., None], 0.0, dev)
    assert np.abs(f - f_expect).max() <= 1e-5 * np.abs(f_expect).max()
    perm = np.random.default_rng(0).permutation(8192)
    en_p, f_p, _ = k.execute_host(np.ascontiguousarray(pos[perm]))
    # energies are sums of per-warp partials added atomically: equal up to FP64 re-association
    assert np.allclose(en_p, en[perm], rtol=1e-12, atol=1e-9) and np.array_equal(f_p, f[perm])
    en_a, _, _ = k.execute_host(np.ascontiguousarray(pos[:4096]))
    en_b, _, _ = k.execute_host(np.ascontiguousarray(pos[4096:]))
    assert np.allclose(np.concatenate([en_a, en_b]), en, rtol=1e-12, atol=1e-9)
    k.close()
    grid.close()


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
def test_single_replica_atom_range_chunks(gpu_device, oracle_built, precision, monkeypatch):
    """One large replica goes through the host pipeline in ATOM ranges (C3's e2e path). Forced to 7 uneven chunks here:
    energies of all ranges accumulate into the one replica entry, per-grid energies too, ADD keeps the caller's forces."""
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(12)
    counts, sp, og = (40, 36, 44), (0.05, 0.06, 0.045), (-0.3, 0.2, 1.0)
    n = 100_003
    grids = [(rng.normal(size=counts) * 4).astype(np.float32).astype(np.float64) for _ in range(2)]
    length = np.array(sp) * (np.array(counts) - 1)
    pos = np.array(og) + rng.uniform(-0.02, 1.02, size=(n, 3)) * length
    sc = rng.uniform(0.5, 1.5, size=(2, n))
    port = oracle_built.PortOracle(counts, sp, og, grids, sc)
    ge_ref, f_ref = port.execute_batched(pos[None], n_threads=4)
    monkeypatch.setenv("GFB_HOST_CHUNKS", "7")
    gs = [gf.Grid(gpu_device, counts, sp, og, g, precision) for g in grids]
    k = gf.Kernel(gpu_device, gs, sc)
    base = rng.normal(size=(1, n, 3))
    forces = base.copy()
    en, _, ge = k.execute_host(pos, forces=forces, force_mode=gf.FORCE_F64_ADD, want_grid_energies=True)
    te, tf = ((1e-6, 1e-5), (1e-12, 1e-12))[precision]
    assert np.abs(ge[0] - ge_ref[0]).max() <= te * np.abs(ge_ref).max()
    assert abs(en[0] - ge_ref.sum()) <= te * abs(ge_ref.sum())
    assert _rel(forces - base, f_ref) <= tf + 1e-15
    k.close()
    for g in gs:
        g.close()


def _per_replica_energy_ok(w, pos, en, e_ref):
    """north_star: energies within 1e-6 relative, PER REPLICA — with the floor the FP32-stored grid values impose
    (6e-8 * sum over atoms and grids of |s| * interpolated |V|, workloads.mixed_energy_bound)."""
    from openmmgridforce_b200 import workloads as W
    bound = np.maximum(1e-6 * np.abs(e_ref), 6e-8 * W.mixed_energy_bound(w, pos))
    err = np.abs(en - e_ref)
    return bool((err <= bound).all()), float((err / np.maximum(bound, 1e-300)).max())


@pytest.mark.parametrize("pdl", [False, True], ids=["serial", "pdl"])
def test_c5_bench_workload_every_replica_vs_oracle(gpu_device, oracle_built, pdl):
    """The workload bench.py times — all 65,536 replicas of configs[4] on the noise grids, device path, OpenMM fixed-point
    forces accumulated over two launches, with and without programmatic dependent launch — compared with the oracle
    replica by replica: energies per replica (1e-6 relative or the FP32-storage floor), forces 1e-5 relative (max-norm)."""
    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c5_sharded_replicas()
    port = oracle_built.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling, oob_k=w.oob_k)
    ge_ref, f_ref = port.execute_batched(w.pos, n_threads=16)
    e_ref = ge_ref.sum(axis=1)
    grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(gpu_device, grids, w.scaling, oob_k=w.oob_k)
    assert k.uses_lines_kernel()
    k.set_launch_overlap(pdl)
    tdev = torch.device("cuda:0")
    r, a = w.n_replicas, w.n_atoms
    n = r * a
    stride = ((n + 31) // 32) * 32
    d_pos = torch.from_numpy(w.pos).to(tdev)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = [torch.zeros(r, dtype=torch.float64, device=tdev) for _ in range(2)]
    d_out = torch.empty(n, 3, dtype=torch.float64, device=tdev)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    for i in range(2):
        k.execute_device(r, a, d_pos.data_ptr(), d_e[i % 2].data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None,
                         stream.cuda_stream, d_energies_clear=d_e[(i + 1) % 2].data_ptr())
    gpu_device.fixed_to_f64(d_f.data_ptr(), stride, n, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    k.set_launch_overlap(False)
    en = d_e[1].cpu().numpy()
    ok, worst = _per_replica_energy_ok(w, w.pos, en, e_ref)
    assert ok, worst
    f = d_out.cpu().numpy().reshape(r, a, 3) / 2.0
    assert np.abs(f - f_ref).max() <= 1e-5 * np.abs(f_ref).max()
    assert not d_e[0].cpu().numpy().any()              # cleared by the second launch for the next step
    # the same batch in DOUBLE precision through the 256-byte record kernel: 1e-12, every replica
    k.close()
    for g in grids:
        g.close()
    if not pdl:
        grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, gf.PRECISION_DOUBLE) for v in w.grids]
        k = gf.Kernel(gpu_device, grids, w.scaling, oob_k=w.oob_k)
        assert k.eval_path() == 2
        en, f, _ = k.execute_host(w.pos)
        scale = np.abs(ge_ref).max()
        assert np.abs(en - e_ref).max() <= 1e-12 * max(scale, np.abs(e_ref).max())
        assert np.abs(f - f_ref).max() <= 1e-12 * np.abs(f_ref).max()
        k.close()
        for g in grids:
            g.close()


def test_mixed_energy_per_replica_bound_on_cancelling_replicas(gpu_device, oracle_built):
    """Replicas built so that their terms nearly cancel (|E| orders of magnitude below the sum of |terms|): the energy
    cannot meet 1e-6 of |E| with FP32-stored grid values, and must meet the storage floor instead; replicas without
    cancellation must meet 1e-6. Grid values are NOT made FP32-representable here: storage rounding is what is measured."""
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    rng = np.random.default_rng(77)
    counts, sp, og = (40, 40, 40), (0.05, 0.05, 0.05), (0.0, 0.0, 0.0)
    grids_v = [rng.normal(size=counts) * 50.0 + 1000.0 for _ in range(2)]       # large offset: big terms
    a = 48
    sc = np.ones((2, a))
    sc[:, a // 2:] = -1.0                                                        # half the atoms cancel the other half
    sc[1] *= 0.5
    r = 600
    pos = rng.uniform(0.1, 1.8, size=(r, a, 3))
    w = W.Workload("cancelling", counts, sp, og, grids_v, sc, pos, [10000.0] * 2, [0.0] * 2)
    port = oracle_built.PortOracle(counts, sp, og, grids_v, sc, oob_k=w.oob_k)
    ge_ref, f_ref = port.execute_batched(pos, n_threads=4)
    e_ref = ge_ref.sum(axis=1)
    gs = [gf.Grid(gpu_device, counts, sp, og, v, gf.PRECISION_MIXED) for v in grids_v]
    k = gf.Kernel(gpu_device, gs, sc, oob_k=w.oob_k)
    en, f, _ = k.execute_host(pos)
    bound_sum = W.mixed_energy_bound(w, pos)
    assert np.median(np.abs(e_ref) / bound_sum) < 1e-2                           # the construction does cancel
    ok, worst = _per_replica_energy_ok(w, pos, en, e_ref)
    assert ok, worst
    assert (np.abs(en - e_ref) > 1e-6 * np.abs(e_ref)).any()                     # ... and 1e-6 of |E| alone would have failed
    assert np.abs(f - f_ref).max() <= 1e-5 * np.abs(f_ref).max()
    k.close()
    for g in gs:
        g.close()
