"""GPU parity tests (-m gpu) of gf_eval_lines_kernel — the kernel that serves MIXED packed cells of one geometry
(openmmgridforce_b200/csrc/gf_eval_lines.cuh): 1 grid read directly, 2-4 grids through 128-byte records, cp.async +
swizzled shared memory. Every case is compared with the C oracle (oracle/gridforce_oracle.c, bit-exact with the
reference's ReferenceCalcGridForceKernel) on the same inputs.

Tolerances (BASELINE.json north_star): classification bit-exact; energies 1e-6 relative; forces 1e-5 relative, max-norm.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

pytestmark = pytest.mark.gpu

TOL_E, TOL_F = 1e-6, 1e-5


def _case(n_grids, n_replicas, n_atoms, seed, counts=(23, 19, 31), frac_outside=0.08):
    """Random anisotropic grids (FP32-representable values, so the test measures arithmetic and not storage rounding),
    replicas of n_atoms atoms scattered over the box with a few outside, zero scaling factors sprinkled in."""
    rng = np.random.default_rng(seed)
    sp, og = (0.05, 0.07, 0.04), (0.3, -0.2, 1.0)
    grids = [(rng.normal(size=counts) * 4).astype(np.float32).astype(np.float64) for _ in range(n_grids)]
    length = np.array(sp) * (np.array(counts) - 1)
    pos = np.array(og) + rng.uniform(0.0, 1.0, size=(n_replicas, n_atoms, 3)) * length
    out = rng.uniform(size=(n_replicas, n_atoms)) < frac_outside
    pos[out] += rng.choice([-1.0, 1.0], size=(int(out.sum()), 3)) * rng.uniform(0.0, 0.3, size=(int(out.sum()), 3)) * length
    sc = rng.normal(size=(n_grids, n_atoms))
    sc[:, ::7] = 0.0
    if n_grids > 1 and n_atoms > 3:
        sc[1, 3] = 0.0                    # zero on one grid only: that grid skips, the others interpolate
    oob = [10000.0, 1234.0, 777.0, 5000.0][:n_grids]
    return dict(counts=counts, spacing=sp, origin=og, grids=grids, scaling=sc, pos=pos, oob_k=oob)


def _oracle(bindings, c):
    port = bindings.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], oob_k=c["oob_k"])
    return port.execute_batched(c["pos"], n_threads=4)           # ([R][G] energies, [R][A][3] forces)


def _make(gf, dev, c, particles=None):
    grids = [gf.Grid(dev, c["counts"], c["spacing"], c["origin"], g, gf.PRECISION_MIXED, layout=gf.LAYOUT_CELLS) for g in c["grids"]]
    k = gf.Kernel(dev, grids, c["scaling"], particles=particles, oob_k=c["oob_k"])
    return grids, k


def _close(grids, k):
    k.close()
    for g in grids:
        g.close()


def _check(en, ge, forces, ge_ref, f_ref):
    """Per replica: |E - E_ref| <= 1e-6 * max(|E_ref|, largest per-grid term of that replica). The grids of these cases are
    FP32-representable, so no storage-rounding floor is involved (tests/test_gpu_fullsize.py measures that one)."""
    term = np.abs(ge_ref).max(axis=1)
    assert (np.abs(ge - ge_ref) <= TOL_E * np.maximum(np.abs(ge_ref), term[:, None])).all()
    e_ref = ge_ref.sum(axis=1)
    assert (np.abs(en - e_ref) <= TOL_E * np.maximum(np.abs(e_ref), term)).all()
    assert np.abs(forces - f_ref).max() <= TOL_F * np.abs(f_ref).max()


# ragged totals on purpose: 1x1, fewer atoms than a warp, totals that are not multiples of 32 / 128 / 256
SHAPES = [(1, 1), (1, 5), (1, 300), (3, 47), (37, 47), (130, 9), (1, 4099)]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{r}x{a}" for r, a in SHAPES])
@pytest.mark.parametrize("n_grids", [1, 2, 3, 4])
def test_lines_kernel_vs_oracle(gpu_device, oracle_built, n_grids, shape):
    import openmmgridforce_b200 as gf
    r, a = shape
    c = _case(n_grids, r, a, seed=100 * n_grids + r + a)
    ge_ref, f_ref = _oracle(oracle_built, c)
    grids, k = _make(gf, gpu_device, c)
    assert k.uses_lines_kernel()
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    _check(en, ge, forces, ge_ref, f_ref)
    # without per-grid energies (the other template instantiation), accumulating onto existing forces
    f0 = np.random.default_rng(1).normal(size=c["pos"].shape)
    facc = f0.copy()
    en2, _, _ = k.execute_host(c["pos"], forces=facc, force_mode=gf.FORCE_F64_ADD)
    assert np.abs(en2 - ge_ref.sum(axis=1)).max() <= TOL_E * max(np.abs(ge_ref.sum(axis=1)).max(), np.abs(ge_ref).max())
    assert np.abs(facc - f0 - f_ref).max() <= TOL_F * np.abs(f_ref).max()
    # classification through the same device function the kernel uses: bit-exact
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"], oob_k=c["oob_k"])
    for g in range(n_grids):
        got = k.classify_host(c["pos"], g)
        for rep in range(min(r, 3)):
            _, _, want = port.execute(c["pos"][rep], g, classify=True)
            sl = slice(rep * a, (rep + 1) * a)
            assert np.array_equal(got["inside"][sl], want["inside"]) and np.array_equal(got["cell"][sl], want["cell"])
    _close(grids, k)


@pytest.mark.parametrize("n_grids", [1, 3])
def test_lines_kernel_device_fixed_point(gpu_device, oracle_built, n_grids):
    """CUDA-platform style: device positions, OpenMM fixed-point planar forces accumulated over two launches (with the L2
    prefetch of the force lines for one grid), energies accumulated, next step's accumulator cleared by the launch."""
    import torch
    import openmmgridforce_b200 as gf
    r, a = 41, 47
    c = _case(n_grids, r, a, seed=7)
    ge_ref, f_ref = _oracle(oracle_built, c)
    grids, k = _make(gf, gpu_device, c)
    dev = torch.device("cuda:0")
    n = r * a
    stride = ((n + 31) // 32) * 32
    d_pos = torch.from_numpy(c["pos"]).to(dev)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=dev)
    d_e = torch.zeros(r, dtype=torch.float64, device=dev)
    d_next = torch.full((r,), 123.0, dtype=torch.float64, device=dev)
    d_out = torch.empty(n, 3, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(2):
        k.execute_device(r, a, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None,
                         side.cuda_stream, d_energies_clear=d_next.data_ptr())
    gpu_device.fixed_to_f64(d_f.data_ptr(), stride, n, d_out.data_ptr(), side.cuda_stream)
    torch.cuda.synchronize()
    e_ref = ge_ref.sum(axis=1)
    assert np.abs(d_e.cpu().numpy() - 2 * e_ref).max() <= TOL_E * 2 * max(np.abs(e_ref).max(), np.abs(ge_ref).max())
    assert np.abs(d_out.cpu().numpy().reshape(r, a, 3) - 2 * f_ref).max() <= TOL_F * 2 * np.abs(f_ref).max()
    assert not d_next.cpu().numpy().any()
    _close(grids, k)


def test_lines_kernel_particle_subset_and_slots(gpu_device, oracle_built):
    """setParticles-style subset (the non-contiguous position path) plus energy slots whose atoms INTERLEAVE
    (slot pattern 0,1,2,0,1,2,...): every run of equal keys inside a warp has length 1, and equal keys that are not
    adjacent must not be merged twice."""
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(3)
    r, n_particles, a = 5, 90, 60
    c = _case(3, r, n_particles, seed=11)
    particles = np.sort(rng.choice(n_particles, size=a, replace=False)).astype(np.int32)
    sc = rng.normal(size=(3, a))
    slots = (np.arange(a) % 3).astype(np.int32)
    grids = [gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], g, gf.PRECISION_MIXED) for g in c["grids"]]
    k = gf.Kernel(gpu_device, grids, sc, particles=particles, oob_k=c["oob_k"])
    k.set_energy_slots(slots, 3)
    assert k.uses_lines_kernel()
    forces = np.zeros_like(c["pos"])
    en, _, ge = k.execute_host(c["pos"], forces=forces, force_mode=gf.FORCE_F64_ADD, want_grid_energies=True)
    en = en.reshape(r, 3)
    ge = ge.reshape(r, 3, 3)
    want_f = np.zeros_like(forces)
    for s in range(3):
        sel = np.nonzero(slots == s)[0]
        port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], sc[:, sel], oob_k=c["oob_k"])
        ge_ref, f_ref = port.execute_batched(np.ascontiguousarray(c["pos"][:, particles[sel]]), n_threads=2)
        scale = np.abs(ge_ref).max()
        assert np.abs(ge[:, s, :] - ge_ref).max() <= TOL_E * scale, s
        assert np.abs(en[:, s] - ge_ref.sum(axis=1)).max() <= TOL_E * scale * 3, s
        want_f[:, particles[sel]] += f_ref
    assert np.abs(forces - want_f).max() <= TOL_F * np.abs(want_f).max()
    untouched = np.setdiff1d(np.arange(n_particles), particles)
    assert not forces[:, untouched].any()
    _close(grids, k)


def test_general_kernel_interleaved_slots_double(gpu_device, oracle_built):
    """The same interleaved-slot pattern through the general kernel (DOUBLE precision): 1e-12."""
    import openmmgridforce_b200 as gf
    r, a = 2, 70
    c = _case(1, r, a, seed=21)
    slots = (np.arange(a) % 2).astype(np.int32)
    g = gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], c["grids"][0], gf.PRECISION_DOUBLE)
    k = gf.Kernel(gpu_device, [g], c["scaling"], oob_k=c["oob_k"])
    k.set_energy_slots(slots, 2)
    assert not k.uses_lines_kernel()
    en, _, _ = k.execute_host(c["pos"])
    en = en.reshape(r, 2)
    for s in range(2):
        sel = np.nonzero(slots == s)[0]
        port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], c["scaling"][:, sel], oob_k=c["oob_k"])
        ge_ref, _ = port.execute_batched(np.ascontiguousarray(c["pos"][:, sel]), n_threads=2)
        assert np.abs(en[:, s] - ge_ref[:, 0]).max() <= 1e-12 * np.abs(ge_ref).max()
    _close([g], k)


def test_lines_kernel_adversarial_quotients_three_grids(gpu_device, oracle_built):
    """Positions a few ulps either side of grid nodes, evaluated (not only classified) by the 3-grid lines kernel: a wrong
    cell would show up as a wrong energy only when the field is discontinuous, so the check is on the classification the
    kernel's own device function reports, plus energies/forces against the oracle."""
    import openmmgridforce_b200 as gf
    rng = np.random.default_rng(17)
    counts, sp, og = (120, 41, 57), (0.0125, 0.1 / 3, 0.07), (1.00175115, -0.3, 0.0)
    grids = [rng.normal(size=counts).astype(np.float32).astype(np.float64) for _ in range(3)]
    n = 20000
    node = np.stack([rng.integers(1, cc - 1, size=n) for cc in counts], 1).astype(np.float64)
    pos = np.array(og) + node * np.array(sp)
    for _ in range(3):
        step = rng.integers(-3, 4, size=pos.shape)
        pos = np.where(step > 0, np.nextafter(pos, np.inf), np.where(step < 0, np.nextafter(pos, -np.inf), pos))
    sc = rng.uniform(0.5, 1.5, size=(3, n))
    port = oracle_built.PortOracle(counts, sp, og, grids, sc)
    g3 = [gf.Grid(gpu_device, counts, sp, og, g, gf.PRECISION_MIXED) for g in grids]
    k = gf.Kernel(gpu_device, g3, sc)
    assert k.uses_lines_kernel()
    _, _, want = port.execute(pos, 0, classify=True)
    got = k.classify_host(pos, 0)
    assert np.array_equal(got["inside"], want["inside"])
    assert np.array_equal(got["cell"], want["cell"]), int((got["cell"] != want["cell"]).any(axis=1).sum())
    ge_ref, f_ref = port.execute_batched(pos[None], n_threads=4)
    en, f, ge = k.execute_host(pos, want_grid_energies=True)
    _check(en, ge, f, ge_ref, f_ref)
    _close(g3, k)


def test_identity_order_matches_no_order(gpu_device):
    """The lines kernel with an (identity) evaluation order — gathered positions, scattered force stores — and without
    one (warp-staged positions and stores) agree on the same state."""
    import torch
    import openmmgridforce_b200 as gf
    r, a = 64, 47
    c = _case(3, r, a, seed=5)
    grids, k = _make(gf, gpu_device, c)
    dev = torch.device("cuda:0")
    n = r * a
    d_pos = torch.from_numpy(c["pos"]).to(dev)
    side = torch.cuda.Stream()
    res = []
    for use_order in (False, True):
        d_f = torch.zeros(n, 3, dtype=torch.float64, device=dev)
        d_e = torch.zeros(r, dtype=torch.float64, device=dev)
        d_order = torch.arange(n, dtype=torch.int32, device=dev) if use_order else None
        torch.cuda.synchronize()
        k.execute_device(r, a, d_pos.data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), gf.FORCE_F64_STORE, 0,
                         d_order.data_ptr() if use_order else None, side.cuda_stream)
        torch.cuda.synchronize()
        res.append((d_e.cpu().numpy(), d_f.cpu().numpy()))
    (e0, f0), (e1, f1) = res
    assert np.abs(e0 - e1).max() <= 1e-9 * np.abs(e1).max()
    assert np.abs(f0 - f1).max() <= 2e-6 * np.abs(f1).max()
    _close(grids, k)


def test_host_path_pinned_and_pageable_buffers_agree(gpu_device):
    """gfb_kernel_execute_host: with pinned caller buffers the kernels store forces straight into them (zero-copy), with
    pageable ones into the handle's pinned staging first; both must give bit-identical results, for odd replica counts
    (chunk boundaries), and an ADD call must take the copy pipeline and accumulate."""
    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c5_sharded_replicas(n_local=9001, n=64)
    grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    k = gf.Kernel(gpu_device, grids, w.scaling, oob_k=w.oob_k)
    assert k.uses_lines_kernel()
    e_page, f_page, _ = k.execute_host(w.pos)
    pos_pin = torch.from_numpy(w.pos.copy()).pin_memory()
    f_pin = torch.full(w.pos.shape, 7.0, dtype=torch.float64).pin_memory()
    e_pin = torch.zeros(w.n_replicas, dtype=torch.float64).pin_memory()
    k.execute_host(pos_pin.numpy(), forces=f_pin.numpy(), energies_out=e_pin.numpy())
    assert np.array_equal(f_pin.numpy(), f_page), int((f_pin.numpy() != f_page).sum())
    # a replica's atoms sit in two warps: two FP64 atomics whose order is not fixed
    assert np.abs(e_pin.numpy() - e_page).max() <= 1e-13 * np.abs(e_page).max()
    acc = f_page.copy()
    k.execute_host(w.pos, forces=acc, force_mode=gf.FORCE_F64_ADD)
    assert np.abs(acc - 2.0 * f_page).max() <= 1e-12 * np.abs(f_page).max()
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("n_grids", [1, 3])
def test_launch_overlap_pdl_gives_same_results(gpu_device, n_grids):
    """gfb_kernel_set_launch_overlap: back-to-back device launches made with programmatic stream serialization (a launch's
    blocks fetch their inputs during the previous launch's tail and wait for it before their first write) must leave
    exactly the fixed-point forces and (to summation order) the energies that serialized launches leave — with rotating
    energy accumulators cleared by the previous launch, forces accumulated over all launches into ONE buffer."""
    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c5_sharded_replicas(n_local=6000, n=96)
    grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids[:n_grids]]
    k = gf.Kernel(gpu_device, grids, w.scaling[:n_grids], oob_k=w.oob_k[:n_grids])
    tdev = torch.device("cuda:0")
    r, a = w.n_replicas, w.n_atoms
    stride = ((r * a + 31) // 32) * 32
    rng = np.random.default_rng(3)
    sets = [torch.from_numpy(w.pos + rng.uniform(-0.01, 0.01, size=3)).to(tdev) for _ in range(5)]
    stream = torch.cuda.Stream()
    out = {}
    for pdl in (False, True):
        k.set_launch_overlap(pdl)
        d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
        d_e = [torch.zeros(r, dtype=torch.float64, device=tdev) for _ in range(3)]
        kept = []
        torch.cuda.synchronize()      # the zero fills ran on torch's default stream; `stream` does not order after it
        with torch.cuda.stream(stream):
            for i in range(15):
                k.execute_device(r, a, sets[i % 5].data_ptr(), d_e[i % 3].data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride,
                                 None, stream.cuda_stream, d_energies_clear=d_e[(i + 1) % 3].data_ptr())
                if i >= 12:
                    kept.append(d_e[i % 3].clone())        # stream-ordered copy of step i's energies
        stream.synchronize()
        out[pdl] = (d_f.cpu().numpy(), [x.cpu().numpy() for x in kept])
    k.set_launch_overlap(False)
    assert np.array_equal(out[False][0], out[True][0])             # integer accumulation: order-independent, exact
    for e0, e1 in zip(out[False][1], out[True][1]):
        assert np.abs(e0 - e1).max() <= 1e-13 * np.abs(e0).max()
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("n_grids,n_replicas,mode", [
    (3, 3000, "fixed"),      # 2 204 tiles of 64 atoms: pipeline depth 2 (half of the resident grid), ~1.9 tiles per block
    (3, 9000, "fixed"),      # 6 610 tiles: 2.8 per block, every tile parked
    (3, 12500, "fixed"),     # 9 180 tiles: 3.9 per block — blocks with a 4th tile take the wait inside the loop
    (3, 9000, "none"),       # energy only
    (3, 9000, "f64_add"),
    (2, 9000, "fixed"),
    (4, 9000, "fixed"),
    (1, 9000, "fixed"),      # one grid, batched: 256-thread blocks
])
def test_tile_striding_launches_match_one_block_per_tile(gpu_device, n_grids, n_replicas, mode):
    """Small launches under launch overlap run the tile-striding instantiation of the lines kernel (a resident grid, energy
    sums parked in shared memory, one griddepcontrol.wait at the block's end; gf_launch_lines.cu picks grid and depth).
    Over a sequence of back-to-back launches on rotating accumulators it must leave what serialized one-block-per-tile
    launches leave: fixed-point forces bit for bit, energies to summation order, cleared accumulators cleared."""
    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W
    w = W.c5_sharded_replicas(n_local=n_replicas, n=64)
    g4 = list(w.grids) + [w.grids[0][::-1].copy()]
    s4 = np.concatenate([w.scaling, w.scaling[:1] * 0.5])
    k4 = list(w.oob_k) + [w.oob_k[0]]
    grids = [gf.Grid(gpu_device, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in g4[:n_grids]]
    k = gf.Kernel(gpu_device, grids, s4[:n_grids], oob_k=k4[:n_grids])
    assert k.uses_lines_kernel()
    tdev = torch.device("cuda:0")
    r, a = w.n_replicas, w.n_atoms
    n = r * a
    stride = ((n + 31) // 32) * 32
    rng = np.random.default_rng(11)
    sets = [torch.from_numpy(w.pos + rng.uniform(-0.02, 0.02, size=3)).to(tdev) for _ in range(4)]
    stream = torch.cuda.Stream()
    fmode = {"fixed": gf.FORCE_FIXED_ADD, "f64_add": gf.FORCE_F64_ADD, "none": 0}[mode]
    out = {}
    for pdl in (False, True):
        k.set_launch_overlap(pdl)
        if mode == "fixed":
            d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
        elif mode == "f64_add":
            d_f = torch.zeros(3 * n, dtype=torch.float64, device=tdev)
        else:
            d_f = None
        d_e = [torch.zeros(r, dtype=torch.float64, device=tdev) for _ in range(3)]
        kept = []
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for i in range(9):
                k.execute_device(r, a, sets[i % 4].data_ptr(), d_e[i % 3].data_ptr(), None, d_f.data_ptr() if d_f is not None else None,
                                 fmode, stride, None, stream.cuda_stream, d_energies_clear=d_e[(i + 1) % 3].data_ptr())
                if i >= 6:
                    kept.append(d_e[i % 3].clone())
        stream.synchronize()
        out[pdl] = (d_f.cpu().numpy() if d_f is not None else None, [x.cpu().numpy() for x in kept], d_e[0].cpu().numpy())
    k.set_launch_overlap(False)
    if mode == "fixed":
        assert out[True][0].any() and np.array_equal(out[False][0], out[True][0])
    elif mode == "f64_add":
        assert np.abs(out[False][0] - out[True][0]).max() <= 1e-12 * np.abs(out[False][0]).max()
    for e0, e1 in zip(out[False][1], out[True][1]):
        assert np.abs(e0).max() > 0 and np.abs(e0 - e1).max() <= 1e-13 * np.abs(e0).max()
    assert not out[True][2].any()          # accumulator 0 was cleared by the last launch (step 8 clears (8 + 1) % 3 = 0)
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("atoms", [3, 33])
def test_tile_striding_launch_many_replicas_per_warp_vs_oracle(gpu_device, oracle_built, atoms):
    """The tile-striding instantiation with replicas shorter than (3 atoms: ten run heads per warp, every one parked) and
    barely longer than a warp (33), out-of-grid atoms and zero scaling factors in the mix: device path under launch
    overlap, energies and fixed-point forces of the LAST of four launches against the oracle."""
    import torch
    import openmmgridforce_b200 as gf
    r = 36000 if atoms == 3 else 3300                  # ~108k atoms: more tiles than half the resident grid, fewer than six per block
    c = _case(3, r, atoms, seed=21 + atoms)
    ge_ref, f_ref = _oracle(oracle_built, c)
    grids, k = _make(gf, gpu_device, c)
    assert k.uses_lines_kernel()
    k.set_launch_overlap(True)
    tdev = torch.device("cuda:0")
    n = r * atoms
    stride = ((n + 31) // 32) * 32
    d_pos = torch.from_numpy(c["pos"]).to(tdev)
    d_f = torch.zeros(3 * stride, dtype=torch.int64, device=tdev)
    d_e = [torch.zeros(r, dtype=torch.float64, device=tdev) for _ in range(2)]
    d_out = torch.empty(n, 3, dtype=torch.float64, device=tdev)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    launches = 4
    for i in range(launches):
        k.execute_device(r, atoms, d_pos.data_ptr(), d_e[i % 2].data_ptr(), None, d_f.data_ptr(), gf.FORCE_FIXED_ADD, stride, None,
                         stream.cuda_stream, d_energies_clear=d_e[(i + 1) % 2].data_ptr())
    gpu_device.fixed_to_f64(d_f.data_ptr(), stride, n, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    k.set_launch_overlap(False)
    en = d_e[(launches - 1) % 2].cpu().numpy()
    e_ref = ge_ref.sum(axis=1)
    term = np.abs(ge_ref).max(axis=1)
    assert (np.abs(en - e_ref) <= TOL_E * np.maximum(np.abs(e_ref), term)).all()
    f = d_out.cpu().numpy().reshape(r, atoms, 3) / launches
    assert np.abs(f - f_ref).max() <= TOL_F * np.abs(f_ref).max()
    assert not d_e[launches % 2].cpu().numpy().any()             # cleared by the last launch
    _close(grids, k)
