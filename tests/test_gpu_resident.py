"""GPU tests of the resident evaluator (gfb_kernel_set_resident, csrc/gf_resident.cu): one ligand per MD step served by a
block that stays on the GPU — the same answers as the launch-per-step path and as the oracle
(ReferenceCalcGridForceKernel::execute, platforms/reference/src/ReferenceGridForceKernels.cpp:646-1121), step after step,
across idle time-outs, parameter updates, particle subsets, released cells, and through the platform property.

Tolerances as everywhere: MIXED energy 1e-6 / forces 1e-5 (max-norm), DOUBLE 1e-12.
"""
import os
import sys
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {0: (1e-6, 1e-5), 1: (1e-12, 1e-12)}


def _rel_f(f, ref):
    return np.abs(f - ref).max() / max(np.abs(ref).max(), 1e-300)


def _make(gf, dev, c, precision, particles=None, n_grids=None):
    vals = c["grids"][:n_grids] if n_grids else c["grids"]
    grids = [gf.Grid(dev, c["counts"], c["spacing"], c["origin"], g, precision, layout=gf.LAYOUT_CELLS) for g in vals]
    ng = len(grids)
    k = gf.Kernel(dev, grids, c["scaling"][:ng], particles=particles, inv_power=c["inv_power"][:ng], oob_k=c["oob_k"][:ng])
    return grids, k


def _close(grids, k):
    k.close()
    for g in grids:
        g.close()


@pytest.mark.parametrize("precision", [0, 1], ids=["mixed", "double"])
@pytest.mark.parametrize("n_grids", [1, 3])
def test_resident_steps_match_oracle_and_launch_path(gpu_device, oracle_built, precision, n_grids):
    """40 steps of a moving ligand: every step's energies, per-grid energies and forces against the oracle and against
    the launch-per-step path; the block is launched once."""
    import openmmgridforce_b200 as gf
    c, _ = cases.load_golden("ligand_three_grids")
    grids, k = _make(gf, gpu_device, c, precision, n_grids=n_grids)
    grids2, k2 = _make(gf, gpu_device, c, precision, n_grids=n_grids)
    port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"][:n_grids], c["scaling"][:n_grids],
                                   oob_k=c["oob_k"][:n_grids])
    k2.execute_host(c["pos"], want_grid_energies=True)   # the launch path's kernel is loaded before the block takes up residence: loading a
    k.set_resident(True)             # module while it spins would wait for its idle time-out (a one-off 100 ms, harmless)
    rng = np.random.default_rng(1)
    tol_e, tol_f = TOL[precision]
    for step in range(40):
        shift = rng.uniform(-0.4, 0.4, size=3) if step % 7 else rng.uniform(-1.2, 1.2, size=3)   # every 7th: partly outside
        pos = c["pos"] + shift
        ge_ref, f_ref = port.execute_batched(pos[None])
        en, forces, ge = k.execute_host(pos, want_grid_energies=True)
        en2, forces2, ge2 = k2.execute_host(pos, want_grid_energies=True)
        scale = max(np.abs(ge_ref).max(), 1e-300)
        assert np.abs(ge[0] - ge_ref[0]).max() <= tol_e * scale, step
        assert abs(en[0] - ge_ref[0].sum()) <= tol_e * max(abs(ge_ref[0].sum()), scale), step
        assert _rel_f(forces[0], f_ref[0]) <= tol_f, step
        assert np.abs(ge[0] - ge2[0]).max() <= tol_e * scale and _rel_f(forces[0], forces2[0]) <= tol_f
    assert k.resident_launches() <= 2        # 1, unless the driver loaded another module meanwhile
    tl = k.resident_timeline()
    assert tl.shape == (2,) and np.all(tl >= 0.0) and tl.sum() < 1000.0        # microseconds on the GPU per step
    # energy only, and forces only for a caller that wants no energies
    en, none, _ = k.execute_host(c["pos"], want_forces=False)
    assert none is None and abs(en[0] - k2.execute_host(c["pos"])[0][0]) <= tol_e * abs(en[0])
    k.set_resident(False)
    assert k.resident_launches() == 0
    en3, f3, _ = k.execute_host(c["pos"])
    assert abs(en3[0] - en[0]) <= tol_e * abs(en[0])
    _close(grids, k)
    _close(grids2, k2)


def test_resident_idle_timeout_relaunches(gpu_device, oracle_built):
    """The block leaves the GPU after the idle time and the next step brings it back; the answers do not change."""
    import openmmgridforce_b200 as gf
    c, ref = cases.load_golden("ligand_three_grids")
    grids, k = _make(gf, gpu_device, c, 1)
    k.set_resident(True, idle_us=10000)
    for n in range(1, 4):
        en, forces, _ = k.execute_host(c["pos"])
        assert abs(en[0] - ref["energy"]) <= 1e-12 * abs(ref["energy"])
        assert _rel_f(forces[0], ref["forces"]) <= 1e-12
        assert k.resident_launches() == n
        time.sleep(0.15)                 # 15x the idle time: the block is gone
    for _ in range(200):                 # back-to-back: no further launches
        k.execute_host(c["pos"])
    launches = k.resident_launches()
    assert launches in (4, 5)            # one more after the last sleep, then none (5: the interpreter paused for > 10 ms once)
    k.resident_stop()
    en, _, _ = k.execute_host(c["pos"])
    assert k.resident_launches() == launches + 1 and abs(en[0] - ref["energy"]) <= 1e-12 * abs(ref["energy"])
    _close(grids, k)


def test_resident_particle_subset_store_add_and_update(gpu_device, oracle_built):
    """A particle map (forces of the other particles stay / accumulate as in the launch path), F64_ADD, and
    gfb_kernel_update_parameters while the block is up (it is stopped and picks up the new factors)."""
    import openmmgridforce_b200 as gf
    c, _ = cases.load_golden("random_aniso")
    rng = np.random.default_rng(5)
    n_particles = 300
    particles = rng.permutation(n_particles)[:120].astype(np.int32)
    pos = c["pos"][:n_particles]
    sc = c["scaling"][:, :120].copy()
    cc = dict(c, scaling=sc)
    grids, k = _make(gf, gpu_device, cc, 1, particles=particles)
    k.set_resident(True)

    def oracle(scaling):
        port = oracle_built.PortOracle(c["counts"], c["spacing"], c["origin"], c["grids"], scaling, oob_k=c["oob_k"])
        f_ref = np.zeros((n_particles, 3))
        e_ref = 0.0
        for g in range(2):
            e, f, _ = port.execute(pos, g, ligand_atoms=particles)
            f_ref[particles] += f
            e_ref += e
        return e_ref, f_ref

    e_ref, f_ref = oracle(sc)
    marker = rng.normal(size=(1, n_particles, 3))
    forces = marker.copy()
    en, forces, _ = k.execute_host(pos, forces=forces)                              # STORE
    untouched = np.setdiff1d(np.arange(n_particles), particles)
    assert abs(en[0] - e_ref) <= 1e-12 * abs(e_ref)
    assert np.array_equal(forces[0, untouched], marker[0, untouched])
    assert np.abs(forces[0, particles] - f_ref[particles]).max() <= 1e-12 * np.abs(f_ref).max()
    acc = marker.copy()
    k.execute_host(pos, forces=acc, force_mode=gf.FORCE_F64_ADD)                    # ADD
    assert np.abs(acc[0] - (marker[0] + f_ref)).max() <= 1e-12 * np.abs(f_ref).max()
    launches = k.resident_launches()
    k.update_parameters(scaling=2.0 * sc)
    e2, f2 = oracle(2.0 * sc)
    en, forces, _ = k.execute_host(pos)
    assert abs(en[0] - e2) <= 1e-12 * abs(e2)
    assert np.abs(forces[0] - f2).max() <= 1e-12 * np.abs(f2).max()
    assert k.resident_launches() == launches + 1
    _close(grids, k)


def test_resident_released_cells_and_refusals(gpu_device, oracle_built):
    """After gfb_grid_release_cells the block reads the interleaved records; layouts / sizes that do not qualify are
    refused with GFB_ERR_UNSUPPORTED and keep the launch path."""
    import openmmgridforce_b200 as gf
    c, ref = cases.load_golden("ligand_three_grids")
    grids, k = _make(gf, gpu_device, c, 0)
    for g in grids:
        g.release_cells()
    k.set_resident(True)
    en, forces, ge = k.execute_host(c["pos"], want_grid_energies=True)
    assert abs(en[0] - ref["energy"]) <= 1e-6 * abs(ref["energy"])
    assert np.abs(ge[0] - ref["grid_energies"]).max() <= 1e-6 * np.abs(ref["grid_energies"]).max()
    assert _rel_f(forces[0], ref["forces"]) <= 1e-5
    _close(grids, k)
    b = [gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], c["grids"][0], 0, layout=gf.LAYOUT_BSPLINE)]
    kb = gf.Kernel(gpu_device, b, c["scaling"][:1])
    with pytest.raises(gf.GridForceB200Error, match="trilinear packed cells"):
        kb.set_resident(True)
    _close(b, kb)
    big = [gf.Grid(gpu_device, c["counts"], c["spacing"], c["origin"], c["grids"][0], 0, layout=gf.LAYOUT_CELLS)]
    kbig = gf.Kernel(gpu_device, big, np.ones((1, 300)))
    with pytest.raises(gf.GridForceB200Error):
        kbig.set_resident(True)
    _close(big, kbig)


@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_plugin_resident_kernel_property(precision):
    """Context(..., {"ResidentKernel": "true"}): the golden ligand through the platform plugin, three fused GridForces,
    every force group, repeated evaluations."""
    import openmmgridforce_b200.gridforceplugin as gfp
    from test_plugin import _build_system
    c, ref = cases.load_golden("ligand_three_grids")
    platform = gfp.Platform.getPlatformByName("B200")
    system, forces = _build_system(gfp, c)
    context = gfp.Context(system, platform, {"Precision": precision, "ResidentKernel": "true", "ResidentIdleMicroseconds": "50000"})
    te, tf = (1e-6, 1e-5) if precision == "mixed" else (1e-12, 1e-12)
    for rep in range(5):
        context.setPositions(c["pos"])
        state = context.getState(getEnergy=True, getForces=True)
        assert abs(state.getPotentialEnergy() - ref["energy"]) <= te * abs(ref["energy"])
        assert np.abs(state.getForces() - ref["forces"]).max() <= tf * np.abs(ref["forces"]).max()
        for g in range(len(forces)):
            s = context.getState(getEnergy=True, getForces=True, groups=1 << g)
            assert abs(s.getPotentialEnergy() - ref["grid_energies"][g]) <= te * max(abs(ref["grid_energies"][g]), 1e-300)
            assert np.abs(s.getForces() - ref["grid_forces"][g]).max() <= tf * np.abs(ref["grid_forces"][g]).max()
    assert platform.getPropertyValue(context, "ResidentKernel") == "true"
    with pytest.raises(RuntimeError, match="ResidentKernel must be"):
        gfp.Context(system, platform, {"ResidentKernel": "maybe"})
