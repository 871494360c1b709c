"""Grid generation from receptor atoms: the oracle restatement against the reference's own
auto-generation path (CPU), and the FP64 CUDA kernel against the oracle (GPU)."""
import numpy as np
import pytest

COUNTS, SPACING, ORIGIN = (13, 11, 9), (0.125, 0.15, 0.2), (0.1, -0.05, 0.05)


def _receptor(n=300, seed=0):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-0.3, 2.0, size=(n, 3))
    pos[0] = np.array(ORIGIN) + np.array(SPACING) * (3, 2, 1)        # an atom exactly on a grid point: r clamps to 1e-6
    return pos, rng.normal(size=n) * 0.4, rng.uniform(0.1, 0.2, n), rng.uniform(0.1, 1.0, n)


@pytest.mark.parametrize("grid_type", ["charge", "ljr", "lja"])
def test_port_matches_reference_generation(oracle_built, grid_type):
    if not oracle_built.ref_available():
        pytest.skip("needs oracle/_ref (the reference's auto-generation path)")
    pos, q, sg, ep = _receptor()
    ref = oracle_built.ref_generate_grid(COUNTS, SPACING, ORIGIN, grid_type, pos, q, sg, ep, grid_cap=41840.0)
    port = oracle_built.port_generate_grid(COUNTS, SPACING, ORIGIN, grid_type, pos, q, sg, ep, grid_cap=41840.0, n_threads=3)
    assert np.array_equal(ref, port)
    assert np.abs(ref).max() <= 41840.0 and np.isfinite(ref).all()


def test_generation_known_answer(oracle_built):
    """One unit charge at distance d from a point: U*tanh(138.935456/d / U)."""
    pos = np.array([[0.0, 0.0, 0.0]])
    v = oracle_built.port_generate_grid((3, 1, 1), (0.5, 1, 1), (0.5, 0, 0), "charge", pos, [1.0], [0.1], [0.1], grid_cap=41840.0)
    expect = 41840.0 * np.tanh(138.935456 / np.array([0.5, 1.0, 1.5]) / 41840.0)
    assert np.allclose(v.ravel(), expect, rtol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("grid_type", ["charge", "ljr", "lja"])
def test_cuda_generation_matches_oracle(gpu_device, oracle_built, grid_type):
    import openmmgridforce_b200 as gf
    pos, q, sg, ep = _receptor(n=700, seed=3)                  # 3 shared-memory tiles, the last one partial
    want = oracle_built.port_generate_grid(COUNTS, SPACING, ORIGIN, grid_type, pos, q, sg, ep, grid_cap=41840.0, n_threads=4)
    grid, got = gf.Grid.generate(gpu_device, COUNTS, SPACING, ORIGIN, grid_type, pos, q, sg, ep, precision=gf.PRECISION_DOUBLE)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    assert np.median(rel) < 1e-14
    # the generated grid is directly usable: evaluate a few atoms on it and on an upload of the oracle's values
    rng = np.random.default_rng(1)
    atoms = np.array(ORIGIN) + rng.uniform(0, 1, size=(50, 3)) * (np.array(SPACING) * (np.array(COUNTS) - 1))
    sc = rng.normal(size=(1, 50))
    k1 = gf.Kernel(gpu_device, [grid], sc)
    g2 = gf.Grid(gpu_device, COUNTS, SPACING, ORIGIN, want, gf.PRECISION_DOUBLE)
    k2 = gf.Kernel(gpu_device, [g2], sc)
    e1, f1, _ = k1.execute_host(atoms)
    e2, f2, _ = k2.execute_host(atoms)
    assert abs(e1[0] - e2[0]) <= 1e-9 * abs(e2[0]) and np.abs(f1 - f2).max() <= 1e-9 * np.abs(f2).max()
    for o in (k1, k2, grid, g2):
        o.close()


@pytest.mark.gpu
def test_cuda_generation_errors(gpu_device):
    import openmmgridforce_b200 as gf
    with pytest.raises(gf.GridForceB200Error, match="Invalid grid type"):
        gf.Grid.generate(gpu_device, COUNTS, SPACING, ORIGIN, "dipole", np.zeros((1, 3)), [0.0], [0.1], [0.1])
    with pytest.raises(gf.GridForceB200Error):
        gf.Grid.generate(gpu_device, COUNTS, SPACING, ORIGIN, "ljr", np.zeros((1, 3)), charges=[0.0])   # needs sigma/epsilon
