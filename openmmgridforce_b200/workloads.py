"""Seeded synthetic inputs for the configurations BASELINE.json names.

Pure input generation (numpy): grids, positions, scaling factors. No evaluation happens here.

  C1  10^3 all-ones grid (0.1 nm) + the 47 ligand atoms translated inside  -> E = sum(s) exactly, F = 0
  C2  47 ligand atoms x 3 grids of 208x278x231 @ 0.0125 nm (python/tests/test_simple_grid_energy.py:29-31)
  C3  1,000,000 atoms x one 256^3 grid
  C4  4096 ligand replicas x 3 grids (208x278x231)
  C5  65,536 ligand replicas x 3 grids of 192^3
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np

TEST_GRID_COUNTS = (208, 278, 231)
TEST_GRID_SPACING = 0.0125
TEST_GRID_ORIGIN = (1.00175115, 0.5328844699999999, 0.8606374500000002)
SEED = 1234


@dataclass
class Workload:
    name: str
    counts: tuple
    spacing: tuple
    origin: tuple
    grids: list            # G arrays, each [nx, ny, nz] float64 (x-major, z fastest when ravelled)
    scaling: np.ndarray    # [G, A]
    pos: np.ndarray        # [R, A, 3] nm
    oob_k: list = field(default_factory=list)
    inv_power: list = field(default_factory=list)

    @property
    def n_grids(self):
        return len(self.grids)

    @property
    def n_replicas(self):
        return self.pos.shape[0]

    @property
    def n_atoms(self):
        return self.pos.shape[1]

    @property
    def evals(self):
        """atom-grid evaluations in one pass over the workload"""
        return self.n_replicas * self.n_atoms * self.n_grids


def ligand47():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ligand47.json")
    doc = json.load(open(path))
    return np.array(doc["positions_nm"], dtype=np.float64), np.array(doc["charges_e"], dtype=np.float64)


def synthetic_grid(counts, spacing, seed=SEED, amplitude=10.0, wavelength=0.4, dtype=np.float64):
    """V = A sin(kx) cos(ky) sin(kz) + U(-1, 1): smooth field plus noise, so neighbouring corners differ."""
    nx, ny, nz = counts
    k = 2.0 * np.pi / wavelength
    x = np.sin(k * spacing[0] * np.arange(nx))[:, None, None]
    y = np.cos(k * spacing[1] * np.arange(ny))[None, :, None]
    z = np.sin(k * spacing[2] * np.arange(nz))[None, None, :]
    rng = np.random.default_rng(seed)
    v = rng.uniform(-1.0, 1.0, size=(nx, ny, nz))
    v += amplitude * (x * y * z)
    return v.astype(dtype, copy=False)


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ], 1)


def ligand_replicas(n_replicas, center, seed=SEED, jitter=0.3, escape_every=50, escape_shift=None):
    """[R, 47, 3]: the ligand rigidly rotated and translated by U(-jitter, jitter) nm around `center`.
    Every `escape_every`-th replica is pushed by `escape_shift` so that part of it leaves the grid and the
    out-of-grid restraint branch is exercised (~1-2 % of atoms)."""
    lig, _ = ligand47()
    rng = np.random.default_rng(seed)
    local = lig - lig.mean(axis=0)
    rot = _random_rotations(rng, n_replicas)
    pos = np.einsum("rij,aj->rai", rot, local)
    pos += np.asarray(center)[None, None, :] + rng.uniform(-jitter, jitter, size=(n_replicas, 1, 3))
    if escape_every and escape_shift is not None:
        pos[::escape_every] += np.asarray(escape_shift)[None, None, :]
    return np.ascontiguousarray(pos)


def _ligand_scaling(n_grids, seed=SEED):
    _, q = ligand47()
    rng = np.random.default_rng(seed + 1)
    rows = [q] + [rng.uniform(0.5, 1.5, size=q.size) for _ in range(n_grids - 1)]
    return np.stack(rows[:n_grids])


def c1_ones_grid():
    counts, sp = (10, 10, 10), (0.1, 0.1, 0.1)
    lig, q = ligand47()
    pos = (lig - lig.min(axis=0)) * 0.3 + 0.05        # squeezed into the 0.9 nm box
    return Workload("C1 10^3 ones grid, 47 atoms", counts, sp, (0.0, 0.0, 0.0), [np.ones(counts)], q[None, :].copy(),
                    pos[None].copy(), [10000.0], [0.0])


def c2_single_ligand(counts=TEST_GRID_COUNTS):
    sp = (TEST_GRID_SPACING,) * 3
    lig, _ = ligand47()
    grids = [synthetic_grid(counts, sp, seed=SEED + g) for g in range(3)]
    return Workload("C2 single ligand x 3 grids", tuple(counts), sp, TEST_GRID_ORIGIN, grids, _ligand_scaling(3),
                    lig[None].copy(), [10000.0] * 3, [0.0] * 3)


def c3_million_atoms(n_atoms=1_000_000, n=256, seed=SEED):
    counts, sp = (n, n, n), (TEST_GRID_SPACING,) * 3
    rng = np.random.default_rng(seed)
    length = sp[0] * (n - 1)
    pos = rng.uniform(0.0, 0.999 * length, size=(1, n_atoms, 3))
    scaling = rng.uniform(0.5, 1.5, size=(1, n_atoms))
    return Workload(f"C3 {n_atoms} atoms x one {n}^3 grid", counts, sp, (0.0, 0.0, 0.0), [synthetic_grid(counts, sp, seed)],
                    scaling, pos, [10000.0], [0.0])


def c4_batched_replicas(n_replicas=4096, counts=TEST_GRID_COUNTS):
    sp = (TEST_GRID_SPACING,) * 3
    lig, _ = ligand47()
    grids = [synthetic_grid(counts, sp, seed=SEED + g) for g in range(3)]
    pos = ligand_replicas(n_replicas, lig.mean(axis=0), escape_shift=(1.0, 0.0, 0.0))
    return Workload(f"C4 {n_replicas} replicas x 47 atoms x 3 grids", tuple(counts), sp, TEST_GRID_ORIGIN, grids,
                    _ligand_scaling(3), pos, [10000.0] * 3, [0.0] * 3)


def c5_sharded_replicas(n_replicas=65536, n=192, replica_offset=0, n_local=None, pose_seed=SEED):
    """Replicas [replica_offset, replica_offset + n_local) of the 65,536-replica batch (a rank's shard).
    The whole batch is generated from one seed and sliced, so shards are identical however they are cut.
    `pose_seed` changes the replica poses only (grids and scaling factors stay): bench.py's weak-scaling runs give
    every rank its own 65,536 poses with pose_seed = SEED + rank instead of generating N x 65,536 and slicing."""
    counts, sp = (n, n, n), (TEST_GRID_SPACING,) * 3
    grids = [synthetic_grid(counts, sp, seed=SEED + g) for g in range(3)]
    half = 0.5 * sp[0] * (n - 1)
    pos = ligand_replicas(n_replicas, (half, half, half), seed=pose_seed, escape_shift=(0.9, 0.0, 0.0))
    if n_local is not None:
        pos = np.ascontiguousarray(pos[replica_offset:replica_offset + n_local])
    return Workload(f"C5 {n_replicas} replicas x 47 atoms x 3 grids of {n}^3", counts, sp, (0.0, 0.0, 0.0), grids,
                    _ligand_scaling(3), pos, [10000.0] * 3, [0.0] * 3)


def mixed_energy_bound(w, pos):
    """Per-replica sum over atoms and grids of |s| * trilinear(|V|): the scale of the irreducible MIXED-precision energy
    error. Grid values are stored in FP32 (relative rounding 2^-24 = 6e-8 per corner) and the trilinear weights are
    non-negative and sum to one, so however exact the arithmetic, a replica's energy can differ from the FP64 reference by
    up to 6e-8 times this bound — which is why a replica whose terms cancel (|E| much smaller than the sum of |terms|)
    cannot meet 1e-6 of |E| and the parity tests assert  |E - E_ref| <= max(1e-6 |E_ref|, 6e-8 * bound).
    Input-side helper (numpy only, no evaluation path involved). pos: [R, A, 3]; atoms outside the grid contribute 0."""
    pos = np.asarray(pos, dtype=np.float64)
    sp, og = np.asarray(w.spacing), np.asarray(w.origin)
    counts = np.asarray(w.counts)
    rel = (pos - og) / sp
    inside = ((pos - og >= 0) & (pos - og <= sp * (counts - 1))).all(axis=-1)
    idx = np.clip(np.floor(rel).astype(np.int64), 0, counts - 2)
    f = np.clip(rel - idx, 0.0, 1.0)
    out = np.zeros(pos.shape[0])
    for g in range(w.n_grids):
        v = np.abs(np.asarray(w.grids[g]))
        acc = np.zeros(pos.shape[:2])
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    wgt = (f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1]) * (f[..., 2] if dz else 1 - f[..., 2])
                    acc += wgt * v[idx[..., 0] + dx, idx[..., 1] + dy, idx[..., 2] + dz]
        out += (np.abs(w.scaling[g])[None, :] * acc * inside).sum(axis=1)
    return out
