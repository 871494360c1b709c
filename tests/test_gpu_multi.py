"""Replica-sharded multi-GPU entry points of the C ABI (-m gpu): gfb_multi_* (one process, N devices) and gfb_comm_* (one
process per GPU), including the fused in-kernel energy gather over peer memory, against the oracle and against each
other. World size 1 runs on any GPU box (the tail-gather code path — ticket, copy, flag, wait — is the same); world
size 2 needs `gpurun --gpus 2` and is skipped elsewhere. Replaces the reference's sequential replica loop
(example/sampler.py:130-164); the sharding rule is SURVEY.md §8(e)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import openmmgridforce_b200 as gf
    return gf.Device.count()


def _workload(n_total):
    from openmmgridforce_b200 import workloads as W
    return W.c5_sharded_replicas(n_replicas=n_total, n=48)


def _oracle(bindings, w, pos):
    port = bindings.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling, oob_k=w.oob_k)
    ge, f = port.execute_batched(pos, n_threads=8)
    return ge.sum(axis=1), f


@pytest.mark.parametrize("n_dev", [1, 2])
def test_multi_host_path_matches_oracle(oracle_built, n_dev):
    import openmmgridforce_b200 as gf
    if _n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    w = _workload(3001)
    e_ref, f_ref = _oracle(oracle_built, w, w.pos)
    m = gf.Multi(list(range(n_dev)))
    for v in w.grids:
        m.add_grid(w.counts, w.spacing, w.origin, v)
    m.build(w.scaling, oob_k=w.oob_k)
    en, f = m.execute_host(w.pos)
    assert np.abs(en - e_ref).max() <= 1e-6 * np.abs(e_ref).max()
    assert np.abs(f - f_ref).max() <= 1e-5 * np.abs(f_ref).max()
    en32, f32 = m.execute_host(w.pos, force_mode=gf.FORCE_F32_STORE)
    assert f32.dtype == np.float32 and np.array_equal(f32, f.astype(np.float32))
    en0, none = m.execute_host(w.pos, want_forces=False)
    assert none is None and np.abs(en0 - en).max() <= 1e-13 * np.abs(en).max()
    m.close()


@pytest.mark.parametrize("n_dev", [1, 2])
def test_multi_resident_steps_and_gathers(oracle_built, n_dev):
    """Device-resident shards: one launch per device per step, energies gathered by ncclAllGather (1) and by the fused
    in-kernel gather over peer memory (2), the stand-alone push kernel (3) or the one-kernel flag-in-data gather (4); every
    device ends up with all energies; forces accumulate in fixed point."""
    import openmmgridforce_b200 as gf
    if _n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    w = _workload(2001)
    e_ref, f_ref = _oracle(oracle_built, w, w.pos)
    m = gf.Multi(list(range(n_dev)))
    for v in w.grids:
        m.add_grid(w.counts, w.spacing, w.origin, v)
    m.build(w.scaling, oob_k=w.oob_k)
    m.upload(w.pos)
    steps = 0
    for gather in (0, 2, 1, 3, 4, 2, 4, 4, 3):
        m.step(gather)
        steps += 1
        for d in range(n_dev):
            en, f = m.download(from_device=d, want_forces=(d == 0))
            assert np.abs(en - e_ref).max() <= 1e-6 * np.abs(e_ref).max(), (gather, d)
            if f is not None:
                assert np.abs(f - steps * f_ref).max() <= 1e-5 * steps * np.abs(f_ref).max()
    m.close()


@pytest.mark.parametrize("world", [1, 2])
def test_comm_fused_gather_matches_nccl_and_oracle(oracle_built, world):
    """One process per GPU: every rank evaluates its shard with gfb_kernel_execute_device_gather (the launch's last block
    stores the energies into every rank's gathered array and raises its flag) and waits with gfb_comm_gather_wait; the
    result must equal ncclAllGather of the same energies bit for bit, and the oracle within tolerance, on every rank."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    n_total = 1001
    with tempfile.TemporaryDirectory() as xdir:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "helpers", "comm_worker.py"), str(r), str(world), xdir,
                                   str(n_total)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=420)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for r, p in enumerate(procs):
            assert p.returncode == 0, f"rank {r}:\n{outs[r][-3000:]}"
        w = _workload(n_total)
        res = [np.load(os.path.join(xdir, f"out{r}.npz")) for r in range(world)]
    shifts = res[0]["shifts"]
    for step in range(3):
        e_ref, _ = _oracle(oracle_built, w, w.pos + shifts[step])
        for r in range(world):
            fused = res[r]["fused"][step]
            width = int(res[r]["width"])
            padded = res[r]["nccl"][step]
            via_nccl = np.concatenate([padded[q * width:q * width + int(res[q]["hi"]) - int(res[q]["lo"])] for q in range(world)])
            assert np.array_equal(fused, via_nccl), (step, r)
            assert np.array_equal(res[r]["ll"][step], via_nccl), (step, r)      # gfb_comm_gather (flag-in-data, one kernel)
            assert np.abs(fused - e_ref).max() <= 1e-6 * np.abs(e_ref).max(), (step, r)
