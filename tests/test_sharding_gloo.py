"""World-size-2 (and 3, ragged) test of the multi-GPU host logic on CPU with the gloo backend: every rank builds its
shard of the replica batch exactly as bench.py does, evaluates it (here with the CPU oracle standing in for the GPU),
all-gathers the per-replica energies, and every rank must end up with the energies of the whole batch in global
replica order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmmgridforce_b200 import sharding
from openmmgridforce_b200 import workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_replicas, ragged, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import bindings
        lo, hi = sharding.shard_bounds(n_replicas, world, rank)
        w = W.c5_sharded_replicas(n_replicas=n_replicas, n=24, replica_offset=lo, n_local=hi - lo)
        port_oracle = bindings.PortOracle(w.counts, w.spacing, w.origin, w.grids, w.scaling, oob_k=w.oob_k)
        ge, _ = port_oracle.execute_batched(w.pos, want_forces=False)
        local = torch.from_numpy(ge.sum(axis=1))
        if ragged:
            full = sharding.gather_energies_ragged(dist, local, n_replicas)
        else:
            full = sharding.gather_energies(dist, local)
        ret[rank] = full.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_replicas,ragged", [(2, 64, False), (3, 50, True)])
def test_sharded_energies_match_whole_batch(oracle_built, world, n_replicas, ragged):
    whole = W.c5_sharded_replicas(n_replicas=n_replicas, n=24)
    port_oracle = oracle_built.PortOracle(whole.counts, whole.spacing, whole.origin, whole.grids, whole.scaling, oob_k=whole.oob_k)
    ge, _ = port_oracle.execute_batched(whole.pos, want_forces=False)
    expect = ge.sum(axis=1)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_replicas, ragged, ret), nprocs=world, join=True)
    for rank in range(world):
        assert np.array_equal(ret[rank], expect), f"rank {rank} does not hold the whole batch's energies in order"


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            bounds = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_workload_shards_are_slices_of_the_whole():
    whole = W.c5_sharded_replicas(n_replicas=40, n=16)
    lo, hi = sharding.shard_bounds(40, 3, 1)
    part = W.c5_sharded_replicas(n_replicas=40, n=16, replica_offset=lo, n_local=hi - lo)
    assert np.array_equal(part.pos, whole.pos[lo:hi]) and np.array_equal(part.scaling, whole.scaling)
