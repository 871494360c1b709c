"""Replica sharding across GPUs (one process per GPU) and the one collective of the path.

Replicas are independent (each atom's loop iteration is, ReferenceGridForceKernels.cpp:682-1118; replicas are separate
Contexts in example/sampler.py:130-151), so rank g simply owns a contiguous block of replicas, grids are replicated on
every GPU, and nothing is exchanged on the force path. The only communication is the collection of per-replica
energies: an all-gather of R_local doubles per rank over NCCL (gloo in the CPU tests)."""


def shard_bounds(n_units, world_size, rank):
    """Block partition: rank g owns [g*n/N, (g+1)*n/N). Sizes differ by at most one; order is preserved."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    lo = n_units * rank // world_size
    hi = n_units * (rank + 1) // world_size
    return lo, hi


def gather_energies(dist, local_energies, out=None, group=None):
    """All-gather equally sized per-replica energy vectors; result[r*R_local + i] is replica i of rank r, i.e. the
    global replica order of shard_bounds. `local_energies` is a 1-D torch tensor on the rank's device."""
    import torch
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_energies
    if out is None:
        out = torch.empty(world * local_energies.numel(), dtype=local_energies.dtype, device=local_energies.device)
    dist.all_gather_into_tensor(out, local_energies, group=group)
    return out


def gather_energies_ragged(dist, local_energies, n_units, group=None):
    """Same for block partitions whose sizes differ by one (n_units not divisible by the world size): pads to the
    largest shard, gathers, and strips the padding."""
    import torch
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_units, world, r)[1] - shard_bounds(n_units, world, r)[0] for r in range(world)]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local_energies.dtype, device=local_energies.device)
    padded[:local_energies.numel()] = local_energies
    out = torch.empty(world * width, dtype=local_energies.dtype, device=local_energies.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width:r * width + sizes[r]] for r in range(world)])
