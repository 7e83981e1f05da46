"""Sharding of independent annealing initialisations over the GPUs of one box.

The reference parallelises over initial paths with an SGE array job, one OS process per
initialisation and no communication (examples/nnet_barimages/SGEcluster/submit_multiM.sh:18-30).
Here: one process per GPU (torchrun), rank r anneals the contiguous block
``shard_bounds(B, world, r)`` of the batch on its own device with no collective on the hot path;
a single ``all_gather`` at the very end (NCCL over NVLink, or gloo in the CPU tests) collects the
per-beta action tables -- and, if asked, the estimated parameters -- on every rank.
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """[lo, hi) of the contiguous block of items owned by ``rank`` (blocks of ceil(n/world))."""
    per = -(-int(n_items) // int(world))
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def gather_blocks(local, n_items, group=None):
    """all_gather of per-rank blocks along axis 0: ``local`` is this rank's
    (hi - lo, ...) float64 array; returns the full (n_items, ...) array on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-int(n_items) // world)
    local = np.ascontiguousarray(local, dtype=np.float64)
    tail = local.shape[1:]
    lo, hi = shard_bounds(n_items, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d items, expected %d" % (rank, local.shape[0], hi - lo))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros((per,) + tail, dtype=torch.float64, device=dev)
    if hi > lo:
        buf[:hi - lo] = torch.from_numpy(local).to(dev)
    out = torch.empty((world * per,) + tail, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:n_items].cpu().numpy()


def anneal_sharded(annealer, X0, P0, *args, group=None, **kwargs):
    """Anneal rank-local slices of a batch (X0 (B, N, D), P0 (B, NP)) and gather the
    (B, Nbeta, 5) action tables [beta, A, me, fe, fe/RF] and the (B, Nbeta, NP) parameters.
    Positional / keyword arguments after P0 are those of ``Annealer.anneal``."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = X0.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    if hi > lo:
        annealer.anneal(X0[lo:hi], P0[lo:hi], *args, **kwargs)
        tables = np.stack([annealer.action_errors_table(init=i) for i in range(hi - lo)])
        params = annealer.minpaths[:, :, annealer._nX:]
    else:
        tables = np.zeros((0, len(args[1]), 5))
        params = np.zeros((0, len(args[1]), P0.shape[1]))
    return gather_blocks(tables, B, group), gather_blocks(params, B, group)
