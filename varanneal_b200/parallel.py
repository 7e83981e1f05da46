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


def anneal_sharded(annealer, X0, P0, *args, group=None, gather_paths='none', **kwargs):
    """Anneal rank-local slices of a batch (X0 (B, N, D), P0 (B, NP) -- or (B, N, NP) for parameter
    time series) and gather the (B, Nbeta, 5) action tables [beta, A, me, fe, fe/RF] and the
    (B, Nbeta, NP) [(B, Nbeta, N, NP)] parameters.
    Positional / keyword arguments after P0 are those of ``Annealer.anneal``.

    ``gather_paths``: 'none' (default: the minimising paths stay with the rank that computed them
    -- at D = 1000, N = 100000 they are 0.8 GB per path and rung and must not be gathered,
    SURVEY.md 8(e)), 'last' (the last rung's paths, (B, n_states + NP)) or 'all'
    ((B, Nbeta, n_states + NP); needs keep_paths == 'all').  With 'last' / 'all' a third array is
    returned."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = P0.shape[0] if callable(X0) else X0.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    nb = len(args[1])
    if hi > lo:
        x_loc = (lambda b0, b1: X0(lo + b0, lo + b1)) if callable(X0) else X0[lo:hi]
        annealer.anneal(x_loc, P0[lo:hi], *args, **kwargs)
        tables = np.stack([annealer.action_errors_table(init=i) for i in range(hi - lo)])
        params = annealer.params_array if hasattr(annealer, "params_array") else annealer.minpaths[:, :, annealer._nX:]
        paths = annealer.minpaths
    else:
        tables = np.zeros((0, nb, 5))
        params = np.zeros((0, nb) + tuple(P0.shape[1:]))       # (NP,) or, for parameter time series, (N_model, NP)
        paths = None
    out = (gather_blocks(tables, B, group), gather_blocks(params, B, group))
    if gather_paths == 'none':
        return out
    if gather_paths not in ('last', 'all'):
        raise ValueError("gather_paths must be 'none', 'last' or 'all'")
    import torch
    width = torch.tensor([0 if paths is None else paths.shape[-1]], dtype=torch.int64)
    if dist.is_initialized():
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        width = width.to(dev)
        dist.all_reduce(width, op=dist.ReduceOp.MAX, group=group)
    w = int(width.item())
    if paths is None:
        loc = np.zeros((0, w) if gather_paths == 'last' else (0, nb, w))
    elif gather_paths == 'last':
        if paths.shape[-2] == 0:
            raise ValueError("gather_paths='last' needs keep_paths 'all' or 'last'")
        loc = paths[:, -1]
    else:
        if paths.shape[-2] != nb:
            raise ValueError("gather_paths='all' needs keep_paths='all'")
        loc = paths
    return out + (gather_blocks(loc, B, group),)
