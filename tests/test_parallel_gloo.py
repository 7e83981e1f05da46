"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: contiguous sharding of the
initialisations and the final all_gather of the per-beta tables (the only collective of the
design).  The per-rank annealing itself is replaced by a deterministic stand-in, since the CUDA
path cannot run here."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeAnnealer(object):
    """Stands in for va_ode.Annealer: 'anneals' by a closed-form function of the inputs."""
    _nX = 6

    def anneal(self, X0, P0, alpha, beta_array, *a, **k):
        B, nb = X0.shape[0], len(beta_array)
        self.tab = np.zeros((B, nb, 5))
        self.minpaths = np.zeros((B, nb, self._nX + P0.shape[1]))
        for b in range(B):
            for i, beta in enumerate(beta_array):
                A = X0[b].sum() * alpha ** beta
                self.tab[b, i] = [beta, A, 0.25 * A, 0.75 * A, 0.75 * A / alpha ** beta]
                self.minpaths[b, i, self._nX:] = P0[b] + i

    def action_errors_table(self, cmpt=0, init=None):
        return self.tab[init]


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    from varanneal_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    X0 = rng.rand(B, 2, 3)
    P0 = rng.rand(B, 2)
    tables, params = parallel.anneal_sharded(_FakeAnnealer(), X0, P0, 1.5, np.arange(4))
    ref = _FakeAnnealer()
    ref.anneal(X0, P0, 1.5, np.arange(4))
    ok = np.array_equal(tables, ref.tab) and np.array_equal(params, ref.minpaths[:, :, 6:])
    lo, hi = parallel.shard_bounds(B, world, rank)
    q.put((rank, bool(ok), lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 8), (2, 5), (3, 4)])
def test_sharded_anneal_equals_single_process(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    # blocks are contiguous, disjoint and cover the batch
    assert res[0][2] == 0 and res[-1][3] == B
    for a, b in zip(res[:-1], res[1:]):
        assert a[3] == b[2]


def test_shard_bounds_edge_cases():
    from varanneal_b200.parallel import shard_bounds
    assert [shard_bounds(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert [shard_bounds(3, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert shard_bounds(0, 2, 1) == (0, 0)
