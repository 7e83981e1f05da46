"""GPU test of the multi-GPU path: a 2-rank sharded anneal -- the real
va_ode.Annealer on every rank, contiguous blocks of initial paths, no collective until the final
gather -- equals the 1-rank anneal of the whole batch bit for bit (tables, parameters, last-rung
paths).  With two visible GPUs every rank takes its own device and the gather runs over NCCL;
on a one-GPU box both ranks share cuda:0 and gather over gloo (the arithmetic per path does not
depend on which paths share a launch, so the comparison is just as strict).  Also: the wave
scheduler (initialisations that do not fit the device at once) reproduces the all-resident run."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.RandomState(31)
    D, N, B = 20, 61, 6
    Lidx = [0, 3, 6, 9, 12, 15, 18]
    Y = rng.randn(N, len(Lidx))
    t = 0.025 * np.arange(N)
    X0 = rng.randn(B, N, D)
    P0 = 8.0 + 0.2 * rng.randn(B, 1)
    return D, N, B, Lidx, Y, t, X0, P0


ARGS = dict(disc="SimpsonHermite", init_to_data=True, opt_args={"gtol": 1e-9, "ftol": 1e-12, "maxiter": 300})


def _anneal_whole(device=None, wave=None, keep='all'):
    from varanneal_b200 import va_ode
    D, N, B, Lidx, Y, t, X0, P0 = _problem()
    an = va_ode.Annealer(device=device)
    an.wave_size, an.keep_paths = wave, keep
    an.set_model("lorenz96", D)
    an.set_data(Y, t=t)
    an.anneal(X0.copy(), P0.copy(), 1.5, [8, 12, 16], 4.0, 4e-6, Lidx, [0], **ARGS)
    return an


def _worker(rank, world, port, ngpu, q):
    import torch
    import torch.distributed as dist
    from varanneal_b200 import parallel, va_ode
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if ngpu >= world else 0
    torch.cuda.set_device(dev)
    backend = "nccl" if ngpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    D, N, B, Lidx, Y, t, X0, P0 = _problem()
    an = va_ode.Annealer(device=dev)
    an.set_model("lorenz96", D)
    an.set_data(Y, t=t)
    tables, params, last = parallel.anneal_sharded(an, X0.copy(), P0.copy(), 1.5, [8, 12, 16], 4.0, 4e-6, Lidx, [0],
                                                   gather_paths='last', **ARGS)
    q.put((rank, backend, tables, params, last))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_anneal_equals_single_rank_bit_for_bit():
    import torch
    ref = _anneal_whole()
    B = ref.A_array.shape[0]
    ref_tab = np.stack([ref.action_errors_table(init=i) for i in range(B)])
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ngpu = torch.cuda.device_count()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ngpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for rank, backend, tables, params, last in res:
        assert np.array_equal(tables, ref_tab), (rank, backend)
        assert np.array_equal(params, ref.params_array)
        assert np.array_equal(last, ref.minpaths[:, -1])


@pytest.mark.parametrize("wave,keep", [(4, 'all'), (1, 'last'), (5, 'none')])
def test_waves_reproduce_the_all_resident_run(wave, keep):
    """6 initialisations with 4 / 1 / 5 resident at a time (ragged last wave) and the three
    keep_paths policies: same tables, parameters and kept paths as the single-wave run."""
    ref = _anneal_whole()
    an = _anneal_whole(wave=wave, keep=keep)
    assert an.n_waves == -(-6 // wave) and an._B == wave
    for name in ("A_array", "me_array", "fe_array", "exitflags", "nit_array", "nfev_array", "params_array", "P"):
        assert np.array_equal(getattr(an, name), getattr(ref, name)), name
    if keep == 'all':
        assert np.array_equal(an.minpaths, ref.minpaths)
    elif keep == 'last':
        assert an.minpaths.shape == (6, 1, ref.minpaths.shape[-1])
        assert np.array_equal(an.minpaths[:, 0], ref.minpaths[:, -1])
    else:
        assert an.minpaths.shape[:2] == (6, 0)
    with pytest.raises(NotImplementedError):          # rung-by-rung driving needs every path resident
        an.anneal_step()


def test_lazy_initial_paths_and_memmap_sink(tmp_path):
    """X0 as a callable producing blocks of initial paths on demand (NumPy or CUDA tensors) and
    minpaths backed by a .npy memory map: same results as the in-memory run."""
    import torch
    from varanneal_b200 import va_ode
    D, N, B, Lidx, Y, t, X0, P0 = _problem()
    ref = _anneal_whole()
    for kind in ("numpy", "cuda"):
        an = va_ode.Annealer()
        an.wave_size = 4
        an.paths_file = str(tmp_path / ("paths_%s.npy" % kind))
        an.set_model("lorenz96", D)
        an.set_data(Y, t=t)
        fn = (lambda b0, b1: X0[b0:b1].copy()) if kind == "numpy" else (lambda b0, b1: torch.from_numpy(X0[b0:b1]).cuda())
        an.anneal(fn, P0.copy(), 1.5, [8, 12, 16], 4.0, 4e-6, Lidx, [0], **ARGS)
        assert np.array_equal(an.A_array, ref.A_array) and np.array_equal(an.params_array, ref.params_array)
        an.minpaths.flush()
        disk = np.load(an.paths_file)
        # rung 0 of the lazy run holds only minimisers (the initial paths were never parked on the host)
        assert np.array_equal(disk, ref.minpaths)
