"""Saved-array layouts of both annealers -- pure host code, exercised on CPU by
filling the result attributes by hand (the annealing itself needs a GPU and is covered by the
`gpu` tests; tests/test_gpu_ladder.py::test_saved_layouts_match_reference runs the ODE savers after
a real anneal).

Reference layouts: va_ode.py:794-889 (save_paths (Nbeta, N_model, 1+D) with t in column 0,
save_params (Nbeta, NP), save_action_errors (Nbeta, 5) = [beta, A, me, fe, fe/(RF0 alpha**beta)],
save_as_minAone rows [beta, exitflag, A, path...]); va_nnet.py:528-663 (save_io object array
(M, Nbeta, 2), save_Wb object arrays (Nbeta, N-1), save_states, save_params, save_action_errors;
the reference slices the parameters at NDnet instead of NDens, :596 -- fixed here)."""
import numpy as np
import pytest

from varanneal_b200 import va_nnet, va_ode


def _nn(B=None):
    an = va_nnet.Annealer()
    an.set_structure([3, 4, 2])
    an.M = 5
    an._sizes()
    Nbeta, NDens, NP = 3, 5 * 9, 3 * 4 + 4 + 4 * 2 + 2
    rng = np.random.RandomState(0)
    shape = (Nbeta,) if B is None else (B, Nbeta)
    an.batched = B is not None
    an.Nbeta, an.NP, an.NDens = Nbeta, NP, NDens
    an.minpaths = rng.randn(*(shape + (NDens + NP,)))
    an.A_array, an.me_array, an.fe_array = rng.rand(*shape), rng.rand(*shape), rng.rand(*shape)
    an.beta_array = np.array([0.0, 1.0, 2.0])
    an.alpha, an.RF0 = 1.1, 2.0
    return an, Nbeta, NDens, NP


def test_nnet_savers_layouts(tmp_path):
    an, Nbeta, NDens, NP = _nn()
    mp = an.minpaths
    an.save_states(str(tmp_path / "states.npy"))
    st = np.load(tmp_path / "states.npy")
    assert st.shape == (Nbeta, 5, 9) and np.array_equal(st[1, 2], mp[1, 2 * 9:3 * 9])
    an.save_io(str(tmp_path / "io.npy"))
    io = np.load(tmp_path / "io.npy", allow_pickle=True)
    assert io.shape == (5, Nbeta, 2)
    assert np.array_equal(io[4, 2, 0], mp[2, 4 * 9:4 * 9 + 3])            # input layer of example 4 at beta 2
    assert np.array_equal(io[4, 2, 1], mp[2, 4 * 9 + 7:4 * 9 + 9])        # output layer (last 2 of 9)
    an.save_Wb(str(tmp_path / "W.npy"), str(tmp_path / "b.npy"))
    W = np.load(tmp_path / "W.npy", allow_pickle=True)
    b = np.load(tmp_path / "b.npy", allow_pickle=True)
    assert W.shape == (Nbeta, 2) and b.shape == (Nbeta, 2)
    p = mp[1, NDens:]                                                     # parameters start at NDens, not NDnet
    assert W[1, 0].shape == (4, 3) and np.array_equal(W[1, 0], p[:12].reshape(4, 3))
    assert np.array_equal(b[1, 0], p[12:16])
    assert W[1, 1].shape == (2, 4) and np.array_equal(W[1, 1], p[16:24].reshape(2, 4))
    assert np.array_equal(b[1, 1], p[24:26])
    an.save_params(str(tmp_path / "params.npy"))
    assert np.array_equal(np.load(tmp_path / "params.npy"), mp[:, NDens:])
    an.save_action_errors(str(tmp_path / "ae.npy"))
    ae = np.load(tmp_path / "ae.npy")
    assert ae.shape == (Nbeta, 5) and np.array_equal(ae[:, 1], an.A_array)
    assert np.allclose(ae[:, 4], an.fe_array / (2.0 * 1.1 ** an.beta_array))
    an.save_action_errors(str(tmp_path / "ae.txt"))
    assert np.allclose(np.loadtxt(tmp_path / "ae.txt"), ae, rtol=1e-7)


def test_nnet_savers_batched_need_init(tmp_path):
    an, Nbeta, NDens, NP = _nn(B=2)
    with pytest.raises(ValueError, match="init"):
        an.save_params(str(tmp_path / "p.npy"))
    an.save_params(str(tmp_path / "p.npy"), init=1)
    assert np.array_equal(np.load(tmp_path / "p.npy"), an.minpaths[1, :, NDens:])
    an.save_states(str(tmp_path / "s.npy"), init=0)
    assert np.load(tmp_path / "s.npy").shape == (Nbeta, 5, 9)


def test_ode_savers_layouts_cpu(tmp_path):
    an = va_ode.Annealer()
    N, D, NP, Nbeta = 7, 4, 2, 3
    rng = np.random.RandomState(1)
    an.batched = False
    an.N_model, an.D, an.NP, an.Nbeta, an._nX = N, D, NP, Nbeta, N * D
    an.t_model = 0.1 * np.arange(N)
    an.minpaths = rng.randn(Nbeta, N * D + NP)
    an.A_array, an.me_array, an.fe_array = rng.rand(Nbeta), rng.rand(Nbeta), rng.rand(Nbeta)
    an.exitflags = np.array([0, 1, 0], dtype=np.int8)
    an.beta_array = np.array([0, 2, 5], dtype=np.uint16)
    an.alpha, an.RF0, an.L, an.adolcID = 1.5, np.full((N - 1, D), 3e-3), 2, 7
    an.save_paths(str(tmp_path / "paths.npy"))
    paths = np.load(tmp_path / "paths.npy")
    assert paths.shape == (Nbeta, N, 1 + D) and np.allclose(paths[:, :, 0], an.t_model)
    assert np.array_equal(paths[2, :, 1:].ravel(), an.minpaths[2, :N * D])
    an.save_paths(str(tmp_path / "paths.txt"))
    assert np.loadtxt(tmp_path / "paths.txt").shape == (Nbeta * N, 1 + D)
    an.save_params(str(tmp_path / "params.npy"))
    assert np.array_equal(np.load(tmp_path / "params.npy"), an.minpaths[:, N * D:])
    an.save_action_errors(str(tmp_path / "ae.npy"), cmpt=1)
    ae = np.load(tmp_path / "ae.npy")                  # array RF0: column 4 uses RF0[0, cmpt] (va_ode.py:866-870)
    assert np.allclose(ae[:, 4], an.fe_array / (3e-3 * 1.5 ** np.array([0.0, 2.0, 5.0])))
    an.save_as_minAone(str(tmp_path))
    m1 = np.loadtxt(tmp_path / "D4_M2_PATH7.dat")
    assert m1.shape == (Nbeta, 3 + N * D + NP) and np.allclose(m1[:, 0], [0, 2, 5]) and np.allclose(m1[:, 1], [0, 1, 0])
    assert np.allclose(m1[:, 2], an.A_array) and np.allclose(m1[:, 3:], an.minpaths)


def test_pipeline_spans_cover_the_batch_with_short_ends():
    """Groups of paths of the pipelined host-to-host evaluation (DeviceMin._pipeline_spans): a
    partition of [0, B) in order, with small groups at both ends (the first upload and the last
    download overlap nothing) for batches large enough to ramp."""
    from varanneal_b200._devicemin import DeviceMin
    for B in (1, 2, 3, 7, 8, 16, 31, 64, 100, 257):
        spans = DeviceMin._pipeline_spans(B)
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(hi > lo for lo, hi in spans)
    sizes = [hi - lo for lo, hi in DeviceMin._pipeline_spans(64)]
    assert sizes[0] == 1 and sizes[-1] == 1 and max(sizes) == 8 and sizes[:3] == [1, 2, 4]
