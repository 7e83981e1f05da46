"""GPU parity of the neural-network action+gradient (C ABI: vab_nn_action_grad through
va_nnet.Annealer) against the reference-generated golden vectors (tests/golden/
nnet_action_golden.npz: value by the reference's own loops, gradient by complex step through
them) and against the NumPy oracle on larger seeded cases; plus a short device ladder against
SciPy L-BFGS-B on the oracle action.  Tolerance 1e-10 relative."""
import numpy as np
import pytest
import scipy.optimize as opt

import golden_util
from oracle import nnet_port

pytestmark = pytest.mark.gpu
TOL = 1e-10
NN_CASES = golden_util.nnet_cases()


def _annealer(structure, data_in, data_out, X0, P0, alpha, beta, RM, RF0, Pidx, act="sigmoid", opt_args=None,
              Lidx=None):
    from varanneal_b200 import va_nnet
    an = va_nnet.Annealer()
    an.set_structure(structure)
    an.set_activation(act)
    an.set_input_data(data_in)
    an.set_output_data(data_out)
    an.anneal_init(X0, P0, alpha, beta, RM, RF0, Pidx, Lidx=Lidx, init_to_data=False, opt_args=opt_args)
    return an


@pytest.mark.parametrize("c", NN_CASES, ids=[c["name"] for c in NN_CASES])
def test_nn_action_grad_vs_reference_golden(c):
    an = _annealer(c["structure"], c["data_in"], c["data_out"], c["X0"].copy(), c["P0"].copy(), c["alpha"],
                   [c["beta"]], c["RM"], c["RF0"], c["Pidx"])
    XP = np.append(c["X0"], c["P0"][c["Pidx"]])
    A, g = an.A_gradA_taped(XP)
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.me_gaussian(XP) - c["A"][1]) <= TOL * abs(c["A"][1])
    assert abs(an.fe_gaussian(XP) - c["A"][2]) <= TOL * abs(c["A"][2])


NN_RM_CASES = golden_util.nnet_rm_matrix_cases()


@pytest.mark.parametrize("c", NN_RM_CASES, ids=[c["name"] for c in NN_RM_CASES])
@pytest.mark.parametrize("flags", ["", "F", "S"])
def test_nn_matrix_rm_vs_reference_golden(c, flags, monkeypatch):
    """RM = [RM_in, RM_out] (va_nnet.py:135-139), not symmetric, through every kernel family (split
    kernels, fused kernel, register-tiled small-network kernel) and for a batch with unequal paths."""
    if "F" in flags:
        monkeypatch.setenv("VAB_NN_SPLIT", "0")
    if "S" in flags:
        monkeypatch.setenv("VAB_NN_SMALL", "1")
    an = _annealer(c["structure"], c["data_in"], c["data_out"], c["X0"].copy(), c["P0"].copy(), c["alpha"],
                   [c["beta"]], c["RM"].copy(), c["RF0"], c["Pidx"], Lidx=c["Lidx"])
    XP = np.append(c["X0"], c["P0"][c["Pidx"]])
    A, g = an.A_gradA_taped(XP)
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.me_gaussian(XP) - c["A"][1]) <= TOL * abs(c["A"][1])
    # a batch of three different paths, differently sized matrices on the two sides (extension)
    rng = np.random.default_rng(2)
    B = 3
    X0 = np.tile(c["X0"], (B, 1)) + 0.1 * rng.standard_normal((B, c["X0"].size))
    P0 = np.tile(c["P0"], (B, 1)) + 0.05 * rng.standard_normal((B, c["P0"].size))
    Lidx = [c["Lidx"][0], c["Lidx"][1][:-1]]
    RM = [c["RM"][0], c["RM"][1][:-1, :-1]]
    anb = _annealer(c["structure"], c["data_in"], c["data_out"][:, :-1], X0.copy(), P0.copy(), c["alpha"], [c["beta"]], RM,
                    c["RF0"], c["Pidx"], Lidx=Lidx)
    XPb = np.concatenate([X0, P0[:, c["Pidx"]]], axis=1)
    Ab, gb = anb.A_gradA_taped(XPb)
    for b in range(B):
        prob = nnet_port.NnetProblem(c["structure"], c["data_in"], c["data_out"][:, :-1], Lidx, P0[b], c["Pidx"], RM)
        Ao, go = prob.action_grad(XPb[b], c["RF0"] * c["alpha"] ** c["beta"])
        assert abs(Ab[b] - Ao) <= TOL * abs(Ao)
        assert np.max(np.abs(gb[b] - go)) <= TOL * np.max(np.abs(go))


@pytest.mark.parametrize("structure,M,act,RM,partial,B", [
    ([25, 30, 4], 1000, "sigmoid", 1.0, False, 2),          # bar-images shape, many example tiles
    ([100, 100, 100, 100], 130, "sigmoid", [3.0, 0.5], False, 2),   # nnet_twin widths ~100, RM (2,)
    ([300, 17, 5], 7, "tanh", 2.0, True, 3),                # wide first layer -> chunked weights
    ([6, 9, 9, 3], 40, "linear", 1.5, True, 1),
])
def test_nn_action_grad_vs_oracle(structure, M, act, RM, partial, B):
    rng = np.random.RandomState(3)
    st = np.array(structure)
    NDnet = int(st.sum())
    NP = int(sum(st[n] * st[n + 1] + st[n + 1] for n in range(len(st) - 1)))
    Lidx = [np.arange(st[0])[::2], np.arange(st[-1])]
    data_in = rng.rand(M, len(Lidx[0]))
    data_out = rng.rand(M, len(Lidx[1]))
    X0 = rng.rand(B, M * NDnet)
    P0 = 0.3 * rng.randn(B, NP)
    Pidx = np.arange(0, NP, 3) if partial else np.arange(NP)
    alpha, beta, RF0 = 1.1, 30.0, 1e-2
    an = _annealer(st, data_in, data_out, X0.copy(), P0.copy(), alpha, [beta], RM, RF0, Pidx, act=act, Lidx=Lidx)
    XP = np.concatenate([X0, P0[:, Pidx]], axis=1)
    A, G = an.A_gradA(XP)
    for b in range(B):
        prob = nnet_port.NnetProblem(st, data_in, data_out, Lidx, P0[b], Pidx, RM, act=act)
        Ar, gr = prob.action_grad(XP[b], RF0 * alpha ** beta)
        assert abs(A[b] - Ar) <= TOL * abs(Ar)
        assert np.max(np.abs(G[b] - gr)) <= TOL * np.max(np.abs(gr))


@pytest.mark.parametrize("tcgen05", [False, True], ids=["dmma", "tcgen05"])
def test_nn_ladder_vs_scipy(tcgen05, monkeypatch):
    """Short annealing ladder over the neuron states of a small twin network with the weights held
    at (perturbed) teacher values, at RF values where the model error shapes the minimum.  (With
    the weights free, or at small RF, the net fits anything: A -> 0 along a flat valley, SciPy
    itself runs into maxfun = 2e5 without converging, and two correct optimisers stop anywhere on
    it -- measured while writing this test.)  Per-beta minimum action vs SciPy L-BFGS-B on the
    oracle action to 1e-6 relative; and SciPy started at the device's minimiser has nothing left
    to do."""
    if tcgen05:                            # the same ladder with every contraction on tcgen05 (graph-replayed cycles)
        monkeypatch.setenv("VAB_NN_TCGEN05", "1")
    rng = np.random.RandomState(8)
    st = np.array([4, 6, 3])
    M = 12
    NDnet, NP = int(st.sum()), int(4 * 6 + 6 + 6 * 3 + 3)
    Wt = [rng.randn(6, 4), rng.randn(3, 6)]
    xin = rng.rand(M, 4)
    h = 1 / (1 + np.exp(-(xin @ Wt[0].T)))
    yout = 1 / (1 + np.exp(-(h @ Wt[1].T))) + 0.05 * rng.randn(M, 3)
    X0 = rng.rand(M * NDnet)
    P0 = np.concatenate([Wt[0].ravel(), np.zeros(6), Wt[1].ravel(), np.zeros(3)]) * (1 + 0.05 * rng.randn(NP))
    Pidx = np.arange(0)
    alpha, betas, RM, RF0 = 2.0, np.arange(0.0, 12.0, 2.0), 400.0, 1.0
    opts = {"gtol": 1e-11, "ftol": 1e-15, "maxfun": 200000, "maxiter": 200000}
    from varanneal_b200 import va_nnet
    an = va_nnet.Annealer()
    an.set_structure(st); an.set_activation(va_nnet.sigmoid)
    an.set_input_data(xin); an.set_output_data(yout)
    X0d = X0.copy()
    an.anneal(X0d, P0.copy(), alpha, betas, RM, RF0, Pidx, opt_args=opts)
    prob = nnet_port.NnetProblem(st, xin, yout, None, P0, Pidx, RM)
    xp = X0d.copy()                        # X0d carries the init_to_data overwrite
    for i, beta in enumerate(betas):
        rf = RF0 * alpha ** beta
        res = opt.minimize(lambda z: prob.action_grad(z, rf), xp, method="L-BFGS-B", jac=True, options=opts)
        xp = res.x
        print("beta %4.1f scipy %.10e (nit %d) device %.10e (nit %d)" % (beta, res.fun, res.nit, an.A_array[i], an.nit_array[i]))
        assert abs(an.A_array[i] - res.fun) <= 1e-6 * abs(res.fun), (i, an.A_array[i], res.fun)
        # secondary acceptance: SciPy, started at the device's minimiser, has nothing left to do
        res2 = opt.minimize(lambda z: prob.action_grad(z, rf), an.minpaths[i][:M * NDnet], method="L-BFGS-B", jac=True,
                            options={"gtol": 1e-9, "ftol": 1e-13})
        assert res2.nit <= 2 and abs(res2.fun - an.A_array[i]) <= 1e-9 * abs(res2.fun)
    assert an.minpaths.shape == (len(betas), M * NDnet + NP)
    if tcgen05:
        # the split contractions carry rounding noise of ~1e-14 of the row maxima: with ftol = 1e-15 the
        # last line search of a rung may find no decrease left and end "abnormally" (status 2) *at* the
        # minimum -- the two checks above hold on every rung either way
        assert np.all(np.isin(an.exitflags, (0, 2))), an.exitflags
    else:
        assert np.all(an.exitflags == 0)
    assert an._ctx.nn_kernel_family == (5 if tcgen05 else 3)


@pytest.mark.parametrize("structure,M,B", [([25, 30, 4], 333, 2), ([100, 100, 100], 129, 2), ([12, 40], 70, 3)])
def test_nn_split_kernels_agree_with_fused_kernel(structure, M, B, monkeypatch):
    """The split design (nn_fb_kernel / nn_fix_kernel / nn_gw_kernel) and the fused example-tile
    kernel are two decompositions of the same sums: value and gradient agree to rounding (and both
    meet the oracle, see the tests above).  M is not a multiple of any tile size; the last case
    has no middle layer (nn_fix_kernel is skipped)."""
    rng = np.random.RandomState(11)
    st = np.array(structure)
    NDnet = int(st.sum())
    NP = int(sum(st[n] * st[n + 1] + st[n + 1] for n in range(len(st) - 1)))
    Lidx = [np.arange(st[0])[1::2], np.arange(st[-1])[::2]]
    data_in, data_out = rng.rand(M, len(Lidx[0])), rng.rand(M, len(Lidx[1]))
    X0 = rng.rand(B, M * NDnet)
    P0 = 0.3 * rng.randn(B, NP)
    Pidx = np.arange(1, NP, 2)
    XP = np.concatenate([X0, P0[:, Pidx]], axis=1)
    res = {}
    # "1": split design (small networks: the all-layers tile kernel), "L": split design with the
    # per-layer kernels forced, "0": fused example-tile kernel, "S": the register-tiled CUDA-core
    # kernel for narrow networks (nn_small.cuh; the default where it applies, i.e. widths <= 64)
    for flag, split, alll, small in (("1", "1", "1", "0"), ("L", "1", "0", "0"), ("0", "0", "1", "0"), ("S", "1", "1", "1")):
        monkeypatch.setenv("VAB_NN_SPLIT", split)
        monkeypatch.setenv("VAB_NN_ALL_LAYERS", alll)
        monkeypatch.setenv("VAB_NN_SMALL", small)
        an = _annealer(st, data_in, data_out, X0.copy(), P0.copy(), 1.1, [25.0], [2.0, 0.7], 1e-2, Pidx, Lidx=Lidx)
        l0 = an.gpu_launches
        A, G = an.A_gradA(XP)
        res[flag] = (A.copy(), G.copy(), an.gpu_launches - l0, an.me_gaussian(XP), an.fe_gaussian(XP))
    assert res["1"][2] != res["0"][2]                    # really two different launch sequences
    for k in ("1", "L", "S"):
        assert np.max(np.abs(res[k][0] - res["0"][0]) / np.abs(res["0"][0])) <= 1e-13
        assert np.max(np.abs(res[k][1] - res["0"][1])) <= 1e-12 * np.max(np.abs(res["0"][1]))
        assert np.allclose(res[k][3], res["0"][3], rtol=1e-13) and np.allclose(res[k][4], res["0"][4], rtol=1e-13)


def test_nn_tnc_method_minimises_the_oracle_action():
    """va_nnet with method='TNC' (min_tnc_scipy): the device truncated Newton reaches the minimum
    SciPy's TNC reaches on the oracle action (1e-6 relative) on the well-conditioned twin problem
    of the ladder test above (weights held at perturbed teacher values, RF large enough for the
    model error to shape the minimum; with free weights the valley is flat and SciPy's own
    L-BFGS-B runs 1e5 iterations without converging)."""
    rng = np.random.RandomState(8)
    st = np.array([4, 6, 3])
    M = 12
    NDnet, NP = int(st.sum()), int(4 * 6 + 6 + 6 * 3 + 3)
    Wt = [rng.randn(6, 4), rng.randn(3, 6)]
    xin = rng.rand(M, 4)
    h = 1 / (1 + np.exp(-(xin @ Wt[0].T)))
    yout = 1 / (1 + np.exp(-(h @ Wt[1].T))) + 0.05 * rng.randn(M, 3)
    X0 = rng.rand(M * NDnet)
    P0 = np.concatenate([Wt[0].ravel(), np.zeros(6), Wt[1].ravel(), np.zeros(3)]) * (1 + 0.05 * rng.randn(NP))
    Pidx = np.arange(0)
    alpha, betas, RM, RF0 = 2.0, [4.0, 8.0], 400.0, 1.0
    from varanneal_b200 import va_nnet
    an = va_nnet.Annealer()
    an.set_structure(st); an.set_activation(va_nnet.sigmoid)
    an.set_input_data(xin); an.set_output_data(yout)
    X0d = X0.copy()
    an.anneal(X0d, P0.copy(), alpha, betas, RM, RF0, Pidx, method="TNC", opt_args={"gtol": 1e-9, "maxfun": 200000})
    assert np.all(an.exitflags <= 2), an.exitflags           # local minimum / f converged / x converged
    prob = nnet_port.NnetProblem(st, xin, yout, None, P0, Pidx, RM)
    xp = X0d.copy()
    for i, beta in enumerate(betas):
        rf = RF0 * alpha ** beta
        res = opt.minimize(lambda z: prob.action_grad(z, rf), xp, method="TNC", jac=True,
                           options={"gtol": 1e-10, "maxfun": 200000})
        xp = res.x
        assert abs(an.A_array[i] - res.fun) <= 1e-6 * abs(res.fun), (i, an.A_array[i], res.fun)
        A, g = prob.action_grad(an.minpaths[i][:M * NDnet], rf)
        assert abs(A - an.A_array[i]) <= 1e-10 * abs(A) and np.max(np.abs(g)) <= 1e-5 * max(1.0, abs(A))


# ---------------------------------------------------------------------------------------------
# tcgen05 / TMEM / TMA path (VAB_NN_TCGEN05=1): the forward and the backward-to-states contractions of
# every layer as Ozaki-split int8 GEMMs on the 5th-generation tensor cores (csrc/ozaki_gemm.cu), the
# same 1e-10 bar as the fp64 tensor-pipe kernels.
@pytest.mark.parametrize("c", [c for c in NN_CASES if max(c["structure"]) <= 128], ids=lambda c: c["name"])
def test_nn_tcgen05_path_vs_reference_golden(c, monkeypatch):
    monkeypatch.setenv("VAB_NN_TCGEN05", "1")
    an = _annealer(c["structure"], c["data_in"], c["data_out"], c["X0"].copy(), c["P0"].copy(), c["alpha"],
                   [c["beta"]], c["RM"], c["RF0"], c["Pidx"])
    XP = np.append(c["X0"], c["P0"][c["Pidx"]])
    A, g = an.A_gradA_taped(XP)
    assert an._ctx.nn_kernel_family == 5, "the tcgen05 kernels did not run"
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.fe_gaussian(XP) - c["A"][2]) <= TOL * abs(c["A"][2])


@pytest.mark.parametrize("structure,M,act,RM,partial,B", [
    ([25, 30, 4], 1000, "sigmoid", 1.0, False, 2),                  # bar-images shape
    ([100, 100, 100, 100], 130, "sigmoid", [3.0, 0.5], False, 2),   # nnet_twin widths
    ([100, 100, 100, 100, 100], 1000, "sigmoid", 1.0, False, 3),    # C4: 5 x 100, M = 1000
    ([128, 17, 5], 300, "tanh", 2.0, True, 3),                      # the widest layer the planes hold
    ([6, 9, 9, 3], 40, "linear", 1.5, True, 1),
])
def test_nn_tcgen05_path_vs_oracle(structure, M, act, RM, partial, B, monkeypatch):
    monkeypatch.setenv("VAB_NN_TCGEN05", "1")
    rng = np.random.RandomState(3)
    st = np.array(structure)
    NDnet = int(st.sum())
    NP = int(sum(st[n] * st[n + 1] + st[n + 1] for n in range(len(st) - 1)))
    Lidx = [np.arange(st[0])[::2], np.arange(st[-1])]
    data_in = rng.rand(M, len(Lidx[0]))
    data_out = rng.rand(M, len(Lidx[1]))
    # states and weights with three decades of spread inside a row: the digit planes are scaled by
    # the row maximum, so small entries next to large ones are what they resolve least well
    X0 = rng.rand(B, M * NDnet) * 10.0 ** rng.uniform(-3, 0, size=(B, M * NDnet))
    P0 = 0.3 * rng.randn(B, NP) * 10.0 ** rng.uniform(-3, 0, size=(B, NP))
    Pidx = np.arange(0, NP, 3) if partial else np.arange(NP)
    alpha, beta, RF0 = 1.1, 30.0, 1e-2
    an = _annealer(st, data_in, data_out, X0.copy(), P0.copy(), alpha, [beta], RM, RF0, Pidx, act=act, Lidx=Lidx)
    XP = np.concatenate([X0, P0[:, Pidx]], axis=1)
    A, G = an.A_gradA(XP)
    assert an._ctx.nn_kernel_family == 5, "the tcgen05 kernels did not run"
    for b in range(B):
        prob = nnet_port.NnetProblem(st, data_in, data_out, Lidx, P0[b], Pidx, RM, act=act)
        Ar, gr = prob.action_grad(XP[b], RF0 * alpha ** beta)
        assert abs(A[b] - Ar) <= TOL * abs(Ar)
        assert np.max(np.abs(G[b] - gr)) <= TOL * np.max(np.abs(gr))
    # and the default (fp64 tensor-pipe) kernels on the same input agree with it to the same bar
    monkeypatch.delenv("VAB_NN_TCGEN05")
    an2 = _annealer(st, data_in, data_out, X0.copy(), P0.copy(), alpha, [beta], RM, RF0, Pidx, act=act, Lidx=Lidx)
    A2, G2 = an2.A_gradA(XP)
    assert an2._ctx.nn_kernel_family in (2, 3)
    assert np.max(np.abs(A - A2) / np.abs(A2)) <= TOL
    assert np.max(np.abs(G - G2)) <= TOL * np.max(np.abs(G2))
