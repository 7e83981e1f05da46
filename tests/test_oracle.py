"""CPU tests of the oracle (the checker the GPU parity tests rely on).

1. Against the committed golden vectors (tests/golden/*.npz): values produced by the *reference's
   own* NumPy action code and gradients by complex-step differentiation through it
   (tests/golden/make_golden.py).  These run everywhere.
2. Against the verbatim reference loaded from /root/reference (marker `reference`; skipped on
   machines without the reference tree, e.g. the GPU box).
Tolerances: 1e-13 relative on values (same arithmetic, summation order differs only in the NN
port), 1e-12 on gradients (complex-step is exact to rounding).
"""
import numpy as np
import pytest

import golden_util
from oracle import nnet_port, ref_shim
from oracle.models_np import MODELS
from oracle.ode_port import OdeProblem, rf_ladder

ODE_CASES = golden_util.ode_cases()
NN_CASES = golden_util.nnet_cases()


def _problem(c):
    nskip = 1
    dt_data = c["t"][1] - c["t"][0]
    dt = dt_data
    if c["dt_model"] is not None:
        nskip = int(round(dt_data / c["dt_model"]))
        dt = c["dt_model"]
    return OdeProblem(c["model"], c["X0"].shape[1], c["Y"], c["Lidx"], dt, c["disc"], c["P0"], c["Pidx"],
                      c["RM"], nskip=nskip, stim=c["stim"])


def _rf(c, prob):
    RF0 = c["RF0"]
    if not np.isscalar(RF0):
        RF0 = np.resize(RF0, (prob.N - 1, prob.D))
    return RF0 * c["alpha"] ** c["beta"]


@pytest.mark.parametrize("c", ODE_CASES, ids=[c["name"] for c in ODE_CASES])
def test_ode_port_matches_reference_golden(c):
    prob = _problem(c)
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]]) if len(c["Pidx"]) else c["X0"].ravel()
    A, me, fe, g = prob.action_grad(XP, _rf(c, prob), parts=True)
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(me - c["A"][1]) <= 1e-13 * abs(c["A"][1])
    assert abs(fe - c["A"][2]) <= 1e-13 * abs(c["A"][2])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))


PTIME_CASES = golden_util.ptime_cases()
RM_CASES = golden_util.rm_matrix_cases()
RF_CASES = golden_util.rf_matrix_cases()


@pytest.mark.parametrize("c", RF_CASES, ids=[c["name"] for c in RF_CASES])
def test_ode_port_matrix_rf_matches_reference_golden(c):
    """RF0 of shape (D, D) / (N-1, D, D) with SimpsonHermite (va_ode.py:211-218)."""
    N, D = c["X0"].shape
    prob = _problem(c)
    RF = (c["RF0"] if c["RF0"].ndim == 3 else np.resize(c["RF0"], (N - 1, D, D))) * c["alpha"] ** c["beta"]
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]])
    A, me, fe, g = prob.action_grad(XP, RF, parts=True)
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(fe - c["A"][2]) <= 1e-13 * abs(c["A"][2])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))
    with pytest.raises(ValueError):
        OdeProblem(c["model"], D, c["Y"], c["Lidx"], 0.01, "trapezoid", c["P0"], c["Pidx"], 1.0).fe(XP, RF)


@pytest.mark.parametrize("c", RM_CASES, ids=[c["name"] for c in RM_CASES])
def test_ode_port_matrix_rm_matches_reference_golden(c):
    """RM of shape (L, L) / (N_data, L, L): sum_i diff_i . (RM_i diff_i) (va_ode.py:149-152)."""
    nd = c["Y"].shape[0]
    L = len(c["Lidx"])
    cc = dict(c, RM=c["RM"] if c["RM"].ndim == 3 else np.resize(c["RM"], (nd, L, L)))
    prob = _problem(cc)
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]])
    A, me, fe, g = prob.action_grad(XP, _rf(c, prob), parts=True)
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(me - c["A"][1]) <= 1e-13 * abs(c["A"][1])
    assert abs(fe - c["A"][2]) <= 1e-12 * abs(c["A"][2])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))


@pytest.mark.parametrize("c", PTIME_CASES, ids=[c["name"] for c in PTIME_CASES])
def test_ode_port_time_dependent_parameters_match_reference_golden(c):
    """P0 of shape (N_model, NP): XP = X.flatten() ++ P[:, Pidx].flatten() (va_ode.py:170-188)."""
    prob = _problem(c)
    assert prob.ptime and prob.n == c["grad"].size
    XP = np.append(c["X0"].ravel(), c["P0"][:, c["Pidx"]].ravel())
    A, me, fe, g = prob.action_grad(XP, _rf(c, prob), parts=True)
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(me - c["A"][1]) <= 1e-13 * abs(c["A"][1])
    assert abs(fe - c["A"][2]) <= 1e-13 * abs(c["A"][2])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))


NN_RM_CASES = golden_util.nnet_rm_matrix_cases()


@pytest.mark.parametrize("c", NN_RM_CASES, ids=[c["name"] for c in NN_RM_CASES])
def test_nnet_port_matrix_rm_matches_reference_golden(c):
    """va_nnet with RM = [RM_in, RM_out] (va_nnet.py:135-139)."""
    prob = nnet_port.NnetProblem(c["structure"], c["data_in"], c["data_out"], c["Lidx"], c["P0"], c["Pidx"], c["RM"])
    XP = np.append(c["X0"], c["P0"][c["Pidx"]])
    A, g = prob.action_grad(XP, c["RF0"] * c["alpha"] ** c["beta"])
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(prob.me(XP) - c["A"][1]) <= 1e-13 * abs(c["A"][1])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))


@pytest.mark.parametrize("c", NN_CASES, ids=[c["name"] for c in NN_CASES])
def test_nnet_port_matches_reference_golden(c):
    st = c["structure"]
    Lidx = [np.arange(st[0]), np.arange(st[-1])]
    prob = nnet_port.NnetProblem(st, c["data_in"], c["data_out"], Lidx, c["P0"], c["Pidx"], c["RM"])
    XP = np.append(c["X0"], c["P0"][c["Pidx"]])
    A, me, fe, g = prob.action_grad(XP, c["RF0"] * c["alpha"] ** c["beta"], parts=True)
    assert abs(A - c["A"][0]) <= 1e-13 * abs(c["A"][0])
    assert abs(me - c["A"][1]) <= 1e-13 * abs(c["A"][1])
    assert abs(fe - c["A"][2]) <= 1e-13 * abs(c["A"][2])
    assert np.max(np.abs(g - c["grad"])) <= 1e-12 * np.max(np.abs(c["grad"]))


def test_rk4_extension_gradient_by_complex_step():
    """rk4 does not exist in the reference (va_ode.py:382-402 is commented out): the port's
    adjoint is pinned by complex-step differentiation through the port's own value."""
    rng = np.random.RandomState(0)
    D, N = 8, 9
    Y = rng.randn(N, 3)
    prob = OdeProblem("lorenz96", D, Y, [0, 3, 5], 0.02, "rk4", [8.17], [0], 2.0)
    XP = np.append(rng.randn(N * D), 8.0)
    A, g = prob.action_grad(XP, 0.7)
    gc = ref_shim.complex_step_grad(lambda z: prob.action(z, 0.7), XP)
    assert np.max(np.abs(g - gc)) <= 1e-12 * np.max(np.abs(gc))
    for name, P in (("lorenz63", [10.0, 28.0, 8.0 / 3.0]),):
        prob = OdeProblem(name, 3, rng.randn(N, 1), [1], 0.01, "rk4", P, [0, 1, 2], 1.0)
        XP = np.append(rng.randn(N * 3), P)
        A, g = prob.action_grad(XP, 0.3)
        gc = ref_shim.complex_step_grad(lambda z: prob.action(z, 0.3), XP)
        assert np.max(np.abs(g - gc)) <= 1e-12 * np.max(np.abs(gc))


def test_lorenz63_port_gradient_by_complex_step():
    """Lorenz63 is an extension (not in the reference): pinned by complex step, all four
    reference discretisations."""
    rng = np.random.RandomState(1)
    N = 11
    for disc in ("euler", "trapezoid", "SimpsonHermite", "forwardmap"):
        prob = OdeProblem("lorenz63", 3, rng.randn(N, 2), [0, 2], 0.01, disc, [10.0, 28.0, 8.0 / 3.0], [1], 1.5)
        XP = np.append(rng.randn(N * 3), 27.0)
        A, g = prob.action_grad(XP, 0.2)
        gc = ref_shim.complex_step_grad(lambda z: prob.action(z, 0.2), XP)
        assert np.max(np.abs(g - gc)) <= 1e-12 * np.max(np.abs(gc))


def test_rf_ladder_uint16_truncation():
    rf, beta = rf_ladder(4e-6, 1.5, [0, 1.9, 3])
    assert list(beta) == [0, 1, 3]
    assert rf[1] == 4e-6 * 1.5 ** 1


# ------------------------------------------------------------------ against the live reference
@pytest.mark.reference
def test_port_vs_live_reference_random_problem():
    ShimOde, _ = ref_shim.make_shim_classes()
    rng = np.random.RandomState(42)
    D, N = 12, 15
    Lidx = [1, 4, 7, 10]
    t = 0.05 * np.arange(N)
    Y = rng.randn(N, len(Lidx))
    for disc in ("euler", "trapezoid", "SimpsonHermite", "forwardmap"):
        X0 = rng.randn(N, D)
        P0 = np.array([8.0])
        an = ShimOde()
        an.set_model(MODELS["lorenz96"], D)
        an.set_data(Y, t=t)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            an.anneal_init(X0.copy(), P0.copy(), 1.3, [9], 2.0, 1e-3, np.array(Lidx), [0], disc=disc,
                           init_to_data=False)
        XP = np.append(X0.ravel(), P0)
        prob = OdeProblem("lorenz96", D, Y, Lidx, 0.05, disc, P0, [0], 2.0)
        A, g = prob.action_grad(XP, an.RF)
        assert abs(A - an.A(XP)) <= 1e-13 * abs(A)
        gc = ref_shim.complex_step_grad(an.A, XP)
        assert np.max(np.abs(g - gc)) <= 1e-12 * np.max(np.abs(gc))
