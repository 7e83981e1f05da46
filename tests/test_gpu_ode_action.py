"""GPU parity of the fused ODE action+gradient (C ABI: vab_ode_action_grad, reached through
va_ode.Annealer.A_gradA) against the CPU oracle on the same seeded inputs.

Tolerance: 1e-10 relative on the action and on the gradient (max-norm), the figure
BASELINE.json's north_star states for fp64; the measured error is ~1e-14 (summation order only).
"""
import numpy as np
import pytest

from oracle.ode_port import OdeProblem

pytestmark = pytest.mark.gpu

TOL = 1e-10
NAKL_P = [120.0, 20.0, 0.3, 50.0, -77.0, -54.0, -40.0, 15.0, 0.1, 0.4, -60.0, -15.0, 1.0, 7.0,
          -55.0, 30.0, 1.0, 5.0]
DISCS = ["euler", "trapezoid", "SimpsonHermite", "forwardmap", "rk4"]


def _annealer(model, D, Y, t, stim, X0, P0, RM, RF0, Lidx, Pidx, disc, dt_model=None, beta=7,
              alpha=1.5):
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    an.set_model(model, D)
    an.set_data(Y, stim=stim, t=t)
    an.anneal_init(X0, P0, alpha, [beta], RM, RF0, Lidx, Pidx, dt_model=dt_model, disc=disc,
                   init_to_data=False)
    return an


def _check(model, D, Nd, nskip, disc, Lidx, P, Pidx, RM, RF0, B=2, stim=False, seed=0):
    rng = np.random.RandomState(seed)
    dt_model = 0.01
    N = (Nd - 1) * nskip + 1
    t = dt_model * nskip * np.arange(Nd)
    Y = rng.randn(Nd, len(Lidx))
    st = rng.randn(N) if stim else None
    if isinstance(RM, str):
        RM = rng.rand(Nd, len(Lidx)) + 0.5
    if isinstance(RF0, str):
        RF0 = (rng.rand(N - 1, D) + 0.5) * 1e-2
    P = np.array(P, dtype=np.float64)
    prob = OdeProblem(model, D, Y, Lidx, dt_model, disc, P, Pidx, RM, nskip=nskip, stim=st)
    XP = rng.randn(B, prob.n) * 2
    if model == "nakl":
        XP[:, :prob.nX] = 0.2 * rng.rand(B, prob.nX) + 0.4
        XP[:, 0:prob.nX:4] = -70 + 20 * rng.randn(B, N)
        XP[:, prob.nX:] = P[Pidx] * (1 + 0.01 * rng.randn(B, len(Pidx)))
    else:
        XP[:, prob.nX:] = P[Pidx] + 0.1 * rng.randn(B, len(Pidx))
    X0 = XP[:, :prob.nX].reshape(B, N, D).copy()
    P0 = np.tile(P, (B, 1))
    P0[:, Pidx] = XP[:, prob.nX:]
    an = _annealer(model, D, Y, t, st, X0, P0, RM, RF0, Lidx, Pidx, disc,
                   dt_model=None if nskip == 1 else dt_model)
    A, G = an.A_gradA(XP)
    scale = 1.5 ** 7
    for b in range(B):
        Ar, mer, fer, gr = prob.action_grad(XP[b], RF0 * scale, parts=True)
        assert abs(A[b] - Ar) <= TOL * abs(Ar)
        assert np.max(np.abs(G[b] - gr)) <= TOL * np.max(np.abs(gr))
    me = an._me.cpu().numpy(); fe = an._fe.cpu().numpy()
    assert abs(me[B - 1] - mer) <= TOL * max(abs(mer), 1e-300)
    assert abs(fe[B - 1] - fer) <= TOL * abs(fer)
    assert an.gpu_launches >= 2


@pytest.mark.parametrize("disc", DISCS)
def test_l96_d20_shipped_shape(disc):
    _check("lorenz96", 20, 161, 1, disc, [0, 2, 4, 6, 8, 10, 14, 16], [8.17], [0], 4.0, 4e-3, B=3)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_arrays_nskip_noparams(disc):
    _check("lorenz96", 10, 21, 2, disc, [1, 3, 9], [8.17], [], "arr", "arr")


@pytest.mark.parametrize("disc", DISCS)
def test_l96_odd_D(disc):
    _check("lorenz96", 7, 21, 3, disc, [0, 6], [8.17], [0], 2.0, 1e-2)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_d100_n1001_batch5(disc):
    _check("lorenz96", 100, 1001, 1, disc, [i for i in range(100) if i % 5 in (0, 2)], [8.17],
           [0], 4.0, 4e-3, B=5)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_d1000(disc):
    _check("lorenz96", 1000, 201, 1, disc, list(range(0, 1000, 7)), [8.17], [0], 2.0, 1e-2, B=2)


@pytest.mark.parametrize("disc", DISCS)
def test_l63(disc):
    _check("lorenz63", 3, 301, 1, disc, [0], [10, 28, 8 / 3], [0, 1, 2], 1.0, 0.3)
    _check("lorenz63", 3, 31, 2, disc, [0, 2], [10, 28, 8 / 3], [1], 1.0, 0.3, B=3)


@pytest.mark.parametrize("disc", DISCS[:4])
def test_nakl_stimulus(disc):
    _check("nakl", 4, 401, 1, disc, [0], NAKL_P, list(range(18)), 1.0,
           np.resize([1e-2, 1e2, 1e2, 1e2], (400, 4)), stim=True)
    _check("nakl", 4, 41, 1, disc, [0], NAKL_P, [0, 3, 17], 1.0, 0.5, stim=True)


def test_single_path_reference_shapes():
    """Non-batched call keeps the reference's return shapes (_autodiffmin.py:57-58)."""
    rng = np.random.RandomState(3)
    D, N = 20, 41
    Y = rng.randn(N, 3)
    X0 = rng.randn(N, D)
    an = _annealer("lorenz96", D, Y, 0.01 * np.arange(N), None, X0, np.array([8.17]), 4.0, 4e-3,
                   [0, 5, 7], [0], "trapezoid")
    XP = np.append(X0.ravel(), 8.17)
    A, g = an.A_gradA_taped(XP)
    assert isinstance(A, float) and g.shape == (N * D + 1,)
    prob = OdeProblem("lorenz96", D, Y, [0, 5, 7], 0.01, "trapezoid", [8.17], [0], 4.0)
    Ar, gr = prob.action_grad(XP, 4e-3 * 1.5 ** 7)
    assert abs(A - Ar) <= TOL * abs(Ar)
    assert np.max(np.abs(g - gr)) <= TOL * np.max(np.abs(gr))


def test_large_linearity_property_full_size():
    """BASELINE config-2 size (D=100, N=5001, Simpson): size-independent properties instead of
    the (slow) oracle: (i) the action is quadratic along X for fixed... no -- instead check the
    directional derivative: (A(x+h d) - A(x-h d)) / 2h == g.d to 1e-7, and determinism (two
    evaluations are bit-identical)."""
    rng = np.random.RandomState(11)
    D, N, B = 100, 5001, 4
    Lidx = [i for i in range(D) if i % 5 in (0, 2)]
    Y = rng.randn(N, len(Lidx))
    X0 = rng.randn(B, N, D)
    P0 = np.full((B, 1), 8.17)
    an = _annealer("lorenz96", D, Y, 0.025 * np.arange(N), None, X0, P0, 4.0, 4e-3, Lidx, [0],
                   "SimpsonHermite", beta=3)
    XP = np.concatenate([X0.reshape(B, -1), P0], axis=1)
    A1, G1 = an.A_gradA(XP)
    A2, G2 = an.A_gradA(XP)
    assert np.array_equal(A1, A2) and np.array_equal(G1, G2)
    d = rng.randn(*XP.shape)
    h = 1e-5
    Ap, _ = an.A_gradA(XP + h * d)
    Am, _ = an.A_gradA(XP - h * d)
    fd = (Ap - Am) / (2 * h)
    gd = np.sum(G1 * d, axis=1)
    assert np.all(np.abs(fd - gd) <= 1e-6 * np.abs(gd))


@pytest.mark.parametrize("disc", ["SimpsonHermite", "rk4"])
def test_headline_c2_shape_against_oracle(disc):
    """The exact launch bench.py times (BASELINE.json configs[1]: Lorenz96 D=100, N=5001, L=40,
    SimpsonHermite, RF = RF0 alpha**10, the twin data and initial paths of bench.py, init_to_data)
    compared with the oracle on two paths of a 64-path batch: A, me, fe and the gradient to
    1e-10.  rk4 rides along at the same shape."""
    import bench
    from varanneal_b200 import va_ode
    B = 64
    _, Y = bench.twin_data()
    X0, P0 = bench.initial_paths(B, 1000)
    an = va_ode.Annealer()
    an.set_model("lorenz96", bench.D)
    an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
    an.anneal_init(X0, P0, bench.ALPHA, [bench.BETA_EVAL], bench.RM, bench.RF0, bench.LIDX, [0], disc=disc,
                   init_to_data=True)
    XP = np.concatenate([X0.reshape(B, -1), P0], axis=1)          # X0 now carries the data (init_to_data)
    A, G = an.A_gradA(XP)
    me, fe = an._me.cpu().numpy(), an._fe.cpu().numpy()
    prob = OdeProblem("lorenz96", bench.D, Y, bench.LIDX, bench.DT, disc, [bench.K_FORCING], [0], bench.RM)
    for b in (0, B - 1):
        Ar, mer, fer, gr = prob.action_grad(XP[b], bench.RF0 * bench.ALPHA ** bench.BETA_EVAL, parts=True)
        assert abs(A[b] - Ar) <= TOL * abs(Ar)
        assert abs(fe[b] - fer) <= TOL * abs(fer) and abs(me[b] - mer) <= TOL * max(abs(mer), 1e-300)
        assert np.max(np.abs(G[b] - gr)) <= TOL * np.max(np.abs(gr))


@pytest.mark.parametrize("disc", DISCS)
def test_edge_shapes_minimal_and_unobserved(disc):
    """Smallest legal problems: N_model = 3 (one Simpson pair), a single observed component, and
    batch sizes that leave most lanes of the last warp idle."""
    _check("lorenz96", 8, 3, 1, disc, [5], [8.17], [0], 1.0, 0.5, B=1, seed=4)
    _check("lorenz96", 4, 5, 1, disc, [0, 1, 2, 3], [8.17], [], 1.0, 0.5, B=7, seed=5)
    _check("lorenz96", 12, 7, 2, disc, [11], [8.17], [0], 3.0, 0.1, B=2, seed=6)


@pytest.mark.parametrize("disc", ["trapezoid", "SimpsonHermite", "rk4"])
def test_l96_window_mode_wide_rows_long_paths(disc):
    """D = 1000 rows are cut into column windows with redundant halo lanes; long paths exercise
    64-bit indexing (n = 4e6 unknowns per path)."""
    rng = np.random.RandomState(9)
    D, N, B = 1000, 4001, 2
    Lidx = list(range(0, D, 3))
    Y = rng.randn(N, len(Lidx))
    X0 = rng.randn(B, N, D)
    P0 = np.full((B, 1), 8.17)
    an = _annealer("lorenz96", D, Y, 0.01 * np.arange(N), None, X0, P0, 2.0, 1e-2, Lidx, [0], disc, beta=2)
    XP = np.concatenate([X0.reshape(B, -1), P0], axis=1)
    A, G = an.A_gradA(XP)
    prob = OdeProblem("lorenz96", D, Y, Lidx, 0.01, disc, [8.17], [0], 2.0)
    Ar, gr = prob.action_grad(XP[1], 1e-2 * 1.5 ** 2)
    assert abs(A[1] - Ar) <= TOL * abs(Ar)
    assert np.max(np.abs(G[1] - gr)) <= TOL * np.max(np.abs(gr))


@pytest.mark.parametrize("B", [1, 5, 37])
def test_pinned_pipelined_eval_matches_plain_eval(B):
    """A_gradA(pinned host tensor) -- the end-to-end seam bench.py times: strided DMAs on two copy
    streams pipelined against the kernels of groups of paths (ramped group sizes) -- returns
    exactly what the plain NumPy-in / NumPy-out call returns, for batches that do and do not
    divide into the groups, with per-path fixed parameters."""
    import torch
    rng = np.random.RandomState(17)
    D, N = 20, 75
    Lidx = [1, 4, 9, 16]
    Y = rng.randn(N, len(Lidx))
    X0 = rng.randn(B, N, D)
    P0 = 8.0 + rng.rand(B, 1)
    an = _annealer("lorenz96", D, Y, 0.02 * np.arange(N), None, X0, P0, 4.0, 4e-3, Lidx, [], "SimpsonHermite")
    XP = X0.reshape(B, -1).copy()
    A0, G0 = an.A_gradA(XP)
    XP_pin = torch.from_numpy(XP).pin_memory()
    A1, G1 = an.A_gradA(XP_pin)
    assert np.array_equal(A1.numpy(), A0) and np.array_equal(G1.numpy(), G0)
    prob = OdeProblem("lorenz96", D, Y, Lidx, 0.02, "SimpsonHermite", P0[B - 1], [], 4.0)
    Ar, gr = prob.action_grad(XP[B - 1], 4e-3 * 1.5 ** 7)
    assert abs(A0[B - 1] - Ar) <= TOL * abs(Ar) and np.max(np.abs(G0[B - 1] - gr)) <= TOL * np.max(np.abs(gr))


def test_jacobian_and_hessian_seams():
    """ADmin.jacA_taped / A_jacaA_taped / hessianA_taped (_autodiffmin.py:60-67): the (1, n)
    Jacobian is the gradient; the dense Hessian (central differences of the device gradient)
    matches the complex-step Hessian of the oracle's gradient... here: finite differences of the
    oracle gradient at a tighter step, to 1e-6."""
    rng = np.random.RandomState(4)
    D, N = 4, 6
    Y = rng.randn(N, 2)
    X0 = rng.randn(N, D)
    an = _annealer("lorenz96", D, Y, 0.05 * np.arange(N), None, X0, np.array([8.0]), 4.0, 0.5, [0, 2], [0],
                   "trapezoid")
    XP = np.append(X0.ravel(), 8.0)
    A, g = an.A_gradA(XP)
    J = an.jacA_taped(XP)
    assert J.shape == (1, XP.size) and np.array_equal(J[0], g)
    A2, J2 = an.A_jacaA_taped(XP)
    assert A2 == A and np.array_equal(J2, J)
    H = an.hessianA_taped(XP)
    assert H.shape == (XP.size, XP.size) and np.array_equal(H, H.T)
    prob = OdeProblem("lorenz96", D, Y, [0, 2], 0.05, "trapezoid", [8.0], [0], 4.0)
    rf = 0.5 * 1.5 ** 7
    Href = np.empty_like(H)
    for j in range(XP.size):
        e = np.zeros(XP.size); e[j] = 1e-6 * (1 + abs(XP[j]))
        Href[:, j] = (prob.action_grad(XP + e, rf)[1] - prob.action_grad(XP - e, rf)[1]) / (2 * e[j])
    assert np.max(np.abs(H - 0.5 * (Href + Href.T))) <= 1e-6 * np.max(np.abs(Href))
