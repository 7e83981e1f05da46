"""GPU parity of the device's bounded L-BFGS-B (generalised Cauchy point + subspace minimisation,
csrc/lbfgsb_bounded.cuh) against the reference's minimiser -- SciPy L-BFGS-B with ``bounds``
(_autodiffmin.py:85-86; bounds expansion va_ode.py:582-605) -- driving the oracle action, and
against the restatement oracle/lbfgsb_port.py that specifies it.

On well-conditioned problems the device takes the *same number of iterations and evaluations* as
SciPy (identical algorithmic decisions: which bounds the Cauchy search fixes, the subspace step,
the projection, every line-search trial); minima agree to 1e-6 relative (north-star tolerance)
-- measured far below.
"""
import numpy as np
import pytest
import scipy.optimize as opt

import golden_util
from oracle import lbfgsb_port
from oracle.ode_port import OdeProblem

pytestmark = pytest.mark.gpu


def _l96(B, bounds, seed=5, D=10, N=41, opts=None, beta=6):
    from varanneal_b200 import va_ode
    rng = np.random.RandomState(seed)
    Lidx = list(range(0, D, 2))
    t = 0.02 * np.arange(N)
    Y = 3.0 * rng.randn(N, len(Lidx))
    X0 = 2.0 * rng.randn(B, N, D)
    P0 = np.tile([8.0], (B, 1))
    an = va_ode.Annealer()
    an.set_model("lorenz96", D)
    an.set_data(Y, t=t)
    opts = opts or {"gtol": 1e-9, "ftol": 1e-14, "maxiter": 5000}
    an.anneal(X0.copy(), P0.copy(), 2.0, [beta], 1.0, 1e-2, Lidx, [0], disc="trapezoid", bounds=bounds,
              init_to_data=False, opt_args=opts)
    prob = OdeProblem("lorenz96", D, Y, Lidx, 0.02, "trapezoid", [8.0], [0], 1.0)
    return an, prob, X0, P0, 1e-2 * 2.0 ** beta, opts


@pytest.mark.parametrize("box", ["two-sided", "lower-only", "upper-only+param"])
def test_bounded_minimise_follows_scipy(box):
    D, N, B = 10, 41, 3
    if box == "two-sided":
        bounds = [[-1.5, 1.5]] * D + [[7.0, 9.0]]
    elif box == "lower-only":
        bounds = [[-0.5, None]] * D + [[None, None]]
    else:
        bounds = [[None, 1.0]] * D + [[8.5, 12.0]]             # the parameter's lower bound is active
    an, prob, X0, P0, rf, opts = _l96(B, bounds)
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds])
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds])
    lo = np.concatenate([np.tile(lo[:D], N), lo[D:]])
    hi = np.concatenate([np.tile(hi[:D], N), hi[D:]])
    sb = list(zip(np.where(np.isfinite(lo), lo, None), np.where(np.isfinite(hi), hi, None)))
    assert np.all(an.exitflags == 0)
    for b in range(B):
        xp0 = np.append(X0[b].ravel(), P0[b])
        r = opt.minimize(lambda z: prob.action_grad(z, rf), xp0, method="L-BFGS-B", jac=True, bounds=sb,
                         options={"gtol": opts["gtol"], "ftol": opts["ftol"], "maxiter": 5000})
        xmin = an.minpaths[b, 0]
        assert np.all(xmin >= lo) and np.all(xmin <= hi)                    # exactly inside the box
        assert abs(an.A_array[b, 0] - r.fun) <= 1e-6 * abs(r.fun), (an.A_array[b, 0], r.fun)
        # (both runs stop on the relative-reduction test in a shallow valley: the minimisers agree
        # to ~1e-4 while A agrees to 1e-6; what is checked instead is that the device point is a
        # constrained stationary point of the oracle's action)
        assert np.max(np.abs(xmin - r.x)) <= 2e-3
        Ao, go = prob.action_grad(xmin, rf)
        pg = np.where(go < 0, np.maximum(xmin - hi, go), np.minimum(xmin - lo, go))
        assert abs(Ao - an.A_array[b, 0]) <= 1e-10 * abs(Ao) and np.max(np.abs(pg)) <= 1e-5
        nact_dev = int(np.sum(xmin <= lo) + np.sum(xmin >= hi))
        nact_ref = int(np.sum(r.x <= lo) + np.sum(r.x >= hi))
        assert nact_dev == nact_ref and nact_ref > 0                        # bounds really are active
        # same algorithm, same decisions: iteration / evaluation counts within a few of SciPy's
        print("bounded[%s] path %d: device nit %d nfev %d | scipy nit %d nfev %d | active %d" % (
            box, b, an.nit_array[b, 0], an.nfev_array[b, 0], r.nit, r.nfev, nact_ref))
        assert abs(int(an.nit_array[b, 0]) - r.nit) <= max(5, r.nit // 4), (an.nit_array[b, 0], r.nit)
        assert abs(int(an.nfev_array[b, 0]) - r.nfev) <= max(5, r.nfev // 4), (an.nfev_array[b, 0], r.nfev)
        if b == 0:                                                          # ... and the restatement's
            q = lbfgsb_port.minimize(lambda z: prob.action_grad(z, rf), xp0, lo, hi, ftol=opts["ftol"],
                                     gtol=opts["gtol"], maxiter=5000, maxfun=15000)
            assert abs(q["fun"] - an.A_array[b, 0]) <= 1e-8 * abs(q["fun"])


def test_bounded_first_iterations_match_port_exactly():
    """A short horizon (5 iterations) from a start with most state variables outside the box: the
    device's iterate after maxiter iterations equals the restatement's to rounding, i.e. the
    Cauchy search crossed the same breakpoints and the subspace step / projection agree."""
    D, N = 10, 41
    bounds = [[-1.0, 1.0]] * D + [[7.5, 8.5]]
    for maxiter in (1, 2, 5):
        an, prob, X0, P0, rf, opts = _l96(1, bounds, seed=11, opts={"gtol": 1e-12, "ftol": 1e-16, "maxiter": maxiter})
        lo = np.concatenate([np.tile([-1.0] * D, N), [7.5]])
        hi = np.concatenate([np.tile([1.0] * D, N), [8.5]])
        xp0 = np.append(X0[0].ravel(), P0[0])
        q = lbfgsb_port.minimize(lambda z: prob.action_grad(z, rf), xp0, lo, hi, ftol=1e-16, gtol=1e-12,
                                 maxiter=maxiter, maxfun=15000)
        assert int(an.nit_array[0, 0]) == q["nit"] == maxiter and int(an.nfev_array[0, 0]) == q["nfev"]
        assert np.max(np.abs(an.minpaths[0, 0] - q["x"])) <= 1e-10, maxiter
        assert abs(an.A_array[0, 0] - q["fun"]) <= 1e-12 * abs(q["fun"])


def test_bounds_never_active_equals_unbounded_run():
    """A box far away from every iterate: L-BFGS-B's bounded code path must reproduce the
    unbounded minimisation (same minimum; counts within a few -- the first step differs: 1/||d||
    is only used when the problem is not boxed, lnsrlb)."""
    D = 10
    an_b, prob, X0, P0, rf, opts = _l96(2, [[-1e3, 1e3]] * D + [[-1e3, 1e3]])
    an_u, _, _, _, _, _ = _l96(2, None)
    assert np.all(an_b.exitflags == 0) and np.all(an_u.exitflags == 0)
    assert np.max(np.abs(an_b.A_array - an_u.A_array) / np.abs(an_u.A_array)) <= 1e-9
    assert np.max(np.abs(an_b.minpaths - an_u.minpaths)) <= 2e-3      # flat valley: A to 1e-9, x to ~1e-3


def test_nakl_tutorial_box_single_rung_vs_scipy():
    """The tutorial's bounded NaKL problem (ipynb:3069-3087, 3139-3141), one rung, 18 parameters and
    all states boxed, trapezoid and SimpsonHermite: device vs SciPy on the oracle action."""
    from varanneal_b200 import va_ode
    c = [c for c in golden_util.ode_cases() if c["name"] == "nakl_trapezoid_18p"][0]
    Pb = [[60.0, 180.0], [10.0, 30.0], [0.15, 0.45], [47.5, 52.5], [-80.85, -73.15], [-56.7, -51.3],
          [-42.0, -38.0], [14.25, 15.75], [0.095, 0.105], [0.38, 0.42], [-63.0, -57.0], [-15.75, -14.25],
          [0.95, 1.05], [6.65, 7.35], [-57.75, -52.25], [28.5, 31.5], [0.95, 1.05], [4.75, 5.25]]
    Pb[0] = [60.0, 110.0]
    Pb[1] = [22.0, 30.0]
    bounds = [[-100.0, 100.0], [0.0, 1.0], [0.0, 1.0], [0.0, 1.0]] + Pb
    N = c["X0"].shape[0]
    rng = np.random.RandomState(8)
    P0 = np.array([(b[1] - b[0]) * rng.rand() + b[0] for b in Pb])
    RF0 = [1e-8, 1e-4, 1e-4, 1e-4]
    lo = np.concatenate([np.tile(np.array(bounds)[:4, 0], N), np.array(Pb)[:, 0]])
    hi = np.concatenate([np.tile(np.array(bounds)[:4, 1], N), np.array(Pb)[:, 1]])
    for disc, beta in (("trapezoid", 150), ("SimpsonHermite", 170)):
        opts = {"gtol": 1e-11, "ftol": 1e-13, "maxiter": 200000, "maxfun": 400000}
        an = va_ode.Annealer()
        an.set_model("nakl", 4)
        an.set_data(c["Y"], stim=c["stim"], t=c["t"])
        X0 = c["X0"].copy()
        an.anneal(X0, P0.copy(), 1.1, [beta], 1.0, RF0, [0], list(range(18)), disc=disc, bounds=bounds,
                  opt_args=opts)
        prob = OdeProblem("nakl", 4, c["Y"], [0], an.dt_model, disc, P0, list(range(18)), 1.0, stim=c["stim"])
        rf = np.resize(np.asarray(RF0), (N - 1, 4)) * 1.1 ** beta
        xp0 = np.append(X0.ravel(), P0)                       # X0 carries the data now (init_to_data)
        r = opt.minimize(lambda z: prob.action_grad(z, rf), xp0, method="L-BFGS-B", jac=True,
                         bounds=list(zip(lo, hi)), options=opts)
        xmin = an.minpaths[0]
        assert np.all(xmin >= lo) and np.all(xmin <= hi)
        assert an.exitflags[0] == 0 and r.status == 0
        A, g = prob.action_grad(xmin, rf)
        assert abs(A - an.A_array[0]) <= 1e-10 * abs(A)
        # ~1500 iterations in a shallow valley, both stop on the relative-reduction test; which of the
        # two ends lower is decided by rounding (measured: 2e-5 below SciPy with one build of the action
        # kernel, 3e-5 above with the next; the reference's own ulp-perturbed twin runs of this problem
        # differ by 4.6e-5 / 3.3e-3, tests/golden/nakl_bounded_ladder_golden.npz).  Asserted: within
        # 1e-3 of SciPy, and SciPy restarted at the device's point gains next to nothing.
        # (SciPy on the oracle is just as sensitive: re-associating one product in the oracle's Simpson
        # adjoint moved *its* end point from 0.03824563 after 1342 iterations to 0.03823652 after 3786.)
        assert abs(an.A_array[0] - r.fun) <= 1e-3 * abs(r.fun), (disc, an.A_array[0], r.fun, an.nit_array[0], r.nit)
        r2 = opt.minimize(lambda z: prob.action_grad(z, rf), xmin, method="L-BFGS-B", jac=True,
                          bounds=list(zip(lo, hi)), options=dict(opts, maxiter=60))
        assert A - r2.fun <= 1e-4 * abs(A), (disc, A, r2.fun, r2.nit)
        assert int(np.sum(xmin <= lo) + np.sum(xmin >= hi)) > 0
