"""Pins oracle/lbfgsb_port.py -- the restatement of L-BFGS-B 3.0 that specifies the device's
bounded minimiser -- behaviourally against this image's SciPy (the reference's own minimiser,
_autodiffmin.py:85-86): same number of iterations and evaluations, the same iterates to rounding,
on bounded problems with active lower / upper / two-sided bounds, and on an unbounded one."""
import numpy as np
import pytest
import scipy.optimize as opt

from oracle import lbfgsb_port as lp


def _scipy(fun, x0, lo, hi, **o):
    its = []
    b = list(zip(np.where(np.isfinite(lo), lo, None), np.where(np.isfinite(hi), hi, None)))
    r = opt.minimize(fun, x0, method="L-BFGS-B", jac=True, bounds=b, options=o, callback=lambda xk: its.append(xk.copy()))
    return r, its


def _port(fun, x0, lo, hi, **o):
    its = []
    q = lp.minimize(fun, x0, lo, hi, ftol=o["ftol"], gtol=o["gtol"], maxiter=o["maxiter"], maxfun=o["maxfun"],
                    callback=lambda xk, f: its.append(xk))
    return q, its


def _rosen(x):
    f = np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
    g[1:] += 200 * (x[1:] - x[:-1] ** 2)
    return f, g


def _cases():
    rng = np.random.RandomState(0)
    n = 50
    A = rng.randn(n, n)
    A = A @ A.T + 0.1 * np.eye(n)
    b = rng.randn(n) * 5
    quad = lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b)       # noqa: E731
    lo = np.full(n, -np.inf); lo[::3] = 0.0
    hi = np.full(n, np.inf); hi[1::4] = 0.1
    x0 = rng.randn(n)
    yield "quadratic box", quad, x0, np.full(n, -0.3), np.full(n, 0.3), None
    yield "quadratic mixed", quad, x0, lo, hi, None
    m = 30
    yield "rosenbrock lower", _rosen, np.full(m, 1.5) + 0.1 * rng.randn(m), np.full(m, 1.1), np.full(m, np.inf), None
    yield "rosenbrock box", _rosen, np.full(m, -0.5) + 0.1 * rng.randn(m), np.full(m, -1.0), np.full(m, 0.8), 40
    yield "rosenbrock free", _rosen, np.full(m, -0.5) + 0.1 * rng.randn(m), np.full(m, -np.inf), np.full(m, np.inf), 25


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_port_follows_scipy_iterate_for_iterate(case):
    name, fun, x0, lo, hi, horizon = case
    o = dict(ftol=1e-12, gtol=1e-10, maxiter=15000, maxfun=15000)
    r, its = _scipy(fun, x0, lo, hi, **o)
    q, mine = _port(fun, x0, lo, hi, **o)
    assert np.all(q["x"] >= lo) and np.all(q["x"] <= hi)
    assert abs(q["fun"] - r.fun) <= 1e-9 * max(1.0, abs(r.fun))
    if horizon is None:                      # well conditioned: the whole run coincides
        assert (q["nit"], q["nfev"]) == (r.nit, r.nfev)
        horizon = len(its)
    # (nonconvex problems decorrelate after some tens of iterations: rounding, not algorithm)
    k = min(horizon, len(its), len(mine))
    dev = max(np.max(np.abs(its[i] - mine[i])) / max(1.0, np.max(np.abs(its[i]))) for i in range(k))
    assert dev <= 1e-8, dev


def test_port_stops_like_scipy_on_limits():
    rng = np.random.RandomState(3)
    x0 = np.full(20, -0.5) + 0.1 * rng.randn(20)
    lo, hi = np.full(20, -1.0), np.full(20, 0.8)
    o = dict(ftol=1e-14, gtol=1e-12, maxiter=7, maxfun=15000)
    r, _ = _scipy(_rosen, x0, lo, hi, **o)
    q, _ = _port(_rosen, x0, lo, hi, **o)
    assert r.status == 1 and q["status"] == 1 and q["nit"] == r.nit == 7 and q["nfev"] == r.nfev
    assert np.allclose(q["x"], r.x, rtol=0, atol=1e-10)
