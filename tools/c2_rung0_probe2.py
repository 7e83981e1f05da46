"""Diagnostic for the open observation of DESIGN 2.1 (rung 0 of the C2 slice: the device's first
accepted step is 3e-8 shorter than SciPy's): evaluates the device action at the *port's* trial points
and compares f and g.d with the oracle there."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from oracle import lbfgsb_port                 # noqa: E402
from oracle.ode_port import OdeProblem        # noqa: E402
from varanneal_b200 import va_ode              # noqa: E402

_, Y = bench.twin_data()
X0, P0 = bench.initial_paths(1, 1000)
X0[0][:, bench.LIDX] = Y
prob = OdeProblem("lorenz96", bench.D, Y, bench.LIDX, bench.DT, "SimpsonHermite", [8.0], [0], bench.RM)
rf = bench.RF0
xp0 = np.append(X0[0].ravel(), P0[0])
pts = []


def fun(z):
    a, g = prob.action_grad(z, rf)
    pts.append((z.copy(), a, g))
    return a, g


q = lbfgsb_port.minimize(fun, xp0, None, None, ftol=1e-8, gtol=1e-8, maxiter=1, maxfun=15000)
an = va_ode.Annealer()
an.set_model("lorenz96", bench.D)
an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
an.anneal_init(X0[0].copy(), P0[0].copy(), bench.ALPHA, [0], bench.RM, bench.RF0, bench.LIDX, [0], dt_model=bench.DT,
               init_to_data=True, disc="SimpsonHermite", opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxiter": 1})
d = -pts[0][2]
for k, (z, a, g) in enumerate(pts):
    A, G = an.A_gradA(z)
    stp = float((z - xp0) @ d / (d @ d))
    print("eval %d stp %.17e | f oracle %.17e device %.17e rel %.2e | g.d oracle %.17e device %.17e rel %.2e | max|dg|/max|g| %.2e | dg_k rel %.2e"
          % (k, stp, a, A, abs(A - a) / abs(a), g @ d, G @ d, abs(G @ d - g @ d) / abs(g @ d), np.max(np.abs(G - g)) / np.max(np.abs(g)),
             abs(G[-1] - g[-1]) / abs(g[-1])))
_, Amin, st = an.min_lbfgs_scipy(xp0)
x = an._download_paths()[0] if hasattr(an, "_download_paths") else None
if x is not None:
    sd = x[:xp0.size] - xp0
    print("device accepted step %.17e ; port %.17e ; ratio-1 %.3e" % (float(sd @ d / (d @ d)), float((q["x"] - xp0) @ d / (d @ d)),
                                                                   float(sd @ (q["x"] - xp0) / ((q["x"] - xp0) @ (q["x"] - xp0))) - 1.0))
