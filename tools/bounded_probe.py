"""Diagnostic: device bounded L-BFGS-B vs oracle/lbfgsb_port.py after 1, 2, 3, ... iterations on a
small boxed Lorenz96 problem (prints where the iterates start to differ)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lbfgsb_port                 # noqa: E402
from oracle.ode_port import OdeProblem        # noqa: E402
from varanneal_b200 import va_ode              # noqa: E402

D, N = 10, 41
rng = np.random.RandomState(11)
Lidx = list(range(0, D, 2))
t = 0.02 * np.arange(N)
Y = 3.0 * rng.randn(N, len(Lidx))
X0 = 2.0 * rng.randn(1, N, D)
P0 = np.array([[8.0]])
prob = OdeProblem("lorenz96", D, Y, Lidx, 0.02, "trapezoid", [8.0], [0], 1.0)
rf = 1e-2 * 2.0 ** 6
for name, bounds in (("box[-1,1]", [[-1.0, 1.0]] * D + [[7.5, 8.5]]), ("lower-only", [[-0.5, None]] * D + [[None, None]]),
                     ("far box", [[-1e3, 1e3]] * D + [[-1e3, 1e3]])):
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds])
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds])
    lo = np.concatenate([np.tile(lo[:D], N), lo[D:]])
    hi = np.concatenate([np.tile(hi[:D], N), hi[D:]])
    xp0 = np.append(X0[0].ravel(), P0[0])
    for maxiter in (1, 2, 3, 5, 10, 30, 5000):
        an = va_ode.Annealer()
        an.set_model("lorenz96", D)
        an.set_data(Y, t=t)
        an.anneal(X0.copy(), P0.copy(), 2.0, [6], 1.0, 1e-2, Lidx, [0], disc="trapezoid", bounds=bounds,
                  init_to_data=False, opt_args={"gtol": 1e-10, "ftol": 1e-15, "maxiter": maxiter})
        q = lbfgsb_port.minimize(lambda z: prob.action_grad(z, rf), xp0, lo, hi, ftol=1e-15, gtol=1e-10,
                                 maxiter=maxiter, maxfun=15000)
        x = an.minpaths[0, 0]
        print("%-10s maxiter %4d | device nit %4d nfev %4d st %d A %.12e | port nit %4d nfev %4d st %d A %.12e | max|dx| %.2e  nact dev %d port %d inside %s"
              % (name, maxiter, an.nit_array[0, 0], an.nfev_array[0, 0], an.exitflags[0, 0], an.A_array[0, 0],
                 q["nit"], q["nfev"], q["status"], q["fun"], np.max(np.abs(x - q["x"])),
                 np.sum(x <= lo) + np.sum(x >= hi), np.sum(q["x"] <= lo) + np.sum(q["x"] >= hi),
                 bool(np.all(x >= lo) and np.all(x <= hi))), flush=True)
