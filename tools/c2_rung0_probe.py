"""Diagnostic: rung 0 of the C2 slice, device (unbounded TMA path) vs oracle/lbfgsb_port.py after k iterations."""
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
for maxiter in (1, 2, 3):
    an = va_ode.Annealer()
    an.set_model("lorenz96", bench.D)
    an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
    an.anneal(X0[0].copy(), P0[0].copy(), bench.ALPHA, [0], bench.RM, bench.RF0, bench.LIDX, [0], dt_model=bench.DT,
              init_to_data=True, disc="SimpsonHermite", opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxiter": maxiter})
    q = lbfgsb_port.minimize(lambda z: prob.action_grad(z, rf), xp0, None, None, ftol=1e-8, gtol=1e-8, maxiter=maxiter, maxfun=15000)
    x = an.minpaths[0]
    g = prob.action_grad(x, rf)[1]
    gq = prob.action_grad(q["x"], rf)[1]
    dx = x - q["x"]
    step_p, step_d = q["x"] - xp0, x - xp0
    i = int(np.argmax(np.abs(dx)))
    ratio = float(step_d @ step_p / (step_p @ step_p))
    print("   argmax|dx| at %d of %d (last = parameter); step ratio dev/port - 1 = %.3e; |dx - (ratio-1) step_p|max = %.2e; |step| max %.3e"
          % (i, x.size, ratio - 1.0, np.max(np.abs(dx - (ratio - 1.0) * step_p)), np.max(np.abs(step_p))))
    print("maxiter %4d | device nit %3d nfev %3d st %d A %.12e max|g| %.6e | port nit %3d nfev %3d st %d A %.12e max|g| %.6e | max|dx| %.2e"
          % (maxiter, an.nit_array[0], an.nfev_array[0], an.exitflags[0], an.A_array[0], np.max(np.abs(g)),
             q["nit"], q["nfev"], q["status"], q["fun"], np.max(np.abs(gq)), np.max(np.abs(x - q["x"]))), flush=True)
