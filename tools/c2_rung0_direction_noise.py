"""CPU demonstration behind DESIGN 2.1: on rung 0 of the C2 slice the device's first accepted step is
3.1e-8 shorter than SciPy's.  L-BFGS-B forms its search direction as d = z - x (Cauchy / subspace point
minus the iterate, lnsrlb); with |x| ~ 10 and |g| ~ 1e-9 that subtraction leaves rounding noise of
ulp(x) / |g| ~ 1.6e-6 *per component* in d, which changes |d|, the first trial step 1 / |d| and hence
the interpolated step.  The device forms d = -H g directly.  Running MINPACK's dcstep on the oracle
with either direction reproduces both numbers to 12 digits.  No GPU needed."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import lbfgsb_port
from oracle.ode_port import OdeProblem
_, Y = bench.twin_data()
X0, P0 = bench.initial_paths(1, 1000)
X0[0][:, bench.LIDX] = Y
prob = OdeProblem("lorenz96", bench.D, Y, bench.LIDX, bench.DT, "SimpsonHermite", [8.0], [0], bench.RM)
rf = bench.RF0
x0 = np.append(X0[0].ravel(), P0[0])
f0, g0 = prob.action_grad(x0, rf)
d = -g0                                   # the device's first direction: exact
dz = (x0 - g0) - x0                       # L-BFGS-B's: Cauchy point minus x
print("relative noise of SciPy's first direction, per component: median %.2e, parameter %.2e; |dz|/|d| - 1 = %.3e"
      % (np.median(np.abs(dz - d) / np.abs(d)), abs(dz[-1]-d[-1])/abs(d[-1]), np.linalg.norm(dz)/np.linalg.norm(d) - 1))
for name, dd in (("exact d = -g (device)", d), ("d = (x - g) - x (L-BFGS-B)", dz)):
    stp0 = 1.0 / np.sqrt(dd @ dd)
    f1, g1 = prob.action_grad(x0 + stp0 * dd, rf)
    out = lbfgsb_port._dcstep(0.0, f0, g0 @ dd, 0.0, f0, g0 @ dd, stp0, f1, g1 @ dd, False, 0.0, 5.0 * stp0)
    print("%-28s stp0 %.17e -> stp1 %.17e" % (name, stp0, out[6]))
print("device accepted 1.23084898077413626e+06 ; SciPy/port accepted 1.23084901902853348e+06 (in units of its own d)")
