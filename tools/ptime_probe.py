import numpy as np, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import golden_util, scipy.optimize as opt
from oracle.ode_port import OdeProblem
from varanneal_b200 import va_ode
z = golden_util.load("ode_ptime_golden.npz")
c = [c for c in golden_util.ptime_cases() if c["name"]=="l96_D20_trapezoid"][0]
alpha, RM, RF0, gtol, ftol = z["ladder/meta"]
N,D = c["X0"].shape
beta = z["ladder/beta"]; tab = z["ladder/table"]
an = va_ode.Annealer(); an.set_model("lorenz96", D); an.set_data(c["Y"], t=c["t"])
an.anneal(c["X0"].copy(), c["P0"].copy(), alpha, beta, RM, RF0, c["Lidx"], [0], dt_model=c["dt_model"], init_to_data=True, disc="trapezoid",
          opt_args={"gtol": gtol, "ftol": ftol, "maxfun": 100000, "maxiter": 100000})
prob = OdeProblem("lorenz96", D, c["Y"], c["Lidx"], c["dt_model"], "trapezoid", c["P0"], [0], RM)
print("exit", an.exitflags, "nit", an.nit_array, "nfev", an.nfev_array)
for i in range(len(beta)):
    rf = RF0*alpha**float(beta[i])
    xp = np.concatenate([an.minpaths[i,:N*D], an.minpaths[i,N*D:]])
    A0,g0 = prob.action_grad(xp, rf)
    r = opt.minimize(lambda x: prob.action_grad(x, rf), xp, jac=True, method="L-BFGS-B", options=dict(gtol=gtol, ftol=ftol, maxiter=100000, maxfun=100000))
    print(i, "dev %.12e ref %.12e oracle@dev %.12e |g| %.2e restart-> %.12e nit %d" % (an.A_array[i], tab[i,1], A0, np.max(np.abs(g0)), r.fun, r.nit))
