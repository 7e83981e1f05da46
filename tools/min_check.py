"""Development aid: device L-BFGS-B vs SciPy L-BFGS-B (oracle action) on a small twin problem."""
import os, sys, time
import numpy as np
import scipy.optimize as opt
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ode_port import OdeProblem
from varanneal_b200 import va_ode

D = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 161
disc = sys.argv[3] if len(sys.argv) > 3 else "trapezoid"
nbeta = int(sys.argv[4]) if len(sys.argv) > 4 else 12
B = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dt, k = 0.025, 8.17
def l96(x): return np.roll(x, 1) * (np.roll(x, -1) - np.roll(x, 2)) - x + k
rng = np.random.RandomState(5)
x = k + rng.randn(D)
rows = []
for n in range(500 + N):
    k1 = l96(x); k2 = l96(x + 0.5*dt*k1); k3 = l96(x + 0.5*dt*k2); k4 = l96(x + dt*k3)
    x = x + dt/6*(k1 + 2*k2 + 2*k3 + k4)
    if n >= 500: rows.append(x.copy())
truth = np.array(rows)
Lidx = [i for i in range(D) if i % 5 in (0, 2)]
Y = truth[:, Lidx] + 0.5 * rng.randn(N, len(Lidx))
t = dt * np.arange(N)
X0 = 20 * rng.rand(B, N, D) - 10
P0 = 4 * rng.rand(B, 1) + 6
alpha, RM, RF0 = 2.0, 4.0, 4e-6
betas = np.arange(nbeta)
opts = {"gtol": float(os.environ.get("GTOL", 1e-8)), "ftol": float(os.environ.get("FTOL", 1e-8)), "maxfun": 1000000, "maxiter": 1000000}

an = va_ode.Annealer()
an.set_model("lorenz96", D)
an.set_data(Y, t=t)
X0g = X0.copy()
t0 = time.time()
an.anneal(X0g, P0.copy(), alpha, betas, RM, RF0, Lidx, [0], disc=disc, opt_args=opts)
tg = time.time() - t0
print("device ladder: %.2f s, nfev total %d, iters %d, launches %d" % (tg, an.nfev_array.sum(), an.nit_array.sum(), an.gpu_launches))

prob = OdeProblem("lorenz96", D, Y, Lidx, dt, disc, [8.0], [0], RM)
for b in range(B):
    X = X0[b].copy(); X[:, Lidx] = Y
    xp = np.append(X.ravel(), P0[b])
    t0 = time.time(); nf = 0
    for ib, beta in enumerate(betas):
        rf = RF0 * alpha ** beta
        res = opt.minimize(lambda z: prob.action_grad(z, rf), xp, method="L-BFGS-B", jac=True, options=opts)
        xp = res.x; nf += res.nfev
        Ag = an.A_array[b, ib]
        print("b=%d beta=%2d scipy A=%.10e nit=%5d st=%d | device A=%.10e nit=%5d st=%d rel=%.2e P=%.6f/%.6f" % (
            b, beta, res.fun, res.nit, res.status, Ag, an.nit_array[b, ib], an.exitflags[b, ib],
            abs(Ag - res.fun) / abs(res.fun), xp[-1], an.minpaths[b, ib, -1]))
    print("scipy ladder %.2f s, nfev %d" % (time.time() - t0, nf))
