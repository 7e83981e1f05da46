"""Development aid: device-only ladder on the shipped-size problem (D=20, N=161). args: B nbeta maxiter"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import va_ode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
rng = np.random.RandomState(5)
D, N = 20, 161
Lidx = [i for i in range(D) if i % 5 in (0, 2)]
Y = rng.randn(N, len(Lidx))
X0 = 20 * rng.rand(B, N, D) - 10
P0 = 4 * rng.rand(B, 1) + 6
an = va_ode.Annealer()
an.set_model("lorenz96", D)
an.set_data(Y, t=0.025 * np.arange(N))
an.anneal_init(X0, P0, 2.0, np.arange(nb), 4.0, 4e-6, Lidx, [0], disc="trapezoid",
               opt_args={"gtol": 1e-11, "ftol": 1e-15, "maxfun": 1000000, "maxiter": maxiter})
import torch
torch.cuda.synchronize()
t0 = time.time()
for i in range(nb):
    an.anneal_step()
dt = time.time() - t0
cyc = an.nfev_array.max(axis=0).sum()
print("B=%d: %.3f s, cycles ~%d -> %.1f us/cycle, nfev total %d -> %.0f evals/s" % (
    B, dt, cyc, 1e6 * dt / cyc, an.nfev_array.sum(), an.nfev_array.sum() / dt))
