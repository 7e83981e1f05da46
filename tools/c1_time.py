"""Development aid: wall time per minimiser cycle of the shipped Lorenz96 example (C1), one path or a
batch.  python tools/c1_time.py [nbeta] [paths]      (VAB_FUSED_TIMING=1 prints the phase split)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util                      # noqa: E402
from varanneal_b200 import va_ode       # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
z = golden_util.load("c1_shipped_ladder_golden.npz")
data = golden_util.load("l96_ladder_golden.npz")["data"]
LIDX = [0, 2, 4, 6, 8, 10, 14, 16]
disc = "trapezoid"
alpha, RM, RF0, gtol, ftol = z[disc + "/meta"][:5]
beta = z[disc + "/table"][:nb, 0]
rng = np.random.default_rng(1)
X0 = z[disc + "/X0"].copy()
P0 = z[disc + "/P0"].copy()
if B > 1:
    X0 = X0[None] + 0.5 * rng.standard_normal((B,) + X0.shape)
    P0 = np.tile(P0, (B, 1))
for rep in range(2):
    an = va_ode.Annealer()
    an.set_model("lorenz96", 20)
    an.set_data(data[:, 1:][:, LIDX], t=data[:, 0])
    t0 = time.time()
    an.anneal(X0.copy(), P0.copy(), alpha, beta, RM, RF0, LIDX, [0], dt_model=0.025, init_to_data=True, disc=disc,
              opt_args={"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000})
    wall = time.time() - t0
    cyc = an._ctx.graph_launches
    print("C1 %d betas, %d path(s): %.3f s, %d graph cycles, %.1f us per cycle, nfev %d, A_last %.10e"
          % (nb, B, wall, cyc, 1e6 * wall / max(cyc, 1), int(np.sum(an.nfev_array)), np.ravel(an.A_array)[-1]))
