"""Development aid: C2 ladder timing on the device. python tools/c2_ladder.py B nbeta"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from varanneal_b200 import va_ode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 20
_, Y = bench.twin_data()
X0, P0 = bench.initial_paths(B, 1000)
an = va_ode.Annealer()
an.set_model("lorenz96", bench.D)
an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
t0 = time.time()
an.anneal_init(X0, P0, bench.ALPHA, np.arange(nb), bench.RM, bench.RF0, bench.LIDX, [0], disc="SimpsonHermite",
               init_to_data=True, opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
print("init %.2f s" % (time.time() - t0))
for i in range(nb):
    t0 = time.time()
    an.anneal_step()
    dt = time.time() - t0
    print("beta %2d  %.2f s  nfev mean %.0f max %d  nit mean %.0f  A mean %.4e  k mean %.4f  flags %s" % (
        i, dt, an.nfev_array[:, i].mean(), an.nfev_array[:, i].max(), an.nit_array[:, i].mean(),
        an.A_array[:, i].mean(), an.minpaths[:, i, -1].mean(), np.bincount(an.exitflags[:, i], minlength=3)))
