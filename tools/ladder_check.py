"""Development aid: device ladder vs the golden reference ladder (tests/golden/l96_ladder_golden.npz)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_util
from varanneal_b200 import va_ode
z = golden_util.load("l96_ladder_golden.npz")
data = z["data"]
Lidx = [0, 2, 4, 6, 8, 10, 14, 16]
for disc in ("trapezoid", "SimpsonHermite"):
    alpha, RM, RF0, gtol, ftol = z[disc + "/meta"]
    an = va_ode.Annealer()
    an.set_model("lorenz96", 20)
    an.set_data(data[:, 1:][:, Lidx], t=data[:, 0])
    t0 = time.time()
    an.anneal(z[disc + "/X0"].copy(), z[disc + "/P0"].copy(), alpha, z[disc + "/beta"], RM, RF0, Lidx, [0],
              dt_model=0.025, init_to_data=True, disc=disc,
              opt_args={"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000})
    print(disc, "device ladder %.2f s nfev %d" % (time.time() - t0, an.nfev_array.sum()))
    tab = z[disc + "/table"]
    for i in range(len(tab)):
        print("beta %3d ref A %.10e dev A %.10e rel %.2e | k ref %.6f dev %.6f | nit %d st %d" % (
            tab[i, 0], tab[i, 1], an.A_array[i], abs(an.A_array[i] - tab[i, 1]) / tab[i, 1],
            z[disc + "/params"][i], an.minpaths[i, -1], an.nit_array[i], an.exitflags[i]))
