"""Development aid: the shared-memory-resident ladder kernel (lb_resident_kernel) against the
per-phase kernels on short ladders (few iterations, so that rounding differences cannot grow), the
oracle at its minimisers, and its wall time on the shipped example.
    python tools/resident_probe.py [time]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util                      # noqa: E402
from oracle.ode_port import OdeProblem  # noqa: E402
from varanneal_b200 import va_ode       # noqa: E402

LIDX = [0, 2, 4, 6, 8, 10, 14, 16]
data = golden_util.load("l96_ladder_golden.npz")["data"]


def run(mode, disc, Pidx, B, nb, maxiter, N=161, nskip=1, seed=3):
    os.environ["VAB_LBFGS_RESIDENT"] = mode
    rng = np.random.RandomState(seed)
    an = va_ode.Annealer()
    an.set_model("lorenz96", 20)
    Y = data[:N:nskip, 1:][:, LIDX]
    an.set_data(Y, t=data[:N:nskip, 0])
    N = nskip * (len(Y) - 1) + 1
    X0 = 20.0 * rng.rand(B, N, 20) - 10.0
    P0 = 8.0 + 0.5 * rng.randn(B, 1)
    if B == 1:
        X0, P0 = X0[0], P0[0]
    an.anneal(X0, P0, 2.0, np.arange(0, 3 * nb, 3), 4.0, 4e-6, LIDX, Pidx, dt_model=0.025,
              init_to_data=True, disc=disc, opt_args={"gtol": 1e-11, "ftol": 1e-15, "maxfun": 1000000, "maxiter": maxiter})
    return an


quick = len(sys.argv) > 1 and sys.argv[1] == "time1"
if len(sys.argv) < 2 or sys.argv[1] not in ("time", "time1"):
    for disc in ("trapezoid", "SimpsonHermite", "euler", "forwardmap"):
        for Pidx in ([0], []):
            for (B, N, nskip) in ((2, 161, 1), (3, 41, 2)):
                a = run("1", disc, Pidx, B, 3, 6, N=N, nskip=nskip)
                b = run("0", disc, Pidx, B, 3, 6, N=N, nskip=nskip)
                dA = np.max(np.abs(a.A_array - b.A_array) / np.abs(b.A_array))
                dx = np.max(np.abs(a.minpaths - b.minpaths))
                same = np.array_equal(a.nit_array, b.nit_array) and np.array_equal(a.nfev_array, b.nfev_array)
                print("%-15s Pidx=%-4s B=%d N=%3d nskip=%d: rel dA %.2e  max|dx| %.2e  same counts %s  nit %s" % (
                    disc, Pidx, B, N, nskip, dA, dx, same, np.ravel(a.nit_array)[:3]))
else:
    z = golden_util.load("c1_shipped_ladder_golden.npz")
    for disc in (("trapezoid",) if quick else ("trapezoid", "SimpsonHermite")):
        alpha, RM, RF0, gtol, ftol = z[disc + "/meta"][:5]
        beta = z[disc + "/table"][:, 0]
        plan = ((1, "8"), (1, "4")) if quick else \
            ((1, "8"), (1, "4"), (1, "0"), (16, "8"), (16, "4"), (16, "0"), (64, "4"), (64, "0"))
        for (B, mode) in plan:
            if True:
                if mode == "-1":
                    os.environ.pop("VAB_LBFGS_RESIDENT", None)
                else:
                    os.environ["VAB_LBFGS_RESIDENT"] = mode
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
                    an.anneal(X0.copy(), P0.copy(), alpha, beta, RM, RF0, LIDX, [0], dt_model=0.025, init_to_data=True,
                              disc=disc, opt_args={"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000})
                    wall = time.time() - t0
                nfev = int(np.sum(an.nfev_array))
                cyc = int(np.max(np.sum(np.atleast_2d(an.nfev_array), axis=-1)))
                print("C1 %-15s %3d path(s) resident=%s: %.3f s, nfev %d, longest path %d cycles -> %.2f us per cycle, A_last %.10e"
                      % (disc, B, mode, wall, nfev, cyc, 1e6 * wall / cyc, np.ravel(an.A_array)[-1]), flush=True)
