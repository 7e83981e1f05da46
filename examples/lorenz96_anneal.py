"""State and parameter estimation in Lorenz96 by variational annealing -- the flow of the
reference's examples/Lorenz96_D20/Lorenz96_anneal.py on the B200 engine.

    python examples/lorenz96_anneal.py [--D 20] [--N 161] [--inits 8] [--nbeta 101] [--out DIR]

What changes with respect to the reference script: the import, the model (a registered device
model instead of a Python callable), and -- optionally -- a batch of initial paths annealed at
once (X0 of shape (B, N, D)).  The twin-experiment data are generated here
(varanneal_b200.datagen.lorenz96_twin) instead of being loaded from the shipped .npy file;
pass --data FILE to use a file with time in column 0, as the reference does.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import datagen, va_ode            # noqa: E402  (reference: from varanneal import va_ode)
from varanneal_b200.models import lorenz96            # noqa: E402  (reference: def l96(t, x, k): ...)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=20)
    ap.add_argument("--N", type=int, default=161)
    ap.add_argument("--inits", type=int, default=8, help="initial paths annealed concurrently (1 = reference shapes)")
    ap.add_argument("--nbeta", type=int, default=101)
    ap.add_argument("--data", default=None, help=".npy with time in column 0 and all D components after it")
    ap.add_argument("--disc", default="SimpsonHermite")
    ap.add_argument("--out", default=None, help="directory for paths.npy / params.npy / action_errors.npy")
    a = ap.parse_args()

    D = a.D
    Lidx = [i for i in [0, 2, 4, 6, 8, 10, 14, 16] if i < D] if D <= 20 else [i for i in range(D) if i % 5 in (0, 2)]
    RM, RF0 = 1.0 / 0.5 ** 2, 4.0e-6
    alpha, beta_array = 1.5, np.linspace(0, a.nbeta - 1, a.nbeta)

    if a.data is None:
        times_data, _, data = datagen.lorenz96_twin(D=D, N=a.N, dt=0.025, k=8.17, sigma=0.5, Lidx=Lidx, seed=100)
    else:
        full = np.load(a.data)
        times_data, data = full[:, 0], full[:, 1:][:, Lidx]
    N_model = len(times_data)

    np.random.seed(12345)
    B = a.inits
    X0 = 20.0 * np.random.rand(B, N_model, D) - 10.0
    P0 = 4.0 * np.random.rand(B, 1) + 6.0
    if B == 1:                                         # exactly the reference's shapes
        X0, P0 = X0[0], P0[0]

    anneal1 = va_ode.Annealer()
    anneal1.set_model(lorenz96, D)
    anneal1.set_data(data, t=times_data)
    BFGS_options = {'gtol': 1.0e-8, 'ftol': 1.0e-8, 'maxfun': 1000000, 'maxiter': 1000000}
    tstart = time.time()
    anneal1.anneal(X0, P0, alpha, beta_array, RM, RF0, Lidx, [0], dt_model=times_data[1] - times_data[0],
                   init_to_data=True, disc=a.disc, method='L-BFGS-B', opt_args=BFGS_options, adolcID=0)
    print("annealing of %d path(s) over %d betas completed in %.2f s (%d evaluations)"
          % (B, a.nbeta, time.time() - tstart, int(anneal1.nfev_array.sum())))
    A = anneal1.A_array if B > 1 else anneal1.A_array[None]
    k = (anneal1.minpaths[:, -1, -1] if B > 1 else anneal1.minpaths[-1:, -1])
    print("last rung: action min %.6e  median %.6e  max %.6e;  forcing estimate of the best path %.4f (truth 8.17)"
          % (A[:, -1].min(), np.median(A[:, -1]), A[:, -1].max(), k[np.argmin(A[:, -1])]))
    if a.out:
        os.makedirs(a.out, exist_ok=True)
        best = int(np.argmin(A[:, -1]))
        kw = {"init": best} if B > 1 else {}
        anneal1.save_paths(os.path.join(a.out, "paths.npy"), **kw)
        anneal1.save_params(os.path.join(a.out, "params.npy"), **kw)
        anneal1.save_action_errors(os.path.join(a.out, "action_errors.npy"), **kw)
        print("saved paths / params / action_errors of initialisation %d to %s" % (best, a.out))


if __name__ == "__main__":
    main()
