"""Training a network on data from a teacher network by variational annealing -- the flow of the
reference's examples/nnet_twin/nnet_twin_anneal.py (+ its data/gen_params.py, gen_io_pairs.py) on
the B200 engine.

    python examples/nnet_twin_anneal.py [--layers 5] [--width 10] [--M 100] [--inits 4] [--nbeta 60]

Only the weights are estimated (biases fixed at zero), as in the reference script.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import datagen, va_nnet           # noqa: E402  (reference: from varanneal import va_nnet)
from varanneal_b200.va_nnet import sigmoid            # noqa: E402  (reference: def sigmoid(x, W, b): ...)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--width", type=int, default=10)
    ap.add_argument("--M", type=int, default=100)
    ap.add_argument("--inits", type=int, default=4)
    ap.add_argument("--nbeta", type=int, default=60)
    a = ap.parse_args()

    N, d, M, B = a.layers, a.width, a.M, a.inits
    structure = np.array([d] * N)
    (W, b), = datagen.nnet_twin_params(structure, seed=17439860)
    data_in, data_out, _ = datagen.nnet_twin_io(W, b, M, sigma=0.005, seed=43650832)

    RM = 1.0 / 0.005 ** 2
    RF0 = 1.0e-8 * RM * float(np.sum(structure) - structure[0]) / float(structure[0] + structure[-1])
    alpha, beta_array = 1.1, np.linspace(0, a.nbeta - 1, a.nbeta)

    # initial guesses (nnet_twin_anneal.py:71-119): inputs standardised noise, hidden states in
    # [0.4, 0.6], weights U(-1, 1) / fan_in, biases zero and not estimated
    rng = np.random.RandomState(89072545)
    NDnet = int(structure.sum())
    X0 = np.zeros((B, M, NDnet))
    for bb in range(B):
        for m in range(M):
            xin = rng.randn(d)
            X0[bb, m, :d] = (xin - np.average(xin)) / np.std(xin)
            X0[bb, m, d:] = 0.2 * rng.rand(NDnet - d) + 0.4
    X0 = X0.reshape(B, M * NDnet)
    Pidx, P0, off = [], [], 0
    for n in range(N - 1):
        nw = int(structure[n] * structure[n + 1])
        Pidx += list(range(off, off + nw))
        P0.append((2.0 * rng.rand(B, nw) - 1.0) / d)
        P0.append(np.zeros((B, int(structure[n + 1]))))
        off += nw + int(structure[n + 1])
    P0 = np.concatenate(P0, axis=1)

    anneal1 = va_nnet.Annealer()
    anneal1.set_structure(structure)
    anneal1.set_activation(sigmoid)
    anneal1.set_input_data(data_in)
    anneal1.set_output_data(data_out)
    BFGS_options = {'gtol': 1.0e-12, 'ftol': 1.0e-12, 'maxfun': 1000000, 'maxiter': 1000000}
    tstart = time.time()
    anneal1.anneal(X0, P0, alpha, beta_array, RM, RF0, Pidx, method='L-BFGS-B', opt_args=BFGS_options)
    print("annealing of %d initialisation(s), %d examples, %d betas completed in %.2f s (%d evaluations)"
          % (B, M, a.nbeta, time.time() - tstart, int(anneal1.nfev_array.sum())))
    tab = anneal1.action_errors_table(init=0)
    for row in tab[:: max(1, a.nbeta // 6)]:
        print("beta %5.0f  A %.6e  me %.6e  fe %.6e" % (row[0], row[1], row[2], row[3]))


if __name__ == "__main__":
    main()
