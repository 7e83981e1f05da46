"""Golden per-beta ladders of the *real* configurations, from the reference's own anneal() loop
driving SciPy L-BFGS-B (run in the build container; needs /root/reference):

    python tests/golden/make_ladder_golden.py [c1] [c2] [nakl] [nnet]

  c1    BASELINE.json configs[0] as shipped (examples/Lorenz96_D20/Lorenz96_anneal.py:24-30, 51,
        66-68, 82-86): shipped data file, 8 observed components, 101 betas, alpha 1.5,
        gtol = ftol = 1e-8, trapezoid (the signature default) and SimpsonHermite (the script's),
        np.random.seed(12345), X0 drawn before P0.
  c2    a one-initialisation slice of configs[1]: Lorenz96 D=100, N=5001, L=40, SimpsonHermite,
        alpha 2.5, 20 betas, the path bench.py seeds with 1000.
  nakl  bounded problem (va_ode.py:582-605 -> scipy bounds, _autodiffmin.py:85-86): the tutorial's
        NaKL neuron (first 201 rows of its data), the tutorial's box (ipynb:3069-3087, 3139-3141)
        with two parameter intervals shrunk so that bounds are *active* at the minimisers.
  nnet  va_nnet ladder with every weight estimated (examples/nnet_twin/nnet_twin_anneal.py:101-119
        estimates all weights; biases fixed at 0), structure [10]*6.

The reference's anneal_init / anneal_step / RF ladder / warm starts / bounds expansion run
verbatim (oracle.ref_shim); the gradient handed to SciPy is the NumPy adjoint of oracle.*_port,
pinned against complex-step through the reference (tests/test_oracle.py).  Stored per rung:
[beta, A, me, fe, fe/RF-scale], SciPy's nit / nfev / status, the estimated parameters; inputs that
cannot be regenerated on the GPU box are stored too.
"""
import os
import sys
import time

import numpy as np
import scipy.optimize as opt

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.models_np import MODELS              # noqa: E402
from oracle.ode_port import OdeProblem           # noqa: E402
from oracle import nnet_port                     # noqa: E402

REF = ref_shim.REFERENCE_ROOT
L96_FILE = os.path.join(REF, "examples", "Lorenz96_D20", "l96_D20_dt0p025_N161_sm0p5_sec1_mem1.npy")
NAKL_DIR = os.path.join(REF, "examples", "jupyter-tutorial", "NaKL", "data")
ShimOde, ShimNnet = ref_shim.make_shim_classes()


class _Counting(object):
    """Wraps scipy.optimize.minimize to keep nit / nfev / status of every rung (the reference
    drops them)."""

    def __init__(self):
        self.rows = []
        self._orig = opt.minimize

    def __enter__(self):
        def wrapped(*a, **k):
            r = self._orig(*a, **k)
            self.rows.append((int(r.nit), int(r.nfev), int(r.status)))
            return r
        opt.minimize = wrapped
        return self

    def __exit__(self, *exc):
        opt.minimize = self._orig


def _self_restart(action_grad, minpaths, rfs, opts, bounds=None):
    """What SciPy L-BFGS-B does when it is started again at its *own* minimiser of every rung
    (empty memory): (nit, nfev, further decrease of A).  This is the yardstick for the acceptance
    test of SURVEY.md 7.4(2) -- with ftol-type stops a restart is not always immediate even from
    the reference's own result."""
    rows = []
    for xp, rf in zip(minpaths, rfs):
        A0 = action_grad(xp, rf)[0]
        r = opt.minimize(lambda z: action_grad(z, rf), xp, method="L-BFGS-B", jac=True, bounds=bounds,
                         options=dict(opts))
        rows.append((int(r.nit), int(r.nfev), float(A0 - r.fun)))
    return np.array(rows)


def _perturb(X0, k=1):
    """X0 moved by k units in the last place: the same problem to 16 digits."""
    return X0 * (1.0 + k * 2.0 ** -52)


def _table(an, RF0_scalar, alpha):
    beta = np.asarray(an.beta_array, dtype=float)
    return np.column_stack([beta, an.A_array, an.me_array, an.fe_array,
                            an.fe_array / (RF0_scalar * alpha ** beta)])


def make_c1():
    data = np.load(L96_FILE)
    t, Yall = data[:, 0], data[:, 1:]
    Lidx = [0, 2, 4, 6, 8, 10, 14, 16]
    out = {}
    for disc in ("trapezoid", "SimpsonHermite"):
        np.random.seed(12345)
        X0 = (20.0 * np.random.rand(161 * 20) - 10.0).reshape((161, 20))
        P0 = np.array([4.0 * np.random.rand() + 6.0])
        beta_array = np.linspace(0.0, 100.0, 101)
        alpha, RM, RF0 = 1.5, 4.0, 4e-6
        opts = {"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000}
        an = ShimOde()
        an.set_model(MODELS["lorenz96"], 20)
        an.set_data(Yall[:, Lidx], t=t)
        prob = OdeProblem("lorenz96", 20, Yall[:, Lidx], Lidx, 0.025, disc, P0, [0], RM)
        an.grad_fn = lambda XP, an=an, prob=prob: prob.action_grad(XP, an.RF)[1]
        t0 = time.time()
        with _Counting() as cnt:
            an.anneal_quiet(X0.copy(), P0.copy(), alpha, beta_array, RM, RF0, np.array(Lidx), [0],
                            dt_model=0.025, init_to_data=True, disc=disc, method="L-BFGS-B", opt_args=opts)
        wall = time.time() - t0
        tab = _table(an, RF0, alpha)
        print("c1 %s: %.1f s, nfev %d; A first/last %.6e %.6e" % (
            disc, wall, sum(r[1] for r in cnt.rows), tab[0, 1], tab[-1, 1]), flush=True)
        out[disc + "/X0"] = X0
        out[disc + "/P0"] = P0
        out[disc + "/table"] = tab
        out[disc + "/counts"] = np.array(cnt.rows)
        out[disc + "/params"] = an.minpaths[:, -1]
        out[disc + "/meta"] = np.array([alpha, RM, RF0, opts["gtol"], opts["ftol"], wall])
        out[disc + "/minpaths"] = an.minpaths.copy()
        rfs = RF0 * alpha ** beta_array
        out[disc + "/self_restart"] = _self_restart(prob.action_grad, an.minpaths, rfs, opts)
        # the same ladder from an initial path moved by one unit in the last place: the reference's
        # own reproducibility (how far two runs of the *same* code drift apart, rung by rung)
        for k in (1, 2):
            an2 = ShimOde()
            an2.set_model(MODELS["lorenz96"], 20)
            an2.set_data(Yall[:, Lidx], t=t)
            an2.grad_fn = lambda XP, an2=an2, prob=prob: prob.action_grad(XP, an2.RF)[1]
            an2.anneal_quiet(_perturb(X0, k), P0.copy(), alpha, beta_array, RM, RF0, np.array(Lidx), [0],
                             dt_model=0.025, init_to_data=True, disc=disc, method="L-BFGS-B", opt_args=opts)
            out[disc + "/table_ulp%d" % k] = _table(an2, RF0, alpha)
        sr = out[disc + "/self_restart"]
        band = np.abs(out[disc + "/table_ulp1"][:, 1] - tab[:, 1]) / np.abs(tab[:, 1])
        print("   self-restart: %d of %d rungs need > 2 iterations; ulp-perturbed run: %d rungs differ by > 1e-6, max %.1e"
              % (np.sum(sr[:, 0] > 2), len(sr), np.sum(band > 1e-6), band.max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "c1_shipped_ladder_golden.npz"), **out)


def make_c2():
    sys.path.insert(0, ROOT)
    import bench
    _, Y = bench.twin_data()
    X0, P0 = bench.initial_paths(1, 1000)
    X0, P0 = X0[0], P0[0]
    D, N, Lidx = bench.D, bench.N_MODEL, bench.LIDX
    beta_array = np.arange(bench.N_BETA)
    opts = {"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000}
    an = ShimOde()
    an.set_model(MODELS["lorenz96"], D)
    an.set_data(Y, t=bench.DT * np.arange(N))
    prob = OdeProblem("lorenz96", D, Y, Lidx, bench.DT, "SimpsonHermite", P0, [0], bench.RM)
    an.grad_fn = lambda XP: prob.action_grad(XP, an.RF)[1]
    t0 = time.time()
    with _Counting() as cnt:
        an.anneal_quiet(X0.copy(), P0.copy(), bench.ALPHA, beta_array, bench.RM, bench.RF0, np.array(Lidx), [0],
                        dt_model=bench.DT, init_to_data=True, disc="SimpsonHermite", method="L-BFGS-B",
                        opt_args=opts)
    wall = time.time() - t0
    tab = _table(an, bench.RF0, bench.ALPHA)
    print("c2 slice: %.1f s, nfev %d; A first/last %.6e %.6e" % (
        wall, sum(r[1] for r in cnt.rows), tab[0, 1], tab[-1, 1]), flush=True)
    rfs = bench.RF0 * bench.ALPHA ** beta_array.astype(float)
    sr = _self_restart(prob.action_grad, an.minpaths, rfs, opts)
    an2 = ShimOde()
    an2.set_model(MODELS["lorenz96"], D)
    an2.set_data(Y, t=bench.DT * np.arange(N))
    an2.grad_fn = lambda XP: prob.action_grad(XP, an2.RF)[1]
    an2.anneal_quiet(_perturb(X0), P0.copy(), bench.ALPHA, beta_array, bench.RM, bench.RF0, np.array(Lidx), [0],
                     dt_model=bench.DT, init_to_data=True, disc="SimpsonHermite", method="L-BFGS-B", opt_args=opts)
    tab2 = _table(an2, bench.RF0, bench.ALPHA)
    band = np.abs(tab2[:, 1] - tab[:, 1]) / np.abs(tab[:, 1])
    print("   self-restart: %d of %d rungs need > 2 iterations; ulp-perturbed run: %d rungs differ by > 1e-6, max %.1e"
          % (np.sum(sr[:, 0] > 2), len(sr), np.sum(band > 1e-6), band.max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "c2_slice_ladder_golden.npz"), table=tab, counts=np.array(cnt.rows),
                        params=an.minpaths[:, -1], Ysum=np.array([Y.sum(), np.abs(Y).sum()]),
                        X0sum=np.array([X0.sum(), P0[0]]), self_restart=sr, table_ulp1=tab2,
                        meta=np.array([bench.ALPHA, bench.RM, bench.RF0, opts["gtol"], opts["ftol"], wall]))


NAKL_PB = [[60.0, 180.0], [10.0, 30.0], [0.15, 0.45], [47.5, 52.5], [-80.85, -73.15], [-56.7, -51.3],
           [-42.0, -38.0], [14.25, 15.75], [0.095, 0.105], [0.38, 0.42], [-63.0, -57.0], [-15.75, -14.25],
           [0.95, 1.05], [6.65, 7.35], [-57.75, -52.25], [28.5, 31.5], [0.95, 1.05], [4.75, 5.25]]


def make_nakl():
    Nrows = 201
    V = np.load(os.path.join(NAKL_DIR, "NaKL_Vdata_dt0p02_N6001_sm1p0.npy"))[:Nrows]
    stim = np.load(os.path.join(NAKL_DIR, "NaKL_stim_dt0p02_N6001.npy"))[:Nrows]
    Pb = [list(b) for b in NAKL_PB]
    Pb[0] = [60.0, 110.0]       # gNa: true 120 lies outside -> upper bound active
    Pb[1] = [22.0, 30.0]        # gK: true 20 outside -> lower bound active
    bounds = [[-100.0, 100.0], [0.0, 1.0], [0.0, 1.0], [0.0, 1.0]] + Pb
    rng = np.random.RandomState(20260)
    X0 = np.column_stack([200.0 * rng.rand(Nrows) - 100.0, rng.rand(Nrows), rng.rand(Nrows), rng.rand(Nrows)])
    P0 = np.array([(b[1] - b[0]) * rng.rand() + b[0] for b in Pb])
    alpha, RM, RF0 = 1.1, 1.0, [1e-8, 1e-4, 1e-4, 1e-4]
    # the rungs where the action is resolvable (A ~ 6e-5 ... 0.8; below beta = 100 it is < 1e-5 and
    # every stopping test fires after one iteration) and bounds become active (0 -> 11 components)
    beta_array = np.arange(100, 250, 10)
    out = {"V": V, "stim": stim, "bounds": np.array(bounds), "X0": X0, "P0": P0, "beta": beta_array}
    for disc in ("trapezoid", "SimpsonHermite"):
        opts = {"gtol": 1e-11, "ftol": 1e-13, "maxfun": 1000000, "maxiter": 1000000}
        an = ShimOde()
        an.set_model(MODELS["nakl"], 4)
        an.set_data(V[:, 1:2], stim=stim[:, 1], t=V[:, 0])
        prob = OdeProblem("nakl", 4, V[:, 1:2], [0], 0.02, disc, P0, list(range(18)), RM, stim=stim[:, 1])
        an.grad_fn = lambda XP, an=an, prob=prob: prob.action_grad(XP, np.asarray(an.RF))[1]
        t0 = time.time()
        with _Counting() as cnt:
            an.anneal_quiet(X0.copy(), P0.copy(), alpha, beta_array, RM, RF0, np.array([0]), list(range(18)),
                            dt_model=None, init_to_data=True, disc=disc, method="L-BFGS-B", opt_args=opts,
                            bounds=bounds)
        wall = time.time() - t0
        tab = _table(an, RF0[0], alpha)
        lo = np.concatenate([np.tile(np.array(bounds)[:4, 0], Nrows), np.array(Pb)[:, 0]])
        hi = np.concatenate([np.tile(np.array(bounds)[:4, 1], Nrows), np.array(Pb)[:, 1]])
        nact = [(int(np.sum(an.minpaths[i] <= lo)), int(np.sum(an.minpaths[i] >= hi))) for i in range(len(beta_array))]
        print("nakl %s: %.1f s, nfev %d; A first/last %.6e %.6e; active (lo,hi) first/last %s %s; status %s" % (
            disc, wall, sum(r[1] for r in cnt.rows), tab[0, 1], tab[-1, 1], nact[0], nact[-1],
            sorted(set(r[2] for r in cnt.rows))), flush=True)
        out[disc + "/table"] = tab
        out[disc + "/counts"] = np.array(cnt.rows)
        out[disc + "/params"] = an.minpaths[:, 4 * Nrows:]
        out[disc + "/nactive"] = np.array(nact)
        out[disc + "/meta"] = np.array([alpha, RM, opts["gtol"], opts["ftol"], wall])
    out["RF0"] = np.array(RF0)
    np.savez_compressed(os.path.join(HERE, "nakl_bounded_ladder_golden.npz"), **out)


def make_nnet():
    sys.path.insert(0, ROOT)
    from varanneal_b200 import datagen
    structure = np.array([10] * 6)
    M = 24
    (W, b), = datagen.nnet_twin_params(structure, seed=17439860)
    data_in, data_out, _ = datagen.nnet_twin_io(W, b, M, sigma=0.005, seed=43650832)
    NDnet = int(structure.sum())
    NP = int(sum(structure[n] * structure[n + 1] + structure[n + 1] for n in range(len(structure) - 1)))
    # initial guesses as nnet_twin_anneal.py:71-119 (seed 89072545): states U(0,1), weights U(-.1,.1)-ish, biases 0
    rng = np.random.RandomState(89072545)
    X0 = rng.rand(M * NDnet)
    P0 = np.zeros(NP)
    Pidx = []
    off = 0
    for n in range(len(structure) - 1):
        nw = structure[n] * structure[n + 1]
        P0[off:off + nw] = (2.0 * rng.rand(nw) - 1.0) / structure[n]
        Pidx.extend(range(off, off + nw))
        off += nw + structure[n + 1]
    Pidx = np.array(Pidx)
    RM = 1.0 / 0.005 ** 2
    RF0 = 1e-8 * RM * float(NDnet - structure[0]) / float(structure[0] + structure[-1])
    alpha = 1.1
    beta_array = np.arange(0.0, 436.0, 12.0)
    opts = {"gtol": 1e-12, "ftol": 1e-12, "maxfun": 1000000, "maxiter": 1000000}
    Lidx = [np.arange(structure[0]), np.arange(structure[-1])]
    an = ShimNnet()
    an.set_structure(structure)
    an.set_activation(nnet_port.sigmoid)
    an.set_input_data(data_in)
    an.set_output_data(data_out)
    prob = nnet_port.NnetProblem(structure, data_in, data_out, Lidx, P0, Pidx, RM)
    an.grad_fn = lambda XP: prob.action_grad(XP, an.RF)[1]
    t0 = time.time()
    with _Counting() as cnt:
        an.anneal_quiet(X0.copy(), P0.copy(), alpha, beta_array, RM, RF0, Pidx, Lidx=Lidx, init_to_data=True,
                        method="L-BFGS-B", opt_args=opts)
    wall = time.time() - t0
    tab = _table(an, RF0, alpha)
    print("nnet free weights: %.1f s, nfev %d; A first/last %.6e %.6e; status %s" % (
        wall, sum(r[1] for r in cnt.rows), tab[0, 1], tab[-1, 1], sorted(set(r[2] for r in cnt.rows))), flush=True)
    np.savez_compressed(os.path.join(HERE, "nnet_freeweights_ladder_golden.npz"), structure=structure,
                        data_in=data_in, data_out=data_out, X0=X0, P0=P0, Pidx=Pidx, beta=beta_array, table=tab,
                        counts=np.array(cnt.rows), meta=np.array([RM, RF0, alpha, opts["gtol"], opts["ftol"], wall]))


if __name__ == "__main__":
    what = sys.argv[1:] or ["c1", "c2", "nakl", "nnet"]
    for w in what:
        {"c1": make_c1, "c2": make_c2, "nakl": make_nakl, "nnet": make_nnet}[w]()
