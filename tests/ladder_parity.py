"""Shared helpers of the per-beta ladder parity tests (tests/test_gpu_ladder_configs.py) and of
tools/parity_probe.py: run the device ladder of one of the golden configurations
(tests/golden/make_ladder_golden.py), compare the per-beta minimum actions with the reference +
SciPy table, and apply the acceptance test of SURVEY.md 7.4(2) to every rung:

    SciPy L-BFGS-B on the *oracle* action, started at the device's minimiser with the run's own
    options, must terminate at once (<= 2 iterations) at the same action.

The acceptance test is what "the same minimum" can mean for an optimiser that stops on
(f_k - f_{k+1}) <= ftol * max(|f_k|, |f_{k+1}|, 1): two correct L-BFGS-B implementations that
differ in the rounding of their dot products stop a few iterations apart, i.e. at actions that
differ by up to a few ftol * max(|A|, 1) -- far above 1e-6 * A on the early, flat rungs where
A ~ 1e-5.
"""
import os
import sys

import numpy as np
import scipy.optimize as opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import golden_util                                   # noqa: E402
from oracle.ode_port import OdeProblem               # noqa: E402
from oracle import nnet_port                         # noqa: E402

LIDX_C1 = [0, 2, 4, 6, 8, 10, 14, 16]


RESTART_MAXITER = 60      # enough to tell "stops at once" from "keeps going"; keeps the GPU suite short


def restart(action_grad, XP, rf, opts, bounds=None):
    """SciPy L-BFGS-B on the oracle action from XP: (nit, nfev, A_end, A_start, max|proj g| at XP);
    at most RESTART_MAXITER iterations (a restart that has not stopped by then counts as slow and its
    gain is what those iterations achieved)."""
    opts = dict(opts)
    opts["maxiter"] = min(int(opts.get("maxiter", RESTART_MAXITER)), RESTART_MAXITER)
    A0, g0 = action_grad(XP, rf)
    if bounds is not None:
        lo, hi = bounds
        pg = np.where(g0 < 0, np.maximum(XP - hi, g0), np.minimum(XP - lo, g0))
        b = list(zip(np.where(np.isfinite(lo), lo, None), np.where(np.isfinite(hi), hi, None)))
    else:
        pg, b = g0, None
    r = opt.minimize(lambda z: action_grad(z, rf), XP, method="L-BFGS-B", jac=True, bounds=b,
                     options=dict(opts))
    return int(r.nit), int(r.nfev), float(r.fun), float(A0), float(np.max(np.abs(pg)))


def summarize(name, beta, A_dev, A_ref, rows):
    """rows: per rung (nit, nfev, A_end, A_start, pg).  Returns a dict of arrays for printing /
    asserting."""
    rows = np.array(rows, dtype=float)
    rel = np.abs(A_dev - A_ref) / np.abs(A_ref)
    out = dict(name=name, beta=np.asarray(beta, dtype=float), A_dev=A_dev, A_ref=A_ref, rel=rel,
               nit=rows[:, 0], nfev=rows[:, 1], A_end=rows[:, 2], A_oracle=rows[:, 3], pg=rows[:, 4])
    out["drop"] = (rows[:, 3] - rows[:, 2])                   # what SciPy still gains from the device point
    out["oracle_rel"] = np.abs(rows[:, 3] - A_dev) / np.abs(A_dev)
    return out


def fmt(s):
    lines = ["%s: beta  A_dev  A_ref  rel  | restart nit nfev drop drop/max(A,1) drop/A  pg | A_dev vs oracle" % s["name"]]
    for i in range(len(s["beta"])):
        lines.append("%6.1f %.10e %.10e %.1e | %2d %3d %.2e %.1e %.1e %.1e | %.1e" % (
            s["beta"][i], s["A_dev"][i], s["A_ref"][i], s["rel"][i], s["nit"][i], s["nfev"][i], s["drop"][i],
            s["drop"][i] / max(abs(s["A_dev"][i]), 1.0), s["drop"][i] / abs(s["A_dev"][i]), s["pg"][i],
            s["oracle_rel"][i]))
    return "\n".join(lines)


# ------------------------------------------------------------------------------------------------
def run_c1(disc, nbeta=None):
    """BASELINE.json configs[0] as shipped (101 betas, gtol = ftol = 1e-8)."""
    from varanneal_b200 import va_ode
    z = golden_util.load("c1_shipped_ladder_golden.npz")
    old = golden_util.load("l96_ladder_golden.npz")             # holds the shipped data file
    data = old["data"]
    alpha, RM, RF0, gtol, ftol = z[disc + "/meta"][:5]
    tab = z[disc + "/table"]
    nb = len(tab) if nbeta is None else nbeta
    beta = tab[:nb, 0]
    opts = {"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000}
    an = va_ode.Annealer()
    an.set_model("lorenz96", 20)
    an.set_data(data[:, 1:][:, LIDX_C1], t=data[:, 0])
    an.anneal(z[disc + "/X0"].copy(), z[disc + "/P0"].copy(), alpha, beta, RM, RF0, LIDX_C1, [0],
              dt_model=0.025, init_to_data=True, disc=disc, opt_args=opts)
    prob = OdeProblem("lorenz96", 20, data[:, 1:][:, LIDX_C1], LIDX_C1, 0.025, disc, [8.0], [0], RM)
    rows = [restart(prob.action_grad, an.minpaths[i], RF0 * alpha ** float(beta[i]), opts) for i in range(nb)]
    return an, summarize("c1/" + disc, beta, an.A_array.copy(), tab[:nb, 1], rows), z


def run_c2_slice(nbeta=None):
    """One initialisation of BASELINE.json configs[1] (the path bench.py seeds with 1000)."""
    import bench
    from varanneal_b200 import va_ode
    z = golden_util.load("c2_slice_ladder_golden.npz")
    _, Y = bench.twin_data()
    assert np.allclose([Y.sum(), np.abs(Y).sum()], z["Ysum"], rtol=1e-13), "twin data differs from the golden run's"
    X0, P0 = bench.initial_paths(1, 1000)
    tab = z["table"]
    nb = len(tab) if nbeta is None else nbeta
    beta = tab[:nb, 0]
    opts = {"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000}
    an = va_ode.Annealer()
    an.set_model("lorenz96", bench.D)
    an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
    an.anneal(X0[0].copy(), P0[0].copy(), bench.ALPHA, beta, bench.RM, bench.RF0, bench.LIDX, [0],
              dt_model=bench.DT, init_to_data=True, disc="SimpsonHermite", opt_args=opts)
    prob = OdeProblem("lorenz96", bench.D, Y, bench.LIDX, bench.DT, "SimpsonHermite", [8.0], [0], bench.RM)
    rows = [restart(prob.action_grad, an.minpaths[i], bench.RF0 * bench.ALPHA ** float(beta[i]), opts)
            for i in range(nb)]
    return an, summarize("c2-slice", beta, an.A_array.copy(), tab[:nb, 1], rows), z


def run_nakl(disc, nbeta=None):
    """Bounded NaKL ladder (tutorial box, two parameter intervals active)."""
    from varanneal_b200 import va_ode
    z = golden_util.load("nakl_bounded_ladder_golden.npz")
    V, stim, bounds = z["V"], z["stim"], z["bounds"]
    alpha, RM, gtol, ftol = z[disc + "/meta"][:4]
    RF0 = list(z["RF0"])
    tab = z[disc + "/table"]
    nb = len(tab) if nbeta is None else nbeta
    beta = tab[:nb, 0]
    opts = {"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000}
    N = V.shape[0]
    an = va_ode.Annealer()
    an.set_model("nakl", 4)
    an.set_data(V[:, 1:2], stim=stim[:, 1], t=V[:, 0])
    an.anneal(z["X0"].copy(), z["P0"].copy(), alpha, beta, RM, RF0, [0], list(range(18)), dt_model=None,
              init_to_data=True, disc=disc, bounds=[list(b) for b in bounds], opt_args=opts)
    prob = OdeProblem("nakl", 4, V[:, 1:2], [0], an.dt_model, disc, z["P0"], list(range(18)), RM, stim=stim[:, 1])
    lo = np.concatenate([np.tile(bounds[:4, 0], N), bounds[4:, 0]])
    hi = np.concatenate([np.tile(bounds[:4, 1], N), bounds[4:, 1]])
    rf = lambda b: np.resize(np.asarray(RF0), (N - 1, 4)) * alpha ** float(b)   # noqa: E731
    rows = [restart(prob.action_grad, an.minpaths[i], rf(beta[i]), opts, bounds=(lo, hi)) for i in range(nb)]
    s = summarize("nakl/" + disc, beta, an.A_array.copy(), tab[:nb, 1], rows)
    s["nactive_dev"] = np.array([[np.sum(an.minpaths[i] <= lo), np.sum(an.minpaths[i] >= hi)] for i in range(nb)])
    s["nactive_ref"] = z[disc + "/nactive"][:nb]
    s["inside"] = bool(np.all(an.minpaths[:nb] >= lo - 0.0) and np.all(an.minpaths[:nb] <= hi + 0.0))
    return an, s, z


def run_nnet(nbeta=None):
    """va_nnet ladder with every weight estimated ([10]*6, M = 24)."""
    from varanneal_b200 import va_nnet
    z = golden_util.load("nnet_freeweights_ladder_golden.npz")
    RM, RF0, alpha, gtol, ftol = z["meta"][:5]
    st = z["structure"]
    tab = z["table"]
    nb = len(tab) if nbeta is None else nbeta
    beta = z["beta"][:nb]
    opts = {"gtol": gtol, "ftol": ftol, "maxfun": 1000000, "maxiter": 1000000}
    Lidx = [np.arange(st[0]), np.arange(st[-1])]
    an = va_nnet.Annealer()
    an.set_structure(st)
    an.set_activation("sigmoid")
    an.set_input_data(z["data_in"])
    an.set_output_data(z["data_out"])
    an.anneal(z["X0"].copy(), z["P0"].copy(), alpha, beta, RM, RF0, z["Pidx"], Lidx=Lidx, init_to_data=True,
              opt_args=opts)
    prob = nnet_port.NnetProblem(st, z["data_in"], z["data_out"], Lidx, z["P0"], z["Pidx"], RM)
    NDens = an.NDens
    est = lambda row: np.concatenate([row[:NDens], row[NDens:][z["Pidx"]]])     # noqa: E731
    rows = [restart(prob.action_grad, est(an.minpaths[i]), RF0 * alpha ** float(beta[i]), opts) for i in range(nb)]
    return an, summarize("nnet-freeweights", beta, an.A_array.copy(), tab[:nb, 1], rows), z
