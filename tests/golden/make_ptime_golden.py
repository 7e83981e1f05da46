"""Golden fixtures for *time-dependent parameters* (P0 of shape (N_model, NP), va_ode.py:568-570),
generated from the reference itself.  Run in the build container (needs /root/reference):

    python tests/golden/make_ptime_golden.py

For the branches of the reference that run (trapezoid and SimpsonHermite, va_ode.py:170-188,
:368-369, :416-418) the action is the verbatim reference's and the gradient is complex-step
differentiation through it; a short L-BFGS-B ladder through the reference's own anneal() pins the
XP layout and the write-back (va_ode.py:758-774).  The reference's euler / forwardmap branches of
this mode fail (a (N-1)-row parameter array is sliced a second time, :346-349 after :172-173); those
two cases are stored as *extension* vectors from the NumPy port (N parameter rows, the last never
read), complex-step checked, and flagged `ref=0`.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.models_np import MODELS              # noqa: E402
from oracle.ode_port import OdeProblem           # noqa: E402

REF = ref_shim.REFERENCE_ROOT
NAKL_DIR = os.path.join(REF, "examples", "jupyter-tutorial", "NaKL", "data")
ShimOde, _ = ref_shim.make_shim_classes()

NAKL_P = [120., 20., .3, 50., -77., -54.4, -40., 15., .1, .4, -60., -15., 1., 7., -55., 30., 1., 5.]


def cases():
    V = np.load(os.path.join(NAKL_DIR, "NaKL_Vdata_dt0p02_N6001_sm1p0.npy"))[:41]
    stim = np.load(os.path.join(NAKL_DIR, "NaKL_stim_dt0p02_N6001.npy"))[:41]
    out = []
    for disc in ("trapezoid", "SimpsonHermite", "euler", "forwardmap"):
        out.append(dict(name="l96_D20_" + disc, model="lorenz96", D=20, N=41, dt=0.025, Lidx=[0, 3, 6, 9, 12, 15, 18],
                        P=[8.17], Pidx=[0], stim=None, Y=None, t=None, seed=5, RM=4.0, RF0=4e-3, alpha=1.5, beta=6))
        out.append(dict(name="l63_" + disc, model="lorenz63", D=3, N=31, dt=0.01, Lidx=[0], P=[10.0, 28.0, 8.0 / 3.0],
                        Pidx=[0, 1, 2], stim=None, Y=None, t=None, seed=6, RM=1.0, RF0=[1e-2, 2e-2, 3e-2], alpha=2.0,
                        beta=3))
    for disc, Pidx in (("trapezoid", list(range(18))), ("SimpsonHermite", [0, 3, 17]), ("euler", [1, 7])):
        out.append(dict(name="nakl_%s_%dp" % (disc, len(Pidx)), model="nakl", D=4, N=41, dt=None, Lidx=[0],
                        P=NAKL_P, Pidx=Pidx, stim=stim[:, 1], Y=V[:, 1:2], t=V[:, 0], seed=7, RM=1.0,
                        RF0=[1e-8, 1e-4, 1e-4, 1e-4], alpha=1.1, beta=40))
    return out


def inputs(c):
    rng = np.random.RandomState(c["seed"])
    N, D = c["N"], c["D"]
    if c["model"] == "nakl":
        X0 = np.column_stack([-70 + 20 * rng.randn(N), 0.2 * rng.rand(N) + 0.4,
                              0.2 * rng.rand(N) + 0.4, 0.2 * rng.rand(N) + 0.4])
        Y, t = c["Y"], c["t"]
    else:
        X0 = 20.0 * rng.rand(N, D) - 10.0 if c["model"] == "lorenz96" else 10.0 * rng.randn(N, D)
        Y = rng.randn(N, len(c["Lidx"])) * 3.0
        t = c["dt"] * np.arange(N)
    P0 = np.asarray(c["P"])[None, :] * (1.0 + 0.05 * rng.randn(N, len(c["P"])))
    return X0, P0, Y, t


def main():
    out, names = {}, []
    for c in cases():
        X0, P0, Y, t = inputs(c)
        N, D, Pidx = c["N"], c["D"], c["Pidx"]
        runs = c["name"].split("_")[-1] in ("trapezoid", "SimpsonHermite") or (
            c["model"] == "nakl" and "euler" not in c["name"])
        disc = [d for d in ("trapezoid", "SimpsonHermite", "euler", "forwardmap") if d in c["name"]][0]
        XP = np.append(X0.ravel(), P0[:, Pidx].ravel())
        rf_beta = np.asarray(c["RF0"]) * c["alpha"] ** c["beta"] if not np.isscalar(c["RF0"]) else c["RF0"] * c["alpha"] ** c["beta"]
        dt = c["dt"] if c["dt"] is not None else float(t[1] - t[0])
        prob = OdeProblem(c["model"], D, Y, c["Lidx"], dt, disc, P0, Pidx, c["RM"], stim=c["stim"])
        rf = rf_beta if np.isscalar(rf_beta) else np.resize(rf_beta, (N - 1, D))
        Ap, mep, fep, gp = prob.action_grad(XP, rf, parts=True)
        t0 = time.time()
        if runs:
            an = ShimOde()
            an.set_model(MODELS[c["model"]], D)
            an.set_data(Y, stim=c["stim"], t=t)
            with contextlib.redirect_stdout(io.StringIO()):
                an.anneal_init(X0.copy(), P0.copy(), c["alpha"], [c["beta"]], c["RM"], c["RF0"], np.array(c["Lidx"]),
                               Pidx, dt_model=None, init_to_data=False, disc=disc)
            A = float(an.A(XP))
            me = float(an.me_gaussian(XP[:N * D]))
            fe = float(an.fe_gaussian(XP))
            g = ref_shim.complex_step_grad(an.A, XP)
        else:
            A, me, fe = Ap, mep, fep
            g = ref_shim.complex_step_grad(lambda z: prob.action(z, rf), XP)
        print("%-26s ref=%d n=%5d A=%.16e port rel %.1e grad rel %.1e (%.1fs)" % (
            c["name"], runs, XP.size, A, abs(Ap - A) / abs(A), np.max(np.abs(gp - g)) / np.max(np.abs(g)),
            time.time() - t0))
        n = c["name"]
        names.append(n)
        for k, v in (("X0", X0), ("P0", P0), ("t", t), ("Y", Y),
                     ("stim", np.zeros(0) if c["stim"] is None else c["stim"]),
                     ("Lidx", np.asarray(c["Lidx"], dtype=np.int64)), ("Pidx", np.asarray(Pidx, dtype=np.int64)),
                     ("RM", c["RM"]), ("RF0", np.asarray(c["RF0"], dtype=np.float64)),
                     ("meta", np.array([c["alpha"], c["beta"], dt, float(runs)])),
                     ("model_disc", np.array([c["model"], disc])), ("A", np.array([A, me, fe])), ("grad", g)):
            out[n + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)

    # a short ladder through the reference's anneal(): trapezoid, every parameter a time series
    c = [c for c in cases() if c["name"] == "l96_D20_trapezoid"][0]
    X0, P0, Y, t = inputs(c)
    beta = np.arange(0, 16, 2)
    opts = {"gtol": 1e-10, "ftol": 1e-13, "maxfun": 100000, "maxiter": 100000}
    an = ShimOde()
    an.set_model(MODELS["lorenz96"], c["D"])
    an.set_data(Y, t=t)
    prob = OdeProblem("lorenz96", c["D"], Y, c["Lidx"], c["dt"], "trapezoid", P0, [0], c["RM"])
    an.grad_fn = lambda XP, an=an, prob=prob: prob.action_grad(XP, an.RF)[1]
    t0 = time.time()
    an.anneal_quiet(X0.copy(), P0.copy(), c["alpha"], beta, c["RM"], c["RF0"], np.array(c["Lidx"]), [0],
                    dt_model=c["dt"], init_to_data=True, disc="trapezoid", method="L-BFGS-B", opt_args=opts)
    print("ptime ladder: %.1f s; A %s" % (time.time() - t0, an.A_array))
    out["ladder/beta"] = beta
    out["ladder/table"] = np.column_stack([an.beta_array, an.A_array, an.me_array, an.fe_array])
    # the reference's own reproducibility: the same run from X0 * (1 + 2^-52) (one unit in the last place)
    an2 = ShimOde()
    an2.set_model(MODELS["lorenz96"], c["D"])
    an2.set_data(Y, t=t)
    an2.grad_fn = lambda XP, an=an2, prob=prob: prob.action_grad(XP, an.RF)[1]
    an2.anneal_quiet(X0 * (1.0 + 2.0 ** -52), P0.copy(), c["alpha"], beta, c["RM"], c["RF0"], np.array(c["Lidx"]), [0],
                     dt_model=c["dt"], init_to_data=True, disc="trapezoid", method="L-BFGS-B", opt_args=opts)
    out["ladder/table_ulp1"] = np.column_stack([an2.beta_array, an2.A_array, an2.me_array, an2.fe_array])
    print("ptime ladder twin: rel %s" % (np.abs(an2.A_array - an.A_array) / an.A_array))
    out["ladder/minpaths"] = an.minpaths
    out["ladder/P_final"] = np.array(an.P)
    out["ladder/meta"] = np.array([c["alpha"], c["RM"], c["RF0"], opts["gtol"], opts["ftol"]])
    np.savez_compressed(os.path.join(HERE, "ode_ptime_golden.npz"), **out)


if __name__ == "__main__":
    main()
