"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def ode_case(z, n):
    """Dict of one ODE action case of ode_action_golden.npz."""
    alpha, beta, dtm = z[n + "/meta"]
    model, disc = [str(s) for s in z[n + "/model_disc"]]
    RM = z[n + "/RM"]
    RF0 = z[n + "/RF0"]
    stim = z[n + "/stim"]
    return dict(name=n, model=model, disc=disc, X0=z[n + "/X0"], P0=z[n + "/P0"], t=z[n + "/t"], Y=z[n + "/Y"],
                stim=None if stim.size == 0 else stim, Lidx=z[n + "/Lidx"], Pidx=z[n + "/Pidx"],
                RM=float(RM) if RM.ndim == 0 else RM, RF0=float(RF0) if RF0.ndim == 0 else RF0,
                alpha=float(alpha), beta=int(beta), dt_model=None if dtm < 0 else float(dtm),
                A=z[n + "/A"], grad=z[n + "/grad"])


def ode_cases():
    z = load("ode_action_golden.npz")
    return [ode_case(z, str(n)) for n in z["names"]]


def ptime_cases():
    """Cases of ode_ptime_golden.npz (time-dependent parameters; tests/golden/make_ptime_golden.py).
    `ref` says whether the values come from the verbatim reference (trapezoid / SimpsonHermite) or
    are extension vectors of the port (euler / forwardmap: the reference's branches do not run)."""
    z = load("ode_ptime_golden.npz")
    out = []
    for n in z["names"]:
        n = str(n)
        alpha, beta, dt, ref = z[n + "/meta"]
        c = ode_case({k: (z[k] if not k.endswith("/meta") else z[k][:3]) for k in z.files if k.startswith(n + "/")}, n)
        c["dt_model"], c["ref"] = float(dt), bool(ref)
        out.append(c)
    return out


def rm_matrix_cases():
    """Cases of ode_rm_matrix_golden.npz (matrix RM, tests/golden/make_rm_matrix_golden.py)."""
    z = load("ode_rm_matrix_golden.npz")
    out = []
    for n in z["names"]:
        n = str(n)
        alpha, beta, RF0, dtm = z[n + "/meta"]
        model, disc = [str(s) for s in z[n + "/model_disc"]]
        out.append(dict(name=n, model=model, disc=disc, X0=z[n + "/X0"], P0=z[n + "/P0"], t=z[n + "/t"], Y=z[n + "/Y"],
                        stim=None, Lidx=z[n + "/Lidx"], Pidx=z[n + "/Pidx"], RM=z[n + "/RM"], RF0=float(RF0),
                        alpha=float(alpha), beta=int(beta), dt_model=None if dtm < 0 else float(dtm),
                        A=z[n + "/A"], grad=z[n + "/grad"]))
    return out


def nnet_cases():
    z = load("nnet_action_golden.npz")
    out = []
    for n in z["names"]:
        n = str(n)
        RM, RF0, alpha, beta = z[n + "/meta"]
        out.append(dict(name=n, structure=z[n + "/structure"], data_in=z[n + "/data_in"],
                        data_out=z[n + "/data_out"], X0=z[n + "/X0"], P0=z[n + "/P0"], Pidx=z[n + "/Pidx"],
                        RM=float(RM), RF0=float(RF0), alpha=float(alpha), beta=float(beta),
                        A=z[n + "/A"], grad=z[n + "/grad"]))
    return out


def nnet_rm_matrix_cases():
    """Cases of nnet_rm_matrix_golden.npz (va_nnet with RM = [RM_in, RM_out], make_rm_matrix_golden.py)."""
    z = load("nnet_rm_matrix_golden.npz")
    out = []
    for n in z["names"]:
        n = str(n)
        RF0, alpha, beta = z[n + "/meta"]
        out.append(dict(name=n, structure=z[n + "/structure"], data_in=z[n + "/data_in"], data_out=z[n + "/data_out"],
                        X0=z[n + "/X0"], P0=z[n + "/P0"], Pidx=z[n + "/Pidx"], Lidx=[z[n + "/Lin"], z[n + "/Lout"]],
                        RM=z[n + "/RM"], RF0=float(RF0), alpha=float(alpha), beta=float(beta), A=z[n + "/A"],
                        grad=z[n + "/grad"]))
    return out


def rf_matrix_cases():
    """Cases of ode_rf_matrix_golden.npz (matrix RF0 with SimpsonHermite, make_rm_matrix_golden.py)."""
    z = load("ode_rf_matrix_golden.npz")
    out = []
    for n in z["names"]:
        n = str(n)
        alpha, beta, RM = z[n + "/meta"]
        out.append(dict(name=n, model=str(z[n + "/model"][0]), disc="SimpsonHermite", X0=z[n + "/X0"], P0=z[n + "/P0"],
                        t=z[n + "/t"], Y=z[n + "/Y"], stim=None, Lidx=z[n + "/Lidx"], Pidx=z[n + "/Pidx"], RM=float(RM),
                        RF0=z[n + "/RF0"], alpha=float(alpha), beta=int(beta), dt_model=None, A=z[n + "/A"],
                        grad=z[n + "/grad"]))
    return out
