"""ctypes wrapper for the test-only kernel emulator (tests/emul/libvab_emul.so)."""
import ctypes as ct
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emul"))
import build as _build  # noqa: E402

MODEL_ID = {"lorenz96": 0, "lorenz63": 1, "nakl": 2}
DISC_ID = {"euler": 0, "trapezoid": 1, "SimpsonHermite": 2, "forwardmap": 3, "rk4": 4}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ct.CDLL(_build.build())
    return _lib


def _p(a, t=ct.c_double):
    return None if a is None else a.ctypes.data_as(ct.POINTER(t))


def emul_action_grad(prob, XP, rf_scale, RF0, active=None, tseg=0, want_grad=True):
    """prob: oracle.ode_port.OdeProblem; XP (B, n); RF0 scalar or (N-1, D) array."""
    XP = np.ascontiguousarray(np.atleast_2d(XP), dtype=np.float64)
    B, n = XP.shape
    ld = n + (n % 2)
    XPp = np.zeros((B, ld))
    XPp[:, :n] = XP
    G = np.full((B, ld), np.nan)
    A = np.zeros(B); me = np.zeros(B); fe = np.zeros(B)
    Lidx = np.ascontiguousarray(prob.Lidx, dtype=np.int32)
    Pidx = np.ascontiguousarray(prob.Pidx, dtype=np.int32)
    Y = np.ascontiguousarray(prob.Y)
    stim = None
    S = 0
    if prob.stim is not None:
        stim = np.ascontiguousarray(prob.stim.reshape(prob.N, -1))
        S = stim.shape[1]
    rm_arr = None if np.isscalar(prob.RM) else np.ascontiguousarray(prob.RM)
    rm_s = float(prob.RM) if np.isscalar(prob.RM) else 0.0
    rf_arr = None if np.isscalar(RF0) else np.ascontiguousarray(RF0, dtype=np.float64)
    rf_s = float(RF0) if np.isscalar(RF0) else 0.0
    pfix = np.ascontiguousarray(prob.P, dtype=np.float64)
    act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
    rc = lib().emul_ode_action_grad(
        MODEL_ID[prob.f.model_name], DISC_ID[prob.disc], prob.D, prob.N, prob.N_data, prob.nskip,
        prob.L, prob.NP, prob.NPest, S, ct.c_double(prob.dt), _p(Lidx, ct.c_int), _p(Pidx, ct.c_int),
        _p(Y), _p(stim), ct.c_double(rm_s), _p(rm_arr), ct.c_double(rf_s), _p(rf_arr),
        _p(pfix), ct.c_longlong(0), B, _p(XPp), ct.c_longlong(ld), ct.c_double(rf_scale),
        _p(act, ct.c_int), tseg, _p(A), _p(me), _p(fe), _p(G) if want_grad else None, ct.c_longlong(ld))
    if rc != 0:
        raise RuntimeError("emulator rc=%d" % rc)
    return A, me, fe, G[:, :n]
