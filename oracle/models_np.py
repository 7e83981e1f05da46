"""TEST INFRASTRUCTURE ONLY -- NumPy vector fields for the oracle.

Each model is written as the reference expects a user model to be written
(``f(t, x, p)`` or ``f(t, x, (p, stim))`` acting on a whole slab of time rows,
va_ode.py:56-67) so the *same callable* can be handed to the verbatim reference through
``oracle.ref_shim``.  In addition every model carries the two adjoint products the analytic
gradient needs:

    jtv(x, v, p, stim)  ->  J_f(x)^T v      row-wise, shape (N', D)
    ptv(x, v, p, stim)  ->  sum_rows (df/dp)^T v, shape (NP,)
    ptv_rows(x, v, p, stim) -> the same product row by row, shape (N', NP)  (time-dependent
                            parameters: ``p`` is then (N', NP), one parameter row per time row,
                            va_ode.py:170-188)

Sources of the model equations:
  * Lorenz96 : examples/Lorenz96_D20/Lorenz96_anneal.py:15-16
  * NaKL     : examples/jupyter-tutorial/VarAnneal_tutorial.ipynb cell 36 (raw lines 2919-2949)
  * Lorenz63 : not in the reference (BASELINE.json north_star asks for it); textbook form.
All arithmetic is dtype-generic so complex-step differentiation passes through.
"""
import numpy as np


def _split_pstim(p):
    if isinstance(p, tuple):
        return p[0], p[1]
    return p, None


def _pc(p, k):
    """Parameter k: a scalar for static parameters, one value per time row for a (rows, NP) series."""
    return p[k] if np.ndim(p) == 1 else p[:, k]


# --------------------------------------------------------------------------- Lorenz 96
def lorenz96(t, x, p):
    p, _ = _split_pstim(p)
    k = p if np.ndim(p) == 0 else (p[0] if np.ndim(p) == 1 else p[:, 0:1])
    return np.roll(x, 1, 1) * (np.roll(x, -1, 1) - np.roll(x, 2, 1)) - x + k


def _l96_jtv(x, v, p, stim=None):
    # f_i = x_{i-1}(x_{i+1} - x_{i-2}) - x_i + k
    # (J^T v)_j = v_{j+1}(x_{j+2} - x_{j-1}) + v_{j-1} x_{j-2} - v_{j+2} x_{j+1} - v_j
    r = lambda a, s: np.roll(a, s, 1)  # noqa: E731   r(a, 1)[j] = a[j-1]
    return (r(v, -1) * (r(x, -2) - r(x, 1)) + r(v, 1) * r(x, 2)
            - r(v, -2) * r(x, -1) - v)


def _l96_ptv_rows(x, v, p, stim=None):
    return np.sum(v, axis=1)[:, None]


def _l96_ptv(x, v, p, stim=None):
    return np.sum(_l96_ptv_rows(x, v, p, stim), axis=0)


lorenz96.jtv = _l96_jtv
lorenz96.ptv = _l96_ptv
lorenz96.ptv_rows = _l96_ptv_rows
lorenz96.NP = 1
lorenz96.model_name = "lorenz96"


# --------------------------------------------------------------------------- Lorenz 63
def lorenz63(t, x, p):
    p, _ = _split_pstim(p)
    s, r, b = _pc(p, 0), _pc(p, 1), _pc(p, 2)
    out = np.zeros_like(x)
    out[:, 0] = s * (x[:, 1] - x[:, 0])
    out[:, 1] = x[:, 0] * (r - x[:, 2]) - x[:, 1]
    out[:, 2] = x[:, 0] * x[:, 1] - b * x[:, 2]
    return out


def _l63_jtv(x, v, p, stim=None):
    s, r, b = _pc(p, 0), _pc(p, 1), _pc(p, 2)
    out = np.zeros_like(v)
    out[:, 0] = -s * v[:, 0] + (r - x[:, 2]) * v[:, 1] + x[:, 1] * v[:, 2]
    out[:, 1] = s * v[:, 0] - v[:, 1] + x[:, 0] * v[:, 2]
    out[:, 2] = -x[:, 0] * v[:, 1] - b * v[:, 2]
    return out


def _l63_ptv_rows(x, v, p, stim=None):
    return np.stack([(x[:, 1] - x[:, 0]) * v[:, 0], x[:, 0] * v[:, 1], -x[:, 2] * v[:, 2]], axis=1)


def _l63_ptv(x, v, p, stim=None):
    return np.sum(_l63_ptv_rows(x, v, p, stim), axis=0)


lorenz63.jtv = _l63_jtv
lorenz63.ptv = _l63_ptv
lorenz63.ptv_rows = _l63_ptv_rows
lorenz63.NP = 3
lorenz63.model_name = "lorenz63"


# --------------------------------------------------------------------------- NaKL
def _gate(V, z, Vt, Vs, t1, t2):
    a = (V - Vt) / Vs
    T = np.tanh(a)
    zinf = 0.5 * (1.0 + T)
    tau = t1 + t2 * (1.0 - T * T)
    return a, T, zinf, tau


def _cols(p):
    """p[k] -> parameter k for both layouts: (NP,) static, (rows, NP) time series."""
    return p if np.ndim(p) == 1 else np.asarray(p).T


def nakl(t, x, pstim):
    p, Iext = _split_pstim(pstim)
    p = _cols(p)
    if Iext is None:
        Iext = 0.0
    else:
        Iext = np.asarray(Iext)
        if Iext.ndim == 2:
            Iext = Iext[:, 0]
    V, m, h, n = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
    out = np.zeros_like(x)
    out[:, 0] = (p[0] * m ** 3 * h * (p[3] - V) + p[1] * n ** 4 * (p[4] - V)
                 + p[2] * (p[5] - V) + Iext)
    for c, z in ((1, m), (2, h), (3, n)):
        Vt, Vs, t1, t2 = p[2 + 4 * c], p[3 + 4 * c], p[4 + 4 * c], p[5 + 4 * c]
        _, _, zinf, tau = _gate(V, z, Vt, Vs, t1, t2)
        out[:, c] = (zinf - z) / tau
    return out


def _nakl_partials(x, p):
    """Per-row partial derivatives shared by jtv and ptv."""
    V, m, h, n = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
    d = {}
    d["dV_V"] = -p[0] * m ** 3 * h - p[1] * n ** 4 - p[2]
    d["dV_m"] = 3.0 * p[0] * m * m * h * (p[3] - V)
    d["dV_h"] = p[0] * m ** 3 * (p[3] - V)
    d["dV_n"] = 4.0 * p[1] * n ** 3 * (p[4] - V)
    for c, z in ((1, m), (2, h), (3, n)):
        Vt, Vs, t1, t2 = p[2 + 4 * c], p[3 + 4 * c], p[4 + 4 * c], p[5 + 4 * c]
        a, T, zinf, tau = _gate(V, z, Vt, Vs, t1, t2)
        sech2 = 1.0 - T * T
        zinf_a = 0.5 * sech2
        tau_a = -2.0 * t2 * T * sech2
        dz_a = (zinf_a * tau - (zinf - z) * tau_a) / (tau * tau)
        d[("a", c)] = (a, dz_a, Vs)
        d[("z", c)] = -1.0 / tau
        d[("t1", c)] = -(zinf - z) / (tau * tau)
        d[("t2", c)] = -(zinf - z) * sech2 / (tau * tau)
    return d


def _nakl_jtv(x, v, pstim, stim=None):
    p, _ = _split_pstim(pstim)
    p = _cols(p)
    d = _nakl_partials(x, p)
    out = np.zeros_like(v)
    out[:, 0] = d["dV_V"] * v[:, 0]
    out[:, 1] = d["dV_m"] * v[:, 0]
    out[:, 2] = d["dV_h"] * v[:, 0]
    out[:, 3] = d["dV_n"] * v[:, 0]
    for c in (1, 2, 3):
        a, dz_a, Vs = d[("a", c)]
        out[:, 0] = out[:, 0] + (dz_a / Vs) * v[:, c]
        out[:, c] = out[:, c] + d[("z", c)] * v[:, c]
    return out


def _nakl_ptv_rows(x, v, pstim, stim=None):
    p, _ = _split_pstim(pstim)
    p = _cols(p)
    V, m, h, n = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
    d = _nakl_partials(x, p)
    g = np.zeros((x.shape[0], 18), dtype=v.dtype)
    g[:, 0] = m ** 3 * h * (p[3] - V) * v[:, 0]
    g[:, 1] = n ** 4 * (p[4] - V) * v[:, 0]
    g[:, 2] = (p[5] - V) * v[:, 0]
    g[:, 3] = p[0] * m ** 3 * h * v[:, 0]
    g[:, 4] = p[1] * n ** 4 * v[:, 0]
    g[:, 5] = p[2] * v[:, 0]
    for c in (1, 2, 3):
        a, dz_a, Vs = d[("a", c)]
        g[:, 2 + 4 * c] = dz_a * (-1.0 / Vs) * v[:, c]
        g[:, 3 + 4 * c] = dz_a * (-a / Vs) * v[:, c]
        g[:, 4 + 4 * c] = d[("t1", c)] * v[:, c]
        g[:, 5 + 4 * c] = d[("t2", c)] * v[:, c]
    return g


def _nakl_ptv(x, v, pstim, stim=None):
    return np.sum(_nakl_ptv_rows(x, v, pstim, stim), axis=0)


nakl.jtv = _nakl_jtv
nakl.ptv = _nakl_ptv
nakl.ptv_rows = _nakl_ptv_rows
nakl.NP = 18
nakl.model_name = "nakl"

MODELS = {"lorenz96": lorenz96, "lorenz63": lorenz63, "nakl": nakl}
