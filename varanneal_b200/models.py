"""Model registry behind ``va_ode.Annealer.set_model``.

The reference lets the user pass any Python callable ``f(t, x, p)`` (va_ode.py:56-67) because
ADOL-C tapes whatever NumPy code it runs.  The device path evaluates the vector field, its
Jacobian-transpose product and its parameter derivatives inside hand-written CUDA kernels
(``csrc/ode_models.cuh``), so ``set_model`` accepts *registry entries* instead: either the name
(``'lorenz96'``, ``'lorenz63'``, ``'nakl'``) or one of the callables below.  The callables are
also valid reference-style models (same signature and vectorisation over time rows), so a user
script written for the reference keeps working when it imports its model from here, and twin
data can be generated with them.  There is no CPU fallback: an unknown callable raises.

Equations: Lorenz96 -- reference examples/Lorenz96_D20/Lorenz96_anneal.py:15-16;
NaKL -- reference tutorial notebook cell 36; Lorenz63 -- extension named by BASELINE.json.
"""
import numpy as np

MODEL_IDS = {"lorenz96": 0, "lorenz63": 1, "nakl": 2}
MODEL_NP = {"lorenz96": 1, "lorenz63": 3, "nakl": 18}
MODEL_D = {"lorenz96": None, "lorenz63": 3, "nakl": 4}


def _pstim(p):
    return (p[0], p[1]) if isinstance(p, tuple) else (p, None)


def lorenz96(t, x, p):
    """dx_i/dt = x_{i-1} (x_{i+1} - x_{i-2}) - x_i + k, indices mod D; p = [k]."""
    p, _ = _pstim(p)
    k = p[0] if np.ndim(p) >= 1 else p
    return np.roll(x, 1, 1) * (np.roll(x, -1, 1) - np.roll(x, 2, 1)) - x + k


def lorenz63(t, x, p):
    """Classic Lorenz 1963 system; p = [sigma, rho, beta]."""
    p, _ = _pstim(p)
    out = np.zeros_like(x)
    out[:, 0] = p[0] * (x[:, 1] - x[:, 0])
    out[:, 1] = x[:, 0] * (p[1] - x[:, 2]) - x[:, 1]
    out[:, 2] = x[:, 0] * x[:, 1] - p[2] * x[:, 2]
    return out


def nakl(t, x, pstim):
    """Hodgkin-Huxley Na/K/leak neuron, states (V, m, h, n), 18 parameters
    [gNa gK gL ENa EK EL | Vt Vs t1 t2 for m, h, n], injected current as stimulus."""
    p, Iext = _pstim(pstim)
    Iext = 0.0 if Iext is None else (Iext[:, 0] if np.ndim(Iext) == 2 else Iext)
    V = x[:, 0]
    out = np.zeros_like(x)
    out[:, 0] = (p[0] * x[:, 1] ** 3 * x[:, 2] * (p[3] - V) + p[1] * x[:, 3] ** 4 * (p[4] - V)
                 + p[2] * (p[5] - V) + Iext)
    for c in (1, 2, 3):
        Vt, Vs, t1, t2 = p[2 + 4 * c], p[3 + 4 * c], p[4 + 4 * c], p[5 + 4 * c]
        T = np.tanh((V - Vt) / Vs)
        out[:, c] = (0.5 * (1.0 + T) - x[:, c]) / (t1 + t2 * (1.0 - T * T))
    return out


for _f, _name in ((lorenz96, "lorenz96"), (lorenz63, "lorenz63"), (nakl, "nakl")):
    _f.vab_model = _name

REGISTRY = {"lorenz96": lorenz96, "lorenz63": lorenz63, "nakl": nakl}


def resolve(f):
    """Registry name of ``f`` (a name or a registry callable); raises for anything else."""
    if isinstance(f, str):
        if f in MODEL_IDS:
            return f
        raise ValueError("unknown model %r; registered device models: %s" % (f, sorted(MODEL_IDS)))
    name = getattr(f, "vab_model", None)
    if name in MODEL_IDS:
        return name
    raise ValueError(
        "set_model needs a registered device model (%s) -- arbitrary Python callables cannot run "
        "inside the CUDA kernels and there is no CPU fallback. Import the model from "
        "varanneal_b200.models or pass its name." % ", ".join(sorted(MODEL_IDS)))
