"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference source.

The reference (paulrozdeba/varanneal) is Python 2 and imports ``adolc``; neither exists in
this image.  This module reads the three reference files as text from ``/root/reference``
(never copied into this repo), applies the purely syntactic py2->py3 substitutions listed in
SURVEY.md App. C (plus ``/`` -> ``//`` in the one loop bound that relies on Python 2's integer
division, ``va_ode.py:215``), and ``exec``s them into fresh module objects with a stub ``adolc`` module.
The reference's own ``Annealer.anneal_init / A_gaussian / me_gaussian / fe_gaussian /
disc_* / anneal_step / save_*`` then run verbatim under NumPy 2.x.

Only two methods need replacing because they call ADOL-C
(``_autodiffmin.py:32-49`` ``tape_A`` and ``:57-58`` ``A_gradA_taped``): ``ShimOde`` /
``ShimNnet`` override them with a caller-supplied gradient (complex-step through the
reference's own action, or the validated NumPy adjoint of ``oracle.ode_port``).

``/root/reference`` exists only in the build container, not on the GPU box: nothing that runs
under ``-m gpu``, ``smoke()`` or ``bench.py`` may import this module.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that pin
``oracle.ode_port`` / ``oracle.nnet_port`` against the reference (skipped when the reference
tree is absent).
"""
import os
import re
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("VARANNEAL_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "varanneal", "va_ode.py"))


def _stub_adolc():
    """Module object standing in for pyadolc: only ``isinstance`` targets are needed
    (va_ode.py:752,763; va_nnet.py:500)."""
    mod = types.ModuleType("adolc")
    inner = types.ModuleType("adolc._adolc")

    class adouble(object):  # noqa: N801  (name fixed by the reference's isinstance checks)
        pass

    class adub(object):  # noqa: N801
        pass

    inner.adouble = adouble
    inner.adub = adub
    mod._adolc = inner
    return mod


_EXEC_RE = re.compile(r"exec '(self\.\w+) = self\.(\w*)%s'%\((\w+),?\)")


def _py3(src):
    src = _EXEC_RE.sub(r"exec('\1 = self.\2%s'%(\3,))", src)
    src = src.replace(".im_func.", ".__func__.")
    src = src.replace("xrange((self.N_model - 1) / 2)", "xrange((self.N_model - 1) // 2)")   # py2 integer division (va_ode.py:215)
    src = src.replace("xrange", "range")
    return src


_cache = {}


def load_reference():
    """Returns (admin_module, va_ode_module, va_nnet_module) built from the reference text."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    saved = sys.modules.get("adolc")
    sys.modules["adolc"] = _stub_adolc()
    try:
        mods = []
        for name in ("_autodiffmin", "va_ode", "va_nnet"):
            path = os.path.join(REFERENCE_ROOT, "varanneal", name + ".py")
            with open(path, "r") as fh:
                src = _py3(fh.read())
            m = types.ModuleType("refshim_" + name)
            m.__file__ = path
            if name != "_autodiffmin":
                src = src.replace("from _autodiffmin import ADmin", "")
                m.ADmin = mods[0].ADmin
                if name == "va_nnet":
                    m.sys = sys  # va_nnet.py:492 uses sys without importing it
            exec(compile(src, path, "exec"), m.__dict__)
            mods.append(m)
    finally:
        if saved is None:
            del sys.modules["adolc"]
        else:
            sys.modules["adolc"] = saved
    _cache["mods"] = tuple(mods)
    return _cache["mods"]


def complex_step_grad(fun, xp, h=1e-30):
    """Gradient of a real-analytic scalar function by complex-step differentiation;
    exact to rounding (no subtractive cancellation).  O(n) evaluations: small n only."""
    xp = np.asarray(xp, dtype=np.float64)
    g = np.empty_like(xp)
    z = xp.astype(np.complex128)
    for j in range(xp.size):
        z[j] = complex(xp[j], h)
        g[j] = np.imag(fun(z)) / h
        z[j] = xp[j]
    return g


def _quiet(fn):
    def wrapped(*a, **k):
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(*a, **k)
    return wrapped


def make_shim_classes():
    """Subclasses of the reference Annealers with the two ADOL-C methods replaced."""
    _, va_ode, va_nnet = load_reference()

    class ShimOde(va_ode.Annealer):
        grad_fn = None  # callable(self, XP) -> gradient; default complex-step

        def tape_A(self, xtrace):
            self.taped = True

        def A_gradA_taped(self, XP):
            XP = np.asarray(XP, dtype=np.float64)
            if self.grad_fn is None:
                g = complex_step_grad(self.A, XP)
            else:
                g = self.grad_fn(XP)
            return float(self.A(XP)), g

        anneal_quiet = _quiet(va_ode.Annealer.anneal)
        anneal_step_quiet = _quiet(va_ode.Annealer.anneal_step)

    class ShimNnet(va_nnet.Annealer):
        grad_fn = None

        def tape_A(self, xtrace):
            self.taped = True

        def A_gradA_taped(self, XP):
            XP = np.asarray(XP, dtype=np.float64)
            if self.grad_fn is None:
                g = complex_step_grad(self.A, XP)
            else:
                g = self.grad_fn(XP)
            return float(self.A(XP)), g

        anneal_quiet = _quiet(va_nnet.Annealer.anneal)
        anneal_step_quiet = _quiet(va_nnet.Annealer.anneal_step)

    return ShimOde, ShimNnet
