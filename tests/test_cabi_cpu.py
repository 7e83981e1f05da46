"""CPU-only checks of the boundary: the shared library loads, exports every symbol that
include/varanneal_b200.h declares, and the host layer fails loudly without a GPU (no CPU
fallback exists)."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "varanneal_b200.h")).read()
    return sorted(set(re.findall(r"VAB_API\s+[\w\s\*]+?\b(vab_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from varanneal_b200 import _lib
    lib = ct.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.EXPORTS) == names, "ctypes signatures out of sync with the header"
    assert lib.vab_abi_version() == 1


def test_library_has_no_torch_or_cudart_dependency():
    import subprocess
    from varanneal_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out


def test_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from varanneal_b200 import _lib
    lib = _lib.load()
    h = ct.c_void_p()
    rc = lib.vab_ctx_create(0, None, ct.byref(h))
    assert rc != 0
    assert b"no CUDA device" in lib.vab_last_error(None) or b"CPU" in lib.vab_last_error(None)


def test_annealer_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    an.set_model("lorenz96", 20)
    an.set_data(np.zeros((11, 3)), t=0.1 * np.arange(11))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        an.anneal_init(np.zeros((11, 20)), np.array([8.0]), 1.5, [0], 1.0, 1e-3, [0, 1, 2], [0])


def test_model_registry_rejects_arbitrary_callables():
    from varanneal_b200 import models, va_ode
    assert models.resolve("lorenz96") == "lorenz96"
    assert models.resolve(models.nakl) == "nakl"
    with pytest.raises(ValueError, match="registered device model"):
        va_ode.Annealer().set_model(lambda t, x, p: x, 3)
    with pytest.raises(ValueError):
        va_ode.Annealer().set_model("lorenz63", 5)


def test_set_data_layouts_match_reference():
    """va_ode.py:98-124: time in column 0 unless t is given; nstart / N window."""
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    data = np.column_stack([0.5 * np.arange(10), np.arange(10) * 1.0, np.arange(10) * 2.0])
    an.set_data(data, nstart=2, N=5)
    assert an.N_data == 5 and an.Y.shape == (5, 2) and an.t_data[0] == 1.0 and an.dt_data == 0.5
    an.set_data(data[:, 1:], t=data[:, 0])
    assert an.N_data == 10 and an.Y.shape == (10, 2)


def test_registry_models_are_reference_style_callables():
    from oracle.models_np import MODELS
    from varanneal_b200 import models
    rng = np.random.RandomState(0)
    x = rng.randn(7, 20)
    assert np.allclose(models.lorenz96(None, x, [8.17]), MODELS["lorenz96"](None, x, [8.17]))
    x3 = rng.randn(5, 3)
    assert np.allclose(models.lorenz63(None, x3, [10, 28, 8 / 3]), MODELS["lorenz63"](None, x3, [10, 28, 8 / 3]))
