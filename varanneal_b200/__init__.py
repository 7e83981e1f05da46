"""varanneal_b200 -- B200-native drop-in for the annealing hot path of paulrozdeba/varanneal.

Import surface mirrors the reference package (varanneal/__init__.py:1-2):
``from varanneal_b200 import va_ode`` / ``va_nnet``; each exposes ``Annealer``.
The numeric work is done by libvarannealb200.so (hand-written sm_100a CUDA behind the C ABI of
include/varanneal_b200.h); there is no CPU fallback.
"""
from . import models  # noqa: F401
from . import va_ode  # noqa: F401
from . import va_nnet  # noqa: F401

__version__ = "0.1.0"
