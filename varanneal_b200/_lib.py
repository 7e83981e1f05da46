"""ctypes binding of libvarannealb200.so (include/varanneal_b200.h).

The product has no CPU path: if the library is missing, or no B200 is visible when a context is
created, this module raises -- it never falls back to NumPy.
"""
import ctypes as ct
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvarannealb200.so")

c_double_p = ct.POINTER(ct.c_double)
c_int_p = ct.POINTER(ct.c_int32)


class OdeDesc(ct.Structure):
    _fields_ = [("model", ct.c_int32), ("disc", ct.c_int32), ("D", ct.c_int32),
                ("N_model", ct.c_int32), ("N_data", ct.c_int32), ("nskip", ct.c_int32),
                ("L", ct.c_int32), ("NP", ct.c_int32), ("NPest", ct.c_int32),
                ("n_stim", ct.c_int32), ("dt_model", ct.c_double)]


class LbfgsOpts(ct.Structure):
    _fields_ = [("m", ct.c_int32), ("maxls", ct.c_int32), ("maxfun", ct.c_int64),
                ("maxiter", ct.c_int64), ("ftol", ct.c_double), ("pgtol", ct.c_double),
                ("poll_every", ct.c_int32), ("method", ct.c_int32)]


DISC_IDS = {"euler": 0, "trapezoid": 1, "SimpsonHermite": 2, "forwardmap": 3, "rk4": 4}
ACT_IDS = {"sigmoid": 0, "tanh": 1, "linear": 2}

_VP = ct.c_void_p
_SIGS = {
    "vab_abi_version": (ct.c_int, []),
    "vab_ctx_create": (ct.c_int, [ct.c_int, _VP, ct.POINTER(_VP)]),
    "vab_ctx_destroy": (ct.c_int, [_VP]),
    "vab_sync": (ct.c_int, [_VP]),
    "vab_last_error": (ct.c_char_p, [_VP]),
    "vab_launch_count": (ct.c_longlong, [_VP]),
    "vab_graph_launch_count": (ct.c_longlong, [_VP]),
    "vab_nn_kernel_family": (ct.c_int, [_VP]),
    "vab_measure_fp64_peak": (ct.c_int, [_VP, c_double_p]),
    "vab_measure_fp64_dmma_peak": (ct.c_int, [_VP, c_double_p]),
    "vab_ozaki_gemm_probe": (ct.c_int, [_VP, ct.c_int32, ct.c_int32, ct.c_int32, ct.c_int32, ct.c_int32, ct.c_double, c_double_p]),
    "vab_ode_problem_set": (ct.c_int, [_VP, ct.POINTER(OdeDesc), c_int_p, c_int_p, _VP, _VP]),
    "vab_ode_set_weights": (ct.c_int, [_VP, ct.c_double, _VP, ct.c_double, _VP]),
    "vab_ode_set_fixed_params": (ct.c_int, [_VP, _VP, ct.c_int64]),
    "vab_ode_set_rm_matrix": (ct.c_int, [_VP, _VP]),
    "vab_ode_set_rf_matrix": (ct.c_int, [_VP, _VP]),
    "vab_ode_set_time_dependent": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64]),
    "vab_ode_action_grad": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64, ct.c_double,
                                       _VP, _VP, _VP, _VP, ct.c_int64]),
    "vab_nn_problem_set": (ct.c_int, [_VP, ct.c_int32, c_int_p, ct.c_int32, ct.c_int32,
                                      ct.c_int32, c_int_p, ct.c_int32, c_int_p, _VP, _VP,
                                      ct.c_int32, c_int_p]),
    "vab_nn_set_weights": (ct.c_int, [_VP, ct.c_double, ct.c_double, ct.c_double]),
    "vab_nn_set_rm_matrices": (ct.c_int, [_VP, _VP, _VP]),
    "vab_nn_set_fixed_params": (ct.c_int, [_VP, _VP, ct.c_int64]),
    "vab_nn_action_grad": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64, ct.c_double,
                                      _VP, _VP, _VP, _VP, ct.c_int64]),
    "vab_minimize": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64, ct.c_double,
                                ct.POINTER(LbfgsOpts), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "vab_anneal": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64, ct.c_double, c_double_p,
                              ct.c_int32, ct.POINTER(LbfgsOpts), _VP, _VP, _VP, _VP, _VP, _VP,
                              _VP]),
    "vab_copy_rows_async": (ct.c_int, [_VP, ct.c_int32, _VP, ct.c_int64, _VP, ct.c_int64, ct.c_int64, ct.c_int64, _VP]),
    "vab_set_path_sink": (ct.c_int, [_VP, _VP, ct.c_int64, ct.c_int64]),
    "vab_set_path_window": (ct.c_int, [_VP, ct.c_int64, ct.c_int64]),
    "vab_copy_rows_to_host": (ct.c_int, [_VP, _VP, ct.c_int64, _VP, ct.c_int64, ct.c_int64, ct.c_int64]),
}
EXPORTS = sorted(_SIGS)

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -m varanneal_b200.build` (nvcc, sm_100a). "
            "varanneal_b200 has no CPU fallback." % LIB_PATH)
    lib = ct.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)      # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.vab_abi_version() != 1:
        raise ImportError("libvarannealb200.so ABI version mismatch")
    _lib = lib
    return lib


class VabError(RuntimeError):
    pass


def check(rc, ctx=None):
    if rc == 0:
        return
    msg = load().vab_last_error(ctx)
    msg = msg.decode() if msg else "error %d" % rc
    if rc == -1:
        raise ValueError(msg)
    raise VabError("libvarannealb200: %s (code %d)" % (msg, rc))


def int_array(seq):
    seq = [int(v) for v in seq]
    return (ct.c_int32 * max(len(seq), 1))(*seq)


class Context(object):
    """One native context bound to one CUDA device and the current torch stream."""

    def __init__(self, device_index=0, stream_ptr=0):
        self.lib = load()
        h = _VP()
        check(self.lib.vab_ctx_create(int(device_index), _VP(stream_ptr or None), ct.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.vab_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.vab_sync(self.h), self.h)

    @property
    def launches(self):
        return int(self.lib.vab_launch_count(self.h))

    def fp64_peak_tflops(self):
        v = ct.c_double(0.0)
        check(self.lib.vab_measure_fp64_peak(self.h, ct.byref(v)), self.h)
        return float(v.value)

    def fp64_dmma_peak_tflops(self):
        v = ct.c_double(0.0)
        check(self.lib.vab_measure_fp64_dmma_peak(self.h, ct.byref(v)), self.h)
        return float(v.value)

    def ozaki_gemm_probe(self, P, M, N, K, reps=5, spread=8.0):
        out = (ct.c_double * 12)()
        check(self.lib.vab_ozaki_gemm_probe(self.h, P, M, N, K, reps, float(spread), out), self.h)
        keys = ("max_rel_err", "ms_planes", "ms_tcgen05", "ms_total", "tflops_equiv_total", "tflops_equiv_tcgen05",
                "ms_fp64_reference", "max_abs_C", "v2_max_rel_err", "v2_ms_tcgen05", "v2_tflops_equiv_tcgen05",
                "v2_tflops_equiv_total")
        return dict(zip(keys, [float(v) for v in out]))

    @property
    def graph_launches(self):
        return int(self.lib.vab_graph_launch_count(self.h))

    @property
    def nn_kernel_family(self):
        """1 fused, 2 per-layer DMMA, 3 all-layer DMMA, 4 CUDA cores, 5 tcgen05 (include/varanneal_b200.h)."""
        return int(self.lib.vab_nn_kernel_family(self.h))
