"""``va_ode.Annealer`` -- drop-in for the reference's ODE annealer (varanneal/va_ode.py:43-905).

Same class name, method names, positional order and defaults as the reference; the numeric body
(ADOL-C tape of ``A_gaussian`` + SciPy L-BFGS-B) is replaced by libvarannealb200.so.  Host code
here only normalises arguments the way ``anneal_init`` does (va_ode.py:531-705), owns the device
tensors through PyTorch, and lays the results out exactly like the reference
(``minpaths``, ``A_array``, ``me_array``, ``fe_array``, ``P``; savers at va_ode.py:794-889).

Differences from the reference, all deliberate:
  * ``set_model(f, D)``: ``f`` must be a registered device model (name or callable from
    ``varanneal_b200.models``); arbitrary callables raise -- there is no CPU fallback.
  * errors raise ``ValueError`` instead of ``print`` + ``sys.exit(1)``.
  * ``exitflags`` is filled with the minimiser's status (the reference never writes it, B3).
  * ``set_data_fromfile`` works (the reference's is broken, B5).
  * ``method``: 'L-BFGS-B', 'NCG' and 'TNC' run on the device ('TNC' rung by rung; bounds honoured by
    'L-BFGS-B' as SciPy does and by 'TNC' through an active set);
    'LM' is dead code in the reference (B10).
  * batches: ``X0`` of shape (B, N, D) [+ ``P0`` (B, NP) or (NP,)] anneals B independent
    initialisations concurrently; every result array gains a leading B axis.  With the
    reference's shapes the results have exactly the reference's shapes.
  * time-dependent parameters: ``P0`` of shape (N_model, NP) with a 2-D ``X0`` (va_ode.py:568-570),
    or (B, N_model, NP) with a batch.  ``XP = X.flatten() ++ P[:, Pidx].flatten()`` and row n of P
    enters f at model time n, as the reference's trapezoid / SimpsonHermite branches do
    (va_ode.py:170-188, 368-369, 416-418).  euler / forwardmap -- whose reference branches fail on a
    doubly dropped row -- use the same N_model rows and never read the last one; rk4 and rows wider
    than one lane group (lorenz96 D > 128) raise.
"""
import ctypes as ct
import time

import numpy as np

from . import _lib
from . import models as _models
from ._devicemin import DeviceMin, ptr

_DISCS = ("euler", "trapezoid", "SimpsonHermite", "forwardmap", "rk4")


class Annealer(DeviceMin):
    def __init__(self, device=None, verbose=False):
        self.taped = False
        self.annealing_initialized = False     # never set True, as in the reference (App. B2)
        self._device_arg = device
        self.verbose = verbose
        self.stim = None
        self._ptime = False                    # parameters are a time series (set by anneal_init)

    # ------------------------------------------------------------------ problem definition
    def set_model(self, f, D):
        """va_ode.py:56-67.  ``f``: 'lorenz96' | 'lorenz63' | 'nakl' or the matching callable
        from ``varanneal_b200.models``."""
        self.model_name = _models.resolve(f)
        self.f = _models.REGISTRY[self.model_name]
        fixed_D = _models.MODEL_D[self.model_name]
        if fixed_D is not None and int(D) != fixed_D:
            raise ValueError("model %s has D = %d" % (self.model_name, fixed_D))
        self.D = int(D)

    def set_data_fromfile(self, data_file, stim_file=None, nstart=0, N=None):
        """va_ode.py:69-96 (signature kept; loads then defers to set_data)."""
        load = lambda fn: np.load(fn) if fn.endswith("npy") else np.loadtxt(fn)  # noqa: E731
        data = load(data_file)
        stim = None if stim_file is None else load(stim_file)
        self.set_data(data, stim=stim, nstart=nstart, N=N)

    def set_data(self, data, stim=None, t=None, nstart=0, N=None):
        """va_ode.py:98-124: time in column 0 unless ``t`` is given."""
        data = np.asarray(data)
        self.N_data = data.shape[0] if N is None else int(N)
        sl = slice(nstart, nstart + self.N_data)
        if t is None:
            self.t_data = np.array(data[sl, 0], dtype=np.float64)
            self.Y = np.array(data[sl, 1:], dtype=np.float64)
            self.stim = None if stim is None else np.array(np.asarray(stim)[sl, 1:], dtype=np.float64)
        else:
            self.t_data = np.array(np.asarray(t)[sl], dtype=np.float64)
            self.Y = np.array(data[sl], dtype=np.float64)
            self.stim = None if stim is None else np.array(np.asarray(stim)[sl], dtype=np.float64)
        if self.Y.ndim == 1:
            self.Y = self.Y.reshape(-1, 1)
        self.N_data = self.Y.shape[0]
        self.dt_data = self.t_data[1] - self.t_data[0]

    # ------------------------------------------------------------------ annealing
    def anneal(self, X0, P0, alpha, beta_array, RM, RF0, Lidx, Pidx, dt_model=None,
               init_to_data=True, action='A_gaussian', disc='trapezoid',
               method='L-BFGS-B', bounds=None, opt_args=None, adolcID=0,
               track_paths=None, track_params=None, track_action_errors=None):
        """Full annealing run over ``beta_array`` (va_ode.py:459-528)."""
        if not self.annealing_initialized:
            self.anneal_init(X0, P0, alpha, beta_array, RM, RF0, Lidx, Pidx, dt_model,
                             init_to_data, action, disc, method, bounds, opt_args, adolcID)
        tracked = not (track_paths is None and track_params is None and track_action_errors is None)
        if not tracked and not self.verbose and self.betaidx == 0 and self.method != 'TNC':
            self._anneal_device()           # the whole ladder in native calls, one per wave of paths
            return
        self._require_resident("anneal() with track_* / verbose / method='TNC'")
        for _ in range(self.Nbeta):
            if self.verbose:
                print('------------------------------')
                print('Step %d of %d' % (self.betaidx + 1, self.Nbeta))
                rf = self.RF if np.isscalar(self.RF) else np.asarray(self.RF).flat[0]
                print('beta = %d, RF = %.8e' % (self.beta, rf))
            self.anneal_step()
            if track_paths is not None:
                self.save_paths(track_paths['filename'], track_paths.get('dtype', np.float64),
                                track_paths.get('fmt', "%.8e"))
            if track_params is not None:
                self.save_params(track_params['filename'], track_params.get('dtype', np.float64),
                                 track_params.get('fmt', "%.8e"))
            if track_action_errors is not None:
                self.save_action_errors(track_action_errors['filename'],
                                        track_action_errors.get('cmpt', 0),
                                        track_action_errors.get('dtype', np.float64),
                                        track_action_errors.get('fmt', "%.8e"))

    def anneal_init(self, X0, P0, alpha, beta_array, RM, RF0, Lidx, Pidx, dt_model=None,
                    init_to_data=True, action='A_gaussian', disc='trapezoid',
                    method='L-BFGS-B', bounds=None, opt_args=None, adolcID=0):
        """Argument normalisation of va_ode.py:531-705, then problem upload to the device."""
        if method == 'LM':
            raise ValueError("method='LM' is dead code in the reference (SURVEY.md App. B10)")
        if method not in ('L-BFGS-B', 'NCG', 'TNC'):
            raise ValueError("Optimization routine not recognized: %r" % (method,))
        self.method = method
        if action != 'A_gaussian':
            raise ValueError("only action='A_gaussian' exists (va_ode.py:130-136)")
        if disc not in _DISCS:
            raise ValueError("unknown disc %r; choose from %s" % (disc, _DISCS))
        self.disc_name = disc

        # time grids (va_ode.py:544-558)
        if dt_model is not None and dt_model != self.dt_data and self.stim is not None:
            raise ValueError("Separate dt_data and dt_model currently not supported with an "
                             "external stimulus.")
        if dt_model is None:
            self.dt_model = float(self.dt_data)
            self.N_model = self.N_data
            self.merr_nskip = 1
            self.t_model = np.copy(self.t_data)
        else:
            self.dt_model = float(dt_model)
            ratio = self.dt_data / self.dt_model
            self.merr_nskip = int(round(ratio))
            if self.merr_nskip < 1 or abs(ratio - self.merr_nskip) > 1e-6 * ratio:
                raise ValueError("dt_data must be an integer multiple of dt_model")
            self.N_model = (self.N_data - 1) * self.merr_nskip + 1
            self.t_model = np.linspace(self.t_data[0], self.t_data[-1], self.N_model)
        if disc == 'SimpsonHermite' and self.N_model % 2 == 0:
            raise ValueError("disc='SimpsonHermite' needs an odd number of model time points "
                             "(va_ode.py:430-435), got N_model = %d" % self.N_model)
        if disc == 'rk4' and self.stim is not None:
            raise ValueError("disc='rk4' (extension) does not take a stimulus")

        self.opt_args = opt_args
        N, D = self.N_model, self.D

        # batch detection (extension): X0 (B, N, D), or a callable X0(b0, b1) -> (b1 - b0, N, D)
        # (NumPy array or CUDA tensor) that produces the initial paths of a block on demand, for
        # batches whose initial paths do not fit host memory (C3: 0.8 GB each); P0 (B, NP) then
        # defines the batch size
        P0 = np.asarray(P0, dtype=np.float64) if not isinstance(P0, np.ndarray) else P0
        lazy = callable(X0)
        self._ptime = False
        if lazy:
            if P0.ndim != 2:
                raise ValueError("a callable X0 needs P0 of shape (B, NP)")
            self._X0_fn, self._init_to_data = X0, bool(init_to_data)
            X0 = np.empty((P0.shape[0], 0, 0))
        else:
            self._X0_fn = None
            X0 = np.asarray(X0) if not isinstance(X0, np.ndarray) else X0
        self.batched = (X0.ndim == 3)
        if lazy:
            B = P0.shape[0]
        elif self.batched:
            B = X0.shape[0]
            if X0.shape[1:] != (N, D):
                raise ValueError("X0 must have shape (B, %d, %d)" % (N, D))
            if P0.ndim == 1:
                P0 = np.tile(P0, (B, 1))
            if P0.ndim == 3:
                self._ptime = True                       # (B, N_model, NP): a time series per path
                if P0.shape[:2] != (B, N):
                    raise ValueError("batched anneal with time-dependent parameters: P0 must be "
                                     "(B, N_model, NP) = (%d, %d, NP)" % (B, N))
            elif P0.ndim != 2 or P0.shape[0] != B:
                raise ValueError("batched anneal: P0 must be (NP,), (B, NP) or (B, N_model, NP)")
        else:
            B = 1
            if X0.shape != (N, D):
                raise ValueError("X0 must have shape (%d, %d)" % (N, D))
            if P0.ndim == 2:
                self._ptime = True                       # va_ode.py:568-570
                if P0.shape[0] != N:
                    raise ValueError("time-dependent parameters: P0 must have one row per model time "
                                     "point, (%d, NP)" % N)
            elif P0.ndim != 1:
                raise ValueError("P0 must be (NP,) or (N_model, NP)")
        if self._ptime and disc == 'rk4':
            raise ValueError("disc='rk4' (extension) takes static parameters only")
        self.P = np.array(P0, dtype=np.float64)          # (NP,) / (N, NP) [batch: leading B]; updated per beta
        self.NP = self.P.shape[-1]
        if self.NP != _models.MODEL_NP[self.model_name]:
            raise ValueError("model %s takes %d parameters, P0 has %d"
                             % (self.model_name, _models.MODEL_NP[self.model_name], self.NP))
        self.Pidx = np.asarray(Pidx, dtype=np.int64).reshape(-1)
        self.NPest = len(self.Pidx)
        self.Lidx = np.asarray(Lidx, dtype=np.int64).reshape(-1)
        self.L = len(self.Lidx)
        if self.L != self.Y.shape[1]:
            raise ValueError("data has %d measured columns but len(Lidx) = %d" % (self.Y.shape[1], self.L))

        nX = N * D
        NPl, Pidxl, NPestl, _ = self._lay()
        n = nX + NPestl
        self._nX = nX

        # bounds (va_ode.py:582-605): D + NPest [lo, hi] pairs -> one pair per unknown
        if bounds is not None:
            bounds = list(bounds)
            if len(bounds) != D + self.NPest:
                raise ValueError("bounds must hold D + len(Pidx) = %d [lo, hi] pairs" % (D + self.NPest))
            lohi = np.array([[-np.inf if b[0] is None else b[0], np.inf if b[1] is None else b[1]]
                             for b in bounds], dtype=np.float64)
            nrow = N if self._ptime else 1            # parameter rows (va_ode.py:592-605)
            self.bounds = [list(b) for b in bounds[:D]] * N + [list(b) for b in bounds[D:]] * nrow
            lo = np.concatenate([np.tile(lohi[:D, 0], N), np.tile(lohi[D:, 0], nrow)])
            hi = np.concatenate([np.tile(lohi[:D, 1], N), np.tile(lohi[D:, 1], nrow)])
        else:
            self.bounds = None
            lo = hi = None

        # RM / RF0 (va_ode.py:612-640)
        if isinstance(RM, list):
            RM = np.array(RM)
        if isinstance(RM, np.ndarray) and RM.ndim > 0:
            if RM.shape == (self.L,):
                self.RM = np.resize(RM, (self.N_data, self.L)).astype(np.float64)
            elif RM.shape == (self.L, self.L):
                self.RM = np.resize(RM, (self.N_data, self.L, self.L)).astype(np.float64)   # va_ode.py:616-617
            elif RM.shape in ((self.N_data, self.L), (self.N_data, self.L, self.L)):
                self.RM = RM.astype(np.float64)
            else:
                raise ValueError("RM must be a scalar, (L,), (N_data, L), (L, L) or (N_data, L, L)")
        else:
            self.RM = float(RM)
        if isinstance(RF0, list):
            RF0 = np.array(RF0)
        if isinstance(RF0, np.ndarray) and RF0.ndim > 0:
            if RF0.shape == (D,):
                self.RF0 = np.resize(RF0, (N - 1, D)).astype(np.float64)
            elif RF0.shape == (D, D):
                self.RF0 = np.resize(RF0, (N - 1, D, D)).astype(np.float64)          # va_ode.py:631-632
            elif RF0.shape in ((N - 1, D), (N - 1, D, D)):
                self.RF0 = RF0.astype(np.float64)
            else:
                raise ValueError("RF0 must be a scalar, (D,), (N_model-1, D), (D, D) or (N_model-1, D, D)")
            if self.RF0.ndim == 3:
                if disc != 'SimpsonHermite':
                    raise ValueError("a matrix RF0 works with disc='SimpsonHermite' only: the reference's branch "
                                     "for the other discretisations does not run (va_ode.py:222)")
                if self._ptime:
                    raise ValueError("a matrix RF0 cannot be combined with time-dependent parameters")
        else:
            self.RF0 = float(RF0)

        # beta ladder (va_ode.py:643-650), with the reference's uint16 truncation (App. B1)
        self.alpha = alpha
        self.beta_array = np.array(beta_array, dtype=np.uint16)
        self.Nbeta = len(self.beta_array)
        self.betaidx = 0
        self.beta = self.beta_array[0]
        self.RF = self.RF0 * self.alpha ** float(self.beta)

        # ---- device side: context first (the wave plan needs the free memory of the device)
        ctx = self._open_context(self._device_arg)
        self._Btot = B
        Bw = self._plan_wave(B, n, self.Nbeta)
        self._alloc_results(B, self.Nbeta, nX, self.batched)

        # initialise observed components to the data (va_ode.py:677-678): on a float64 working
        # copy, then written back into the caller's array, which the reference mutates in place
        self._Xw = None
        if not lazy:
            Xw = np.array(X0, dtype=np.float64).reshape(B, N, D)
            if init_to_data:
                Xw[:, ::self.merr_nskip, self.Lidx] = self.Y
                if X0.flags.writeable:
                    X0[...] = Xw.reshape(X0.shape)
            if self.keep_paths == 'all':         # the reference parks XP0 in minpaths[0] (va_ode.py:667)
                self.minpaths.reshape(B, self.Nbeta, nX + NPl)[:, 0] = np.concatenate(
                    [Xw.reshape(B, nX), self.P.reshape(B, NPl)], axis=1)
            else:
                self._Xw = Xw.reshape(B, nX)
        self.adolcID = adolcID               # accepted and ignored: nothing is taped
        self._alloc_paths(Bw, n)
        self._Y_dev = self._to_dev(self.Y)
        self._stim_dev = None
        n_stim = 0
        if self.stim is not None:
            st = np.asarray(self.stim, dtype=np.float64).reshape(self.N_data, -1)
            n_stim = st.shape[1]
            self._stim_dev = self._to_dev(st)
        desc = _lib.OdeDesc(_models.MODEL_IDS[self.model_name], _lib.DISC_IDS[disc], D, N,
                            self.N_data, self.merr_nskip, self.L, self.NP, self.NPest, n_stim,
                            self.dt_model)
        _lib.check(ctx.lib.vab_ode_problem_set(
            ctx.h, desc, _lib.int_array(self.Lidx), _lib.int_array(self.Pidx),
            ptr(self._Y_dev), ptr(self._stim_dev)), ctx.h)
        rm_matrix = (not np.isscalar(self.RM)) and self.RM.ndim == 3          # va_ode.py:149-152
        self._rm_dev = None if np.isscalar(self.RM) else self._to_dev(self.RM)
        self._rf0_dev = None if np.isscalar(self.RF0) else self._to_dev(self.RF0)
        rf_matrix = (not np.isscalar(self.RF0)) and self.RF0.ndim == 3         # va_ode.py:211-218
        _lib.check(ctx.lib.vab_ode_set_weights(
            ctx.h, self.RM if np.isscalar(self.RM) else 0.0, None if rm_matrix else ptr(self._rm_dev),
            self.RF0 if np.isscalar(self.RF0) else 1.0, None if rf_matrix else ptr(self._rf0_dev)), ctx.h)
        if rm_matrix:
            _lib.check(ctx.lib.vab_ode_set_rm_matrix(ctx.h, ptr(self._rm_dev)), ctx.h)
        if rf_matrix:
            _lib.check(ctx.lib.vab_ode_set_rf_matrix(ctx.h, ptr(self._rf0_dev)), ctx.h)
        self._pfix_dev = self._to_dev(self.P.reshape(B, NPl)[:Bw])
        if self._ptime:
            _lib.check(ctx.lib.vab_ode_set_time_dependent(ctx.h, 1, ptr(self._pfix_dev), NPl), ctx.h)
        else:
            _lib.check(ctx.lib.vab_ode_set_fixed_params(ctx.h, ptr(self._pfix_dev), NPl), ctx.h)
        self._lo_dev = self._hi_dev = None
        if lo is not None:
            pad = self._ld - n
            self._lo_dev = self._to_dev(np.concatenate([lo, np.full(pad, -np.inf)]))
            self._hi_dev = self._to_dev(np.concatenate([hi, np.full(pad, np.inf)]))
        self._dev_paths_current = False
        self.initalized = True               # sic (va_ode.py:705)

    def _lay(self):
        """Parameter block of a path row: (NP,) static, or the flattened (N_model, NP) time series
        with the estimated entries P[:, Pidx] in row-major order (va_ode.py:688-689)."""
        if not self._ptime:
            return self.NP, self.Pidx, self.NPest, (self.NP,)
        N = self.N_model
        idx = (np.arange(N)[:, None] * self.NP + self.Pidx[None, :]).reshape(-1)
        return N * self.NP, idx, N * self.NPest, (N, self.NP)

    def _action_grad_native(self, rf_scale, b0=0, nb=None):
        """Evaluates paths [b0, b0 + nb) of the device batch (default: all)."""
        ctx = self._ctx
        nb = self._B - b0 if nb is None else nb
        sub = (b0 != 0 or nb != self._B)
        off = b0 * self._ld * 8

        def at(t, per_path):
            return ct.c_void_p(t.data_ptr() + b0 * per_path)

        NPl = self._lay()[0]
        if sub:
            _lib.check(ctx.lib.vab_ode_set_fixed_params(
                ctx.h, ct.c_void_p(self._pfix_dev.data_ptr() + b0 * NPl * 8), NPl), ctx.h)
        try:
            _lib.check(ctx.lib.vab_ode_action_grad(
                ctx.h, nb, ct.c_void_p(self._XP.data_ptr() + off), self._ld, float(rf_scale),
                at(self._A, 8), at(self._me, 8), at(self._fe, 8),
                ct.c_void_p(self._G.data_ptr() + off), self._ld), ctx.h)
        finally:
            if sub:
                _lib.check(ctx.lib.vab_ode_set_fixed_params(ctx.h, ptr(self._pfix_dev), NPl), ctx.h)

    def _wave_rows(self, w0, bw):
        """Initial XP rows of initialisations [w0, w0 + bw) (see DeviceMin._anneal_device)."""
        NPl, Pidxl, _, _ = self._lay()
        Pw = self.P.reshape(self._Btot, NPl)[w0:w0 + bw][:, Pidxl]
        if self._X0_fn is None:
            if self._Xw is not None:
                return np.concatenate([self._Xw[w0:w0 + bw], Pw], axis=1)
            mp = self.minpaths.reshape(self._Btot, self.Nbeta, self._nX + NPl)
            return self._est_slice(mp[w0:w0 + bw, 0])
        import torch
        X = self._X0_fn(w0, w0 + bw)
        N, D = self.N_model, self.D
        if isinstance(X, torch.Tensor):
            X = X.to(device=self._device, dtype=torch.float64).reshape(bw, N, D)
            if self._init_to_data:
                Li = torch.as_tensor(self.Lidx, device=self._device)
                X[:, ::self.merr_nskip, Li] = self._Y_dev
            return torch.cat([X.reshape(bw, N * D), torch.from_numpy(np.ascontiguousarray(Pw)).to(self._device)], dim=1)
        X = np.array(X, dtype=np.float64).reshape(bw, N, D)
        if self._init_to_data:
            X[:, ::self.merr_nskip, self.Lidx] = self.Y
        return np.concatenate([X.reshape(bw, N * D), Pw], axis=1)

    def _est_slice(self, full):
        """(B, nX+NP) rows X ++ full P  ->  (B, nX+NPest) rows X ++ P[Pidx] (va_ode.py:715-732)."""
        return np.concatenate([full[:, :self._nX], full[:, self._nX:][:, self._lay()[1]]], axis=1)

    def anneal_step(self):
        """One rung of the ladder (va_ode.py:707-789): minimise from the previous minimiser,
        record A / me / fe / path / parameters, then raise RF."""
        self._require_resident("anneal_step()")
        B, b = self._B, self.betaidx
        prev = max(b - 1, 0)
        if not self._dev_paths_current:
            src = self.minpaths[:, prev] if self.batched else self.minpaths[prev][None, :]
            self._upload_paths(self._est_slice(src))
        t0 = time.time()
        self._minimize_device(self._rf_scale())
        XPmin = self._download_paths()
        A = self._A.cpu().numpy()
        me = self._me.cpu().numpy()
        fe = self._fe.cpu().numpy()
        st = self._status.cpu().numpy()
        nit = self._nit.cpu().numpy()
        nfev = self._nfev.cpu().numpy()
        self._dev_paths_current = True
        if self.verbose:
            print("Optimization complete!  Time = %.3f s" % (time.time() - t0))
            print("Exit flag = %s  Iterations = %s  Obj. function value = %s" % (st, nit, A))
        NPl, Pidxl, _, pshape = self._lay()
        P = self.P.reshape(B, NPl)                      # a view: self.P is updated (va_ode.py:758-774)
        if self.NPest > 0:
            P[:, Pidxl] = XPmin[:, self._nX:]
        if self.batched:
            self.A_array[:, b], self.me_array[:, b], self.fe_array[:, b] = A, me, fe
            self.exitflags[:, b], self.nit_array[:, b], self.nfev_array[:, b] = st, nit, nfev
            self.minpaths[:, b, :self._nX] = XPmin[:, :self._nX]
            self.minpaths[:, b, self._nX:] = P
            self.params_array[:, b] = P.reshape((B,) + pshape)
        else:
            self.params_array[b] = P[0].reshape(pshape)
            self.A_array[b], self.me_array[b], self.fe_array[b] = A[0], me[0], fe[0]
            self.exitflags[b], self.nit_array[b], self.nfev_array[b] = st[0], nit[0], nfev[0]
            self.minpaths[b, :self._nX] = XPmin[0, :self._nX]
            self.minpaths[b, self._nX:] = P[0]
        if b < self.Nbeta - 1:
            self.betaidx += 1
            self.beta = self.beta_array[self.betaidx]
            self.RF = self.RF0 * self.alpha ** float(self.beta)
        self.taped = False

    def _upload_paths(self, XP):
        DeviceMin._upload_paths(self, XP)
        self._dev_paths_current = False

    # ------------------------------------------------------------------ host-side error terms
    def _eval_parts(self, XP):
        XP = np.asarray(XP, dtype=np.float64)
        self._upload_paths(XP)
        self._action_grad_native(self._rf_scale())
        return self._A.cpu().numpy(), self._me.cpu().numpy(), self._fe.cpu().numpy()

    def A_gaussian(self, XP):
        """Action at XP for the current RF (va_ode.py:130-136), evaluated on the device."""
        A = self._eval_parts(XP)[0]
        return float(A[0]) if np.ndim(XP) == 1 else A

    def me_gaussian(self, X):
        """Measurement error (va_ode.py:138-158); accepts X or XP like the reference."""
        X = np.asarray(X, dtype=np.float64)
        if X.shape[-1] == self._nX:
            NPl, Pidxl, NPestl, _ = self._lay()
            pad = np.zeros(X.shape[:-1] + (NPestl,))
            pad[...] = self.P.reshape(self._B, NPl)[:, Pidxl] if X.ndim > 1 else self.P.reshape(-1)[Pidxl]
            X = np.concatenate([X, pad], axis=-1)
        me = self._eval_parts(X)[1]
        return float(me[0]) if X.ndim == 1 else me

    def fe_gaussian(self, XP):
        """RF-weighted model error (va_ode.py:160-234)."""
        fe = self._eval_parts(XP)[2]
        return float(fe[0]) if np.ndim(XP) == 1 else fe

    # ------------------------------------------------------------------ savers (va_ode.py:794-889)
    def _per_init(self, arr, init):
        if not self.batched:
            return arr
        if init is None:
            raise ValueError("batched run: pass init=<index of the initialisation to save>")
        return arr[init]

    def save_paths(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """(Nbeta, N_model, 1 + D) with t_model in column 0 (va_ode.py:794-809)."""
        mp = self._per_init(self.minpaths, init)
        X = mp[:, :self._nX].reshape(self.Nbeta, self.N_model, self.D)
        t = np.broadcast_to(self.t_model.reshape(1, self.N_model, 1), (self.Nbeta, self.N_model, 1))
        out = np.concatenate([t, X], axis=2)
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out.reshape(self.Nbeta * self.N_model, 1 + self.D), fmt=fmt)

    def save_params(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """(Nbeta, NP) -- time-dependent parameters: (Nbeta, N_model, NP) -- fixed values with the
        per-beta estimates written in (va_ode.py:811-845).  As text a time series is written as
        Nbeta * N_model rows (the reference's np.savetxt call fails on the 3-D array)."""
        mp = self._per_init(self.minpaths, init)
        pshape = (self.N_model, self.NP) if self._ptime else (self.NP,)
        out = np.array(mp[:, self._nX:], dtype=np.float64).reshape((self.Nbeta,) + pshape)
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out.reshape(-1, self.NP), fmt=fmt)

    def action_errors_table(self, cmpt=0, init=None):
        """(Nbeta, 5) rows [beta, A, me, fe, fe / (RF0 alpha**beta)] (va_ode.py:847-873)."""
        out = np.zeros((self.Nbeta, 5))
        out[:, 0] = self.beta_array
        out[:, 1] = self._per_init(self.A_array, init)
        out[:, 2] = self._per_init(self.me_array, init)
        out[:, 3] = self._per_init(self.fe_array, init)
        # (va_ode.py:859-866: RF0[0, 0] for an array, RF0[0, 0, 0] for the matrix form; cmpt picks the column)
        rf0 = self.RF0 if np.isscalar(self.RF0) else (self.RF0[0, cmpt] if self.RF0.ndim == 2 else self.RF0[0, cmpt, cmpt])
        out[:, 4] = out[:, 3] / (rf0 * float(self.alpha) ** self.beta_array.astype(np.float64))
        return out

    def save_action_errors(self, filename, cmpt=0, dtype=np.float64, fmt="%.8e", init=None):
        out = self.action_errors_table(cmpt, init)
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out, fmt=fmt)

    def save_as_minAone(self, savedir='', savefile=None, init=None):
        """minAone-style text file: rows [beta, exitflag, A, minpath...] (va_ode.py:875-889)."""
        if not savedir.endswith('/'):
            savedir += '/'
        if savefile is None:
            savefile = 'D%d_M%d_PATH%d.dat' % (self.D, self.L, self.adolcID)
        out = np.hstack([self.beta_array.reshape(-1, 1).astype(np.float64),
                         self._per_init(self.exitflags, init).reshape(-1, 1).astype(np.float64),
                         self._per_init(self.A_array, init).reshape(-1, 1),
                         self._per_init(self.minpaths, init)])
        np.savetxt(savedir + savefile, out)

    def gen_xtrace(self):
        """Kept for API compatibility (va_ode.py:894-905); nothing is taped."""
        return np.random.rand(self._n)
