"""Device-side stand-in for the reference's ``ADmin`` base class (_autodiffmin.py:16-168).

The reference tapes ``self.A`` with ADOL-C once per beta and hands ``A_gradA_taped`` to
``scipy.optimize.minimize``.  Here both halves live in libvarannealb200.so: the fused analytic
action+gradient kernels replace the tape, and a batched, device-resident L-BFGS-B replaces the
SciPy driver.  The method names of the two seams are kept (``A_gradA_taped``,
``min_lbfgs_scipy``) so code written against the reference keeps working; ``tape_A`` is a no-op
because RF is a kernel argument, not a tape constant.
"""
import ctypes as ct

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("varanneal_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def ptr(t):
    return ct.c_void_p(t.data_ptr()) if t is not None else None


def round_up(n, k):
    return (n + k - 1) // k * k


class DeviceMin(object):
    """State shared by the ODE and NN annealers: native context, device buffers for the batch of
    paths, and the eval / minimise seams."""

    _ctx = None
    _B = 1
    _n = 0          # unknowns per path
    _ld = 0         # leading dimension (doubles) of the path buffers
    opt_args = None
    bounds = None
    RF0_scalar = None
    verbose = False
    # ---- extensions for runs that do not fit the reference's "keep everything" layout
    # keep_paths: which minimising paths anneal() keeps on the host.
    #   'all'  (reference, va_ode.py:666-667): minpaths (.., Nbeta, n_states + NP)
    #   'last' : only the final rung's path, minpaths (.., 1, n_states + NP)
    #   'none' : no paths, minpaths (.., 0, n_states + NP)
    # The per-rung *parameter* estimates are always kept (``params_array`` (.., Nbeta, NP)), as are
    # A / me / fe / exitflags.  paths_file: with 'all', back minpaths by a .npy memory map so the
    # finished rungs stream to disk instead of host RAM.
    keep_paths = 'all'
    paths_file = None
    # wave_size: initialisations resident on the device at once (None: sized from free device memory)
    wave_size = None
    _Btot = 1

    # ------------------------------------------------------------------ plumbing
    def _open_context(self, device=None):
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device()
        self._device = torch.device("cuda", int(device))
        if self._ctx is not None:
            self._ctx.close()
        stream = torch.cuda.current_stream(self._device).cuda_stream
        self._ctx = _lib.Context(self._device.index, stream)
        return self._ctx

    def _to_dev(self, a, dtype=np.float64):
        torch = _torch()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self._device)

    def _alloc_paths(self, B, n):
        torch = _torch()
        self._B, self._n, self._ld = int(B), int(n), round_up(int(n), 16)
        z = lambda *s, **k: torch.zeros(*s, dtype=k.get("dtype", torch.float64), device=self._device)  # noqa: E731
        self._XP = z(self._B, self._ld)
        self._G = z(self._B, self._ld)
        self._A = z(self._B)
        self._me = z(self._B)
        self._fe = z(self._B)
        self._status = z(self._B, dtype=torch.int32)
        self._nit = z(self._B, dtype=torch.int32)
        self._nfev = z(self._B, dtype=torch.int32)

    def _upload_paths(self, XP):
        torch = _torch()
        XP = np.ascontiguousarray(np.atleast_2d(XP), dtype=np.float64)
        if XP.shape != (self._B, self._n):
            raise ValueError("XP has shape %s, expected (%d, %d)" % (XP.shape, self._B, self._n))
        self._XP[:, :self._n].copy_(torch.from_numpy(XP))

    def _action_grad_native(self, rf_scale):
        raise NotImplementedError

    def _download_paths(self):
        """Device paths -> NumPy view (B, n) of a pinned staging buffer (valid until the next call):
        one contiguous DMA instead of a strided pageable copy."""
        torch = _torch()
        if getattr(self, "_XP_full_pin", None) is None or self._XP_full_pin.shape != self._XP.shape:
            self._XP_full_pin = torch.empty(self._XP.shape, dtype=torch.float64, pin_memory=True)
        self._XP_full_pin.copy_(self._XP, non_blocking=True)
        torch.cuda.current_stream(self._device).synchronize()
        return self._XP_full_pin.numpy()[:, :self._n]

    # ------------------------------------------------------------------ eval seam
    def _rf_scale(self):
        return float(self.alpha) ** float(self.beta)

    def _pinned(self):
        """Pinned host staging buffers for the eval seam (allocated on first use)."""
        torch = _torch()
        if getattr(self, "_XP_pin", None) is None or self._XP_pin.shape != (self._B, self._n):
            self._XP_pin = torch.empty(self._B, self._n, dtype=torch.float64, pin_memory=True)
            self._G_pin = torch.empty(self._B, self._n, dtype=torch.float64, pin_memory=True)
            self._A_pin = torch.empty(self._B, dtype=torch.float64, pin_memory=True)
        return self._XP_pin, self._G_pin, self._A_pin

    def A_gradA(self, XP):
        """(A, grad A) at XP for the current RF.  XP flat (n,) -> (float, (n,) array); a batch
        (B, n) -> ((B,), (B, n)).  Replaces ADmin.A_gradA_taped (_autodiffmin.py:57-58).
        A pinned host ``torch.Tensor`` (B, n) is copied to the device without staging and the
        results come back as pinned host tensors (valid until the next call)."""
        torch = _torch()
        if isinstance(XP, torch.Tensor):
            if XP.shape != (self._B, self._n) or XP.dtype != torch.float64:
                raise ValueError("XP tensor must be float64 of shape (%d, %d)" % (self._B, self._n))
            _, G_pin, A_pin = self._pinned()
            self._dev_paths_current = False
            self._pipelined_eval(XP, G_pin, A_pin)
            return A_pin, G_pin
        XP = np.asarray(XP, dtype=np.float64)
        single = XP.ndim == 1
        if single and self._B != 1:
            raise ValueError("this Annealer was initialised with a batch of %d paths" % self._B)
        self._upload_paths(XP)
        self._action_grad_native(self._rf_scale())
        A = self._A.cpu().numpy()
        G = self._G[:, :self._n].cpu().numpy()
        if single:
            return float(A[0]), G[0]
        return A, G

    @staticmethod
    def _pipeline_spans(B, chunks=8):
        """Groups of paths for the pipelined eval: small groups at both ends (the first upload and
        the last download overlap nothing, so they are kept short), groups of B/chunks between."""
        big = max(1, -(-B // max(1, chunks)))
        if B < 4 * big or big < 4:
            sizes = [big] * (B // big) + ([B % big] if B % big else [])
        else:
            ramp = []
            k = 1
            while k < big:
                ramp.append(k)
                k *= 2
            mid = B - 2 * sum(ramp)
            sizes = ramp + [big] * (mid // big) + ([mid % big] if mid % big else []) + ramp[::-1]
        spans, lo = [], 0
        for sz in sizes:
            spans.append((lo, lo + sz))
            lo += sz
        return spans

    def _pipelined_eval(self, XP_pin, G_pin, A_pin, chunks=8):
        """Host -> device -> host evaluation of a pinned batch, software-pipelined over groups of
        paths: the upload of group i+1 and the download of group i-1 overlap the kernel of group i
        (both PCIe directions busy).  The copies are single strided DMAs (vab_copy_rows_async:
        host pitch n, device pitch ld) on their own streams."""
        torch = _torch()
        main = torch.cuda.current_stream(self._device)
        if getattr(self, "_s_h2d", None) is None:
            self._s_h2d = torch.cuda.Stream(self._device)
            self._s_d2h = torch.cuda.Stream(self._device)
        B, n, ld = self._B, self._n, self._ld
        lib, h = self._ctx.lib, self._ctx.h
        spans = self._pipeline_spans(B, chunks)
        scale = self._rf_scale()
        self._s_h2d.wait_stream(main)
        self._s_d2h.wait_stream(main)
        s_in, s_out = ct.c_void_p(self._s_h2d.cuda_stream), ct.c_void_p(self._s_d2h.cuda_stream)
        xp_h, g_h = XP_pin.data_ptr(), G_pin.data_ptr()
        xp_d, g_d = self._XP.data_ptr(), self._G.data_ptr()
        ev_in = []
        for lo, hi in spans:
            _lib.check(lib.vab_copy_rows_async(h, 1, ct.c_void_p(xp_d + lo * ld * 8), ld,
                                               ct.c_void_p(xp_h + lo * n * 8), n, n, hi - lo, s_in), h)
            e = torch.cuda.Event()
            e.record(self._s_h2d)
            ev_in.append(e)
        for (lo, hi), e in zip(spans, ev_in):
            main.wait_event(e)
            self._action_grad_native(scale, lo, hi - lo)
            eo = torch.cuda.Event()
            eo.record(main)
            self._s_d2h.wait_event(eo)
            _lib.check(lib.vab_copy_rows_async(h, 0, ct.c_void_p(g_d + lo * ld * 8), ld,
                                               ct.c_void_p(g_h + lo * n * 8), n, n, hi - lo, s_out), h)
        with torch.cuda.stream(self._s_d2h):
            A_pin.copy_(self._A, non_blocking=True)
        main.wait_stream(self._s_d2h)
        main.synchronize()

    def A_gradA_taped(self, XP):
        return self.A_gradA(XP)

    def A_taped(self, XP):
        return self.A_gradA(XP)[0]

    def gradA_taped(self, XP):
        return self.A_gradA(XP)[1]

    def jacA_taped(self, XP):
        """ADmin.jacA_taped (_autodiffmin.py:60-61): the Jacobian of the scalar action, i.e. the
        gradient as a (1, n) array (per path: (B, 1, n))."""
        g = np.asarray(self.A_gradA(XP)[1])
        return g.reshape(g.shape[:-1] + (1, g.shape[-1]))

    def A_jacaA_taped(self, XP):
        """ADmin.A_jacaA_taped (sic, _autodiffmin.py:63-64)."""
        A, g = self.A_gradA(XP)
        g = np.asarray(g)
        return A, g.reshape(g.shape[:-1] + (1, g.shape[-1]))

    def hessianA_taped(self, XP, rel_step=1e-6):
        """ADmin.hessianA_taped (_autodiffmin.py:66-67): dense (n, n) Hessian of the action at a
        single path.  ADOL-C differentiates the tape twice; here the analytic device gradient is
        differenced centrally, column by column (2 n evaluations, steps rel_step (1 + |x_j|)) and
        symmetrised -- accurate to ~1e-9 relative, meant for the small problems a dense Hessian
        makes sense for."""
        XP = np.asarray(XP, dtype=np.float64)
        if XP.ndim != 1 or self._B != 1:
            raise ValueError("hessianA_taped takes one flat path on a single-path Annealer")
        n = XP.shape[0]
        H = np.empty((n, n))
        x = XP.copy()
        for j in range(n):
            h = rel_step * (1.0 + abs(XP[j]))
            x[j] = XP[j] + h
            gp = self.A_gradA(x)[1]
            x[j] = XP[j] - h
            gm = self.A_gradA(x)[1]
            x[j] = XP[j]
            H[:, j] = (gp - gm) / (2.0 * h)
        return 0.5 * (H + H.T)

    def tape_A(self, xtrace=None):
        """No-op: nothing is taped (kept for API compatibility, _autodiffmin.py:32-49)."""
        self.taped = True

    # ------------------------------------------------------------------ minimise seam
    def _lbfgs_opts(self, method=0):
        o = dict(self.opt_args or {})
        known = {"gtol", "ftol", "maxfun", "maxiter", "maxcor", "maxls", "poll_every",
                 "disp", "iprint", "eps", "maxiter_per_beta"}
        if method == 2:     # scipy.optimize.minimize(method='TNC') options; the ones the device driver has no
            known |= {"maxCGit", "eta", "stepmx", "accuracy", "minfev", "rescale", "xtol", "scale", "offset"}   # use for are accepted and ignored
        bad = set(o) - known
        if bad:
            raise ValueError("unsupported opt_args keys for the device L-BFGS-B: %s" % sorted(bad))
        opts = _lib.LbfgsOpts()
        opts.m = int(o.get("maxcor", 10))
        opts.maxls = int(o.get("maxls", 20))
        opts.maxfun = int(min(float(o.get("maxfun", 15000)), 2 ** 62))
        opts.maxiter = int(min(float(o.get("maxiter", 15000)), 2 ** 62))
        opts.ftol = float(o.get("ftol", 2.220446049250313e-09))
        opts.pgtol = float(o.get("gtol", 1e-5))
        opts.poll_every = int(o.get("poll_every", 0))
        opts.method = int(method)
        if method == 1 and "maxiter" not in o:
            opts.maxiter = 200 * self._n                     # SciPy CG default
        if method == 2:
            # SciPy's TNC: maxfun (alias maxiter) bounds the evaluations, default max(100, 10 n);
            # ftol < 0 -> 0 (off), gtol < 0 -> 1e-2 sqrt(accuracy) with accuracy = sqrt(eps)
            opts.m = int(o.get("maxCGit", 0))
            lim = o.get("maxfun", o.get("maxiter", max(100, 10 * self._n)))
            opts.maxfun = int(min(float(lim), 2 ** 62))
            opts.maxiter = opts.maxfun
            opts.ftol = max(0.0, float(o.get("ftol", -1.0)))
            g = float(o.get("gtol", -1.0))
            opts.pgtol = g if g >= 0.0 else 1e-2 * (2.220446049250313e-16 ** 0.25)
        return opts

    def _minimize_device(self, rf_scale, method=None):
        """Runs the device minimiser on the paths currently in self._XP (in place)."""
        if method is None:
            method = {"NCG": 1, "TNC": 2}.get(getattr(self, "method", "L-BFGS-B"), 0)
        opts = self._lbfgs_opts(method)
        lo = ptr(getattr(self, "_lo_dev", None)) if method in (0, 2) else None     # SciPy's CG ignores bounds
        hi = ptr(getattr(self, "_hi_dev", None)) if method in (0, 2) else None
        _lib.check(self._ctx.lib.vab_minimize(
            self._ctx.h, self._B, ptr(self._XP), self._ld, float(rf_scale), ct.byref(opts),
            lo, hi, ptr(self._A), ptr(self._me), ptr(self._fe), ptr(self._status),
            ptr(self._nit), ptr(self._nfev)), self._ctx.h)

    # ------------------------------------------------------------------ waves
    def _per_path_extra_bytes(self):
        """Device bytes per resident path besides the n-vectors (problem-specific workspaces)."""
        return 0

    def _plan_wave(self, B, n, Nbeta):
        """How many of the B initialisations are resident at once.  Per path the device holds
        XP and the gradient, the minimiser's 2m+4 vectors and -- with keep_paths == 'all' -- the
        Nbeta minimisers of the ladder; a wave is as many paths as fit in 85 % of the free device
        memory (C3, n = 1e8: 26 vectors x 0.8 GB -> 7 paths of 128 per GPU at a time).
        ``wave_size`` / VAB_WAVE override."""
        import os
        torch = _torch()
        if self.keep_paths not in ('all', 'last', 'none'):
            raise ValueError("keep_paths must be 'all', 'last' or 'none'")
        forced = self.wave_size or int(os.environ.get("VAB_WAVE", "0"))
        if forced:
            return max(1, min(int(B), int(forced)))
        free, _ = torch.cuda.mem_get_info(self._device)
        m = int((self.opt_args or {}).get("maxcor", 10))
        ld = round_up(int(n), 16)
        nvec = 2 + (2 * m + 4) + (Nbeta if self.keep_paths == 'all' else 0)
        per = nvec * ld * 8 + self._per_path_extra_bytes()
        fit = int(0.85 * free // per)
        if fit < 1:
            raise MemoryError("one path needs %.1f GB on the device (%d vectors of %d doubles); %.1f GB are free"
                              % (per / 1e9, nvec, ld, free / 1e9))
        return min(int(B), fit)

    def _ladder_fits_device(self):
        """anneal() runs the ladder in native calls, wave by wave (waves are sized by _plan_wave
        to fit); kept for callers of the round-1 name."""
        return True

    def _wave_rows(self, w0, bw):
        """(bw, n) rows X0 ++ P0[Pidx] of initialisations [w0, w0 + bw): from the working copy of
        X0 made by anneal_init, or from the lazy X0 callable."""
        raise NotImplementedError

    def _lay(self):
        """(NP, Pidx, NPest, block shape) of the parameter block that follows the states in
        minpaths rows (full block) and in XP (estimated entries).  A parameter time series
        (va_ode: P0 of shape (N_model, NP)) overrides this with the flattened (N_model, NP) block."""
        return self.NP, self.Pidx, self.NPest, (self.NP,)

    def _set_wave_fixed_params(self, w0, bw):
        """Fixed (non-estimated) parameter values of the wave's paths -> the device block the
        kernels read (va_ode.py:178-181: fixed entries come from self.P)."""
        torch = _torch()
        P = np.ascontiguousarray(self.P.reshape(self._Btot, self._lay()[0])[w0:w0 + bw], dtype=np.float64)
        self._pfix_dev[:bw].copy_(torch.from_numpy(P))

    def _load_wave(self, w0, bw):
        torch = _torch()
        rows = self._wave_rows(w0, bw)
        if not isinstance(rows, torch.Tensor):
            rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64))
        self._XP[:bw, :self._n].copy_(rows)
        self._set_wave_fixed_params(w0, bw)
        self._dev_paths_current = False

    def _alloc_results(self, B, Nbeta, n_states, batched):
        """Host result arrays with the reference's layout (va_ode.py:666-699, va_nnet.py:419-452);
        a batch adds a leading axis; keep_paths shrinks the rung axis of minpaths."""
        lead = (B,) if batched else ()
        nkeep = {'all': Nbeta, 'last': 1, 'none': 0}[self.keep_paths]
        NPl, _, _, pshape = self._lay()
        width = n_states + NPl
        if self.paths_file is not None and self.keep_paths == 'all':
            self.minpaths = np.lib.format.open_memmap(self.paths_file, mode='w+', dtype=np.float64,
                                                      shape=lead + (nkeep, width))
        else:
            self.minpaths = np.zeros(lead + (nkeep, width), dtype=np.float64)
        shape = lead + (Nbeta,)
        self.A_array = np.zeros(shape, dtype=np.float64)
        self.me_array = np.zeros(shape, dtype=np.float64)
        self.fe_array = np.zeros(shape, dtype=np.float64)
        self.exitflags = np.zeros(shape, dtype=np.int8)
        self.nit_array = np.zeros(shape, dtype=np.int64)
        self.nfev_array = np.zeros(shape, dtype=np.int64)
        self.params_array = np.zeros(shape + (NPl,), dtype=np.float64)
        self.params_array[...] = self.P.reshape(lead + (1, NPl))
        self.params_array = self.params_array.reshape(shape + pshape)

    def _require_resident(self, what):
        if self._Btot != self._B or self.keep_paths != 'all':
            raise NotImplementedError(
                "%s drives the device rung by rung and needs every initialisation resident and "
                "keep_paths='all' (%d of %d fit); use anneal() without track_* / verbose"
                % (what, self._B, self._Btot))

    def _anneal_device(self):
        """The whole beta loop (reference: anneal + anneal_step, va_ode.py:473-490, 707-789;
        va_nnet.py:281-286, 459-523) in native calls: ``vab_anneal`` runs every resident path down
        the ladder on the device (paths advance rung by rung independently of each other); the
        initialisations are processed in waves of ``self._B`` resident paths (one wave when they all
        fit), and the result table and the minimising paths are laid out exactly as the stepwise
        loop does."""
        torch = _torch()
        Btot, Bw, nb, n, nX = self._Btot, self._B, self.Nbeta, self._n, self._nX
        method = 1 if getattr(self, "method", "L-BFGS-B") == "NCG" else 0     # TNC never gets here (anneal())
        opts = self._lbfgs_opts(method)
        lo = ptr(getattr(self, "_lo_dev", None)) if method == 0 else None
        hi = ptr(getattr(self, "_hi_dev", None)) if method == 0 else None
        dev = self._device
        keep = self.keep_paths
        table = torch.zeros(Bw, nb, 5, dtype=torch.float64, device=dev)
        stat = torch.zeros(Bw, nb, dtype=torch.int32, device=dev)
        nit = torch.zeros_like(stat)
        nfev = torch.zeros_like(stat)
        NPl, Pidxl, npest, _ = self._lay()
        ppitch = max(2, (npest + 1) // 2 * 2)
        if keep == 'all':
            paths = torch.empty(Bw, nb, self._ld, dtype=torch.float64, device=dev)
        else:
            paths = torch.zeros(Bw, nb, ppitch, dtype=torch.float64, device=dev)   # parameter window only
        betas = np.asarray(self.beta_array, dtype=np.float64)
        beta_c = (ct.c_double * nb)(*betas)
        lib, h = self._ctx.lib, self._ctx.h
        nkeep = self.minpaths.shape[-2]
        mp_all = self.minpaths.reshape(Btot, nkeep, nX + NPl)
        shape = (Btot, nb)
        A_all, me_all, fe_all = (a.reshape(shape) for a in (self.A_array, self.me_array, self.fe_array))
        ef_all, nit_all, nfev_all = (a.reshape(shape) for a in (self.exitflags, self.nit_array, self.nfev_array))
        P_all = self.P.reshape(Btot, NPl)
        par_all = self.params_array.reshape(Btot, nb, NPl)
        self.n_waves = 0
        for w0 in range(0, Btot, Bw):
            bw = min(Bw, Btot - w0)
            self._load_wave(w0, bw)
            if keep == 'all':
                # states go home while the ladder runs: finished rungs are copied from the device
                # rows (pitch ld) straight into the host rows (pitch nX + NP) of self.minpaths
                mp = mp_all[w0:w0 + bw].reshape(bw * nb, nX + NPl)
                _lib.check(lib.vab_set_path_sink(h, ct.c_void_p(mp.ctypes.data), nX + NPl, nX), h)
            else:
                _lib.check(lib.vab_set_path_window(h, nX, npest), h)
            _lib.check(lib.vab_anneal(h, bw, ptr(self._XP), self._ld, float(self.alpha), beta_c, nb,
                                      ct.byref(opts), lo, hi, ptr(table), ptr(paths), ptr(stat), ptr(nit),
                                      ptr(nfev)), h)
            tab = table[:bw].cpu().numpy()
            A_all[w0:w0 + bw] = tab[:, :, 1]
            me_all[w0:w0 + bw] = tab[:, :, 2]
            fe_all[w0:w0 + bw] = tab[:, :, 3]
            ef_all[w0:w0 + bw] = stat[:bw].cpu().numpy()
            nit_all[w0:w0 + bw] = nit[:bw].cpu().numpy()
            nfev_all[w0:w0 + bw] = nfev[:bw].cpu().numpy()
            # parameters: the fixed values with the estimates of every rung written in
            P = np.broadcast_to(P_all[w0:w0 + bw].reshape(bw, 1, NPl), (bw, nb, NPl)).copy()
            if npest > 0:
                if keep == 'all':
                    P[:, :, Pidxl] = paths[:bw, :, nX:n].cpu().numpy()
                else:
                    P[:, :, Pidxl] = paths[:bw, :, :npest].cpu().numpy()
            par_all[w0:w0 + bw] = P
            if keep == 'all':
                mp_all[w0:w0 + bw, :, nX:] = P
            elif keep == 'last':
                # vab_anneal leaves the last rung's minimiser in XP
                _lib.check(lib.vab_copy_rows_to_host(h, ct.c_void_p(mp_all[w0:w0 + bw].ctypes.data), nX + NPl,
                                                     ptr(self._XP), self._ld, nX, bw), h)
                mp_all[w0:w0 + bw, 0, nX:] = P[:, -1]
            P_all[w0:w0 + bw] = P[:, -1]
            self.n_waves += 1
        del paths
        self._dev_paths_current = (Btot == Bw)
        self.betaidx = nb - 1
        self.beta = self.beta_array[-1]
        self.RF = self.RF0 * self.alpha ** float(self.beta)
        self.taped = False

    def min_lbfgs_scipy(self, XP0, xtrace=None):
        """Same contract as ADmin.min_lbfgs_scipy (_autodiffmin.py:72-95): returns
        (XPmin, Amin, status) -- but the whole minimisation runs on the device.  With a batch,
        the three are arrays over the paths."""
        XP0 = np.asarray(XP0, dtype=np.float64)
        single = XP0.ndim == 1
        self._upload_paths(XP0)
        self._minimize_device(self._rf_scale())
        XPmin = self._XP[:, :self._n].cpu().numpy()
        A = self._A.cpu().numpy()
        st = self._status.cpu().numpy()
        self.last_nit = self._nit.cpu().numpy()
        self.last_nfev = self._nfev.cpu().numpy()
        if single:
            return XPmin[0], float(A[0]), int(st[0])
        return XPmin, A, st

    min_lbfgs = min_lbfgs_scipy

    def min_cg_scipy(self, XP0, xtrace=None):
        """Same contract as ADmin.min_cg_scipy (_autodiffmin.py:97-119), on the device: nonlinear
        conjugate gradients (Polak-Ribiere+, More'-Thuente search with SciPy's c1 = 1e-4,
        c2 = 0.4, stop on max|g| <= gtol or maxiter).  Bounds are ignored, as by SciPy's CG."""
        XP0 = np.asarray(XP0, dtype=np.float64)
        single = XP0.ndim == 1
        self._upload_paths(XP0)
        lo, hi = getattr(self, "_lo_dev", None), getattr(self, "_hi_dev", None)
        self._lo_dev = self._hi_dev = None
        try:
            self._minimize_device(self._rf_scale(), method=1)
        finally:
            self._lo_dev, self._hi_dev = lo, hi
        XPmin = self._XP[:, :self._n].cpu().numpy()
        A = self._A.cpu().numpy()
        st = self._status.cpu().numpy()
        self.last_nit = self._nit.cpu().numpy()
        self.last_nfev = self._nfev.cpu().numpy()
        if single:
            return XPmin[0], float(A[0]), int(st[0])
        return XPmin, A, st

    def min_tnc_scipy(self, XP0, xtrace=None):
        """Same contract as ADmin.min_tnc_scipy (_autodiffmin.py:121-143), on the device: truncated
        Newton (CG inner solve on gradient-difference Hessian-vector products, More'-Thuente
        search); status = SciPy's TNC return code.  See csrc/tnc.cu for what differs from Nash's
        TNC.  Bounds are honoured (active-set version of the same iteration)."""
        XP0 = np.asarray(XP0, dtype=np.float64)
        single = XP0.ndim == 1
        self._upload_paths(XP0)
        self._minimize_device(self._rf_scale(), method=2)
        XPmin = self._XP[:, :self._n].cpu().numpy()
        A = self._A.cpu().numpy()
        st = self._status.cpu().numpy()
        self.last_nit = self._nit.cpu().numpy()
        self.last_nfev = self._nfev.cpu().numpy()
        if single:
            return XPmin[0], float(A[0]), int(st[0])
        return XPmin, A, st

    @property
    def gpu_launches(self):
        return self._ctx.launches if self._ctx is not None else 0
