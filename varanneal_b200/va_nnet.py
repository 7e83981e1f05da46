"""``va_nnet.Annealer`` -- drop-in for the reference's neural-network annealer
(varanneal/va_nnet.py:44-673): the layers of a feed-forward net play the role of time, the M
training examples are independent copies sharing the parameters.

Same class / method names, positional order (``Pidx`` *before* ``Lidx``, no ``dt_model`` --
va_nnet.py:269-271) and result-array layouts as the reference; the ADOL-C tape and the SciPy loop
are replaced by libvarannealb200.so (csrc/nn_action.cu, csrc/lbfgs.cu).

Deliberate differences:
  * ``set_activation(f)`` takes 'sigmoid' | 'tanh' | 'linear' (or the reference-style callables
    ``varanneal_b200.va_nnet.sigmoid`` ...): arbitrary Python callables cannot run in a kernel.
  * the default ``Lidx=None`` observes every input and output neuron with integer indices (the
    reference builds float ``np.linspace`` indices that no longer index under modern NumPy).
  * ``RM`` scalar, shape (2,) (input / output weights, va_nnet.py:131-144) or a pair of matrices
    [RM_in, RM_out] (va_nnet.py:135-139).
  * ``exitflags`` is filled; errors raise.
  * batches: ``X0`` (B, M*NDnet) + ``P0`` (B, NP) anneal B initialisations at once.
"""
import ctypes as ct
import time

import numpy as np

from . import _lib
from ._devicemin import DeviceMin, ptr


def sigmoid(x, W, b):
    """Reference-style activation (examples/nnet_twin/nnet_twin_anneal.py:20-22)."""
    return 1.0 / (1.0 + np.exp(-(np.dot(W, x) + b)))


def tanh(x, W, b):
    return np.tanh(np.dot(W, x) + b)


def linear(x, W, b):
    return np.dot(W, x) + b


for _f, _n in ((sigmoid, "sigmoid"), (tanh, "tanh"), (linear, "linear")):
    _f.vab_activation = _n


class Annealer(DeviceMin):
    def __init__(self, device=None, verbose=False):
        self.taped = False
        self.annealing_initialized = False
        self.M = 0
        self.structure = None
        self.act_name = "sigmoid"
        self._device_arg = device
        self.verbose = verbose

    # ------------------------------------------------------------------ problem definition
    def _sizes(self):
        if self.structure is not None and self.M > 0:
            self.NDnet = int(np.sum(self.structure))
            self.NDens = self.NDnet * self.M

    def set_structure(self, structure):
        """va_nnet.py:59-69."""
        self.structure = np.asarray(structure, dtype=np.int64)
        self.N = len(self.structure)
        self._sizes()

    def set_activation(self, f):
        """va_nnet.py:71-76; ``f`` a name or one of this module's activation callables."""
        name = f if isinstance(f, str) else getattr(f, "vab_activation", None)
        if name not in _lib.ACT_IDS:
            raise ValueError("set_activation needs 'sigmoid', 'tanh' or 'linear' (or the callables of "
                             "varanneal_b200.va_nnet); arbitrary Python callables cannot run in the kernels")
        self.act_name = name
        self.f = {"sigmoid": sigmoid, "tanh": tanh, "linear": linear}[name]

    def set_input_data(self, data):
        """va_nnet.py:78-91."""
        data = np.asarray(data, dtype=np.float64)
        self.data_in = data.reshape(1, -1) if data.ndim == 1 else data
        if self.M == 0:
            self.M = self.data_in.shape[0]
        self._sizes()

    def set_output_data(self, data):
        """va_nnet.py:93-106."""
        data = np.asarray(data, dtype=np.float64)
        self.data_out = data.reshape(1, -1) if data.ndim == 1 else data
        if self.M == 0:
            self.M = self.data_out.shape[0]
        self._sizes()

    # ------------------------------------------------------------------ annealing
    def anneal(self, X0, P0, alpha, beta_array, RM, RF0, Pidx, Lidx=None,
               init_to_data=True, action='A_gaussian', disc='forwardmap',
               method='L-BFGS-B', bounds=None, opt_args=None, adolcID=0):
        """va_nnet.py:269-286."""
        if not self.annealing_initialized:
            self.anneal_init(X0, P0, alpha, beta_array, RM, RF0, Pidx, Lidx, init_to_data, action,
                             disc, method, bounds, opt_args, adolcID)
        if not self.verbose and self.betaidx == 0 and self.method != 'TNC':
            self._anneal_device()           # the whole ladder in native calls, one per wave of paths
            return
        self._require_resident("anneal() with verbose / method='TNC'")
        for _ in range(self.Nbeta):
            if self.verbose:
                print('------------------------------')
                print('Step %d of %d' % (self.betaidx + 1, self.Nbeta))
                print('beta = %d, RF = %.8e' % (self.beta, self.RF))
            self.anneal_step()

    def anneal_init(self, X0, P0, alpha, beta_array, RM, RF0, Pidx, Lidx=None,
                    init_to_data=True, action='A_gaussian', disc='forwardmap',
                    method='L-BFGS-B', bounds=None, opt_args=None, adolcID=0):
        """va_nnet.py:288-457."""
        if method == 'LM':
            raise ValueError("method='LM' is dead code in the reference (SURVEY.md App. B10)")
        if method not in ('L-BFGS-B', 'NCG', 'TNC'):
            raise ValueError("Optimization routine not recognized: %r" % (method,))
        if action != 'A_gaussian' or disc != 'forwardmap':
            raise ValueError("va_nnet supports action='A_gaussian', disc='forwardmap' (va_nnet.py:260-264)")
        if self.structure is None or self.M == 0:
            raise ValueError("set_structure / set_input_data / set_output_data first")
        self.method = method
        self.opt_args = opt_args
        st = self.structure
        NP = int(sum(st[n] * st[n + 1] + st[n + 1] for n in range(self.N - 1)))

        X0 = X0 if isinstance(X0, np.ndarray) else np.asarray(X0, dtype=np.float64)
        P0 = P0 if isinstance(P0, np.ndarray) else np.asarray(P0, dtype=np.float64)
        self.batched = (X0.ndim == 2)
        B = X0.shape[0] if self.batched else 1
        if X0.shape[-1] != self.NDens:
            raise ValueError("X0 must have M*NDnet = %d entries per path" % self.NDens)
        if self.batched and P0.ndim == 1:
            P0 = np.tile(P0, (B, 1))
        if P0.shape[-1] != NP:
            raise ValueError("P0 must have %d entries ([W_0, b_0, W_1, b_1, ...], va_nnet.py:194-207)" % NP)
        self.P = np.array(P0, dtype=np.float64)
        self.NP = NP
        self.Pidx = np.asarray(Pidx, dtype=np.int64).reshape(-1)
        self.NPest = len(self.Pidx)
        if Lidx is None:
            Lidx = [np.arange(st[0]), np.arange(st[-1])]
        self.Lidx = [np.asarray(Lidx[0], dtype=np.int64), np.asarray(Lidx[1], dtype=np.int64)]
        self.L = [len(self.Lidx[0]), len(self.Lidx[1])]
        self.Ltot = self.L[0] + self.L[1]
        if self.data_in.shape != (self.M, self.L[0]) or self.data_out.shape != (self.M, self.L[1]):
            raise ValueError("data_in / data_out must be (M, len(Lidx[0])) / (M, len(Lidx[1]))")
        rm_mats = None
        if np.isscalar(RM) or (isinstance(RM, np.ndarray) and RM.ndim == 0):
            self.RM = float(RM)
            rm_in = rm_out = self.RM
        elif len(RM) == 2 and np.ndim(RM[0]) == 2 and np.ndim(RM[1]) == 2:
            # [RM_in, RM_out] matrices (va_nnet.py:135-139; the reference reaches this branch only
            # with a (2, L, L) array, i.e. as many measured inputs as outputs -- a pair of
            # differently sized matrices is accepted here as well)
            rm_mats = (np.ascontiguousarray(RM[0], dtype=np.float64), np.ascontiguousarray(RM[1], dtype=np.float64))
            if rm_mats[0].shape != (self.L[0],) * 2 or rm_mats[1].shape != (self.L[1],) * 2:
                raise ValueError("RM matrices must be (len(Lidx[0]),)*2 and (len(Lidx[1]),)*2")
            self.RM = RM
            rm_in = rm_out = 0.0
        else:
            RM = np.asarray(RM, dtype=np.float64)
            if RM.shape != (2,):
                raise ValueError("RM must be a scalar, of shape (2,), or a pair of matrices (va_nnet.py:131-144)")
            self.RM = RM
            rm_in, rm_out = float(RM[0]), float(RM[1])
        if bounds is not None:
            bounds = np.asarray(bounds, dtype=np.float64)
            if bounds.shape != (self.NDens + self.NPest, 2):
                raise ValueError("bounds must hold one [lo, hi] pair per unknown (va_nnet.py:302-305)")
        self.bounds = bounds

        self.alpha = alpha
        self.beta_array = np.asarray(beta_array, dtype=np.float64)     # floats kept (va_nnet.py:396)
        self.Nbeta = len(self.beta_array)
        self.betaidx = 0
        self.beta = self.beta_array[0]
        if RF0 is not None:
            self.RF0 = float(RF0)
        self.RF = self.RF0 * self.alpha ** self.beta

        # float64 working copy of the initial states; init_to_data acts on it and is then written
        # back into the caller's array (the reference mutates X0 in place, va_nnet.py:427-434) --
        # a strided or integer X0 can no longer lose the clamp on the way to the device
        Xw = np.array(X0, dtype=np.float64).reshape(B, self.M, self.NDnet)
        if init_to_data:
            Xw[:, :, self.Lidx[0]] = self.data_in
            Xw[:, :, self.NDnet - st[-1] + self.Lidx[1]] = self.data_out
            if isinstance(X0, np.ndarray) and X0.flags.writeable:
                X0[...] = Xw.reshape(X0.shape)

        self.adolcID = adolcID
        self._nX = self.NDens
        ctx = self._open_context(self._device_arg)
        self._Btot = B
        Bw = self._plan_wave(B, self.NDens + self.NPest, self.Nbeta)
        self._alloc_results(B, self.Nbeta, self.NDens, self.batched)
        self._Xw = None
        if self.keep_paths == 'all':             # the reference parks XP0 in minpaths[0] (va_nnet.py:419-421)
            self.minpaths.reshape(B, self.Nbeta, self.NDens + NP)[:, 0] = np.concatenate(
                [Xw.reshape(B, self.NDens), self.P.reshape(B, NP)], axis=1)
        else:
            self._Xw = Xw.reshape(B, self.NDens)
        self._alloc_paths(Bw, self.NDens + self.NPest)
        self._din_dev = self._to_dev(self.data_in)
        self._dout_dev = self._to_dev(self.data_out)
        _lib.check(ctx.lib.vab_nn_problem_set(
            ctx.h, self.N, _lib.int_array(st), self.M, _lib.ACT_IDS[self.act_name],
            self.L[0], _lib.int_array(self.Lidx[0]), self.L[1], _lib.int_array(self.Lidx[1]),
            ptr(self._din_dev), ptr(self._dout_dev), self.NPest, _lib.int_array(self.Pidx)), ctx.h)
        _lib.check(ctx.lib.vab_nn_set_weights(ctx.h, rm_in, rm_out, self.RF0), ctx.h)
        if rm_mats is not None:
            self._rm_mats_dev = (self._to_dev(rm_mats[0]), self._to_dev(rm_mats[1]))
            _lib.check(ctx.lib.vab_nn_set_rm_matrices(ctx.h, ptr(self._rm_mats_dev[0]), ptr(self._rm_mats_dev[1])), ctx.h)
        self._pfix_dev = self._to_dev(self.P.reshape(B, NP)[:Bw])
        _lib.check(ctx.lib.vab_nn_set_fixed_params(ctx.h, ptr(self._pfix_dev), NP), ctx.h)
        self._lo_dev = self._hi_dev = None
        if bounds is not None:
            pad = self._ld - self._n
            self._lo_dev = self._to_dev(np.concatenate([bounds[:, 0], np.full(pad, -np.inf)]))
            self._hi_dev = self._to_dev(np.concatenate([bounds[:, 1], np.full(pad, np.inf)]))
        self._dev_paths_current = False
        self.initalized = True

    def _action_grad_native(self, rf_scale, b0=0, nb=None):
        ctx = self._ctx
        nb = self._B - b0 if nb is None else nb
        sub = (b0 != 0 or nb != self._B)
        off = b0 * self._ld * 8
        if sub:
            _lib.check(ctx.lib.vab_nn_set_fixed_params(
                ctx.h, ct.c_void_p(self._pfix_dev.data_ptr() + b0 * self.NP * 8), self.NP), ctx.h)
        try:
            _lib.check(ctx.lib.vab_nn_action_grad(
                ctx.h, nb, ct.c_void_p(self._XP.data_ptr() + off), self._ld, float(rf_scale),
                ct.c_void_p(self._A.data_ptr() + b0 * 8), ct.c_void_p(self._me.data_ptr() + b0 * 8),
                ct.c_void_p(self._fe.data_ptr() + b0 * 8), ct.c_void_p(self._G.data_ptr() + off), self._ld), ctx.h)
        finally:
            if sub:
                _lib.check(ctx.lib.vab_nn_set_fixed_params(ctx.h, ptr(self._pfix_dev), self.NP), ctx.h)

    def _upload_paths(self, XP):
        DeviceMin._upload_paths(self, XP)
        self._dev_paths_current = False

    def _est_slice(self, full):
        return np.concatenate([full[:, :self.NDens], full[:, self.NDens:][:, self.Pidx]], axis=1)

    def _per_path_extra_bytes(self):
        # Delta / lambda side buffers of the split kernels: M x sum(d_1..) doubles each per path
        return 2 * 8 * self.M * int(np.sum(self.structure[1:]))

    def _wave_rows(self, w0, bw):
        """Initial XP rows of initialisations [w0, w0 + bw) (see DeviceMin._anneal_device)."""
        if self._Xw is not None:
            Pw = self.P.reshape(self._Btot, self.NP)[w0:w0 + bw][:, self.Pidx]
            return np.concatenate([self._Xw[w0:w0 + bw], Pw], axis=1)
        mp = self.minpaths.reshape(self._Btot, self.Nbeta, self.NDens + self.NP)
        return self._est_slice(mp[w0:w0 + bw, 0])

    def anneal_step(self):
        """va_nnet.py:459-523."""
        self._require_resident("anneal_step()")
        B, b = self._B, self.betaidx
        prev = max(b - 1, 0)
        if not self._dev_paths_current:
            src = self.minpaths[:, prev] if self.batched else self.minpaths[prev][None, :]
            self._upload_paths(self._est_slice(src))
        t0 = time.time()
        self._minimize_device(self._rf_scale())
        XPmin = self._download_paths()
        A, me, fe = self._A.cpu().numpy(), self._me.cpu().numpy(), self._fe.cpu().numpy()
        st, nit, nfev = self._status.cpu().numpy(), self._nit.cpu().numpy(), self._nfev.cpu().numpy()
        self._dev_paths_current = True
        if self.verbose:
            print("Optimization complete!  Time = %.3f s  Exit flag = %s  Iterations = %s  Obj = %s"
                  % (time.time() - t0, st, nit, A))
        P = self.P.reshape(B, self.NP)
        if self.NPest > 0:
            P[:, self.Pidx] = XPmin[:, self.NDens:]
        if self.batched:
            self.A_array[:, b], self.me_array[:, b], self.fe_array[:, b] = A, me, fe
            self.exitflags[:, b], self.nit_array[:, b], self.nfev_array[:, b] = st, nit, nfev
            self.minpaths[:, b, :self.NDens] = XPmin[:, :self.NDens]
            self.minpaths[:, b, self.NDens:] = P
            self.params_array[:, b] = P
        else:
            self.params_array[b] = P[0]
            self.A_array[b], self.me_array[b], self.fe_array[b] = A[0], me[0], fe[0]
            self.exitflags[b], self.nit_array[b], self.nfev_array[b] = st[0], nit[0], nfev[0]
            self.minpaths[b, :self.NDens] = XPmin[0, :self.NDens]
            self.minpaths[b, self.NDens:] = P[0]
        if b < self.Nbeta - 1:
            self.betaidx += 1
            self.beta = self.beta_array[self.betaidx]
            self.RF = self.RF0 * self.alpha ** self.beta
        self.taped = False

    # ------------------------------------------------------------------ error terms on the device
    def _eval_parts(self, XP):
        self._upload_paths(np.asarray(XP, dtype=np.float64))
        self._action_grad_native(self._rf_scale())
        return self._A.cpu().numpy(), self._me.cpu().numpy(), self._fe.cpu().numpy()

    def A_gaussian(self, XP):
        A = self._eval_parts(XP)[0]
        return float(A[0]) if np.ndim(XP) == 1 else A

    def me_gaussian(self, XP):
        XP = np.asarray(XP, dtype=np.float64)
        if XP.shape[-1] == self.NDens:
            XP = np.concatenate([XP, np.zeros(XP.shape[:-1] + (self.NPest,))], axis=-1)
        me = self._eval_parts(XP)[1]
        return float(me[0]) if XP.ndim == 1 else me

    def fe_gaussian(self, XP):
        fe = self._eval_parts(XP)[2]
        return float(fe[0]) if np.ndim(XP) == 1 else fe

    # ------------------------------------------------------------------ savers (va_nnet.py:528-663)
    def _per_init(self, arr, init):
        if not self.batched:
            return arr
        if init is None:
            raise ValueError("batched run: pass init=<index of the initialisation to save>")
        return arr[init]

    def _layer_slices(self):
        off = np.concatenate([[0], np.cumsum(self.structure)])
        return [(int(off[n]), int(off[n + 1])) for n in range(self.N)]

    def save_states(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """Minimising neuron states, (Nbeta, M, NDnet) (va_nnet.py:528-546 stores the same values as
        nested object arrays; a dense array is written here)."""
        mp = self._per_init(self.minpaths, init)
        out = mp[:, :self.NDens].reshape(self.Nbeta, self.M, self.NDnet)
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out.reshape(self.Nbeta, -1), fmt=fmt)

    def save_io(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """Input and output layer states per example and beta: object array (M, Nbeta, 2) as in
        va_nnet.py:548-577."""
        mp = self._per_init(self.minpaths, init)
        X = mp[:, :self.NDens].reshape(self.Nbeta, self.M, self.NDnet)
        d0, dl = int(self.structure[0]), int(self.structure[-1])
        out = np.empty((self.M, self.Nbeta, 2), dtype=object)
        for m in range(self.M):
            for i in range(self.Nbeta):
                out[m, i, 0] = X[i, m, :d0].astype(dtype)
                out[m, i, 1] = X[i, m, self.NDnet - dl:].astype(dtype)
        np.save(filename, out, allow_pickle=True)

    def save_Wb(self, W_filename, b_filename, dtype=np.float64, fmt="%.8e", init=None):
        """Weights / biases per beta and layer: object arrays (Nbeta, N-1) as in va_nnet.py:579-634
        (with the slicing done at NDens, which the reference gets wrong)."""
        mp = self._per_init(self.minpaths, init)
        W = np.empty((self.Nbeta, self.N - 1), dtype=object)
        bb = np.empty((self.Nbeta, self.N - 1), dtype=object)
        st = self.structure
        for i in range(self.Nbeta):
            p = mp[i, self.NDens:]
            o = 0
            for n in range(self.N - 1):
                W[i, n] = p[o:o + st[n] * st[n + 1]].reshape(st[n + 1], st[n]).astype(dtype)
                o += st[n] * st[n + 1]
                bb[i, n] = p[o:o + st[n + 1]].astype(dtype)
                o += st[n + 1]
        np.save(W_filename, W, allow_pickle=True)
        np.save(b_filename, bb, allow_pickle=True)

    def save_params(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """(Nbeta, NP) (va_nnet.py:636-647)."""
        mp = self._per_init(self.minpaths, init)
        out = np.array(mp[:, self.NDens:])
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out, fmt=fmt)

    def action_errors_table(self, init=None):
        out = np.zeros((self.Nbeta, 5))
        out[:, 0] = self.beta_array
        out[:, 1] = self._per_init(self.A_array, init)
        out[:, 2] = self._per_init(self.me_array, init)
        out[:, 3] = self._per_init(self.fe_array, init)
        out[:, 4] = out[:, 3] / (self.RF0 * float(self.alpha) ** self.beta_array)
        return out

    def save_action_errors(self, filename, dtype=np.float64, fmt="%.8e", init=None):
        """(Nbeta, 5) rows [beta, A, me, fe, fe/RF] (va_nnet.py:649-663)."""
        out = self.action_errors_table(init)
        if filename.endswith('.npy'):
            np.save(filename, out.astype(dtype))
        else:
            np.savetxt(filename, out, fmt=fmt)

    def gen_xtrace(self):
        return np.random.rand(self._n)
