"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference's neural-network
action (va_nnet) and its analytic adjoint.  Checker for the CUDA path; never imported by the
product.

Follows (file:line under /root/reference/varanneal):
  measurement error   va_nnet.py:117-173   ``me_gaussian`` (scalar RM, RM.shape == (2,), or two matrices)
  model error         va_nnet.py:175-255   ``fe_gaussian`` + ``disc_forwardmap`` :260-264
  parameter layout    va_nnet.py:194-207   P = [W_0 (d_1 x d_0 row-major), b_0, W_1, b_1, ...]
  state layout        va_nnet.py:149-152,212   X.reshape(M, NDnet), layers concatenated
  XP layout           X ++ P[Pidx]          va_nnet.py:468-473
  activation          examples/nnet_twin/nnet_twin_anneal.py:20-22   sigmoid(W x + b)

The reference loops over examples m and layers n in Python; this port batches the examples
into one (M x d_n)(d_n x d_{n+1}) product per layer, which is the same arithmetic in a
different summation order (pinned against the verbatim reference to ~1e-15 relative by
tests/test_oracle.py; gradient pinned by complex-step through the reference).
"""
import numpy as np


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


ACTIVATIONS = {
    # name: (f(z), f'(z) expressed through s = f(z) and z)
    "sigmoid": (_sigmoid, lambda s, z: s * (1.0 - s)),
    "tanh": (np.tanh, lambda s, z: 1.0 - s * s),
    "linear": (lambda z: z, lambda s, z: np.ones_like(s)),
}


def sigmoid(x, W, b):
    """Reference-style activation callable (nnet_twin_anneal.py:20-22)."""
    return 1.0 / (1.0 + np.exp(-(np.dot(W, x) + b)))


class NnetProblem(object):
    def __init__(self, structure, data_in, data_out, Lidx, P, Pidx, RM, act="sigmoid"):
        self.structure = np.asarray(structure, dtype=np.int64)
        self.NL = len(self.structure)
        self.data_in = np.atleast_2d(np.asarray(data_in, dtype=np.float64))
        self.data_out = np.atleast_2d(np.asarray(data_out, dtype=np.float64))
        self.M = self.data_in.shape[0]
        self.NDnet = int(self.structure.sum())
        self.NDens = self.NDnet * self.M
        if Lidx is None:
            Lidx = [np.arange(self.structure[0]), np.arange(self.structure[-1])]
        self.Lin = np.asarray(Lidx[0], dtype=np.int64)
        self.Lout = np.asarray(Lidx[1], dtype=np.int64)
        self.Ltot = self.Lin.size + self.Lout.size
        self.P = np.array(P, dtype=np.float64)
        self.NP = self.P.size
        self.Pidx = np.asarray(Pidx, dtype=np.int64)
        self.NPest = self.Pidx.size
        # RM: scalar, shape (2,), or a pair of matrices [RM_in (Lin, Lin), RM_out (Lout, Lout)]
        # (va_nnet.py:131-140); everything is carried as one matrix per side
        if np.isscalar(RM):
            rin, rout = float(RM) * np.eye(self.Lin.size), float(RM) * np.eye(self.Lout.size)
        elif np.ndim(RM[0]) == 0:
            if len(RM) != 2:
                raise ValueError("RM must be scalar, shape (2,) or a pair of matrices")
            rin, rout = float(RM[0]) * np.eye(self.Lin.size), float(RM[1]) * np.eye(self.Lout.size)
        else:
            rin, rout = np.asarray(RM[0], dtype=np.float64), np.asarray(RM[1], dtype=np.float64)
            if rin.shape != (self.Lin.size,) * 2 or rout.shape != (self.Lout.size,) * 2:
                raise ValueError("RM matrices must be (Lin, Lin) and (Lout, Lout)")
        self.RMin, self.RMout = rin, rout
        self.act, self.dact = ACTIVATIONS[act]
        self.xoff = np.concatenate([[0], np.cumsum(self.structure)])
        d = self.structure
        self.woff, self.boff = [], []
        o = 0
        for n in range(self.NL - 1):
            self.woff.append(o)
            o += d[n] * d[n + 1]
            self.boff.append(o)
            o += d[n + 1]
        assert o == self.NP, "P length does not match structure"
        self.n = self.NDens + self.NPest

    def unpack(self, XP):
        X = XP[:self.NDens].reshape(self.M, self.NDnet)
        if self.NPest == 0:
            p = self.P
        else:
            p = np.array(self.P, dtype=XP.dtype)
            p[self.Pidx] = XP[self.NDens:]
        return X, p

    def layer(self, X, n):
        return X[:, self.xoff[n]:self.xoff[n + 1]]

    def Wb(self, p, n):
        d = self.structure
        W = p[self.woff[n]:self.woff[n] + d[n] * d[n + 1]].reshape(d[n + 1], d[n])
        b = p[self.boff[n]:self.boff[n] + d[n + 1]]
        return W, b

    def me(self, XP):
        X, _ = self.unpack(XP)
        din = self.layer(X, 0)[:, self.Lin] - self.data_in
        dout = self.layer(X, self.NL - 1)[:, self.Lout] - self.data_out
        q = np.einsum("ml,lk,mk->", din, self.RMin, din) + np.einsum("ml,lk,mk->", dout, self.RMout, dout)
        return q / (self.Ltot * self.M)

    def fe(self, XP, RF):
        X, p = self.unpack(XP)
        s = 0.0
        for n in range(self.NL - 1):
            W, b = self.Wb(p, n)
            E = self.layer(X, n + 1) - self.act(self.layer(X, n) @ W.T + b)
            s = s + np.sum(E * E)
        return RF * s / ((self.NDnet - self.structure[0]) * self.M)

    def action(self, XP, RF):
        return self.me(XP) + self.fe(XP, RF)

    def action_grad(self, XP, RF, parts=False):
        XP = np.asarray(XP, dtype=np.float64)
        X, p = self.unpack(XP)
        GX = np.zeros_like(X)
        gp = np.zeros(self.NP)
        cm = 1.0 / (self.Ltot * self.M)
        din = self.layer(X, 0)[:, self.Lin] - self.data_in
        dout = self.layer(X, self.NL - 1)[:, self.Lout] - self.data_out
        me = cm * (np.einsum("ml,lk,mk->", din, self.RMin, din) + np.einsum("ml,lk,mk->", dout, self.RMout, dout))
        g0 = np.zeros((self.M, self.structure[0]))
        g0[:, self.Lin] = cm * din @ (self.RMin + self.RMin.T)
        GX[:, self.xoff[0]:self.xoff[1]] += g0
        gl = np.zeros((self.M, self.structure[-1]))
        gl[:, self.Lout] = cm * dout @ (self.RMout + self.RMout.T)
        GX[:, self.xoff[self.NL - 1]:self.xoff[self.NL]] += gl

        cf = RF / ((self.NDnet - self.structure[0]) * self.M)
        s = 0.0
        d = self.structure
        for n in range(self.NL - 1):
            W, b = self.Wb(p, n)
            Xn = self.layer(X, n)
            Z = Xn @ W.T + b
            S = self.act(Z)
            E = self.layer(X, n + 1) - S
            s += np.sum(E * E)
            lam = 2.0 * cf * E
            GX[:, self.xoff[n + 1]:self.xoff[n + 2]] += lam
            delta = -lam * self.dact(S, Z)
            GX[:, self.xoff[n]:self.xoff[n + 1]] += delta @ W
            gp[self.woff[n]:self.woff[n] + d[n] * d[n + 1]] = (delta.T @ Xn).ravel()
            gp[self.boff[n]:self.boff[n] + d[n + 1]] = delta.sum(axis=0)
        fe = cf * s
        g = np.concatenate([GX.ravel(), gp[self.Pidx]])
        if parts:
            return me + fe, me, fe, g
        return me + fe, g
