"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu 1995;
Zhu, Byrd, Lu, Nocedal 1997; Morales, Nocedal 2011), the algorithm behind
``scipy.optimize.minimize(method='L-BFGS-B')`` that the reference calls with ``bounds=self.bounds``
(_autodiffmin.py:85-86; bounds expansion va_ode.py:582-605).

The algorithm lives in a third-party dependency of the reference (SciPy, unpinned upstream; this
image: SciPy 1.18.1, a C translation of the Fortran L-BFGS-B 3.0 of which only the compiled
extension is present).  This file restates the *published* algorithm -- generalised Cauchy point
(``cauchy``), free-variable subspace minimisation by the direct primal method with the projection
/ backtracking refinement of version 3.0 (``subsm``), More'-Thuente line search with the step
limited to the box (``lnsrlb`` / ``dcsrch``), compact-representation update (``matupd``) with the
``s'y <= eps * (-g's)`` skip, the stopping tests of ``mainlb`` -- and is pinned *behaviourally*:
tests/test_lbfgsb_port.py requires iterate-for-iterate agreement with SciPy (same number of
iterations and function evaluations, same iterates to rounding) on bounded problems.  It is the
specification the device implementation (varanneal_b200/csrc/lbfgsb_bounded.cuh) is checked
against; nothing in the product imports it.

Differences from the Fortran that do not change the mathematics: the 2m x 2m middle-matrix
products (``bmv``) and the subspace system (``formk`` + ``dtrsl``) are solved by dense LU instead
of the Cholesky / LEL^T factorisations, and the reduced matrices over the free set are formed
directly instead of incrementally.
"""
import numpy as np

EPSMCH = 2.220446049250313e-16
BIG = 1.0e10


# ---------------------------------------------------------------------------------------- dcsrch
class _Search(object):
    """MINPACK-2 dcsrch / dcstep state machine (ftol 1e-3, gtol 0.9, xtol 0.1 in lnsrlb)."""

    def __init__(self, f, g, stp, ftol, gtol, xtol, stpmin, stpmax):
        self.ftol, self.gtol, self.xtol, self.stpmin, self.stpmax = ftol, gtol, xtol, stpmin, stpmax
        self.brackt = False
        self.stage = 1
        self.finit, self.ginit = f, g
        self.gtest = ftol * g
        self.width = stpmax - stpmin
        self.width1 = self.width / 0.5
        self.stx, self.fx, self.gx = 0.0, f, g
        self.sty, self.fy, self.gy = 0.0, f, g
        self.stmin = 0.0
        self.stmax = stp + 4.0 * stp
        self.stp = stp

    def step(self, f, g):
        """Returns True on convergence / warning, else leaves the next trial step in self.stp."""
        stp = self.stp
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and g >= 0.0:
            self.stage = 2
        if self.brackt and (stp <= self.stmin or stp >= self.stmax):
            return True
        if self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax:
            return True
        if stp == self.stpmax and f <= ftest and g <= self.gtest:
            return True
        if stp == self.stpmin and (f > ftest or g >= self.gtest):
            return True
        if f <= ftest and abs(g) <= self.gtol * (-self.ginit):
            return True
        if self.stage == 1 and f <= self.fx and f > ftest:
            fm = f - stp * self.gtest
            fxm = self.fx - self.stx * self.gtest
            fym = self.fy - self.sty * self.gtest
            gm = g - self.gtest
            gxm = self.gx - self.gtest
            gym = self.gy - self.gtest
            (self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt) = _dcstep(
                self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm, self.brackt, self.stmin, self.stmax)
            self.fx = fxm + self.stx * self.gtest
            self.fy = fym + self.sty * self.gtest
            self.gx = gxm + self.gtest
            self.gy = gym + self.gtest
        else:
            (self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, self.brackt) = _dcstep(
                self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, g, self.brackt,
                self.stmin, self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= 0.66 * self.width1:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.width1 = self.width
            self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin = min(self.stx, self.sty)
            self.stmax = max(self.stx, self.sty)
        else:
            self.stmin = stp + 1.1 * (stp - self.stx)
            self.stmax = stp + 4.0 * (stp - self.stx)
        stp = max(stp, self.stpmin)
        stp = min(stp, self.stpmax)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or \
                (self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax):
            stp = self.stx
        self.stp = stp
        return False


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = min(stpmax, stpf)
            stpf = max(stpmin, stpf)
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * np.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


# ---------------------------------------------------------------------------------------- pieces
def projgr(x, g, lo, hi):
    """Infinity norm of the projected gradient (projgr)."""
    pg = np.where(g < 0.0, np.maximum(x - hi, g), np.minimum(x - lo, g))
    return float(np.max(np.abs(pg))) if pg.size else 0.0


class _Memory(object):
    """Compact representation: S, Y (n x col, oldest first), S'S, S'Y, theta."""

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.reset()

    def reset(self):
        self.S = np.zeros((self.n, 0))
        self.Y = np.zeros((self.n, 0))
        self.SS = np.zeros((0, 0))
        self.SY = np.zeros((0, 0))
        self.theta = 1.0

    @property
    def col(self):
        return self.S.shape[1]

    def update(self, s, y, rr, dr, stp, dtd):
        if self.col == self.m:
            self.S, self.Y = self.S[:, 1:], self.Y[:, 1:]
            self.SS, self.SY = self.SS[1:, 1:], self.SY[1:, 1:]
        c = self.col
        SS = np.zeros((c + 1, c + 1))
        SY = np.zeros((c + 1, c + 1))
        SS[:c, :c], SY[:c, :c] = self.SS, self.SY
        SY[c, :c] = s @ self.Y              # last row of S'Y
        SY[:c, c] = self.S.T @ y            # (upper part, not used by the Fortran; kept for N)
        SS[:c, c] = self.S.T @ s
        SS[c, :c] = SS[:c, c]
        SS[c, c] = dtd if stp == 1.0 else stp * stp * dtd
        SY[c, c] = dr
        self.S = np.column_stack([self.S, s])
        self.Y = np.column_stack([self.Y, y])
        self.SS, self.SY = SS, SY
        self.theta = rr / dr

    def Minv(self):
        """[[-D, L'], [L, theta S'S]] (the inverse of the middle matrix M of B = theta I - W M W')."""
        c = self.col
        D = np.diag(np.diag(self.SY))
        L = np.tril(self.SY, -1)
        return np.block([[-D, L.T], [L, self.theta * self.SS]])

    def formt_ok(self):
        """formt: T = theta S'S + L D^-1 L' must be positive definite (else the memory is reset)."""
        c = self.col
        if c == 0:
            return True
        D = np.diag(self.SY)
        if np.any(D <= 0.0):
            return False
        L = np.tril(self.SY, -1)
        T = self.theta * self.SS + (L / D) @ L.T
        try:
            np.linalg.cholesky(T)
            return True
        except np.linalg.LinAlgError:
            return False

    def bmv(self, v):
        """M v."""
        return np.linalg.solve(self.Minv(), v)

    def Wrow(self, i):
        return np.concatenate([self.Y[i], self.theta * self.S[i]])


def cauchy(x, g, lo, hi, mem, sbgnrm, iwhere):
    """Generalised Cauchy point.  Returns (xcp, c = W'(xcp - x), iwhere updated, nseg)."""
    n = x.size
    col = mem.col
    theta = mem.theta
    if sbgnrm <= 0.0:
        return x.copy(), np.zeros(2 * col), iwhere, 0
    neg = -g
    has_lo, has_hi = np.isfinite(lo), np.isfinite(hi)
    tl, tu = x - lo, hi - x
    iw = iwhere.copy()
    mov = (iw != 3) & (iw != -1)                 # bounded, not fixed
    xlower = has_lo & (tl <= 0.0)
    xupper = has_hi & (tu <= 0.0)
    new = np.zeros(n, dtype=np.int64)
    new[mov & xlower & (neg <= 0.0)] = 1
    new[mov & ~xlower & xupper & (neg >= 0.0)] = 2
    new[mov & ~xlower & ~xupper & (np.abs(neg) <= 0.0)] = -3
    iw[mov] = new[mov]
    d = np.where((iw != 0) & (iw != -1), 0.0, neg)
    moving = (iw == 0) | (iw == -1)
    f1 = -float(np.sum(d * d))
    p = np.zeros(2 * col)
    if col > 0:
        p[:col] = mem.Y.T @ d
        p[col:] = theta * (mem.S.T @ d)
    # breakpoints
    tb = np.full(n, np.inf)
    m1 = moving & has_lo & (neg < 0.0)
    m2 = moving & has_hi & (neg > 0.0)
    tb[m1] = tl[m1] / (-neg[m1])
    tb[m2] = tu[m2] / neg[m2]
    brk = np.where(np.isfinite(tb))[0]
    nbreak = brk.size
    nfree_nb = int(np.sum(moving & ~np.isfinite(tb)))          # moving variables without a breakpoint
    bnded = not np.any(moving & ~np.isfinite(tb) & (np.abs(neg) > 0.0))
    xcp = x.copy()
    c = np.zeros(2 * col)
    if nbreak == 0 and nfree_nb == 0:
        return xcp, c, iw, 0
    f2 = -theta * f1
    f2_org = f2
    if col > 0:
        f2 -= float(mem.bmv(p) @ p)
    dtm = -f1 / f2
    tsum = 0.0
    nseg = 1
    order = brk[np.argsort(tb[brk], kind="stable")]
    # NOTE: ties are broken by index here; the Fortran heap may pick another order among *equal*
    # breakpoints, which changes nothing (the updates commute when dt = 0).
    tj = 0.0
    nleft = nbreak
    done_all = False
    k = 0
    while nleft > 0:
        ibp = order[k]
        tj0 = tj
        tj = tb[ibp]
        dt = tj - tj0
        if dtm < dt:
            break
        tsum += dt
        nleft -= 1
        k += 1
        dibp = d[ibp]
        d[ibp] = 0.0
        if dibp > 0.0:
            zibp = hi[ibp] - x[ibp]
            xcp[ibp] = hi[ibp]
            iw[ibp] = 2
        else:
            zibp = lo[ibp] - x[ibp]
            xcp[ibp] = lo[ibp]
            iw[ibp] = 1
        if nleft == 0 and nbreak == n:
            dtm = dt
            done_all = True
            break
        nseg += 1
        dibp2 = dibp * dibp
        f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp
        f2 = f2 - theta * dibp2
        if col > 0:
            c += dt * p
            wbp = mem.Wrow(ibp)
            v = mem.bmv(wbp)
            wmc, wmp, wmw = float(c @ v), float(p @ v), float(wbp @ v)
            p -= dibp * wbp
            f1 += dibp * wmc
            f2 += 2.0 * dibp * wmp - dibp2 * wmw
        f2 = max(EPSMCH * f2_org, f2)
        if nleft > 0:
            dtm = -f1 / f2
        elif bnded:
            f1 = f2 = dtm = 0.0
        else:
            dtm = -f1 / f2
    if not done_all:
        if dtm <= 0.0:
            dtm = 0.0
        tsum += dtm
        xcp = xcp + tsum * d
    if col > 0:
        c += dtm * p
    return xcp, c, iw, nseg


def subsm(x, g, lo, hi, xcp, c, free, mem):
    """Subspace minimisation over the free variables at the Cauchy point (cmprlb + formk + subsm),
    direct primal method, then the projection / backtracking step of L-BFGS-B 3.0.  Returns z."""
    theta = mem.theta
    col = mem.col
    idx = np.where(free)[0]
    Wf = np.column_stack([mem.Y[idx], theta * mem.S[idx]])            # Z'W
    r = -theta * (xcp[idx] - x[idx]) - g[idx] + Wf @ mem.bmv(c)       # cmprlb
    N = mem.Minv() - (Wf.T @ Wf) / theta                              # formk
    wv = np.linalg.solve(N, Wf.T @ r)
    d = r / theta + (Wf @ wv) / theta ** 2
    # projection (Morales-Nocedal): clip xcp + d to the box
    z = xcp.copy()
    xk = xcp[idx] + d
    z[idx] = np.minimum(hi[idx], np.maximum(lo[idx], xk))
    projected = bool(np.any((z[idx] == lo[idx]) & np.isfinite(lo[idx])) or
                     np.any((z[idx] == hi[idx]) & np.isfinite(hi[idx])))
    if not projected:
        return z
    dd_p = float((z - x) @ g)
    if dd_p > 0.0:
        # backtrack from the Cauchy point along d to the first bound
        z = xcp.copy()
        alpha = 1.0
        ibd = -1
        for j, k in enumerate(idx):
            dk = d[j]
            temp1 = alpha
            if dk < 0.0 and np.isfinite(lo[k]):
                t2 = lo[k] - z[k]
                if t2 >= 0.0:
                    temp1 = 0.0
                elif dk * alpha < t2:
                    temp1 = t2 / dk
            elif dk > 0.0 and np.isfinite(hi[k]):
                t2 = hi[k] - z[k]
                if t2 <= 0.0:
                    temp1 = 0.0
                elif dk * alpha > t2:
                    temp1 = t2 / dk
            if temp1 < alpha:
                alpha = temp1
                ibd = j
        if alpha < 1.0:
            dk = d[ibd]
            k = idx[ibd]
            if dk > 0.0:
                z[k] = hi[k]
                d[ibd] = 0.0
            elif dk < 0.0:
                z[k] = lo[k]
                d[ibd] = 0.0
        z[idx] = z[idx] + alpha * d
    return z


def minimize(fun, x0, lo=None, hi=None, m=10, ftol=2.220446049250313e-09, gtol=1e-5,
             maxfun=15000, maxiter=15000, maxls=20, callback=None):
    """L-BFGS-B 3.0 as driven by scipy.optimize.minimize (options ftol -> factr*epsmch, gtol ->
    pgtol).  fun(x) -> (f, g).  Returns dict(x, fun, nit, nfev, status, message)."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    lo = np.full(n, -np.inf) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = np.full(n, np.inf) if hi is None else np.asarray(hi, dtype=np.float64)
    has_lo, has_hi = np.isfinite(lo), np.isfinite(hi)
    cnstnd = bool(np.any(has_lo | has_hi))
    boxed = bool(np.all(has_lo & has_hi))
    x = np.minimum(hi, np.maximum(lo, x))                      # active(): project the start point
    iwhere = np.where(has_lo | has_hi, np.where(has_lo & has_hi & (hi - lo <= 0.0), 3, 0), -1)
    mem = _Memory(n, m)
    f, g = fun(x)
    g = np.asarray(g, dtype=np.float64)
    nfev, nit = 1, 0
    sbgnrm = projgr(x, g, lo, hi)
    if sbgnrm <= gtol:
        return dict(x=x, fun=f, nit=0, nfev=1, status=0, message="CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL")
    while True:
        # ---- search direction: Cauchy point + subspace minimisation
        if not cnstnd and mem.col > 0:
            z = x.copy()
            cvec = np.zeros(2 * mem.col)
            free = np.ones(n, dtype=bool)
        else:
            z, cvec, iwhere, _ = cauchy(x, g, lo, hi, mem, sbgnrm, iwhere)
            free = iwhere <= 0
        if np.any(free) and mem.col > 0:
            try:
                z = subsm(x, g, lo, hi, z, cvec, free, mem)
            except np.linalg.LinAlgError:
                mem.reset()
                continue
        d = z - x
        # ---- line search (lnsrlb)
        dtd = float(d @ d)
        dnorm = np.sqrt(dtd)
        stpmx = BIG
        if cnstnd:
            if nit == 0:
                stpmx = 1.0
            else:
                for i in np.where(d != 0.0)[0]:
                    a1 = d[i]
                    if a1 < 0.0 and has_lo[i]:
                        a2 = lo[i] - x[i]
                        if a2 >= 0.0:
                            stpmx = 0.0
                        elif a1 * stpmx < a2:
                            stpmx = a2 / a1
                    elif a1 > 0.0 and has_hi[i]:
                        a2 = hi[i] - x[i]
                        if a2 <= 0.0:
                            stpmx = 0.0
                        elif a1 * stpmx > a2:
                            stpmx = a2 / a1
        stp = min(1.0 / dnorm, stpmx) if (nit == 0 and not boxed) else 1.0
        t, r, fold = x.copy(), g.copy(), f
        gd = float(g @ d)
        gdold = gd
        restart = False
        if gd >= 0.0:
            # not a descent direction (info = -4)
            if mem.col == 0:
                return dict(x=x, fun=f, nit=nit, nfev=nfev, status=2, message="ABNORMAL_TERMINATION_IN_LNSRCH")
            mem.reset()
            continue
        srch = _Search(f, gd, stp, 1e-3, 0.9, 0.1, 0.0, stpmx)
        ifun = 0
        while True:
            stp = srch.stp
            ifun += 1
            iback = ifun - 1
            if iback >= maxls:
                restart = True
                break
            x = z.copy() if stp == 1.0 else stp * d + t
            f, g = fun(x)
            g = np.asarray(g, dtype=np.float64)
            nfev += 1
            gd = float(g @ d)
            if srch.step(f, gd):
                break
        if restart:
            x, g, f = t, r, fold
            if mem.col == 0:
                return dict(x=x, fun=f, nit=nit, nfev=nfev, status=2, message="ABNORMAL_TERMINATION_IN_LNSRCH")
            mem.reset()
            continue
        stp = srch.stp
        nit += 1
        if callback is not None:
            callback(x.copy(), f)
        sbgnrm = projgr(x, g, lo, hi)
        # SciPy's driver looks at its limits when setulb hands back NEW_X, i.e. before setulb's own
        # convergence tests run on re-entry (_lbfgsb_py.py: n_iterations >= maxiter, nfev > maxfun)
        if nit >= maxiter or nfev > maxfun:
            return dict(x=x, fun=f, nit=nit, nfev=nfev, status=1, message="STOP: TOTAL NO. OF F,G EVALUATIONS / ITERATIONS EXCEEDS LIMIT")
        if sbgnrm <= gtol:
            return dict(x=x, fun=f, nit=nit, nfev=nfev, status=0, message="CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL")
        ddum = max(abs(fold), abs(f), 1.0)
        if (fold - f) <= ftol * ddum:
            return dict(x=x, fun=f, nit=nit, nfev=nfev, status=0, message="CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH")
        # ---- update
        y = g - r
        rr = float(y @ y)
        if stp == 1.0:
            dr = gd - gdold
            ddum = -gdold
            s = d.copy()
        else:
            dr = (gd - gdold) * stp
            s = stp * d
            ddum = -gdold * stp
        if dr <= EPSMCH * ddum:
            continue                                            # skip the update
        mem.update(s, y, rr, dr, stp, dtd)
        if not mem.formt_ok():
            mem.reset()
