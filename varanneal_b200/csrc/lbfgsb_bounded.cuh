// Bounded problems: the "B" of L-BFGS-B on the device.
//
// The reference hands its box to SciPy (scipy.optimize.minimize(method='L-BFGS-B',
// bounds=self.bounds), _autodiffmin.py:85-86; expansion of the D + NPest pairs at
// va_ode.py:582-605; the NaKL tutorial uses it).  This file is L-BFGS-B 3.0's treatment of the
// bounds -- generalised Cauchy point along the projected steepest-descent path, subspace
// minimisation over the variables free at that point (direct primal method) with the projection /
// backtracking refinement of version 3.0, line search limited to the box -- batched over paths.
// Specification: oracle/lbfgsb_port.py, which tests/test_lbfgsb_port.py pins iterate for iterate
// against SciPy; the device path is pinned against both in tests/test_gpu_bounded.py.
//
// Work per iteration and path, all inside the minimiser's cycle (lbfgs.cu), nothing on the host:
//   lbb_update        x <- xt, g <- gt, (s, y) into the history; s.Y_j, s.S_j, y.y
//   lbb_hist          S'S, S'Y, theta, col / head; T = theta S'S + L D^-1 L' must be SPD (formt)
//   lbb_cauchy_prep   classify variables, d = -g on the moving ones, breakpoints t_i,
//                     p = W'd, f' = -d'd                                  (one history pass)
//   lbb_cauchy_loop   one CTA per path walks the breakpoints in increasing order (block-wide
//                     arg-min, per-thread cached minima) updating p, c, f', f''; stops at the
//                     first segment that holds the minimiser of the quadratic model
//   lbb_gram_free     W'ZZ'W over the variables free at the Cauchy point   (one history pass)
//   lbb_resid         x^c, r = -Z'(g + theta (x^c - x) - W M c), W'Z r      (one history pass)
//   lbb_subsolve      N = M^-1 - W'ZZ'W / theta (2m x 2m, LU), v = N^-1 W'Z r
//   lbb_step          d^ = r / theta + Z'W v / theta^2, z = P(x^c + d^); (z - x).g and the
//                     feasible step for the fall-back                       (one history pass)
//   lbb_decide / lbb_backtrack   positive directional derivative after projection: z = x^c +
//                     alpha d^ instead (first bound hit), as L-BFGS-B 3.0
//   lbb_dir           d = z - x, d.d, g.d, largest feasible step
// Work vectors besides the minimiser's: Z (x^c, then z), T (breakpoints; afterwards the markers
// -1 = fixed by the Cauchy search, -2 = held at a bound from the start), R (r, then d^).
#pragma once

namespace {

constexpr int M2 = 2 * MMAX;
constexpr int NPAIR = M2 * (M2 + 1) / 2;          // upper triangle of the 2m x 2m reduced Gram matrix
constexpr int NPB = NPAIR + M2 + 8;               // per-chunk partials of the bounded passes
constexpr double TB_CROSSED = -1.0, TB_HELD = -2.0;
#define LBB_INF (__longlong_as_double(0x7ff0000000000000LL))

// ---- small dense algebra, one thread, matrices of order <= 2m with leading dimension M2
__device__ bool lbb_lu_factor(double* A, int* piv, int n) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[k * M2 + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(A[i * M2 + k]);
      if (v > best) { best = v; p = i; }
    }
    if (!(best > 0.0) || !isfinite(best)) return false;
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; ++j) { const double t = A[k * M2 + j]; A[k * M2 + j] = A[p * M2 + j]; A[p * M2 + j] = t; }
    const double inv = 1.0 / A[k * M2 + k];
    for (int i = k + 1; i < n; ++i) {
      const double l = A[i * M2 + k] * inv;
      A[i * M2 + k] = l;
      for (int j = k + 1; j < n; ++j) A[i * M2 + j] = fma(-l, A[k * M2 + j], A[i * M2 + j]);
    }
  }
  return true;
}
__device__ void lbb_lu_solve(const double* A, const int* piv, int n, double* b) {
  // the factorisation swaps whole rows (multipliers included), so P b comes first, then L, then U
  for (int k = 0; k < n; ++k) {
    const int p = piv[k];
    if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
  }
  for (int k = 0; k < n; ++k)
    for (int i = k + 1; i < n; ++i) b[i] = fma(-A[i * M2 + k], b[k], b[i]);
  for (int k = n - 1; k >= 0; --k) {
    double v = b[k];
    for (int j = k + 1; j < n; ++j) v = fma(-A[k * M2 + j], b[j], v);
    b[k] = v / A[k * M2 + k];
  }
}
// [[-D, L'], [L, theta S'S]] in history order (oldest first): the inverse of the middle matrix M
__device__ void lbb_minv(const LbPath& s, int m, double* A) {
  const int col = s.col;
  for (int a = 0; a < col; ++a) {
    const int ia = (s.head + a) % m;
    for (int b = 0; b < col; ++b) {
      const int ib = (s.head + b) % m;
      A[a * M2 + b] = (a == b) ? -s.SY[ia * MMAX + ia] : 0.0;
      A[(col + a) * M2 + (col + b)] = s.theta * s.SS[ia * MMAX + ib];
      const double l = (a > b) ? s.SY[ia * MMAX + ib] : 0.0;      // L: s_newer . y_older
      A[(col + a) * M2 + b] = l;
      A[b * M2 + (col + a)] = l;
    }
  }
}
__device__ __forceinline__ void lbb_forget(LbPath& s) { s.col = 0; s.head = 0; s.theta = 1.0; }

__device__ __forceinline__ bool lbb_has_lo(double l) { return l > -DBL_MAX; }
__device__ __forceinline__ bool lbb_has_hi(double h) { return h < DBL_MAX; }

// are all variables bounded on both sides?  (lnsrlb: first step 1/||d|| unless the problem is boxed)
__global__ void lbb_boxed_kernel(const double* __restrict__ lo, const double* __restrict__ hi, long long n, int* boxed) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  int mine = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    if (!(lbb_has_lo(lo[i]) && lbb_has_hi(hi[i]))) mine = 1;
  if (mine) atomicOr(&bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) *boxed = bad ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// accepted step: x <- xt, g <- gt; (s, y) into slot p; partials [0..M) s.Y_j, [M..2M) s.S_j, 2M: y.y
__global__ void __launch_bounds__(NT) lbb_update_kernel(
    double* __restrict__ X, double* __restrict__ G, const double* __restrict__ XT, const double* __restrict__ GT,
    const double* __restrict__ Dv, double* __restrict__ S, double* __restrict__ Y, long long ld, long long n,
    long long hstride, const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[M2 + 1];
  __shared__ int op[M2 + 1];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!s.accepted) return;
  const bool upd = s.do_update != 0;
  const int p = s.pslot, col = s.col;
  const double stp = s.stp;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double acc[M2 + 1];
#pragma unroll
  for (int k = 0; k <= M2; ++k) acc[k] = 0.0;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double xt = XT[off + i], gt = GT[off + i];
    if (upd) {
      const double dv = Dv[off + i];
      const double sv = (stp == 1.0) ? dv : stp * dv;
      const double yv = gt - G[off + i];
#pragma unroll
      for (int j = 0; j < MMAX; ++j) {
        if (j < m && (j < col || col == m) && j != p) {
          acc[j] = fma(sv, Y[(long long)j * hstride + off + i], acc[j]);
          acc[MMAX + j] = fma(sv, S[(long long)j * hstride + off + i], acc[MMAX + j]);
        }
      }
      acc[M2] = fma(yv, yv, acc[M2]);
      S[(long long)p * hstride + off + i] = sv;
      Y[(long long)p * hstride + off + i] = yv;
    }
    X[off + i] = xt;
    G[off + i] = gt;
  }
  for (int k = threadIdx.x; k <= M2; k += NT) op[k] = RED_SUM;
  block_reduce<M2 + 1>(acc, op, res, scratch);
  __syncthreads();
  for (int k = threadIdx.x; k <= M2; k += NT) part[((long long)b * nchunk + blockIdx.x) * NPB + k] = res[k];
}

// matupd + formt for the bounded path; one CTA (a warp) per path, thread 0 works
__global__ void lbb_hist_kernel(LbPath* st, const double* part, int nchunk, int m) {
  const int b = blockIdx.x;
  LbPath& s = st[b];
  if (!s.accepted || s.done || !s.do_update) return;
  __shared__ double sum[M2 + 1];
  for (int k = threadIdx.x; k <= M2; k += blockDim.x) {
    double a = 0.0;
    for (int c = 0; c < nchunk; ++c) a += part[((long long)b * nchunk + c) * NPB + k];
    sum[k] = a;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int p = s.pslot;
  for (int j = 0; j < m; ++j) {
    if (j == p) continue;
    s.SY[p * MMAX + j] = sum[j];                       // s_p . y_j  (new row of S'Y)
    s.SS[p * MMAX + j] = sum[MMAX + j];
    s.SS[j * MMAX + p] = sum[MMAX + j];
  }
  s.SY[p * MMAX + p] = s.dr;                           // from the line search, as matupd does
  s.SS[p * MMAX + p] = (s.stp == 1.0) ? s.dtd : s.stp * s.stp * s.dtd;
  s.theta = sum[M2] / s.dr;
  if (s.col < m) s.col += 1;
  else s.head = (s.head + 1) % m;
  // formt: T = theta S'S + L D^-1 L' must be positive definite, else the memory is dropped
  const int col = s.col;
  double T[MMAX * MMAX];
  bool ok = true;
  for (int a = 0; a < col && ok; ++a) {
    const int ia = (s.head + a) % m;
    if (!(s.SY[ia * MMAX + ia] > 0.0)) ok = false;
    for (int c2 = a; c2 < col; ++c2) {
      const int ic = (s.head + c2) % m;
      double v = s.theta * s.SS[ia * MMAX + ic];
      for (int k = 0; k < a; ++k) {                    // k < min(a, c2) = a
        const int ik = (s.head + k) % m;
        v += s.SY[ia * MMAX + ik] * s.SY[ic * MMAX + ik] / s.SY[ik * MMAX + ik];
      }
      T[a * MMAX + c2] = v;
      T[c2 * MMAX + a] = v;
    }
  }
  for (int k = 0; k < col && ok; ++k) {                // Cholesky
    double dkk = T[k * MMAX + k];
    for (int j = 0; j < k; ++j) dkk -= T[k * MMAX + j] * T[k * MMAX + j];
    if (!(dkk > 0.0) || !isfinite(dkk)) { ok = false; break; }
    dkk = sqrt(dkk);
    T[k * MMAX + k] = dkk;
    for (int i = k + 1; i < col; ++i) {
      double v = T[i * MMAX + k];
      for (int j = 0; j < k; ++j) v -= T[i * MMAX + j] * T[k * MMAX + j];
      T[i * MMAX + k] = v / dkk;
    }
  }
  if (!ok) lbb_forget(s);
}

// ---------------------------------------------------------------------------------------------
// Cauchy search, part 1: classification, d, breakpoints, p = W'd, f' = -d'd.
// partials: [0..2M) p (Y part, then S part without theta), 2M: f1, 2M+1: #breakpoints,
//           2M+2: 1 if some moving variable has no breakpoint (the path is unbounded)
__global__ void __launch_bounds__(NT) lbb_cauchy_prep_kernel(
    const double* __restrict__ X, const double* __restrict__ G, double* __restrict__ Dv, double* __restrict__ T,
    const double* __restrict__ S, const double* __restrict__ Y, long long ld, long long n, long long hstride,
    const double* __restrict__ lo, const double* __restrict__ hi, const LbPath* __restrict__ st, int m, int nchunk,
    double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[M2 + 3];
  __shared__ int op[M2 + 3];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) return;
  const int col = s.col, head = s.head;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double acc[M2 + 3];
#pragma unroll
  for (int k = 0; k < M2 + 3; ++k) acc[k] = 0.0;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double x = X[off + i], g = G[off + i], l = lo[i], h = hi[i];
    const bool hl = lbb_has_lo(l), hh = lbb_has_hi(h);
    const double neg = -g;
    int iw = (hl || hh) ? 0 : -1;
    const double tl = x - l, tu = h - x;
    if (hl && hh && h - l <= 0.0) iw = 3;
    else if (hl || hh) {
      const bool xlower = hl && tl <= 0.0, xupper = hh && tu <= 0.0;
      if (xlower && neg <= 0.0) iw = 1;
      else if (xupper && neg >= 0.0) iw = 2;
      else if (fabs(neg) <= 0.0) iw = -3;
    }
    double d = 0.0, tb = LBB_INF;
    if (iw > 0) tb = TB_HELD;
    else if (iw != -3) {
      d = neg;
      acc[M2] = fma(-neg, neg, acc[M2]);
#pragma unroll
      for (int k = 0; k < MMAX; ++k) {
        if (k < col) {
          const int j = (head + k) % m;
          acc[k] = fma(Y[(long long)j * hstride + off + i], neg, acc[k]);
          acc[MMAX + k] = fma(S[(long long)j * hstride + off + i], neg, acc[MMAX + k]);
        }
      }
      if (hl && neg < 0.0) tb = tl / (-neg);
      else if (hh && neg > 0.0) tb = tu / neg;
      else if (fabs(neg) > 0.0) acc[M2 + 2] = 1.0;
      if (tb < LBB_INF) acc[M2 + 1] += 1.0;
    }
    Dv[off + i] = d;
    T[off + i] = tb;
  }
  for (int k = threadIdx.x; k < M2 + 3; k += NT) op[k] = (k == M2 + 2) ? RED_MAX : RED_SUM;
  __syncthreads();
  block_reduce<M2 + 3>(acc, op, res, scratch);
  __syncthreads();
  for (int k = threadIdx.x; k < M2 + 3; k += NT) part[((long long)b * nchunk + blockIdx.x) * NPB + k] = res[k];
}

// Cauchy search, part 2: walk the breakpoints.  One CTA per path.
constexpr int CLT = 256;
__global__ void __launch_bounds__(CLT) lbb_cauchy_loop_kernel(
    const double* __restrict__ X, const double* __restrict__ Dv, double* __restrict__ T,
    const double* __restrict__ S, const double* __restrict__ Y, long long ld, long long n, long long hstride,
    const double* __restrict__ lo, const double* __restrict__ hi, LbPath* st, int m, int nchunk,
    const double* __restrict__ part) {
  const int b = blockIdx.x, tid = threadIdx.x;
  LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) return;
  __shared__ double A[M2 * M2];                // LU of M^-1
  __shared__ int piv[M2];
  __shared__ double p[M2], c[M2], wbp[M2], v[M2];
  __shared__ double red_v[CLT / 32];
  __shared__ long long red_i[CLT / 32];
  __shared__ double sh_t;
  __shared__ long long sh_i;
  __shared__ int sh_stop, sh_ok;
  __shared__ double sums[M2 + 3];
  const int col = s.col, head = s.head, n2 = 2 * col;
  const double theta = s.theta;
  const long long off = (long long)b * ld;
  for (int k = tid; k < M2 + 3; k += CLT) {
    double a = (k == M2 + 2) ? 0.0 : 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
      const double u = part[((long long)b * nchunk + ch) * NPB + k];
      a = (k == M2 + 2) ? fmax(a, u) : a + u;
    }
    sums[k] = a;
  }
  __syncthreads();
  double f1 = sums[M2], f2 = 0.0, f2_org = 0.0, dtm = 0.0, tsum = 0.0, tj = 0.0;
  long long nleft = (long long)(sums[M2 + 1] + 0.5);
  const long long nbreak = nleft;
  const bool bnded = sums[M2 + 2] == 0.0;
  if (tid == 0) {
    sh_ok = 1;
    for (int k = 0; k < col; ++k) { p[k] = sums[k]; p[col + k] = theta * sums[MMAX + k]; }
    for (int k = 0; k < n2; ++k) c[k] = 0.0;
    f2 = -theta * f1;
    f2_org = f2;
    if (col > 0) {
      lbb_minv(s, m, A);
      if (!lbb_lu_factor(A, piv, n2)) sh_ok = 0;
      else {
        for (int k = 0; k < n2; ++k) v[k] = p[k];
        lbb_lu_solve(A, piv, n2, v);
        for (int k = 0; k < n2; ++k) f2 -= v[k] * p[k];
      }
    }
    dtm = -f1 / f2;
  }
  __syncthreads();
  if (!sh_ok) {                                // singular middle matrix: drop the memory, redo the direction
    if (tid == 0) { lbb_forget(s); s.accepted = 0; s.redo_dir = 1; s.abort_dir = 1; }
    return;
  }
  // per-thread minimum over its own (strided) breakpoints still to come
  double my_t = LBB_INF;
  long long my_i = -1;
  auto rescan = [&]() {
    my_t = LBB_INF; my_i = -1;
    for (long long i = tid; i < n; i += CLT) {
      const double t = T[off + i];
      if (t >= 0.0 && t < my_t) { my_t = t; my_i = i; }
    }
  };
  if (nleft > 0) rescan();
  bool all_fixed = false;
  while (nleft > 0) {
    // block-wide arg-min (smallest t, then smallest index)
    double bt = my_t;
    long long bi = my_i;
    for (int sft = 16; sft > 0; sft >>= 1) {
      const double ot = __shfl_down_sync(0xffffffffu, bt, sft);
      const long long oi = __shfl_down_sync(0xffffffffu, bi, sft);
      if (ot < bt || (ot == bt && oi >= 0 && (bi < 0 || oi < bi))) { bt = ot; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bt; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < CLT / 32; ++w)
        if (red_v[w] < bt || (red_v[w] == bt && red_i[w] >= 0 && (bi < 0 || red_i[w] < bi))) { bt = red_v[w]; bi = red_i[w]; }
      sh_t = bt; sh_i = bi;
      sh_stop = (bi < 0 || dtm < bt - tj) ? 1 : 0;      // the minimiser lies inside this segment
    }
    __syncthreads();
    if (sh_stop) break;
    const long long ibp = sh_i;
    if (tid < n2) {                                      // row of W at the breakpoint variable
      const int k = tid < col ? tid : tid - col;
      const int j = (head + k) % m;
      wbp[tid] = tid < col ? Y[(long long)j * hstride + off + ibp] : theta * S[(long long)j * hstride + off + ibp];
    }
    if (tid == (int)(ibp % CLT)) { T[off + ibp] = TB_CROSSED; rescan(); }
    __syncthreads();
    if (tid == 0) {
      const double dt = sh_t - tj;
      tj = sh_t;
      tsum += dt;
      nleft -= 1;
      const double dibp = Dv[off + ibp];
      const double zibp = (dibp > 0.0) ? hi[ibp] - X[off + ibp] : lo[ibp] - X[off + ibp];
      if (nleft == 0 && nbreak == n) { dtm = dt; all_fixed = true; }
      else {
        const double dibp2 = dibp * dibp;
        f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
        f2 = f2 - theta * dibp2;
        if (col > 0) {
          for (int k = 0; k < n2; ++k) { c[k] = fma(dt, p[k], c[k]); v[k] = wbp[k]; }
          lbb_lu_solve(A, piv, n2, v);
          double wmc = 0.0, wmp = 0.0, wmw = 0.0;
          for (int k = 0; k < n2; ++k) { wmc += c[k] * v[k]; wmp += p[k] * v[k]; wmw += wbp[k] * v[k]; }
          for (int k = 0; k < n2; ++k) p[k] -= dibp * wbp[k];
          f1 += dibp * wmc;
          f2 += 2.0 * dibp * wmp - dibp2 * wmw;
        }
        f2 = fmax(EPSMCH * f2_org, f2);
        if (nleft > 0) dtm = -f1 / f2;
        else if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
        else dtm = -f1 / f2;
      }
      sh_stop = all_fixed ? 1 : 0;
    }
    nleft -= (tid == 0) ? 0 : 1;                         // every thread tracks the count
    __syncthreads();
    if (sh_stop) break;
  }
  if (tid == 0) {
    if (!all_fixed) {
      if (dtm <= 0.0) dtm = 0.0;
      tsum += dtm;
    }
    for (int k = 0; k < n2; ++k) { c[k] = fma(dtm, p[k], c[k]); v[k] = c[k]; }
    if (col > 0) lbb_lu_solve(A, piv, n2, v);
    for (int k = 0; k < n2; ++k) { s.cvec[k] = c[k]; s.Mc[k] = v[k]; }
    s.tsum = tsum;
  }
  // how many variables are free at the Cauchy point
  int cnt = 0;
  __syncthreads();
  for (long long i = tid; i < n; i += CLT) cnt += (T[off + i] >= 0.0) ? 1 : 0;
  for (int sft = 16; sft > 0; sft >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, sft);
  __shared__ int cred[CLT / 32];
  if ((tid & 31) == 0) cred[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    long long nf = 0;
    for (int w = 0; w < CLT / 32; ++w) nf += cred[w];
    s.nfree = nf > 2000000000LL ? 2000000000 : (int)nf;
    s.skip_sub = (nf == 0 || col == 0) ? 1 : 0;
    s.need_bt = 0;
  }
}

// value of variable i at the Cauchy point
__device__ __forceinline__ double lbb_xcp(double x, double d, double tb, double tsum, double l, double h) {
  if (tb >= 0.0) return fma(tsum, d, x);
  if (tb == TB_CROSSED) return d > 0.0 ? h : l;
  return x;
}

// W'ZZ'W (unscaled: Y and S columns in history order) over the free variables; upper triangle,
// pair q = (a, b >= a) -> partial q
constexpr int GTE = 32;                           // elements per tile
__global__ void __launch_bounds__(NT) lbb_gram_free_kernel(
    const double* __restrict__ T, const double* __restrict__ S, const double* __restrict__ Y, long long ld,
    long long n, long long hstride, const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part) {
  __shared__ double tile[GTE][M2 + 1];
  const int b = blockIdx.y, tid = threadIdx.x;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir || s.skip_sub) return;
  const int col = s.col, head = s.head, n2 = 2 * col;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  // this thread's pair
  int pa = -1, pb = -1;
  {
    int q = tid, a = 0;
    while (a < n2 && q >= n2 - a) { q -= n2 - a; ++a; }
    if (a < n2) { pa = a; pb = a + q; }
  }
  double acc = 0.0;
  for (long long e0 = r.i0; e0 < r.i1; e0 += GTE) {
    __syncthreads();
    for (int t = tid; t < GTE * n2; t += NT) {
      const int vec = t / GTE, e = t % GTE;
      const long long i = e0 + e;
      double val = 0.0;
      if (i < r.i1 && T[off + i] >= 0.0) {
        const int k = vec < col ? vec : vec - col;
        const int j = (head + k) % m;
        val = vec < col ? Y[(long long)j * hstride + off + i] : S[(long long)j * hstride + off + i];
      }
      tile[e][vec] = val;
    }
    __syncthreads();
    if (pa >= 0) {
#pragma unroll 8
      for (int e = 0; e < GTE; ++e) acc = fma(tile[e][pa], tile[e][pb], acc);
    }
  }
  if (pa >= 0) {
    // position of (pa, pb) in the packed upper triangle of order M2 (fixed layout whatever col is)
    const int q = pa * M2 - pa * (pa - 1) / 2 + (pb - pa);
    part[((long long)b * nchunk + blockIdx.x) * NPB + q] = acc;
  }
}

// x^c -> Z; r = -theta (x^c - x) - g + W (M c) on the free variables -> R; partials W'Z r at
// [NPAIR .. NPAIR + 2M) (Y part, then S part without theta)
__global__ void __launch_bounds__(NT) lbb_resid_kernel(
    const double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Dv,
    const double* __restrict__ T, double* __restrict__ Z, double* __restrict__ R, const double* __restrict__ S,
    const double* __restrict__ Y, long long ld, long long n, long long hstride, const double* __restrict__ lo,
    const double* __restrict__ hi, const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[M2];
  __shared__ int op[M2];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir) return;
  const int col = s.col, head = s.head;
  const double theta = s.theta, tsum = s.tsum;
  const bool sub = !s.skip_sub;
  double mc[M2];
#pragma unroll
  for (int k = 0; k < M2; ++k) mc[k] = (sub && k < 2 * col) ? s.Mc[k] : 0.0;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double acc[M2];
#pragma unroll
  for (int k = 0; k < M2; ++k) acc[k] = 0.0;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double x = X[off + i], tb = T[off + i];
    const double xc = lbb_xcp(x, Dv[off + i], tb, tsum, lo[i], hi[i]);
    Z[off + i] = xc;
    if (!sub) continue;
    double rv = 0.0;
    if (tb >= 0.0) {
      double yk[MMAX], sk[MMAX];
      rv = -theta * (xc - x) - G[off + i];
#pragma unroll
      for (int k = 0; k < MMAX; ++k) {
        if (k < col) {
          const int j = (head + k) % m;
          yk[k] = Y[(long long)j * hstride + off + i];
          sk[k] = S[(long long)j * hstride + off + i];
          rv = fma(yk[k], mc[k], rv);
          rv = fma(theta * sk[k], mc[col + k], rv);
        }
      }
#pragma unroll
      for (int k = 0; k < MMAX; ++k) {
        if (k < col) {
          acc[k] = fma(yk[k], rv, acc[k]);
          acc[MMAX + k] = fma(sk[k], rv, acc[MMAX + k]);
        }
      }
    }
    R[off + i] = rv;
  }
  if (!sub) return;
  for (int k = threadIdx.x; k < M2; k += NT) op[k] = RED_SUM;
  __syncthreads();
  block_reduce<M2>(acc, op, res, scratch);
  __syncthreads();
  for (int k = threadIdx.x; k < M2; k += NT) part[((long long)b * nchunk + blockIdx.x) * NPB + NPAIR + k] = res[k];
}

// N = M^-1 - W'ZZ'W / theta, v = N^-1 W'Z r; one CTA per path
__global__ void lbb_subsolve_kernel(LbPath* st, const double* part, int nchunk, int m) {
  const int b = blockIdx.x;
  LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir || s.skip_sub) return;
  __shared__ double gram[NPAIR], wz[M2];
  __shared__ double A[M2 * M2];
  __shared__ int piv[M2];
  const int col = s.col, n2 = 2 * col;
  for (int q = threadIdx.x; q < NPAIR + M2; q += blockDim.x) {
    // only entries of the leading 2 col x 2 col block were written
    int a = 0, rem = q;
    bool live;
    if (q < NPAIR) {
      while (rem >= M2 - a) { rem -= M2 - a; ++a; }
      live = a < n2 && a + rem < n2;
    } else {
      const int k = q - NPAIR;
      live = (k < MMAX ? k : k - MMAX) < col;
    }
    double acc = 0.0;
    if (live)
      for (int c = 0; c < nchunk; ++c) acc += part[((long long)b * nchunk + c) * NPB + q];
    if (q < NPAIR) gram[q] = acc; else wz[q - NPAIR] = acc;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double theta = s.theta;
  lbb_minv(s, m, A);
  double rhs[M2];
  for (int a = 0; a < n2; ++a) {
    const double sa = a < col ? 1.0 : theta;             // W = [Y, theta S]
    rhs[a] = sa * (a < col ? wz[a] : wz[MMAX + (a - col)]);
    for (int c2 = a; c2 < n2; ++c2) {
      const double sc = c2 < col ? 1.0 : theta;
      const double g = sa * sc * gram[a * M2 - a * (a - 1) / 2 + (c2 - a)] / theta;
      A[a * M2 + c2] -= g;
      if (c2 != a) A[c2 * M2 + a] -= g;
    }
  }
  if (!lbb_lu_factor(A, piv, n2)) {
    lbb_forget(s); s.accepted = 0; s.redo_dir = 1; s.abort_dir = 1;
    return;
  }
  lbb_lu_solve(A, piv, n2, rhs);
  for (int k = 0; k < n2; ++k) s.wv[k] = rhs[k];
}

// d^ = r / theta + Z'W v / theta^2 -> R; z = P(x^c + d^) -> Z;
// partials at [NPAIR + 2M ..): +0 projected (max), +1 (z - x).g (sum), +2 alpha (min), +3 index of that minimum
__global__ void __launch_bounds__(NT) lbb_step_kernel(
    const double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ T,
    double* __restrict__ Z, double* __restrict__ R, const double* __restrict__ S, const double* __restrict__ Y,
    long long ld, long long n, long long hstride, const double* __restrict__ lo, const double* __restrict__ hi,
    const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[3];
  __shared__ double amin[NT];
  __shared__ long long aidx[NT];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir || s.skip_sub) return;
  const int col = s.col, head = s.head;
  const double theta = s.theta;
  double wv[M2];
#pragma unroll
  for (int k = 0; k < M2; ++k) wv[k] = (k < 2 * col) ? s.wv[k] : 0.0;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double v[3] = {0.0, 0.0, 1.0};
  long long my_idx = -1;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double x = X[off + i], xc = Z[off + i];
    double z = xc;
    if (T[off + i] >= 0.0) {
      double wvi = 0.0;
#pragma unroll
      for (int k = 0; k < MMAX; ++k) {
        if (k < col) {
          const int j = (head + k) % m;
          wvi = fma(Y[(long long)j * hstride + off + i], wv[k], wvi);
          wvi = fma(theta * S[(long long)j * hstride + off + i], wv[col + k], wvi);
        }
      }
      const double dk = (R[off + i] + wvi / theta) / theta;
      R[off + i] = dk;
      const double l = lo[i], h = hi[i];
      const bool hl = lbb_has_lo(l), hh = lbb_has_hi(h);
      z = xc + dk;
      if (hl) z = fmax(l, z);
      if (hh) z = fmin(h, z);
      if ((hl && z == l) || (hh && z == h)) v[0] = 1.0;
      // feasible fraction of d^ from x^c (the fall-back of subsm)
      double t1 = 1.0;
      if (dk < 0.0 && hl) {
        const double t2 = l - xc;
        if (t2 >= 0.0) t1 = 0.0;
        else if (dk < t2) t1 = t2 / dk;
      } else if (dk > 0.0 && hh) {
        const double t2 = h - xc;
        if (t2 <= 0.0) t1 = 0.0;
        else if (dk > t2) t1 = t2 / dk;
      }
      if (t1 < v[2]) { v[2] = t1; my_idx = i; }
      Z[off + i] = z;
    }
    v[1] = fma(z - x, G[off + i], v[1]);
  }
  // arg-min of alpha: smallest value, then smallest index
  amin[threadIdx.x] = v[2];
  aidx[threadIdx.x] = my_idx;
  const int op[3] = {RED_MAX, RED_SUM, RED_MIN};
  block_reduce<3>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 1.0;
    long long ai = -1;
    for (int t = 0; t < NT; ++t)
      if (aidx[t] >= 0 && (amin[t] < a || (amin[t] == a && (ai < 0 || aidx[t] < ai)))) { a = amin[t]; ai = aidx[t]; }
    double* o = part + ((long long)b * nchunk + blockIdx.x) * NPB + NPAIR + M2;
    o[0] = res[0]; o[1] = res[1]; o[2] = a; o[3] = (double)ai;
  }
}

__global__ void lbb_decide_kernel(LbPath* st, const double* part, int nchunk) {
  const int b = blockIdx.x;
  LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir || s.skip_sub) return;
  if (threadIdx.x != 0) return;
  double proj = 0.0, ddp = 0.0, alpha = 1.0, idx = -1.0;
  for (int c = 0; c < nchunk; ++c) {
    const double* o = part + ((long long)b * nchunk + c) * NPB + NPAIR + M2;
    proj = fmax(proj, o[0]);
    ddp += o[1];
    if (o[3] >= 0.0 && o[2] < alpha) { alpha = o[2]; idx = o[3]; }
  }
  s.need_bt = (proj > 0.0 && ddp > 0.0) ? 1 : 0;
  s.alpha_bt = alpha;
  s.ibd = (long long)idx;
}

// positive directional derivative after the projection: z = x^c + alpha d^, the variable that
// limits alpha exactly on its bound
__global__ void __launch_bounds__(NT) lbb_backtrack_kernel(
    const double* __restrict__ X, const double* __restrict__ Dv, const double* __restrict__ T,
    double* __restrict__ Z, const double* __restrict__ R, long long ld, long long n,
    const double* __restrict__ lo, const double* __restrict__ hi, const LbPath* __restrict__ st, int nchunk) {
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir || s.skip_sub || !s.need_bt) return;
  const double alpha = s.alpha_bt, tsum = s.tsum;
  const long long ibd = s.ibd;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double tb = T[off + i];
    const double xc = lbb_xcp(X[off + i], Dv[off + i], tb, tsum, lo[i], hi[i]);
    double z = xc;
    if (tb >= 0.0) {
      const double dk = R[off + i];
      if (alpha < 1.0 && i == ibd && dk != 0.0) z = dk > 0.0 ? hi[i] : lo[i];
      else z = fma(alpha, dk, xc);
    }
    Z[off + i] = z;
  }
}

// d = z - x; partials [d.d, g.d, largest feasible step (1 at the first iteration)]
__global__ void __launch_bounds__(NT) lbb_dir_kernel(
    const double* __restrict__ X, const double* __restrict__ G, double* __restrict__ Dv,
    const double* __restrict__ Z, long long ld, long long n, const double* __restrict__ lo,
    const double* __restrict__ hi, const LbPath* __restrict__ st, int nchunk, double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[3];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done || s.abort_dir) return;
  const bool first_it = s.iter == 0;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double v[3] = {0.0, 0.0, first_it ? 1.0 : BIG};
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double x = X[off + i];
    const double d = Z[off + i] - x;
    Dv[off + i] = d;
    v[0] = fma(d, d, v[0]);
    v[1] = fma(G[off + i], d, v[1]);
    if (!first_it) {
      const double l = lo[i], h = hi[i];
      if (d < 0.0 && lbb_has_lo(l)) {
        const double a2 = l - x;
        if (a2 >= 0.0) v[2] = 0.0;
        else if (d * v[2] < a2) v[2] = a2 / d;
      } else if (d > 0.0 && lbb_has_hi(h)) {
        const double a2 = h - x;
        if (a2 <= 0.0) v[2] = 0.0;
        else if (d * v[2] > a2) v[2] = a2 / d;
      }
    }
  }
  const int op[3] = {RED_SUM, RED_SUM, RED_MIN};
  block_reduce<3>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x < 3) part[((long long)b * nchunk + blockIdx.x) * 3 + threadIdx.x] = res[threadIdx.x];
}

}  // namespace
