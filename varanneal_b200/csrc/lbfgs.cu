// Device-resident, batched L-BFGS-B: replaces ADmin.min_lbfgs_scipy (_autodiffmin.py:72-95), i.e.
// scipy.optimize.minimize(method='L-BFGS-B', jac=True) driving A_gradA_taped, for B independent
// paths at once -- and the beta loop around it (va_ode.py:473-490, 707-789).
//
// What is mirrored from L-BFGS-B 3.0 / SciPy: limited-memory BFGS direction
// with H0 = I/theta, theta = y'y / s'y, m = maxcor pairs; More'-Thuente line search (dcsrch /
// dcstep, ftol 1e-3, gtol 0.9, xtol 0.1) with first step 1/||d|| at iteration 0 and 1 afterwards;
// the update is skipped when s'y <= eps * (-g'd * stp); at most maxls evaluations per line search,
// after which the memory is dropped and the iteration restarted (abnormal termination if it was
// already empty); stopping tests max|proj g| <= pgtol and
// (f_k - f_{k+1}) <= ftol * max(|f_k|, |f_{k+1}|, 1); maxiter / maxfun checked after each
// iteration; status = SciPy's warnflag (0 converged, 1 limit reached, 2 abnormal).
// Bounds: active-set projection (variables sitting on a bound with the gradient pushing outward
// are frozen for the iteration, the step is limited to the largest feasible one) instead of the
// generalised Cauchy point + subspace minimisation -- same minimisers, different iterates.
//
// Execution model: nothing per-iteration ever reaches the host.  Every decision (line search
// state machine, convergence tests, history management, the two-loop recursion carried out in
// coefficient space on a small Gram matrix) is taken by single-CTA kernels on per-path state in
// device memory; the n-vector work is three fused passes per iteration:
//     trial      xt = x + stp d
//     [f, g](xt) the fused action kernel (ode_stream.cuh / ode_sweep.cuh / nn_action.cu)
//     gd         g(xt).d and max |proj g|                               (2 vector reads)
//     update     s, y into the history, x <- xt, g <- gt, and every dot product the
//                Gram matrix needs, in one sweep                         (2m+4 reads, 4 writes)
//     direction  d = -H g as a linear combination of the history + d'd, g'd, max feasible step
//                                                                        (2m+1 reads, 1 write)
// The host only enqueues fixed "cycles" of these kernels and polls a counter of running paths
// every few cycles; finished paths are masked out inside the kernels.  All reductions are
// fixed-order (bit-reproducible runs).
//
// The two history passes are HBM-bound streams over 2m+4 vectors: for the unbounded L-BFGS case
// they are fed by the bulk-copy engine (cp.async.bulk into a 4-stage shared-memory ring per CTA,
// completion on mbarriers), so the bytes in flight do not depend on registers or occupancy.
//
// Asynchronous ladder (vab_anneal): every path carries its own rung index.  A path that has
// converged on rung i has its row of the result table and its minimiser saved by the cycle's last
// two kernels and starts rung i+1 (RF0 alpha**beta[i+1], handed to the action kernels through a
// per-path scale array) in the very next cycle -- paths never wait for the slowest one of a rung.
// Per-path arithmetic does not depend on what the other paths do, so the results are identical
// to running the rungs one after the other.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cfloat>
#include <cstdio>
#include <chrono>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lb_common.cuh"
#include "ode_params.h"
#include "vab_ctx.h"
#include "vab_tma.cuh"

using namespace vabmin;

namespace {

constexpr int MMAX = 10;          // largest history size (SciPy's default maxcor)
constexpr int NACC_U = 5 * MMAX + 4;

struct LbPath {
  // control flags (written by the single-CTA kernels, read by everything)
  int done, need_eval, accepted, do_update, redo_dir, first;
  int iter, nfev, col, head, pslot, ifun, iback, nskip, status;
  int ib, finished;         // rung of the ladder this path is on; ladder complete
  int xt_ready;             // the direction pass has already written the first trial point xt = x + d
  double stp_at, restore;   // in-place trials (TMA path): x currently sits at x0 + stp_at d; a failed
                            // search leaves restore = that step for lb_update_tma_kernel to undo
  double f, fold, me, fe;
  double stp, gd, gdold, dnorm, stpmx, theta, sbgnrm, dr;
  double ls_ftol, ls_gtol, ls_xtol, cd, fprev;   // line-search constants; CG: coefficient of the old direction, f two iterates back
  // dcsrch
  int brackt, stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  // Gram blocks over the history slots and the coefficients of the direction
  double SY[MMAX * MMAX];   // SY[i][j] = s_i . y_j
  double YY[MMAX * MMAX];   // y_i . y_j
  // bounded problems (lbfgsb_bounded.cuh): S'S, the Cauchy search's c = W'(x^c - x) and M c, the
  // subspace solution, bookkeeping of the projection / backtracking step
  double SS[MMAX * MMAX];
  double cvec[2 * MMAX], Mc[2 * MMAX], wv[2 * MMAX];
  double tsum, alpha_bt, dtd;
  long long ibd;
  int nfree, skip_sub, need_bt, abort_dir;
  double gS[MMAX], gY[MMAX], gg;
  double cs[MMAX], cy[MMAX], cg;
};

struct LbOpts {
  int m, maxls;
  long long maxfun, maxiter;
  double ftol, pgtol;
  int method;            // 0 = L-BFGS-B, 1 = nonlinear CG (Polak-Ribiere+)
  double ls_ftol, ls_gtol, ls_xtol;   // sufficient-decrease / curvature / bracket constants of the line search
};

}  // namespace

struct LbfgsWork {
  int B = 0;
  long long ld = 0;
  int m = 0;
  double* vec = nullptr;        // XT, GT, G, Dv, S[m], Y[m]  each (B, ld)
  size_t vec_cap = 0;
  LbPath* st = nullptr;
  int st_cap = 0;
  int* act_eval = nullptr;      // (B) mask handed to the action kernel
  double* ft = nullptr;         // (B) trial f / me / fe
  double* met = nullptr;
  double* fet = nullptr;
  double* part = nullptr;       // partial sums of the vector kernels
  size_t part_cap = 0;
  int* n_running_dev = nullptr;     // [2]
  int* n_running_host = nullptr;    // pinned [2]
  cudaEvent_t ev[2] = {nullptr, nullptr};
  double* lad = nullptr;            // scales[Nbeta] | betas[Nbeta] | rf_path[B]
  size_t lad_cap = 0;
  int* prog_dev = nullptr;          // [2][B] rungs completed by each path, per poll buffer
  int* prog_host = nullptr;         // pinned [2][B]
  int prog_cap = 0;
  cudaStream_t copy_stream = nullptr;   // device -> host sink of finished rungs
  cudaStream_t cap_stream = nullptr;    // records the cycle graph when the context runs on the legacy default stream
};

namespace {

// ---------------------------------------------------------------------------------------------
// helpers of the vector kernels

// projected gradient component (L-BFGS-B projgr)
__device__ __forceinline__ double proj_g(double x, double g, double lo, double hi) {
  if (g < 0.0) return fmax(x - hi, g);
  return fmin(x - lo, g);
}
// variable frozen for this iteration: on a bound with the gradient pushing outward
__device__ __forceinline__ bool frozen(double x, double g, double lo, double hi) {
  return (x <= lo && g > 0.0) || (x >= hi && g < 0.0);
}

// ---------------------------------------------------------------------------------------------
// state of a path at the start of a minimisation (one rung)
__device__ void lb_reset(LbPath& s, double ls_ftol, double ls_gtol, double ls_xtol) {
  s.done = 0; s.need_eval = 1; s.accepted = 0; s.do_update = 0; s.redo_dir = 0; s.first = 1;
  s.xt_ready = 0; s.stp_at = 0.0; s.restore = 0.0;
  s.tsum = 0.0; s.alpha_bt = 1.0; s.dtd = 0.0; s.ibd = -1; s.nfree = 0; s.skip_sub = 1; s.need_bt = 0; s.abort_dir = 0;
  s.iter = 0; s.nfev = 0; s.col = 0; s.head = 0; s.pslot = 0; s.ifun = 0; s.iback = 0; s.nskip = 0;
  s.status = 2;
  s.f = 0.0; s.fold = 0.0; s.me = 0.0; s.fe = 0.0;
  s.stp = 0.0; s.gd = 0.0; s.gdold = 0.0; s.dnorm = 0.0; s.stpmx = BIG; s.theta = 1.0;
  s.sbgnrm = 0.0; s.dr = 0.0;
  s.cg = 1.0; s.cd = 0.0; s.ls_ftol = ls_ftol; s.ls_gtol = ls_gtol; s.ls_xtol = ls_xtol; s.fprev = 0.0; s.gg = 0.0;
  for (int j = 0; j < MMAX; ++j) { s.cs[j] = 0.0; s.cy[j] = 0.0; s.gS[j] = 0.0; s.gY[j] = 0.0; }
}

__global__ void lb_init_kernel(LbPath* st, int* act_eval, int B, double ls_ftol, double ls_gtol, double ls_xtol,
                               const double* scales, double* rf_path) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LbPath& s = st[b];
  lb_reset(s, ls_ftol, ls_gtol, ls_xtol);
  s.ib = 0; s.finished = 0;
  rf_path[b] = scales[0];
  act_eval[b] = 1;
}

// clip the start point into the bounds (SciPy does this before the first evaluation)
__global__ void lb_clip_kernel(double* X, long long ld, long long n, const double* lo, const double* hi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* x = X + (long long)blockIdx.y * ld;
  x[i] = fmin(fmax(x[i], lo[i]), hi[i]);
}

// xt = x + stp d.  inplace (TMA path, unbounded): there is no separate trial buffer -- x itself moves
// along d, from the step it sits at (stp_at) to the new one; the accepted point then needs no copy.
__global__ void __launch_bounds__(NT) lb_trial_kernel(double* __restrict__ XT, double* __restrict__ X,
                                                      const double* __restrict__ Dv, long long ld, long long n,
                                                      const LbPath* __restrict__ st, int nchunk, int inplace,
                                                      const double* __restrict__ Zc) {
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (s.done || !s.need_eval || s.xt_ready) return;
  const double stp = s.stp;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  double* x = X + (long long)b * ld;
  const double* d = Dv + (long long)b * ld;
  if (inplace) {
    if (s.first) return;              // first evaluation of a minimisation: at x itself
    const double a = stp - s.stp_at;
    for (long long i = r.i0 + 2LL * threadIdx.x; i < r.i1; i += 2LL * NT) {
      if (i + 1 < r.i1) {
        const double2 xv = *reinterpret_cast<const double2*>(x + i);
        const double2 dv = *reinterpret_cast<const double2*>(d + i);
        *reinterpret_cast<double2*>(x + i) = make_double2(fma(a, dv.x, xv.x), fma(a, dv.y, xv.y));
      } else {
        x[i] = fma(a, d[i], x[i]);
      }
    }
    return;
  }
  double* xt = XT + (long long)b * ld;
  if (s.first) {                      // first evaluation of a minimisation: at x itself
    for (long long i = r.i0 + 2LL * threadIdx.x; i < r.i1; i += 2LL * NT) {
      if (i + 1 < r.i1) *reinterpret_cast<double2*>(xt + i) = *reinterpret_cast<const double2*>(x + i);
      else xt[i] = x[i];
    }
    return;
  }
  if (Zc != nullptr && stp == 1.0) {   // bounded: the full step is the projected point z itself (lnsrlb: x = z)
    const double* z = Zc + (long long)b * ld;
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) xt[i] = z[i];
    return;
  }
  for (long long i = r.i0 + 2LL * threadIdx.x; i < r.i1; i += 2LL * NT) {
    if (i + 1 < r.i1) {
      const double2 xv = *reinterpret_cast<const double2*>(x + i);
      const double2 dv = *reinterpret_cast<const double2*>(d + i);
      *reinterpret_cast<double2*>(xt + i) = make_double2(fma(stp, dv.x, xv.x), fma(stp, dv.y, xv.y));
    } else {
      xt[i] = fma(stp, d[i], x[i]);
    }
  }
}

// partials [gd, max |proj g|] of the trial point
template <bool BOUNDED>
__global__ void __launch_bounds__(NT) lb_gd_kernel(const double* __restrict__ XT, const double* __restrict__ GT,
                                                   const double* __restrict__ Dv, long long ld, long long n,
                                                   const double* __restrict__ lo, const double* __restrict__ hi,
                                                   const LbPath* __restrict__ st, int nchunk,
                                                   double* __restrict__ part) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[2];
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (s.done || !s.need_eval) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const double* g = GT + (long long)b * ld;
  const double* d = Dv + (long long)b * ld;
  const double* x = XT + (long long)b * ld;
  double v[2] = {0.0, 0.0};
  const bool first = s.first != 0;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double gi = g[i];
    if (!first) v[0] = fma(gi, d[i], v[0]);
    const double pg = BOUNDED ? proj_g(x[i], gi, lo[i], hi[i]) : gi;
    v[1] = fmax(v[1], fabs(pg));
  }
  const int op[2] = {RED_SUM, RED_MAX};
  block_reduce<2>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x < 2) part[((long long)b * nchunk + blockIdx.x) * 2 + threadIdx.x] = res[threadIdx.x];
}

// line-search state machine + stopping tests of one path (one thread); gd = g.d and sbg = max |proj g|
// of the trial point, fb / meb / feb its action, act = the path's entry of the evaluation mask
__device__ void lb_linesearch_body(LbPath& s, int* act, double fb, double meb, double feb, double gd, double sbg,
                                   const LbOpts& o) {
  s.nfev += 1;
  s.xt_ready = 0;                  // any further trial of this search goes through lb_trial_kernel
  const double f = fb;
  if (s.first) {
    s.first = 0;
    s.f = f; s.me = meb; s.fe = feb;
    s.sbgnrm = sbg;
    s.need_eval = 0;
    s.accepted = 1; s.do_update = 0;
    if (sbg <= o.pgtol) { s.done = 1; s.status = 0; *act = 0; }
    return;
  }
  // an evaluation that produced a NaN / Inf cannot be used by the search: treat as a failed step
  const bool finite = isfinite(f) && isfinite(gd);
  const double stp_eval = s.stp;                       // the step this evaluation was taken at
  int conv = 0;
  if (finite) conv = dcsrch_step(s, f, gd, 0.0, s.stpmx);
  else s.iback = o.maxls;      // force the restart branch below
  if (finite && conv) {
    // NEW_X
    s.f = f; s.me = meb; s.fe = feb;
    s.gd = gd;
    s.iter += 1;
    s.sbgnrm = sbg;
    s.need_eval = 0;
    s.accepted = 1;
    s.do_update = 0;
    if (s.iter >= o.maxiter) { s.done = 1; s.status = 1; }
    else if (s.nfev > o.maxfun) { s.done = 1; s.status = 1; }
    else if (sbg <= o.pgtol) { s.done = 1; s.status = 0; }
    else if (o.method == 0) {
      const double ddum = fmax(fabs(s.fold), fmax(fabs(f), 1.0));
      if ((s.fold - f) <= o.ftol * ddum) { s.done = 1; s.status = 0; }
    }
    s.fprev = s.fold;
    if (!s.done && o.method == 1) {
      s.do_update = 1;             // CG keeps no history; the update pass supplies g.y and g.g
      s.pslot = 0;
      s.dr = 1.0;
    } else if (!s.done) {
      double dr, dd;
      if (s.stp == 1.0) { dr = gd - s.gdold; dd = -s.gdold; }
      else { dr = (gd - s.gdold) * s.stp; dd = -s.gdold * s.stp; }
      s.dr = dr;
      if (dr <= EPSMCH * dd) { s.nskip += 1; s.do_update = 0; }
      else {
        s.do_update = 1;
        s.pslot = (s.col < o.m) ? (s.head + s.col) % o.m : s.head;
      }
    }
    if (s.done) *act = 0;
    return;
  }
  s.stp_at = stp_eval;                                 // in-place trials: where x sits now
  if (finite) {
    s.ifun += 1;
    s.iback = s.ifun - 1;
  }
  if (s.iback >= o.maxls) {
    // line search failed: x, g, f still hold the start of the search (in-place trials: x is put
    // back at the top of lb_update_tma_kernel)
    s.need_eval = 0;
    s.restore = stp_eval;
    if (s.col == 0) {
      s.done = 1; s.status = 2; *act = 0;      // ABNORMAL_TERMINATION_IN_LNSRCH
    } else {
      s.col = 0; s.head = 0; s.theta = 1.0;
      s.redo_dir = 1;                                  // RESTART_FROM_LNSRCH
    }
  }
}

// one warp per path
__global__ void lb_linesearch_kernel(LbPath* st, int* act_eval, const double* ft, const double* met,
                                     const double* fet, const double* part, int nchunk, LbOpts o,
                                     int bounded) {
  const int b = blockIdx.x;
  LbPath& s = st[b];
  if (s.done || !s.need_eval) return;
  if (threadIdx.x != 0) return;
  double gd = 0.0, sbg = 0.0;
  for (int c = 0; c < nchunk; ++c) {
    gd += part[((long long)b * nchunk + c) * 2 + 0];
    sbg = fmax(sbg, part[((long long)b * nchunk + c) * 2 + 1]);
  }
  lb_linesearch_body(s, act_eval + b, ft[b], met[b], fet[b], gd, sbg, o);
  (void)bounded;
}

// sum over the 16 lanes of a half warp (MMAX <= 16 columns, one per lane), fixed butterfly order
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
  for (int sft = 8; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
  return v;
}

// lb_gram_body for L-BFGS executed by one warp: lane j owns column j of the coefficient vectors, the
// inner products of the two-loop recursion are half-warp reductions instead of a serial chain of
// dependent multiply-adds (same algebra; the sums associate differently).  s, SYs, YYs, sum in
// shared memory.  All 32 lanes must call.
__device__ void lb_gram_warp(LbPath& s, const double* sum, double* SYs, double* YYs, int m) {
  const int j = threadIdx.x & 31;
  const bool on = j < m;
  int col = s.col, head = s.head;
  double theta = s.theta;
  double gSj = on ? s.gS[j] : 0.0, gYj = on ? s.gY[j] : 0.0;
  const bool accepted = s.accepted != 0, upd = s.do_update != 0;
  const int p = s.pslot;
  const double dr = s.dr;
  __syncwarp();
  if (accepted) {
    if (upd) {
      if (on && j != p) {
        SYs[p * MMAX + j] = sum[2 * MMAX + j];      // s_p . y_j
        SYs[j * MMAX + p] = sum[3 * MMAX + j];      // s_j . y_p
        YYs[p * MMAX + j] = sum[4 * MMAX + j];
        YYs[j * MMAX + p] = sum[4 * MMAX + j];
      }
      if (j == p) {
        SYs[p * MMAX + p] = dr;                     // s'y from the line search, as L-BFGS-B does
        YYs[p * MMAX + p] = sum[5 * MMAX];
      }
      theta = sum[5 * MMAX] / dr;
      if (col < m) col += 1;
      else head = (head + 1) % m;
      if (on) { gSj = (j == p) ? sum[5 * MMAX + 1] : sum[j]; gYj = (j == p) ? sum[5 * MMAX + 2] : sum[MMAX + j]; }
      __syncwarp();
      if (on) {                                     // write back the row / column that changed
        s.SY[p * MMAX + j] = SYs[p * MMAX + j]; s.SY[j * MMAX + p] = SYs[j * MMAX + p];
        s.YY[p * MMAX + j] = YYs[p * MMAX + j]; s.YY[j * MMAX + p] = YYs[j * MMAX + p];
      }
      if (j == 0) { s.theta = theta; s.col = col; s.head = head; }
    } else if (on) {
      gSj = sum[j]; gYj = sum[MMAX + j];
    }
    if (on) { s.gS[j] = gSj; s.gY[j] = gYj; }
    if (j == 0) s.gg = sum[5 * MMAX + 3];
  }
  __syncwarp();
  // r = H ghat  as  cg * ghat + sum_j cs_j s_j + cy_j y_j   (d = -r)
  double cg = 1.0, csj = 0.0, cyj = 0.0, alphaj = 0.0;
  for (int k = col - 1; k >= 0; --k) {               // newest -> oldest
    const int i = (head + k) % m;
    const double gi = __shfl_sync(0xffffffffu, gSj, i);
    const double sq = gi * cg + half_sum(on ? cyj * SYs[i * MMAX + j] : 0.0);
    const double a = sq / SYs[i * MMAX + i];
    if (j == i) { alphaj = a; cyj -= a; }
  }
  const double gamma = (col > 0) ? 1.0 / theta : 1.0;
  cg *= gamma;
  cyj *= gamma;
  for (int k = 0; k < col; ++k) {                    // oldest -> newest
    const int i = (head + k) % m;
    const double gi = __shfl_sync(0xffffffffu, gYj, i);
    const double yr = gi * cg + half_sum(on ? cyj * YYs[i * MMAX + j] + csj * SYs[j * MMAX + i] : 0.0);
    const double beta = yr / SYs[i * MMAX + i];
    if (j == i) csj += alphaj - beta;
  }
  if (j == 0) s.cg = cg;
  if (j < MMAX) { s.cs[j] = csj; s.cy[j] = cyj; }
}

// Reduction of NV <= 64 per-thread sums over the CTA without NV rounds through shared memory: a
// halving butterfly inside each warp (at every step a lane hands half of its values to its partner
// and adds the half it receives: 62 exchanges instead of 5 NV; lane l ends up with the warp sums of
// entries 2l and 2l + 1), then the warps in order.  wred: (NT / 32) * 64 doubles.  Fixed order.
template <int NV>
__device__ void cta_reduce_sum(const double* acc, double* out, double* wred) {
  static_assert(NV <= 64, "halving butterfly over 64 slots");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) v[k] = (k < NV) ? acc[k] : 0.0;
#pragma unroll
  for (int half = 32, mask = 16; half >= 2; half >>= 1, mask >>= 1) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const double send = up ? v[k] : v[k + half];
      const double keep = up ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
  // one more step (mask 1 handled above when half == 2 -> values 0, 1 remain)
  wred[warp * 64 + 2 * lane] = v[0];
  wred[warp * 64 + 2 * lane + 1] = v[1];
  __syncthreads();
  for (int k = threadIdx.x; k < NV; k += NT) {
    double a = wred[k];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) a += wred[w * 64 + k];
    out[k] = a;
  }
}


// accepted step: x <- xt, g <- gt; if the pair is kept, s = stp d and y = gt - g go to slot p;
// partial dot products (NACC_U per chunk):
//   [0..M)    ghat . S_j      [M..2M)  ghat . Y_j     [2M..3M)  s . Y_j
//   [3M..4M)  y . S_j         [4M..5M) y . Y_j        5M: y.y   5M+1: ghat.s   5M+2: ghat.y  5M+3: ghat.ghat
// (ghat = new gradient with the frozen components zeroed; entries of slot p refer to its old
// contents and are ignored by lb_gram_kernel).
template <bool BOUNDED>
__global__ void __launch_bounds__(NT, 1) lb_update_kernel(
    double* __restrict__ X, double* __restrict__ G, const double* __restrict__ XT,
    const double* __restrict__ GT, const double* __restrict__ Dv, double* __restrict__ S,
    double* __restrict__ Y, long long ld, long long n, long long hstride, const double* __restrict__ lo,
    const double* __restrict__ hi, const LbPath* __restrict__ st, int m, int nchunk,
    double* __restrict__ part, int b0) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[NACC_U];
  const int b = b0 + blockIdx.y;
  const LbPath& s = st[b];
  if (!s.accepted) return;
  const bool upd = s.do_update != 0;
  const int p = s.pslot;
  const double stp = s.stp;
  const int col = s.col;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double acc[NACC_U];
#pragma unroll
  for (int k = 0; k < NACC_U; ++k) acc[k] = 0.0;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double xt = XT[off + i], gt = GT[off + i];
    const double gold = G[off + i];
    double sv = 0.0, yv = 0.0;
    if (upd) {
      sv = stp * Dv[off + i];
      yv = gt - gold;
    }
    double gh = gt;
    if (BOUNDED) {
      if (frozen(xt, gt, lo[i], hi[i])) gh = 0.0;
    }
#pragma unroll
    for (int j = 0; j < MMAX; ++j) {
      if (j < m && (j < col || col == m) && !(upd && j == p)) {   // slot p is about to be replaced
        const double sj = S[(long long)j * hstride + off + i];
        const double yj = Y[(long long)j * hstride + off + i];
        acc[j] = fma(gh, sj, acc[j]);
        acc[MMAX + j] = fma(gh, yj, acc[MMAX + j]);
        acc[2 * MMAX + j] = fma(sv, yj, acc[2 * MMAX + j]);
        acc[3 * MMAX + j] = fma(yv, sj, acc[3 * MMAX + j]);
        acc[4 * MMAX + j] = fma(yv, yj, acc[4 * MMAX + j]);
      }
    }
    acc[5 * MMAX] = fma(yv, yv, acc[5 * MMAX]);
    acc[5 * MMAX + 1] = fma(gh, sv, acc[5 * MMAX + 1]);
    acc[5 * MMAX + 2] = fma(gh, yv, acc[5 * MMAX + 2]);
    acc[5 * MMAX + 3] = fma(gh, gh, acc[5 * MMAX + 3]);
    X[off + i] = xt;
    G[off + i] = gt;
    if (upd) {
      S[(long long)p * hstride + off + i] = sv;
      Y[(long long)p * hstride + off + i] = yv;
    }
  }
  cta_reduce_sum<NACC_U>(acc, res, scratch);
  __syncthreads();
  for (int k = threadIdx.x; k < NACC_U; k += NT)
    part[((long long)b * nchunk + blockIdx.x) * NACC_U + k] = res[k];
}

// history bookkeeping + the two-loop recursion in coefficient space of one path (one thread).
// sum: the NACC_U dot products of the update pass (read when s.accepted); SYs / YYs: the path's
// Gram blocks staged in shared memory (updated in place and written back to s)
__device__ void lb_gram_body(LbPath& s, const double* sum, double* SYs, double* YYs, int m, int method) {
  if (method == 1) {
    // Polak-Ribiere+:  d = -g + max(0, g_new.(g_new - g_old) / g_old.g_old) d_old
    double beta = 0.0;
    if (s.accepted && s.do_update && s.gg > 0.0) beta = fmax(0.0, sum[5 * MMAX + 2] / s.gg);
    if (s.accepted) s.gg = sum[5 * MMAX + 3];
    s.cd = beta;
    s.cg = 1.0;
    for (int j = 0; j < MMAX; ++j) { s.cs[j] = 0.0; s.cy[j] = 0.0; }
    return;
  }
  double gS[MMAX], gY[MMAX];
  for (int j = 0; j < MMAX; ++j) { gS[j] = s.gS[j]; gY[j] = s.gY[j]; }
  int col = s.col, head = s.head;
  double theta = s.theta;
  if (s.accepted) {
    if (s.do_update) {
      const int p = s.pslot;
      for (int j = 0; j < m; ++j) {
        if (j == p) continue;
        SYs[p * MMAX + j] = sum[2 * MMAX + j];      // s_p . y_j
        SYs[j * MMAX + p] = sum[3 * MMAX + j];      // s_j . y_p
        YYs[p * MMAX + j] = sum[4 * MMAX + j];
        YYs[j * MMAX + p] = sum[4 * MMAX + j];
      }
      SYs[p * MMAX + p] = s.dr;                     // s'y from the line search, as L-BFGS-B does
      YYs[p * MMAX + p] = sum[5 * MMAX];
      theta = sum[5 * MMAX] / s.dr;
      if (col < m) col += 1;
      else head = (head + 1) % m;
      for (int j = 0; j < m; ++j) { gS[j] = sum[j]; gY[j] = sum[MMAX + j]; }
      gS[p] = sum[5 * MMAX + 1];
      gY[p] = sum[5 * MMAX + 2];
      for (int j = 0; j < m; ++j) {                 // write back the row / column that changed
        s.SY[p * MMAX + j] = SYs[p * MMAX + j]; s.SY[j * MMAX + p] = SYs[j * MMAX + p];
        s.YY[p * MMAX + j] = YYs[p * MMAX + j]; s.YY[j * MMAX + p] = YYs[j * MMAX + p];
      }
      s.theta = theta; s.col = col; s.head = head;
    } else {
      for (int j = 0; j < m; ++j) { gS[j] = sum[j]; gY[j] = sum[MMAX + j]; }
    }
    for (int j = 0; j < m; ++j) { s.gS[j] = gS[j]; s.gY[j] = gY[j]; }
    s.gg = sum[5 * MMAX + 3];
  }
  // r = H ghat  as  cg * ghat + sum_j cs_j s_j + cy_j y_j   (d = -r)
  double cg = 1.0, cs[MMAX], cy[MMAX], alpha[MMAX];
  for (int j = 0; j < MMAX; ++j) { cs[j] = 0.0; cy[j] = 0.0; alpha[j] = 0.0; }
  for (int k = col - 1; k >= 0; --k) {               // newest -> oldest
    const int i = (head + k) % m;
    double sq = gS[i] * cg;
    for (int j = 0; j < m; ++j) sq += cy[j] * SYs[i * MMAX + j];
    alpha[i] = sq / SYs[i * MMAX + i];
    cy[i] -= alpha[i];
  }
  const double gamma = (col > 0) ? 1.0 / theta : 1.0;
  cg *= gamma;
  for (int j = 0; j < m; ++j) cy[j] *= gamma;
  for (int k = 0; k < col; ++k) {                    // oldest -> newest
    const int i = (head + k) % m;
    double yr = gY[i] * cg;
    for (int j = 0; j < m; ++j) yr += cy[j] * YYs[i * MMAX + j] + cs[j] * SYs[j * MMAX + i];
    const double beta = yr / SYs[i * MMAX + i];
    cs[i] += alpha[i] - beta;
  }
  s.cg = cg;
  for (int j = 0; j < MMAX; ++j) { s.cs[j] = cs[j]; s.cy[j] = cy[j]; }
}

// history bookkeeping + the two-loop recursion in coefficient space; one thread per path
__global__ void lb_gram_kernel(LbPath* st, const double* part, int nchunk, int m, int b0, int method) {
  const int b = b0 + blockIdx.x;
  LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) return;
  __shared__ double sum[NACC_U];
  if (s.accepted) {
    for (int k = threadIdx.x; k < NACC_U; k += blockDim.x) {
      double a = 0.0;
      for (int c = 0; c < nchunk; ++c) a += part[((long long)b * nchunk + c) * NACC_U + k];
      sum[k] = a;
    }
  }
  // the Gram blocks are walked by one thread: stage them in shared memory (global latency would
  // dominate the ~m^2 dependent loads of the two-loop recursion)
  __shared__ double SYs[MMAX * MMAX], YYs[MMAX * MMAX];
  for (int k = threadIdx.x; k < MMAX * MMAX; k += blockDim.x) { SYs[k] = s.SY[k]; YYs[k] = s.YY[k]; }
  __syncthreads();
  if (method == 1) { if (threadIdx.x == 0) lb_gram_body(s, sum, SYs, YYs, m, method); }
  else if (threadIdx.x < 32) lb_gram_warp(s, sum, SYs, YYs, m);
}

// d = -(cg ghat + sum_j cs_j S_j + cy_j Y_j), frozen components zero; partials [d.d, g.d, stpmx]
template <bool BOUNDED>
__global__ void __launch_bounds__(NT) lb_direction_kernel(
    const double* __restrict__ X, const double* __restrict__ G, double* __restrict__ Dv,
    const double* __restrict__ S, const double* __restrict__ Y, long long ld, long long n,
    long long hstride, const double* __restrict__ lo, const double* __restrict__ hi,
    const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part, int b0) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[3];
  const int b = b0 + blockIdx.y;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  const int col = s.col;
  double cs[MMAX], cy[MMAX];
#pragma unroll
  for (int j = 0; j < MMAX; ++j) { cs[j] = s.cs[j]; cy[j] = s.cy[j]; }
  const double cg = s.cg, cd = s.cd;
  double v[3] = {0.0, 0.0, BIG};
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double g = G[off + i];
    bool fr = false;
    double x = 0.0, l = 0.0, h = 0.0;
    if (BOUNDED) {
      x = X[off + i]; l = lo[i]; h = hi[i];
      fr = frozen(x, g, l, h);
    }
    double rr = fr ? 0.0 : cg * g;
#pragma unroll
    for (int j = 0; j < MMAX; ++j) {
      if (j < m && (j < col || col == m)) {
        rr = fma(cs[j], S[(long long)j * hstride + off + i], rr);
        rr = fma(cy[j], Y[(long long)j * hstride + off + i], rr);
      }
    }
    double d = fr ? 0.0 : -rr;
    if (cd != 0.0 && !fr) d = fma(cd, Dv[off + i], d);
    if (BOUNDED) {
      // keep x + stp d feasible: largest step before a bound is hit (lnsrlb)
      if (d < 0.0 && l > -DBL_MAX) {
        const double a2 = l - x;
        if (a2 >= 0.0) d = 0.0;
        else if (d * v[2] < a2) v[2] = a2 / d;
      } else if (d > 0.0 && h < DBL_MAX) {
        const double a2 = h - x;
        if (a2 <= 0.0) d = 0.0;
        else if (d * v[2] > a2) v[2] = a2 / d;
      }
    }
    Dv[off + i] = d;
    v[0] = fma(d, d, v[0]);
    v[1] = fma(g, d, v[1]);
  }
  const int op[3] = {RED_SUM, RED_SUM, RED_MIN};
  block_reduce<3>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x < 3) part[((long long)b * nchunk + blockIdx.x) * 3 + threadIdx.x] = res[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variants of the two history passes (unbounded L-BFGS; same arithmetic per element and
// the same per-chunk partial layout as lb_update_kernel / lb_direction_kernel above).
//
// A CTA owns one chunk of one path and walks it in tiles of HTILE elements.  A stage of its
// shared-memory ring holds the tile of every input vector (up to 2m+4 of them, 2 KB each).
// Warp specialisation: a producer warp -- lane k owns vector k -- enqueues one cp.async.bulk per
// vector, all completing on the stage's "full" mbarrier, up to HNS tiles ahead; HW consumer warps
// do the arithmetic from shared memory and hand the stage back through its "empty" mbarrier.
// No __syncthreads in the loop; the bytes in flight (up to HNS x 48 KB per SM) do not depend on
// the register count of the 5m+4 running dot products.
constexpr int HW = 4;                   // consumer warps
constexpr int HT = HW * 32;             // consumer threads
constexpr int HTILE = 2 * HT;           // elements per tile (two per consumer thread)
constexpr int U_NSTR = 3 + 2 * MMAX;    // GT, G, D, S_0.., Y_0..  (x moves in place: no trial buffer to copy)
constexpr int D_NSTR = 2 + 2 * MMAX;    // G, X, S_0.., Y_0..
static_assert(U_NSTR <= 32 && D_NSTR <= 32, "one producer lane per vector");
// ring stages HNS (template parameter): 4 with one CTA per SM, or 2 with two CTAs per SM
constexpr size_t hist_smem(int nstr, int hns) { return (size_t)hns * nstr * HTILE * sizeof(double) + 2 * hns * 8; }

// producer warp: stream the chunk [0, len) of every vector with a non-null source (this lane's:
// src) through the ring.  The last tile may be short; a bulk copy moves whole 16-byte units, the
// odd element read past the end lies inside the row padding and is masked by the consumers.
template <int NSTR, int HNS>
__device__ __forceinline__ void hist_produce(uint32_t ring, uint32_t full, uint32_t empty,
                                             const double* src, int nact, int ntile, long long len) {
  const int lane = threadIdx.x & 31;
  for (int t = 0; t < ntile; ++t) {
    const int stage = t % HNS;
    if (t >= HNS) vabs::mbar_wait(empty + 8u * (uint32_t)stage, (uint32_t)(((t / HNS) - 1) & 1));
    const long long e0 = (long long)t * HTILE;
    long long cnt = len - e0;
    if (cnt > HTILE) cnt = HTILE;
    cnt = (cnt + 1) & ~1LL;
    const uint32_t bytes = (uint32_t)cnt * 8u;
    const uint32_t bar = full + 8u * (uint32_t)stage;
    if (lane == 0) vabs::mbar_expect_tx(bar, bytes * (uint32_t)nact);
    __syncwarp();
    if (src != nullptr)
      vabs::tma_load(ring + (uint32_t)((stage * NSTR + lane) * (HTILE * 8)), src + e0, bytes, bar);
  }
}

template <int HNS>
__device__ __forceinline__ void hist_ring_init(uint32_t full, uint32_t empty) {
#pragma unroll
  for (int q = 0; q < HNS; ++q) {
    vabs::mbar_init(full + 8u * q, 1);
    vabs::mbar_init(empty + 8u * q, HW);
  }
  vabs::mbar_fence_init();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void hist_consumer_sync() {          // the HT consumer threads only
  asm volatile("bar.sync 1, %0;" ::"n"(HT) : "memory");
}

// fixed-order reduction of NV per-thread sums over the consumer threads: lanes by shuffle, then
// the warps in order
template <int NV>
__device__ __forceinline__ void hist_reduce_store(const double* acc, double (*wred)[NV], double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double v = acc[k];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_down_sync(0xffffffffu, v, sft);
    if (lane == 0) wred[warp][k] = v;
  }
  hist_consumer_sync();
  for (int k = threadIdx.x; k < NV; k += HT) {
    double a = wred[0][k];
#pragma unroll
    for (int w = 1; w < HW; ++w) a += wred[w][k];
    out[k] = a;
  }
}

template <int HNS>
__global__ void __launch_bounds__(HT + 32, 1) lb_update_tma_kernel(
    double* __restrict__ X, double* __restrict__ G,
    const double* __restrict__ GT, const double* __restrict__ Dv, double* __restrict__ S,
    double* __restrict__ Y, long long ld, long long n, long long hstride,
    const LbPath* __restrict__ st, int m, int nchunk, double* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char hist_sm[];
  __shared__ double wred[HW][NACC_U];
  const int b = blockIdx.y, tid = threadIdx.x;
  const LbPath& s = st[b];
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  if (s.restore != 0.0) {
    // in-place trials: a failed line search (never an accepted step) leaves x at x0 + restore d;
    // put it back before the direction pass forms the next trial point
    const double a = s.restore;
    double* x = X + (long long)b * ld;
    const double* d = Dv + (long long)b * ld;
    for (long long i = r.i0 + tid; i < r.i1; i += HT + 32) x[i] = fma(-a, d[i], x[i]);
    return;
  }
  if (!s.accepted) return;
  const bool upd = s.do_update != 0;
  const int p = s.pslot;
  const double stp = s.stp;
  const int col = s.col;
  const long long len = r.i1 - r.i0;
  double* pout = part + ((long long)b * nchunk + blockIdx.x) * NACC_U;
  if (len <= 0) {
    for (int k = tid; k < NACC_U; k += HT + 32) pout[k] = 0.0;
    return;
  }
  unsigned hmask = 0;                       // history slots whose old contents are read
  for (int j = 0; j < MMAX; ++j)
    if (j < m && (j < col || col == m) && !(upd && j == p)) hmask |= 1u << j;
  const long long base = (long long)b * ld + r.i0;
  double* tiles = reinterpret_cast<double*>(hist_sm);
  const uint32_t ring = vabs::s32(tiles);
  const uint32_t full = ring + (uint32_t)(HNS * U_NSTR * HTILE * 8);
  const uint32_t empty = full + HNS * 8u;
  const int ntile = (int)((len + HTILE - 1) / HTILE);
  if (tid == 0) hist_ring_init<HNS>(full, empty);
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == HW) {                         // ---- producer warp
    const int k = tid & 31;
    const double* src = nullptr;
    if (k == 0) src = GT + base;
    else if (k == 1) src = G + base;
    else if (k == 2) src = upd ? Dv + base : nullptr;
    else if (k < 3 + MMAX) src = ((hmask >> (k - 3)) & 1u) ? S + (long long)(k - 3) * hstride + base : nullptr;
    else if (k < U_NSTR) src = ((hmask >> (k - 3 - MMAX)) & 1u) ? Y + (long long)(k - 3 - MMAX) * hstride + base : nullptr;
    const int nact = 2 + (upd ? 1 : 0) + 2 * __popc(hmask);
    hist_produce<U_NSTR, HNS>(ring, full, empty, src, nact, ntile, len);
    return;
  }
  // ---- consumer warps
  double acc[NACC_U];
#pragma unroll
  for (int k = 0; k < NACC_U; ++k) acc[k] = 0.0;
  double* go = G + base;
  double* so = S + (long long)p * hstride + base;
  double* yo = Y + (long long)p * hstride + base;
  for (int t = 0; t < ntile; ++t) {
    const int stage = t % HNS;
    vabs::mbar_wait(full + 8u * (uint32_t)stage, (uint32_t)((t / HNS) & 1));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double* tl = tiles + (size_t)stage * (U_NSTR * HTILE) + h * HT + tid;
      const long long e = (long long)t * HTILE + h * HT + tid;
      const bool valid = e < len;
      const double gt = valid ? tl[0] : 0.0;
      const double gold = valid ? tl[HTILE] : 0.0;
      double sv = 0.0, yv = 0.0;
      if (upd) {
        const double dv = valid ? tl[2 * HTILE] : 0.0;
        sv = stp * dv;
        yv = gt - gold;
      }
      const double gh = gt;
#pragma unroll
      for (int j = 0; j < MMAX; ++j) {
        if ((hmask >> j) & 1u) {
          const double sj = valid ? tl[(3 + j) * HTILE] : 0.0;
          const double yj = valid ? tl[(3 + MMAX + j) * HTILE] : 0.0;
          acc[j] = fma(gh, sj, acc[j]);
          acc[MMAX + j] = fma(gh, yj, acc[MMAX + j]);
          acc[2 * MMAX + j] = fma(sv, yj, acc[2 * MMAX + j]);
          acc[3 * MMAX + j] = fma(yv, sj, acc[3 * MMAX + j]);
          acc[4 * MMAX + j] = fma(yv, yj, acc[4 * MMAX + j]);
        }
      }
      acc[5 * MMAX] = fma(yv, yv, acc[5 * MMAX]);
      acc[5 * MMAX + 1] = fma(gh, sv, acc[5 * MMAX + 1]);
      acc[5 * MMAX + 2] = fma(gh, yv, acc[5 * MMAX + 2]);
      acc[5 * MMAX + 3] = fma(gh, gh, acc[5 * MMAX + 3]);
      if (valid) {
        go[e] = gt;
        if (upd) {
          so[e] = sv;
          yo[e] = yv;
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) vabs::mbar_arrive(empty + 8u * (uint32_t)stage);   // stage may be refilled
  }
  hist_reduce_store<NACC_U>(acc, wred, pout);
}

template <int HNS>
__global__ void __launch_bounds__(HT + 32, 1) lb_direction_tma_kernel(
    const double* __restrict__ G, double* __restrict__ Dv, const double* __restrict__ S,
    const double* __restrict__ Y, double* __restrict__ X, long long ld,
    long long n, long long hstride, const LbPath* __restrict__ st, int m, int nchunk,
    double* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char hist_sm[];
  __shared__ double wred[HW][2];
  const int b = blockIdx.y, tid = threadIdx.x;
  const LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long len = r.i1 - r.i0;
  double* pout = part + ((long long)b * nchunk + blockIdx.x) * 3;
  if (len <= 0) {
    if (tid == 0) { pout[0] = 0.0; pout[1] = 0.0; pout[2] = BIG; }
    return;
  }
  const int col = s.col;
  unsigned hmask = 0;
  for (int j = 0; j < MMAX; ++j)
    if (j < m && (j < col || col == m)) hmask |= 1u << j;
  // After the first iteration the search starts at stp = 1 (lb_start_kernel): x is moved to the first
  // trial point x + d here, in the same pass (trials are in place), and lb_trial_kernel is skipped.
  const bool wxt = s.iter > 0;
  const long long base = (long long)b * ld + r.i0;
  double* tiles = reinterpret_cast<double*>(hist_sm);
  const uint32_t ring = vabs::s32(tiles);
  const uint32_t full = ring + (uint32_t)(HNS * D_NSTR * HTILE * 8);
  const uint32_t empty = full + HNS * 8u;
  const int ntile = (int)((len + HTILE - 1) / HTILE);
  if (tid == 0) hist_ring_init<HNS>(full, empty);
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == HW) {                         // ---- producer warp
    const int k = tid & 31;
    const double* src = nullptr;
    if (k == 0) src = G + base;
    else if (k == 1) src = wxt ? X + base : nullptr;
    else if (k < 2 + MMAX) src = ((hmask >> (k - 2)) & 1u) ? S + (long long)(k - 2) * hstride + base : nullptr;
    else if (k < D_NSTR) src = ((hmask >> (k - 2 - MMAX)) & 1u) ? Y + (long long)(k - 2 - MMAX) * hstride + base : nullptr;
    const int nact = 1 + (wxt ? 1 : 0) + 2 * __popc(hmask);
    hist_produce<D_NSTR, HNS>(ring, full, empty, src, nact, ntile, len);
    return;
  }
  double cs[MMAX], cy[MMAX];
#pragma unroll
  for (int j = 0; j < MMAX; ++j) { cs[j] = s.cs[j]; cy[j] = s.cy[j]; }
  const double cg = s.cg;
  double v[2] = {0.0, 0.0};
  double* dout = Dv + base;
  double* xtout = X + base;
  for (int t = 0; t < ntile; ++t) {
    const int stage = t % HNS;
    vabs::mbar_wait(full + 8u * (uint32_t)stage, (uint32_t)((t / HNS) & 1));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double* tl = tiles + (size_t)stage * (D_NSTR * HTILE) + h * HT + tid;
      const long long e = (long long)t * HTILE + h * HT + tid;
      const bool valid = e < len;
      const double g = valid ? tl[0] : 0.0;
      double rr = cg * g;
#pragma unroll
      for (int j = 0; j < MMAX; ++j) {
        if ((hmask >> j) & 1u) {
          const double sj = valid ? tl[(2 + j) * HTILE] : 0.0;
          const double yj = valid ? tl[(2 + MMAX + j) * HTILE] : 0.0;
          rr = fma(cs[j], sj, rr);
          rr = fma(cy[j], yj, rr);
        }
      }
      const double d = -rr;
      if (valid) {
        dout[e] = d;
        if (wxt) xtout[e] = fma(1.0, d, tl[HTILE]);        // = lb_trial_kernel's fma(stp, d, x) at stp = 1
      }
      v[0] = fma(d, d, v[0]);
      v[1] = fma(g, d, v[1]);
    }
    __syncwarp();
    if ((tid & 31) == 0) vabs::mbar_arrive(empty + 8u * (uint32_t)stage);
  }
  hist_reduce_store<2>(v, wred, pout);
  if (tid == 0) pout[2] = BIG;
}

}  // namespace
#include "lbfgsb_bounded.cuh"
namespace {

// start of a line search (lnsrlb, task = START)
// (dd = d.d, gd = g.d, stpmx = largest feasible step of the direction just formed; one thread)
__device__ void lb_start_body(LbPath& s, int* act, double dd, double gd, double stpmx, const LbOpts& o,
                              int bounded, int xt_fused, int boxed) {
  s.restore = 0.0;
  s.accepted = 0;
  s.redo_dir = 0;
  s.do_update = 0;
  s.dnorm = sqrt(dd);
  s.dtd = dd;
  s.gdold = gd;
  if (!(gd < 0.0) || !(dd > 0.0)) {
    // not a descent direction (info = -4): drop the memory and restart, or give up
    if (s.col == 0) { s.done = 1; s.status = (dd == 0.0) ? 0 : 2; *act = 0; }
    else {
      s.col = 0; s.head = 0; s.theta = 1.0; s.redo_dir = 1;
      // in-place trials: the direction pass has already moved x by this (rejected) direction;
      // lb_update_tma_kernel takes it back in the next cycle, before the new direction is formed
      if (xt_fused && s.iter > 0) s.restore = 1.0;
    }
    return;
  }
  if (bounded && s.iter == 0) stpmx = fmin(stpmx, 1.0);
  if (!(stpmx > 0.0)) { s.done = 1; s.status = 2; *act = 0; return; }
  s.stpmx = stpmx;
  if (o.method == 1) {
    // SciPy's CG: alpha_1 = min(1, 1.01 * 2 (f_k - f_{k-1}) / g.d), with f_{-1} = f_0 + |g|/2
    double a1 = (s.iter == 0) ? 1.01 / s.dnorm : 1.01 * 2.0 * (s.f - s.fprev) / gd;
    if (!(a1 > 0.0)) a1 = 1.0;
    s.stp = fmin(fmin(1.0, a1), stpmx);
  } else if (s.iter == 0 && !(bounded && boxed)) s.stp = fmin(1.0 / s.dnorm, stpmx);   // lnsrlb: iter == 0 and not boxed
  else s.stp = fmin(1.0, stpmx);
  s.fold = s.f;
  s.ifun = 1;
  s.iback = 0;
  dcsrch_start(s, s.f, gd, stpmx);
  s.need_eval = 1;
  s.xt_ready = (xt_fused && s.iter > 0 && s.stp == 1.0) ? 1 : 0;   // lb_direction_tma_kernel wrote xt = x + d
  s.stp_at = s.xt_ready ? 1.0 : 0.0;
  *act = 1;
}

__global__ void lb_start_kernel(LbPath* st, int* act_eval, const double* part, int nchunk, LbOpts o,
                                int bounded, int b0, int xt_fused, const int* boxed_flag) {
  const int b = b0 + blockIdx.x;
  LbPath& s = st[b];
  if (!(s.accepted || s.redo_dir) || s.done) { if (s.done) { s.accepted = 0; s.redo_dir = 0; s.restore = 0.0; } return; }
  if (s.abort_dir) { s.abort_dir = 0; return; }       // the direction is formed again next cycle (memory dropped)
  const int boxed = boxed_flag ? *boxed_flag : 0;
  double dd = 0.0, gd = 0.0, stpmx = BIG;
  for (int c = 0; c < nchunk; ++c) {
    dd += part[((long long)b * nchunk + c) * 3 + 0];
    gd += part[((long long)b * nchunk + c) * 3 + 1];
    stpmx = fmin(stpmx, part[((long long)b * nchunk + c) * 3 + 2]);
  }
  lb_start_body(s, act_eval + b, dd, gd, stpmx, o, bounded, xt_fused, boxed);
}

// ---------------------------------------------------------------------------------------------
// end of a rung.  A path whose minimisation has just stopped (done, ladder not finished) has its
// minimiser copied to minpaths[b][ib] by lb_save_kernel; lb_advance_kernel then records the row
// of the result table and either restarts the path on the next rung or retires it.
struct LbLadder {
  int Nbeta;
  const double* scales;      // (Nbeta) alpha**beta
  const double* betas;       // (Nbeta)
  double* rf_path;           // (B) scale of the rung each path is on (read by the action kernels)
  double* table;             // (B, Nbeta, 5) or nullptr
  double* minpaths;          // (B, Nbeta, mp_pitch) or nullptr
  long long win0, winw;      // columns [win0, win0 + winw) of each minimiser are stored (vab_set_path_window)
  long long mp_pitch;        // doubles between stored rows (ld for the full path)
  int *status2, *nit2, *nfev2;   // (B, Nbeta) or nullptr
};

__global__ void __launch_bounds__(NT) lb_save_kernel(const double* __restrict__ X, long long ld, long long n,
                                                     const LbPath* __restrict__ st, int nchunk, LbLadder L) {
  const int b = blockIdx.y;
  const LbPath& s = st[b];
  if (!s.done || s.finished) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const double* x = X + (long long)b * ld;
  double* o = L.minpaths + ((long long)b * L.Nbeta + s.ib) * L.mp_pitch;
  if (L.winw != n) {                  // a column window (e.g. the estimated parameters only)
    const long long lo = r.i0 > L.win0 ? r.i0 : L.win0;
    const long long hi = r.i1 < L.win0 + L.winw ? r.i1 : L.win0 + L.winw;
    for (long long i = lo + threadIdx.x; i < hi; i += NT) o[i - L.win0] = x[i];
    return;
  }
  for (long long i = r.i0 + 2LL * threadIdx.x; i < r.i1; i += 2LL * NT) {
    if (i + 1 < r.i1) *reinterpret_cast<double2*>(o + i) = *reinterpret_cast<const double2*>(x + i);
    else o[i] = x[i];
  }
}

__device__ void lb_advance_body(LbPath& s, int* act_eval, int b, const LbLadder& L, const LbOpts& o) {
  if (!s.done || s.finished) return;
  const int ib = s.ib;
  if (L.table) {
    double* row = L.table + ((long long)b * L.Nbeta + ib) * 5;
    row[0] = L.betas[ib]; row[1] = s.f; row[2] = s.me; row[3] = s.fe; row[4] = s.fe / L.scales[ib];
  }
  if (L.status2) L.status2[(long long)b * L.Nbeta + ib] = s.status;
  if (L.nit2) L.nit2[(long long)b * L.Nbeta + ib] = s.iter;
  if (L.nfev2) L.nfev2[(long long)b * L.Nbeta + ib] = s.nfev;
  if (ib + 1 < L.Nbeta) {
    lb_reset(s, o.ls_ftol, o.ls_gtol, o.ls_xtol);      // warm start: x stays where it is
    s.ib = ib + 1;
    L.rf_path[b] = L.scales[ib + 1];
    act_eval[b] = 1;
  } else {
    s.finished = 1;
    s.accepted = 0; s.redo_dir = 0;
    act_eval[b] = 0;
  }
}

__global__ void lb_advance_kernel(LbPath* st, int* act_eval, int B, LbLadder L, LbOpts o) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  lb_advance_body(st[b], act_eval, b, L, o);
}

// ---------------------------------------------------------------------------------------------
// Small problems (the shipped Lorenz96 example: n = 3221): a cycle of eleven dependent launches
// costs ~60 us of launch latency for a few microseconds of work.  Everything that follows the
// evaluation -- g.d, the line-search decision, the history update, the two-loop recursion, the new
// direction, the start of the next search, the end-of-rung bookkeeping and the next trial point --
// is done here by ONE thread-block cluster per path: FCS CTAs, each owning one chunk of the
// vectors (read from L2; a CTA only ever touches its own chunk), partial dot products exchanged
// through distributed shared memory, phases separated by cluster barriers, decisions taken by
// thread 0 of CTA 0 with the same code (lb_*_body) as the per-phase kernels and published to the
// other CTAs through CTA 0's shared memory.  A cycle is then three launches (walk, finalize, this).
// Unbounded L-BFGS and CG; separate trial buffer XT.  What it buys is modest -- replayed from a CUDA
// graph the eleven small kernels cost ~1.7 us each, not the 5 us of an eager launch: the fused
// kernel itself takes ~20 us (the update pass and the coefficient-space recursion dominate), the
// walk ~12 us -- so it is used only while every cluster gets its own SMs (B * FCS <= SM count).
constexpr int FCS = 8;                  // CTAs per path (cluster size)

struct FusedFlags {
  int finished, done, need_eval, first, accepted, do_update, redo_dir, pslot, col, ib;
  double stp, cg, cd;
  double cs[MMAX], cy[MMAX];
};
__device__ void fused_publish(FusedFlags& F, const LbPath& s) {
  F.finished = s.finished; F.done = s.done; F.need_eval = s.need_eval; F.first = s.first;
  F.accepted = s.accepted; F.do_update = s.do_update; F.redo_dir = s.redo_dir; F.pslot = s.pslot;
  F.col = s.col; F.ib = s.ib;
  F.stp = s.stp; F.cg = s.cg; F.cd = s.cd;
  for (int j = 0; j < MMAX; ++j) { F.cs[j] = s.cs[j]; F.cy[j] = s.cy[j]; }
}

__global__ void __cluster_dims__(FCS, 1, 1) __launch_bounds__(NT, 1) lb_fused_kernel(
    double* X, double* G, double* XT, const double* GT, double* Dv, double* S, double* Y, long long ld,
    long long n, long long hstride, LbPath* st, int* act_eval, const double* ft, const double* met,
    const double* fet, LbOpts o, LbLadder L, long long* dbg) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double scratch[8 * NT];
  __shared__ double res[NACC_U];
  __shared__ double sum[NACC_U];
  __shared__ double SYs[MMAX * MMAX], YYs[MMAX * MMAX];
  __shared__ int ops[NACC_U];
  __shared__ FusedFlags F;
  const int rank = (int)cl.block_rank(), tid = threadIdx.x;
  const int b = blockIdx.x / FCS;
  const bool lead = (rank == 0 && tid == 0);
  // development aid (VAB_FUSED_TIMING=1): SM cycles spent up to each phase boundary, summed over calls
  long long t_prev = 0;
  int t_slot = 0;
  auto stamp = [&]() {
    if (dbg != nullptr && lead && blockIdx.x == 0) {
      const long long t = clock64();
      if (t_slot > 0) dbg[t_slot] += t - t_prev;
      t_prev = t;
      ++t_slot;
    }
  };
  stamp();
  // the path's state lives in CTA 0's shared memory for the duration of the kernel: the decision
  // code is a single thread chasing a few hundred dependent loads and stores, which costs tens of
  // microseconds against global memory and next to nothing here
  __shared__ LbPath sp;
  static_assert(sizeof(LbPath) % 8 == 0, "LbPath is copied as 8-byte words");
  LbPath& s = sp;                                     // touched by CTA 0 only
  if (rank == 0) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(st + b);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sp);
    for (int k = tid; k < (int)(sizeof(LbPath) / 8); k += NT) dst[k] = src[k];
    __syncthreads();
  }
  const FusedFlags* F0 = cl.map_shared_rank(&F, 0);
  const long long off = (long long)b * ld;
  const Range r = chunk_range(n, FCS, rank);
  const int m = o.m;
  // sum over the CTAs' res[k], in rank order (fixed order: bit-reproducible)
  auto gather = [&](int k, int op) {
    double a = 0.0;
    for (int q = 0; q < FCS; ++q) {
      const double u = *cl.map_shared_rank(&res[k], q);
      a = (q == 0) ? u : ((op == RED_SUM) ? a + u : (op == RED_MAX ? fmax(a, u) : fmin(a, u)));
    }
    return a;
  };

  if (lead) fused_publish(F, s);
  cl.sync();
  stamp();
  const bool fin = F0->finished != 0;
  cl.sync();                                          // (CTA 0 must not leave before the others have read the flag)
  stamp();
  if (fin) return;                                    // uniform over the cluster (nothing to write back)

  // ---- g.d and max |g| of the trial point (lb_gd_kernel), then the line-search decision
  const bool ph1 = !F0->done && F0->need_eval;
  if (ph1) {
    const bool first = F0->first != 0;
    double v[2] = {0.0, 0.0};
    for (long long i = r.i0 + tid; i < r.i1; i += NT) {
      const double gi = GT[off + i];
      if (!first) v[0] = fma(gi, Dv[off + i], v[0]);
      v[1] = fmax(v[1], fabs(gi));
    }
    const int op2[2] = {RED_SUM, RED_MAX};
    block_reduce<2>(v, op2, res, scratch);
  }
  cl.sync();
  stamp();
  if (lead) {
    if (ph1) lb_linesearch_body(s, act_eval + b, ft[b], met[b], fet[b], gather(0, RED_SUM), gather(1, RED_MAX), o);
    fused_publish(F, s);
  }
  cl.sync();
  stamp();

  // ---- accepted step: x <- xt, g <- gt, new pair into slot p, dot products (lb_update_kernel)
  const bool ph2 = F0->accepted != 0;
  const bool ph3 = (F0->accepted || F0->redo_dir) && !F0->done;
  const bool done_now = F0->done != 0;
  if (ph2) {
    const bool upd = F0->do_update != 0;
    const int p = F0->pslot, col = F0->col;
    const double stp = F0->stp;
    double acc[NACC_U];
#pragma unroll
    for (int k = 0; k < NACC_U; ++k) acc[k] = 0.0;
    for (long long i = r.i0 + tid; i < r.i1; i += NT) {
      const double xt = XT[off + i], gt = GT[off + i];
      const double gold = G[off + i];
      double sv = 0.0, yv = 0.0;
      if (upd) {
        sv = stp * Dv[off + i];
        yv = gt - gold;
      }
      const double gh = gt;
#pragma unroll
      for (int j = 0; j < MMAX; ++j) {
        if (j < m && (j < col || col == m) && !(upd && j == p)) {
          const double sj = S[(long long)j * hstride + off + i];
          const double yj = Y[(long long)j * hstride + off + i];
          acc[j] = fma(gh, sj, acc[j]);
          acc[MMAX + j] = fma(gh, yj, acc[MMAX + j]);
          acc[2 * MMAX + j] = fma(sv, yj, acc[2 * MMAX + j]);
          acc[3 * MMAX + j] = fma(yv, sj, acc[3 * MMAX + j]);
          acc[4 * MMAX + j] = fma(yv, yj, acc[4 * MMAX + j]);
        }
      }
      acc[5 * MMAX] = fma(yv, yv, acc[5 * MMAX]);
      acc[5 * MMAX + 1] = fma(gh, sv, acc[5 * MMAX + 1]);
      acc[5 * MMAX + 2] = fma(gh, yv, acc[5 * MMAX + 2]);
      acc[5 * MMAX + 3] = fma(gh, gh, acc[5 * MMAX + 3]);
      X[off + i] = xt;
      G[off + i] = gt;
      if (upd) {
        S[(long long)p * hstride + off + i] = sv;
        Y[(long long)p * hstride + off + i] = yv;
      }
    }
    cta_reduce_sum<NACC_U>(acc, res, scratch);
  }
  cl.sync();
  stamp();
  // ---- history bookkeeping and two-loop recursion (lb_gram_kernel), CTA 0
  if (rank == 0) {
    if (ph3) {
      if (ph2)
        for (int k = tid; k < NACC_U; k += NT) sum[k] = gather(k, RED_SUM);
      for (int k = tid; k < MMAX * MMAX; k += NT) { SYs[k] = s.SY[k]; YYs[k] = s.YY[k]; }
      __syncthreads();
      if (o.method == 1) { if (tid == 0) lb_gram_body(s, sum, SYs, YYs, m, o.method); }
      else if (tid < 32) lb_gram_warp(s, sum, SYs, YYs, m);
      __syncthreads();
    }
    if (tid == 0) fused_publish(F, s);
  }
  cl.sync();
  stamp();

  // ---- d = -(cg g + sum_j cs_j S_j + cy_j Y_j) [+ cd d_old]; d.d, g.d (lb_direction_kernel)
  if (ph3) {
    const int col = F0->col;
    double cs[MMAX], cy[MMAX];
#pragma unroll
    for (int j = 0; j < MMAX; ++j) { cs[j] = F0->cs[j]; cy[j] = F0->cy[j]; }
    const double cgc = F0->cg, cd = F0->cd;
    double v[3] = {0.0, 0.0, BIG};
    for (long long i = r.i0 + tid; i < r.i1; i += NT) {
      const double g = G[off + i];
      double rr = cgc * g;
#pragma unroll
      for (int j = 0; j < MMAX; ++j) {
        if (j < m && (j < col || col == m)) {
          rr = fma(cs[j], S[(long long)j * hstride + off + i], rr);
          rr = fma(cy[j], Y[(long long)j * hstride + off + i], rr);
        }
      }
      double d = -rr;
      if (cd != 0.0) d = fma(cd, Dv[off + i], d);
      Dv[off + i] = d;
      v[0] = fma(d, d, v[0]);
      v[1] = fma(g, d, v[1]);
    }
    const int op3[3] = {RED_SUM, RED_SUM, RED_MIN};
    block_reduce<3>(v, op3, res, scratch);
  }
  cl.sync();
  stamp();
  // ---- start of the next line search (lb_start_kernel)
  if (lead) {
    if (!ph3) { if (done_now) { s.accepted = 0; s.redo_dir = 0; s.restore = 0.0; } }
    else if (s.abort_dir) s.abort_dir = 0;
    else lb_start_body(s, act_eval + b, gather(0, RED_SUM), gather(1, RED_SUM), gather(2, RED_MIN), o, 0, 0, 0);
    fused_publish(F, s);
  }
  cl.sync();
  stamp();

  // ---- end of a rung: minimiser to minpaths, table row, next rung (lb_save_kernel + lb_advance_kernel)
  const bool ph5 = F0->done && !F0->finished;
  const int ib = F0->ib;
  cl.sync();                                          // everyone has read the flags of this phase
  stamp();
  if (ph5) {
    if (L.minpaths != nullptr) {
      double* out = L.minpaths + ((long long)b * L.Nbeta + ib) * L.mp_pitch;
      if (L.winw != n) {
        const long long w0 = r.i0 > L.win0 ? r.i0 : L.win0;
        const long long w1 = r.i1 < L.win0 + L.winw ? r.i1 : L.win0 + L.winw;
        for (long long i = w0 + tid; i < w1; i += NT) out[i - L.win0] = X[off + i];
      } else {
        for (long long i = r.i0 + tid; i < r.i1; i += NT) out[i] = X[off + i];
      }
    }
    if (lead) { lb_advance_body(s, act_eval, b, L, o); fused_publish(F, s); }
  }
  cl.sync();
  stamp();

  // ---- trial point of the next evaluation (lb_trial_kernel)
  if (!F0->done && F0->need_eval) {
    if (F0->first) {
      for (long long i = r.i0 + tid; i < r.i1; i += NT) XT[off + i] = X[off + i];
    } else {
      const double stp = F0->stp;
      for (long long i = r.i0 + tid; i < r.i1; i += NT) XT[off + i] = fma(stp, Dv[off + i], X[off + i]);
    }
  }
  cl.sync();                                          // no CTA leaves while its shared memory may still be read
  stamp();
  if (dbg != nullptr && lead && blockIdx.x == 0) dbg[0] += 1;
  if (rank == 0) {                                    // state back to global memory
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&sp);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(st + b);
    for (int k = tid; k < (int)(sizeof(LbPath) / 8); k += NT) dst[k] = src[k];
  }
}

}  // namespace
#include "lb_resident.cuh"
namespace {

__global__ void lb_count_kernel(const LbPath* st, int B, int* n_running, int* prog, int Nbeta) {
  int c = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int fin = st[b].finished;
    c += fin ? 0 : 1;
    if (prog) prog[b] = fin ? Nbeta : st[b].ib;        // rungs this path has completed
  }
  for (int sft = 16; sft > 0; sft >>= 1) c += __shfl_down_sync(0xffffffffu, c, sft);
  __shared__ int w[8];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < (int)blockDim.x / 32; ++k) t += w[k];
    *n_running = t;
  }
}

// results of the rung a path finished last (vab_minimize: the only one)
__global__ void lb_export_kernel(const LbPath* st, int B, double* A, double* me, double* fe,
                                 int* status, int* nit, int* nfev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const LbPath& s = st[b];
  if (A) A[b] = s.f;
  if (me) me[b] = s.me;
  if (fe) fe[b] = s.fe;
  if (status) status[b] = s.status;
  if (nit) nit[b] = s.iter;
  if (nfev) nfev[b] = s.nfev;
}

#define LB_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return vab_cuda_fail(ctx, e_, #call);              \
  } while (0)

int lb_reserve(vab_ctx* ctx, int B, long long ld, int m, int Nbeta, bool need_xt, bool gcp) {
  if (!ctx->lb) ctx->lb = new LbfgsWork();
  LbfgsWork* w = ctx->lb;
  // GT, G, D, S[m], Y[m] (+ XT: the trial buffer of the bounded / CG paths; the unbounded L-BFGS
  // path moves x in place and does without it -- one vector per path less at the C3 scale)
  // (+ Z, T, R: Cauchy point / projected point, breakpoints, subspace residual of the bounded path)
  const size_t nvec = (size_t)(2 * m + 3 + (need_xt ? 1 : 0) + (gcp ? 3 : 0));
  const size_t need = nvec * (size_t)B * (size_t)ld;
  int rc = vab_reserve(ctx, &w->vec, &w->vec_cap, need);
  if (rc != VAB_OK) return rc;
  if (B > w->st_cap) {
    cudaFree(w->st); cudaFree(w->act_eval); cudaFree(w->ft); cudaFree(w->met); cudaFree(w->fet);
    w->st = nullptr; w->act_eval = nullptr; w->ft = w->met = w->fet = nullptr;
    LB_CUDA(cudaMalloc((void**)&w->st, sizeof(LbPath) * B));
    LB_CUDA(cudaMalloc((void**)&w->act_eval, sizeof(int) * B));
    LB_CUDA(cudaMalloc((void**)&w->ft, sizeof(double) * B));
    LB_CUDA(cudaMalloc((void**)&w->met, sizeof(double) * B));
    LB_CUDA(cudaMalloc((void**)&w->fet, sizeof(double) * B));
    w->st_cap = B;
  }
  int nchunk = lb_nchunk(ctx->n_unknowns(), B);
  if (nchunk < FCS) nchunk = FCS;                 // small problems use FCS fixed chunks (lb_run)
  rc = vab_reserve(ctx, &w->part, &w->part_cap, (size_t)B * nchunk * (gcp ? NPB : NACC_U));
  if (rc != VAB_OK) return rc;
  rc = vab_reserve(ctx, &w->lad, &w->lad_cap, (size_t)2 * Nbeta + B);
  if (rc != VAB_OK) return rc;
  if (B > w->prog_cap) {
    cudaFree(w->prog_dev);
    if (w->prog_host) cudaFreeHost(w->prog_host);
    w->prog_dev = nullptr; w->prog_host = nullptr;
    LB_CUDA(cudaMalloc((void**)&w->prog_dev, 2 * sizeof(int) * B));
    LB_CUDA(cudaMallocHost((void**)&w->prog_host, 2 * sizeof(int) * B));
    w->prog_cap = B;
  }
  if (!w->copy_stream) LB_CUDA(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
  if (!w->n_running_dev) LB_CUDA(cudaMalloc((void**)&w->n_running_dev, 3 * sizeof(int)));   // [2] = boxed flag
  if (!w->n_running_host) LB_CUDA(cudaMallocHost((void**)&w->n_running_host, 2 * sizeof(int)));
  for (int k = 0; k < 2; ++k)
    if (!w->ev[k]) LB_CUDA(cudaEventCreateWithFlags(&w->ev[k], cudaEventDisableTiming));
  w->B = B; w->ld = ld; w->m = m;
  return VAB_OK;
}

// Runs every path down the ladder scales_host[0..Nbeta) (Nbeta = 1: a plain minimisation) from XP,
// in place; leaves the per-path state of the last rung in w->st.
struct LbSink { double* host = nullptr; long long pitch = 0, width = 0; long long win0 = 0, winw = -1; };
// the sink armed by vab_set_path_sink belongs to exactly one call: taken (and cleared) before that
// call validates anything, so no failure path can leave a stale host pointer behind
LbSink lb_take_sink(vab_ctx* ctx) {
  LbSink k;
  k.host = ctx->sink_host; k.pitch = ctx->sink_pitch; k.width = ctx->sink_width;
  k.win0 = ctx->win0; k.winw = ctx->winw;
  ctx->sink_host = nullptr; ctx->sink_pitch = 0; ctx->sink_width = 0;
  ctx->win0 = 0; ctx->winw = -1;
  return k;
}

int lb_run(vab_ctx* ctx, int B, double* XP, long long ld, const double* scales_host,
           const double* betas_host, int Nbeta, const vab_lbfgs_opts* uo, const double* lo,
           const double* hi, double* table, double* minpaths, int* status2, int* nit2, int* nfev2,
           LbSink sink_req) {
  const long long n = ctx->n_unknowns();
  if (n <= 0) return vab_fail(ctx, VAB_ERR_STATE, "minimize: no problem set on this context");
  if (B < 1 || !XP || ld < n || (ld & 1)) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: bad batch / XP / ldxp");
  if (((uintptr_t)XP & 15) || ((uintptr_t)minpaths & 15))
    return vab_fail(ctx, VAB_ERR_INVALID, "minimize: XP / minpaths must be 16-byte aligned");
  const bool windowed = sink_req.winw >= 0 && !(sink_req.win0 == 0 && sink_req.winw == n);
  if (windowed && (sink_req.win0 < 0 || sink_req.win0 + sink_req.winw > n))
    return vab_fail(ctx, VAB_ERR_INVALID, "anneal: path window outside [0, n)");
  if (windowed && sink_req.host) return vab_fail(ctx, VAB_ERR_INVALID, "anneal: a path window and a host sink exclude each other");
  if ((lo == nullptr) != (hi == nullptr)) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: give both bounds or none");
  LbOpts o;
  o.m = uo && uo->m > 0 ? uo->m : 10;
  if (o.m > MMAX) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: maxcor > 10 is not supported");
  o.maxls = uo && uo->maxls > 0 ? uo->maxls : 20;
  o.maxfun = uo && uo->maxfun > 0 ? uo->maxfun : 15000;
  o.maxiter = uo && uo->maxiter > 0 ? uo->maxiter : 15000;
  o.ftol = uo ? uo->ftol : 2.220446049250313e-09;
  o.pgtol = uo ? uo->pgtol : 1e-5;
  o.method = uo ? uo->method : 0;
  if (o.method != 0 && o.method != 1) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: method must be 0 (L-BFGS-B) or 1 (CG)");
  o.ls_ftol = 1e-3; o.ls_gtol = 0.9; o.ls_xtol = 0.1;         // L-BFGS-B's lnsrlb
  if (o.method == 1) {
    o.m = 1;
    o.ls_ftol = 1e-4; o.ls_gtol = 0.4; o.ls_xtol = 1e-14;      // scipy.optimize fmin_cg (c1, c2)
    if (!(uo && uo->maxls > 0)) o.maxls = 100;
  }
  int poll = uo && uo->poll_every > 0 ? uo->poll_every : 0;
  const bool bounded = lo != nullptr;
  // the TMA-fed history passes serve the unbounded L-BFGS case (VAB_LBFGS_TMA=0: plain kernels)
  bool use_tma = !bounded && o.method == 0;
  if (const char* e = getenv("VAB_LBFGS_TMA")) use_tma = use_tma && atoi(e) != 0;
  // small unbounded problems are launch-latency bound: one cluster per path does everything
  // behind the evaluation in a single launch (lb_fused_kernel); VAB_LBFGS_FUSED=0/1 overrides
  // (measured on the shipped Lorenz96 example, n = 3221: 37 us per cycle against 50 us for one path,
  // 50 against 58 us for 16; with more clusters than SMs the per-phase kernels win again)
  // Small problems (n <= 32768) run the plain kernels on FCS fixed chunks per path; the fused kernel
  // does exactly the same arithmetic in the same order (chunks, reduction trees, recursion), so the
  // choice between the two is pure scheduling and the results are bit-identical
  // (tests/test_gpu_ladder.py::test_fused_cycle_kernel_is_bit_identical).
  const bool small = !bounded && n <= 32768;
  // (a cluster needs its FCS SMs inside one GPC -- 18 SMs on B200, i.e. two clusters per GPC: stay a
  //  little below SM count / FCS so that every cluster is resident at once)
  bool fused = small && B <= ctx->num_sms / FCS - 2;
  if (const char* e = getenv("VAB_LBFGS_FUSED")) fused = small && atoi(e) != 0;
  if (small) use_tma = false;
  // Lorenz96 problems small enough for an 8-CTA cluster's shared memory run their whole ladder inside
  // one launch (lb_resident.cuh) while every cluster is resident at once (two per GPC);
  // VAB_LBFGS_RESIDENT=0/1 overrides.  The action is evaluated by a second implementation there, so
  // its results agree with the other paths to rounding, not bit for bit.
  bool resident = false;
  int res_rpc = 0, res_cs = 0;
  size_t res_smem = 0;
  ResKernel res_kernel = nullptr;
  if (small && o.method == 0 && ctx->problem == VAB_PROBLEM_ODE) {
    const vab_ode_desc& d = ctx->od;
    if (d.model == 0 && d.NP == 1 && d.n_stim == 0 && !ctx->ptime && !ctx->rm_dev && !ctx->rm_matrix &&
        !ctx->rf0_dev && !ctx->rf0_mat && d.disc != VAB_DISC_RK4) {
      // clusters of 8 CTAs give the shortest cycle; clusters of 4 let twice as many paths be resident
      // at once.  Larger batches run in rounds of clusters, scheduled as clusters finish; measured on C1
      // that stays ahead of the general path at every batch size tried (128 paths: 3.5 s against 8.6 s;
      // the general path costs ~70 ms per path-ladder at 256 paths, a round of 32 resident clusters 0.87 s).
      int want = -1;                                  // -1: decide here, 0: off, 1: on (size decided here), 4 / 8 / 16: that size
      if (const char* e = getenv("VAB_LBFGS_RESIDENT")) want = atoi(e);
      // (16 CTAs: a non-portable cluster size, for paths too long for 8 CTAs' shared memory)
      const int sizes[3] = {8, 4, 16};
      bool fits[3] = {false, false, false};
      int cap[3] = {0, 0, 0}, rpc[3] = {0, 0, 0};
      size_t smem[3] = {0, 0, 0};
      ResKernel kern[3] = {lb_resident_pick<8>(d.disc), lb_resident_pick<4>(d.disc), lb_resident_pick<16>(d.disc)};
      for (int q = 0; q < 3 && want != 0; ++q) {
        const int cs = sizes[q];
        rpc[q] = lb_resident_rows(d.N_model, cs);
        smem[q] = lb_resident_smem(rpc[q], d.D);
        if (smem[q] > (size_t)226000 || (long long)rpc[q] * cs < d.N_model) continue;
        if (cudaFuncSetAttribute(kern[q], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem[q]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (cs > 8 && cudaFuncSetAttribute(kern[q], cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(B * cs); cfg.blockDim = dim3(RNT); cfg.dynamicSmemBytes = smem[q]; cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&cap[q], kern[q], &cfg) != cudaSuccess) { cudaGetLastError(); cap[q] = 0; }
        fits[q] = cap[q] >= 1;
      }
      int pick = -1;
      if (want == 8 || want == 4 || want == 16) {
        const int q = want == 8 ? 0 : (want == 4 ? 1 : 2);
        pick = fits[q] ? q : -1;
      } else if (want != 0) {
        if (fits[0] && B <= cap[0]) pick = 0;
        else if (fits[1]) pick = 1;
        else if (fits[0]) pick = 0;
        else if (fits[2]) pick = 2;
      }
      if (pick >= 0) {
        resident = true; res_cs = sizes[pick]; res_rpc = rpc[pick]; res_smem = smem[pick]; res_kernel = kern[pick];
      }
    }
  }
  if (resident) fused = false;
  // bounded L-BFGS-B: generalised Cauchy point + subspace minimisation (lbfgsb_bounded.cuh);
  // VAB_BOUNDS=projection selects round 1's active-set projection for comparison
  bool gcp = bounded && o.method == 0;
  if (const char* e = getenv("VAB_BOUNDS")) gcp = gcp && strcmp(e, "projection") != 0;
  int rc = lb_reserve(ctx, B, ld, o.m, Nbeta, !use_tma, gcp);
  if (rc != VAB_OK) return rc;
  LbfgsWork* w = ctx->lb;
  cudaStream_t st = ctx->stream;
  const size_t vs = (size_t)B * (size_t)ld;
  double* GT = w->vec;
  double* G = w->vec + vs;
  double* Dv = w->vec + 2 * vs;
  double* S = w->vec + 3 * vs;
  double* Y = w->vec + (3 + (size_t)o.m) * vs;
  double* XT = use_tma ? XP : w->vec + (3 + 2 * (size_t)o.m) * vs;   // in-place trials: never dereferenced as a separate buffer
  double* Zc = gcp ? w->vec + (4 + 2 * (size_t)o.m) * vs : nullptr;
  double* Tb = gcp ? w->vec + (5 + 2 * (size_t)o.m) * vs : nullptr;
  double* Rs = gcp ? w->vec + (6 + 2 * (size_t)o.m) * vs : nullptr;
  int* boxed_dev = w->n_running_dev + 2;
  const long long hstride = (long long)vs;
  const int nchunk = small ? FCS : lb_nchunk(n, B);
  const dim3 vgrid(nchunk, B);
  // ring depth: 2 stages x 2 CTAs per SM (measured best on B200: C2 update 0.93 ms = 6.7 TB/s,
  // direction 0.75 ms) or 4 stages x 1 CTA per SM (VAB_LBFGS_NS=4: 0.99 / 0.78 ms)
  int hns = 2;
  if (const char* e = getenv("VAB_LBFGS_NS")) hns = atoi(e) == 4 ? 4 : 2;
  if (use_tma) {
    if (!ctx->attr_lbfgs) {
      LB_CUDA(cudaFuncSetAttribute(lb_update_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem(U_NSTR, 4)));
      LB_CUDA(cudaFuncSetAttribute(lb_direction_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem(D_NSTR, 4)));
      LB_CUDA(cudaFuncSetAttribute(lb_update_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem(U_NSTR, 2)));
      LB_CUDA(cudaFuncSetAttribute(lb_direction_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem(D_NSTR, 2)));
      ctx->attr_lbfgs = true;
    }
  }

  LbLadder L;
  L.Nbeta = Nbeta;
  L.scales = w->lad; L.betas = w->lad + Nbeta; L.rf_path = w->lad + 2 * (size_t)Nbeta;
  L.table = table; L.minpaths = minpaths; L.status2 = status2; L.nit2 = nit2; L.nfev2 = nfev2;
  L.win0 = windowed ? sink_req.win0 : 0;
  L.winw = windowed ? sink_req.winw : n;
  L.mp_pitch = windowed ? ((sink_req.winw + 1) & ~1LL) : ld;
  if (L.mp_pitch < 2) L.mp_pitch = 2;
  {
    std::string tmp((size_t)2 * Nbeta * sizeof(double), '\0');
    double* t = reinterpret_cast<double*>(&tmp[0]);
    for (int i = 0; i < Nbeta; ++i) { t[i] = scales_host[i]; t[Nbeta + i] = betas_host ? betas_host[i] : 0.0; }
    LB_CUDA(cudaMemcpyAsync(w->lad, t, tmp.size(), cudaMemcpyHostToDevice, st));
    LB_CUDA(cudaStreamSynchronize(st));           // tmp dies with this scope
  }
  lb_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(w->st, w->act_eval, B, o.ls_ftol, o.ls_gtol, o.ls_xtol,
                                                   L.scales, L.rf_path);
  if (bounded) lb_clip_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, st>>>(XP, ld, n, lo, hi);
  if (bounded) lbb_boxed_kernel<<<1, 256, 0, st>>>(lo, hi, n, boxed_dev);
  else LB_CUDA(cudaMemsetAsync(boxed_dev, 0, sizeof(int), st));
  LB_CUDA(cudaMemsetAsync(Dv, 0, vs * sizeof(double), st));
  LB_CUDA(cudaMemsetAsync(G, 0, vs * sizeof(double), st));
  ctx->launches += 2;

  if (poll <= 0) {
    // small problems are launch-bound (a cycle is ~30 us): poll rarely; large ones take ms per cycle
    poll = (n * (long long)B < (1LL << 22)) ? 64 : 8;
  }
  long long* fused_dbg = nullptr;
  if ((fused && getenv("VAB_FUSED_TIMING")) || (resident && getenv("VAB_RESIDENT_TIMING"))) {
    if (cudaMalloc((void**)&fused_dbg, 32 * sizeof(long long)) == cudaSuccess) cudaMemset(fused_dbg, 0, 32 * sizeof(long long));
    else fused_dbg = nullptr;
  }
  if (fused) {                                 // the first trial point (x itself); later ones come from lb_fused_kernel
    lb_trial_kernel<<<vgrid, NT, 0, st>>>(XT, XP, Dv, ld, n, w->st, nchunk, 0, nullptr);
    ctx->launches += 1;
  }
  auto enqueue_cycle = [&](cudaStream_t st) -> int {
    if (fused) {
      int r = vab_eval(ctx, B, XT, ld, 1.0, L.rf_path, w->act_eval, w->ft, w->met, w->fet, GT, ld);
      if (r != VAB_OK) return r;
      lb_fused_kernel<<<B * FCS, NT, 0, st>>>(XP, G, XT, GT, Dv, S, Y, ld, n, hstride, w->st, w->act_eval,
                                              w->ft, w->met, w->fet, o, L, fused_dbg);
      ctx->launches += 1;
      return VAB_OK;
    }
    const int inplace = use_tma ? 1 : 0;       // unbounded L-BFGS: x moves in place, no trial buffer
    lb_trial_kernel<<<vgrid, NT, 0, st>>>(XT, XP, Dv, ld, n, w->st, nchunk, inplace, Zc);
    int r = vab_eval(ctx, B, inplace ? XP : XT, ld, 1.0, L.rf_path, w->act_eval, w->ft, w->met, w->fet, GT, ld);
    if (r != VAB_OK) return r;
    if (bounded) lb_gd_kernel<true><<<vgrid, NT, 0, st>>>(XT, GT, Dv, ld, n, lo, hi, w->st, nchunk, w->part);
    else lb_gd_kernel<false><<<vgrid, NT, 0, st>>>(XT, GT, Dv, ld, n, lo, hi, w->st, nchunk, w->part);
    lb_linesearch_kernel<<<B, 32, 0, st>>>(w->st, w->act_eval, w->ft, w->met, w->fet, w->part, nchunk, o, bounded ? 1 : 0);
    if (gcp) {
      // bounded: update, Cauchy point, subspace minimisation, projection, direction d = z - x
      lbb_update_kernel<<<vgrid, NT, 0, st>>>(XP, G, XT, GT, Dv, S, Y, ld, n, hstride, w->st, o.m, nchunk, w->part);
      lbb_hist_kernel<<<B, 32, 0, st>>>(w->st, w->part, nchunk, o.m);
      lbb_cauchy_prep_kernel<<<vgrid, NT, 0, st>>>(XP, G, Dv, Tb, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part);
      lbb_cauchy_loop_kernel<<<B, CLT, 0, st>>>(XP, Dv, Tb, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part);
      lbb_gram_free_kernel<<<vgrid, NT, 0, st>>>(Tb, S, Y, ld, n, hstride, w->st, o.m, nchunk, w->part);
      lbb_resid_kernel<<<vgrid, NT, 0, st>>>(XP, G, Dv, Tb, Zc, Rs, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part);
      lbb_subsolve_kernel<<<B, 32, 0, st>>>(w->st, w->part, nchunk, o.m);
      lbb_step_kernel<<<vgrid, NT, 0, st>>>(XP, G, Tb, Zc, Rs, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part);
      lbb_decide_kernel<<<B, 32, 0, st>>>(w->st, w->part, nchunk);
      lbb_backtrack_kernel<<<vgrid, NT, 0, st>>>(XP, Dv, Tb, Zc, Rs, ld, n, lo, hi, w->st, nchunk);
      lbb_dir_kernel<<<vgrid, NT, 0, st>>>(XP, G, Dv, Zc, ld, n, lo, hi, w->st, nchunk, w->part);
      ctx->launches += 9;
    } else
    if (use_tma && hns == 4) lb_update_tma_kernel<4><<<vgrid, HT + 32, hist_smem(U_NSTR, 4), st>>>(XP, G, GT, Dv, S, Y, ld, n, hstride, w->st, o.m, nchunk, w->part);
    else if (use_tma) lb_update_tma_kernel<2><<<vgrid, HT + 32, hist_smem(U_NSTR, 2), st>>>(XP, G, GT, Dv, S, Y, ld, n, hstride, w->st, o.m, nchunk, w->part);
    else if (bounded) lb_update_kernel<true><<<vgrid, NT, 0, st>>>(XP, G, XT, GT, Dv, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part, 0);
    else lb_update_kernel<false><<<vgrid, NT, 0, st>>>(XP, G, XT, GT, Dv, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part, 0);
    if (!gcp) lb_gram_kernel<<<B, 64, 0, st>>>(w->st, w->part, nchunk, o.m, 0, o.method);
    if (gcp) { /* direction formed above */ } else
    if (use_tma && hns == 4) lb_direction_tma_kernel<4><<<vgrid, HT + 32, hist_smem(D_NSTR, 4), st>>>(G, Dv, S, Y, XP, ld, n, hstride, w->st, o.m, nchunk, w->part);
    else if (use_tma) lb_direction_tma_kernel<2><<<vgrid, HT + 32, hist_smem(D_NSTR, 2), st>>>(G, Dv, S, Y, XP, ld, n, hstride, w->st, o.m, nchunk, w->part);
    else if (bounded) lb_direction_kernel<true><<<vgrid, NT, 0, st>>>(XP, G, Dv, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part, 0);
    else lb_direction_kernel<false><<<vgrid, NT, 0, st>>>(XP, G, Dv, S, Y, ld, n, hstride, lo, hi, w->st, o.m, nchunk, w->part, 0);
    lb_start_kernel<<<B, 1, 0, st>>>(w->st, w->act_eval, w->part, nchunk, o, bounded ? 1 : 0, 0, use_tma ? 1 : 0, boxed_dev);
    if (minpaths) { lb_save_kernel<<<vgrid, NT, 0, st>>>(XP, ld, n, w->st, nchunk, L); ctx->launches += 1; }
    lb_advance_kernel<<<(B + 127) / 128, 128, 0, st>>>(w->st, w->act_eval, B, L, o);
    ctx->launches += 8;
    return VAB_OK;
  };
  // Everything this call owns besides the workspaces: the captured graph, and the guarantee that
  // neither the context's stream nor the sink's copy stream still touches caller memory when the
  // call returns -- on every exit path.
  struct Cleanup {
    cudaStream_t st, copy;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    long long* dbg = nullptr;
    ~Cleanup() {
      cudaStreamSynchronize(st);
      if (copy) cudaStreamSynchronize(copy);
      if (dbg) cudaFree(dbg);
      if (gexec) cudaGraphExecDestroy(gexec);
      if (graph) cudaGraphDestroy(graph);
    }
  } own{st, w->copy_stream};
  own.dbg = fused_dbg;
  // The first cycle runs eagerly (it may allocate workspaces and opt kernels in to large shared
  // memory); the steady-state cycle is then captured once into a CUDA graph and replayed.  Capture
  // is illegal on the legacy default stream (which is what PyTorch's current stream is unless the
  // caller set one): the cycle is then recorded on a private capture stream -- recording executes
  // nothing -- and the instantiated graph is launched into the context's stream like any kernel.
  if (resident) {
    const vab_ode_desc& d = ctx->od;
    ResArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.XP = XP; ra.ld = ld; ra.N = d.N_model; ra.D = d.D; ra.nskip = d.nskip; ra.nobs = d.L; ra.RPC = res_rpc; ra.disc = d.disc;
    ra.dt = d.dt_model; ra.cf = 1.0 / ((double)d.D * (d.N_model - 1)); ra.rf0 = ctx->rf0_scalar;
    ra.Y = ctx->Y_dense; ra.wobs = ctx->wobs_dev; ra.k_est = d.NPest == 1 ? 1 : 0;
    ra.pfix = ctx->pfix_dev; ra.pfix_stride = ctx->pfix_stride;
    ra.st = w->st; ra.o = o; ra.L = L;
    {
      double mc0 = (double)Nbeta * ((double)o.maxfun + (double)o.maxiter + 64.0);
      ra.max_cycles = mc0 < 9.0e18 ? (long long)mc0 : (long long)9.0e18;
    }
    ra.dbg = fused_dbg;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = res_cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(B * res_cs); cfg.blockDim = dim3(RNT); cfg.dynamicSmemBytes = res_smem; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = 1;
    LB_CUDA(cudaLaunchKernelEx(&cfg, res_kernel, ra));
    LB_CUDA(cudaGetLastError());
    ctx->launches += 1;
  } else {
    rc = enqueue_cycle(st);
    if (rc != VAB_OK) return rc;
    LB_CUDA(cudaGetLastError());
  }
  bool use_graph = !resident;
  if (const char* e = getenv("VAB_LBFGS_GRAPH")) use_graph = use_graph && atoi(e) != 0;
  long long launches_per_cycle = 0;
  if (use_graph) {
    cudaStream_t cap = st;
    if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
      if (!w->cap_stream && cudaStreamCreateWithFlags(&w->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        w->cap_stream = nullptr;
        cudaGetLastError();
      }
      cap = w->cap_stream;
    }
    const long long l0 = ctx->launches;
    cudaError_t ce = cap ? cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) : cudaErrorStreamCaptureUnsupported;
    if (ce == cudaSuccess) {
      ctx->stream = cap;                        // vab_eval launches on the context's stream
      rc = enqueue_cycle(cap);
      ctx->stream = st;
      cudaError_t ce2 = cudaStreamEndCapture(cap, &own.graph);
      launches_per_cycle = ctx->launches - l0;
      ctx->launches = l0;
      if (rc != VAB_OK || ce2 != cudaSuccess || own.graph == nullptr ||
          cudaGraphInstantiate(&own.gexec, own.graph, 0) != cudaSuccess) {
        if (own.graph) cudaGraphDestroy(own.graph);
        own.graph = nullptr; own.gexec = nullptr;
        cudaGetLastError();
        if (rc != VAB_OK) return rc;
      }
    } else {
      cudaGetLastError();
    }
  }
  // Polling is double-buffered: the next group of cycles is enqueued before the host waits for
  // the counter of the previous group, so the device never idles on the host (the price is at
  // most one group of empty cycles at the end).
  // host sink (vab_set_path_sink): rows of minpaths are sent home as the paths finish their rungs
  double* sink = (minpaths != nullptr) ? sink_req.host : nullptr;
  const long long sink_pitch = sink_req.pitch, sink_width = sink_req.width;
  if (sink != nullptr && sink_width > ld) return vab_fail(ctx, VAB_ERR_INVALID, "anneal: sink width > ldxp");
  std::vector<int> sent(sink ? B : 0, 0);
  auto send_rows = [&](const int* prog) -> cudaError_t {
    for (int b = 0; b < B; ++b) {
      const int upto = prog ? prog[b] : Nbeta;
      for (; sent[b] < upto; ++sent[b]) {
        const size_t row = (size_t)b * Nbeta + sent[b];
        cudaError_t ce = cudaMemcpyAsync(sink + row * (size_t)sink_pitch, minpaths + row * (size_t)ld,
                                         (size_t)sink_width * sizeof(double), cudaMemcpyDeviceToHost, w->copy_stream);
        if (ce != cudaSuccess) return ce;
      }
    }
    return cudaSuccess;
  };
  int rc_loop = VAB_OK;
  long long groups = 0, cycles = 0;
  double mc = (double)Nbeta * ((double)o.maxfun + (double)o.maxiter + 64.0) + 4.0 * poll;
  const long long max_cycles = mc < 9.0e18 ? (long long)mc : (long long)9.0e18;
  const double t_limit = getenv("VAB_LBFGS_MAX_SECONDS") ? atof(getenv("VAB_LBFGS_MAX_SECONDS")) : 0.0;
  const auto t0 = std::chrono::steady_clock::now();
  int gsize = poll < 4 ? poll : 4;              // groups grow 4, 8, ... poll: short runs waste little
  while (!resident) {
    const int k = (int)(groups & 1);
    for (int c = 0; c < gsize; ++c) {
      if (own.gexec) {
        cudaError_t ge = cudaGraphLaunch(own.gexec, st);
        if (ge != cudaSuccess) { rc_loop = vab_cuda_fail(ctx, ge, "cudaGraphLaunch"); break; }
        ctx->launches += launches_per_cycle;
        ctx->graph_launches += 1;
      } else {
        rc_loop = enqueue_cycle(st);
        if (rc_loop != VAB_OK) break;
      }
    }
    if (rc_loop != VAB_OK) break;
    cycles += gsize;
    if (gsize < poll) gsize = (2 * gsize < poll) ? 2 * gsize : poll;
    lb_count_kernel<<<1, 256, 0, st>>>(w->st, B, w->n_running_dev + k, sink ? w->prog_dev + (size_t)k * B : nullptr, Nbeta);
    ctx->launches += 1;
    LB_CUDA(cudaMemcpyAsync(w->n_running_host + k, w->n_running_dev + k, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (sink) LB_CUDA(cudaMemcpyAsync(w->prog_host + (size_t)k * B, w->prog_dev + (size_t)k * B, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    LB_CUDA(cudaEventRecord(w->ev[k], st));
    groups += 1;
    if (groups >= 2) {
      LB_CUDA(cudaEventSynchronize(w->ev[1 - k]));
      LB_CUDA(cudaGetLastError());
      if (sink) LB_CUDA(send_rows(w->prog_host + (size_t)(1 - k) * B));
      if (w->n_running_host[1 - k] == 0) break;
    }
    if (cycles > max_cycles) { rc_loop = vab_fail(ctx, VAB_ERR_STATE, "minimize: cycle limit exceeded (internal error)"); break; }
    if (t_limit > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > t_limit) {
      rc_loop = vab_fail(ctx, VAB_ERR_STATE, "minimize: VAB_LBFGS_MAX_SECONDS exceeded"); break;
    }
  }
  cudaStreamSynchronize(st);
  if (fused_dbg) {
    long long h[32];
    if (cudaMemcpy(h, fused_dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && h[0] > 0) {
      fprintf(stderr, "%s: %lld cycles; mean SM clocks per phase:", resident ? "lb_resident_kernel" : "lb_fused_kernel", h[0]);
      for (int q = 1; q < 16; ++q) fprintf(stderr, " %.0f", (double)h[q] / (double)h[0]);
      fprintf(stderr, "\n");
    }
  }
  if (sink && rc_loop == VAB_OK) {
    cudaError_t ce = send_rows(nullptr);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(w->copy_stream);
    if (ce != cudaSuccess) rc_loop = vab_cuda_fail(ctx, ce, "anneal: host sink");
  }
  return rc_loop;
}

}  // namespace

int tnc_minimize(vab_ctx* ctx, int B, double* XP, long long ld, double rf_scale, const vab_lbfgs_opts* uo,
                 const double* lo, const double* hi, double* A, double* me, double* fe, int* status, int* nit,
                 int* nfev);   // tnc.cu

void lbfgs_destroy(vab_ctx* ctx) {
  LbfgsWork* w = ctx->lb;
  if (!w) return;
  cudaFree(w->vec); cudaFree(w->st); cudaFree(w->act_eval); cudaFree(w->ft); cudaFree(w->met);
  cudaFree(w->fet); cudaFree(w->part); cudaFree(w->n_running_dev); cudaFree(w->lad); cudaFree(w->prog_dev);
  if (w->prog_host) cudaFreeHost(w->prog_host);
  if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
  if (w->cap_stream) cudaStreamDestroy(w->cap_stream);
  if (w->n_running_host) cudaFreeHost(w->n_running_host);
  for (int k = 0; k < 2; ++k)
    if (w->ev[k]) cudaEventDestroy(w->ev[k]);
  delete w;
  ctx->lb = nullptr;
}

extern "C" {

int vab_minimize(vab_ctx* ctx, int32_t B, double* XP_dev, int64_t ldxp, double rf_scale,
                 const vab_lbfgs_opts* opts, const double* lo_dev, const double* hi_dev,
                 double* A_dev, double* me_dev, double* fe_dev, int32_t* status_dev,
                 int32_t* nit_dev, int32_t* nfev_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  (void)lb_take_sink(ctx);                            // a sink only serves vab_anneal
  cudaSetDevice(ctx->device);
  if (opts && opts->method == 2) {                    // truncated Newton (min_tnc_scipy), tnc.cu
    return tnc_minimize(ctx, B, XP_dev, ldxp, rf_scale, opts, lo_dev, hi_dev, A_dev, me_dev, fe_dev, status_dev, nit_dev, nfev_dev);
  }
  int rc = lb_run(ctx, B, XP_dev, ldxp, &rf_scale, nullptr, 1, opts, lo_dev, hi_dev, nullptr, nullptr,
                  nullptr, nullptr, nullptr, LbSink());
  if (rc != VAB_OK) return rc;
  lb_export_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(ctx->lb->st, B, A_dev, me_dev, fe_dev,
                                                              status_dev, nit_dev, nfev_dev);
  ctx->launches += 1;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "vab_minimize");
  return VAB_OK;
}

int vab_anneal(vab_ctx* ctx, int32_t B, double* XP_dev, int64_t ldxp, double alpha,
               const double* beta_host, int32_t Nbeta, const vab_lbfgs_opts* opts,
               const double* lo_dev, const double* hi_dev, double* table_dev, double* minpaths_dev,
               int32_t* status_dev, int32_t* nit_dev, int32_t* nfev_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  const LbSink sink = lb_take_sink(ctx);              // cleared whatever happens below
  if (!beta_host || Nbeta < 1) return vab_fail(ctx, VAB_ERR_INVALID, "anneal: empty beta ladder");
  if (opts && opts->method == 2)
    return vab_fail(ctx, VAB_ERR_INVALID, "anneal: method 2 (truncated Newton) runs rung by rung through vab_minimize");
  cudaSetDevice(ctx->device);
  std::string buf((size_t)Nbeta * sizeof(double), '\0');
  double* scales = reinterpret_cast<double*>(&buf[0]);
  for (int ib = 0; ib < Nbeta; ++ib) scales[ib] = pow(alpha, beta_host[ib]);
  int rc = lb_run(ctx, B, XP_dev, ldxp, scales, beta_host, Nbeta, opts, lo_dev, hi_dev, table_dev,
                  minpaths_dev, status_dev, nit_dev, nfev_dev, sink);
  if (rc != VAB_OK) return rc;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "vab_anneal");
  return VAB_OK;
}

}  // extern "C"
