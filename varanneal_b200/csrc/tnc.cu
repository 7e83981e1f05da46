// Device-resident truncated Newton: the method='TNC' seam of the reference
// (ADmin.min_tnc_scipy, _autodiffmin.py:121-143 -> scipy.optimize.minimize(method='TNC', jac=True)),
// for B independent paths at once, on the same cycle machinery as the L-BFGS-B driver
// (lbfgs.cu): every path is a small state machine in device memory, a cycle performs exactly one
// fused action+gradient evaluation per running path, the host only enqueues cycles and polls.
//
// What is kept of SciPy's TNC (Nash's truncated Newton): an outer Newton iteration whose step
// solves H p = -g approximately by conjugate gradients, Hessian-vector products taken by
// differencing the gradient (one evaluation each), at most maxCGit = max(1, min(50, n/2)) inner
// iterations, a line search along p, stopping on the gradient (gtol), on the decrease of f (ftol,
// off by default: f no longer changing), on the step length (xtol = sqrt(eps)) and on the
// evaluation budget; SciPy's return codes (0 local minimum, 1 f converged, 2 x converged,
// 3 evaluation limit, 4 line search failed).  What differs (documented deviation, like
// the bounds of the L-BFGS-B driver): the inner solve is plain CG truncated by the
// Dembo-Steihaug rule ||r|| <= min(0.5, sqrt||g||) ||g|| without Nash's diagonal/BFGS
// preconditioner and variable rescaling, the line search is More'-Thuente (dcsrch, ftol 1e-4,
// gtol 0.25 = TNC's eta) instead of Nash's getptc.  Same minima on well-conditioned problems
// (tests/test_gpu_ladder.py), different iterates.
// Bounds (the reference forwards them, _autodiffmin.py:133-134): an active-set version of the same
// iteration.  At every outer iterate the variables sitting on a bound with the gradient pushing
// outward are frozen: the inner CG runs in the subspace of the others (r, v, p are zero on the
// frozen ones), the line search is limited to the largest feasible step (a free variable that
// would leave the box at once is dropped from the step), trial points are clipped, and the
// stopping test uses the projected gradient.  Frozen variables are released as soon as their
// gradient turns inward at a later outer iterate.
//
// Per cycle and path (phase 1 = inner CG, phase 2 = line search):
//   trial     xt = x + delta v  |  x + stp p                                  (3 vector passes)
//   [f, g](xt)                                                                 (2)
//   dots      v.(gt - g), gt.p, max|gt|, gt.gt                                 (4 reads)
//   update    CG: p += alpha v, r -= alpha/delta (gt - g), r.r   |  accepted step: x <- xt, g <- gt,
//             r = v = -g, p = 0, x.x                                           (<= 6 reads, 5 writes)
//   direction CG: v = r + beta v, v.v   |  end of CG: g.p, p.p                  (<= 3 reads, 1 write)
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "lb_common.cuh"
#include "vab_ctx.h"

using namespace vabmin;

namespace {

enum { PH_FIRST = 0, PH_CG = 1, PH_LS = 2, PH_DONE = 3 };
enum { ACT_NONE = 0, ACT_INIT = 1, ACT_INIT_MOVE = 2, ACT_CG = 3, ACT_PV = 4, ACT_END = 5 };
enum { ACT2_NONE = 0, ACT2_DIR = 1, ACT2_ENDCG = 2 };

struct TnPath {
  int phase, act, act2;
  int iter, nfev, ncg, status, ifun;
  double f, fold, me, fe, sbgnrm, gg, rr, vv, xx, pp, alpha, beta, delta, tol2;
  double stp, gd, stpmx;
  double ls_ftol, ls_gtol, ls_xtol;
  int brackt, stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

struct TnOpts {
  int maxcg, maxls;
  long long maxfun;
  double ftol, pgtol;
};

}  // namespace

struct TncWork {
  int B = 0;
  double* vec = nullptr;        // XT, GT, G, R, V, P  each (B, ld)
  size_t vec_cap = 0;
  TnPath* st = nullptr;
  int st_cap = 0;
  int* act_eval = nullptr;
  double *ft = nullptr, *met = nullptr, *fet = nullptr;
  double* part = nullptr;
  size_t part_cap = 0;
  int* n_running_dev = nullptr;
  int* n_running_host = nullptr;
};

namespace {

__device__ __forceinline__ bool tn_frozen(double x, double g, double lo, double hi) {
  return (x <= lo && g > 0.0) || (x >= hi && g < 0.0);
}
__device__ __forceinline__ double tn_projg(double x, double g, double lo, double hi) {
  return g < 0.0 ? fmax(x - hi, g) : fmin(x - lo, g);
}
__global__ void tn_clip_kernel(double* X, long long ld, long long n, const double* lo, const double* hi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* x = X + (long long)blockIdx.y * ld;
  x[i] = fmin(fmax(x[i], lo[i]), hi[i]);
}

__global__ void tn_init_kernel(TnPath* st, int* act_eval, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  TnPath& s = st[b];
  memset(&s, 0, sizeof(TnPath));
  s.phase = PH_FIRST;
  s.status = 4;
  s.stpmx = BIG;
  s.ls_ftol = 1e-4; s.ls_gtol = 0.25; s.ls_xtol = 0.1;
  act_eval[b] = 1;
}

__global__ void __launch_bounds__(NT) tn_trial_kernel(double* __restrict__ XT, const double* __restrict__ X,
                                                      const double* __restrict__ V, const double* __restrict__ Pv,
                                                      long long ld, long long n, const TnPath* __restrict__ st,
                                                      int nchunk, const double* __restrict__ lo,
                                                      const double* __restrict__ hi) {
  const int b = blockIdx.y;
  const TnPath& s = st[b];
  if (s.phase == PH_DONE) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  const double a = (s.phase == PH_CG) ? s.delta : (s.phase == PH_LS ? s.stp : 0.0);
  const double* w = (s.phase == PH_CG) ? V + off : Pv + off;
  const bool clip = lo != nullptr && s.phase == PH_LS;      // line-search points stay inside the box
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    double xt = (s.phase == PH_FIRST) ? X[off + i] : fma(a, w[i], X[off + i]);
    if (clip) xt = fmin(fmax(xt, lo[i]), hi[i]);
    XT[off + i] = xt;
  }
}

// partials per chunk: [0] v.(gt - g)  [1] gt.p  [2] max|gt|  [3] gt.gt
__global__ void __launch_bounds__(NT) tn_dots_kernel(const double* __restrict__ GT, const double* __restrict__ G,
                                                     const double* __restrict__ V, const double* __restrict__ Pv,
                                                     long long ld, long long n, const TnPath* __restrict__ st,
                                                     int nchunk, double* __restrict__ part,
                                                     const double* __restrict__ XT, const double* __restrict__ lo,
                                                     const double* __restrict__ hi) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[4];
  const int b = blockIdx.y;
  const TnPath& s = st[b];
  if (s.phase == PH_DONE) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  const int ph = s.phase;
  for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
    const double gt = GT[off + i];
    if (ph == PH_CG) v[0] = fma(V[off + i], gt - G[off + i], v[0]);
    if (ph == PH_LS) v[1] = fma(gt, Pv[off + i], v[1]);
    if (lo != nullptr) {                              // bounded: projected gradient, |g|^2 over the free variables
      const double xt = XT[off + i];
      v[2] = fmax(v[2], fabs(tn_projg(xt, gt, lo[i], hi[i])));
      if (!tn_frozen(xt, gt, lo[i], hi[i])) v[3] = fma(gt, gt, v[3]);
    } else {
      v[2] = fmax(v[2], fabs(gt));
      v[3] = fma(gt, gt, v[3]);
    }
  }
  const int op[4] = {RED_SUM, RED_SUM, RED_MAX, RED_SUM};
  block_reduce<4>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x < 4) part[((long long)b * nchunk + blockIdx.x) * 4 + threadIdx.x] = res[threadIdx.x];
}

__global__ void tn_decide1_kernel(TnPath* st, int* act_eval, const double* ft, const double* met,
                                  const double* fet, const double* part, int nchunk, TnOpts o, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  TnPath& s = st[b];
  s.act = ACT_NONE;
  s.act2 = ACT2_NONE;
  if (s.phase == PH_DONE) return;
  double d0 = 0.0, d1 = 0.0, sbg = 0.0, gg = 0.0;
  for (int c = 0; c < nchunk; ++c) {
    const double* q = part + ((long long)b * nchunk + c) * 4;
    d0 += q[0]; d1 += q[1]; sbg = fmax(sbg, q[2]); gg += q[3];
  }
  s.nfev += 1;
  const double f = ft[b];
  if (s.phase == PH_FIRST) {
    s.f = f; s.me = met[b]; s.fe = fet[b]; s.sbgnrm = sbg; s.gg = gg;
    if (!isfinite(f)) { s.phase = PH_DONE; s.status = 4; act_eval[b] = 0; return; }
    if (sbg <= o.pgtol) { s.phase = PH_DONE; s.status = 0; act_eval[b] = 0; return; }
    s.act = ACT_INIT;
    return;
  }
  if (s.phase == PH_CG) {
    const double vHv = d0 / s.delta;
    if (!(vHv > 0.0) || !isfinite(vHv)) {            // non-positive curvature: stop the inner solve
      s.act = (s.ncg == 0) ? ACT_PV : ACT_END;
    } else {
      s.alpha = s.rr / vHv;
      s.act = ACT_CG;
    }
    if (s.nfev >= o.maxfun) { s.phase = PH_DONE; s.status = 3; act_eval[b] = 0; s.act = ACT_NONE; }
    return;
  }
  // line search
  const bool finite = isfinite(f) && isfinite(d1);
  int conv = 0;
  if (finite) conv = dcsrch_step(s, f, d1, 0.0, s.stpmx);
  if (finite && conv) {
    s.f = f; s.me = met[b]; s.fe = fet[b]; s.sbgnrm = sbg; s.gg = gg;
    s.iter += 1;
    s.act = ACT_INIT_MOVE;
    // stopping tests of TNC: projected gradient (0), decrease of f (1; with ftol = 0 this is "f no
    // longer changes", here: by less than a few ulps), step length against xtol = sqrt(eps) (2),
    // evaluation budget (3)
    const double fscale = fmax(fabs(s.fold), fmax(fabs(f), 1.0));
    if (sbg <= o.pgtol) { s.phase = PH_DONE; s.status = 0; }
    else if ((s.fold - f) <= fmax(o.ftol, 8.0 * EPSMCH) * fscale) { s.phase = PH_DONE; s.status = 1; }
    else if (s.stp * sqrt(s.pp) <= sqrt(EPSMCH) * fmax(1.0, sqrt(s.xx))) { s.phase = PH_DONE; s.status = 2; }
    else if (s.nfev >= o.maxfun) { s.phase = PH_DONE; s.status = 3; }
    if (s.phase == PH_DONE) act_eval[b] = 0;          // the update still moves x to the accepted point
    return;
  }
  s.ifun += 1;
  if (!finite) s.stp = 0.5 * s.stp;                  // step into a non-finite region: back off
  if (s.ifun >= o.maxls || s.nfev >= o.maxfun) {
    s.phase = PH_DONE; s.status = (s.nfev >= o.maxfun) ? 3 : 4; act_eval[b] = 0;
  }
}

// partial per chunk: [0] x.x (after an accepted step / the first evaluation) or r.r (CG step)
__global__ void __launch_bounds__(NT) tn_update_kernel(double* __restrict__ X, const double* __restrict__ XT,
                                                       double* __restrict__ G, const double* __restrict__ GT,
                                                       double* __restrict__ R, double* __restrict__ V,
                                                       double* __restrict__ Pv, long long ld, long long n,
                                                       const TnPath* __restrict__ st, int nchunk,
                                                       double* __restrict__ part, const double* __restrict__ lo,
                                                       const double* __restrict__ hi) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[1];
  const int b = blockIdx.y;
  const TnPath& s = st[b];
  const int act = s.act;
  if (act == ACT_NONE || act == ACT_END) return;
  const bool bnd = lo != nullptr;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double v[1] = {0.0};
  if (act == ACT_INIT || act == ACT_INIT_MOVE) {
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
      const double x = (act == ACT_INIT_MOVE) ? XT[off + i] : X[off + i];
      const double g = GT[off + i];
      if (act == ACT_INIT_MOVE) X[off + i] = x;
      G[off + i] = g;
      const double ng = (bnd && tn_frozen(x, g, lo[i], hi[i])) ? 0.0 : -g;   // frozen variables stay out of the inner solve
      R[off + i] = ng;
      V[off + i] = ng;
      Pv[off + i] = 0.0;
      v[0] = fma(x, x, v[0]);
    }
  } else if (act == ACT_CG) {
    const double a = s.alpha, ad = s.alpha / s.delta;
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
      const double vi = V[off + i];
      Pv[off + i] = fma(a, vi, Pv[off + i]);
      double ri = fma(-ad, GT[off + i] - G[off + i], R[off + i]);
      if (bnd && tn_frozen(X[off + i], G[off + i], lo[i], hi[i])) ri = 0.0;  // Z'(H v): rows of the frozen variables dropped
      R[off + i] = ri;
      v[0] = fma(ri, ri, v[0]);
    }
  } else {                                            // ACT_PV: steepest-descent-like step p = v
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) Pv[off + i] = V[off + i];
  }
  const int op[1] = {RED_SUM};
  block_reduce<1>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x == 0) part[(long long)b * nchunk + blockIdx.x] = res[0];
}

__device__ double tn_delta(double xx, double vv) {
  return sqrt(EPSMCH) * (1.0 + sqrt(xx)) / sqrt(vv);
}

__global__ void tn_decide2_kernel(TnPath* st, const double* part, int nchunk, TnOpts o, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  TnPath& s = st[b];
  const int act = s.act;
  if (act == ACT_NONE) return;
  double a = 0.0;
  if (act != ACT_END && act != ACT_PV)
    for (int c = 0; c < nchunk; ++c) a += part[(long long)b * nchunk + c];
  if (act == ACT_INIT || act == ACT_INIT_MOVE) {
    if (s.phase == PH_DONE) return;                   // converged on this step: x has been moved, nothing else
    s.xx = a;
    s.rr = s.gg; s.vv = s.gg;
    s.ncg = 0;
    const double gn = sqrt(s.gg);
    const double eta = fmin(0.5, sqrt(gn));
    s.tol2 = eta * eta * s.gg;
    s.delta = tn_delta(s.xx, s.vv);
    s.phase = PH_CG;
    return;
  }
  if (act == ACT_CG) {
    s.ncg += 1;
    if (a <= s.tol2 || s.ncg >= o.maxcg) { s.act2 = ACT2_ENDCG; return; }
    s.beta = a / s.rr;
    s.rr = a;
    s.act2 = ACT2_DIR;
    return;
  }
  s.act2 = ACT2_ENDCG;                                // ACT_PV, ACT_END
}

// partials per chunk: [0] v.v (new CG direction) | g.p (end of CG)   [1] p.p (end of CG)
__global__ void __launch_bounds__(NT) tn_dir_kernel(const double* __restrict__ G, const double* __restrict__ R,
                                                    double* __restrict__ V, double* __restrict__ Pv,
                                                    long long ld, long long n, const TnPath* __restrict__ st,
                                                    int nchunk, double* __restrict__ part, const double* __restrict__ X,
                                                    const double* __restrict__ lo, const double* __restrict__ hi) {
  __shared__ double scratch[8 * NT];
  __shared__ double res[3];
  const int b = blockIdx.y;
  const TnPath& s = st[b];
  const int act2 = s.act2;
  if (act2 == ACT2_NONE) return;
  const Range r = chunk_range(n, nchunk, blockIdx.x);
  const long long off = (long long)b * ld;
  double v[3] = {0.0, 0.0, BIG};
  if (act2 == ACT2_DIR) {
    const double be = s.beta;
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
      const double vi = fma(be, V[off + i], R[off + i]);
      V[off + i] = vi;
      v[0] = fma(vi, vi, v[0]);
    }
  } else {
    for (long long i = r.i0 + threadIdx.x; i < r.i1; i += NT) {
      double p = Pv[off + i];
      if (lo != nullptr && p != 0.0) {
        // largest step that keeps x + stp p inside the box; a variable that would leave at once is dropped
        const double x = X[off + i];
        if (p < 0.0 && lo[i] > -DBL_MAX) {
          const double a2 = lo[i] - x;
          if (a2 >= 0.0) { p = 0.0; Pv[off + i] = 0.0; }
          else if (p * v[2] < a2) v[2] = a2 / p;
        } else if (p > 0.0 && hi[i] < DBL_MAX) {
          const double a2 = hi[i] - x;
          if (a2 <= 0.0) { p = 0.0; Pv[off + i] = 0.0; }
          else if (p * v[2] > a2) v[2] = a2 / p;
        }
      }
      v[0] = fma(G[off + i], p, v[0]);
      v[1] = fma(p, p, v[1]);
    }
  }
  const int op[3] = {RED_SUM, RED_SUM, RED_MIN};
  block_reduce<3>(v, op, res, scratch);
  __syncthreads();
  if (threadIdx.x < 3) part[((long long)b * nchunk + blockIdx.x) * 3 + threadIdx.x] = res[threadIdx.x];
}

__global__ void tn_decide3_kernel(TnPath* st, int* act_eval, const double* part, int nchunk, TnOpts o, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  TnPath& s = st[b];
  const int act2 = s.act2;
  if (act2 == ACT2_NONE) return;
  double a0 = 0.0, a1 = 0.0, smax = BIG;
  for (int c = 0; c < nchunk; ++c) {
    a0 += part[((long long)b * nchunk + c) * 3 + 0];
    a1 += part[((long long)b * nchunk + c) * 3 + 1];
    smax = fmin(smax, part[((long long)b * nchunk + c) * 3 + 2]);
  }
  if (act2 == ACT2_DIR) {
    s.vv = a0;
    s.delta = tn_delta(s.xx, s.vv);
    return;
  }
  // end of the inner solve: line search along p from x
  const double gp = a0, pp = a1;
  if (!(gp < 0.0) || !(pp > 0.0)) {                   // not a descent direction (cannot happen with vHv > 0)
    s.phase = PH_DONE; s.status = 4; act_eval[b] = 0;
    return;
  }
  s.stpmx = smax;                                     // BIG without bounds
  s.pp = pp;
  s.stp = fmin((s.iter == 0) ? fmin(1.0, 1.0 / sqrt(pp)) : 1.0, smax);
  s.fold = s.f;
  s.ifun = 0;
  dcsrch_start(s, s.f, gp, s.stpmx);
  s.phase = PH_LS;
  (void)o;
}

__global__ void tn_count_kernel(const TnPath* st, int B, int* n_running) {
  int c = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) c += (st[b].phase == PH_DONE) ? 0 : 1;
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

__global__ void tn_export_kernel(const TnPath* st, int B, double* A, double* me, double* fe, int* status,
                                 int* nit, int* nfev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const TnPath& s = st[b];
  if (A) A[b] = s.f;
  if (me) me[b] = s.me;
  if (fe) fe[b] = s.fe;
  if (status) status[b] = s.status;
  if (nit) nit[b] = s.iter;
  if (nfev) nfev[b] = s.nfev;
}

#define TN_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return vab_cuda_fail(ctx, e_, #call);              \
  } while (0)

}  // namespace

void tnc_destroy(vab_ctx* ctx) {
  TncWork* w = ctx->tn;
  if (!w) return;
  cudaFree(w->vec); cudaFree(w->st); cudaFree(w->act_eval); cudaFree(w->ft); cudaFree(w->met);
  cudaFree(w->fet); cudaFree(w->part); cudaFree(w->n_running_dev);
  if (w->n_running_host) cudaFreeHost(w->n_running_host);
  delete w;
  ctx->tn = nullptr;
}

int tnc_minimize(vab_ctx* ctx, int B, double* XP, long long ld, double rf_scale, const vab_lbfgs_opts* uo,
                 const double* lo, const double* hi, double* A, double* me, double* fe, int* status, int* nit,
                 int* nfev) {
  const long long n = ctx->n_unknowns();
  if ((lo == nullptr) != (hi == nullptr)) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: give both bounds or none");
  if (n <= 0) return vab_fail(ctx, VAB_ERR_STATE, "minimize: no problem set on this context");
  if (B < 1 || !XP || ld < n || (ld & 1)) return vab_fail(ctx, VAB_ERR_INVALID, "minimize: bad batch / XP / ldxp");
  TnOpts o;
  const long long half = n / 2 < 50 ? n / 2 : 50;
  o.maxcg = (uo && uo->m > 0) ? uo->m : (int)(half > 1 ? half : 1);
  o.maxls = (uo && uo->maxls > 0) ? uo->maxls : 20;
  const long long dflt = 10 * n > 100 ? 10 * n : 100;
  o.maxfun = (uo && uo->maxfun > 0) ? uo->maxfun : dflt;
  if (uo && uo->maxiter > 0 && uo->maxiter < o.maxfun) o.maxfun = uo->maxiter;   // SciPy: maxiter bounds the evaluations
  o.ftol = uo ? uo->ftol : 0.0;
  o.pgtol = uo ? uo->pgtol : 1.2207e-6;
  if (!ctx->tn) ctx->tn = new TncWork();
  TncWork* w = ctx->tn;
  const size_t vs = (size_t)B * (size_t)ld;
  int rc = vab_reserve(ctx, &w->vec, &w->vec_cap, 6 * vs);
  if (rc != VAB_OK) return rc;
  if (B > w->st_cap) {
    cudaFree(w->st); cudaFree(w->act_eval); cudaFree(w->ft); cudaFree(w->met); cudaFree(w->fet);
    w->st = nullptr; w->act_eval = nullptr; w->ft = w->met = w->fet = nullptr;
    TN_CUDA(cudaMalloc((void**)&w->st, sizeof(TnPath) * B));
    TN_CUDA(cudaMalloc((void**)&w->act_eval, sizeof(int) * B));
    TN_CUDA(cudaMalloc((void**)&w->ft, sizeof(double) * B));
    TN_CUDA(cudaMalloc((void**)&w->met, sizeof(double) * B));
    TN_CUDA(cudaMalloc((void**)&w->fet, sizeof(double) * B));
    w->st_cap = B;
  }
  const int nchunk = lb_nchunk(n, B);
  rc = vab_reserve(ctx, &w->part, &w->part_cap, (size_t)B * nchunk * 4);
  if (rc != VAB_OK) return rc;
  if (!w->n_running_dev) TN_CUDA(cudaMalloc((void**)&w->n_running_dev, sizeof(int)));
  if (!w->n_running_host) TN_CUDA(cudaMallocHost((void**)&w->n_running_host, sizeof(int)));
  cudaStream_t st = ctx->stream;
  double* XT = w->vec;
  double* GT = w->vec + vs;
  double* G = w->vec + 2 * vs;
  double* R = w->vec + 3 * vs;
  double* V = w->vec + 4 * vs;
  double* Pv = w->vec + 5 * vs;
  const dim3 vgrid(nchunk, B);
  const int tb = (B + 127) / 128;
  tn_init_kernel<<<tb, 128, 0, st>>>(w->st, w->act_eval, B);
  if (lo) tn_clip_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, st>>>(XP, ld, n, lo, hi);
  TN_CUDA(cudaMemsetAsync(w->vec, 0, 6 * vs * sizeof(double), st));
  const int poll = (n * (long long)B < (1LL << 22)) ? 32 : 8;
  long long cycles = 0;
  const long long max_cycles = o.maxfun + 64;
  while (true) {
    for (int c = 0; c < poll; ++c) {
      tn_trial_kernel<<<vgrid, NT, 0, st>>>(XT, XP, V, Pv, ld, n, w->st, nchunk, lo, hi);
      rc = vab_eval(ctx, B, XT, ld, rf_scale, nullptr, w->act_eval, w->ft, w->met, w->fet, GT, ld);
      if (rc != VAB_OK) return rc;
      tn_dots_kernel<<<vgrid, NT, 0, st>>>(GT, G, V, Pv, ld, n, w->st, nchunk, w->part, XT, lo, hi);
      tn_decide1_kernel<<<tb, 128, 0, st>>>(w->st, w->act_eval, w->ft, w->met, w->fet, w->part, nchunk, o, B);
      tn_update_kernel<<<vgrid, NT, 0, st>>>(XP, XT, G, GT, R, V, Pv, ld, n, w->st, nchunk, w->part, lo, hi);
      tn_decide2_kernel<<<tb, 128, 0, st>>>(w->st, w->part, nchunk, o, B);
      tn_dir_kernel<<<vgrid, NT, 0, st>>>(G, R, V, Pv, ld, n, w->st, nchunk, w->part, XP, lo, hi);
      tn_decide3_kernel<<<tb, 128, 0, st>>>(w->st, w->act_eval, w->part, nchunk, o, B);
      ctx->launches += 7;
    }
    cycles += poll;
    tn_count_kernel<<<1, 256, 0, st>>>(w->st, B, w->n_running_dev);
    ctx->launches += 1;
    TN_CUDA(cudaMemcpyAsync(w->n_running_host, w->n_running_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    TN_CUDA(cudaGetLastError());
    if (*w->n_running_host == 0) break;
    if (cycles > max_cycles) return vab_fail(ctx, VAB_ERR_STATE, "minimize (TNC): cycle limit exceeded (internal error)");
  }
  tn_export_kernel<<<tb, 128, 0, st>>>(w->st, B, A, me, fe, status, nit, nfev);
  ctx->launches += 1;
  TN_CUDA(cudaStreamSynchronize(st));
  return VAB_OK;
}
