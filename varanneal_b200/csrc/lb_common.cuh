// Pieces shared by the device minimisers (lbfgs.cu: L-BFGS-B / CG; tnc.cu: truncated Newton):
// the More'-Thuente line search (MINPACK-2 dcsrch / dcstep as used by L-BFGS-B's lnsrlb), the
// chunking of an n-vector over CTAs, and the fixed-order block reduction of the vector kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace vabmin {

constexpr int NT = 256;           // threads per CTA of the vector kernels
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double BIG = 1.0e10;    // stpmx of an unconstrained line search (lnsrlb)

// ---------------------------------------------------------------------------------------------
// More'-Thuente step (MINPACK-2 dcstep)
__device__ inline void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy,
                       double& stp, double fp, double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// dcsrch with task = 'START': initialise the search for the step already stored in s.stp
template <class PathState>
__device__ void dcsrch_start(PathState& s, double f, double g, double stpmax) {
  s.brackt = 0;
  s.stage = 1;
  s.finit = f;
  s.ginit = g;
  s.gtest = s.ls_ftol * g;
  s.width = stpmax - 0.0;
  s.width1 = s.width / 0.5;
  s.stx = 0.0; s.fx = f; s.gx = g;
  s.sty = 0.0; s.fy = f; s.gy = g;
  s.stmin = 0.0;
  s.stmax = s.stp + 4.0 * s.stp;
}

// dcsrch with task = 'FG': returns 0 = evaluate again at the new s.stp, 1 = convergence / warning
template <class PathState>
__device__ int dcsrch_step(PathState& s, double f, double g, double stpmin, double stpmax) {
  const double ftol = s.ls_ftol, gtol = s.ls_gtol, xtol = s.ls_xtol;
  const double ftest = s.finit + s.stp * s.gtest;
  if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
  int stop = 0;
  if (s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) stop = 1;           // rounding errors
  if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) stop = 1;               // xtol test
  if (s.stp == stpmax && f <= ftest && g <= s.gtest) stop = 1;                 // stp = stpmax
  if (s.stp == stpmin && (f > ftest || g >= s.gtest)) stop = 1;                // stp = stpmin
  if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) stop = 1;                    // convergence
  if (stop) return 1;
  (void)ftol;
  if (s.stage == 1 && f <= s.fx && f > ftest) {
    const double fm = f - s.stp * s.gtest;
    double fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
    const double gm = g - s.gtest;
    double gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
    dcstep(s.stx, fxm, gxm, s.sty, fym, gym, s.stp, fm, gm, s.brackt, s.stmin, s.stmax);
    s.fx = fxm + s.stx * s.gtest;
    s.fy = fym + s.sty * s.gtest;
    s.gx = gxm + s.gtest;
    s.gy = gym + s.gtest;
  } else {
    dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, s.stp, f, g, s.brackt, s.stmin, s.stmax);
  }
  if (s.brackt) {
    if (fabs(s.sty - s.stx) >= 0.66 * s.width1) s.stp = s.stx + 0.5 * (s.sty - s.stx);
    s.width1 = s.width;
    s.width = fabs(s.sty - s.stx);
  }
  if (s.brackt) {
    s.stmin = fmin(s.stx, s.sty);
    s.stmax = fmax(s.stx, s.sty);
  } else {
    s.stmin = s.stp + 1.1 * (s.stp - s.stx);
    s.stmax = s.stp + 4.0 * (s.stp - s.stx);
  }
  s.stp = fmax(s.stp, stpmin);
  s.stp = fmin(s.stp, stpmax);
  if ((s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) ||
      (s.brackt && s.stmax - s.stmin <= xtol * s.stmax))
    s.stp = s.stx;
  return 0;
}


struct Range { long long i0, i1; };
__device__ __forceinline__ Range chunk_range(long long n, int nchunk, int c) {
  long long len = (n + nchunk - 1) / nchunk;
  len = (len + 1) & ~1LL;
  Range r;
  r.i0 = (long long)c * len;
  r.i1 = r.i0 + len;
  if (r.i1 > n) r.i1 = n;
  if (r.i0 > n) r.i0 = n;
  return r;
}

// block reduction of NV per-thread values (sum, or max / min per entry), fixed order; result valid in thread 0..NV-1? no:
// every value k ends up in out[k] (written by one thread).  scratch: 8 * NT doubles.
enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
template <int NV>
__device__ void block_reduce(const double* v, const int* op, double* out, double* scratch) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int base = 0; base < NV; base += 8) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (base + k < NV) scratch[k * NT + tid] = v[base + k];
    __syncthreads();
    const int k = warp;                       // NT / 32 == 8 warps: warp k reduces entry base + k
    if (base + k < NV) {
      const int o = op[base + k];
      double acc = scratch[k * NT + lane];
      for (int t = lane + 32; t < NT; t += 32) {
        const double u = scratch[k * NT + t];
        acc = (o == RED_SUM) ? acc + u : (o == RED_MAX ? fmax(acc, u) : fmin(acc, u));
      }
      for (int sft = 16; sft > 0; sft >>= 1) {
        const double u = __shfl_down_sync(0xffffffffu, acc, sft);
        acc = (o == RED_SUM) ? acc + u : (o == RED_MAX ? fmax(acc, u) : fmin(acc, u));
      }
      if (lane == 0) out[base + k] = acc;
    }
  }
}

// chunks (CTAs) per path of the vector kernels: 8192 elements each for large batches, smaller
// chunks when the whole batch would not fill the machine (launch-latency-bound small problems)
inline int lb_nchunk(long long n, int B = 1 << 20) {
  static long long chunk = 0;
  if (chunk == 0) {
    const char* e = getenv("VAB_LBFGS_CHUNK");           // tuning knob
    chunk = (e && atoll(e) >= 512) ? atoll(e) : 8192;
  }
  long long c = (n + chunk - 1) / chunk;
  const long long want = (592 + B - 1) / B;            // ~4 CTAs per SM over the batch
  if (c < want) c = want;
  const long long cmax = (n + 511) / 512;              // at least 512 elements per chunk
  if (c > cmax) c = cmax;
  if (c < 1) c = 1;
  if (c > 4096) c = 4096;
  return (int)c;
}

}  // namespace vabmin
