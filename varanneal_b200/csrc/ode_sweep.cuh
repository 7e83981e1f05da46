// Fused ODE action + adjoint gradient: the register "sweep" kernels (device code only).
//
// Replaces, for one RF value and a batch of paths, what the reference does with an ADOL-C tape of
// va_ode.Annealer.A_gaussian (va_ode.py:130-234 taped by _autodiffmin.py:32-49 and replayed by
// :57-58): one pass over X computes the measurement error, the model error under the chosen
// discretisation, and the analytic gradient.
//
// Work decomposition
//   strip   = C consecutive state components owned by one lane (C = 4 -> one 32-byte load per row).
//   group   = GW adjacent lanes of a warp covering one *window* of a row: either the whole row
//             (rows of <= 32 strips: halos wrap around inside the group, no redundant lanes) or
//             WS output strips plus NHL halo lanes either side (wider rows, e.g. D = 1000).
//   unit    = (path b, segment sg of Tseg time rows, window w); one group walks one unit forward
//             in time.  GPW groups share a warp and run in lockstep; warps never synchronise with
//             each other -- there is no __syncthreads and no shared memory in the time loop.
//   Each lane streams its own strip straight from HBM into registers (16-byte loads, software
//   prefetch PD steps ahead), keeps the few rows the time stencil still needs in registers, and
//   gets the +-2-component halo of the periodic Lorenz96 stencil from its neighbour lanes with
//   warp shuffles (once for x, once for the adjoint seed v).  Gradient rows are written once, as
//   soon as their seed is complete.  Per-unit partial sums (me, fe, parameter gradient) go to a
//   partials buffer that ode_finalize reduces in a fixed order: bit-reproducible, no atomics.
#pragma once
#include "ode_models.cuh"
#include "ode_params.h"
#include "vab_hd.h"

#define VAB_FULL 0xffffffffu

template <class M>
struct SLane {
  static constexpr int C = M::C, H = M::H, W = M::C + 2 * M::H, NPM = M::NPM;
  int j, gbase, unit, b, r0, r1, i0, N, D;
  bool act, out;
  int srcM1, srcM2, srcP1, srcP2;      // lanes holding the strips at offsets -1, -2, +1, +2
  const double* xpath;
  double* gpath;
  double p[NPM];
  double wob[C];                       // 2 cm RM of the own components (0 = unobserved)
  double rfs, rsc;                     // RF (scalar form) and the scale of an RF0 array, of this path
  double me_acc, fe_acc, pacc[NPM];

  __device__ __forceinline__ void init(const OdeParams& P) {
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int g = lane / P.GW;
    j = lane - g * P.GW;
    gbase = g * P.GW;
    const bool ingroup = g < P.GPW;
    unit = gwarp * P.GPW + g;
    const bool uok = ingroup && unit < P.nunits;
    int w = 0, sg = 0;
    b = 0;
    if (uok) {
      w = unit % P.nwin;
      const int t = unit / P.nwin;
      sg = t % P.nseg;
      b = t / P.nseg;
    }
    N = P.N;
    D = P.D;
    act = uok && (P.active == nullptr || __ldg(P.active + b) != 0);
    r0 = sg * P.Tseg;
    r1 = min(r0 + P.Tseg, P.N);
    if (!uok) { r0 = 0; r1 = 0; }
    const int TPR = P.TPR;
    int st;
    if (P.NHL == 0) {                       // whole row in the group, periodic wrap inside it
      st = j;
      out = act && (j < TPR);
      const int gw = P.GW;
      srcM1 = gbase + (j + gw - 1) % gw;
      srcM2 = gbase + (j + 2 * gw - 2) % gw;
      srcP1 = gbase + (j + 1) % gw;
      srcP2 = gbase + (j + 2) % gw;
    } else {                                // window of WS strips + NHL halo lanes either side
      const int rel = w * P.WS - P.NHL + j;
      st = ((rel % TPR) + TPR) % TPR;
      out = act && j >= P.NHL && j < P.NHL + P.WS && (w * P.WS + j - P.NHL) < TPR;
      srcM1 = gbase + max(j - 1, 0);
      srcM2 = gbase + max(j - 2, 0);
      srcP1 = gbase + min(j + 1, P.GW - 1);
      srcP2 = gbase + min(j + 2, P.GW - 1);
    }
    if (!ingroup) { srcM1 = srcM2 = srcP1 = srcP2 = lane; }
    i0 = st * C;
    xpath = P.XP + (long long)b * P.ldxp;
    gpath = P.G ? P.G + (long long)b * P.ldg : nullptr;
    const long long nX = (long long)P.N * P.D;
#pragma unroll
    for (int k = 0; k < NPM; ++k) {
      double v = 0.0;
      if (act) {
        const int e = __ldg(P.pmap + k);
        v = (e >= 0) ? __ldg(xpath + nX + e) : __ldg(P.pfix + (long long)b * P.pfix_stride + k);
      }
      p[k] = v;
      pacc[k] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) wob[c] = act ? __ldg(P.wobs + i0 + c) : 0.0;
    rsc = (P.rf_path != nullptr) ? __ldg(P.rf_path + b) : P.rf_scale;
    rfs = (P.rf_path != nullptr) ? P.rf0 * rsc : P.rf_scalar;
    me_acc = 0.0;
    fe_acc = 0.0;
  }

  __device__ __forceinline__ bool rowvalid(int r) const { return act && r >= 0 && r < N; }
  __device__ __forceinline__ bool owned(int r) const { return out && r >= r0 && r < r1; }
  __device__ __forceinline__ bool unit_owns(int r) const { return act && r >= r0 && r < r1; }

  // ---- parameter time series (P.ptime; va_ode.py:170-188) -------------------------------------
  // parameters of time row `row` (zeros when the row is not needed)
  __device__ __forceinline__ void loadp(const OdeParams& P, int row, bool need, double* q) const {
    const long long nX = (long long)N * D;
#pragma unroll
    for (int k = 0; k < NPM; ++k) {
      double v = 0.0;
      if (need) {
        const int e = __ldg(P.pmap + k);
        v = (e >= 0) ? __ldg(xpath + nX + (long long)row * P.NPest + e)
                     : __ldg(P.pfix + (long long)b * P.pfix_stride + (long long)row * P.NP + k);
      }
      q[k] = v;
    }
  }
  // gradient of the parameters of row `row`: minus the sum of the lanes' (df/dp)^T v over the
  // group, in lane order (bit-reproducible).  Must be executed by all 32 lanes.
  __device__ __forceinline__ void storep(const OdeParams& P, int row, const double* pl) const {
    const bool wr = unit_owns(row) && j == 0 && gpath != nullptr;
    const long long nX = (long long)N * D;
#pragma unroll
    for (int k = 0; k < NPM; ++k) {
      double acc = 0.0;
      for (int s = 0; s < P.GW; ++s) acc += __shfl_sync(VAB_FULL, pl[k], gbase + s);
      if (wr) {
        const int e = __ldg(P.pmap + k);
        if (e >= 0) gpath[nX + (long long)row * P.NPest + e] = -acc;
      }
    }
  }

  // own strip of row `row` (zeros when the row is not needed)
  __device__ __forceinline__ void load(int row, bool need, double* dst) const {
    if (need) {
      vab_load_strip<C>(xpath + (long long)row * D + i0, dst);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) dst[c] = 0.0;
    }
  }
  // full[H..H+C) holds the own values; fills the halo from the neighbour lanes.  Must be executed
  // by all 32 lanes.
  __device__ __forceinline__ void halo(double* full) const {
    if constexpr (H > 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const int lo = (H - h + C - 1) / C;            // lanes to the left (1 or 2)
        const int cl = C * lo - (H - h);               // component inside that lane's strip
        const int ro = 1 + h / C;                      // lanes to the right
        const int cr = h % C;
        const double a = __shfl_sync(VAB_FULL, full[H + cl], lo == 1 ? srcM1 : srcM2);
        const double bb = __shfl_sync(VAB_FULL, full[H + cr], ro == 1 ? srcP1 : srcP2);
        full[h] = a;
        full[H + C + h] = bb;
      }
    }
  }
  __device__ __forceinline__ const double* stim_row(const OdeParams& P, int row) const {
    return (M::NSTIM > 0 && P.stim != nullptr) ? P.stim + (long long)row * P.S : nullptr;
  }
  __device__ __forceinline__ double wgt(const OdeParams& P, int row, int c) const {
    return P.rf_arr ? __ldg(P.rf_arr + (long long)row * D + i0 + c) * rsc : rfs;
  }
  // matrix RF (P.rf_mat, va_ode.py:211-223): out = scale (R_row + R_row') e for the own components;
  // e is exchanged over the group's lanes strip by strip.  e' R e = 1/2 e . out.  Whole rows inside
  // one lane group only.  Must be executed by all 32 lanes.
  __device__ __forceinline__ void matrf(const OdeParams& P, int row, bool valid, const double* e, double* res) const {
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] = 0.0;
    const double* R = P.rf_mat + (long long)row * D * D;
    const bool ld = valid && out;
    for (int s = 0; s < P.TPR; ++s) {
      double es[C];
#pragma unroll
      for (int k = 0; k < C; ++k) es[k] = __shfl_sync(VAB_FULL, e[k], gbase + s);
      if (ld) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const long long i = i0 + c;
#pragma unroll
          for (int k = 0; k < C; ++k) {
            const long long jj = (long long)s * C + k;
            if (i < D && jj < D) res[c] = fma(__ldg(R + i * D + jj) + __ldg(R + jj * D + i), es[k], res[c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] *= rsc;
  }
  // measurement term of row r (va_ode.py:138-158): adds to the direct gradient and to me_acc
  // (me_acc collects sum 2 cm RM diff^2 = 2 me)
  __device__ __forceinline__ void measure(const OdeParams& P, int r, const double* xown, double* dir) {
    if (P.L == 0 || (P.nskip != 1 && (r % P.nskip) != 0)) return;
    const long long o = (long long)((P.nskip == 1) ? r : r / P.nskip) * D + i0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double w = P.rmd ? __ldg(P.rmd + o + c) : wob[c];
      if (w != 0.0) {
        const double diff = xown[c] - __ldg(P.Y + o + c);
        const double wd = w * diff;
        me_acc = fma(wd, diff, me_acc);
        dir[c] += wd;
      }
    }
  }
  __device__ __forceinline__ void store(int r, const double* dir, const double* jt) const {
    if (gpath == nullptr) return;
    double g[C];
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = dir[c] - jt[c];
    vab_store_strip<C>(gpath + (long long)r * D + i0, g);
  }
  // per-unit partials: lane partials -> shared memory -> fixed-order sum by the group's lanes
  __device__ __forceinline__ void finish(const OdeParams& P, double* smem, double psign) const {
    double* mine = smem + (long long)threadIdx.x * P.K;
    mine[0] = 0.5 * me_acc;
    mine[1] = fe_acc * P.cf;
#pragma unroll
    for (int k = 0; k < NPM; ++k) mine[2 + k] = psign * pacc[k];
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const bool ingroup = (lane / P.GW) < P.GPW;
    if (ingroup && unit < P.nunits) {
      const double* grp = smem + (long long)(threadIdx.x - j) * P.K;
      for (int k = j; k < P.K; k += P.GW) {
        double acc = 0.0;
        for (int s = 0; s < P.GW; ++s) acc += grp[(long long)s * P.K + k];
        P.partials[(long long)unit * P.K + k] = acc;
      }
    }
  }
};

// --------------------------------------------------------------------------------------------
// euler / trapezoid / forwardmap (va_ode.py:341-380, 439-454):
//   e_m = x_{m+1} - AL x_m - (CA f_m + CB f_{m+1})
//   g_r = [lam_{r-1} - AL lam_r] + meas_r - J^T(x_r) (CB lam_{r-1} + CA lam_r),  lam = 2 cf w e
// PT: the parameters are a time series -- row m's parameters are loaded with row m, the adjoint call
// of a row yields that row's parameter gradient (one call per row: both uses of f(x_r, p_r) share
// the seed V_r), reduced over the group's lanes and written next to the row's state gradient.
template <class M, int DISC, int PD, int MINB, bool PT = false>
__global__ void __launch_bounds__(128, MINB) sweep_twopoint_kernel(const __grid_constant__ OdeParams P) {
  using LN = SLane<M>;
  constexpr int C = LN::C, H = LN::H, W = LN::W, NPM = LN::NPM;
  extern __shared__ double smem[];
  LN L;
  vab_pdl_trigger();
  L.init(P);
  const double dt = P.dt;
  const double ca = (DISC == DISC_EULER) ? dt : (DISC == DISC_TRAPEZOID ? 0.5 * dt : 1.0);
  const double cb = (DISC == DISC_TRAPEZOID) ? 0.5 * dt : 0.0;
  const double al = (DISC == DISC_FORWARDMAP) ? 0.0 : 1.0;
  const double cf2 = 2.0 * P.cf;
  double X1[W], F1[C], lamp[C], pf[PD][C];
  [[maybe_unused]] double P1[NPM];                    // PT: parameters of row m-1
  if constexpr (PT) {
#pragma unroll
    for (int k = 0; k < NPM; ++k) P1[k] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < W; ++c) X1[c] = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) { F1[c] = 0.0; lamp[c] = 0.0; }
  const int m0 = L.r0 - 1;
  auto need = [&](int r) { return L.rowvalid(r) && r >= L.r0 - 1 && r <= L.r1; };
#pragma unroll
  for (int u = 0; u < PD; ++u) L.load(m0 + u, need(m0 + u), pf[u]);
  const int nsteps = ((P.Tseg + 2 + PD - 1) / PD) * PD;
  for (int s0 = 0; s0 < nsteps; s0 += PD) {
#pragma unroll
    for (int u = 0; u < PD; ++u) {
      const int t = s0 + u, m = m0 + t;
      double Xm[W], Fm[C];
#pragma unroll
      for (int c = 0; c < C; ++c) Xm[H + c] = pf[u][c];
      L.load(m + PD, need(m + PD), pf[u]);
      L.halo(Xm);
      const bool vm = need(m);
      [[maybe_unused]] double Pm[NPM];
      if constexpr (PT) L.loadp(P, m, vm, Pm);
      if (vm) {
        if constexpr (PT) M::f(Xm, Pm, L.stim_row(P, m), Fm);
        else M::f(Xm, L.p, L.stim_row(P, m), Fm);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) Fm[c] = 0.0;
      }
      const bool ve = vm && t >= 1 && m >= 1;
      const bool own = L.owned(m - 1);
      double lam[C], V[W], d[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        double l = 0.0;
        if (ve) {
          const double w = L.wgt(P, m - 1, c);
          const double e = Xm[H + c] - al * X1[H + c] - (ca * F1[c] + cb * Fm[c]);
          l = cf2 * w * e;
          if (own) L.fe_acc = fma(w * e, e, L.fe_acc);
        }
        lam[c] = l;
        V[H + c] = cb * lamp[c] + ca * l;
        d[c] = lamp[c] - al * l;
      }
      L.halo(V);
      if constexpr (PT) {
        double pl[NPM];
#pragma unroll
        for (int k = 0; k < NPM; ++k) pl[k] = 0.0;
        if (own) {
          double jt[C];
          L.measure(P, m - 1, X1 + H, d);
          M::adj(X1, V, P1, jt, pl);
          L.store(m - 1, d, jt);
        }
        L.storep(P, m - 1, pl);
#pragma unroll
        for (int k = 0; k < NPM; ++k) P1[k] = Pm[k];
      } else {
        if (own) {
          double jt[C];
          L.measure(P, m - 1, X1 + H, d);
          M::adj(X1, V, L.p, jt, L.pacc);
          L.store(m - 1, d, jt);
        }
      }
#pragma unroll
      for (int c = 0; c < W; ++c) X1[c] = Xm[c];
#pragma unroll
      for (int c = 0; c < C; ++c) { F1[c] = Fm[c]; lamp[c] = lam[c]; }
    }
  }
  vab_pdl_wait();
  L.finish(P, smem, -1.0);
}

// --------------------------------------------------------------------------------------------
// Simpson-Hermite (va_ode.py:404-437 + :192-195).  Pair k = rows (a, b, c) = (2k, 2k+1, 2k+2):
//   e1 = x_c - x_a - dt/3 (f_a + 4 f_b + f_c),  e2 = x_b - (x_a + x_c)/2 - dt/4 (f_a - f_c)
// One step per pair; rows a and b get their gradient in the step of their pair, row c's partial
// seed is carried into the next pair (where it is row a).  Segments start on even rows and the
// walk begins one pair early so that the c-part of row r0 is available.
// MRF: RF is one (D, D) matrix per residual row (P.rf_mat): the seeds are (R + R') e instead of 2 w e.
template <class M, int PD, int MINB, bool PT = false, bool MRF = false>
__global__ void __launch_bounds__(128, MINB) sweep_simpson_kernel(const __grid_constant__ OdeParams P) {
  using LN = SLane<M>;
  constexpr int C = LN::C, H = LN::H, W = LN::W, NPM = LN::NPM;
  extern __shared__ double smem[];
  LN L;
  vab_pdl_trigger();
  L.init(P);
  const double dt = P.dt;
  const double cf2 = 2.0 * P.cf;
  const double dt3 = dt / 3.0, dt4 = dt / 4.0;
  auto need = [&](int r) { return L.rowvalid(r) && r >= L.r0 - 2 && r <= L.r1; };
  double Xa[W], Fa[C], vcp[C], dcp[C], pf[PD][2][C];
#pragma unroll
  for (int c = 0; c < C; ++c) { vcp[c] = 0.0; dcp[c] = 0.0; }
  const int a0 = L.r0 - 2;
  [[maybe_unused]] double Pa[NPM];                    // PT: parameters of row a
  {
    double own[C];
    L.load(a0, need(a0), own);
#pragma unroll
    for (int c = 0; c < C; ++c) Xa[H + c] = own[c];
    L.halo(Xa);
    if constexpr (PT) L.loadp(P, a0, need(a0), Pa);
    if (need(a0)) {
      if constexpr (PT) M::f(Xa, Pa, L.stim_row(P, a0), Fa);
      else M::f(Xa, L.p, L.stim_row(P, a0), Fa);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) Fa[c] = 0.0;
    }
  }
#pragma unroll
  for (int u = 0; u < PD; ++u) {
    L.load(a0 + 2 * u + 1, need(a0 + 2 * u + 1), pf[u][0]);
    L.load(a0 + 2 * u + 2, need(a0 + 2 * u + 2), pf[u][1]);
  }
  const int nsteps = ((P.Tseg / 2 + 1 + PD - 1) / PD) * PD;
  for (int s0 = 0; s0 < nsteps; s0 += PD) {
#pragma unroll
    for (int u = 0; u < PD; ++u) {
      const int a = a0 + 2 * (s0 + u), bq = a + 1, c2 = a + 2;
      double Xb[W], Xc[W], Fb[C], Fc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { Xb[H + c] = pf[u][0][c]; Xc[H + c] = pf[u][1][c]; }
      L.load(bq + 2 * PD, need(bq + 2 * PD), pf[u][0]);
      L.load(c2 + 2 * PD, need(c2 + 2 * PD), pf[u][1]);
      L.halo(Xb);
      L.halo(Xc);
      const bool vc = need(c2);
      const bool vp = need(a) && vc;                 // the pair exists (a >= 0, c <= N-1)
      [[maybe_unused]] double Pb[NPM], Pc[NPM];
      if constexpr (PT) {
        L.loadp(P, bq, vp, Pb);
        L.loadp(P, c2, vc, Pc);
      }
      if (vp) {
        if constexpr (PT) M::f(Xb, Pb, L.stim_row(P, bq), Fb);
        else M::f(Xb, L.p, L.stim_row(P, bq), Fb);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) Fb[c] = 0.0;
      }
      if (vc) {
        if constexpr (PT) M::f(Xc, Pc, L.stim_row(P, c2), Fc);
        else M::f(Xc, L.p, L.stim_row(P, c2), Fc);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) Fc[c] = 0.0;
      }
      const bool own_p = L.owned(bq);
      double Va[W], Vb[W], da[C], db[C], vcn[C], dcn[C];
      [[maybe_unused]] double m1[C], m2[C];
      if constexpr (MRF) {
        double e1[C], e2[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          e1[c] = vp ? Xc[H + c] - Xa[H + c] - dt3 * (Fa[c] + 4.0 * Fb[c] + Fc[c]) : 0.0;
          e2[c] = vp ? Xb[H + c] - 0.5 * (Xa[H + c] + Xc[H + c]) - dt4 * (Fa[c] - Fc[c]) : 0.0;
        }
        L.matrf(P, a, vp, e1, m1);
        L.matrf(P, bq, vp, e2, m2);
        if (vp && own_p) {
#pragma unroll
          for (int c = 0; c < C; ++c) L.fe_acc = fma(0.5 * e1[c], m1[c], fma(0.5 * e2[c], m2[c], L.fe_acc));
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        double l1 = 0.0, l2 = 0.0;
        if constexpr (MRF) {
          l1 = 0.5 * cf2 * m1[c];
          l2 = 0.5 * cf2 * m2[c];
        } else if (vp) {
          const double w1 = L.wgt(P, a, c), w2 = L.wgt(P, bq, c);
          const double e1 = Xc[H + c] - Xa[H + c] - dt3 * (Fa[c] + 4.0 * Fb[c] + Fc[c]);
          const double e2 = Xb[H + c] - 0.5 * (Xa[H + c] + Xc[H + c]) - dt4 * (Fa[c] - Fc[c]);
          l1 = cf2 * w1 * e1;
          l2 = cf2 * w2 * e2;
          if (own_p) L.fe_acc = fma(w1 * e1, e1, fma(w2 * e2, e2, L.fe_acc));
        }
        Vb[H + c] = (4.0 * dt3) * l1;
        db[c] = l2;
        Va[H + c] = vcp[c] + dt3 * l1 + dt4 * l2;
        da[c] = dcp[c] - l1 - 0.5 * l2;
        vcn[c] = dt3 * l1 - dt4 * l2;
        dcn[c] = l1 - 0.5 * l2;
      }
      L.halo(Va);
      L.halo(Vb);
      if constexpr (PT) {
        double pl[NPM];
#pragma unroll
        for (int k = 0; k < NPM; ++k) pl[k] = 0.0;
        if (L.owned(a)) {
          double jt[C];
          L.measure(P, a, Xa + H, da);
          M::adj(Xa, Va, Pa, jt, pl);
          L.store(a, da, jt);
        }
        L.storep(P, a, pl);
#pragma unroll
        for (int k = 0; k < NPM; ++k) pl[k] = 0.0;
        if (own_p) {
          double jt[C];
          L.measure(P, bq, Xb + H, db);
          M::adj(Xb, Vb, Pb, jt, pl);
          L.store(bq, db, jt);
        }
        L.storep(P, bq, pl);
#pragma unroll
        for (int k = 0; k < NPM; ++k) Pa[k] = Pc[k];
      } else {
        if (L.owned(a)) {
          double jt[C];
          L.measure(P, a, Xa + H, da);
          M::adj(Xa, Va, L.p, jt, L.pacc);
          L.store(a, da, jt);
        }
        if (own_p) {
          double jt[C];
          L.measure(P, bq, Xb + H, db);
          M::adj(Xb, Vb, L.p, jt, L.pacc);
          L.store(bq, db, jt);
        }
      }
#pragma unroll
      for (int c = 0; c < W; ++c) Xa[c] = Xc[c];
#pragma unroll
      for (int c = 0; c < C; ++c) { Fa[c] = Fc[c]; vcp[c] = vcn[c]; dcp[c] = dcn[c]; }
    }
  }
  vab_pdl_wait();
  L.finish(P, smem, -1.0);
}

// --------------------------------------------------------------------------------------------
// RK4 (extension; intent at va_ode.py:382-402):  e_m = x_{m+1} - x_m - dt/6 (k1 + 2k2 + 2k3 + k4).
// Per row: three forward stage exchanges (y2, y3, y4), then the discrete adjoint walks the stages
// backwards with one seed exchange per stage.  g_m = lam_{m-1} + xbar_m + meas_m.
template <class M, int PD, int MINB>
__global__ void __launch_bounds__(128, MINB) sweep_rk4_kernel(const __grid_constant__ OdeParams P) {
  using LN = SLane<M>;
  constexpr int C = LN::C, H = LN::H, W = LN::W, NPM = LN::NPM;
  extern __shared__ double smem[];
  LN L;
  vab_pdl_trigger();
  L.init(P);
  const double dt = P.dt;
  const double cf2 = 2.0 * P.cf;
  auto need = [&](int r) { return L.rowvalid(r) && r >= L.r0 - 1 && r <= L.r1; };
  double x0[C], lamp[C], pf[PD][C];
#pragma unroll
  for (int c = 0; c < C; ++c) lamp[c] = 0.0;
  const int m0 = L.r0 - 1;
  L.load(m0, need(m0), x0);
#pragma unroll
  for (int u = 0; u < PD; ++u) L.load(m0 + 1 + u, need(m0 + 1 + u), pf[u]);
  const int nsteps = ((P.Tseg + 1 + PD - 1) / PD) * PD;
  for (int s0 = 0; s0 < nsteps; s0 += PD) {
#pragma unroll
    for (int u = 0; u < PD; ++u) {
      const int m = m0 + s0 + u;
      double xn[C];                                   // row m+1 (own strip)
#pragma unroll
      for (int c = 0; c < C; ++c) xn[c] = pf[u][c];
      L.load(m + 1 + PD, need(m + 1 + PD), pf[u]);
      const bool vr = need(m) && need(m + 1);         // residual m exists
      const bool own = L.owned(m);
      const bool va = vr && own;
      double Y1[W], Y2[W], Y3[W], Y4[W], k[C], ks[C];
#pragma unroll
      for (int c = 0; c < C; ++c) Y1[H + c] = x0[c];
      L.halo(Y1);
      if (vr) { M::f(Y1, L.p, nullptr, k); }
      else {
#pragma unroll
        for (int c = 0; c < C; ++c) k[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] = k[c]; Y2[H + c] = x0[c] + 0.5 * dt * k[c]; }
      L.halo(Y2);
      if (vr) M::f(Y2, L.p, nullptr, k);
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] += 2.0 * k[c]; Y3[H + c] = x0[c] + 0.5 * dt * k[c]; }
      L.halo(Y3);
      if (vr) M::f(Y3, L.p, nullptr, k);
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] += 2.0 * k[c]; Y4[H + c] = x0[c] + dt * k[c]; }
      L.halo(Y4);
      if (vr) M::f(Y4, L.p, nullptr, k);
      double lam[C], xb[C], KB[W], jt[C], pl[NPM];
#pragma unroll
      for (int q = 0; q < NPM; ++q) pl[q] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        double l = 0.0;
        if (vr) {
          const double w = L.wgt(P, m, c);
          const double e = xn[c] - x0[c] - (dt / 6.0) * (ks[c] + k[c]);
          l = cf2 * w * e;
          if (own) L.fe_acc = fma(w * e, e, L.fe_acc);
        }
        lam[c] = l;
        xb[c] = -l;
        KB[H + c] = -(dt / 6.0) * l;
      }
      L.halo(KB);
      M::adj(Y4, KB, L.p, jt, pl);
#pragma unroll
      for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = -(dt / 3.0) * lam[c] + dt * jt[c]; }
      L.halo(KB);
      M::adj(Y3, KB, L.p, jt, pl);
#pragma unroll
      for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = -(dt / 3.0) * lam[c] + 0.5 * dt * jt[c]; }
      L.halo(KB);
      M::adj(Y2, KB, L.p, jt, pl);
#pragma unroll
      for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = -(dt / 6.0) * lam[c] + 0.5 * dt * jt[c]; }
      L.halo(KB);
      M::adj(Y1, KB, L.p, jt, pl);
      if (va) {
#pragma unroll
        for (int q = 0; q < NPM; ++q) L.pacc[q] += pl[q];
      }
      if (own) {
        double dir[C], z[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { dir[c] = lamp[c] + (vr ? xb[c] + jt[c] : 0.0); z[c] = 0.0; }
        L.measure(P, m, x0, dir);
        L.store(m, dir, z);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) { lamp[c] = lam[c]; x0[c] = xn[c]; }
    }
  }
  vab_pdl_wait();
  L.finish(P, smem, 1.0);
}
