// Fused ODE action + adjoint gradient: the "strip walk".
//
// Replaces, for one RF value and a batch of paths, what the reference does with an ADOL-C tape of
// va_ode.Annealer.A_gaussian (va_ode.py:130-234 taped by _autodiffmin.py:32-49 and replayed by
// :57-58): one pass over X computes the measurement error, the model error under the chosen
// discretisation, and the analytic gradient.
//
// Work decomposition
//   unit      = (path b, segment sg): Tseg consecutive time rows of one path.
//   row-group = TPR = D/C threads; thread s of the group owns components [s*C, s*C+C) ("strip").
//   A CTA holds RG row-groups, each walking its own unit forward in time, all in lockstep.
//   A thread loads only its own strip from HBM (16-byte vector loads, software-prefetched PD
//   rows ahead, every X element read once apart from the few lead-in rows per segment) and keeps
//   the rows it still needs in registers; the only inter-thread traffic is the 2-component halo
//   of the periodic Lorenz96 stencil, exchanged through double-buffered shared-memory rows with
//   one __syncthreads per phase.  Dense models (Lorenz63, NaKL: C = D, H = 0) need no exchange.
//   Gradient rows are written once, as soon as their adjoint seed is complete.
//   Per-unit partial sums (me, fe, parameter gradients) go to a partials buffer and are reduced
//   in a fixed order by ode_finalize -> bit-reproducible results, no atomics.
//
// Every walker is a plain struct with init / prologue / phase<U,PH>(step) / finish methods made of
// VAB_HD code, so tests/emul/ can run the identical code thread-by-thread on the host.
#pragma once
#include "ode_models.cuh"
#include "vab_hd.h"

enum { DISC_EULER = 0, DISC_TRAPEZOID = 1, DISC_SIMPSON = 2, DISC_FORWARDMAP = 3, DISC_RK4 = 4 };

struct OdeParams {
  const double* XP;       // (B, ldxp)
  long long ldxp;
  double* G;              // (B, ldg) or nullptr (value only)
  long long ldg;
  int B, D, N, N_data, nskip, L;
  double dt;
  const int* obs_slot;    // (D) -> column of Y (library layout: columns sorted by component), or -1
  const double* Y;        // (N_data, Lp) library copy of the observations, row pitch Lp (even)
  int Lp;
  double rm_scalar;
  const double* rm_arr;   // (N_data, Lp) same layout as Y, or nullptr
  const int* win_y0;      // (nwin) first Y column (even) a window's stage copies; stream kernels
  int Lw;                 // Y columns copied per row and window (even)
  double rf_scalar;       // RF0*scale when rf_arr == nullptr
  const double* rf_arr;   // RF0 (N-1, D) or nullptr
  double rf_scale;
  const double* stim;     // (N, S) or nullptr
  int S;
  int NP, NPest;
  const int* pmap;        // (NP) -> index among estimated parameters, or -1
  const double* pfix;     // fixed parameter values
  long long pfix_stride;  // 0 (shared) or NP
  int Tseg, nseg;         // rows per segment, segments per path
  int TPR, RG;            // threads per row-group, row-groups per CTA
  int nunits;             // units in the launch
  int upp;                // units per path (their partials are contiguous and summed in order)
  int wpb;                // stream kernels: warps per (segment, window) = ceil(B / GPW)
  int GW, GPW, WS, nwin, NHL;   // sweep kernels: lanes per group, groups per warp, output strips
                                // per window, windows per row, halo lanes either side (0 = wrap)
  int K;                  // partial slots per unit = 2 + NPM
  double* partials;       // (nunits, K)
  const int* active;      // (B) or nullptr: skip paths with active[b] == 0
  double cm, cf;          // 1/(L*N_data), 1/(D*(N-1))
};

// --------------------------------------------------------------------------------------------
template <class M>
struct WalkBase {
  static constexpr int C = M::C, H = M::H, W = M::C + 2 * M::H, NPM = M::NPM;
  const OdeParams* Pp;
  int q, s, i0, unit, b;
  bool act, lane_in;     // lane_in: thread belongs to a row-group (idle tail lanes must not touch smem rows)
  int r0, r1;
  const double* xpath;
  double* gpath;
  double* sm;             // this row-group's exchange rows
  double* red;            // CTA reduction area
  double p[NPM];
  int hidx[2 * H + 1];    // wrapped component index of each halo slot (+1: no zero-size array)
  int slot[C];
  double me_acc, fe_acc, pacc[NPM];

  VAB_HD void base_init(const OdeParams& prm, int bid, int tid, double* smem, int nbuf) {
    Pp = &prm;
    const OdeParams& P = prm;
    q = tid / P.TPR;
    s = tid - q * P.TPR;
    i0 = s * C;
    unit = bid * P.RG + q;
    lane_in = (q < P.RG);
    const bool lane_ok = lane_in && (unit < P.nunits);
    b = lane_ok ? unit / P.nseg : 0;
    const int sg = lane_ok ? unit - b * P.nseg : 0;
    act = lane_ok && (P.active == nullptr || vab_ldg(P.active + b) != 0);
    r0 = sg * P.Tseg;
    r1 = r0 + P.Tseg;
    if (r1 > P.N) r1 = P.N;
    if (!lane_ok) { r0 = 0; r1 = 0; }
    xpath = P.XP + (long long)b * P.ldxp;
    gpath = P.G ? P.G + (long long)b * P.ldg : nullptr;
    sm = smem + (long long)(q < P.RG ? q : 0) * nbuf * P.D;
    red = smem;
    const long long nX = (long long)P.N * P.D;
#pragma unroll
    for (int k = 0; k < NPM; ++k) {
      double v = 0.0;
      if (act) {
        const int e = vab_ldg(P.pmap + k);
        v = (e >= 0) ? vab_ldg(xpath + nX + e) : vab_ldg(P.pfix + (long long)b * P.pfix_stride + k);
      }
      p[k] = v;
      pacc[k] = 0.0;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      hidx[h] = (i0 - H + h + P.D) % P.D;
      hidx[H + h] = (i0 + C + h) % P.D;
    }
#pragma unroll
    for (int j = 0; j < C; ++j) slot[j] = act ? vab_ldg(P.obs_slot + i0 + j) : -1;
    me_acc = 0.0;
    fe_acc = 0.0;
  }

  VAB_HD bool owned(int r) const { return act && r >= r0 && r < r1; }

  VAB_HD void load_own(int row, bool need, double* dst) const {
    if (need) {
      vab_load_strip<C>(xpath + (long long)row * Pp->D + i0, dst);
    } else {
#pragma unroll
      for (int j = 0; j < C; ++j) dst[j] = 0.0;
    }
  }
  VAB_HD void put_row(double* buf, const double* own) const {
    if constexpr (H > 0) {
      if (lane_in) {
#pragma unroll
        for (int j = 0; j < C; ++j) buf[i0 + j] = own[j];
      }
    }
  }
  // full[H..H+C) must already hold the own values
  VAB_HD void get_halo(const double* buf, double* full) const {
    if constexpr (H > 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        full[h] = buf[hidx[h]];
        full[H + C + h] = buf[hidx[H + h]];
      }
    }
  }
  VAB_HD const double* stim_row(int row) const {
    return (M::NSTIM > 0 && Pp->stim != nullptr) ? Pp->stim + (long long)row * Pp->S : nullptr;
  }
  VAB_HD double wgt(int row, int j) const {
    return Pp->rf_arr ? vab_ldg(Pp->rf_arr + (long long)row * Pp->D + i0 + j) * Pp->rf_scale : Pp->rf_scalar;
  }
  // measurement term of row r (va_ode.py:138-158): adds to the direct gradient and to me_acc
  VAB_HD void measure(int r, const double* xown, double* dir) {
    if (r % Pp->nskip != 0) return;
    const long long nd = r / Pp->nskip;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      if (slot[j] >= 0) {
        const long long o = nd * Pp->Lp + slot[j];
        const double rm = Pp->rm_arr ? vab_ldg(Pp->rm_arr + o) : Pp->rm_scalar;
        const double diff = xown[j] - vab_ldg(Pp->Y + o);
        me_acc = fma(rm * diff, diff, me_acc);
        dir[j] = fma(2.0 * Pp->cm * rm, diff, dir[j]);
      }
    }
  }
  VAB_HD void store_grad(int r, const double* dir, const double* jt) {
    if (gpath == nullptr) return;
    double g[C];
#pragma unroll
    for (int j = 0; j < C; ++j) g[j] = dir[j] - jt[j];
    vab_store_strip<C>(gpath + (long long)r * Pp->D + i0, g);
  }
  // thread partials -> smem (call after the last phase + a barrier)
  VAB_HD void finish_write(int tid, double psign) const {
    double* dst = red + (long long)tid * Pp->K;
    const bool lane = (q < Pp->RG);
    dst[0] = lane ? me_acc * Pp->cm : 0.0;
    dst[1] = lane ? fe_acc * Pp->cf : 0.0;
#pragma unroll
    for (int k = 0; k < NPM; ++k) dst[2 + k] = lane ? psign * pacc[k] : 0.0;
  }
};

// fixed-order reduction of the thread partials of each row-group (runs after finish_write + barrier)
VAB_HD void walk_reduce(const OdeParams& P, int bid, int tid, int nthreads, const double* red) {
  for (int idx = tid; idx < P.RG * P.K; idx += nthreads) {
    const int q = idx / P.K, k = idx - q * P.K;
    const int unit = bid * P.RG + q;
    if (unit >= P.nunits) continue;
    double acc = 0.0;
    for (int s = 0; s < P.TPR; ++s) acc += red[(long long)(q * P.TPR + s) * P.K + k];
    P.partials[(long long)unit * P.K + k] = acc;
  }
}

// --------------------------------------------------------------------------------------------
// euler / trapezoid / forwardmap:  e_m = x_{m+1} - AL x_m - (a f_m + bb f_{m+1})
//   direct_m = lam_{m-1} - AL lam_m,   v_m = bb lam_{m-1} + a lam_m,   g_m = direct_m - J^T(x_m) v_m
template <class M, int DISC, int PD_>
struct WalkTwoPoint : WalkBase<M> {
  using Base = WalkBase<M>;
  using Base::C; using Base::H; using Base::W; using Base::NPM;
  static constexpr int PD = PD_;
  static constexpr int NPH = 1;
  static constexpr int NBUF = (H > 0) ? 4 : 0;     // XB[2], VB[2]
  static constexpr double PSIGN = -1.0;
  double ca, cb, al;
  double X1[W], X2[W], F1[C], lam2[C], v2[C], d2[C];
  double pf[PD][C];

  VAB_HD static int nsteps(const OdeParams& P) { return ((P.Tseg + 3 + PD - 1) / PD) * PD; }
  VAB_HD bool need(int row) const {
    return this->act && row >= 0 && row >= this->r0 - 1 && row <= this->r1 && row < this->Pp->N;
  }
  VAB_HD void init(const OdeParams& prm, int bid, int tid, double* smem) {
    this->base_init(prm, bid, tid, smem, NBUF);
    const double dt = prm.dt;
    ca = (DISC == DISC_EULER) ? dt : (DISC == DISC_TRAPEZOID ? 0.5 * dt : 1.0);
    cb = (DISC == DISC_TRAPEZOID) ? 0.5 * dt : 0.0;
    al = (DISC == DISC_FORWARDMAP) ? 0.0 : 1.0;
#pragma unroll
    for (int j = 0; j < W; ++j) { X1[j] = 0.0; X2[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < C; ++j) { F1[j] = 0.0; lam2[j] = 0.0; v2[j] = 0.0; d2[j] = 0.0; }
  }
  VAB_HD void prologue() {
#pragma unroll
    for (int u = 0; u < PD; ++u) this->load_own(this->r0 - 1 + u, need(this->r0 - 1 + u), pf[u]);
    this->put_row(this->sm + 0 * this->Pp->D, pf[0]);       // XB[0], read by phase 0
    double z[C];
#pragma unroll
    for (int j = 0; j < C; ++j) z[j] = 0.0;
    this->put_row(this->sm + 2 * this->Pp->D, z);           // VB[0]
  }
  template <int U, int PH>
  VAB_HD void phase(int step) {
    constexpr int rp = U & 1, wp = rp ^ 1;
    const int D = this->Pp->D;
    const int m = this->r0 - 1 + step;
    const double* xb_r = this->sm + rp * D;
    const double* vb_r = this->sm + (2 + rp) * D;
    double* xb_w = this->sm + wp * D;
    double* vb_w = this->sm + (2 + wp) * D;

    // row m arrives
    double Xm[W], Fm[C];
#pragma unroll
    for (int j = 0; j < C; ++j) Xm[H + j] = pf[U][j];
    this->get_halo(xb_r, Xm);
    const bool vm = need(m);
    if (vm) {
      M::f(Xm, this->p, this->stim_row(m), Fm);
    } else {
#pragma unroll
      for (int j = 0; j < C; ++j) Fm[j] = 0.0;
    }
    // residual m-1
    double lam[C];
    const bool ve = vm && step >= 1 && m >= 1;
    const bool own_e = this->owned(m - 1);
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double l = 0.0;
      if (ve) {
        const double w = this->wgt(m - 1, j);
        const double e = Xm[H + j] - al * X1[H + j] - (ca * F1[j] + cb * Fm[j]);
        l = 2.0 * this->Pp->cf * w * e;
        if (own_e) this->fe_acc = fma(w * e, e, this->fe_acc);
      }
      lam[j] = l;
    }
    // seed + direct term of row m-1
    double v1[C], d1[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      v1[j] = cb * lam2[j] + ca * lam[j];
      d1[j] = lam2[j] - al * lam[j];
    }
    if (own_e) this->measure(m - 1, X1 + H, d1);
    // gradient of row m-2
    {
      double V2[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) V2[H + j] = v2[j];
      this->get_halo(vb_r, V2);
      if (this->owned(m - 2)) {
        M::adj(X2, V2, this->p, jt, this->pacc);
        this->store_grad(m - 2, d2, jt);
      }
    }
    // publish own strips for the next phase, refill the prefetch slot
    this->put_row(xb_w, pf[(U + 1) % PD]);
    this->put_row(vb_w, v1);
    this->load_own(m + PD, need(m + PD), pf[U]);
#pragma unroll
    for (int j = 0; j < W; ++j) { X2[j] = X1[j]; X1[j] = Xm[j]; }
#pragma unroll
    for (int j = 0; j < C; ++j) { F1[j] = Fm[j]; lam2[j] = lam[j]; v2[j] = v1[j]; d2[j] = d1[j]; }
  }
};

// --------------------------------------------------------------------------------------------
// Simpson-Hermite (va_ode.py:404-437 + :192-195).  Pair k = rows (a, b, c) = (2k, 2k+1, 2k+2):
//   e1 = x_c - x_a - dt/3 (f_a + 4 f_b + f_c),  e2 = x_b - (x_a + x_c)/2 - dt/4 (f_a - f_c)
// One phase per pair; rows a, b of a pair get their gradient one phase later (after the seed halo
// exchange).  Segments start on even rows; the walk begins two pairs early so that the c-part of
// row r0 is available.
template <class M, int PD_>
struct WalkSimpson : WalkBase<M> {
  using Base = WalkBase<M>;
  using Base::C; using Base::H; using Base::W; using Base::NPM;
  static constexpr int PD = PD_;
  static constexpr int NPH = 1;
  static constexpr int NBUF = (H > 0) ? 8 : 0;     // XB[2][2 rows], VB[2][2 rows]
  static constexpr double PSIGN = -1.0;
  double Xa[W], Fa[C];                 // row a of the current pair (= c of the previous one)
  double Xao[W], Xbo[W];               // rows a, b of the previous pair (awaiting their gradient)
  double vao[C], vbo[C], dao[C], dbo[C];
  double vcp[C], dcp[C];               // c-part carried into the next pair's row a
  double pf[PD][2][C];                 // prefetched rows (b, c) of upcoming pairs

  VAB_HD static int nsteps(const OdeParams& P) { return ((P.Tseg / 2 + 4 + PD - 1) / PD) * PD; }
  VAB_HD bool need(int row) const {
    return this->act && row >= 0 && row >= this->r0 - 2 && row <= this->r1 && row < this->Pp->N;
  }
  VAB_HD void init(const OdeParams& prm, int bid, int tid, double* smem) {
    this->base_init(prm, bid, tid, smem, NBUF);
#pragma unroll
    for (int j = 0; j < W; ++j) { Xa[j] = 0.0; Xao[j] = 0.0; Xbo[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < C; ++j) {
      Fa[j] = 0.0; vao[j] = 0.0; vbo[j] = 0.0; dao[j] = 0.0; dbo[j] = 0.0; vcp[j] = 0.0; dcp[j] = 0.0;
    }
  }
  VAB_HD void prologue() {
    // pair index of phase 0 is k0 = r0/2 - 2 -> rows b = r0-3, c = r0-2
#pragma unroll
    for (int u = 0; u < PD; ++u) {
      const int a = this->r0 - 4 + 2 * u;
      this->load_own(a + 1, need(a + 1), pf[u][0]);
      this->load_own(a + 2, need(a + 2), pf[u][1]);
    }
    const int D = this->Pp->D;
    this->put_row(this->sm + 0 * D, pf[0][0]);      // XB[0] row b
    this->put_row(this->sm + 1 * D, pf[0][1]);      // XB[0] row c
    double z[C];
#pragma unroll
    for (int j = 0; j < C; ++j) z[j] = 0.0;
    this->put_row(this->sm + 4 * D, z);             // VB[0] rows
    this->put_row(this->sm + 5 * D, z);
  }
  template <int U, int PH>
  VAB_HD void phase(int step) {
    constexpr int rp = U & 1, wp = rp ^ 1;
    const int D = this->Pp->D;
    const double dt = this->Pp->dt;
    const int a = this->r0 - 4 + 2 * step, bq = a + 1, c = a + 2;
    const double* xb_r = this->sm + (2 * rp) * D;
    const double* vb_r = this->sm + (4 + 2 * rp) * D;
    double* xb_w = this->sm + (2 * wp) * D;
    double* vb_w = this->sm + (4 + 2 * wp) * D;

    // gradients of the previous pair's rows a-2, b-2 (their seeds were published last phase)
    {
      double V[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) V[H + j] = vao[j];
      this->get_halo(vb_r, V);
      if (this->owned(a - 2)) {
        M::adj(Xao, V, this->p, jt, this->pacc);
        this->store_grad(a - 2, dao, jt);
      }
#pragma unroll
      for (int j = 0; j < C; ++j) V[H + j] = vbo[j];
      this->get_halo(vb_r + D, V);
      if (this->owned(bq - 2)) {
        M::adj(Xbo, V, this->p, jt, this->pacc);
        this->store_grad(bq - 2, dbo, jt);
      }
    }
    // rows b, c arrive
    double Xb[W], Xc[W], Fb[C], Fc[C];
#pragma unroll
    for (int j = 0; j < C; ++j) { Xb[H + j] = pf[U][0][j]; Xc[H + j] = pf[U][1][j]; }
    this->get_halo(xb_r, Xb);
    this->get_halo(xb_r + D, Xc);
    const bool nb = need(bq), nc = need(c);
    if (nb) { M::f(Xb, this->p, this->stim_row(bq), Fb); }
    else {
#pragma unroll
      for (int j = 0; j < C; ++j) Fb[j] = 0.0;
    }
    if (nc) { M::f(Xc, this->p, this->stim_row(c), Fc); }
    else {
#pragma unroll
      for (int j = 0; j < C; ++j) Fc[j] = 0.0;
    }
    // pair residuals
    const bool vp = need(a) && nb && nc && step >= 1;
    const bool own_p = this->owned(bq);
    double va[C], vb[C], da[C], db[C], vcn[C], dcn[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double l1 = 0.0, l2 = 0.0;
      if (vp) {
        const double w1 = this->wgt(a, j), w2 = this->wgt(bq, j);
        const double e1 = Xc[H + j] - Xa[H + j] - (dt / 3.0) * (Fa[j] + 4.0 * Fb[j] + Fc[j]);
        const double e2 = Xb[H + j] - 0.5 * (Xa[H + j] + Xc[H + j]) - (dt / 4.0) * (Fa[j] - Fc[j]);
        l1 = 2.0 * this->Pp->cf * w1 * e1;
        l2 = 2.0 * this->Pp->cf * w2 * e2;
        if (own_p) this->fe_acc = fma(w1 * e1, e1, fma(w2 * e2, e2, this->fe_acc));
      }
      vb[j] = (4.0 * dt / 3.0) * l1;
      db[j] = l2;
      va[j] = vcp[j] + (dt / 3.0) * l1 + (dt / 4.0) * l2;
      da[j] = dcp[j] - l1 - 0.5 * l2;
      vcn[j] = (dt / 3.0) * l1 - (dt / 4.0) * l2;
      dcn[j] = l1 - 0.5 * l2;
    }
    if (this->owned(a)) this->measure(a, Xa + H, da);
    if (own_p) this->measure(bq, Xb + H, db);
    // publish, prefetch, rotate
    this->put_row(xb_w, pf[(U + 1) % PD][0]);
    this->put_row(xb_w + D, pf[(U + 1) % PD][1]);
    this->put_row(vb_w, va);
    this->put_row(vb_w + D, vb);
    {
      const int an = a + 2 * PD;
      this->load_own(an + 1, need(an + 1), pf[U][0]);
      this->load_own(an + 2, need(an + 2), pf[U][1]);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) { Xao[j] = Xa[j]; Xbo[j] = Xb[j]; Xa[j] = Xc[j]; }
#pragma unroll
    for (int j = 0; j < C; ++j) {
      Fa[j] = Fc[j];
      vao[j] = va[j]; vbo[j] = vb[j]; dao[j] = da[j]; dbo[j] = db[j];
      vcp[j] = vcn[j]; dcp[j] = dcn[j];
    }
  }
};

// --------------------------------------------------------------------------------------------
// RK4 (extension; intent at va_ode.py:382-402):  e_m = x_{m+1} - x_m - dt/6 (k1 + 2k2 + 2k3 + k4).
// Eight phases per row: three forward stage exchanges (y2, y3, y4), then the discrete adjoint
// walks the stages backwards with one seed exchange per stage.
template <class M, int PD_>
struct WalkRK4 : WalkBase<M> {
  using Base = WalkBase<M>;
  using Base::C; using Base::H; using Base::W; using Base::NPM;
  static constexpr int PD = PD_;
  static constexpr int NPH = 8;
  static constexpr int NBUF = (H > 0) ? 2 : 0;     // A, B alternate every phase
  static constexpr double PSIGN = 1.0;
  double Y1[W], Y2[W], Y3[W], Y4[W];
  double k1[C], k2[C], k3[C];
  double lam[C], lamp[C], xb[C], kb[C];
  double pf[PD][C];
  bool vr, va;            // residual m valid / its adjoint belongs to this segment

  VAB_HD static int nsteps(const OdeParams& P) { return ((P.Tseg + 1 + PD - 1) / PD) * PD; }
  VAB_HD bool need(int row) const {
    return this->act && row >= 0 && row >= this->r0 - 1 && row <= this->r1 && row < this->Pp->N;
  }
  VAB_HD void init(const OdeParams& prm, int bid, int tid, double* smem) {
    this->base_init(prm, bid, tid, smem, NBUF);
#pragma unroll
    for (int j = 0; j < C; ++j) { lamp[j] = 0.0; lam[j] = 0.0; }
    vr = false;
    va = false;
  }
  VAB_HD void prologue() {
    // pf[u] holds row (r0-1+u)+1; the row entering phase 0 of step 0 goes straight to buffer B
    double x0[C];
    this->load_own(this->r0 - 1, need(this->r0 - 1), x0);
#pragma unroll
    for (int j = 0; j < C; ++j) Y1[H + j] = x0[j];
    this->put_row(this->sm + this->Pp->D, x0);
#pragma unroll
    for (int u = 0; u < PD; ++u) this->load_own(this->r0 + u, need(this->r0 + u), pf[u]);
  }
  template <int U, int PH>
  VAB_HD void phase(int step) {
    const int D = this->Pp->D;
    const double dt = this->Pp->dt;
    const int m = this->r0 - 1 + step;
    double* bufA = this->sm;
    double* bufB = this->sm + D;
    const double* rd = (PH % 2 == 0) ? bufB : bufA;
    double* wr = (PH % 2 == 0) ? bufA : bufB;
    if constexpr (PH == 0) {
      vr = need(m) && (m + 1 < this->Pp->N) && (m + 1 <= this->r1);
      va = vr && this->owned(m);
      this->get_halo(rd, Y1);
      if (vr) M::f(Y1, this->p, nullptr, k1);
      double own[C];
#pragma unroll
      for (int j = 0; j < C; ++j) { if (!vr) k1[j] = 0.0; own[j] = Y1[H + j] + 0.5 * dt * k1[j]; Y2[H + j] = own[j]; }
      this->put_row(wr, own);
    } else if constexpr (PH == 1) {
      this->get_halo(rd, Y2);
      if (vr) M::f(Y2, this->p, nullptr, k2);
      double own[C];
#pragma unroll
      for (int j = 0; j < C; ++j) { if (!vr) k2[j] = 0.0; own[j] = Y1[H + j] + 0.5 * dt * k2[j]; Y3[H + j] = own[j]; }
      this->put_row(wr, own);
    } else if constexpr (PH == 2) {
      this->get_halo(rd, Y3);
      if (vr) M::f(Y3, this->p, nullptr, k3);
      double own[C];
#pragma unroll
      for (int j = 0; j < C; ++j) { if (!vr) k3[j] = 0.0; own[j] = Y1[H + j] + dt * k3[j]; Y4[H + j] = own[j]; }
      this->put_row(wr, own);
    } else if constexpr (PH == 3) {
      this->get_halo(rd, Y4);
      double k4[C];
      if (vr) M::f(Y4, this->p, nullptr, k4);
      const bool own_e = this->owned(m);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        double l = 0.0;
        if (vr) {
          const double w = this->wgt(m, j);
          const double e = pf[U][j] - Y1[H + j] - (dt / 6.0) * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j]);
          l = 2.0 * this->Pp->cf * w * e;
          if (own_e) this->fe_acc = fma(w * e, e, this->fe_acc);
        }
        lam[j] = l;
        xb[j] = -l;
        kb[j] = -(dt / 6.0) * l;            // k4 adjoint
      }
      this->put_row(wr, kb);
    } else if constexpr (PH == 4) {
      double Kf[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) Kf[H + j] = kb[j];
      this->get_halo(rd, Kf);
#pragma unroll
      for (int j = 0; j < C; ++j) jt[j] = 0.0;
      if (va) M::adj(Y4, Kf, this->p, jt, this->pacc);
#pragma unroll
      for (int j = 0; j < C; ++j) { xb[j] += jt[j]; kb[j] = -(dt / 3.0) * lam[j] + dt * jt[j]; }
      this->put_row(wr, kb);
    } else if constexpr (PH == 5) {
      double Kf[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) Kf[H + j] = kb[j];
      this->get_halo(rd, Kf);
#pragma unroll
      for (int j = 0; j < C; ++j) jt[j] = 0.0;
      if (va) M::adj(Y3, Kf, this->p, jt, this->pacc);
#pragma unroll
      for (int j = 0; j < C; ++j) { xb[j] += jt[j]; kb[j] = -(dt / 3.0) * lam[j] + 0.5 * dt * jt[j]; }
      this->put_row(wr, kb);
    } else if constexpr (PH == 6) {
      double Kf[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) Kf[H + j] = kb[j];
      this->get_halo(rd, Kf);
#pragma unroll
      for (int j = 0; j < C; ++j) jt[j] = 0.0;
      if (va) M::adj(Y2, Kf, this->p, jt, this->pacc);
#pragma unroll
      for (int j = 0; j < C; ++j) { xb[j] += jt[j]; kb[j] = -(dt / 6.0) * lam[j] + 0.5 * dt * jt[j]; }
      this->put_row(wr, kb);
    } else {
      double Kf[W], jt[C];
#pragma unroll
      for (int j = 0; j < C; ++j) Kf[H + j] = kb[j];
      this->get_halo(rd, Kf);
#pragma unroll
      for (int j = 0; j < C; ++j) jt[j] = 0.0;
      if (va) M::adj(Y1, Kf, this->p, jt, this->pacc);
      if (this->owned(m)) {
        double dir[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { dir[j] = lamp[j] + xb[j] + jt[j]; jt[j] = 0.0; }
        this->measure(m, Y1 + H, dir);
        this->store_grad(m, dir, jt);
      }
      // next row enters: own strip to buffer B, refill prefetch slot
#pragma unroll
      for (int j = 0; j < C; ++j) { lamp[j] = lam[j]; Y1[H + j] = pf[U][j]; }
      this->put_row(wr, pf[U]);
      this->load_own(m + 1 + PD, need(m + 1 + PD), pf[U]);
    }
  }
};

// --------------------------------------------------------------------------------------------
template <class M, int DISC, int PD>
struct WalkSelect { using type = WalkTwoPoint<M, DISC, PD>; };
template <class M, int PD>
struct WalkSelect<M, DISC_SIMPSON, PD> { using type = WalkSimpson<M, PD>; };
template <class M, int PD>
struct WalkSelect<M, DISC_RK4, PD> { using type = WalkRK4<M, PD>; };

// shared-memory doubles a CTA needs: exchange rows or the reduction area, whichever is larger
template <class WK>
VAB_HD long long walk_smem_doubles(const OdeParams& P, int nthreads) {
  const long long ex = (long long)P.RG * WK::NBUF * P.D;
  const long long rd = (long long)nthreads * P.K;
  return ex > rd ? ex : rd;
}
