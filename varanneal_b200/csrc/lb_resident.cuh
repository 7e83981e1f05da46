// Shared-memory-resident L-BFGS ladder for small Lorenz96 problems (the shipped example, C1:
// D = 20, N = 161, n = 3221).  Included by lbfgs.cu (needs LbPath, the lb_*_body decision code,
// lb_gram_warp, cta_reduce_sum, LbLadder).
//
// At this size a minimiser cycle is a chain of dependent L2 round trips (lb_fused_kernel, DESIGN
// 4.3c: ~1-2 us per phase whatever the work).  Here nothing but the observations is read from
// global memory inside the loop: one cluster of RCS CTAs per path keeps x, g, d, the trial point and
// its gradient, and the 2m history vectors in *distributed shared memory* -- CTA c owns the time rows
// [c RPC, (c+1) RPC) of every vector (C1: 25 vectors x 22 rows x 20 components = 88 KB per CTA) --
// and the whole ladder of the path runs inside one launch:
//
//   trial point -> cluster barrier -> halo rows of the trial point from the neighbour CTAs (DSMEM)
//   -> action + adjoint gradient of the own rows, *all rows in parallel* (f, residuals / seeds,
//      J^T products: three block-local phases; the rows next to a slice boundary are recomputed,
//      not exchanged) -> partial sums -> cluster barrier -> every CTA gathers the partials of all
//      CTAs in rank order, so every CTA holds bit-identical totals and runs the decision code
//      (line search, two-loop recursion, start of the next search, end of rung) redundantly on its
//      own copy of the path state: no broadcast, four cluster barriers per cycle.
//
// Same algorithm and the same decision functions as the per-phase kernels; the action is a second
// implementation of va_ode.A_gaussian for Lorenz96 (va_ode.py:130-234, 341-380, 404-454) and is
// checked against the oracle like the others.  Scope: Lorenz96, static parameter (estimated or not),
// scalar RM and RF0, no stimulus, unbounded L-BFGS; everything else takes the general path.
#pragma once

namespace {

// CTAs per path: a template parameter (8: shortest cycle; 4: twice as many paths resident at once)
constexpr int RNT = 256;                // threads per CTA
static_assert(RNT == NT, "block_reduce / cta_reduce_sum are written for NT threads");

struct ResArgs {
  double* XP;             // (B, ld) paths, in / out
  long long ld;
  int N, D, nskip, nobs;  // model rows, state dimension, merr_nskip, observed components
  int RPC;                // rows per CTA (even)
  int disc;
  double dt, cf, rf0;
  const double* Y;        // (N_data, D) dense observations
  const double* wobs;     // (D) 2 cm RM at the observed components
  int k_est;              // 1: the forcing is the last unknown of the path; 0: fixed
  const double* pfix;     // fixed forcing per path (stride pfix_stride) when !k_est
  long long pfix_stride;
  LbPath* st;             // (B) state in / out
  LbOpts o;
  LbLadder L;
  long long max_cycles;
  long long* dbg;         // development aid (VAB_RESIDENT_TIMING=1): SM cycles per phase, cluster 0
};

// vector slots of the per-CTA store
// (the trial point lives apart from them, between its halo rows: XTH)
enum { RV_X = 0, RV_G = 1, RV_D = 2, RV_GT = 3, RV_S = 4, RV_Y = 4 + MMAX, RV_N = 4 + 2 * MMAX };

// History bookkeeping of an accepted step and the two-loop recursion r = H g in coefficient space
// (r = cg g + sum_j cs_j s_j + cy_j y_j, d = -r) by one warp, on the path state in shared memory.
// Same algebra as lb_gram_warp / lb_gram_body, but *column oriented*: lane k owns the pair of age k
// (0 = oldest) and keeps its own running inner product (s_k . q in the first loop, y_k . r in the
// second); when the coefficient of pair c becomes known it is broadcast and every lane that still
// needs it applies one multiply-add.  The dependent chain per step is a multiply, a shuffle and a
// multiply-add instead of a half-warp reduction and a division (~100 instead of ~600 clocks).
//   first loop   a_c = rho_c (s_c.g - sum_{k newer} a_k s_c.y_k)
//   second loop  b_c = rho_c (gamma (y_c.g - sum_k a_k y_c.y_k) + sum_{k older} (a_k - b_k) s_k.y_c)
//   cg = gamma, cy_k = -gamma a_k, cs_k = a_k - b_k.                          All 32 lanes must call.
__device__ void res_gram_warp(LbPath& s, const double* sum, int m, long long* tacc = nullptr) {
  const unsigned full = 0xffffffffu;
  const int j = threadIdx.x & 31;
  long long tq = tacc ? clock64() : 0;
  auto lap = [&](int slot) {                            // development aid: clocks per section (lane 0)
    if (tacc != nullptr && j == 0) { const long long t = clock64(); tacc[slot] += t - tq; tq = t; }
  };
  const bool on = j < m;
  int col = s.col, head = s.head;
  const bool accepted = s.accepted != 0, upd = s.do_update != 0;
  const bool newpair = accepted && upd;
  const int p = s.pslot;
  const double dr = s.dr, yy = sum[5 * MMAX];
  const double theta_old = s.theta;
  if (newpair) {
    if (col < m) col += 1;
    else head = (head + 1 == m) ? 0 : head + 1;
  }
  const int k = j;                                   // age of this lane's pair
  const bool act = k < col;
  int ik = head + k;
  if (ik >= m) ik -= m;
  if (!act) ik = 0;
  // The divisions of the phase -- rho_k = 1 / s_k.y_k per lane, theta = y.y / s.y, gamma = 1 / theta --
  // as ONE division instruction: lanes MMAX and MMAX + 1 carry the operands of theta and gamma (a
  // dependent fp64 division costs ~500 clocks here; three in a row were a third of the phase)
  const double diag = (newpair && ik == p) ? dr : s.SY[ik * MMAX + ik];
  double num = 1.0, den = act ? diag : 1.0;
  if (j == MMAX) { num = newpair ? yy : theta_old; den = newpair ? dr : 1.0; }
  if (j == MMAX + 1) { num = newpair ? dr : 1.0; den = newpair ? yy : theta_old; }
  const double quo = num / den;
  const double rho = act ? quo : 0.0;
  const double theta = __shfl_sync(full, quo, MMAX);
  const double gamma = (col > 0) ? __shfl_sync(full, quo, MMAX + 1) : 1.0;
  __syncwarp();
  if (accepted) {
    if (upd) {
      if (on && j != p) {
        s.SY[p * MMAX + j] = sum[2 * MMAX + j];      // s_p . y_j
        s.SY[j * MMAX + p] = sum[3 * MMAX + j];      // s_j . y_p
        s.YY[p * MMAX + j] = sum[4 * MMAX + j];
        s.YY[j * MMAX + p] = sum[4 * MMAX + j];
      }
      if (j == p) {
        s.SY[p * MMAX + p] = dr;                     // s'y from the line search, as L-BFGS-B does
        s.YY[p * MMAX + p] = yy;
      }
      if (on) { s.gS[j] = (j == p) ? sum[5 * MMAX + 1] : sum[j]; s.gY[j] = (j == p) ? sum[5 * MMAX + 2] : sum[MMAX + j]; }
      if (j == 0) { s.theta = theta; s.col = col; s.head = head; }
    } else if (on) {
      s.gS[j] = sum[j]; s.gY[j] = sum[MMAX + j];
    }
    if (j == 0) s.gg = sum[5 * MMAX + 3];
  }
  __syncwarp();
  lap(10);
  // this lane's row / column of the Gram blocks in age order, in registers before the chains start
  // (entries of pairs that do not exist yet are loaded too -- always inside the arrays -- and never
  // used); the row of the first loop is pre-scaled by rho_k, which takes the multiply out of its chain
  double syr[MMAX], syc[MMAX], yyr[MMAX];
  {
    const double* syrow = s.SY + ik * MMAX;
    const double* yyrow = s.YY + ik * MMAX;
    const double* sycol = s.SY + ik;
    int ic = head < m ? head : 0;
#pragma unroll
    for (int c = 0; c < MMAX; ++c) {
      // masked here, off the chains: the first loop updates only older pairs (k < c), the second
      // only newer ones (k > c); pairs that do not exist contribute nothing
      // (pairs that do not exist are skipped by the loops; lanes without a pair compute garbage
      // nobody reads)
      syr[c] = (k < c) ? rho * syrow[ic] : 0.0;      // rho_k s_k . y_c
      syc[c] = (k > c) ? sycol[ic * MMAX] : 0.0;     // s_c . y_k
      yyr[c] = yyrow[ic];
      ic = (ic + 1 >= m) ? 0 : ic + 1;
    }
  }
  double t = act ? rho * s.gS[ik] : 0.0;             // rho_k s_k . q
  double u = act ? s.gY[ik] : 0.0;
  lap(11);
  double a = 0.0;
#pragma unroll
  for (int c = MMAX - 1; c >= 0; --c) {              // newest -> oldest
    if (c < col) {
      const double ac = __shfl_sync(full, t, c);
      t = fma(-ac, syr[c], t);                       // the chain: shuffle, multiply-add
      if (k == c) a = ac;
      u = fma(-ac, yyr[c], u);                       // y_k . (g - sum_c a_c y_c), off the chain
    }
  }
  u *= gamma;
  lap(12);
  double cs = 0.0;
#pragma unroll
  for (int c = 0; c < MMAX; ++c) {                   // oldest -> newest
    if (c < col) {
      const double cc = __shfl_sync(full, fma(-rho, u, a), c);
      u = fma(cc, syc[c], u);                        // the chain: multiply-add, shuffle, multiply-add
      if (k == c) cs = cc;
    }
  }
  lap(13);
  if (j == 0) s.cg = gamma;
  if (j < MMAX) { s.cs[j] = 0.0; s.cy[j] = 0.0; }
  __syncwarp();
  if (act) { s.cs[ik] = cs; s.cy[ik] = -gamma * a; }
}

// Sums of 16 per-lane values over a warp with a halving butterfly (16 exchanges instead of 80): on
// return every lane l holds the warp total of entry l >> 1.  Fixed order.
__device__ __forceinline__ double warp_sum16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const double send = up ? v[k] : v[k + half];
      const double keep = up ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// NV per-thread values over the CTA: butterflies inside the warps, then the warps in order by NV
// threads.  wred: (RNT / 32) * NV doubles that no other reduction in flight uses.  One block barrier
// inside; out[] is complete behind the *next* barrier of the caller.  Fixed order.
template <int NV>
__device__ __forceinline__ void res_reduce(double (&v)[NV], const int (&op)[NV], double* out, double* wred) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double u = __shfl_xor_sync(0xffffffffu, v[k], sft);
      v[k] = (op[k] == RED_SUM) ? v[k] + u : (op[k] == RED_MAX ? fmax(v[k], u) : fmin(v[k], u));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) wred[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int k = threadIdx.x;
    int o = op[0];
#pragma unroll
    for (int kk = 1; kk < NV; ++kk)
      if (kk == k) o = op[kk];
    double a = wred[k];
#pragma unroll
    for (int w = 1; w < RNT / 32; ++w) {
      const double u = wred[w * NV + k];
      a = (o == RED_SUM) ? a + u : (o == RED_MAX ? fmax(a, u) : fmin(a, u));
    }
    out[k] = a;
  }
}

template <int DISC, int RCS>
__global__ void __launch_bounds__(RNT, 1) lb_resident_kernel(const ResArgs A) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ double rs[];
  const int rank = (int)cl.block_rank(), tid = threadIdx.x;
  const int b = blockIdx.x / RCS;
  const int N = A.N, D = A.D, RPC = A.RPC;
  const int R0 = rank * RPC, R1 = min(R0 + RPC, N);
  const int nrow = max(R1 - R0, 0);
  const int VS = RPC * D;                               // doubles per vector slice
  const int nloc = nrow * D;                            // own elements
  double* vec = rs;                                     // RV_N slices
  double* XTH = vec + (size_t)RV_N * VS;                // trial point with its halo rows: rows R0-2 .. R0+RPC  [(RPC + 3)][D]
  double* F = XTH + (size_t)(RPC + 3) * D;              // f of rows R0-2 .. R1           [(RPC + 3)][D]
  double* E1 = F + (size_t)(RPC + 3) * D;               // seeds per residual row / pair
  double* E2 = E1 + (size_t)(RPC + 3) * D;
  double* scratch = E2 + (size_t)(RPC + 3) * D;         // 8 * NT (block_reduce / cta_reduce_sum)
  // partial sums of this CTA, read by the others after a cluster barrier; one buffer per phase, so a
  // buffer is rewritten only after a later barrier that every reader has passed
  double* partB = scratch + 8 * NT;                     // NACC_U: update pass
  double* tot = partB + NACC_U;                         // NACC_U gathered totals
  __shared__ double partA[8];                           // evaluation: fe, me, dA/dk, g.d, max|g|
  __shared__ double partC[4];                           // direction: d.d, g.d
  double* Yl = tot + NACC_U;                            // observations of the own rows (0 where there is none)
  double* Wl = Yl + VS;                                 // their weights 2 cm RM (0 where there is none)
  __shared__ LbPath s;                                  // this CTA's copy of the path state
  __shared__ double pk[8 + 2 * MMAX];                   // the forcing: x, g, d, xt, gt, -, -, -, S_j, Y_j (replicated)
  __shared__ int dummy_act;
  __shared__ double rf_cur;                             // RF of the rung the path is on
  double* X = vec + RV_X * VS;
  double* G = vec + RV_G * VS;
  double* Dv = vec + RV_D * VS;
  double* XT = XTH + 2 * D;                             // own rows of the trial point
  double* GT = vec + RV_GT * VS;
  const long long nX = (long long)N * D;
  const long long n = nX + (A.k_est ? 1 : 0);
  double* xg = A.XP + (long long)b * A.ld;
  const bool last = rank == RCS - 1;                    // owns the forcing's terms of every dot product
  const int m = A.o.m;
  constexpr double al = (DISC == DISC_FORWARDMAP) ? 0.0 : 1.0;
  const double dt = A.dt;
  const double ca = (DISC == DISC_EULER) ? dt : (DISC == DISC_TRAPEZOID ? 0.5 * dt : 1.0);
  const double cb = (DISC == DISC_TRAPEZOID) ? 0.5 * dt : 0.0;

  // ---- load the path and its state
  for (int e = tid; e < nloc; e += RNT) X[e] = xg[(long long)R0 * D + e];
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(A.st + b);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&s);
    for (int k = tid; k < (int)(sizeof(LbPath) / 8); k += RNT) dst[k] = src[k];
  }
  if (tid == 0) {
    for (int k = 0; k < 8 + 2 * MMAX; ++k) pk[k] = 0.0;
    pk[0] = A.k_est ? xg[nX] : A.pfix[(long long)b * A.pfix_stride];
    dummy_act = 1;
  }
  for (int e = tid; e < (RPC + 3) * D; e += RNT) XTH[e] = 0.0;
  for (int e = tid; e < VS; e += RNT) {
    G[e] = 0.0; Dv[e] = 0.0;
    double w = 0.0, y = 0.0;
    if (e < nloc && A.nobs > 0) {
      const int lr = e / D, i = e - lr * D, r = R0 + lr;
      if (A.nskip == 1 || (r % A.nskip) == 0) {
        w = A.wobs[i];
        if (w != 0.0) y = A.Y[(long long)(A.nskip == 1 ? r : r / A.nskip) * D + i];
      }
    }
    Yl[e] = y; Wl[e] = w;
  }
  __syncthreads();
  if (tid == 0) rf_cur = A.rf0 * A.L.scales[s.ib];
  __syncthreads();
  // Halo rows are *pushed*: whoever writes a row of the trial point that a neighbour's stencil needs
  // stores it into that neighbour's XTH as well (remote shared-memory stores, visible behind the
  // cluster barrier that follows the trial point anyway).  rank - 1 needs my first row (its row R1),
  // rank + 1 my last two (its rows R0-2, R0-1).
  double* prevH = (rank > 0 && nrow > 0) ? cl.map_shared_rank(XTH, rank - 1) + (size_t)(RPC + 2) * D : nullptr;
  double* nextH = (R0 + RPC < N) ? cl.map_shared_rank(XTH, rank + 1) - (size_t)(RPC - 2) * D : nullptr;
  const int e_next = (RPC - 2) * D;                     // first element of the last two rows of a full slice
  cl.sync();                                            // every XTH is zeroed before the first push arrives

  // sum / max over the CTAs' partial k, in rank order
  auto gather_sum = [&](double* part, int k) {
    double v[RCS];
#pragma unroll
    for (int q = 0; q < RCS; ++q) v[q] = *cl.map_shared_rank(&part[k], q);
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < RCS; ++q) a += v[q];
    return a;
  };
  auto gather_max = [&](double* part, int k) {
    double v[RCS];
#pragma unroll
    for (int q = 0; q < RCS; ++q) v[q] = *cl.map_shared_rank(&part[k], q);
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < RCS; ++q) a = fmax(a, v[q]);
    return a;
  };
  // trial-point row r, component i (periodic): own slice or halo
  auto xt_at = [&](int r, int i) -> double {            // R0 - 2 <= r <= R0 + RPC
    if (i < 0) i += D; else if (i >= D) i -= D;
    return XTH[(r - (R0 - 2)) * D + i];
  };
  auto f_at = [&](int r, int i) -> double { return F[(r - (R0 - 2)) * D + i]; };

  __shared__ long long tacc[16];                        // phase clocks (development aid), flushed at the end
  if (tid < 16) tacc[tid] = 0;
  const bool timing = A.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  long long t_prev = 0;
  int t_slot = 0;
  auto stamp = [&]() {
    if (timing) {
      const long long t = clock64();
      if (t_slot > 0) tacc[t_slot] += t - t_prev;
      t_prev = t;
      ++t_slot;
    }
  };
  // The first trial point of a search is x + d (stp = 1) except in the first iteration of a rung: the
  // direction pass writes it (and pushes its halo rows) in the same sweep, its barrier serves as the
  // trial point's, and the next cycle skips the trial phase if the search did start with stp = 1.
  bool spec = false;
  long long cycles = 0;
  while (!s.finished && cycles < A.max_cycles) {
    ++cycles;
    t_slot = 0;
    stamp();
    const double rfs = rf_cur;
    const double cf2 = 2.0 * A.cf;
    const double fk = s.first ? pk[0] : fma(s.stp, pk[2], pk[0]);       // trial forcing
    // ---- trial point (lb_trial_kernel)
    if (!(spec && !s.first && s.stp == 1.0)) {
      const bool fst = s.first != 0;
      const double stp = s.stp;
      for (int e = tid; e < nloc; e += RNT) {
        const double v = fst ? X[e] : fma(stp, Dv[e], X[e]);
        XT[e] = v;
        if (prevH != nullptr && e < D) prevH[e] = v;
        if (nextH != nullptr && e >= e_next) nextH[e] = v;
      }
      cl.sync();
    }
    spec = false;
    stamp();                                            // 1: trial point, halo rows pushed, barrier
    // ---- f of the rows R0-2 .. R1 (the rows beside the slice are recomputed, not exchanged)
    for (int e = tid; e < (RPC + 3) * D; e += RNT) {
      const int lr = e / D, i = e - lr * D;
      const int r = R0 - 2 + lr;
      double v = 0.0;
      if (r >= 0 && r < N && r <= R1 && nrow > 0)
        v = xt_at(r, i - 1) * (xt_at(r, i + 1) - xt_at(r, i - 2)) - xt_at(r, i) + fk;
      F[e] = v;
    }
    __syncthreads();
    // ---- residuals -> seeds
    double acc_fe = 0.0;
    if (DISC == DISC_SIMPSON) {
      // pair k = rows (a, a+1, a+2), a even; slot (a - (R0 - 2)) / 2; needed: a = R0-2, R0, ..., R1-2
      const double dt3 = dt / 3.0, dt4 = dt / 4.0;
      const int npair = nrow > 0 ? (nrow + 1) / 2 + 1 : 0;          // slots of every pair an own row belongs to
      for (int e = tid; e < npair * D; e += RNT) {
        const int kp = e / D, i = e - kp * D;
        const int a = R0 - 2 + 2 * kp;
        double l1 = 0.0, l2 = 0.0;
        if (a >= 0 && a + 2 <= N - 1) {
          const double xa = xt_at(a, i), xb = xt_at(a + 1, i), xc = xt_at(a + 2, i);
          const double fa = f_at(a, i), fb = f_at(a + 1, i), fc = f_at(a + 2, i);
          const double e1 = xc - xa - dt3 * (fa + 4.0 * fb + fc);
          const double e2 = xb - 0.5 * (xa + xc) - dt4 * (fa - fc);
          l1 = cf2 * rfs * e1;
          l2 = cf2 * rfs * e2;
          if (a >= R0) acc_fe = fma(rfs * e1, e1, fma(rfs * e2, e2, acc_fe));      // own pair
        }
        E1[e] = l1;
        E2[e] = l2;
      }
    } else {
      // residual mm between rows mm and mm+1; slot mm - (R0 - 1); needed: mm = R0-1 .. R1-1
      const int nres = nrow > 0 ? nrow + 1 : 0;
      for (int e = tid; e < nres * D; e += RNT) {
        const int kr = e / D, i = e - kr * D;
        const int mm = R0 - 1 + kr;
        double l = 0.0;
        if (mm >= 0 && mm + 1 <= N - 1) {
          const double ev = xt_at(mm + 1, i) - al * xt_at(mm, i) - (ca * f_at(mm, i) + cb * f_at(mm + 1, i));
          l = cf2 * rfs * ev;
          if (mm >= R0) acc_fe = fma(rfs * ev, ev, acc_fe);
        }
        E1[e] = l;
      }
    }
    __syncthreads();
    // ---- gradient rows: direct terms, measurement term, J^T(x_r) V_r; forcing gradient -sum V
    double acc_me = 0.0, acc_pk = 0.0, acc_gd = 0.0, acc_mx = 0.0;
    const bool firstev = s.first != 0;
    for (int e = tid; e < nloc; e += RNT) {
      const int lr = e / D, i = e - lr * D;
      const int r = R0 + lr;
      // V_r at components i-2 .. i+2 and the direct term at i
      double V[5], dir;
      if (DISC == DISC_SIMPSON) {
        const double dt3 = dt / 3.0, dt4 = dt / 4.0;
        if (r & 1) {                                    // row b of pair a = r - 1
          const int kp = (r - 1 - (R0 - 2)) / 2;
#pragma unroll
          for (int q = 0; q < 5; ++q) { int j = i - 2 + q; if (j < 0) j += D; else if (j >= D) j -= D; V[q] = (4.0 * dt3) * E1[kp * D + j]; }
          dir = E2[kp * D + i];
        } else {                                        // row a of pair r, row c of pair r - 2
          const int ka = (r - (R0 - 2)) / 2, kc = ka - 1;
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            int j = i - 2 + q; if (j < 0) j += D; else if (j >= D) j -= D;
            V[q] = (dt3 * E1[ka * D + j] + dt4 * E2[ka * D + j]) + (dt3 * E1[kc * D + j] - dt4 * E2[kc * D + j]);
          }
          dir = (-E1[ka * D + i] - 0.5 * E2[ka * D + i]) + (E1[kc * D + i] - 0.5 * E2[kc * D + i]);
        }
      } else {
        const int kr = r - (R0 - 1);                    // residual r; residual r - 1 is slot kr - 1
#pragma unroll
        for (int q = 0; q < 5; ++q) { int j = i - 2 + q; if (j < 0) j += D; else if (j >= D) j -= D; V[q] = cb * E1[(kr - 1) * D + j] + ca * E1[kr * D + j]; }
        dir = E1[(kr - 1) * D + i] - al * E1[kr * D + i];
      }
      const double xm2 = xt_at(r, i - 2), xm1 = xt_at(r, i - 1), x0 = XT[e], xp1 = xt_at(r, i + 1), xp2 = xt_at(r, i + 2);
      // (J^T v)_i = v_{i+1} (x_{i+2} - x_{i-1}) + v_{i-1} x_{i-2} - v_{i+2} x_{i+1} - v_i
      const double jt = V[3] * (xp2 - xm1) + V[1] * xm2 - V[4] * xp1 - V[2];
      double g = dir - jt;
      {
        const double w = Wl[e];
        if (w != 0.0) {
          const double diff = x0 - Yl[e];
          const double wd = w * diff;
          acc_me = fma(wd, diff, acc_me);
          g += wd;
        }
      }
      GT[e] = g;
      acc_pk += V[2];
      if (!firstev) acc_gd = fma(g, Dv[e], acc_gd);
      acc_mx = fmax(acc_mx, fabs(g));
    }
    {
      double v[5] = {acc_fe, acc_me, acc_pk, acc_gd, acc_mx};
      const int op5[5] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_MAX};
      stamp();                                          // 2: evaluation (halo, f, seeds, gradient rows)
      res_reduce<5>(v, op5, partA, scratch);
    }
    cl.sync();
    stamp();                                            // 3: partial sums + barrier
    // ---- totals in rank order and the line-search decision (thread 0 of every CTA, identically)
    if (tid < 5) tot[tid] = (tid == 4) ? gather_max(partA, 4) : gather_sum(partA, tid);
    __syncthreads();
    if (tid == 0) {
      const double fetot = tot[0] * A.cf;
      const double metot = 0.5 * tot[1];
      const double gk = A.k_est ? -tot[2] : 0.0;
      double gdtot = tot[3];
      double sbg = tot[4];
      if (A.k_est) {
        if (!firstev) gdtot = fma(gk, pk[2], gdtot);
        sbg = fmax(sbg, fabs(gk));
      }
      pk[3] = fk;
      pk[4] = gk;
      if (!s.done && s.need_eval) lb_linesearch_body(s, &dummy_act, metot + fetot, metot, fetot, gdtot, sbg, A.o);
    }
    __syncthreads();
    stamp();                                            // 4: totals + line-search decision

    // ---- accepted step: x <- xt, g <- gt, new pair, dot products (lb_update_kernel)
    const bool ph2 = s.accepted != 0;
    const bool ph3 = (s.accepted || s.redo_dir) && !s.done;
    const bool upd = s.do_update != 0;
    const int p = s.pslot;
    const double stp = s.stp;
    if (ph2) {
      const int col = s.col;
      // Dot products by warp: warp w owns the history slots w and w + 8 (five products each: g.s_j,
      // g.y_j, s.y_j, y.s_j, y.y_j), warp 7 also y.y, g.s, g.y, g.g; a lane walks the own elements with
      // stride 32 and the 16 sums of a warp are reduced by one halving butterfly -- no pass through
      // shared memory and no combination across warps.
      const int lane = tid & 31, warp = tid >> 5;
      const int j0 = warp, j1 = warp + 8;
      const bool u0 = j0 < m && (j0 < col || col == m) && !(upd && j0 == p);
      const bool u1 = j1 < m && (j1 < col || col == m) && !(upd && j1 == p);
      const double* S0 = vec + (size_t)(RV_S + j0) * VS;
      const double* Y0 = vec + (size_t)(RV_Y + j0) * VS;
      const double* S1 = vec + (size_t)(RV_S + (j1 < MMAX ? j1 : 0)) * VS;
      const double* Y1 = vec + (size_t)(RV_Y + (j1 < MMAX ? j1 : 0)) * VS;
      double a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.0;
      auto one = [&](double gt, double gold, double dv, double s0, double y0, double s1, double y1) {
        double sv = 0.0, yv = 0.0;
        if (upd) { sv = stp * dv; yv = gt - gold; }
        if (u0) {
          a[0] = fma(gt, s0, a[0]); a[1] = fma(gt, y0, a[1]); a[2] = fma(sv, y0, a[2]);
          a[3] = fma(yv, s0, a[3]); a[4] = fma(yv, y0, a[4]);
        }
        if (u1) {
          a[5] = fma(gt, s1, a[5]); a[6] = fma(gt, y1, a[6]); a[7] = fma(sv, y1, a[7]);
          a[8] = fma(yv, s1, a[8]); a[9] = fma(yv, y1, a[9]);
        }
        if (warp == 7) {
          a[10] = fma(yv, yv, a[10]); a[11] = fma(gt, sv, a[11]); a[12] = fma(gt, yv, a[12]); a[13] = fma(gt, gt, a[13]);
        }
      };
#pragma unroll 4
      for (int e = lane; e < nloc; e += 32)
        one(GT[e], G[e], Dv[e], u0 ? S0[e] : 0.0, u0 ? Y0[e] : 0.0, u1 ? S1[e] : 0.0, u1 ? Y1[e] : 0.0);
      if (A.k_est && last && lane == 0)                 // the forcing's terms are counted once, by the last CTA
        one(pk[4], pk[1], pk[2], pk[8 + j0], pk[8 + MMAX + j0], j1 < MMAX ? pk[8 + j1] : 0.0, j1 < MMAX ? pk[8 + MMAX + j1] : 0.0);
      const double tsum = warp_sum16(a);
      if (!(lane & 1)) {
        const int q = lane >> 1;                        // entry 0..15 of this warp
        if (q < 5) partB[q * MMAX + j0] = tsum;
        else if (q < 10) { if (j1 < MMAX) partB[(q - 5) * MMAX + j1] = tsum; }
        else if (q < 14 && warp == 7) partB[5 * MMAX + (q - 10)] = tsum;
      }
    }
    cl.sync();                                          // (also: every warp has read g and the history)
    stamp();                                            // 5: dot products + barrier
    // warps 0-1 gather the totals, then warp 0 runs the recursion while warps 1-7 commit the step
    // (x <- xt, g <- gt, the new pair into slot p): neither touches what the other reads
    {
      const int warp = tid >> 5;
      if (warp < 2) {
        if (ph3 && ph2 && tid < NACC_U) tot[tid] = gather_sum(partB, tid);
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (warp == 0 && ph3) res_gram_warp(s, tot, m, timing ? tacc : nullptr);
      }
      if (ph2 && warp >= 1) {
        for (int e = tid - 32; e < nloc; e += RNT - 32) {
          const double gt = GT[e];
          if (upd) {
            vec[(size_t)(RV_S + p) * VS + e] = stp * Dv[e];
            vec[(size_t)(RV_Y + p) * VS + e] = gt - G[e];
          }
          X[e] = XT[e];
          G[e] = gt;
        }
        if (A.k_est && tid == 32) {                     // every CTA updates its replica of the forcing
          if (upd) { pk[8 + p] = stp * pk[2]; pk[8 + MMAX + p] = pk[4] - pk[1]; }
          pk[0] = pk[3];
          pk[1] = pk[4];
        }
      }
      __syncthreads();
    }
    stamp();                                            // 6: totals, two-loop recursion | commit

    // ---- direction (lb_direction_kernel) and the start of the next search
    if (ph3) {
      const int col = s.col;
      double cs[MMAX], cy[MMAX];
#pragma unroll
      for (int j = 0; j < MMAX; ++j) { cs[j] = s.cs[j]; cy[j] = s.cy[j]; }
      const double cgc = s.cg, cd = s.cd;
      double v[3] = {0.0, 0.0, BIG};
      for (int e = tid; e < nloc; e += RNT) {
        const double g = G[e];
        double rr = cgc * g, ry = 0.0;                  // two chains: the S and the Y terms
#pragma unroll
        for (int j = 0; j < MMAX; ++j) {
          if (j < m && (j < col || col == m)) {
            rr = fma(cs[j], vec[(size_t)(RV_S + j) * VS + e], rr);
            ry = fma(cy[j], vec[(size_t)(RV_Y + j) * VS + e], ry);
          }
        }
        double d = -(rr + ry);
        if (cd != 0.0) d = fma(cd, Dv[e], d);
        Dv[e] = d;
        {                                               // the trial point x + d of the next cycle
          const double v = fma(1.0, d, X[e]);
          XT[e] = v;
          if (prevH != nullptr && e < D) prevH[e] = v;
          if (nextH != nullptr && e >= e_next) nextH[e] = v;
        }
        v[0] = fma(d, d, v[0]);
        v[1] = fma(g, d, v[1]);
      }
      double dkk = 0.0, gkk = 0.0;
      if (A.k_est) {                                    // the forcing's entry: every CTA computes it; the last one counts it
        gkk = pk[1];
        double rr = cgc * gkk;
#pragma unroll
        for (int j = 0; j < MMAX; ++j) {
          if (j < m && (j < col || col == m)) {
            rr = fma(cs[j], pk[8 + j], rr);
            rr = fma(cy[j], pk[8 + MMAX + j], rr);
          }
        }
        dkk = -rr;
        if (cd != 0.0) dkk = fma(cd, pk[2], dkk);
        if (last && tid == 0) { v[0] = fma(dkk, dkk, v[0]); v[1] = fma(gkk, dkk, v[1]); }
      }
      const int op3[3] = {RED_SUM, RED_SUM, RED_MIN};
      res_reduce<3>(v, op3, partC, scratch + 64);
      __syncthreads();
      if (A.k_est && tid == 0) pk[2] = dkk;
    }
    spec = ph3 && !s.abort_dir;                         // x + d is in place (thread 0 clears abort_dir behind the barrier)
    cl.sync();
    stamp();                                            // 7: direction pass + barrier
    if (tid == 0) {
      if (!ph3) { if (s.done) { s.accepted = 0; s.redo_dir = 0; s.restore = 0.0; } }
      else if (s.abort_dir) s.abort_dir = 0;
      else {
        double dd = 0.0, gd = 0.0;
        dd = gather_sum(partC, 0);
        gd = gather_sum(partC, 1);
        lb_start_body(s, &dummy_act, dd, gd, BIG, A.o, 0, 0, 0);
      }
    }
    __syncthreads();
    stamp();                                            // 8: start of the next search

    // ---- end of a rung: minimiser to minpaths, table row, next rung
    if (s.done && !s.finished) {
      if (A.L.minpaths != nullptr) {
        double* out = A.L.minpaths + ((long long)b * A.L.Nbeta + s.ib) * A.L.mp_pitch;
        if (A.L.winw != n) {
          for (int e = tid; e < nloc; e += RNT) {
            const long long i = (long long)R0 * D + e;
            if (i >= A.L.win0 && i < A.L.win0 + A.L.winw) out[i - A.L.win0] = X[e];
          }
          if (A.k_est && last && tid == 0 && nX >= A.L.win0 && nX < A.L.win0 + A.L.winw) out[nX - A.L.win0] = pk[0];
        } else {
          for (int e = tid; e < nloc; e += RNT) out[(long long)R0 * D + e] = X[e];
          if (A.k_est && last && tid == 0) out[nX] = pk[0];
        }
      }
      __syncthreads();
      if (tid == 0) {
        if (rank == 0) {
          lb_advance_body(s, &dummy_act, 0, LbLadder{A.L.Nbeta, A.L.scales, A.L.betas, A.L.rf_path + b,
                                                    A.L.table ? A.L.table + (long long)b * A.L.Nbeta * 5 : nullptr, nullptr, 0, 0, 0,
                                                    A.L.status2 ? A.L.status2 + (long long)b * A.L.Nbeta : nullptr,
                                                    A.L.nit2 ? A.L.nit2 + (long long)b * A.L.Nbeta : nullptr,
                                                    A.L.nfev2 ? A.L.nfev2 + (long long)b * A.L.Nbeta : nullptr}, A.o);
        } else {
          double rfdummy[1];
          lb_advance_body(s, &dummy_act, 0, LbLadder{A.L.Nbeta, A.L.scales, A.L.betas, rfdummy, nullptr, nullptr, 0, 0, 0,
                                                    nullptr, nullptr, nullptr}, A.o);
        }
      }
      if (tid == 0) rf_cur = A.rf0 * A.L.scales[s.ib];
      __syncthreads();
    }
    stamp();                                            // 9: end of rung
    if (timing) tacc[0] += 1;
  }
  if (timing)
    for (int k = 0; k < 16; ++k) A.dbg[k] += tacc[k];
  // ---- results back to global memory
  for (int e = tid; e < nloc; e += RNT) xg[(long long)R0 * D + e] = X[e];
  if (A.k_est && last && tid == 0) xg[nX] = pk[0];
  if (rank == 0) {
    __syncthreads();
    if (tid == 0 && !s.finished) { s.status = 2; s.finished = 1; }        // cycle cap hit (internal error)
    __syncthreads();
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&s);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(A.st + b);
    for (int k = tid; k < (int)(sizeof(LbPath) / 8); k += RNT) dst[k] = src[k];
  }
  cl.sync();
}

typedef void (*ResKernel)(const ResArgs);
template <int RCS>
inline ResKernel lb_resident_pick(int disc) {
  switch (disc) {
    case VAB_DISC_EULER: return lb_resident_kernel<DISC_EULER, RCS>;
    case VAB_DISC_TRAPEZOID: return lb_resident_kernel<DISC_TRAPEZOID, RCS>;
    case VAB_DISC_SIMPSON_HERMITE: return lb_resident_kernel<DISC_SIMPSON, RCS>;
    default: return lb_resident_kernel<DISC_FORWARDMAP, RCS>;
  }
}

inline int lb_resident_rows(int N, int cs) {            // rows per CTA: even, at least 2
  int r = (((N + cs - 1) / cs) + 1) & ~1;
  return r < 2 ? 2 : r;
}

inline size_t lb_resident_smem(int RPC, int D) {
  return ((size_t)RV_N * RPC * D + 4 * (size_t)(RPC + 3) * D + 8 * NT + 2 * NACC_U + 2 * (size_t)RPC * D) * sizeof(double);
}

}  // namespace
