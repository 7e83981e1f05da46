// Fused ODE action + adjoint gradient: the TMA-fed streaming kernels (the hot path for Lorenz96).
//
// Same mathematics and the same (path, time segment, column window) work units as the register
// sweep kernels in ode_sweep.cuh, with the data movement rebuilt around Blackwell's bulk-copy
// engine:
//   * every warp owns a private ring of NS stages in shared memory.  A stage holds two
//     consecutive time rows of the warp's window of X (for each of the GPW paths the warp works
//     on) and the same two rows of the dense observation matrix Y.  One lane per path issues
//     `cp.async.bulk` (TMA, 1-D) copies global -> shared that complete on the stage's mbarrier;
//     the other lanes only wait on the barrier.  No registers are spent on prefetching, loads run
//     NS-2 stages (2(NS-2) rows) ahead of the compute, and the measurement data arrives with the
//     state instead of stalling the gradient on an L2 round trip.
//   * lanes read their own strip and its two-component halo straight from the staged row
//     (LDS.128 at precomputed 32-bit shared addresses); only the adjoint seed v is exchanged
//     between lanes, with warp shuffles.
//   * all paths of a warp share (segment, window), so row validity is warp-uniform: the steady
//     state runs without per-lane predicates except the store mask; ragged ends are peeled.
//   * there is no __syncthreads anywhere; warps run decoupled.
// Gradient rows are written once with streaming 16-byte stores; per-unit partial sums are reduced
// in a fixed order (bit-reproducible).  Reference semantics: va_ode.py:130-234 (action),
// :341-380, :404-454 (discretisations); adjoint formulas SURVEY.md App. A.3.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "ode_models.cuh"
#include "ode_params.h"
#include "vab_hd.h"
#include "vab_tma.cuh"

#ifndef VAB_FULL
#define VAB_FULL 0xffffffffu
#endif

namespace vabs {

__device__ __forceinline__ void lds2(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

// MODE 1 (FAST): nskip == 1, scalar RM, scalar RF (the common case), the RF weight a launch
// constant; MODE 2: the same with one RF per path (P.rf_path, asynchronous ladder: the weight
// becomes a per-lane register); MODE 0: every weight / row test is looked up at run time.
template <class M, int NS, int MODE>
struct Stream {
  static constexpr bool FAST = (MODE != 0), PERPATH = (MODE == 2);
  static constexpr int C = M::C, H = M::H, W = M::C + 2 * M::H, NPM = M::NPM;
  static_assert(H == 2 && (C == 4 || C == 2), "stream kernels: Lorenz96-type stencil, strips of 4 or 2");
  const OdeParams& P;
  // warp-uniform
  int sg, w, r0, r1, N, rowbase, qmax;
  bool wrap;
  uint32_t ring0, bar0;        // shared addresses: this warp's first stage / first barrier
  uint32_t stage_b, row_b;     // bytes per stage (all paths of the warp) / per staged window row
  uint32_t full_bytes;         // bytes a full stage brings in, all paths of the warp
  // per lane
  int g, j, bidx, i0, D;
  bool pact, out, leader, bvalid;
  bool swz;                    // lanes 4..7 of every quarter-warp read their 16-byte pieces in the
                               // opposite order: conflict-free LDS.128 at a 32-byte lane stride
  uint32_t a_o1, a_o2;         // shared addresses (stage 0, row 0) of the two halves of the own strip
  uint32_t a_h1, a_h2;         // ... of the left / right halo pair (swapped when sw)
  int srcM1, srcP1;
  const double* xsrc;          // leader: first global element of this path's window
  const double* ysrc;
  double* grow;                // &G[b][0][i0]
  double p[NPM];
  double wob[C];               // 2 cm RM of the own components (0 = unobserved)
  double wfs;                  // 2 cf RF (scalar RF)
  uint32_t a_wfs;              // MODE 2: shared address of this lane's copy of wfs (a live 64-bit
                               // value would not fit the 128-register budget: it is re-read at use)
  double rsc;                  // scale of an RF0 array (per path under the asynchronous ladder)
  double me_acc, fe_acc, pacc[NPM];

  __device__ __forceinline__ Stream(const OdeParams& P_) : P(P_) {}

  // returns false if this warp has no work
  __device__ __forceinline__ bool init(double* smem) {
    const int warp = __shfl_sync(VAB_FULL, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * 4 + warp;
    const int sw = gwarp / P.wpb, bgrp = gwarp - sw * P.wpb;
    if (sw >= P.nseg * P.nwin) return false;
    sg = sw / P.nwin;
    w = sw - sg * P.nwin;
    N = P.N;
    D = P.D;
    r0 = sg * P.Tseg;
    r1 = min(r0 + P.Tseg, N);
    wrap = (P.NHL == 0);
    const int Wd = P.GW * C;                               // doubles per staged window row
    row_b = (uint32_t)Wd * 8u;
    const uint32_t grp_b = 4u * row_b;                     // X rows a, b then Y rows a, b
    stage_b = (uint32_t)P.GPW * grp_b;
    ring0 = s32(smem) + (uint32_t)warp * NS * stage_b;
    bar0 = s32(smem) + 4u * NS * stage_b + (uint32_t)warp * NS * 8u;
    int gg = lane / P.GW;
    j = lane - gg * P.GW;
    const bool ingroup = gg < P.GPW;
    int b = bgrp * P.GPW + gg;
    bidx = b;
    bvalid = ingroup && b < P.B;
    pact = bvalid && (P.active == nullptr || __ldg(P.active + b) != 0);
    if (!bvalid) b = 0;
    if (!ingroup) gg = 0;
    g = gg;
    leader = pact && j == 0;
    const int nact = __popc(__ballot_sync(VAB_FULL, leader));
    full_bytes = (uint32_t)nact * grp_b;
    const int TPR = P.TPR;
    int st, st0 = 0, jo = j * C;
    uint32_t o_l, o_r;
    if (wrap) {
      st = j;
      out = pact;
      o_l = (uint32_t)((jo - 2 + Wd) % Wd);
      o_r = (uint32_t)((jo + C) % Wd);
      srcM1 = gg * P.GW + (j + P.GW - 1) % P.GW;
      srcP1 = gg * P.GW + (j + 1) % P.GW;
    } else {
      const int rel = w * P.WS - P.NHL;
      st0 = ((rel % TPR) + TPR) % TPR;
      st = (st0 + j) % TPR;
      out = pact && j >= P.NHL && j < P.NHL + P.WS && (w * P.WS + j - P.NHL) < TPR;
      o_l = (uint32_t)max(jo - 2, 0);
      o_r = (uint32_t)min(jo + C, Wd - 2);
      srcM1 = gg * P.GW + max(j - 1, 0);
      srcP1 = gg * P.GW + min(j + 1, P.GW - 1);
    }
    if (!ingroup) { srcM1 = srcP1 = lane; jo = 0; o_l = o_r = 0; }
    const uint32_t gbase = ring0 + (uint32_t)gg * grp_b;
    swz = (C == 4) && ((lane >> 2) & 1);
    const uint32_t a_own = gbase + (uint32_t)jo * 8u, a_l = gbase + o_l * 8u, a_r = gbase + o_r * 8u;
    a_o1 = a_own + (swz ? 16u : 0u);
    a_o2 = a_own + (swz ? 0u : 16u);
    a_h1 = swz ? a_r : a_l;
    a_h2 = swz ? a_l : a_r;
    i0 = st * C;
    const double* xpath = P.XP + (long long)b * P.ldxp;
    xsrc = xpath + st0 * C;
    ysrc = P.Y + st0 * C;
    grow = (out && P.G != nullptr) ? P.G + (long long)b * P.ldg + i0 : nullptr;
    const long long nX = (long long)N * D;
#pragma unroll
    for (int k = 0; k < NPM; ++k) {
      double v = 0.0;
      if (pact) {
        const int e = __ldg(P.pmap + k);
        v = (e >= 0) ? __ldg(xpath + nX + e) : __ldg(P.pfix + (long long)b * P.pfix_stride + k);
      }
      p[k] = v;
      pacc[k] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) wob[c] = out ? __ldg(P.wobs + i0 + c) : 0.0;
    if constexpr (MODE == 1) {
      rsc = P.rf_scale;
      wfs = 2.0 * P.cf * P.rf_scalar;
    } else {
      rsc = (P.rf_path != nullptr) ? __ldg(P.rf_path + b) : P.rf_scale;
      wfs = 2.0 * P.cf * ((P.rf_path != nullptr) ? P.rf0 * rsc : P.rf_scalar);
    }
    a_wfs = 0;
    if constexpr (MODE == 2) {
      // slot behind the ring, the barriers and the reduction scratch (128 K doubles)
      a_wfs = s32(smem) + 4u * NS * stage_b + 4u * NS * 8u + (uint32_t)(128 * P.K) * 8u + (uint32_t)threadIdx.x * 8u;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(a_wfs), "d"(wfs) : "memory");
    }
    me_acc = 0.0;
    fe_acc = 0.0;
    // zero this warp's ring: rows that are never copied (path ends, rows without observations)
    // must read as finite numbers
    for (uint32_t o = lane * 16u; o < NS * stage_b; o += 32u * 16u)
      asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(ring0 + o), "d"(0.0) : "memory");
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8u * s, 1);
      mbar_fence_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero fill before TMA writes
    __syncwarp();
    return true;
  }

  __device__ __forceinline__ uint32_t slot_off(int q) const { return (uint32_t)(q % NS) * stage_b; }
  __device__ __forceinline__ bool obs_row(int r) const { return FAST || P.nskip == 1 || (r % P.nskip) == 0; }
  __device__ __forceinline__ int obs_index(int r) const { return FAST ? r : r / P.nskip; }

  // Enqueue the copies of stage q: rows rowbase + 2q and rowbase + 2q + 1.  All lanes call it.
  __device__ __forceinline__ void issue(int q) {
    if (q > qmax) return;
    const int ra = rowbase + 2 * q, rb = ra + 1;
    const uint32_t bar = bar0 + 8u * (q % NS);
    const uint32_t so = slot_off(q);
    if (FAST && wrap && ra >= 0 && rb < N) {       // steady state: two valid, contiguous rows
      if ((threadIdx.x & 31) == 0) mbar_expect_tx(bar, full_bytes);
      __syncwarp();
      if (leader) {
        const uint32_t dst = ring0 + (uint32_t)g * 4u * row_b + so;   // the path's stage base
        const long long off = (long long)ra * D;
        tma_load(dst, xsrc + off, 2u * row_b, bar);
        tma_load(dst + 2u * row_b, ysrc + off, 2u * row_b, bar);
      }
      return;
    }
    const bool va = ra >= 0 && ra < N, vb = rb >= 0 && rb < N;
    const bool ya = va && P.L > 0 && obs_row(ra), yb = vb && P.L > 0 && obs_row(rb);
    const uint32_t per = (uint32_t)((va ? 1 : 0) + (vb ? 1 : 0) + (ya ? 1 : 0) + (yb ? 1 : 0)) * row_b;
    const int nact = (int)(full_bytes / (4u * row_b));
    if ((threadIdx.x & 31) == 0) mbar_expect_tx(bar, per * (uint32_t)nact);
    __syncwarp();
    if (leader) {
      const uint32_t dst = ring0 + (uint32_t)g * 4u * row_b + so;
      // a window that runs past the end of the row continues at its beginning (periodic model)
      const int st0 = (int)((xsrc - (P.XP + (long long)bidx * P.ldxp)) / C);
      const int n1 = wrap ? P.GW : min(P.GW, P.TPR - st0);
      const uint32_t b1 = (uint32_t)n1 * C * 8u, b2 = row_b - b1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = ra + h;
        if (h == 0 ? va : vb) {
          const double* src = xsrc + (long long)r * D;
          tma_load(dst + h * row_b, src, b1, bar);
          if (b2) tma_load(dst + h * row_b + b1, src - st0 * C, b2, bar);
        }
        if (h == 0 ? ya : yb) {
          const double* src = ysrc + (long long)obs_index(r) * D;
          tma_load(dst + (2 + h) * row_b, src, b1, bar);
          if (b2) tma_load(dst + (2 + h) * row_b + b1, src - st0 * C, b2, bar);
        }
      }
    }
  }
  __device__ __forceinline__ void wait(int q) const {
    if (q > qmax) return;
    mbar_wait(bar0 + 8u * (q % NS), (uint32_t)((q / NS) & 1));
  }

  // staged row (byte offset `off` from stage 0 / row 0) -> own strip with halo
  __device__ __forceinline__ void read_own(uint32_t off, double* x) const {
    if constexpr (C == 4) {
      double o0, o1, o2, o3;
      lds2(a_o1 + off, o0, o1);
      lds2(a_o2 + off, o2, o3);
      x[0] = swz ? o2 : o0; x[1] = swz ? o3 : o1;
      x[2] = swz ? o0 : o2; x[3] = swz ? o1 : o3;
    } else {
      lds2(a_o1 + off, x[0], x[1]);
    }
  }
  __device__ __forceinline__ void read_halo(uint32_t off, double* X) const {
    double h0, h1, h2, h3;
    lds2(a_h1 + off, h0, h1);
    lds2(a_h2 + off, h2, h3);
    X[0] = swz ? h2 : h0; X[1] = swz ? h3 : h1;
    X[H + C] = swz ? h0 : h2; X[H + C + 1] = swz ? h1 : h3;
  }
  __device__ __forceinline__ void read_row(uint32_t off, double* X) const {
    read_halo(off, X);
    read_own(off, X + H);
  }
  // own values in V[H..H+C): fetch the halo from the neighbour lanes (all lanes must call)
  __device__ __forceinline__ void halo(double* V) const {
    V[0] = __shfl_sync(VAB_FULL, V[H + C - 2], srcM1);
    V[1] = __shfl_sync(VAB_FULL, V[H + C - 1], srcM1);
    V[H + C] = __shfl_sync(VAB_FULL, V[H], srcP1);
    V[H + C + 1] = __shfl_sync(VAB_FULL, V[H + 1], srcP1);
  }
  __device__ __forceinline__ double wgt(int row, int c) const {   // 2 cf RF for residual (row, c)
    if constexpr (MODE == 2) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a_wfs));
      return v;
    }
    if (FAST) return wfs;
    return P.rf_arr ? 2.0 * P.cf * rsc * __ldg(P.rf_arr + (long long)row * D + i0 + c) : wfs;
  }
  // measurement term of row r (va_ode.py:138-158), y staged at byte offset yoff:
  // d += 2 cm RM (x - y);  me_acc += 2 cm RM (x - y)^2.  Branch-free: unobserved components have
  // weight 0 and a zero y.
  __device__ __forceinline__ void measure(int r, uint32_t yoff, const double* xown, double* d) {
    if (!FAST) {
      if (P.L == 0 || !obs_row(r)) return;                    // warp-uniform
    }
    double y[C];
    read_own(yoff, y);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double wo = wob[c];
      if (!FAST) {
        if (P.rmd != nullptr) wo = out ? __ldg(P.rmd + (long long)obs_index(r) * D + i0 + c) : 0.0;
      }
      const double diff = xown[c] - y[c];
      const double wd = wo * diff;
      me_acc = fma(wd, diff, me_acc);
      d[c] += wd;
    }
  }
  __device__ __forceinline__ void store(int r, const double* gr) const {
    if (grow != nullptr) vab_store_strip<C>(grow + (long long)r * D, gr);
  }
  __device__ __forceinline__ void finish(double* red) {
    if (!out) {
      me_acc = 0.0;
      fe_acc = 0.0;
#pragma unroll
      for (int k = 0; k < NPM; ++k) pacc[k] = 0.0;
    }
    double* mine = red + (size_t)threadIdx.x * P.K;
    mine[0] = 0.5 * me_acc;                       // me_acc holds sum 2 cm RM diff^2
    mine[1] = 0.5 * fe_acc;                       // fe_acc holds sum lam e = 2 cf RF e^2
#pragma unroll
    for (int k = 0; k < NPM; ++k) mine[2 + k] = -pacc[k];
    __syncwarp();
    if (bvalid) {
      const long long unit = (long long)bidx * P.upp + (long long)sg * P.nwin + w;
      const double* grp = red + (size_t)(threadIdx.x - j) * P.K;
      for (int k = j; k < P.K; k += P.GW) {
        double acc = 0.0;
        for (int s = 0; s < P.GW; ++s) acc += grp[(size_t)s * P.K + k];
        P.partials[unit * P.K + k] = acc;
      }
    }
  }
};

}  // namespace vabs

// --------------------------------------------------------------------------------------------
// Simpson-Hermite (va_ode.py:404-437 + :192-195).  Pair k = rows (a, b, c) = (2k, 2k+1, 2k+2):
//   e1 = x_c - x_a - dt/3 (f_a + 4 f_b + f_c),  e2 = x_b - (x_a + x_c)/2 - dt/4 (f_a - f_c)
// One step per pair = one ring stage (rows b, c).  Rows a and b get their gradient in the step of
// their pair; row c's partial seed is carried into the next pair, where it is row a.  Row a itself
// is re-read from the previous stage (kept resident one step longer) instead of being carried.
template <class M, int NS, int MINB, int MODE>
__global__ void __launch_bounds__(128, MINB) stream_simpson_kernel(const __grid_constant__ OdeParams P) {
  using ST = vabs::Stream<M, NS, MODE>;
  constexpr int C = ST::C, H = ST::H, W = ST::W;
  extern __shared__ __align__(16) double smem[];
  ST S(P);
  vab_pdl_trigger();          // the finalize kernel may be scheduled behind this one right away
  if (!S.init(smem)) return;
  double* red = smem + ((size_t)4 * NS * S.stage_b) / 8 + 4 * NS;
  const double dt = P.dt, dt3 = dt / 3.0, dt4 = dt / 4.0, dt43 = 4.0 * dt / 3.0;
  const int r0 = S.r0, N = S.N;
  const uint32_t rb = S.row_b;
  const bool last = (S.r1 == N);
  const int nfull = last ? (N - 1 - r0) / 2 : P.Tseg / 2;
  S.rowbase = r0 - 3;                       // stage q holds rows (r0 - 3 + 2q, r0 - 2 + 2q)
  S.qmax = nfull + 1;
#pragma unroll
  for (int q = 0; q < NS; ++q) S.issue(q);

  double xa[C], Fa[C], vcp[C], dcp[C];
  // one pair: sq = slot offset of its stage (rows b, c), sp = slot offset of the previous stage
  // (whose second row is row a).  STORE = false for the lead-in pair.
  auto pair = [&](int a, uint32_t sq, uint32_t sp, auto store_tag) {
    constexpr bool STORE = decltype(store_tag)::value;
    double Xb[W], l1[C], l2[C], Fc[C], xc[C];
    {
      double Xc[W], Fb[C];
      S.read_row(sq, Xb);
      S.read_row(sq + rb, Xc);
      M::f(Xb, S.p, nullptr, Fb);
      M::f(Xc, S.p, nullptr, Fc);
      double wq = 0.0;                      // scalar-RF modes: one weight for the whole pair
      if constexpr (ST::FAST) wq = S.wgt(a, 0);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        xc[c] = Xc[H + c];
        const double s = fma(4.0, Fb[c], Fa[c]) + Fc[c];
        const double e1 = fma(-dt3, s, xc[c] - xa[c]);
        const double e2 = fma(-dt4, Fa[c] - Fc[c], fma(-0.5, xa[c] + xc[c], Xb[H + c]));
        l1[c] = (ST::FAST ? wq : S.wgt(a, c)) * e1;
        l2[c] = (ST::FAST ? wq : S.wgt(a + 1, c)) * e2;
        if (STORE) S.fe_acc = fma(l1[c], e1, fma(l2[c], e2, S.fe_acc));
      }
    }
    if (STORE) {
      {   // row b:  v = 4dt/3 l1,  d = l2
        double V[W], gr[C], t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { V[H + c] = dt43 * l1[c]; gr[c] = l2[c]; }
        S.halo(V);
        S.measure(a + 1, sq + 2u * rb, Xb + H, gr);
        M::adj_t(Xb, V, S.p, t, S.pacc);
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = (gr[c] + V[H + c]) - t[c];
        S.store(a + 1, gr);
      }
      {   // row a:  v = carried c-part + dt/3 l1 + dt/4 l2,  d = carried - l1 - l2/2
        double Xa[W], V[W], gr[C], t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          V[H + c] = fma(dt4, l2[c], fma(dt3, l1[c], vcp[c]));
          gr[c] = fma(-0.5, l2[c], dcp[c] - l1[c]);
        }
        S.halo(V);
        S.read_halo(sp + rb, Xa);
#pragma unroll
        for (int c = 0; c < C; ++c) Xa[H + c] = xa[c];
        S.measure(a, sp + 3u * rb, xa, gr);
        M::adj_t(Xa, V, S.p, t, S.pacc);
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = (gr[c] + V[H + c]) - t[c];
        S.store(a, gr);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      vcp[c] = fma(-dt4, l2[c], dt3 * l1[c]);
      dcp[c] = fma(-0.5, l2[c], l1[c]);
      Fa[c] = Fc[c];
      xa[c] = xc[c];
    }
  };

  // ---- lead-in: pair (r0-2, r0-1, r0), or just row 0 when the segment starts the path
  S.wait(0);
  S.wait(1);
  {
    double Xa[W];
    if (r0 >= 2) {
      S.read_row(S.slot_off(0) + rb, Xa);
      M::f(Xa, S.p, nullptr, Fa);
#pragma unroll
      for (int c = 0; c < C; ++c) { xa[c] = Xa[H + c]; vcp[c] = 0.0; dcp[c] = 0.0; }
      pair(r0 - 2, S.slot_off(1), S.slot_off(0), std::false_type{});
    } else {
      S.read_row(S.slot_off(1) + rb, Xa);
      M::f(Xa, S.p, nullptr, Fa);
#pragma unroll
      for (int c = 0; c < C; ++c) { xa[c] = Xa[H + c]; vcp[c] = 0.0; dcp[c] = 0.0; }
    }
  }
  __syncwarp();
  S.issue(NS);
  // ---- steady state
  for (int t = 1; t <= nfull; ++t) {
    const int q = t + 1;
    S.wait(q);
    pair(r0 + 2 * (t - 1), S.slot_off(q), S.slot_off(q - 1), std::true_type{});
    __syncwarp();
    S.issue(q - 1 + NS);
  }
  // ---- the last row of the path (even, only the c-part of the last pair and its measurement)
  if (last) {
    double Xa[W], V[W], gr[C], t[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { V[H + c] = vcp[c]; gr[c] = dcp[c]; }
    S.halo(V);
    const uint32_t sl = S.slot_off(nfull + 1);
    S.read_halo(sl + rb, Xa);
#pragma unroll
    for (int c = 0; c < C; ++c) Xa[H + c] = xa[c];
    S.measure(N - 1, sl + 3u * rb, xa, gr);
    M::adj_t(Xa, V, S.p, t, S.pacc);
#pragma unroll
    for (int c = 0; c < C; ++c) gr[c] = (gr[c] + V[H + c]) - t[c];
    S.store(N - 1, gr);
  }
  vab_pdl_wait();             // partials of the previous evaluation have been consumed by its finalize kernel
  S.finish(red);
}

// --------------------------------------------------------------------------------------------
// euler / trapezoid / forwardmap (va_ode.py:341-380, 439-454):
//   e_m = x_{m+1} - AL x_m - (CA f_m + CB f_{m+1}),  lam = 2 cf RF e
//   g_r = [lam_{r-1} - AL lam_r] + meas_r - J^T(x_r) (CB lam_{r-1} + CA lam_r)
// One step = one ring stage = two arriving rows.
template <class M, int DISC, int NS, int MINB, int MODE>
__global__ void __launch_bounds__(128, MINB) stream_twopoint_kernel(const __grid_constant__ OdeParams P) {
  using ST = vabs::Stream<M, NS, MODE>;
  constexpr int C = ST::C, H = ST::H, W = ST::W;
  extern __shared__ __align__(16) double smem[];
  ST S(P);
  vab_pdl_trigger();          // the finalize kernel may be scheduled behind this one right away
  if (!S.init(smem)) return;
  double* red = smem + ((size_t)4 * NS * S.stage_b) / 8 + 4 * NS;
  const double dt = P.dt;
  const double ca = (DISC == DISC_EULER) ? dt : (DISC == DISC_TRAPEZOID ? 0.5 * dt : 1.0);
  const double cb = (DISC == DISC_TRAPEZOID) ? 0.5 * dt : 0.0;
  const double al = (DISC == DISC_FORWARDMAP) ? 0.0 : 1.0;
  const int r0 = S.r0, r1 = S.r1, N = S.N;
  const uint32_t rb = S.row_b;
  S.rowbase = r0 - 2;                       // stage q holds rows (r0 - 2 + 2q, r0 - 1 + 2q)
  S.qmax = (min(r1, N - 1) - r0 + 2) / 2;
#pragma unroll
  for (int q = 0; q < NS; ++q) S.issue(q);

  double X1[W], F1[C], lamp[C];
  bool valid1;
  // row m arrives (staged at byte offset xoff); finalises row m-1 (its Y row at yoff) when fin
  auto sub = [&](int m, uint32_t xoff, uint32_t yoff, bool vm, bool fin) {
    double Xm[W], Fm[C], lam[C];
    if (vm) {
      S.read_row(xoff, Xm);
      M::f(Xm, S.p, nullptr, Fm);
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) Xm[c] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) Fm[c] = 0.0;
    }
    const bool ve = vm && valid1;
    const bool own = (m - 1 >= r0) && (m - 1 < r1);
    double wq = 0.0;                        // scalar-RF modes: one weight for the whole row
    if constexpr (ST::FAST) wq = S.wgt(m - 1, 0);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double l = 0.0;
      if (ve) {
        const double e = Xm[H + c] - al * X1[H + c] - fma(ca, F1[c], cb * Fm[c]);
        l = (ST::FAST ? wq : S.wgt(m - 1, c)) * e;
        if (own) S.fe_acc = fma(l, e, S.fe_acc);
      }
      lam[c] = l;
    }
    if (fin) {
      double V[W], gr[C], t[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        V[H + c] = fma(cb, lamp[c], ca * lam[c]);
        gr[c] = lamp[c] - al * lam[c];
      }
      S.halo(V);
      S.measure(m - 1, yoff, X1 + H, gr);
      M::adj_t(X1, V, S.p, t, S.pacc);
#pragma unroll
      for (int c = 0; c < C; ++c) gr[c] = (gr[c] + V[H + c]) - t[c];
      S.store(m - 1, gr);
    }
#pragma unroll
    for (int c = 0; c < W; ++c) X1[c] = Xm[c];
#pragma unroll
    for (int c = 0; c < C; ++c) { F1[c] = Fm[c]; lamp[c] = lam[c]; }
    valid1 = vm;
  };

  // ---- prologue: row r0 - 1 (second row of stage 0)
  S.wait(0);
  valid1 = (r0 >= 1);
  if (valid1) {
    S.read_row(S.slot_off(0) + rb, X1);
    M::f(X1, S.p, nullptr, F1);
  } else {
#pragma unroll
    for (int c = 0; c < W; ++c) X1[c] = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) F1[c] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) lamp[c] = 0.0;
  const int tmax = (r1 - r0) / 2;                               // last step (m1 = r0 + 2 tmax <= r1)
  const int mtop = min(r1, N - 1);
  int tfull = (mtop - r0 - 1) / 2;                              // steps with both rows valid + final
  if (tfull > tmax) tfull = tmax;
  for (int t = 0; t <= tmax; ++t) {
    const int q = t + 1, m1 = r0 + 2 * t;
    S.wait(q);
    const uint32_t sq = S.slot_off(q), sp = S.slot_off(q - 1);
    if (t >= 1 && t <= tfull) {
      sub(m1, sq, sp + 3u * rb, true, true);
      sub(m1 + 1, sq + rb, sq + 2u * rb, true, true);
    } else {
      sub(m1, sq, sp + 3u * rb, m1 < N, t >= 1);
      if (m1 + 1 <= r1) sub(m1 + 1, sq + rb, sq + 2u * rb, m1 + 1 < N, true);
    }
    __syncwarp();
    S.issue(q - 1 + NS);
  }
  vab_pdl_wait();             // partials of the previous evaluation have been consumed by its finalize kernel
  S.finish(red);
}

// --------------------------------------------------------------------------------------------
// rk4 (extension; intent of the commented-out va_ode.py:383-402):
//   e_m = x_{m+1} - x_m - dt/6 (k1 + 2 k2 + 2 k3 + k4),  k1 = f(x_m), k2 = f(x_m + dt/2 k1), ...
// Same ring / row schedule as the two-point kernel: when row m arrives, residual m-1 is formed
// from the carried row m-1 (its x halo comes from the staged row; only the three stage states and
// the four adjoint seeds are exchanged by shuffles) and row m-1's gradient
//   g_{m-1} = lam_{m-2} - lam_{m-1} + meas_{m-1} + [adjoint of -dt/6 K(x_{m-1}) lam_{m-1}]
// is finalised by reverse accumulation through the stages (discrete adjoint of RK4).
template <class M, int NS, int MINB, int MODE>
__global__ void __launch_bounds__(128, MINB) stream_rk4_kernel(const __grid_constant__ OdeParams P) {
  using ST = vabs::Stream<M, NS, MODE>;
  constexpr int C = ST::C, H = ST::H, W = ST::W;
  extern __shared__ __align__(16) double smem[];
  ST S(P);
  vab_pdl_trigger();          // the finalize kernel may be scheduled behind this one right away
  if (!S.init(smem)) return;
  double* red = smem + ((size_t)4 * NS * S.stage_b) / 8 + 4 * NS;
  const double dt = P.dt, hdt = 0.5 * dt, dt6 = dt / 6.0, dt3 = dt / 3.0;
  const int r0 = S.r0, r1 = S.r1, N = S.N;
  const uint32_t rb = S.row_b;
  S.rowbase = r0 - 2;                       // stage q holds rows (r0 - 2 + 2q, r0 - 1 + 2q)
  S.qmax = (min(r1, N - 1) - r0 + 2) / 2;
#pragma unroll
  for (int q = 0; q < NS; ++q) S.issue(q);

  double X1[W], lamp[C];
  bool valid1;
  // row m arrives (staged at byte offset xoff); finalises row m-1 (its Y row at yoff) when fin
  auto sub = [&](int m, uint32_t xoff, uint32_t yoff, bool vm, bool fin) {
    double xm[C], lam[C];
    if (vm) {
      S.read_own(xoff, xm);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) xm[c] = 0.0;
    }
    const bool ve = vm && valid1;                 // residual m-1 exists
    const bool own = (m - 1 >= r0) && (m - 1 < r1);
    if (ve) {                                     // warp-uniform
      double Y2[W], Y3[W], Y4[W], k[C], ks[C];
      M::f(X1, S.p, nullptr, k);
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] = k[c]; Y2[H + c] = fma(hdt, k[c], X1[H + c]); }
      S.halo(Y2);
      M::f(Y2, S.p, nullptr, k);
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] = fma(2.0, k[c], ks[c]); Y3[H + c] = fma(hdt, k[c], X1[H + c]); }
      S.halo(Y3);
      M::f(Y3, S.p, nullptr, k);
#pragma unroll
      for (int c = 0; c < C; ++c) { ks[c] = fma(2.0, k[c], ks[c]); Y4[H + c] = fma(dt, k[c], X1[H + c]); }
      S.halo(Y4);
      M::f(Y4, S.p, nullptr, k);
      double wq = 0.0;                      // scalar-RF modes: one weight for the whole row
      if constexpr (ST::FAST) wq = S.wgt(m - 1, 0);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const double e = (xm[c] - X1[H + c]) - dt6 * (ks[c] + k[c]);
        const double l = (ST::FAST ? wq : S.wgt(m - 1, c)) * e;
        if (own) S.fe_acc = fma(l, e, S.fe_acc);
        lam[c] = l;
      }
      if (fin) {
        // reverse sweep through the stages: xb collects d e_{m-1} / d x_{m-1} applied to lam
        // (the seeds carry the minus sign of the residual, finish() negates pacc: collect locally)
        double KB[W], jt[C], xb[C], gr[C], pl[ST::NPM];
#pragma unroll
        for (int q = 0; q < ST::NPM; ++q) pl[q] = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) { xb[c] = -lam[c]; KB[H + c] = -dt6 * lam[c]; }
        S.halo(KB);
        M::adj(Y4, KB, S.p, jt, pl);
#pragma unroll
        for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = fma(dt, jt[c], -dt3 * lam[c]); }
        S.halo(KB);
        M::adj(Y3, KB, S.p, jt, pl);
#pragma unroll
        for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = fma(hdt, jt[c], -dt3 * lam[c]); }
        S.halo(KB);
        M::adj(Y2, KB, S.p, jt, pl);
#pragma unroll
        for (int c = 0; c < C; ++c) { xb[c] += jt[c]; KB[H + c] = fma(hdt, jt[c], -dt6 * lam[c]); }
        S.halo(KB);
        M::adj(X1, KB, S.p, jt, pl);
#pragma unroll
        for (int q = 0; q < ST::NPM; ++q) S.pacc[q] -= pl[q];
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = lamp[c] + (xb[c] + jt[c]);
        S.measure(m - 1, yoff, X1 + H, gr);
        S.store(m - 1, gr);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) lam[c] = 0.0;
      if (fin) {                                  // last row of the path: no residual starts here
        double gr[C];
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = lamp[c];
        S.measure(m - 1, yoff, X1 + H, gr);
        S.store(m - 1, gr);
      }
    }
    if (vm) {
      S.read_halo(xoff, X1);
#pragma unroll
      for (int c = 0; c < C; ++c) X1[H + c] = xm[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) lamp[c] = lam[c];
    valid1 = vm;
  };

  // ---- prologue: row r0 - 1 (second row of stage 0)
  S.wait(0);
  valid1 = (r0 >= 1);
  if (valid1) {
    S.read_row(S.slot_off(0) + rb, X1);
  } else {
#pragma unroll
    for (int c = 0; c < W; ++c) X1[c] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) lamp[c] = 0.0;
  const int tmax = (r1 - r0) / 2;                               // last step (m1 = r0 + 2 tmax <= r1)
  for (int t = 0; t <= tmax; ++t) {
    const int q = t + 1, m1 = r0 + 2 * t;
    S.wait(q);
    const uint32_t sq = S.slot_off(q), sp = S.slot_off(q - 1);
    sub(m1, sq, sp + 3u * rb, m1 < N, t >= 1);
    if (m1 + 1 <= r1) sub(m1 + 1, sq + rb, sq + 2u * rb, m1 + 1 < N, true);
    __syncwarp();
    S.issue(q - 1 + NS);
  }
  vab_pdl_wait();             // partials of the previous evaluation have been consumed by its finalize kernel
  S.finish(red);
}
