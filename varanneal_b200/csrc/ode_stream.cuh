// Fused ODE action + adjoint gradient: the TMA-fed streaming kernels (the hot path for Lorenz96).
//
// Same mathematics and the same (path, time segment, column window) work units as the register
// sweep kernels in ode_sweep.cuh, with the data movement rebuilt around Blackwell's bulk-copy
// engine:
//   * every warp owns a private ring of NS stages in shared memory.  A stage holds two
//     consecutive time rows of the warp's window of X (for each of the GPW paths the warp works
//     on) plus the matching rows of the observations Y.  One lane per path issues
//     `cp.async.bulk` (TMA, 1-D) copies global -> shared that complete on the stage's mbarrier;
//     the other lanes only wait on the barrier.  No registers are spent on prefetching, loads are
//     NS-2 stages (2(NS-2) rows) ahead of the compute, and the measurement data arrives with the
//     state instead of stalling the gradient on an L2 round trip.
//   * lanes read their own strip and its two-component halo straight from the staged row
//     (4 x LDS.128); only the adjoint seed v is exchanged between lanes, with warp shuffles.
//   * all warps of a (segment, window) share row validity, so the time loop is uniform: the
//     steady state runs without any per-lane predicate except the store mask; ragged ends are
//     peeled into separate code.
//   * there is no __syncthreads anywhere; warps run decoupled.
// Gradient rows are written once with streaming 16-byte stores; per-unit partial sums are reduced
// in a fixed order (bit-reproducible).  Reference semantics: va_ode.py:130-234 (action),
// :341-380, :404-454 (discretisations); adjoint formulas SURVEY.md App. A.3.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "ode_models.cuh"
#include "ode_walk.cuh"

#ifndef VAB_FULL
#define VAB_FULL 0xffffffffu
#endif

namespace vabs {

__device__ __forceinline__ uint32_t s32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared (TMA); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// FAST: nskip == 1, scalar RM, scalar RF (the common case); otherwise every weight / row test is
// looked up at run time.
template <class M, int NS, bool FAST>
struct Stream {
  static constexpr int C = M::C, H = M::H, W = M::C + 2 * M::H, NPM = M::NPM;
  static_assert(H == 2 && (C == 4 || C == 2), "stream kernels: Lorenz96-type stencil, strips of 4 or 2");
  const OdeParams& P;
  // warp-uniform
  int lane, sg, w, r0, r1, N, D, Wd, Lw, nact, rowbase, qmax, y0w;
  bool wrap, has_obs;
  uint32_t full_bytes;         // bytes a full stage (two valid rows) brings in, all paths of the warp
  int st0, n1;                 // window mode: first strip of the window, strips before the wrap
  double* ring;                // this warp's stages
  uint32_t bar0;               // shared address of this warp's first barrier
  int stage_d, grp_d;
  // per lane
  int g, j, b, bidx, i0;
  bool pact, out, leader, bvalid;
  double wfs;                  // 2 cf RF (scalar RF)
  int o_own, o_l, o_r;         // offsets (doubles) inside a staged window row
  int srcM1, srcP1;
  const double* xpath;
  double* gpath;
  double p[NPM];
  int ys[C];                   // Y column inside the staged Y row (0 if unobserved)
  int slot[C];                 // Y column in the library layout, -1 if unobserved
  double wobs[C];              // 2 cm RM for observed components (scalar RM), else 0
  double me_acc, fe_acc, pacc[NPM];

  __device__ __forceinline__ Stream(const OdeParams& P_) : P(P_) {}

  // returns false if this warp has no work
  __device__ __forceinline__ bool init(double* smem) {
    const int warp = __shfl_sync(VAB_FULL, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    lane = threadIdx.x & 31;
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
    Wd = P.GW * C;
    Lw = P.Lw;
    grp_d = 2 * Wd + 2 * Lw;
    stage_d = P.GPW * grp_d;
    ring = smem + (size_t)warp * NS * stage_d;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)4 * NS * stage_d) + warp * NS;
    bar0 = s32(bars);
    y0w = __ldg(P.win_y0 + w);
    g = lane / P.GW;
    j = lane - g * P.GW;
    const bool ingroup = g < P.GPW;
    b = bgrp * P.GPW + g;
    bidx = b;
    bvalid = ingroup && b < P.B;
    wfs = 2.0 * P.cf * P.rf_scalar;
    pact = ingroup && b < P.B && (P.active == nullptr || __ldg(P.active + b) != 0);
    if (!ingroup || b >= P.B) b = 0;
    if (!ingroup) g = 0;
    leader = pact && j == 0;
    nact = __popc(__ballot_sync(VAB_FULL, leader));
    has_obs = P.L > 0;
    full_bytes = (uint32_t)nact * (uint32_t)(2 * Wd + (has_obs ? 2 * Lw : 0)) * 8u;
    const int TPR = P.TPR;
    int st;
    if (wrap) {
      st = j;
      out = pact;
      o_own = j * C;
      o_l = (j * C - 2 + Wd) % Wd;
      o_r = (j * C + C) % Wd;
      const int gb = g * P.GW;
      srcM1 = gb + (j + P.GW - 1) % P.GW;
      srcP1 = gb + (j + 1) % P.GW;
      st0 = 0;
      n1 = P.GW;
    } else {
      const int rel = w * P.WS - P.NHL;
      st0 = ((rel % TPR) + TPR) % TPR;
      n1 = min(P.GW, TPR - st0);
      st = (st0 + j) % TPR;
      out = pact && j >= P.NHL && j < P.NHL + P.WS && (w * P.WS + j - P.NHL) < TPR;
      o_own = j * C;
      o_l = max(j * C - 2, 0);
      o_r = min(j * C + C, Wd - 2);
      const int gb = g * P.GW;
      srcM1 = gb + max(j - 1, 0);
      srcP1 = gb + min(j + 1, P.GW - 1);
    }
    if (!ingroup) { srcM1 = srcP1 = lane; o_own = o_l = o_r = 0; }
    i0 = st * C;
    xpath = P.XP + (long long)b * P.ldxp;
    gpath = P.G ? P.G + (long long)b * P.ldg + (long long)st * C : nullptr;
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
    for (int c = 0; c < C; ++c) {
      const int s = out ? __ldg(P.obs_slot + st * C + c) : -1;
      slot[c] = s;
      ys[c] = (s >= 0) ? s - y0w : 0;
      wobs[c] = (s >= 0) ? 2.0 * P.cm * P.rm_scalar : 0.0;
    }
    me_acc = 0.0;
    fe_acc = 0.0;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8 * s, 1);
      mbar_fence_init();
    }
    __syncwarp();
    return true;
  }

  __device__ __forceinline__ double* stage(int q) const { return ring + (size_t)(q % NS) * stage_d + g * grp_d; }
  __device__ __forceinline__ bool obs_row(int r) const { return FAST || P.nskip == 1 || (r % P.nskip) == 0; }
  __device__ __forceinline__ int obs_index(int r) const { return FAST ? r : r / P.nskip; }

  // Enqueue the copies of stage q: rows rowbase + 2q and rowbase + 2q + 1.  All lanes call it.
  __device__ __forceinline__ void issue(int q) {
    if (q > qmax) return;
    const int ra = rowbase + 2 * q, rb = ra + 1;
    const uint32_t bar = bar0 + 8 * (q % NS);
    if (FAST && wrap && ra >= 0 && rb < N) {       // steady state: two valid rows, rows contiguous
      if (lane == 0) mbar_expect_tx(bar, full_bytes);
      __syncwarp();
      if (leader) {
        double* dst = stage(q);
        tma_load(s32(dst), xpath + (long long)ra * D, 2u * Wd * 8u, bar);
        if (has_obs) tma_load(s32(dst + 2 * Wd), P.Y + (long long)ra * P.Lp, 2u * Lw * 8u, bar);
      }
      return;
    }
    const bool va = ra >= 0 && ra < N, vb = rb >= 0 && rb < N;
    const bool ya = va && has_obs && obs_row(ra), yb = vb && has_obs && obs_row(rb);
    const uint32_t per = (uint32_t)((va ? 1 : 0) + (vb ? 1 : 0)) * Wd * 8u +
                         (uint32_t)((ya ? 1 : 0) + (yb ? 1 : 0)) * Lw * 8u;
    if (lane == 0) mbar_expect_tx(bar, per * (uint32_t)nact);
    __syncwarp();
    if (leader) {
      double* dst = stage(q);
      if (wrap) {
        if (va && vb) {
          tma_load(s32(dst), xpath + (long long)ra * D, 2u * Wd * 8u, bar);
        } else if (va) {
          tma_load(s32(dst), xpath + (long long)ra * D, Wd * 8u, bar);
        } else if (vb) {
          tma_load(s32(dst + Wd), xpath + (long long)rb * D, Wd * 8u, bar);
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0 ? va : vb) {
            const double* src = xpath + (long long)(ra + h) * D;
            tma_load(s32(dst + h * Wd), src + st0 * C, (uint32_t)n1 * C * 8u, bar);
            if (n1 < P.GW) tma_load(s32(dst + h * Wd + n1 * C), src, (uint32_t)(P.GW - n1) * C * 8u, bar);
          }
        }
      }
      if (ya) tma_load(s32(dst + 2 * Wd), P.Y + (long long)obs_index(ra) * P.Lp + y0w, Lw * 8u, bar);
      if (yb) tma_load(s32(dst + 2 * Wd + Lw), P.Y + (long long)obs_index(rb) * P.Lp + y0w, Lw * 8u, bar);
    }
  }
  __device__ __forceinline__ void wait(int q) const {
    if (q > qmax) return;
    mbar_wait(bar0 + 8 * (q % NS), (uint32_t)((q / NS) & 1));
  }

  // staged row -> own strip with halo
  __device__ __forceinline__ void read_row(const double* row, double* X) const {
    const double2 l = *reinterpret_cast<const double2*>(row + o_l);
    const double2 r = *reinterpret_cast<const double2*>(row + o_r);
    X[0] = l.x; X[1] = l.y;
    X[H + C] = r.x; X[H + C + 1] = r.y;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      const double2 o = *reinterpret_cast<const double2*>(row + o_own + c);
      X[H + c] = o.x; X[H + c + 1] = o.y;
    }
  }
  // own values in V[H..H+C): fetch the halo from the neighbour lanes (all lanes must call)
  __device__ __forceinline__ void halo(double* V) const {
    V[0] = __shfl_sync(VAB_FULL, V[H + C - 2], srcM1);
    V[1] = __shfl_sync(VAB_FULL, V[H + C - 1], srcM1);
    V[H + C] = __shfl_sync(VAB_FULL, V[H], srcP1);
    V[H + C + 1] = __shfl_sync(VAB_FULL, V[H + 1], srcP1);
  }
  __device__ __forceinline__ double wgt(int row, int c) const {   // 2 cf RF for residual (row, c)
    if (FAST) return wfs;
    return P.rf_arr ? 2.0 * P.cf * P.rf_scale * __ldg(P.rf_arr + (long long)row * D + i0 + c) : wfs;
  }
  // measurement term of row r (va_ode.py:138-158) from the staged Y row: d += 2 cm RM (x - y)
  __device__ __forceinline__ void measure(int r, const double* yrow, const double* xown, double* d) {
    if (!has_obs || !obs_row(r)) return;                       // warp-uniform
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double wo = wobs[c];
      if (!FAST) {
        if (P.rm_arr != nullptr && slot[c] >= 0)
          wo = 2.0 * P.cm * __ldg(P.rm_arr + (long long)obs_index(r) * P.Lp + slot[c]);
      }
      const double diff = xown[c] - yrow[ys[c]];
      const double wd = wo * diff;
      me_acc = fma(wd, diff, me_acc);
      d[c] += wd;
    }
  }
  // g = d + v - t  (t = adjoint product without its -v term); streaming 16-byte stores
  __device__ __forceinline__ void store(int r, const double* g) const {
    if (out && gpath != nullptr) vab_store_strip<C>(gpath + (long long)r * D, g);
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
// their pair; row c's partial seed is carried into the next pair, where it is row a.
template <class M, int NS, int MINB, bool FAST>
__global__ void __launch_bounds__(128, MINB) stream_simpson_kernel(const __grid_constant__ OdeParams P) {
  using ST = vabs::Stream<M, NS, FAST>;
  constexpr int C = ST::C, H = ST::H, W = ST::W;
  extern __shared__ __align__(16) double smem[];
  ST S(P);
  if (!S.init(smem)) return;
  double* red = smem + (size_t)4 * NS * S.stage_d + 4 * NS;
  const double dt = P.dt, dt3 = dt / 3.0, dt4 = dt / 4.0, dt43 = 4.0 * dt / 3.0;
  const int r0 = S.r0, N = S.N, Wd = S.Wd, Lw = S.Lw;
  const bool last = (S.r1 == N);
  const int nfull = last ? (N - 1 - r0) / 2 : P.Tseg / 2;
  S.rowbase = r0 - 3;                       // stage q holds rows (r0 - 3 + 2q, r0 - 2 + 2q)
  S.qmax = nfull + 1;
#pragma unroll
  for (int q = 0; q < NS; ++q) S.issue(q);

  double Xa[W], Fa[C], vcp[C], dcp[C];
  // one pair; STORE = false for the lead-in pair (only the carried c-part is wanted)
  auto pair = [&](int a, const double* stq, const double* yarow, auto store_tag) {
    constexpr bool STORE = decltype(store_tag)::value;
    double Xb[W], Xc[W], Fb[C], Fc[C], l1[C], l2[C];
    S.read_row(stq, Xb);
    S.read_row(stq + Wd, Xc);
    M::f(Xb, S.p, nullptr, Fb);
    M::f(Xc, S.p, nullptr, Fc);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double s = fma(4.0, Fb[c], Fa[c]) + Fc[c];
      const double e1 = fma(-dt3, s, Xc[H + c] - Xa[H + c]);
      const double e2 = fma(-dt4, Fa[c] - Fc[c], fma(-0.5, Xa[H + c] + Xc[H + c], Xb[H + c]));
      l1[c] = S.wgt(a, c) * e1;
      l2[c] = S.wgt(a + 1, c) * e2;
      if (STORE) S.fe_acc = fma(l1[c], e1, fma(l2[c], e2, S.fe_acc));
    }
    if (STORE) {
      {   // row b:  v = 4dt/3 l1,  d = l2
        double V[W], g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { V[H + c] = dt43 * l1[c]; g[c] = l2[c]; }
        S.halo(V);
        S.measure(a + 1, stq + 2 * Wd, Xb + H, g);
        double t[C];
        M::adj_t(Xb, V, S.p, t, S.pacc);
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = (g[c] + V[H + c]) - t[c];
        S.store(a + 1, g);
      }
      {   // row a:  v = carried c-part + dt/3 l1 + dt/4 l2,  d = carried - l1 - l2/2
        double V[W], g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          V[H + c] = fma(dt4, l2[c], fma(dt3, l1[c], vcp[c]));
          g[c] = fma(-0.5, l2[c], dcp[c] - l1[c]);
        }
        S.halo(V);
        S.measure(a, yarow, Xa + H, g);
        double t[C];
        M::adj_t(Xa, V, S.p, t, S.pacc);
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = (g[c] + V[H + c]) - t[c];
        S.store(a, g);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      vcp[c] = fma(-dt4, l2[c], dt3 * l1[c]);
      dcp[c] = fma(-0.5, l2[c], l1[c]);
      Fa[c] = Fc[c];
    }
#pragma unroll
    for (int c = 0; c < W; ++c) Xa[c] = Xc[c];
  };

  // ---- lead-in: pair (r0-2, r0-1, r0), or just row 0 when the segment starts the path
  S.wait(0);
  S.wait(1);
  {
    const double* st1 = S.stage(1);
    if (r0 >= 2) {
      S.read_row(S.stage(0) + Wd, Xa);
      M::f(Xa, S.p, nullptr, Fa);
      pair(r0 - 2, st1, nullptr, std::false_type{});
    } else {
      S.read_row(st1 + Wd, Xa);
      M::f(Xa, S.p, nullptr, Fa);
#pragma unroll
      for (int c = 0; c < C; ++c) { vcp[c] = 0.0; dcp[c] = 0.0; }
    }
  }
  __syncwarp();
  S.issue(NS);
  // ---- steady state
  for (int t = 1; t <= nfull; ++t) {
    const int q = t + 1;
    S.wait(q);
    pair(r0 + 2 * (t - 1), S.stage(q), S.stage(q - 1) + 2 * Wd + Lw, std::true_type{});
    __syncwarp();
    S.issue(q - 1 + NS);
  }
  // ---- the last row of the path (even, only the c-part of the last pair and its measurement)
  if (last) {
    double V[W], g[C], t[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { V[H + c] = vcp[c]; g[c] = dcp[c]; }
    S.halo(V);
    S.measure(N - 1, S.stage(nfull + 1) + 2 * Wd + Lw, Xa + H, g);
    M::adj_t(Xa, V, S.p, t, S.pacc);
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = (g[c] + V[H + c]) - t[c];
    S.store(N - 1, g);
  }
  S.finish(red);
}

// --------------------------------------------------------------------------------------------
// euler / trapezoid / forwardmap (va_ode.py:341-380, 439-454):
//   e_m = x_{m+1} - AL x_m - (CA f_m + CB f_{m+1}),  lam = 2 cf RF e
//   g_r = [lam_{r-1} - AL lam_r] + meas_r - J^T(x_r) (CB lam_{r-1} + CA lam_r)
// One step = one ring stage = two arriving rows.
template <class M, int DISC, int NS, int MINB, bool FAST>
__global__ void __launch_bounds__(128, MINB) stream_twopoint_kernel(const __grid_constant__ OdeParams P) {
  using ST = vabs::Stream<M, NS, FAST>;
  constexpr int C = ST::C, H = ST::H, W = ST::W;
  extern __shared__ __align__(16) double smem[];
  ST S(P);
  if (!S.init(smem)) return;
  double* red = smem + (size_t)4 * NS * S.stage_d + 4 * NS;
  const double dt = P.dt;
  const double ca = (DISC == DISC_EULER) ? dt : (DISC == DISC_TRAPEZOID ? 0.5 * dt : 1.0);
  const double cb = (DISC == DISC_TRAPEZOID) ? 0.5 * dt : 0.0;
  const double al = (DISC == DISC_FORWARDMAP) ? 0.0 : 1.0;
  const int r0 = S.r0, r1 = S.r1, N = S.N, Wd = S.Wd, Lw = S.Lw;
  S.rowbase = r0 - 2;                       // stage q holds rows (r0 - 2 + 2q, r0 - 1 + 2q)
  S.qmax = (min(r1, N - 1) - r0 + 2) / 2;
#pragma unroll
  for (int q = 0; q < NS; ++q) S.issue(q);

  double X1[W], F1[C], lamp[C];
  bool valid1;
  // row m arrives (staged at xrow); finalises row m-1 (its Y row at yrow) when fin
  auto sub = [&](int m, const double* xrow, const double* yrow, bool vm, bool fin) {
    double Xm[W], Fm[C], lam[C];
    if (vm) {
      S.read_row(xrow, Xm);
      M::f(Xm, S.p, nullptr, Fm);
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) Xm[c] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) Fm[c] = 0.0;
    }
    const bool ve = vm && valid1;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double l = 0.0;
      if (ve) {
        const double e = Xm[H + c] - al * X1[H + c] - fma(ca, F1[c], cb * Fm[c]);
        l = S.wgt(m - 1, c) * e;
        if (m - 1 >= r0 && m - 1 < r1) S.fe_acc = fma(l, e, S.fe_acc);
      }
      lam[c] = l;
    }
    if (fin) {
      double V[W], g[C], t[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        V[H + c] = fma(cb, lamp[c], ca * lam[c]);
        g[c] = lamp[c] - al * lam[c];
      }
      S.halo(V);
      S.measure(m - 1, yrow, X1 + H, g);
      M::adj_t(X1, V, S.p, t, S.pacc);
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = (g[c] + V[H + c]) - t[c];
      S.store(m - 1, g);
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
    S.read_row(S.stage(0) + Wd, X1);
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
    const double* st = S.stage(q);
    const double* stp = S.stage(q - 1);
    if (t >= 1 && t <= tfull) {
      sub(m1, st, stp + 2 * Wd + Lw, true, true);
      sub(m1 + 1, st + Wd, st + 2 * Wd, true, true);
    } else {
      sub(m1, st, stp + 2 * Wd + Lw, m1 < N, t >= 1);
      if (m1 + 1 <= r1) sub(m1 + 1, st + Wd, st + 2 * Wd, m1 + 1 < N, true);
    }
    __syncwarp();
    S.issue(q - 1 + NS);
  }
  S.finish(red);
}
