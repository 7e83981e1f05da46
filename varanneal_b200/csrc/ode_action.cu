// Launch planning and dispatch of the fused ODE action+gradient kernels
// (register sweep: ode_sweep.cuh; TMA stream: ode_stream.cuh).
#include <cuda_runtime.h>

#include <mutex>

#include <cstdlib>

#include "ode_action.h"
#include "ode_stream.cuh"
#include "ode_sweep.cuh"

namespace {

// one warp per path: lane k sums partial slot k over the path's units in unit order
__global__ void ode_finalize_kernel(const __grid_constant__ OdeParams P, double* A, double* me,
                                    double* fe) {
  const int b = blockIdx.x, k = threadIdx.x;
  vab_pdl_wait();              // the walk kernel has completed: its partials are visible
  vab_pdl_trigger();           // the next evaluation's walk kernel may start its prologue (it waits
                               // again before it overwrites the partials read here)
  if (P.active != nullptr && P.active[b] == 0) return;
  double v = 0.0;
  if (k < P.K)
    for (int u = 0; u < P.upp; ++u) v += P.partials[((long long)b * P.upp + u) * P.K + k];
  const double m = __shfl_sync(0xffffffffu, v, 0), f = __shfl_sync(0xffffffffu, v, 1);
  if (k == 0) {
    if (me) me[b] = m;
    if (fe) fe[b] = f;
    if (A) A[b] = m + f;
  }
  if (k >= 2 && k < P.K && P.G != nullptr && !P.ptime) {   // (a parameter time series gets its
    const int e = P.pmap[k - 2];                            //  gradient row by row in the walk kernel)
    if (e >= 0) P.G[(long long)b * P.ldg + (long long)P.N * P.D + e] = v;
  }
}

}  // namespace

// ============================================================================================
// sweep kernels: plan + dispatch
// ============================================================================================
#ifndef VAB_SW_PD2
#define VAB_SW_PD2 2     // prefetch distance (rows) of the two-point kernels
#endif
#ifndef VAB_SW_PDS
#define VAB_SW_PDS 1     // prefetch distance (pairs) of the Simpson kernel
#endif
#ifndef VAB_SW_PDR
#define VAB_SW_PDR 1     // prefetch distance (rows) of the RK4 kernel
#endif
#ifndef VAB_SW_MINB
#define VAB_SW_MINB 4    // resident CTAs per SM the register allocator must allow
#endif

namespace {

typedef void (*SweepKernel)(const OdeParams);

template <class M, int MINB>
SweepKernel sweep_kernel_for(int disc, bool window) {
  switch (disc) {
    case DISC_EULER: return sweep_twopoint_kernel<M, DISC_EULER, VAB_SW_PD2, MINB>;
    case DISC_TRAPEZOID: return sweep_twopoint_kernel<M, DISC_TRAPEZOID, VAB_SW_PD2, MINB>;
    case DISC_FORWARDMAP: return sweep_twopoint_kernel<M, DISC_FORWARDMAP, VAB_SW_PD2, MINB>;
    case DISC_SIMPSON: return sweep_simpson_kernel<M, VAB_SW_PDS, MINB>;
    case DISC_RK4: {
      // The RK4 body (4 stage states + adjoint seeds live at once) spills at 128 registers.
      // Measured on B200: rows that fit one lane group (D = 100, B = 64) still prefer 4 resident
      // CTAs per SM (0.357 ms vs 0.562 ms at 3); window mode (D = 1000, B = 8) prefers the
      // spill-free 164-register build at 3 CTAs per SM (1.67 ms vs 2.65 ms).  VAB_RK4_MINB overrides.
      static int forced = -1;
      if (forced < 0) {
        const char* e = getenv("VAB_RK4_MINB");
        forced = e ? atoi(e) : 0;
      }
      const int minb = forced ? forced : (window ? 3 : 4);
      if (minb == 2) return sweep_rk4_kernel<M, VAB_SW_PDR, 2>;
      if (minb == 3) return sweep_rk4_kernel<M, VAB_SW_PDR, 3>;
      return sweep_rk4_kernel<M, VAB_SW_PDR, 4>;
    }
  }
  return nullptr;
}

// parameter time series (OdeParams::ptime): the PT instantiations of the two register kernels that
// carry a per-row parameter load and gradient; whole rows inside one lane group only, no rk4
template <class M, int MINB>
SweepKernel sweep_kernel_ptime(int disc) {
  switch (disc) {
    case DISC_EULER: return sweep_twopoint_kernel<M, DISC_EULER, VAB_SW_PD2, MINB, true>;
    case DISC_TRAPEZOID: return sweep_twopoint_kernel<M, DISC_TRAPEZOID, VAB_SW_PD2, MINB, true>;
    case DISC_FORWARDMAP: return sweep_twopoint_kernel<M, DISC_FORWARDMAP, VAB_SW_PD2, MINB, true>;
    case DISC_SIMPSON: return sweep_simpson_kernel<M, VAB_SW_PDS, MINB, true>;
  }
  return nullptr;
}

SweepKernel sweep_kernel(int model, int C, int disc, bool window, bool ptime = false, bool mrf = false) {
  if (mrf) {                                  // matrix RF: SimpsonHermite, whole rows in one lane group
    if (window || ptime || disc != DISC_SIMPSON) return nullptr;
    if (model == 0 && C == 4) return sweep_simpson_kernel<ModelL96<4>, VAB_SW_PDS, 3, false, true>;
    if (model == 0 && C == 2) return sweep_simpson_kernel<ModelL96<2>, VAB_SW_PDS, VAB_SW_MINB, false, true>;
    if (model == 0 && C == 1) return sweep_simpson_kernel<ModelL96<1>, VAB_SW_PDS, VAB_SW_MINB, false, true>;
    if (model == 1) return sweep_simpson_kernel<ModelL63, VAB_SW_PDS, VAB_SW_MINB, false, true>;
    if (model == 2) return sweep_simpson_kernel<ModelNaKL, VAB_SW_PDS, 2, false, true>;
    return nullptr;
  }
  if (ptime) {
    if (window) return nullptr;
    if (model == 0 && C == 4) return sweep_kernel_ptime<ModelL96<4>, 3>(disc);   // (spills at 128 registers)
    if (model == 0 && C == 2) return sweep_kernel_ptime<ModelL96<2>, VAB_SW_MINB>(disc);
    if (model == 0 && C == 1) return sweep_kernel_ptime<ModelL96<1>, VAB_SW_MINB>(disc);
    if (model == 1) return sweep_kernel_ptime<ModelL63, VAB_SW_MINB>(disc);
    if (model == 2) return sweep_kernel_ptime<ModelNaKL, 2>(disc);
    return nullptr;
  }
  if (model == 0) {
    if (C == 4) {
      if (disc == DISC_SIMPSON) {                 // tuning variants of the flagship kernel
        static int variant = -1;
        if (variant < 0) {
          const char* v = getenv("VAB_SIMPSON_VARIANT");
          variant = v ? atoi(v) : 0;
        }
        switch (variant) {
          case 1: return sweep_simpson_kernel<ModelL96<4>, 1, 3>;
          case 2: return sweep_simpson_kernel<ModelL96<4>, 2, 3>;
          case 3: return sweep_simpson_kernel<ModelL96<4>, 2, 2>;
          default: return sweep_simpson_kernel<ModelL96<4>, 1, 4>;
        }
      }
      return sweep_kernel_for<ModelL96<4>, VAB_SW_MINB>(disc, window);
    }
    if (C == 2) return sweep_kernel_for<ModelL96<2>, VAB_SW_MINB>(disc, window);
    if (C == 1) return sweep_kernel_for<ModelL96<1>, VAB_SW_MINB>(disc, window);
    return nullptr;
  }
  if (model == 1) return sweep_kernel_for<ModelL63, VAB_SW_MINB>(disc, window);
  if (model == 2) return sweep_kernel_for<ModelNaKL, 2>(disc, window);
  return nullptr;
}

#ifndef VAB_ST_NS
#define VAB_ST_NS 4      // ring stages per warp of the stream kernels (prefetch = NS-2 stages)
#endif
#ifndef VAB_ST_MINB
#define VAB_ST_MINB 4
#endif

// MODE: 0 = run-time weights, 1 = scalar RM / RF (launch constants), 2 = the same with one RF per
// path (ode_stream.cuh)
template <class M, int MODE>
SweepKernel stream_kernel_for(int disc, bool window) {
  switch (disc) {
    case DISC_EULER: return stream_twopoint_kernel<M, DISC_EULER, VAB_ST_NS, VAB_ST_MINB, MODE>;
    case DISC_TRAPEZOID: return stream_twopoint_kernel<M, DISC_TRAPEZOID, VAB_ST_NS, VAB_ST_MINB, MODE>;
    case DISC_FORWARDMAP: return stream_twopoint_kernel<M, DISC_FORWARDMAP, VAB_ST_NS, VAB_ST_MINB, MODE>;
    case DISC_SIMPSON: return stream_simpson_kernel<M, VAB_ST_NS, VAB_ST_MINB, MODE>;
    case DISC_RK4: {
      // resident CTAs per SM the register budget allows.  Measured on B200: one lane group per
      // row (D = 100, B = 64) 0.280 ms at 4 (128 registers, spills) vs 0.304 ms at 3; window mode
      // (D = 1000, B = 8) 1.34 ms at 3 (168 registers, no spills) vs 1.76 ms at 4.
      static int forced = -1;
      if (forced < 0) {
        const char* e = getenv("VAB_RK4_STREAM_MINB");
        forced = e ? atoi(e) : 0;
      }
      const int minb = forced ? forced : (window ? 3 : 4);
      if (minb == 4) return stream_rk4_kernel<M, VAB_ST_NS, 4, MODE>;
      if (minb == 2) return stream_rk4_kernel<M, VAB_ST_NS, 2, MODE>;
      return stream_rk4_kernel<M, VAB_ST_NS, 3, MODE>;
    }
  }
  return nullptr;
}
template <class M>
SweepKernel stream_kernel_mode(int disc, int mode, bool window) {
  if (mode == 1) return stream_kernel_for<M, 1>(disc, window);
  if (mode == 2) return stream_kernel_for<M, 2>(disc, window);
  return stream_kernel_for<M, 0>(disc, window);
}
int stream_variant() {
  static int variant = -1;
  if (variant < 0) {
    const char* v = getenv("VAB_STREAM_VARIANT");
    variant = v ? atoi(v) : 0;
  }
  return variant;
}
int stream_ns(int C, int disc, bool fast) {
  if (fast && C == 4 && disc == DISC_SIMPSON) {
    switch (stream_variant()) {
      case 2: return 3;
      case 3: return 3;
      default: return 4;
    }
  }
  return VAB_ST_NS;
}
SweepKernel stream_kernel(int C, int disc, bool fast, bool perpath, bool window) {
  const int mode = fast ? (perpath ? 2 : 1) : 0;
  if (mode == 1 && C == 4 && disc == DISC_SIMPSON) {       // tuning variants of the flagship kernel
    switch (stream_variant()) {
      case 1: return stream_simpson_kernel<ModelL96<4>, 4, 5, 1>;
      case 2: return stream_simpson_kernel<ModelL96<4>, 3, 5, 1>;
      case 3: return stream_simpson_kernel<ModelL96<4>, 3, 4, 1>;
      default: return stream_simpson_kernel<ModelL96<4>, 4, 4, 1>;
    }
  }
  if (C == 4) return stream_kernel_mode<ModelL96<4>>(disc, mode, window);
  if (C == 2) return stream_kernel_mode<ModelL96<2>>(disc, mode, window);
  return nullptr;
}

// resident CTAs per SM of each kernel at a given dynamic shared memory size (queried once per
// device: cudaFuncSetAttribute acts on the current device only, so a second context on another
// GPU of the same process must opt in again)
int blocks_per_sm(SweepKernel k, size_t smem, cudaError_t* cerr) {
  struct Entry { SweepKernel k; size_t smem; int nb; int dev; };
  static Entry seen[512];
  static int nseen = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < nseen; ++i)
    if (seen[i].k == k && seen[i].smem == smem && seen[i].dev == dev) return seen[i].nb;
  // opt in to the largest dynamic shared memory size once per kernel (a later, smaller request
  // must not lower the limit of an earlier one)
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int nb = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 128, smem);
  if (e != cudaSuccess) {
    if (cerr) *cerr = e;
    return -1;
  }
  if (nb < 1) return 0;
  if (nseen < 512) seen[nseen++] = Entry{k, smem, nb, dev};
  return nb;
}

}  // namespace

int ode_sweep_prepare(int model, int disc, int D, int N, int B, int num_sms, int tseg_override,
                      bool allow_stream, OdeParams* P, SweepLaunch* sl, cudaError_t* cerr) {
  OdeGeo g;
  if (ode_geometry(model, disc, D, &g) != 0) return -1;
  const int K = 2 + ode_model_npm(model);
  P->TPR = g.TPR; P->GW = g.GW; P->GPW = g.GPW; P->WS = g.WS; P->nwin = g.nwin; P->NHL = g.NHL;
  P->K = K;
  sl->C = g.C;
  sl->stream = false;
  sl->smem = (size_t)128 * K * sizeof(double);
  SweepKernel k = nullptr;
  int nb = 0;
  if (P->ptime && (g.nwin > 1 || disc == DISC_RK4)) return -3;
  if (P->rf_mat && (g.nwin > 1 || disc != DISC_SIMPSON || P->ptime)) return -3;
  if (allow_stream && !P->ptime && !P->rf_mat && ode_stream_supported(model, disc, g)) {
    const size_t stage_b = (size_t)g.GPW * 4 * (size_t)g.GW * g.C * sizeof(double);
    const bool fast = (P->nskip == 1 && P->rmd == nullptr && P->rf_arr == nullptr && P->L > 0);
    const int ns = stream_ns(g.C, disc, fast);
    // ring + barriers + reduction scratch (+ one double per thread: the per-path RF weight of MODE 2)
    const size_t smem = (size_t)4 * ns * stage_b + ((size_t)4 * ns + (size_t)128 * K + 128) * sizeof(double);
    if (smem <= 200 * 1024) {
      k = stream_kernel(g.C, disc, fast, P->rf_path != nullptr, g.nwin > 1);
      nb = k ? blocks_per_sm(k, smem, cerr) : 0;
      if (nb < 0) return -2;
      if (nb > 0) { sl->stream = true; sl->smem = smem; }
    }
  }
  if (!sl->stream) {
    k = sweep_kernel(model, g.C, disc, g.nwin > 1, P->ptime != 0, P->rf_mat != nullptr);
    if (!k) return -1;
    nb = blocks_per_sm(k, sl->smem, cerr);
    if (nb < 0) return -2;
    if (nb == 0) nb = 1;
  }
  // one wave of resident warps, every warp walking an equally long segment
  const int lead = (disc == DISC_SIMPSON) ? 3 : 2;
  long long target_warps = (long long)num_sms * nb * 4;
  if (const char* w = getenv("VAB_WAVES")) target_warps *= (atoi(w) > 0 ? atoi(w) : 1);
  const long long wpb = ((long long)B + g.GPW - 1) / g.GPW;      // warps per (segment, window)
  long long per_seg = wpb * g.nwin;
  long long nseg = (target_warps + per_seg / 2) / per_seg;
  if (nseg < 1) nseg = 1;
  long long tseg = (N + nseg - 1) / nseg;
  const int tmin = 8 * lead;                              // keep the lead-in rows under ~12 %
  if (tseg < tmin) tseg = tmin;
  if (tseg > N) tseg = N;
  if (tseg_override > 0) tseg = tseg_override;
  if (tseg % 2) tseg += 1;
  P->Tseg = (int)tseg;
  P->nseg = (N + P->Tseg - 1) / P->Tseg;
  P->upp = P->nseg * g.nwin;
  P->nunits = B * P->upp;
  P->wpb = (int)wpb;
  long long warps;
  if (sl->stream) {
    warps = wpb * P->nseg * g.nwin;
  } else {
    warps = ((long long)P->nunits + g.GPW - 1) / g.GPW;
  }
  sl->grid = (int)((warps + 3) / 4);
  return 0;
}

int ode_sweep_launch(const OdeParams& P, const SweepLaunch& sl, int model, int disc,
                     cudaStream_t st, double* A, double* me, double* fe, cudaError_t* cerr, bool pdl) {
  const bool fast = (P.nskip == 1 && P.rmd == nullptr && P.rf_arr == nullptr && P.L > 0);
  SweepKernel k = sl.stream ? stream_kernel(sl.C, disc, fast, P.rf_path != nullptr, P.nwin > 1) : sweep_kernel(model, sl.C, disc, P.nwin > 1, P.ptime != 0, P.rf_mat != nullptr);
  if (!k) return -1;
  cudaError_t e;
  if (pdl) {
    // back-to-back evaluations (vab_ode_action_grad): programmatic dependent launch hides the
    // launch latency of the finalize kernel behind the walk kernel and lets the next walk kernel
    // run its prologue while this finalize kernel sums the partials
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sl.grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sl.smem; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, k, P);
    if (e == cudaSuccess) {
      cudaLaunchConfig_t cf2 = {};
      cf2.gridDim = dim3(P.B); cf2.blockDim = dim3(32); cf2.dynamicSmemBytes = 0; cf2.stream = st;
      cf2.attrs = at; cf2.numAttrs = 1;
      e = cudaLaunchKernelEx(&cf2, ode_finalize_kernel, P, A, me, fe);
    }
  } else {
    k<<<sl.grid, 128, sl.smem, st>>>(P);
    e = cudaGetLastError();
    if (e == cudaSuccess) {
      ode_finalize_kernel<<<P.B, 32, 0, st>>>(P, A, me, fe);
      e = cudaGetLastError();
    }
  }
  if (cerr) *cerr = e;
  return e == cudaSuccess ? 0 : -2;
}
