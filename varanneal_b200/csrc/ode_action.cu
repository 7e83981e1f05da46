// Kernel wrappers and launcher for the fused ODE action+gradient (see ode_walk.cuh for the design).
#include <cuda_runtime.h>

#include "ode_action.h"
#include "ode_dispatch.h"

namespace {

template <class WK, int U, int PH>
__device__ __forceinline__ void dev_phases(WK& w, int step) {
  w.template phase<U, PH>(step);
  if constexpr (WK::H > 0) __syncthreads();
  if constexpr (PH + 1 < WK::NPH) dev_phases<WK, U, PH + 1>(w, step);
}
template <class WK, int U>
__device__ __forceinline__ void dev_steps(WK& w, int s0) {
  dev_phases<WK, U, 0>(w, s0 + U);
  if constexpr (U + 1 < WK::PD) dev_steps<WK, U + 1>(w, s0);
}

template <class WK, int MAXT>
__global__ void __launch_bounds__(MAXT) ode_walk_kernel(const __grid_constant__ OdeParams P) {
  extern __shared__ double smem[];
  WK w;
  w.init(P, blockIdx.x, threadIdx.x, smem);
  w.prologue();
  if constexpr (WK::H > 0) __syncthreads();
  const int ns = WK::nsteps(P);
  for (int s0 = 0; s0 < ns; s0 += WK::PD) dev_steps<WK, 0>(w, s0);
  __syncthreads();                       // exchange rows are dead; reuse smem for the reduction
  w.finish_write(threadIdx.x, WK::PSIGN);
  __syncthreads();
  walk_reduce(P, blockIdx.x, threadIdx.x, blockDim.x, smem);
}

// one warp per path: lane k sums partial slot k over the path's segments in segment order
__global__ void ode_finalize_kernel(const __grid_constant__ OdeParams P, double* A, double* me,
                                    double* fe) {
  const int b = blockIdx.x, k = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  double v = (k < P.K) ? ode_partial_sum(P, b, k) : 0.0;
  const double m = __shfl_sync(0xffffffffu, v, 0), f = __shfl_sync(0xffffffffu, v, 1);
  if (k == 0) {
    if (me) me[b] = m;
    if (fe) fe[b] = f;
    if (A) A[b] = m + f;
  }
  if (k >= 2 && k < P.K && P.G != nullptr) {
    const int e = P.pmap[k - 2];
    if (e >= 0) P.G[(long long)b * P.ldg + (long long)P.N * P.D + e] = v;
  }
}

struct DevRun {
  const OdeParams* P;
  const OdePlan* pl;
  cudaStream_t st;
  cudaError_t err = cudaSuccess;
  template <class WK>
  int run() {
    const size_t smem = (size_t)walk_smem_doubles<WK>(*P, pl->NT) * sizeof(double);
    if (pl->NT <= 128) {
      return launch<WK, 128>(smem);
    }
    return launch<WK, 256>(smem);
  }
  template <class WK, int MAXT>
  int launch(size_t smem) {
    auto kern = ode_walk_kernel<WK, MAXT>;
    if (smem > 48 * 1024) {
      err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return -2;
    }
    kern<<<pl->grid, pl->NT, smem, st>>>(*P);
    err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -2;
  }
};

}  // namespace

int ode_launch_action(const OdeParams& P, const OdePlan& pl, int model, int disc,
                      cudaStream_t st, double* A, double* me, double* fe, cudaError_t* cerr) {
  DevRun dr{&P, &pl, st};
  int rc = ode_dispatch(model, pl.C, disc, dr);
  if (rc != 0) {
    if (cerr) *cerr = dr.err;
    return rc;
  }
  ode_finalize_kernel<<<P.B, 32, 0, st>>>(P, A, me, fe);
  cudaError_t e = cudaGetLastError();
  if (cerr) *cerr = e;
  return e == cudaSuccess ? 0 : -2;
}
