#pragma once
#include <cuda_runtime.h>

#include "ode_params.h"
#include "ode_plan.h"

// Register-sweep kernels (ode_sweep.cuh).  ode_sweep_prepare fills the mapping fields of P
// (TPR, GW, GPW, WS, nwin, NHL, Tseg, nseg, nunits, upp) and the launch shape; the caller then
// sizes P.partials (nunits * K doubles) and calls ode_sweep_launch (walk kernel + finalize).
struct SweepLaunch {
  int C;         // strip width
  int grid;      // CTAs of 128 threads
  bool stream;   // TMA stream kernel (ode_stream.cuh) instead of the register sweep
  size_t smem;   // dynamic shared memory per CTA
};
int ode_sweep_prepare(int model, int disc, int D, int N, int B, int num_sms, int tseg_override,
                      bool allow_stream, OdeParams* P, SweepLaunch* sl, cudaError_t* cerr);
// pdl: launch both kernels with programmatic stream serialization (stand-alone evaluations only)
int ode_sweep_launch(const OdeParams& P, const SweepLaunch& sl, int model, int disc,
                     cudaStream_t st, double* A, double* me, double* fe, cudaError_t* cerr, bool pdl);
