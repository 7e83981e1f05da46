#pragma once
#include <cuda_runtime.h>

#include "ode_plan.h"
#include "ode_walk.cuh"

// Launches the strip-walk kernel selected by (model, plan.C, disc) followed by the per-path
// finalize kernel (2 launches).  Returns 0, -1 (unsupported combination) or -2 (CUDA error, code
// in *cerr).  A / me / fe may be nullptr.
int ode_launch_action(const OdeParams& P, const OdePlan& pl, int model, int disc,
                      cudaStream_t st, double* A, double* me, double* fe, cudaError_t* cerr);
