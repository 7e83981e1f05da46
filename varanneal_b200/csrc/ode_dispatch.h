// Runtime (model, strip width, disc) -> compile-time walker type.  The callee supplies a functor
// with `template <class WK> int run()`; the product launches a kernel there, the test-only
// emulator runs the same walker on the host.
#pragma once
#include "ode_walk.cuh"

#ifndef VAB_PD
#define VAB_PD 2   // software prefetch distance (rows / pairs in flight per thread); must be even
#endif

template <class M, class F>
int ode_dispatch_disc(int disc, F& fn) {
  switch (disc) {
    case DISC_EULER: return fn.template run<typename WalkSelect<M, DISC_EULER, VAB_PD>::type>();
    case DISC_TRAPEZOID: return fn.template run<typename WalkSelect<M, DISC_TRAPEZOID, VAB_PD>::type>();
    case DISC_SIMPSON: return fn.template run<typename WalkSelect<M, DISC_SIMPSON, VAB_PD>::type>();
    case DISC_FORWARDMAP: return fn.template run<typename WalkSelect<M, DISC_FORWARDMAP, VAB_PD>::type>();
    case DISC_RK4: return fn.template run<typename WalkSelect<M, DISC_RK4, VAB_PD>::type>();
  }
  return -1;
}

template <class F>
int ode_dispatch(int model, int C, int disc, F& fn) {
  if (model == 0) {
    if (C == 4) return ode_dispatch_disc<ModelL96<4>>(disc, fn);
    if (C == 2) return ode_dispatch_disc<ModelL96<2>>(disc, fn);
    if (C == 1) return ode_dispatch_disc<ModelL96<1>>(disc, fn);
    return -1;
  }
  if (model == 1) return ode_dispatch_disc<ModelL63>(disc, fn);
  if (model == 2) return ode_dispatch_disc<ModelNaKL>(disc, fn);
  return -1;
}

// Sum of one partial slot over the segments of path b, in segment order (deterministic).
VAB_HD double ode_partial_sum(const OdeParams& P, int b, int k) {
  double acc = 0.0;
  for (int u = 0; u < P.upp; ++u) acc += P.partials[((long long)b * P.upp + u) * P.K + k];
  return acc;
}
