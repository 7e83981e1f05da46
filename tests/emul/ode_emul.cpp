// TEST-ONLY host emulation of the strip-walk kernels (varanneal_b200/csrc/ode_walk.cuh).
// Compiled with g++ (no CUDA) by tests/emul/build.py; runs every CTA / thread / phase of the same
// walker code serially so that index arithmetic, halo exchange, segment ownership and the
// fixed-order reductions can be checked against the oracle on a machine without a GPU.
// The product never loads this library.
#include <cstring>
#include <utility>
#include <vector>

#include "ode_dispatch.h"
#include "ode_plan.h"

namespace {
template <class WK, int U, int PH>
void emu_phase(std::vector<WK>& ws, int step) {
  for (auto& w : ws) w.template phase<U, PH>(step);
}
template <class WK, int U, int... PHs>
void emu_step(std::vector<WK>& ws, int step, std::integer_sequence<int, PHs...>) {
  (emu_phase<WK, U, PHs>(ws, step), ...);
}
template <class WK, int... Us>
void emu_steps(std::vector<WK>& ws, int s0, std::integer_sequence<int, Us...>) {
  (emu_step<WK, Us>(ws, s0 + Us, std::make_integer_sequence<int, WK::NPH>{}), ...);
}

struct EmulRun {
  const OdeParams* P;
  int NT, grid;
  template <class WK>
  int run() {
    std::vector<double> smem(walk_smem_doubles<WK>(*P, NT) + 16);
    for (int bid = 0; bid < grid; ++bid) {
      std::fill(smem.begin(), smem.end(), 0.0);
      std::vector<WK> ws(NT);
      for (int t = 0; t < NT; ++t) ws[t].init(*P, bid, t, smem.data());
      for (int t = 0; t < NT; ++t) ws[t].prologue();
      const int ns = WK::nsteps(*P);
      for (int s0 = 0; s0 < ns; s0 += WK::PD)
        emu_steps<WK>(ws, s0, std::make_integer_sequence<int, WK::PD>{});
      for (int t = 0; t < NT; ++t) ws[t].finish_write(t, WK::PSIGN);
      for (int t = 0; t < NT; ++t) walk_reduce(*P, bid, t, NT, smem.data());
    }
    return 0;
  }
};
}  // namespace

extern "C" int emul_ode_action_grad(
    int model, int disc, int D, int N, int N_data, int nskip, int L, int NP, int NPest, int S,
    double dt, const int* Lidx, const int* Pidx, const double* Y, const double* stim,
    double rm_scalar, const double* rm_arr, double rf0_scalar, const double* rf0_arr,
    const double* pfix, long long pfix_stride, int B, const double* XP, long long ldxp,
    double rf_scale, const int* active, int tseg_override, double* A, double* me, double* fe,
    double* G, long long ldg) {
  OdePlan pl;
  if (ode_make_plan(model, disc, D, N, B, 148, tseg_override, &pl) != 0) return -1;
  if (NP != ode_model_npm(model)) return -2;
  std::vector<int> obs(D, -1), pmap(NP, -1);
  for (int l = 0; l < L; ++l) obs[Lidx[l]] = l;
  for (int e = 0; e < NPest; ++e) pmap[Pidx[e]] = e;
  OdeParams P;
  std::memset(&P, 0, sizeof(P));
  P.XP = XP; P.ldxp = ldxp; P.G = G; P.ldg = ldg;
  P.B = B; P.D = D; P.N = N; P.N_data = N_data; P.nskip = nskip; P.L = L; P.dt = dt;
  P.obs_slot = obs.data(); P.Y = Y;
  P.rm_scalar = rm_scalar; P.rm_arr = rm_arr;
  P.rf_scalar = rf0_scalar * rf_scale; P.rf_arr = rf0_arr; P.rf_scale = rf_scale;
  P.stim = stim; P.S = S; P.NP = NP; P.NPest = NPest; P.pmap = pmap.data();
  P.pfix = pfix; P.pfix_stride = pfix_stride;
  P.Tseg = pl.Tseg; P.nseg = pl.nseg; P.TPR = pl.TPR; P.RG = pl.RG; P.nunits = pl.nunits;
  P.K = 2 + NP;
  P.upp = pl.nseg;
  P.Lp = L;
  std::vector<double> partials((size_t)pl.nunits * P.K, 0.0);
  P.partials = partials.data();
  P.active = active;
  P.cm = 1.0 / ((double)L * N_data);
  P.cf = 1.0 / ((double)D * (N - 1));
  EmulRun er{&P, pl.NT, pl.grid};
  int rc = ode_dispatch(model, pl.C, disc, er);
  if (rc != 0) return rc;
  const long long nX = (long long)N * D;
  for (int b = 0; b < B; ++b) {
    if (active && !active[b]) continue;
    const double m = ode_partial_sum(P, b, 0), f = ode_partial_sum(P, b, 1);
    if (me) me[b] = m;
    if (fe) fe[b] = f;
    if (A) A[b] = m + f;
    for (int k = 0; k < NP; ++k)
      if (pmap[k] >= 0 && G) G[b * ldg + nX + pmap[k]] = ode_partial_sum(P, b, 2 + k);
  }
  return 0;
}
