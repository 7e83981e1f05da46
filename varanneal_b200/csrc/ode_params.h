// Kernel argument block of the fused ODE action+gradient kernels (ode_sweep.cuh, ode_stream.cuh).
#pragma once

enum { DISC_EULER = 0, DISC_TRAPEZOID = 1, DISC_SIMPSON = 2, DISC_FORWARDMAP = 3, DISC_RK4 = 4 };

struct OdeParams {
  const double* XP;       // (B, ldxp) paths: X (N, D) row-major ++ estimated parameters
                          // (ptime: ++ (N, NPest) row-major, one parameter row per model time)
  long long ldxp;
  double* G;              // (B, ldg) gradient, same layout; nullptr = value only
  long long ldg;
  int B, D, N, N_data, nskip, L;
  double dt;
  // measurement data in the library's dense layout (vab_ode_problem_set / vab_ode_set_weights):
  const double* Y;        // (N_data, D): observations scattered to their state component, else 0
  const double* wobs;     // (D): 2 cm RM for observed components (scalar RM), else 0
  const double* rmd;      // (N_data, D): 2 cm RM[n, l] at observed components, else 0; or nullptr
  double rf_scalar;       // RF0 * scale when rf_arr == nullptr
  const double* rf_arr;   // RF0 (N-1, D) or nullptr
  const double* rf_mat;   // RF0 (N-1, D, D), matrix form (va_ode.py:211-223; SimpsonHermite only), or nullptr
  double rf_scale;
  double rf0;             // RF0 when it is a scalar (rf_scalar = rf0 * rf_scale)
  const double* rf_path;  // (B) or nullptr: per-path scale replacing rf_scale (asynchronous ladder:
                          // every path sits on its own rung, lbfgs.cu)
  const double* stim;     // (N, S) or nullptr
  int S;
  int NP, NPest;
  const int* pmap;        // (NP) -> index among the estimated parameters, or -1
  const double* pfix;     // values of the parameters that are not estimated
  long long pfix_stride;  // 0 (shared by all paths) or NP (ptime: N * NP)
  int ptime;              // 1 = the parameters are a time series (va_ode.py:568-570): pfix is
                          // (N, NP) per path, row n enters every evaluation of f at time n
  const int* active;      // (B) or nullptr: paths with active[b] == 0 are skipped
  double cm, cf;          // 1/(L N_data), 1/(D (N-1))
  // work decomposition (ode_plan.h)
  int TPR;                // strips per row
  int GW, GPW;            // lanes per group, groups (= paths) per warp
  int WS, nwin, NHL;      // output strips per window, windows per row, halo lanes either side
  int Tseg, nseg;         // time rows per segment (even), segments per path
  int wpb;                // warps per (segment, window) = ceil(B / GPW)
  int upp;                // units (segment x window) per path: their partials are contiguous
  int nunits;             // B * upp
  int K;                  // partial sums per unit: me, fe, NPM parameter-gradient entries
  double* partials;       // (nunits, K)
};
