// Launch geometry of the fused ODE kernels, shared by the plan (ode_action.cu) and the problem
// set-up (vab_api.cu).
#pragma once

// model ids as in include/varanneal_b200.h
inline int ode_model_npm(int model) { return model == 0 ? 1 : (model == 1 ? 3 : (model == 2 ? 18 : -1)); }
inline int ode_model_nstim(int model) { return model == 2 ? 1 : 0; }

// ---- lane geometry shared by the sweep and stream kernels (see ode_sweep.cuh / ode_stream.cuh)
struct OdeGeo {
  int C, H;      // strip width, stencil halo (components)
  int TPR;       // strips per row
  int GW, GPW;   // lanes per group, groups per warp
  int WS, nwin;  // output strips per window, windows per row
  int NHL;       // halo lanes either side of a window (0: whole row in one group, periodic wrap)
};

inline int ode_geometry(int model, int disc, int D, OdeGeo* g) {
  if (model == 0) {
    if (D < 4) return -1;
    g->C = (D % 4 == 0) ? 4 : ((D % 2 == 0) ? 2 : 1);
    g->H = 2;
  } else if (model == 1) {
    if (D != 3) return -1;
    g->C = 3; g->H = 0;
  } else if (model == 2) {
    if (D != 4) return -1;
    g->C = 4; g->H = 0;
  } else {
    return -1;
  }
  g->TPR = D / g->C;
  if (g->H == 0) {
    g->GW = 1; g->WS = 1; g->nwin = 1; g->NHL = 0;
  } else if (g->TPR <= 32) {
    g->GW = g->TPR; g->WS = g->TPR; g->nwin = 1; g->NHL = 0;
  } else {
    const int need = (disc == 4) ? 12 : 3;     // components of validity lost per side (rk4 : others)
    g->NHL = (need + g->C - 1) / g->C;
    const int usable = 32 - 2 * g->NHL;
    if (usable < 1) return -1;
    g->nwin = (g->TPR + usable - 1) / usable;
    g->WS = (g->TPR + g->nwin - 1) / g->nwin;
    g->GW = g->WS + 2 * g->NHL;
  }
  g->GPW = 32 / g->GW;
  return 0;
}

// the TMA stream kernels exist for the Lorenz96 stencil with 16-byte strips (all discretisations;
// rk4 since the ring version of the kernel, stream_rk4_kernel)
inline bool ode_stream_supported(int model, int disc, const OdeGeo& g) {
  (void)disc;
  return model == 0 && (g.C == 4 || g.C == 2);
}
