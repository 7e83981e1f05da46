// Device vector fields, written on register "strips": a thread owns C consecutive components of
// one time row and sees them, plus H halo components either side, as a small local array
//   xh[0 .. C+2H)  <->  components i0-H .. i0+C+H-1   (indices mod D for the periodic model).
// Each model supplies
//   f   (xh, p, stim, out[C])            out_j = f_{i0+j}(x)
//   adj (xh, vh, p, jt[C], pacc[NPM])    jt_j = (J_f(x)^T v)_{i0+j};  pacc[k] += (df/dp_k)^T v
//                                        restricted to the strip's own components.
// Equations: Lorenz96  reference examples/Lorenz96_D20/Lorenz96_anneal.py:15-16
//            NaKL      reference examples/jupyter-tutorial/VarAnneal_tutorial.ipynb cell 36
//            Lorenz63  extension named by BASELINE.json (not in the reference)
// Adjoint formulas: SURVEY.md App. A.3, checked against complex-step through the reference's own
// NumPy action by the oracle (oracle/models_np.py carries the same formulas in NumPy).
#pragma once
#include "vab_hd.h"

template <int C_>
struct ModelL96 {
  static constexpr int C = C_;
  static constexpr int H = 2;      // x_{i-2} .. x_{i+2} are touched
  static constexpr int NPM = 1;    // forcing k
  static constexpr int NSTIM = 0;

  VAB_HD static void f(const double* xh, const double* p, const double* /*stim*/, double* out) {
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = j + H;
      out[j] = fma(xh[c - 1], xh[c + 1] - xh[c - 2], p[0] - xh[c]);
    }
  }
  VAB_HD static void adj(const double* xh, const double* vh, const double* /*p*/, double* jt,
                         double* pacc) {
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = j + H;
      double t = vh[c + 1] * (xh[c + 2] - xh[c - 1]);
      t = fma(vh[c - 1], xh[c - 2], t);
      t = fma(-vh[c + 2], xh[c + 1], t);
      jt[j] = t - vh[c];
      pacc[0] += vh[c];
    }
  }
  // the same without the diagonal term:  t_j = (J^T v)_j + v_j  (callers fold v into the store)
  VAB_HD static void adj_t(const double* xh, const double* vh, const double* /*p*/, double* t,
                           double* pacc) {
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = j + H;
      double u = vh[c + 1] * (xh[c + 2] - xh[c - 1]);
      u = fma(vh[c - 1], xh[c - 2], u);
      t[j] = fma(-vh[c + 2], xh[c + 1], u);
      pacc[0] += vh[c];
    }
  }
};

struct ModelL63 {
  static constexpr int C = 3;
  static constexpr int H = 0;
  static constexpr int NPM = 3;    // sigma, rho, beta
  static constexpr int NSTIM = 0;

  VAB_HD static void f(const double* x, const double* p, const double* /*stim*/, double* out) {
    out[0] = p[0] * (x[1] - x[0]);
    out[1] = x[0] * (p[1] - x[2]) - x[1];
    out[2] = x[0] * x[1] - p[2] * x[2];
  }
  VAB_HD static void adj(const double* x, const double* v, const double* p, double* jt,
                         double* pacc) {
    jt[0] = -p[0] * v[0] + (p[1] - x[2]) * v[1] + x[1] * v[2];
    jt[1] = p[0] * v[0] - v[1] + x[0] * v[2];
    jt[2] = -x[0] * v[1] - p[2] * v[2];
    pacc[0] += (x[1] - x[0]) * v[0];
    pacc[1] += x[0] * v[1];
    pacc[2] += -x[2] * v[2];
  }
};

struct ModelNaKL {
  static constexpr int C = 4;      // V, m, h, n
  static constexpr int H = 0;
  static constexpr int NPM = 18;   // gNa gK gL ENa EK EL | Vt Vs t1 t2 for m, h, n
  static constexpr int NSTIM = 1;  // injected current

  VAB_HD static void f(const double* x, const double* p, const double* stim, double* out) {
    const double V = x[0], m = x[1], h = x[2], n = x[3];
    const double n2 = n * n;
    out[0] = p[0] * (m * m * m) * h * (p[3] - V) + p[1] * (n2 * n2) * (p[4] - V) +
             p[2] * (p[5] - V) + (stim ? vab_ldg(stim) : 0.0);
#pragma unroll
    for (int c = 1; c <= 3; ++c) {
      const double* q = p + 2 + 4 * c;   // Vt, Vs, t1, t2
      const double T = tanh((V - q[0]) / q[1]);
      const double zinf = 0.5 * (1.0 + T);
      const double tau = q[2] + q[3] * (1.0 - T * T);
      out[c] = (zinf - x[c]) / tau;
    }
  }
  VAB_HD static void adj(const double* x, const double* v, const double* p, double* jt,
                         double* pacc) {
    const double V = x[0], m = x[1], h = x[2], n = x[3];
    const double m2 = m * m, m3 = m2 * m, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
    const double v0 = v[0];
    jt[0] = (-p[0] * m3 * h - p[1] * n4 - p[2]) * v0;
    jt[1] = 3.0 * p[0] * m2 * h * (p[3] - V) * v0;
    jt[2] = p[0] * m3 * (p[3] - V) * v0;
    jt[3] = 4.0 * p[1] * n3 * (p[4] - V) * v0;
    pacc[0] += m3 * h * (p[3] - V) * v0;
    pacc[1] += n4 * (p[4] - V) * v0;
    pacc[2] += (p[5] - V) * v0;
    pacc[3] += p[0] * m3 * h * v0;
    pacc[4] += p[1] * n4 * v0;
    pacc[5] += p[2] * v0;
#pragma unroll
    for (int c = 1; c <= 3; ++c) {
      const double* q = p + 2 + 4 * c;
      const double a = (V - q[0]) / q[1];
      const double T = tanh(a);
      const double sech2 = 1.0 - T * T;
      const double zinf = 0.5 * (1.0 + T);
      const double tau = q[2] + q[3] * sech2;
      const double itau2 = 1.0 / (tau * tau);
      const double dz = zinf - x[c];
      const double dz_a = (0.5 * sech2 * tau + dz * 2.0 * q[3] * T * sech2) * itau2;
      const double vc = v[c];
      jt[0] += (dz_a / q[1]) * vc;
      jt[c] += (-1.0 / tau) * vc;
      double* g = pacc + 2 + 4 * c;
      g[0] += dz_a * (-1.0 / q[1]) * vc;
      g[1] += dz_a * (-a / q[1]) * vc;
      g[2] += -dz * itau2 * vc;
      g[3] += -dz * sech2 * itau2 * vc;
    }
  }
};
