// Internal context layout of libvarannealb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/varanneal_b200.h"
#include "ode_params.h"
#include "ode_plan.h"

struct NnProblem;   // nn_action.h
struct LbfgsWork;   // lbfgs.cu
struct TncWork;     // tnc.cu
struct OzakiWork;   // ozaki_gemm.cu

enum { VAB_PROBLEM_NONE = 0, VAB_PROBLEM_ODE = 1, VAB_PROBLEM_NN = 2 };

struct vab_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  std::string err;
  long long launches = 0;
  long long graph_launches = 0;     // cycles replayed from the captured CUDA graph (vab_graph_launch_count)
  // cudaFuncSetAttribute opt-ins act on the current device: remembered per context, not per process
  bool attr_lbfgs = false, attr_nn_split = false, attr_nn_fused = false, attr_nn_small = false;
  int problem = VAB_PROBLEM_NONE;

  // ---- ODE problem (vab_ode_problem_set / set_weights / set_fixed_params)
  vab_ode_desc od{};
  int* pmap_dev = nullptr;          // (NP)
  int* lcomp_dev = nullptr;         // (L) state component of each observed column
  const double* Y_src = nullptr;    // caller's (N_data, L) observations (read during set-up only)
  double* Y_dense = nullptr;        // (N_data, D) library copy, zero where unobserved
  double* rm_dense = nullptr;       // (N_data, D) 2 cm RM[n, l] when RM is an array
  double* wobs_dev = nullptr;       // (D) 2 cm RM for observed components (scalar RM)
  size_t Y_cap = 0, rm_cap = 0;     // doubles
  const double* stim_dev = nullptr;
  double rm_scalar = 1.0;
  const double* rm_dev = nullptr;   // = rm_dense when RM is an array
  const double* rf0_mat = nullptr;    // (N_model-1, D, D) matrix form of RF0 (vab_ode_set_rf_matrix), caller-owned
  const double* rm_matrix = nullptr;  // (N_data, L, L) matrix form of RM (vab_ode_set_rm_matrix), caller-owned
  double rf0_scalar = 1.0;
  const double* rf0_dev = nullptr;
  const double* pfix_dev = nullptr;
  long long pfix_stride = 0;
  int ptime = 0;                    // parameters are a time series (vab_ode_set_time_dependent)
  double* pfix_zero = nullptr;      // default fixed-parameter block (zeros)
  int tseg_override = 0;            // tuning knob (env VAB_TSEG)
  bool use_sweep = false;           // env VAB_KERNEL=sweep: register sweep kernels even where the stream kernels apply

  // ---- NN problem
  NnProblem* nn = nullptr;

  // ---- host sink of vab_anneal's minimising paths (vab_set_path_sink); consumed by the next call
  double* sink_host = nullptr;
  long long sink_pitch = 0, sink_width = 0;
  long long win0 = 0, winw = -1;    // vab_set_path_window (winw < 0: whole paths); consumed by the next call too

  // ---- workspaces
  double* partials = nullptr;
  size_t partials_cap = 0;          // doubles
  LbfgsWork* lb = nullptr;
  TncWork* tn = nullptr;
  OzakiWork* oz = nullptr;          // digit planes of the tcgen05 contraction path (VAB_NN_TCGEN05=1)
  int nn_family = 0;                // kernels of the last NN evaluation: 1 fused, 2 per-layer DMMA, 3 all-layer DMMA, 4 CUDA cores, 5 tcgen05

  long long n_unknowns() const;     // per path, for the problem currently set
};

int vab_fail(vab_ctx* ctx, int code, const std::string& msg);
int vab_cuda_fail(vab_ctx* ctx, cudaError_t e, const char* where);
// grow-only device buffer
int vab_reserve(vab_ctx* ctx, double** buf, size_t* cap, size_t need);

// objective evaluation for the problem currently set (ODE or NN); used by the minimiser.
// active_dev: (B) int mask or nullptr.  rf_path_dev: (B) per-path scale of RF0 (replaces rf_scale)
// or nullptr.
// fp64 contraction on tcgen05 / TMEM / TMA (ozaki_gemm.cu): C[p][r][c] = sum_k A[p][r][k] B[p][c][k], K <= 128
int ozaki_gemm(vab_ctx* ctx, int P, int M, int N, int K,
               const double* A, long long lda, long long aks, long long aps,
               const double* B, long long ldb, long long bks, long long bps,
               double* C, long long ldc, long long cps, const int* active_dev);
int ozaki_gemm_chunked(vab_ctx* ctx, int paths, int nchunk, int M, int N, int K, int Ktotal,
                       const double* A, long long lda, long long aks, long long aps, long long apc,
                       const double* B, long long ldb, long long bks, long long bps, long long bpc,
                       double* C, long long ldc, long long cps, const int* active_dev);
void ozaki_destroy(vab_ctx* ctx);

int vab_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
             const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
             double* G, long long ldg);
