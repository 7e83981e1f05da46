/*
 * varanneal_b200 -- C ABI of libvarannealb200.so
 *
 * The reference (paulrozdeba/varanneal) is pure Python and has no FFI of its own; its seams are
 * Python methods.  This header declares the native entry points that replace the numeric body of
 * those seams, and for each one cites the reference interface it stands in for
 * (paths relative to the reference tree, varanneal/...).  The Python classes in
 * varanneal_b200/va_ode.py and varanneal_b200/va_nnet.py keep the reference's method names and
 * argument order and call these functions through ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain C: pointers, sizes, ints, doubles.  No torch / C++ types cross this boundary.
 *   - every function returns int: 0 = ok, <0 = vab_status error.  vab_last_error() gives text.
 *   - pointers named *_dev are DEVICE pointers owned by the caller (PyTorch allocates them);
 *     pointers named *_host are host pointers read during the call only.  The library never
 *     frees caller memory.  Internal workspaces belong to the context.
 *   - all work is enqueued on the context's CUDA stream.  Only vab_sync, vab_*_minimize and
 *     vab_*_anneal block the host (they poll a pinned completion flag).
 *   - a context is bound to one device and is not thread-safe; distinct contexts are independent
 *     (no global state -- unlike the reference's process-global ADOL-C tape ids).
 *   - batches: B independent paths ("annealing initialisations") share the problem data.
 *     Path b lives at XP_dev + b*ldxp (doubles), laid out exactly like the reference's flat
 *     XP = X.flatten() ++ P[Pidx]   (va_ode.py:715-732, va_nnet.py:468-473).
 *     ldxp must be even (16-byte aligned paths).
 */
#ifndef VARANNEAL_B200_H
#define VARANNEAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAB_ABI_VERSION 1

#if defined(__GNUC__)
#define VAB_API __attribute__((visibility("default")))
#else
#define VAB_API
#endif

typedef struct vab_ctx vab_ctx;

typedef enum {
  VAB_OK = 0,
  VAB_ERR_INVALID = -1,   /* bad argument / unsupported combination */
  VAB_ERR_CUDA = -2,      /* CUDA runtime error (text in vab_last_error) */
  VAB_ERR_STATE = -3,     /* call order (e.g. action before problem_set) */
  VAB_ERR_NOMEM = -4
} vab_status;

/* Device vector fields selectable through set_model (replaces the arbitrary Python callable of
 * va_ode.py:56-67; equations: examples/Lorenz96_D20/Lorenz96_anneal.py:15-16, tutorial
 * notebook cell 36 for NaKL; Lorenz63 is an extension named by BASELINE.json). */
typedef enum { VAB_MODEL_LORENZ96 = 0, VAB_MODEL_LORENZ63 = 1, VAB_MODEL_NAKL = 2 } vab_model;

/* Time discretisations: va_ode.py:341-356 (euler), :358-380 (trapezoid), :404-437
 * (SimpsonHermite, needs odd N_model), :439-454 (forwardmap); rk4 is an extension following the
 * commented-out intent at :382-402 (static parameters, no stimulus). */
typedef enum {
  VAB_DISC_EULER = 0, VAB_DISC_TRAPEZOID = 1, VAB_DISC_SIMPSON_HERMITE = 2,
  VAB_DISC_FORWARDMAP = 3, VAB_DISC_RK4 = 4
} vab_disc;

typedef enum { VAB_ACT_SIGMOID = 0, VAB_ACT_TANH = 1, VAB_ACT_LINEAR = 2 } vab_activation;

/* ---- context ------------------------------------------------------------------------------ */
VAB_API int vab_abi_version(void);
/* stream: a cudaStream_t passed as void* (NULL = the device's default stream). */
VAB_API int vab_ctx_create(int device, void* stream, vab_ctx** out);
VAB_API int vab_ctx_destroy(vab_ctx* ctx);
VAB_API int vab_sync(vab_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL: last error of a failed create). */
VAB_API const char* vab_last_error(const vab_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
VAB_API long long vab_launch_count(const vab_ctx* ctx);
/* Number of minimiser cycles (one line-search evaluation of every running path: 11-12 kernels) that
 * were replayed from the CUDA graph captured by vab_minimize / vab_anneal, as opposed to enqueued
 * kernel by kernel.  0 means the capture was refused or switched off (VAB_LBFGS_GRAPH=0). */
VAB_API long long vab_graph_launch_count(const vab_ctx* ctx);
/* Which kernels evaluated the last neural-network action on this context: 0 none yet, 1 fused
 * example-tile kernel, 2 per-layer fp64 tensor-pipe (DMMA) kernels, 3 all-layer DMMA tile kernel,
 * 4 CUDA-core register-tiled kernel (VAB_NN_SMALL=1), 5 tcgen05 / TMEM / TMA Ozaki-split contractions
 * (VAB_NN_TCGEN05=1; layers up to 128 wide).  No reference counterpart: a diagnostic the parity tests
 * use to assert that the path they mean to check is the one that ran. */
VAB_API int vab_nn_kernel_family(const vab_ctx* ctx);

/* Measures the device's double-precision FMA rate on the CUDA cores (TFLOP/s, 2 flops per FMA) with
 * a register-resident micro-kernel timed by CUDA events: the roofline denominator of the
 * fp64-bound kernels (rk4, NaKL, neural-network contractions), which MEASURED_PEAKS.json lacks. */
VAB_API int vab_measure_fp64_peak(vab_ctx* ctx, double* tflops_host);
/* The same for the fp64 tensor pipe: sustained rate of mma.sync.m8n8k4.f64 (DMMA), the instruction
 * the neural-network contractions issue -- their roofline denominator. */
VAB_API int vab_measure_fp64_dmma_peak(vab_ctx* ctx, double* tflops_host);

/* Measured prototype of the neural-network contraction on the 5th-generation tensor cores
 * (csrc/ozaki_gemm.cu): P independent products C_p = A_p B_p^T (A_p: M x K, B_p: N x K, fp64, K <= 128)
 * computed as an Ozaki split -- 7 int8 digit planes per operand, 28 exact int8 x int8 -> int32 plane
 * products on tcgen05.mma.kind::i8 with TMEM accumulators, operands brought in by TMA tensor maps,
 * recombined in fp64 -- on generated data with `spread` octaves of dynamic range inside a row, and
 * compared with an fp64 FMA reference.  out_host[12]: [max |C - Cref| / max |Cref|, ms digit planes,
 * ms tcgen05 kernel, ms total, fp64-equivalent TFLOP/s (total), the same for the tcgen05 kernel alone,
 * ms of the fp64 reference kernel, max |Cref|; then for the persistent, pipelined version of the
 * kernel: max relative error, ms, TFLOP/s-equivalent of the kernel, of kernel + digit planes].
 * Reference math: the layer contractions of
 * va_nnet.py:210-255. */
VAB_API int vab_ozaki_gemm_probe(vab_ctx* ctx, int32_t P, int32_t M, int32_t N, int32_t K, int32_t reps,
                                 double spread, double* out_host);

/* ---- ODE problem -------------------------------------------------------------------------- */
/* Everything va_ode.Annealer.anneal_init fixes for a run (va_ode.py:531-705). */
typedef struct {
  int32_t model;        /* vab_model */
  int32_t disc;         /* vab_disc */
  int32_t D;            /* state dimension (set_model, va_ode.py:56-67) */
  int32_t N_model;      /* (N_data-1)*nskip+1 (va_ode.py:557) */
  int32_t N_data;       /* rows of Y (va_ode.py:104-107) */
  int32_t nskip;        /* merr_nskip (va_ode.py:556) */
  int32_t L;            /* len(Lidx) (va_ode.py:577-578) */
  int32_t NP;           /* len(P0) (va_ode.py:564-570); must equal the model's parameter count */
  int32_t NPest;        /* len(Pidx) (va_ode.py:573-574) */
  int32_t n_stim;       /* columns of the stimulus (0 = none) (va_ode.py:113-124) */
  double dt_model;      /* va_ode.py:549-555 */
} vab_ode_desc;

/* Lidx_host[L], Pidx_host[NPest]: index lists as passed to anneal().
 * Y_dev: (N_data, L) row-major observations (va_ode.py:112,120).
 * stim_dev: (N_model, n_stim) row-major or NULL.
 * Y is copied into a library-owned buffer during the call (padded rows, columns sorted by
 * state component, the layout the kernels stream); the array behind stim_dev must stay alive
 * while the problem is set. */
VAB_API int vab_ode_problem_set(vab_ctx* ctx, const vab_ode_desc* desc,
                        const int32_t* Lidx_host, const int32_t* Pidx_host,
                        const double* Y_dev, const double* stim_dev);

/* RM / RF0 exactly as anneal_init normalises them (va_ode.py:612-640): a scalar, or a per-entry
 * array already broadcast to (N_data, L) / (N_model-1, D).  rm_dev / rf0_dev NULL = scalar.
 * An RM array is copied by the call; the array behind rf0_dev must stay alive.
 * The (.,L,L)/(.,D,D) matrix forms are unsupported (the reference's own branch is broken,
 * va_ode.py:222). */
VAB_API int vab_ode_set_weights(vab_ctx* ctx, double rm_scalar, const double* rm_dev,
                        double rf0_scalar, const double* rf0_dev);

/* Matrix form of RF0: rf0_mat_dev is (N_model-1, D, D) row-major -- a (D, D) matrix is repeated over
 * time by the caller, as va_ode.py:631-632 does; model error sum over the Simpson pairs of
 * e1_i . (RF[2i] e1_i) + e2_i . (RF[2i+1] e2_i), RF = RF0 * scale (va_ode.py:211-218).  SimpsonHermite
 * only (the reference's branch for the other discretisations does not run), rows inside one lane
 * group, static parameters.  Replaces the RF0 of vab_ode_set_weights (which clears it).  The array
 * must stay alive while it is set. */
VAB_API int vab_ode_set_rf_matrix(vab_ctx* ctx, const double* rf0_mat_dev);

/* Matrix form of RM: rm_dev is (N_data, L, L) row-major -- an (L, L) matrix is repeated over time by
 * the caller, as va_ode.py:616-617 does -- and the measurement error is
 * sum_i diff_i . (RM_i diff_i) / (L N_data) (va_ode.py:149-152; the matrix need not be symmetric).
 * Replaces the scalar / per-entry RM of vab_ode_set_weights (which in turn clears the matrix); RF0 is
 * left as it is.  The array must stay alive while it is set.  The term costs L^2 work per observed
 * entry in a kernel of its own behind the fused action kernels. */
VAB_API int vab_ode_set_rm_matrix(vab_ctx* ctx, const double* rm_dev);

/* Values of the parameters that are NOT estimated (va_ode.py:178-181): pfix_dev is (NP) shared by
 * all paths (pfix_stride = 0) or (B, NP) (pfix_stride = NP).  Estimated entries are ignored. */
VAB_API int vab_ode_set_fixed_params(vab_ctx* ctx, const double* pfix_dev, int64_t pfix_stride);

/* Parameter time series: P0 of shape (N_model, NP) (va_ode.py:568-570).  With enabled != 0 every
 * path is X (N_model, D) ++ P[:, Pidx] (N_model, NPest), both row-major (va_ode.py:170-188, 688-689),
 * n = N_model * (D + NPest); row n of the parameters enters every evaluation of f at model time n
 * (trapezoid va_ode.py:368-369, SimpsonHermite :416-418; euler / forwardmap, whose reference branches
 * do not run, use the same N_model rows and never read the last).  pfix_dev: (N_model, NP) values of
 * the parameters that are not estimated, shared (pfix_stride = 0) or per path (N_model * NP); may be
 * NULL when NPest == NP.  The array must stay alive while the problem is set.  Call after
 * vab_ode_problem_set (which resets to static parameters); rk4 and rows wider than one lane group
 * (lorenz96 D > 128) are refused.  enabled == 0 returns to static parameters (fixed values zeroed:
 * call vab_ode_set_fixed_params again). */
VAB_API int vab_ode_set_time_dependent(vab_ctx* ctx, int32_t enabled, const double* pfix_dev,
                                       int64_t pfix_stride);

/* Replaces ADmin.A_gradA_taped (_autodiffmin.py:57-58) -- and tape_A (:32-49), which has no
 * analogue -- for B paths at once, with RF = RF0 * rf_scale (rf_scale = alpha**beta,
 * va_ode.py:650,782).  Outputs (any may be NULL): A_dev[B] action, me_dev[B] measurement error
 * (va_ode.py:138-158), fe_dev[B] RF-weighted model error (:160-234), G_dev gradient with the same
 * layout / leading dimension convention as XP (ldg doubles between paths). */
VAB_API int vab_ode_action_grad(vab_ctx* ctx, int32_t B, const double* XP_dev, int64_t ldxp,
                        double rf_scale, double* A_dev, double* me_dev, double* fe_dev,
                        double* G_dev, int64_t ldg);

/* ---- neural-network problem ---------------------------------------------------------------- */
/* va_nnet.Annealer.set_structure / set_activation / set_input_data / set_output_data
 * (va_nnet.py:59-106) + the parts of anneal_init that fix the problem (:288-457).
 * structure_host[n_layers]; Lin_host[n_Lin] / Lout_host[n_Lout] = Lidx[0] / Lidx[1];
 * data_in_dev (M, n_Lin), data_out_dev (M, n_Lout) row-major; Pidx_host[NPest] indexes the flat
 * parameter vector [W_0 (d_1 x d_0 row-major), b_0, W_1, b_1, ...] (va_nnet.py:194-207). */
VAB_API int vab_nn_problem_set(vab_ctx* ctx, int32_t n_layers, const int32_t* structure_host, int32_t M,
                       int32_t activation,
                       int32_t n_Lin, const int32_t* Lin_host,
                       int32_t n_Lout, const int32_t* Lout_host,
                       const double* data_in_dev, const double* data_out_dev,
                       int32_t NPest, const int32_t* Pidx_host);
/* RM scalar or (2,) (va_nnet.py:131-144): pass rm_in == rm_out for the scalar form. RF0 scalar. */
VAB_API int vab_nn_set_weights(vab_ctx* ctx, double rm_in, double rm_out, double rf0);
/* Matrix form of RM for the network (va_nnet.py:135-139): rm_in_dev (n_Lin, n_Lin), rm_out_dev
 * (n_Lout, n_Lout), row-major, not necessarily symmetric; me = sum_m [din_m . (RM_in din_m) +
 * dout_m . (RM_out dout_m)] / (Ltot M).  Replaces the scalar weights of vab_nn_set_weights (which in
 * turn clears the matrices; RF0 is kept).  The arrays must stay alive while they are set. */
VAB_API int vab_nn_set_rm_matrices(vab_ctx* ctx, const double* rm_in_dev, const double* rm_out_dev);
VAB_API int vab_nn_set_fixed_params(vab_ctx* ctx, const double* pfix_dev, int64_t pfix_stride);
/* Replaces A_gradA_taped for va_nnet.A_gaussian (va_nnet.py:111-255). Same conventions. */
VAB_API int vab_nn_action_grad(vab_ctx* ctx, int32_t B, const double* XP_dev, int64_t ldxp,
                       double rf_scale, double* A_dev, double* me_dev, double* fe_dev,
                       double* G_dev, int64_t ldg);

/* ---- minimiser + ladder -------------------------------------------------------------------- */
/* Options forwarded from opt_args to scipy.optimize.minimize(method='L-BFGS-B')
 * (_autodiffmin.py:85-86): gtol -> pgtol, ftol -> ftol (factr = ftol/eps), maxfun, maxiter,
 * maxcor -> m, maxls. */
typedef struct {
  int32_t m;          /* history size, SciPy default 10; method 2: maxCGit (0 = max(1, min(50, n/2))) */
  int32_t maxls;      /* line-search steps per iteration, SciPy default 20 */
  int64_t maxfun;     /* SciPy default 15000 */
  int64_t maxiter;    /* SciPy default 15000 */
  double ftol;        /* stop when (f_k - f_{k+1})/max(|f_k|,|f_{k+1}|,1) <= ftol */
  double pgtol;       /* stop when max|proj g| <= pgtol */
  int32_t poll_every; /* evaluations enqueued between host polls of the done flag (>=1) */
  int32_t method;     /* 0 = L-BFGS-B (min_lbfgs_scipy), 1 = nonlinear CG, Polak-Ribiere+ (min_cg_scipy,
                         _autodiffmin.py:97-119; ftol / m are ignored), 2 = truncated Newton
                         (min_tnc_scipy, _autodiffmin.py:121-143; vab_minimize only; bounds through an active set;
                         status = SciPy's TNC return code: 0 local minimum, 1 f converged,
                         2 x converged, 3 evaluation limit, 4 line search failed) */
} vab_lbfgs_opts;

/* Replaces ADmin.min_lbfgs_scipy (_autodiffmin.py:72-95) for whichever problem (ODE or NN) was
 * set last on this context: minimise A(.; RF0*rf_scale) from XP_dev (overwritten by the
 * minimiser) for B paths concurrently, entirely on the device (batched L-BFGS-B with per-path
 * convergence masks).  lo_dev / hi_dev: (n) bounds shared by all paths or NULL (= unbounded),
 * with +-inf for one-sided (va_ode.py:582-605, _autodiffmin.py:86).
 * Outputs [B] each (device, may be NULL): A/me/fe at the minimiser, status (SciPy warnflag:
 * 0 converged, 1 maxfun/maxiter, 2 abnormal), nit iterations, nfev evaluations. */
VAB_API int vab_minimize(vab_ctx* ctx, int32_t B, double* XP_dev, int64_t ldxp, double rf_scale,
                 const vab_lbfgs_opts* opts, const double* lo_dev, const double* hi_dev,
                 double* A_dev, double* me_dev, double* fe_dev,
                 int32_t* status_dev, int32_t* nit_dev, int32_t* nfev_dev);

/* Replaces the beta loop of Annealer.anneal + anneal_step (va_ode.py:473-490, 707-789;
 * va_nnet.py:281-286, 459-523): for i in range(Nbeta): minimise at RF0*alpha**beta[i] warm-started
 * from the previous minimiser.  table_dev: (B, Nbeta, 5) rows [beta, A, me, fe, fe/(alpha**beta)]
 * -- the caller divides column 4 by RF0 (va_ode.py:847-873).  minpaths_dev: (B, Nbeta, ldxp) or
 * NULL; status/nit/nfev: (B, Nbeta) or NULL.
 * The ladder is asynchronous across paths: a path that has converged on rung i starts rung i+1
 * at once instead of waiting for the slowest path of the batch (the paths are independent, so the
 * results are those of the rung-by-rung loop, bit for bit). */
VAB_API int vab_anneal(vab_ctx* ctx, int32_t B, double* XP_dev, int64_t ldxp, double alpha,
               const double* beta_host, int32_t Nbeta, const vab_lbfgs_opts* opts,
               const double* lo_dev, const double* hi_dev,
               double* table_dev, double* minpaths_dev,
               int32_t* status_dev, int32_t* nit_dev, int32_t* nfev_dev);

/* Host sink for the minimising paths of the next vab_anneal call (which must be given a
 * minpaths_dev buffer): as soon as a path has finished rung i, its minimiser (the first `width`
 * doubles of the row) is copied to host_dst + (b * Nbeta + i) * host_pitch on a second stream while
 * the other paths keep annealing, so the host array the reference fills rung by rung
 * (minpaths, va_ode.py:776) is complete when vab_anneal returns and the transfer costs no wall
 * time.  host_dst may be pageable.  The sink is cleared by the call that used it; NULL clears it. */
VAB_API int vab_set_path_sink(vab_ctx* ctx, double* host_dst, int64_t host_pitch, int64_t width);

/* Column window of the minimising paths kept by the next vab_anneal call: minpaths_dev then holds
 * only columns [first, first + width) of every minimiser, as (B, Nbeta, pitch) with
 * pitch = max(2, width rounded up to even) doubles.  The reference keeps every path of every rung
 * (minpaths (Nbeta, N*D+NP), va_ode.py:666-667) -- 16 GB per initialisation at
 * D = 1000, N = 100000, 20 betas -- so a run of that size keeps the per-rung *parameter* estimates
 * (first = N*D, width = NPest) and only the last rung's path, which vab_anneal leaves in XP_dev.
 * width < 0 restores whole paths.  Consumed (cleared) by the next vab_anneal / vab_minimize; cannot
 * be combined with a host sink. */
VAB_API int vab_set_path_window(vab_ctx* ctx, int64_t first, int64_t width);

/* Strided device -> host copy of `rows` rows of `width` doubles (pitches in doubles), on the
 * context's stream, synchronous for the caller.  Used by the host mirror to lay the device
 * result buffers of vab_anneal out as the reference's minpaths array (va_ode.py:666-667, 776:
 * rows X ++ full P, so the host pitch differs from ldxp).  host_dst may be pageable. */
VAB_API int vab_copy_rows_to_host(vab_ctx* ctx, double* host_dst, int64_t host_pitch,
                          const double* src_dev, int64_t dev_pitch, int64_t width, int64_t rows);

/* Asynchronous strided copy of `rows` rows of `width` doubles between host and device (pitches in
 * doubles) on `stream` (a cudaStream_t passed as void*; NULL = the context's stream).
 * to_device != 0: host_ptr -> dev_ptr, else dev_ptr -> host_ptr.  One DMA with both pitches, no
 * staging: the host mirror uses it to move (B, n) host batches to / from the (B, ldxp) device
 * buffers of the eval seam (A_gradA_taped, _autodiffmin.py:57-58) while the kernels of the
 * neighbouring groups of paths run.  host_ptr should be pinned for the copy to be asynchronous. */
VAB_API int vab_copy_rows_async(vab_ctx* ctx, int32_t to_device, double* dev_ptr, int64_t dev_pitch,
                        double* host_ptr, int64_t host_pitch, int64_t width, int64_t rows,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VARANNEAL_B200_H */
