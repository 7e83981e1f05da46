// C ABI of libvarannealb200.so: context, ODE problem set-up, ODE action+gradient.
// (NN entry points: nn_action.cu; minimiser / ladder: lbfgs.cu.)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ode_action.h"
#include "vab_ctx.h"

void nn_destroy(vab_ctx* ctx);       // nn_action.cu
void lbfgs_destroy(vab_ctx* ctx);    // lbfgs.cu
void tnc_destroy(vab_ctx* ctx);      // tnc.cu
int nn_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
            const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
            double* G, long long ldg);
long long nn_unknowns(const vab_ctx* ctx);

static thread_local std::string g_create_error;

int vab_fail(vab_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}
int vab_cuda_fail(vab_ctx* ctx, cudaError_t e, const char* where) {
  return vab_fail(ctx, VAB_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
int vab_reserve(vab_ctx* ctx, double** buf, size_t* cap, size_t need) {
  if (need <= *cap) return VAB_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void**)buf, need * sizeof(double));
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "cudaMalloc(workspace)");
  *cap = need;
  return VAB_OK;
}

// dense (rows, D) <- compact (rows, L): column l goes to state component comp[l], times `scale`
// (padding zeroed beforehand)
__global__ void vab_scatter_obs_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                       const int* __restrict__ comp, long long rows, int L, int D,
                                       double scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const long long r = i / L;
  const int l = (int)(i - r * L);
  dst[r * D + comp[l]] = scale * src[i];
}

// Matrix RM (va_ode.py:149-152): me = cm sum_i diff_i . (RM_i diff_i), gradient cm (RM_i + RM_i') diff_i.
// The quadratic form couples the observed components of a row, which the walk kernels spread over
// lanes; it is added here, after the walk + finalize kernels (which run with zero measurement
// weights): one CTA per path, thread per (data row, observed column), fixed-order reduction.
// rm: (N_data, L, L) row-major.  Meant for the small problems this form of RM is used on (L^2 work
// per observed entry).
__global__ void __launch_bounds__(256) ode_me_matrix_kernel(const double* XP, long long ldxp, double* G,
                                                            long long ldg, const double* Yd, const int* lcomp,
                                                            const double* rm, int N_data, int L, int D, int nskip,
                                                            double cm, const int* active, double* A, double* me) {
  const int b = blockIdx.x;
  if (active != nullptr && active[b] == 0) return;
  const double* x = XP + (long long)b * ldxp;
  double* g = G ? G + (long long)b * ldg : nullptr;
  double acc = 0.0;
  for (long long t = threadIdx.x; t < (long long)N_data * L; t += 256) {
    const long long i = t / L;
    const int l = (int)(t - i * L);
    const long long xrow = i * nskip * D, yrow = i * D;
    const double* R = rm + i * L * L;
    double w = 0.0, wt = 0.0;
    for (int k = 0; k < L; ++k) {
      const int c = lcomp[k];
      const double dk = x[xrow + c] - Yd[yrow + c];
      w = fma(R[(long long)l * L + k], dk, w);
      wt = fma(R[(long long)k * L + l], dk, wt);
    }
    const int cl = lcomp[l];
    const double dl = x[xrow + cl] - Yd[yrow + cl];
    acc = fma(dl, w, acc);
    if (g) g[xrow + cl] += cm * (w + wt);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int sft = 128; sft > 0; sft >>= 1) {
    if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double v = cm * red[0];
    if (A) A[b] += v;
    if (me) me[b] += v;
  }
}

// src == nullptr: fill one row with `scale` at the observed components (the scalar-RM weights)
static int vab_scatter_obs(vab_ctx* ctx, const double* src, long long rows, double scale,
                           double** dst, size_t* cap) {
  const vab_ode_desc& d = ctx->od;
  const size_t need = (size_t)rows * d.D + 2;
  int rc = vab_reserve(ctx, dst, cap, need);
  if (rc != VAB_OK) return rc;
  cudaError_t e = cudaMemsetAsync(*dst, 0, need * sizeof(double), ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "dense observation layout (memset)");
  const long long n = rows * d.L;
  if (n > 0 && src != nullptr) {
    vab_scatter_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        src, *dst, ctx->lcomp_dev, rows, d.L, d.D, scale);
    e = cudaGetLastError();
    if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "dense observation layout");
    ctx->launches += 1;
  }
  return VAB_OK;
}

long long vab_ctx::n_unknowns() const {
  if (problem == VAB_PROBLEM_ODE) return (long long)od.N_model * od.D + (long long)od.NPest * (ptime ? od.N_model : 1);
  if (problem == VAB_PROBLEM_NN) return nn_unknowns(this);
  return 0;
}

extern "C" {

int vab_abi_version(void) { return VAB_ABI_VERSION; }

const char* vab_last_error(const vab_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

long long vab_launch_count(const vab_ctx* ctx) { return ctx ? ctx->launches : 0; }

long long vab_graph_launch_count(const vab_ctx* ctx) { return ctx ? ctx->graph_launches : 0; }

int vab_nn_kernel_family(const vab_ctx* ctx) { return ctx ? ctx->nn_family : 0; }

int vab_ctx_create(int device, void* stream, vab_ctx** out) {
  if (!out) return vab_fail(nullptr, VAB_ERR_INVALID, "vab_ctx_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return vab_fail(nullptr, VAB_ERR_CUDA,
                    std::string("vab_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                        "); this library has no CPU path");
  if (device < 0 || device >= ndev)
    return vab_fail(nullptr, VAB_ERR_INVALID, "vab_ctx_create: bad device index");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return vab_cuda_fail(nullptr, e, "cudaSetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return vab_cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major < 10)
    return vab_fail(nullptr, VAB_ERR_INVALID,
                    "vab_ctx_create: built for sm_100a (B200); device is sm_" +
                        std::to_string(prop.major) + std::to_string(prop.minor));
  vab_ctx* c = new vab_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->num_sms = prop.multiProcessorCount;
  if (const char* t = getenv("VAB_TSEG")) c->tseg_override = atoi(t);
  if (const char* t = getenv("VAB_KERNEL")) c->use_sweep = (strcmp(t, "sweep") == 0);
  e = cudaMalloc((void**)&c->pfix_zero, 64 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(c->pfix_zero, 0, 64 * sizeof(double), c->stream);
  if (e != cudaSuccess) {
    delete c;
    return vab_cuda_fail(nullptr, e, "cudaMalloc(ctx)");
  }
  *out = c;
  return VAB_OK;
}

int vab_ctx_destroy(vab_ctx* ctx) {
  if (!ctx) return VAB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nn_destroy(ctx);
  lbfgs_destroy(ctx);
  ozaki_destroy(ctx);
  tnc_destroy(ctx);
  cudaFree(ctx->pmap_dev);
  cudaFree(ctx->lcomp_dev);
  cudaFree(ctx->Y_dense);
  cudaFree(ctx->rm_dense);
  cudaFree(ctx->wobs_dev);
  cudaFree(ctx->pfix_zero);
  cudaFree(ctx->partials);
  delete ctx;
  return VAB_OK;
}

int vab_sync(vab_ctx* ctx) {
  if (!ctx) return VAB_ERR_INVALID;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "cudaStreamSynchronize");
  return VAB_OK;
}

int vab_copy_rows_async(vab_ctx* ctx, int32_t to_device, double* dev_ptr, int64_t dev_pitch,
                        double* host_ptr, int64_t host_pitch, int64_t width, int64_t rows, void* stream) {
  if (!ctx) return VAB_ERR_INVALID;
  if (!dev_ptr || !host_ptr || width < 0 || rows < 0 || host_pitch < width || dev_pitch < width)
    return vab_fail(ctx, VAB_ERR_INVALID, "copy_rows_async: bad arguments");
  if (width == 0 || rows == 0) return VAB_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  cudaError_t e;
  if (to_device)
    e = cudaMemcpy2DAsync(dev_ptr, (size_t)dev_pitch * sizeof(double), host_ptr, (size_t)host_pitch * sizeof(double),
                          (size_t)width * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice, st);
  else
    e = cudaMemcpy2DAsync(host_ptr, (size_t)host_pitch * sizeof(double), dev_ptr, (size_t)dev_pitch * sizeof(double),
                          (size_t)width * sizeof(double), (size_t)rows, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "copy_rows_async");
  return VAB_OK;
}

int vab_set_path_sink(vab_ctx* ctx, double* host_dst, int64_t host_pitch, int64_t width) {
  if (!ctx) return VAB_ERR_INVALID;
  if (host_dst != nullptr && (width < 1 || host_pitch < width))
    return vab_fail(ctx, VAB_ERR_INVALID, "set_path_sink: bad pitch / width");
  ctx->sink_host = host_dst;
  ctx->sink_pitch = host_dst ? host_pitch : 0;
  ctx->sink_width = host_dst ? width : 0;
  return VAB_OK;
}

int vab_set_path_window(vab_ctx* ctx, int64_t first, int64_t width) {
  if (!ctx) return VAB_ERR_INVALID;
  if (width >= 0 && first < 0) return vab_fail(ctx, VAB_ERR_INVALID, "set_path_window: negative first column");
  ctx->win0 = width >= 0 ? first : 0;
  ctx->winw = width >= 0 ? width : -1;
  return VAB_OK;
}

int vab_copy_rows_to_host(vab_ctx* ctx, double* host_dst, int64_t host_pitch, const double* src_dev,
                          int64_t dev_pitch, int64_t width, int64_t rows) {
  if (!ctx) return VAB_ERR_INVALID;
  if (!host_dst || !src_dev || width < 0 || rows < 0 || host_pitch < width || dev_pitch < width)
    return vab_fail(ctx, VAB_ERR_INVALID, "copy_rows_to_host: bad arguments");
  if (width == 0 || rows == 0) return VAB_OK;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMemcpy2DAsync(host_dst, (size_t)host_pitch * sizeof(double), src_dev,
                                    (size_t)dev_pitch * sizeof(double), (size_t)width * sizeof(double),
                                    (size_t)rows, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "copy_rows_to_host");
  return VAB_OK;
}

int vab_ode_problem_set(vab_ctx* ctx, const vab_ode_desc* d, const int32_t* Lidx_host,
                        const int32_t* Pidx_host, const double* Y_dev, const double* stim_dev) {
  if (!ctx || !d) return VAB_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (d->D < 1 || d->N_model < 2 || d->N_data < 1 || d->nskip < 1 || d->L < 0 || d->NPest < 0 ||
      d->NPest > d->NP)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: bad sizes");
  if (d->N_model != (d->N_data - 1) * d->nskip + 1)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: N_model != (N_data-1)*nskip+1");
  if (d->disc < VAB_DISC_EULER || d->disc > VAB_DISC_RK4)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: unknown discretisation");
  if (d->disc == VAB_DISC_SIMPSON_HERMITE && d->N_model % 2 == 0)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: SimpsonHermite needs odd N_model");
  const int npm = ode_model_npm(d->model);
  if (npm < 0) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: unknown model");
  if (d->NP != npm)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: model expects " + std::to_string(npm) +
                                              " parameters, got " + std::to_string(d->NP));
  if (d->n_stim != 0 && d->n_stim < ode_model_nstim(d->model))
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: stimulus has too few columns");
  if (d->n_stim != 0 && !stim_dev)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: n_stim > 0 but stim_dev is NULL");
  if (d->disc == VAB_DISC_RK4 && d->n_stim != 0)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: rk4 (extension) takes no stimulus");
  if (d->L > 0 && (!Lidx_host || !Y_dev))
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: Lidx/Y missing");
  OdeGeo geo;
  if (ode_geometry(d->model, d->disc, d->D, &geo) != 0)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: D not valid for this model");
  std::vector<int> seen(d->D, 0), pmap(d->NP, -1), lcomp(d->L > 0 ? d->L : 1, 0);
  for (int l = 0; l < d->L; ++l) {
    const int i = Lidx_host[l];
    if (i < 0 || i >= d->D) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: Lidx out of range");
    if (seen[i]) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: duplicate Lidx entry");
    seen[i] = 1;
    lcomp[l] = i;
  }
  for (int e = 0; e < d->NPest; ++e) {
    const int k = Pidx_host[e];
    if (k < 0 || k >= d->NP) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: Pidx out of range");
    if (pmap[k] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: duplicate Pidx entry");
    pmap[k] = e;
  }
  cudaStreamSynchronize(ctx->stream);
  // from here on the old problem is being taken apart: the context holds no ODE problem until the
  // new one is complete (a failure below leaves later calls answering VAB_ERR_STATE)
  if (ctx->problem == VAB_PROBLEM_ODE) ctx->problem = VAB_PROBLEM_NONE;
  cudaFree(ctx->pmap_dev);
  cudaFree(ctx->lcomp_dev);
  cudaFree(ctx->wobs_dev);
  ctx->pmap_dev = nullptr;
  ctx->lcomp_dev = nullptr;
  ctx->wobs_dev = nullptr;
  cudaError_t e = cudaMalloc((void**)&ctx->pmap_dev, sizeof(int) * (d->NP > 0 ? d->NP : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->lcomp_dev, sizeof(int) * lcomp.size());
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->wobs_dev, sizeof(double) * (d->D + 2));
  if (e == cudaSuccess && d->NP > 0)
    e = cudaMemcpy(ctx->pmap_dev, pmap.data(), sizeof(int) * d->NP, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->lcomp_dev, lcomp.data(), sizeof(int) * lcomp.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ode_problem_set");
  ctx->od = *d;
  // dense copy of the observations: the layout the kernels stream (same columns as X)
  {
    int rc = vab_scatter_obs(ctx, Y_dev, d->N_data, 1.0, &ctx->Y_dense, &ctx->Y_cap);
    if (rc != VAB_OK) return rc;
  }
  ctx->stim_dev = (d->n_stim > 0) ? stim_dev : nullptr;
  ctx->pfix_dev = ctx->pfix_zero;
  ctx->pfix_stride = 0;
  ctx->ptime = 0;
  ctx->rm_matrix = nullptr;
  ctx->rf0_mat = nullptr;
  ctx->rf0_scalar = 1.0; ctx->rf0_dev = nullptr;
  ctx->problem = VAB_PROBLEM_ODE;
  const int rcw = vab_ode_set_weights(ctx, 1.0, nullptr, 1.0, nullptr);
  if (rcw != VAB_OK) ctx->problem = VAB_PROBLEM_NONE;
  return rcw;
}

int vab_ode_set_weights(vab_ctx* ctx, double rm_scalar, const double* rm_dev, double rf0_scalar,
                        const double* rf0_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_weights: no ODE problem set");
  cudaSetDevice(ctx->device);
  const vab_ode_desc& d = ctx->od;
  const double cm2 = d.L > 0 ? 2.0 / ((double)d.L * d.N_data) : 0.0;
  ctx->rm_scalar = rm_scalar; ctx->rm_dev = nullptr;
  ctx->rm_matrix = nullptr;
  {  // weights of the measurement term in the dense layout: 2 cm RM at the observed components
    size_t cap = (size_t)d.D + 2;
    int rc = vab_scatter_obs(ctx, nullptr, 1, 0.0, &ctx->wobs_dev, &cap);   // zero fill
    if (rc != VAB_OK) return rc;
    if (d.L > 0) {
      std::vector<double> ones(d.L, cm2 * rm_scalar);
      double* tmp = nullptr;
      cudaError_t e = cudaMalloc((void**)&tmp, sizeof(double) * d.L);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, ones.data(), sizeof(double) * d.L, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) { cudaFree(tmp); return vab_cuda_fail(ctx, e, "set_weights"); }
      rc = vab_scatter_obs(ctx, tmp, 1, 1.0, &ctx->wobs_dev, &cap);
      cudaStreamSynchronize(ctx->stream);
      cudaFree(tmp);
      if (rc != VAB_OK) return rc;
    }
  }
  if (rm_dev) {
    int rc = vab_scatter_obs(ctx, rm_dev, d.N_data, cm2, &ctx->rm_dense, &ctx->rm_cap);
    if (rc != VAB_OK) return rc;
    ctx->rm_dev = ctx->rm_dense;
  }
  ctx->rf0_scalar = rf0_scalar; ctx->rf0_dev = rf0_dev;
  ctx->rf0_mat = nullptr;
  return VAB_OK;
}

int vab_ode_set_rf_matrix(vab_ctx* ctx, const double* rf0_mat_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_rf_matrix: no ODE problem set");
  if (!rf0_mat_dev) return vab_fail(ctx, VAB_ERR_INVALID, "set_rf_matrix: NULL");
  const vab_ode_desc& d = ctx->od;
  if (d.disc != VAB_DISC_SIMPSON_HERMITE)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_rf_matrix: the matrix form of RF exists for SimpsonHermite only "
                                          "(the reference's branch for the other discretisations does not run, va_ode.py:222)");
  OdeGeo geo;
  if (ode_geometry(d.model, d.disc, d.D, &geo) != 0 || geo.nwin > 1)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_rf_matrix: needs a row that fits one lane group (D <= 128 for lorenz96)");
  if (ctx->ptime) return vab_fail(ctx, VAB_ERR_INVALID, "set_rf_matrix: not together with a parameter time series");
  ctx->rf0_mat = rf0_mat_dev;
  ctx->rf0_scalar = 1.0; ctx->rf0_dev = nullptr;
  return VAB_OK;
}

int vab_ode_set_fixed_params(vab_ctx* ctx, const double* pfix_dev, int64_t pfix_stride) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_fixed_params: no ODE problem set");
  if (!pfix_dev) return vab_fail(ctx, VAB_ERR_INVALID, "set_fixed_params: NULL");
  const long long per_path = (long long)ctx->od.NP * (ctx->ptime ? ctx->od.N_model : 1);
  if (pfix_stride != 0 && pfix_stride != per_path)
    return vab_fail(ctx, VAB_ERR_INVALID, ctx->ptime ? "set_fixed_params: stride must be 0 or N_model * NP"
                                                     : "set_fixed_params: stride must be 0 or NP");
  ctx->pfix_dev = pfix_dev;
  ctx->pfix_stride = pfix_stride;
  return VAB_OK;
}

int vab_ode_set_rm_matrix(vab_ctx* ctx, const double* rm_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_rm_matrix: no ODE problem set");
  if (!rm_dev) return vab_fail(ctx, VAB_ERR_INVALID, "set_rm_matrix: NULL");
  if (ctx->od.L < 1) return vab_fail(ctx, VAB_ERR_INVALID, "set_rm_matrix: nothing is observed");
  // the walk kernels carry no measurement term any more; RF0 stays what it was
  const double rf0 = ctx->rf0_scalar;
  const double* rf0d = ctx->rf0_dev;
  const double* rf0m = ctx->rf0_mat;
  const int rc = vab_ode_set_weights(ctx, 0.0, nullptr, rf0, rf0d);
  if (rc != VAB_OK) return rc;
  ctx->rf0_mat = rf0m;
  ctx->rm_matrix = rm_dev;
  return VAB_OK;
}

int vab_ode_set_time_dependent(vab_ctx* ctx, int32_t enabled, const double* pfix_dev, int64_t pfix_stride) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_time_dependent: no ODE problem set");
  const vab_ode_desc& d = ctx->od;
  if (!enabled) {
    ctx->ptime = 0;
    ctx->pfix_dev = ctx->pfix_zero;
    ctx->pfix_stride = 0;
    return VAB_OK;
  }
  if (d.disc == VAB_DISC_RK4)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_time_dependent: rk4 (extension) takes static parameters only");
  if (ctx->rf0_mat) return vab_fail(ctx, VAB_ERR_INVALID, "set_time_dependent: not together with a matrix RF");
  OdeGeo geo;
  if (ode_geometry(d.model, d.disc, d.D, &geo) != 0 || geo.nwin > 1)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_time_dependent: a parameter time series needs a row that fits "
                                          "one lane group (D <= 128 for lorenz96)");
  const long long per_path = (long long)d.NP * d.N_model;
  if (!pfix_dev && d.NPest < d.NP)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_time_dependent: the parameters that are not estimated need "
                                          "their (N_model, NP) values");
  if (pfix_dev && pfix_stride != 0 && pfix_stride != per_path)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_time_dependent: stride must be 0 or N_model * NP");
  ctx->ptime = 1;
  ctx->pfix_dev = pfix_dev ? pfix_dev : ctx->pfix_zero;   // (never read when every parameter is estimated)
  ctx->pfix_stride = pfix_dev ? pfix_stride : 0;
  return VAB_OK;
}

}  // extern "C"

static int ode_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
                    const double* rf_path_dev, const int* active_dev, double* A, double* me,
                    double* fe, double* G, long long ldg) {
  const vab_ode_desc& d = ctx->od;
  const long long n = ctx->n_unknowns();
  if (B < 1 || !XP) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: bad batch / XP");
  if (ldxp < n || (ldxp & 1)) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: ldxp must be even and >= n");
  if (G && (ldg < n || (ldg & 1))) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: ldg must be even and >= n");
  if (((uintptr_t)XP & 15) || ((uintptr_t)G & 15))
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: XP/G must be 16-byte aligned");
  OdeParams P;
  memset(&P, 0, sizeof(P));
  P.XP = XP; P.ldxp = ldxp; P.G = G; P.ldg = ldg;
  P.B = B; P.D = d.D; P.N = d.N_model; P.N_data = d.N_data; P.nskip = d.nskip; P.L = d.L;
  P.dt = d.dt_model;
  P.Y = ctx->Y_dense; P.wobs = ctx->wobs_dev; P.rmd = ctx->rm_dev;
  P.rf_scalar = ctx->rf0_scalar * rf_scale; P.rf_arr = ctx->rf0_dev; P.rf_scale = rf_scale;
  P.rf0 = ctx->rf0_scalar; P.rf_path = rf_path_dev;
  P.rf_mat = ctx->rf0_mat;
  P.stim = ctx->stim_dev; P.S = d.n_stim;
  P.NP = d.NP; P.NPest = d.NPest; P.pmap = ctx->pmap_dev;
  P.pfix = ctx->pfix_dev; P.pfix_stride = ctx->pfix_stride;
  P.ptime = ctx->ptime;
  P.K = 2 + d.NP;
  P.active = active_dev;
  P.cm = d.L > 0 ? 1.0 / ((double)d.L * d.N_data) : 0.0;
  P.cf = 1.0 / ((double)d.D * (d.N_model - 1));
  cudaError_t cerr = cudaSuccess;
  int rc;
  {
    SweepLaunch sl;
    rc = ode_sweep_prepare(d.model, d.disc, d.D, d.N_model, B, ctx->num_sms, ctx->tseg_override,
                           !ctx->use_sweep, &P, &sl, &cerr);
    if (rc == -1) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: unsupported shape");
    if (rc == -3) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: parameter time series: unsupported discretisation / row width");
    if (rc != 0) return vab_cuda_fail(ctx, cerr, "ode_action_grad occupancy query");
    rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)P.nunits * P.K);
    if (rc != VAB_OK) return rc;
    P.partials = ctx->partials;
    // stand-alone evaluations (no mask, no per-path RF: vab_ode_action_grad) use dependent launch;
    // inside the minimiser's captured cycles the kernels are launched the ordinary way
    static int pdl_on = -1;
    if (pdl_on < 0) {
      const char* e = getenv("VAB_PDL");
      pdl_on = e ? (atoi(e) != 0) : 1;
    }
    const bool pdl = pdl_on && active_dev == nullptr && rf_path_dev == nullptr;
    rc = ode_sweep_launch(P, sl, d.model, d.disc, ctx->stream, A, me, fe, &cerr, pdl);
  }
  if (rc == -1) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: no kernel for this model/disc");
  if (rc != 0) return vab_cuda_fail(ctx, cerr, "ode_action_grad launch");
  ctx->launches += 2;
  if (ctx->rm_matrix != nullptr) {
    ode_me_matrix_kernel<<<B, 256, 0, ctx->stream>>>(XP, ldxp, G, ldg, ctx->Y_dense, ctx->lcomp_dev, ctx->rm_matrix,
                                                     d.N_data, d.L, d.D, d.nskip, P.cm, active_dev, A, me);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ode_me_matrix_kernel");
    ctx->launches += 1;
  }
  return VAB_OK;
}

int vab_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
             const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
             double* G, long long ldg) {
  if (ctx->problem == VAB_PROBLEM_ODE)
    return ode_eval(ctx, B, XP, ldxp, rf_scale, rf_path_dev, active_dev, A, me, fe, G, ldg);
  if (ctx->problem == VAB_PROBLEM_NN)
    return nn_eval(ctx, B, XP, ldxp, rf_scale, rf_path_dev, active_dev, A, me, fe, G, ldg);
  return vab_fail(ctx, VAB_ERR_STATE, "no problem set on this context");
}

extern "C" int vab_ode_action_grad(vab_ctx* ctx, int32_t B, const double* XP_dev, int64_t ldxp,
                                   double rf_scale, double* A_dev, double* me_dev, double* fe_dev,
                                   double* G_dev, int64_t ldg) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "ode_action_grad: no ODE problem set");
  cudaSetDevice(ctx->device);
  return ode_eval(ctx, B, XP_dev, ldxp, rf_scale, nullptr, nullptr, A_dev, me_dev, fe_dev, G_dev, ldg);
}
