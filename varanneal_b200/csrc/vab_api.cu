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
int nn_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
            const int* active_dev, double* A, double* me, double* fe, double* G, long long ldg);
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

// dst (rows, Lp) <- src (rows, L) with column l going to perm[l] (padding zeroed beforehand)
__global__ void vab_pack_obs_scatter(const double* __restrict__ src, double* __restrict__ dst,
                                     const int* __restrict__ perm, long long rows, int L, int Lp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const long long r = i / L;
  const int l = (int)(i - r * L);
  dst[r * Lp + perm[l]] = src[i];
}

static int vab_pack_obs(vab_ctx* ctx, const double* src, double** dst, size_t* cap) {
  const vab_ode_desc& d = ctx->od;
  const size_t need = (size_t)d.N_data * ctx->Lp + ctx->Lw + 2;
  int rc = vab_reserve(ctx, dst, cap, need);
  if (rc != VAB_OK) return rc;
  cudaError_t e = cudaMemsetAsync(*dst, 0, need * sizeof(double), ctx->stream);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "pack observations (memset)");
  const long long n = (long long)d.N_data * d.L;
  if (n > 0)
    vab_pack_obs_scatter<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, *dst, ctx->lperm_dev, d.N_data, d.L, ctx->Lp);
  e = cudaGetLastError();
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "pack observations");
  ctx->launches += 1;
  return VAB_OK;
}

long long vab_ctx::n_unknowns() const {
  if (problem == VAB_PROBLEM_ODE) return (long long)od.N_model * od.D + od.NPest;
  if (problem == VAB_PROBLEM_NN) return nn_unknowns(this);
  return 0;
}

extern "C" {

int vab_abi_version(void) { return VAB_ABI_VERSION; }

const char* vab_last_error(const vab_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

long long vab_launch_count(const vab_ctx* ctx) { return ctx ? ctx->launches : 0; }

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
  if (const char* t = getenv("VAB_KERNEL")) {
    c->use_walk = (strcmp(t, "walk") == 0);
    c->use_sweep = (strcmp(t, "sweep") == 0);
  }
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
  cudaFree(ctx->obs_slot_dev);
  cudaFree(ctx->pmap_dev);
  cudaFree(ctx->Y_pad);
  cudaFree(ctx->rm_pad);
  cudaFree(ctx->lperm_dev);
  cudaFree(ctx->win_y0_dev);
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
  if (ctx->use_walk) {
    OdePlan pl;
    int prc = ode_make_plan(d->model, d->disc, d->D, d->N_model, 1, ctx->num_sms, 0, &pl);
    if (prc != 0)
      return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: D not supported by the legacy walk kernels");
  }
  if ((d->model == VAB_MODEL_LORENZ96 && d->D < 4) || (d->model == VAB_MODEL_LORENZ63 && d->D != 3) ||
      (d->model == VAB_MODEL_NAKL && d->D != 4))
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: D not valid for this model");

  std::vector<int> obs(d->D, -1), pmap(d->NP, -1), lperm(d->L > 0 ? d->L : 1, 0);
  for (int l = 0; l < d->L; ++l) {
    const int i = Lidx_host[l];
    if (i < 0 || i >= d->D) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: Lidx out of range");
    if (obs[i] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: duplicate Lidx entry");
    obs[i] = l;
  }
  // library layout of Y / RM: columns sorted by state component (rank), row pitch Lp (even)
  std::vector<int> sorted_comp;
  for (int i = 0, rank = 0; i < d->D; ++i)
    if (obs[i] >= 0) {
      lperm[obs[i]] = rank;
      obs[i] = rank++;
      sorted_comp.push_back(i);
    }
  for (int e = 0; e < d->NPest; ++e) {
    const int k = Pidx_host[e];
    if (k < 0 || k >= d->NP) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: Pidx out of range");
    if (pmap[k] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: duplicate Pidx entry");
    pmap[k] = e;
  }
  // per-window Y column ranges of the stream kernels
  OdeGeo geo;
  if (ode_geometry(d->model, d->disc, d->D, &geo) != 0)
    return vab_fail(ctx, VAB_ERR_INVALID, "ode_problem_set: D not valid for this model");
  const int Lp = (d->L + 1) & ~1;
  std::vector<int> wy0(geo.nwin, 0);
  int Lw = 2;
  for (int w = 0; w < geo.nwin; ++w) {
    const int c_lo = w * geo.WS * geo.C;
    int c_hi = (w + 1) * geo.WS * geo.C;
    if (c_hi > d->D) c_hi = d->D;
    int s0 = 0, s1 = 0;
    for (int c : sorted_comp) { if (c < c_lo) ++s0; if (c < c_hi) ++s1; }
    const int y0 = s0 & ~1;
    int len = (s1 - y0 + 1) & ~1;
    if (geo.nwin == 1) len = Lp;
    wy0[w] = y0;
    if (len > Lw) Lw = len;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->obs_slot_dev);
  cudaFree(ctx->pmap_dev);
  cudaFree(ctx->lperm_dev);
  cudaFree(ctx->win_y0_dev);
  ctx->obs_slot_dev = nullptr;
  ctx->pmap_dev = nullptr;
  ctx->lperm_dev = nullptr;
  ctx->win_y0_dev = nullptr;
  cudaError_t e = cudaMalloc((void**)&ctx->obs_slot_dev, sizeof(int) * d->D);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->pmap_dev, sizeof(int) * (d->NP > 0 ? d->NP : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->lperm_dev, sizeof(int) * lperm.size());
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->win_y0_dev, sizeof(int) * wy0.size());
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->obs_slot_dev, obs.data(), sizeof(int) * d->D, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && d->NP > 0)
    e = cudaMemcpy(ctx->pmap_dev, pmap.data(), sizeof(int) * d->NP, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->lperm_dev, lperm.data(), sizeof(int) * lperm.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->win_y0_dev, wy0.data(), sizeof(int) * wy0.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ode_problem_set");
  ctx->od = *d;
  ctx->Lp = Lp;
  ctx->Lw = Lw;
  {
    int rc = vab_pack_obs(ctx, Y_dev, &ctx->Y_pad, &ctx->Y_cap);
    if (rc != VAB_OK) return rc;
  }
  ctx->Y_dev = ctx->Y_pad;
  ctx->stim_dev = (d->n_stim > 0) ? stim_dev : nullptr;
  ctx->pfix_dev = ctx->pfix_zero;
  ctx->pfix_stride = 0;
  ctx->rm_scalar = 1.0; ctx->rm_dev = nullptr;
  ctx->rf0_scalar = 1.0; ctx->rf0_dev = nullptr;
  ctx->problem = VAB_PROBLEM_ODE;
  return VAB_OK;
}

int vab_ode_set_weights(vab_ctx* ctx, double rm_scalar, const double* rm_dev, double rf0_scalar,
                        const double* rf0_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_weights: no ODE problem set");
  ctx->rm_scalar = rm_scalar; ctx->rm_dev = nullptr;
  if (rm_dev) {
    cudaSetDevice(ctx->device);
    int rc = vab_pack_obs(ctx, rm_dev, &ctx->rm_pad, &ctx->rm_cap);
    if (rc != VAB_OK) return rc;
    ctx->rm_dev = ctx->rm_pad;
  }
  ctx->rf0_scalar = rf0_scalar; ctx->rf0_dev = rf0_dev;
  return VAB_OK;
}

int vab_ode_set_fixed_params(vab_ctx* ctx, const double* pfix_dev, int64_t pfix_stride) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "set_fixed_params: no ODE problem set");
  if (!pfix_dev) return vab_fail(ctx, VAB_ERR_INVALID, "set_fixed_params: NULL");
  if (pfix_stride != 0 && pfix_stride != ctx->od.NP)
    return vab_fail(ctx, VAB_ERR_INVALID, "set_fixed_params: stride must be 0 or NP");
  ctx->pfix_dev = pfix_dev;
  ctx->pfix_stride = pfix_stride;
  return VAB_OK;
}

}  // extern "C"

static int ode_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
                    const int* active_dev, double* A, double* me, double* fe, double* G,
                    long long ldg) {
  const vab_ode_desc& d = ctx->od;
  const long long n = (long long)d.N_model * d.D + d.NPest;
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
  P.obs_slot = ctx->obs_slot_dev; P.Y = ctx->Y_dev; P.Lp = ctx->Lp; P.Lw = ctx->Lw;
  P.win_y0 = ctx->win_y0_dev;
  P.rm_scalar = ctx->rm_scalar; P.rm_arr = ctx->rm_dev;
  P.rf_scalar = ctx->rf0_scalar * rf_scale; P.rf_arr = ctx->rf0_dev; P.rf_scale = rf_scale;
  P.stim = ctx->stim_dev; P.S = d.n_stim;
  P.NP = d.NP; P.NPest = d.NPest; P.pmap = ctx->pmap_dev;
  P.pfix = ctx->pfix_dev; P.pfix_stride = ctx->pfix_stride;
  P.K = 2 + d.NP;
  P.active = active_dev;
  P.cm = d.L > 0 ? 1.0 / ((double)d.L * d.N_data) : 0.0;
  P.cf = 1.0 / ((double)d.D * (d.N_model - 1));
  cudaError_t cerr = cudaSuccess;
  int rc;
  if (ctx->use_walk) {
    OdePlan pl;
    if (ode_make_plan(d.model, d.disc, d.D, d.N_model, B, ctx->num_sms, ctx->tseg_override, &pl) != 0)
      return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: unsupported shape");
    P.Tseg = pl.Tseg; P.nseg = pl.nseg; P.TPR = pl.TPR; P.RG = pl.RG; P.nunits = pl.nunits;
    P.upp = pl.nseg;
    rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)pl.nunits * P.K);
    if (rc != VAB_OK) return rc;
    P.partials = ctx->partials;
    rc = ode_launch_action(P, pl, d.model, d.disc, ctx->stream, A, me, fe, &cerr);
  } else {
    SweepLaunch sl;
    rc = ode_sweep_prepare(d.model, d.disc, d.D, d.N_model, B, ctx->num_sms, ctx->tseg_override,
                           !ctx->use_sweep, &P, &sl, &cerr);
    if (rc == -1) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: unsupported shape");
    if (rc != 0) return vab_cuda_fail(ctx, cerr, "ode_action_grad occupancy query");
    rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)P.nunits * P.K);
    if (rc != VAB_OK) return rc;
    P.partials = ctx->partials;
    rc = ode_sweep_launch(P, sl, d.model, d.disc, ctx->stream, A, me, fe, &cerr);
  }
  if (rc == -1) return vab_fail(ctx, VAB_ERR_INVALID, "ode_action_grad: no kernel for this model/disc");
  if (rc != 0) return vab_cuda_fail(ctx, cerr, "ode_action_grad launch");
  ctx->launches += 2;
  return VAB_OK;
}

int vab_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
             const int* active_dev, double* A, double* me, double* fe, double* G, long long ldg) {
  if (ctx->problem == VAB_PROBLEM_ODE)
    return ode_eval(ctx, B, XP, ldxp, rf_scale, active_dev, A, me, fe, G, ldg);
  if (ctx->problem == VAB_PROBLEM_NN)
    return nn_eval(ctx, B, XP, ldxp, rf_scale, active_dev, A, me, fe, G, ldg);
  return vab_fail(ctx, VAB_ERR_STATE, "no problem set on this context");
}

extern "C" int vab_ode_action_grad(vab_ctx* ctx, int32_t B, const double* XP_dev, int64_t ldxp,
                                   double rf_scale, double* A_dev, double* me_dev, double* fe_dev,
                                   double* G_dev, int64_t ldg) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_ODE) return vab_fail(ctx, VAB_ERR_STATE, "ode_action_grad: no ODE problem set");
  cudaSetDevice(ctx->device);
  return ode_eval(ctx, B, XP_dev, ldxp, rf_scale, nullptr, A_dev, me_dev, fe_dev, G_dev, ldg);
}
