// Measured denominators for the roofline of the fp64-bound kernels (rk4, NaKL, the neural-network
// contractions): the device's double-precision FMA throughput on the CUDA cores, timed with CUDA
// events.  MEASURED_PEAKS.json (driver-written) holds the HBM copy bandwidth and the bf16 tensor
// rate but no fp64 figure, so bench.py measures this one itself and says so.
#include <cuda_runtime.h>

#include "vab_ctx.h"

namespace {

// 16 independent FMA chains per thread: enough instruction-level parallelism to cover the DFMA
// latency with 8 warps per scheduler
__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double a, double b) {
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = fma(v[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the loop alive
}

// the same for the fp64 tensor pipe: 8 independent DMMA.8x8x4 accumulator chains per warp (the
// instruction the neural-network kernels issue, nn_action.cu)
__global__ void __launch_bounds__(256) fp64_dmma_kernel(double* out, int iters, double a, double b) {
  double c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) c[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[2 * k]), "+d"(c[2 * k + 1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += c[k];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// which: 0 = DFMA on the CUDA cores, 1 = DMMA.8x8x4 on the tensor pipe
static int measure_peak(vab_ctx* ctx, int which, double* tflops_host);

extern "C" int vab_measure_fp64_dmma_peak(vab_ctx* ctx, double* tflops_host) { return measure_peak(ctx, 1, tflops_host); }

extern "C" int vab_measure_fp64_peak(vab_ctx* ctx, double* tflops_host) { return measure_peak(ctx, 0, tflops_host); }

static int measure_peak(vab_ctx* ctx, int which, double* tflops_host) {
  if (!ctx || !tflops_host) return VAB_ERR_INVALID;
  cudaSetDevice(ctx->device);
  double* out = nullptr;
  const int blocks = ctx->num_sms * 8, threads = 256, iters = 4096;
  cudaError_t e = cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double));
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "measure_fp64_peak");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    if (which == 1) fp64_dmma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 0.999999, 1e-9);
    else fp64_fma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, ctx->stream);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // DFMA: 16 FMAs per thread and iteration; DMMA: 8 instructions per warp and iteration, 8x8x4 MACs each
    const double flops = which == 1 ? 2.0 * 8.0 * 256.0 * iters * (double)blocks * (threads / 32)
                                    : 2.0 * 16.0 * iters * (double)blocks * threads;
    if (rep > 0 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);   // first launch = warm-up
  }
  ctx->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "measure_fp64_peak");
  *tflops_host = best;
  return VAB_OK;
}
