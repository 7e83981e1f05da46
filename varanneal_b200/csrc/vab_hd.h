// Shared host/device helpers.  Everything numeric lives in VAB_HD functions so that the same
// code can be compiled by g++ into the test-only kernel emulator (tests/emul/) and by nvcc into
// the product kernels.  The emulator exists to check index arithmetic without a GPU; it is never
// loaded by the product.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VAB_HD __host__ __device__ __forceinline__
#else
#define VAB_HD inline
#endif

// read-only global load (ld.global.nc on the device)
VAB_HD double vab_ldg(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
VAB_HD int vab_ldg(const int* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// C consecutive doubles; 16-byte vector accesses when C is even (the host guarantees alignment:
// ld even, D multiple of C).
template <int C>
VAB_HD void vab_load_strip(const double* src, double* dst) {
#if defined(__CUDA_ARCH__)
  if constexpr (C % 2 == 0) {
#pragma unroll
    for (int j = 0; j < C; j += 2) {
      double2 t = __ldg(reinterpret_cast<const double2*>(src + j));
      dst[j] = t.x;
      dst[j + 1] = t.y;
    }
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < C; ++j) dst[j] = vab_ldg(src + j);
}

template <int C>
VAB_HD void vab_store_strip(double* dst, const double* src) {
#if defined(__CUDA_ARCH__)
  if constexpr (C % 2 == 0) {
#pragma unroll
    for (int j = 0; j < C; j += 2)
      __stcs(reinterpret_cast<double2*>(dst + j), make_double2(src[j], src[j + 1]));
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < C; ++j) dst[j] = src[j];
}

// Programmatic dependent launch (PDL).  Kernels launched back to back with the
// programmatic-stream-serialization attribute may start before their predecessor has finished:
// vab_pdl_trigger() lets the next kernel in the stream begin launching, vab_pdl_wait() blocks
// until the previous kernel has completed and its writes are visible.  Both are no-ops for a
// kernel launched the ordinary way.
#if defined(__CUDACC__)
__device__ __forceinline__ void vab_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void vab_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
