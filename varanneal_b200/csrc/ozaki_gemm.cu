// fp64 contractions on the 5th-generation tensor cores: Ozaki-split GEMM over tcgen05 / TMEM / TMA.
//
// The neural-network action (va_nnet.py:175-255) is three dense contractions per layer over the
// example batch; BASELINE.json's north star asks for them on tcgen05 with TMEM accumulators fed by
// TMA.  tcgen05.mma has no f64 kind, and parity is 1e-10 in fp64, so the operands are split
// (Ozaki scheme): every row of A and of B is scaled by a power of two to (-1, 1) and cut into
// NS = 7 signed 7-bit digits q_s (x = 2^e sum_s q_s 2^(-6 - 7 s), exact in fp64 arithmetic).  The
// product of two digits planes is an int8 x int8 -> int32 GEMM, *exact* on the tensor cores
// (kind::i8); planes with the same order t = i + j accumulate into the same TMEM accumulator
// (|sum| <= 7 * 128 * 64^2 < 2^22), the orders t <= 6 are kept (28 plane pairs) and recombined in
// fp64 in the epilogue: C = 2^(ea + eb - 12) sum_t 2^(-7 t) C_t.  Dropped pairs (i + j >= 7) and the
// digits beyond the 7th bound the error by ~(6 K + 2) 2^-49 of (row max of A) x (row max of B):
// 3e-13 for K = 100 -- measured by vab_ozaki_gemm_probe against an fp64 FMA reference.
//
// Kernels
//   ozaki_slice_kernel   rows of an fp64 matrix -> 7 int8 digit planes, K padded to 128 (one
//                        128-byte swizzle row), + the row exponent
//   ozaki_mma_kernel     one CTA per (problem, 128-row tile, 64-column tile): one thread issues 14
//                        TMA tensor-map loads (7 planes of A: 128 x 128 B, 7 of B: 64 x 128 B,
//                        SWIZZLE_128B) completing on an mbarrier; one thread issues the 112
//                        tcgen05.mma.kind::i8 (M 128, N 64, K 32) into 7 accumulators = 448 TMEM
//                        columns and commits to a second mbarrier; four warps read their 32 TMEM
//                        lanes back with tcgen05.ld, recombine in fp64 and store C.
// Status: measured prototype of the forward contraction (C4 shape: 256 problems 1000 x 100 x 100)
// behind the C ABI probe below; the product's NN kernels (nn_action.cu) stay on the fp64 tensor
// pipe (DMMA) until the backward contractions and the sigmoid epilogue are fused here too --
// DESIGN.md section 4.4 has the numbers and the decision.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vab_ctx.h"

namespace {

constexpr int NS = 7;                 // digit planes per operand
constexpr int KP = 128;               // padded K (bytes per plane row) = one swizzle-128B row
constexpr int TM = 128, TN = 64;      // output tile
constexpr int TMEM_COLS = 512;        // 7 accumulators x 64 columns, rounded up to a power of two
constexpr uint32_t A_PLANE = TM * KP, B_PLANE = TN * KP;
constexpr uint32_t SMEM_BYTES = NS * (A_PLANE + B_PLANE) + 1024 /*alignment*/ + 64;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- digit planes ---------------------------------------------------------------------------
// one warp per row: row max -> exponent, then the digits of every element
// (element k of row r of problem p = src[p pstride + r ld + k kstride]: kstride != 1 slices a transposed operand)
__global__ void __launch_bounds__(256) ozaki_slice_kernel(const double* __restrict__ src, long long ld, long long pstride,
                                                          int rows, int K, int rows_pad, int8_t* __restrict__ planes,
                                                          long long plane_stride, int* __restrict__ expo,
                                                          long long kstride = 1, const int* __restrict__ active = nullptr,
                                                          int pdiv = 1, long long cstride = 0, int ktotal = 0) {
  // problem p = (path p / pdiv, chunk p % pdiv of the contraction axis): chunk c starts cstride doubles
  // further and holds min(K, ktotal - c K) elements (the split-K form of a contraction longer than 128)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.y;
  if (warp >= rows_pad) return;
  const int po = p / pdiv, pc = p - po * pdiv;
  if (active != nullptr && active[po] == 0) return;
  if (ktotal > 0) K = min(K, ktotal - pc * K);
  const long long orow = (long long)p * rows_pad + warp;
  int8_t* out = planes + orow * KP;
  if (warp >= rows) {                                  // padding rows: zeros
    for (int s = 0; s < NS; ++s)
      for (int k = lane; k < KP; k += 32) out[(long long)s * plane_stride + k] = 0;
    if (lane == 0) expo[orow] = 0;
    return;
  }
  const double* x = src + (long long)po * pstride + (long long)pc * cstride + (long long)warp * ld;
  double amax = 0.0;
  for (int k = lane; k < K; k += 32) amax = fmax(amax, fabs(x[(long long)k * kstride]));
  for (int sft = 16; sft > 0; sft >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, sft));
  int e = 0;
  if (amax > 0.0) { (void)frexp(amax, &e); }           // amax = m 2^e, m in [0.5, 1)  ->  |x| / 2^e < 1
  if (lane == 0) expo[orow] = e;
  // lane l owns bytes 4 l .. 4 l + 3 of the 128-byte plane row: one 4-byte store per plane, 128 B per warp
  {
    const int k0 = 4 * lane;
    // 2^(6 - e) assembled from its bits (exact scaling; e is within the normal range for finite data)
    const double sc = __longlong_as_double((long long)(6 - e + 1023) << 52);
    double t[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) t[c] = (k0 + c < K) ? x[(long long)(k0 + c) * kstride] * sc : 0.0;     // |t| < 64
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      unsigned pack = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double qd = rint(t[c]);
        pack |= ((unsigned)(int)qd & 0xFFu) << (8 * c);
        t[c] = (t[c] - qd) * 128.0;                    // exact: the remainder has fewer significant bits
      }
      *reinterpret_cast<unsigned*>(out + (long long)s * plane_stride + k0) = pack;
    }
  }
}

// ---- the tcgen05 kernel ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptor, K-major operand in the SWIZZLE_128B layout (cute::UMMA::SmemDescriptor:
// start address >> 4 in bits [0,14), leading byte offset (1: unused for swizzled K-major) in [16,30),
// stride byte offset = 8 rows x 128 B = 1024 >> 4 in [32,46), version 1 in [46,48), layout type 2 in [61,64))
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) bits [4,6), a / b format INT8 (1)
// bits [7,10) / [10,13), K-major A and B, n_dim = N >> 3 bits [17,23), m_dim = M >> 4 bits [24,29)
constexpr uint32_t IDESC_S8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC_S8), "r"(accumulate) : "memory");
}

struct OzakiParams {
  int M, N, P;               // rows of A / of B per problem, problems
  int m_tiles, n_tiles;      // per problem
  int Mpad, Npad;
  const int* ea;             // (P * Mpad) row exponents of A
  const int* eb;             // (P * Npad)
  double* C;                 // element (p, r, c) at C[p cps + r ldc + c]
  long long ldc, cps;
  const int* active;         // (P / pdiv) or nullptr: paths to skip
  int pdiv;                  // problems per path (chunks of a split contraction)
};

__global__ void __launch_bounds__(128, 1) ozaki_mma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB,
                                                           const OzakiParams q) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (s32(smem_raw) + 1023u) & ~1023u;           // swizzle-128B tiles want 1024-byte alignment
  const uint32_t sA = sbase, sB = sbase + NS * A_PLANE;
  const uint32_t bar_tma = s32(&bars[0]), bar_mma = s32(&bars[1]);
  int tile = blockIdx.x;
  const int nt = tile % q.n_tiles; tile /= q.n_tiles;
  const int mt = tile % q.m_tiles;
  const int p = tile / q.m_tiles;
  if (q.active != nullptr && q.active[p / q.pdiv] == 0) return;

  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_sh)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_sh;

  if (tid == 0) {                                   // ---- TMA producer: all 14 planes of this tile
    mbar_expect_tx(bar_tma, NS * (A_PLANE + B_PLANE));
    const int rowA = p * q.Mpad + mt * TM, rowB = p * q.Npad + nt * TN;
    for (int s = 0; s < NS; ++s) {
      tma_load_3d(sA + s * A_PLANE, &mapA, bar_tma, 0, rowA, s);
      tma_load_3d(sB + s * B_PLANE, &mapB, bar_tma, 0, rowB, s);
    }
  }
  if (tid == 32) {                                  // ---- MMA issuer (one thread)
    mbar_wait(bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int t = 0; t < NS; ++t) {                  // order t = i + j -> accumulator t
      uint32_t acc = 0;
      for (int i = 0; i <= t; ++i) {
        const int j = t - i;
#pragma unroll
        for (int k = 0; k < KP / 32; ++k) {         // UMMA_K = 32 bytes: +2 sixteen-byte units per step
          const uint64_t da = umma_desc_sw128(sA + i * A_PLANE) + (uint64_t)(2 * k);
          const uint64_t db = umma_desc_sw128(sB + j * B_PLANE) + (uint64_t)(2 * k);
          umma_s8(tmem + (uint32_t)(t * TN), da, db, acc);
          acc = 1;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
  }
  __syncwarp();

  // ---- epilogue: warp w owns TMEM lanes (= tile rows) 32 w .. 32 w + 31
  mbar_wait(bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = mt * TM + warp * 32 + lane;
  const int ea = q.ea[(long long)p * q.Mpad + row];
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < TN; c0 += 16) {
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0;
#pragma unroll
    for (int t = NS - 1; t >= 0; --t) {             // Horner in 2^-7: acc = C_t + acc / 128 (small terms first)
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(lane_addr + (uint32_t)(t * TN + c0)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fma(acc[c], 1.0 / 128.0, (double)(int)v[c]);
    }
    if (row < q.M) {
      double* out = q.C + (long long)p * q.cps + (long long)row * q.ldc;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int col = nt * TN + c0 + c;
        if (col < q.N) out[col] = ldexp(acc[c], ea + q.eb[(long long)p * q.Npad + col] - 12);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}


// ---- version 2: persistent, warp-specialised, pipelined ---------------------------------------
// 148 CTAs walk the tiles (128 x 32 outputs).  Warp 0 = TMA producer, warp 1 = MMA issuer, warps
// 2-5 = epilogue.  Operand ring: 3 stages of one K-half each (7 planes of A: 128 rows x 64 B, 7 of B:
// 32 rows x 64 B, SWIZZLE_64B; 70 KB per stage) with full / empty mbarriers; two accumulator sets in
// TMEM (2 x 7 x 32 = 448 columns) with tmem_full / tmem_empty mbarriers, so the TMA loads of the
// next K-half / tile, the MMAs of this one and the fp64 recombination of the previous tile overlap.
constexpr int TN2 = 32, KH = 64, NSTG = 3;
constexpr uint32_t A_PLANE2 = TM * KH, B_PLANE2 = TN2 * KH;
constexpr uint32_t STAGE2 = NS * (A_PLANE2 + B_PLANE2);                 // 71 680 B
constexpr uint32_t SMEM2_BYTES = NSTG * STAGE2 + 1024 + 64;
constexpr uint32_t IDESC_S8_N32 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN2 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// K-major operand in the SWIZZLE_64B layout: stride byte offset = 8 rows x 64 B = 512 >> 4, layout type 4
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void umma_s8_n32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC_S8_N32), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(192, 1) ozaki_mma2_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const OzakiParams q) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * NSTG + 4];
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (s32(smem_raw) + 1023u) & ~1023u;
  auto full = [&](int s) { return s32(&bars[s]); };
  auto empty = [&](int s) { return s32(&bars[NSTG + s]); };
  auto tfull = [&](int a) { return s32(&bars[2 * NSTG + a]); };
  auto tempty = [&](int a) { return s32(&bars[2 * NSTG + 2 + a]); };
  const int ntiles = q.P * q.m_tiles * q.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }   // 4 epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_sh)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {                                 // ---- TMA producer
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int t = tile;
        const int nt = t % q.n_tiles; t /= q.n_tiles;
        const int mt = t % q.m_tiles;
        const int p = t / q.m_tiles;
        const int rowA = p * q.Mpad + mt * TM, rowB = p * q.Npad + nt * TN2;
        for (int h = 0; h < 2; ++h, ++it) {
          const int s = it % NSTG;
          if (it >= NSTG) mbar_wait(empty(s), (uint32_t)(((it / NSTG) - 1) & 1));
          mbar_expect_tx(full(s), STAGE2);
          const uint32_t base = sbase + s * STAGE2;
          for (int pl = 0; pl < NS; ++pl) {
            tma_load_3d(base + pl * A_PLANE2, &mapA, full(s), h * KH, rowA, pl);
            tma_load_3d(base + NS * A_PLANE2 + pl * B_PLANE2, &mapB, full(s), h * KH, rowB, pl);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                 // ---- MMA issuer
      int it = 0, nt_done = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt_done) {
        const int a = nt_done & 1;
        if (nt_done >= 2) mbar_wait(tempty(a), (uint32_t)(((nt_done >> 1) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc0 = tmem + (uint32_t)(a * NS * TN2);
        for (int h = 0; h < 2; ++h, ++it) {
          const int s = it % NSTG;
          mbar_wait(full(s), (uint32_t)((it / NSTG) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = sbase + s * STAGE2;
          for (int t = 0; t < NS; ++t) {
            for (int i = 0; i <= t; ++i) {
              const int j = t - i;
#pragma unroll
              for (int k = 0; k < KH / 32; ++k) {
                const uint64_t da = umma_desc_sw64(base + i * A_PLANE2) + (uint64_t)(2 * k);
                const uint64_t db = umma_desc_sw64(base + NS * A_PLANE2 + j * B_PLANE2) + (uint64_t)(2 * k);
                umma_s8_n32(acc0 + (uint32_t)(t * TN2), da, db, (h | i | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(empty(s));                     // the stage may be refilled once these MMAs have read it
        }
        umma_commit(tfull(a));                       // accumulators of this tile complete
      }
    }
  } else {                                           // ---- epilogue warps 2..5: TMEM lanes 32 (warp % 4) ..
    const int lg = warp & 3;
    int nt_done = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt_done) {
      int t = tile;
      const int nt = t % q.n_tiles; t /= q.n_tiles;
      const int mt = t % q.m_tiles;
      const int p = t / q.m_tiles;
      const int a = nt_done & 1;
      mbar_wait(tfull(a), (uint32_t)((nt_done >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = mt * TM + lg * 32 + lane;
      const int ea = q.ea[(long long)p * q.Mpad + row];
      const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * NS * TN2);
      double acc[TN2];
#pragma unroll
      for (int c = 0; c < TN2; ++c) acc[c] = 0.0;
#pragma unroll
      for (int tt = NS - 1; tt >= 0; --tt) {
        uint32_t v[TN2];
#pragma unroll
        for (int c0 = 0; c0 < TN2; c0 += 16) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
              : "r"(lane_addr + (uint32_t)(tt * TN2 + c0)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < TN2; ++c) acc[c] = fma(acc[c], 1.0 / 128.0, (double)(int)v[c]);
      }
      // the accumulators are in registers: hand the TMEM set back before the global stores
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(a));
      if (row < q.M) {
        double* out = q.C + (long long)p * q.cps + (long long)row * q.ldc;
        const int* ebp = q.eb + (long long)p * q.Npad + nt * TN2;
#pragma unroll
        for (int c = 0; c < TN2; ++c) {
          const int col = nt * TN2 + c;
          if (col < q.N) {
            // 2^(ea + eb - 12) assembled from its bits: exact scaling, no ldexp call
            const long long ex = (long long)(ea + ebp[c] - 12 + 1023);
            const double sc = __longlong_as_double(ex << 52);
            out[col] = acc[c] * sc;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ---- reference + data for the probe ---------------------------------------------------------
__global__ void ozaki_fill_kernel(double* x, long long n, unsigned long long seed, double spread) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  const double mag = exp2(spread * ((double)((z >> 3) & 1023) / 1023.0 - 0.5));    // wide dynamic range inside a row
  x[i] = u * mag;
}
// C = A B^T in fp64 FMAs (one thread per output element; the yardstick for the error, not for speed)
__global__ void ozaki_ref_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                 int M, int N, int K) {
  const int p = blockIdx.z;
  const int r = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M || c >= N) return;
  const double* a = A + ((long long)p * M + r) * K;
  const double* b = B + ((long long)p * N + c) * K;
  double s = 0.0;
  for (int k = 0; k < K; ++k) s = fma(a[k], b[k], s);
  C[((long long)p * M + r) * N + c] = s;
}
__global__ void ozaki_err_kernel(const double* __restrict__ C, const double* __restrict__ Cref, long long n, double* out) {
  __shared__ double se[256], sm[256];
  double e = 0.0, m = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    e = fmax(e, fabs(C[i] - Cref[i]));
    m = fmax(m, fabs(Cref[i]));
  }
  se[threadIdx.x] = e; sm[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { se[threadIdx.x] = fmax(se[threadIdx.x], se[threadIdx.x + s]); sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = se[0]; out[2 * blockIdx.x + 1] = sm[0]; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(vab_ctx* ctx, EncodeTiledFn enc, CUtensorMap* map, void* base, long long rows_total, int box_rows,
                   int box_k = KP, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  const cuuint64_t dims[3] = {(cuuint64_t)KP, (cuuint64_t)rows_total, (cuuint64_t)NS};
  const cuuint64_t strides[2] = {(cuuint64_t)KP, (cuuint64_t)rows_total * KP};          // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return vab_fail(ctx, VAB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return VAB_OK;
}

#define OZ_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) { rc = vab_cuda_fail(ctx, e_, #call); goto done; } \
  } while (0)

}  // namespace

// ---- product entry point (nn_action.cu, VAB_NN_TCGEN05=1) -----------------------------------------
// C[p][r][c] = sum_k A[p][r][k] B[p][c][k] for P problems with K <= 128, operands addressed by
// (problem stride, row stride, element stride), the output by (problem stride, row pitch):
// digit planes of both operands (ozaki_slice_kernel), then one tcgen05 tile per CTA
// (ozaki_mma_kernel).  The plane / exponent workspaces belong to the context and only grow.
struct OzakiWork {
  int8_t *pa = nullptr, *pb = nullptr;
  int *ea = nullptr, *eb = nullptr;
  size_t pa_cap = 0, pb_cap = 0, ea_cap = 0, eb_cap = 0;      // bytes / ints
  EncodeTiledFn enc = nullptr;
  bool attr = false;
};

void ozaki_destroy(vab_ctx* ctx) {
  OzakiWork* w = ctx->oz;
  if (!w) return;
  cudaFree(w->pa); cudaFree(w->pb); cudaFree(w->ea); cudaFree(w->eb);
  delete w;
  ctx->oz = nullptr;
}

template <typename T>
static int oz_reserve(vab_ctx* ctx, T** buf, size_t* cap, size_t need) {
  if (*cap >= need) return VAB_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(*buf);
  *buf = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc((void**)buf, need * sizeof(T));
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ozaki workspace");
  *cap = need;
  return VAB_OK;
}

// Split-K form: every path holds nchunk problems, chunk c contracting over elements
// [c K, min((c + 1) K, Ktotal)) of an axis of length Ktotal (apc / bpc: doubles between chunks);
// problem (path b, chunk c) writes C[(b nchunk + c) cps + r ldc + col] -- partial products summed by
// the caller in chunk order.
int ozaki_gemm_chunked(vab_ctx* ctx, int paths, int nchunk, int M, int N, int K, int Ktotal,
                       const double* A, long long lda, long long aks, long long aps, long long apc,
                       const double* B, long long ldb, long long bks, long long bps, long long bpc,
                       double* C, long long ldc, long long cps, const int* active_dev) {
  const int P = paths * nchunk;
  if (paths < 1 || nchunk < 1 || M < 1 || N < 1 || K < 1 || K > KP) return vab_fail(ctx, VAB_ERR_INVALID, "ozaki_gemm: bad sizes (K <= 128)");
  if (!ctx->oz) ctx->oz = new OzakiWork();
  OzakiWork* w = ctx->oz;
  if (!w->enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) return vab_fail(ctx, VAB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    w->enc = (EncodeTiledFn)fn;
  }
  if (!w->attr) {
    cudaError_t e = cudaFuncSetAttribute(ozaki_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ozaki_gemm smem opt-in");
    w->attr = true;
  }
  const int m_tiles = (M + TM - 1) / TM, n_tiles = (N + TN - 1) / TN;
  const int Mpad = m_tiles * TM, Npad = n_tiles * TN;
  const long long rowsA = (long long)P * Mpad, rowsB = (long long)P * Npad;
  int rc = oz_reserve(ctx, &w->pa, &w->pa_cap, (size_t)NS * rowsA * KP);
  if (rc == VAB_OK) rc = oz_reserve(ctx, &w->pb, &w->pb_cap, (size_t)NS * rowsB * KP);
  if (rc == VAB_OK) rc = oz_reserve(ctx, &w->ea, &w->ea_cap, (size_t)rowsA);
  if (rc == VAB_OK) rc = oz_reserve(ctx, &w->eb, &w->eb_cap, (size_t)rowsB);
  if (rc != VAB_OK) return rc;
  CUtensorMap mapA, mapB;
  rc = make_plane_map(ctx, w->enc, &mapA, w->pa, rowsA, TM);
  if (rc == VAB_OK) rc = make_plane_map(ctx, w->enc, &mapB, w->pb, rowsB, TN);
  if (rc != VAB_OK) return rc;
  cudaStream_t st = ctx->stream;
  const int kt = Ktotal;
  ozaki_slice_kernel<<<dim3((Mpad * 32 + 255) / 256, P), 256, 0, st>>>(A, lda, aps, M, K, Mpad, w->pa, rowsA * KP, w->ea, aks, active_dev, nchunk, apc, kt);
  ozaki_slice_kernel<<<dim3((Npad * 32 + 255) / 256, P), 256, 0, st>>>(B, ldb, bps, N, K, Npad, w->pb, rowsB * KP, w->eb, bks, active_dev, nchunk, bpc, kt);
  OzakiParams q;
  q.M = M; q.N = N; q.P = P; q.m_tiles = m_tiles; q.n_tiles = n_tiles; q.Mpad = Mpad; q.Npad = Npad;
  q.ea = w->ea; q.eb = w->eb; q.C = C; q.ldc = ldc; q.cps = cps; q.active = active_dev; q.pdiv = nchunk;
  ozaki_mma_kernel<<<P * m_tiles * n_tiles, 128, SMEM_BYTES, st>>>(mapA, mapB, q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "ozaki_gemm launch");
  ctx->launches += 3;
  return VAB_OK;
}

int ozaki_gemm(vab_ctx* ctx, int P, int M, int N, int K,
               const double* A, long long lda, long long aks, long long aps,
               const double* B, long long ldb, long long bks, long long bps,
               double* C, long long ldc, long long cps, const int* active_dev) {
  return ozaki_gemm_chunked(ctx, P, 1, M, N, K, K, A, lda, aks, aps, 0, B, ldb, bks, bps, 0, C, ldc, cps, active_dev);
}

// out_host[12] (8..11: version 2 -- max rel err, ms, TFLOP/s-equivalent of the kernel, of kernel + planes): 0 max |C - Cref| / max |Cref|, 1 ms digit planes (both operands), 2 ms tcgen05 kernel,
//              3 ms total, 4 fp64-equivalent TFLOP/s of the total, 5 the same for the tcgen05 kernel alone,
//              6 ms of the fp64 FMA reference kernel, 7 max |Cref|
extern "C" int vab_ozaki_gemm_probe(vab_ctx* ctx, int32_t P, int32_t M, int32_t N, int32_t K, int32_t reps,
                                    double spread, double* out_host) {
  if (!ctx || !out_host) return VAB_ERR_INVALID;
  if (P < 1 || M < 1 || N < 1 || K < 1 || K > KP || reps < 1) return vab_fail(ctx, VAB_ERR_INVALID, "ozaki_gemm_probe: bad sizes (K <= 128)");
  cudaSetDevice(ctx->device);
  int rc = VAB_OK;
  cudaStream_t st = ctx->stream;
  const int m_tiles = (M + TM - 1) / TM, n_tiles = (N + TN - 1) / TN;
  const int Mpad = m_tiles * TM, Npad = n_tiles * TN;
  const long long rowsA = (long long)P * Mpad, rowsB = (long long)P * Npad;
  double *A = nullptr, *B = nullptr, *C = nullptr, *Cref = nullptr, *errbuf = nullptr;
  int8_t *pa = nullptr, *pb = nullptr;
  int *ea = nullptr, *eb = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  EncodeTiledFn enc = nullptr;
  CUtensorMap mapA, mapB, mapA2, mapB2;
  OzakiParams q, q2;
  float ms_slice = 0.f, ms_mma = 0.f, ms_ref = 0.f, ms_mma2 = 0.f;
  double err2 = 0.0;
  int grid2 = ctx->num_sms;
  std::vector<double> herr(2 * 256);
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    OZ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { rc = vab_fail(ctx, VAB_ERR_CUDA, "cuTensorMapEncodeTiled not available"); goto done; }
    enc = (EncodeTiledFn)fn;
  }
  OZ_CUDA(cudaMalloc((void**)&A, sizeof(double) * (size_t)P * M * K));
  OZ_CUDA(cudaMalloc((void**)&B, sizeof(double) * (size_t)P * N * K));
  OZ_CUDA(cudaMalloc((void**)&C, sizeof(double) * (size_t)P * M * N));
  OZ_CUDA(cudaMalloc((void**)&Cref, sizeof(double) * (size_t)P * M * N));
  OZ_CUDA(cudaMalloc((void**)&errbuf, sizeof(double) * 2 * 256));
  OZ_CUDA(cudaMalloc((void**)&pa, (size_t)NS * rowsA * KP));
  OZ_CUDA(cudaMalloc((void**)&pb, (size_t)NS * rowsB * KP));
  OZ_CUDA(cudaMalloc((void**)&ea, sizeof(int) * rowsA));
  OZ_CUDA(cudaMalloc((void**)&eb, sizeof(int) * rowsB));
  for (int k = 0; k < 4; ++k) OZ_CUDA(cudaEventCreate(&ev[k]));
  {
    const long long na = (long long)P * M * K, nb = (long long)P * N * K;
    ozaki_fill_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(A, na, 0x1234ull, spread);
    ozaki_fill_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(B, nb, 0x9876ull, spread);
  }
  rc = make_plane_map(ctx, enc, &mapA, pa, rowsA, TM);
  if (rc != VAB_OK) goto done;
  rc = make_plane_map(ctx, enc, &mapB, pb, rowsB, TN);
  if (rc != VAB_OK) goto done;
  rc = make_plane_map(ctx, enc, &mapA2, pa, rowsA, TM, KH, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != VAB_OK) goto done;
  rc = make_plane_map(ctx, enc, &mapB2, pb, rowsB, TN2, KH, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != VAB_OK) goto done;
  OZ_CUDA(cudaFuncSetAttribute(ozaki_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
  OZ_CUDA(cudaFuncSetAttribute(ozaki_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  q.M = M; q.N = N; q.P = P; q.m_tiles = m_tiles; q.n_tiles = n_tiles; q.Mpad = Mpad; q.Npad = Npad;
  q.ea = ea; q.eb = eb; q.C = C; q.ldc = N; q.cps = (long long)M * N; q.active = nullptr; q.pdiv = 1;
  for (int rep = 0; rep <= reps; ++rep) {            // rep 0 = warm-up
    if (rep == 1) OZ_CUDA(cudaEventRecord(ev[0], st));
    ozaki_slice_kernel<<<dim3((Mpad * 32 + 255) / 256, P), 256, 0, st>>>(A, K, (long long)M * K, M, K, Mpad, pa, rowsA * KP, ea);
    ozaki_slice_kernel<<<dim3((Npad * 32 + 255) / 256, P), 256, 0, st>>>(B, K, (long long)N * K, N, K, Npad, pb, rowsB * KP, eb);
    if (rep == reps) OZ_CUDA(cudaEventRecord(ev[1], st));
  }
  for (int rep = 0; rep <= reps; ++rep) {
    if (rep == 1) OZ_CUDA(cudaEventRecord(ev[2], st));
    ozaki_mma_kernel<<<P * m_tiles * n_tiles, 128, SMEM_BYTES, st>>>(mapA, mapB, q);
    if (rep == reps) OZ_CUDA(cudaEventRecord(ev[3], st));
  }
  OZ_CUDA(cudaGetLastError());
  OZ_CUDA(cudaEventSynchronize(ev[3]));
  OZ_CUDA(cudaEventElapsedTime(&ms_slice, ev[0], ev[1]));
  OZ_CUDA(cudaEventElapsedTime(&ms_mma, ev[2], ev[3]));
  ms_slice /= reps; ms_mma /= reps;
  OZ_CUDA(cudaEventRecord(ev[0], st));
  ozaki_ref_kernel<<<dim3((N + 31) / 32, (M + 7) / 8, P), dim3(32, 8), 0, st>>>(A, B, Cref, M, N, K);
  OZ_CUDA(cudaEventRecord(ev[1], st));
  ozaki_err_kernel<<<256, 256, 0, st>>>(C, Cref, (long long)P * M * N, errbuf);
  OZ_CUDA(cudaMemcpyAsync(herr.data(), errbuf, sizeof(double) * 2 * 256, cudaMemcpyDeviceToHost, st));
  OZ_CUDA(cudaStreamSynchronize(st));
  OZ_CUDA(cudaEventElapsedTime(&ms_ref, ev[0], ev[1]));
  // ---- version 2 (persistent, pipelined) on the same planes; its result replaces C
  q2 = q; q2.n_tiles = Npad / TN2;
  if (const char* g = getenv("VAB_OZAKI_GRID")) grid2 = atoi(g) > 0 ? atoi(g) : grid2;
  if (grid2 > P * m_tiles * q2.n_tiles) grid2 = P * m_tiles * q2.n_tiles;
  if (getenv("VAB_OZAKI_V2") == nullptr || atoi(getenv("VAB_OZAKI_V2")) != 0) {
    OZ_CUDA(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)P * M * N, st));
    for (int rep = 0; rep <= reps; ++rep) {
      if (rep == 1) OZ_CUDA(cudaEventRecord(ev[2], st));
      ozaki_mma2_kernel<<<grid2, 192, SMEM2_BYTES, st>>>(mapA2, mapB2, q2);
      if (rep == reps) OZ_CUDA(cudaEventRecord(ev[3], st));
    }
    OZ_CUDA(cudaGetLastError());
    OZ_CUDA(cudaEventSynchronize(ev[3]));
    OZ_CUDA(cudaEventElapsedTime(&ms_mma2, ev[2], ev[3]));
    ms_mma2 /= reps;
    std::vector<double> herr2(2 * 256);
    ozaki_err_kernel<<<256, 256, 0, st>>>(C, Cref, (long long)P * M * N, errbuf);
    OZ_CUDA(cudaMemcpyAsync(herr2.data(), errbuf, sizeof(double) * 2 * 256, cudaMemcpyDeviceToHost, st));
    OZ_CUDA(cudaStreamSynchronize(st));
    double e2 = 0.0, m2 = 0.0;
    for (int k = 0; k < 256; ++k) { e2 = fmax(e2, herr2[2 * k]); m2 = fmax(m2, herr2[2 * k + 1]); }
    err2 = m2 > 0.0 ? e2 / m2 : 0.0;
    ctx->launches += reps + 2;
  }
  ctx->launches += 3 * (reps + 1) + 4;
  {
    double e = 0.0, m = 0.0;
    for (int k = 0; k < 256; ++k) { e = fmax(e, herr[2 * k]); m = fmax(m, herr[2 * k + 1]); }
    const double flops = 2.0 * P * (double)M * N * K;
    out_host[0] = m > 0.0 ? e / m : 0.0;
    out_host[1] = ms_slice; out_host[2] = ms_mma; out_host[3] = ms_slice + ms_mma;
    out_host[4] = flops / ((ms_slice + ms_mma) * 1e-3) / 1e12;
    out_host[5] = flops / (ms_mma * 1e-3) / 1e12;
    out_host[6] = ms_ref; out_host[7] = m;
    out_host[8] = err2; out_host[9] = ms_mma2;
    out_host[10] = ms_mma2 > 0.f ? flops / (ms_mma2 * 1e-3) / 1e12 : 0.0;
    out_host[11] = ms_mma2 > 0.f ? flops / ((ms_slice + ms_mma2) * 1e-3) / 1e12 : 0.0;
  }
done:
  for (int k = 0; k < 4; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cref); cudaFree(errbuf); cudaFree(pa); cudaFree(pb); cudaFree(ea); cudaFree(eb);
  return rc;
}
