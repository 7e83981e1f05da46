// Neural-network action + gradient: replaces the ADOL-C tape of va_nnet.Annealer.A_gaussian
// (va_nnet.py:111-255: me_gaussian :117-173, fe_gaussian :175-255, disc_forwardmap :260-264).
//
// Layout (reference): X flat = for each example m the layers concatenated (va_nnet.py:149-152),
// i.e. X.reshape(M, NDnet); P flat = [W_0 (d_1 x d_0 row-major), b_0, W_1, b_1, ...]
// (va_nnet.py:194-207); XP = X ++ P[Pidx] (va_nnet.py:468-473).
//
// One CTA owns a tile of TM examples of one path and walks all layers: the examples of a tile are
// independent across layers, so the three contractions of a layer
//     Z  = X_n W_n^T + b_n              -> S = act(Z), E = X_{n+1} - S, fe += sum E^2
//     GX_n += Delta W_n                    Delta = -2 cf E act'(Z)
//     GW_n  = Delta^T X_n  (contraction over the examples), Gb_n = sum_m Delta
// are done on shared-memory tiles without leaving the CTA; every gradient row of X is written
// once.  Weight gradients are per-tile partials reduced in a fixed order by nn_reduce_kernel
// (bit-reproducible, no atomics).  fp64 on the CUDA cores: tcgen05 has no f64 kind, and parity is
// 1e-10.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vab_ctx.h"

namespace {

constexpr int NT = 256;

struct NnParams {
  const double* XP; long long ldxp;
  double* G; long long ldg;
  int B, M, NL, NDnet, NP, NPest, act, TM, ntiles, dpitch, wcap;
  long long NDens;
  const int* structure;   // (NL)
  const int* xoff;        // (NL+1) offsets of the layers inside an example
  const int* woff;        // (NL-1)
  const int* boff;        // (NL-1)
  const int* pmap;        // (NP) -> estimated index or -1
  const double* pfix; long long pfix_stride;
  double* pfull;          // (B, NP) dense parameter vectors gathered before the fused kernel
  const int* slot_in;     // (d_0) column of data_in or -1
  const int* slot_out;    // (d_last)
  int n_Lin, n_Lout;
  const double* data_in;  // (M, n_Lin)
  const double* data_out; // (M, n_Lout)
  double wm_in, wm_out;   // 2 cm RM_in, 2 cm RM_out
  double cf2;             // 2 RF / ((NDnet - d_0) M)
  double cf2_num, cf2_den; // 2 RF0 and (NDnet - d_0) M: cf2 of a path on its own rung = cf2_num * rf_path[b] / cf2_den
  const double* rf_path;  // (B) or nullptr
  const int* active;
  double* partials;       // (B, nparts, 2): me, fe
  double* gwpart;         // (B, ngw, NP)
  int nparts, ngw;        // partials per path summed by nn_reduce_kernel (fused: ntiles, ntiles)
  // split design (nn_fb_kernel / nn_gw_kernel)
  int TMF, nmt;           // examples per forward/backward tile, tiles per layer
  int gw_klen, gw_nsplit; // examples per split of the weight-gradient GEMM, number of splits
  int gw_njs;             // CTAs sharing the row tiles of one W block
  int gw_kp;              // examples per panel of the weight-gradient GEMM (multiple of 8)
  const double* one;      // device constant 1.0 (source of the ones column)
  double* me_parts;       // (B, n_me) per-block measurement-error sums of nn_fix_kernel, or nullptr
  int n_me;
  int fba_pxn, fba_pd;    // nn_fba_kernel: pitch of the all-layer state / gradient tiles, of the Delta tile
  int fb_T, fb_nbuf;      // nn_fb_kernel: example tiles per CTA (W_n resident), tile buffers (1 or 2)
  double* dbuf;           // (B, M, NDnet - d_0) Delta of every layer
  double* lam;            // (B, M, NDnet - d_0) lambda = direct term of the next layer's gradient rows
};

__device__ __forceinline__ double act_f(int act, double z) {
  if (act == VAB_ACT_SIGMOID) return 1.0 / (1.0 + exp(-z));
  if (act == VAB_ACT_TANH) return tanh(z);
  return z;
}
__device__ __forceinline__ double act_d(int act, double s) {   // derivative through the output s
  if (act == VAB_ACT_SIGMOID) return s * (1.0 - s);
  if (act == VAB_ACT_TANH) return 1.0 - s * s;
  return 1.0;
}

// dense parameter vector of every path: estimated entries from XP, the others from pfix
// (va_nnet.py:180-190); one coalesced pass so that the fused kernel stages weights with plain
// independent loads instead of two dependent ones per element
__global__ void nn_gather_params_kernel(const __grid_constant__ NnParams P) {
  const int b = blockIdx.y;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.NP) return;
  const int e = P.pmap[k];
  P.pfull[(long long)b * P.NP + k] =
      e >= 0 ? P.XP[(long long)b * P.ldxp + P.NDens + e] : P.pfix[(long long)b * P.pfix_stride + k];
}

// D(8x8) += A(8x4, row) * B(4x8, col) in fp64 on the tensor pipe (DMMA).  Fragment layout
// (lane l): a = A[l/4][l%4], b = B[l%4][l/4], c0/c1 = C[l/4][2(l%4)], C[l/4][2(l%4)+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double bb) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(bb));
}

constexpr int NTILE = 7;    // 8x8 output tiles a warp accumulates at once (A fragment reused)

// One CTA = one tile of TM examples (multiple of 8) of one path, all layers.  Shared-memory
// tiles have pitches = 4 (mod 8) doubles so that the 8x4 / 4x8 fragment loads are bank-conflict
// free, and are zero-padded to multiples of 8 so that the MMAs need no edge handling.
__global__ void __launch_bounds__(NT) nn_fused_kernel(const __grid_constant__ NnParams P) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;            // fragment coordinates
  const int TM = P.TM, dp = P.dpitch, wcap = P.wcap;
  double* Xs = sm;                       // [TM][dp] states of layer n
  double* Xn = Xs + TM * dp;             // layer n+1
  double* GXc = Xn + TM * dp;            // gradient rows of layer n
  double* GXn = GXc + TM * dp;           // gradient rows of layer n+1
  double* Ds = GXn + TM * dp;            // Delta chunk [TM][dp]
  double* Ws = Ds + TM * dp;             // [JBP][dnp]
  double* bs = Ws + wcap;                // [JBP]
  __shared__ double red[2][NT / 32];

  const int m0 = tile * TM;
  const int rows = min(TM, P.M - m0);
  const double* xp = P.XP + (long long)b * P.ldxp;
  double* gp = P.G ? P.G + (long long)b * P.ldg : nullptr;
  const double* pfull = P.pfull + (long long)b * P.NP;
  double* gw = P.gwpart + ((long long)b * P.ntiles + tile) * P.NP;
  auto param = [&](int k) -> double { return __ldg(pfull + k); };
  double me_acc = 0.0, fe_acc = 0.0;
  const double cf2 = (P.rf_path != nullptr) ? P.cf2_num * __ldg(P.rf_path + b) / P.cf2_den : P.cf2;
  const int MTL = TM >> 3;                            // 8-row tiles of examples

  // layer 0 states (zero-padded) + measurement term of the input layer
  {
    const int d0 = P.structure[0], d0P = (d0 + 7) & ~7;
    for (int idx = tid; idx < TM * d0P; idx += NT) {
      const int m = idx / d0P, i = idx - m * d0P;
      double x = 0.0, gx = 0.0;
      if (m < rows && i < d0) {
        x = xp[(long long)(m0 + m) * P.NDnet + i];
        const int s = P.slot_in[i];
        if (s >= 0) {
          const double diff = x - P.data_in[(long long)(m0 + m) * P.n_Lin + s];
          me_acc = fma(P.wm_in * diff, diff, me_acc);
          gx = P.wm_in * diff;
        }
      }
      Xs[m * dp + i] = x;
      GXc[m * dp + i] = gx;
    }
  }
  for (int n = 0; n < P.NL - 1; ++n) {
    const int dn = P.structure[n], dn1 = P.structure[n + 1];
    const int dnP = (dn + 7) & ~7, dn1P = (dn1 + 7) & ~7;
    const int dnp = dnP + 4;                              // pitch of the staged weight rows
    const int xo1 = P.xoff[n + 1];
    const bool lastl = (n + 1 == P.NL - 1);
    __syncthreads();
    for (int idx = tid; idx < TM * dn1P; idx += NT) {
      const int m = idx / dn1P, j = idx - m * dn1P;
      Xn[m * dp + j] = (m < rows && j < dn1) ? xp[(long long)(m0 + m) * P.NDnet + xo1 + j] : 0.0;
    }
    int JB = (wcap / dnp) & ~7;
    if (JB > dn1P) JB = dn1P;
    for (int j0 = 0; j0 < dn1; j0 += JB) {
      const int jb = min(JB, dn1 - j0);                   // real rows of W in this chunk
      const int jbP = (jb + 7) & ~7;
      __syncthreads();
      for (int idx = tid; idx < jbP * dnP; idx += NT) {
        const int j = idx / dnP, k = idx - j * dnP;
        Ws[j * dnp + k] = (j < jb && k < dn) ? param(P.woff[n] + (j0 + j) * dn + k) : 0.0;
      }
      for (int j = tid; j < jbP; j += NT) bs[j] = (j < jb) ? param(P.boff[n] + j0 + j) : 0.0;
      __syncthreads();
      // ---- (1) Z = X W^T: warp tasks = (8-row tile of examples) x (group of NTILE column tiles)
      {
        const int NTL = jbP >> 3, NG = (NTL + NTILE - 1) / NTILE;
        for (int task = warp; task < MTL * NG; task += NT / 32) {
          const int mt = task / NG, g = task - mt * NG;
          const int nt0 = g * NTILE, ntn = min(NTILE, NTL - nt0);
          double c0[NTILE], c1[NTILE];
#pragma unroll
          for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
          const double* arow = Xs + (mt * 8 + lr) * dp + lc;
          const double* brow = Ws + (nt0 * 8 + lr) * dnp + lc;
          for (int k = 0; k < dnP; k += 4) {
            const double a = arow[k];
#pragma unroll
            for (int t = 0; t < NTILE; ++t)
              if (t < ntn) dmma(c0[t], c1[t], a, brow[t * 8 * dnp + k]);
          }
          // epilogue on the accumulator fragments: residual, lambda, Delta
          const int m = mt * 8 + lr;
#pragma unroll
          for (int t = 0; t < NTILE; ++t) {
            if (t >= ntn) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = (nt0 + t) * 8 + 2 * lc + h;
              const double z = (h ? c1[t] : c0[t]) + bs[j];
              double lam = 0.0, dl = 0.0, gx = 0.0;
              if (m < rows && j < jb) {
                const double sv = act_f(P.act, z);
                const double xn1 = Xn[m * dp + j0 + j];
                const double e = xn1 - sv;
                lam = cf2 * e;
                fe_acc = fma(lam, e, fe_acc);
                dl = -lam * act_d(P.act, sv);
                gx = lam;
                if (lastl) {
                  const int so = P.slot_out[j0 + j];
                  if (so >= 0) {
                    const double diff = xn1 - P.data_out[(long long)(m0 + m) * P.n_Lout + so];
                    me_acc = fma(P.wm_out * diff, diff, me_acc);
                    gx += P.wm_out * diff;
                  }
                }
              }
              if (j < jb) GXn[m * dp + j0 + j] = gx;
              Ds[m * dp + j] = dl;                         // zero in the padding
            }
          }
        }
      }
      __syncthreads();
      // ---- (2) GX_n += Delta W: (8-row tile of examples) x (column tiles over d_n)
      {
        const int NTL = dnP >> 3, NG = (NTL + NTILE - 1) / NTILE;
        for (int task = warp; task < MTL * NG; task += NT / 32) {
          const int mt = task / NG, g = task - mt * NG;
          const int nt0 = g * NTILE, ntn = min(NTILE, NTL - nt0);
          double c0[NTILE], c1[NTILE];
#pragma unroll
          for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
          const double* arow = Ds + (mt * 8 + lr) * dp + lc;          // A[m][j]
          const double* brow = Ws + lc * dnp + nt0 * 8 + lr;          // B[j][k] = W[j][k]
          for (int j = 0; j < jbP; j += 4) {
            const double a = arow[j];
#pragma unroll
            for (int t = 0; t < NTILE; ++t)
              if (t < ntn) dmma(c0[t], c1[t], a, brow[j * dnp + t * 8]);
          }
          const int m = mt * 8 + lr;
#pragma unroll
          for (int t = 0; t < NTILE; ++t) {
            if (t >= ntn) continue;
            const int k = (nt0 + t) * 8 + 2 * lc;
            GXc[m * dp + k] += c0[t];
            GXc[m * dp + k + 1] += c1[t];
          }
        }
      }
      // ---- (3) per-tile partial of GW = Delta^T X: (8-row tile over j) x (column tiles over d_n)
      {
        const int JTL = jbP >> 3, NTL = dnP >> 3, NG = (NTL + NTILE - 1) / NTILE;
        for (int task = warp; task < JTL * NG; task += NT / 32) {
          const int jt = task / NG, g = task - jt * NG;
          const int nt0 = g * NTILE, ntn = min(NTILE, NTL - nt0);
          double c0[NTILE], c1[NTILE];
#pragma unroll
          for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
          const double* arow = Ds + lc * dp + jt * 8 + lr;            // A[j][m] = Delta[m][j]
          const double* brow = Xs + lc * dp + nt0 * 8 + lr;           // B[m][k] = X[m][k]
          for (int m = 0; m < TM; m += 4) {
            const double a = arow[m * dp];
#pragma unroll
            for (int t = 0; t < NTILE; ++t)
              if (t < ntn) dmma(c0[t], c1[t], a, brow[m * dp + t * 8]);
          }
          const int j = jt * 8 + lr;
          if (j < jb) {
            double* grow = gw + P.woff[n] + (long long)(j0 + j) * dn;
#pragma unroll
            for (int t = 0; t < NTILE; ++t) {
              if (t >= ntn) continue;
              const int k = (nt0 + t) * 8 + 2 * lc;
              if (k < dn) grow[k] = c0[t];
              if (k + 1 < dn) grow[k + 1] = c1[t];
            }
          }
        }
        for (int j = tid; j < jb; j += NT) {
          double acc = 0.0;
          for (int m = 0; m < TM; ++m) acc += Ds[m * dp + j];
          gw[P.boff[n] + j0 + j] = acc;
        }
      }
    }
    __syncthreads();
    // gradient rows of layer n are complete
    if (gp) {
      const int xo = P.xoff[n];
      for (int idx = tid; idx < rows * dn; idx += NT) {
        const int m = idx / dn, k = idx - m * dn;
        gp[(long long)(m0 + m) * P.NDnet + xo + k] = GXc[m * dp + k];
      }
    }
    __syncthreads();
    double* t = Xs; Xs = Xn; Xn = t;
    t = GXc; GXc = GXn; GXn = t;
    // the new GXn will be accumulated into (as GXc) two layers on: clear its padding columns
  }
  if (gp) {
    const int dl = P.structure[P.NL - 1], xo = P.xoff[P.NL - 1];
    for (int idx = tid; idx < rows * dl; idx += NT) {
      const int m = idx / dl, k = idx - m * dl;
      gp[(long long)(m0 + m) * P.NDnet + xo + k] = GXc[m * dp + k];
    }
  }
  // per-tile me / fe (fixed-order block reduction)
  for (int sft = 16; sft > 0; sft >>= 1) {
    me_acc += __shfl_down_sync(0xffffffffu, me_acc, sft);
    fe_acc += __shfl_down_sync(0xffffffffu, fe_acc, sft);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = me_acc; red[1][tid >> 5] = fe_acc; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < NT / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    P.partials[((long long)b * P.ntiles + tile) * 2 + 0] = 0.5 * a;
    P.partials[((long long)b * P.ntiles + tile) * 2 + 1] = 0.5 * c;
  }
}

// =============================================================================================
// Split design for layers whose weight matrix fits in shared memory (the twin and bar-image
// networks of the examples): three kernels instead of one example-tile kernel walking all layers.
//   nn_fb_kernel   one CTA = (layer n, tile of TMF examples): Z = X_n W_n^T, residual / lambda /
//                  Delta in the accumulator epilogue, then Delta W_n with the same staged W_n.
//                  Every (layer, tile) is independent (the states of all layers are unknowns), so
//                  the grid is (tiles, layers, paths).  Delta goes to a global buffer for the
//                  weight gradient; lambda (the direct term of layer n+1's gradient rows) to a
//                  second one.
//   nn_fix_kernel  G rows of the middle layers += lambda (two addends per entry: order-free).
//   nn_gw_kernel   GW_n = Delta_n^T X_n as a split-K GEMM over the example axis: one CTA =
//                  (split, layer, path) keeps the whole d_{n+1} x d_n block in accumulator
//                  fragments while panels of 32 examples stream through shared memory; the few
//                  per-split partials are summed in split order by nn_reduce_kernel.
// This removes the per-tile weight-gradient partials (C4: 1.3 GB of traffic per batch) and the
// chain of barrier-separated phases per layer of the fused kernel.
// 8-byte asynchronous copy global -> shared (LDGSTS); nbytes = 0 writes zeros (padding)
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, int nbytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// tile[r][c] (pitch) <- src[(row0 + r) * ld + c] for r < nr, c < ncP; entries with row0 + r >= row_end
// or c >= nc are zero-filled.  One warp per row, lanes across the columns; all loads in flight.
__device__ __forceinline__ void stage_tile_async(double* tile, int pitch, const double* src, long long ld,
                                                 int row0, int row_end, int nr, int nc, int ncP,
                                                 const double* one = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nr; r += nwarps) {
    const bool rv = row0 + r < row_end;
    const double* srow = src + (long long)(rv ? row0 + r : row0) * ld;
    double* drow = tile + r * pitch;
    for (int c = lane; c < ncP; c += 32) {
      const bool v = rv && c < nc;
      if (one != nullptr && rv && c == nc) cp_async8(drow + c, one, 8);     // column of ones (bias gradient)
      else cp_async8(drow + c, srow + (v ? c : 0), v ? 8 : 0);
    }
  }
}

constexpr int GW_MAXTASK = 2;      // (row tile, column group) tasks a warp may own; wider blocks are
                                   // split over the rows of W among several CTAs (P.gw_njs)

constexpr int NTF = 512;           // largest CTA of nn_fb_kernel (launched with 256 or 512 threads)
// One CTA = (layer n, P.fb_T consecutive tiles of TMF examples): W_n and b_n are staged once and stay
// resident; the state tiles are double-buffered (P.fb_nbuf = 2) so that the cp.async copies of tile
// i+1 run under the MMAs of tile i.
__global__ void __launch_bounds__(NTF, 1) nn_fb_kernel(const __grid_constant__ NnParams P) {
  extern __shared__ double sm[];
  const int n = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, nthr = blockDim.x;
  const int lr = lane >> 2, lc = lane & 3;
  const int TM = P.TMF;
  const int dn = P.structure[n], dn1 = P.structure[n + 1];
  const int dnP = (dn + 7) & ~7, dn1P = (dn1 + 7) & ~7;
  const int px = dnP + 4, pd = dn1P + 4;
  const int tile_sz = TM * (px + pd);    // one buffer: [TM][px] states of layer n, then [TM][pd] x_{n+1} / Delta
  double* Ws = sm + P.fb_nbuf * tile_sz; // [dn1P][px]  W_n
  double* bs = Ws + dn1P * px;           // [dn1P]
  __shared__ double red[2][NTF / 32];
  const int xo = P.xoff[n], xo1 = P.xoff[n + 1], d0 = P.structure[0];
  const int ND1 = P.NDnet - d0;
  const bool lastl = (n + 1 == P.NL - 1);
  const double* xp = P.XP + (long long)b * P.ldxp;
  double* gp = P.G + (long long)b * P.ldg;
  double* dbuf = P.dbuf + (long long)b * P.M * ND1;
  double* lam = P.lam + (long long)b * P.M * ND1;
  const double* pfull = P.pfull + (long long)b * P.NP;
  const double cf2 = (P.rf_path != nullptr) ? P.cf2_num * __ldg(P.rf_path + b) / P.cf2_den : P.cf2;
  const int t_lo = blockIdx.x * P.fb_T, t_hi = min(P.nmt, t_lo + P.fb_T);
  // stage X_n, X_{n+1} (into the Delta tile: the epilogue of (1) reads x_{n+1}[m][j] and overwrites
  // the same entry with Delta[m][j]) with asynchronous copies
  auto stage = [&](int buf, int mtb) {
    double* Xs = sm + buf * tile_sz;
    stage_tile_async(Xs, px, xp + xo, P.NDnet, mtb * TM, P.M, TM, dn, dnP);
    stage_tile_async(Xs + TM * px, pd, xp + xo1, P.NDnet, mtb * TM, P.M, TM, dn1, dn1P);
  };
  stage_tile_async(Ws, px, pfull + P.woff[n], dn, 0, dn1, dn1P, dn, dnP);
  for (int j = tid; j < dn1P; j += nthr) cp_async8(bs + j, pfull + P.boff[n] + (j < dn1 ? j : 0), j < dn1 ? 8 : 0);
  stage(0, t_lo);
  cp_async_commit();
  const int MTL = TM >> 3;
  for (int mtb = t_lo; mtb < t_hi; ++mtb) {
    const int buf = (P.fb_nbuf == 2) ? ((mtb - t_lo) & 1) : 0;
    if (P.fb_nbuf == 2 && mtb + 1 < t_hi) {
      stage(buf ^ 1, mtb + 1);                              // next tile streams in under this tile's MMAs
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    double* Xs = sm + buf * tile_sz;
    double* Ds = Xs + TM * px;
    const int m0 = mtb * TM;
    const int rows = min(TM, P.M - m0);
    double me_acc = 0.0, fe_acc = 0.0;
    // ---- (1) Z = X W^T, epilogue: residual, lambda, Delta
    {
      // column tiles per task: NTILE, or fewer when that would leave warps without a task
      const int NTL = dn1P >> 3;
      int gsz = NTILE;
      if (MTL * ((NTL + NTILE - 1) / NTILE) < nwarps) gsz = max(1, (MTL * NTL) / nwarps);
      const int NG = (NTL + gsz - 1) / gsz;
      for (int task = warp; task < MTL * NG; task += nwarps) {
        const int mt = task / NG, g = task - mt * NG;
        const int nt0 = g * gsz, ntn = min(gsz, NTL - nt0);
        double c0[NTILE], c1[NTILE];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
        const double* arow = Xs + (mt * 8 + lr) * px + lc;
        const double* brow = Ws + (nt0 * 8 + lr) * px + lc;
        for (int k = 0; k < dnP; k += 4) {
          const double a = arow[k];
#pragma unroll
          for (int t = 0; t < NTILE; ++t)
            if (t < ntn) dmma(c0[t], c1[t], a, brow[t * 8 * px + k]);
        }
        const int m = mt * 8 + lr;
        const long long grow = (long long)(m0 + m);
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
          if (t >= ntn) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = (nt0 + t) * 8 + 2 * lc + h;
            const double z = (h ? c1[t] : c0[t]) + bs[j];
            double dl = 0.0;
            if (m < rows && j < dn1) {
              const double sv = act_f(P.act, z);
              const double xn1 = Ds[m * pd + j];
              const double e = xn1 - sv;
              const double lm = cf2 * e;
              fe_acc = fma(lm, e, fe_acc);
              dl = -lm * act_d(P.act, sv);
              // direct term of layer n+1's gradient rows: the last layer has no back term and is
              // written here; the others are added to the back term by nn_fix_kernel
              if (lastl) gp[grow * P.NDnet + xo1 + j] = lm;
              else lam[grow * ND1 + (xo1 - d0) + j] = lm;
              dbuf[grow * ND1 + (xo1 - d0) + j] = dl;
            }
            Ds[m * pd + j] = dl;                            // zero in the padding
          }
        }
      }
    }
    __syncthreads();
    // ---- (2) back term of layer n's gradient rows: Delta W
    {
      const int NTL = dnP >> 3;
      int gsz = NTILE;
      if (MTL * ((NTL + NTILE - 1) / NTILE) < nwarps) gsz = max(1, (MTL * NTL) / nwarps);
      const int NG = (NTL + gsz - 1) / gsz;
      for (int task = warp; task < MTL * NG; task += nwarps) {
        const int mt = task / NG, g = task - mt * NG;
        const int nt0 = g * gsz, ntn = min(gsz, NTL - nt0);
        double c0[NTILE], c1[NTILE];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
        const double* arow = Ds + (mt * 8 + lr) * pd + lc;        // A[m][j]
        const double* brow = Ws + lc * px + nt0 * 8 + lr;         // B[j][k] = W[j][k]
        for (int j = 0; j < dn1P; j += 4) {
          const double a = arow[j];
#pragma unroll
          for (int t = 0; t < NTILE; ++t)
            if (t < ntn) dmma(c0[t], c1[t], a, brow[j * px + t * 8]);
        }
        const int m = mt * 8 + lr;
        const long long grow = (long long)(m0 + m);
        if (m < rows) {
#pragma unroll
          for (int t = 0; t < NTILE; ++t) {
            if (t >= ntn) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int k = (nt0 + t) * 8 + 2 * lc + h;
              if (k >= dn) continue;
              gp[grow * P.NDnet + xo + k] = h ? c1[t] : c0[t];
            }
          }
        }
      }
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
      me_acc += __shfl_down_sync(0xffffffffu, me_acc, sft);
      fe_acc += __shfl_down_sync(0xffffffffu, fe_acc, sft);
    }
    if (lane == 0) { red[0][warp] = me_acc; red[1][warp] = fe_acc; }
    __syncthreads();            // also: every warp is done with this buffer before it is refilled
    if (tid == 0) {
      double a = 0.0, c = 0.0;
      for (int w = 0; w < nwarps; ++w) { a += red[0][w]; c += red[1][w]; }
      const long long item = ((long long)b * (P.NL - 1) + n) * P.nmt + mtb;
      P.partials[item * 2 + 0] = 0.5 * a;
      P.partials[item * 2 + 1] = 0.5 * c;
    }
    if (P.fb_nbuf == 1 && mtb + 1 < t_hi) {                 // single buffer: the next tile can only be staged now
      stage(0, mtb + 1);
      cp_async_commit();
    }
  }
}

// Small networks (all layers of an example tile and every W fit in shared memory together, e.g.
// the bar-image classifier [25, 30, 4]): one CTA = one tile of TMF examples walks ALL layers with
// the states, the gradient rows and the weights resident.  X is read once and G written once per
// example (no lambda buffer, no elementwise pass); Delta still goes to the global buffer of the
// split-K weight-gradient kernel.  grid (tiles, paths).
//   Xs [TMF][pxn]  all layers of the tile's examples, pxn >= NDnet, = 4 (mod 8)
//   Gs [TMF][pxn]  gradient rows: measurement terms, then per layer lambda (direct) and Delta W (back)
//   Ds [TMF][pd]   Delta of the layer in flight
//   Ws             W_n [dn1P][dnP + 4] of every layer, then the biases
__global__ void __launch_bounds__(NT, 2) nn_fba_kernel(const __grid_constant__ NnParams P) {
  extern __shared__ double sm[];
  const int mtb = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int TM = P.TMF, pxn = P.fba_pxn, pd = P.fba_pd, ND = P.NDnet;
  double* Xs = sm;
  double* Gs = Xs + TM * pxn + 16;       // 16 doubles of zero slack behind the X tile (fragment
  double* Ds = Gs + TM * pxn;            // loads of the last layer run past its width)
  double* Ws = Ds + TM * pd;
  __shared__ double red[2][NT / 32];
  __shared__ int wofs[64], bofs[64];     // shared-memory offsets of W_n / b_n (NL - 1 <= 64)
  const int m0 = mtb * TM;
  const int rows = min(TM, P.M - m0);
  const int d0 = P.structure[0];
  const int ND1 = ND - d0;
  const double* xp = P.XP + (long long)b * P.ldxp;
  double* gp = P.G + (long long)b * P.ldg;
  double* dbuf = P.dbuf + (long long)b * P.M * ND1;
  const double* pfull = P.pfull + (long long)b * P.NP;
  const double cf2 = (P.rf_path != nullptr) ? P.cf2_num * __ldg(P.rf_path + b) / P.cf2_den : P.cf2;
  double me_acc = 0.0, fe_acc = 0.0;
  if (tid == 0) {
    int o = 0;
    for (int n = 0; n + 1 < P.NL; ++n) {
      const int dnP = (P.structure[n] + 7) & ~7, dn1P = (P.structure[n + 1] + 7) & ~7;
      wofs[n] = o;
      o += dn1P * (dnP + 4);
    }
    for (int n = 0; n + 1 < P.NL; ++n) {
      bofs[n] = o;
      o += (P.structure[n + 1] + 7) & ~7;
    }
  }
  __syncthreads();
  // stage the states of the tile (rows are contiguous in X), all weights and biases
  stage_tile_async(Xs, pxn, xp, ND, m0, P.M, TM, ND, pxn);
  for (int n = 0; n + 1 < P.NL; ++n) {
    const int dn = P.structure[n], dn1 = P.structure[n + 1];
    const int dnP = (dn + 7) & ~7, dn1P = (dn1 + 7) & ~7;
    stage_tile_async(Ws + wofs[n], dnP + 4, pfull + P.woff[n], dn, 0, dn1, dn1P, dn, dnP);
    for (int j = tid; j < dn1P; j += NT) cp_async8(Ws + bofs[n] + j, pfull + P.boff[n] + (j < dn1 ? j : 0), j < dn1 ? 8 : 0);
  }
  cp_async_commit();
  for (int i = tid; i < TM * pxn; i += NT) Gs[i] = 0.0;
  if (tid < 16) Xs[TM * pxn + tid] = 0.0;
  cp_async_wait<0>();
  __syncthreads();
  // measurement terms of the input and output layer (va_nnet.py:117-173)
  {
    const int dl = P.structure[P.NL - 1], clast = ND - dl;
    for (int idx = tid; idx < rows * d0; idx += NT) {
      const int m = idx / d0, c = idx - m * d0;
      const int s = P.slot_in[c];
      if (s >= 0) {
        const double diff = Xs[m * pxn + c] - P.data_in[(long long)(m0 + m) * P.n_Lin + s];
        me_acc = fma(P.wm_in * diff, diff, me_acc);
        Gs[m * pxn + c] += P.wm_in * diff;
      }
    }
    __syncthreads();                      // a one-layer-pair network has input and output columns only
    for (int idx = tid; idx < rows * dl; idx += NT) {
      const int m = idx / dl, c = idx - m * dl;
      const int s = P.slot_out[c];
      if (s >= 0) {
        const double diff = Xs[m * pxn + clast + c] - P.data_out[(long long)(m0 + m) * P.n_Lout + s];
        me_acc = fma(P.wm_out * diff, diff, me_acc);
        Gs[m * pxn + clast + c] += P.wm_out * diff;
      }
    }
  }
  const int MTL = TM >> 3;
  for (int n = 0; n + 1 < P.NL; ++n) {
    const int dn = P.structure[n], dn1 = P.structure[n + 1];
    const int dnP = (dn + 7) & ~7, dn1P = (dn1 + 7) & ~7;
    const int pw = dnP + 4;
    const int xo = P.xoff[n], xo1 = P.xoff[n + 1];
    const double* Wn = Ws + wofs[n];
    const double* bn = Ws + bofs[n];
    __syncthreads();
    // ---- (1) Z = X_n W_n^T, epilogue: residual, lambda (into the gradient tile), Delta
    {
      const int NTL = dn1P >> 3, NG = (NTL + NTILE - 1) / NTILE;
      for (int task = warp; task < MTL * NG; task += NT / 32) {
        const int mt = task / NG, g = task - mt * NG;
        const int nt0 = g * NTILE, ntn = min(NTILE, NTL - nt0);
        double c0[NTILE], c1[NTILE];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
        const double* arow = Xs + (mt * 8 + lr) * pxn + xo + lc;
        const double* brow = Wn + (nt0 * 8 + lr) * pw + lc;
        for (int k = 0; k < dnP; k += 4) {
          const double a = arow[k];
#pragma unroll
          for (int t = 0; t < NTILE; ++t)
            if (t < ntn) dmma(c0[t], c1[t], a, brow[t * 8 * pw + k]);
        }
        const int m = mt * 8 + lr;
        const long long grow = (long long)(m0 + m);
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
          if (t >= ntn) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = (nt0 + t) * 8 + 2 * lc + h;
            const double z = (h ? c1[t] : c0[t]) + bn[j];
            double dl = 0.0;
            if (m < rows && j < dn1) {
              const double sv = act_f(P.act, z);
              const double e = Xs[m * pxn + xo1 + j] - sv;
              const double lm = cf2 * e;
              fe_acc = fma(lm, e, fe_acc);
              dl = -lm * act_d(P.act, sv);
              Gs[m * pxn + xo1 + j] += lm;
              dbuf[grow * ND1 + (xo1 - d0) + j] = dl;
            }
            Ds[m * pd + j] = dl;
          }
        }
      }
    }
    __syncthreads();
    // ---- (2) gradient rows of layer n += Delta W_n
    {
      const int NTL = dnP >> 3, NG = (NTL + NTILE - 1) / NTILE;
      for (int task = warp; task < MTL * NG; task += NT / 32) {
        const int mt = task / NG, g = task - mt * NG;
        const int nt0 = g * NTILE, ntn = min(NTILE, NTL - nt0);
        double c0[NTILE], c1[NTILE];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
        const double* arow = Ds + (mt * 8 + lr) * pd + lc;
        const double* brow = Wn + lc * pw + nt0 * 8 + lr;
        for (int j = 0; j < dn1P; j += 4) {
          const double a = arow[j];
#pragma unroll
          for (int t = 0; t < NTILE; ++t)
            if (t < ntn) dmma(c0[t], c1[t], a, brow[j * pw + t * 8]);
        }
        const int m = mt * 8 + lr;
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
          if (t >= ntn) continue;
          const int k = (nt0 + t) * 8 + 2 * lc;
          if (k < dn) Gs[m * pxn + xo + k] += c0[t];
          if (k + 1 < dn) Gs[m * pxn + xo + k + 1] += c1[t];
        }
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < rows * ND; idx += NT) {
    const int m = idx / ND, c = idx - m * ND;
    gp[(long long)(m0 + m) * ND + c] = Gs[m * pxn + c];
  }
  for (int sft = 16; sft > 0; sft >>= 1) {
    me_acc += __shfl_down_sync(0xffffffffu, me_acc, sft);
    fe_acc += __shfl_down_sync(0xffffffffu, fe_acc, sft);
  }
  if (lane == 0) { red[0][warp] = me_acc; red[1][warp] = fe_acc; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < NT / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    const long long item = (long long)b * P.nmt + mtb;
    P.partials[item * 2 + 0] = 0.5 * a;
    P.partials[item * 2 + 1] = 0.5 * c;
  }
}

// Elementwise pass over the gradient rows after nn_fb_kernel: columns of the middle layers
// += lambda of the layer below; observed components of the input / output layer += the
// measurement term (va_nnet.py:117-173), whose per-block sums are the me partials.
// tcgen05 path (VAB_NN_TCGEN05=1): the contractions Z = X_n W_n^T and Delta W_n of nn_fb_kernel run as
// Ozaki-split int8 GEMMs on the 5th-generation tensor cores (ozaki_gemm.cu: TMA tensor maps, TMEM
// accumulators); this kernel is the epilogue between the two -- the same arithmetic as the
// accumulator epilogue of nn_fb_kernel (va_nnet.py:210-255: activation, residual, lambda, Delta), one
// block per tile of TMF examples so that the fe partials keep the layout nn_reduce_kernel sums.
__global__ void __launch_bounds__(256) nn_tc_epilogue_kernel(const __grid_constant__ NnParams P, int n,
                                                             const double* __restrict__ Z) {
  const int mtb = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int dn1 = P.structure[n + 1], d0 = P.structure[0];
  const int xo1 = P.xoff[n + 1];
  const int ND1 = P.NDnet - d0;
  const bool lastl = (n + 1 == P.NL - 1);
  const double* xp = P.XP + (long long)b * P.ldxp;
  double* gp = P.G + (long long)b * P.ldg;
  double* dbuf = P.dbuf + (long long)b * P.M * ND1;
  double* lam = P.lam + (long long)b * P.M * ND1;
  const double* bias = P.pfull + (long long)b * P.NP + P.boff[n];
  const double* z = Z + (long long)b * P.M * dn1;
  const double cf2 = (P.rf_path != nullptr) ? P.cf2_num * __ldg(P.rf_path + b) / P.cf2_den : P.cf2;
  const int m0 = mtb * P.TMF, rows = min(P.TMF, P.M - m0);
  double fe_acc = 0.0;
  for (int e = tid; e < rows * dn1; e += 256) {
    const int m = e / dn1, j = e - m * dn1;
    const long long grow = (long long)(m0 + m);
    const double sv = act_f(P.act, z[grow * dn1 + j] + bias[j]);
    const double ev = xp[grow * P.NDnet + xo1 + j] - sv;
    const double lm = cf2 * ev;
    fe_acc = fma(lm, ev, fe_acc);
    if (lastl) gp[grow * P.NDnet + xo1 + j] = lm;
    else lam[grow * ND1 + (xo1 - d0) + j] = lm;
    dbuf[grow * ND1 + (xo1 - d0) + j] = -lm * act_d(P.act, sv);
  }
  __shared__ double red[8];
  for (int sft = 16; sft > 0; sft >>= 1) fe_acc += __shfl_down_sync(0xffffffffu, fe_acc, sft);
  if ((tid & 31) == 0) red[tid >> 5] = fe_acc;
  __syncthreads();
  if (tid == 0) {
    double c = 0.0;
    for (int w = 0; w < 8; ++w) c += red[w];
    const long long item = ((long long)b * (P.NL - 1) + n) * P.nmt + mtb;
    P.partials[item * 2 + 0] = 0.0;
    P.partials[item * 2 + 1] = 0.5 * c;
  }
}

// tcgen05 path: bias gradient of layer n, Gb_n[j] = sum_m Delta[m][j], per chunk of TC_KC examples (the
// chunks of the split-K weight-gradient GEMM), summed in chunk order by nn_reduce_kernel.
constexpr int TC_KC = 128;
__global__ void __launch_bounds__(128) nn_tc_bias_kernel(const __grid_constant__ NnParams P, int n) {
  const int c = blockIdx.x, b = blockIdx.y, j = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int dn1 = P.structure[n + 1], d0 = P.structure[0];
  if (j >= dn1) return;
  const int ND1 = P.NDnet - d0;
  const double* d = P.dbuf + (long long)b * P.M * ND1 + (P.xoff[n + 1] - d0) + j;
  const int m1 = min(P.M, (c + 1) * TC_KC);
  double acc = 0.0;
  for (int m = c * TC_KC; m < m1; ++m) acc += d[(long long)m * ND1];
  P.gwpart[((long long)b * P.ngw + c) * P.NP + P.boff[n] + j] = acc;
}

constexpr int FIX_NT = 256;
__global__ void __launch_bounds__(FIX_NT) nn_fix_kernel(const __grid_constant__ NnParams P) {
  const int b = blockIdx.y;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int d0 = P.structure[0], dl = P.structure[P.NL - 1];
  const int ND = P.NDnet, ND1 = ND - d0, clast = ND - dl;
  const long long total = (long long)P.M * ND;
  double* gp = P.G + (long long)b * P.ldg;
  const double* xp = P.XP + (long long)b * P.ldxp;
  const double* lam = P.lam + (long long)b * P.M * ND1;
  double me_acc = 0.0;
  // (m, c) advance by a constant stride: no division in the loop
  const long long stride = (long long)gridDim.x * FIX_NT;
  const long long dm = stride / ND;
  const int dc = (int)(stride - dm * ND);
  long long i = (long long)blockIdx.x * FIX_NT + threadIdx.x;
  long long m = i / ND;
  int c = (int)(i - m * ND);
  for (; i < total; i += stride) {
    if (c < d0) {
      const int s = P.slot_in[c];
      if (s >= 0) {
        const double diff = xp[i] - P.data_in[m * P.n_Lin + s];
        me_acc = fma(P.wm_in * diff, diff, me_acc);
        gp[i] += P.wm_in * diff;
      }
    } else if (c >= clast) {
      const int s = P.slot_out[c - clast];
      if (s >= 0) {
        const double diff = xp[i] - P.data_out[m * P.n_Lout + s];
        me_acc = fma(P.wm_out * diff, diff, me_acc);
        gp[i] += P.wm_out * diff;
      }
    } else {
      gp[i] += lam[m * ND1 + (c - d0)];
    }
    m += dm;
    c += dc;
    if (c >= ND) { c -= ND; m += 1; }
  }
  __shared__ double red[FIX_NT / 32];
  for (int sft = 16; sft > 0; sft >>= 1) me_acc += __shfl_down_sync(0xffffffffu, me_acc, sft);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = me_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < FIX_NT / 32; ++w) a += red[w];
    P.me_parts[(long long)b * gridDim.x + blockIdx.x] = 0.5 * a;
  }
}

__global__ void __launch_bounds__(NT, 2) nn_gw_kernel(const __grid_constant__ NnParams P) {
  extern __shared__ double sm[];
  const int sp = blockIdx.x / P.gw_njs, js = blockIdx.x - sp * P.gw_njs;
  const int n = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int dn = P.structure[n], dn1 = P.structure[n + 1];
  // X is augmented by a column of ones: column dn of the product is the bias gradient sum_m Delta[m][j]
  const int dnP = (dn + 1 + 7) & ~7, dn1P = (dn1 + 7) & ~7;
  const int px = dnP + 4, pd = dn1P + 4;
  const int KP = P.gw_kp;
  const int pstride = KP * (px + pd);       // doubles per panel buffer: X panel then Delta panel
  const int xo = P.xoff[n], d0 = P.structure[0];
  const int ND1 = P.NDnet - d0, co = P.xoff[n + 1] - d0;
  const double* xp = P.XP + (long long)b * P.ldxp;
  const double* dbuf = P.dbuf + (long long)b * P.M * ND1;
  const int mlo = sp * P.gw_klen, mhi = min(P.M, mlo + P.gw_klen);
  const int JTL = dn1P >> 3, NTL = dnP >> 3;
  const int jtper = (JTL + P.gw_njs - 1) / P.gw_njs;          // row tiles of W owned by this CTA
  const int jt_lo = js * jtper, jt_n = max(0, min(JTL, jt_lo + jtper) - jt_lo);
  // column tiles per task: NTILE, or fewer when that would leave warps without a task (tiny layers)
  int gsz = NTILE;
  if (jt_n * ((NTL + NTILE - 1) / NTILE) < NT / 32) gsz = max(1, (jt_n * NTL) / (NT / 32));
  while (gsz < NTILE && jt_n * ((NTL + gsz - 1) / gsz) > GW_MAXTASK * (NT / 32)) ++gsz;
  const int NG = (NTL + gsz - 1) / gsz;
  const int ntask = jt_n * NG;
  double c0[GW_MAXTASK][NTILE], c1[GW_MAXTASK][NTILE];
#pragma unroll
  for (int q = 0; q < GW_MAXTASK; ++q)
#pragma unroll
    for (int t = 0; t < NTILE; ++t) { c0[q][t] = 0.0; c1[q][t] = 0.0; }
  auto load_panel = [&](int buf, int mp) {
    double* Xp = sm + buf * pstride;
    stage_tile_async(Xp, px, xp + xo, P.NDnet, mp, mhi, KP, dn, dnP, P.one);
    stage_tile_async(Xp + KP * px, pd, dbuf + co, ND1, mp, mhi, KP, dn1, dn1P);
    cp_async_commit();
  };
  const int npanel = (mhi - mlo + KP - 1) / KP;
  load_panel(0, mlo);
  for (int pi = 0; pi < npanel; ++pi) {
    if (pi + 1 < npanel) {
      load_panel((pi + 1) & 1, mlo + (pi + 1) * KP);          // next panel streams in during the MMAs
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Xp = sm + (pi & 1) * pstride;
    const double* Dp = Xp + KP * px;
#pragma unroll
    for (int q = 0; q < GW_MAXTASK; ++q) {
      const int task = warp + q * (NT / 32);
      if (task < ntask) {
        const int jt = jt_lo + task / NG, g = task % NG;
        const int nt0 = g * gsz, ntn = min(gsz, NTL - nt0);
        const double* arow = Dp + lc * pd + jt * 8 + lr;            // A[j][m] = Delta[m][j]
        const double* brow = Xp + lc * px + nt0 * 8 + lr;           // B[m][k] = X[m][k]
#pragma unroll 2
        for (int mm = 0; mm < KP; mm += 4) {
          const double a = arow[mm * pd];
#pragma unroll
          for (int t = 0; t < NTILE; ++t)
            if (t < ntn) dmma(c0[q][t], c1[q][t], a, brow[mm * px + t * 8]);
        }
      }
    }
    __syncthreads();                                              // buffer (pi & 1) is refilled next
  }
  double* gw = P.gwpart + ((long long)b * P.gw_nsplit + sp) * P.NP;
#pragma unroll
  for (int q = 0; q < GW_MAXTASK; ++q) {
    const int task = warp + q * (NT / 32);
    if (task < ntask) {
      const int jt = jt_lo + task / NG, g = task % NG;
      const int nt0 = g * gsz, ntn = min(gsz, NTL - nt0);
      const int j = jt * 8 + lr;
      if (j < dn1) {
        double* grow = gw + P.woff[n] + (long long)j * dn;
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
          if (t >= ntn) continue;
          const int k = (nt0 + t) * 8 + 2 * lc;
          if (k < dn) grow[k] = c0[q][t];
          else if (k == dn) gw[P.boff[n] + j] = c0[q][t];
          if (k + 1 < dn) grow[k + 1] = c1[q][t];
          else if (k + 1 == dn) gw[P.boff[n] + j] = c1[q][t];
        }
      }
    }
  }
}

// sums the per-tile partials in tile order: A / me / fe and the gradient of the estimated params
__global__ void nn_reduce_kernel(const __grid_constant__ NnParams P, double* A, double* me, double* fe) {
  const int b = blockIdx.y;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < P.NP && P.G != nullptr) {
    const int e = P.pmap[k];
    if (e >= 0) {
      double acc = 0.0;
      const double* src = P.gwpart + (long long)b * P.ngw * P.NP + k;
      for (int t = 0; t < P.ngw; ++t) acc += src[(long long)t * P.NP];
      P.G[(long long)b * P.ldg + P.NDens + e] = acc;
    }
  }
  if (k == 0) {
    double m = 0.0, f = 0.0;
    for (int t = 0; t < P.nparts; ++t) {
      m += P.partials[((long long)b * P.nparts + t) * 2 + 0];
      f += P.partials[((long long)b * P.nparts + t) * 2 + 1];
    }
    if (P.me_parts != nullptr)
      for (int t = 0; t < P.n_me; ++t) m += P.me_parts[(long long)b * P.n_me + t];
    if (me) me[b] = m;
    if (fe) fe[b] = f;
    if (A) A[b] = m + f;
  }
}

}  // namespace
#include "nn_small.cuh"

struct NnProblem {
  int NL = 0, M = 0, NDnet = 0, NP = 0, NPest = 0, act = 0, n_Lin = 0, n_Lout = 0, dmax = 0, d0 = 0;
  long long NDens = 0;
  int Ltot = 0;
  int* ints = nullptr;            // structure | xoff | woff | boff | pmap | slot_in | slot_out
  int *structure = nullptr, *xoff = nullptr, *woff = nullptr, *boff = nullptr, *pmap = nullptr,
      *slot_in = nullptr, *slot_out = nullptr;
  const double* data_in = nullptr;
  const double* data_out = nullptr;
  double rm_in = 1.0, rm_out = 1.0, rf0 = 1.0;
  const double* rm_in_mat = nullptr;   // (n_Lin, n_Lin) / (n_Lout, n_Lout) matrix form of RM (vab_nn_set_rm_matrices),
  const double* rm_out_mat = nullptr;  // caller-owned; rm_in = rm_out = 0 while set
  const double* pfix = nullptr;
  long long pfix_stride = 0;
  double* pfix_zero = nullptr;
  double* gwpart = nullptr;
  size_t gwpart_cap = 0;
  double* pfull = nullptr;
  size_t pfull_cap = 0;
  double* dbuf = nullptr;          // split design: Delta and lambda buffers (B, M, NDnet - d_0)
  size_t dbuf_cap = 0;
  double* lam = nullptr;
  size_t lam_cap = 0;
  double* zbuf = nullptr;          // tcgen05 path: Z = X_n W_n^T of the layer in flight (B, M, d_{n+1})
  size_t zbuf_cap = 0;
  double* one_dev = nullptr;       // device constant 1.0
  std::vector<int> st_host;        // layer widths (host copy, for launch planning)
};

void nn_destroy(vab_ctx* ctx) {
  NnProblem* p = ctx->nn;
  if (!p) return;
  cudaFree(p->ints);
  cudaFree(p->pfix_zero);
  cudaFree(p->gwpart);
  cudaFree(p->pfull);
  cudaFree(p->dbuf);
  cudaFree(p->lam);
  cudaFree(p->zbuf);
  cudaFree(p->one_dev);
  delete p;
  ctx->nn = nullptr;
}

long long nn_unknowns(const vab_ctx* ctx) { return ctx->nn ? ctx->nn->NDens + ctx->nn->NPest : 0; }

// Matrix RM (va_nnet.py:135-139): me = cm sum_m [din_m . (RM_in din_m) + dout_m . (RM_out dout_m)], added
// behind the action kernels (which then run with zero measurement weights): one CTA per path, thread
// per (example, measured component), fixed-order reduction.  cin / cout: state column of each measured
// input / output component, rebuilt from the slot tables.
__global__ void __launch_bounds__(256) nn_me_matrix_kernel(const __grid_constant__ NnParams P, const double* rm_in,
                                                           const double* rm_out, double cm, double* A, double* me) {
  extern __shared__ int cols[];
  const int b = blockIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  const int d0 = P.structure[0], dl = P.structure[P.NL - 1], ND = P.NDnet, clast = ND - dl;
  int* cin = cols;
  int* cout = cols + P.n_Lin;
  for (int c = threadIdx.x; c < d0; c += 256) { const int sl = P.slot_in[c]; if (sl >= 0) cin[sl] = c; }
  for (int c = threadIdx.x; c < dl; c += 256) { const int sl = P.slot_out[c]; if (sl >= 0) cout[sl] = clast + c; }
  __syncthreads();
  const double* x = P.XP + (long long)b * P.ldxp;
  double* g = P.G ? P.G + (long long)b * P.ldg : nullptr;
  const int Lt = P.n_Lin + P.n_Lout;
  double acc = 0.0;
  for (long long t = threadIdx.x; t < (long long)P.M * Lt; t += 256) {
    const long long m = t / Lt;
    int l = (int)(t - m * Lt);
    const bool in = l < P.n_Lin;
    if (!in) l -= P.n_Lin;
    const int L = in ? P.n_Lin : P.n_Lout;
    const int* col = in ? cin : cout;
    const double* dat = in ? P.data_in + m * P.n_Lin : P.data_out + m * P.n_Lout;
    const double* R = in ? rm_in : rm_out;
    const long long base = m * ND;
    double w = 0.0, wt = 0.0;
    for (int k = 0; k < L; ++k) {
      const double dk = x[base + col[k]] - dat[k];
      w = fma(R[(long long)l * L + k], dk, w);
      wt = fma(R[(long long)k * L + l], dk, wt);
    }
    const double dv = x[base + col[l]] - dat[l];
    acc = fma(dv, w, acc);
    if (g) g[base + col[l]] += cm * (w + wt);
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

static int nn_eval_core(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
                        const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
                        double* G, long long ldg);

int nn_eval(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
            const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
            double* G, long long ldg) {
  int rc = nn_eval_core(ctx, B, XP, ldxp, rf_scale, rf_path_dev, active_dev, A, me, fe, G, ldg);
  if (rc != VAB_OK) return rc;
  NnProblem* p = ctx->nn;
  if (p->rm_in_mat != nullptr) {
    NnParams P;
    memset(&P, 0, sizeof(P));
    P.XP = XP; P.ldxp = ldxp; P.G = G; P.ldg = ldg; P.B = B; P.M = p->M; P.NL = p->NL; P.NDnet = p->NDnet;
    P.structure = p->structure; P.slot_in = p->slot_in; P.slot_out = p->slot_out; P.n_Lin = p->n_Lin; P.n_Lout = p->n_Lout;
    P.data_in = p->data_in; P.data_out = p->data_out; P.active = active_dev;
    const double cm = p->Ltot > 0 ? 1.0 / ((double)p->Ltot * p->M) : 0.0;
    nn_me_matrix_kernel<<<B, 256, (size_t)(p->n_Lin + p->n_Lout + 1) * sizeof(int), ctx->stream>>>(P, p->rm_in_mat, p->rm_out_mat, cm, A, me);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_me_matrix_kernel launch");
    ctx->launches += 1;
  }
  return VAB_OK;
}

static int nn_eval_core(vab_ctx* ctx, int B, const double* XP, long long ldxp, double rf_scale,
                        const double* rf_path_dev, const int* active_dev, double* A, double* me, double* fe,
                        double* G, long long ldg) {
  NnProblem* p = ctx->nn;
  if (!p) return vab_fail(ctx, VAB_ERR_STATE, "nn_action_grad: no NN problem set");
  const long long n = p->NDens + p->NPest;
  if (B < 1 || !XP || ldxp < n) return vab_fail(ctx, VAB_ERR_INVALID, "nn_action_grad: bad batch / XP / ldxp");
  if (G && ldg < n) return vab_fail(ctx, VAB_ERR_INVALID, "nn_action_grad: ldg < n");
  NnParams P;
  memset(&P, 0, sizeof(P));
  P.XP = XP; P.ldxp = ldxp; P.G = G; P.ldg = ldg;
  P.B = B; P.M = p->M; P.NL = p->NL; P.NDnet = p->NDnet; P.NP = p->NP; P.NPest = p->NPest;
  P.act = p->act; P.NDens = p->NDens;
  P.structure = p->structure; P.xoff = p->xoff; P.woff = p->woff; P.boff = p->boff; P.pmap = p->pmap;
  P.pfix = p->pfix; P.pfix_stride = p->pfix_stride;
  P.slot_in = p->slot_in; P.slot_out = p->slot_out; P.n_Lin = p->n_Lin; P.n_Lout = p->n_Lout;
  P.data_in = p->data_in; P.data_out = p->data_out;
  const double cm = p->Ltot > 0 ? 1.0 / ((double)p->Ltot * p->M) : 0.0;
  P.wm_in = 2.0 * cm * p->rm_in;
  P.wm_out = 2.0 * cm * p->rm_out;
  P.cf2 = 2.0 * p->rf0 * rf_scale / ((double)(p->NDnet - p->d0) * p->M);
  P.cf2_num = 2.0 * p->rf0;
  P.cf2_den = (double)(p->NDnet - p->d0) * p->M;
  P.rf_path = rf_path_dev;
  P.active = active_dev;
  // ---- small networks (bar-image class): register-tiled CUDA-core kernel, nn_small.cuh
  {
    // Opt-in (VAB_NN_SMALL=1): measured on B200 at the bar-image shape (M = 10 000, 32 paths) it only ties
    // the tensor-pipe tile kernels -- 0.56 ms vs 0.55 ms per evaluation of the batch -- because half of its
    // shared-memory wavefronts are bank conflicts on the transposed tiles (profiles/r02_nn_small_ncu.json,
    // DESIGN.md section 4.4); it stays in the tree as the starting point for that work and is covered by
    // test_nn_split_kernels_agree_with_fused_kernel.
    const char* env_small = getenv("VAB_NN_SMALL");
    NnSmallPlan L;
    size_t smem_small = 0;
    if (env_small && atoi(env_small) != 0 && nn_small_plan(p->st_host, p->M, B, ctx->num_sms, ldxp, XP, &L, &smem_small)) {
      int rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)B * L.ncta * 2);
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->gwpart, &p->gwpart_cap, (size_t)B * L.ncta * (p->NP > 0 ? p->NP : 1));
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->pfull, &p->pfull_cap, (size_t)B * (p->NP > 0 ? p->NP : 1));
      if (rc != VAB_OK) return rc;
      P.partials = ctx->partials; P.gwpart = p->gwpart; P.pfull = p->pfull;
      P.nparts = L.ncta; P.ngw = L.ncta;
      if (!ctx->attr_nn_small) {
        cudaError_t e = cudaFuncSetAttribute(nn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_action_grad smem opt-in (small)");
        ctx->attr_nn_small = true;
      }
      if (p->NP > 0) nn_gather_params_kernel<<<dim3((p->NP + 255) / 256, B), 256, 0, ctx->stream>>>(P);
      ctx->nn_family = 4;
      nn_small_kernel<<<dim3(L.ncta, B), SM_NT, smem_small, ctx->stream>>>(P, L);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_small_kernel launch");
      const int nk = p->NP > 1 ? p->NP : 1;
      nn_reduce_kernel<<<dim3((nk + 255) / 256, B), 256, 0, ctx->stream>>>(P, A, me, fe);
      e = cudaGetLastError();
      if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_reduce_kernel launch");
      ctx->launches += 3;
      return VAB_OK;
    }
  }
  // ---- split design when every layer's W fits next to the tiles (and a gradient is wanted)
  {
    const char* env_split = getenv("VAB_NN_SPLIT");            // 0: always the fused kernel
    const bool use_split = env_split ? (atoi(env_split) != 0) : true;
    const size_t budget = 200 * 1024;
    int TMF = 0;
    size_t smem_fb = 0, smem_gw = 0;
    bool fits = use_split && G != nullptr && p->NL >= 2;
    // VAB_NN_TCGEN05=1: the forward and backward-to-states contractions on tcgen05 / TMEM / TMA
    // (Ozaki-split, ozaki_gemm.cu) instead of the fp64 tensor pipe; layers up to 128 wide
    bool use_tc = false;
    if (const char* env_tc = getenv("VAB_NN_TCGEN05")) {
      use_tc = fits && atoi(env_tc) != 0;
      for (int n = 0; n < p->NL && use_tc; ++n)
        if (p->st_host[n] > 128) use_tc = false;
    }
    // small networks: all layers of an example tile + every W resident (nn_fba_kernel)
    bool all_layers = false;
    int fba_pxn = 0, fba_pd = 0;
    size_t smem_fba = 0;
    if (fits && p->NL - 1 <= 64 && !use_tc) {
      const char* env_fba = getenv("VAB_NN_ALL_LAYERS");      // 0: per-layer kernels even for small networks
      if (!(env_fba && atoi(env_fba) == 0)) {
        fba_pxn = ((p->NDnet + 7) & ~7) + 4;
        size_t wsz = 0;
        int d1max = 0;
        for (int n = 0; n + 1 < p->NL; ++n) {
          const int dnP = (p->st_host[n] + 7) & ~7, dn1P = (p->st_host[n + 1] + 7) & ~7;
          wsz += (size_t)dn1P * (dnP + 4) + dn1P;
          if (dn1P > d1max) d1max = dn1P;
        }
        fba_pd = d1max + 4;
        for (int tm = 128; tm >= 32 && !all_layers; tm >>= 1) {
          const size_t s = ((size_t)2 * tm * fba_pxn + 16 + (size_t)tm * fba_pd + wsz) * sizeof(double);
          if (s <= (size_t)100 * 1024 && (tm <= 32 || tm * 4 <= p->M)) { all_layers = true; TMF = tm; smem_fba = s; }
        }
      }
    }
    int fb_T = 1, fb_nbuf = 1;
    if (fits && !all_layers) {
      auto fb_smem = [&](int tm, int nbuf) {
        size_t worst = 0;
        for (int n = 0; n + 1 < p->NL; ++n) {
          const int dnP = (p->st_host[n] + 7) & ~7, dn1P = (p->st_host[n + 1] + 7) & ~7;
          const size_t s = ((size_t)nbuf * tm * ((dnP + 4) + (dn1P + 4)) + (size_t)dn1P * (dnP + 4) + dn1P) * sizeof(double);
          if (s > worst) worst = s;
        }
        return worst;
      };
      // first choice: W resident over several tiles of 64 examples with double-buffered staging.
      // (Measured on B200: with 32-example tiles -- all that fits next to a 100-wide W -- the 8-warp
      // CTA is slower than one 64-example tile on 16 warps, C4 1.36 vs 1.29 ms; small layers gain,
      // 20 x 10: 66k -> 73k evals/s.)
      const char* env_db = getenv("VAB_NN_FB_DB");              // 0: one tile per CTA, single buffer
      if (!(env_db && atoi(env_db) == 0)) {
        const char* env_tm = getenv("VAB_NN_FB_TM32");          // 1: also try 32-example tiles (A/B knob)
        const int tm_min = (env_tm && atoi(env_tm) != 0) ? 32 : 64;
        for (int tm = 64; tm >= tm_min && TMF == 0; tm >>= 1) {
          const int nmt = (p->M + tm - 1) / tm;
          if (nmt >= 8 && fb_smem(tm, 2) <= budget) {
            TMF = tm; smem_fb = fb_smem(tm, 2); fb_nbuf = 2;
            fb_T = nmt >= 32 ? 8 : 4;
          }
        }
      }
      for (int tm = 256; tm >= 16 && TMF == 0; tm >>= 1) {
        if (tm > 64 && tm * 4 > p->M) continue;                 // keep at least a few tiles per layer
        const size_t worst = fb_smem(tm, 1);
        // big tiles only while two CTAs still share an SM
        if (worst <= (tm > 64 ? (size_t)100 * 1024 : budget)) { TMF = tm; smem_fb = worst; }
      }
      if (TMF == 0) fits = false;
    }
    int njs = 1, kp = 128;
    if (fits) {
      for (int n = 0; n + 1 < p->NL; ++n) {
        const int dnP = (p->st_host[n] + 1 + 7) & ~7, dn1P = (p->st_host[n + 1] + 7) & ~7;   // X carries a ones column
        while (kp > 32 && (size_t)2 * kp * ((dnP + 4) + (dn1P + 4)) * sizeof(double) > (size_t)100 * 1024) kp >>= 1;
      }
      for (int n = 0; n + 1 < p->NL; ++n) {
        const int dnP = (p->st_host[n] + 1 + 7) & ~7, dn1P = (p->st_host[n + 1] + 7) & ~7;
        const int NG = ((dnP >> 3) + NTILE - 1) / NTILE;
        if (NG > GW_MAXTASK * (NT / 32)) fits = false;
        // row tiles per CTA so that a warp owns at most GW_MAXTASK (row tile, column group) tasks
        const int jtper = (GW_MAXTASK * (NT / 32)) / NG;
        if (jtper >= 1) {
          const int need = ((dn1P >> 3) + jtper - 1) / jtper;
          if (need > njs) njs = need;
        }
        const size_t s = (size_t)2 * kp * ((dnP + 4) + (dn1P + 4)) * sizeof(double);   // two panel buffers
        if (s > smem_gw) smem_gw = s;
      }
    }
    if (fits) {
      P.TMF = TMF;
      P.nmt = (p->M + TMF - 1) / TMF;
      P.gw_njs = njs;
      P.gw_kp = kp;
      P.one = p->one_dev;
      const int per = (p->NL - 1) * B * njs;
      int want = (8 * ctx->num_sms + per - 1) / per;               // ~8 CTAs per SM over the batch
      if (want < 1) want = 1;
      int klen = (p->M + want - 1) / want;
      klen = ((klen + kp - 1) / kp) * kp;
      P.gw_klen = klen;
      P.gw_nsplit = (p->M + klen - 1) / klen;
      P.nparts = all_layers ? P.nmt : (p->NL - 1) * P.nmt;
      P.ngw = use_tc ? (p->M + TC_KC - 1) / TC_KC : P.gw_nsplit;
      P.fba_pxn = fba_pxn; P.fba_pd = fba_pd;
      P.fb_T = fb_T; P.fb_nbuf = fb_nbuf;
      const size_t nd1 = (size_t)(p->NDnet - p->d0);
      long long nfix = ((long long)p->M * p->NDnet + FIX_NT * 8 - 1) / (FIX_NT * 8);     // ~8 entries per thread
      if (nfix > 1024) nfix = 1024;
      if (nfix < 1) nfix = 1;
      P.n_me = (int)nfix;
      int rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)B * P.nparts * 2 + (size_t)B * nfix);
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->gwpart, &p->gwpart_cap, (size_t)B * P.ngw * (p->NP > 0 ? p->NP : 1));
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->pfull, &p->pfull_cap, (size_t)B * (p->NP > 0 ? p->NP : 1));
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->dbuf, &p->dbuf_cap, (size_t)B * p->M * nd1);
      if (rc != VAB_OK) return rc;
      rc = vab_reserve(ctx, &p->lam, &p->lam_cap, (size_t)B * p->M * nd1);
      if (rc != VAB_OK) return rc;
      P.partials = ctx->partials; P.gwpart = p->gwpart; P.pfull = p->pfull; P.dbuf = p->dbuf; P.lam = p->lam;
      P.me_parts = all_layers ? nullptr : ctx->partials + (size_t)B * P.nparts * 2;
      if (!ctx->attr_nn_split) {
        cudaError_t e = cudaFuncSetAttribute(nn_fb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nn_gw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nn_fba_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_action_grad smem opt-in (split)");
        ctx->attr_nn_split = true;
      }
      if (smem_gw > 200 * 1024) return vab_fail(ctx, VAB_ERR_INVALID, "nn_action_grad: internal plan error (gw smem)");
      if (p->NP > 0) nn_gather_params_kernel<<<dim3((p->NP + 255) / 256, B), 256, 0, ctx->stream>>>(P);
      int nl = 3;
      cudaError_t e = cudaSuccess;
      ctx->nn_family = all_layers ? 3 : (use_tc ? 5 : 2);
      if (all_layers) {
        nn_fba_kernel<<<dim3(P.nmt, B), NT, smem_fba, ctx->stream>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_fba_kernel launch");
        nl = 2;
      } else if (use_tc) {
        int d1max = 0;
        for (int n = 1; n < p->NL; ++n) d1max = p->st_host[n] > d1max ? p->st_host[n] : d1max;
        rc = vab_reserve(ctx, &p->zbuf, &p->zbuf_cap, (size_t)B * p->M * d1max);
        if (rc != VAB_OK) return rc;
        const int d0 = p->d0;
        const long long ND1 = p->NDnet - d0;
        int xo = 0, wo = 0;
        for (int n = 0; n + 1 < p->NL; ++n) {
          const int dn = p->st_host[n], dn1 = p->st_host[n + 1];
          const int xo1 = xo + dn;
          // Z[b] = X_n[b] W_n[b]^T
          rc = ozaki_gemm(ctx, B, p->M, dn1, dn, XP + xo, p->NDnet, 1, ldxp, p->pfull + wo, dn, 1, p->NP,
                          p->zbuf, dn1, (long long)p->M * dn1, active_dev);
          if (rc != VAB_OK) return rc;
          nn_tc_epilogue_kernel<<<dim3(P.nmt, B), 256, 0, ctx->stream>>>(P, n, p->zbuf);
          e = cudaGetLastError();
          if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_tc_epilogue_kernel launch");
          // back term of layer n's gradient rows: G[b][:, xo : xo + dn] = Delta[b] W_n[b]
          rc = ozaki_gemm(ctx, B, p->M, dn, dn1, p->dbuf + (xo1 - d0), ND1, 1, (long long)p->M * ND1,
                          p->pfull + wo, 1, dn, p->NP, G + xo, p->NDnet, ldg, active_dev);
          if (rc != VAB_OK) return rc;
          ctx->launches += 1;
          xo = xo1;
          wo += dn * dn1 + dn1;
        }
        nn_fix_kernel<<<dim3((unsigned)nfix, B), FIX_NT, 0, ctx->stream>>>(P);
        // weight gradients GW_n = Delta_n^T X_n: split-K over chunks of TC_KC examples, one tcgen05
        // problem per (path, chunk), partial blocks to gwpart (summed in chunk order by nn_reduce_kernel)
        xo = 0; wo = 0;
        for (int n = 0; n + 1 < p->NL; ++n) {
          const int dn = p->st_host[n], dn1 = p->st_host[n + 1];
          const int xo1 = xo + dn;
          rc = ozaki_gemm_chunked(ctx, B, P.ngw, dn1, dn, TC_KC, p->M,
                                  p->dbuf + (xo1 - d0), 1, ND1, (long long)p->M * ND1, (long long)TC_KC * ND1,
                                  XP + xo, 1, p->NDnet, ldxp, (long long)TC_KC * p->NDnet,
                                  p->gwpart + wo, dn, p->NP, active_dev);
          if (rc != VAB_OK) return rc;
          nn_tc_bias_kernel<<<dim3(P.ngw, B), 128, 0, ctx->stream>>>(P, n);
          ctx->launches += 1;
          xo = xo1;
          wo += dn * dn1 + dn1;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn tcgen05 weight-gradient launch");
        const int nk = p->NP > 1 ? p->NP : 1;
        nn_reduce_kernel<<<dim3((nk + 255) / 256, B), 256, 0, ctx->stream>>>(P, A, me, fe);
        e = cudaGetLastError();
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_reduce_kernel launch");
        ctx->launches += 4;
        return VAB_OK;
      } else {
        // 16 warps when every warp still gets a (row tile, column group) task of the smallest layer pair
        int min_tasks = 1 << 30;
        for (int n = 0; n + 1 < p->NL; ++n) {
          const int a = (((p->st_host[n] + 7) >> 3) + NTILE - 1) / NTILE, c = (((p->st_host[n + 1] + 7) >> 3) + NTILE - 1) / NTILE;
          const int t = (TMF >> 3) * (a < c ? a : c);
          if (t < min_tasks) min_tasks = t;
        }
        // (with fewer coarse tasks the kernel splits the column groups, so 16 warps still pay off
        // as long as there are enough 8x8 output tiles: VAB_NN_FB_THREADS forces 256 / 512)
        int nthr = (min_tasks >= 16) ? NTF : NT;
        if (const char* et = getenv("VAB_NN_FB_THREADS")) nthr = atoi(et) == 512 ? NTF : NT;
        nn_fb_kernel<<<dim3((P.nmt + fb_T - 1) / fb_T, p->NL - 1, B), nthr, smem_fb, ctx->stream>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_fb_kernel launch");
        nn_fix_kernel<<<dim3((unsigned)nfix, B), FIX_NT, 0, ctx->stream>>>(P);
      }
      nn_gw_kernel<<<dim3(P.gw_nsplit * njs, p->NL - 1, B), NT, smem_gw, ctx->stream>>>(P);
      e = cudaGetLastError();
      if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_gw_kernel launch");
      const int nk = p->NP > 1 ? p->NP : 1;
      nn_reduce_kernel<<<dim3((nk + 255) / 256, B), 256, 0, ctx->stream>>>(P, A, me, fe);
      e = cudaGetLastError();
      if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_reduce_kernel launch");
      ctx->launches += nl + 2;
      return VAB_OK;
    }
  }
  // Shared-memory plan: 5 [TM][dpitch] tiles + a chunk of weight rows.  Pitches are = 4 (mod 8)
  // doubles; TM is a multiple of 8.  Prefer a footprint that lets two CTAs share an SM.
  const int dP = (p->dmax + 7) & ~7;
  const int dpitch = dP + 4;
  if (dpitch * 8 > 8192) return vab_fail(ctx, VAB_ERR_INVALID, "nn_action_grad: layers wider than 1016 neurons are not supported");
  const int Mp = (p->M + 7) & ~7;
  const size_t budget2 = 110 * 1024, budget1 = 220 * 1024;
  const size_t extra = (size_t)(dP + 8) * sizeof(double);          // bias chunk
  int TM = 0, wcap = 0;
  size_t smem = 0;
  // first choice: two CTAs per SM with at least min(dP, 32) weight rows per chunk
  for (int pass = 0; pass < 2 && TM == 0; ++pass) {
    const size_t budget = pass == 0 ? budget2 : budget1;
    const int jmin = pass == 0 ? (dP < 32 ? dP : 32) : 8;
    for (int tm = (Mp < 32 ? Mp : 32); tm >= 8; tm -= 8) {
      const size_t tiles = (size_t)5 * tm * dpitch * sizeof(double);
      if (tiles + extra + (size_t)jmin * dpitch * sizeof(double) > budget) continue;
      size_t w = (budget - tiles - extra) / sizeof(double);
      if (w > (size_t)dP * dpitch) w = (size_t)dP * dpitch;         // the whole widest layer fits
      TM = tm;
      wcap = (int)w;
      smem = tiles + extra + (size_t)wcap * sizeof(double);
      break;
    }
  }
  if (TM == 0) return vab_fail(ctx, VAB_ERR_INVALID, "nn_action_grad: layer too wide for the shared-memory tiles");
  P.TM = TM;
  P.dpitch = dpitch;
  P.wcap = wcap;
  P.ntiles = (p->M + TM - 1) / TM;
  int rc = vab_reserve(ctx, &ctx->partials, &ctx->partials_cap, (size_t)B * P.ntiles * 2);
  if (rc != VAB_OK) return rc;
  rc = vab_reserve(ctx, &p->gwpart, &p->gwpart_cap, (size_t)B * P.ntiles * p->NP);
  if (rc != VAB_OK) return rc;
  rc = vab_reserve(ctx, &p->pfull, &p->pfull_cap, (size_t)B * (p->NP > 0 ? p->NP : 1));
  if (rc != VAB_OK) return rc;
  P.partials = ctx->partials;
  P.gwpart = p->gwpart;
  P.pfull = p->pfull;
  P.nparts = P.ntiles;
  P.ngw = P.ntiles;
  if (p->NP > 0) nn_gather_params_kernel<<<dim3((p->NP + 255) / 256, B), 256, 0, ctx->stream>>>(P);
  if (!ctx->attr_nn_fused) {
    cudaError_t e = cudaFuncSetAttribute(nn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_action_grad smem opt-in");
    ctx->attr_nn_fused = true;
  }
  ctx->nn_family = 1;
  nn_fused_kernel<<<dim3(P.ntiles, B), NT, smem, ctx->stream>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_fused_kernel launch");
  const int nk = p->NP > 1 ? p->NP : 1;
  nn_reduce_kernel<<<dim3((nk + 255) / 256, B), 256, 0, ctx->stream>>>(P, A, me, fe);
  e = cudaGetLastError();
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_reduce_kernel launch");
  ctx->launches += 3;
  return VAB_OK;
}

extern "C" {

int vab_nn_problem_set(vab_ctx* ctx, int32_t n_layers, const int32_t* structure_host, int32_t M,
                       int32_t activation, int32_t n_Lin, const int32_t* Lin_host, int32_t n_Lout,
                       const int32_t* Lout_host, const double* data_in_dev, const double* data_out_dev,
                       int32_t NPest, const int32_t* Pidx_host) {
  if (!ctx) return VAB_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (n_layers < 2 || !structure_host || M < 1) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: bad structure / M");
  if (activation < VAB_ACT_SIGMOID || activation > VAB_ACT_LINEAR)
    return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: unknown activation");
  if ((n_Lin > 0 && (!Lin_host || !data_in_dev)) || (n_Lout > 0 && (!Lout_host || !data_out_dev)))
    return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: Lidx / data missing");
  cudaStreamSynchronize(ctx->stream);
  // the old problem goes first; until the new one is complete the context holds no NN problem, so a
  // failure below leaves it in a state where every later call answers VAB_ERR_STATE
  if (ctx->problem == VAB_PROBLEM_NN) ctx->problem = VAB_PROBLEM_NONE;
  nn_destroy(ctx);
  NnProblem* p = new NnProblem();
  ctx->nn = p;
  struct Guard {                       // a half-built problem never survives a failed call
    vab_ctx* c; bool ok = false;
    ~Guard() { if (!ok) nn_destroy(c); }
  } guard{ctx};
  p->NL = n_layers; p->M = M; p->act = activation; p->n_Lin = n_Lin; p->n_Lout = n_Lout;
  p->Ltot = n_Lin + n_Lout;
  std::vector<int> st(structure_host, structure_host + n_layers), xoff(n_layers + 1, 0), woff(n_layers - 1),
      boff(n_layers - 1);
  int np = 0;
  for (int n = 0; n < n_layers; ++n) {
    if (st[n] < 1) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: empty layer");
    xoff[n + 1] = xoff[n] + st[n];
    if (st[n] > p->dmax) p->dmax = st[n];
  }
  for (int n = 0; n + 1 < n_layers; ++n) {
    woff[n] = np; np += st[n] * st[n + 1];
    boff[n] = np; np += st[n + 1];
  }
  p->st_host = st;
  p->NDnet = xoff[n_layers]; p->NDens = (long long)p->NDnet * M; p->NP = np; p->d0 = st[0];
  if (NPest < 0 || NPest > np) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: bad NPest");
  p->NPest = NPest;
  std::vector<int> pmap(np > 0 ? np : 1, -1), sin(st[0], -1), sout(st[n_layers - 1], -1);
  for (int e = 0; e < NPest; ++e) {
    const int k = Pidx_host[e];
    if (k < 0 || k >= np || pmap[k] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: bad Pidx");
    pmap[k] = e;
  }
  for (int l = 0; l < n_Lin; ++l) {
    const int i = Lin_host[l];
    if (i < 0 || i >= st[0] || sin[i] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: bad Lidx[0]");
    sin[i] = l;
  }
  for (int l = 0; l < n_Lout; ++l) {
    const int i = Lout_host[l];
    if (i < 0 || i >= st[n_layers - 1] || sout[i] >= 0) return vab_fail(ctx, VAB_ERR_INVALID, "nn_problem_set: bad Lidx[1]");
    sout[i] = l;
  }
  std::vector<int> all;
  auto push = [&](const std::vector<int>& v) { size_t o = all.size(); all.insert(all.end(), v.begin(), v.end()); return o; };
  const size_t o_st = push(st), o_x = push(xoff), o_w = push(woff), o_b = push(boff), o_p = push(pmap),
               o_si = push(sin), o_so = push(sout);
  cudaError_t e = cudaMalloc((void**)&p->ints, all.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(p->ints, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->pfix_zero, (size_t)(np > 0 ? np : 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(p->pfix_zero, 0, (size_t)(np > 0 ? np : 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->one_dev, sizeof(double));
  if (e == cudaSuccess) { const double one = 1.0; e = cudaMemcpy(p->one_dev, &one, sizeof(double), cudaMemcpyHostToDevice); }
  if (e != cudaSuccess) return vab_cuda_fail(ctx, e, "nn_problem_set");
  p->structure = p->ints + o_st; p->xoff = p->ints + o_x; p->woff = p->ints + o_w; p->boff = p->ints + o_b;
  p->pmap = p->ints + o_p; p->slot_in = p->ints + o_si; p->slot_out = p->ints + o_so;
  p->data_in = data_in_dev; p->data_out = data_out_dev;
  p->pfix = p->pfix_zero; p->pfix_stride = 0;
  guard.ok = true;
  ctx->problem = VAB_PROBLEM_NN;
  return VAB_OK;
}

int vab_nn_set_weights(vab_ctx* ctx, double rm_in, double rm_out, double rf0) {
  if (!ctx) return VAB_ERR_INVALID;
  if (!ctx->nn) return vab_fail(ctx, VAB_ERR_STATE, "nn_set_weights: no NN problem set");
  ctx->nn->rm_in = rm_in; ctx->nn->rm_out = rm_out; ctx->nn->rf0 = rf0;
  ctx->nn->rm_in_mat = nullptr; ctx->nn->rm_out_mat = nullptr;
  return VAB_OK;
}

int vab_nn_set_rm_matrices(vab_ctx* ctx, const double* rm_in_dev, const double* rm_out_dev) {
  if (!ctx) return VAB_ERR_INVALID;
  if (!ctx->nn) return vab_fail(ctx, VAB_ERR_STATE, "nn_set_rm_matrices: no NN problem set");
  if (!rm_in_dev || !rm_out_dev) return vab_fail(ctx, VAB_ERR_INVALID, "nn_set_rm_matrices: NULL");
  ctx->nn->rm_in = 0.0; ctx->nn->rm_out = 0.0;       // the action kernels carry no measurement term any more
  ctx->nn->rm_in_mat = rm_in_dev; ctx->nn->rm_out_mat = rm_out_dev;
  return VAB_OK;
}

int vab_nn_set_fixed_params(vab_ctx* ctx, const double* pfix_dev, int64_t pfix_stride) {
  if (!ctx) return VAB_ERR_INVALID;
  if (!ctx->nn) return vab_fail(ctx, VAB_ERR_STATE, "nn_set_fixed_params: no NN problem set");
  if (!pfix_dev || (pfix_stride != 0 && pfix_stride != ctx->nn->NP))
    return vab_fail(ctx, VAB_ERR_INVALID, "nn_set_fixed_params: NULL or stride not 0 / NP");
  ctx->nn->pfix = pfix_dev; ctx->nn->pfix_stride = pfix_stride;
  return VAB_OK;
}

int vab_nn_action_grad(vab_ctx* ctx, int32_t B, const double* XP_dev, int64_t ldxp, double rf_scale,
                       double* A_dev, double* me_dev, double* fe_dev, double* G_dev, int64_t ldg) {
  if (!ctx) return VAB_ERR_INVALID;
  if (ctx->problem != VAB_PROBLEM_NN) return vab_fail(ctx, VAB_ERR_STATE, "nn_action_grad: no NN problem set");
  cudaSetDevice(ctx->device);
  return nn_eval(ctx, B, XP_dev, ldxp, rf_scale, nullptr, nullptr, A_dev, me_dev, fe_dev, G_dev, ldg);
}

}  // extern "C"
