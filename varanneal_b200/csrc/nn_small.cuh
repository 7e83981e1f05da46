// Small networks on the CUDA cores (the bar-image classifier [25, 30, 4] of
// examples/nnet_barimages, BASELINE.json configs[4]): layers this narrow fill an 8 x 8 tensor-pipe
// tile mostly with padding and the evaluation is bound by HBM and by fp64 FMA issue at the same
// time (5 220 flop and 944 B per example: 5.5 flop / B, the ridge of the machine), so the three
// contractions of a layer (va_nnet.py:210-255) run as register-tiled 4 x 4 micro-GEMMs from
// *transposed* shared-memory tiles -- one 16-byte shared load feeds 8 FMAs -- instead of DMMA
// fragments with 8-byte staging copies.
//
// One CTA walks TPC tiles of T = 32 examples of one path (three CTAs per SM).  Per tile: the
// contiguous rows x NDnet block of the path arrives with one cp.async.bulk, is transposed in shared
// memory (Xt[column][example]); per layer
//   F  Z = X_n W_n^T + b, sigma, residual, lambda (-> gradient rows of layer n+1), Delta (-> Dt)
//   B  gradient rows of layer n += Delta W_n
//   W  GW_n += Delta^T X_n, Gb_n += sum Delta   -- accumulated in registers over all tiles of the
//      CTA (every thread owns one 4 x 4 block of one layer's weight gradient), written once per CTA
// and the gradient rows leave through a coalesced transposing store.  Same outputs and the same
// fixed-order reduction (nn_reduce_kernel) as the other NN kernels: bit-reproducible.
#pragma once

namespace {

constexpr int SM_NT = 128;        // threads
constexpr int SM_T = 32;          // examples per tile
constexpr int SM_TP = 34;         // pitch of the transposed tiles (even: 16-byte aligned 4-example groups)
constexpr int SM_MAXL = 8;        // layers - 1

struct NnSmallPlan {
  int nl1;                        // layers - 1
  int woffs[SM_MAXL];             // offset (doubles) of W_n (row-major [j][i], pitch wp[n]) in the weight area
  int wtoffs[SM_MAXL];            // offset of Wt_n ([i][j], pitch wtp[n])
  int boffs[SM_MAXL];
  int wp[SM_MAXL], wtp[SM_MAXL];
  int blk0[SM_MAXL];              // first thread owning a weight-gradient block (2 outputs x 4 inputs) of layer n
  int wtotal;                     // doubles of the weight area
  int dmax1;                      // widest layer after the first
  int doff[SM_MAXL];              // first row of layer n's Delta in Dt
  int drows;                      // rows of Dt
  int tpc;                        // tiles per CTA
  int ncta;                       // CTAs per path
  int bulk;                       // 1: the raw tile is fetched with one cp.async.bulk (16-byte aligned rows of tiles)
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

__global__ void __launch_bounds__(SM_NT, 4) nn_small_kernel(const __grid_constant__ NnParams P, const __grid_constant__ NnSmallPlan L) {
  extern __shared__ __align__(128) double sm[];
  const int b = blockIdx.y, cta = blockIdx.x, tid = threadIdx.x;
  if (P.active != nullptr && P.active[b] == 0) return;
  double* Xt = sm;                                   // [NDnet][TP]
  double* Gt = Xt + (size_t)P.NDnet * SM_TP;         // [NDnet][TP]; first the landing zone of the raw tile
  double* Dt = Gt + (size_t)P.NDnet * SM_TP;         // Delta of every layer: [sum of d_{n+1} rounded to 2][TP]
  double* Wa = Dt + (size_t)L.drows * SM_TP;
  __shared__ double red[2][SM_NT / 32];
  __shared__ __align__(8) unsigned long long bar;
  const double* xp = P.XP + (long long)b * P.ldxp;
  double* gp = P.G ? P.G + (long long)b * P.ldg : nullptr;
  const double* pfull = P.pfull + (long long)b * P.NP;
  const double cf2 = (P.rf_path != nullptr) ? P.cf2_num * __ldg(P.rf_path + b) / P.cf2_den : P.cf2;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- weights: both layouts, zero-padded to multiples of 4
  for (int k = tid; k < L.wtotal; k += SM_NT) Wa[k] = 0.0;
  __syncthreads();
  for (int n = 0; n < L.nl1; ++n) {
    const int dn = P.structure[n], dn1 = P.structure[n + 1];
    for (int k = tid; k < dn * dn1; k += SM_NT) {
      const int j = k / dn, i = k - j * dn;
      const double w = __ldg(pfull + P.woff[n] + k);
      Wa[L.woffs[n] + j * L.wp[n] + i] = w;
      Wa[L.wtoffs[n] + i * L.wtp[n] + j] = w;
    }
    for (int j = tid; j < dn1; j += SM_NT) Wa[L.boffs[n] + j] = __ldg(pfull + P.boff[n] + j);
  }
  // ---- which weight-gradient block (2 outputs x 4 inputs) this thread accumulates over the CTA's tiles
  int my_n = -1, my_jb = 0, my_ib = 0;
  for (int n = 0; n < L.nl1; ++n) {
    const int nib = (P.structure[n] + 3) >> 2, njb = (P.structure[n + 1] + 1) >> 1;
    if (tid >= L.blk0[n] && tid < L.blk0[n] + nib * njb) { my_n = n; my_jb = (tid - L.blk0[n]) / nib; my_ib = (tid - L.blk0[n]) - my_jb * nib; }
  }
  double gwacc[2][4], gbacc[2] = {0.0, 0.0}, me_acc = 0.0, fe_acc = 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) gwacc[a][c] = 0.0;

  const int ntiles = (P.M + SM_T - 1) / SM_T;
  const int d0 = P.structure[0], ND = P.NDnet;
  uint32_t phase = 0;
  for (int tt = 0; tt < L.tpc; ++tt) {
    const int tile = cta * L.tpc + tt;
    if (tile >= ntiles) break;
    const int m0 = tile * SM_T;
    const int rows = min(SM_T, P.M - m0);
    __syncthreads();                                   // previous tile fully consumed (Gt is the landing zone)
    // ---- the tile is contiguous in the path (rows x NDnet doubles): one bulk copy into the landing zone
    if (L.bulk) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)rows * ND * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(Gt)), "l"(xp + (long long)m0 * ND), "r"(bytes), "r"(bar_a) : "memory");
      }
      asm volatile(
          "{\n\t.reg .pred p;\n\tWAITB_%=:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONEB_%=;\n\tbra WAITB_%=;\n\tDONEB_%=:\n\t}"
          ::"r"(bar_a), "r"(phase) : "memory");
      phase ^= 1u;
    } else {
      for (int idx = tid; idx < rows * ND; idx += SM_NT) Gt[idx] = xp[(long long)m0 * ND + idx];
      __syncthreads();
    }
    // transpose: thread = (example m, column group); consecutive lanes = consecutive examples -> conflict-free stores
    for (int idx = tid; idx < SM_T * ND; idx += SM_NT) {
      const int c = idx / SM_T, m = idx - c * SM_T;
      Xt[c * SM_TP + m] = (m < rows) ? Gt[m * ND + c] : 0.0;
    }
    __syncthreads();
    // gradient rows start from the measurement term of the input layer, zero elsewhere
    for (int idx = tid; idx < SM_T * ND; idx += SM_NT) {
      const int c = idx / SM_T, m = idx - c * SM_T;
      double gx = 0.0;
      if (m < rows && c < d0) {
        const int s = P.slot_in[c];
        if (s >= 0) {
          const double diff = Xt[c * SM_TP + m] - P.data_in[(long long)(m0 + m) * P.n_Lin + s];
          me_acc = fma(P.wm_in * diff, diff, me_acc);
          gx = P.wm_in * diff;
        }
      }
      Gt[c * SM_TP + m] = gx;
    }
    for (int n = 0; n < L.nl1; ++n) {
      const int dn = P.structure[n], dn1 = P.structure[n + 1];
      const int xo = P.xoff[n], xo1 = P.xoff[n + 1];
      const bool lastl = (n + 1 == P.NL - 1);
      const double* W = Wa + L.woffs[n];
      const double* Wt = Wa + L.wtoffs[n];
      const double* bs = Wa + L.boffs[n];
      const int wp = L.wp[n], wtp = L.wtp[n];
      __syncthreads();
      double* Dn = Dt + (size_t)L.doff[n] * SM_TP;        // this layer's Z, then Delta
      // ---- F: Z = X_n W_n^T.  4 examples x 2 outputs per thread, or one output per thread when the layer is so
      //      narrow that the 4 x 2 blocks would leave most of the CTA idle
      {
        const int njb = (dn1 + 1) >> 1;
        if ((SM_T / 4) * njb * 2 > SM_NT) {
          for (int blk = tid; blk < (SM_T / 4) * njb; blk += SM_NT) {
            const int jb = blk / (SM_T / 4), mb = blk - jb * (SM_T / 4);
            double acc[4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a) { acc[a][0] = 0.0; acc[a][1] = 0.0; }
            const double* xcol = Xt + (size_t)xo * SM_TP + 4 * mb;
            const double* wrow = Wt + 2 * jb;
#pragma unroll 5
            for (int i = 0; i < dn; ++i) {
              const double2 x01 = ld2(xcol + (size_t)i * SM_TP), x23 = ld2(xcol + (size_t)i * SM_TP + 2);
              const double2 w = ld2(wrow + i * wtp);
              acc[0][0] = fma(x01.x, w.x, acc[0][0]); acc[0][1] = fma(x01.x, w.y, acc[0][1]);
              acc[1][0] = fma(x01.y, w.x, acc[1][0]); acc[1][1] = fma(x01.y, w.y, acc[1][1]);
              acc[2][0] = fma(x23.x, w.x, acc[2][0]); acc[2][1] = fma(x23.x, w.y, acc[2][1]);
              acc[3][0] = fma(x23.y, w.x, acc[3][0]); acc[3][1] = fma(x23.y, w.y, acc[3][1]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              double* zrow = Dn + (size_t)(2 * jb + c) * SM_TP + 4 * mb;
              *reinterpret_cast<double2*>(zrow) = make_double2(acc[0][c], acc[1][c]);
              *reinterpret_cast<double2*>(zrow + 2) = make_double2(acc[2][c], acc[3][c]);
            }
          }
        } else {
          for (int o = tid; o < SM_T * dn1; o += SM_NT) {
            const int j = o / SM_T, m = o - j * SM_T;
            const double* xcol = Xt + (size_t)xo * SM_TP + m;
            double a0 = 0.0, a1 = 0.0;
            int i = 0;
            for (; i + 1 < dn; i += 2) {
              a0 = fma(xcol[(size_t)i * SM_TP], Wt[i * wtp + j], a0);
              a1 = fma(xcol[(size_t)(i + 1) * SM_TP], Wt[(i + 1) * wtp + j], a1);
            }
            if (i < dn) a0 = fma(xcol[(size_t)i * SM_TP], Wt[i * wtp + j], a0);
            Dn[(size_t)j * SM_TP + m] = a0 + a1;
          }
        }
      }
      __syncthreads();
      // ---- epilogue, one (output, example) per thread and step: sigma, residual, lambda, Delta
      {
        const int dn1e = (dn1 + 1) & ~1;                    // the odd row of the last 2-output block is cleared
        for (int o = tid; o < SM_T * dn1e; o += SM_NT) {
          const int j = o / SM_T, m = o - j * SM_T;
          double dl = 0.0;
          if (j < dn1) {
            double gx = 0.0;
            if (m < rows) {
              const double sv = act_f(P.act, Dn[(size_t)j * SM_TP + m] + bs[j]);
              const double xn1 = Xt[(size_t)(xo1 + j) * SM_TP + m];
              const double e = xn1 - sv;
              const double lam = cf2 * e;
              fe_acc = fma(lam, e, fe_acc);
              dl = -lam * act_d(P.act, sv);
              gx = lam;
              if (lastl) {
                const int so = P.slot_out[j];
                if (so >= 0) {
                  const double diff = xn1 - P.data_out[(long long)(m0 + m) * P.n_Lout + so];
                  me_acc = fma(P.wm_out * diff, diff, me_acc);
                  gx += P.wm_out * diff;
                }
              }
            }
            Gt[(size_t)(xo1 + j) * SM_TP + m] = gx;
          }
          Dn[(size_t)j * SM_TP + m] = dl;
        }
      }
      __syncthreads();
      // ---- B: gradient rows of layer n += Delta W_n (4 examples x 2 inputs per thread)
      {
        const int nib = (dn + 1) >> 1;
        for (int blk = tid; blk < (SM_T / 4) * nib; blk += SM_NT) {
          const int ib = blk / (SM_T / 4), mb = blk - ib * (SM_T / 4);
          double acc[4][2];
#pragma unroll
          for (int a = 0; a < 4; ++a) { acc[a][0] = 0.0; acc[a][1] = 0.0; }
#pragma unroll 5
          for (int j = 0; j < dn1; ++j) {
            const double2 d01 = ld2(Dn + (size_t)j * SM_TP + 4 * mb), d23 = ld2(Dn + (size_t)j * SM_TP + 4 * mb + 2);
            const double2 w = ld2(W + j * wp + 2 * ib);
            acc[0][0] = fma(d01.x, w.x, acc[0][0]); acc[0][1] = fma(d01.x, w.y, acc[0][1]);
            acc[1][0] = fma(d01.y, w.x, acc[1][0]); acc[1][1] = fma(d01.y, w.y, acc[1][1]);
            acc[2][0] = fma(d23.x, w.x, acc[2][0]); acc[2][1] = fma(d23.x, w.y, acc[2][1]);
            acc[3][0] = fma(d23.y, w.x, acc[3][0]); acc[3][1] = fma(d23.y, w.y, acc[3][1]);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int i = 2 * ib + c;
            if (i < dn) {
#pragma unroll
              for (int a = 0; a < 4; ++a) Gt[(size_t)(xo + i) * SM_TP + 4 * mb + a] += acc[a][c];
            }
          }
        }
      }
    }
    // ---- W (all layers at once, every Delta is still in Dt): this thread's 2 x 4 block of GW_n, summed over the
    //      tile's examples (+ the bias entries of its 2 outputs when it owns input block 0).  Only reads: no barrier
    //      between the last layer's B phase and here.
    if (my_n >= 0) {
      const double* dcol = Dt + (size_t)(L.doff[my_n] + 2 * my_jb) * SM_TP;
      const double* xcol = Xt + (size_t)(P.xoff[my_n] + 4 * my_ib) * SM_TP;
#pragma unroll 4
      for (int m = 0; m < SM_T; m += 2) {
        const double2 da = ld2(dcol + m), db = ld2(dcol + SM_TP + m);
        double2 xv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xv[c] = ld2(xcol + (size_t)c * SM_TP + m);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          gwacc[0][c] = fma(da.y, xv[c].y, fma(da.x, xv[c].x, gwacc[0][c]));
          gwacc[1][c] = fma(db.y, xv[c].y, fma(db.x, xv[c].x, gwacc[1][c]));
        }
        if (my_ib == 0) { gbacc[0] += da.x + da.y; gbacc[1] += db.x + db.y; }
      }
    }
    __syncthreads();
    // ---- gradient rows of the tile: transposing read (consecutive lanes = consecutive columns), coalesced store
    if (gp) {
      for (int idx = tid; idx < rows * ND; idx += SM_NT) {
        const int m = idx / ND, c = idx - m * ND;
        gp[(long long)(m0 + m) * ND + c] = Gt[c * SM_TP + m];
      }
    }
  }
  // ---- weight-gradient partial of this CTA
  double* gw = P.gwpart + ((long long)b * L.ncta + cta) * P.NP;
  if (my_n >= 0) {
    const int dn = P.structure[my_n], dn1 = P.structure[my_n + 1];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int j = 2 * my_jb + a;
      if (j < dn1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = 4 * my_ib + c;
          if (i < dn) gw[P.woff[my_n] + j * dn + i] = gwacc[a][c];
        }
        if (my_ib == 0) gw[P.boff[my_n] + j] = gbacc[a];
      }
    }
  }
  for (int sft = 16; sft > 0; sft >>= 1) {
    me_acc += __shfl_down_sync(0xffffffffu, me_acc, sft);
    fe_acc += __shfl_down_sync(0xffffffffu, fe_acc, sft);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = me_acc; red[1][tid >> 5] = fe_acc; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < SM_NT / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    P.partials[((long long)b * L.ncta + cta) * 2 + 0] = 0.5 * a;
    P.partials[((long long)b * L.ncta + cta) * 2 + 1] = 0.5 * c;
  }
}

// host side: is the network small enough, and where does everything go
inline bool nn_small_plan(const std::vector<int>& st, int M, int B, int num_sms, long long ldxp, const double* xp,
                          NnSmallPlan* L, size_t* smem) {
  const int nl1 = (int)st.size() - 1;
  if (nl1 < 1 || nl1 > SM_MAXL) return false;
  int nd = 0, dmax1 = 0, off = 0, thr = 0;
  for (int n = 0; n <= nl1; ++n) { if (st[n] > 64) return false; nd += st[n]; if (n > 0 && st[n] > dmax1) dmax1 = st[n]; }
  L->nl1 = nl1;
  for (int n = 0; n < nl1; ++n) {
    L->wp[n] = (st[n] + 3) & ~3;
    L->wtp[n] = (st[n + 1] + 3) & ~3;
    L->woffs[n] = off; off += ((st[n + 1] + 3) & ~3) * L->wp[n];
    L->wtoffs[n] = off; off += ((st[n] + 3) & ~3) * L->wtp[n];
    L->boffs[n] = off; off += (st[n + 1] + 3) & ~3;
    L->blk0[n] = thr; thr += ((st[n] + 3) >> 2) * ((st[n + 1] + 1) >> 1);
  }
  if (thr > SM_NT) return false;
  L->wtotal = off;
  L->dmax1 = dmax1;
  int dr = 0;
  for (int n = 0; n < nl1; ++n) { L->doff[n] = dr; dr += (st[n + 1] + 1) & ~1; }
  L->drows = dr;
  // Gt doubles as the landing zone of the raw tile: T x NDnet <= NDnet x TP always holds (T < TP)
  *smem = ((size_t)2 * nd * SM_TP + (size_t)dr * SM_TP + off) * sizeof(double);
  if (*smem > (size_t)72 * 1024) return false;            // three CTAs per SM (the bar-image network: 57.5 KB)
  const int ntiles = (M + SM_T - 1) / SM_T;
  // enough CTAs for ~8 per SM over the batch, at most 16 tiles each (fewer weight-gradient partials)
  int tpc = 16;
  while (tpc > 1 && (long long)B * ((ntiles + tpc - 1) / tpc) < 8LL * num_sms) tpc >>= 1;
  L->tpc = tpc;
  L->ncta = (ntiles + tpc - 1) / tpc;
  // one bulk copy per tile needs 16-byte aligned tile starts and sizes: paths are 16-byte aligned (ldxp even),
  // a tile of T = 32 examples is 32 NDnet doubles (a multiple of 16 bytes); the ragged last tile too if its row count is even
  L->bulk = ((ldxp & 1) == 0 && (((uintptr_t)xp) & 15) == 0 && ((M % SM_T) * nd) % 2 == 0) ? 1 : 0;
  return true;
}

}  // namespace
