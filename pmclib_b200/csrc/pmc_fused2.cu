// pmc_fused2.cu — second-generation fused PMC iteration kernel (the one bench.py times).
//
// Same contract and outputs as pmc_fused.cu (draw + mixture log-density + target + log-weight + RB statistics in
// one pass), re-shaped after the first ncu capture (profiles/r01_v1_fused_ncu.txt) showed the first kernel bound by
// shared-memory wavefronts (65 % of peak) with the FP64 pipe only 34 % busy:
//
//   * (D, K) are template parameters and the mixture lives in __constant__ memory.  With the component loop fully
//     unrolled every Cholesky-factor entry is an immediate constant-bank operand of a DFMA  (DFMA R, R, c[3][imm], R):
//     no load instruction, no shared-memory wavefront, warp-uniform by construction.
//   * per-component log-densities stay in registers (no shared-memory round trip).
//   * phase 2 (the K x E statistic GEMM  C[k][e] += g[s][k] * P[s][e]) runs on the FP64 tensor path:
//     mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  DMMA shares the FP64 pipe with DFMA (no extra flops, see
//     profiles/r01_fp64_peak.md) but one instruction does 256 FMAs from two operands per lane, which removes the
//     operand-bandwidth wall of a per-lane DFMA register tile (2 B of shared memory per FMA instead of 4.7 B).
//     Tiles are stored with odd row strides and k-steps take samples {t, t+4, t+8, t+12}: fragment loads and row
//     stores are both bank-conflict-free.
#include "pmc_kernels.h"

namespace pmcb200 {

constexpr int F2_WARPS = 8;
constexpr int F2_CMIX = 4096, F2_CTGT = 2048;
__constant__ double c_mix[F2_CMIX];
__constant__ double c_tgt[F2_CTGT];

__host__ __device__ constexpr int f2_ev2(int n) { return (n + 1) & ~1; }
__host__ __device__ constexpr int f2_oRinv(int D) { return f2_ev2(D); }
__host__ __device__ constexpr int f2_oLc(int D) { return f2_oRinv(D) + f2_ev2(D); }
__host__ __device__ constexpr int f2_oLr(int D) { return f2_oLc(D) + f2_ev2(D * (D - 1) / 2); }
__host__ __device__ constexpr int f2_oSc(int D) { return f2_oLr(D) + f2_ev2(D * (D + 1) / 2); }
__host__ __device__ constexpr int f2_stride(int D) {
  return ((f2_oSc(D) + 2) / 2) % 2 == 1 ? f2_oSc(D) + 2 : f2_oSc(D) + 4;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// |L^-1 (x - mu)|^2 with every parameter read from a constant bank at a compile-time offset
template <int D, int SPL, bool TGT>
__device__ __forceinline__ void whiten_c(int base, const double (&x)[SPL][D], double (&q)[SPL]) {
  const double *cp = TGT ? c_tgt : c_mix;
  double v[SPL][D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int s = 0; s < SPL; ++s) v[s][a] = x[s][a] - cp[base + a];
  }
#pragma unroll
  for (int s = 0; s < SPL; ++s) q[s] = 0.0;
  int off = base + f2_oLc(D);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double y[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      y[s] = v[s][j] * cp[base + f2_oRinv(D) + j];
      q[s] = fma(y[s], y[s], q[s]);
    }
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) v[s][i] = fma(-cp[off + (i - j - 1)], y[s], v[s][i]);
    }
    off += D - 1 - j;
  }
}

template <int D>
struct RegVec2 {
  const double (&v)[D];
  __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};

template <int D, int K, int KT, int SPL>
struct F2Cfg {
  static constexpr int E = 1 + D + D * (D + 1) / 2;
  static constexpr int NT = (E + 7) / 8;   // entry tiles (n8)
  static constexpr int MT = (K + 7) / 8;   // component tiles (m8)
  static constexpr int PS = (NT * 8) | 1;  // odd row strides: conflict-free row stores and fragment loads
  static constexpr int GS = (MT * 8) | 1;
  static constexpr int TS = 32 * SPL;
  static constexpr int ACC = MT * NT * 64;  // doubles per warp / per block partial
  static constexpr int ST = f2_stride(D);
  static constexpr int WARP_DOUBLES = 32 * PS + 32 * GS;
};

template <int D, int K, int KT, int SPL>
__global__ void __launch_bounds__(F2_WARPS * 32, 1) k_fused2(const FusedArgs a) {
  using C = F2Cfg<D, K, KT, SPL>;
  constexpr int E = C::E, NT = C::NT, MT = C::MT, PS = C::PS, GS = C::GS, TS = C::TS, ST = C::ST;
  static_assert(K * ST <= F2_CMIX && KT * ST <= F2_CTGT, "constant memory budget");
  static_assert(TS * D <= 32 * PS, "X staging must fit in the P tile");
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  double *mixs = smem;                      // copy of the packed mixture for the draw (per-lane component)
  double *cws = mixs + K * ST;
  double *warp0 = cws + f2_ev2(K);
  double *ptile = warp0 + (size_t)wid * C::WARP_DOUBLES;
  double *gtile = ptile + 32 * PS;
  __shared__ unsigned int s_count[64];
  __shared__ double s_m[F2_WARPS], s_S[F2_WARPS], s_r[F2_WARPS];

  for (int i = threadIdx.x; i < K * ST; i += blockDim.x) mixs[i] = c_mix[i];
  if (threadIdx.x < 64) s_count[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {  // cumulative weights exactly as mvdens.c:748-750 (sequential sum)
    double c = 0.0;
    for (int k = 0; k < K; ++k) { c = (k == 0) ? mixs[f2_oSc(D)] : c + mixs[k * ST + f2_oSc(D)]; cws[k] = c; }
  }
  // zero the padding rows/columns of this warp's tiles once (entries E..8NT-1, components K..8MT-1)
  for (int i = lane; i < C::WARP_DOUBLES; i += 32) ptile[i] = 0.0;
  __syncthreads();

  double acc[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  double m_run = -INFINITY, S_lane = 0.0, maxr_lane = -INFINITY;
  unsigned int nok_lane = 0, rej_lane = 0, fail_lane = 0;

  const long ntiles = (a.N + TS - 1) / TS;
  const Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const int fr = lane >> 2, fc = lane & 3;  // fragment row (component / entry within a tile), k-step column (sample)

  for (long tile = (long)blockIdx.x * F2_WARPS + wid; tile < ntiles; tile += (long)gridDim.x * F2_WARPS) {
    const long tbase = tile * TS;
    const int tn = (a.N - tbase) < TS ? (int)(a.N - tbase) : TS;
    double x[SPL][D];
    bool valid[SPL];
    int kdraw[SPL];
    double *xt = ptile;  // X staging aliases the P tile

    // ---------------- draw or load --------------------------------------------------------------------------
    if (a.do_draw) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        valid[s] = i < a.N;
        kdraw[s] = 0;
        bool ok = false;
        if (valid[s]) {
          const uint64_t gi = (uint64_t)(a.first_index + i);
          const uint32_t ilo = (uint32_t)gi, ihi = (uint32_t)(gi >> 32);
          for (int attempt = 0; attempt < a.max_attempts; ++attempt) {
            const uint32_t cb = (uint32_t)attempt << 12;
            const uint4 r0 = ph(ilo, ihi, cb, a.iter);
            const int k = ranindx(cws, K, u01_co(r0.x, r0.y));
            double g[D + 1];
#pragma unroll
            for (int j = 0; j < D; j += 2) box_muller(ph(ilo, ihi, cb | (1 + j / 2), a.iter), g[j], g[j + 1]);
            const double *cp = mixs + k * ST;
            bool inb = true;
#pragma unroll
            for (int i2 = 0; i2 < D; ++i2) {
              double t = 0.0;
#pragma unroll
              for (int j = 0; j <= i2; ++j) t = fma(g[j], cp[f2_oLr(D) + i2 * (i2 + 1) / 2 + j], t);
              const double xv = t + cp[i2];
              x[s][i2] = xv;
              inb = inb && isfinite(xv);
              if (a.has_box) inb = inb && !(a.lo_thr[i2] + xv < 0) && !(a.hi_thr[i2] - xv < 0);
            }
            kdraw[s] = k;
            if (inb) { ok = true; break; }
            ++rej_lane;
          }
          if (!ok) ++fail_lane;
        } else {
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) x[s][i2] = 0.0;
        }
        valid[s] = valid[s] && ok;
        if (i < a.N) {
          a.indices[i] = (unsigned long long)kdraw[s];
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) xt[(s * 32 + lane) * D + i2] = x[s][i2];
        }
      }
      __syncwarp();
      for (int idx = lane; idx < tn * D; idx += 32) a.X[(size_t)tbase * D + idx] = xt[idx];
      __syncwarp();
    } else {
      for (int idx = lane; idx < tn * D; idx += 32) xt[idx] = a.X[(size_t)tbase * D + idx];
      __syncwarp();
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        valid[s] = i < a.N;
#pragma unroll
        for (int i2 = 0; i2 < D; ++i2) x[s][i2] = valid[s] ? xt[(s * 32 + lane) * D + i2] : 0.0;
        kdraw[s] = valid[s] ? (int)a.indices[i] : 0;
      }
      __syncwarp();
    }

    // ---------------- target ------------------------------------------------------------------------------------
    double post[SPL];
    bool tok[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) { post[s] = 0.0; tok[s] = true; }
    if (a.tgt_kind == PMCGPU_TGT_BANANA) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) post[s] = bent_gauss(RegVec2<D>{x[s]}, D, a.b, a.sig1);
    } else if (KT > 0 && a.tgt_kind == PMCGPU_TGT_MVDENS) {
      double q[SPL];
      whiten_c<D, SPL, true>(0, x, q);
#pragma unroll
      for (int s = 0; s < SPL; ++s) post[s] = -0.5 * (D * kLn2Pi + q[s]) - c_tgt[f2_oSc(D) + 1];
    } else if (KT > 0 && a.tgt_kind == PMCGPU_TGT_MIX) {
      double tl[SPL][KT > 0 ? KT : 1], mx[SPL];
      bool first = true;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        if (c_tgt[k * ST + f2_oSc(D)] > 0.0) {
          double q[SPL];
          whiten_c<D, SPL, true>(k * ST, x, q);
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            tl[s][k] = -0.5 * (D * kLn2Pi + q[s]) - c_tgt[k * ST + f2_oSc(D) + 1];
            if (first || mx[s] < tl[s][k]) mx[s] = tl[s][k];
          }
          first = false;
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        double lp = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (c_tgt[k * ST + f2_oSc(D)] > 0.0) lp = lp + exp(tl[s][k] - mx[s] + kDynMax) * c_tgt[k * ST + f2_oSc(D)];
        post[s] = mx[s] + log(lp) - kDynMax;
      }
    } else if (a.tgt_kind == PMCGPU_TGT_HOST) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        if (i < a.N) { post[s] = a.post_in[i]; tok[s] = a.ok_in ? (a.ok_in[i] != 0) : true; }
      }
    }

    // ---------------- proposal log-density (mvdens.c:784-825), component loop fully unrolled ----------------------
    double ev[SPL][K], mx[SPL], lpos[SPL];
    {
      bool first = true;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (c_mix[k * ST + f2_oSc(D)] > 0.0) {
          double q[SPL];
          whiten_c<D, SPL, false>(k * ST, x, q);
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            ev[s][k] = -0.5 * (D * kLn2Pi + q[s]) - c_mix[k * ST + f2_oSc(D) + 1];
            if (first || mx[s] < ev[s][k]) mx[s] = ev[s][k];
          }
          first = false;
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        lpos[s] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          double e = 0.0;
          if (c_mix[k * ST + f2_oSc(D)] > 0.0) { e = exp(ev[s][k] - mx[s] + kDynMax) * c_mix[k * ST + f2_oSc(D)]; lpos[s] = lpos[s] + e; }
          ev[s][k] = e;
        }
      }
    }

    // ---------------- log-weight, flags, maxima ----------------------------------------------------------------
    double lw[SPL], w[SPL];
    bool wfin[SPL];
    double tmax = -INFINITY;
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      const long i = tbase + s * 32 + lane;
      const double rloc = mx[s] + log(lpos[s]) - kDynMax;
      const bool rfin = !(isnan(rloc) || isinf(rloc));
      const bool reached = valid[s] && rfin && tok[s];  // pmc.c:546-583
      lw[s] = reached ? (a.t_temper * post[s] - rloc) : 0.0;
      if (i < a.N) {
        a.log_rho[i] = rloc;
        a.lw[i] = lw[s];
        a.flg[i] = reached ? 1 : 0;
        if (a.first_index + i == 0) {
          a.first_vals[0] = rloc; a.first_vals[1] = 1.0; a.first_vals[2] = lw[s]; a.first_vals[3] = reached ? 1.0 : 0.0;
        }
        if (rloc > maxr_lane) maxr_lane = rloc;
      }
      wfin[s] = reached && !isnan(lw[s]) && !(lw[s] == INFINITY);
      if (reached && lw[s] == INFINITY) atomicAdd(&a.counters[CNT_POSINF], 1ull);
      if (wfin[s] && lw[s] > tmax) tmax = lw[s];
    }
    tmax = warp_max_nan_ignoring(tmax);
    if (tmax > m_run) {  // online rescaling of everything accumulated so far
      const double sc = (m_run == -INFINITY) ? 0.0 : exp(m_run - tmax);
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) { acc[m][n][0] *= sc; acc[m][n][1] *= sc; }
      S_lane *= sc;
      m_run = tmax;
    }
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      w[s] = (wfin[s] && lw[s] > -INFINITY) ? exp(lw[s] - m_run) : 0.0;
      S_lane += w[s];
      if (wfin[s]) { ++nok_lane; atomicAdd(&s_count[kdraw[s] & 63], 1u); }
    }

    // ---------------- phase 2: statistics on the FP64 tensor path ---------------------------------------------------
    if (a.do_stats) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        {
          const double rs = (w[s] != 0.0) ? w[s] / lpos[s] : 0.0;
          double *grow = gtile + lane * GS;
#pragma unroll
          for (int k = 0; k < K; ++k) grow[k] = ev[s][k] * rs;  // g_ik = w_i alpha_k phi_k(x_i) / q(x_i)
          double xc[D];
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) xc[i2] = (w[s] != 0.0) ? x[s][i2] - a.pivot[i2] : 0.0;
          double *prow = ptile + lane * PS;
          prow[0] = 1.0;
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) prow[1 + i2] = xc[i2];
          int e = 1 + D;
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2)
#pragma unroll
            for (int j = i2; j < D; ++j) prow[e++] = xc[i2] * xc[j];
          // entries E..8NT-1 and components K..8MT-1 are padding: whatever they hold only reaches accumulators
          // that the combine step never reads
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int smp = 16 * h + t + 4 * fc;
            double af[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) af[m] = gtile[smp * GS + 8 * m + fr];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const double bf = ptile[smp * PS + 8 * n + fr];
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma884(acc[m][n][0], acc[m][n][1], af[m], bf);
            }
          }
        __syncwarp();
      }
    }
  }

  // ---------------- block combine (fixed order -> deterministic) ----------------------------------------------
  __syncthreads();
  if (lane == 0) s_m[wid] = m_run;
  {
    const double Sw = warp_sum(S_lane);
    const double Rw = warp_max_nan_ignoring(maxr_lane);
    if (lane == 0) { s_S[wid] = Sw; s_r[wid] = Rw; }
  }
  __syncthreads();
  double Mb = -INFINITY, Rb = -INFINITY;
#pragma unroll
  for (int w2 = 0; w2 < F2_WARPS; ++w2) { if (s_m[w2] > Mb) Mb = s_m[w2]; if (s_r[w2] > Rb) Rb = s_r[w2]; }
  const double scale = (m_run == -INFINITY) ? 0.0 : exp(m_run - Mb);
  double *comb = warp0;
  for (int w2 = 0; w2 < F2_WARPS; ++w2) {
    if (wid == w2) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int idx = ((m * NT + n) * 32 + lane) * 2 + c;
            comb[idx] = (w2 == 0 ? 0.0 : comb[idx]) + acc[m][n][c] * scale;
          }
    }
    __syncthreads();
  }
  double *out = a.partials + (size_t)blockIdx.x * a.partial_len;
  for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = comb[idx];
  if (threadIdx.x == 0) {
    double Sb = 0.0;
    for (int w2 = 0; w2 < F2_WARPS; ++w2) Sb += (s_m[w2] == -INFINITY) ? 0.0 : s_S[w2] * exp(s_m[w2] - Mb);
    out[C::ACC + 0] = Mb;
    out[C::ACC + 1] = Sb;
    out[C::ACC + 2] = Rb;
    out[C::ACC + 3] = 0.0;
  }
  {
    const unsigned long long n1 = (unsigned long long)warp_sum_ll(nok_lane);
    const unsigned long long n2 = (unsigned long long)warp_sum_ll(rej_lane);
    const unsigned long long n3 = (unsigned long long)warp_sum_ll(fail_lane);
    if (lane == 0) {
      if (n1) atomicAdd(&a.counters[CNT_NOK], n1);
      if (n2) atomicAdd(&a.counters[CNT_REJECT], n2);
      if (n3) atomicAdd(&a.counters[CNT_DRAWFAIL], n3);
    }
  }
  if (threadIdx.x < 64 && threadIdx.x < K && s_count[threadIdx.x]) atomicAdd(&a.count_k[threadIdx.x], (unsigned long long)s_count[threadIdx.x]);
}

// canonical statistics stats[k*E + e] from the DMMA accumulator fragments of all blocks, rescaled to the global max
__global__ void k_fused2_combine(const double *__restrict__ partials, int nblocks, int partial_len, int acc_len, int K,
                                 int E, int NT, double *__restrict__ out) {
  extern __shared__ double sc[];
  __shared__ double sM;
  if (threadIdx.x == 0) {
    double M = -INFINITY;
    for (int b = 0; b < nblocks; ++b) { const double mb = partials[(size_t)b * partial_len + acc_len]; if (mb > M) M = mb; }
    sM = M;
  }
  __syncthreads();
  const double M = sM;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double mb = partials[(size_t)b * partial_len + acc_len];
    sc[b] = (mb == -INFINITY) ? 0.0 : exp(mb - M);
  }
  __syncthreads();
  const int nstat = K * E;
  for (int t = threadIdx.x; t < nstat; t += blockDim.x) {
    const int k = t / E, e = t - k * E;
    const int m = k >> 3, r = k & 7, n = e >> 3, cc = e & 7;
    const int idx = ((m * NT + n) * 32 + (r * 4 + (cc >> 1))) * 2 + (cc & 1);  // C fragment: row = lane/4, cols 2(lane%4)+{0,1}
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * partial_len + idx] * sc[b];
    out[t] = s;
  }
  if (threadIdx.x == 0) {
    double S = 0.0, R = -INFINITY;
    for (int b = 0; b < nblocks; ++b) {
      S += partials[(size_t)b * partial_len + acc_len + 1] * sc[b];
      const double rb = partials[(size_t)b * partial_len + acc_len + 2];
      if (rb > R) R = rb;
    }
    out[nstat + 0] = M;
    out[nstat + 1] = S;
    out[nstat + 2] = R;
    out[nstat + 3] = 0.0;
  }
}

// ---- instantiation table: (D, K, KT, SPL) ------------------------------------------------------------------------
#define PMC_FUSED2_TABLE(X) \
  X(10, 10, 0, 1)           \
  X(5, 20, 4, 1)            \
  X(2, 9, 0, 1)

template <int D, int K, int KT, int SPL>
static int launch_one2(const FusedPlan &p, const FusedArgs &a, const double *h_mix, const double *h_tgt, cudaStream_t s) {
  auto kern = k_fused2<D, K, KT, SPL>;
  static size_t configured = 0;
  if (p.smem_bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess) return -1;
    configured = p.smem_bytes;
  }
  if (h_mix && cudaMemcpyToSymbolAsync(c_mix, h_mix, sizeof(double) * K * f2_stride(D), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (KT > 0 && h_tgt && cudaMemcpyToSymbolAsync(c_tgt, h_tgt, sizeof(double) * KT * f2_stride(D), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  kern<<<p.grid, p.threads, p.smem_bytes, s>>>(a);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

FusedPlan fused2_plan(int d, int K, int Kt, int tgt_kind, int sm_count, int want_spl) {
  FusedPlan best{};
  best.valid = 0;
  const int kt_needed = (tgt_kind == PMCGPU_TGT_MVDENS || tgt_kind == PMCGPU_TGT_MIX) ? Kt : 0;
#define X(D_, K_, KT_, SPL_)                                                                             \
  if (!best.valid && d == D_ && K == K_ && (kt_needed == 0 || kt_needed == KT_) && SPL_ == want_spl) {   \
    using C = F2Cfg<D_, K_, KT_, SPL_>;                                                                  \
    const size_t dbl = (size_t)K_ * C::ST + f2_ev2(K_) + (size_t)F2_WARPS * C::WARP_DOUBLES;             \
    if (dbl * 8 <= 227 * 1024) {                                                                         \
      best.valid = 2; best.D = D_; best.RK = K_; best.RE = KT_; best.NR = 0; best.SPL = SPL_;            \
      best.grid = sm_count; best.threads = F2_WARPS * 32; best.smem_bytes = dbl * 8;                     \
      best.partial_len = C::ACC + 4; best.nstat = K_ * C::E;                                             \
    }                                                                                                    \
  }
  PMC_FUSED2_TABLE(X)
#undef X
  return best;
}

int launch_fused2(const FusedPlan &p, const FusedArgs &a, const double *h_mix, const double *h_tgt, cudaStream_t s) {
#define X(D_, K_, KT_, SPL_) \
  if (p.D == D_ && p.RK == K_ && p.RE == KT_ && p.SPL == SPL_) return launch_one2<D_, K_, KT_, SPL_>(p, a, h_mix, h_tgt, s);
  PMC_FUSED2_TABLE(X)
#undef X
  return -1;
}

int launch_fused2_combine(const FusedPlan &p, const double *partials, double *stats_out, cudaStream_t s) {
  const int E = 1 + p.D + p.D * (p.D + 1) / 2;
  const int NT = (E + 7) / 8;
  k_fused2_combine<<<1, 256, p.grid * sizeof(double), s>>>(partials, p.grid, p.partial_len, p.partial_len - 4, p.RK, E, NT, stats_out);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace pmcb200
