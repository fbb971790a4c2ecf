// pmc_fused.cu — the B200 hot path: ONE pass over the samples for a whole PMC iteration at small d.
//
//   draw (Philox4x32-10 -> component pick -> L g + mu -> box test)            simulate_mix_mvdens   pmc.c:266-298
//   proposal mixture log-density (K triangular solves + max-shifted LSE)      mix_mvdens_log_pdf    mvdens.c:784-825
//   target log-density, log-weight, flags, running maxima                     ..._and_deduced_verb  pmc.c:484-600
//   Rao-Blackwellised sufficient statistics                                   update_prop_rb        pmc.c:1376-1453
//
// Work decomposition (sm_100a, FP64-pipe bound; see DESIGN.md §4 and profiles/r01_fp64_peak.md):
//   phase 1  one sample per lane (SPL=2: two, sharing every Cholesky-factor operand load): everything that is
//            per-sample lives in registers; the mixture sits in shared memory and is read with warp-uniform
//            (broadcast) LDS.128; divisions are multiplications by host-precomputed 1/L_ii.
//   phase 2  the K x E statistic block of a 32-sample tile is a small GEMM  C[k][e] += g[s][k] * P[s][e]
//            (P = 1, x~_a, x~_a x~_b with x~ = x - pivot, computed once per sample and shared by all K components).
//            Lanes switch ownership from "a sample" to "an RK x RE register tile of C"; g and P are exchanged
//            through bank-conflict-free shared-memory tiles.  No atomics: accumulators stay in registers for the
//            whole kernel, weights are kept relative to a warp-running maximum (online rescaling, as in a
//            streaming softmax) and block partials are combined in a fixed order -> deterministic results.
//   X is staged through shared memory so the global store of a tile is one contiguous, fully coalesced block.
#include "pmc_kernels.h"

namespace pmcb200 {

constexpr int FUSED_WARPS = 8;

// packed component layout for dimension D, in doubles (every section starts 16-byte aligned):
//   mu[D] | rinv[D] = 1/L_ii | strict-lower L, column-major packed (column j holds rows j+1..D-1) |
//   lower L incl. diagonal, row-major packed (row i holds columns 0..i) | alpha, logdet
// stride/2 is odd, so lanes that read different components at the same offset (the draw) hit different banks.
__host__ __device__ constexpr int ev2(int n) { return (n + 1) & ~1; }
__host__ __device__ constexpr int pk_oRinv(int D) { return ev2(D); }
__host__ __device__ constexpr int pk_oLc(int D) { return pk_oRinv(D) + ev2(D); }
__host__ __device__ constexpr int pk_oLr(int D) { return pk_oLc(D) + ev2(D * (D - 1) / 2); }
__host__ __device__ constexpr int pk_oSc(int D) { return pk_oLr(D) + ev2(D * (D + 1) / 2); }
__host__ __device__ constexpr int pk_stride(int D) {
  return ((pk_oSc(D) + 2) / 2) % 2 == 1 ? pk_oSc(D) + 2 : pk_oSc(D) + 4;
}

int fused_comp_stride(int D) { return pk_stride(D); }

int fused_pack_mixture(int K, int D, const double *wght, const double *mean, const double *L, const double *logdet,
                       double *out) {
  const int st = pk_stride(D);
  for (int k = 0; k < K; ++k) {
    double *o = out + (size_t)k * st;
    const double *Lk = L + (size_t)k * D * D;
    for (int i = 0; i < st; ++i) o[i] = 0.0;
    if (!(wght[k] > 0.0)) {  // dead component: neutral, finite parameters (unit factor, weight 0); kernels skip it by weight
      for (int a = 0; a < D; ++a) o[pk_oRinv(D) + a] = 1.0;
      int e0 = pk_oLr(D);
      for (int i = 0; i < D; ++i)
        for (int j = 0; j <= i; ++j) o[e0++] = (i == j) ? 1.0 : 0.0;
      continue;
    }
    for (int a = 0; a < D; ++a) {
      o[a] = mean[(size_t)k * D + a];
      o[pk_oRinv(D) + a] = 1.0 / Lk[a * D + a];
    }
    int e = pk_oLc(D);
    for (int j = 0; j < D; ++j)
      for (int i = j + 1; i < D; ++i) o[e++] = Lk[i * D + j];
    e = pk_oLr(D);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) o[e++] = Lk[i * D + j];
    o[pk_oSc(D)] = wght[k];
    o[pk_oSc(D) + 1] = logdet[k];
  }
  return K * st;
}

// Sequential reader of a packed parameter block through 16-byte shared-memory loads: entries are consumed in
// increasing order inside fully unrolled loops, so `have` is a compile-time constant and every pair is loaded once.
struct PairReader {
  const double2 *p;
  double2 cur;
  int have;
  __device__ __forceinline__ explicit PairReader(const double *base) : p(reinterpret_cast<const double2 *>(base)), have(-1) {}
  __device__ __forceinline__ double get(int e) {
    if ((e >> 1) != have) { cur = p[e >> 1]; have = e >> 1; }
    return (e & 1) ? cur.y : cur.x;
  }
};

// |L^-1 (x - mu)|^2 for SPL samples at once; column-oriented forward substitution: after y_j is known every
// remaining row is updated (independent FMAs), each v_i sees the same sequence of updates as cblas dtrsv applies.
template <int D, int SPL>
__device__ __forceinline__ void whiten(const double *__restrict__ cp, const double (&x)[SPL][D], double (&q)[SPL]) {
  double v[SPL][D];
  PairReader rd(cp);
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const double mu = rd.get(a);
#pragma unroll
    for (int s = 0; s < SPL; ++s) v[s][a] = x[s][a] - mu;
  }
#pragma unroll
  for (int s = 0; s < SPL; ++s) q[s] = 0.0;
  PairReader rr(cp), rl(cp);
  int off = pk_oLc(D);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const double ri = rr.get(pk_oRinv(D) + j);
    double y[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      y[s] = v[s][j] * ri;
      q[s] = fma(y[s], y[s], q[s]);
    }
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      const double l = rl.get(off + (i - j - 1));
#pragma unroll
      for (int s = 0; s < SPL; ++s) v[s][i] = fma(-l, y[s], v[s][i]);
    }
    off += D - 1 - j;
  }
}

template <int D>
struct RegVec {  // lets bent_gauss index a register array with compile-time-unrolled loops
  const double (&v)[D];
  __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};

template <int D, int RK, int RE, int NR, int SPL>
struct FusedCfg {
  static constexpr int E = 1 + D + D * (D + 1) / 2;  // statistics per component: 1, x~_a, x~_a x~_b (a<=b)
  static constexpr int NCH = (E + RE - 1) / RE;       // chunks of RE entries
  static constexpr int CSP = RE + 2;                  // chunk stride inside a P row: spreads the NCH chunk heads over banks
  static constexpr int PROW = NCH * CSP;
  static constexpr int PS = PROW + ((PROW % 4 == 2) ? 0 : ((PROW % 4 == 0) ? 2 : ((PROW % 4 == 1) ? 1 : 3)));  // row stride == 2 (mod 4)
  static constexpr int TS = 32 * SPL;                 // samples per warp tile
  static constexpr int ACC = NR * 32 * RK * RE;       // accumulators per warp (= per block partial)
};

template <int D, int RK, int RE, int NR, int SPL>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 1) k_fused(const FusedArgs a) {
  using C = FusedCfg<D, RK, RE, NR, SPL>;
  constexpr int E = C::E, NCH = C::NCH, CSP = C::CSP, PS = C::PS, TS = C::TS;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int K = a.K, Kt = a.Kt;
  const int GS = ev2((K > Kt ? K : Kt) + (RK - 1)) ;  // g-row stride (even, >= K rounded up to RK)

  double *mixs = smem;
  double *tgts = mixs + a.mix_doubles;
  double *cws = tgts + a.tgt_doubles;
  double *warp0 = cws + ev2(K);
  double *ptile = warp0 + (size_t)wid * (32 * PS + TS * GS);
  double *gtile = ptile + 32 * PS;
  __shared__ unsigned int s_count[64];
  __shared__ double s_m[FUSED_WARPS], s_S[FUSED_WARPS], s_r[FUSED_WARPS];

  for (int i = threadIdx.x; i < a.mix_doubles; i += blockDim.x) mixs[i] = a.mixpack[i];
  for (int i = threadIdx.x; i < a.tgt_doubles; i += blockDim.x) tgts[i] = a.tgtpack[i];
  if (threadIdx.x < 64) s_count[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {  // cumulative weights exactly as mvdens.c:748-750 (sequential sum)
    double c = 0.0;
    for (int k = 0; k < K; ++k) { c = (k == 0) ? mixs[pk_oSc(D)] : c + mixs[(size_t)k * a.comp_stride + pk_oSc(D)]; cws[k] = c; }
  }
  __syncthreads();

  double acc[NR][RK][RE];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int i = 0; i < RK; ++i)
#pragma unroll
      for (int j = 0; j < RE; ++j) acc[r][i][j] = 0.0;
  double m_run = -INFINITY, S_lane = 0.0, maxr_lane = -INFINITY;
  unsigned int nok_lane = 0, rej_lane = 0, fail_lane = 0;

  // phase-2 ownership of this lane
  int p2_g[NR], p2_p[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int u = lane + 32 * r;
    int kp = u / NCH;
    const int ch = u - kp * NCH;
    if (kp * RK >= K) kp = 0;  // idle unit: reads valid memory, result ignored by the combine step
    p2_g[r] = kp * RK;
    p2_p[r] = ch * CSP;
  }

  const long ntiles = (a.N + TS - 1) / TS;
  const Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};

  for (long tile = (long)blockIdx.x * FUSED_WARPS + wid; tile < ntiles; tile += (long)gridDim.x * FUSED_WARPS) {
    const long tbase = tile * TS;
    const int tn = (a.N - tbase) < TS ? (int)(a.N - tbase) : TS;  // samples in this tile
    double x[SPL][D];
    bool valid[SPL];
    int kdraw[SPL];
    double *xt = ptile;  // X staging aliases the P tile (P is written later)

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
            const double *cp = mixs + (size_t)k * a.comp_stride;
            PairReader rl(cp), rm(cp);
            bool inb = true;
#pragma unroll
            for (int i2 = 0; i2 < D; ++i2) {
              double t = 0.0;
#pragma unroll
              for (int j = 0; j <= i2; ++j) t = fma(g[j], rl.get(pk_oLr(D) + i2 * (i2 + 1) / 2 + j), t);
              const double xv = t + rm.get(i2);
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

    // ---------------- target (before the proposal: it may use the g row as scratch) -----------------------------
    double post[SPL];
    bool tok[SPL];
    double *grow[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) { grow[s] = gtile + (size_t)(s * 32 + lane) * GS; post[s] = 0.0; tok[s] = true; }
    if (a.tgt_kind == PMCGPU_TGT_BANANA) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) post[s] = bent_gauss(RegVec<D>{x[s]}, D, a.b, a.sig1);
    } else if (a.tgt_kind == PMCGPU_TGT_MVDENS) {
      double q[SPL];
      whiten<D, SPL>(tgts, x, q);
#pragma unroll
      for (int s = 0; s < SPL; ++s) post[s] = -0.5 * (D * kLn2Pi + q[s]) - tgts[pk_oSc(D) + 1];
    } else if (a.tgt_kind == PMCGPU_TGT_MIX) {
      double mx[SPL];
      bool first = true;
      for (int k = 0; k < Kt; ++k) {
        const double *cp = tgts + (size_t)k * a.tgt_stride;
        if (cp[pk_oSc(D)] > 0.0) {
          double q[SPL];
          whiten<D, SPL>(cp, x, q);
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            const double ld = -0.5 * (D * kLn2Pi + q[s]) - cp[pk_oSc(D) + 1];
            grow[s][k] = ld;
            if (first || mx[s] < ld) mx[s] = ld;
          }
          first = false;
        }
      }
      double lpos[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) lpos[s] = 0.0;
      for (int k = 0; k < Kt; ++k) {
        const double al = tgts[(size_t)k * a.tgt_stride + pk_oSc(D)];
        if (al > 0.0) {
#pragma unroll
          for (int s = 0; s < SPL; ++s) lpos[s] = lpos[s] + exp(grow[s][k] - mx[s] + kDynMax) * al;
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) post[s] = mx[s] + log(lpos[s]) - kDynMax;
    } else {  // PMCGPU_TGT_HOST
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        if (i < a.N) { post[s] = a.post_in[i]; tok[s] = a.ok_in ? (a.ok_in[i] != 0) : true; }
      }
    }

    // ---------------- proposal log-density -----------------------------------------------------------------
    double mx[SPL], lpos[SPL];
    {
      bool first = true;
      for (int k = 0; k < K; ++k) {
        const double *cp = mixs + (size_t)k * a.comp_stride;
        if (cp[pk_oSc(D)] > 0.0) {
          double q[SPL];
          whiten<D, SPL>(cp, x, q);
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            const double ld = -0.5 * (D * kLn2Pi + q[s]) - cp[pk_oSc(D) + 1];
            grow[s][k] = ld;
            if (first || mx[s] < ld) mx[s] = ld;
          }
          first = false;
        } else {
#pragma unroll
          for (int s = 0; s < SPL; ++s) grow[s][k] = -INFINITY;
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) lpos[s] = 0.0;
      for (int k = 0; k < K; ++k) {
        const double al = mixs[(size_t)k * a.comp_stride + pk_oSc(D)];
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
          double e = 0.0;
          if (al > 0.0) { e = exp(grow[s][k] - mx[s] + kDynMax) * al; lpos[s] = lpos[s] + e; }
          grow[s][k] = e;
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
        if (rloc > maxr_lane) maxr_lane = rloc;  // NaN never wins (pmc.c:552)
      }
      // exp(lw - maxW + 500) is non-finite only for NaN / +inf log-weights (pmc.c:663-669)
      wfin[s] = reached && !isnan(lw[s]) && !(lw[s] == INFINITY);
      if (reached && lw[s] == INFINITY) atomicAdd(&a.counters[CNT_POSINF], 1ull);
      if (wfin[s] && lw[s] > tmax) tmax = lw[s];
    }
    tmax = warp_max_nan_ignoring(tmax);
    if (tmax > m_run) {  // online rescaling of everything accumulated so far
      const double sc = (m_run == -INFINITY) ? 0.0 : exp(m_run - tmax);
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < RK; ++i)
#pragma unroll
          for (int j = 0; j < RE; ++j) acc[r][i][j] *= sc;
      S_lane *= sc;
      m_run = tmax;
    }
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      w[s] = (wfin[s] && lw[s] > -INFINITY) ? exp(lw[s] - m_run) : 0.0;
      S_lane += w[s];
      if (wfin[s]) { ++nok_lane; atomicAdd(&s_count[kdraw[s] & 63], 1u); }
    }
    if (a.do_stats) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const double rs = (w[s] != 0.0) ? w[s] / lpos[s] : 0.0;
        for (int k = 0; k < K; ++k) grow[s][k] = (rs != 0.0) ? grow[s][k] * rs : 0.0;  // g_ik = w_i alpha_k phi_k(x_i)/q(x_i); NaN densities of flagged-out samples must not leak
        for (int k = K; k < GS; ++k) grow[s][k] = 0.0;
      }

      // ---------------- phase 2: statistics ------------------------------------------------------------------
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        {  // this lane's sample -> one row of the P tile
          double xc[D];
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) xc[i2] = (w[s] != 0.0) ? x[s][i2] - a.pivot[i2] : 0.0;
          double *prow = ptile + (size_t)lane * PS;
          double pe[NCH * RE];
          pe[0] = 1.0;
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2) pe[1 + i2] = xc[i2];
          int e = 1 + D;
#pragma unroll
          for (int i2 = 0; i2 < D; ++i2)
#pragma unroll
            for (int j = i2; j < D; ++j) pe[e++] = xc[i2] * xc[j];
#pragma unroll
          for (int e2 = E; e2 < NCH * RE; ++e2) pe[e2] = 0.0;
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int j = 0; j < RE; j += 2)
              *reinterpret_cast<double2 *>(prow + ch * CSP + j) = make_double2(pe[ch * RE + j], pe[ch * RE + j + 1]);
        }
        __syncwarp();
        const double *gbase = gtile + (size_t)(s * 32) * GS;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            double gk[RK], pv[RE];
            if (RK == 2) {
              const double2 t = *reinterpret_cast<const double2 *>(gbase + (size_t)j * GS + p2_g[r]);
              gk[0] = t.x; gk[RK - 1] = t.y;
            } else {
#pragma unroll
              for (int i = 0; i < RK; ++i) gk[i] = gbase[(size_t)j * GS + p2_g[r] + i];
            }
#pragma unroll
            for (int e2 = 0; e2 < RE; e2 += 2) {
              const double2 t = *reinterpret_cast<const double2 *>(ptile + (size_t)j * PS + p2_p[r] + e2);
              pv[e2] = t.x; pv[e2 + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < RK; ++i)
#pragma unroll
              for (int e2 = 0; e2 < RE; ++e2) acc[r][i][e2] = fma(gk[i], pv[e2], acc[r][i][e2]);
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
  for (int w2 = 0; w2 < FUSED_WARPS; ++w2) { if (s_m[w2] > Mb) Mb = s_m[w2]; if (s_r[w2] > Rb) Rb = s_r[w2]; }
  const double scale = (m_run == -INFINITY) ? 0.0 : exp(m_run - Mb);
  double *comb = warp0;  // aliases the (now idle) warp tiles: ACC doubles
  for (int w2 = 0; w2 < FUSED_WARPS; ++w2) {
    if (wid == w2) {
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < RK; ++i)
#pragma unroll
          for (int j = 0; j < RE; ++j) {
            const int idx = ((r * 32 + lane) * RK + i) * RE + j;
            comb[idx] = (w2 == 0 ? 0.0 : comb[idx]) + acc[r][i][j] * scale;
          }
    }
    __syncthreads();
  }
  double *out = a.partials + (size_t)blockIdx.x * a.partial_len;
  for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = comb[idx];
  if (threadIdx.x == 0) {
    double Sb = 0.0;
    for (int w2 = 0; w2 < FUSED_WARPS; ++w2) Sb += (s_m[w2] == -INFINITY) ? 0.0 : s_S[w2] * exp(s_m[w2] - Mb);
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

// Combine: canonical statistics stats[k*E + e], rescaled to the global maximum M.
__global__ void k_fused_combine(const double *__restrict__ partials, int nblocks, int partial_len, int acc_len, int K,
                                int E, int RK, int RE, int NCH, double *__restrict__ out) {
  extern __shared__ double sc[];  // per-block scale
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
    const int u = (k / RK) * NCH + e / RE;
    const int idx = ((u / 32 * 32 + (u & 31)) * RK + (k % RK)) * RE + (e % RE);
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

// ---- instantiation table ---------------------------------------------------------------------------------------
template <int D, int RK, int RE, int NR, int SPL>
static int launch_one(const FusedPlan &p, const FusedArgs &a, cudaStream_t s) {
  auto kern = k_fused<D, RK, RE, NR, SPL>;
  // per device and per mixture size: set on every launch (a few hundred nanoseconds against a multi-millisecond kernel)
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess) return -1;
  kern<<<p.grid, p.threads, p.smem_bytes, s>>>(a);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// (D, RK, RE, NR, SPL): E = 1 + D + D(D+1)/2 statistics per component are cut into chunks of RE; a lane owns RK
// components x one chunk; NR rounds of 32 lanes must cover ceil(K/RK) * ceil(E/RE) units.
#define PMC_FUSED_TABLE(X) \
  X(10, 2, 12, 1, 1)       \
  X(10, 2, 12, 1, 2)       \
  X(10, 2, 12, 2, 1)       \
  X(8, 2, 12, 1, 1)        \
  X(8, 2, 12, 2, 1)        \
  X(6, 2, 10, 1, 1)        \
  X(6, 2, 10, 2, 1)        \
  X(5, 2, 8, 1, 1)         \
  X(5, 2, 8, 1, 2)         \
  X(4, 2, 8, 1, 1)         \
  X(4, 2, 8, 2, 1)         \
  X(3, 2, 6, 1, 1)         \
  X(3, 2, 6, 2, 1)         \
  X(2, 2, 6, 1, 1)         \
  X(2, 2, 6, 1, 2)

FusedPlan fused_plan(int d, int K, int Kt, int sm_count, int want_spl) {
  FusedPlan best{};
  best.valid = 0;
  if (K > 64 || K < 1) return best;
#define X(D_, RK_, RE_, NR_, SPL_)                                                                        \
  if (!best.valid && d == D_ && SPL_ == want_spl) {                                                       \
    using C = FusedCfg<D_, RK_, RE_, NR_, SPL_>;                                                          \
    const int units = ((K + RK_ - 1) / RK_) * C::NCH;                                                     \
    if (units <= 32 * NR_) {                                                                              \
      const int GS = ev2((K > Kt ? K : Kt) + (RK_ - 1));                                                  \
      const size_t dbl = (size_t)K * pk_stride(D_) + (size_t)Kt * pk_stride(D_) + ev2(K) +                \
                         (size_t)FUSED_WARPS * (32 * C::PS + C::TS * GS);                                 \
      if (dbl * 8 <= 227 * 1024 && (size_t)C::ACC <= (size_t)FUSED_WARPS * (32 * C::PS + C::TS * GS)) {   \
        best.valid = 1; best.D = D_; best.RK = RK_; best.RE = RE_; best.NR = NR_; best.SPL = SPL_;        \
        best.grid = sm_count; best.threads = FUSED_WARPS * 32; best.smem_bytes = dbl * 8;                 \
        best.partial_len = C::ACC + 4; best.nstat = K * C::E;                                             \
      }                                                                                                   \
    }                                                                                                     \
  }
  PMC_FUSED_TABLE(X)
#undef X
  return best;
}

int launch_fused(const FusedPlan &p, const FusedArgs &a, cudaStream_t s) {
#define X(D_, RK_, RE_, NR_, SPL_) \
  if (p.D == D_ && p.RK == RK_ && p.RE == RE_ && p.NR == NR_ && p.SPL == SPL_) return launch_one<D_, RK_, RE_, NR_, SPL_>(p, a, s);
  PMC_FUSED_TABLE(X)
#undef X
  return -1;
}

int launch_fused_combine(const FusedPlan &p, const double *partials, int K, double *stats_out, cudaStream_t s) {
  const int E = 1 + p.D + p.D * (p.D + 1) / 2;
  const int NCH = (E + p.RE - 1) / p.RE;
  k_fused_combine<<<1, 256, p.grid * sizeof(double), s>>>(partials, p.grid, p.partial_len, p.partial_len - 4, K, E,
                                                            p.RK, p.RE, NCH, stats_out);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace pmcb200
