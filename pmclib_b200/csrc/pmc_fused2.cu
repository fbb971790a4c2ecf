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

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem);
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem);

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

// MODE 0: draw/weights only; 1: + responsibilities r[k][i] = alpha_k phi_k(x_i)/q(x_i) to global memory (for k_stats2);
// 2: statistics fused in the same kernel (per-warp DMMA accumulators); 3: draw only (X, indices; integer-heavy, small code).
template <int D, int K, int KT, int SPL, int MODE, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_fused2(const FusedArgs a) {
  using C = F2Cfg<D, K, KT, SPL>;
  constexpr int F2_WARPS = WARPS;
  constexpr int E = C::E, NT = C::NT, MT = C::MT, PS = C::PS, GS = C::GS, TS = C::TS, ST = C::ST;
  static_assert(K * ST <= F2_CMIX && KT * ST <= F2_CTGT, "constant memory budget");
  static_assert(TS * D <= 32 * PS, "X staging must fit in the P tile");
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  double *mixs = smem;                      // copy of the packed mixture for the draw (per-lane component)
  double *cws = mixs + K * ST;
  double2 *zts = reinterpret_cast<double2 *>(cws + f2_ev2(K));  // ziggurat table (x_i, x_{i+1}/x_i), 64 KB
  double *warp0 = cws + f2_ev2(K) + 2 * kZigLayers;
  constexpr int WD = (MODE == 2) ? C::WARP_DOUBLES : f2_ev2(TS * D);
  double *ptile = warp0 + (size_t)wid * WD;
  double *gtile = ptile + 32 * PS;
  __shared__ unsigned int s_count[64];
  __shared__ double s_m[F2_WARPS], s_S[F2_WARPS], s_r[F2_WARPS];

  for (int i = threadIdx.x; i < K * ST; i += blockDim.x) mixs[i] = c_mix[i];
  for (int i = threadIdx.x; i < kZigLayers; i += blockDim.x) zts[i] = a.zig[i];
  if (threadIdx.x < 64) s_count[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {  // cumulative weights exactly as mvdens.c:748-750 (sequential sum)
    double c = 0.0;
    for (int k = 0; k < K; ++k) { c = (k == 0) ? mixs[f2_oSc(D)] : c + mixs[k * ST + f2_oSc(D)]; cws[k] = c; }
    // the mixture may have fewer components than this instantiation (a.K <= K, the rest packed as zeros): ranindx stops at
    // the last component of the MIXTURE (mvdens.c:776, "index < K - 1"), so the pick never looks past cw[a.K - 2]
    for (int k = (a.K > 0 ? a.K - 1 : 0); k < K; ++k) cws[k] = INFINITY;
  }
  // zero the padding rows/columns of this warp's tiles once (entries E..8NT-1, components K..8MT-1)
  for (int i = lane; i < WD; i += 32) ptile[i] = 0.0;
  __syncthreads();

  constexpr int AM = (MODE == 2) ? MT : 1, AN = (MODE == 2) ? NT : 1;
  double acc[AM][AN][2];
#pragma unroll
  for (int m = 0; m < AM; ++m)
#pragma unroll
    for (int n = 0; n < AN; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  double m_run = -INFINITY, S_lane = 0.0, maxr_lane = -INFINITY;
  unsigned int nok_lane = 0, rej_lane = 0, fail_lane = 0;

  const long ntiles = (a.N + TS - 1) / TS;
  const Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const int fr = lane >> 2, fc = lane & 3;  // fragment row (component / entry within a tile), k-step column (sample)

  for (long tile = (long)blockIdx.x * F2_WARPS + wid; tile < ntiles; tile += (long)gridDim.x * F2_WARPS) {
    const long tbase = tile * TS;
    double x[SPL][D];
    bool valid[SPL];
    int kdraw[SPL];

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
            double g[D + 1];
            uint32_t spare;
            zig_normals<D>(ph, ilo, ihi, cb, a.iter, zts, g, spare);
            // uniform for the component pick: the 3 spare bits of each of the D candidates when that gives >= 30 bits
            // (the reference's own gsl_rng_uniform has 32-bit resolution), else one more Philox call
            double zc;
            if (3 * D >= 30) {
              zc = (double)(spare & 0x3fffffffu) * (1.0 / 1073741824.0);
            } else {
              const uint4 r0 = ph(ilo, ihi, cb, a.iter);
              zc = u01_co(r0.x, r0.y);
            }
            // first index with cw >= z == number of entries below z (cw is non-decreasing), mvdens.c:776
            int k = 0;
#pragma unroll
            for (int kk = 0; kk < K - 1; ++kk) k += (cws[kk] < zc) ? 1 : 0;
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
              // thr - c*x < 0 (parabox.c:125-131) for c = -1 / +1: IEEE addition keeps the sign of the exact sum, so the
              // decision equals a direct comparison with the (negated) threshold
              if (a.has_box) inb = inb && !(xv < -a.lo_thr[i2]) && !(xv > a.hi_thr[i2]);
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
          if ((D & 1) == 0) {
            double2 *xo = reinterpret_cast<double2 *>(a.X + (size_t)i * D);
#pragma unroll
            for (int i2 = 0; i2 < D; i2 += 2) xo[i2 / 2] = make_double2(x[s][i2], x[s][i2 + 1]);
          } else {
#pragma unroll
            for (int i2 = 0; i2 < D; ++i2) a.X[(size_t)i * D + i2] = x[s][i2];
          }
        }
      }
    } else {
      // samples already in HBM: the tile was prefetched into this warp's stage with cp.async (LDGSTS) while the previous
      // tile was being computed, so the first FP64 instruction does not wait for a global load
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        valid[s] = i < a.N;
        kdraw[s] = 0;
#pragma unroll
        for (int i2 = 0; i2 < D; ++i2) x[s][i2] = 0.0;
        if (valid[s]) {  // each lane owns 8D contiguous bytes of X (16-byte aligned rows when D is even)
          if ((D & 1) == 0) {
            const double2 *xi = reinterpret_cast<const double2 *>(a.X + (size_t)i * D);
#pragma unroll
            for (int i2 = 0; i2 < D; i2 += 2) { const double2 t = xi[i2 / 2]; x[s][i2] = t.x; x[s][i2 + 1] = t.y; }
          } else {
#pragma unroll
            for (int i2 = 0; i2 < D; ++i2) x[s][i2] = a.X[(size_t)i * D + i2];
          }
          kdraw[s] = (int)a.indices[i];
        }
      }
    }

    if (MODE == 3) continue;  // draw-only kernel: the FP64 part runs as a second launch on the stored samples

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
#pragma unroll
      for (int s = 0; s < SPL; ++s) mx[s] = -INFINITY;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        double q[SPL];
        whiten_c<D, SPL, true>(k * ST, x, q);
        const bool alive = c_tgt[k * ST + f2_oSc(D)] > 0.0;
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
          tl[s][k] = -0.5 * (D * kLn2Pi + q[s]) - c_tgt[k * ST + f2_oSc(D) + 1];
          if (alive && mx[s] < tl[s][k]) mx[s] = tl[s][k];
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        double arg[KT > 0 ? KT : 1], ex[KT > 0 ? KT : 1];
#pragma unroll
        for (int k = 0; k < KT; ++k) arg[k] = (c_tgt[k * ST + f2_oSc(D)] > 0.0) ? tl[s][k] - mx[s] + kDynMax : -INFINITY;
        exp_batch<(KT > 0 ? KT : 1)>(arg, ex);
        double lp = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (c_tgt[k * ST + f2_oSc(D)] > 0.0) lp = lp + ex[k] * c_tgt[k * ST + f2_oSc(D)];
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
    // One straight-line block for all K components: dead components (weight <= 0) carry neutral parameters in the
    // pack (fused2 host packing) and are excluded by selects, so ptxas can schedule constant loads and FMA chains
    // across component boundaries (a branch per component cost a constant-load latency bubble each, ncu r01_v4).
    double ev[SPL][K], mx[SPL], lpos[SPL];
    {
#pragma unroll
      for (int s = 0; s < SPL; ++s) mx[s] = -INFINITY;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double q[SPL];
        whiten_c<D, SPL, false>(k * ST, x, q);
        const bool alive = c_mix[k * ST + f2_oSc(D)] > 0.0;
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
          ev[s][k] = -0.5 * (D * kLn2Pi + q[s]) - c_mix[k * ST + f2_oSc(D) + 1];
          if (alive && mx[s] < ev[s][k]) mx[s] = ev[s][k];
        }
      }
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        // exponentials in sub-batches of EB: each batch keeps 7 registers per element live, so the batch size sets the
        // register peak of the kernel (EB = 5: fits 128 registers without spills, enough chains to hide the FMA latency)
        constexpr int EB = (K % 5 == 0) ? 5 : ((K % 4 == 0) ? 4 : ((K % 3 == 0) ? 3 : 1));
#pragma unroll
        for (int k0 = 0; k0 < K; k0 += EB) {
          double arg[EB], ex[EB];
#pragma unroll
          for (int j = 0; j < EB; ++j)
            arg[j] = (c_mix[(k0 + j) * ST + f2_oSc(D)] > 0.0) ? ev[s][k0 + j] - mx[s] + kDynMax : -INFINITY;
          exp_batch<EB>(arg, ex);
#pragma unroll
          for (int j = 0; j < EB; ++j) ev[s][k0 + j] = ex[j];
        }
        lpos[s] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {  // same summation order as mvdens.c:814-819
          const double e = (c_mix[k * ST + f2_oSc(D)] > 0.0) ? ev[s][k] * c_mix[k * ST + f2_oSc(D)] : 0.0;
          lpos[s] = lpos[s] + e;
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
      for (int m = 0; m < AM; ++m)
#pragma unroll
        for (int n = 0; n < AN; ++n) { acc[m][n][0] *= sc; acc[m][n][1] *= sc; }
      S_lane *= sc;
      m_run = tmax;
    }
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      w[s] = 0.0;
      if (MODE == 2) {
        w[s] = (wfin[s] && lw[s] > -INFINITY) ? exp_one(lw[s] - m_run) : 0.0;
        S_lane += w[s];
      }
      if (wfin[s]) { ++nok_lane; atomicAdd(&s_count[kdraw[s] & 63], 1u); }
    }

    if (MODE == 1) {  // responsibilities, component-major so that every store is one coalesced 256-byte row
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const long i = tbase + s * 32 + lane;
        if (i < a.N) {
          const double rs = 1.0 / lpos[s];
#pragma unroll
          for (int k = 0; k < K; ++k) a.resp[(size_t)k * a.N + i] = ev[s][k] * rs;
        }
      }
    }
    // ---------------- phase 2: statistics on the FP64 tensor path ---------------------------------------------------
    if (MODE == 2 && a.do_stats) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        {
          const double rs = (w[s] != 0.0) ? w[s] / lpos[s] : 0.0;
          double *grow = gtile + lane * GS;
#pragma unroll
          for (int k = 0; k < K; ++k) grow[k] = (rs != 0.0) ? ev[s][k] * rs : 0.0;  // g_ik = w_i alpha_k phi_k(x_i) / q(x_i)
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
  double *out = a.partials + (size_t)blockIdx.x * a.partial_len;
  if (MODE == 2) {
    double *comb = warp0;
    for (int w2 = 0; w2 < F2_WARPS; ++w2) {
      if (wid == w2) {
#pragma unroll
        for (int m = 0; m < AM; ++m)
#pragma unroll
          for (int n = 0; n < AN; ++n)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int idx = ((m * NT + n) * 32 + lane) * 2 + c;
              comb[idx] = (w2 == 0 ? 0.0 : comb[idx]) + acc[m][n][c] * scale;
            }
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = comb[idx];
  } else {
    (void)scale;
    for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = 0.0;
  }
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
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nstat; t += gridDim.x * blockDim.x) {
    const int k = t / E, e = t - k * E;
    const int m = k >> 3, r = k & 7, n = e >> 3, cc = e & 7;
    const int idx = ((m * NT + n) * 32 + (r * 4 + (cc >> 1))) * 2 + (cc & 1);  // C fragment: row = lane/4, cols 2(lane%4)+{0,1}
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * partial_len + idx] * sc[b];
    out[t] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
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

// ---- statistics kernel of the split pipeline ----------------------------------------------------------------------
// C[k][e] += g[i][k] * P[i][e] over all samples, g = exp(lw_i - M) * r[k][i], P = products of the augmented vector
// (1, x - pivot).  Pure DMMA work: per warp and 32-sample tile 8 k-steps x MT x NT mma.m8n8k4.f64; the B fragments are
// formed on the fly from the 11-double augmented rows (two conflict-free LDS + one DMUL), so the tile is 7 KB per warp
// and 16 warps fit per SM.
struct StatsArgs {
  const double *X, *lw, *resp;
  const short *flg;
  const double *M;  // device scalar: global maximum log-weight (this launch)
  double pivot[FUSED_MAXD];
  long N;
  double *partials;
  int partial_len;
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

template <int D, int K>
struct S2Cfg {
  static constexpr int XS = (D + 1) | 1;                          // odd stride of the augmented row (1, x~)
  static constexpr int GS = F2Cfg<D, K, 0, 1>::GS;
  static constexpr int STAGE = 32 * D + 32 * K + 32 + 8;          // raw X rows | resp rows (k-major) | lw | flg (32 shorts)
  static constexpr int WARP_DOUBLES = 32 * XS + 32 * GS + STAGE;  // compute tiles + one prefetch stage
};

template <int D, int K>
__global__ void __launch_bounds__(256, 2) k_stats2(const StatsArgs a) {
  using C = F2Cfg<D, K, 0, 1>;
  using S2 = S2Cfg<D, K>;
  constexpr int NT = C::NT, MT = C::MT, GS = C::GS, XS = S2::XS;
  constexpr int WARPS = 8;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *xh = smem + (size_t)wid * S2::WARP_DOUBLES;
  double *gt = xh + 32 * XS;
  double *stx = gt + 32 * GS;        // [32][D] raw sample rows of the NEXT tile (LDGSTS prefetch)
  double *str = stx + 32 * D;        // [K][32] responsibilities
  double *stw = str + 32 * K;        // [32] log-weights
  short *stf = reinterpret_cast<short *>(stw + 32);  // [32] flags
  __shared__ double s_S[WARPS];
  for (int i = lane; i < S2::WARP_DOUBLES; i += 32) xh[i] = 0.0;
  __syncwarp();
  const int fr = lane >> 2, fc = lane & 3;
  // which two augmented coordinates multiply to entry e = 8n + fr
  int oa[NT], ob[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int e = 8 * n + fr;
    int pa = 0, pb = 0;
    if (e >= 1 && e <= D) pb = e;
    else if (e > D && e < C::E) {
      int q = e - 1 - D, r = 0;
      while (q >= D - r) { q -= D - r; ++r; }
      pa = 1 + r;
      pb = 1 + r + q;
    }
    oa[n] = pa;
    ob[n] = pb;
  }
  double acc[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  double S_lane = 0.0;
  const double M = a.M[0];
  const long ntiles = (a.N + 31) / 32;
  const long tstep = (long)gridDim.x * WARPS;

  // asynchronous copy of one tile's inputs into the stage (cp.async = SASS LDGSTS): overlaps with the DMMA loop
  auto prefetch = [&](long tile) {
    const long base = tile * 32;
    const int tn = (a.N - base) < 32 ? (int)(a.N - base) : 32;
    if ((D & 1) == 0) {
      for (int c = lane; c < tn * D / 2; c += 32) cp_async16(stx + 2 * c, a.X + (size_t)base * D + 2 * c);
    } else {
      for (int c = lane; c < tn * D; c += 32) cp_async8(stx + c, a.X + (size_t)base * D + c);
    }
    if (lane < tn) {
#pragma unroll
      for (int k = 0; k < K; ++k) cp_async8(str + k * 32 + lane, a.resp + (size_t)k * a.N + base + lane);
      cp_async8(stw + lane, a.lw + base + lane);
    }
    if (2 * lane + 1 < tn) cp_async4(stf + 2 * lane, a.flg + base + 2 * lane);
    else if (2 * lane < tn) stf[2 * lane] = a.flg[base + 2 * lane];
    asm volatile("cp.async.commit_group;\n" ::);
  };

  long tile = (long)blockIdx.x * WARPS + wid;
  if (tile < ntiles) prefetch(tile);
  for (; tile < ntiles; tile += tstep) {
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncwarp();
    const long i = tile * 32 + lane;
    double w = 0.0;
    double xr[D];
#pragma unroll
    for (int j = 0; j < D; ++j) xr[j] = 0.0;
    if (i < a.N) {
      const double lw = stw[lane];
      const bool wfin = stf[lane] != 0 && !isnan(lw) && !(lw == INFINITY);
      w = (wfin && lw > -INFINITY) ? exp_one(lw - M) : 0.0;
      if (w != 0.0) {
#pragma unroll
        for (int j = 0; j < D; ++j) xr[j] = stx[lane * D + j] - a.pivot[j];
      }
    }
    S_lane += w;
    double *xrow = xh + lane * XS, *grow = gt + lane * GS;
    xrow[0] = 1.0;
#pragma unroll
    for (int j = 0; j < D; ++j) xrow[1 + j] = xr[j];
#pragma unroll
    for (int k = 0; k < K; ++k) grow[k] = (w != 0.0) ? w * str[k * 32 + lane] : 0.0;
    __syncwarp();  // tiles complete, stage consumed
    if (tile + tstep < ntiles) prefetch(tile + tstep);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int smp = 16 * h + t + 4 * fc;
        double af[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) af[m] = gt[smp * GS + 8 * m + fr];
        const double *xs = xh + smp * XS;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const double bf = xs[oa[n]] * xs[ob[n]];
#pragma unroll
          for (int m = 0; m < MT; ++m) dmma884(acc[m][n][0], acc[m][n][1], af[m], bf);
        }
      }
    __syncwarp();
  }
  // block combine in warp order
  __syncthreads();
  {
    const double Sw = warp_sum(S_lane);
    if (lane == 0) s_S[wid] = Sw;
  }
  double *comb = smem;
  for (int w2 = 0; w2 < WARPS; ++w2) {
    __syncthreads();
    if (wid == w2) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int idx = ((m * NT + n) * 32 + lane) * 2 + c;
            comb[idx] = (w2 == 0 ? 0.0 : comb[idx]) + acc[m][n][c];
          }
    }
  }
  __syncthreads();
  double *out = a.partials + (size_t)blockIdx.x * a.partial_len;
  for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = comb[idx];
  if (threadIdx.x == 0) {
    double Sb = 0.0;
    for (int w2 = 0; w2 < WARPS; ++w2) Sb += s_S[w2];
    out[C::ACC + 0] = M;
    out[C::ACC + 1] = Sb;
    out[C::ACC + 2] = -INFINITY;
    out[C::ACC + 3] = 0.0;
  }
}

// global maxima from the block partials of a MODE 0/1 launch: out[0] = M, out[1] = maxR
__global__ void k_f2_max(const double *__restrict__ partials, int nblocks, int partial_len, int acc_len, double *out) {
  double M = -INFINITY, R = -INFINITY;
  for (int b = threadIdx.x; b < nblocks; b += 32) {
    const double mb = partials[(size_t)b * partial_len + acc_len], rb = partials[(size_t)b * partial_len + acc_len + 2];
    if (mb > M) M = mb;
    if (rb > R) R = rb;
  }
  M = warp_max_nan_ignoring(M);
  R = warp_max_nan_ignoring(R);
  if (threadIdx.x == 0) { out[0] = M; out[1] = R; }
}

// ---- instantiation table: (D, K, KT, SPL) ------------------------------------------------------------------------
#define PMC_FUSED2_TABLE(X) \
  X(10, 10, 0, 1)           \
  X(10, 20, 0, 1)           \
  X(10, 5, 0, 1)            \
  X(5, 20, 4, 1)            \
  X(5, 10, 0, 1)            \
  X(2, 9, 0, 1)

// launch shapes of the draw/weights kernel (MODE 0/1): variant -> (warps per block, min blocks per SM)
template <int D, int K, int KT, int SPL, int MODE, int WARPS, int MINB>
static int launch_k1(const FusedPlan &p, const FusedArgs &a, int *nblocks_out, cudaStream_t s) {
  auto kern = k_fused2<D, K, KT, SPL, MODE, WARPS, MINB>;
  using C = F2Cfg<D, K, KT, SPL>;
  const size_t wd = (MODE == 2) ? C::WARP_DOUBLES : f2_ev2(C::TS * D);
  const size_t smem = ((size_t)K * C::ST + f2_ev2(K) + 2 * kZigLayers + (size_t)WARPS * wd) * sizeof(double);
  static std::atomic<unsigned long long> smem_set{0};
  if (!set_max_smem_once(kern, smem, smem_set)) return -1;
  static std::atomic<int> nb_cached{0};  // the occupancy query costs several microseconds: once per instantiation
  int nb = nb_cached.load(std::memory_order_relaxed);
  if (nb == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, smem) != cudaSuccess || nb < 1) nb = 1;
    if (nb > 5) nb = 5;
    nb_cached.store(nb, std::memory_order_relaxed);
  }
  *nblocks_out = p.grid * nb;
  kern<<<p.grid * nb, WARPS * 32, smem, s>>>(a);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int D, int K, int KT, int SPL>
static int launch_one2(const FusedPlan &p, const FusedArgs &a, const double *h_mix, const double *h_tgt, int mode,
                       int *nb, cudaStream_t s) {
  if (h_mix && cudaMemcpyToSymbolAsync(c_mix, h_mix, sizeof(double) * K * f2_stride(D), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (KT > 0 && h_tgt && cudaMemcpyToSymbolAsync(c_tgt, h_tgt, sizeof(double) * KT * f2_stride(D), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  // launch shapes (warps per block, min blocks per SM) chosen by A/B timing on B200
  static_assert(SPL == 1, "two samples per lane measured slower (register spills); only SPL = 1 is instantiated");
  if (mode == 3) return launch_k1<D, K, KT, SPL, 3, 8, 2>(p, a, nb, s);
  if (mode == 2) return launch_k1<D, K, KT, SPL, 2, 8, 1>(p, a, nb, s);
  if (mode == 0) return launch_k1<D, K, KT, SPL, 0, 4, 4>(p, a, nb, s);
  return launch_k1<D, K, KT, SPL, 1, 4, 4>(p, a, nb, s);
}

template <int D, int K>
static int launch_stats2(const FusedPlan &p, const StatsArgs &a, cudaStream_t s) {
  auto kern = k_stats2<D, K>;
  using C = F2Cfg<D, K, 0, 1>;
  size_t smem = (size_t)8 * S2Cfg<D, K>::WARP_DOUBLES * sizeof(double);
  if (smem < (size_t)C::ACC * sizeof(double)) smem = (size_t)C::ACC * sizeof(double);
  static std::atomic<unsigned long long> smem_set{0};
  if (!set_max_smem_once(kern, smem, smem_set)) return -1;
  kern<<<p.grid * 2, 256, smem, s>>>(a);
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
    const size_t dbl = (size_t)K_ * C::ST + f2_ev2(K_) + 2 * kZigLayers + (size_t)8 * C::WARP_DOUBLES;   \
    best.valid = 2; best.D = D_; best.RK = K_; best.RE = KT_; best.SPL = SPL_;                           \
    best.NR = (dbl * 8 <= 227 * 1024) ? 1 : 0; /* does the single-kernel mode (MODE 2) fit in shared memory? */ \
    best.grid = sm_count; best.threads = 256; best.smem_bytes = dbl * 8;                                 \
    best.partial_len = C::ACC + 4; best.nstat = K_ * C::E;                                               \
  }
  PMC_FUSED2_TABLE(X)
#undef X
  return best;
}

// the draw-only kernel (MODE 3) of the smallest instantiation with at least K components
FusedPlan fused2_plan_draw(int d, int K, int sm_count) {
  FusedPlan best{};
  best.valid = 0;
#define X(D_, K_, KT_, SPL_)                                                                  \
  if (d == D_ && K <= K_ && SPL_ == 1 && (!best.valid || K_ < best.RK)) {                     \
    using C = F2Cfg<D_, K_, KT_, SPL_>;                                                       \
    best.valid = 2; best.D = D_; best.RK = K_; best.RE = KT_; best.SPL = SPL_; best.NR = 0;   \
    best.grid = sm_count; best.threads = 256; best.smem_bytes = 0;                            \
    best.partial_len = C::ACC + 4; best.nstat = K_ * C::E;                                    \
  }
  PMC_FUSED2_TABLE(X)
#undef X
  return best;
}

int fused2_max_blocks(const FusedPlan &p) { return p.grid * 5; }

int launch_fused2(const FusedPlan &p, const FusedArgs &a, const double *h_mix, const double *h_tgt, int mode,
                  int *nblocks_out, cudaStream_t s) {
#define X(D_, K_, KT_, SPL_) \
  if (p.D == D_ && p.RK == K_ && p.RE == KT_ && p.SPL == SPL_) return launch_one2<D_, K_, KT_, SPL_>(p, a, h_mix, h_tgt, mode, nblocks_out, s);
  PMC_FUSED2_TABLE(X)
#undef X
  return -1;
}

int launch_fused2_max(const FusedPlan &p, const double *partials, int nblocks, double *out2, cudaStream_t s) {
  k_f2_max<<<1, 32, 0, s>>>(partials, nblocks, p.partial_len, p.partial_len - 4, out2);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_fused2_stats(const FusedPlan &p, const double *X, const double *lw, const double *resp, const short *flg,
                        const double *Mdev, const double *pivot, long N, double *partials, cudaStream_t s) {
  StatsArgs a{};
  a.X = X; a.lw = lw; a.resp = resp; a.flg = flg; a.M = Mdev; a.N = N; a.partials = partials; a.partial_len = p.partial_len;
  for (int j = 0; j < FUSED_MAXD; ++j) a.pivot[j] = pivot[j];
#define X(D_, K_, KT_, SPL_) \
  if (p.D == D_ && p.RK == K_ && p.RE == KT_ && p.SPL == SPL_) return launch_stats2<D_, K_>(p, a, s);
  PMC_FUSED2_TABLE(X)
#undef X
  return -1;
}

int launch_fused2_combine(const FusedPlan &p, const double *partials, int nblocks, double *stats_out, cudaStream_t s) {
  const int E = 1 + p.D + p.D * (p.D + 1) / 2;
  const int NT = (E + 7) / 8;
  const int nstat = p.RK * E;
  k_fused2_combine<<<(nstat + 63) / 64, 64, nblocks * sizeof(double), s>>>(partials, nblocks, p.partial_len, p.partial_len - 4, p.RK, E, NT, stats_out);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace pmcb200
