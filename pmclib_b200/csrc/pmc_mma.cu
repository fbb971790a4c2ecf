// pmc_mma.cu — the tensor path: whitening, Rao-Blackwell statistics and the grouped draw as dense FP64 contractions
// (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) for every shape d <= 128 that has no fused small-d instantiation — above all the
// large dimensions (configs[3] d = 30, configs[4] d = 128).
//
// Why a separate path: with one sample per lane (pmc_fused2.cu) a d x d triangular factor no longer fits the constant
// bank / registers, and the generic kernels (one thread per sample, factor streamed from L2) run at ~4 % of the FP64
// peak.  For large d the whitening  y = L_k^-1 (x - mu_k)  of MANY samples against ONE component is a true matrix
// product, so it is tiled the GEMM way:
//
//   k_q_mma<NT>      q[k][i] = | Linv_k (x_i - mu_k) |^2 .  M = 8 samples per DMMA, N = 8 output rows, K = 4 input
//                    columns; only the tiles on or below the diagonal of Linv_k are issued (NT(NT+1) of 2 NT^2).  The
//                    sample fragments stay in registers for all components; one component's factor (already laid out
//                    in B-fragment order by the host) is staged in shared memory by cp.async while the previous one is
//                    being used, so every LDS is a conflict-free 256-byte row.
//   k_weights_q      literal mix_mvdens_log_pdf / importance-weight tail (mvdens.c:784-825, pmc.c:528-590) from q.
//   k_resp_q         responsibilities g[k][i] = w_i alpha_k phi_k(x_i) / q(x_i)   (pmc.c:1384-1390, scale-free form)
//   k_stats_mma<NT>  S_k = sum_i g[k][i] (x_i - c)(x_i - c)^T, s_k = sum_i g[k][i] (x_i - c), W_k = sum_i g[k][i]
//                    about one pivot c: per component a (d x n)(n x d) product, lower tiles only, warps own row-tile
//                    pairs (r, NT-1-r) so every warp issues NT+1 DMMAs per 4 samples.
//
// The inverse factors are formed on the host in extended precision (mma_pack_mixture); a product with Linv differs from
// the reference's forward substitution (mvdens.c:341) by O(eps cond(L)) — inside the 1e-12 parity tolerance for the
// well-conditioned factors a PMC proposal has (tests/test_gpu_mma.py measures it against the oracle).
//
//   k_draw_pick/perm/mma  x = mu_k + L_k g with the samples grouped by component (same Philox counters as the literal
//                    kernel); inside a slab box the rejected first attempts are handed to the literal kernel.
#include <algorithm>
#include <cmath>
#include <vector>
#include "pmc_kernels.h"

namespace pmcb200 {

namespace {

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cpa16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cpa_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__host__ __device__ constexpr int q_tiles(int NT) {  // tiles (nt, ks) with 4 ks <= 8 nt + 7
  int t = 0;
  for (int ks = 0; ks < 2 * NT; ++ks) t += NT - ks / 2;
  return t;
}
__host__ __device__ constexpr int q_comp_stride(int NT) { return 8 * NT + 32 * q_tiles(NT); }
__host__ __device__ constexpr int q_mt(int NT) { return NT <= 4 ? 4 : (NT <= 8 ? 2 : 1); }
// components per shared-memory stage of the whitening kernel: small factors are staged several at a time so that the
// block-wide barrier and the cp.async round trip are paid once per 640+ DMMAs per warp group
__host__ __device__ constexpr int q_cps(int NT) { return NT <= 4 ? 4 : (NT <= 8 ? 2 : 1); }
// draw kernel: samples per warp (x 8) and blocks per SM.  Each sample costs ~1800 FP64 issue slots of Box-Muller latency chains,
// so the kernel wants warps, not register tiles: two blocks of 8 warps per SM where the fragments fit in 128 registers (one
// block per SM ran the configs[3] draw at 3.95 ms per 1e7 samples with the FP64 pipe a quarter busy)
__host__ __device__ constexpr int d_mt(int NT) { return NT <= 2 ? 4 : (NT <= 4 ? 2 : 1); }
__host__ __device__ constexpr int d_blocks(int NT) { return NT <= 8 ? 2 : 1; }
// warps per block of the whitening kernel: as many as the register file allows (NT = 16 needs 162 registers: 12 warps)
__host__ __device__ constexpr int q_warps(int NT) { return NT == 16 ? 12 : 8; }

// ---- whitening ----------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(32 * q_warps(NT), 1)
k_q_mma(const double *__restrict__ X, long n, int d, const double *__restrict__ pack, int J,
        const int *__restrict__ kmap, double *__restrict__ Q, long qs) {
  constexpr int KS = 2 * NT, MT = q_mt(NT), CS = q_comp_stride(NT), QW = q_warps(NT), TS = 8 * QW * MT, CPS = q_cps(NT);
  extern __shared__ double sm[];  // [2][CPS][CS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const long ntile = (n + TS - 1) / TS;
  if ((long)blockIdx.x >= ntile) return;
  const int nstage = (J + CPS - 1) / CPS;  // stages per sample tile
  auto issue = [&](int buf, int st) {      // components [st*CPS, min(J, st*CPS+CPS)) are contiguous in the pack
    const int j0 = st * CPS, nj = (J - j0) < CPS ? (J - j0) : CPS;
    const double *src = pack + (size_t)j0 * CS;
    double *dst = sm + (size_t)buf * CPS * CS;
    for (int c = threadIdx.x; c < nj * CS / 2; c += 32 * QW) cpa16(dst + 2 * c, src + 2 * c);
    cpa_commit();
  };
  issue(0, 0);
  unsigned s = 0;
  for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const long base = tile * TS + warp * (8 * MT);
    double xa[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const long row = base + mt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int col = 4 * ks + t4;
        xa[mt][ks] = (row < n && col < d) ? X[(size_t)row * d + col] : 0.0;
      }
    }
    const bool last_tile = tile + gridDim.x >= ntile;
    for (int st = 0; st < nstage; ++st, ++s) {
      cpa_wait0();
      __syncthreads();
      if (!(last_tile && st == nstage - 1)) issue((s + 1) & 1, (st + 1 == nstage) ? 0 : st + 1);
      const int j0 = st * CPS, nj = (J - j0) < CPS ? (J - j0) : CPS;
      for (int jj = 0; jj < nj; ++jj) {
        const double *cb = sm + ((size_t)(s & 1) * CPS + jj) * CS;
        double C[MT][NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) { C[mt][nt][0] = 0.0; C[mt][nt][1] = 0.0; }
        const double *tp = cb + 8 * NT + lane;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double mu = cb[4 * ks + t4];
          double a[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) a[mt] = xa[mt][ks] - mu;
#pragma unroll
          for (int nt = ks / 2; nt < NT; ++nt) {
            const double b = *tp;
            tp += 32;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma(C[mt][nt][0], C[mt][nt][1], a[mt], b);
          }
        }
        const int k = kmap[j0 + jj];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          double q = 0.0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) { q = fma(C[mt][nt][0], C[mt][nt][0], q); q = fma(C[mt][nt][1], C[mt][nt][1], q); }
          q += __shfl_xor_sync(0xffffffffu, q, 1);
          q += __shfl_xor_sync(0xffffffffu, q, 2);
          const long row = base + mt * 8 + g;
          if (t4 == 0 && row < n) Q[(size_t)k * qs + row] = q;
        }
      }
    }
  }
}

// ---- literal weight tail from q ---------------------------------------------------------------------------------------
// mix_mvdens_log_pdf (mvdens.c:784-825) with ld_k = tail(q_k).  The column of q values is overwritten by the
// component log-densities ld_k (k_resp_q reads them back; the Student-t tail costs a log, so it is evaluated once).
__device__ __forceinline__ double mix_lse_q(const MixDev &m, double *__restrict__ Q, long qs, long i) {
  int iloc = 0;
  while (iloc < m.K && m.alpha[iloc] <= 0) iloc++;
  if (iloc >= m.K) return nan("");
  double mx = logpdf_from_q(Q[(size_t)iloc * qs + i], m.d, m.logdet[iloc], m.df[iloc], m.gam[iloc], m.norm[iloc]);
  Q[(size_t)iloc * qs + i] = mx;
  for (int k = iloc + 1; k < m.K; ++k)
    if (m.alpha[k] > 0.0) {
      const double v = logpdf_from_q(Q[(size_t)k * qs + i], m.d, m.logdet[k], m.df[k], m.gam[k], m.norm[k]);
      Q[(size_t)k * qs + i] = v;
      if (mx < v) mx = v;
    }
  double lpos = 0.0;
  for (int k = 0; k < m.K; ++k)
    if (m.alpha[k] > 0.0) lpos = lpos + exp(Q[(size_t)k * qs + i] - mx + kDynMax) * m.alpha[k];
  return mx + log(lpos) - kDynMax;
}

// The same value without storing the component log-densities (importance-only iterations: nothing reads them afterwards),
// with the table-driven exp / log of pmc_dev.cuh and the per-component constants hoisted into shared memory by the block:
//     ld_k = c0_k - hq_k q - hnp_k log(1 + q invn_k)
// Gaussian component: c0 = -d ln2pi / 2 - log detL, hq = 1/2, hnp = invn = 0; Student-t (mvdens.c:376-379): c0 = lgamma((n+d)/2)
// - lgamma(n/2) - d/2 log(n pi) - log detL, hq = 0, hnp = (n+d)/2, invn = 1/n.  Pass 1 finds the maximum, pass 2 recomputes each
// log-density from q and accumulates.  ncu of the literal tail at configs[3] (K = 50, nu = 9): 206 warp instructions per
// (sample, component) pair - a libm log, a libm exp, a double division 1/n, int-to-double conversions and five parameter
// loads with 64-bit address arithmetic per pair - issue-bound at 4.2 ms per 1e7 samples.
// Returns false if a q is not finite or negative: the caller then takes the literal function (the reference's NaN flow).
struct LseTabParams { const double *c0, *hq, *hnp, *invn, *alpha; };
__device__ __forceinline__ bool mix_lse_q_tab(int K, const LseTabParams &P, const double *__restrict__ Q, long qs, long i,
                                              const double *s_exp, const double2 *s_log, double &out) {
  constexpr int U = 8;
  double mx = -INFINITY;
  bool fin = true;
  for (int k0 = 0; k0 < K; k0 += U) {
    double q[U];
#pragma unroll
    for (int j = 0; j < U; ++j) q[j] = (k0 + j < K) ? Q[(size_t)(k0 + j) * qs + i] : 0.0;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int k = k0 + j;
      if (k < K && P.alpha[k] > 0.0) {
        fin = fin && (q[j] >= 0.0) && (q[j] < INFINITY);
        const double v = fma(-P.hnp[k], log_tab(fma(q[j], P.invn[k], 1.0), s_log), fma(-P.hq[k], q[j], P.c0[k]));
        mx = (mx < v) ? v : mx;
      }
    }
  }
  if (!fin || !(mx > -INFINITY) || !(mx < INFINITY)) return false;
  double lpos = 0.0;
  for (int k0 = 0; k0 < K; k0 += U) {
    double q[U];
#pragma unroll
    for (int j = 0; j < U; ++j) q[j] = (k0 + j < K) ? Q[(size_t)(k0 + j) * qs + i] : 0.0;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int k = k0 + j;
      if (k < K && P.alpha[k] > 0.0) {
        const double v = fma(-P.hnp[k], log_tab(fma(q[j], P.invn[k], 1.0), s_log), fma(-P.hq[k], q[j], P.c0[k]));
        lpos = fma(exp_tab(v - mx + kDynMax, s_exp), P.alpha[k], lpos);
      }
    }
  }
  out = mx + log_tab(lpos, s_log) - kDynMax;
  return true;
}

// same outputs and flag/maximum semantics as k_weights_generic (pmc.c:528-590); store_ld = 0: the component log-densities are
// not written back over q (importance-only iterations)
__global__ void k_weights_q(int store_ld, MixDev m, TargetDev tg, int mode, const double *__restrict__ post_in,
                            const short *__restrict__ ok_in, double t_temper, long first_index, long N,
                            const double *__restrict__ X, double *__restrict__ Q, double *__restrict__ Qt,
                            long qs, double *__restrict__ lw, double *__restrict__ log_rho, short *__restrict__ flg,
                            double *blk_maxw, double *blk_maxr, unsigned long long *counters, double *first_vals) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  __shared__ double s_exp[64];
  __shared__ double2 s_log[128];
  extern __shared__ double s_par[];  // store_ld == 0: [5][K] per-component constants of mix_lse_q_tab
  LseTabParams P{s_par, s_par + m.K, s_par + 2 * m.K, s_par + 3 * m.K, s_par + 4 * m.K};
  if (!store_ld) {  // tables of exp_tab / log_tab (pmc_dev.cuh), built by the block: 2^(j/64); (rc_j, -log rc_j)
    for (int t = threadIdx.x; t < 128; t += blockDim.x) {
      if (t < 64) s_exp[t] = exp2((double)t / 64.0);
      const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
      s_log[t] = make_double2(rc, -log(rc));
    }
    for (int k = threadIdx.x; k < m.K; k += blockDim.x) {
      const bool st = m.df[k] != -1;
      const double n = (double)m.df[k];
      s_par[k] = st ? (m.gam[k] - m.norm[k]) - m.logdet[k] : -0.5 * (m.d * kLn2Pi) - m.logdet[k];
      s_par[m.K + k] = st ? 0.0 : 0.5;
      s_par[2 * m.K + k] = st ? 0.5 * (n + m.d) : 0.0;
      s_par[3 * m.K + k] = st ? 1.0 / n : 0.0;
      s_par[4 * m.K + k] = m.alpha[k];
    }
    __syncthreads();
  }
  double my_w = -INFINITY, my_r = -INFINITY;
  bool okflag = false;
  if (i < N) {
    flg[i] = 0;
    lw[i] = 0;
    double rloc;
    if (store_ld || !mix_lse_q_tab(m.K, P, Q, qs, i, s_exp, s_log, rloc)) rloc = mix_lse_q(m, Q, qs, i);
    my_r = rloc;
    log_rho[i] = rloc;
    bool reached_w = false;
    double w = 0.0;
    if (!(isnan(rloc) || isinf(rloc))) {
      bool tok = true;
      double post;
      if (mode == 0) {
        if (tg.kind == PMCGPU_TGT_BANANA) post = bent_gauss(X + (size_t)i * m.d, tg.d, tg.b, tg.sig1);
        else if (tg.kind == PMCGPU_TGT_MVDENS)
          post = logpdf_from_q(Qt[i], tg.mix.d, tg.mix.logdet[0], tg.mix.df[0], tg.mix.gam[0], tg.mix.norm[0]);
        else post = mix_lse_q(tg.mix, Qt, qs, i);
      } else {
        post = post_in[i];
        tok = ok_in ? (ok_in[i] != 0) : true;
      }
      if (tok) {
        w = t_temper * post - rloc;
        reached_w = true;
        lw[i] = w;
        flg[i] = 1;
        okflag = true;
        my_w = w;
        if (w == INFINITY) atomicAdd(&counters[CNT_POSINF], 1ull);
      }
    }
    if (first_index + i == 0) {
      first_vals[0] = rloc;
      first_vals[1] = 1.0;
      first_vals[2] = w;
      first_vals[3] = reached_w ? 1.0 : 0.0;
    }
  }
  weight_block_epilogue(my_w, my_r, okflag, blk_maxw, blk_maxr, counters);
}

// mixture log-density only (pmcgpu_proposal_log_pdf on the large-d path)
__global__ void k_lse_q(MixDev m, long n, double *__restrict__ Q, long qs, double *__restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = mix_lse_q(m, Q, qs, i);
}

// responsibilities in place of ld: g[k][i] = w_i alpha_k exp(ld_k(x_i) - log_rho_i) for ok samples and components with
// count > MINCOUNT (pmc.c:1382-1390; the reference's common factor exp(maxR - 500) is dropped, it cancels).
// Student-t components (pmc.c:1391-1397): the stored value is gamma g with gamma = (nu + d) / (nu + q_k(x_i)); q is
// recovered from the log-density (expm1 keeps it accurate for small q) and sum_i g, which the reference divides the
// covariance by (weight_d, pmc.c:1459-1460), is reduced per block in a fixed order into wpart[block][K].
constexpr int RESP_THREADS = 1024;
__global__ void __launch_bounds__(RESP_THREADS)
k_resp_q(MixDev m, long N, const double *__restrict__ w, const double *__restrict__ log_rho,
         const short *__restrict__ flg, const unsigned long long *__restrict__ count, double *__restrict__ Q, long qs,
         double *__restrict__ wpart /* nullptr for Gaussian mixtures */) {
  extern __shared__ double sw[];  // [32 warps][K] when wpart != nullptr
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool live = i < N;
  const bool ok = live && flg[i] != 0;
  const double wi = live ? w[i] : 0.0, lr = live ? log_rho[i] : 0.0;
  for (int k = 0; k < m.K; ++k) {
    double g = 0.0, gg = 0.0;
    const bool alive = m.alpha[k] > 0.0;
    if (alive && ok && count[k] > kMinCount) {
      const double ld = Q[(size_t)k * qs + i];  // the column holds ld_k since k_weights_q
      g = wi * (m.alpha[k] * exp(ld - lr));
      if (!(g == g) || isinf(g)) g = 0.0;
      gg = g;
      if (m.df[k] != -1) {
        const double n = (double)m.df[k];
        const double q = n * expm1((m.gam[k] - m.norm[k] - m.logdet[k] - ld) / (0.5 * (n + m.d)));
        gg = ((n + m.d) / (n + q)) * g;
        if (!(gg == gg) || isinf(gg)) gg = 0.0;
      }
    }
    if (alive && live) Q[(size_t)k * qs + i] = gg;
    if (wpart) {
      const double s = warp_sum(g);
      if (lane == 0) sw[wid * m.K + k] = s;
    }
  }
  if (wpart) {
    __syncthreads();
    for (int k = threadIdx.x; k < m.K; k += blockDim.x) {
      double s = 0.0;
      for (int v = 0; v < RESP_THREADS / 32; ++v) s += sw[v * m.K + k];
      wpart[(size_t)blockIdx.x * m.K + k] = s;
    }
  }
}

// ---- statistics ------------------------------------------------------------------------------------------------------
constexpr int ST_SS = 64;  // samples per shared-memory stage
constexpr int ST_CK = 2;   // components per block: the centred tile and its B fragments are shared by both
__host__ __device__ constexpr int st_tw(int NT) { return NT / 2; }
__host__ __device__ constexpr int st_sw(int NT) { return 8 / st_tw(NT) < 1 ? 1 : 8 / st_tw(NT); }
__host__ __device__ constexpr int st_warps(int NT) { return st_tw(NT) * st_sw(NT); }
__host__ __device__ constexpr int st_dpad(int NT) { return 8 * NT + 4; }
// doubles one (warp, component) writes: (NT+1) tiles x 64, two s1 fragments x 32, one W fragment x 32
__host__ __device__ constexpr int st_warp_len(int NT) { return (NT + 1) * 64 + 96; }
__host__ __device__ constexpr int st_unit_len(int NT) { return st_warps(NT) * ST_CK * st_warp_len(NT); }

// The DMMA loop of one warp that owns row tiles R0 = TWI and R1 = NT-1-TWI (compile-time, so every accumulator and
// fragment index is static): per 4 samples R1+1 B-fragment loads feed CK (NT+1) DMMAs.
template <int NT, int TWI>
__device__ __forceinline__ void stats_warp(const double *__restrict__ dx, const double *__restrict__ dg, int sw, int g,
                                           int t4, double (&C)[ST_CK][NT + 1][2], double (&s1)[ST_CK][2],
                                           double (&wacc)[ST_CK]) {
  constexpr int SW = st_sw(NT), DPAD = st_dpad(NT), SS = ST_SS, R0 = TWI, R1 = NT - 1 - TWI;
#pragma unroll 2
  for (int ks = sw; ks < SS / 4; ks += SW) {
    const int srow = 4 * ks + t4;
    const double *xr = dx + srow * DPAD + g;
    double bf[R1 + 1];
#pragma unroll
    for (int bt = 0; bt <= R1; ++bt) bf[bt] = xr[8 * bt];
    double a0[ST_CK], a1[ST_CK];
#pragma unroll
    for (int c = 0; c < ST_CK; ++c) {
      const double gv = dg[c * SS + srow];
      a0[c] = gv * bf[R0];
      a1[c] = gv * bf[R1];
      s1[c][0] += a0[c];
      s1[c][1] += a1[c];
      wacc[c] += gv;
    }
#pragma unroll
    for (int bt = 0; bt <= R1; ++bt) {
#pragma unroll
      for (int c = 0; c < ST_CK; ++c) {
        dmma(C[c][R0 + 1 + bt][0], C[c][R0 + 1 + bt][1], a1[c], bf[bt]);
        if (bt <= R0) dmma(C[c][bt][0], C[c][bt][1], a0[c], bf[bt]);
      }
    }
  }
}

// Xc = X - pivot for samples that carry weight, exact zeros otherwise (so that 0 * NaN never reaches a DMMA)
__global__ void k_centre(const double *__restrict__ X, const short *__restrict__ flg, const double *__restrict__ pivot,
                         long n, int d, double *__restrict__ Xc) {
  const long tot = n * d;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
    const long i = e / d;
    const int a = (int)(e - i * d);
    Xc[e] = flg[i] != 0 ? X[e] - pivot[a] : 0.0;
  }
}

template <int NT>
__global__ void __launch_bounds__(32 * st_warps(NT), 1)
k_stats_mma(const double *__restrict__ X /* centred, rows without weight zeroed (k_centre) */, long n, int d,
            const double *__restrict__ G, long qs, int J, const int *__restrict__ kmap, int nchunk, long chunk_len,
            double *__restrict__ part) {
  constexpr int TW = st_tw(NT), SW = st_sw(NT), NWARP = TW * SW, NTHR = 32 * NWARP, DPAD = st_dpad(NT), SS = ST_SS,
                CK = ST_CK;
  extern __shared__ double sm[];
  double *xs = sm;                       // [2][SS][DPAD]
  double *gs = sm + 2 * SS * DPAD;       // [2][CK][SS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const int tw = warp % TW, sw = warp / TW;
  const int jg = blockIdx.x / nchunk, ch = blockIdx.x % nchunk;
  const double *Gk[CK];
#pragma unroll
  for (int c = 0; c < CK; ++c) Gk[c] = (jg * CK + c < J) ? G + (size_t)kmap[jg * CK + c] * qs : nullptr;
  const long lo = ch * chunk_len, hi = (lo + chunk_len < n) ? lo + chunk_len : n;
  for (int c = threadIdx.x; c < 2 * SS * DPAD; c += NTHR) xs[c] = 0.0;
  for (int c = threadIdx.x; c < 2 * CK * SS; c += NTHR) gs[c] = 0.0;
  __syncthreads();
  const bool even = (d & 1) == 0;
  auto issue = [&](int buf, long s0) {
    double *dx = xs + (size_t)buf * SS * DPAD;
    const int rows = (int)((hi - s0) < SS ? (hi - s0) : SS);
    if (even) {
      const int per = d / 2;
      for (int c = threadIdx.x; c < rows * per; c += NTHR) {
        const int r = c / per, q2 = c - r * per;
        cpa16(dx + r * DPAD + 2 * q2, X + (size_t)(s0 + r) * d + 2 * q2);
      }
    } else {
      for (int c = threadIdx.x; c < rows * d; c += NTHR) {
        const int r = c / d, q1 = c - r * d;
        cpa8(dx + r * DPAD + q1, X + (size_t)(s0 + r) * d + q1);
      }
    }
    for (int c = threadIdx.x; c < CK * SS; c += NTHR) {
      const int cc = c / SS, r = c - cc * SS;
      if (r < rows && Gk[cc]) cpa8(gs + (buf * CK + cc) * SS + r, Gk[cc] + s0 + r);
      else gs[(buf * CK + cc) * SS + r] = 0.0;  // tail rows carry no weight (their tile rows keep older, finite data)
    }
    cpa_commit();
  };
  double C[CK][NT + 1][2];
  double s1[CK][2], wacc[CK];
#pragma unroll
  for (int c = 0; c < CK; ++c) {
#pragma unroll
    for (int t = 0; t <= NT; ++t) { C[c][t][0] = 0.0; C[c][t][1] = 0.0; }
    s1[c][0] = s1[c][1] = wacc[c] = 0.0;
  }
  if (lo < hi) issue(0, lo);
  unsigned st = 0;
  for (long s0 = lo; s0 < hi; s0 += SS, ++st) {
    const int buf = st & 1;
    cpa_wait0();
    __syncthreads();  // stage st landed; everybody is done with stage st-1 (the other buffer)
    if (s0 + SS < hi) issue(buf ^ 1, s0 + SS);
    const double *dx = xs + (size_t)buf * SS * DPAD;
    const double *dg = gs + buf * CK * SS;
    switch (tw) {
      case 0: stats_warp<NT, 0>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 1: if constexpr (TW > 1) stats_warp<NT, 1>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 2: if constexpr (TW > 2) stats_warp<NT, 2>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 3: if constexpr (TW > 3) stats_warp<NT, 3>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 4: if constexpr (TW > 4) stats_warp<NT, 4>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 5: if constexpr (TW > 5) stats_warp<NT, 5>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 6: if constexpr (TW > 6) stats_warp<NT, 6>(dx, dg, sw, g, t4, C, s1, wacc); break;
      case 7: if constexpr (TW > 7) stats_warp<NT, 7>(dx, dg, sw, g, t4, C, s1, wacc); break;
    }
  }
#pragma unroll
  for (int c = 0; c < CK; ++c) {
    double *o = part + (size_t)blockIdx.x * st_unit_len(NT) + (size_t)(warp * CK + c) * st_warp_len(NT);
#pragma unroll
    for (int t = 0; t <= NT; ++t) { o[t * 64 + 2 * lane] = C[c][t][0]; o[t * 64 + 2 * lane + 1] = C[c][t][1]; }
    o[(NT + 1) * 64 + lane] = s1[c][0];
    o[(NT + 1) * 64 + 32 + lane] = s1[c][1];
    o[(NT + 1) * 64 + 64 + lane] = wacc[c];
  }
}

// partials -> canonical statistics  out[k] = { W, s1[d], S2[d*d] (full symmetric) }, fixed summation order
template <int NT>
__global__ void k_stats_mma_combine(const double *__restrict__ part, int J, const int *__restrict__ kmap, int nchunk,
                                    int d, double *__restrict__ out) {
  constexpr int TW = st_tw(NT), SW = st_sw(NT), WL = st_warp_len(NT), UL = st_unit_len(NT), CK = ST_CK;
  const int j = blockIdx.y;
  const int k = kmap[j];
  const int jg = j / CK, cc = j % CK;
  const long len = 1 + d + (long)d * d;
  double *ok = out + (size_t)k * len;
  auto blk = [&](int c, int w) { return part + (size_t)(jg * nchunk + c) * UL + (size_t)(w * CK + cc) * WL; };
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < 1 + d + (long)d * (d + 1) / 2; e += (long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    if (e == 0) {  // W: warps with tw == 0, lanes g == 0 (t4 = 0..3 carry distinct samples)
      for (int c = 0; c < nchunk; ++c)
        for (int sw = 0; sw < SW; ++sw) {
          const double *p = blk(c, sw * TW) + (NT + 1) * 64 + 64;
          acc += (p[0] + p[1]) + (p[2] + p[3]);
        }
      ok[0] = acc;
    } else if (e <= d) {
      const int a = (int)e - 1, rt = a >> 3, g = a & 7;
      const bool first = rt < TW;
      const int tw = first ? rt : NT - 1 - rt;
      for (int c = 0; c < nchunk; ++c)
        for (int sw = 0; sw < SW; ++sw) {
          const double *p = blk(c, sw * TW + tw) + (NT + 1) * 64 + (first ? 0 : 32) + 4 * g;
          acc += (p[0] + p[1]) + (p[2] + p[3]);
        }
      ok[1 + a] = acc;
    } else {
      // lower-triangle element (a >= b), enumerated row by row
      long t = e - 1 - d;
      int a = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((long)(a + 1) * (a + 2) / 2 <= t) ++a;
      while ((long)a * (a + 1) / 2 > t) --a;
      const int b = (int)(t - (long)a * (a + 1) / 2);
      const int rt = a >> 3, bt = b >> 3, g = a & 7, col = b & 7;
      const bool first = rt < TW;
      const int tw = first ? rt : NT - 1 - rt;
      const int slot = first ? bt : tw + 1 + bt;
      const int off = slot * 64 + 2 * (g * 4 + (col >> 1)) + (col & 1);
      for (int c = 0; c < nchunk; ++c)
        for (int sw = 0; sw < SW; ++sw) acc += blk(c, sw * TW + tw)[off];
      ok[1 + d + (size_t)a * d + b] = acc;
      ok[1 + d + (size_t)b * d + a] = acc;
    }
  }
}

// ---- draw ------------------------------------------------------------------------------------------------------------
// simulate_mix_mvdens without a box (pmc.c:266-298): x = mu_k + L_k g.  Samples are grouped by their component so that
// L_k g is a product of one factor with many normal vectors.  The Philox counters are those of k_draw_generic
// (call 0: component, calls 1..ceil(d/2): Box-Muller pairs, then the chi-square of a Student-t component), so both
// kernels produce the same sample up to the rounding of the different summation order.
__global__ void k_draw_pick(const double *__restrict__ cw, int K, uint64_t seed, uint32_t iter, long first_index, long N,
                            unsigned long long *__restrict__ indices, unsigned int *__restrict__ cnt) {
  extern __shared__ unsigned int sc[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) sc[k] = 0;
  __syncthreads();
  const Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
    const uint64_t gi = (uint64_t)(first_index + i);
    const uint4 r0 = ph((uint32_t)gi, (uint32_t)(gi >> 32), 0u, iter);
    const int k = ranindx(cw, K, u01_co(r0.x, r0.y));
    indices[i] = (unsigned long long)k;
    atomicAdd(&sc[k], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (sc[k]) atomicAdd(&cnt[k], sc[k]);
}
// perm = sample ids grouped by component (order inside a group is irrelevant: every sample is computed from its own id).
// A block ranks its samples per component with shared-memory atomics and reserves one range per component with a single
// global atomic: K instead of PERM_CHUNK global atomics per block (1e7 atomics on 50 addresses cost 2.7 ms before).
constexpr int PERM_PER_THREAD = 8, PERM_CHUNK = 256 * PERM_PER_THREAD;
__global__ void __launch_bounds__(256)
k_draw_perm(long N, int K, const unsigned long long *__restrict__ indices, unsigned int *__restrict__ cursor,
            unsigned int *__restrict__ perm) {
  extern __shared__ unsigned int sc[];  // [K] counts, then [K] bases
  unsigned int *sb = sc + K;
  for (long c0 = (long)blockIdx.x * PERM_CHUNK; c0 < N; c0 += (long)gridDim.x * PERM_CHUNK) {
    for (int k = threadIdx.x; k < K; k += 256) sc[k] = 0;
    __syncthreads();
    unsigned int rank[PERM_PER_THREAD];
    int comp[PERM_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PERM_PER_THREAD; ++u) {
      const long i = c0 + u * 256 + threadIdx.x;
      comp[u] = -1;
      if (i < N) { comp[u] = (int)indices[i]; rank[u] = atomicAdd(&sc[comp[u]], 1u); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) sb[k] = sc[k] ? atomicAdd(&cursor[k], sc[k]) : 0u;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PERM_PER_THREAD; ++u)
      if (comp[u] >= 0) perm[sb[comp[u]] + rank[u]] = (unsigned int)(c0 + u * 256 + threadIdx.x);
    __syncthreads();
  }
}

template <int NT>
__global__ void __launch_bounds__(256, d_blocks(NT))
k_draw_mma(const double *__restrict__ pack, const int4 *__restrict__ tiles, int ntiles, const unsigned int *__restrict__ perm,
           const int *__restrict__ dfj, int d, uint64_t seed, uint32_t iter, long first_index, double *__restrict__ X,
           short *__restrict__ flg, unsigned long long *counters, const double *__restrict__ thr,
           unsigned int *__restrict__ redo, unsigned int *__restrict__ nredo) {
  constexpr int KS = 2 * NT, MT = d_mt(NT), CS = q_comp_stride(NT);
  extern __shared__ double sm[];  // [CS]: mean and factor tiles of the current component
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const int per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int tb = blockIdx.x * per, te = min(ntiles, tb + per);
  const Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  int cur = -1;
  for (int t = tb; t < te; ++t) {
    const int4 td = tiles[t];  // x = packed component, y = first position in perm, z = samples in this tile
    if (td.x != cur) {
      __syncthreads();
      const double *src = pack + (size_t)td.x * CS;
      for (int c = threadIdx.x; c < CS / 2; c += 256) cpa16(sm + 2 * c, src + 2 * c);
      cpa_commit();
      cur = td.x;
    }
    // standard normals in A-fragment order: over a pair of k-steps (8 columns) each lane of a quad evaluates one
    // Box-Muller pair, the quad exchanges them by shuffles
    double xa[MT][KS];
    long sid[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int pos = warp * (8 * MT) + mt * 8 + g;
      sid[mt] = pos < td.z ? (long)perm[td.y + pos] : -1;
      const uint64_t gi = (uint64_t)(first_index + (sid[mt] < 0 ? 0 : sid[mt]));
#pragma unroll
      for (int m = 0; m < NT; ++m) {
        const int c0 = 8 * m + 2 * t4;  // this lane's pair of columns
        double g0 = 0.0, g1 = 0.0;
        if (c0 < d) box_muller(ph((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)(1 + 4 * m + t4), iter), g0, g1);
        const int qb = lane & ~3;
        const double e0 = __shfl_sync(0xffffffffu, g0, qb + (t4 >> 1)), e1 = __shfl_sync(0xffffffffu, g1, qb + (t4 >> 1));
        const double f0 = __shfl_sync(0xffffffffu, g0, qb + 2 + (t4 >> 1)), f1 = __shfl_sync(0xffffffffu, g1, qb + 2 + (t4 >> 1));
        const double va = (t4 & 1) ? e1 : e0, vb = (t4 & 1) ? f1 : f0;
        xa[mt][2 * m] = (8 * m + t4 < d) ? va : 0.0;
        xa[mt][2 * m + 1] = (8 * m + 4 + t4 < d) ? vb : 0.0;
      }
    }
    cpa_wait0();
    __syncthreads();
    double C[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { C[mt][nt][0] = 0.0; C[mt][nt][1] = 0.0; }
    const double *tp = sm + 8 * NT + lane;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nt = ks / 2; nt < NT; ++nt) {
        const double b = *tp;
        tp += 32;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(C[mt][nt][0], C[mt][nt][1], xa[mt][ks], b);
      }
    }
    const int df = dfj[td.x];
    // Student-t: y *= sqrt(df / chi2_df)  (mvdens.c:306-312).  The rejection sampler is divergent scalar work: lane t4
    // of a quad draws the factor of the quad's sample mt == t4, the quad then exchanges the factors.
    double corr_mine = 1.0;
    if (df != -1) {
      long mysid = -1;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) if (t4 == mt) mysid = sid[mt];
      if (mysid >= 0) {
        const uint64_t gi = (uint64_t)(first_index + mysid);
        DrawStream st{ph, (uint32_t)gi, (uint32_t)(gi >> 32), iter, 0u, (uint32_t)(1 + (d + 1) / 2)};
        corr_mine = sqrt(df / chisq_mt(st, (double)df));
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const double corr = (df != -1) ? __shfl_sync(0xffffffffu, corr_mine, (lane & ~3) + (mt & 3)) : 1.0;
      bool fin = true, inb = true;
      if (sid[mt] >= 0) {
        double *xr = X + (size_t)sid[mt] * d;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int a = 8 * nt + 2 * t4 + e;
            if (a < d) {
              const double v = C[mt][nt][e] * corr + sm[a];
              fin = fin && isfinite(v);
              if (thr) inb = inb && !(v < -thr[a]) && !(v > thr[d + a]);  // slab faces -e_a and +e_a (parabox.c:112-136)
              xr[a] = v;
            }
          }
        }
      }
      const unsigned bad = __ballot_sync(0xffffffffu, !fin), out = __ballot_sync(0xffffffffu, !inb);
      if (t4 == 0 && sid[mt] >= 0) {
        const bool ok = ((bad >> (lane & ~3)) & 0xfu) == 0, inside = ((out >> (lane & ~3)) & 0xfu) == 0;
        flg[sid[mt]] = (ok && inside) ? 1 : 0;
        if (thr && !(ok && inside)) {  // redrawn by the literal kernel from attempt 1 on
          redo[atomicAdd(nredo, 1u)] = (unsigned int)sid[mt];
          atomicAdd(&counters[CNT_REJECT], 1ull);
        }
        else if (!ok) atomicAdd(&counters[CNT_DRAWFAIL], 1ull);
      }
    }
  }
}

template <int NT>
int launch_draw_t(const double *pack, const int4 *tiles, int ntiles, const unsigned int *perm, const int *dfj, int d,
                  uint64_t seed, uint32_t iter, long first_index, double *X, short *flg, unsigned long long *counters,
                  const double *thr, unsigned int *redo, unsigned int *nredo, int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)q_comp_stride(NT) * sizeof(double);
  static std::atomic<unsigned long long> smem_set{0};
  if (!set_max_smem_once(k_draw_mma<NT>, smem, smem_set)) return 1;
  const int grid = ntiles < sm_count * d_blocks(NT) ? ntiles : sm_count * d_blocks(NT);
  k_draw_mma<NT><<<grid, 256, smem, s>>>(pack, tiles, ntiles, perm, dfj, d, seed, iter, first_index, X, flg, counters, thr, redo, nredo);
  count_launch();
  return cudaGetLastError() != cudaSuccess;
}

template <int NT>
int launch_q(const double *X, long n, int d, const double *pack, int J, const int *kmap, double *Q, long qs, int sm_count,
             cudaStream_t s) {
  constexpr int TS = 8 * q_warps(NT) * q_mt(NT);
  const size_t smem = (size_t)2 * q_cps(NT) * q_comp_stride(NT) * sizeof(double);
  static std::atomic<unsigned long long> smem_set{0};
  if (!set_max_smem_once(k_q_mma<NT>, smem, smem_set)) return 1;
  const long ntile = (n + TS - 1) / TS;
  const int grid = (int)(ntile < sm_count ? ntile : sm_count);
  k_q_mma<NT><<<grid, 32 * q_warps(NT), smem, s>>>(X, n, d, pack, J, kmap, Q, qs);
  count_launch();
  return cudaGetLastError() != cudaSuccess;
}

template <int NT>
int launch_stats(const double *X, long n, int d, const double *G, long qs, int J, const int *kmap, int nchunk,
                 double *part, double *out, cudaStream_t s) {
  const size_t smem = ((size_t)2 * ST_SS * st_dpad(NT) + 2 * ST_CK * ST_SS) * sizeof(double);
  static std::atomic<unsigned long long> smem_set{0};
  if (!set_max_smem_once(k_stats_mma<NT>, smem, smem_set)) return 1;
  long chunk_len = (n + nchunk - 1) / nchunk;
  chunk_len = ((chunk_len + ST_SS - 1) / ST_SS) * ST_SS;
  k_stats_mma<NT><<<((J + ST_CK - 1) / ST_CK) * nchunk, 32 * st_warps(NT), smem, s>>>(X, n, d, G, qs, J, kmap, nchunk, chunk_len, part);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return 1;
  k_stats_mma_combine<NT><<<dim3(8, J), 256, 0, s>>>(part, J, kmap, nchunk, d, out);
  count_launch();
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace

// ---- host interface ----------------------------------------------------------------------------------------------------
int mma_nt(int d) { return d <= 16 ? 2 : (d <= 32 ? 4 : (d <= 64 ? 8 : (d <= 128 ? 16 : 0))); }
size_t mma_comp_stride(int d) {
  switch (mma_nt(d)) {
    case 2: return q_comp_stride(2);
    case 4: return q_comp_stride(4);
    case 8: return q_comp_stride(8);
    case 16: return q_comp_stride(16);
  }
  return 0;
}

// One component in kernel order: mu[8 NT] (zero padded), then for ks ascending, nt = ks/2 .. NT-1, the 32 B-fragment
// values  Linv[8 nt + lane/4][4 ks + lane%4].  Linv is formed column by column in extended precision.
// Conditioning probe of the explicit inverse (host, O(d^2)): for a few vectors b = L z the whitened vector is computed
// both ways in plain double - the product Linv b the tensor kernel forms, and the forward substitution of mvdens.c:341
// the reference (and the literal kernels) use - and the largest difference relative to |y| is returned.  Both have the
// forward error bound d eps cond(L); for an ill-conditioned factor their difference can exceed the 1e-12 parity
// tolerance, and the caller then keeps such a mixture off the tensor path.
static double inverse_discrepancy(int d, const double *L, const std::vector<long double> &inv) {
  double worst = 0.0;
  std::vector<double> z(d), b(d), y1(d), y2(d);
  for (int probe = 0; probe < 3; ++probe) {
    unsigned long long st = 0x9E3779B97F4A7C15ull * (unsigned long long)(probe + 1);
    for (int i = 0; i < d; ++i) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      const double u = (double)(st >> 11) / 9007199254740992.0;
      z[i] = probe == 0 ? ((i & 1) ? -1.0 : 1.0) : (probe == 1 ? 2.0 * u - 1.0 : ((u < 0.5) ? -3.0 * u : 3.0 * u));
    }
    for (int r = 0; r < d; ++r) {
      long double acc = 0.0L;
      for (int c = 0; c <= r; ++c) acc += (long double)L[(size_t)r * d + c] * (long double)z[c];
      b[r] = (double)acc;
    }
    double ymax = 0.0;
    for (int r = 0; r < d; ++r) {
      double a1 = 0.0, t = b[r];
      for (int c = 0; c <= r; ++c) a1 += (double)inv[(size_t)r * d + c] * b[c];
      for (int c = 0; c < r; ++c) t -= L[(size_t)r * d + c] * y2[c];
      y1[r] = a1;
      y2[r] = t / L[(size_t)r * d + r];
      if (std::fabs(y2[r]) > ymax) ymax = std::fabs(y2[r]);
    }
    for (int r = 0; r < d; ++r) {
      const double e = std::fabs(y1[r] - y2[r]) / (ymax > 0.0 ? ymax : 1.0);
      if (!(e <= worst)) worst = e;
    }
  }
  return worst;
}

void mma_pack_component(int d, const double *mean, const double *L, int invert, double *out, double *inv_discrepancy) {
  const int NT = mma_nt(d), DP = 8 * NT;
  // Linv L = I, row r of Linv from the diagonal leftwards: inv[r][c] = -(sum_{m=c+1..r} inv[r][m] L[m][c]) / L[c][c];
  // Lt holds L transposed so that both operands of the dot product are contiguous
  std::vector<double> Lt((size_t)d * d, 0.0);
  for (int r = 0; r < d; ++r)
    for (int c = 0; c <= r; ++c) Lt[(size_t)c * d + r] = L[(size_t)r * d + c];
  std::vector<long double> inv((size_t)d * d, 0.0L);
  for (int r = 0; r < d && !invert; ++r)
    for (int c = 0; c <= r; ++c) inv[(size_t)r * d + c] = (long double)L[(size_t)r * d + c];  // the factor itself (draw)
  for (int r = 0; r < d && invert; ++r) {
    long double *ir = inv.data() + (size_t)r * d;
    ir[r] = 1.0L / (long double)L[(size_t)r * d + r];
    for (int c = r - 1; c >= 0; --c) {
      const double *lc = Lt.data() + (size_t)c * d;
      long double acc0 = 0.0L, acc1 = 0.0L;
      int m = c + 1;
      for (; m + 1 <= r; m += 2) { acc0 += ir[m] * (long double)lc[m]; acc1 += ir[m + 1] * (long double)lc[m + 1]; }
      if (m <= r) acc0 += ir[m] * (long double)lc[m];
      ir[c] = -(acc0 + acc1) / (long double)lc[c];
    }
  }
  if (invert && inv_discrepancy) *inv_discrepancy = inverse_discrepancy(d, L, inv);
  for (int a = 0; a < DP; ++a) out[a] = a < d ? mean[a] : 0.0;
  double *tp = out + DP;
  for (int ks = 0; ks < 2 * NT; ++ks)
    for (int nt = ks / 2; nt < NT; ++nt) {
      for (int lane = 0; lane < 32; ++lane) {
        const int r = 8 * nt + lane / 4, c = 4 * ks + lane % 4;
        tp[lane] = (r < d && c < d && c <= r) ? (double)inv[(size_t)r * d + c] : 0.0;
      }
      tp += 32;
    }
}

int launch_q_mma(const double *X, long n, int d, const double *pack, int J, const int *kmap, double *Q, long qs,
                 int sm_count, cudaStream_t s) {
  if (n <= 0 || J <= 0) return 0;
  switch (mma_nt(d)) {
    case 2: return launch_q<2>(X, n, d, pack, J, kmap, Q, qs, sm_count, s);
    case 4: return launch_q<4>(X, n, d, pack, J, kmap, Q, qs, sm_count, s);
    case 8: return launch_q<8>(X, n, d, pack, J, kmap, Q, qs, sm_count, s);
    case 16: return launch_q<16>(X, n, d, pack, J, kmap, Q, qs, sm_count, s);
  }
  return 1;
}

void launch_weights_q(const MixDev &m, const TargetDev &tg, int mode, const double *post_in, const short *ok_in,
                      double t_temper, long first_index, long N, const double *X, double *Q, double *Qt,
                      long qs, double *lw, double *log_rho, short *flg, double *blk_maxw, double *blk_maxr,
                      unsigned long long *counters, double *first_vals, int store_ld, cudaStream_t s) {
  if (N <= 0) return;
  const int nb = (int)((N + GEN_WEIGHT_THREADS - 1) / GEN_WEIGHT_THREADS);
  k_weights_q<<<nb, GEN_WEIGHT_THREADS, store_ld ? 0 : (size_t)5 * m.K * sizeof(double), s>>>(store_ld, m, tg, mode, post_in, ok_in, t_temper, first_index, N, X, Q, Qt, qs, lw,
                                                log_rho, flg, blk_maxw, blk_maxr, counters, first_vals);
  count_launch();
}

void launch_lse_q(const MixDev &m, long n, double *Q, long qs, double *out, cudaStream_t s) {
  if (n <= 0) return;
  k_lse_q<<<(int)((n + 127) / 128), 128, 0, s>>>(m, n, Q, qs, out);
  count_launch();
}

int resp_q_blocks(long N) { return (int)((N + RESP_THREADS - 1) / RESP_THREADS); }
void launch_resp_q(const MixDev &m, long N, const double *w, const double *log_rho, const short *flg,
                   const unsigned long long *count, double *Q, long qs, double *wpart, cudaStream_t s) {
  if (N <= 0) return;
  const size_t smem = wpart ? (size_t)(RESP_THREADS / 32) * m.K * sizeof(double) : 0;
  k_resp_q<<<resp_q_blocks(N), RESP_THREADS, smem, s>>>(m, N, w, log_rho, flg, count, Q, qs, wpart);
  count_launch();
}

size_t mma_stats_partial_doubles(int d, int J, int nchunk) {
  size_t ul = 0;
  switch (mma_nt(d)) {
    case 2: ul = st_unit_len(2); break;
    case 4: ul = st_unit_len(4); break;
    case 8: ul = st_unit_len(8); break;
    case 16: ul = st_unit_len(16); break;
  }
  return ul * (size_t)((J + ST_CK - 1) / ST_CK) * nchunk;
}

// number of sample chunks per component: fill whole waves of the grid, at least four stages per block
int mma_stats_chunks(long n, int J, int sm_count) {
  const long max_c = std::max<long>(1, n / (4 * ST_SS));
  int best = 1;
  double best_eff = 0.0;
  for (int c = 1; c <= 64 && c <= max_c; ++c) {
    const double waves = (double)((J + ST_CK - 1) / ST_CK) * c / sm_count;
    const double eff = waves / std::ceil(waves);
    if (eff > best_eff + 1e-3) { best_eff = eff; best = c; }
  }
  return best;
}

int mma_draw_tile_samples(int d) {
  switch (mma_nt(d)) {
    case 2: return 64 * d_mt(2);
    case 4: return 64 * d_mt(4);
    case 8: return 64 * d_mt(8);
    case 16: return 64 * d_mt(16);
  }
  return 0;
}
void launch_draw_pick(const double *cw, int K, uint64_t seed, uint32_t iter, long first_index, long N,
                      unsigned long long *indices, unsigned int *cnt, int nblocks, cudaStream_t s) {
  k_draw_pick<<<nblocks, 256, K * sizeof(unsigned int), s>>>(cw, K, seed, iter, first_index, N, indices, cnt);
  count_launch();
}
void launch_draw_perm(long N, int K, const unsigned long long *indices, unsigned int *cursor, unsigned int *perm,
                      int nblocks, cudaStream_t s) {
  const long chunks = (N + PERM_CHUNK - 1) / PERM_CHUNK;
  if (chunks < nblocks) nblocks = (int)chunks;
  k_draw_perm<<<nblocks, 256, 2 * K * sizeof(unsigned int), s>>>(N, K, indices, cursor, perm);
  count_launch();
}
int launch_draw_mma(const double *pack, const void *tiles, int ntiles, const unsigned int *perm, const int *dfj, int d,
                    uint64_t seed, uint32_t iter, long first_index, double *X, short *flg, unsigned long long *counters,
                    const double *thr, unsigned int *redo, unsigned int *nredo, int sm_count, cudaStream_t s) {
  if (ntiles <= 0) return 0;
  const int4 *t4 = reinterpret_cast<const int4 *>(tiles);
  switch (mma_nt(d)) {
    case 2: return launch_draw_t<2>(pack, t4, ntiles, perm, dfj, d, seed, iter, first_index, X, flg, counters, thr, redo, nredo, sm_count, s);
    case 4: return launch_draw_t<4>(pack, t4, ntiles, perm, dfj, d, seed, iter, first_index, X, flg, counters, thr, redo, nredo, sm_count, s);
    case 8: return launch_draw_t<8>(pack, t4, ntiles, perm, dfj, d, seed, iter, first_index, X, flg, counters, thr, redo, nredo, sm_count, s);
    case 16: return launch_draw_t<16>(pack, t4, ntiles, perm, dfj, d, seed, iter, first_index, X, flg, counters, thr, redo, nredo, sm_count, s);
  }
  return 1;
}

int launch_stats_mma(const double *X, const short *flg, long n, int d, const double *G, long qs, int J, const int *kmap,
                     const double *pivot, int nchunk, double *Xc, double *part, double *out, int sm_count, cudaStream_t s) {
  if (n <= 0 || J <= 0) return 0;
  k_centre<<<sm_count * 8, 256, 0, s>>>(X, flg, pivot, n, d, Xc);
  count_launch();
  switch (mma_nt(d)) {
    case 2: return launch_stats<2>(Xc, n, d, G, qs, J, kmap, nchunk, part, out, s);
    case 4: return launch_stats<4>(Xc, n, d, G, qs, J, kmap, nchunk, part, out, s);
    case 8: return launch_stats<8>(Xc, n, d, G, qs, J, kmap, nchunk, part, out, s);
    case 16: return launch_stats<16>(Xc, n, d, G, qs, J, kmap, nchunk, part, out, s);
  }
  return 1;
}

}  // namespace pmcb200
