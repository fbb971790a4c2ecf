// pmc_iter3.cu — third-generation kernels of the small-d PMC iteration (the ones bench.py times).
//
// What changed against pmc_fused2.cu, and which round-1 ncu finding each change answers (profiles/r01_final_ncu.txt):
//
//   k_draw3     The second-generation draw picked the component per lane, so every load of a Cholesky-factor entry was a
//               divergent LDS (62 % of the shared-memory wavefront peak, 2.3e8 bank conflicts, FP64 pipe 22 %).  Here a
//               block picks the components of a chunk of samples first (Philox call 0, ranindx as mvdens.c:766-782),
//               buckets the chunk by component in shared memory, and every warp then draws 32 samples of ONE component:
//               mu_k and L_k are warp-uniform constant-bank operands of the DFMAs, no shared-memory traffic besides the
//               ziggurat table.  Sample i keeps its position, its Philox counters and its component: the output is the
//               same i.i.d. sequence whatever the chunking or the sharding.
//   k_weights3  (mvdens.c:784-825, 354-385, 319-349; pmc.c:537-592)  Same forward substitution as before (75 FP64
//               instructions per component at d = 10, each with exactly one constant-bank operand: a second constant in
//               one instruction costs a register load whose latency the scoreboard showed).  exp() is table-driven
//               (2^(j/64) in shared memory + degree-5 polynomial, 11 FP64 instructions instead of ~23 slots), log() likewise
//               (128-entry reciprocal table, 10 instead of ~53), dead components are neutralised in the pack instead of by
//               selects, the kernel carries no draw code, the next tile's samples are loaded while the current one is
//               computed, and the kernel parameters are __grid_constant__ (ptxas had copied them to local memory).  It
//               also accumulates sum exp(lw - max) with per-lane online rescaling, so the normaliser is known when the
//               kernel ends and no separate pass over the log-weights is needed.
//   k_stats3    (pmc.c:641-697, 1041-1077, 1323-1469)  The DMMA statistics kernel now also writes the normalised weights and
//               the final flags and accumulates sum w log w, sum w^2, the ok count and the component histogram: the two
//               normalisation passes (0.83 ms and 2.8 GB of traffic per 1e8 samples) are gone.
//   k_red3      block partials -> maxima, normaliser, the i==0 quirk of pmc.c:552,579 and the "clean" flag, all on the
//               device: the host synchronises once per iteration, and with ranks attached the three all-reduces run on
//               device buffers between these launches.
//
// Non-finite corner cases keep the reference's behaviour through two escape hatches: a sample whose fast log-density
// is not finite is re-evaluated by a literal restatement of mix_mvdens_log_pdf, and an iteration whose maxima are
// poisoned (NaN at sample 0, +inf log-weight, no usable sample) is flagged "not clean": k_stats3 then leaves the store
// untouched and the host runs the literal normalise/update sequence of pmc_generic.cu instead.
#include <type_traits>
#include "pmc_kernels.h"

// ---- constant-bank parameter blocks (C linkage: the whitening loads address them by symbol from inline PTX) ----------
constexpr int I3_CW = 2048, I3_CT = 1024, I3_CD = 2048;
extern "C" {
__constant__ __align__(16) double pmcb200_c3_w[I3_CW];   // whitening pack of the proposal
__constant__ __align__(16) double pmcb200_c3_t[I3_CT];   // whitening pack of a Gaussian / Gaussian-mixture target
}
#define c3_w pmcb200_c3_w
#define c3_t pmcb200_c3_t

namespace pmcb200 {

__constant__ double c3_d[I3_CD];   // draw pack of the proposal
__constant__ double c3_cw[64];     // cumulative weights (mvdens.c:748-750)
__constant__ double c3_box[2 * FUSED_MAXD];  // slab box: (-lo_thr_j, hi_thr_j) pairs, +-inf when absent
__constant__ double c3_exptab[64];     // 2^(j/64)
__constant__ double2 c3_logtab[128];   // (rc_j, -log rc_j), rc_j ~ 1/(1 + (j+0.5)/128)

__host__ __device__ constexpr int i3_ev2(int n) { return (n + 1) & ~1; }
// whitening pack, per component: -mu_j | 1/L_jj | -L_ij column by column | c' = -0.5 d ln2pi - log detL | alpha
// Every FP64 instruction of the solve takes exactly one of these as its constant-bank operand; entries that are consumed
// together sit in the same 16-byte pair (one LDCU.128 fills two uniform-register pairs: a pair whose halves are needed far
// apart costs two UMOVs to keep the second half alive).
__host__ __device__ constexpr int i3_woR(int D) { return i3_ev2(D); }
__host__ __device__ constexpr int i3_woL(int D) { return 2 * i3_ev2(D); }
__host__ __device__ constexpr int i3_woC(int D) { return i3_woL(D) + D * (D - 1) / 2; }
__host__ __device__ constexpr int i3_wstride(int D) { return i3_ev2(i3_woC(D) + 2); }
// draw pack, per component: rows (mu_i, L_i0 .. L_ii)
__host__ __device__ constexpr int i3_dstride(int D) { return i3_ev2(D + D * (D + 1) / 2); }

int iter3_wstride(int D) { return i3_wstride(D); }
int iter3_dstride(int D) { return i3_dstride(D); }

// pre = 1: row-scaled form for the PRE kernels — m_j = -(mu_j - shift_j)/L_jj | 1/L_jj | -L_ij/L_ii — so that
// y_i = x~_i/L_ii + m_i - sum_j (L_ij/L_ii) y_j needs no separate subtraction and scaling (65 instead of 75 FP64
// instructions per component at d = 10); x~ = x - shift is formed once per sample.
void iter3_pack_w(int K, int D, const double *wght, const double *mean, const double *L, const double *logdet,
                  const double *shift, double *out, int pre) {
  const int st = i3_wstride(D);
  for (int k = 0; k < K; ++k) {
    double *o = out + (size_t)k * st;
    for (int i = 0; i < st; ++i) o[i] = 0.0;
    if (!(wght[k] > 0.0)) {  // dead component: finite q, log-density -1e300, weight 0 -> contributes exactly nothing
      o[i3_woC(D)] = -1.0e300;
      continue;
    }
    const double *Lk = L + (size_t)k * D * D;
    for (int j = 0; j < D; ++j) {
      const double dj = 1.0 / Lk[j * D + j];
      o[j] = pre ? -(mean[(size_t)k * D + j] - shift[j]) * dj : -(mean[(size_t)k * D + j] - shift[j]);
      o[i3_woR(D) + j] = dj;
    }
    int e = i3_woL(D);
    for (int j = 0; j < D; ++j)
      for (int i = j + 1; i < D; ++i) o[e++] = pre ? -Lk[i * D + j] * (1.0 / Lk[i * D + i]) : -Lk[i * D + j];
    // pre: log alpha_k is folded into the constant, so the log-sum-exp runs over log(alpha_k phi_k) without the K products
    // alpha_k exp(.) and without the DYNMAX offsets (mvdens.c:805-822 shifts by max ld_k + 500: any pivot gives the same sum)
    o[i3_woC(D)] = -0.5 * (D * kLn2Pi) - logdet[k] + (pre ? log(wght[k]) : 0.0);
    o[i3_woC(D) + 1] = wght[k];
  }
}

void iter3_pack_d(int K, int D, const double *mean, const double *L, double *out) {
  const int st = i3_dstride(D);
  for (int k = 0; k < K; ++k) {
    double *o = out + (size_t)k * st;
    for (int i = 0; i < st; ++i) o[i] = 0.0;
    int e = 0;
    for (int i = 0; i < D; ++i) {
      o[e++] = mean[(size_t)k * D + i];
      for (int j = 0; j <= i; ++j) o[e++] = L[(size_t)k * D * D + (size_t)i * D + j];
    }
  }
}

// the largest |mu'_j / L_jj| of the live components: subtracting the common shift first costs ~eps times this in
// relative accuracy of the whitened coordinates
double iter3_shift_ratio(int K, int D, const double *wght, const double *mean, const double *L, const double *shift) {
  double worst = 0.0;
  for (int k = 0; k < K; ++k) {
    if (!(wght[k] > 0.0)) continue;
    for (int j = 0; j < D; ++j) {
      const double v = fabs((mean[(size_t)k * D + j] - shift[j]) / L[(size_t)k * D * D + (size_t)j * D + j]);
      if (!(v <= worst)) worst = v;
    }
  }
  return worst;
}

static std::atomic<unsigned long long> g_tables_done{0};
static int upload_tables() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  const unsigned long long bit = 1ull << (dev & 63);
  if (g_tables_done.load(std::memory_order_acquire) & bit) return 0;
  double et[64];
  double2 lt[128];
  for (int j = 0; j < 64; ++j) et[j] = (double)exp2l((long double)j / 64.0L);
  for (int j = 0; j < 128; ++j) {
    const double rc = (double)(1.0L / (1.0L + ((long double)j + 0.5L) / 128.0L));
    lt[j].x = rc;
    lt[j].y = (double)(-logl((long double)rc));
  }
  if (cudaMemcpyToSymbol(c3_exptab, et, sizeof et) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c3_logtab, lt, sizeof lt) != cudaSuccess) return -1;
  g_tables_done.fetch_or(bit, std::memory_order_release);
  return 0;
}

// table-driven exp_tab / log_tab: pmc_dev.cuh (shared with the tensor path's importance tail)

// ---- pre-scaled triangular solve: q = |L^-1 (x - mu)|^2, every parameter a constant-bank operand -----------------------
// compile-time loop: f(std::integral_constant<int, i>) for i in [B, E)
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

// Two consecutive pack entries as ONE constant load (LDCU.128 into uniform registers).  The tile body that uses these lives
// in a __noinline__ function (tile3) because both nvvm and ptxas treat constant loads as loop-invariant: inlined into the
// persistent tile loop, hundreds of them were hoisted in front of the loop and spilled, so that the hot loop read its
// parameters back from local memory (long-scoreboard stalls on 20 % of the warp samples, ncu r2_prof4).
template <bool TGT, int OFF>
__device__ __forceinline__ void ldc2(double &a, double &b) {
  static_assert((OFF & 1) == 0, "pairs are 16-byte aligned");
  if (TGT) asm volatile("ld.const.v2.f64 {%0, %1}, [pmcb200_c3_t + %2];" : "=d"(a), "=d"(b) : "n"(OFF * 8));
  else asm volatile("ld.const.v2.f64 {%0, %1}, [pmcb200_c3_w + %2];" : "=d"(a), "=d"(b) : "n"(OFF * 8));
}

// one component: q = |L^-1 (x - mu)|^2 by forward substitution, lv = -q/2 + c'; al = alpha.  Every FP64 instruction has
// exactly one pack entry as operand (ptxas keeps them in uniform registers: DFMA R, R, UR, R)
// PRE: the row-scaled pack (iter3_pack_w, pre = 1).  The offsets m_j are the one operand that cannot come from the constant
// bank (an FP64 instruction takes one constant operand): they are read from shared memory, two per warp-uniform LDS.128.
template <int D, int SPL, bool TGT, int BASE, bool PRE = false>
__device__ __forceinline__ void comp3(const double (&x)[SPL][D], double (&lv)[SPL], double &al, const double2 *s_mu = nullptr) {
  constexpr int ST = i3_wstride(D);
  constexpr int H = i3_ev2(D) / 2;
  double c[ST];
  static_for<(PRE ? H : 0), ST / 2>([&](auto pc) {
    constexpr int p = decltype(pc)::value;
    ldc2<TGT, BASE + 2 * p>(c[2 * p], c[2 * p + 1]);
  });
  double v[SPL][D], q[SPL];
  if constexpr (PRE) {
    double mm[2 * H];
#pragma unroll
    for (int p = 0; p < H; ++p) { const double2 t2 = s_mu[(BASE / ST) * H + p]; mm[2 * p] = t2.x; mm[2 * p + 1] = t2.y; }
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) v[s][j] = fma(x[s][j], c[i3_woR(D) + j], mm[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) v[s][j] = x[s][j] + c[j];
    }
  }
  int off = i3_woL(D);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double y[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      y[s] = PRE ? v[s][j] : v[s][j] * c[i3_woR(D) + j];
      q[s] = (j == 0) ? y[s] * y[s] : fma(y[s], y[s], q[s]);
    }
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
#pragma unroll
      for (int s = 0; s < SPL; ++s) v[s][i] = fma(c[off + (i - j - 1)], y[s], v[s][i]);
    }
    off += D - 1 - j;
  }
#pragma unroll
  for (int s = 0; s < SPL; ++s) lv[s] = fma(-0.5, q[s], c[i3_woC(D)]);
  al = c[i3_woC(D) + 1];
}

// mix_mvdens_log_pdf (mvdens.c:784-825) of KK Gaussian components packed at c3_w / c3_t: e[k] = alpha_k exp(ld_k - max + 500),
// lpos = sum_k e[k] in component order, returns (max + log lpos) - 500
template <int D, int KK, int SPL, bool TGT, bool PRE = false>
__device__ __forceinline__ void mix_fast(const double (&xs)[SPL][D], const double *s_exp,
                                         const double2 *s_log, double (&e)[SPL][KK], double (&lpos)[SPL], double (&out)[SPL],
                                         const double2 *s_mu = nullptr) {
  constexpr int ST = i3_wstride(D);
  double mx[SPL], al[KK];
  static_for<0, KK>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    double lv[SPL];
    comp3<D, SPL, TGT, k * ST, PRE>(xs, lv, al[k], s_mu);
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      e[s][k] = lv[s];
      mx[s] = (k == 0) ? lv[s] : ((mx[s] < lv[s]) ? lv[s] : mx[s]);
    }
  });
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    lpos[s] = 0.0;
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      if constexpr (PRE) e[s][k] = exp_tab(e[s][k] - mx[s], s_exp);  // log alpha_k is part of the pack's constant
      else e[s][k] = exp_tab((e[s][k] - mx[s]) + kDynMax, s_exp) * al[k];
      lpos[s] = lpos[s] + e[s][k];
    }
    out[s] = PRE ? mx[s] + log_tab(lpos[s], s_log) : (mx[s] + log_tab(lpos[s], s_log)) - kDynMax;
  }
}

// The same for SPL samples per lane with the K component log-densities STREAMED through shared memory instead of held in
// registers: during the component loop a lane keeps only x and the running solve of its SPL samples (one LDCU.128 now
// feeds 2 SPL FP64 instructions), afterwards one sample at a time is finished from the stored values.  lvs: this warp's
// [KK][SPL*32] doubles; on return it holds e[k] = alpha_k exp(ld_k - max + 500) of every sample.
template <int D, int KK, int SPL, bool TGT>
__device__ __forceinline__ void mix_stream(const double (&xs)[SPL][D], const double *s_exp, const double2 *s_log, double *lvs, int lane,
                                           double (&lpos)[SPL], double (&out)[SPL]) {
  constexpr int ST = i3_wstride(D);
  double mx[SPL];
  static_for<0, KK>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    double lv[SPL], al;
    comp3<D, SPL, TGT, k * ST>(xs, lv, al);
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      lvs[(k * SPL + s) * 32 + lane] = lv[s];
      mx[s] = (k == 0) ? lv[s] : ((mx[s] < lv[s]) ? lv[s] : mx[s]);
    }
  });
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    double lp = 0.0;
    static_for<0, KK>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const double al = TGT ? c3_t[k * ST + i3_woC(D) + 1] : c3_w[k * ST + i3_woC(D) + 1];
      const double e = exp_tab((lvs[(k * SPL + s) * 32 + lane] - mx[s]) + kDynMax, s_exp) * al;
      lvs[(k * SPL + s) * 32 + lane] = e;
      lp = lp + e;
    });
    lpos[s] = lp;
    out[s] = (mx[s] + log_tab(lp, s_log)) - kDynMax;
  }
}

// ---- literal restatements for the non-finite corner (reference layout in global memory) --------------------------------
static __device__ __noinline__ double mix_log_pdf_literal(const MixDev m, const double *x) {
  double ld[64], y[FUSED_MAXD];
  int iloc = 0;
  while (iloc < m.K && m.alpha[iloc] <= 0) iloc++;
  if (iloc >= m.K) return nan("");
  double mx = 0.0;
  for (int k = iloc; k < m.K; ++k) {
    if (!(m.alpha[k] > 0.0)) continue;
    const double *L = m.L + (size_t)k * m.d * m.d, *mu = m.mean + (size_t)k * m.d;
    for (int i = 0; i < m.d; ++i) y[i] = x[i] - mu[i];
    y[0] = y[0] / L[0];
    for (int i = 1; i < m.d; ++i) {
      double t = y[i];
      for (int j = 0; j < i; ++j) t -= L[(size_t)i * m.d + j] * y[j];
      y[i] = t / L[(size_t)i * m.d + i];
    }
    double q = 0.0;
    for (int i = 0; i < m.d; ++i) q += y[i] * y[i];
    const double v = -0.5 * (m.d * kLn2Pi + q) - m.logdet[k];
    ld[k] = v;
    if (k == iloc || mx < v) mx = v;
  }
  double lpos = 0.0;
  for (int k = 0; k < m.K; ++k)
    if (m.alpha[k] > 0.0) lpos = lpos + exp(ld[k] - mx + kDynMax) * m.alpha[k];
  return mx + log(lpos) - kDynMax;
}

template <int D>
struct RegVec3 {
  const double (&v)[D];
  __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};

// =====================================================================================================================
// weights
// =====================================================================================================================
struct W3Args {
  const double *X;
  long N, first_index;
  int tgt_kind;
  double b, sig1, t_temper;
  const double *post_in;
  const short *ok_in;
  double *lw, *log_rho, *resp;
  short *flg;
  double *partials;  // [grid][4]: max log-weight over usable samples, sum exp(lw - max), max log_rho, spare
  unsigned long long *counters;
  double *first_vals;
  MixDev mix, tmix;  // reference layout, for the literal re-evaluation of non-finite samples
  double shift[FUSED_MAXD];  // PRE kernels: the common shift of the row-scaled pack
};
__constant__ W3Args c3_wargs;  // the launch's arguments: read at compile-time offsets from every function of the kernel

struct TileAcc {  // per-lane running values of the kernel, updated by tile3 through local memory
  double m, S, r;
  unsigned int nok, posinf;
};

// One tile of 32*SPL samples: target, proposal density, log-weight, flags, responsibilities (pmc.c:537-592).
template <int D, int K, int KT, int SPL, bool RESP, bool PRE>
static __device__ __noinline__ void tile3(long tile, int lane, const double *s_exp, const double2 *s_log, TileAcc *acc, double *lvs,
                                          const double2 *s_mu) {
  const W3Args &a = c3_wargs;
  constexpr int TS = 32 * SPL;
  const long tbase = tile * TS;
  double x[SPL][D];
  bool valid[SPL];
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    const long i = tbase + s * 32 + lane;
    valid[s] = i < a.N;
    const long il = valid[s] ? i : a.N - 1;
    if ((D & 1) == 0) {
      const double2 *xi = reinterpret_cast<const double2 *>(a.X + (size_t)il * D);
#pragma unroll
      for (int j = 0; j < D; j += 2) { const double2 t2 = xi[j / 2]; x[s][j] = t2.x; x[s][j + 1] = t2.y; }
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) x[s][j] = a.X[(size_t)il * D + j];
    }
  }

  // ---- target ----------------------------------------------------------------------------------------------------
  double post[SPL];
  bool tok[SPL];
#pragma unroll
  for (int s = 0; s < SPL; ++s) { post[s] = 0.0; tok[s] = true; }
  if (a.tgt_kind == PMCGPU_TGT_BANANA) {
#pragma unroll
    for (int s = 0; s < SPL; ++s) post[s] = bent_gauss(RegVec3<D>{x[s]}, D, a.b, a.sig1);
  } else if (KT > 0 && a.tgt_kind == PMCGPU_TGT_MVDENS) {
    double talpha;
    comp3<D, SPL, true, 0>(x, post, talpha);
  } else if (KT > 0 && a.tgt_kind == PMCGPU_TGT_MIX) {
    double te[SPL][KT > 0 ? KT : 1], tl[SPL];
    mix_fast<D, (KT > 0 ? KT : 1), SPL, true>(x, s_exp, s_log, te, tl, post);
#pragma unroll
    for (int s = 0; s < SPL; ++s)
      if (!(fabs(post[s]) < INFINITY)) post[s] = mix_log_pdf_literal(a.tmix, a.X + (size_t)(valid[s] ? tbase + s * 32 + lane : a.N - 1) * D);
  } else if (a.tgt_kind == PMCGPU_TGT_HOST) {
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      const long i = tbase + s * 32 + lane;
      if (i < a.N) { post[s] = a.post_in[i]; tok[s] = a.ok_in ? (a.ok_in[i] != 0) : true; }
    }
  }

  // ---- proposal log-density ------------------------------------------------------------------------------------------
  constexpr bool STREAM = SPL > 1;  // two samples per lane: the component log-densities live in shared memory (mix_stream)
  double e[STREAM ? 1 : SPL][STREAM ? 1 : K], lpos[SPL], rloc[SPL];
  if constexpr (PRE) {
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) x[s][j] -= a.shift[j];
    }
  }
  if constexpr (STREAM) mix_stream<D, K, SPL, false>(x, s_exp, s_log, lvs, lane, lpos, rloc);
  else mix_fast<D, K, SPL, false, PRE>(x, s_exp, s_log, e, lpos, rloc, s_mu);

  // ---- log-weight, flags, maxima (pmc.c:546-583) ---------------------------------------------------------------------
  double m_lane = acc->m, S_lane = acc->S, r_lane = acc->r;
  unsigned int nok_lane = acc->nok, posinf_lane = acc->posinf;
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    const long i = tbase + s * 32 + lane;
    bool fastok = fabs(rloc[s]) < INFINITY;
    if (!fastok) rloc[s] = mix_log_pdf_literal(a.mix, a.X + (size_t)(valid[s] ? i : a.N - 1) * D);  // non-finite: the reference's own operation sequence decides
    const bool rfin = !(isnan(rloc[s]) || isinf(rloc[s]));
    const bool reached = valid[s] && rfin && tok[s];
    const double lw = reached ? (a.t_temper * post[s] - rloc[s]) : 0.0;
    if (valid[s]) {
      a.log_rho[i] = rloc[s];
      a.lw[i] = lw;
      a.flg[i] = reached ? 1 : 0;
      if (a.first_index + i == 0) {
        a.first_vals[0] = rloc[s]; a.first_vals[1] = 1.0; a.first_vals[2] = lw; a.first_vals[3] = reached ? 1.0 : 0.0;
      }
      if (rloc[s] > r_lane) r_lane = rloc[s];
    }
    const bool wfin = reached && !isnan(lw) && !(lw == INFINITY);
    if (reached && lw == INFINITY) ++posinf_lane;
    if (wfin) {
      ++nok_lane;
      if (lw > -INFINITY) {  // online sum of exp(lw - running max): one exp per sample
        const double dlt = lw - m_lane;
        const double ex = exp_tab(-fabs(dlt), s_exp);
        if (dlt > 0.0) { S_lane = fma(S_lane, ex, 1.0); m_lane = lw; }
        else S_lane += ex;
      }
    }
    if (RESP && valid[s]) {  // responsibilities alpha_k phi_k(x)/q(x), component-major: coalesced 256-byte rows
      const double rs = fastok ? 1.0 / lpos[s] : 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) a.resp[(size_t)k * a.N + i] = (STREAM ? lvs[(k * SPL + s) * 32 + lane] : e[STREAM ? 0 : s][STREAM ? 0 : k]) * rs;
    }
  }
  acc->m = m_lane; acc->S = S_lane; acc->r = r_lane; acc->nok = nok_lane; acc->posinf = posinf_lane;
}

template <int D, int K, int KT, int SPL, bool RESP, int MINB, bool PRE = false>
__global__ void __launch_bounds__(128, MINB) k_weights3() {
  constexpr int ST = i3_wstride(D);
  static_assert(K * ST <= I3_CW && KT * ST <= I3_CT, "constant memory budget");
  static_assert(!PRE || SPL == 1, "the row-scaled solve is built for one sample per lane");
  const W3Args &a = c3_wargs;
  __shared__ double s_exp[64];
  __shared__ double2 s_log[128];
  __shared__ double s_m[4], s_S[4], s_r[4];
  constexpr int H = i3_ev2(D) / 2;
  __shared__ double2 s_mu[PRE ? K * H : 1];  // PRE: the offsets m_j of every component (the pack's first ev2(D) entries)
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_exp[i] = c3_exptab[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_log[i] = c3_logtab[i];
  if constexpr (PRE) {
    for (int i = threadIdx.x; i < K * H; i += blockDim.x) {
      const int k = i / H, p = i - k * H;
      s_mu[i] = make_double2(c3_w[k * ST + 2 * p], c3_w[k * ST + 2 * p + 1]);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int TS = 32 * SPL;
  extern __shared__ __align__(16) double w3_dyn[];  // SPL > 1: [4 warps][K][SPL*32] component log-densities (mix_stream)
  double *lvs = w3_dyn + (size_t)wid * K * SPL * 32;
  const long ntiles = (a.N + TS - 1) / TS;
  TileAcc acc{-INFINITY, 0.0, -INFINITY, 0u, 0u};
  const long tstep = (long)gridDim.x * 4;
  for (long tile = (long)blockIdx.x * 4 + wid; tile < ntiles; tile += tstep) {
    if (tile + tstep < ntiles) {  // the next tile's rows into L1/L2 while this one is computed: 80-byte rows, two lines at most each
      const char *nx = reinterpret_cast<const char *>(a.X + (size_t)((tile + tstep) * TS + lane) * D);
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + (size_t)s * 32 * D * 8));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + (size_t)s * 32 * D * 8 + D * 8 - 8));
      }
    }
    tile3<D, K, KT, SPL, RESP, PRE>(tile, lane, s_exp, s_log, &acc, lvs, s_mu);
  }
  const double m_lane = acc.m, S_lane = acc.S, r_lane = acc.r;
  const unsigned int nok_lane = acc.nok, posinf_lane = acc.posinf;

  // ---- block partial: (max, sum relative to it, max log_rho), combined in lane/warp order -> deterministic ------------
  double mw = warp_max_nan_ignoring(m_lane);
  double sw = (m_lane == -INFINITY) ? 0.0 : S_lane * exp(m_lane - mw);
  sw = warp_sum(sw);
  const double rw = warp_max_nan_ignoring(r_lane);
  if (lane == 0) { s_m[wid] = mw; s_S[wid] = sw; s_r[wid] = rw; }
  const unsigned long long n1 = (unsigned long long)warp_sum_ll(nok_lane), n2 = (unsigned long long)warp_sum_ll(posinf_lane);
  if (lane == 0) {
    if (n1) atomicAdd(&a.counters[CNT_NOK], n1);
    if (n2) atomicAdd(&a.counters[CNT_POSINF], n2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double Mb = -INFINITY, Rb = -INFINITY, Sb = 0.0;
    for (int w = 0; w < 4; ++w) { if (s_m[w] > Mb) Mb = s_m[w]; if (s_r[w] > Rb) Rb = s_r[w]; }
    for (int w = 0; w < 4; ++w) Sb += (s_m[w] == -INFINITY) ? 0.0 : s_S[w] * exp(s_m[w] - Mb);
    double *out = a.partials + (size_t)blockIdx.x * 4;
    out[0] = Mb; out[1] = Sb; out[2] = Rb; out[3] = 0.0;
  }
}

// =====================================================================================================================
// reductions between the kernels (one block): scal[] is the device-resident scalar block of the iteration
// =====================================================================================================================
// coll[0..7]  (all-reduce max):  M, maxR-candidate, NaN-at-0 flags (W, R), reached-at-0 flags (W, R), +inf seen, spare
// coll[8..15] (all-reduce sum):  S relative to the global M, usable samples, box rejections, failed draws
// scal: I3_* below
// With ranks attached the two blocks travel in ONE collective: phases 1|2|16 leave the rank-local values (S relative to the
// LOCAL maximum) in this rank's 16-double slot of gath[nranks][16], all other slots zero, so that a sum all-reduce of gath is
// an all-gather; phase 8 then folds the slots in rank order - M = max M_r, S = sum_r S_r exp(M_r - M), flags by max, counters
// by sum - identically on every rank, and phase 4 finishes as in the single-rank case.
__global__ void k_red3(const double *__restrict__ partials, int nblocks, const unsigned long long *__restrict__ counters,
                       const double *__restrict__ first_vals, int has_first, double *coll, double *scal, int phases,
                       double n_total, double *gath, int rank, int nranks) {
  __shared__ double sa[256], sb[256];
  const int t = threadIdx.x;
  if (phases & 1) {
    double M = -INFINITY, R = -INFINITY;
    for (int b = t; b < nblocks; b += blockDim.x) {
      const double mb = partials[(size_t)b * 4], rb = partials[(size_t)b * 4 + 2];
      if (mb > M) M = mb;
      if (rb > R) R = rb;
    }
    sa[t] = M; sb[t] = R;
    __syncthreads();
    if (t == 0) {
      for (int i = 1; i < (int)blockDim.x; ++i) { if (sa[i] > M) M = sa[i]; if (sb[i] > R) R = sb[i]; }
      const bool r0w = has_first && first_vals[3] != 0.0, r0r = has_first && first_vals[1] != 0.0;
      scal[I3_MLOC] = M;
      coll[0] = M;
      coll[1] = R;
      coll[2] = (r0w && isnan(first_vals[2])) ? 1.0 : 0.0;
      coll[3] = (r0r && isnan(first_vals[0])) ? 1.0 : 0.0;
      coll[4] = r0w ? 1.0 : 0.0;
      coll[5] = r0r ? 1.0 : 0.0;
      coll[6] = counters[CNT_POSINF] ? 1.0 : 0.0;
      coll[7] = 0.0;
    }
    __syncthreads();
  }
  if (phases & 2) {
    const double Mg = coll[0];
    double S = 0.0;
    // the terms S_b exp(m_b - M) by all threads (one serial thread took 85 us for 592 blocks: two dependent loads and an exp
    // each), then summed by one thread in block order from shared memory: same fixed order, 2 us
    __shared__ double st[1024];
    for (int base = 0; base < nblocks; base += 1024) {
      const int nb = (nblocks - base) < 1024 ? (nblocks - base) : 1024;
      for (int b = t; b < nb; b += blockDim.x) {
        const double mb = partials[(size_t)(base + b) * 4];
        st[b] = (mb > -INFINITY) ? partials[(size_t)(base + b) * 4 + 1] * exp(mb - Mg) : 0.0;
      }
      __syncthreads();
      if (t == 0)
        for (int b = 0; b < nb; ++b) S += st[b];
      __syncthreads();
    }
    if (t == 0) {
      coll[8] = S;
      coll[9] = (double)counters[CNT_NOK];
      coll[10] = (double)counters[CNT_REJECT];
      coll[11] = (double)counters[CNT_DRAWFAIL];
      coll[12] = coll[13] = coll[14] = coll[15] = 0.0;
    }
    __syncthreads();
  }
  if (phases & 16) {  // publish: own slot = coll[0..15], the others zero
    for (int j = t; j < nranks * 16; j += blockDim.x) gath[j] = (j >> 4) == rank ? coll[j & 15] : 0.0;
  }
  if (phases & 8) {  // combine the gathered slots (after the all-reduce), rank order
    if (t == 0) {
      double M = -INFINITY, R = -INFINITY, f[5] = {0, 0, 0, 0, 0}, S = 0.0, cnt[3] = {0, 0, 0};
      for (int r = 0; r < nranks; ++r) {
        const double *g = gath + 16 * r;
        if (g[0] > M) M = g[0];
        if (g[1] > R) R = g[1];
        for (int j = 0; j < 5; ++j) if (g[2 + j] > f[j]) f[j] = g[2 + j];
      }
      for (int r = 0; r < nranks; ++r) {
        const double *g = gath + 16 * r;
        if (g[0] > -INFINITY) S += g[8] * exp(g[0] - M);
        for (int j = 0; j < 3; ++j) cnt[j] += g[9 + j];
      }
      coll[0] = M; coll[1] = R;
      for (int j = 0; j < 5; ++j) coll[2 + j] = f[j];
      coll[7] = 0.0;
      coll[8] = S; coll[9] = cnt[0]; coll[10] = cnt[1]; coll[11] = cnt[2];
    }
    __syncthreads();
  }
  if ((phases & 4) && t == 0) {
    const double M = coll[0], R = coll[1], S = coll[8];
    // the reference's running maxima with their unconditional assignment at i == 0 (pmc.c:535,552,579)
    const double maxW = coll[2] != 0.0 ? NAN : ((coll[6] != 0.0) ? INFINITY : (coll[4] != 0.0 ? M : (M > kNegInit ? M : kNegInit)));
    const double maxR = coll[3] != 0.0 ? NAN : (coll[5] != 0.0 ? R : (R > kNegInit ? R : kNegInit));
    const bool draw_ok = !(coll[11] > 0.0) && !(coll[10] > 5.0 * n_total);
    const bool clean = draw_ok && coll[2] == 0.0 && coll[3] == 0.0 && coll[6] == 0.0 && M > kNegInit && M < INFINITY && S > 0.0 && S < INFINITY &&
                       maxW == M;
    const double Sref = S * exp(kDynMax);  // sum exp(lw - maxW + 500), what pmc.c:660-674 accumulates
    scal[I3_MAXW] = maxW;
    scal[I3_MAXR] = maxR;
    scal[I3_M] = M;
    scal[I3_SREF] = Sref;
    scal[I3_INVS] = 1.0 / Sref;
    scal[I3_LOGSREF] = log(S) + kDynMax;
    scal[I3_CLEAN] = clean ? 1.0 : 0.0;
    scal[I3_NOK_W] = coll[9];
    scal[I3_NREJ] = coll[10];
    scal[I3_NFAIL] = coll[11];
  }
}

// =====================================================================================================================
// statistics + normalisation
// =====================================================================================================================
__device__ __forceinline__ void dmma884_3(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cpa16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cpa4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

struct S3Args {
  const double *X, *resp;
  double *W;   // in: log-weights, out: normalised weights (pmc.c:660-682)
  short *flg;  // in/out: samples whose weight is not finite are flagged out (pmc.c:667-672)
  const unsigned long long *indices;
  const double *scal;
  double pivot[FUSED_MAXD];
  long N;
  double *partials;  // [grid][partial_len]: DMMA accumulator fragments | H, Q, ok, flagged-out | count_k[64]
  int partial_len;
};

// Tiling of the statistics product C[k][e] += g[i][k] P[i][e] with DMMA.8x8x4 (M = components, N = features, K = samples).
// Full 8-component row tiles cost NT DMMAs per 4 samples.  A remainder of R = K mod 8 components would pad to a whole
// row tile (K = 10: 37 % of the tensor work was padding, ncu r01).  For R = 1 or 2 the remainder instead uses the product
// structure of the features: with the augmented row z = (1, x~_1 .. x~_D), P[(a,b)] = z_a z_b, so
//     D[m = b][n = (r, a)] += sum_i z_i[b] * (g_i[r] z_i[a])
// gives W, s and S of the remaining components from blocks (a-group of AG = 8/R, b-group of 8); by symmetry only blocks
// that contain an entry with a >= b are issued (d = 10, K = 10: 4 DMMAs instead of 9 -> 13 instead of 18 per 4 samples).
template <int D, int K>
struct S3Cfg {
  static constexpr int E = 1 + D + D * (D + 1) / 2;
  static constexpr int NT = (E + 7) / 8;
  static constexpr int R = K % 8;
  static constexpr bool XT = (R == 1 || R == 2) && K > 8;
  static constexpr int MT = XT ? K / 8 : (K + 7) / 8;  // full row tiles
  static constexpr int AG = XT ? 8 / R : 8;
  static constexpr int NGA = (D + AG) / AG, NGB = (D + 8) / 8;  // ceil((D+1)/AG), ceil((D+1)/8)
  __host__ __device__ static constexpr bool xinc(int ga, int gb) { return AG * (ga + 1) - 1 >= 8 * gb; }
  static constexpr int xcount() {
    int n = 0;
    for (int ga = 0; ga < NGA; ++ga)
      for (int gb = 0; gb < NGB; ++gb) n += xinc(ga, gb) ? 1 : 0;
    return n;
  }
  static constexpr int NX = XT ? xcount() : 0;
  // cyclic feature layout: the NI (NI + 1) / 2 products z_a z_b as (a, s) = z_a z_{(a+s) mod NI} with s < SH = NI/2 + 1;
  // tile s holds a = 0..7 (lane row fr = a), the a >= 8 take NHI further tiles with lane row fr = s.  A lane then needs
  // only the SH consecutive values z_fr .. z_{fr+SH-1} for its first SH products (6 loads instead of 12 at d = 10).
  // Used when it needs no more tiles than the row-major layout (d = 6, 8, 10).  The rows are then stored periodically
  // extended (z_{i mod NI}, i < RL), so that every fragment operand of every k-step is ONE per-lane base address plus a
  // compile-time offset: no address arithmetic in the DMMA loop and no registers spent on per-tile offsets.
  static constexpr int NI = D + 1, SH = NI / 2 + 1, NHI = NI > 8 ? NI - 8 : 0;
  static constexpr bool CYC = (SH + NHI == (1 + D + D * (D + 1) / 2 + 7) / 8) && SH <= 8;
  static constexpr int RL = CYC ? (NI > 8 ? NI + 7 : 7 + SH) : D + 2;  // row length: largest index read + 1
  static constexpr int XS = RL | 1;  // odd stride: conflict-free fragment loads (non-cyclic: 1, x~_1 .. x~_D, 0)
  static constexpr int GS = (XT ? K : ((K + 7) / 8) * 8) | 1;  // g row stride (odd): full row tiles read 8 MT entries, the product tiles K
  static constexpr int ACC = (MT * NT + NX) * 64;
  static constexpr int STAGE = 32 * D + 32 * K + 32 + 32 + 8;  // next tile, by cp.async: X rows | resp (k-major) | lw | indices | flg
  static constexpr int WARP_DOUBLES = 32 * XS + 32 * GS + STAGE;

  static constexpr int PLEN = ACC + 4 + 64;
};

// per-sample normalisation shared by k_stats3 and k_norm3; returns the normalised weight if the sample is usable, else 0
__device__ __forceinline__ double norm_sample(double lwv, short f, double maxW, double invS, double logSref,
                                              const double *s_exp, double *Wout, short *Fout, double &H, double &Q,
                                              unsigned int &nok, unsigned int &nflag, bool &ok) {
  ok = false;
  if (f != 0 && lwv == lwv) {
    const double t2 = (lwv - maxW) + kDynMax;
    const double w = exp_tab(t2, s_exp) * invS;
    *Wout = w;
    ok = true;
    ++nok;
    if (w >= kPerpZero) { H = fma(w, t2 - logSref, H); Q = fma(w, w, Q); }
    return w;
  }
  if (f != 0) *Fout = 0;  // NaN log-weight: exp is not finite, the reference flags the sample out and skips it
  *Wout = lwv * invS;     // pmc.c:680-682 divides every entry, flagged or not
  ++nflag;
  return 0.0;
}

template <int D, int K, bool STATS>
__global__ void __launch_bounds__(256, 2) k_stats3(const __grid_constant__ S3Args a) {
  using C = S3Cfg<D, K>;
  constexpr int NT = C::NT, MT = C::MT, GS = C::GS, XS = C::XS, NI = C::NI, SH = C::SH, NHI = C::NHI;
  constexpr int WARPS = 8;
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_exp[64];
  __shared__ double s_red[WARPS][2];
  __shared__ unsigned int s_cnt[WARPS][2];
  __shared__ unsigned int s_count[64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (a.scal[I3_CLEAN] == 0.0) return;  // poisoned maxima: the host runs the literal sequence instead
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_exp[i] = c3_exptab[i]; s_count[i] = 0; }
  const double maxW = a.scal[I3_MAXW], invS = a.scal[I3_INVS], logSref = a.scal[I3_LOGSREF];
  double *xh = smem + (size_t)wid * C::WARP_DOUBLES;  // augmented rows z = (1, x~_1 .. x~_D, 0) of the warp's 32 samples
  double *gt = xh + 32 * XS;                          // g rows: w_i * responsibility_k
  double *stx = gt + 32 * GS;                         // staging of the NEXT tile (cp.async, no registers held across the DMMA phase)
  double *str = stx + 32 * D;
  double *stw = str + 32 * K;
  unsigned long long *sti = reinterpret_cast<unsigned long long *>(stw + 32);
  short *stf = reinterpret_cast<short *>(sti + 32);
  for (int i = lane; i < C::WARP_DOUBLES; i += 32) xh[i] = 0.0;
  __syncthreads();
  const int fr = lane >> 2, fc = lane & 3;
  // B-fragment operands.  Cyclic layout (C::CYC): feature (a, s) = z_a z_{(a+s) mod NI}, s < SH; a lane loads the SH values
  // z_{fr}, z_{fr+1}, .. once per k-step and forms SH products from them, the features with a >= 8 take NHI more tiles
  // (lane fr <-> shift s = fr).  Otherwise the row-major upper-triangular layout with two loads per product.
  int oa[NT], ob[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    int pa = 0, pb = 0;
    if (C::CYC) {
      if (n < SH) { pa = fr % NI; pb = (fr + n) % NI; }
      else { pa = 8 + (n - SH); pb = (pa + fr) % NI; }
    } else {
      const int e = 8 * n + fr;
      if (e >= 1 && e <= D) pb = e;
      else if (e > D && e < C::E) {
        int q = e - 1 - D, r = 0;
        while (q >= D - r) { q -= D - r; ++r; }
        pa = 1 + r;
        pb = 1 + r + q;
      }
    }
    oa[n] = pa;
    ob[n] = pb;
  }
  double acc[STATS ? MT : 1][STATS ? NT : 1][2];
#pragma unroll
  for (int m = 0; m < (STATS ? MT : 1); ++m)
#pragma unroll
    for (int n = 0; n < (STATS ? NT : 1); ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  // remainder components (C::XT): product tiles, see S3Cfg
  constexpr int NXA = (STATS && C::XT) ? C::NX : 1;
  double accx[NXA][2];
#pragma unroll
  for (int b = 0; b < NXA; ++b) accx[b][0] = accx[b][1] = 0.0;
  int xao[C::NGA], xbo[C::NGB];
#pragma unroll
  for (int g = 0; g < C::NGA; ++g) { const int ai = C::AG * g + (fr % C::AG); xao[g] = ai <= D ? ai : D + 1; }
#pragma unroll
  for (int g = 0; g < C::NGB; ++g) { const int bi = 8 * g + fr; xbo[g] = bi <= D ? bi : D + 1; }
  const int gxo = 8 * MT + fr / C::AG;
  double H = 0.0, Q = 0.0;
  unsigned int nok = 0, nflag = 0;
  const long ntiles = (a.N + 31) / 32;
  const long tstep = (long)gridDim.x * WARPS;

  auto prefetch = [&](long tile) {
    const long base = tile * 32;
    const int tn = (a.N - base) < 32 ? (int)(a.N - base) : 32;
    if (STATS) {
      if ((D & 1) == 0) {
        for (int c = lane; c < tn * D / 2; c += 32) cpa16(stx + 2 * c, a.X + (size_t)base * D + 2 * c);
      } else {
        for (int c = lane; c < tn * D; c += 32) cpa8(stx + c, a.X + (size_t)base * D + c);
      }
    }
    if (lane < tn) {
      if (STATS) {
#pragma unroll
        for (int k = 0; k < K; ++k) cpa8(str + k * 32 + lane, a.resp + (size_t)k * a.N + base + lane);
      }
      cpa8(stw + lane, a.W + base + lane);
      cpa8(sti + lane, a.indices + base + lane);
    }
    if (2 * lane + 1 < tn) cpa4(stf + 2 * lane, a.flg + base + 2 * lane);
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
    if (i < a.N) {
      bool ok;
      w = norm_sample(stw[lane], stf[lane], maxW, invS, logSref, s_exp, a.W + i, a.flg + i, H, Q, nok, nflag, ok);
      if (ok) atomicAdd(&s_count[(int)sti[lane] & 63], 1u);
    }
    if (STATS) {
      double *xrow = xh + lane * XS, *grow = gt + lane * GS;
      const bool use = w != 0.0;
      auto put = [&](int idx, double v) {  // z_idx, and its periodic copies in the cyclic layout
        xrow[idx] = v;
        if (C::CYC && idx + NI < C::RL) xrow[idx + NI] = v;
      };
      put(0, 1.0);
      if ((D & 1) == 0) {
        const double2 *xi = reinterpret_cast<const double2 *>(stx + lane * D);  // 80-byte rows: conflict-free 16-byte loads
#pragma unroll
        for (int j = 0; j < D; j += 2) {
          const double2 t2 = xi[j / 2];
          put(1 + j, use ? t2.x - a.pivot[j] : 0.0);
          put(2 + j, use ? t2.y - a.pivot[j + 1] : 0.0);
        }
      } else {
#pragma unroll
        for (int j = 0; j < D; ++j) put(1 + j, use ? stx[lane * D + j] - a.pivot[j] : 0.0);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) grow[k] = use ? w * str[k * 32 + lane] : 0.0;
    }
    __syncwarp();
    if (tile + tstep < ntiles) prefetch(tile + tstep);  // lands while the DMMA phase below runs
    if (STATS) {
      // per-lane bases: sample 4 fc of the k-step, row element fr; every operand below is base + compile-time offset
      const double *xb0 = xh + (4 * fc) * XS + fr, *xb1 = xh + (4 * fc) * XS + (fr % C::AG);
      const double *gb0 = gt + (4 * fc) * GS + fr, *gb1 = gt + (4 * fc) * GS + gxo;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int so = 16 * h + t;  // sample offset of the k-step (added to 4 fc)
          const int smp = so + 4 * fc;
          double af[MT];
#pragma unroll
          for (int m = 0; m < MT; ++m) af[m] = gb0[so * GS + 8 * m];
          const double *xs = xh + smp * XS;
          double u0 = 0.0;
          if constexpr (C::CYC) {
            double u[SH];
#pragma unroll
            for (int n = 0; n < SH; ++n) u[n] = xb0[so * XS + n];
            u0 = u[0];
#pragma unroll
            for (int n = 0; n < SH; ++n) {
              const double bf = u[0] * u[n];
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma884_3(acc[m][n][0], acc[m][n][1], af[m], bf);
            }
#pragma unroll
            for (int n = SH; n < NT; ++n) {
              const double bf = xs[8 + (n - SH)] * xb0[so * XS + 8 + (n - SH)];
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma884_3(acc[m][n][0], acc[m][n][1], af[m], bf);
            }
          } else {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const double bf = xs[oa[n]] * xs[ob[n]];
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma884_3(acc[m][n][0], acc[m][n][1], af[m], bf);
            }
          }
          if constexpr (C::XT) {
            const double gx = gb1[so * GS];
            double a2[C::NGB];
#pragma unroll
            for (int gb = 0; gb < C::NGB; ++gb) {
              if (C::CYC) a2[gb] = (gb == 0) ? u0 : xb0[so * XS + 8 * gb];  // rows beyond D carry periodic copies: those output rows are never read
              else a2[gb] = xs[xbo[gb]];
            }
            int xb = 0;
#pragma unroll
            for (int ga = 0; ga < C::NGA; ++ga) {
              const double b2 = gx * (C::CYC ? xb1[so * XS + C::AG * ga] : xs[xao[ga]]);
#pragma unroll
              for (int gb = 0; gb < C::NGB; ++gb)
                if (C::xinc(ga, gb)) { dmma884_3(accx[xb][0], accx[xb][1], a2[gb], b2); ++xb; }
            }
          }
        }
      __syncwarp();
    }
  }
  // ---- block combine in warp order -------------------------------------------------------------------------------
  __syncthreads();
  {
    const double Hw = warp_sum(H), Qw = warp_sum(Q);
    const unsigned int c1 = (unsigned int)warp_sum_ll(nok), c2 = (unsigned int)warp_sum_ll(nflag);
    if (lane == 0) { s_red[wid][0] = Hw; s_red[wid][1] = Qw; s_cnt[wid][0] = c1; s_cnt[wid][1] = c2; }
  }
  double *out = a.partials + (size_t)blockIdx.x * a.partial_len;
  if (STATS) {
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
        if constexpr (C::XT) {
#pragma unroll
          for (int b = 0; b < C::NX; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int idx = ((MT * NT + b) * 32 + lane) * 2 + c;
              comb[idx] = (w2 == 0 ? 0.0 : comb[idx]) + accx[b][c];
            }
        }
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < C::ACC; idx += blockDim.x) out[idx] = comb[idx];
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double Hb = 0.0, Qb = 0.0, n1 = 0.0, n2 = 0.0;
    for (int w2 = 0; w2 < WARPS; ++w2) { Hb += s_red[w2][0]; Qb += s_red[w2][1]; n1 += (double)s_cnt[w2][0]; n2 += (double)s_cnt[w2][1]; }
    out[C::ACC + 0] = Hb;
    out[C::ACC + 1] = Qb;
    out[C::ACC + 2] = n1;
    out[C::ACC + 3] = n2;
  }
  if (threadIdx.x < 64) out[C::ACC + 4 + threadIdx.x] = (double)s_count[threadIdx.x];
}

// canonical statistics from the block partials, in block order: out = stats[K*E] | H, Q, ok, flagged-out | count_k[K]
// Components below 8*MT sit in the (component, feature) row tiles; the others (ag > 0) in the product tiles of S3Cfg.
// One WARP per output value: lane l adds the partials of blocks l, l + 32, .. in that order, the 32 lane sums are combined by
// a fixed shuffle tree -> the same association order every run (one thread per value walking all blocks took 55 us).
__global__ void k_comb3(const double *__restrict__ partials, int nblocks, int partial_len, int acc_len, int K, int E, int NT,
                        int MT, int D, int ag, int cyc, int with_stats, const double *__restrict__ scal, double *__restrict__ out) {
  const int nstat = with_stats ? K * E : 0;
  const int total = nstat + 4 + K;
  const bool clean = scal[I3_CLEAN] != 0.0;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < total; t += gridDim.x * wpb) {
    int idx;
    if (t < nstat) {
      const int k = t / E, e = t - k * E;
      int pa = 0, pb = e;  // canonical feature e = z_pa z_pb, pa <= pb (z_0 = 1): W, s_1..s_D, then the upper triangle row by row
      if (e > D) {
        int q = e - 1 - D, r = 0;
        while (q >= D - r) { q -= D - r; ++r; }
        pa = 1 + r;
        pb = 1 + r + q;
      }
      if (ag == 0 || k < 8 * MT) {
        const int m = k >> 3, r = k & 7;
        int n = e >> 3, cc = e & 7;
        if (cyc) {  // cyclic layout of S3Cfg: (a, s) with z_b = z_{(a+s) mod NI}
          const int NI = D + 1, SH = NI / 2 + 1, s1 = pb - pa;
          const int a_ = s1 < SH ? pa : pb, s_ = s1 < SH ? s1 : NI - s1;
          n = a_ < 8 ? s_ : SH + a_ - 8;
          cc = a_ < 8 ? a_ : s_;
        }
        idx = ((m * NT + n) * 32 + (r * 4 + (cc >> 1))) * 2 + (cc & 1);
      } else {
        const int a = pb, b = pa;  // a >= b
        const int ga = a / ag, gb = b >> 3, ngb = (D + 8) / 8;
        int xb = 0;  // issued blocks before (ga, gb): a-groups outer, b-groups inner, block issued iff ag*(ga+1)-1 >= 8*gb
        for (int g1 = 0; g1 <= ga; ++g1)
          for (int g2 = 0; g2 < (g1 < ga ? ngb : gb); ++g2) xb += (ag * (g1 + 1) - 1 >= 8 * g2) ? 1 : 0;
        const int row = b & 7, col = (k - 8 * MT) * ag + a % ag;
        idx = ((MT * NT + xb) * 32 + (row * 4 + (col >> 1))) * 2 + (col & 1);
      }
    } else {
      idx = acc_len + (t - nstat);
    }
    double s = 0.0;
    if (clean)
      for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * partial_len + idx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[t] = s;
  }
}

// =====================================================================================================================
// draw
// =====================================================================================================================
struct D3Args {
  long N, first_index;
  uint64_t seed;
  uint32_t iter;
  int max_attempts, chunk;
  int has_box;
  double *X;
  unsigned long long *indices;
  short *flg;
  const double2 *zig;
  const double *gpack;  // the draw pack in global memory: per-lane component of a redraw
  unsigned long long *counters;
};

// attempts >= 1 of a sample whose first draw fell outside the box (pmc.c:281-295: a fresh component every time)
template <int D, int K>
static __device__ __noinline__ bool redraw3(const D3Args &a, const Philox ph, uint32_t ilo, uint32_t ihi, const double2 *zts,
                                            double *x, int &kout, unsigned int &rej) {
  constexpr int DS = i3_dstride(D);
  for (int attempt = 1; attempt < a.max_attempts; ++attempt) {
    const uint32_t cb = (uint32_t)attempt << 12;
    const uint4 r0 = ph(ilo, ihi, cb, a.iter);
    const double u = u01_co(r0.x, r0.y);
    int k = 0;
    for (int kk = 0; kk < K - 1; ++kk) k += (c3_cw[kk] < u) ? 1 : 0;
    double g[D + 1];
    uint32_t spare;
    zig_normals<D>(ph, ilo, ihi, cb, a.iter, zts, g, spare);
    const double *cp = a.gpack + (size_t)k * DS;
    bool inb = true;
    int e = 0;
    for (int i = 0; i < D; ++i) {
      double t = cp[e++];
      for (int j = 0; j <= i; ++j) t = fma(g[j], cp[e++], t);
      x[i] = t;
      inb = inb && isfinite(t);
      if (a.has_box) inb = inb && !(t < c3_box[2 * i]) && !(t > c3_box[2 * i + 1]);
    }
    kout = k;
    if (inb) return true;
    ++rej;
  }
  return false;
}

// x = mu_k + L_k g for a warp-uniform k: one straight-line copy per component so that every parameter is a constant-bank
// operand at a compile-time offset (a register-indexed constant load per entry cost an instruction and two registers each)
template <int D, int K>
__device__ __forceinline__ void apply_L3(int k, const double (&g)[D + 1], double (&x)[D]) {
  constexpr int DS = i3_dstride(D);
#pragma unroll
  for (int kk = 0; kk < K; ++kk) {
    if (k == kk) {
      int e = kk * DS;
#pragma unroll
      for (int i2 = 0; i2 < D; ++i2) {
        const int em = e++;
        double t = g[0] * c3_d[e++];  // one constant-bank operand per instruction (a second one costs a register load)
#pragma unroll
        for (int j2 = 1; j2 <= i2; ++j2) t = fma(g[j2], c3_d[e++], t);
        x[i2] = t + c3_d[em];
      }
    }
  }
}

template <int D, int K>
__global__ void __launch_bounds__(256, 2) k_draw3(const __grid_constant__ D3Args a) {
  constexpr int DS = i3_dstride(D);
  static_assert(K * DS <= I3_CD && K <= 64, "constant memory budget");
  extern __shared__ __align__(16) unsigned char sm3[];
  double2 *zts = reinterpret_cast<double2 *>(sm3);
  unsigned int *tmp = reinterpret_cast<unsigned int *>(zts + kZigLayers);
  unsigned short *perm = reinterpret_cast<unsigned short *>(tmp + a.chunk);
  __shared__ unsigned int s_hist[64], s_base[65], s_tile[65];
  __shared__ int s_next;
  int tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));  // read once: ptxas otherwise re-reads the special register in the loop
  const int lane = tid & 31;
  for (int i = tid; i < kZigLayers; i += blockDim.x) zts[i] = a.zig[i];
  const Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  unsigned int rej_lane = 0, fail_lane = 0;
  const long nchunks = (a.N + a.chunk - 1) / a.chunk;
  for (long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long cbase = c * a.chunk;
    const int cn = (a.N - cbase) < a.chunk ? (int)(a.N - cbase) : a.chunk;
    if (tid < 64) s_hist[tid] = 0;
    if (tid == 64) s_next = 0;
    __syncthreads();
    // ---- component of the first attempt of every sample of the chunk (mvdens.c:775-776), rank inside its bucket --------
    for (int o = tid; o < cn; o += blockDim.x) {
      const uint64_t gi = (uint64_t)(a.first_index + cbase + o);
      const uint4 r0 = ph((uint32_t)gi, (uint32_t)(gi >> 32), 0u, a.iter);
      const double u = u01_co(r0.x, r0.y);
      int k = 0;
#pragma unroll
      for (int kk = 0; kk < K - 1; ++kk) k += (c3_cw[kk] < u) ? 1 : 0;
      const unsigned int rank = atomicAdd(&s_hist[k], 1u);
      tmp[o] = ((unsigned int)k << 16) | rank;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int acc = 0, tacc = 0;
      for (int k = 0; k < K; ++k) { s_base[k] = acc; acc += s_hist[k]; s_tile[k] = tacc; tacc += (s_hist[k] + 31) >> 5; }
      s_base[K] = acc;
      s_tile[K] = tacc;
    }
    __syncthreads();
    for (int o = tid; o < cn; o += blockDim.x) {
      const unsigned int v = tmp[o];
      perm[s_base[v >> 16] + (v & 0xffffu)] = (unsigned short)o;
    }
    __syncthreads();
    // ---- one component per warp tile: x = mu_k + L_k g with warp-uniform constant-bank operands.  Tiles are handed out
    // by a shared counter, so the warps of the block stay busy until the chunk is finished.
    const int ntile = (int)s_tile[K];
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&s_next, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= ntile) break;
      int k = 0;
#pragma unroll
      for (int kk = 1; kk < K; ++kk) k += ((int)s_tile[kk] <= t) ? 1 : 0;
      const int nk = (int)s_hist[k], b = (int)s_base[k];
      {
        const int j = (t - (int)s_tile[k]) * 32 + lane;
        const bool active = j < nk;
        const int o = (int)perm[b + (active ? j : 0)];
        const long i = cbase + o;
        const uint64_t gi = (uint64_t)(a.first_index + i);
        const uint32_t ilo = (uint32_t)gi, ihi = (uint32_t)(gi >> 32);
        double g[D + 1];
        uint32_t spare;
        zig_normals<D>(ph, ilo, ihi, 0u, a.iter, zts, g, spare);
        double x[D];
        apply_L3<D, K>(k, g, x);
        bool inb = true;
#pragma unroll
        for (int i2 = 0; i2 < D; ++i2) {
          inb = inb && isfinite(x[i2]);
          // thr - c*x < 0 (parabox.c:125-131) for c = -1 / +1 equals a direct comparison with the (negated) threshold
          if (a.has_box) inb = inb && !(x[i2] < c3_box[2 * i2]) && !(x[i2] > c3_box[2 * i2 + 1]);
        }
        if (active) {
          int kk = k;
          bool ok = inb;
          if (!inb) {
            ++rej_lane;
            ok = redraw3<D, K>(a, ph, ilo, ihi, zts, x, kk, rej_lane);
            if (!ok) ++fail_lane;
          }
          a.indices[i] = (unsigned long long)kk;
          a.flg[i] = ok ? 1 : 0;
          if ((D & 1) == 0) {
            double2 *xo = reinterpret_cast<double2 *>(a.X + (size_t)i * D);
#pragma unroll
            for (int i2 = 0; i2 < D; i2 += 2) xo[i2 / 2] = make_double2(x[i2], x[i2 + 1]);
          } else {
#pragma unroll
            for (int i2 = 0; i2 < D; ++i2) a.X[(size_t)i * D + i2] = x[i2];
          }
        }
      }
    }
    __syncthreads();
  }
  const unsigned long long n2 = (unsigned long long)warp_sum_ll(rej_lane), n3 = (unsigned long long)warp_sum_ll(fail_lane);
  if (lane == 0) {
    if (n2) atomicAdd(&a.counters[CNT_REJECT], n2);
    if (n3) atomicAdd(&a.counters[CNT_DRAWFAIL], n3);
  }
}

// =====================================================================================================================
// host side: instantiation table and launchers
// =====================================================================================================================
#define PMC_ITER3_TABLE(X) \
  X(10, 10, 0)             \
  X(10, 20, 0)             \
  X(10, 5, 0)              \
  X(5, 20, 4)              \
  X(5, 10, 0)              \
  X(2, 9, 0)

Iter3Plan iter3_plan(int d, int K, int Kt, int tgt_kind, int sm_count) {
  Iter3Plan best{};
  const int kt_needed = (tgt_kind == PMCGPU_TGT_MVDENS || tgt_kind == PMCGPU_TGT_MIX) ? Kt : 0;
  // the smallest instantiation with K_ >= K serves the mixture: the extra components are packed as dead ones
#define X(D_, K_, KT_)                                                                                          \
  if (d == D_ && K <= K_ && (kt_needed == 0 || (KT_ > 0 && kt_needed <= KT_)) && (!best.valid || K_ < best.K)) { \
    using C = S3Cfg<D_, K_>;                                                                                     \
    best.valid = 1; best.D = D_; best.K = K_; best.KT = KT_; best.sm_count = sm_count;                           \
    best.E = C::E; best.NT = C::NT; best.MT = C::MT; best.AG = C::XT ? C::AG : 0; best.CYC = C::CYC ? 1 : 0;                                \
    best.acc_len = C::ACC; best.partial_len = C::PLEN;                                                          \
    best.stats_blocks = sm_count * 2; best.weights_blocks = sm_count * 4; /* upper bound over the variants */                                        \
  }
  PMC_ITER3_TABLE(X)
#undef X
  return best;
}

// variant 0: one sample per lane, 4 blocks/SM (128 registers); 1: one sample per lane, 3 blocks/SM (168 registers);
// 2: two samples per lane with the component log-densities streamed through shared memory, 4 blocks/SM (128 registers).
// Variants 1 and 2 are instantiated for the headline shape only.
template <int D, int K, int KT>
static int launch_w3(const Iter3Plan &p, const W3Args &a, int variant, bool resp, int *nblocks, cudaStream_t s) {
  constexpr bool kHead = (D == 10 && K == 10 && KT == 0);
  if (!kHead && variant != 3) variant = 0;
  const int spl = variant == 2 ? 2 : 1, minb = variant == 1 ? 3 : 4;
  const size_t dyn2 = (size_t)4 * K * 2 * 32 * sizeof(double);  // variant 2: the streamed component log-densities
  const long ntiles = (a.N + 32 * spl - 1) / (32 * spl);
  int grid = p.sm_count * minb;
  if ((long)grid * 4 > ntiles) grid = (int)((ntiles + 3) / 4);
  if (grid < 1) grid = 1;
  *nblocks = grid;
  if (cudaMemcpyToSymbolAsync(c3_wargs, &a, sizeof a, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (variant == 0) {
    if (resp) k_weights3<D, K, KT, 1, true, 4><<<grid, 128, 0, s>>>();
    else k_weights3<D, K, KT, 1, false, 4><<<grid, 128, 0, s>>>();
  } else if (variant == 3) {  // row-scaled solve (the pack was built with pre = 1)
    if (resp) k_weights3<D, K, KT, 1, true, 4, true><<<grid, 128, 0, s>>>();
    else k_weights3<D, K, KT, 1, false, 4, true><<<grid, 128, 0, s>>>();
  }
  if constexpr (kHead) {
    if (variant == 1) {
      if (resp) k_weights3<D, K, KT, 1, true, 3><<<grid, 128, 0, s>>>();
      else k_weights3<D, K, KT, 1, false, 3><<<grid, 128, 0, s>>>();
    } else if (variant == 2) {
      if (resp) k_weights3<D, K, KT, 2, true, 4><<<grid, 128, dyn2, s>>>();
      else k_weights3<D, K, KT, 2, false, 4><<<grid, 128, dyn2, s>>>();
    }
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int iter3_upload(const Iter3Plan &p, const double *h_w, int nw, const double *h_t, int nt, const double *h_d, int nd,
                 const double *h_cw, int ncw, cudaStream_t s) {
  (void)p;
  if (upload_tables()) return -1;
  if (h_w && cudaMemcpyToSymbolAsync(c3_w, h_w, sizeof(double) * nw, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (h_t && nt && cudaMemcpyToSymbolAsync(c3_t, h_t, sizeof(double) * nt, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (h_d && cudaMemcpyToSymbolAsync(c3_d, h_d, sizeof(double) * nd, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  if (h_cw && cudaMemcpyToSymbolAsync(c3_cw, h_cw, sizeof(double) * ncw, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  return 0;
}

int launch_weights3(const Iter3Plan &p, const Iter3Weights &w, int *nblocks, cudaStream_t s) {
  W3Args a{};
  a.X = w.X; a.N = w.N; a.first_index = w.first_index; a.tgt_kind = w.tgt_kind; a.b = w.b; a.sig1 = w.sig1; a.t_temper = w.t_temper;
  a.post_in = w.post_in; a.ok_in = w.ok_in;
  a.lw = w.lw; a.log_rho = w.log_rho; a.resp = w.resp; a.flg = w.flg; a.partials = w.partials; a.counters = w.counters;
  a.first_vals = w.first_vals; a.mix = w.mix; a.tmix = w.tmix;
  for (int j = 0; j < FUSED_MAXD; ++j) a.shift[j] = w.shift[j];
  const bool resp = w.resp != nullptr;
#define X(D_, K_, KT_) \
  if (p.D == D_ && p.K == K_ && p.KT == KT_) return launch_w3<D_, K_, KT_>(p, a, w.variant, resp, nblocks, s);
  PMC_ITER3_TABLE(X)
#undef X
  return -1;
}

int launch_red3(const double *partials, int nblocks, const unsigned long long *counters, const double *first_vals,
                int has_first, double *coll, double *scal, int phases, double n_total, cudaStream_t s, double *gath, int rank,
                int nranks) {
  k_red3<<<1, 256, 0, s>>>(partials, nblocks, counters, first_vals, has_first, coll, scal, phases, n_total, gath, rank, nranks);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int D, int K>
static int launch_s3(const Iter3Plan &p, const S3Args &a, bool stats, cudaStream_t s) {
  using C = S3Cfg<D, K>;
  size_t smem = (size_t)8 * C::WARP_DOUBLES * sizeof(double);
  if (smem < (size_t)C::ACC * sizeof(double)) smem = (size_t)C::ACC * sizeof(double);
  static std::atomic<unsigned long long> set1{0}, set0{0};
  if (stats) {
    if (!set_max_smem_once(k_stats3<D, K, true>, smem, set1)) return -1;
    k_stats3<D, K, true><<<p.stats_blocks, 256, smem, s>>>(a);
  } else {
    if (!set_max_smem_once(k_stats3<D, K, false>, smem, set0)) return -1;
    k_stats3<D, K, false><<<p.stats_blocks, 256, smem, s>>>(a);
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_stats3(const Iter3Plan &p, const Iter3Stats &st, cudaStream_t s) {
  S3Args a{};
  a.X = st.X; a.resp = st.resp; a.W = st.W; a.flg = st.flg; a.indices = st.indices; a.scal = st.scal; a.N = st.N;
  a.partials = st.partials; a.partial_len = p.partial_len;
  for (int j = 0; j < FUSED_MAXD; ++j) a.pivot[j] = st.pivot[j];
#define X(D_, K_, KT_) \
  if (p.D == D_ && p.K == K_ && p.KT == KT_) return launch_s3<D_, K_>(p, a, st.with_stats != 0, s);
  PMC_ITER3_TABLE(X)
#undef X
  return -1;
}

int launch_comb3(const Iter3Plan &p, const double *partials, int with_stats, const double *scal, double *out, cudaStream_t s) {
  const int total = (with_stats ? p.K * p.E : 0) + 4 + p.K;
  k_comb3<<<(total + 7) / 8, 256, 0, s>>>(partials, p.stats_blocks, p.partial_len, p.acc_len, p.K, p.E, p.NT, p.MT, p.D, p.AG, p.CYC, with_stats, scal, out);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

struct D3Host { D3Args a; double lo_thr[FUSED_MAXD], hi_thr[FUSED_MAXD]; };
template <int D, int K>
static int launch_d3(const Iter3Plan &p, const D3Host &a0, cudaStream_t s) {
  D3Args a = a0.a;
  // chunk: as large as shared memory comfortably allows (few partially filled warp tiles), but small enough that
  // every SM gets work when N is small
  long chunk = 8192;
  const long want = (a.N + (long)p.sm_count * 2 - 1) / ((long)p.sm_count * 2);
  if (want < chunk) chunk = ((want + 255) / 256) * 256;
  if (chunk < 256) chunk = 256;
  a.chunk = (int)chunk;
  const size_t smem = (size_t)kZigLayers * sizeof(double2) + (size_t)chunk * 6;
  static std::atomic<unsigned long long> set{0};
  if (!set_max_smem_once(k_draw3<D, K>, (size_t)kZigLayers * sizeof(double2) + 8192 * 6, set)) return -1;
  const long nchunks = (a.N + chunk - 1) / chunk;
  long grid = (long)p.sm_count * 2;
  if (grid > nchunks) grid = nchunks;
  {
    double hb[2 * FUSED_MAXD];
    for (int j = 0; j < FUSED_MAXD; ++j) { hb[2 * j] = -a0.lo_thr[j]; hb[2 * j + 1] = a0.hi_thr[j]; }
    if (cudaMemcpyToSymbolAsync(c3_box, hb, sizeof hb, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
  }
  k_draw3<D, K><<<(int)grid, 256, smem, s>>>(a);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_draw3(const Iter3Plan &p, const Iter3Draw &d, cudaStream_t s) {
  D3Host h{};
  D3Args &a = h.a;
  a.N = d.N; a.first_index = d.first_index; a.seed = d.seed; a.iter = d.iter; a.max_attempts = d.max_attempts;
  a.has_box = d.has_box;
  for (int j = 0; j < FUSED_MAXD; ++j) { h.lo_thr[j] = d.lo_thr[j]; h.hi_thr[j] = d.hi_thr[j]; }
  a.X = d.X; a.indices = d.indices; a.flg = d.flg; a.zig = d.zig; a.gpack = d.gpack; a.counters = d.counters;
#define X(D_, K_, KT_) \
  if (p.D == D_ && p.K == K_ && p.KT == KT_) return launch_d3<D_, K_>(p, h, s);
  PMC_ITER3_TABLE(X)
#undef X
  return -1;
}

}  // namespace pmcb200
