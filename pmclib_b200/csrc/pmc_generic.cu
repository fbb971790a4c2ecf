// pmc_generic.cu — the any-(d,K) kernels: one thread per sample, mixture read from global memory in reference
// layout, every step the literal operation sequence of the reference (divisions kept, two-pass RB update).
// They are the parity backstop for shapes the fused small-d kernel does not cover (Student-t, d > 16, large K)
// and the "precise path" the fused iteration falls back to when its pivot guard trips.
#include "pmc_kernels.h"

namespace pmcb200 {

constexpr int MAXD = PMCB200_MAXD;

// ---- literal pieces ------------------------------------------------------------------------------------------
// scalar_product (mvdens.c:319-349) with cblas dtrsv's operation order; y is scratch of length d
__device__ __forceinline__ double whiten_q(const double *__restrict__ L, const double *__restrict__ mu, const double *x,
                                           double *y, int d) {
  for (int i = 0; i < d; ++i) y[i] = x[i] - mu[i];
  y[0] = y[0] / L[0];
  for (int i = 1; i < d; ++i) {
    double t = y[i];
    const double *row = L + (size_t)i * d;
    for (int j = 0; j < i; ++j) t -= row[j] * y[j];
    y[i] = t / row[i];
  }
  double q = 0.0;
  for (int i = 0; i < d; ++i) q += y[i] * y[i];
  return q;
}

__device__ __forceinline__ double comp_log_pdf(const MixDev &m, int k, const double *x, double *y, double *q_out) {
  const double q = whiten_q(m.L + (size_t)k * m.d * m.d, m.mean + (size_t)k * m.d, x, y, m.d);
  if (q_out) *q_out = q;
  return logpdf_from_q(q, m.d, m.logdet[k], m.df[k], m.gam[k], m.norm[k]);
}

// mix_mvdens_log_pdf (mvdens.c:784-825).  ld = scratch[K] in global memory (row of this thread)
__device__ double mix_log_pdf_generic(const MixDev &m, const double *x, double *y, double *ld) {
  int iloc = 0;
  while (iloc < m.K && m.alpha[iloc] <= 0) iloc++;
  if (iloc >= m.K) return nan("");
  double mx = comp_log_pdf(m, iloc, x, y, nullptr);
  ld[iloc] = mx;
  for (int i = iloc + 1; i < m.K; ++i) {
    if (m.alpha[i] > 0.0) {
      const double v = comp_log_pdf(m, i, x, y, nullptr);
      ld[i] = v;
      if (mx < v) mx = v;
    }
  }
  double lpos = 0.0;
  for (int i = 0; i < m.K; ++i) {
    if (m.alpha[i] > 0.0) lpos = lpos + exp(ld[i] - mx + kDynMax) * m.alpha[i];
  }
  return mx + log(lpos) - kDynMax;
}

__device__ double target_log_pdf_generic(const TargetDev &t, const double *xfree, double *xfull, double *y, double *ld) {
  const double *x = xfree;
  if (t.nfull > 0) {  // distribution_fill_pars (distribution.c:193-209)
    int ir = 0;
    for (int ik = 0; ik < t.nfull; ++ik) xfull[ik] = t.is_def[ik] ? t.def_val[ik] : xfree[ir++];
    x = xfull;
  }
  if (t.kind == PMCGPU_TGT_BANANA) return bent_gauss(x, t.d, t.b, t.sig1);
  if (t.kind == PMCGPU_TGT_MVDENS) return comp_log_pdf(t.mix, 0, x, y, nullptr);
  return mix_log_pdf_generic(t.mix, x, y, ld);
}

// ---- draw (pmc.c:266-298, mvdens.c:730-782, 279-317) ---------------------------------------------------------
// list != nullptr: only the samples list[0..N) are (re)drawn, starting at attempt number attempt0 (the tensor-path draw
// hands over the samples whose first attempt fell outside the box)
__global__ void k_draw_generic(MixDev m, BoxDev bx, uint64_t seed, uint32_t iter, long first_index, long N,
                               int max_attempts, double *__restrict__ X, unsigned long long *__restrict__ indices,
                               short *__restrict__ flg, unsigned long long *counters,
                               const unsigned int *__restrict__ list, int attempt0) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= N) return;
  const long i = list ? (long)list[t] : t;
  const int d = m.d;
  double x[MAXD];
  const uint64_t gi = (uint64_t)(first_index + i);
  DrawStream st{Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}, (uint32_t)gi, (uint32_t)(gi >> 32), iter, 0u, 0u};
  unsigned long long rejects = 0;
  int k = 0;
  bool ok = false;
  for (int attempt = attempt0; attempt < max_attempts; ++attempt) {
    st.attempt = attempt;
    st.call = 0;
    const uint4 r0 = st.next();
    k = ranindx(m.cw, m.K, u01_co(r0.x, r0.y));
    for (int j = 0; j < d; j += 2) {
      double g0, g1;
      box_muller(st.next(), g0, g1);
      x[j] = g0;
      if (j + 1 < d) x[j + 1] = g1;
    }
    // res = L g, cblas dtrmv order (i descending, in place)
    const double *L = m.L + (size_t)k * d * d;
    for (int a = d; a > 0 && a--;) {
      double t = 0.0;
      const double *row = L + (size_t)a * d;
      for (int j = 0; j < a; ++j) t += x[j] * row[j];
      x[a] = t + x[a] * row[a];
    }
    double corr = 1.0;
    if (m.df[k] != -1) corr = sqrt(m.df[k] / chisq_mt(st, (double)m.df[k]));
    const double *mu = m.mean + (size_t)k * d;
    bool fin = true;
    for (int a = 0; a < d; ++a) {
      x[a] *= corr;
      x[a] += mu[a];
      fin = fin && isfinite(x[a]);
    }
    if (fin && (bx.nfaces == 0 || in_box(bx, x))) { ok = true; break; }
    ++rejects;
  }
  for (int a = 0; a < d; ++a) X[(size_t)i * d + a] = x[a];
  indices[i] = (unsigned long long)k;
  flg[i] = ok ? 1 : 0;
  if (rejects) atomicAdd(&counters[CNT_REJECT], rejects);
  if (!ok) atomicAdd(&counters[CNT_DRAWFAIL], 1ull);
}

// ---- weights (pmc.c:484-600) ---------------------------------------------------------------------------------
// mode 0: device target; mode 1: target values supplied (post_in / ok_in)
__global__ void k_weights_generic(MixDev m, TargetDev tg, int mode, const double *__restrict__ post_in,
                                  const short *__restrict__ ok_in, double t_temper, long first_index, long N,
                                  const double *__restrict__ X, double *__restrict__ lw, double *__restrict__ log_rho,
                                  short *__restrict__ flg, double *__restrict__ ldscratch, int ldstride,
                                  double *blk_maxw, double *blk_maxr, unsigned long long *counters, double *first_vals) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  double my_w = -INFINITY, my_r = -INFINITY;
  bool okflag = false;
  if (i < N) {
    double x[MAXD], xf[MAXD], y[MAXD];
    const int d = m.d;
    for (int a = 0; a < d; ++a) x[a] = X[(size_t)i * d + a];
    double *ld = ldscratch + (size_t)i * ldstride;
    flg[i] = 0;
    lw[i] = 0;
    const double rloc = mix_log_pdf_generic(m, x, y, ld);
    // pmc.c:552-554: maxR and log_rho are updated before the non-finite rloc is caught at :563
    my_r = rloc;
    log_rho[i] = rloc;
    bool reached_w = false;
    double w = 0.0;
    if (!(isnan(rloc) || isinf(rloc))) {
      bool tok = true;
      double post;
      if (mode == 0) post = target_log_pdf_generic(tg, x, xf, y, ld);
      else { post = post_in[i]; tok = ok_in ? (ok_in[i] != 0) : true; }
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
    if (first_index + i == 0) {  // global sample 0: the host applies the unconditional i==0 assignment (pmc.c:552,579)
      first_vals[0] = rloc;
      first_vals[1] = 1.0;  // proposal evaluated
      first_vals[2] = w;
      first_vals[3] = reached_w ? 1.0 : 0.0;
    }
  }
  weight_block_epilogue(my_w, my_r, okflag, blk_maxw, blk_maxr, counters);
}

// log-pdf of a mixture (proposal or mix target) on n points
__global__ void k_mix_log_pdf(MixDev m, long n, const double *__restrict__ X, double *__restrict__ out,
                              double *__restrict__ ldscratch, int ldstride) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[MAXD], y[MAXD];
  for (int a = 0; a < m.d; ++a) x[a] = X[(size_t)i * m.d + a];
  out[i] = mix_log_pdf_generic(m, x, y, ldscratch + (size_t)i * ldstride);
}
// scalar_product (mvdens.c:319-349) of component k of a mixture on n points: q = |L^-1 (x - mu)|^2
__global__ void k_scalar_product(MixDev m, int k, long n, const double *__restrict__ X, double *__restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[MAXD], y[MAXD];
  for (int a = 0; a < m.d; ++a) x[a] = X[(size_t)i * m.d + a];
  out[i] = whiten_q(m.L + (size_t)k * m.d * m.d, m.mean + (size_t)k * m.d, x, y, m.d);
}
__global__ void k_target_log_pdf(TargetDev tg, int dfree, long n, const double *__restrict__ X, double *__restrict__ out,
                                 double *__restrict__ ldscratch, int ldstride) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[MAXD], xf[MAXD], y[MAXD];
  for (int a = 0; a < dfree; ++a) x[a] = X[(size_t)i * dfree + a];
  out[i] = target_log_pdf_generic(tg, x, xf, y, ldscratch + (size_t)i * ldstride);
}

// combine_lkl (distribution.c:350-396) on n points: the terms' log-densities, each on its index-mapped sub-vector, added in
// term order.  Every term is one of the densities this library evaluates itself (bentGauss, mvdens, mix_mvdens).
__global__ void k_combine_target(const CombTermDev *__restrict__ terms, int nterms, int d, long n, const double *__restrict__ X,
                                 double *__restrict__ out, double *__restrict__ ldscratch, int ldstride) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[MAXD], xs[MAXD], y[MAXD];
  for (int a = 0; a < d; ++a) x[a] = X[(size_t)i * d + a];
  double res = 0.0;
  for (int t = 0; t < nterms; ++t) {
    const CombTermDev tm = terms[t];
    for (int a = 0; a < tm.nd; ++a) xs[a] = x[tm.dim[a]];
    double v;
    if (tm.kind == PMCGPU_TGT_BANANA) v = bent_gauss(xs, tm.nd, tm.b, tm.sig1);
    else if (tm.kind == PMCGPU_TGT_MVDENS) v = comp_log_pdf(tm.mix, 0, xs, y, nullptr);
    else v = mix_log_pdf_generic(tm.mix, xs, y, ldscratch + (size_t)i * ldstride);
    res += v;
  }
  out[i] = res;
}

__global__ void k_ranindx(const double *__restrict__ cw, int K, long n, const double *__restrict__ z,
                          unsigned long long *__restrict__ idx) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (unsigned long long)ranindx(cw, K, z[i]);
}
__global__ void k_isinbox(BoxDev bx, long n, const double *__restrict__ X, int *__restrict__ inside) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[MAXD];
  bool fin = true;
  for (int a = 0; a < bx.d; ++a) { x[a] = X[(size_t)i * bx.d + a]; fin = fin && isfinite(x[a]); }
  inside[i] = !fin ? -1 : (in_box(bx, x) ? 1 : 0);
}

// ---- deterministic block-partial sums -------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double *out /* [gridDim.x*NV] */) {
  __shared__ double sh[NV][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double s = warp_sum(v[j]);
    if (lane == 0) sh[j][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double s = warp_sum(lane < nw ? sh[j][lane] : 0.0);
      if (lane == 0) out[(size_t)blockIdx.x * NV + j] = s;
    }
  }
}

// normalize_importance_weight pass 1 (pmc.c:659-674): w_i <- exp(lw_i - maxW + 500), non-finite flagged out, sum
__global__ void k_norm_exp_sum(long N, double maxW, double *__restrict__ w, short *__restrict__ flg,
                               double *__restrict__ partial, unsigned long long *counters) {
  double v[1] = {0.0};
  unsigned int cnt = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int U = 4;  // four independent loads in flight per thread (HBM-bound pass: memory-level parallelism)
  for (long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x; i0 < N; i0 += U * stride) {
    double lw[U];
    short f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * stride;
      f[u] = (i < N) ? flg[i] : (short)0;
      lw[u] = (i < N) ? w[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * stride;
      if (i < N && f[u] == 1) {
        const double e = exp(lw[u] - maxW + kDynMax);
        w[i] = e;
        if (!isfinite(e)) { flg[i] = 0; continue; }
        v[0] += e;
        ++cnt;
      }
    }
  }
  block_sum_store<1>(v, partial);
  const unsigned int c = (unsigned int)warp_sum_ll(cnt);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[CNT_NORMCOUNT], (unsigned long long)c);
}
// pass 2 (pmc.c:684-686): every entry, flagged or not, is divided by the sum
__global__ void k_norm_scale(long N, double S, double *__restrict__ w) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) w[i] /= S;
}
// perplexity_and_ess (pmc.c:1041-1077): partial[2b] = sum w log w, partial[2b+1] = sum w^2 over flg!=0, w >= 1e-20
__global__ void k_perplexity(long N, const double *__restrict__ w, const short *__restrict__ flg,
                             double *__restrict__ partial, unsigned long long *counters) {
  double v[2] = {0.0, 0.0};
  unsigned int nexit = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
    if (flg[i] == 0) { ++nexit; continue; }
    const double wi = w[i];
    if (wi < kPerpZero) continue;
    v[0] += wi * log(wi);
    v[1] += wi * wi;
  }
  block_sum_store<2>(v, partial);
  const unsigned int c = (unsigned int)warp_sum_ll(nexit);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[CNT_FLAGGED], (unsigned long long)c);
}
// prepare_update's histogram (pmc.c:1559-1564)
__global__ void k_count_indices(long N, int K, const unsigned long long *__restrict__ indices,
                                const short *__restrict__ flg, unsigned long long *__restrict__ count) {
  extern __shared__ unsigned int sc[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) sc[k] = 0;
  __syncthreads();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
    if (flg[i] == 1 && indices[i] < (unsigned long long)K) atomicAdd(&sc[indices[i]], 1u);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (sc[k]) atomicAdd(&count[k], (unsigned long long)sc[k]);
}

// ---- two-pass RB update, literal (pmc.c:1376-1453) ----------------------------------------------------------
// pass A, per sample: rho_d, w8, gw8 for every component with count > MINCOUNT; gw8 and w8 go to scratch [n*K]
// (the reference keeps gamma_rho_d[N*K] the same way, pmc.c:1358).  hard = 1 gives update_prop's indicator
// responsibilities (pmc.c:1246-1268).
__global__ void k_rb_resp(MixDev m, long n, long base, const double *__restrict__ X, const double *__restrict__ w,
                          const double *__restrict__ log_rho, const short *__restrict__ flg,
                          const unsigned long long *__restrict__ indices, const unsigned long long *__restrict__ count,
                          double maxR, int hard, double *__restrict__ gw8, double *__restrict__ w8) {
  const long il = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (il >= n) return;
  const long i = base + il;
  const int d = m.d, K = m.K;
  double *g = gw8 + (size_t)il * K, *ww = w8 + (size_t)il * K;
  for (int k = 0; k < K; ++k) { g[k] = 0.0; ww[k] = 0.0; }
  if (flg[i] == 0) return;
  double x[MAXD], y[MAXD];
  for (int a = 0; a < d; ++a) x[a] = X[(size_t)i * d + a];
  if (hard) {
    const unsigned long long k = indices[i];
    if (k < (unsigned long long)K && count[k] > kMinCount) {
      const double wi = w[i];
      const double gw = wi;
      // NB the reference evaluates the Student-t gamma against the already zeroed component (pmc.c:1241-1258),
      // which yields NaN; the hard update is only supported for Gaussian components here.
      ww[k] = wi;
      g[k] = gw;
    }
    return;
  }
  const double rho_d0 = log_rho[i] - maxR + kDynMax;
  for (int k = 0; k < K; ++k) {
    if (count[k] > kMinCount) {
      double q;
      double rho_d = exp(comp_log_pdf(m, k, x, y, &q) - rho_d0);
      rho_d *= m.alpha[k];
      const double wk = w[i] * rho_d;
      double gk = wk;
      if (m.df[k] != -1) gk = ((m.df[k] + d) / (m.df[k] + q)) * wk;
      ww[k] = wk;
      g[k] = gk;
    }
  }
}
// Reductions over samples.  Each thread walks its samples in increasing order, i.e. the reference's own summation
// order (pmc.c:1376-1405, 1432-1453); with nsplit > 1 the sample range is cut into nsplit contiguous pieces whose
// partial sums are added in piece order by k_rb_sum_splits.
// block (k, split): thread j < d -> sum gw8*x_j ; j == d -> sum gw8 ; j == d+1 -> sum w8
__global__ void k_rb_reduce_mean(int K, int d, long n, long base, int nsplit, const double *__restrict__ X,
                                 const double *__restrict__ gw8, const double *__restrict__ w8,
                                 double *__restrict__ part /* [nsplit][K*(d+2)] */) {
  const int k = blockIdx.x, sp = blockIdx.y, j = threadIdx.x;
  if (j >= d + 2) return;
  const long per = (n + nsplit - 1) / nsplit;
  const long lo = sp * per, hi = (lo + per < n) ? lo + per : n;
  double acc = 0.0;
  for (long il = lo; il < hi; ++il) {
    const double g = (j == d + 1) ? w8[(size_t)il * K + k] : gw8[(size_t)il * K + k];
    if (g == 0.0) continue;
    acc += (j < d) ? X[(size_t)(base + il) * d + j] * g : g;
  }
  part[(size_t)sp * K * (d + 2) + (size_t)k * (d + 2) + j] = acc;
}
// pass B (pmc.c:1432-1453): block (k, row tile of RB_ROWS, split); thread b accumulates rows a0..a0+RB_ROWS-1, b >= a
constexpr int RB_ROWS = 8;
__global__ void k_rb_reduce_cov(int K, int d, long n, long base, int nsplit, const double *__restrict__ X,
                                const double *__restrict__ gw8, const double *__restrict__ newmean,
                                double *__restrict__ part /* [nsplit][K*d*d] */) {
  const int k = blockIdx.x, a0 = blockIdx.y * RB_ROWS, sp = blockIdx.z, b = threadIdx.x;
  if (b >= d) return;
  const long per = (n + nsplit - 1) / nsplit;
  const long lo = sp * per, hi = (lo + per < n) ? lo + per : n;
  double acc[RB_ROWS], ma[RB_ROWS];
#pragma unroll
  for (int r = 0; r < RB_ROWS; ++r) { acc[r] = 0.0; ma[r] = (a0 + r < d) ? newmean[(size_t)k * d + a0 + r] : 0.0; }
  const double mb = newmean[(size_t)k * d + b];
  for (long il = lo; il < hi; ++il) {
    const double g = gw8[(size_t)il * K + k];
    if (g == 0.0) continue;
    const double *x = X + (size_t)(base + il) * d;
    const double xb = x[b] - mb;
#pragma unroll
    for (int r = 0; r < RB_ROWS; ++r)
      if (a0 + r < d && a0 + r <= b) acc[r] += ((x[a0 + r] - ma[r]) * xb) * g;
  }
  double *o = part + (size_t)sp * K * d * d + (size_t)k * d * d;
#pragma unroll
  for (int r = 0; r < RB_ROWS; ++r)
    if (a0 + r < d && a0 + r <= b) o[(size_t)(a0 + r) * d + b] = acc[r];
}
// out[t] += sum over splits, in split order
__global__ void k_rb_sum_splits(long len, int nsplit, const double *__restrict__ part, double *__restrict__ out) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= len) return;
  double s = out[t];
  for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * len + t];
  out[t] = s;
}
// normalize pass 2 fused with perplexity_and_ess: w_i /= S for every i (pmc.c:684-686), then the sums of
// pmc.c:1059-1069 over flg != 0
__global__ void k_norm_scale_perp(long N, double S, double *__restrict__ w, const short *__restrict__ flg,
                                  double *__restrict__ partial, unsigned long long *counters) {
  double v[2] = {0.0, 0.0};
  unsigned int nexit = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x; i0 < N; i0 += U * stride) {
    double wu[U];
    short f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * stride;
      f[u] = (i < N) ? flg[i] : (short)1;
      wu[u] = (i < N) ? w[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * stride;
      if (i >= N) continue;
      const double wi = wu[u] / S;
      w[i] = wi;
      if (f[u] == 0) { ++nexit; continue; }
      if (wi < kPerpZero) continue;
      v[0] += wi * log(wi);
      v[1] += wi * wi;
    }
  }
  block_sum_store<2>(v, partial);
  const unsigned int c = (unsigned int)warp_sum_ll(nexit);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[CNT_FLAGGED], (unsigned long long)c);
}


// ---- weight post-processing hooks (pmc.c:699-770, 1108-1120) -----------------------------------------------------------
// filter_histogram, step 1: minimum and maximum of the weights of ok samples ("w > max" / "w < min": NaN never wins)
__global__ void k_minmax_ok(long N, const double *__restrict__ w, const short *__restrict__ flg,
                            double *__restrict__ partial /* [2*gridDim.x]: max, -min */, unsigned long long *cnt) {
  double mx = -INFINITY, mn = INFINITY;
  unsigned int n = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
    if (flg[i] == 1) {
      const double v = w[i];
      if (v > mx) mx = v;
      if (v < mn) mn = v;
      ++n;
    }
  double a = warp_max_nan_ignoring(mx), b = warp_max_nan_ignoring(-mn);
  __shared__ double sa[32], sb[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sa[wid] = a; sb[wid] = b; }
  __syncthreads();
  if (wid == 0) {
    a = warp_max_nan_ignoring(lane < nw ? sa[lane] : -INFINITY);
    b = warp_max_nan_ignoring(lane < nw ? sb[lane] : -INFINITY);
    if (lane == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
  }
  const unsigned int c = (unsigned int)warp_sum_ll(n);
  if (lane == 0 && c) atomicAdd(cnt, (unsigned long long)c);
}
// step 2: ok samples outside [bl, br] are flagged out (pmc.c:756-765)
__global__ void k_flag_outside(long N, const double *__restrict__ w, short *__restrict__ flg, double bl, double br,
                               unsigned long long *cnt) {
  unsigned int n = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
    if (flg[i] == 1) {
      const double v = w[i];
      if (v < bl || br < v) { flg[i] = 0; ++n; }
    }
  const unsigned int c = (unsigned int)warp_sum_ll(n);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, (unsigned long long)c);
}
// mean_from_psim (pmc.c:1108-1120): sum over ok samples of w_i X[i][a]
__global__ void k_weighted_coord(long N, int d, int a, const double *__restrict__ X, const double *__restrict__ w,
                                 const short *__restrict__ flg, double *__restrict__ partial) {
  double v[1] = {0.0};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
    if (flg[i] != 0) v[0] += w[i] * X[(size_t)i * d + a];
  block_sum_store<1>(v, partial);
}

// ---- launch wrappers -----------------------------------------------------------------------------------------
static inline int nblk(long n, int t) { return (int)((n + t - 1) / t); }
static long g_launches = 0;
long launch_count() { return g_launches; }
void count_launch(int n) { g_launches += n; }

void launch_draw_generic(const MixDev &m, const BoxDev &bx, uint64_t seed, uint32_t iter, long first_index, long N,
                         int max_attempts, double *X, unsigned long long *indices, short *flg,
                         unsigned long long *counters, cudaStream_t s, const unsigned int *list, int attempt0) {
  if (N > 0) { k_draw_generic<<<nblk(N, 128), 128, 0, s>>>(m, bx, seed, iter, first_index, N, max_attempts, X, indices, flg, counters, list, attempt0); count_launch(); }
}
void launch_weights_generic(const MixDev &m, const TargetDev &tg, int mode, const double *post_in, const short *ok_in,
                            double t_temper, long first_index, long N, const double *X, double *lw, double *log_rho,
                            short *flg, double *ldscratch, int ldstride, double *blk_maxw, double *blk_maxr,
                            unsigned long long *counters, double *first_vals, cudaStream_t s) {
  if (N > 0)
    { k_weights_generic<<<nblk(N, GEN_WEIGHT_THREADS), GEN_WEIGHT_THREADS, 0, s>>>(m, tg, mode, post_in, ok_in, t_temper, first_index, N, X, lw, log_rho, flg,
                                                 ldscratch, ldstride, blk_maxw, blk_maxr, counters, first_vals); count_launch(); }
}
void launch_mix_log_pdf(const MixDev &m, long n, const double *X, double *out, double *ld, int ldstride, cudaStream_t s) {
  if (n > 0) { k_mix_log_pdf<<<nblk(n, 128), 128, 0, s>>>(m, n, X, out, ld, ldstride); count_launch(); }
}
void launch_combine_target(const CombTermDev *terms, int nterms, int d, long n, const double *X, double *out, double *ld,
                           int ldstride, cudaStream_t s) {
  if (n > 0) { k_combine_target<<<nblk(n, 128), 128, 0, s>>>(terms, nterms, d, n, X, out, ld, ldstride); count_launch(); }
}
void launch_scalar_product(const MixDev &m, int k, long n, const double *X, double *out, cudaStream_t s) {
  if (n > 0) { k_scalar_product<<<nblk(n, 128), 128, 0, s>>>(m, k, n, X, out); count_launch(); }
}
void launch_target_log_pdf(const TargetDev &tg, int dfree, long n, const double *X, double *out, double *ld, int ldstride, cudaStream_t s) {
  if (n > 0) { k_target_log_pdf<<<nblk(n, 128), 128, 0, s>>>(tg, dfree, n, X, out, ld, ldstride); count_launch(); }
}
void launch_ranindx(const double *cw, int K, long n, const double *z, unsigned long long *idx, cudaStream_t s) {
  if (n > 0) { k_ranindx<<<nblk(n, 256), 256, 0, s>>>(cw, K, n, z, idx); count_launch(); }
}
void launch_isinbox(const BoxDev &bx, long n, const double *X, int *inside, cudaStream_t s) {
  if (n > 0) { k_isinbox<<<nblk(n, 128), 128, 0, s>>>(bx, n, X, inside); count_launch(); }
}
void launch_norm_exp_sum(long N, double maxW, double *w, short *flg, double *partial, int nblocks,
                         unsigned long long *counters, cudaStream_t s) {
  { k_norm_exp_sum<<<nblocks, 256, 0, s>>>(N, maxW, w, flg, partial, counters); count_launch(); }
}
void launch_norm_scale(long N, double S, double *w, int nblocks, cudaStream_t s) {
  { k_norm_scale<<<nblocks, 256, 0, s>>>(N, S, w); count_launch(); }
}
void launch_perplexity(long N, const double *w, const short *flg, double *partial, int nblocks,
                       unsigned long long *counters, cudaStream_t s) {
  { k_perplexity<<<nblocks, 256, 0, s>>>(N, w, flg, partial, counters); count_launch(); }
}
void launch_count_indices(long N, int K, const unsigned long long *indices, const short *flg,
                          unsigned long long *count, int nblocks, cudaStream_t s) {
  { k_count_indices<<<nblocks, 256, K * sizeof(unsigned int), s>>>(N, K, indices, flg, count); count_launch(); }
}
void launch_rb_resp(const MixDev &m, long n, long base, const double *X, const double *w, const double *log_rho,
                    const short *flg, const unsigned long long *indices, const unsigned long long *count, double maxR,
                    int hard, double *gw8, double *w8, cudaStream_t s) {
  if (n > 0) { k_rb_resp<<<nblk(n, 128), 128, 0, s>>>(m, n, base, X, w, log_rho, flg, indices, count, maxR, hard, gw8, w8); count_launch(); }
}
void launch_rb_reduce_mean(int K, int d, long n, long base, int nsplit, const double *X, const double *gw8,
                           const double *w8, double *part, double *out, cudaStream_t s) {
  { k_rb_reduce_mean<<<dim3(K, nsplit), ((d + 2 + 31) / 32) * 32, 0, s>>>(K, d, n, base, nsplit, X, gw8, w8, part); count_launch(); }
  const long len = (long)K * (d + 2);
  { k_rb_sum_splits<<<nblk(len, 256), 256, 0, s>>>(len, nsplit, part, out); count_launch(); }
}
void launch_rb_reduce_cov(int K, int d, long n, long base, int nsplit, const double *X, const double *gw8,
                          const double *newmean, double *part, double *out, cudaStream_t s) {
  cudaMemsetAsync(part, 0, sizeof(double) * (size_t)nsplit * K * d * d, s);
  { k_rb_reduce_cov<<<dim3(K, (d + RB_ROWS - 1) / RB_ROWS, nsplit), ((d + 31) / 32) * 32, 0, s>>>(K, d, n, base, nsplit, X, gw8, newmean, part); count_launch(); }
  const long len = (long)K * d * d;
  { k_rb_sum_splits<<<nblk(len, 256), 256, 0, s>>>(len, nsplit, part, out); count_launch(); }
}
void launch_norm_scale_perp(long N, double S, double *w, const short *flg, double *partial, int nblocks,
                            unsigned long long *counters, cudaStream_t s) {
  { k_norm_scale_perp<<<nblocks, 256, 0, s>>>(N, S, w, flg, partial, counters); count_launch(); }
}

void launch_minmax_ok(long N, const double *w, const short *flg, double *partial, int nblocks, unsigned long long *cnt,
                      cudaStream_t s) {
  { k_minmax_ok<<<nblocks, 256, 0, s>>>(N, w, flg, partial, cnt); count_launch(); }
}
void launch_flag_outside(long N, const double *w, short *flg, double bl, double br, int nblocks, unsigned long long *cnt,
                         cudaStream_t s) {
  { k_flag_outside<<<nblocks, 256, 0, s>>>(N, w, flg, bl, br, cnt); count_launch(); }
}
void launch_weighted_coord(long N, int d, int a, const double *X, const double *w, const short *flg, double *partial,
                           int nblocks, cudaStream_t s) {
  { k_weighted_coord<<<nblocks, 256, 0, s>>>(N, d, a, X, w, flg, partial); count_launch(); }
}

}  // namespace pmcb200
