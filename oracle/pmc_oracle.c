/* pmc_oracle.c — CPU restatement of pmclib's PMC iteration.  TEST INFRASTRUCTURE ONLY (see pmc_oracle.h).
 *
 * Each function follows the cited reference lines operation by operation (same order of floating-point
 * operations, same constants, same branch conditions) so that it reproduces the reference to rounding.
 * GSL pieces (Cholesky, dtrsv, dtrmv) restate the published gsl-1.14 / reference-CBLAS algorithms, because
 * GSL is an un-vendored dependency of the reference (wscript:131-133).
 */
#include "pmc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- GSL pieces -------------------------------------------------------------------------------- */
static double nrm2(const double *x, int n) { /* reference BLAS dnrm2 (scaled sum of squares) */
  double scale = 0.0, ssq = 1.0;
  if (n <= 0) return 0.0;
  if (n == 1) return fabs(x[0]);
  for (int i = 0; i < n; ++i) {
    if (x[i] != 0.0) {
      const double ax = fabs(x[i]);
      if (scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else ssq += (ax / scale) * (ax / scale);
    }
  }
  return scale * sqrt(ssq);
}

/* mvdens.c:229-277: factor in place, on failure restore the input and report mv_cholesky; detL = prod L_ii
 * (determinant(), mvdens.c:204-213).  The factorisation is gsl-1.14's: row by row, diagonal through dnrm2,
 * L^T mirrored into the upper triangle, failure when any pivot <= 0. */
int orc_cholesky(int d, double *a, double *detL) {
  double *save = (double *)malloc(sizeof(double) * d * d);
  int bad = 0;
  memcpy(save, a, sizeof(double) * d * d);
  {
    const double a00 = a[0];
    a[0] = a00 >= 0 ? sqrt(a00) : NAN;
    if (a00 <= 0) bad = 1;
  }
  if (d > 1) {
    const double l10 = a[d] / a[0];
    const double diag = a[d + 1] - l10 * l10;
    a[d] = l10;
    a[d + 1] = diag >= 0 ? sqrt(diag) : NAN;
    if (diag <= 0) bad = 1;
  }
  for (int k = 2; k < d; ++k) {
    const double akk = a[k * d + k];
    for (int i = 0; i < k; ++i) {
      double sum = 0.0;
      for (int j = 0; j < i; ++j) sum += a[i * d + j] * a[k * d + j];
      a[k * d + i] = (a[k * d + i] - sum) / a[i * d + i];
    }
    {
      const double nr = nrm2(a + k * d, k);
      const double diag = akk - nr * nr;
      a[k * d + k] = diag >= 0 ? sqrt(diag) : NAN;
      if (diag <= 0) bad = 1;
    }
  }
  for (int i = 1; i < d; ++i)
    for (int j = 0; j < i; ++j) a[j * d + i] = a[i * d + j];
  if (bad) {
    memcpy(a, save, sizeof(double) * d * d);
    free(save);
    return ORC_MV_CHOLESKY;
  }
  free(save);
  if (detL) {
    double det = 1.0;
    for (int i = 0; i < d * d; i += d + 1) det *= a[i];
    *detL = det;
  }
  return ORC_OK;
}

/* mvdens.c:319-349: tmp = x - mean; forward substitution (cblas dtrsv row-major/lower/no-trans/non-unit:
 * x0 /= L00; xi = (xi - sum_{j<i} L_ij x_j) / L_ii with the subtraction applied term by term); sum of squares */
double orc_scalar_product(int d, const double *mean, const double *L, const double *x, double *tmp) {
  double q = 0.0;
  for (int i = 0; i < d; ++i) tmp[i] = x[i];
  for (int i = 0; i < d; ++i) tmp[i] -= mean[i];
  tmp[0] = tmp[0] / L[0];
  for (int i = 1; i < d; ++i) {
    double t = tmp[i];
    for (int j = 0; j < i; ++j) t -= L[i * d + j] * tmp[j];
    tmp[i] = t / L[i * d + i];
  }
  for (int i = 0; i < d; ++i) q += tmp[i] * tmp[i];
  return q;
}

/* mvdens.c:354-385 */
double orc_mvdens_log_pdf(int d, const double *mean, const double *L, double detL, int df, const double *x, double *tmp) {
  const double logdet = log(detL);
  const double log_g = orc_scalar_product(d, mean, L, x, tmp);
  if (df != -1) {
    const size_t n = df, p = d;
    const double gamma = lgamma(0.5 * (n + p));
    const double norm = lgamma(0.5 * n) + 0.5 * p * log(n * ORC_PI);
    const double r = 0.5 * (n + p) * log(1.0 + (1.0 / n) * log_g);
    return gamma - (norm + logdet + r);
  }
  return -0.5 * (d * ORC_LN2PI + log_g) - logdet;
}

/* mvdens.c:784-825: components with wght<=0 are skipped; max-shifted sum with the DYNMAX offset */
double orc_mix_log_pdf(const orc_mix *m, const double *x, double *tmp, double *ld, int *rc) {
  const int K = m->K, d = m->d;
  int iloc = 0;
  double mx, lpos;
  while (iloc < K && m->wght[iloc] <= 0) iloc++;
  if (iloc >= K) { if (rc) *rc = ORC_MV_OUTOFBOUND; return 0.0; }
  mx = orc_mvdens_log_pdf(d, m->mean + (size_t)iloc * d, m->L + (size_t)iloc * d * d, m->detL[iloc], m->df[iloc], x, tmp);
  ld[iloc] = mx;
  for (int i = iloc + 1; i < K; ++i) {
    if (m->wght[i] > 0.0) {
      const double v = orc_mvdens_log_pdf(d, m->mean + (size_t)i * d, m->L + (size_t)i * d * d, m->detL[i], m->df[i], x, tmp);
      ld[i] = v;
      if (mx < v) mx = v;
    }
  }
  lpos = 0.0;
  for (int i = 0; i < K; ++i) {
    if (m->wght[i] > 0.0) {
      ld[i] = ld[i] - mx + ORC_DYNMAX;
      lpos = lpos + exp(ld[i]) * m->wght[i];
    }
  }
  return mx + log(lpos) - ORC_DYNMAX;
}

/* mvdens.c:736-751 */
int orc_cwght(int K, double *wght, double *cw) {
  double s = 0.0;
  for (int i = 0; i < K; ++i) s += wght[i];
  if (fabs(s - 1.0) > 1e-6) {
    if (s == 0) return ORC_MV_NEGWEIGHT;
    for (int i = 0; i < K; ++i) wght[i] *= 1.0 / s;
  }
  cw[0] = wght[0];
  for (int i = 1; i < K; ++i) cw[i] = cw[i - 1] + wght[i];
  return ORC_OK;
}

/* mvdens.c:766-782 */
size_t orc_ranindx(const double *cw, size_t K, double z) {
  size_t index = 0;
  while (cw[index] < z && index < K - 1) index++;
  return index;
}

/* mvdens.c:279-317: res = L g by cblas dtrmv (row-major/lower/no-trans/non-unit: i descending,
 * x_i = sum_{j<i} x_j L_ij + x_i L_ii), Student-t correction sqrt(df/u), then res*corr + mean */
void orc_mvdens_ran_from(int d, const double *mean, const double *L, int df, const double *g, double u, double *x) {
  double corr = 1;
  for (int i = 0; i < d; ++i) x[i] = g[i];
  for (int i = d; i > 0 && i--;) {
    double t = 0.0;
    for (int j = 0; j < i; ++j) t += x[j] * L[i * d + j];
    x[i] = t + x[i] * L[i * d + i];
  }
  if (df != -1) corr = sqrt(df / u);
  for (int i = 0; i < d; ++i) { x[i] *= corr; x[i] += mean[i]; }
}

/* parabox.c:112-136 (the finiteness test raises pb_infnan in the reference; here -1) */
int orc_isinbox(int d, int nfaces, const double *faces, const double *x) {
  for (int j = 0; j < d; ++j) if (!isfinite(x[j])) return -1;
  for (int i = 0; i < nfaces; ++i) {
    const double *c = faces + (size_t)i * (d + 1);
    double thr = c[d];
    for (int j = 0; j < d; ++j) thr -= c[j] * x[j];
    if (thr < 0) return 0;
  }
  return 1;
}

/* examples/target/sampleTarget.c:89-110 */
double orc_bentgauss(int d, double b, double sig1, const double *x) {
  double res = 0, x2 = x[0] * x[0];
  res += x2 / sig1;
  x2 -= sig1;
  x2 *= b;
  x2 += x[1];
  res += x2 * x2;
  for (int i = 2; i < d; ++i) res += x[i] * x[i];
  return -res / 2.;
}

double orc_target_log_pdf(const orc_target *t, const double *x, double *tmp, double *ld, int *rc) {
  if (t->kind == ORC_TGT_COMBINE) { /* distribution.c:350-396: res += distribution_lkl(term, gathered parameters) */
    double res = 0.0;
    for (int id = 0; id < t->nterms; ++id) {
      const orc_target *a = &t->terms[id];
      double *xs = (double *)malloc(sizeof(double) * a->d), *tm = (double *)malloc(sizeof(double) * a->d);
      double *l2 = (double *)malloc(sizeof(double) * (a->kind == ORC_TGT_BANANA ? 1 : (a->mix.K > 0 ? a->mix.K : 1)));
      for (int ip = 0; ip < a->d; ++ip) xs[ip] = x[t->dims[id][ip]];
      res += orc_target_log_pdf(a, xs, tm, l2, rc);
      free(xs); free(tm); free(l2);
    }
    return res;
  }
  if (t->kind == ORC_TGT_BANANA) return orc_bentgauss(t->d, t->b, t->sig1, x);
  if (t->kind == ORC_TGT_MVDENS)
    return orc_mvdens_log_pdf(t->d, t->mix.mean, t->mix.L, t->mix.detL[0], t->mix.df[0], x, tmp);
  return orc_mix_log_pdf(&t->mix, x, tmp, ld, rc);
}

/* pmc.c:484-600.  Per sample: flg=0, w=0; rloc = proposal log-pdf; a non-finite rloc raises pmc_infinite which
 * the next ParameterErrorVerb swallows -> sample stays flagged out but maxR/log_rho were already updated
 * (pmc.c:549-554 run before the `continue` at :563); then post, log-weight t*post - rloc, running maxima with the
 * unconditional i==0 assignment, flg=1. */
long orc_weights(long N, int d, const double *X, const orc_mix *prop, const orc_target *tgt, double t,
                 double *weights, double *log_rho, short *flg, double *maxW, double *maxR) {
  double MR = -1.0e30, MW = -1.0e30;
  long cnt = 0;
  const int tk = (tgt->kind == ORC_TGT_BANANA || tgt->kind == ORC_TGT_COMBINE) ? 0 : tgt->mix.K;
  const int kmax = prop->K > tk ? prop->K : tk;
  double *tmp = (double *)malloc(sizeof(double) * d);
  double *ld = (double *)malloc(sizeof(double) * (kmax > 0 ? kmax : 1));
  if (t < 0) { free(tmp); free(ld); return ORC_PMC_INFINITE; } /* pmc.c:508 */
  for (long i = 0; i < N; ++i) {
    const double *x = X + (size_t)i * d;
    int rc = 0;
    double rloc, post;
    flg[i] = 0;
    weights[i] = 0;
    rloc = orc_mix_log_pdf(prop, x, tmp, ld, &rc);
    if (rc) continue; /* error inside the proposal density: swallowed, sample flagged out (pmc.c:546) */
    if ((i == 0) || rloc > MR) MR = rloc;
    log_rho[i] = rloc;
    post = orc_target_log_pdf(tgt, x, tmp, ld, &rc);
    if (rc || isnan(rloc) || isinf(rloc)) continue; /* pmc.c:548-549 error caught at :563 */
    rloc = t * post - rloc;
    if ((i == 0) || rloc > MW) MW = rloc;
    weights[i] = rloc;
    flg[i] = 1;
    cnt += 1;
  }
  *maxW = MW;
  *maxR = MR;
  free(tmp);
  free(ld);
  return cnt;
}

/* pmc.c:641-697 */
int orc_normalize(long N, double *weights, short *flg, double maxW, double *logSum) {
  double sum = 0.0;
  long count = 0;
  for (long i = 0; i < N; ++i) {
    if (flg[i] == 1) {
      const double w = weights[i];
      weights[i] = exp(w - maxW + ORC_DYNMAX);
      if (!isfinite(weights[i])) { flg[i] = 0; continue; }
      sum += weights[i];
      count++;
    }
  }
  if (count == 0) return ORC_PMC_NOSAMPLEP;
  if (sum <= 0) return ORC_PMC_NEGWEIGHT;
  for (long i = 0; i < N; ++i) weights[i] /= sum;
  *logSum = log(sum) + maxW - ORC_DYNMAX;
  return ORC_OK;
}

/* pmc.c:1041-1077 */
int orc_perplexity_ess(long N, const double *weights, const short *flg, double *perp, double *ess) {
  double p = 0, e = 0;
  long nexit = 0;
  for (long i = 0; i < N; ++i) {
    double wi;
    if (flg[i] == 0) { nexit++; continue; }
    wi = weights[i];
    if (wi < ORC_PERP_ZERO) continue;
    p += wi * log(wi);
    e += wi * wi;
  }
  if (nexit == N) return ORC_PMC_UNDEF;
  *ess = 1 / e;
  *perp = exp(-p) / (N - nexit);
  return ORC_OK;
}

/* pmc.c:1088-1105 */
int orc_evidence(long N, const short *flg, double logSum, double *ln_evi) {
  long n = 0;
  for (long i = 0; i < N; ++i) if (flg[i] != 0) n++;
  if (n == 0) return ORC_PMC_UNDEF;
  *ln_evi = logSum - log((double)n);
  return ORC_OK;
}

/* pmc.c:1471-1533 */
int orc_cleanup_after_update(orc_mix *m, const size_t *count, double minwgt, int *retry, const double *snapshot) {
  const int K = m->K, d = m->d;
  double s = 0.0;
  int nok = 0;
  for (int i = 0; i < K; ++i) if (count[i] > ORC_MINCOUNT) s += m->wght[i];
  for (int i = 0; i < K; ++i) m->wght[i] *= 1.0 / s;
  for (int i = 0; i < K; ++i) {
    if (m->wght[i] > minwgt) {
      double *a = m->L + (size_t)i * d * d;
      if (orc_cholesky(d, a, &m->detL[i]) != ORC_OK) {
        if (*retry == 0) { memcpy(a, snapshot + (size_t)i * d * d, sizeof(double) * d * d); nok = 1; }
        else m->wght[i] = 0;
      }
    }
  }
  *retry = nok;
  for (int i = 0; i < K; ++i) {
    if (count[i] < ORC_MINCOUNT || m->wght[i] < minwgt) {
      m->wght[i] = 0.0;
      memset(m->mean + (size_t)i * d, 0, sizeof(double) * d);
      memset(m->L + (size_t)i * d * d, 0, sizeof(double) * d * d);
    }
  }
  s = 0.0;
  for (int i = 0; i < K; ++i) s += m->wght[i];
  if (s <= 0.0) return ORC_PMC_NEGWEIGHT;
  for (int i = 0; i < K; ++i) m->wght[i] *= 1.0 / s;
  return ORC_OK;
}

static size_t band_of(const orc_mix *m, int k) { return m->band_limit ? m->band_limit[k] : (size_t)m->d; }

/* pmc.c:1323-1469 with prepare_update (:1535-1579: histogram of indices over flg==1; snapshot of the factors when
 * retry==0) and cleanup_after_update.  Weights must already be normalised (isLog==0 state). */
int orc_update_rb(long N, const double *X, const size_t *indices, const double *weights, const double *log_rho,
                  const short *flg, double maxR, orc_mix *m, int *retry) {
  const int K = m->K, d = m->d;
  const double minwgt = 1.0 / N;
  size_t *count = (size_t *)calloc(K, sizeof(size_t));
  double *grd = (double *)calloc((size_t)K * N, sizeof(double));
  double *wgd = (double *)calloc(K, sizeof(double));
  double *wd = (double *)calloc(K, sizeof(double));
  double *md = (double *)calloc((size_t)K * d, sizeof(double));
  double *snap = (double *)malloc(sizeof(double) * K * d * d);
  double *tmp = (double *)malloc(sizeof(double) * d);
  int rc;
  for (long i = 0; i < N; ++i) if (flg[i] == 1) count[indices[i]]++;
  memcpy(snap, m->L, sizeof(double) * K * d * d);

  for (long i = 0; i < N; ++i) {
    const double *x;
    double rho_d0;
    if (flg[i] == 0) continue;
    x = X + (size_t)i * d;
    rho_d0 = log_rho[i] - maxR + ORC_DYNMAX;
    for (int k = 0; k < K; ++k) {
      if (count[k] > ORC_MINCOUNT) {
        const double *mu = m->mean + (size_t)k * d, *L = m->L + (size_t)k * d * d;
        double rho_d = exp(orc_mvdens_log_pdf(d, mu, L, m->detL[k], m->df[k], x, tmp) - rho_d0);
        double w8, gw8;
        rho_d *= m->wght[k];
        w8 = weights[i] * rho_d;
        if (m->df[k] != -1) {
          const double gamma_d = (m->df[k] + d) / (m->df[k] + orc_scalar_product(d, mu, L, x, tmp));
          gw8 = gamma_d * w8;
        } else gw8 = w8;
        grd[(size_t)i * K + k] = gw8;
        wd[k] += w8;
        wgd[k] += gw8;
        for (int j = 0; j < d; ++j) md[(size_t)k * d + j] += x[j] * gw8;
      }
    }
  }
  for (int k = 0; k < K; ++k) {
    m->wght[k] = wd[k];
    memset(m->L + (size_t)k * d * d, 0, sizeof(double) * d * d);
    if (count[k] > ORC_MINCOUNT) for (int j = 0; j < d; ++j) m->mean[(size_t)k * d + j] = md[(size_t)k * d + j] / wgd[k];
    else for (int j = 0; j < d; ++j) m->mean[(size_t)k * d + j] = 0;
  }
  for (long i = 0; i < N; ++i) {
    const double *x;
    if (flg[i] == 0) continue;
    x = X + (size_t)i * d;
    for (int k = 0; k < K; ++k) {
      if (count[k] > ORC_MINCOUNT) {
        const double gw8 = grd[k + (size_t)i * K];
        double *S = m->L + (size_t)k * d * d;
        const double *mu = m->mean + (size_t)k * d;
        const size_t bl = band_of(m, k);
        for (int a = 0; a < d; ++a) {
          const double vloc = x[a] - mu[a];
          S[a * d + a] += vloc * vloc * gw8;
          for (int b = a + 1; b < d; ++b) {
            double rloc;
            if ((size_t)abs(a - b) >= bl) break;
            rloc = vloc * (x[b] - mu[b]) * gw8;
            S[a * d + b] += rloc;
            S[b * d + a] += rloc;
          }
        }
      }
    }
  }
  for (int k = 0; k < K; ++k)
    if (count[k] > ORC_MINCOUNT) {
      const double sc = 1.0 / m->wght[k];
      for (int j = 0; j < d * d; ++j) m->L[(size_t)k * d * d + j] *= sc;
    }
  rc = orc_cleanup_after_update(m, count, minwgt, retry, snap);
  free(count); free(grd); free(wgd); free(wd); free(md); free(snap); free(tmp);
  return rc;
}

/* pmc.c:1194-1310 */
int orc_update(long N, const double *X, const size_t *indices, const double *weights, const short *flg,
               orc_mix *m, int *retry) {
  const int K = m->K, d = m->d;
  const double minwgt = 1.0 / N;
  size_t *count = (size_t *)calloc(K, sizeof(size_t));
  double *grd = (double *)calloc(N, sizeof(double));
  double *wgd = (double *)calloc(K, sizeof(double));
  double *snap = (double *)malloc(sizeof(double) * K * d * d);
  double *oldmean = (double *)malloc(sizeof(double) * K * d);
  double *tmp = (double *)malloc(sizeof(double) * d);
  int rc;
  for (long i = 0; i < N; ++i) if (flg[i] == 1) count[indices[i]]++;
  memcpy(snap, m->L, sizeof(double) * K * d * d);
  memcpy(oldmean, m->mean, sizeof(double) * K * d);
  /* the reference zeroes wght, means and covariances first and evaluates the Student-t gamma with the zeroed
   * component (pmc.c:1241-1258) */
  for (int k = 0; k < K; ++k) m->wght[k] = 0;
  memset(m->mean, 0, sizeof(double) * K * d);
  memset(m->L, 0, sizeof(double) * K * d * d);
  for (long i = 0; i < N; ++i) {
    size_t k;
    if (flg[i] == 0) continue;
    k = indices[i];
    if (count[k] > ORC_MINCOUNT) {
      const double *x = X + (size_t)i * d;
      const double w8 = weights[i];
      double gw8;
      if (m->df[k] != -1) {
        const double gamma_d = (m->df[k] + d) / (m->df[k] + orc_scalar_product(d, m->mean + k * d, m->L + k * d * d, x, tmp));
        gw8 = gamma_d * w8;
      } else gw8 = w8;
      grd[i] = gw8;
      m->wght[k] += w8;
      wgd[k] += gw8;
      for (int j = 0; j < d; ++j) m->mean[k * d + j] += x[j] * gw8;
    }
  }
  for (int k = 0; k < K; ++k)
    if (count[k] > ORC_MINCOUNT) for (int j = 0; j < d; ++j) m->mean[(size_t)k * d + j] *= 1.0 / wgd[k];
  for (long i = 0; i < N; ++i) {
    size_t k;
    if (flg[i] == 0) continue;
    k = indices[i];
    if (count[k] > ORC_MINCOUNT) {
      const double *x = X + (size_t)i * d, *mu = m->mean + k * d;
      double *S = m->L + k * d * d;
      const double gw8 = grd[i];
      const size_t bl = band_of(m, (int)k);
      for (int a = 0; a < d; ++a) {
        const double vloc = x[a] - mu[a];
        S[a * d + a] += vloc * vloc * gw8;
        for (int b = a + 1; b < d; ++b) {
          double rloc;
          if ((size_t)abs(a - b) >= bl) break;
          rloc = vloc * (x[b] - mu[b]) * gw8;
          S[a * d + b] += rloc;
          S[b * d + a] += rloc;
        }
      }
    }
  }
  for (int k = 0; k < K; ++k)
    if (count[k] > ORC_MINCOUNT) {
      const double sc = 1.0 / m->wght[k];
      for (int j = 0; j < d * d; ++j) m->L[(size_t)k * d * d + j] *= sc;
    }
  rc = orc_cleanup_after_update(m, count, minwgt, retry, snap);
  free(count); free(grd); free(wgd); free(snap); free(oldmean); free(tmp);
  return rc;
}

/* ---- draw (pmc.c:266-298) with a private stream ------------------------------------------------- */
static unsigned long long sm64(unsigned long long *s) {
  unsigned long long z = (*s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static double u01(unsigned long long *s) { return (sm64(s) >> 11) * (1.0 / 9007199254740992.0); }
static double u01pos(unsigned long long *s) { double x; do { x = u01(s); } while (x == 0); return x; }
static double gauss(unsigned long long *s) { /* polar method, as gsl_ran_gaussian */
  double x, y, r2;
  do { x = -1 + 2 * u01pos(s); y = -1 + 2 * u01pos(s); r2 = x * x + y * y; } while (r2 > 1.0 || r2 == 0);
  return y * sqrt(-2.0 * log(r2) / r2);
}
static double gamma_mt(unsigned long long *s, double a) { /* Marsaglia-Tsang, a >= 1 */
  const double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d);
  for (;;) {
    double x, v, u;
    do { x = gauss(s); v = 1.0 + c * x; } while (v <= 0);
    v = v * v * v;
    u = u01pos(s);
    if (u < 1 - 0.0331 * x * x * x * x) return d * v;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return d * v;
  }
}
static double chisq(unsigned long long *s, double nu) {
  const double a = nu / 2;
  if (a < 1) { const double u = u01pos(s); return 2 * gamma_mt(s, 1.0 + a) * pow(u, 1.0 / a); }
  return 2 * gamma_mt(s, a);
}

long orc_simulate(long N, orc_mix *m, int nfaces, const double *faces, unsigned long long seed,
                  double *X, size_t *indices, short *flg, double *cw) {
  const int d = m->d;
  unsigned long long s = seed;
  long i = 0, inok = 0;
  double *g = (double *)malloc(sizeof(double) * d);
  int rc = orc_cwght(m->K, m->wght, cw);
  if (rc) { free(g); return rc; }
  while (i < N) {
    const size_t k = orc_ranindx(cw, m->K, u01(&s));
    double u = 1.0;
    for (int j = 0; j < d; ++j) g[j] = gauss(&s);
    if (m->df[k] != -1) u = chisq(&s, m->df[k]);
    orc_mvdens_ran_from(d, m->mean + k * d, m->L + k * d * d, m->df[k], g, u, X + (size_t)i * d);
    if (nfaces > 0 && orc_isinbox(d, nfaces, faces, X + (size_t)i * d) != 1) {
      inok++;
      if (inok > 5 * N) { free(g); return ORC_PMC_TOOMANYSTEPS; }
      continue;
    }
    indices[i] = k;
    flg[i] = 1;
    i++;
  }
  free(g);
  return i;
}

/* filter_histogram (pmc.c:709-770): the shipped pmc_filter hook.  min and max start from weights[0] (not from the
 * first ok sample, pmc.c:734-735) as soon as any sample is ok; returns the number of samples flagged out. */
long orc_filter_histogram(long N, const double *weights, short *flg, int nbins, int cl, int cr) {
  double minW = 0, maxW = 0;
  long i, nfilt = 0;
  for (i = 0; i < N; i++)
    if (flg[i] == 1) { minW = weights[0]; maxW = weights[0]; break; }
  for (; i < N; i++)
    if (flg[i] == 1) {
      const double w = weights[i];
      if (w > maxW) maxW = w;
      if (w < minW) minW = w;
    }
  const double step = (maxW - minW) * 1. / nbins;
  const double bl = minW + step * cl, br = maxW - step * cr;
  for (i = 0; i < N; i++)
    if (flg[i] == 1) {
      const double w = weights[i];
      if (w < bl || br < w) { flg[i] = 0; nfilt++; }
    }
  return nfilt;
}

/* mean_from_psim (pmc.c:1108-1120) */
double orc_mean_from_psim(const double *X, const double *w, const short *flg, long N, int d, int a) {
  double m = 0.0;
  for (long i = 0; i < N; i++) {
    if (flg[i] == 0) continue;
    m += w[i] * X[i * d + a];
  }
  return m;
}

/* clip_weights (pmc.c:1134-1188), literally, including the initialisation loop that tests flg[i] but copies weights[j]
 * with j stuck at 0 (pmc.c:1144-1151).  Returns 0, or -6015 (pmc_infnan) when no weight is left; *thr = the threshold. */
static int orc_dcmp(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x < y) ? -1 : (x > y) ? 1 : 0;
}
int orc_clip_weights(long N, double *weights, short *flg, int n, double *thr) {
  if (n == 0) return 0;
  double *mw = (double *)calloc(n, sizeof(double));
  long i = 0, j = 0;
  while (i < n && j < N) {
    if (flg[i] == 0) { j++; continue; }
    mw[i] = weights[j];
    i++;
  }
  qsort(mw, n, sizeof(double), orc_dcmp);
  for (j = n; j < N; j++) {
    if (flg[j] == 0) continue;
    if (mw[0] < weights[j]) { mw[0] = weights[j]; qsort(mw, n, sizeof(double), orc_dcmp); }
  }
  double sum = 0.0;
  for (j = 0; j < N; j++) {
    if (flg[j] == 0) continue;
    if (weights[j] >= mw[0]) { weights[j] = 0.0; flg[j] = 0; }
    else sum += weights[j];
  }
  if (thr) *thr = mw[0];
  free(mw);
  if (sum <= 0) return -6015;
  for (j = 0; j < N; j++) {
    if (flg[j] == 0) continue;
    weights[j] /= sum;
  }
  return 0;
}
