// pmc_host.cpp — the O(K d^3) host side of the proposal update: from reduced sufficient statistics to the new
// mixture.  This is deliberately NOT a kernel (4.5e7 flop at the largest config) and it follows
// the reference's control flow literally, because component deaths, the retry flag and the weight floor are
// integer/branch behaviour that must match exactly:
//   mvdens_cholesky_decomp      mvdens.c:229-277  (+ gsl-1.14 gsl_linalg_cholesky_decomp semantics, un-vendored GSL)
//   cleanup_after_update        pmc.c:1471-1533
//   end of update_prop_rb       pmc.c:1409-1426, 1455-1462
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/pmcb200.h"
#include "pmc_hostonly.h"

namespace {

constexpr int kMinCount = 20;  // pmc.h:71
constexpr int kMvCholesky = -706;     // mv_cholesky, mvdens.h:60
constexpr int kPmcNegWeight = -6005;  // pmc_negWeight, pmc.h:28

// Euclidean norm the way the BLAS reference dnrm2 forms it (running scale and scaled sum of squares); GSL's
// Cholesky uses it for the diagonal, so the pivots round the same way.
double scaled_norm(const double *x, int n) {
  if (n < 1) return 0.0;
  if (n == 1) return std::fabs(x[0]);
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < n; ++i) {
    if (x[i] == 0.0) continue;
    const double ax = std::fabs(x[i]);
    if (scale < ax) {
      const double r = scale / ax;
      ssq = 1.0 + ssq * r * r;
      scale = ax;
    } else {
      const double r = ax / scale;
      ssq += r * r;
    }
  }
  return scale * std::sqrt(ssq);
}

inline double sqrt_or_nan(double v) { return v >= 0 ? std::sqrt(v) : std::nan(""); }

}  // namespace

namespace pmcb200 {

// Row-by-row factorisation; lower triangle receives L, upper triangle its mirror; any non-positive pivot -> failure
// (after finishing the sweep, as GSL does), input restored.
int host_cholesky(int d, double *a, double *detL) {
  std::vector<double> keep(a, a + (size_t)d * d);
  bool pd = true;
  for (int r = 0; r < d; ++r) {
    double *row = a + (size_t)r * d;
    const double arr = row[r];
    if (r == 1) {
      // GSL treats the second row specially (explicit product instead of the norm)
      const double l10 = row[0] / a[0];
      const double piv = arr - l10 * l10;
      row[0] = l10;
      row[1] = sqrt_or_nan(piv);
      if (piv <= 0) pd = false;
      continue;
    }
    for (int c = 0; c < r; ++c) {
      const double *rc = a + (size_t)c * d;
      double dot = 0.0;
      for (int j = 0; j < c; ++j) dot += rc[j] * row[j];
      row[c] = (row[c] - dot) / rc[c];
    }
    const double nr = (r == 0) ? 0.0 : scaled_norm(row, r);
    const double piv = (r == 0) ? arr : arr - nr * nr;
    row[r] = sqrt_or_nan(piv);
    if (piv <= 0) pd = false;
  }
  for (int r = 1; r < d; ++r)
    for (int c = 0; c < r; ++c) a[(size_t)c * d + r] = a[(size_t)r * d + c];
  if (!pd) {
    std::memcpy(a, keep.data(), sizeof(double) * d * d);
    return kMvCholesky;
  }
  if (detL) {
    double det = 1.0;
    for (int r = 0; r < d; ++r) det *= a[(size_t)r * d + r];  // determinant(), mvdens.c:204-213
    *detL = det;
  }
  return 0;
}

// cleanup_after_update (pmc.c:1471-1533).  On entry L[k] holds the updated COVARIANCE of component k and wght the
// un-normalised new weights; snapshot = the factors before the update (restored once on a failed factorisation).
int host_cleanup_after_update(int K, int d, const size_t *count, double minwgt, int *retry, const double *snapshot,
                              double *wght, double *mean, double *L, double *detL) {
  const size_t dd = (size_t)d * d;
  double total = 0.0;
  for (int k = 0; k < K; ++k)
    if (count[k] > (size_t)kMinCount) total += wght[k];
  {
    const double inv = 1.0 / total;
    for (int k = 0; k < K; ++k) wght[k] *= inv;
  }
  int need_retry = 0;
  std::vector<int> failed(K, 0);
  parallel_components(K, (double)d * d * d / 3.0, [&](int k) {
    if (!(wght[k] > minwgt)) return;
    failed[k] = host_cholesky(d, L + k * dd, &detL[k]) != 0;
  });
  for (int k = 0; k < K; ++k) {
    if (!failed[k]) continue;
    if (*retry == 0) {
      std::memcpy(L + k * dd, snapshot + k * dd, sizeof(double) * dd);
      need_retry = 1;
    } else {
      wght[k] = 0;
    }
  }
  *retry = need_retry;
  for (int k = 0; k < K; ++k) {
    if (count[k] < (size_t)kMinCount || wght[k] < minwgt) {
      wght[k] = 0.0;
      std::memset(mean + (size_t)k * d, 0, sizeof(double) * d);
      std::memset(L + k * dd, 0, sizeof(double) * dd);
    }
  }
  total = 0.0;
  for (int k = 0; k < K; ++k) total += wght[k];
  if (total <= 0.0) return kPmcNegWeight;
  {
    const double inv = 1.0 / total;
    for (int k = 0; k < K; ++k) wght[k] *= inv;
  }
  return 0;
}

}  // namespace pmcb200

namespace pmcb200 {

// The work of pmcgpu_finalize_update with explicit component strides of the statistics (so that the reduced buffer
// { W, s1[d], S2[d*d] } x K of the tensor path is read in place) and the old factors as a separate read-only array.
int finalize_update_strided(int K, int d, const double *W, size_t sW, const double *G, size_t sG, const double *s1,
                            size_t ss1, const double *S2, size_t sS2, const double *pivot, int pivot_per_component,
                            const size_t *count, long N_total, const size_t *band_limit, int *retry, const double *oldL,
                            double *wght, double *mean, double *L, double *detL, int *alive) {
  const size_t dd = (size_t)d * d;
  static const bool trace = getenv("PMCB200_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[pmcb200]   finalize: %-16s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  parallel_components(K, (double)d * d * 16.0, [&](int k) {
    double *cov = L + k * dd;
    double *mu = mean + (size_t)k * d;
    if (count[k] > (size_t)kMinCount) {
      const double *c = pivot + (pivot_per_component ? (size_t)k * d : 0);
      const double *s = s1 + (size_t)k * ss1;
      const double *S = S2 + (size_t)k * sS2;
      const double Wk = W[(size_t)k * sW], Gk = G[(size_t)k * sG];
      const size_t bl = band_limit ? band_limit[k] : (size_t)d;
      wght[k] = Wk;
      // scale-free form: the statistics may carry any common factor (the reference's exp(maxR-500) is ~1e-217), so
      // never multiply two of them: shift = s1/G is O(1), S2/W is O(1), G/W is O(1)
      std::vector<double> shift(d);
      for (int a = 0; a < d; ++a) { shift[a] = s[a] / Gk; mu[a] = c[a] + shift[a]; }
      const double invW = 1.0 / Wk;
      const double gw = Gk * invW;
      for (int a = 0; a < d; ++a)
        for (int b = a; b < d; ++b) {
          double v = 0.0;
          if ((size_t)(b - a) < bl) v = S[(size_t)a * d + b] * invW - (shift[a] * shift[b]) * gw;  // band cut, pmc.c:1444
          cov[(size_t)a * d + b] = v;
          cov[(size_t)b * d + a] = v;
        }
    } else {
      wght[k] = 0.0;
      std::memset(mu, 0, sizeof(double) * d);
      std::memset(cov, 0, sizeof(double) * dd);
    }
  });
  tick("covariances");
  const int rc = host_cleanup_after_update(K, d, count, 1.0 / (double)N_total, retry, oldL, wght, mean, L, detL);
  tick("cleanup+cholesky");
  if (alive)
    for (int k = 0; k < K; ++k) alive[k] = wght[k] > 0.0 ? 1 : 0;
  return rc;
}

}  // namespace pmcb200

extern "C" {

int pmcgpu_cholesky(int d, double *a, double *detL) { return pmcb200::host_cholesky(d, a, detL); }

long pmcgpu_stats_len(int K, int d) { return (long)K * (2 + d + (long)d * d) + 8; }

// New proposal from pivoted one-pass statistics:
//   mu'_k = c + s1_k / G_k ;  Sigma'_k = (S2_k - s1_k s1_k^T / G_k) / W_k ;  alpha'_k = W_k (normalised in cleanup)
// Components with count <= MINCOUNT get weight 0, mean 0, zero matrix exactly as pmc.c:1410-1425.
int pmcgpu_finalize_update(int K, int d, const double *W, const double *G, const double *s1, const double *S2,
                           const double *pivot, int pivot_per_component, const size_t *count, long N_total,
                           const size_t *band_limit, int *retry, double *wght, double *mean, double *L, double *detL,
                           int *alive) {
  const size_t dd = (size_t)d * d;
  // old factors, restored by cleanup_after_update on a first failed factorisation; the buffer persists across calls
  // (a fresh 8 MB vector per iteration costs more in page faults than the copy itself)
  static thread_local std::vector<double> snapshot;
  snapshot.resize(K * dd);
  double *const snap = snapshot.data();  // (the worker threads must not name the thread_local object itself)
  pmcb200::parallel_components(K, (double)d * d * 16.0, [=](int k) { std::memcpy(snap + k * dd, L + k * dd, sizeof(double) * dd); });
  return pmcb200::finalize_update_strided(K, d, W, 1, G, 1, s1, (size_t)d, S2, dd, pivot, pivot_per_component, count, N_total,
                                          band_limit, retry, snap, wght, mean, L, detL, alive);
}

}  // extern "C"
