// pmc_hostonly.h — host-only helpers shared between pmc_host.cpp and the C-ABI layer (internal).
#pragma once
#include <stddef.h>
#include <algorithm>
#include <thread>
#include <vector>
namespace pmcb200 {
// Components are independent in every O(d^3) host step (inversion, covariance assembly, factorisation); at d = 128,
// K = 64 one such step is 20-50 MFLOP of latency-bound scalar code, so it is spread over a few host threads.  Each index
// is processed by exactly one thread and results are combined by the caller in index order: deterministic.
template <typename F>
void parallel_components(int K, double work_per_component, F fn) {
  int nthr = 1;
  if (work_per_component * K > 4e6) {
    const unsigned hw = std::thread::hardware_concurrency();
    nthr = std::min<int>(K, std::min<int>(16, hw ? (int)hw : 4));
  }
  if (nthr <= 1) {
    for (int k = 0; k < K; ++k) fn(k);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nthr);
  for (int t = 0; t < nthr; ++t)
    th.emplace_back([=]() { for (int k = t; k < K; k += nthr) fn(k); });
  for (auto &x : th) x.join();
}

int host_cholesky(int d, double *a, double *detL);
int finalize_update_strided(int K, int d, const double *W, size_t sW, const double *G, size_t sG, const double *s1,
                            size_t ss1, const double *S2, size_t sS2, const double *pivot, int pivot_per_component,
                            const size_t *count, long N_total, const size_t *band_limit, int *retry, const double *oldL,
                            double *wght, double *mean, double *L, double *detL, int *alive);
int host_cleanup_after_update(int K, int d, const size_t *count, double minwgt, int *retry, const double *snapshot,
                              double *wght, double *mean, double *L, double *detL);
}  // namespace pmcb200
