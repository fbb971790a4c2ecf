// pmc_hostonly.h — host-only helpers shared between pmc_host.cpp and the C-ABI layer (internal).
#pragma once
#include <stddef.h>
namespace pmcb200 {
int host_cholesky(int d, double *a, double *detL);
int host_cleanup_after_update(int K, int d, const size_t *count, double minwgt, int *retry, const double *snapshot,
                              double *wght, double *mean, double *L, double *detL);
}  // namespace pmcb200
