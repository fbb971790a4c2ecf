// pmc_dropin_internal.h — what pmc_dropin.cpp shares with pmc_dropin2.cpp / pmc_dropin_mpi.cpp (not part of the ABI)
#pragma once
#include <cstdint>
#include <mutex>
#include "../../include/pmcb200.h"
#include "../../include/pmclib_abi.h"

namespace pmcb200 {
// the context behind the point-wise entry points (mvdens_log_pdf, mvdens_ran, ...) and the lock that serialises them
std::mutex &dropin_point_mutex();
pmcgpu_ctx *dropin_point_ctx(error **err);
// Philox key of one call from the caller's gsl_rng (two 32-bit words), or a process-local counter
uint64_t dropin_seed_from(gsl_rng *r);
// whoever installs another proposal / target on the point context (holding the mutex) calls this afterwards
void dropin_point_invalidate();
// proposal + target (may be null) + box onto a context; mixture <- flat arrays; which device density evaluates a target
bool dropin_push_model(pmcgpu_ctx *ctx, mix_mvdens *m, distribution *target, parabox *pb, int d, error **err);
void dropin_unflatten(mix_mvdens *m, const double *w, const double *mean, const double *L, const double *detL);
int dropin_device_kind(const distribution *t);
int dropin_device_ordinal();
// per-sample loop over an opaque target callback with the reference's error contract (pmc.c:557-577); false = abort
bool dropin_host_target_values(posterior_log_pdf_func *f, retrieve_ded_func *ret, void *target_data, const double *X, double *X_ded,
                               long N, int d, int nd, int quiet, double *post, short *ok, error **err);
}  // namespace pmcb200
