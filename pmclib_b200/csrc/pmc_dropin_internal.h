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
}  // namespace pmcb200
