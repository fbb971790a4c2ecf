/* pmcb200.h — the C-ABI of libpmcb200.so: pmclib's Population Monte Carlo iteration on one B200.
 *
 * Plain pointers and sizes only; no torch, no C++ types.  This is what a binding (cgo / ctypes / the drop-in
 * C layer in pmclib_abi.h) calls.  Every entry point cites the reference interface it replaces; paths are
 * relative to the reference tree (CosmoStat/pmclib).
 *
 * Conventions
 *   - every function returns 0 on success or a negative code: the reference's own codes where one applies
 *     (pmc.h:23-43, mvdens.h:54-68), PMCGPU_ERR_* below for CUDA/NCCL/usage failures.  pmcgpu_last_error()
 *     gives the text.  Nothing aborts; there is no CPU fallback: without a CUDA device pmcgpu_create fails.
 *   - a mixture is passed as the reference stores it after mvdens_cholesky_decomp (mvdens.h:83-110):
 *     wght[K], mean[K*d], L[K*d*d] row-major lower Cholesky factors (upper triangle ignored), detL[K] = prod L_ii,
 *     df[K] (-1 = Gaussian, else Student-t degrees of freedom).
 *   - the sample store mirrors pmc_simu (pmc.h:76-108): X[N*d] row-major, weights[N], log_rho[N],
 *     indices[N] (size_t), flg[N] (short); it lives in HBM and is copied to/from host buffers only by
 *     pmcgpu_upload / pmcgpu_download, or streamed to an attached host sink while an iteration runs
 *     (pmcgpu_set_host_sink).
 *   - one context = one GPU = one caller thread (the reference is not re-entrant either).
 */
#ifndef PMCB200_H
#define PMCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmcgpu_ctx pmcgpu_ctx;

#define PMCGPU_ERR_CUDA (-6021)    /* pmc_base - 21: CUDA runtime failure */
#define PMCGPU_ERR_NCCL (-6022)    /* pmc_base - 22: NCCL failure / NCCL not loadable */
#define PMCGPU_ERR_USAGE (-6023)   /* pmc_base - 23: bad argument / call order */
#define PMCGPU_ERR_NODEVICE (-6024)

enum { PMCGPU_TGT_NONE = 0, PMCGPU_TGT_BANANA = 1, PMCGPU_TGT_MVDENS = 2, PMCGPU_TGT_MIX = 3, PMCGPU_TGT_HOST = 4, PMCGPU_TGT_COMBINE = 5 };

/* ---- life cycle ------------------------------------------------------------------------------- */
int pmcgpu_create(int device, pmcgpu_ctx **out);
void pmcgpu_destroy(pmcgpu_ctx *ctx);
const char *pmcgpu_last_error(const pmcgpu_ctx *ctx);
int pmcgpu_version(void);
int pmcgpu_device_count(void); /* CUDA devices visible to this process (0 without a driver) */
/* the CUDA stream all work of this context is enqueued on (a cudaStream_t), for callers that time with events */
void *pmcgpu_stream(pmcgpu_ctx *ctx);

/* tuning / test knobs: "force_generic" (1 = only the literal any-shape kernels), "fused_version" (0 = newest small-d
 * kernel that has an instantiation, 1 = first generation, 2 = second generation only), "spl" (samples per lane of the
 * first-generation kernel, 1 or 2), "split_stats" / "split_draw" (0 = single-kernel variants of the second generation),
 * "mma" (0 = do not use the large-d tensor path), "mma_min_d" (smallest dimension routed to it when no fused small-d kernel is instantiated for the shape, default 2),
 * "mma_max_chunk" (test knob: samples whitened per pass; 0 = as many as memory allows),
 * "t_temper" (t_iter_tempered used by pmcgpu_pmc_step, pmc.c:570; default 1), "max_attempts" (per-sample cap on box rejections), "force_precise" (1 = always take the two-pass update),
 * "guard_ratio" (pivot guard threshold of the one-pass update), "timing" (see pmcgpu_get_timing),
 * "gen3" (0 = without the third-generation small-d kernels), "w3_variant" (third-generation weights kernel: 3 = row-scaled
 * triangular solve, the default; 0 = exact per-component mean subtraction; 1, 2 = launch-shape experiments), "gen3_max_ratio"
 * (largest |mu_j - c_j| / L_jj the row-scaled solve is used for, default 100; beyond it the exact form runs), "draw3" (1 = the
 * block-grouped draw kernel), "sink_idx8" (0 = the host sink receives the component indices as size_t over PCIe instead of as
 * bytes widened on the host) */
int pmcgpu_set_option(pmcgpu_ctx *ctx, const char *name, double value);
/* read-only diagnostics: "mma_active", "mma_inverse_discrepancy" (conditioning guard of the tensor path: explicit inverse
 * factors are used only while they agree with the reference's forward substitution to option "mma_inv_tol", default 1e-13),
 * "w3_rowscaled" (1 if the current proposal is packed for the row-scaled solve), "w3_shift_ratio" (the guard's ratio),
 * "sink_bytes" (device-to-host bytes the host-sink copies of the last call moved) */
int pmcgpu_get_info(pmcgpu_ctx *ctx, const char *name, double *value);

/* timing of the dominant (fused) kernel, measured with CUDA events on the launching stream while option "timing" = 1:
 * accumulated milliseconds and launches since the option was set; total_launches = kernels launched by the library
 * in this process */
int pmcgpu_get_timing(pmcgpu_ctx *ctx, double *fused_ms, long *fused_launches, long *total_launches);
/* accumulated milliseconds of the statistics kernel (split pipeline) over the same launches */
int pmcgpu_get_stats_timing(pmcgpu_ctx *ctx, double *stats_ms);
/* accumulated milliseconds of the draw kernel when it runs as its own launch (option "split_draw") */
int pmcgpu_get_draw_timing(pmcgpu_ctx *ctx, double *draw_ms);

/* ---- model: proposal, target, box ------------------------------------------------------------- */
/* replaces the host-side mix_mvdens the reference passes around (mvdens.h:102-110) */
int pmcgpu_set_proposal(pmcgpu_ctx *ctx, int K, int d, const double *wght, const double *mean, const double *L,
                        const double *detL, const int *df);
/* mvdens.band_limit (mvdens.h:96, used at pmc.c:1444): band_limit[K], NULL = full matrices; reset by set_proposal */
int pmcgpu_set_band_limit(pmcgpu_ctx *ctx, const size_t *band_limit);
/* device targets; replace the posterior_log_pdf_func callback (allmc.h:21) for the densities the reference ships:
 * bentGauss_lkl (examples/target/sampleTarget.c:89-110), mvdens_log_pdf (mvdens.c:354-385),
 * mix_mvdens_log_pdf (mvdens.c:784-825) */
int pmcgpu_set_target_banana(pmcgpu_ctx *ctx, int d, double b, double sig1);
int pmcgpu_set_target_mix(pmcgpu_ctx *ctx, int kind, int K, int d, const double *wght, const double *mean,
                          const double *L, const double *detL, const int *df);
/* conditional ("default") parameters of distribution_fill_pars (distribution.c:193-209): the target sees a vector of
 * nfull = d + ndef entries, full[i] = is_def[i] ? value[i] : x[next free].  nfull = 0 clears the map. */
int pmcgpu_set_target_defaults(pmcgpu_ctx *ctx, int nfull, const int *is_def, const double *value);
/* opaque host callback: the caller evaluates the target on the host and passes post[N] to pmcgpu_weights_host_post */
int pmcgpu_set_target_host(pmcgpu_ctx *ctx, int d);
/* combined target, replaces combine_lkl (pmclib/src/distribution.c:350-396) and the Gaussian priors built on it
 * (add_gaussian_prior*, distribution.c:548-606): the target is the SUM of nterms log-densities; term t is a density of
 * kind PMCGPU_TGT_BANANA (b, sig1), _MVDENS (K = 1) or _MIX evaluated on the ndim coordinates dim_idx[0..ndim) of the
 * d-dimensional sample.  Arrays as in pmcgpu_set_target_mix.  Evaluated on the device, term by term in the given order. */
typedef struct {
  int kind, ndim;
  const int *dim_idx;
  double b, sig1;
  int K;
  const double *wght, *mean, *L, *detL;
  const int *df;
} pmcgpu_combine_term;
int pmcgpu_set_target_combine(pmcgpu_ctx *ctx, int d, int nterms, const pmcgpu_combine_term *terms);
/* parabox (parabox.h:22-26): faces[nfaces*(d+1)]; nfaces = 0 removes the box */
int pmcgpu_set_box(pmcgpu_ctx *ctx, int d, int nfaces, const double *faces);

/* ---- sample store ----------------------------------------------------------------------------- */
/* pmc_simu_init / pmc_simu_realloc (pmc.c:29-74, 222-246): N local samples of dimension d */
int pmcgpu_resize(pmcgpu_ctx *ctx, long N, int d);
/* any pointer may be NULL (that array is left untouched) */
int pmcgpu_upload(pmcgpu_ctx *ctx, const double *X, const size_t *indices, const double *weights,
                  const double *log_rho, const short *flg);
int pmcgpu_download(pmcgpu_ctx *ctx, double *X, size_t *indices, double *weights, double *log_rho, short *flg);
/* page-lock / unlock a caller-owned host block (e.g. pmc_simu.data) so pmcgpu_upload/download run as DMA */
int pmcgpu_pin_host(void *p, size_t bytes);
int pmcgpu_unpin_host(void *p);
/* device addresses of the store, for zero-copy consumers: out[5] = X, indices, weights, log_rho, flg */
int pmcgpu_device_arrays(pmcgpu_ctx *ctx, void **out);

/* ---- the phases of one iteration, on the resident store ------------------------------------------ */
/* simulate_mix_mvdens (pmc.c:266-298): draws N samples from the proposal inside the box; Philox4x32-10 keyed by
 * (seed, iter) and counted by the global sample index first_index + i, so the stream does not depend on how
 * samples are sharded.  Sets X, indices, flg=1.  n_reject may be NULL. */
int pmcgpu_draw(pmcgpu_ctx *ctx, uint64_t seed, uint32_t iter, long first_index, long *n_reject);
/* generic_get_importance_weight_and_deduced_verb (pmc.c:484-600) with the device target: fills log_rho, weights
 * (log-weights t*post - log_rho), flg; returns the ok count and the running maxima exactly as the reference
 * defines them (the i==0 quirk applies to the sample whose global index is 0: pass first_index). */
int pmcgpu_weights(pmcgpu_ctx *ctx, double t_temper, long first_index, long *nok, double *maxW, double *maxR);
/* pmc_simu_importance (pmc.c:175-194) in one kernel: draw + weights; the log-weights are left un-normalised */
int pmcgpu_draw_weights(pmcgpu_ctx *ctx, uint64_t seed, uint32_t iter, long first_index, double t_temper, long *nok,
                        double *maxW, double *maxR, long *n_reject);
/* same with the target evaluated by the caller on the host (PMCGPU_TGT_HOST): post[N]; ok[N] (may be NULL) = 0
 * where the callback raised an error (sample is flagged out, pmc.c:563) */
int pmcgpu_weights_host_post(pmcgpu_ctx *ctx, const double *post, const short *ok, double t_temper,
                             long first_index, long *nok, double *maxW, double *maxR);
/* mix_mvdens_log_pdf (mvdens.c:784-825) of the PROPOSAL / the device TARGET on n host points: out[n] */
int pmcgpu_proposal_log_pdf(pmcgpu_ctx *ctx, long n, const double *x, double *out);
int pmcgpu_target_log_pdf(pmcgpu_ctx *ctx, long n, const double *x, double *out);
/* scalar_product (pmctools/src/mvdens.c:319-349): q = (x-mu)^T Sigma^-1 (x-mu) of the mvdens installed as the target */
int pmcgpu_target_scalar_product(pmcgpu_ctx *ctx, long n, const double *x, double *out);
/* normalize_importance_weight (pmc.c:641-697): weights <- exp(lw - maxW + 500)/S over flg==1, non-finite
 * entries flagged out; logSum = log S + maxW - 500.  count = number of samples summed. */
int pmcgpu_normalize(pmcgpu_ctx *ctx, double maxW, double *logSum, long *count);
/* perplexity_and_ess (pmc.c:1041-1077, MC_UNORM) and the flagged-out count evidence() needs (pmc.c:1088-1105) */
int pmcgpu_perplexity_ess(pmcgpu_ctx *ctx, double *perp, double *ess, long *n_flagged_out);
/* Host sink for pmcgpu_pmc_step and pmcgpu_draw_weights: the caller's pmc_simu arrays (any may be NULL; all NULL detaches).  While attached, the
 * iteration copies X and indices to the host as soon as the draw has finished, log_rho after the weight kernel and the
 * normalised weights and flags after the normalisation, on a second stream, overlapped with the kernels that follow;
 * pmcgpu_pmc_step returns when the arrays are complete.  Page-locked host memory (pmcgpu_pin_host) is needed for the
 * overlap.  The error exits of the normalisation (no ok sample, sum <= 0) may leave weights/flg unwritten. */
int pmcgpu_set_host_sink(pmcgpu_ctx *ctx, double *X, size_t *indices, double *weights, double *log_rho, short *flg);
/* filter_histogram (pmc.c:709-770), the shipped pmc_filter hook of pmc_simu_pmc_step (pmc.c:203-206), on the resident
 * log-weights: ok samples whose weight lies in the clipleft lowest or clipright highest of nbins equal bins between the
 * minimum and maximum are flagged out.  Reproduces the reference's initialisation of min/max with weights[0] (whatever
 * its flag).  With ranks attached, min/max are reduced over all ranks (first_index = global index of local sample 0). */
int pmcgpu_filter_histogram(pmcgpu_ctx *ctx, int nbins, int clipleft, int clipright, long first_index, long *n_filtered);
/* the same filter applied inside pmcgpu_pmc_step / pmcgpu_step_given between the weights and the normalisation
 * (nbins <= 0 switches it off).  With the fused small-d kernels the update then runs through the literal two-pass path. */
int pmcgpu_set_filter_histogram(pmcgpu_ctx *ctx, int nbins, int clipleft, int clipright);
/* mean_from_psim (pmc.c:1108-1120): sum over flg != 0 of weights[i] * X[i][a] */
int pmcgpu_weighted_mean(pmcgpu_ctx *ctx, int a, double *mean);
/* Resampling of the resident weighted sample.  sample_from_pmc_simu (pmc.c:1585-1601): isample[N], each entry an index
 * drawn with probability proportional to weights[i] (all entries, whatever their flag, as the reference).
 * resample_residual (pmc.c:1605-1626): multiplicity[N] = multinomial(N trials) over the residual weights
 * N w_i - int(N w_i).  Philox is keyed by (seed, stream) and counted by the draw number: distributionally equivalent to
 * the reference's gsl_ran_discrete / gsl_ran_multinomial, not stream-identical. */
int pmcgpu_sample_indices(pmcgpu_ctx *ctx, uint64_t seed, uint32_t stream, size_t *isample);
int pmcgpu_resample_residual(pmcgpu_ctx *ctx, uint64_t seed, uint32_t stream, unsigned int *multiplicity);
/* clip_weights (pmc.c:1134-1188) on the resident normalised weights: the samples carrying the n largest weights are
 * zeroed and flagged out, the remaining ok weights re-normalised to one.  Reproduces the reference's initialisation of
 * its candidate array (copies of weights[0], samples j < n never offered).  Error -6015 (pmc_infnan) if nothing is left. */
int pmcgpu_clip_weights(pmcgpu_ctx *ctx, int n, double *threshold, long *n_clipped, double *sum_after);
/* prepare_update's histogram (pmc.c:1559-1564): count[K] of indices over flg==1 */
int pmcgpu_count_indices(pmcgpu_ctx *ctx, size_t *count);
/* update_prop_rb (pmc.c:1323-1469) + prepare_update + cleanup_after_update on the resident, normalised sample:
 * two passes over the samples exactly as the reference (responsibilities against the stored log_rho and maxR, then
 * covariances centred on the new means).  N_total = psim->nsamples of the whole (all-rank) sample, for the
 * 1/N weight floor.  The proposal held by the context is replaced; outputs (any may be NULL) receive it in
 * reference layout: o_wght[K], o_mean[K*d], o_L[K*d*d] (factor with the upper triangle mirrored, zeros for dead
 * components), o_detL[K], o_alive[K].  *retry is psim->retry, in/out. */
int pmcgpu_update_rb(pmcgpu_ctx *ctx, double maxR, long N_total, int *retry, double *o_wght, double *o_mean,
                     double *o_L, double *o_detL, int *o_alive);
/* update_prop (pmc.c:1194-1310), the hard-assignment variant */
int pmcgpu_update_hard(pmcgpu_ctx *ctx, long N_total, int *retry, double *o_wght, double *o_mean, double *o_L,
                       double *o_detL, int *o_alive);
/* ranindx (mvdens.c:766-782) on injected uniforms z[n] with the proposal's cumulative weights: idx[n] */
int pmcgpu_ranindx(pmcgpu_ctx *ctx, long n, const double *z, size_t *idx);
/* isinBox (parabox.c:112-136) on n host points: inside[n] */
int pmcgpu_isinbox(pmcgpu_ctx *ctx, long n, const double *x, int *inside);

/* ---- the fused iteration ---------------------------------------------------------------------- */
typedef struct {
  double perplexity;   /* perplexity(psim, MC_UNORM), what pmc_simu_pmc_step returns (pmc.c:216) */
  double ess;
  double logSum;       /* psim->logSum */
  double ln_evidence;  /* evidence() */
  double maxW, maxR;
  long nok;            /* samples with flg==1 after normalisation (all ranks) */
  long n_reject;       /* box rejections during the draw (all ranks) */
  int retry;           /* psim->retry after the update */
  int n_alive;         /* components with weight > 0 after the update */
  int used_precise_path;  /* 1 if the pivot guard sent the update through the two-pass path */
} pmcgpu_step_result;

/* pmc_simu_pmc_step (pmc.c:196-219) with t_iter_tempered = 1 (the evident intent; the reference's own call passes
 * -1 and fails): draw + weights + normalise + Rao-Blackwellised update + perplexity, one pass over
 * the samples for everything that scales with N.  do_update = 0 gives pmc_simu_importance + perplexity/evidence
 * (vanilla_importance.c:56-82).  N_total / first_index describe the shard (N_total = local N and first_index = 0
 * on one GPU); with a communicator attached the sufficient statistics are all-reduced.  The updated proposal
 * replaces the context's and is returned like pmcgpu_update_rb's. */
int pmcgpu_pmc_step(pmcgpu_ctx *ctx, uint64_t seed, uint32_t iter, long first_index, long N_total, int do_update,
                    int *retry, pmcgpu_step_result *res, double *o_wght, double *o_mean, double *o_L,
                    double *o_detL, int *o_alive);
/* same on samples already in the store (no draw): indices must be valid for the update */
int pmcgpu_step_given(pmcgpu_ctx *ctx, long first_index, long N_total, double t_temper, int do_update, int *retry,
                      pmcgpu_step_result *res, double *o_wght, double *o_mean, double *o_L, double *o_detL,
                      int *o_alive);
/* the same for an OPAQUE host target (allmc.h:21 callbacks): the samples are in the store (pmcgpu_draw), the caller has
 * evaluated post[N_local] (ok[i] = 0 where the callback raised an error, may be NULL) on the host; weights, reductions over
 * ranks, normalisation, update and perplexity then run as in pmcgpu_pmc_step.  This is what a rank of vanilla_pmc_mpi
 * does after its likelihood loop. */
int pmcgpu_step_given_post(pmcgpu_ctx *ctx, const double *post, const short *ok, long first_index, long N_total,
                           double t_temper, int do_update, int *retry, pmcgpu_step_result *res, double *o_wght,
                           double *o_mean, double *o_L, double *o_detL, int *o_alive);
/* read back the proposal currently installed (after a step): reference layout as above */
int pmcgpu_get_proposal(pmcgpu_ctx *ctx, double *o_wght, double *o_mean, double *o_L, double *o_detL, int *o_alive);

/* ---- multi-GPU: one context per GPU, samples sharded, statistics all-reduced (pmc_mpi.c:44-125) --------- */
/* shard layout of send_simulation (pmc_mpi.c:344-352) generalised to "every rank works": rank r of P owns
 * [first, first+len) with len = N/P (+1 for the first N%P ranks) */
void pmcgpu_shard_range(long N_total, int rank, int nranks, long *first, long *len);
/* NCCL (dlopen'ed libnccl.so.2): id is PMCGPU_NCCL_ID_BYTES bytes obtained on rank 0 and broadcast by the caller */
#define PMCGPU_NCCL_ID_BYTES 128
int pmcgpu_nccl_unique_id(void *id_out);
int pmcgpu_comm_init(pmcgpu_ctx *ctx, const void *id, int rank, int nranks);
/* alternative for callers that own the collective (tests on CPU, MPI hosts): called on a HOST buffer of n doubles,
 * op 0 = sum, 1 = max; must return 0 on success */
typedef int (*pmcgpu_allreduce_fn)(void *user, double *buf, long n, int op);
int pmcgpu_comm_set_callback(pmcgpu_ctx *ctx, pmcgpu_allreduce_fn fn, void *user, int rank, int nranks);

/* Collectives for hosts that shard a pmc_simu over ranks the way libpmc_mpi does (pmc_mpi.c:44-125, 334-566), over the
 * attached communicator (NCCL or the callback).  With one rank they are no-ops / plain copies.
 *   allreduce: n host doubles in place, op 0 = sum, 1 = max.
 *   bcast: `bytes` host bytes from rank `root` to every rank (replaces send_mix_mvdens / receive_mix_mvdens and the
 *          size messages of send_simulation).
 *   gather_store: the five arrays of this rank's HBM shard [first, first + N_local) land in root's host arrays of the
 *          whole N_total-sample pmc_simu (replaces send_importance_weight / receive_importance_weight, and delivers the
 *          samples the master of the reference would have drawn itself).  Null arrays are skipped; non-root ranks may
 *          pass nulls. */
int pmcgpu_comm_allreduce(pmcgpu_ctx *ctx, double *buf, long n, int op);
int pmcgpu_comm_bcast(pmcgpu_ctx *ctx, void *buf, size_t bytes, int root);
/* every rank owns the bytes [my_off, my_off + my_len) of a host buffer of `total` bytes (disjoint ranges); afterwards all
 * ranks hold the whole buffer.  The scatter / gather messages of pmc_mpi.c:334-520 are expressed with it. */
int pmcgpu_comm_allgather_bytes(pmcgpu_ctx *ctx, void *buf, size_t total, size_t my_off, size_t my_len);
int pmcgpu_comm_gather_store(pmcgpu_ctx *ctx, int root, long N_total, long first, double *X, size_t *indices,
                             double *weights, double *log_rho, short *flg);
int pmcgpu_comm_rank(const pmcgpu_ctx *ctx, int *rank, int *nranks);

/* ---- host-side finalisation, exposed for tests and for hosts that reduce the statistics themselves -------- */
/* number of doubles in the sufficient-statistics block for (K, d): K*(2 + d + d*d) + 8 */
long pmcgpu_stats_len(int K, int d);
/* new proposal from reduced statistics: W[K], G[K], s1[K*d] = sum g (x - pivot), S2[K*d*d] = sum g (x-pivot)(x-pivot)^T,
 * count[K]; old mixture (for the retry snapshot) in reference layout; writes the new one in place.  Host only: it is
 * the literal cleanup_after_update logic (pmc.c:1471-1533) around a GSL-1.14-semantics Cholesky. */
int pmcgpu_finalize_update(int K, int d, const double *W, const double *G, const double *s1, const double *S2,
                           const double *pivot, int pivot_per_component, const size_t *count, long N_total,
                           const size_t *band_limit, int *retry, double *wght, double *mean, double *L, double *detL,
                           int *alive);
/* mvdens_cholesky_decomp (mvdens.c:229-277): in place; returns mv_cholesky (-706) and restores a[] if not PD */
int pmcgpu_cholesky(int d, double *a, double *detL);

#ifdef __cplusplus
}
#endif
#endif
