/* pmclib_abi.h — the part of pmclib's public C API that libpmcb200.so implements as a drop-in.
 *
 * Existing pmclib callers are compiled against pmclib's own headers and only need to LINK against libpmcb200.so
 * instead of libpmc / libmvdens / liberrorio for the PMC path: every struct below has the reference's layout and
 * every function the reference's name, signature and error behaviour.  This header exists for new C callers, for the
 * tests (tests/c/dropin_driver.c) and as the written contract; citations are file:line in CosmoStat/pmclib.
 *
 * What runs where: everything that loops over samples runs in CUDA kernels (see pmcb200.h); the functions here are
 * host orchestration over the reference-layout host structs (psim->X etc. are HOST arrays, copied to and from the
 * GPU inside each call, exactly the memory contract the reference offers).  There is no CPU implementation of the
 * sample loops: without a CUDA device every compute entry point raises pmc_base-24 through the error chain.
 */
#ifndef PMCLIB_ABI_H
#define PMCLIB_ABI_H
#include <stddef.h>
#include <stdio.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errorlist (pmctools/include/errorlist.h:20-53) -------------------------------------------------------- */
#ifndef __ERRORLIST_H
#define TXT_SZ 4192
#define WHR_SZ 2048
typedef struct _err {
  char errWhere[WHR_SZ];
  char errText[TXT_SZ];
  int errValue;
  struct _err *next;
} error;
#define forwardErr -123456789
#define undefErr -987654321
#define noErr 0
#define placeHolderError 1020304050
error *newError(int errV, char *where, char *text, error *linkTo, error *addErr);
error *newErrorVA(int errV, const char *func, char *where, char *text, error *linkTo, error *addErr, ...);
void stringError(char *str, error *err);
void printError(FILE *flog, error *err);
int getErrorValue(error *err);
void purgeError(error **err);
error *initError(void);
void endError(error **err);
int _isError(error *err);
#define isError(err) (_isError(err))
int pickleError(error *err, void **buf);   /* errorlist.c:126-150 */
error *unpickleError(void *buf, int len);  /* errorlist.c:152-169 */
#endif

/* ---- GSL types that appear by value in the ABI (mvdens.h:83-110, pmc.h:133) ---------------------------------- */
#ifndef __GSL_BLOCK_DOUBLE_H__
typedef struct { size_t size; double *data; } gsl_block;
#endif
#ifndef __GSL_VECTOR_DOUBLE_H__
typedef struct { size_t size; size_t stride; double *data; gsl_block *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
#endif
#ifndef __GSL_MATRIX_DOUBLE_H__
typedef struct { size_t size1; size_t size2; size_t tda; double *data; gsl_block *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
#endif
#ifndef __GSL_RNG_H__
typedef struct {
  const char *name;
  unsigned long int max, min;
  size_t size;
  void (*set)(void *state, unsigned long int seed);
  unsigned long int (*get)(void *state);
  double (*get_double)(void *state);
} gsl_rng_type;
typedef struct { const gsl_rng_type *type; void *state; } gsl_rng;
#endif

/* ---- parabox (pmclib/include/parabox.h:22-26) ------------------------------------------------------------------ */
#ifndef __PARABOX__
typedef struct { int ndim; int nfaces; double *faces; } parabox;
parabox *init_parabox(int parasize, error **err);
int isinBox(const parabox *bx, double *pos, error **err);
void add_lobound(parabox *bx, int coord, double vmin, error **err);
void add_hibound(parabox *bx, int coord, double vmax, error **err);
void add_slab(parabox *bx, int coord, double vmin, double vmax, error **err);
void add_face(parabox *bx, double *dirvec, double threshold, error **err);
void add_halfspace(parabox *bx, int coord, double sign, double val, error **err);
void free_parabox(parabox **bx);
void print_parabox(parabox *bx, FILE *mstream);                                        /* parabox.c:147-186 */
#endif

/* ---- mvdens / mix_mvdens (pmctools/include/mvdens.h:83-155) -------------------------------------------------------- */
#ifndef __MVDENS_H
typedef struct {
  size_t ndim;
  void *buf;
  int own_buf;
  double *mean;
  double *std;
  double *x_tmp;
  gsl_vector_view mean_view_container;
  gsl_vector *mean_view;
  gsl_vector_view x_tmp_view_container;
  gsl_vector *x_tmp_view;
  gsl_matrix_view std_view_container;
  gsl_matrix *std_view;
  size_t band_limit;
  int chol;
  int df;
  double detL;
} mvdens;
typedef struct {
  size_t ncomp, ndim;
  int init_cwght;
  double *wght, *cwght;
  gsl_vector_view wght_view_container, cwght_view_container;
  gsl_vector *wght_view, *cwght_view;
  mvdens **comp;
  void *buf_comp;
} mix_mvdens;
mvdens *mvdens_alloc(size_t ndim, error **err);
mvdens *mvdens_init(mvdens *g, size_t nndim, double *mn, double *std, error **err);
void mvdens_empty(mvdens *g);
void mvdens_free(mvdens **g);
void mvdens_free_void(void **g);
void mvdens_print(FILE *where, mvdens *what);
void mvdens_dump(FILE *where, mvdens *what);
void mvdens_set_band_limit(mvdens *self, size_t value);
double determinant(const double *L, size_t ndim);
void mvdens_cholesky_decomp(mvdens *self, error **err);
double scalar_product(mvdens *g, const double *x, error **err);
double mvdens_log_pdf(mvdens *g, const double *x, error **err);
double mvdens_log_pdf_void(void *g, const double *x, error **err);
mix_mvdens *mix_mvdens_alloc(size_t ncomp, size_t size, error **err);
void mix_mvdens_free_void(void **m);
void mix_mvdens_free(mix_mvdens **m);
void mix_mvdens_print(FILE *where, mix_mvdens *what);
void mix_mvdens_dump(FILE *where, mix_mvdens *what);
mix_mvdens *mix_mvdens_dwnp(FILE *where, error **err);
void mix_mvdens_set_band_limit(mix_mvdens *self, size_t value);
void mix_mvdens_cholesky_decomp(mix_mvdens *self, error **err);
size_t ranindx(double *cw, size_t nE, gsl_rng *r, error **err);
double mix_mvdens_log_pdf(mix_mvdens *m, const double *x, error **err);
double mix_mvdens_log_pdf_void(void *m, const double *x, error **err);
/* second half of mvdens.h:113-155 (pmc_dropin2.cpp) */
mvdens *mvdens_dwnp(FILE *where, error **err);                                         /* mvdens.c:146-188 */
mvdens *mvdens_read_and_chol(FILE *where, error **err);                                /* mvdens.c:190-204 */
void mvdens_chdump(const char *name, mvdens *what, error **err);                       /* mvdens.c:136-144 */
void mvdens_from_meanvar(mvdens *m, const double *pmean, const double *pvar, double correct); /* mvdens.c:62-71 */
double ln_determinant(const double *L, size_t ndim, error **err);                      /* mvdens.c:215-227 */
double *mvdens_ran(double *dest, mvdens *g, gsl_rng *r, error **err);                  /* mvdens.c:279-317 */
double mvdens_inverse(mvdens *m, error **err);                                         /* mvdens.c:393-410 */
void mix_mvdens_copy(mix_mvdens *target, const mix_mvdens *source, error **err);       /* mvdens.c:418-431 */
double effective_number_of_components(const mix_mvdens *self, error **err);            /* mvdens.c:500-511 */
size_t mix_mvdens_size(size_t ncomp, size_t ndim);                                     /* mvdens.c:585-594 */
void *serialize_mix_mvdens(mix_mvdens *self, size_t *sz, error **err);                 /* mvdens.c:596-660 */
mix_mvdens *deserialize_mix_mvdens(void *serialized, size_t sz, error **err);          /* mvdens.c:662-720 */
double *mix_mvdens_ran(double *dest, size_t *index, mix_mvdens *m, gsl_rng *r, error **err); /* mvdens.c:730-764 */
#endif

/* ---- plugin interface (pmclib/include/allmc.h:21-33, distribution.h:25-41) ------------------------------------------ */
#ifndef __ALLMC_H
struct _pmc_simu_struct_;
typedef double posterior_log_pdf_func(void *, const double *, error **);
typedef void retrieve_ded_func(const void *, double *, error **);
typedef void(posterior_log_free)(void **);
typedef long simulate_func(struct _pmc_simu_struct_ *, void *, gsl_rng *, parabox *, error **);
typedef void filter_func(struct _pmc_simu_struct_ *, void *, error **);
typedef void update_func(void *, struct _pmc_simu_struct_ *, error **);
typedef void *mpi_exchange_func(void *, error **);
typedef double first_derivative_func(void *, int, const double *, error **err);
typedef double second_derivative_func(void *, int, int, const double *, error **err);
#endif
#ifndef __MC_DIST
typedef char _char_name[1024];
typedef struct _distribution_struct_ {
  int ndim, n_ded, ndef;
  double *pars;
  int *def;
  void *data;
  posterior_log_pdf_func *log_pdf;
  retrieve_ded_func *retrieve;
  posterior_log_free *free;
  simulate_func *simulate;
  mpi_exchange_func *broadcast_mpi;
  first_derivative_func *f_der;
  second_derivative_func *d_der;
  void *dlhandle;
  _char_name *name;
} distribution;
distribution *init_distribution_full(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                     simulate_func *simulate, int nded, retrieve_ded_func *retrieve, error **err);
distribution *init_simple_distribution(int ndim, void *data, posterior_log_pdf_func *log_pdf,
                                       posterior_log_free *freef, error **err);
distribution *init_distribution(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                simulate_func *simulate, error **err);
void free_distribution(distribution **pdist);
double distribution_lkl(void *pdist, const double *pars, error **err);
double *distribution_fill_pars(distribution *dist, double *pars, error **err); /* distribution.c:193-209 (global, undeclared there) */
void distribution_retrieve(const void *pdist, double *pars, error **err);
void distribution_set_default(distribution *dist, int ndef, int *idef, double *vdef, error **err);
/* names, broadcast hook, combined targets and Gaussian priors (distribution.h:66-119, pmc_dropin2.cpp) */
int distribution_get_name(distribution *dist, char *name, error **err);
int *distribution_get_names(distribution *dist, int nname, char **name, int includeded, error **err);
void distribution_set_names(distribution *dist, char **name, error **err);
void distribution_set_broadcast(distribution *dist, mpi_exchange_func *broadcast, error **err);
void distribution_set_default_name(distribution *dist, int ndef, char **idef, double *vdef, error **err);
typedef struct {
  double *pars, *pded, *dummy;
  int *ded_from, **dim, **ded;
  distribution **dist;
  int ndim, nded, ndist, ndummy;
} comb_dist_data;
distribution *combine_distribution_init(int ndim, int nded, error **err);
distribution *combine_distribution_simple_init(int ndim, error **err);
double combine_lkl(void *pcbd, const double *pars, error **err);
void combine_retrieve(const void *pcbd, double *pded, error **err);
void combine_free(void **pcbd);
void add_to_combine_distribution(distribution *comb, distribution *addon, int *dim_idx, int *ded_idx, error **err);
void add_to_combine_distribution_name(distribution *comb, distribution *addon, error **err);
distribution *add_gaussian_prior(distribution *orig, int ndim, int *idim, double *loc, double *var, error **err);
distribution *add_gaussian_prior_2(distribution *orig, int ndim, int *idim, double *loc, double *var, error **err);
distribution *add_gaussian_prior_name(distribution *orig, int ndim, char **idim, double *loc, double *var, error **err);
distribution *add_gaussian_prior_2_name(distribution *orig, int ndim, char **idim, double *loc, double *var, error **err);
#endif

/* ---- pmc_simu (pmclib/include/pmc.h:76-205) --------------------------------------------------------------------------- */
#ifndef __PMC_H
#define MC_NORM 1
#define MC_UNORM 0
typedef struct _pmc_simu_struct_ {
  long nsamples;
  int ndim, n_ded;
  void *data;
  double *X, *X_ded, *weights;
  double *log_rho;
  short *flg;
  size_t *indices;
  int isLog;
  double logSum;
  double maxW, maxR;
  parabox *pb;
  distribution *target;
  distribution *proposal;
  int prop_print_step;
  void *filter_data;
  filter_func *pmc_filter;
  update_func *pmc_update;
  int retry;
  int mpi_rank;
  int mpi_size;
} pmc_simu;
pmc_simu *pmc_simu_init(long nsample, int ndim, error **err);
pmc_simu *pmc_simu_init_plus_ded(long nsample, int ndim, int n_ded, error **err);
void pmc_simu_free(pmc_simu **self);
void pmc_simu_realloc(pmc_simu *psim, long newsamples, error **err);
void pmc_simu_init_target(pmc_simu *psim, distribution *target, parabox *pb, error **err);
void pmc_simu_init_proposal(pmc_simu *psim, distribution *proposal, int prop_print_step, error **err);
void pmc_simu_init_pmc(pmc_simu *psim, void *filter_data, filter_func *pmc_filter, update_func *pmc_update);
void pmc_simu_init_classic_importance(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data,
                                      int prop_print_step, error **err);
void pmc_simu_init_classic_pmc(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                               void *filter_data, filter_func *pmc_filter, error **err);
size_t pmc_simu_importance(pmc_simu *psim, gsl_rng *r, error **err);
double pmc_simu_pmc_step(pmc_simu *psim, gsl_rng *r, error **err);
distribution *mix_mvdens_distribution(int ndim, void *mxmv, error **err);
long simulate_mix_mvdens_void(struct _pmc_simu_struct_ *psim, void *proposal, gsl_rng *r, parabox *pb, error **err);
long simulate_mix_mvdens(pmc_simu *psim, mix_mvdens *m, gsl_rng *r, parabox *pb, error **err);
size_t get_importance_weight(pmc_simu *psim, mix_mvdens *m, posterior_log_pdf_func *posterior_log_pdf, void *extra,
                             error **err);
size_t generic_get_importance_weight(pmc_simu *psim, void *m, posterior_log_pdf_func *proposal_log_pdf,
                                     posterior_log_pdf_func *posterior_log_pdf, void *extra, error **err);
size_t get_importance_weight_plus_ded(pmc_simu *psim, mix_mvdens *m, posterior_log_pdf_func *posterior_log_pdf,
                                      retrieve_ded_func *retrieve_ded, void *extra, error **err);
size_t generic_get_importance_weight_and_deduced(pmc_simu *psim, void *m, posterior_log_pdf_func *proposal_log_pdf,
                                                 posterior_log_pdf_func *posterior_log_pdf,
                                                 retrieve_ded_func *retrieve_ded, void *extra, error **err);
size_t generic_get_importance_weight_and_deduced_verb(pmc_simu *psim, void *proposal_data,
                                                      posterior_log_pdf_func *proposal_log_pdf,
                                                      posterior_log_pdf_func *posterior_log_pdf,
                                                      retrieve_ded_func *retrieve_ded, void *target_data,
                                                      double t_iter_tempered, int quiet, error **err);
double normalize_importance_weight(pmc_simu *psim, error **err);
void update_prop(mix_mvdens *m, pmc_simu *psim, error **err);
void update_prop_rb_void(void *m, pmc_simu *psim, error **err);
void update_prop_rb(mix_mvdens *m, pmc_simu *psim, error **err);
double perplexity(pmc_simu *psim, int normalize, error **err);
double perplexity_and_ess(pmc_simu *psim, int normalize, double *ess, error **err);
double evidence(pmc_simu *psim, double *ln_evi, error **err);
void pmc_simu_dump(FILE *where, pmc_simu *psim, error **err);
void out_pmc_simu(const char *name, const pmc_simu *psim, error **err);                 /* pmc.c:306-325 */
int *filter_histogram_init(int nbins, int clipleft, int clipright, error **err);       /* pmc.c:699-707 */
void filter_histogram(pmc_simu *psim, void *filter_data, error **err);                  /* pmc.c:709-770 */
double mean_from_psim(double *X, double *w, short *flg, int nsamples, int ndim, int a); /* pmc.c:1108-1120 */
FILE *fopen_err(const char *fname, const char *mode, error **err);  /* io.c:103-109 */
void *malloc_err(size_t sz, error **err);                          /* io.c:111-120 */
void *calloc_err(size_t nl, size_t sz1, error **err);              /* io.c:122-131 */
/* the rest of liberrorio's io.h:34-75 (pmc_hostio.cpp): host-side file and memory helpers of drivers and target likelihoods */
#define io_base -200
#define io_file (-2 + io_base)
unsigned int numberoflines(const char *name, error **err);                                   /* io.c:13-27 */
unsigned int numberoflines_comments(const char *name, unsigned int *ncomment, error **err);  /* io.c:30-51 */
int read_double(char **p, double *x);                                                         /* io.c:57-62 */
int read_int(char **p, int *x);                                                               /* io.c:64-69 */
double *readASCII(char *filename, int *fnx, int *fny, error **err);                           /* io.c:72-97 */
void *realloc_err(void *ptr, size_t sz, error **err);                                         /* io.c:136-145 */
void *resize_err(void *orig, size_t sz_orig, size_t sz_new, int free_orig, error **err);      /* io.c:151-174 */
size_t read_line(FILE *fp, char *buff, int bmax, error **err);                                /* io.c:176-201 */
void *read_any_vector(const char *fnm, size_t n, char *prs, size_t sz, error **err);          /* io.c:204-212 */
void *read_any_list(const char *fnm, size_t *n, char *prs, size_t sz, error **err);           /* io.c:218-227 */
void *read_any_list_count(const char *fnm, size_t *n, char *prs, size_t sz, size_t *nlines, error **err); /* io.c:236-323 */
double *read_double_vector(const char *fnm, size_t n, error **err);                           /* io.c:326-387 */
double *read_double_list(const char *fnm, size_t *n, error **err);
int *read_int_vector(const char *fnm, size_t n, error **err);
int *read_int_list(const char *fnm, size_t *n, error **err);
float *read_float_vector(const char *fnm, size_t n, error **err);
float *read_float_list(const char *fnm, size_t *n, error **err);
long *read_long_vector(const char *fnm, size_t n, error **err);
long *read_long_list(const char *fnm, size_t *n, error **err);
void write_bin_vector(void *buf, const char *fnm, size_t n, error **err);                     /* io.c:422-436 */
void *read_bin_vector(const char *fnm, size_t n, error **err);                                /* io.c:389-402 */
void *read_bin_list(const char *fnm, size_t *n, error **err);                                 /* io.c:404-420 */
void chomp(char *x);                                                                          /* io.c:438-442 */
int is_directory(const char *path);                                                           /* io.c:490-502 */
time_t start_time(FILE *FOUT);                                                                /* io.c:444-488 */
void end_time(time_t t_start, FILE *FOUT);
clock_t start_clock(FILE *FOUT);
void end_clock(clock_t c_start, FILE *FOUT);
void prepare_update(mix_mvdens *m, size_t *count, pmc_simu *psim, double **pbuf, error **err);                    /* pmc.c:1535-1579 */
void cleanup_after_update(mix_mvdens *m, size_t *count, double minwgt, pmc_simu *psim, double *buf, error **err); /* pmc.c:1471-1533 */
int double_cmp(const void *av, const void *bv);                                          /* pmc.c:1122-1131 */
void pmc_simu_print(FILE *where, pmc_simu *psim, double nrm);                            /* pmc.c:327-331 */
void clip_weights(pmc_simu *psim, int n, FILE *OUT, error **err);                        /* pmc.c:1134-1188 */
size_t *sample_from_pmc_simu(const pmc_simu *psim, const gsl_rng *rng, error **err);     /* pmc.c:1585-1601 */
unsigned int *resample_residual(const pmc_simu *psim, const gsl_rng *rng, error **err);  /* pmc.c:1605-1626 */
pmc_simu *pmc_simu_from_file(FILE *PMCSIM, int this_nsamples, int npar, int n_ded, mix_mvdens *proposal, int nclipw,
                             error **err);                                               /* pmc.c:373-396 */
double fill_read_pmc_simu(pmc_simu *psim, mix_mvdens *proposal, error **err);            /* pmc.c:400-445 */
size_t *sort_indices(double *list, size_t size, size_t *in_ind, size_t *buf, int order, error **err); /* pmc.c:881-1034 */
void filter_importance_weight(pmc_simu *psim, double nsig, error **err);                 /* pmc.c:829-873 (a stub there too) */
void filter_importance_weight_sig(pmc_simu *psim, double nsig, error **err);             /* pmc.c:772-827 (global, undeclared in pmc.h) */
#endif

/* ---- libpmc_mpi (pmclib/include/pmc_mpi.h:52-101): one process per GPU, NCCL instead of MPI (pmc_dropin_mpi.cpp) ------- */
#ifndef __PMC_MPI_H
pmc_simu *pmc_simu_init_mpi(long nsamples, int ndim, int nded, error **err);
size_t pmc_simu_importance_mpi(pmc_simu *psim, gsl_rng *r, error **err);
double pmc_simu_pmc_step_mpi(pmc_simu *psim, gsl_rng *r, error **err);
void *mvdens_broadcast_mpi(void *, error **err);
void pmc_simu_realloc_mpi(pmc_simu *psim, long newsamples, error **err);
void pmc_simu_init_classic_importance_mpi(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data,
                                          int prop_print_step, error **err);
void pmc_simu_init_classic_pmc_mpi(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                                   void *filter_data, filter_func *pmc_filter, error **err);
distribution *init_distribution_full_mpi(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                         simulate_func *simulate, int nded, retrieve_ded_func *retrieve,
                                         mpi_exchange_func *broadcast_mpi, error **err);
distribution *init_distribution_mpi(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                    simulate_func *simulate, mpi_exchange_func *broadcast_mpi, error **err);
void distribution_broadcast(distribution *dist, error **err);
distribution *mix_mvdens_distribution_mpi(int ndim, void *mxmv, error **err);
void send_mix_mvdens(mix_mvdens *m, int nproc, error **err);
mix_mvdens *receive_mix_mvdens(int myid, int nproc, error **err);
/* the message-level functions (pmc_mpi.c:334-520), for hand-written master / client loops: same slice layout */
size_t send_simulation(pmc_simu *psim, int nproc, error **err);
void receive_simulation(pmc_simu *psim, int basetag, int myid, error **err);
void send_importance_weight(int myid, int basetag, pmc_simu *psim, size_t nok);
size_t receive_importance_weight(pmc_simu *psim, int nproc, int master_cnt, int master_sample, error **err);
#endif
/* rank / size / rendezvous path resolved from the launcher's environment, and the rendezvous primitive (tests) */
int pmcb200_mpi_world(int *rank, int *size, char *rdv, int rdv_len);
int pmcb200_mpi_rendezvous(const char *path, int rank, void *blob, long bytes, double timeout_s);

/* ---- example target shipped with the reference (examples/target/sampleTarget.h) ---------------------------------------- */
#ifndef __SMP_TARGET__
typedef struct { int ndim; double b, sig1; } bentGauss;
bentGauss *init_bentGauss(int ndim, double b, double sig1, error **err);
distribution *init_bentGauss_distribution(int ndim, double b, double sig1, error **err);
double bentGauss_lkl(void *bg, double *pars, error **err);
void bentGauss_free(void **bg);
#endif

/* ---- additions of this library (not in pmclib) ---------------------------------------------------------------------------- */
/* Declare a distribution whose log_pdf is an opaque user function to be one of the densities the GPU evaluates itself
 * (kind = PMCGPU_TGT_* of pmcb200.h; data must then be a bentGauss*, mvdens* or mix_mvdens*). */
int pmcb200_register_target(distribution *dist, int kind);
/* PMCB200_RESIDENT=1 (or this call) keeps X/weights/log_rho/indices/flg in HBM after pmc_simu_pmc_step; this call
 * copies them into the host arrays of psim on demand. */
void pmcb200_set_resident(int on);
int pmcb200_sync_host(pmc_simu *psim, error **err);
/* the pmcgpu context behind a psim (created on first use), for callers that mix both APIs */
void *pmcb200_context(pmc_simu *psim, error **err);

#ifdef __cplusplus
}
#endif
#endif
