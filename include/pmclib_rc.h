/* pmclib_rc.h — the config / driver layer around the PMC path: pmclib's readConf API (pmctools/include/readConf.h:44-89) and
 * pmc_rc API (pmclib/include/pmc_rc.h:66-79) with the reference's names and signatures, implemented WITHOUT Lua
 * (pmclib_b200/csrc/pmc_rc.cpp): parameter files are read by a small evaluator of the Lua subset pmclib's parameter files
 * are written in — assignments to (nested) names, numbers, strings, booleans, table constructors, references to tables
 * defined earlier, arithmetic on numbers, comments.  No functions, no control flow.
 *
 * With this header (or the reference's readConf.h + pmc_rc.h, whose prototypes it mirrors) pmcexec/vanilla_pmc.c:47-133 and
 * vanilla_importance.c compile and link against libpmcb200.so unchanged, Lua not needed.
 */
#ifndef PMCLIB_RC_H
#define PMCLIB_RC_H
#include "pmclib_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __READCONF__
#define RC_FNMLEN (4096 * 4)
#define RC_MAX_DEPTH 32
/* layout of readConf.h:29-41 (lua_State* -> void*); callers treat it as opaque */
typedef struct {
  void *alias;
  char alias_prefix[RC_FNMLEN + 1];
  char fnm[RC_FNMLEN + 1];
  void **mem;
  size_t mem_sz;
  size_t mem_cur;
  void *L;
  char last_key[RC_FNMLEN * RC_MAX_DEPTH + RC_MAX_DEPTH + 1];
  FILE *display;
  char root[RC_FNMLEN];
  char **read_pars;
} confFile;
#define rc_size_error (-201)
#define rc_lua_error (-202)
#define rc_maxdepth_error (-203)
#define rc_bad_type (-204)
#define rc_bad_key (-205)

confFile *rc_open(char *fnm, error **err);
confFile *rc_init(char *fnm, error **err);
confFile *rc_init_from_args(int argc, char **argv, error **err);
void rc_do_args(confFile *rc, int argc, char **argv, char shortopt, char *longopt, error **err);
void rc_close(confFile **self);
void rc_toggle_display(confFile *rc, FILE *display, error **err);
confFile *rc_alias(confFile *rc, char *prefix, error **err);
int rc_has_key(confFile *rc, char *key, error **err);
double rc_get_real(confFile *rc, char *key, error **err);
long rc_get_integer(confFile *rc, char *key, error **err);
char *rc_get_string(confFile *rc, char *key, error **err);
double rc_safeget_real(confFile *rc, char *key, double safeguard, error **err);
long rc_safeget_integer(confFile *rc, char *key, long safeguard, error **err);
char *rc_safeget_string(confFile *rc, char *key, char *safeguard, error **err);
int rc_is_array(confFile *rc, char *key, error **err);
size_t rc_array_size(confFile *rc, char *key, error **err);
size_t rc_get_real_array(confFile *rc, char *key, double **parray, error **err);
size_t rc_get_integer_array(confFile *rc, char *key, long **parray, error **err);
size_t rc_get_integer_array_asint(confFile *rc, char *key, int **parray, error **err);
size_t rc_get_string_array(confFile *rc, char *key, char ***parray, error **err);
size_t rc_get_real_asarray(confFile *rc, char *key, double **parray, int sz, error **err);
size_t rc_get_real_assquarematrix(confFile *rc, char *key, double **res, int sz, error **err);
void rc_free(confFile *rc, void **smthng, error **err);
void rc_add_mem(confFile *rc, void *ptr, error **err);
void *rc_add_malloc(confFile *rc, size_t sz, error **err);
#endif

#ifndef __PMC_RC_H
#define dl_err (-101010)
typedef void *(rcinit_smthng)(confFile *, char *, error **err);
distribution *init_distribution_from_rc(confFile *rc, char *root, error **err);   /* pmc_rc.c:35-82   */
void *rcinit_combine_distributions(confFile *rc, char *root, error **err);        /* pmc_rc.c:84-171  */
void *rcinit_add_gaussian_prior(confFile *rc, char *root, error **err);           /* pmc_rc.c:233-305 */
parabox *parabox_from_rc(confFile *rc, char *root, error **err);                  /* pmc_rc.c:356-389 */
mvdens *rc_mvdens(confFile *rc, mvdens *mv, error **err);                         /* pmc_rc.c:391-433 */
double *rc_get_weight(confFile *rc, int sz, error **err);                         /* pmc_rc.c:435-465 */
mix_mvdens *rc_mix_mvdens(confFile *rc, error **err);                             /* pmc_rc.c:467-599 */
distribution *rcinit_mix_mvdens(confFile *rc, char *root, error **err);           /* pmc_rc.c:601-618 */
pmc_simu *init_importance_from_rc(confFile *rc, char *root, error **err);         /* pmc_rc.c:620-665 */
pmc_simu *init_pmc_from_rc(confFile *rc, char *root, error **err);                /* pmc_rc.c:712-721 */
gsl_rng *rc_gsl_rng(confFile *rc, char *root, error **err);                       /* pmc_rc.c:723-768 */
char *get_itername(confFile *rc, char *key, int iter, error **err);               /* pmc_rc.c:770-794 */
void *rcinit_bentGauss(confFile *rc, char *root, error **err);                    /* examples/target/sampleTarget.c:61-84 */
#endif
#ifndef __PMC_MPI_H
/* libpmc_mpi's config functions (pmc_mpi.h:94-99; the parameter-file name "mix_mvdens_mpi" selects the broadcast proposal) */
mix_mvdens *rc_mix_mvdens_mpi(confFile *rc, error **err);                         /* pmc_mpi.c:236-250 */
distribution *rcinit_mix_mvdens_mpi(confFile *rc, error **err);                   /* pmc_mpi.c:252-272 */
pmc_simu *init_importance_mpi_from_rc(confFile *rc, char *root, error **err);     /* pmc_mpi.c:274-318 */
pmc_simu *init_pmc_mpi_from_rc(confFile *rc, char *root, error **err);            /* pmc_mpi.c:320-330 */
#endif

/* rc_gsl_rng hands out a gsl_rng of GSL's own layout (type, state) backed by an MT19937 of this library (the default
 * generator of pmc_rc.c:733), so hosts without GSL can run the drivers; release it with gsl_rng_free or this function */
void pmcb200_rng_free(gsl_rng *r);

#ifdef __cplusplus
}
#endif
#endif
