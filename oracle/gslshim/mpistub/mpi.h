/* mpi.h stub — TEST INFRASTRUCTURE ONLY.  This image has no MPI; the reference's MPI drivers (pmcexec/vanilla_pmc_mpi.c,
 * vanilla_importance_mpi.c) only use MPI themselves for MPI_Init / MPI_Comm_rank / MPI_Finalize (every exchange goes through
 * libpmc_mpi, which libpmcb200.so implements over NCCL).  These few entry points are provided here from the launcher's
 * environment, so that the drivers can be compiled unmodified and started one process per GPU.  A real MPI host links its
 * own MPI; libpmcb200 reads the ranks mpirun exports (OMPI_COMM_WORLD_RANK, PMI_RANK, ...). */
#ifndef PMCB200_MPI_STUB_H
#define PMCB200_MPI_STUB_H
#include <stdlib.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int dummy; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_LONG 1
#define MPI_DOUBLE 2
#define MPI_SHORT 3
#define MPI_CHAR 4
static inline int pmcb200_stub_env(const char *a, const char *b, int dflt) {
  const char *e = getenv(a);
  if (!e) e = getenv(b);
  return e ? atoi(e) : dflt;
}
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = pmcb200_stub_env("PMCB200_RANK", "RANK", 0); return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = pmcb200_stub_env("PMCB200_WORLD_SIZE", "WORLD_SIZE", 1); return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code ? code : 1); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
#endif
