/* Test-infrastructure stub: a no-op <mpi.h> so that the reference's translation units compile
 * without an MPI installation.  Only the broadcast<> template (reference mpi_util.h:326-354)
 * names MPI symbols and it is never instantiated on the hot path. */
#ifndef KWAGE_ORACLE_STUB_MPI_H
#define KWAGE_ORACLE_STUB_MPI_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_SUCCESS 0
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline double MPI_Wtime() { return 0.0; }
#endif
