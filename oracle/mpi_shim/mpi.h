/* Single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference back end (libcirc/probability.c, innerprod.c, utils/comms.c)
 * is written against MPI.  With world_size == 1 none of its send/recv sites is
 * ever reached (every fan-out loop starts at dest = 1, innerprod.c:69,79,182,192;
 * probability.c:184,204), so a rank-0-of-1 shim is enough to compile and run the
 * reference sources unmodified with plain gcc.  Nothing here is derived from an
 * MPI implementation. */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H
#include <stdlib.h>
#include <stdio.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int unused; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
    (void)b; (void)n; (void)t; (void)dest; (void)tag; (void)c;
    fprintf(stderr, "mpi shim: MPI_Send reached with world_size 1\n"); abort();
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s) {
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s;
    fprintf(stderr, "mpi shim: MPI_Recv reached with world_size 1\n"); abort();
}
#endif
