/* Serial (single-rank) MPI stand-in used ONLY to compile the unmodified
 * reference (picksc/ppic2/pplib2.c, ppush2.c and skeletor/cython/*.pyx) into
 * oracle/_ref/ as a test oracle / CPU baseline.  Test infrastructure, not
 * product code.  rank = 0, size = 1; collectives are memcpy; the point-to-point
 * calls are never reached because cppmove2 / cpptpose take their nvp == 1
 * branches (reference picksc/ppic2/pplib2.c:715-730, 446-453). */
#ifndef SKB_SERIAL_MPI_H
#define SKB_SERIAL_MPI_H
#include <string.h>
#include <stdlib.h>
#include <time.h>

#define MPI_VERSION 3
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Message;
typedef struct { int count; int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0
/* low byte = element size in bytes */
#define MPI_INT 0x104
#define MPI_FLOAT 0x204
#define MPI_DOUBLE 0x308
#define MPI_COMPLEX 0x408
#define MPI_DOUBLE_COMPLEX 0x510
#define MPI_LONG 0x608
#define MPI_SUM 1
#define MPI_MAX 2

static inline int MPI_Initialized(int *flag) { *flag = 1; return 0; }
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline double MPI_Wtime(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t,
                                MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; memcpy(r, s, (size_t)n * (size_t)(t & 0xff)); return 0;
}
/* point-to-point: unreachable with one rank; abort loudly if ever called */
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag,
                            MPI_Comm c, MPI_Request *q) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)q; abort(); return 1;
}
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag,
                            MPI_Comm c, MPI_Request *q) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)q; abort(); return 1;
}
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag,
                           MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; abort(); return 1;
}
static inline int MPI_Wait(MPI_Request *q, MPI_Status *s) { (void)q; (void)s; abort(); return 1; }
static inline int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *n) {
  (void)s; (void)t; *n = 0; abort(); return 1;
}
#endif
