/*
 * ldc_b200_comm.h -- NCCL plumbing for the row-strip decomposition (shared library
 * lid_driven_cavity_problem_b200/csrc/libldc_b200_comm.so).  Kept apart from libldc_b200.so so
 * the compute library has no NCCL dependency; the solver talks to it through the function table
 * `ldc_comm_ops` (include/ldc_b200.h), so any other transport can be plugged in the same way.
 *
 * The reference has no multi-rank path at all (sequential PETSc objects only,
 * petsc_solver_wrapper.py:75,77,117,125); this is the NVLink equivalent SURVEY 8(e) asks for:
 * one cell row (3*nx doubles) exchanged with the strip below and above per operator
 * application, and small all-reduces for norms / Gram-Schmidt coefficients.
 *
 * Bootstrap: rank 0 calls ldc_comm_unique_id, the 128 bytes travel through any host channel
 * (torch.distributed broadcast in this repo), every rank calls ldc_comm_create.
 */
#ifndef LDC_B200_COMM_H
#define LDC_B200_COMM_H

#include <stdint.h>

#include "ldc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldc_comm ldc_comm;

#define LDC_COMM_ID_BYTES 128

LDC_API int ldc_comm_unique_id(void *id_out);
LDC_API int ldc_comm_create(const void *id, int32_t nranks, int32_t rank, ldc_comm **out);
LDC_API void ldc_comm_destroy(ldc_comm *c);
LDC_API const char *ldc_comm_last_error(void);

/* Exchange the one-row halos of a strip: `first_owned_row_dev` points at the first owned row of a
 * buffer laid out [ghost below | row_count owned rows | ghost above], each row `row_doubles`
 * doubles.  Rank r sends its first owned row to r-1 and its last owned row to r+1 and receives
 * the ghost rows from them; ranks 0 / nranks-1 have no neighbour below / above.  One NCCL group. */
LDC_API int ldc_comm_halo_exchange(ldc_comm *c, double *first_owned_row_dev, int64_t row_doubles,
                                   int64_t row_count, void *stream);
/* in-place all-reduce of `count` doubles; op 0 = sum, 1 = max */
LDC_API int ldc_comm_allreduce(ldc_comm *c, double *buf_dev, int64_t count, int32_t op, void *stream);
/* in-place all-gather of `count` doubles per rank (rank r's piece at buf_dev + r*count) */
LDC_API int ldc_comm_allgather(ldc_comm *c, double *buf_dev, int64_t count, void *stream);
/* Peer-mapped device memory for ldc_residual_peer (include/ldc_b200.h): a COLLECTIVE call, every
 * rank passing the same `bytes`.  *local is a zero-filled cudaMalloc region of this rank; *prev /
 * *next are the regions of ranks r-1 / r+1 mapped into this process (CUDA IPC over NVLink; NULL
 * where there is no neighbour), so offset k in *prev is offset k of rank r-1's *local.
 * ldc_comm_peer_free is collective too. */
LDC_API int ldc_comm_peer_alloc(ldc_comm *c, int64_t bytes, void **local, void **prev, void **next);
LDC_API int ldc_comm_peer_free(ldc_comm *c, void *local);
/* function table to hand to ldc_solver_set_comm */
LDC_API const ldc_comm_ops *ldc_comm_get_ops(ldc_comm *c);

#ifdef __cplusplus
}
#endif
#endif
