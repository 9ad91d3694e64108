/*
 * ldc_b200.h -- C-ABI of the B200-native `cuda` backend for the coupled lid-driven-cavity
 * solve (shared library lid_driven_cavity_problem_b200/csrc/libldc_b200.so, sm_100a).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every pointer named *_dev is DEVICE memory owned by the caller (the Python host keeps
 *     them in torch tensors and passes tensor.data_ptr()).  The library owns only opaque
 *     handles (ldc_solver*).
 *   - Every call is stream-ordered on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and returns immediately unless stated otherwise.
 *   - Return value: 0 = ok, <0 = error (ldc_last_error() gives the text).  No exception
 *     crosses the ABI.  One host thread per handle.
 *
 * Unknown vector layout (reference: residual_function/pure_python_residual_function.py:10-22,
 * nonlinear_solver/_common.py:4-18): cell c = i + nx*j, X[3c]=P, X[3c+1]=U (east face; dummy
 * when i==nx-1), X[3c+2]=V (north face; dummy when j==ny-1).  Residual rows use the same slots.
 */
#ifndef LDC_B200_H
#define LDC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDC_API __attribute__((visibility("default")))

#define LDC_OK 0
#define LDC_ERR_INVALID (-1)
#define LDC_ERR_CUDA (-2)
#define LDC_ERR_ALLOC (-3)
#define LDC_ERR_STATE (-4)

/* Description of the grid one call works on.  Replaces the attribute reads the reference's
 * native residual does on the Python `graph` (c++/residual_function.cpp:16-43). */
typedef struct ldc_grid {
    int32_t nx, ny;          /* global cavity grid (pressure mesh), nx >= 3, ny >= 2          */
    int32_t row_begin;       /* first grid row owned by this call (row-strip decomposition)  */
    int32_t row_count;       /* rows owned; full grid: row_begin=0,row_count=ny              */
    int32_t batch;           /* independent cavities stored back to back (Reynolds ensemble) */
    int32_t use_cds;         /* central differencing instead of upwind (graph.use_cds)       */
    double dt, dx, dy, rho, mi, bc;
    const double *dt_dev;    /* optional per-instance dt[batch] (device); NULL -> dt          */
    const double *bc_dev;    /* optional per-instance lid speed bc[batch] (device); NULL -> bc */
} ldc_grid;

/* Strip layout: X_dev/F_dev point at the first OWNED row; when row_begin > 0 the row below
 * (ghost) lives at X_dev - 3*nx, when row_begin+row_count < ny the row above lives at
 * X_dev + 3*nx*row_count.  U_old/V_old hold owned rows only: U_old[i + (nx-1)*jl],
 * V_old[i + nx*jl] with jl the local row.  Instance b of a batch starts at
 * b*3*nx*(row_count) (+ ghosts are not supported together with batch>1). */

LDC_API const char *ldc_last_error(void);
LDC_API int ldc_version(void);
/* number of SMs etc. of the current device; returns 0 or error */
LDC_API int ldc_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, int64_t *l2_bytes);

/* ---- residual: replaces residual_function(X, graph) of c++/residual_function.cpp:10-238 and
 * the five sibling backends.  F_dev is fully overwritten (dummy rows F=X).  When
 * norm2_dev != NULL, *norm2_dev receives sum(F^2) per instance ([batch] doubles), computed by a
 * deterministic two-stage reduction (work_dev must then hold ldc_residual_work_doubles()).
 * variant: 0 = auto (4 when eligible, else 2), 1 = cell-per-thread reference kernel,
 *          2 = flux-form column-marching kernel with register prefetch,
 *          3 = flux form, rows staged per block through a TMA ring (even nx, 16-byte aligned X; else 2),
 *          4 = flux form, one bulk-copy ring per warp, two cells per lane with 128-bit accesses
 *              (even nx and 16-byte aligned X, U_old, V_old, F; else 2).
 */
LDC_API int64_t ldc_residual_work_doubles(const ldc_grid *g);
LDC_API int ldc_residual(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                 const double *V_old_dev, double *F_dev, double *norm2_dev, double *work_dev,
                 int32_t variant, void *stream);

/* ---- the same plug point with HOST buffers (what the reference's callers hold: PETSc's callback
 * and scipy's fsolve pass numpy arrays, petsc_solver_wrapper.py:42-48, scipy_solver_wrapper.py:35;
 * c++/residual_function.cpp:10-43 takes and returns host arrays).  X_host[3*nx*ny],
 * U_old_host[(nx-1)*ny], V_old_host[nx*(ny-1)] are read, F_host[3*nx*ny] is written; any
 * pageable memory.  The call is a copy pipeline over row blocks: multi-threaded staging into
 * pinned memory, H2D, the strip form of ldc_residual per block, D2H and copy-out overlap, with
 * both PCIe directions busy at once; a page-locked F_host (cudaHostAlloc / cudaHostRegister)
 * receives the D2H copies directly.  Synchronous (F_host is complete on return); buffers,
 * streams and copy threads are created once per process; calls are serialised.  batch must be
 * 1, dt_dev / bc_dev NULL.  LDC_HOST_COPY_THREADS overrides the copy-thread count (default:
 * min(8, cores / LOCAL_WORLD_SIZE)). */
LDC_API int ldc_residual_host(const ldc_grid *g, const double *X_host, const double *U_old_host,
                      const double *V_old_host, double *F_host, int32_t variant);

/* ---- residual of one row strip per GPU without a collective (the reference has no
 * multi-rank path).  Buffers use the strip layout above: [ghost row | row_count owned rows | ghost
 * row], and every strip's buffer is mapped into its neighbours' address space (CUDA IPC / peer
 * access, e.g. ldc_comm_peer_alloc).  One evaluation number `epoch` (1, 2, 3, ... identical on all
 * strips) is two launches:
 *   ldc_residual_peer_push   copies the strip's first / last owned row of X into the neighbours'
 *                            ghost rows (NVLink stores) and then stores `epoch` (release, system
 *                            scope) into the neighbours' flag words.  Small kernel; meant for a
 *                            side stream so that it overlaps the residual kernel.  X must be final
 *                            on that stream (the caller orders it after X's producer).
 *   ldc_residual_peer        the residual kernel of ldc_residual; the tiles next to a neighbour
 *                            strip wait (acquire, bounded) until their own flag word shows `epoch`
 *                            and then read the ghost row from LOCAL memory.  It performs no remote
 *                            access: a kernel that stores to peer memory pays a system-scope flush
 *                            of ~2.6 us at its end (measured on B200), the push kernel takes that
 *                            cost off the critical path.
 * Because epoch k's ghost rows overwrite those of the last evaluation on the same buffer, a
 * buffer must not be pushed into while the neighbour still evaluates an older epoch on it: rotate
 * at least two X buffers and order push(k) after the own residual(k-1) (which cannot finish
 * before the neighbours pushed k-1, i.e. before they finished k-2), or separate evaluations by
 * any collective.  A wait that does not complete within a few seconds sets *status_dev = 1 and
 * the kernel finishes anyway (results invalid) -- it never hangs the device.  batch must be 1. */
typedef struct ldc_peer_halo {
    double *prev_ghost;            /* in strip r-1's buffer (peer-mapped): the ghost row ABOVE its last owned row; NULL on the first strip */
    double *next_ghost;            /* in strip r+1's buffer (peer-mapped): the ghost row BELOW its first owned row; NULL on the last strip  */
    unsigned long long *my_flags;  /* local device memory, [0] is written by strip r-1, [1] by strip r+1, start at 0 */
    unsigned long long *prev_flag; /* peer address of strip r-1's my_flags[1] (NULL on the first strip)              */
    unsigned long long *next_flag; /* peer address of strip r+1's my_flags[0] (NULL on the last strip)               */
    unsigned long long epoch;      /* evaluation counter, strictly increasing, same on every strip                  */
    int32_t *status_dev;           /* optional: set to 1 when a wait timed out                                       */
} ldc_peer_halo;
LDC_API int ldc_residual_peer_push(const ldc_grid *g, const ldc_peer_halo *peer, const double *X_dev,
                           void *stream);
LDC_API int ldc_residual_peer(const ldc_grid *g, const ldc_peer_halo *peer, const double *X_dev,
                      const double *U_old_dev, const double *V_old_dev, double *F_dev,
                      double *norm2_dev, double *work_dev, void *stream);
/* One evaluation in one call, overlapped on three streams.  The context owns two high-priority
 * streams: one carries the push, the other a SMALL kernel over the strip's edge tiles (the few
 * rows next to a neighbour strip; it contains the flag waits); `stream` carries the plain kernel
 * of ldc_residual over the rows in between, whose ghost rows are the strip's own edge rows.  (One
 * kernel containing the waits is ~20 % slower on every tile, and its register footprint keeps
 * the push block from becoming resident next to it; thin strips, odd nx or a fused norm fall
 * back to push + ldc_residual_peer.)  F is complete when `stream` AND the edge stream have
 * finished the evaluation: call ldc_peer_ctx_join(ctx, stream) before consuming F on `stream`
 * (bench.py does so once, in front of the event that ends a timed region).  Meant for loops over
 * rotating X buffers whose contents are final before the loop starts: the context's streams are
 * not ordered after work queued on `stream`.  Flow control is amortised over two evaluations:
 * the pushes of evaluations i, i+1 (i even) wait for this strip's newest recorded edge kernel that
 * is at least `lag` evaluations old (evaluation i-lag or i-lag-1).  That edge kernel needed the
 * neighbours' pushes of the same evaluation, which waited for their edge kernels of evaluation
 * >= i-2*lag-2, so a ghost row is never overwritten while a neighbour still reads it as long as
 * at least 2*lag+3 X buffers rotate.  A caller whose X is
 * produced on `stream` between evaluations uses ldc_residual_peer_push / ldc_residual_peer with
 * its own event. */
typedef struct ldc_peer_ctx ldc_peer_ctx;
LDC_API int ldc_peer_ctx_create(int32_t lag, ldc_peer_ctx **out);
LDC_API void ldc_peer_ctx_destroy(ldc_peer_ctx *ctx);
LDC_API int ldc_peer_ctx_join(ldc_peer_ctx *ctx, void *stream);
LDC_API int ldc_residual_peer_step(ldc_peer_ctx *ctx, const ldc_grid *g, const ldc_peer_halo *peer,
                           const double *X_dev, const double *U_old_dev, const double *V_old_dev,
                           double *F_dev, double *norm2_dev, double *work_dev, void *stream);

/* ---- packing: replaces _create_X / _recover_X (nonlinear_solver/_common.py:4-36) */
LDC_API int ldc_pack_X(const ldc_grid *g, const double *U_dev, const double *V_dev, const double *P_dev,
               double *X_dev, void *stream);
LDC_API int ldc_unpack_X(const ldc_grid *g, const double *X_dev, double *U_dev, double *V_dev,
                 double *P_dev, void *stream);

/* ---- Jacobian mask: replaces _calculate_jacobian_mask (nonlinear_solver/_common.py:39-57).
 * Canonical CSR (sorted, int32) of diags(7) (x) ones(dof,dof), wrap-around entries included. */
LDC_API int64_t ldc_mask_nnz(int32_t nx, int32_t ny, int32_t dof);
LDC_API int ldc_mask_csr(int32_t nx, int32_t ny, int32_t dof, int32_t *indptr_dev, int32_t *indices_dev,
                 void *stream);

/* ---- coloured finite-difference Jacobian: replaces SNESComputeJacobianDefaultColor +
 * MatFDColoringApply ('ds') as configured by petsc_solver_wrapper.py:132,140.  One launch
 * evaluates every colour; values_dev[nnz] is the CSR data array of the mask above
 * (per instance: batch*nnz).  Full-grid only (row_begin=0,row_count=ny) in this version. */
LDC_API int ldc_fd_jacobian(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                    const double *V_old_dev, double *values_dev, void *stream);

/* Same Jacobian in the solver's compact layout: the 26 entries per cell a row can really depend
 * on (4 continuity + 11 + 11 momentum; the other 37 mask entries are exact zeros), stored
 * compact_dev[k*N + c], k order documented in csrc/jacobian.cu.  26*N*batch doubles. */
LDC_API int ldc_fd_jacobian_compact(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                                    const double *V_old_dev, double *compact_dev, void *stream);
LDC_API int ldc_spmv_compact(const ldc_grid *g, const double *compact_dev, const double *x_dev,
                             double *y_dev, void *stream);

/* ---- SpMV y = J x on the mask pattern: replaces MatMult_SeqAIJ.
 * ldc_spmv_mask uses the closed-form structure (no index arrays are read);
 * ldc_spmv_csr is the canonical CSR product (reads indptr/indices). */
LDC_API int ldc_spmv_mask(const ldc_grid *g, const double *values_dev, const double *x_dev, double *y_dev,
                  void *stream);
LDC_API int ldc_spmv_csr(int64_t n_rows, const int32_t *indptr_dev, const int32_t *indices_dev,
                 const double *values_dev, const double *x_dev, double *y_dev, void *stream);

/* ---- transport hooks for the row-strip decomposition.  The compute library never
 * links a communication library: a distributed solver calls these two functions, implemented by
 * libldc_b200_comm.so over NCCL (include/ldc_b200_comm.h) or by anything else. */
typedef struct ldc_comm_ops {
    void *ctx;
    int32_t rank, nranks;
    /* one-row halo exchange of a [ghost | row_count owned rows | ghost] buffer, see ldc_b200_comm.h */
    int (*halo_exchange)(void *ctx, double *first_owned_row_dev, int64_t row_doubles,
                         int64_t row_count, void *stream);
    /* in-place all-reduce of device doubles; op 0 = sum, 1 = max */
    int (*allreduce)(void *ctx, double *buf_dev, int64_t count, int32_t op, void *stream);
    /* in-place all-gather: rank r's `count` doubles sit at buf_dev + r*count on entry */
    int (*allgather)(void *ctx, double *buf_dev, int64_t count, void *stream);
} ldc_comm_ops;

/* ---- device-resident Newton-Krylov solve of one pseudo-time step and the time stepper.
 * Replaces PetscSolverWrapper.solve/_setup_snes/_setup_options
 * (nonlinear_solver/petsc_solver_wrapper.py:35-145) and the loop of
 * time_stepper.run_simulation (time_stepper.py:9-56). */
typedef struct ldc_solver ldc_solver;

typedef struct ldc_solver_options {
    double snes_rtol, snes_atol, snes_stol;   /* 1e-4 each  (petsc_solver_wrapper.py:135) */
    double snes_divtol;                       /* 1e4: ||F|| > divtol*||F0|| -> reason -9 (PETSc default) */
    int32_t snes_max_it;                      /* 50 */
    int32_t snes_max_funcs;                   /* 10000 (PETSc default) */
    double ksp_rtol, ksp_atol, ksp_dtol;      /* 1e-3 (PETSc: 1e-5), 1e-50, 1e5 */
    double ksp_stagnation;                    /* 0.9: a restart cycle leaving > 90% of the residual = failed solve */
    int32_t ksp_max_it;                       /* 150 (PETSc: 10000 with ILU) */
    int32_t gmres_restart;                    /* 30 */
    int32_t pc_type;                          /* LDC_PC_* */
    int32_t pc_sweeps;                        /* relaxation sweeps before and after each coarse correction */
    int32_t mg_levels;                        /* 0 = auto (coarsen while both sizes are even) */
    int32_t mg_coarse_sweeps;                 /* relaxation sweeps on the coarsest grid */
    int32_t mg_min_size;                      /* do not coarsen below this many cells per side */
    int32_t residual_variant;                 /* as in ldc_residual */
    int32_t gmres_refine;                     /* 1: second Gram-Schmidt pass (CGS2); 0: PETSc's default (never) */
    int32_t use_cuda_graphs;                  /* 1: replay each Krylov iteration as one CUDA graph */
    int32_t pc_fp32;                          /* 1: run the multigrid cycle (a preconditioner only) in fp32 */
    int32_t reserved;
    double pc_omega_u, pc_omega_p;            /* damping of the cell-block relaxation */
} ldc_solver_options;

#define LDC_PC_NONE 0
#define LDC_PC_VANKA 1      /* additive cell-block (pressure + 4 face velocities) relaxation */
#define LDC_PC_MG_VANKA 2   /* geometric multigrid V-cycle smoothed by LDC_PC_VANKA */

LDC_API void ldc_solver_default_options(ldc_solver_options *opt);

/* result of one Newton solve, per instance; reason codes are PETSc's SNESConvergedReason as
 * tabulated in petsc_solver_wrapper.py:12-31 (2,3,4 converged; -2..-5 diverged). */
typedef struct ldc_newton_report {
    int32_t reason;                /* SNESConvergedReason codes of petsc_solver_wrapper.py:12-31, plus -9 */
    int32_t iterations;            /* Newton iterations */
    int32_t function_evaluations;  /* residual kernel evaluations actually made (1 + iterations); what
                                      snes_max_funcs is compared with.  PETSc would also count the 21
                                      coloured evaluations of each FD Jacobian, which here are ONE launch: */
    int32_t linear_iterations;     /* Krylov iterations, summed over the Newton iterations */
    int32_t ksp_reason;            /* KSPConvergedReason-like code of the last linear solve */
    int32_t jacobian_evaluations;  /* FD Jacobian launches (all colours at once) */
    double fnorm0, fnorm, xnorm, ynorm;
} ldc_newton_report;

LDC_API int ldc_solver_create(const ldc_grid *g, const ldc_solver_options *opt, ldc_solver **out);
LDC_API void ldc_solver_destroy(ldc_solver *s);
/* device pointer of the interleaved state X (3*nx*ny*batch doubles), owned by the solver */
LDC_API double *ldc_solver_state_dev(ldc_solver *s);
/* set per-instance dt (host array of `batch` doubles) */
LDC_API int ldc_solver_set_dt(ldc_solver *s, const double *dt_host);
/* set per-instance lid speed (host array of `batch` doubles) */
LDC_API int ldc_solver_set_bc(ldc_solver *s, const double *bc_host);
/* device-to-device copies of the whole interleaved state */
LDC_API int ldc_solver_set_state(ldc_solver *s, const double *X_dev, void *stream);
LDC_API int ldc_solver_get_state(ldc_solver *s, double *X_dev, void *stream);
/* Row strips: a solver created on a grid with row_begin/row_count holds that strip only (state,
 * Krylov basis, Jacobian rows); with a transport attached, halo rows of the state and of every
 * SpMV input are exchanged with the neighbouring strips and every norm / Gram-Schmidt coefficient
 * is all-reduced, so all ranks take identical decisions.  The preconditioner is block-Jacobi
 * over strips (each strip runs its own multigrid cycle).  COLLECTIVE: the call all-reduces the
 * strip height, the number of multigrid levels and the replication level and fails on EVERY rank
 * unless they are identical everywhere (equal strips: ny divisible by the number of ranks) --
 * ranks with different hierarchies would issue different sequences of collectives and hang. */
LDC_API int ldc_solver_set_comm(ldc_solver *s, const ldc_comm_ops *ops);
/* out3 = {Krylov iterations executed in lock step, sum over cavities of their own Krylov
 * iterations, Newton iterations (Jacobian builds) executed} since creation */
LDC_API int ldc_solver_counters(ldc_solver *s, int64_t *out3);
/* keep != 0: instances whose Newton solve diverged keep their last iterate (what the reference
 * returns under options.IGNORE_DIVERGED, petsc_solver_wrapper.py:90-103) instead of being rolled
 * back to the state they entered ldc_solver_step with (the default: the time stepper retries
 * from the old state, time_stepper.py:23-32).  du_inf is computed on whatever state is kept. */
LDC_API int ldc_solver_set_keep_diverged(ldc_solver *s, int32_t keep);
/* number of multigrid levels the solver built */
LDC_API int ldc_solver_num_levels(ldc_solver *s);
/* Shape of the hierarchy ldc_solver_create would build for the cavity / row strip `g`: number of
 * levels and the first level that is replicated on every rank (-1: none).  Host arithmetic only
 * (no device needed).  All ranks of a distributed solve must obtain the same answer;
 * ldc_solver_set_comm verifies that collectively and refuses the partition otherwise. */
LDC_API int ldc_solver_plan_levels(const ldc_grid *g, const ldc_solver_options *opt, int32_t *levels_out,
                           int32_t *replicated_from_out);
/* Building block exposed for the parity tests: assemble the FD Jacobian (and its multigrid
 * hierarchy) at the current state and solve J y = rhs with the preconditioned GMRES.
 * Host outputs are per instance.  Synchronises the stream. */
LDC_API int ldc_solver_linear_solve(ldc_solver *s, const double *rhs_dev, double *y_dev,
                                    int32_t *its_host, int32_t *reason_host, double *relres_host,
                                    void *stream);
/* Measurement hook (bench.py, row-strip table): average milliseconds of the solver's building
 * blocks on its own data, issued exactly as the Newton / Krylov loops issue them (halo exchanges
 * and all-reduces of a distributed solver included), CUDA events on the solver's stream:
 * ms_out4 = { residual + norm, outer SpMV w = J v, Gram-Schmidt dots against 15 basis vectors
 * (+ all-reduce), one preconditioner application }.  Rebuilds the Jacobian hierarchy at the
 * current state first.  COLLECTIVE for a distributed solver.  Synchronises the stream. */
LDC_API int ldc_solver_time_ops(ldc_solver *s, int32_t iters, double *ms_out4, void *stream);
/* One pseudo-time step for every ACTIVE instance: X_old <- X, Newton solve, X <- solution.
 * active_host[batch] (NULL = all) selects instances; diverged instances keep X = X_old.
 * reports_host[batch] is filled.  Synchronises the stream before returning.
 * du_inf_host[batch] (nullable) receives max|U - U_old| (steady-state test, time_stepper.py:39). */
LDC_API int ldc_solver_step(ldc_solver *s, const int32_t *active_host, ldc_newton_report *reports_host,
                    double *du_inf_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LDC_B200_H */
