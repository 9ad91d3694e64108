// Jacobian sparsity mask and coloured finite-difference Jacobian on the device (sm_100a).
//
// ldc_mask_csr    replaces _calculate_jacobian_mask (nonlinear_solver/_common.py:39-57): the
//                 canonical CSR of diags(ones(7), [-nx+1,-nx,-1,0,1,nx,nx-1]) (x) ones(dof,dof)
//                 from its closed form (MaskGeom), wrap-around entries included, bit-exact.
// ldc_fd_jacobian replaces PETSc's SNESComputeJacobianDefaultColor / MatFDColoringApply with
//                 mat_fd_type 'ds' (petsc_solver_wrapper.py:132,140).  PETSc needs >= 21
//                 residual sweeps (one per colour).  Here ONE launch evaluates every colour:
//                 a residual row depends on at most one unknown of each colour, so
//                 F_r(x + sum_{c in colour} h_c e_c) == F_r(x + h_c e_c) bit for bit, and each
//                 thread perturbs the stencil values of its own three rows in registers.
//                 Entries whose unknown the row never reads (dummy slots, wall-substituted
//                 neighbours, the mask's wrap-around pairs, the 37 structural zeros per cell)
//                 are exactly 0.0, which is what the column-by-column difference gives.
//                 The values land in the CSR data array of the mask (staged through shared
//                 memory so the global writes are fully coalesced).
#include "solver_internal.cuh"

namespace ldc {

// ---------------------------------------------------------------------------------------------
// mask
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_csr_kernel(int nx, int ny, int dof, int32_t *__restrict__ indptr, int32_t *__restrict__ indices)
{
    const MaskGeom m{nx, (long)nx * ny};
    const long rows = m.N * dof;
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    if (r == rows) {
        indptr[rows] = (int32_t)((long)dof * dof * m.pairs_before(m.N));
        return;
    }
    const long c = r / dof;
    const int d = (int)(r - c * dof);
    const int lo = m.lo_missing(c), cnt = m.count(c);
    const long start = (long)dof * dof * m.pairs_before(c) + (long)d * dof * cnt;
    indptr[r] = (int32_t)start;
    for (int k = 0; k < cnt; ++k) {
        const long cc = c + m.offset(lo + k);
        for (int e = 0; e < dof; ++e) indices[start + (long)k * dof + e] = (int32_t)(cc * dof + e);
    }
}

// ---------------------------------------------------------------------------------------------
// finite-difference Jacobian
// ---------------------------------------------------------------------------------------------
static constexpr int kFdCells = 64;  // cells (threads) per block

// PETSc MatFDColoring 'ds': dx = x; |dx| < umin -> +-umin; dx *= error_rel
__device__ __forceinline__ double fd_step(double x)
{
    const double umin = 1.0e-6, eps = 1.4901161193847656e-08;
    double h = x;
    if (fabs(h) < umin) h = (h >= 0.0) ? umin : -umin;
    return dmul(h, eps);
}

// (F(x + h e) - F(x)) * (1/h) with the row formula evaluated on a perturbed copy of the stencil
#define LDC_FD(ROWFN, BASE, FIELD, ACTIVE, F0, OUT)                          \
    do {                                                                     \
        OUT = 0.0;                                                           \
        if (ACTIVE) {                                                        \
            auto t__ = BASE;                                                 \
            const double h__ = fd_step(t__.FIELD);                           \
            t__.FIELD = dadd(t__.FIELD, h__);                                \
            OUT = dmul(dsub(ROWFN(t__, p), F0), ddiv(1.0, h__));             \
        }                                                                    \
    } while (0)

// Compact ("true pattern") layout used inside the solver: the 26 entries per cell a residual
// row can really depend on (4 + 11 + 11), stored SoA as Jc[k*N + c] so that
// consecutive cells are consecutive in memory.  Within each row the entries keep their CSR
// column order; entries that do not exist at a wall are 0.  The numbers are the same ones the
// CSR array holds -- the other 37 mask entries per cell are exact zeros.
//   k  0.. 3  continuity : V_s(c-nx) U_w(c-1) U_e(c) V_n(c)
//   k  4..14  x-momentum : U_S V_SW | V_SE | U_W | P_w U_P V_NW | P_e U_E V_NE | U_N
//   k 15..25  y-momentum : V_S | U_SW V_W | P_s U_SE V_P | V_E | U_NW | P_n U_NE V_N
template <bool POW2, bool CSR, bool COMPACT>
__global__ void __launch_bounds__(kFdCells)
fd_jacobian_kernel(const ldc_grid g, const double *__restrict__ X,
                   const double *__restrict__ U_old, const double *__restrict__ V_old,
                   double *__restrict__ values, long nnz, double *__restrict__ compact)
{
    __shared__ double stage[CSR ? kFdCells * 63 : 1];

    // Row strips (compact output only): cells are the strip's own, c = i + nx*jl with jl the local
    // row; X points at the first owned row with ghost rows around it (include/ldc_b200.h).
    const int inst = blockIdx.y;
    const int nx = g.nx, ny = g.ny;
    const long N = (long)nx * g.row_count;
    X += inst * 3 * N;
    U_old += inst * (long)(nx - 1) * g.row_count;
    V_old += inst * (long)nx * (ny - 1);   // batch > 1 implies the full grid
    if (CSR) values += inst * nnz;
    if (COMPACT) compact += inst * 26 * N;
    const Phys p = make_phys(g, inst);
    const GridView gv{X, nx, ny, g.row_begin};
    const MaskGeom m{nx, N};

    const long c0 = (long)blockIdx.x * kFdCells;
    const long c1 = (c0 + kFdCells < N) ? c0 + kFdCells : N;
    const long chunk_base = m.row_start(c0, 0);
    const int chunk_len = (int)(m.row_start(c1, 0) - chunk_base);  // row_start(N,0) == nnz

    if (CSR) {
        for (int t = threadIdx.x; t < chunk_len; t += kFdCells) stage[t] = 0.0;
        __syncthreads();
    }

    const long c = c0 + threadIdx.x;
    if (c < c1) {
        const int jl = (int)(c / nx), i = (int)(c - (long)jl * nx);
        const int j = g.row_begin + jl;   // grid row
        const int lo = m.lo_missing(c), cnt = m.count(c);
        double *row0 = stage + (CSR ? (m.row_start(c, 0) - chunk_base) : 0);
        double *row1 = row0 + (CSR ? 3 * cnt : 0);
        double *row2 = row1 + (CSR ? 3 * cnt : 0);
        // slot of (offset k, dof e) inside a row; only called for offsets that exist
        auto at = [&](int k, int e) { return 3 * (k - lo) + e; };
        // store one entry: CSR stage slot (row pointer + slot) and/or compact slot kc
        auto put = [&](double *row, int slot, int kc, double v) {
            if (CSR) row[slot] = v;
            if (COMPACT) compact[(long)kc * N + c] = v;
        };
        auto put0 = [&](int kc) { if (COMPACT) compact[(long)kc * N + c] = 0.0; };
        const bool W = i > 0, E = i < nx - 1, S = j > 0, Nn = j < ny - 1;

        // ---- continuity row: linear, 4 dependencies (residual_function.cpp:77-84)
        {
            const double Ue = gv.U(i, j), Uw = gv.U(i - 1, j), Vn = gv.V(i, j), Vs = gv.V(i, j - 1);
            const double f0 = cont_row(Ue, Uw, Vn, Vs, p);
            if (E) {
                const double h = fd_step(Ue);
                put(row0, at(3, 1), 2, dmul(dsub(cont_row(dadd(Ue, h), Uw, Vn, Vs, p), f0), ddiv(1.0, h)));
            } else put0(2);
            if (W) {
                const double h = fd_step(Uw);
                put(row0, at(2, 1), 1, dmul(dsub(cont_row(Ue, dadd(Uw, h), Vn, Vs, p), f0), ddiv(1.0, h)));
            } else put0(1);
            if (Nn) {
                const double h = fd_step(Vn);
                put(row0, at(3, 2), 3, dmul(dsub(cont_row(Ue, Uw, dadd(Vn, h), Vs, p), f0), ddiv(1.0, h)));
            } else put0(3);
            if (S) {
                const double h = fd_step(Vs);
                put(row0, at(0, 2), 0, dmul(dsub(cont_row(Ue, Uw, Vn, dadd(Vs, h), p), f0), ddiv(1.0, h)));
            } else put0(0);
        }

        // ---- x-momentum row of U[i,j] (or the dummy row F = X)
        if (E) {
            XmomIn s;
            gather_xmom(gv, i, j, p.bc, __ldg(U_old + i + (long)(nx - 1) * jl), s);
            const double f0 = xmom_row<POW2>(s, p);
            const bool E2 = i < nx - 2;  // U[i+1,j] is a real unknown
            double v;
            // inactive dependencies give v = 0 and have no CSR slot of their own to fill
            LDC_FD(xmom_row<POW2>, s, U_P, true, f0, v);  put(row1, at(3, 1), 9, v);
            LDC_FD(xmom_row<POW2>, s, P_w, true, f0, v);  put(row1, at(3, 0), 8, v);
            LDC_FD(xmom_row<POW2>, s, P_e, true, f0, v);  put(row1, at(4, 0), 11, v);
            LDC_FD(xmom_row<POW2>, s, U_W, W, f0, v);     if (W) put(row1, at(2, 1), 7, v); else put0(7);
            LDC_FD(xmom_row<POW2>, s, U_E, E2, f0, v);    if (E2) put(row1, at(4, 1), 12, v); else put0(12);
            LDC_FD(xmom_row<POW2>, s, U_N, Nn, f0, v);    if (Nn) put(row1, at(6, 1), 14, v); else put0(14);
            LDC_FD(xmom_row<POW2>, s, U_S, S, f0, v);     if (S) put(row1, at(0, 1), 4, v); else put0(4);
            LDC_FD(xmom_row<POW2>, s, V_NW, Nn, f0, v);   if (Nn) put(row1, at(3, 2), 10, v); else put0(10);
            LDC_FD(xmom_row<POW2>, s, V_NE, Nn, f0, v);   if (Nn) put(row1, at(4, 2), 13, v); else put0(13);
            LDC_FD(xmom_row<POW2>, s, V_SW, S, f0, v);    if (S) put(row1, at(0, 2), 5, v); else put0(5);
            LDC_FD(xmom_row<POW2>, s, V_SE, S, f0, v);    if (S) put(row1, at(1, 2), 6, v); else put0(6);
        } else {
            const double x = X[3 * c + 1];
            const double h = fd_step(x);
            for (int kc = 4; kc <= 14; ++kc) put0(kc);
            put(row1, at(3, 1), 9, dmul(dsub(dadd(x, h), x), ddiv(1.0, h)));
        }

        // ---- y-momentum row of V[i,j] (or the dummy row)
        if (Nn) {
            YmomIn s;
            gather_ymom(gv, i, j, p.bc, __ldg(V_old + c), s);
            const double f0 = ymom_row<POW2>(s, p);
            const bool N2 = j < ny - 2;  // V[i,j+1] / U[.,j+1] are read (not the lid quirk)
            double v;
            LDC_FD(ymom_row<POW2>, s, V_P, true, f0, v);      put(row2, at(3, 2), 20, v);
            LDC_FD(ymom_row<POW2>, s, P_s, true, f0, v);      put(row2, at(3, 0), 18, v);
            LDC_FD(ymom_row<POW2>, s, P_n, true, f0, v);      put(row2, at(6, 0), 23, v);
            LDC_FD(ymom_row<POW2>, s, V_E, E, f0, v);         if (E) put(row2, at(4, 2), 21, v); else put0(21);
            LDC_FD(ymom_row<POW2>, s, V_W, W, f0, v);         if (W) put(row2, at(2, 2), 17, v); else put0(17);
            LDC_FD(ymom_row<POW2>, s, V_N, N2, f0, v);        if (N2) put(row2, at(6, 2), 25, v); else put0(25);
            LDC_FD(ymom_row<POW2>, s, V_S, S, f0, v);         if (S) put(row2, at(0, 2), 15, v); else put0(15);
            LDC_FD(ymom_row<POW2>, s, U_SE, E, f0, v);        if (E) put(row2, at(3, 1), 19, v); else put0(19);
            LDC_FD(ymom_row<POW2>, s, U_SW, W, f0, v);        if (W) put(row2, at(2, 1), 16, v); else put0(16);
            LDC_FD(ymom_row<POW2>, s, U_NE, E && N2, f0, v);  if (E && N2) put(row2, at(6, 1), 24, v); else put0(24);
            LDC_FD(ymom_row<POW2>, s, U_NW, W && N2, f0, v);  if (W && N2) put(row2, at(5, 1), 22, v); else put0(22);
        } else {
            const double x = X[3 * c + 2];
            const double h = fd_step(x);
            for (int kc = 15; kc <= 25; ++kc) put0(kc);
            put(row2, at(3, 2), 20, dmul(dsub(dadd(x, h), x), ddiv(1.0, h)));
        }
    }
    if (CSR) {
        __syncthreads();
        // the block's CSR values are one contiguous global range: coalesced copy-out
        double *out = values + chunk_base;
        for (int t = threadIdx.x; t < chunk_len; t += kFdCells) out[t] = stage[t];
    }
}

int fd_jacobian_launch(const ldc_grid *g, const double *X, const double *U_old, const double *V_old,
                       double *csr_values, double *compact, cudaStream_t st)
{
    const long N = (long)g->nx * g->row_count;
    const long nnz = ldc_mask_nnz(g->nx, g->ny, 3);
    dim3 grid((unsigned)((N + kFdCells - 1) / kFdCells), g->batch);
    const bool full = g->row_begin == 0 && g->row_count == g->ny;
    LDC_REQUIRE(full || csr_values == nullptr, "the CSR(mask) output needs the full grid; strips use the compact layout");
    const bool pow2 = is_pow2_double(g->dx) && is_pow2_double(g->dy);
#define LDC_FD_LAUNCH(P2, C1, C2) \
    fd_jacobian_kernel<P2, C1, C2><<<grid, kFdCells, 0, st>>>(*g, X, U_old, V_old, csr_values, nnz, compact)
    if (csr_values && compact) { if (pow2) LDC_FD_LAUNCH(true, true, true); else LDC_FD_LAUNCH(false, true, true); }
    else if (csr_values) { if (pow2) LDC_FD_LAUNCH(true, true, false); else LDC_FD_LAUNCH(false, true, false); }
    else { if (pow2) LDC_FD_LAUNCH(true, false, true); else LDC_FD_LAUNCH(false, false, true); }
#undef LDC_FD_LAUNCH
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

}  // namespace ldc

using namespace ldc;

extern "C" int64_t ldc_mask_nnz(int32_t nx, int32_t ny, int32_t dof)
{
    if (nx < 3 || ny < 1 || dof < 1) return -1;
    const MaskGeom m{nx, (long)nx * ny};
    return (int64_t)dof * dof * m.pairs_before(m.N);
}

extern "C" int ldc_mask_csr(int32_t nx, int32_t ny, int32_t dof, int32_t *indptr_dev,
                            int32_t *indices_dev, void *stream)
{
    LDC_REQUIRE(nx >= 3 && ny >= 1 && dof >= 1, "ldc_mask_csr: need nx>=3, ny>=1, dof>=1");
    LDC_REQUIRE(indptr_dev && indices_dev, "ldc_mask_csr: null device pointer");
    const int64_t nnz = ldc_mask_nnz(nx, ny, dof);
    LDC_REQUIRE(nnz < 2147483647LL, "ldc_mask_csr: nnz=%lld does not fit the int32 CSR of the reference",
                (long long)nnz);
    const long rows = (long)nx * ny * dof + 1;
    mask_csr_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        nx, ny, dof, indptr_dev, indices_dev);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

extern "C" int ldc_fd_jacobian(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                               const double *V_old_dev, double *values_dev, void *stream)
{
    if (int rc = validate_grid(g, true)) return rc;
    LDC_REQUIRE(X_dev && U_old_dev && V_old_dev && values_dev, "ldc_fd_jacobian: null device pointer");
    return fd_jacobian_launch(g, X_dev, U_old_dev, V_old_dev, values_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int ldc_fd_jacobian_compact(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                                       const double *V_old_dev, double *compact_dev, void *stream)
{
    if (int rc = validate_grid(g, false)) return rc;
    LDC_REQUIRE(X_dev && U_old_dev && V_old_dev && compact_dev, "ldc_fd_jacobian_compact: null device pointer");
    return fd_jacobian_launch(g, X_dev, U_old_dev, V_old_dev, nullptr, compact_dev, (cudaStream_t)stream);
}
