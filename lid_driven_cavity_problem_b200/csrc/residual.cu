// Residual F(X) of the coupled staggered-grid finite-volume lid-driven cavity (sm_100a).
//
// Replaces residual_function(X, graph) of the reference: c++/residual_function.cpp:10-238,
// c++/residual_function_omp.cpp, cython_/residual_function.pyx:16-331 and the numba / numpy /
// pure-Python siblings.  Arithmetic is evaluated in the reference's order with exactly-rounded
// fp64 operations (no FMA contraction), so results are bit-identical to the x86-64 -O2 build.
//
// Two kernels:
//   residual_cell_kernel  : one thread per cell, direct transcription of the three row formulas
//                           (shares them with the FD-Jacobian kernel).
//   residual_march_kernel : flux form.  The scheme is conservative: the advective and diffusive
//                           flux through a face is the SAME expression for the two cells that
//                           share it (the reference evaluates it twice).  Each lane owns one
//                           grid column and marches up a tile of rows; it evaluates only the
//                           east and north faces of its cell, receives the west face from the
//                           neighbouring lane by warp shuffle and carries the south face in
//                           registers.  Each X value is loaded once per tile (+1 halo row), the
//                           4 divisions + ~80 flops per cell are half of the cell kernel's, and
//                           sum(F^2) is reduced by warp shuffles for the Newton norm.
#include "ldc_internal.cuh"

namespace ldc {

static constexpr int kMarchWarps = 4;                 // warps per block
static constexpr int kMarchCols = 30;                 // output columns per warp (lanes 1..30)

struct Strides {
    long x, u, v;  // per-instance strides of X/F, U_old, V_old
};

__host__ __device__ inline Strides instance_strides(const ldc_grid &g)
{
    Strides s;
    s.x = 3L * g.nx * g.row_count;
    s.u = (long)(g.nx - 1) * g.row_count;
    long v_rows = (g.row_begin + g.row_count < g.ny - 1 ? g.row_begin + g.row_count : g.ny - 1) -
                  g.row_begin;
    if (v_rows < 0) v_rows = 0;
    s.v = (long)g.nx * v_rows;
    return s;
}

// ---------------------------------------------------------------------------------------------
// variant 1: one thread per cell
// ---------------------------------------------------------------------------------------------
template <bool POW2>
__global__ void __launch_bounds__(256)
residual_cell_kernel(const ldc_grid g, const double *__restrict__ X,
                     const double *__restrict__ U_old, const double *__restrict__ V_old,
                     double *__restrict__ F, double *__restrict__ partials)
{
    const int inst = blockIdx.z;
    const Strides st = instance_strides(g);
    X += inst * st.x;
    F += inst * st.x;
    U_old += inst * st.u;
    V_old += inst * st.v;
    const Phys p = make_phys(g, inst);
    const GridView gv{X, g.nx, g.ny, g.row_begin};

    const long cells = (long)g.nx * g.row_count;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0;
    if (c < cells) {
        const int jl = (int)(c / g.nx);
        const int i = (int)(c - (long)jl * g.nx);
        const int j = g.row_begin + jl;
        const double f0 = cont_row(gv.U(i, j), gv.U(i - 1, j), gv.V(i, j), gv.V(i, j - 1), p);
        double f1, f2;
        if (i < g.nx - 1) {
            XmomIn s;
            gather_xmom(gv, i, j, p.bc, __ldg(U_old + i + (long)(g.nx - 1) * jl), s);
            f1 = xmom_row<POW2>(s, p);
        } else {
            f1 = X[3 * c + 1];
        }
        if (j < g.ny - 1) {
            YmomIn s;
            gather_ymom(gv, i, j, p.bc, __ldg(V_old + c), s);
            f2 = ymom_row<POW2>(s, p);
        } else {
            f2 = X[3 * c + 2];
        }
        F[3 * c] = f0;
        F[3 * c + 1] = f1;
        F[3 * c + 2] = f2;
        sq = f0 * f0 + f1 * f1 + f2 * f2;
    }
    if (partials) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        __shared__ double warp_sums[8];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) warp_sums[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_sums[w];
            partials[(long)inst * gridDim.x + blockIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// variant 2: flux form, column marching
// ---------------------------------------------------------------------------------------------
struct CellVals {
    double P, U, V;    // raw slots
};

__device__ __forceinline__ CellVals load_cell(const double *__restrict__ X, int i, int jl, int nx,
                                              bool row_ok)
{
    CellVals c;
    if (row_ok && i >= 0 && i < nx) {
        const double *q = X + 3L * ((long)i + (long)nx * jl);
        c.P = __ldg(q);
        c.U = __ldg(q + 1);
        c.V = __ldg(q + 2);
    } else {
        c.P = 0.0;
        c.U = 0.0;
        c.V = 0.0;
    }
    return c;
}

__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

template <bool POW2>
__global__ void __launch_bounds__(kMarchWarps * 32)
residual_march_kernel(const ldc_grid g, const double *__restrict__ X,
                      const double *__restrict__ U_old, const double *__restrict__ V_old,
                      double *__restrict__ F, double *__restrict__ partials, int rows_per_tile)
{
    const int inst = blockIdx.z;
    const Strides st = instance_strides(g);
    X += inst * st.x;
    F += inst * st.x;
    U_old += inst * st.u;
    V_old += inst * st.v;
    const Phys p = make_phys(g, inst);

    const int nx = g.nx, ny = g.ny;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kMarchWarps + warp) * kMarchCols - 1 + lane;  // grid column
    const int jl0 = blockIdx.y * rows_per_tile;
    const int jl1 = min(jl0 + rows_per_tile, g.row_count);
    const bool writer = lane >= 1 && lane <= kMarchCols && i < nx;  // i >= 0 for lane >= 1
    const bool u_col = i >= 0 && i < nx - 1;   // column holds a real U
    const bool side_wall = i < 0 || i >= nx - 1;  // U_NE of the V rows is forced to 0 here

    // a local row jl is readable if it is an owned row or an existing ghost row
    auto row_ok = [&](int jl) {
        const int jg = g.row_begin + jl;
        return jg >= 0 && jg < ny && jl >= -1 && jl <= g.row_count;
    };
    // masked (wall-aware) views of a raw cell at grid row jg
    auto Um = [&](const CellVals &c) { return u_col ? c.U : 0.0; };
    auto Vm = [&](const CellVals &c, int jg) { return (jg < ny - 1) ? c.V : 0.0; };

    // ---- prologue: south faces of the first row = north faces of row jl0-1
    CellVals S = load_cell(X, i, jl0 - 1, nx, row_ok(jl0 - 1));
    CellVals C = load_cell(X, i, jl0, nx, row_ok(jl0));
    CellVals N = load_cell(X, i, jl0 + 1, nx, row_ok(jl0 + 1));
    double AU_s, DU_s, AV_s, DV_s, CV_s;
    {
        const int jg = g.row_begin + jl0 - 1;  // may be -1: everything below the floor is 0
        const double Sm_U = Um(S), Sm_V = (jg >= 0) ? Vm(S, jg) : 0.0;
        const double Cm_U = Um(C), Cm_V = Vm(C, jg + 1);
        const double Sm_Ve = shfl_down1(Sm_V);
        // U rows: north face of (i, jg)   [U_N is never the lid here: jg < ny-1]
        const double dU_n = over_dy<POW2>(dsub(Cm_U, Sm_U), p);
        const double vn = dmul(dadd(Sm_Ve, Sm_V), 0.5);
        AU_s = adv_flux(vn, Cm_U, Sm_U, p.dx, p);
        DU_s = dmul(dmul(p.mi, dU_n), p.dx);
        // V rows
        const double dV_n = over_dy<POW2>(dsub(Cm_V, Sm_V), p);
        const double vvn = dmul(dadd(Sm_V, Cm_V), 0.5);
        AV_s = adv_flux(vvn, Cm_V, Sm_V, p.dx, p);
        DV_s = dmul(dmul(p.mi, dV_n), p.dx);
        CV_s = dmul(Sm_V, p.dx);
    }

    double sq = 0.0;
    for (int jl = jl0; jl < jl1; ++jl) {
        const int jg = g.row_begin + jl;
        // prefetch row jl+2 and this row's old values before the arithmetic
        const CellVals NN = load_cell(X, i, jl + 2, nx, (jl + 2 <= jl1) && row_ok(jl + 2));
        const double Uold = (u_col) ? __ldg(U_old + i + (long)(nx - 1) * jl) : 0.0;
        const double Vold = (i >= 0 && i < nx && jg < ny - 1) ? __ldg(V_old + i + (long)nx * jl) : 0.0;

        const double Cm_U = Um(C), Cm_V = Vm(C, jg);
        const double Nm_U = Um(N), Nm_V = Vm(N, jg + 1);
        const double Cm_Ue = shfl_down1(Cm_U);
        const double Cm_Ve = shfl_down1(Cm_V);
        const double C_Pe = shfl_down1(C.P);

        // ---- north face of (i, jg)
        const double U_N = (jg == ny - 1) ? p.bc : Nm_U;
        const double dU_n = over_dy<POW2>(dsub(U_N, Cm_U), p);
        const double vn = dmul(dadd(Cm_Ve, Cm_V), 0.5);
        const double AU_n = adv_flux(vn, U_N, Cm_U, p.dx, p);
        const double DU_n = dmul(dmul(p.mi, dU_n), p.dx);
        const double dV_n = over_dy<POW2>(dsub(Nm_V, Cm_V), p);
        const double vvn = dmul(dadd(Cm_V, Nm_V), 0.5);
        const double AV_n = adv_flux(vvn, Nm_V, Cm_V, p.dx, p);
        const double DV_n = dmul(dmul(p.mi, dV_n), p.dx);
        const double CV_n = dmul(Cm_V, p.dx);

        // ---- east face of (i, jg)
        const double dU_e = over_dx<POW2>(dsub(Cm_Ue, Cm_U), p);
        const double ue = dmul(dadd(Cm_Ue, Cm_U), 0.5);
        const double AU_e = adv_flux(ue, Cm_Ue, Cm_U, p.dy, p);
        const double DU_e = dmul(dmul(p.mi, dU_e), p.dy);
        const double U_NE = side_wall ? 0.0 : ((jg == ny - 2) ? p.bc : Nm_U);
        const double uue = dmul(dadd(U_NE, Cm_U), 0.5);
        const double dV_e = over_dx<POW2>(dsub(Cm_Ve, Cm_V), p);
        const double AV_e = adv_flux(uue, Cm_Ve, Cm_V, p.dy, p);
        const double DV_e = dmul(dmul(p.mi, dV_e), p.dy);
        const double CU_e = dmul(Cm_U, p.dy);

        // ---- west face = east face of the lane to the left
        const double AU_w = shfl_up1(AU_e);
        const double DU_w = shfl_up1(DU_e);
        const double AV_w = shfl_up1(AV_e);
        const double DV_w = shfl_up1(DV_e);
        const double CU_w = shfl_up1(CU_e);

        if (writer) {
            const double f0 = dadd(dsub(CU_e, CU_w), dsub(CV_n, CV_s));
            double f1, f2;
            if (u_col) {
                const double tr = dmul(dsub(dmul(p.rho, Cm_U), dmul(p.rho, Uold)), p.vol_dt);
                const double adv = dsub(dadd(dsub(AU_e, AU_w), AU_n), AU_s);
                const double dif = dsub(dadd(dsub(DU_e, DU_w), DU_n), DU_s);
                const double src = dmul(-dsub(C_Pe, C.P), p.dy);
                f1 = dsub(dsub(dadd(tr, adv), dif), src);
            } else {
                f1 = C.U;  // dummy row: F = X
            }
            if (jg < ny - 1) {
                const double tr = dmul(dsub(dmul(p.rho, Cm_V), dmul(p.rho, Vold)), p.vol_dt);
                const double adv = dsub(dadd(dsub(AV_e, AV_w), AV_n), AV_s);
                const double dif = dsub(dadd(dsub(DV_e, DV_w), DV_n), DV_s);
                const double src = dmul(-dsub(N.P, C.P), p.dx);
                f2 = dsub(dsub(dadd(tr, adv), dif), src);
            } else {
                f2 = C.V;
            }
            double *out = F + 3L * ((long)i + (long)nx * jl);
            out[0] = f0;
            out[1] = f1;
            out[2] = f2;
            sq += f0 * f0 + f1 * f1 + f2 * f2;
        }
        AU_s = AU_n; DU_s = DU_n; AV_s = AV_n; DV_s = DV_n; CV_s = CV_n;
        C = N;
        N = NN;
    }

    if (partials) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        __shared__ double warp_sums[kMarchWarps];
        if (lane == 0) warp_sums[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kMarchWarps; ++w) t += warp_sums[w];
            const long nblk = (long)gridDim.x * gridDim.y;
            partials[(long)inst * nblk + (long)blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// one block per instance: fixed-order sum of the per-block partials (deterministic)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const double *__restrict__ partials, long per_instance, double *__restrict__ out)
{
    const double *q = partials + (long)blockIdx.x * per_instance;
    double s = 0.0;
    for (long k = threadIdx.x; k < per_instance; k += blockDim.x) s += q[k];
    __shared__ double sm[256];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

struct MarchPlan {
    dim3 grid;
    int rows_per_tile;
};

static MarchPlan plan_march(const ldc_grid &g)
{
    const int col_warps = (g.nx + kMarchCols - 1) / kMarchCols;
    const int col_blocks = (col_warps + kMarchWarps - 1) / kMarchWarps;
    // aim at >= 2 blocks of 128 threads per resident slot (148 SMs x 4 blocks) so the tail is
    // short, while keeping tiles tall enough to amortise the halo row.
    const long target_blocks = 2L * 4 * sm_count();
    long row_tiles = (target_blocks + (long)col_blocks * g.batch - 1) / ((long)col_blocks * g.batch);
    if (row_tiles < 1) row_tiles = 1;
    int rows = (int)((g.row_count + row_tiles - 1) / row_tiles);
    if (rows < 8) rows = 8;
    if (rows > 64) rows = 64;
    if (rows > g.row_count) rows = g.row_count;
    MarchPlan pl;
    pl.rows_per_tile = rows;
    pl.grid = dim3(col_blocks, (g.row_count + rows - 1) / rows, g.batch);
    return pl;
}

static long cell_blocks(const ldc_grid &g) { return ((long)g.nx * g.row_count + 255) / 256; }

}  // namespace ldc

using namespace ldc;

extern "C" int64_t ldc_residual_work_doubles(const ldc_grid *g)
{
    if (!g || validate_grid(g, false) != 0) return -1;
    const MarchPlan pl = plan_march(*g);
    const long a = (long)pl.grid.x * pl.grid.y;
    const long b = cell_blocks(*g);
    return (a > b ? a : b) * g->batch;
}

extern "C" int ldc_residual(const ldc_grid *g, const double *X_dev, const double *U_old_dev,
                            const double *V_old_dev, double *F_dev, double *norm2_dev,
                            double *work_dev, int32_t variant, void *stream)
{
    if (int rc = validate_grid(g, false)) return rc;
    LDC_REQUIRE(X_dev && U_old_dev && V_old_dev && F_dev, "ldc_residual: null device pointer");
    LDC_REQUIRE(X_dev != F_dev, "ldc_residual: F must not alias X");
    LDC_REQUIRE(!norm2_dev || work_dev, "ldc_residual: norm2 requested without work buffer");
    LDC_REQUIRE(variant >= 0 && variant <= 2, "ldc_residual: unknown variant %d", variant);
    cudaStream_t s = (cudaStream_t)stream;
    const bool pow2 = is_pow2_double(g->dx) && is_pow2_double(g->dy);
    double *partials = norm2_dev ? work_dev : nullptr;
    long per_instance;
    if (variant == 0) variant = 2;
    if (variant == 1) {
        dim3 grid((unsigned)cell_blocks(*g), 1, g->batch);
        per_instance = grid.x;
        if (pow2)
            residual_cell_kernel<true><<<grid, 256, 0, s>>>(*g, X_dev, U_old_dev, V_old_dev, F_dev, partials);
        else
            residual_cell_kernel<false><<<grid, 256, 0, s>>>(*g, X_dev, U_old_dev, V_old_dev, F_dev, partials);
    } else {
        const MarchPlan pl = plan_march(*g);
        per_instance = (long)pl.grid.x * pl.grid.y;
        if (pow2)
            residual_march_kernel<true><<<pl.grid, kMarchWarps * 32, 0, s>>>(
                *g, X_dev, U_old_dev, V_old_dev, F_dev, partials, pl.rows_per_tile);
        else
            residual_march_kernel<false><<<pl.grid, kMarchWarps * 32, 0, s>>>(
                *g, X_dev, U_old_dev, V_old_dev, F_dev, partials, pl.rows_per_tile);
    }
    LDC_CHECK_LAUNCH();
    if (norm2_dev) {
        sum_partials_kernel<<<g->batch, 256, 0, s>>>(partials, per_instance, norm2_dev);
        LDC_CHECK_LAUNCH();
    }
    return LDC_OK;
}
