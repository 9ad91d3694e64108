// Internal helpers shared by the sm_100a kernels of libldc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ldc_b200.h"

namespace ldc {

// ---------------------------------------------------------------------------------------------
// error plumbing (no exception crosses the C-ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define LDC_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ldc::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define LDC_CHECK_LAUNCH() LDC_CUDA(cudaGetLastError())

#define LDC_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ldc::set_error(__VA_ARGS__);      \
            return LDC_ERR_INVALID;           \
        }                                     \
    } while (0)

int validate_grid(const ldc_grid *g, bool need_full_grid);
int sm_count();

// ---------------------------------------------------------------------------------------------
// Exactly-rounded fp64 arithmetic.  The reference residual is compiled for baseline x86-64
// (no FMA), so parity-critical kernels spell every operation with the *_rn intrinsics, which
// nvcc never contracts into DFMA regardless of -fmad.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// Physical constants of one cavity, prepared once per thread block.
struct Phys {
    double dx, dy, inv_dx, inv_dy;  // inv_* only used when both spacings are powers of two
    double rho, mi, bc;
    double vol_dt;                  // dx*dy/dt evaluated as (dx*dy)/dt like the reference
    int use_cds;
};

__device__ __forceinline__ Phys make_phys(const ldc_grid &g, int instance)
{
    Phys p;
    p.dx = g.dx;
    p.dy = g.dy;
    p.inv_dx = ddiv(1.0, g.dx);
    p.inv_dy = ddiv(1.0, g.dy);
    p.rho = g.rho;
    p.mi = g.mi;
    p.bc = g.bc_dev ? g.bc_dev[instance] : g.bc;
    const double dt = g.dt_dev ? g.dt_dev[instance] : g.dt;
    p.vol_dt = ddiv(dmul(g.dx, g.dy), dt);
    p.use_cds = g.use_cds;
    return p;
}

// a/dx.  When dx is a power of two, a*(1/dx) is the same correctly rounded value and costs one
// multiply instead of a ~10-instruction IEEE division (POW2 is chosen on the host).
template <bool POW2>
__device__ __forceinline__ double over_dx(double a, const Phys &p)
{
    return POW2 ? dmul(a, p.inv_dx) : ddiv(a, p.dx);
}
template <bool POW2>
__device__ __forceinline__ double over_dy(double a, const Phys &p)
{
    return POW2 ? dmul(a, p.inv_dy) : ddiv(a, p.dy);
}

// Upwind weights (.5-beta, .5+beta) of residual_function.cpp:134-137 (strict '>'), or (.5,.5)
// for central differencing (numba_residual_function.py:128-137).  beta=+-0.5 makes both
// weights exactly 0 or 1, so selecting them is identical to computing .5-+beta.
__device__ __forceinline__ void upwind_weights(double face_velocity, int use_cds, double &w_far,
                                               double &w_near)
{
    if (use_cds) {
        w_far = 0.5;
        w_near = 0.5;
    } else {
        const bool pos = face_velocity > 0.0;
        w_far = pos ? 0.0 : 1.0;   // .5 - beta
        w_near = pos ? 1.0 : 0.0;  // .5 + beta
    }
}

// Advective flux through one face:  rho*vel*((.5-beta)*phi_far + (.5+beta)*phi_near)*len,
// evaluated left to right exactly as written in residual_function.cpp:142-146.
// "near" is the P-side value for east/north faces (U_P) -- see the call sites.
__device__ __forceinline__ double adv_flux(double vel, double phi_a, double phi_b, double len,
                                           const Phys &p)
{
    double wa, wb;
    upwind_weights(vel, p.use_cds, wa, wb);
    const double interp = dadd(dmul(wa, phi_a), dmul(wb, phi_b));
    return dmul(dmul(dmul(p.rho, vel), interp), len);
}

// ---------------------------------------------------------------------------------------------
// Row formulas on explicit stencil values (used by the cell-per-thread residual kernel and by
// the finite-difference Jacobian kernel, which perturbs one argument at a time).
// ---------------------------------------------------------------------------------------------
struct XmomIn {  // residual_function.cpp:113-124
    double U_P, U_W, U_E, U_N, U_S, P_w, P_e, V_NE, V_NW, V_SE, V_SW, U_P_old;
};
struct YmomIn {  // residual_function.cpp:183-194
    double V_P, V_E, V_W, V_N, V_S, P_n, P_s, U_SE, U_SW, U_NE, U_NW, V_P_old;
};

template <bool POW2>
__device__ __forceinline__ double xmom_row(const XmomIn &s, const Phys &p)
{
    const double dU_e = over_dx<POW2>(dsub(s.U_E, s.U_P), p);
    const double dU_w = over_dx<POW2>(dsub(s.U_P, s.U_W), p);
    const double dU_n = over_dy<POW2>(dsub(s.U_N, s.U_P), p);
    const double dU_s = over_dy<POW2>(dsub(s.U_P, s.U_S), p);
    const double U_e = dmul(dadd(s.U_E, s.U_P), 0.5);
    const double U_w = dmul(dadd(s.U_P, s.U_W), 0.5);
    const double V_n = dmul(dadd(s.V_NE, s.V_NW), 0.5);
    const double V_s = dmul(dadd(s.V_SE, s.V_SW), 0.5);
    const double transient = dmul(dsub(dmul(p.rho, s.U_P), dmul(p.rho, s.U_P_old)), p.vol_dt);
    double adv = adv_flux(U_e, s.U_E, s.U_P, p.dy, p);
    adv = dsub(adv, adv_flux(U_w, s.U_P, s.U_W, p.dy, p));
    adv = dadd(adv, adv_flux(V_n, s.U_N, s.U_P, p.dx, p));
    adv = dsub(adv, adv_flux(V_s, s.U_P, s.U_S, p.dx, p));
    double dif = dmul(dmul(p.mi, dU_e), p.dy);
    dif = dsub(dif, dmul(dmul(p.mi, dU_w), p.dy));
    dif = dadd(dif, dmul(dmul(p.mi, dU_n), p.dx));
    dif = dsub(dif, dmul(dmul(p.mi, dU_s), p.dx));
    const double src = dmul(-dsub(s.P_e, s.P_w), p.dy);
    return dsub(dsub(dadd(transient, adv), dif), src);
}

template <bool POW2>
__device__ __forceinline__ double ymom_row(const YmomIn &s, const Phys &p)
{
    const double dV_e = over_dx<POW2>(dsub(s.V_E, s.V_P), p);
    const double dV_w = over_dx<POW2>(dsub(s.V_P, s.V_W), p);
    const double dV_n = over_dy<POW2>(dsub(s.V_N, s.V_P), p);
    const double dV_s = over_dy<POW2>(dsub(s.V_P, s.V_S), p);
    const double U_e = dmul(dadd(s.U_NE, s.U_SE), 0.5);
    const double U_w = dmul(dadd(s.U_NW, s.U_SW), 0.5);
    const double V_n = dmul(dadd(s.V_P, s.V_N), 0.5);
    const double V_s = dmul(dadd(s.V_S, s.V_P), 0.5);
    const double transient = dmul(dsub(dmul(p.rho, s.V_P), dmul(p.rho, s.V_P_old)), p.vol_dt);
    double adv = adv_flux(U_e, s.V_E, s.V_P, p.dy, p);
    adv = dsub(adv, adv_flux(U_w, s.V_P, s.V_W, p.dy, p));
    adv = dadd(adv, adv_flux(V_n, s.V_N, s.V_P, p.dx, p));
    adv = dsub(adv, adv_flux(V_s, s.V_P, s.V_S, p.dx, p));
    double dif = dmul(dmul(p.mi, dV_e), p.dy);
    dif = dsub(dif, dmul(dmul(p.mi, dV_w), p.dy));
    dif = dadd(dif, dmul(dmul(p.mi, dV_n), p.dx));
    dif = dsub(dif, dmul(dmul(p.mi, dV_s), p.dx));
    const double src = dmul(-dsub(s.P_n, s.P_s), p.dx);
    return dsub(dsub(dadd(transient, adv), dif), src);
}

// continuity, residual_function.cpp:77-84
__device__ __forceinline__ double cont_row(double U_e, double U_w, double V_n, double V_s,
                                           const Phys &p)
{
    return dadd(dsub(dmul(U_e, p.dy), dmul(U_w, p.dy)), dsub(dmul(V_n, p.dx), dmul(V_s, p.dx)));
}

// ---------------------------------------------------------------------------------------------
// Grid view: masked fetches from the interleaved vector of ONE instance / strip.
// X points at the first owned row; jl is the local row, jg = row_begin + jl the grid row.
// Out-of-grid neighbours and dummy slots read as 0 (walls), the lid reads as bc where the
// reference does so.  Ghost rows (jl = -1, jl = row_count) are read only when they exist in
// the grid.
// ---------------------------------------------------------------------------------------------
struct GridView {
    const double *X;
    int nx, ny, row_begin;
    __device__ __forceinline__ bool inside(int i, int jg) const
    {
        return i >= 0 && i < nx && jg >= 0 && jg < ny;
    }
    __device__ __forceinline__ long slot(int i, int jg) const
    {
        return 3L * ((long)i + (long)nx * (long)(jg - row_begin));
    }
    __device__ __forceinline__ double P(int i, int jg) const
    {
        return inside(i, jg) ? __ldg(X + slot(i, jg)) : 0.0;
    }
    // U on the east face of cell (i,jg); the slot of column nx-1 is a dummy -> wall value 0
    __device__ __forceinline__ double U(int i, int jg) const
    {
        return (inside(i, jg) && i < nx - 1) ? __ldg(X + slot(i, jg) + 1) : 0.0;
    }
    __device__ __forceinline__ double V(int i, int jg) const
    {
        return (inside(i, jg) && jg < ny - 1) ? __ldg(X + slot(i, jg) + 2) : 0.0;
    }
};

__device__ __forceinline__ void gather_xmom(const GridView &g, int i, int j, double bc,
                                            double U_old, XmomIn &s)
{
    s.U_P = g.U(i, j);
    s.U_W = g.U(i - 1, j);
    s.U_E = g.U(i + 1, j);
    s.U_N = (j == g.ny - 1) ? bc : g.U(i, j + 1);
    s.U_S = g.U(i, j - 1);
    s.P_w = g.P(i, j);
    s.P_e = g.P(i + 1, j);
    s.V_NE = g.V(i + 1, j);
    s.V_NW = g.V(i, j);
    s.V_SE = g.V(i + 1, j - 1);
    s.V_SW = g.V(i, j - 1);
    s.U_P_old = U_old;
}

__device__ __forceinline__ void gather_ymom(const GridView &g, int i, int j, double bc,
                                            double V_old, YmomIn &s)
{
    const bool top = (j == g.ny - 2);
    s.V_P = g.V(i, j);
    s.V_E = g.V(i + 1, j);
    s.V_W = g.V(i - 1, j);
    s.V_N = g.V(i, j + 1);
    s.V_S = g.V(i, j - 1);
    s.P_n = g.P(i, j + 1);
    s.P_s = g.P(i, j);
    s.U_SE = g.U(i, j);
    s.U_SW = g.U(i - 1, j);
    // quirk of residual_function.cpp:192-193: the lid speed replaces U[.,ny-1] on the top V row,
    // but the side walls win over the lid.
    s.U_NE = (i == g.nx - 1) ? 0.0 : (top ? bc : g.U(i, j + 1));
    s.U_NW = (i == 0) ? 0.0 : (top ? bc : g.U(i - 1, j + 1));
    s.V_P_old = V_old;
}

__host__ __device__ __forceinline__ bool is_pow2_double(double v)
{
    // exactly one significand bit set (normal numbers): v = 2^k
    union { double d; unsigned long long u; } c;
    c.d = v;
    return v > 0.0 && (c.u & 0x000FFFFFFFFFFFFFULL) == 0ULL && ((c.u >> 52) & 0x7FF) != 0 &&
           ((c.u >> 52) & 0x7FF) != 0x7FF;
}

// ---------------------------------------------------------------------------------------------
// Closed form of the mask CSR.  Offsets ascending:
//   o0=-nx, o1=-nx+1, o2=-1, o3=0, o4=+1, o5=nx-1, o6=nx ; cell c has the offsets with
//   0 <= c+o < N.  All rows of one cell have the same length 3*cnt(c).
// ---------------------------------------------------------------------------------------------
struct MaskGeom {
    long nx, N;
    __host__ __device__ __forceinline__ int lo_missing(long c) const
    {   // offsets -nx, -nx+1, -1 that fall below 0 (always a prefix of the sorted list)
        return (c < nx ? 1 : 0) + (c < nx - 1 ? 1 : 0) + (c < 1 ? 1 : 0);
    }
    __host__ __device__ __forceinline__ int hi_missing(long c) const
    {   // offsets +1, nx-1, nx that reach N (always a suffix)
        return (c + 1 >= N ? 1 : 0) + (c + nx - 1 >= N ? 1 : 0) + (c + nx >= N ? 1 : 0);
    }
    __host__ __device__ __forceinline__ int count(long c) const
    {
        return 7 - lo_missing(c) - hi_missing(c);
    }
    // number of (cell,offset) pairs of all cells before c
    __host__ __device__ __forceinline__ long pairs_before(long c) const
    {
        auto mn = [](long a, long b) { return a < b ? a : b; };
        auto pos = [](long a) { return a > 0 ? a : 0L; };
        const long lo = mn(c, nx) + mn(c, nx - 1) + mn(c, 1);
        const long hi = pos(c - (N - 1)) + pos(c - (N - nx + 1)) + pos(c - (N - nx));
        return 7 * c - lo - hi;
    }
    __host__ __device__ __forceinline__ long offset(int k) const
    {
        switch (k) {
            case 0: return -nx;
            case 1: return -nx + 1;
            case 2: return -1;
            case 3: return 0;
            case 4: return 1;
            case 5: return nx - 1;
            default: return nx;
        }
    }
    // position of the first value of row 3c+d in the CSR data array (dof = 3)
    __host__ __device__ __forceinline__ long row_start(long c, int d) const
    {
        return 9 * pairs_before(c) + (long)d * 3 * count(c);
    }
};

}  // namespace ldc
