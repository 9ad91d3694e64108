// Residual kernel variant 4 (default when nx is even and the arrays are 16-byte aligned).
// Included by residual.cu inside namespace ldc, after the FluxOps-independent helpers
// (Strides, shfl_up1, smem_u32, mbar_*, bulk_g2s).
//
// Flux form like variant 2 (same arithmetic in the same order -> bit-identical results), different
// data path:
//   * one autonomous WARP per (column strip, row tile); its rows are streamed through a private
//     shared-memory ring by bulk async copies (one per X row slab, one each for the U_old / V_old
//     rows), completing on the warp's own mbarriers.  Bytes in flight do not live in registers:
//     every warp keeps DEPTH-1 rows (~2.6 KB each) outstanding, ~12 warps per SM hold ~60-90 KB,
//     which covers HBM latency without variant 2's register-resident prefetch;
//   * a lane owns TWO adjacent cells (48 contiguous bytes): three conflict-free 128-bit
//     shared-memory loads per row, three 128-bit global stores of F; DRAM/L2 see whole lines;
//   * the west face of a lane's second cell is its own first cell's east face, so five doubles
//     per TWO cells cross lanes (variant 2: eight per cell), and only lane 0 is a halo lane
//     (62 of 64 columns are written; the east neighbour comes from the slab);
//   * the per-cavity constants are computed on the host in the same IEEE operations and travel
//     as a kernel parameter, so fp64 instructions take them straight from the constant bank;
//   * the warp index is broadcast with a shuffle so the compiler keeps tile geometry, ring slots
//     and copy addresses in uniform registers (bulk copies take uniform operands);
//   * no block-level synchronisation: warps of a block are independent work items.
#pragma once

namespace ring {
constexpr int kCols = 62;                               // writer columns per warp (lanes 1..31, two each)
constexpr int kXSlab = 200;                             // doubles: 65 cells * 3 (+1 to keep 16-byte sizes)
constexpr int kUSlab = 68;                              // 62 (+2 alignment slack) U_old values
constexpr int kVSlab = 64;
constexpr int kStage = kXSlab + kUSlab + kVSlab;        // 332 doubles = 2656 B per ring stage
constexpr int kFSlab = 192;                             // 62 cells * 3 residual rows (bulk-store mode)
constexpr int kFBufs = 2;
constexpr int warp_doubles(int depth, bool bulk_store)
{
    return ((depth * kStage + (bulk_store ? kFBufs * kFSlab : 0) + depth + 15) / 16) * 16;
}
}  // namespace ring

struct RingArgs {
    int n_strips, cols;          // column strips per row, writer columns per strip (even, <= 62)
    int n_row_tiles, rows;       // row tiles per instance, rows per tile
    int edge_prev, edge_next;    // rows of the short tiles next to a neighbour strip (peer kernel; else 0)
    long per_instance, total;    // work items (warps) per instance / in total
    long u_total;                // doubles in U_old over the whole batch (over-read guard)
    long u_off;                  // U_old index of the call's first row relative to the (16-byte aligned) U_old base
    long partial_off;            // first slot of this launch in the sum(F^2) partials
    int debug;                   // peer kernel experiments: 1 = no flag waits
    int full_row;                // nx <= 64: ONE strip from wall to wall, all 32 lanes own two cells (no halo lane)
};

// Per-cavity constants, evaluated on the host with the IEEE operations the device helpers use
// (make_phys / FluxOps): 1/dx, (dx*dy)/dt, ... are single correctly rounded operations either way.
struct RingPhys {
    double dx, dy, inv_dx, inv_dy, rho, mi;
    double kx, ky, hx, hy;       // dy/dx, dx/dy, dx/2, dy/2 as inv_dx*dy, inv_dy*dx, .5*dx, .5*dy
    double vol_dt, bc;           // (dx*dy)/dt and the lid speed when they are not per instance
};

// FluxOps of variant 2 on constant-bank operands (see there for the FAST identities).
template <bool POW2, bool FAST, bool CDS>
struct RingFlux {
    const RingPhys &c;
    const double vol_dt;
    __device__ __forceinline__ RingFlux(const RingPhys &rp, double vdt) : c(rp), vol_dt(vdt) {}
    __device__ __forceinline__ double interp(double vel, double phi_a, double phi_b) const
    {
        if (CDS) return dadd(dmul(0.5, phi_a), dmul(0.5, phi_b));
        return vel > 0.0 ? phi_b : phi_a;
    }
    __device__ __forceinline__ double dif_x(double a, double b) const
    {
        if (FAST) return dmul(dsub(a, b), c.kx);
        const double d = dsub(a, b);
        return dmul(dmul(c.mi, POW2 ? dmul(d, c.inv_dx) : ddiv(d, c.dx)), c.dy);
    }
    __device__ __forceinline__ double dif_y(double a, double b) const
    {
        if (FAST) return dmul(dsub(a, b), c.ky);
        const double d = dsub(a, b);
        return dmul(dmul(c.mi, POW2 ? dmul(d, c.inv_dy) : ddiv(d, c.dy)), c.dx);
    }
    __device__ __forceinline__ double adv_x(double va, double vb, double phi_a, double phi_b) const
    {
        if (FAST) {
            const double vl = dmul(dadd(va, vb), c.hy);
            return dmul(vl, interp(vl, phi_a, phi_b));
        }
        const double vel = dmul(dadd(va, vb), 0.5);
        return dmul(dmul(dmul(c.rho, vel), interp(vel, phi_a, phi_b)), c.dy);
    }
    __device__ __forceinline__ double adv_y(double va, double vb, double phi_a, double phi_b) const
    {
        if (FAST) {
            const double vl = dmul(dadd(va, vb), c.hx);
            return dmul(vl, interp(vl, phi_a, phi_b));
        }
        const double vel = dmul(dadd(va, vb), 0.5);
        return dmul(dmul(dmul(c.rho, vel), interp(vel, phi_a, phi_b)), c.dx);
    }
    __device__ __forceinline__ double transient(double phi, double phi_old) const
    {
        if (FAST) return dmul(dsub(phi, phi_old), vol_dt);
        return dmul(dsub(dmul(c.rho, phi), dmul(c.rho, phi_old)), vol_dt);
    }
};

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_f64x2(double *p, double a, double b)
{
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// ldc_residual_peer: the ghost rows X[-1] / X[row_count] of a strip are pushed into the local buffer
// by the neighbour strips (halo_push_kernel, residual.cu), which then store the evaluation number
// into this strip's flag words.  A warp whose tile needs a ghost row waits (bounded) for the word
// before it starts its tile -- local memory in both cases.  (The wait sits in the kernel body, not
// in the tile function: a divergent spin loop next to the ring bookkeeping made ptxas keep the
// ring slots / copy addresses of EVERY tile flavour in vector registers: 168 registers instead of
// 138 and a 15 % slower kernel.)
struct RingGhost {
    const unsigned long long *prev_word, *next_word;   // my_flags[0] / my_flags[1]; nullptr: no neighbour
    unsigned long long epoch;                          // evaluation number the words must have reached
    int32_t *status;
};

__device__ __forceinline__ unsigned long long ld_flag_relaxed(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// bounded wait (a few seconds): a lost neighbour must not hang the device.  The word is polled
// with relaxed loads and ONE system-scope acquire fence follows the successful poll (pairs with the
// release store of halo_push_kernel); then the ghost row is ordered before the bulk copy that
// follows (generic -> async proxy fence).  Measured: neither fence costs anything per step.
__device__ __forceinline__ void wait_ghost(const unsigned long long *word, unsigned long long epoch,
                                           int32_t *status)
{
    for (long it = 0; it < (1L << 24); ++it) {
        if (ld_flag_relaxed(word) >= epoch) {
            asm volatile("fence.acquire.sys;" ::: "memory");
            asm volatile("fence.proxy.async.global;" ::: "memory");
            return;
        }
        __nanosleep(it < 16 ? 100 : 400);
    }
    if (status) atomicExch(status, 1);
}

struct Cell2 {   // the two cells of a lane and the cell east of them: raw slots of one grid row
    double aP, aU, aV, bP, bU, bV, eP, eU, eV;
};

// One (column strip, row tile) for one warp.  INTERIOR: no lane touches a wall, the lid, a dummy
// slot or a missing ghost row, so every mask folds away at compile time.  Everything except
// `lane` and what is derived from it is warp-uniform.
template <bool POW2, bool FAST, bool CDS, bool INTERIOR, bool NORM, bool BULK_STORE, int DEPTH, bool FULL = false>
__device__ __forceinline__ double ring_tile(const ldc_grid &g, const RingPhys &rp, double vol_dt, double bc,
                                            const RingArgs &ra, const double *__restrict__ X,
                                            const double *__restrict__ U_base,
                                            const double *__restrict__ V_old, double *__restrict__ F,
                                            long u_inst_off, double *smem, int lane, int c0, int cw,
                                            int jl0, int jl1)
{
    const RingFlux<POW2, FAST, CDS> fx(rp, vol_dt);
    const int nx = g.nx, ny = g.ny;
    double *stages = smem;
    double *fbuf = smem + DEPTH * ring::kStage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(fbuf + (BULK_STORE ? ring::kFBufs * ring::kFSlab : 0));

    // ---- geometry of this warp's strip
    // Narrow grids (nx <= 64, e.g. the 64x64 cavities of a Reynolds ensemble): one strip from wall to
    // wall, lane L owns cells 2L, 2L+1 and lane 0 evaluates the west-wall face itself instead of
    // being the halo lane -- a 64-column row is one warp instead of two half-empty ones.
    constexpr bool full = FULL;   // compile-time: a run-time switch here cost the wall tiles of wide grids 2.5 %
    const int lo = full ? 2 * lane : 2 * lane - 2;    // lane's first column relative to c0
    const int ia = c0 + lo;                           // first cell of the lane; second ia+1, east ia+2
    const bool writer = full ? (lo < cw) : (lane >= 1 && 2 * lane <= cw);   // lanes that own columns of the strip
    const bool ab_in = INTERIOR || (ia >= 0 && ia < nx);           // nx even: both cells or none
    const bool e_in = INTERIOR || (ia + 2 < nx);
    const bool b_ucol = INTERIOR || (ia + 1 >= 0 && ia + 1 < nx - 1);   // cell b's U slot is a real U
    const bool a_side = !INTERIOR && ia < 0;                            // east face of a lies outside
    const bool b_side = !INTERIOR && (ia + 1 < 0 || ia + 1 >= nx - 1);
    const int row_lo = (g.row_begin > 0) ? -1 : 0;
    const int row_hi = g.row_count + ((g.row_begin + g.row_count < ny) ? 1 : 0);   // exclusive
    const int sc0 = c0 >= 2 ? c0 - 2 : 0;                               // first staged column
    const int sc1 = (c0 + cw + 1 < nx) ? c0 + cw + 1 : nx;              // exclusive
    const uint32_t xbytes = 8u * (uint32_t)((3 * (sc1 - sc0) + 1) & ~1);
    const uint32_t vbytes = 8u * (uint32_t)cw;
    const int xoff = 3 * (ia - sc0);                                    // lane's cell a inside an X slab
    const int ucnt = ((c0 + cw < nx - 1) ? c0 + cw : nx - 1) - c0;      // real U columns of the strip
    const int T = (jl1 - jl0) + 2;                                      // stages: X rows jl0-1 .. jl1

    // ---- producer state: stage t carries X row jl0-1+t and, for t >= 2, U_old / V_old of row
    // jl0+t-2.  Every lane advances the (uniform) state, lane 0 issues the copies.
    const double *x_src = X + 3L * ((long)sc0 + (long)nx * (jl0 - 1));
    const double *v_src = V_old + (long)c0 + (long)nx * jl0;
    long u_gi = u_inst_off + (long)c0 + (long)(nx - 1) * jl0;           // index from the aligned base
    int t_next = 0, slot_next = 0;
    auto issue = [&]() {
        const int jx = jl0 - 1 + t_next;
        const bool has_old = t_next >= 2;
        const bool x_ok = INTERIOR || (jx >= row_lo && jx < row_hi);
        const long ga = u_gi & ~1L, ge = (u_gi + ucnt + 1) & ~1L;
        // an aligned copy that would read past the end of U_old is replaced by direct loads
        const bool u_ok = has_old && ge <= ra.u_total;
        const bool v_ok = has_old && (INTERIOR || g.row_begin + jx - 1 < ny - 1);
        const uint32_t ub = 8u * (uint32_t)(ge - ga);
        if (lane == 0) {
            double *dst = stages + slot_next * ring::kStage;
            uint64_t *bar = bars + slot_next;
            mbar_expect_tx(bar, (x_ok ? xbytes : 0u) + (u_ok ? ub : 0u) + (v_ok ? vbytes : 0u));
            if (x_ok) bulk_g2s(dst, x_src, xbytes, bar);
            if (u_ok) bulk_g2s(dst + ring::kXSlab, U_base + ga, ub, bar);
            if (v_ok) bulk_g2s(dst + ring::kXSlab + ring::kUSlab, v_src, vbytes, bar);
        }
        x_src += 3L * nx;
        if (has_old) {
            u_gi += nx - 1;
            v_src += nx;
        }
        ++t_next;
        slot_next = (slot_next + 1 == DEPTH) ? 0 : slot_next + 1;
    };

    // ---- consumer state
    int slot = 0;
    uint32_t parity = 0;
    auto wait_slot = [&]() { mbar_wait(bars + slot, parity); };
    auto next_slot = [&]() {
        if (++slot == DEPTH) {
            slot = 0;
            parity ^= 1u;
        }
    };
    auto read_row = [&](bool row_ok) {
        Cell2 c;
        const double *q = stages + slot * ring::kStage + xoff;
        if (INTERIOR || (row_ok && ab_in)) {
            const double2 v0 = *reinterpret_cast<const double2 *>(q);
            const double2 v1 = *reinterpret_cast<const double2 *>(q + 2);
            const double2 v2 = *reinterpret_cast<const double2 *>(q + 4);
            c.aP = v0.x; c.aU = v0.y; c.aV = v1.x; c.bP = v1.y; c.bU = v2.x; c.bV = v2.y;
        } else {
            c.aP = c.aU = c.aV = c.bP = c.bU = c.bV = 0.0;
        }
        if (INTERIOR || (row_ok && e_in)) {
            const double2 v3 = *reinterpret_cast<const double2 *>(q + 6);
            c.eP = v3.x; c.eU = v3.y; c.eV = q[8];
        } else {
            c.eP = c.eU = c.eV = 0.0;
        }
        return c;
    };
    auto row_ok = [&](int jl) { return INTERIOR || (jl >= row_lo && jl < row_hi); };
    // masks of one grid row: dummy U slot of column nx-1, V slots of the top row
    auto mask_row = [&](Cell2 &c, int jg) {
        if (!INTERIOR) {
            if (!b_ucol) c.bU = 0.0;
            if (!(jg >= 0 && jg < ny - 1)) { c.aV = 0.0; c.bV = 0.0; c.eV = 0.0; }
        }
    };

    if (lane == 0) {
        for (int s = 0; s < DEPTH; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < DEPTH; ++s)
        if (s < T) issue();

    // ---- prologue: south faces of the first row = north faces of row jl0-1
    double AUs_a, DUs_a, AVs_a, DVs_a, CVs_a, AUs_b, DUs_b, AVs_b, DVs_b, CVs_b;
    Cell2 C;                         // current row, masked
    double rawU_b, rawV_a, rawV_b;   // its raw dummy slots (rows F = X)
    {
        wait_slot();
        Cell2 S = read_row(row_ok(jl0 - 1));
        next_slot();
        wait_slot();
        C = read_row(true);
        next_slot();
        const int jg = g.row_begin + jl0;
        rawU_b = C.bU; rawV_a = C.aV; rawV_b = C.bV;
        mask_row(S, jg - 1);
        mask_row(C, jg);
        AUs_a = fx.adv_y(S.bV, S.aV, C.aU, S.aU);
        DUs_a = fx.dif_y(C.aU, S.aU);
        AVs_a = fx.adv_y(S.aV, C.aV, C.aV, S.aV);
        DVs_a = fx.dif_y(C.aV, S.aV);
        CVs_a = dmul(S.aV, rp.dx);
        AUs_b = fx.adv_y(S.eV, S.bV, C.bU, S.bU);
        DUs_b = fx.dif_y(C.bU, S.bU);
        AVs_b = fx.adv_y(S.bV, C.bV, C.bV, S.bV);
        DVs_b = fx.dif_y(C.bV, S.bV);
        CVs_b = dmul(S.bV, rp.dx);
    }
    __syncwarp();        // every lane has read the two prologue slabs: their slots can be refilled
    if (t_next < T) issue();
    if (t_next < T) issue();

    double sq = 0.0;
    double *frow = F + 3L * ((long)(BULK_STORE ? c0 : ia) + (long)nx * jl0);   // strip / lane start of row jl
    long u_rd = u_inst_off + (long)c0 + (long)(nx - 1) * jl0;                  // U_old index of the row being read
    const uint32_t fbytes = 24u * (uint32_t)cw;
    int fslot = 0;
#pragma unroll 2
    for (int jl = jl0; jl < jl1; ++jl, frow += 3L * nx, u_rd += nx - 1) {
        const int jg = g.row_begin + jl;
        wait_slot();
        const double *stg = stages + slot * ring::kStage;
        Cell2 N = read_row(row_ok(jl + 1));
        const double NrawU_b = N.bU, NrawV_a = N.aV, NrawV_b = N.bV;
        mask_row(N, jg + 1);
        double Uold_a = 0.0, Uold_b = 0.0, Vold_a = 0.0, Vold_b = 0.0;
        if (writer) {
            const bool staged = ((u_rd + ucnt + 1) & ~1L) <= ra.u_total;
            if (staged) {
                const double *uq = stg + ring::kXSlab + lo + (int)(u_rd & 1);
                Uold_a = uq[0];
                if (b_ucol) Uold_b = uq[1];
            } else {
                const double *uq = U_base + u_rd + lo;
                Uold_a = __ldg(uq);
                if (b_ucol) Uold_b = __ldg(uq + 1);
            }
            if (INTERIOR || jg < ny - 1) {
                const double2 vv =
                    *reinterpret_cast<const double2 *>(stg + ring::kXSlab + ring::kUSlab + lo);
                Vold_a = vv.x;
                Vold_b = vv.y;
            }
        }
        next_slot();

        // ---- north faces of a and b
        const bool lid = !INTERIOR && jg == ny - 1;
        const double UN_a = lid ? bc : N.aU, UN_b = lid ? bc : N.bU;
        const double AUn_a = fx.adv_y(C.bV, C.aV, UN_a, C.aU);
        const double DUn_a = fx.dif_y(UN_a, C.aU);
        const double AVn_a = fx.adv_y(C.aV, N.aV, N.aV, C.aV);
        const double DVn_a = fx.dif_y(N.aV, C.aV);
        const double CVn_a = dmul(C.aV, rp.dx);
        const double AUn_b = fx.adv_y(C.eV, C.bV, UN_b, C.bU);
        const double DUn_b = fx.dif_y(UN_b, C.bU);
        const double AVn_b = fx.adv_y(C.bV, N.bV, N.bV, C.bV);
        const double DVn_b = fx.dif_y(N.bV, C.bV);
        const double CVn_b = dmul(C.bV, rp.dx);

        // ---- east faces of a (towards b) and of b (towards e)
        const bool top2 = !INTERIOR && jg == ny - 2;
        const double UNE_a = a_side ? 0.0 : (top2 ? bc : N.aU);
        const double UNE_b = b_side ? 0.0 : (top2 ? bc : N.bU);
        const double AUe_a = fx.adv_x(C.bU, C.aU, C.bU, C.aU);
        const double DUe_a = fx.dif_x(C.bU, C.aU);
        const double AVe_a = fx.adv_x(UNE_a, C.aU, C.bV, C.aV);
        const double DVe_a = fx.dif_x(C.bV, C.aV);
        const double CUe_a = dmul(C.aU, rp.dy);
        const double AUe_b = fx.adv_x(C.eU, C.bU, C.eU, C.bU);
        const double DUe_b = fx.dif_x(C.eU, C.bU);
        const double AVe_b = fx.adv_x(UNE_b, C.bU, C.eV, C.bV);
        const double DVe_b = fx.dif_x(C.eV, C.bV);
        const double CUe_b = dmul(C.bU, rp.dy);

        // ---- west face of a = east face of the previous lane's b
        double AUw_a = shfl_up1(AUe_b);
        double DUw_a = shfl_up1(DUe_b);
        double AVw_a = shfl_up1(AVe_b);
        double DVw_a = shfl_up1(DVe_b);
        double CUw_a = shfl_up1(CUe_b);
        if (full && lane == 0) {
            // the west-wall face of cell 0: what a halo lane holding two zero (outside) cells west
            // of it computes as ITS east face of b -- same expressions, same bits
            AUw_a = fx.adv_x(C.aU, 0.0, C.aU, 0.0);
            DUw_a = fx.dif_x(C.aU, 0.0);
            AVw_a = fx.adv_x(0.0, 0.0, C.aV, 0.0);
            DVw_a = fx.dif_x(C.aV, 0.0);
            CUw_a = dmul(0.0, rp.dy);
        }

        if (writer) {
            const bool v_row = INTERIOR || jg < ny - 1;
            double a0, a1, a2, b0, b1, b2;
            a0 = dadd(dsub(CUe_a, CUw_a), dsub(CVn_a, CVs_a));
            {
                const double tr = fx.transient(C.aU, Uold_a);
                const double adv = dsub(dadd(dsub(AUe_a, AUw_a), AUn_a), AUs_a);
                const double dif = dsub(dadd(dsub(DUe_a, DUw_a), DUn_a), DUs_a);
                const double src = dmul(-dsub(C.bP, C.aP), rp.dy);
                a1 = dsub(dsub(dadd(tr, adv), dif), src);
            }
            if (v_row) {
                const double tr = fx.transient(C.aV, Vold_a);
                const double adv = dsub(dadd(dsub(AVe_a, AVw_a), AVn_a), AVs_a);
                const double dif = dsub(dadd(dsub(DVe_a, DVw_a), DVn_a), DVs_a);
                const double src = dmul(-dsub(N.aP, C.aP), rp.dx);
                a2 = dsub(dsub(dadd(tr, adv), dif), src);
            } else {
                a2 = rawV_a;   // dummy row: F = X
            }
            b0 = dadd(dsub(CUe_b, CUe_a), dsub(CVn_b, CVs_b));
            if (b_ucol) {
                const double tr = fx.transient(C.bU, Uold_b);
                const double adv = dsub(dadd(dsub(AUe_b, AUe_a), AUn_b), AUs_b);
                const double dif = dsub(dadd(dsub(DUe_b, DUe_a), DUn_b), DUs_b);
                const double src = dmul(-dsub(C.eP, C.bP), rp.dy);
                b1 = dsub(dsub(dadd(tr, adv), dif), src);
            } else {
                b1 = rawU_b;
            }
            if (v_row) {
                const double tr = fx.transient(C.bV, Vold_b);
                const double adv = dsub(dadd(dsub(AVe_b, AVe_a), AVn_b), AVs_b);
                const double dif = dsub(dadd(dsub(DVe_b, DVe_a), DVn_b), DVs_b);
                const double src = dmul(-dsub(N.bP, C.bP), rp.dx);
                b2 = dsub(dsub(dadd(tr, adv), dif), src);
            } else {
                b2 = rawV_b;
            }
            if (BULK_STORE) {
                double *fq = fbuf + fslot * ring::kFSlab + 3 * lo;
                *reinterpret_cast<double2 *>(fq) = make_double2(a0, a1);
                *reinterpret_cast<double2 *>(fq + 2) = make_double2(a2, b0);
                *reinterpret_cast<double2 *>(fq + 4) = make_double2(b1, b2);
                fence_async_smem();   // the slab is read by the async proxy (bulk store) next
            } else {
                st_f64x2(frow, a0, a1);
                st_f64x2(frow + 2, a2, b0);
                st_f64x2(frow + 4, b1, b2);
            }
            if (NORM) {
                sq += a0 * a0 + a1 * a1 + a2 * a2;
                sq += b0 * b0 + b1 * b1 + b2 * b2;
            }
        }
        __syncwarp();   // every lane has read this iteration's stage (and written its F slab)
        if (BULK_STORE) {
            if (lane == 0) {
                bulk_s2g(frow, fbuf + fslot * ring::kFSlab, fbytes);
                bulk_commit();
            }
            fslot ^= 1;
        }
        if (t_next < T) issue();   // refill the slot just read
        if (BULK_STORE) {
            if (lane == 0) bulk_wait_read<ring::kFBufs - 1>();   // the other F slab has been drained
            __syncwarp();
        }
        AUs_a = AUn_a; DUs_a = DUn_a; AVs_a = AVn_a; DVs_a = DVn_a; CVs_a = CVn_a;
        AUs_b = AUn_b; DUs_b = DUn_b; AVs_b = AVn_b; DVs_b = DVn_b; CVs_b = CVn_b;
        C = N;
        rawU_b = NrawU_b; rawV_a = NrawV_a; rawV_b = NrawV_b;
    }
    if (BULK_STORE && lane == 0) bulk_wait_all();   // stores complete before the shared memory goes away
    return sq;
}

// work item -> tile.  Row tiles: uniform `rows` tall, except that a strip with a neighbour
// (peer kernel) starts / ends with a SHORT tile of edge_prev / edge_next rows; those two come
// first in the work order so that their NVLink round trips overlap everything else.
struct RingTile {
    int inst, c0, cw, jl0, jl1;
};
__device__ __forceinline__ RingTile ring_decode(const ldc_grid &g, const RingArgs &ra, long w)
{
    RingTile t;
    t.inst = (int)(w / ra.per_instance);
    const int r = (int)(w - (long)t.inst * ra.per_instance);
    int rt = r / ra.n_strips;
    const int strip = r - rt * ra.n_strips;
    t.c0 = strip * ra.cols;
    t.cw = min(ra.cols, g.nx - t.c0);
    const int n_edges = (ra.edge_prev ? 1 : 0) + (ra.edge_next ? 1 : 0);
    if (ra.edge_prev && rt == 0) {
        t.jl0 = 0;
        t.jl1 = ra.edge_prev;
    } else if (ra.edge_next && rt == (ra.edge_prev ? 1 : 0)) {
        t.jl0 = g.row_count - ra.edge_next;
        t.jl1 = g.row_count;
    } else {
        rt -= n_edges;
        t.jl0 = ra.edge_prev + rt * ra.rows;
        t.jl1 = min(t.jl0 + ra.rows, g.row_count - ra.edge_next);
    }
    return t;
}

// PEER = false: the residual kernel of ldc_residual.
// PEER = true : ldc_residual_peer.  The ghost rows next to a neighbour strip arrive in the LOCAL
// buffer through the neighbours' halo_push_kernel; a warp whose tile touches the strip's first /
// last owned row waits (acquire, bounded) until its flag word shows the evaluation number and
// then runs its tile on local memory.  Edge tiles are short and first in the work order.  The
// kernel performs no remote access on purpose: a kernel that stores to peer memory pays a
// system-scope flush of ~2.7 us at its end (measured), and a spin loop inside the tile function
// costs every tile flavour registers (168 instead of 138).
// MAXREG: register cap per thread; up to 168 keep three 4-warp blocks (12 warps) resident per SM.
template <bool POW2, bool FAST, bool CDS, bool NORM, bool PEER, int WARPS, int DEPTH, int MAXREG>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(MAXREG)
residual_ring_kernel(const ldc_grid g, const RingPhys rp, const RingArgs ra, const RingGhost gh,
                     const double *__restrict__ X, const double *__restrict__ U_old,
                     const double *__restrict__ V_old, double *__restrict__ F, double *__restrict__ partials)
{
    constexpr bool BULK_STORE = false;
    extern __shared__ __align__(128) double ring_smem[];
    // let the next kernel of the stream start launching (programmatic dependent launch)
    asm volatile("griddepcontrol.launch_dependents;");
    // the warp index through a shuffle: the compiler then knows it is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * WARPS + warp;      // work item of this warp
    if (w >= ra.total) return;
    const RingTile t = ring_decode(g, ra, w);
    const Strides st = instance_strides(g);
    X += t.inst * st.x;
    F += t.inst * st.x;
    V_old += t.inst * st.v;
    double *smem = ring_smem + (size_t)warp * ring::warp_doubles(DEPTH, BULK_STORE);
    // do not touch global memory before every kernel queued ahead of us has completed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const double bc = g.bc_dev ? g.bc_dev[t.inst] : rp.bc;
    const double vol_dt = g.dt_dev ? ddiv(dmul(rp.dx, rp.dy), g.dt_dev[t.inst]) : rp.vol_dt;
    const int jg0 = g.row_begin + t.jl0, jg1 = g.row_begin + t.jl1;
    bool edge = false;
    if (PEER) {
        const bool needs_prev = t.jl0 == 0 && gh.prev_word != nullptr;
        const bool needs_next = t.jl1 == g.row_count && gh.next_word != nullptr;
        edge = needs_prev || needs_next;
        if (edge && !(ra.debug & 1)) {
            if (lane == 0) {
                if (needs_prev) wait_ghost(gh.prev_word, gh.epoch, gh.status);
                if (needs_next) wait_ghost(gh.next_word, gh.epoch, gh.status);
            }
            __syncwarp();
        }
    }
    // warp-uniform: no lane touches a wall, the lid, a dummy slot or a missing ghost row
    const bool interior = t.c0 >= 2 && t.c0 + t.cw + 1 <= g.nx - 1 && jg0 >= 1 && jg1 <= g.ny - 2 && !edge;
    const long u_off = t.inst * st.u + ra.u_off;
    double sq;
    if (interior)
        sq = ring_tile<POW2, FAST, CDS, true, NORM, BULK_STORE, DEPTH>(g, rp, vol_dt, bc, ra, X, U_old, V_old, F, u_off,
                                                                      smem, lane, t.c0, t.cw, t.jl0, t.jl1);
    else if (!PEER && ra.full_row)
        sq = ring_tile<POW2, FAST, CDS, false, NORM, BULK_STORE, DEPTH, true>(g, rp, vol_dt, bc, ra, X, U_old, V_old, F,
                                                                             u_off, smem, lane, t.c0, t.cw, t.jl0,
                                                                             t.jl1);
    else
        sq = ring_tile<POW2, FAST, CDS, false, NORM, BULK_STORE, DEPTH>(g, rp, vol_dt, bc, ra, X, U_old, V_old, F, u_off,
                                                                       smem, lane, t.c0, t.cw, t.jl0, t.jl1);
    if (NORM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) partials[ra.partial_off + w] = sq;
    }
}
