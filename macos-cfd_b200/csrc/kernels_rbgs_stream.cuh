// kernels_rbgs_stream.cuh -- PressureSolver::smooth (pressure.cpp:64-111) as ONE streaming kernel: every warp marches up
// its own 64-column strip with the four colour passes and the residual following each other two rows apart.
//
// Why (ncu on the tile kernel k_rbgs_fused, profiles/r01_*): it reaches the traffic floor (24 B/cell) but runs at 39 %
// issue utilisation -- 144 instructions per cell, block-wide barriers between passes, a load phase nobody overlaps.
// Here
//   * a warp owns columns [X0, X0+64) (52 of them are output, 6 on either side are halo that is recomputed) and keeps
//     a ring of the last 16 rows of p and rhs in shared memory, private to the warp: no block barrier anywhere, one
//     __syncwarp per row;
//   * rows arrive through cp.async four rows ahead of their use (no registers, no stall on the load; plain loads into
//     registers were measured: they serialise on the scoreboard and stall every step);
//   * at the step that makes row L available, pass k (k = 0..3) updates its colour in row L-1-2k and the residual is
//     taken in row L-9.  With two rows between consecutive passes the four updates of a step touch disjoint rows and
//     are independent of each other (the rows a pass writes, L-1, L-3, L-5, L-7, are never read by another pass in the
//     same step), so a lane has 4 update chains + 2 residual chains in flight;
//   * of the four neighbours of an update only ONE is read from shared memory (the point above): the point above this
//     step's point is next step's horizontal neighbour inside the lane's own column pair, this step's in-pair neighbour
//     is next step's point below -- both stay in registers -- and the neighbour in the adjacent lane's pair is that
//     lane's in-pair neighbour, fetched with a register shuffle; the residual's own columns are pass 3's last three
//     results and its cross-lane neighbours the adjacent lanes' copies of those;
//   * a ring row holds its 64 columns in order with one double of padding after every 16: a lane's column (stride 16 B
//     across the warp) then hits 16 different bank pairs per half-warp, so every access is conflict-free, and the
//     cp.async copies read whole sectors of global memory;
//   * the loop is unrolled over the 16 ring slots: every shared-memory address is register + immediate.
// Measured and dropped (B200, 8192^2, per smooth; this version: 0.42 ms): colour-split ring with 8-byte cp.async 0.42 ms
// (twice the shared-memory wavefronts for the copies); rows through registers (LDG -> STS, 2 or 4 rows in flight) 0.56 ms
// (the loads serialise on the scoreboard); all loads of a step hoisted in front of the arithmetic 0.44 ms; XOR-swizzled
// instead of padded rows 0.59 ms.  ncu: 76 % of the shared-memory wavefront peak, 40 % issue slots, 47 % DRAM.
// Updates are in place: a pass only writes its own colour and reads the other one, exactly like the sequential
// red-black passes, so every update sees the operands the reference's pass would see and the result is bit-identical.
// Halo points (columns and, at strip ends, rows) are recomputed with the same expression; what lies outside the
// "light cone" of the outputs may hold garbage and is never stored.
//
// Non-finite values.  The reference guards every operand with isfinite (pressure.cpp:84-88).  This kernel does not:
// a non-finite operand makes the update non-finite (sums and products with positive constants cannot lose an Inf/NaN),
// so it surfaces in a stored value or a residual term.  Those are tested; a hit sets Scalars::rb_redo and the guarded
// tile kernel (k_rbgs_fused, launched right behind this one, a no-op otherwise) redoes the whole smooth() from the
// untouched input buffer (p is ping-ponged).  For finite data the guards are identities, so the fast path is exact.
#pragma once
#include <type_traits>

#include "kernels_pressure.cuh"

#define RS_R 16     // ring rows (power of two): rows t-9 .. t live at step t, 4 more in flight
#ifndef RS_D
#define RS_D 7      // rows in flight (the row above pass 0 is read one step before its use); <= 7 with a 16-row ring
#endif
#define RS_OUT 52   // output columns per warp
#define RS_HX 6     // halo columns (5 needed; 6 keeps X0 odd = 16-byte aligned pairs in global memory)
#define RS_HY 5     // halo rows
#ifndef RS_WARPS
#define RS_WARPS 4  // warps per block (independent of each other; adjacent strips, so a block reads 4 x 512 contiguous bytes per row)
#endif
#ifndef RS_VC_BLOCKS
#define RS_VC_BLOCKS 2 // resident blocks per SM the V-cycle variant is compiled for
#endif
#define RS_BLOCKS_PER_SM (12 / RS_WARPS) // 12 warps per SM fit (shared memory: 17 KB per warp)
#ifndef RS_PAD
#define RS_PAD 0    // 1: one double of padding after every 16 columns, rows copied with 8-byte cp.async; 0: dense rows, 16-byte cp.async
#endif
#ifndef RS_BULK
#define RS_BULK 0   // 1 (needs RS_PAD 0): rows arrive through cp.async.bulk (the TMA unit; one lane issues two 512-byte copies per
                    // step, completion on an mbarrier per ring slot) instead of one 16-byte cp.async per lane and field
#endif
#define RS_ROW (RS_PAD ? 68 : 64)   // doubles per ring row
#define RS_WDOUBLES (2 + RS_R * RS_ROW + 2 + RS_R * RS_ROW + (RS_BULK ? RS_R : 0)) // per warp: pad, p ring, pad, rhs ring, mbarriers
#define RS_SMEM (RS_WARPS * RS_WDOUBLES * 8)

__device__ __forceinline__ void rs_cp8(double *smem, const double *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void rs_cp16(double *smem, const double *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
// ---- cp.async.bulk + mbarrier (RS_BULK)
__device__ __forceinline__ void rs_mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect(void *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_bulk(double *smem, const double *gmem, int bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(void *bar, int parity) {
    asm volatile("{\n.reg .pred p;\nRS_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra RS_DONE;\nbra RS_WAIT;\nRS_DONE:\n}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void rs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void rs_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// VC = 0: smooth() of the reference (above).  VC = 1: two red-black sweeps of the true V-cycle extension
// (kernels_vcycle.cuh: k_mg_colour x 4 in one launch): p = (sum - rhs) / (denom + wall), where `wall` moves the p = 0
// condition onto the wall faces of boundary cells; no isfinite guards in that algorithm, ghost ring = 0.  What follows the
// sweeps in the cycle rides on the residual stage (`finish`): RS_VC_RESTRICT = residual + 4-cell-mean restriction to the
// coarse right-hand side + coarse initial guess 0 (k_mg_restrict), RS_VC_NORM = ||rhs - L p||_2 and the cycle loop's
// exit test (k_mg_residual_norm).  Strips then start at odd rows and hold an even number of rows (a 2x2 block of fine
// cells never straddles two strips).
enum { RS_VC_SMOOTH = 0, RS_VC_RESTRICT = 1, RS_VC_NORM = 2 };
struct RsCoarse {
    double *rhs, *p; // coarse level arrays (common padded layout)
    int P;
};
template <int POW2, int VC>
__global__ void __launch_bounds__(32 * RS_WARPS, VC ? RS_VC_BLOCKS : RS_BLOCKS_PER_SM) // VC: 2 blocks per SM keep its extra operands out of local memory
    k_rbgs_stream(Geom g, const double *__restrict__ pin, double *__restrict__ pout, const double *__restrict__ rhs,
                  Scalars *sc, double *partials, double tol, int RY, int nstrips, int finish, RsCoarse cz, WaitArgs hw) { pdl_begin();
    // Strip of rows this block works on.  The first strip goes to the LAST blocks of the grid: on a row slab the strips at
    // the slab's ends read the neighbours' rows, and by the time those blocks are scheduled the rows have long arrived
    // (a block that waits at the start of the kernel delays everything queued behind it on its SM).
    const int by = VC ? (int)blockIdx.y : (blockIdx.y + 1 == gridDim.y ? 0 : (int)blockIdx.y + 1);
    if (!VC) { // row slabs: only the strips at the slab's ends read the neighbours' 5 rows of p and rhs
        const int yb0 = (g.bottom ? 0 : 1) + by * RY, yb1 = yb0 + RY - 1;
        p2p_wait_block(hw, yb0 - RS_HY - 1 < 1, yb1 + RS_HY > g.ny, sc);
    }
    extern __shared__ __align__(16) double rs_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *ringp = rs_raw + w * RS_WDOUBLES + 2;   // [slot][RS_ROW]
    double *ringr = ringp + RS_R * RS_ROW + 2;      // [slot][RS_ROW]
    const int wx = blockIdx.x * RS_WARPS + w;
    const bool done = sc->done != 0; // converged earlier: only carry p over to the other buffer
    double acc[1] = {0.0};
    int redo = 0;
    if (wx < nstrips) {
        const int X0 = 1 + RS_OUT * wx - RS_HX; // odd: (X0 + CFD_XOFF) is even, pairs are 16-byte aligned
        const int a = X0 + 2 * lane;            // this lane's columns: a, a + 1
        // rows this launch stores: the slab's cell rows plus the ghost rows the slab owns (carried over unchanged)
        const int ylo = g.bottom ? 0 : 1, yhi = g.top ? g.ny + 1 : g.ny;
        const int y0 = VC ? (by ? 1 + by * RY : 0) : ylo + by * RY;
        const int y1 = VC ? min((by + 1) * RY, yhi) : min(y0 + RY - 1, yhi);
        // rows that may be updated: cells of this slab and, towards a neighbouring slab, its exchanged halo rows
        const int jlo = g.bottom ? 1 : 1 - RS_HY + 1, jhi = g.top ? g.ny : g.ny + RS_HY - 1;
        const bool c0 = a >= 1 && a <= g.nx, c1 = a + 1 >= 1 && a + 1 <= g.nx; // interior columns
        // pairs this lane stores: the 26 output pairs, plus the pair holding ghost column 0 in the first strip
        const bool st_lane = ((lane >= RS_HX / 2 && lane < RS_HX / 2 + RS_OUT / 2) || (wx == 0 && lane == RS_HX / 2 - 1)) && a <= g.nx + 1;
        const bool in_lane = lane >= RS_HX / 2 && lane < RS_HX / 2 + RS_OUT / 2;
        const size_t P = g.P;
        if (done) {
            if (st_lane)
                for (int jj = y0; jj <= y1; ++jj) {
                    const size_t o = gidx(g, a, jj);
                    *reinterpret_cast<double2 *>(pout + o) = *reinterpret_cast<const double2 *>(pin + o);
                }
        } else {
            // first row loaded (relative row 0): RS_HY rows below the strip, one more where that makes column a of relative
            // row 0 colour 0 -- then the column offset E of a pass's point inside the lane's pair is the compile-time value
            // x = colour ^ row parity of the unrolled step, and every select / index on it below folds away
            const int ys = y0 - RS_HY - ((y0 - RS_HY + g.j0 + 1) & 1);
            const int nload = y1 + RS_HY - ys + 1;     // rows loaded
            const int T = y1 + 9 - ys + 1;             // steps: the residual (row L-9) must reach y1
            // cp.async: lane l copies columns X0+l and X0+32+l of the next row (whole sectors per instruction)
            const double *gp_next = pin + gidx(g, X0 + (RS_BULK ? 0 : RS_PAD ? lane : 2 * lane), ys);
            const double *gr_next = rhs + gidx(g, X0 + (RS_BULK ? 0 : RS_PAD ? lane : 2 * lane), ys);
            unsigned long long *mbar = reinterpret_cast<unsigned long long *>(ringr + RS_R * RS_ROW); // [slot], RS_BULK only
            if (RS_BULK) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < RS_R; ++i) rs_mbar_init(mbar + i, 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
                __syncwarp();
            }
            double *po = pout + gidx(g, a, ys - 7);    // where pass 3's row (relative t-7) is stored at step t
            const double idx2 = g.idx2, idy2 = g.idy2;
            // ring column i lives at i + (i >> 4)
            const int cpA = lane + (lane >> 4), cpB = 32 + lane + 2 + (lane >> 4); // cp.async destinations
            const int lb = 2 * lane + (RS_PAD ? (lane >> 3) : 0);                   // column a
            const double *const Bp = ringp + lb; // + x: the pass's point (x = colour ^ row parity = its column offset in the pair)
            const double *const Rp = ringr + lb;
            const bool colok[2] = {c0, c1};
            // VC: the wall term of mg_wall(), column part
            const double wxc[2] = {(double)((a == 1) + (a == g.nx)) * idx2, (double)((a + 1 == 1) + (a + 1 == g.nx)) * idx2}; // by column
            double sprev = 0.0; // RS_VC_RESTRICT: residual sum of the pair's columns in the odd row of a 2x2 block
            const int t_jlo = jlo - ys, t_jhi = jhi - ys;                  // updatable relative rows
            const int t_st0 = y0 - ys, t_st1 = y1 - ys;                    // stored relative rows
            const int rr_first = max(y0, 1) - ys, rr_last = min(y1, g.ny) - ys; // residual rows (relative)
            const bool inA = in_lane && c0, inB = in_lane && c1;
            auto issue = [&](int rt, int slot) { // request relative row rt into ring slot `slot`
                if (RS_BULK) {
                    if (rt < nload && lane == 0) {
                        rs_mbar_expect(mbar + slot, 2 * 64 * 8);
                        rs_bulk(ringp + slot * RS_ROW, gp_next, 64 * 8, mbar + slot);
                        rs_bulk(ringr + slot * RS_ROW, gr_next, 64 * 8, mbar + slot);
                    }
                    gp_next += P, gr_next += P;
                    return;
                }
                if (rt < nload) {
                    if (RS_PAD) {
                        rs_cp8(ringp + slot * RS_ROW + cpA, gp_next);
                        rs_cp8(ringp + slot * RS_ROW + cpB, gp_next + 32);
                        rs_cp8(ringr + slot * RS_ROW + cpA, gr_next);
                        rs_cp8(ringr + slot * RS_ROW + cpB, gr_next + 32);
                    } else { // lane l copies its own pair: columns X0 + 2l, X0 + 2l + 1
                        rs_cp16(ringp + slot * RS_ROW + 2 * lane, gp_next);
                        rs_cp16(ringr + slot * RS_ROW + 2 * lane, gr_next);
                    }
                }
                rs_commit();
                gp_next += P, gr_next += P;
            };
#pragma unroll
            for (int d = 0; d < RS_D; ++d) issue(d, d);
            double ho[4], bo[4];           // per pass: in-pair horizontal neighbour, neighbour below (see header)
            double last[4];                // per pass: the value it left at its point in the previous step = the point
                                           // ABOVE the next pass's point of this step (same column: x flips with the row)
            double r1[2], r2[2], r3[2];    // the lane's final pairs of rows t-8, t-9, t-10 (pass 3's last three rows)
            // operands of the CURRENT step that were fetched during the previous one (software pipeline: every shuffle and
            // shared-memory load of step t+1 is issued at the top of step t, its latency hidden behind step t's arithmetic)
            double ph[4], rv[4], pt0 = 0.0, hl = 0.0, hr = 0.0, rhA = 0.0, rhB = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) ho[k] = bo[k] = last[k] = ph[k] = rv[k] = 0.0;
            r1[0] = r1[1] = r2[0] = r2[1] = r3[0] = r3[1] = 0.0;
            int xbits = 0; // max over the stored values and residual terms of their exponent bits: 0x7ff00000 = non-finite
            // every column of the warp is an interior column: only the row test can switch an update off
            const bool cols_ok = __all_sync(0xffffffffu, c0 && c1);
            // prologue of the pipeline: pass 0 of step 0 listens to relative row 0 (x of step 0 = 0 ^ parity(-1) = 1)
            if (RS_BULK) rs_mbar_wait(mbar, 0);
            else rs_wait<RS_D - 1>();
            __syncwarp();
            pt0 = Bp[1];
            // 16 steps = one turn of the ring.  FILL: the first turn, where the passes start one after the other.
            // A turn always runs all 16 steps (steps past T only touch the ring), so the body is straight-line code.
            auto turn = [&](const int tb, auto fill_c) {
                constexpr bool FILL = decltype(fill_c)::value;
#pragma unroll
                for (int u = 0; u < RS_R; ++u) {
                    const int t = tb + u; // tb is a multiple of 16: slot of relative row t-j is (u-j) & 15, its parity (u-j) & 1
                    if (RS_BULK) { // rows <= t + 1 have arrived (slot (u + 1) & 15 in its (t + 1) / 16-th use)
                        if (t + 1 < nload) rs_mbar_wait(mbar + ((u + 1) & (RS_R - 1)), ((t + 1) >> 4) & 1);
                    } else {
                        rs_wait<RS_D - 2>();
                    }
                    __syncwarp();
                    issue(t + RS_D, (u + RS_D) & (RS_R - 1));
                    // ---- fetch the operands of step t+1 (rows t+1-1-2k, parity of those rows = u & 1)
                    double phn[4], rvn[4], pt0n, hln, hrn, rhAn, rhBn;
                    {
                        const int xn0 = u & 1; // x of pass 0 / 2 in step t+1; passes 1 / 3: xn0 ^ 1
                        pt0n = Bp[((u + 1) & (RS_R - 1)) * RS_ROW + xn0]; // row t+1: above pass 0's point of step t+1 (row t)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int xn = (k & 1) ^ xn0;
                            rvn[k] = Rp[((u - 2 * k) & (RS_R - 1)) * RS_ROW + xn];
                            // ho of step t+1 is pt of this step: the loaded row for pass 0, else what pass k-1 left last step
                            const double hon = k == 0 ? pt0 : last[k - 1];
                            phn[k] = __shfl_sync(0xffffffffu, hon, lane + 2 * xn - 1);
                        }
                        // residual of step t+1 (row t-8): its centre row r2 is today's r1
                        hln = __shfl_sync(0xffffffffu, r1[1], lane - 1); // column a - 1: the lane below's pair
                        hrn = __shfl_sync(0xffffffffu, r1[0], lane + 1); // column a + 2
                        rhAn = Rp[((u - 8) & (RS_R - 1)) * RS_ROW], rhBn = Rp[((u - 8) & (RS_R - 1)) * RS_ROW + 1];
                    }
                    double cur[2] = {0.0, 0.0}; // the lane's pair of row t-7 after pass 3
                    double nv[4] = {0.0, 0.0, 0.0, 0.0};
                    // ---- four colour passes (pressure.cpp:70-94), pass k in relative row t-1-2k.  The p ring is READ-ONLY:
                    //      what a pass produces travels to the next pass in registers (last[] -> pt -> ho -> bo), and the
                    //      ring only serves pass 0 (the freshly loaded row above its point) and the points no pass may
                    //      update (ghost cells, rows outside the slab's light cone), whose value is the loaded one.
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rt = t - 1 - 2 * k;
                        if (FILL && rt < -1) continue; // warp-uniform (compile-time in the unrolled first turn)
                        const int q = k & 1;                          // colour this pass updates
                        const int S = (u - 1 - 2 * k) & (RS_R - 1);   // slot of row rt
                        const int x = q ^ ((u + 1) & 1);              // colour ^ parity of rt = column offset in the pair
                        // the point above this pass's point: the loaded row for pass 0, else what pass k-1 left there
                        const double pt = k == 0 ? pt0 : last[k - 1];
                        if (FILL && rt < 1) { // rows -1, 0: the pass only listens (its ho / bo fill up); row 0 is never updated
                            if (rt == 0) nv[k] = Bp[S * RS_ROW + x];
                            bo[k] = ho[k];
                            ho[k] = pt;
                            continue;
                        }
                        // the horizontal neighbour in the adjacent lane's pair is that lane's in-pair neighbour ho[k]
                        // (same pass, same row): ph[k], shuffled over during the previous step
                        const double sum = (ph[k] + ho[k]) * idx2 + (pt + bo[k]) * idy2;
                        double xv;
                        if (VC) { // kernels_vcycle.cuh, k_mg_colour: (sum - rhs) / (denom + mg_wall)
                            const int jj = ys + rt;
                            const double wall = wxc[x] + (double)((jj == 1) + (jj == g.ny)) * idy2;
                            const double num = sum - rv[k];
                            xv = POW2 ? num * g.inv_denom : num / g.denom; // wall = 0: denom + 0.0 is denom
                            if (wall != 0.0) xv = num / (g.denom + wall);
                        } else {
                            xv = POW2 ? (rv[k] + sum) * g.inv_denom : (rv[k] + sum) / g.denom;
                        }
                        const bool rows_ok = (!FILL || rt >= t_jlo) && rt <= t_jhi; // warp-uniform
                        if (!(rows_ok && cols_ok)) {
                            const bool ok = rows_ok && colok[x];
                            if (!ok) xv = Bp[S * RS_ROW + x]; // not updatable: the value that was loaded
                        }
                        nv[k] = xv;
                        if (k == 3) {                                  // row rt is final: the lane's pair, column order
                            cur[0] = x ? ho[k] : xv;
                            cur[1] = x ? xv : ho[k];
                            if (rt >= t_st0 && rt <= t_st1 && st_lane) {
                                if (!VC) xbits = max(xbits, max(__double2hiint(cur[0]) & 0x7ff00000, __double2hiint(cur[1]) & 0x7ff00000));
                                *reinterpret_cast<double2 *>(po) = make_double2(cur[0], cur[1]);
                            }
                        }
                        bo[k] = ho[k];
                        ho[k] = pt;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) last[k] = nv[k];
                    po += P;
                    // ---- residual (pressure.cpp:96-110) in relative row rr = t-9: r = rhs - L p, sum of r^2.  Rows
                    //      rr-1, rr, rr+1 of the lane's own columns are pass 3's last three results (registers).
                    const int rr = t - 9;
                    if (VC && finish != RS_VC_SMOOTH && (!FILL || rr >= rr_first) && rr <= rr_last) {
                        // kernels_vcycle.cuh, mg_resid: rhs - (L p - wall p), then k_mg_restrict / k_mg_residual_norm
                        const int jj = ys + rr;
                        const double wy = (double)((jj == 1) + (jj == g.ny)) * idy2;
                        double ApA = (r2[1] - 2.0 * r2[0] + hl) * idx2 + (r1[0] - 2.0 * r2[0] + r3[0]) * idy2;
                        double ApB = (hr - 2.0 * r2[1] + r2[0]) * idx2 + (r1[1] - 2.0 * r2[1] + r3[1]) * idy2;
                        ApA = ApA - (wxc[0] + wy) * r2[0];
                        ApB = ApB - (wxc[1] + wy) * r2[1];
                        const double resA = rhA - ApA, resB = rhB - ApB;
                        if (finish == RS_VC_NORM) {
                            acc[0] += inA ? resA * resA : 0.0;
                            acc[0] += inB ? resB * resB : 0.0;
                        } else {
                            const double srow = resA + resB;
                            if (!(jj & 1) && inA) { // second row of the block: coarse cell ((a + 1) / 2, jj / 2)
                                const size_t oc = (size_t)(jj / 2 + CFD_YOFF) * cz.P + ((a + 1) / 2 + CFD_XOFF);
                                cz.rhs[oc] = 0.25 * (sprev + srow);
                                cz.p[oc] = 0.0;
                            }
                            sprev = srow;
                        }
                    }
                    if (!VC && (!FILL || rr >= rr_first) && rr <= rr_last) {
                        const double ApA = (r2[1] - 2.0 * r2[0] + hl) * idx2 + (r1[0] - 2.0 * r2[0] + r3[0]) * idy2;
                        const double ApB = (hr - 2.0 * r2[1] + r2[0]) * idx2 + (r1[1] - 2.0 * r2[1] + r3[1]) * idy2;
                        const double resA = rhA - ApA, resB = rhB - ApB;
                        // (a finite term may square to Inf exactly as in the reference's dot(): test the term, not acc)
                        xbits = max(xbits, max((inA ? __double2hiint(resA) : 0) & 0x7ff00000, (inB ? __double2hiint(resB) : 0) & 0x7ff00000));
                        acc[0] += inA ? resA * resA : 0.0;
                        acc[0] += inB ? resB * resB : 0.0;
                    }
                    r3[0] = r2[0], r3[1] = r2[1];
                    r2[0] = r1[0], r2[1] = r1[1];
                    r1[0] = cur[0], r1[1] = cur[1];
                    // ---- the pipeline moves on
#pragma unroll
                    for (int k = 0; k < 4; ++k) ph[k] = phn[k], rv[k] = rvn[k];
                    pt0 = pt0n, hl = hln, hr = hrn, rhA = rhAn, rhB = rhBn;
                }
            };
            turn(0, std::true_type{});
            for (int tb = RS_R; tb < T; tb += RS_R) turn(tb, std::false_type{});
            if (xbits == 0x7ff00000) redo = 1;
            rs_wait<0>();
        }
    }
    if (done || (VC && finish != RS_VC_NORM)) return;
    if (VC) { // k_mg_residual_norm's tail
        grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
            const double res = sqrt(t[0]);
            sc->res = res;
            sc->iters = sc->iters + 1;
            if (res < tol) sc->done = 1;
        });
        return;
    }
    if (redo) *reinterpret_cast<volatile int *>(&sc->rb_redo) = 1;
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
        if (*reinterpret_cast<volatile int *>(&sc->rb_redo)) return; // the guarded kernel behind us redoes this smooth()
        if (finish) rbgs_tail(sc, t[0], tol);                        // vcycle loop exit, pressure.cpp:120
        else sc->red[0] = t[0];
    });
}
