// kernels_rbgs_fused.cuh -- PressureSolver::smooth (pressure.cpp:64-111) as ONE kernel: the four colour passes
// (2 sweeps x red, black) and the residual run on a tile kept in shared memory (temporal blocking).
//
// The per-colour kernels move every 32-byte sector of p and rhs once per colour pass: 4 x 24 + 16 = 112 B/cell per
// smooth() call (ncu: each pass already runs at 84 % of DRAM peak, so only less traffic helps).  Here a block loads
// a (116+12) x (54+10) tile of p, keeps rhs (and each point's distance to the tile edge) in registers, performs pass k on the tile grown by 5-k cells (so that
// after pass 4 the tile grown by 1 is final, which the residual stencil needs), and writes p once: ~24 B/cell plus
// halo re-reads that hit L2.  p is ping-ponged (neighbouring blocks still read the old values in their halos).
// Results are bit-identical to the sequential colour passes: every update sees exactly the operands the reference's
// pass would see; halo points are recomputed redundantly with the same expression.
//
// Shared layout: two arrays by checkerboard class (local (lx+ly)&1), each RB_H x RB_K, so the 4 neighbours of a
// point (all of the other class) are unit-stride accesses: no bank conflicts.  Shared memory holds fin0(value)
// (the reference guards every neighbour read with isfinite, pressure.cpp:84-88); the unsanitised value goes to
// global memory at a point's last update.
#pragma once
#include "kernels_pressure.cuh"

// Tile geometry.  A thread owns RB_M x RB_N slots (a slot = aligned pair of columns): rows ly = RB_M*warp + m,
// column pairs k = lane + 32 n.  Default: 128 x 64 tile, 512 threads (2 blocks/SM); -DRB_N=1: 64 x 64 tile, 256 threads.
#ifndef RB_N
#define RB_N 2
#endif
#ifndef RB_M
#define RB_M (RB_N == 2 ? 4 : 8)
#endif
#ifndef RB_WARPS
#define RB_WARPS (RB_N == 2 ? 16 : 8)
#endif
#define RB_HX 6            // x halo (5 needed; 6 keeps the tile origin on an even device column)
#define RB_HY 5            // y halo
#define RB_W (64 * RB_N)            // columns of the shared tile
#define RB_H (RB_M * RB_WARPS)      // rows of the shared tile
#define RB_TX (RB_W - 2 * RB_HX)    // output columns per tile (even: keeps double2 alignment)
#define RB_TY (RB_H - 2 * RB_HY)    // output rows per tile
#define RB_K (RB_W / 2)             // points per class per row
#define RB_THREADS (32 * RB_WARPS)
#define RB_MINBLOCKS (RB_N == 2 ? 2 : 4)
#define RB_PAD (RB_K + 8)             // doubles of padding before and after the two class arrays
#define RB_SMEM ((2 * RB_H * RB_K + 2 * RB_PAD) * 8)

// One update of the point (ly, k) with row parity LYP and column parity E during pass PS.  Everything that selects
// registers or shared-memory offsets is a compile-time constant; the only runtime test is the point's margin.
template <int POW2, int PS, int LYP, int E>
__device__ __forceinline__ void rb_update(const Geom &g, double *sm, int ly, int k, int mg, double rvv, double *pout,
                                          size_t o, int &pflag) {
    // Branch-free: the update is always computed (shared memory is padded, so the reads of points outside the pass
    // region are harmless) and only the stores are predicated -- a thread's 8 updates of a pass can then overlap.
    // pass PS updates the tile grown by 4-PS: margin (distance inside the output tile, negative outside) >= PS-4
    constexpr int CLS = (E + LYP) & 1; // checkerboard class of this point
    const double *oth = sm + ((CLS ^ 1) * RB_H + ly) * RB_K + k;
    const double pl = oth[E - 1], pr = oth[E], pb = oth[-RB_K], pt = oth[RB_K];
    const double sum = (pr + pl) * g.idx2 + (pt + pb) * g.idy2;
    const double x = POW2 ? (rvv + sum) * g.inv_denom : (rvv + sum) / g.denom;
    const bool fin = finite_d(x);
    if (mg >= PS - 4) sm[(CLS * RB_H + ly) * RB_K + k] = fin ? x : 0.0;
    if (PS >= 2 && mg >= 0) { // last update of a point of the output tile: the value smooth() leaves in p
        pout[o] = x;
        pflag |= fin ? 0 : 1;
    }
}

// PAR0 = colour of local class 0, i.e. (Xe + Ye + j0) & 1.  Xe + Ye is odd for every tile (RB_TX, RB_TY even,
// RB_HX + RB_HY odd), so PAR0 = (1 + j0) & 1 is one constant per launch.
template <int POW2, int PAR0>
__global__ void __launch_bounds__(RB_THREADS, RB_MINBLOCKS)
    k_rbgs_fused(Geom g, const double *__restrict__ pin, double *__restrict__ pout, const double *__restrict__ rhs,
                 Scalars *sc, double *partials, double tol, int ntx, int ntiles, int finish, int only_if_redo) { pdl_begin();
    // Launched behind k_rbgs_stream as its guarded slow path: nothing to do unless that kernel met a non-finite value.
    if (only_if_redo && !*reinterpret_cast<volatile int *>(&sc->rb_redo)) return;
    extern __shared__ double sm_raw[];
    double *sm = sm_raw + RB_PAD; // [2][RB_H][RB_K], padded on both sides
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool done = sc->done != 0;                    // converged earlier: only carry p over to the other buffer
    double acc[1] = {0.0};
    int pflag = 0;
    // One tile per block when launched with ntiles blocks (the stand-alone smoother); as the slow path behind
    // k_rbgs_stream the grid is two blocks per SM and each block walks over its tiles.
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int ty = tile / ntx, tx = tile - ty * ntx;
    const int X0 = 1 + tx * RB_TX, Y0 = 1 + ty * RB_TY; // first output cell (raw)
    const int Xe = X0 - RB_HX, Ye = Y0 - RB_HY;         // raw index of local (0,0)

    // ---- load: thread (warp, lane) owns slots (ly = 4*warp + m, k = lane + 32 n), m < 4, n < 2; a slot is the
    //      aligned pair of local columns (2k, 2k+1).  rhs stays in registers for all passes; mg = margin of a point:
    //      how far inside the output tile it is (negative = in the halo), -100 when it is not an interior cell.
    double rv[RB_M][RB_N][2]; // [m][n][e] : e = local column parity
    int mgx[RB_N][2], mgy[RB_M]; // margin = min(mgx[n][e], mgy[m])
    const size_t o00 = gidx(g, Xe + 2 * lane, Ye + RB_M * warp); // element offset of slot (m = 0, n = 0)
    // A tile whose halo stays inside the interior needs no ghost / domain logic at all (block-uniform branch).
    // Rows that may be updated: the slab's own cell rows plus, towards a neighbouring slab, the exchanged halo rows
    // (recomputed redundantly, never stored); rows that may be read: one more (the ghost row / outermost halo row).
    const int jlo = g.bottom ? 1 : 1 - RB_HY + 1, jhi = g.top ? g.ny : g.ny + RB_HY - 1;
    const bool edge = Xe < 1 || Xe + RB_W - 1 > g.nx || Ye < jlo || Ye + RB_H - 1 > jhi || done;
#pragma unroll
    for (int n = 0; n < RB_N; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int lx = 2 * (lane + 32 * n) + e, ci = Xe + lx;
            mgx[n][e] = (!edge || (ci >= 1 && ci <= g.nx)) ? min(lx - RB_HX, RB_HX + RB_TX - 1 - lx) : -100;
        }
#pragma unroll
    for (int m = 0; m < RB_M; ++m) {
        const int ly = RB_M * warp + m, jj = Ye + ly;
        // margin in y: distance inside the output tile AND inside the slab's own rows (halo rows of a neighbouring
        // slab are recomputed like tile halo, never stored or counted)
        mgy[m] = !edge ? min(ly - RB_HY, RB_HY + RB_TY - 1 - ly)
                       : (jj >= jlo && jj <= jhi) ? min(min(ly - RB_HY, RB_HY + RB_TY - 1 - ly), min(jj - 1, g.ny - jj)) : -100;
#pragma unroll
        for (int n = 0; n < RB_N; ++n) {
            const int k = lane + 32 * n, ii = Xe + 2 * k;
            const size_t o = o00 + (size_t)m * g.P + 64 * n;
            double2 pv = make_double2(0.0, 0.0), rr = make_double2(0.0, 0.0);
            if (!edge) {
                pv = *reinterpret_cast<const double2 *>(pin + o);
                rr = *reinterpret_cast<const double2 *>(rhs + o);
            } else if (jj >= jlo - 1 && jj <= jhi + 1 && ii >= -1 && ii <= g.nx + 1) { // pair touches the readable box
                pv = *reinterpret_cast<const double2 *>(pin + o);
                rr = *reinterpret_cast<const double2 *>(rhs + o);
                // ghosts never change during a smooth: carry them (and, when already converged, everything this
                // tile owns) over to the output buffer.  Owner = the tile that holds the nearest interior cell.
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ci = ii + e;
                    if (ci < 0 || ci > g.nx + 1) continue;
                    if (jj < 0 || jj > g.ny + 1) continue; // deeper halo rows belong to the neighbouring slab
                    const bool ghost = ci == 0 || ci == g.nx + 1 || (g.bottom && jj == 0) || (g.top && jj == g.ny + 1);
                    if (!ghost && !done) continue;
                    const int ni = min(max(ci, 1), g.nx), nj = min(max(jj, 1), g.ny);
                    const bool mine = ni >= X0 && ni < X0 + RB_TX && nj >= Y0 && nj < Y0 + RB_TY;
                    if (mine) pout[o + e] = e ? pv.y : pv.x;
                }
            }
            const int e0 = m & 1; // class of the even column of this row (ly parity = m parity)
            sm[(e0 * RB_H + ly) * RB_K + k] = fin0(pv.x);
            sm[((e0 ^ 1) * RB_H + ly) * RB_K + k] = fin0(pv.y);
            rv[m][n][0] = fin0(rr.x);
            rv[m][n][1] = fin0(rr.y);
        }
    }
    if (done) continue;
    __syncthreads();

    // ---- four colour passes (pressure.cpp:70-94): pass ps updates colour ps & 1 = class (ps & 1) ^ PAR0, i.e. in
    //      a row of parity LYP the column parity E = class ^ LYP
#define RB_PASS(PS)                                                                                                  \
    {                                                                                                                \
        constexpr int CLS = ((PS) & 1) ^ PAR0;                                                                       \
        constexpr int E0 = CLS, E1 = CLS ^ 1; /* column parity of the class in even / odd rows */                    \
        _Pragma("unroll") for (int m2 = 0; m2 < RB_M / 2; ++m2) {                                                    \
            _Pragma("unroll") for (int n = 0; n < RB_N; ++n) {                                                       \
                const int k = lane + 32 * n, ly = RB_M * warp + 2 * m2;                                              \
                const size_t on = o00 + 64 * n + (size_t)(2 * m2) * g.P;                                             \
                rb_update<POW2, PS, 0, E0>(g, sm, ly, k, min(mgx[n][E0], mgy[2 * m2]), rv[2 * m2][n][E0], pout,      \
                                           on + E0, pflag);                                                          \
                rb_update<POW2, PS, 1, E1>(g, sm, ly + 1, k, min(mgx[n][E1], mgy[2 * m2 + 1]), rv[2 * m2 + 1][n][E1], \
                                           pout, on + g.P + E1, pflag);                                              \
            }                                                                                                        \
        }                                                                                                            \
        __syncthreads();                                                                                             \
    }
    RB_PASS(0)
    RB_PASS(1)
    RB_PASS(2)
    RB_PASS(3)
#undef RB_PASS

    // ---- residual (pressure.cpp:96-110): r = fin0(rhs) - fin0(L p), sum of fin0(r)^2 over the output tile
#pragma unroll
    for (int m = 0; m < RB_M; ++m) {
        const int ly = RB_M * warp + m;
#pragma unroll
        for (int n = 0; n < RB_N; ++n) {
            const int k = lane + 32 * n;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool inside = min(mgx[n][e], mgy[m]) >= 0;
                const int c = (e + m) & 1; // class of this point
                const double *own = sm + (c * RB_H + ly) * RB_K + k, *oth = sm + ((c ^ 1) * RB_H + ly) * RB_K + k;
                const double cc = own[0], l = oth[e - 1], r = oth[e], b = oth[-RB_K], t = oth[RB_K];
                const double Ap = (r - 2.0 * cc + l) * g.idx2 + (t - 2.0 * cc + b) * g.idy2;
                double res = rv[m][n][e] - Ap;
                if (!finite_d(res)) res = fin0(rv[m][n][e] - fin0(Ap)); // fin0(rhs) - fin0(Ap), then dot()'s guard
                acc[0] += inside ? res * res : 0.0;
            }
        }
    }
    __syncthreads(); // the next tile overwrites the shared tile
    } // tiles
    if (done) return;
    if (pflag) sc->nonfinite_p = 1;
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
        sc->rb_redo = 0;
        if (finish) rbgs_tail(sc, t[0], tol); // vcycle loop exit, pressure.cpp:120
        else sc->red[0] = t[0];
    });
}
