// kernels_pcg_march.cuh -- streaming versions of the two PCG kernels (same arithmetic as kernels_pressure.cuh).
//
// ncu on the tile versions (profiles/r01_pcg_tiles.txt): ideal DRAM bytes, but stalled on their own loads
// (long_scoreboard 10-15 warps per issue, load -> barrier -> compute -> barrier per tile, 50 % occupancy).
// Here a thread owns two adjacent columns (one aligned double2 per row) and marches up a strip of rows: the three rows
// of the direction s that the 5-point stencil needs live in registers, the x-neighbours come from the adjacent lanes
// by shuffle (lanes 0 and 31 carry one halo column each: 62 of 64 columns produce output), and the rows of the next
// iteration are requested before the current one is used.  No shared memory, no barriers.
#pragma once
#include "kernels_pressure.cuh"

#ifndef PM_WARPS
#define PM_WARPS 4
#endif
#define PM_COLS 62
#ifndef PM_PF
#define PM_PF 2 // rows in flight per thread (register ring); measured on B200 at 8192^2: PF 1/2/3/4 -> 170.6/164.8/175.6/178.6 ms per 200 iterations
#endif

__device__ __forceinline__ double pm_dn(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double pm_up(double x) { return __shfl_up_sync(0xffffffffu, x, 1); }

// pressure.cpp:20-44 on register values (finite guards included)
__device__ __forceinline__ double lap5r(double c, double l, double r, double b, double t, double idx2, double idy2) {
    double cc = fin0(c), ll = fin0(l), rr = fin0(r), bb = fin0(b), tt = fin0(t);
    return (rr - 2.0 * cc + ll) * idx2 + (tt - 2.0 * cc + bb) * idy2;
}

struct PmGeom {
    int a;          // raw column of this thread's first column (odd => aligned double2); second = a + 1
    bool out_a, out_b, loadable;
};
__device__ __forceinline__ bool pm_setup(const Geom &g, PmGeom &t) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * PM_WARPS + (threadIdx.x >> 5);
    const int c0 = PM_COLS * w; // this warp's output columns: [c0, c0 + 62) intersected with [1, nx]
    if (c0 > g.nx) return false;
    t.a = c0 - 1 + 2 * lane;
    t.out_a = lane >= 1 && t.a >= 1 && t.a <= g.nx;
    t.out_b = lane <= 30 && t.a + 1 >= 1 && t.a + 1 <= g.nx;
    t.loadable = t.a + 1 <= g.P - CFD_XOFF - 1;
    return true;
}

// Phase 1: s_new = z + beta*s_old (z = inv_diag*r; first iteration: s_new = z), dot(s_new, L s_new).
__global__ void __launch_bounds__(32 * PM_WARPS)
    k_pcg_phase1_march(Geom g, const double *__restrict__ r, const double *__restrict__ s_old,
                       double *__restrict__ s_new, Scalars *sc, double *partials, int RY, int finish) { pdl_begin();
    if (sc->done) return;
    PmGeom t;
    double acc[1] = {0.0};
    if (pm_setup(g, t)) {
        const bool first = sc->first != 0;
        const double beta = sc->beta;
        const int jlo = g.bottom ? 1 : 0, jhi = g.top ? g.ny : g.ny + 1; // rows on which s is live data
        const int y0 = 1 + blockIdx.y * RY, y1 = min(y0 + RY - 1, g.ny);
        const size_t P = g.P;
        const double *pr = r + gidx(g, t.a, 0), *ps = s_old + gidx(g, t.a, 0);
        double *pw = s_new + gidx(g, t.a, 0);
        const bool col_a = t.a >= 1 && t.a <= g.nx, col_b = t.a + 1 >= 1 && t.a + 1 <= g.nx; // ghost columns stay 0
        auto ld = [&](const double *base, int jj) {
            return t.loadable ? *reinterpret_cast<const double2 *>(base + (size_t)jj * P) : make_double2(0.0, 0.0);
        };
        auto direction = [&](int jj, double2 rv, double2 so) { // s_new of row jj (0 outside the live region)
            double2 sn = make_double2(0.0, 0.0);
            if (jj >= jlo && jj <= jhi) {
                double zx = g.inv_diag * rv.x, zy = g.inv_diag * rv.y;
                sn.x = first ? zx : zx + beta * so.x;
                sn.y = first ? zy : zy + beta * so.y;
                if (!col_a) sn.x = 0.0;
                if (!col_b) sn.y = 0.0;
            }
            return sn;
        };
        auto store = [&](int jj, double2 sn) {
            if (jj < jlo || jj > jhi) return;
            if (t.out_a && t.out_b) *reinterpret_cast<double2 *>(pw + (size_t)jj * P) = sn;
            else if (t.out_a) pw[(size_t)jj * P] = sn.x;
            else if (t.out_b) pw[(size_t)jj * P + 1] = sn.y;
        };
        // prologue: s_new of rows y0-1 and y0.  The first strip also owns the halo row below (row 0 of a slab).
        double2 sm1 = direction(y0 - 1, ld(pr, y0 - 1), first ? make_double2(0.0, 0.0) : ld(ps, y0 - 1));
        double2 s0 = direction(y0, ld(pr, y0), first ? make_double2(0.0, 0.0) : ld(ps, y0));
        if (blockIdx.y == 0) store(y0 - 1, sm1);
        store(y0, s0);
        // register ring: rows y0+1 .. y0+PM_PF are in flight; slot k is refilled right after it is consumed
        double2 nr[PM_PF], ns[PM_PF];
#pragma unroll
        for (int k = 0; k < PM_PF; ++k) {
            nr[k] = ld(pr, min(y0 + 1 + k, g.ny + 1));
            ns[k] = first ? make_double2(0.0, 0.0) : ld(ps, min(y0 + 1 + k, g.ny + 1));
        }
#pragma unroll 1
        for (int jb = y0; jb <= y1; jb += PM_PF) {
#pragma unroll
            for (int k = 0; k < PM_PF; ++k) {
                const int jj = jb + k;
                if (jj > y1) break;
                const double2 sp1 = direction(jj + 1, nr[k], ns[k]);
                const int nxt = min(jj + 1 + PM_PF, g.ny + 1); // clamped: rows past the slab are never used
                nr[k] = ld(pr, nxt);
                if (!first) ns[k] = ld(ps, nxt);
                if (jj < y1 || jj + 1 == g.ny + 1) store(jj + 1, sp1); // rows of the next strip are written by that strip
                const double left = pm_up(s0.y), right = pm_dn(s0.x);  // s(a-1, jj), s(a+2, jj)
                const double Apa = lap5r(s0.x, left, s0.y, sm1.x, sp1.x, g.idx2, g.idy2);
                const double Apb = lap5r(s0.y, s0.x, right, sm1.y, sp1.y, g.idx2, g.idy2);
                if (t.out_a) acc[0] += fin0(s0.x) * fin0(Apa);
                if (t.out_b) acc[0] += fin0(s0.y) * fin0(Apb);
                sm1 = s0;
                s0 = sp1;
            }
        }
    }
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&tt)[1]) {
        if (finish) pcg_tail1(sc, tt[0]);
        else sc->red[0] = tt[0];
    });
}

// Phase 2: p += alpha*s, r -= alpha*(L s), pin, dot(r,r), dot(r, inv_diag*r) + the scalar tail.
__global__ void __launch_bounds__(32 * PM_WARPS)
    k_pcg_phase2_march(Geom g, double *__restrict__ p, double *__restrict__ r, const double *__restrict__ s,
                       Scalars *sc, double *partials, int RY, double tol, int max_iters, int finish) { pdl_begin();
    if (sc->done) return;
    PmGeom t;
    double acc[2] = {0.0, 0.0};
    int pflag = 0;
    if (pm_setup(g, t)) {
        const double alpha = sc->alpha;
        const int y0 = 1 + blockIdx.y * RY, y1 = min(y0 + RY - 1, g.ny);
        const size_t P = g.P;
        const double *ps = s + gidx(g, t.a, 0);
        double *pp = p + gidx(g, t.a, 0), *pr = r + gidx(g, t.a, 0);
        const bool any = t.out_a || t.out_b;
        auto ld = [&](const double *base, int jj) {
            return t.loadable ? *reinterpret_cast<const double2 *>(base + (size_t)jj * P) : make_double2(0.0, 0.0);
        };
        double2 sm1 = ld(ps, y0 - 1), s0 = ld(ps, y0); // ghost rows / columns of s are 0 in memory
        double2 nsp[PM_PF], np[PM_PF], nr[PM_PF]; // register ring: s of rows y0+1.., p and r of rows y0..
#pragma unroll
        for (int k = 0; k < PM_PF; ++k) {
            nsp[k] = ld(ps, min(y0 + 1 + k, g.ny + 1));
            np[k] = any ? ld(pp, min(y0 + k, g.ny)) : make_double2(0.0, 0.0);
            nr[k] = any ? ld(pr, min(y0 + k, g.ny)) : make_double2(0.0, 0.0);
        }
#pragma unroll 1
        for (int jb = y0; jb <= y1; jb += PM_PF) {
#pragma unroll
            for (int k = 0; k < PM_PF; ++k) {
                const int jj = jb + k;
                if (jj > y1) break;
                const double2 sp1 = nsp[k];
                double2 pv = np[k], rv = nr[k];
                nsp[k] = ld(ps, min(jj + 1 + PM_PF, g.ny + 1));
                if (any) {
                    np[k] = ld(pp, min(jj + PM_PF, g.ny));
                    nr[k] = ld(pr, min(jj + PM_PF, g.ny));
                }
                const double left = pm_up(s0.y), right = pm_dn(s0.x);
                const double Apa = lap5r(s0.x, left, s0.y, sm1.x, sp1.x, g.idx2, g.idy2);
                const double Apb = lap5r(s0.y, s0.x, right, sm1.y, sp1.y, g.idx2, g.idy2);
                if (t.out_a) {
                    pv.x = pv.x + alpha * s0.x;
                    rv.x = rv.x - alpha * Apa;
                    if (g.j0 == 0 && jj == 1 && t.a == 1) pv.x = 0.0; // re-pin p(0,0), pressure.cpp:200
                    if (!finite_d(pv.x)) pflag = 1;
                    acc[0] += fin0(rv.x) * fin0(rv.x);
                    acc[1] += fin0(rv.x) * fin0(g.inv_diag * rv.x);
                }
                if (t.out_b) {
                    pv.y = pv.y + alpha * s0.y;
                    rv.y = rv.y - alpha * Apb;
                    if (g.j0 == 0 && jj == 1 && t.a + 1 == 1) pv.y = 0.0;
                    if (!finite_d(pv.y)) pflag = 1;
                    acc[0] += fin0(rv.y) * fin0(rv.y);
                    acc[1] += fin0(rv.y) * fin0(g.inv_diag * rv.y);
                }
                const size_t o = (size_t)jj * P;
                if (t.out_a && t.out_b) {
                    *reinterpret_cast<double2 *>(pp + o) = pv;
                    *reinterpret_cast<double2 *>(pr + o) = rv;
                } else if (t.out_a) {
                    pp[o] = pv.x;
                    pr[o] = rv.x;
                } else if (t.out_b) {
                    pp[o + 1] = pv.y;
                    pr[o + 1] = rv.y;
                }
                sm1 = s0;
                s0 = sp1;
            }
        }
    }
    if (pflag) sc->nonfinite_p = 1;
    grid_sum<2>(acc, partials, &sc->ticket[2], [=](double(&tt)[2]) {
        if (finish) pcg_tail2(sc, tt[0], tt[1], tol, max_iters);
        else { sc->red[0] = tt[0]; sc->red[1] = tt[1]; }
    });
}
