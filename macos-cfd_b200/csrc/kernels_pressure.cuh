// kernels_pressure.cuh -- PressureSolver::solve (pressure.cpp:113-242) on the device.
//
//  * Jacobi-PCG: the reference's 9 loops per iteration (Laplacian, 3 dots, 2 axpys, preconditioner, direction
//    update, pin) become TWO kernels with one grid-level reduction each; alpha, beta, the residual test, the
//    breakdown test and the iteration counter live in device memory (struct Scalars), so a solve is a
//    fixed sequence of launches with no host round trip -- kernels after convergence return at once.
//  * "Multigrid" in reference-parity mode is what the reference does: fine-grid red-black Gauss-Seidel sweeps
//    followed by a residual (pressure.cpp:64-129).
#pragma once
#include "common.cuh"

// Tile of the pressure kernels: 64 cells x 16 rows per pass, 256 threads, each thread 2 adjacent cells
// (one aligned double2) in 2 rows.  Blocks loop over tiles (grid <= a few waves of 148 SMs).
#define PT_X 64
#define PT_Y 16
#define PT_S (PT_X + 2 + 1) // smem row stride (odd)
#define PT_THREADS 256

// Laplacian of pressure.cpp:20-44 on smem values around c (row stride S), finite guards included.
__device__ __forceinline__ double lap5(const double *c, int S, double idx2, double idy2) {
    double cc = fin0(c[0]), l = fin0(c[-1]), r = fin0(c[1]), b = fin0(c[-S]), t = fin0(c[S]);
    return (r - 2.0 * cc + l) * idx2 + (t - 2.0 * cc + b) * idy2;
}

struct PcgTile {
    int x0, y0; // raw index of the tile's first cell
};
__device__ __forceinline__ PcgTile pcg_tile(const Geom &g, int t, int ntx) {
    PcgTile T;
    int ty = t / ntx, tx = t - ty * ntx;
    T.x0 = 1 + tx * PT_X;
    T.y0 = 1 + ty * PT_Y;
    return T;
}
// rows of s that are live data: cell rows, plus the halo row towards a neighbouring slab
__device__ __forceinline__ bool s_live(const Geom &g, int ii, int jj) {
    return ii >= 1 && ii <= g.nx && jj >= (g.bottom ? 1 : 0) && jj <= (g.top ? g.ny : g.ny + 1);
}

// ---- scalar tails of the solver (thread 0 of the last block on one GPU; a one-thread kernel after the allreduce on
//      row slabs, where every rank evaluates them on the same reduced numbers) ----------------------------------
// after PCG start: rho = dot(r,z), res = sqrt(max(0, dot(r,r))), loop condition (pressure.cpp:175-179)
__device__ __forceinline__ void pcg_tail_init(Scalars *sc, double rz, double rr, double tol, int max_iters) {
    sc->rho = rz;
    sc->first = 1;
    sc->iters = 0;
    sc->breakdown_iter = -1;
    double res = sqrt(max2(0.0, rr));
    sc->res = res;
    sc->done = !(max_iters > 0 && res > tol);
}
// after dot(s, Ap): breakdown test and alpha (pressure.cpp:182-187)
__device__ __forceinline__ void pcg_tail1(Scalars *sc, double denom) {
    if (!finite_d(denom) || fabs(denom) < 1e-20) {
        sc->breakdown_iter = sc->iters;
        sc->done = 1;
    } else {
        sc->alpha = sc->rho / denom;
    }
}
// after the p/r update: residual test (:202-204), rho_new / beta (:216-219, :229), loop condition (:179)
__device__ __forceinline__ void pcg_tail2(Scalars *sc, double rr, double rz, double tol, int max_iters) {
    double res = sqrt(max2(0.0, rr));
    sc->res = res;
    sc->first = 0;
    int it = sc->iters + 1;
    sc->iters = it;
    if (!finite_d(res) || res < tol) {
        sc->done = 1;
    } else if (!finite_d(rz)) {
        sc->done = 1;
    } else {
        sc->beta = rz / sc->rho;
        sc->rho = rz;
        if (!(it < max_iters && res > tol)) sc->done = 1;
    }
}
// after smooth(): res = sqrt(dot(r,r)) and the vcycle loop exit (pressure.cpp:110, :118-122)
__device__ __forceinline__ void rbgs_tail(Scalars *sc, double rr, double tol) {
    double res = sqrt(rr);
    sc->res = res;
    sc->iters = sc->iters + 1;
    if (res < tol) sc->done = 1;
}
__global__ void k_pcg_fin_init(Scalars *sc, double tol, int max_iters) { pdl_begin(); pcg_tail_init(sc, sc->red[0], sc->red[1], tol, max_iters); }
__global__ void k_pcg_fin1(Scalars *sc) { pdl_begin(); if (!sc->done) pcg_tail1(sc, sc->red[0]); }
__global__ void k_pcg_fin2(Scalars *sc, double tol, int max_iters) { pdl_begin(); if (!sc->done) pcg_tail2(sc, sc->red[0], sc->red[1], tol, max_iters); }
__global__ void k_rbgs_fin(Scalars *sc, double tol) { pdl_begin(); if (!sc->done) rbgs_tail(sc, sc->red[0], tol); }

// PCG start (pressure.cpp:133-177): p = 0 everywhere (done by a memset before this kernel), Ap = L(0) = 0,
// r = fin0(rhs) - 0, rho = dot(r, inv_diag*r), res = sqrt(max(0, dot(r,r))).  z and s are not stored: the
// first direction update recomputes z = inv_diag*r (flag `first`).
__global__ void __launch_bounds__(PT_THREADS)
    k_pcg_init(Geom g, const double *__restrict__ rhs, double *__restrict__ r, Scalars *sc, double *partials,
               double tol, int max_iters, int finish) { pdl_begin();
    double acc[2] = {0.0, 0.0};
    const int c2 = threadIdx.x; // 256 threads = 512 cells per row chunk
    for (int jj = 1 + blockIdx.y; jj <= g.ny; jj += gridDim.y) {
        for (int ii = 1 + 2 * (blockIdx.x * PT_THREADS + c2); ii <= g.nx; ii += 2 * PT_THREADS * gridDim.x) {
            const size_t o = gidx(g, ii, jj);
            double2 a = *reinterpret_cast<const double2 *>(rhs + o);
            double Ap = 0.0 * g.idx2 + 0.0 * g.idy2; // L(p = 0)
            double r0 = fin0(a.x) - fin0(Ap);
            double z0 = g.inv_diag * r0;
            acc[0] += fin0(r0) * fin0(z0);
            acc[1] += fin0(r0) * fin0(r0);
            if (ii + 1 <= g.nx) {
                double r1 = fin0(a.y) - fin0(Ap);
                double z1 = g.inv_diag * r1;
                acc[0] += fin0(r1) * fin0(z1);
                acc[1] += fin0(r1) * fin0(r1);
                *reinterpret_cast<double2 *>(r + o) = make_double2(r0, r1);
            } else {
                r[o] = r0;
            }
        }
    }
    grid_sum<2>(acc, partials, &sc->ticket[2], [=](double(&t)[2]) {
        if (finish) pcg_tail_init(sc, t[0], t[1], tol, max_iters);
        else { sc->red[0] = t[0]; sc->red[1] = t[1]; }
    });
}

// Phase 1 of an iteration: direction update + Laplacian + dot(s, Ap)   (pressure.cpp:207-229 of the previous
// iteration, then :181-182).   s_new = z + beta*s_old with z = inv_diag*r (z is never stored); the first
// iteration takes s = z (pressure.cpp:170-171).  s is double-buffered because neighbouring tiles read the old
// direction in their halo.  Ap is not stored either: phase 2 recomputes it from s_new (same expression, same
// inputs => same bits), which saves 16 B/cell of HBM traffic per iteration.
__global__ void __launch_bounds__(PT_THREADS)
    k_pcg_phase1(Geom g, const double *__restrict__ r, const double *__restrict__ s_old, double *__restrict__ s_new,
                 Scalars *sc, double *partials, int ntx, int ntiles, int finish) { pdl_begin();
    if (sc->done) return;
    __shared__ double sm[(PT_Y + 2) * PT_S];
    const bool first = sc->first != 0;
    const double beta = sc->beta;
    const int tid = threadIdx.x;
    const int lx = tid & 31, ly = tid >> 5; // 32 x 8 threads
    const int jlo = g.bottom ? 1 : 0, jhi = g.top ? g.ny : g.ny + 1;
    double acc[1] = {0.0};
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const PcgTile T = pcg_tile(g, t, ntx);
        // interior of the tile: 2 cells x 2 rows per thread
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int jj = T.y0 + ly + 8 * rr, ii = T.x0 + 2 * lx;
            double2 sn = make_double2(0.0, 0.0);
            if (jj >= jlo && jj <= jhi && ii <= g.nx) { // live row: cell row, or the halo row of a slab
                const size_t o = gidx(g, ii, jj);
                double2 rv = *reinterpret_cast<const double2 *>(r + o);
                double zx = g.inv_diag * rv.x, zy = g.inv_diag * rv.y;
                if (first) {
                    sn = make_double2(zx, zy);
                } else {
                    double2 so = *reinterpret_cast<const double2 *>(s_old + o);
                    sn = make_double2(zx + beta * so.x, zy + beta * so.y);
                }
                if (ii + 1 <= g.nx) {
                    *reinterpret_cast<double2 *>(s_new + o) = sn;
                } else { // ghost column stays 0
                    sn.y = 0.0;
                    s_new[o] = sn.x;
                }
            }
            double *d = sm + (ly + 8 * rr + 1) * PT_S + 2 * lx + 1;
            d[0] = sn.x;
            d[1] = sn.y;
        }
        // halo ring: bottom row, top row, left column, right column (2*64 + 2*16 = 160 cells)
        if (tid < 2 * PT_X + 2 * PT_Y) {
            int hx, hy;
            if (tid < PT_X) { hx = tid; hy = -1; }
            else if (tid < 2 * PT_X) { hx = tid - PT_X; hy = PT_Y; }
            else if (tid < 2 * PT_X + PT_Y) { hx = -1; hy = tid - 2 * PT_X; }
            else { hx = PT_X; hy = tid - 2 * PT_X - PT_Y; }
            const int ii = T.x0 + hx, jj = T.y0 + hy;
            double sn = 0.0;
            if (s_live(g, ii, jj)) {
                const size_t o = gidx(g, ii, jj);
                double z = g.inv_diag * r[o];
                sn = first ? z : z + beta * s_old[o];
                // a slab's halo row of s_new is recomputed here from the exchanged halo row of r
                if ((jj == 0 || jj == g.ny + 1) && hx >= 0 && hx < PT_X) s_new[o] = sn;
            }
            sm[(hy + 1) * PT_S + hx + 1] = sn;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int jj = T.y0 + ly + 8 * rr, ii = T.x0 + 2 * lx;
            if (jj <= g.ny) {
                const double *c = sm + (ly + 8 * rr + 1) * PT_S + 2 * lx + 1;
                if (ii <= g.nx) acc[0] += fin0(c[0]) * fin0(lap5(c, PT_S, g.idx2, g.idy2));
                if (ii + 1 <= g.nx) acc[0] += fin0(c[1]) * fin0(lap5(c + 1, PT_S, g.idx2, g.idy2));
            }
        }
        __syncthreads();
    }
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
        if (finish) pcg_tail1(sc, t[0]);
        else sc->red[0] = t[0];
    });
}

// Phase 2: p += alpha*s, r -= alpha*Ap, pin, dot(r,r), dot(r, inv_diag*r)   (pressure.cpp:187-218), then the
// scalar tail of the iteration (residual test :202-204, beta :219, loop condition :179) in the last block.
__global__ void __launch_bounds__(PT_THREADS)
    k_pcg_phase2(Geom g, double *__restrict__ p, double *__restrict__ r, const double *__restrict__ s, Scalars *sc,
                 double *partials, int ntx, int ntiles, double tol, int max_iters, int finish) { pdl_begin();
    if (sc->done) return;
    __shared__ double sm[(PT_Y + 2) * PT_S];
    const double alpha = sc->alpha;
    const int tid = threadIdx.x;
    const int lx = tid & 31, ly = tid >> 5;
    double acc[2] = {0.0, 0.0};
    int pflag = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const PcgTile T = pcg_tile(g, t, ntx);
        double2 sv[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int jj = T.y0 + ly + 8 * rr, ii = T.x0 + 2 * lx;
            sv[rr] = make_double2(0.0, 0.0);
            if (jj <= g.ny + 1 && ii <= g.nx + 1) sv[rr] = *reinterpret_cast<const double2 *>(s + gidx(g, ii, jj));
            double *d = sm + (ly + 8 * rr + 1) * PT_S + 2 * lx + 1;
            d[0] = sv[rr].x;
            d[1] = sv[rr].y;
        }
        if (tid < 2 * PT_X + 2 * PT_Y) {
            int hx, hy;
            if (tid < PT_X) { hx = tid; hy = -1; }
            else if (tid < 2 * PT_X) { hx = tid - PT_X; hy = PT_Y; }
            else if (tid < 2 * PT_X + PT_Y) { hx = -1; hy = tid - 2 * PT_X; }
            else { hx = PT_X; hy = tid - 2 * PT_X - PT_Y; }
            const int ii = T.x0 + hx, jj = T.y0 + hy;
            double x = 0.0;
            if (ii <= g.nx + 1 && jj <= g.ny + 1) x = s[gidx(g, ii, jj)]; // ghosts of s are 0 in memory
            sm[(hy + 1) * PT_S + hx + 1] = x;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int jj = T.y0 + ly + 8 * rr, ii = T.x0 + 2 * lx;
            if (jj <= g.ny && ii <= g.nx) {
                const size_t o = gidx(g, ii, jj);
                const double *c = sm + (ly + 8 * rr + 1) * PT_S + 2 * lx + 1;
                double2 pv = *reinterpret_cast<const double2 *>(p + o);
                double2 rv = *reinterpret_cast<const double2 *>(r + o);
                double Ap0 = lap5(c, PT_S, g.idx2, g.idy2);
                pv.x = pv.x + alpha * c[0];
                rv.x = rv.x - alpha * Ap0;
                if (g.j0 == 0 && jj == 1 && ii == 1) pv.x = 0.0; // re-pin p(0,0), pressure.cpp:200
                if (!finite_d(pv.x)) pflag = 1;
                double z0 = g.inv_diag * rv.x;
                acc[0] += fin0(rv.x) * fin0(rv.x);
                acc[1] += fin0(rv.x) * fin0(z0);
                if (ii + 1 <= g.nx) {
                    double Ap1 = lap5(c + 1, PT_S, g.idx2, g.idy2);
                    pv.y = pv.y + alpha * c[1];
                    rv.y = rv.y - alpha * Ap1;
                    if (!finite_d(pv.y)) pflag = 1;
                    double z1 = g.inv_diag * rv.y;
                    acc[0] += fin0(rv.y) * fin0(rv.y);
                    acc[1] += fin0(rv.y) * fin0(z1);
                    *reinterpret_cast<double2 *>(p + o) = pv;
                    *reinterpret_cast<double2 *>(r + o) = rv;
                } else {
                    p[o] = pv.x;
                    r[o] = rv.x;
                }
            }
        }
        __syncthreads();
    }
    if (pflag) sc->nonfinite_p = 1;
    grid_sum<2>(acc, partials, &sc->ticket[2], [=](double(&t)[2]) {
        if (finish) pcg_tail2(sc, t[0], t[1], tol, max_iters);
        else { sc->red[0] = t[0]; sc->red[1] = t[1]; }
    });
}

// last_residual_ = isfinite(res) ? res : 1e9 (pressure.cpp:123, :232)
__global__ void k_solve_finish(Scalars *sc) { pdl_begin();
    double res = sc->res;
    if (!finite_d(res)) {
        sc->nonfinite_res = 1;
        sc->res = 1e9;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Red-black Gauss-Seidel colour pass (pressure.cpp:70-94), one thread per point of the colour.
//   p = (rhs + (pr+pl)*idx2 + (pt+pb)*idy2) / denom ;  colour = (i+j)%2 on 0-based GLOBAL indices, colour 0 first.
// When denom is a power of two (every power-of-two grid on a power-of-two domain) the division is replaced by
// a multiplication with its exactly representable reciprocal: both are the correctly rounded quotient.
// ---------------------------------------------------------------------------------------------------------
template <int POW2>
__global__ void __launch_bounds__(256)
    k_rbgs_colour(Geom g, double *p, const double *__restrict__ rhs, int colour, const Scalars *sc) { pdl_begin();
    if (sc->done) return;
    const int jj = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (jj > g.ny) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int ii = 1 + 2 * k + ((g.j0 + jj - 1 + colour) & 1);
    if (ii > g.nx) return;
    const size_t o = gidx(g, ii, jj);
    double rv = fin0(rhs[o]);
    double pl = fin0(p[o - 1]), pr = fin0(p[o + 1]), pb = fin0(p[o - g.P]), pt = fin0(p[o + g.P]);
    double sum = (pr + pl) * g.idx2 + (pt + pb) * g.idy2;
    p[o] = POW2 ? (rv + sum) * g.inv_denom : (rv + sum) / g.denom;
}

// Residual of smooth() (pressure.cpp:96-110): Ap = L p, r = fin0(rhs) - fin0(Ap), res = sqrt(dot(r,r)); Ap and r
// are the solver's private scratch in the reference and are not stored here.  The last block also evaluates
// the `res < tol` test of the vcycle loop (pressure.cpp:118-122).
__global__ void __launch_bounds__(256)
    k_rbgs_residual(Geom g, const double *__restrict__ p, const double *__restrict__ rhs, Scalars *sc,
                    double *partials, double tol, int finish) { pdl_begin();
    if (sc->done) return;
    const int ii = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    int pflag = 0;
    if (ii <= g.nx) {
        for (int jj = 1 + blockIdx.y; jj <= g.ny; jj += gridDim.y) {
            const size_t o = gidx(g, ii, jj);
            double c = p[o];
            if (!finite_d(c)) pflag = 1;
            double cc = fin0(c), l = fin0(p[o - 1]), r = fin0(p[o + 1]), b = fin0(p[o - g.P]), t = fin0(p[o + g.P]);
            double Ap = (r - 2.0 * cc + l) * g.idx2 + (t - 2.0 * cc + b) * g.idy2;
            double rr = fin0(rhs[o]) - fin0(Ap);
            acc[0] += fin0(rr) * fin0(rr);
        }
    }
    if (pflag) sc->nonfinite_p = 1;
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
        if (finish) rbgs_tail(sc, t[0], tol);
        else sc->red[0] = t[0];
    });
}

__global__ void k_solve_begin(Scalars *sc) { pdl_begin();
    sc->done = 0;
    sc->iters = 0;
    sc->res = 1e9; // "double res = 1e9" (pressure.cpp:117) survives vcycles == 0
    sc->breakdown_iter = -1;
}
