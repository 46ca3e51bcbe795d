// kernels_pcg_coop.cuh -- the whole Jacobi-PCG solve (pressure.cpp:131-241) as ONE cooperative, persistent kernel
// for grids whose working set fits into the SMs' shared memory (the reference's default 512 x 256 up to ~1024^2).
//
// With one launch pair per iteration such grids are launch-bound (512 x 256: 400 launches of ~3 us work each per
// step).  Here every block owns a band of rows and keeps its part of p, r and s in shared memory for the whole solve;
// an iteration needs two grid-wide barriers (cooperative groups):
//     A  s = z + beta*s (own rows and, from the neighbours' r rows, the two halo rows), Ap = L s, partial dot(s,Ap)
//        -> barrier -> every block adds the partials in the same fixed order: alpha, breakdown test (identical everywhere)
//     B  p += alpha*s, r -= alpha*Ap, pin; the block's first/last r row go to global memory (the neighbours' halo),
//        partial dot(r,r), dot(r,z) -> barrier -> res, convergence tests, beta (identical everywhere)
// Ap is recomputed in B from the same s (same bits).  The arithmetic is that of k_pcg_phase1/2; only the summation
// order of the inner products differs (like every other change of launch geometry): results agree to rounding.
#pragma once
#include <cooperative_groups.h>

#include "kernels_pressure.cuh"

namespace cg = cooperative_groups;

#define PC_THREADS 512

struct PcgCoopArgs {
    Geom g;
    const double *rhs;
    double *p;       // global, written once at the end (zeroed by the caller's memset before)
    double *xrows;   // [2 buffers][nblocks][2 rows][nx+2]: first / last r row of every block
    double *slots;   // [3 buffers][2 values][nblocks]: partial sums
    Scalars *sc;
    double tol;
    int max_iters;
    int rows_per_block;
    int retry_only; // 1: run only if k_pcg_cluster (launched just before) gave up on a non-finite value (Scalars::pcg_redo)
};

// Sum of the blocks' partials, identical in every block: warp 0 loads them strided (independent loads), adds its
// strided share in index order, reduces with a fixed shuffle tree and broadcasts through shared memory.
__device__ __forceinline__ void pc_sum(const double *slots, int nblk, int nv, double *out, double *bcast) {
    if (threadIdx.x < 32) {
        for (int k = 0; k < nv; ++k) {
            double a = 0.0;
            for (int b = threadIdx.x; b < nblk; b += 32) a += __ldcg(&slots[k * nblk + b]);
            for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            if (threadIdx.x == 0) bcast[k] = a;
        }
    }
    __syncthreads();
    for (int k = 0; k < nv; ++k) out[k] = bcast[k];
    __syncthreads();
}

__global__ void __launch_bounds__(PC_THREADS) k_pcg_coop(PcgCoopArgs A) {
    if (A.retry_only && !A.sc->pcg_redo) return; // uniform over the grid: nobody reaches a barrier
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm[];
    const Geom &g = A.g;
    const int nblk = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int R = A.rows_per_block, S = g.nx + 2;
    const int j0 = 1 + b * R, j1 = min(j0 + R - 1, g.ny); // raw rows owned (empty if j0 > ny)
    const int nrows = max(j1 - j0 + 1, 0), ncell = nrows * g.nx;
    double *s = sm;                   // (R + 2) x S, row 0 = halo below, ghost columns 0
    double *r = s + (R + 2) * S;      // R x S
    double *p = r + R * S;            // R x S
    __shared__ double red[2 * 32];
    __shared__ double bc[4];          // broadcast of the block's sums

    auto block_sum2 = [&](double a, double c, double &oa, double &oc) {
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
        if ((tid & 31) == 0) red[tid >> 5] = a, red[32 + (tid >> 5)] = c;
        __syncthreads();
        if (tid < 32) {
            a = tid < PC_THREADS / 32 ? red[tid] : 0.0;
            c = tid < PC_THREADS / 32 ? red[32 + tid] : 0.0;
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, o);
                c += __shfl_down_sync(0xffffffffu, c, o);
            }
            if (tid == 0) bc[0] = a, bc[1] = c;
        }
        __syncthreads();
        oa = bc[0], oc = bc[1];
    };
    auto xrow = [&](int buf, int blk, int which) { return A.xrows + (((size_t)buf * nblk + blk) * 2 + which) * S; };
    auto slot = [&](int buf) { return A.slots + (size_t)buf * 2 * nblk; };

    for (int t = tid; t < (R + 2) * S + 2 * R * S; t += PC_THREADS) sm[t] = 0.0;
    __syncthreads();
    // ---- start (pressure.cpp:133-177): p = 0, r = fin0(rhs), rho = dot(r, z), res = sqrt(max(0, dot(r, r)))
    double a0 = 0.0, a1 = 0.0;
    for (int c = tid; c < ncell; c += PC_THREADS) {
        const int lj = c / g.nx, ii = 1 + c - lj * g.nx;
        const double r0 = fin0(A.rhs[gidx(g, ii, j0 + lj)]) - 0.0;
        r[lj * S + ii] = r0;
        a0 += fin0(r0) * fin0(g.inv_diag * r0);
        a1 += fin0(r0) * fin0(r0);
    }
    __syncthreads();
    for (int t = tid; t < 2 * S; t += PC_THREADS) { // my first / last r row for the neighbours
        const int which = t / S, ii = t - which * S;
        xrow(0, b, which)[ii] = nrows > 0 ? r[(which ? nrows - 1 : 0) * S + ii] : 0.0;
    }
    double sa, sb;
    block_sum2(a0, a1, sa, sb);
    if (tid == 0) slot(0)[b] = sa, slot(0)[nblk + b] = sb;
    grid.sync();
    double tot[2];
    pc_sum(slot(0), nblk, 2, tot, bc + 2);
    double rho = tot[0], res = sqrt(max2(0.0, tot[1])), beta = 0.0;
    int iters = 0, breakdown = -1;
    bool first = true, done = !(A.max_iters > 0 && res > A.tol);
    int xb = 0, sbuf = 1; // exchange-row buffer holding the current r rows; next slot buffer

    while (!done) {
        // ---- A: direction update (own rows + halo rows from the neighbours' r rows), Ap, dot(s, Ap)
        for (int c = tid; c < ncell; c += PC_THREADS) {
            const int lj = c / g.nx, ii = 1 + c - lj * g.nx;
            const double z = g.inv_diag * r[lj * S + ii];
            s[(lj + 1) * S + ii] = first ? z : z + beta * s[(lj + 1) * S + ii];
        }
        for (int t = tid; t < 2 * g.nx; t += PC_THREADS) { // halo rows: row below (which = 0) and above (which = 1)
            const int which = t / g.nx, ii = 1 + t - which * g.nx;
            const int nb = which ? b + 1 : b - 1, row = which ? nrows + 1 : 0;
            double v = 0.0;
            const bool live = nrows > 0 && nb >= 0 && nb < nblk && (which ? j1 < g.ny : j0 > 1);
            if (live) {
                const double z = g.inv_diag * __ldcg(&xrow(xb, nb, which ? 0 : 1)[ii]);
                v = first ? z : z + beta * s[row * S + ii];
            }
            s[row * S + ii] = v;
        }
        __syncthreads();
        double d = 0.0;
        for (int c = tid; c < ncell; c += PC_THREADS) {
            const int lj = c / g.nx, ii = 1 + c - lj * g.nx;
            const double *q = s + (lj + 1) * S + ii;
            d += fin0(q[0]) * fin0(lap5(q, S, g.idx2, g.idy2));
        }
        double dsum, dummy;
        block_sum2(d, 0.0, dsum, dummy);
        if (tid == 0) slot(sbuf)[b] = dsum;
        grid.sync();
        pc_sum(slot(sbuf), nblk, 1, tot, bc + 2);
        sbuf = (sbuf + 1) % 3;
        const double denom = tot[0];
        if (!finite_d(denom) || fabs(denom) < 1e-20) { // "PCG breakdown", pressure.cpp:183-186
            breakdown = iters;
            break;
        }
        const double alpha = rho / denom;
        // ---- B: p += alpha s, r -= alpha Ap, pin, dot(r,r), dot(r,z); publish my boundary r rows
        double rr = 0.0, rz = 0.0;
        for (int c = tid; c < ncell; c += PC_THREADS) {
            const int lj = c / g.nx, ii = 1 + c - lj * g.nx;
            const double *q = s + (lj + 1) * S + ii;
            const double Ap = lap5(q, S, g.idx2, g.idy2);
            double pv = p[lj * S + ii] + alpha * q[0];
            const double rv = r[lj * S + ii] - alpha * Ap;
            if (g.j0 == 0 && j0 + lj == 1 && ii == 1) pv = 0.0; // re-pin p(0,0), pressure.cpp:200
            p[lj * S + ii] = pv;
            r[lj * S + ii] = rv;
            rr += fin0(rv) * fin0(rv);
            rz += fin0(rv) * fin0(g.inv_diag * rv);
        }
        __syncthreads();
        xb ^= 1;
        for (int t = tid; t < 2 * S; t += PC_THREADS) {
            const int which = t / S, ii = t - which * S;
            xrow(xb, b, which)[ii] = nrows > 0 ? r[(which ? nrows - 1 : 0) * S + ii] : 0.0;
        }
        double s0, s1;
        block_sum2(rr, rz, s0, s1);
        if (tid == 0) slot(sbuf)[b] = s0, slot(sbuf)[nblk + b] = s1;
        grid.sync();
        pc_sum(slot(sbuf), nblk, 2, tot, bc + 2);
        sbuf = (sbuf + 1) % 3;
        res = sqrt(max2(0.0, tot[0]));
        first = false;
        ++iters;
        if (!finite_d(res) || res < A.tol) break; // pressure.cpp:203
        if (!finite_d(tot[1])) break;             // :217
        beta = tot[1] / rho;
        rho = tot[1];
        if (!(iters < A.max_iters && res > A.tol)) done = true; // loop condition, :179
    }
    // ---- result: p of my rows; scalars by block 0
    int pflag = 0;
    for (int c = tid; c < ncell; c += PC_THREADS) {
        const int lj = c / g.nx, ii = 1 + c - lj * g.nx;
        const double pv = p[lj * S + ii];
        if (!finite_d(pv)) pflag = 1;
        A.p[gidx(g, ii, j0 + lj)] = pv;
    }
    if (pflag) A.sc->nonfinite_p = 1;
    if (b == 0 && tid == 0) {
        Scalars *sc = A.sc;
        sc->res = res;
        sc->iters = iters;
        sc->breakdown_iter = breakdown;
        sc->rho = rho;
        sc->beta = beta;
        sc->first = first ? 1 : 0;
        sc->done = 1;
        sc->pcg_redo = 0;
    }
}
