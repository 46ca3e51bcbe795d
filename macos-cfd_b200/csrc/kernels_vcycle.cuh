// kernels_vcycle.cuh -- EXTENSION: a true geometric V-cycle (mg_mode = CFD_B200_MG_VCYCLE) honouring mg_levels and
// rbgs_sweeps.  The reference has no multigrid (its "Multigrid" is fine-grid RBGS, pressure.cpp:113-129), so this
// mode has NO reference oracle: parity is UNPINNED.  The algorithm is defined by the CPU restatement cfd_vcycle.c (test tree) and these
// kernels reproduce it bit for bit (same expressions, IEEE division, no FMA); see that file for the method:
// cell-centred coarsening by 2, red-black Gauss-Seidel with the correct sign, p = 0 on the wall faces (modified
// diagonal for wall cells: the same physical boundary on every level), residual + 4-cell-mean restriction fused,
// bilinear prolongation + correction fused, 16 x sweeps on the coarsest level.
//
// Levels that fit into one SM's shared memory are handled by ONE block (k_mg_coarse_cta): the whole lower part of
// the V (smooth / restrict ... coarsest ... prolong / smooth) runs out of shared memory with block barriers only,
// instead of ~10 tiny launches per level.
#pragma once
#include "common.cuh"

struct MgLevel {
    int nx, ny, P;   // ny: LOCAL cell rows of the level (the whole level on a single-GPU handle)
    double idx2, idy2, denom;
    double *p, *rhs; // device arrays in the common padded layout (ghost ring = 0)
    double *alt;     // ping-pong partner of p (streaming sweeps); level 0: the State's p_alt
    // row slabs: global index of local cell row 1 minus 1 (parity of the colours), and whether the slab touches the
    // domain's bottom / top wall (a slab-internal boundary is a halo row, not a wall); 0 / 1 / 1 on a single GPU
    int j0, bottom, top;
};
__host__ __device__ __forceinline__ size_t mg_idx(const MgLevel &L, int ii, int jj) {
    return (size_t)(jj + CFD_YOFF) * (size_t)L.P + (size_t)(ii + CFD_XOFF);
}
__device__ __forceinline__ double mg_wall(int nx, int ny, double idx2, double idy2, int ii, int jj) {
    int bx = (ii == 1) + (ii == nx), by = (jj == 1) + (jj == ny);
    return (double)bx * idx2 + (double)by * idy2;
}
__device__ __forceinline__ double mg_wall(const MgLevel &L, int ii, int jj) { // the same term on one slab of the level
    int bx = (ii == 1) + (ii == L.nx), by = (L.bottom && jj == 1) + (L.top && jj == L.ny);
    return (double)bx * L.idx2 + (double)by * L.idy2;
}

// ---------------------------------------------------------------- generic (any level) global-memory kernels
__global__ void k_mg_zero_ghosts(MgLevel L, const Scalars *sc) { pdl_begin();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= L.nx + 1) {
        if (L.bottom) L.p[mg_idx(L, t, 0)] = 0.0;
        if (L.top) L.p[mg_idx(L, t, L.ny + 1)] = 0.0;
    }
    if (t <= L.ny + 1) {
        L.p[mg_idx(L, 0, t)] = 0.0;
        L.p[mg_idx(L, L.nx + 1, t)] = 0.0;
    }
}

__global__ void __launch_bounds__(256) k_mg_colour(MgLevel L, int colour, const Scalars *sc) { pdl_begin();
    if (sc->done) return;
    const int jj = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (jj > L.ny) return;
    const int ii = 1 + 2 * (blockIdx.x * blockDim.x + threadIdx.x) + ((L.j0 + jj - 1 + colour) & 1);
    if (ii > L.nx) return;
    const size_t o = mg_idx(L, ii, jj);
    const double *q = L.p + o;
    double sum = (q[1] + q[-1]) * L.idx2 + (q[L.P] + q[-L.P]) * L.idy2;
    L.p[o] = (sum - L.rhs[o]) / (L.denom + mg_wall(L, ii, jj));
}

__device__ __forceinline__ double mg_resid(const MgLevel &L, int ii, int jj) {
    const size_t o = mg_idx(L, ii, jj);
    const double *q = L.p + o;
    double Ap = (q[1] - 2.0 * q[0] + q[-1]) * L.idx2 + (q[L.P] - 2.0 * q[0] + q[-L.P]) * L.idy2;
    Ap = Ap - mg_wall(L, ii, jj) * q[0];
    return L.rhs[o] - Ap;
}

// residual of the fine level + restriction to the coarse right-hand side + coarse initial guess 0, fused
__global__ void __launch_bounds__(256) k_mg_restrict(MgLevel F, MgLevel C, const Scalars *sc) { pdl_begin();
    if (sc->done) return;
    const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (I > C.nx || J > C.ny) return;
    const int ii = 2 * I - 1, jj = 2 * J - 1;
    double a = mg_resid(F, ii, jj), b = mg_resid(F, ii + 1, jj), c = mg_resid(F, ii, jj + 1), d = mg_resid(F, ii + 1, jj + 1);
    const size_t o = mg_idx(C, I, J);
    C.rhs[o] = 0.25 * ((a + b) + (c + d));
    C.p[o] = 0.0;
}

// bilinear prolongation + correction, fused
__global__ void __launch_bounds__(256) k_mg_prolong(MgLevel F, MgLevel C, const Scalars *sc) { pdl_begin();
    if (sc->done) return;
    const int ii = 1 + blockIdx.x * blockDim.x + threadIdx.x, jj = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (ii > F.nx || jj > F.ny) return;
    const int I = (ii + 1) / 2, J = (jj + 1) / 2;
    const int di = (ii & 1) ? -1 : 1, dj = (jj & 1) ? -1 : 1;
    const double *c = C.p + mg_idx(C, I, J);
    double e = 0.0625 * ((9.0 * c[0] + 3.0 * (c[di] + c[dj * C.P])) + c[dj * C.P + di]);
    F.p[mg_idx(F, ii, jj)] += e;
}

// ||rhs - L p||_2 of the fine level after a cycle, and the cycle loop's exit test
// (finish = 0 on row slabs: the local sum goes to Scalars::red[0] and the reduction over the slabs runs the same tail)
__device__ __forceinline__ void mg_norm_tail(Scalars *sc, double sum, double tol) {
    double res = sqrt(sum);
    sc->res = res;
    sc->iters = sc->iters + 1;
    if (res < tol) sc->done = 1;
}
__global__ void __launch_bounds__(256) k_mg_residual_norm(MgLevel L, Scalars *sc, double *partials, double tol, int finish) { pdl_begin();
    if (sc->done) return;
    const int ii = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    if (ii <= L.nx)
        for (int jj = 1 + blockIdx.y; jj <= L.ny; jj += gridDim.y) {
            double r = mg_resid(L, ii, jj);
            acc[0] += r * r;
        }
    grid_sum<1>(acc, partials, &sc->ticket[2], [=](double(&t)[1]) {
        if (finish) mg_norm_tail(sc, t[0], tol);
        else sc->red[0] = t[0];
    });
}
__global__ void k_mg_norm_fin(Scalars *sc, double tol) { pdl_begin(); if (!sc->done) mg_norm_tail(sc, sc->red[0], tol); }

// ---------------------------------------------------------------- CTA-resident coarse part of the V
#define MG_CTA_MAXLEV 12
struct MgCtaPlan {
    int nlev;                   // levels handled inside the block (first one = level lc of the hierarchy)
    int nx[MG_CTA_MAXLEV], ny[MG_CTA_MAXLEV];
    int off[MG_CTA_MAXLEV];     // offset (doubles) of the level's p array in shared memory; rhs follows p
    double idx2[MG_CTA_MAXLEV], idy2[MG_CTA_MAXLEV], denom[MG_CTA_MAXLEV];
    int sweeps;
};
// shared-memory arrays have pitch nx+2 and rows ny+2 (ghost ring = 0), exactly the restatement's layout
__device__ __forceinline__ void cta_smooth(double *p, const double *rhs, int nx, int ny, double idx2, double idy2,
                                           double denom, int sweeps) {
    const int S = nx + 2, half = (nx + 1) / 2;
    for (int s = 0; s < sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour) {
            for (int t = threadIdx.x; t < half * ny; t += blockDim.x) {
                const int jj = 1 + t / half, ii = 1 + 2 * (t % half) + ((jj - 1 + colour) & 1);
                if (ii <= nx) {
                    const double *q = p + jj * S + ii;
                    double sum = (q[1] + q[-1]) * idx2 + (q[S] + q[-S]) * idy2;
                    p[jj * S + ii] = (sum - rhs[jj * S + ii]) / (denom + mg_wall(nx, ny, idx2, idy2, ii, jj));
                }
            }
            __syncthreads();
        }
}
__device__ __forceinline__ double cta_resid(const double *p, const double *rhs, int nx, int ny, double idx2,
                                            double idy2, int ii, int jj) {
    const int S = nx + 2;
    const double *q = p + jj * S + ii;
    double Ap = (q[1] - 2.0 * q[0] + q[-1]) * idx2 + (q[S] - 2.0 * q[0] + q[-S]) * idy2;
    Ap = Ap - mg_wall(nx, ny, idx2, idy2, ii, jj) * q[0];
    return rhs[jj * S + ii] - Ap;
}

__global__ void __launch_bounds__(1024) k_mg_coarse_cta(MgLevel G, MgCtaPlan plan, const Scalars *sc) { pdl_begin();
    if (sc->done) return;
    extern __shared__ double smg[];
    auto P = [&](int l) { return smg + plan.off[l]; };
    auto R = [&](int l) { return smg + plan.off[l] + (plan.nx[l] + 2) * (plan.ny[l] + 2); };
    const int total = plan.off[plan.nlev - 1] + 2 * (plan.nx[plan.nlev - 1] + 2) * (plan.ny[plan.nlev - 1] + 2);
    for (int t = threadIdx.x; t < total; t += blockDim.x) smg[t] = 0.0;
    __syncthreads();
    { // level 0 of the plan: rhs from global memory (written by the restriction above it), initial guess 0
        const int nx = plan.nx[0], ny = plan.ny[0], S = nx + 2;
        for (int t = threadIdx.x; t < nx * ny; t += blockDim.x) {
            const int jj = 1 + t / nx, ii = 1 + t % nx;
            R(0)[jj * S + ii] = G.rhs[mg_idx(G, ii, jj)];
        }
    }
    __syncthreads();
    const int last = plan.nlev - 1;
    for (int l = 0; l < last; ++l) { // downward leg
        cta_smooth(P(l), R(l), plan.nx[l], plan.ny[l], plan.idx2[l], plan.idy2[l], plan.denom[l], plan.sweeps);
        const int cx = plan.nx[l + 1], cy = plan.ny[l + 1], CS = cx + 2;
        for (int t = threadIdx.x; t < cx * cy; t += blockDim.x) {
            const int J = 1 + t / cx, I = 1 + t % cx, ii = 2 * I - 1, jj = 2 * J - 1;
            double a = cta_resid(P(l), R(l), plan.nx[l], plan.ny[l], plan.idx2[l], plan.idy2[l], ii, jj);
            double b = cta_resid(P(l), R(l), plan.nx[l], plan.ny[l], plan.idx2[l], plan.idy2[l], ii + 1, jj);
            double c = cta_resid(P(l), R(l), plan.nx[l], plan.ny[l], plan.idx2[l], plan.idy2[l], ii, jj + 1);
            double d = cta_resid(P(l), R(l), plan.nx[l], plan.ny[l], plan.idx2[l], plan.idy2[l], ii + 1, jj + 1);
            R(l + 1)[J * CS + I] = 0.25 * ((a + b) + (c + d));
            P(l + 1)[J * CS + I] = 0.0;
        }
        __syncthreads();
    }
    cta_smooth(P(last), R(last), plan.nx[last], plan.ny[last], plan.idx2[last], plan.idy2[last], plan.denom[last],
               16 * plan.sweeps);
    for (int l = last - 1; l >= 0; --l) { // upward leg
        const int nx = plan.nx[l], ny = plan.ny[l], S = nx + 2, CS = plan.nx[l + 1] + 2;
        for (int t = threadIdx.x; t < nx * ny; t += blockDim.x) {
            const int jj = 1 + t / nx, ii = 1 + t % nx;
            const int I = (ii + 1) / 2, J = (jj + 1) / 2, di = (ii & 1) ? -1 : 1, dj = (jj & 1) ? -1 : 1;
            const double *c = P(l + 1) + J * CS + I;
            double e = 0.0625 * ((9.0 * c[0] + 3.0 * (c[di] + c[dj * CS])) + c[dj * CS + di]);
            P(l)[jj * S + ii] += e;
        }
        __syncthreads();
        cta_smooth(P(l), R(l), nx, ny, plan.idx2[l], plan.idy2[l], plan.denom[l], plan.sweeps);
    }
    { // result: p of the plan's first level back to global memory
        const int nx = plan.nx[0], ny = plan.ny[0], S = nx + 2;
        for (int t = threadIdx.x; t < nx * ny; t += blockDim.x) {
            const int jj = 1 + t / nx, ii = 1 + t % nx;
            G.p[mg_idx(G, ii, jj)] = P(0)[jj * S + ii];
        }
    }
}
