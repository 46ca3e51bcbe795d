// kernels_pcg_cluster.cuh -- the whole Jacobi-PCG solve (pressure.cpp:131-241) inside ONE thread-block cluster of 16 CTAs,
// for grids of at most 512 x 256 cells (the reference's default grid, main.cpp:105): p, r and s of the whole grid live in the
// shared memory of 16 SMs for the entire solve.
//
// Why: on such a grid the cooperative kernel (kernels_pcg_coop.cuh, 148 blocks) spends an iteration -- 9.2 us for 0.9 cells
// per thread -- in its two grid-wide software barriers and in the L2 round trips that carry the partial sums and the
// neighbours' rows.  Here
//   * the two barriers of an iteration are the hardware cluster barrier (barrier.cluster: SASS UCGABAR_ARV / UCGABAR_WAIT);
//   * partial sums and the neighbours' boundary rows of r travel through distributed shared memory: every CTA stores its
//     partial into slot [rank] of all 16 CTAs and its first / last r row into the neighbours' halo buffer, the barrier
//     publishes them, and every CTA adds the 16 partials in the same fixed order (identical decisions everywhere);
//   * a thread owns one column of its CTA's band of <= 16 rows: its s values and Ap stay in REGISTERS for the iteration
//     (the y-neighbours are its own registers, the x-neighbours and the halo cells come from shared memory), so Ap is
//     computed once per iteration and there is no index arithmetic in the loops.
// Measured (512 x 256, 200 iterations): 6.0 us per iteration against 9.2 us of the cooperative kernel (step 1.92 -> 1.27 ms).
// The sweeps are bound by the latency of dependent fp64 operations, so how many rows of a column the compiler may interleave
// decides: a branch per row (chains one after the other) 8.3 us -- 12 300 of 17 300 SM cycles in the sweeps --, groups of
// 2 / 4 / 8 rows behind one test 7.4 / 6.8 / 6.7 us, the whole column behind one test 6.0 us (PK_SWEEP; threads right of the
// grid skip the sweep: masking their stores instead made the compiler spill, 14.0 us).  1024 threads x 8 rows paid more at
// the barriers than they gained (10.9 us).  What is left: ~2.5 us per iteration in the two reductions + cluster barriers.
// The isfinite guards of the reference (pressure.cpp:20-44 on every operand of the Laplacian, :49-60 on every operand of a
// dot product) are identities while every value is finite.  This kernel computes without them and folds the high words of
// Ap, z = inv_diag * r and p into running maxima; the flag rides on every reduction, so all CTAs learn of a non-finite
// value at the same barrier and give up together WITHOUT having written anything: Scalars::pcg_redo is raised and the
// guarded cooperative kernel, always enqueued behind this one, redoes the solve (it returns at once otherwise).
// Arithmetic per cell as k_pcg_phase1/2; only the summation order of the inner products differs: results agree to rounding
// (tests/test_gpu_parity.py: rel-L2 <= 1e-10, iterations +-1).
#pragma once
#include <cooperative_groups.h>

#include "kernels_pcg_coop.cuh"

#define PK_THREADS 512 // one column per thread: nx <= 512
#define PK_COLS 512
#define PK_CTAS 16     // non-portable cluster size (opt-in by attribute; probed before use)
#define PK_ROWS 16     // rows per CTA: ny <= 256

#ifndef PK_GROUP
#define PK_GROUP 16 // rows of a column whose dependency chains may overlap (one basic block per group)
#endif
// One sweep over this thread's column: the body (the macro's last argument) for every row lj (a compile-time constant inside the body) that the CTA owns.  Rows
// are taken in groups of PK_GROUP behind ONE test, so that the compiler can interleave their chains; a branch per row ran
// them one after the other, all 16 rows in one block spilt.
#define PK_SWEEP(lj, ...)                                                                       \
    _Pragma("unroll") for (int g0_ = 0; g0_ < PK_ROWS; g0_ += PK_GROUP) {                       \
        if (g0_ + PK_GROUP <= nrows && col) {                                                   \
            _Pragma("unroll") for (int k_ = 0; k_ < PK_GROUP; ++k_) { const int lj = g0_ + k_; __VA_ARGS__ } \
        } else {                                                                                \
            _Pragma("unroll") for (int k_ = 0; k_ < PK_GROUP; ++k_) { const int lj = g0_ + k_; if (lj < nrows && col) __VA_ARGS__ } \
        }                                                                                       \
    }

struct PcgClusterArgs {
    Geom g;
    const double *rhs;
    double *p; // global, written once at the end (zeroed by the caller's memset before)
    Scalars *sc;
    double tol;
    int max_iters;
    int rows_per_block;
};

__global__ void __launch_bounds__(PK_THREADS, 1) k_pcg_cluster(PcgClusterArgs A) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double sm[];
    const Geom &g = A.g;
    const int b = (int)cluster.block_rank(), tid = threadIdx.x, lane = tid & 31;
    const int R = A.rows_per_block, S = g.nx + 2;
    const int j0 = 1 + b * R, j1 = min(j0 + R - 1, g.ny); // raw rows owned (none if j0 > ny)
    const int nrows = max(j1 - j0 + 1, 0);
    double *s = sm;                 // (R + 2) x S: rows 1..R own (row 0 and R+1 unused here), ghost columns stay 0
    double *r = s + (R + 2) * S;    // R x S
    double *p = r + R * S;          // R x S
    double *rhalo = p + R * S;      // 2 x S: the r row just below ([0]) / above ([1]) my band, stored by the neighbours
    __shared__ double slotA[2 * PK_CTAS], slotB[3 * PK_CTAS]; // [value][rank], stored by every CTA of the cluster
    __shared__ double red[3 * (PK_THREADS / 32)];
    const int ii = tid + 1;
    const bool col = ii <= g.nx;
    const bool live_lo = nrows > 0 && b > 0 && j0 > 1;              // a band with rows lies below mine
    const bool live_hi = nrows > 0 && b + 1 < PK_CTAS && j1 < g.ny; // ... above mine
    const double idx2 = g.idx2, idy2 = g.idy2, inv_diag = g.inv_diag;

    int mxs = 0;       // running maxima of high words (kernels_rhs_march2.cuh): a value is non-finite iff its high word is
    unsigned mxu = 0u; // >= 0x7ff00000 as a signed or >= 0xfff00000 as an unsigned number
    auto track = [&](double x) {
        const int hw = __double2hiint(x);
        mxs = max(mxs, hw);
        mxu = max(mxu, (unsigned)hw);
    };
    auto bad = [&]() { return (mxs >= 0x7ff00000 || mxu >= 0xfff00000u) ? 1.0 : 0.0; };
    // Sum of NV per-thread values over the cluster: shuffle tree -> one value per warp -> warp 0 -> slot [b] of EVERY CTA.
    // The caller's cluster.sync() publishes the slots; cluster_total() then adds them in rank order (butterfly over 16 lanes:
    // the same additions in every warp of every CTA).
    auto cluster_partial = [&](double *slots, double v0, double v1, double v2, int nv) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
            v2 += __shfl_down_sync(0xffffffffu, v2, o);
        }
        constexpr int NW = PK_THREADS / 32;
        if (lane == 0) red[tid >> 5] = v0, red[NW + (tid >> 5)] = v1, red[2 * NW + (tid >> 5)] = v2;
        __syncthreads();
        if (tid < 32) {
            v0 = tid < NW ? red[tid] : 0.0, v1 = tid < NW ? red[NW + tid] : 0.0, v2 = tid < NW ? red[2 * NW + tid] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_down_sync(0xffffffffu, v0, o);
                v1 += __shfl_down_sync(0xffffffffu, v1, o);
                v2 += __shfl_down_sync(0xffffffffu, v2, o);
            }
            v0 = __shfl_sync(0xffffffffu, v0, 0), v1 = __shfl_sync(0xffffffffu, v1, 0), v2 = __shfl_sync(0xffffffffu, v2, 0);
            if (tid < PK_CTAS) { // lane `tid` stores my partials into CTA `tid`
                double *dst = cluster.map_shared_rank(slots, tid);
                dst[b] = v0;
                if (nv > 1) dst[PK_CTAS + b] = v1;
                if (nv > 2) dst[2 * PK_CTAS + b] = v2;
            }
        }
    };
    auto cluster_total = [&](const double *slots, int k) {
        double a = slots[k * PK_CTAS + (lane & (PK_CTAS - 1))];
#pragma unroll
        for (int o = PK_CTAS / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        return a;
    };
    auto send_rows = [&](double first_row, double last_row) { // my boundary r rows into the neighbours' halo buffers
        if (col && live_lo) cluster.map_shared_rank(rhalo, b - 1)[S + ii] = first_row;
        if (col && live_hi) cluster.map_shared_rank(rhalo, b + 1)[ii] = last_row;
    };

    for (int t = tid; t < (3 * R + 4) * S; t += PK_THREADS) sm[t] = 0.0;
    __syncthreads();
    cluster.sync(); // nobody stores into a peer's shared memory before that peer has zeroed it
    // ---- start (pressure.cpp:133-177): p = 0, r = rhs - A p, rho = dot(r, z), res = sqrt(max(0, dot(r, r)))
    double sreg[PK_ROWS], apreg[PK_ROWS]; // this thread's column of s and of A s
    double shb = 0.0, sha = 0.0;          // s just below / above the band in this column
    double a0 = 0.0, a1 = 0.0, rfirst = 0.0, rlast = 0.0;
#pragma unroll
    for (int lj = 0; lj < PK_ROWS; ++lj) {
        sreg[lj] = 0.0, apreg[lj] = 0.0;
        if (lj < nrows && col) {
            const double r0 = A.rhs[gidx(g, ii, j0 + lj)] - 0.0;
            const double z = inv_diag * r0;
            r[lj * S + ii] = r0;
            a0 += r0 * z;
            a1 += r0 * r0;
            track(z);
            if (lj == 0) rfirst = r0;
            if (lj == nrows - 1) rlast = r0;
        }
    }
    send_rows(rfirst, rlast);
    cluster_partial(slotB, a0, a1, bad(), 3);
    cluster.sync();
    bool redo = cluster_total(slotB, 2) > 0.0;
    double rho = cluster_total(slotB, 0), res = sqrt(max2(0.0, cluster_total(slotB, 1))), beta = 0.0;
    int iters = 0, breakdown = -1;
    bool first = true, done = !(A.max_iters > 0 && res > A.tol);

    while (!done && !redo) {
        // ---- A: direction update of my column (+ the two halo cells from the neighbours' r rows), Ap, dot(s, Ap)
        PK_SWEEP(lj, {
            const double z = inv_diag * r[lj * S + ii];
            sreg[lj] = first ? z : z + beta * sreg[lj];
            s[(lj + 1) * S + ii] = sreg[lj];
        })
        if (col) { // the neighbours' direction in the halo cells, recomputed from their r rows (the same bits as they hold)
            double v = 0.0;
            if (live_lo) {
                const double z = inv_diag * rhalo[ii];
                v = first ? z : z + beta * shb;
            }
            shb = v;
            v = 0.0;
            if (live_hi) {
                const double z = inv_diag * rhalo[S + ii];
                v = first ? z : z + beta * sha;
            }
            sha = v;
        }
        __syncthreads();
        double d = 0.0;
        PK_SWEEP(lj, {
            const double c = sreg[lj], l = s[(lj + 1) * S + ii - 1], rt = s[(lj + 1) * S + ii + 1];
            const double bl = lj > 0 ? sreg[lj > 0 ? lj - 1 : 0] : shb;
            const double ab = lj + 1 < nrows ? sreg[lj + 1 < PK_ROWS ? lj + 1 : lj] : sha;
            const double Ap = (rt - 2.0 * c + l) * idx2 + (ab - 2.0 * c + bl) * idy2;
            apreg[lj] = Ap;
            d += c * Ap;
            track(Ap);
        })
        cluster_partial(slotA, d, bad(), 0.0, 2);
        cluster.sync();
        if (cluster_total(slotA, 1) > 0.0) { redo = true; break; }
        const double denom = cluster_total(slotA, 0);
        if (!finite_d(denom) || fabs(denom) < 1e-20) { // "PCG breakdown", pressure.cpp:183-186
            breakdown = iters;
            break;
        }
        const double alpha = rho / denom;
        // ---- B: p += alpha s, r -= alpha Ap, pin, dot(r,r), dot(r,z); my boundary r rows go to the neighbours
        double rr = 0.0, rz = 0.0;
        PK_SWEEP(lj, {
            double pv = p[lj * S + ii] + alpha * sreg[lj];
            const double rv = r[lj * S + ii] - alpha * apreg[lj];
            if (g.j0 == 0 && j0 + lj == 1 && ii == 1) pv = 0.0; // re-pin p(0,0), pressure.cpp:200
            p[lj * S + ii] = pv;
            r[lj * S + ii] = rv;
            const double z = inv_diag * rv;
            rr += rv * rv;
            rz += rv * z;
            track(z);
            track(pv);
            if (lj == 0) rfirst = rv;
            if (lj == nrows - 1) rlast = rv;
        })
        send_rows(rfirst, rlast);
        cluster_partial(slotB, rr, rz, bad(), 3);
        cluster.sync();
        if (cluster_total(slotB, 2) > 0.0) { redo = true; break; }
        const double t0 = cluster_total(slotB, 0), t1 = cluster_total(slotB, 1);
        res = sqrt(max2(0.0, t0));
        first = false;
        ++iters;
        if (!finite_d(res) || res < A.tol) break; // pressure.cpp:203
        if (!finite_d(t1)) break;                 // :217
        beta = t1 / rho;
        rho = t1;
        if (!(iters < A.max_iters && res > A.tol)) done = true; // loop condition, :179
    }
    cluster.sync(); // no CTA leaves while a peer may still store into its shared memory
    if (redo) {     // identical in every CTA: nothing has been written; the guarded kernel behind this one redoes the solve
        if (b == 0 && tid == 0) A.sc->pcg_redo = 1;
        return;
    }
    // ---- result: p of my rows; scalars by CTA 0
#pragma unroll
    for (int lj = 0; lj < PK_ROWS; ++lj)
        if (lj < nrows && col) A.p[gidx(g, ii, j0 + lj)] = p[lj * S + ii];
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
