// kernels_rhs_march.cuh -- the optimised advection + diffusion + Runge-Kutta kernel for INTERIOR faces.
//
// ncu on the simple tile kernel (profiles/r01_v1_*): issue-bound (76 % issue slots, 624 thread-instructions per
// cell, 17 % DRAM): every face flux was computed twice (as "+" flux of one face and "-" flux of its neighbour;
// the two expressions in advection.cpp:59-69 / :72-81 are textually the same function of the same samples), every
// slope twice more, and clip_face_states' 3-way min/max + two clamps cost 13 FP64 compares and 18 32-bit selects
// per flux.  This kernel
//   * marches each thread up a column: the y-stencil (q[j-1..j+2], the slope and the flux of the face below) lives
//     in registers, so each y-flux and y-slope is computed exactly once;
//   * gives each thread 2 adjacent columns (one aligned double2 per row and field) and exchanges the x-neighbours'
//     value / slope / flux with 4 warp shuffles per field: each x-flux and x-slope is computed once per warp
//     (lanes 0 and 31 only feed their neighbours: 60 of 64 columns produce output);
//   * applies clip_face_states only on a slow path (one finiteness test per derivative, not per flux).  Proof that this is exact: with s = minmod(c-m1, p1-c),
//     |s| <= |RN(p1-c)| and s has the sign of p1-c, so c + s/2 lies between c and p1 in exact arithmetic
//     (RN(p1-c)/2 < |p1-c| once the difference is inexact, and the difference is exact when it is small), and
//     rounding is monotone, hence RN(c + s/2) is inside [min(c,p1), max(c,p1)], a subset of the clip interval
//     [min(m1,c,p1), max(m1,c,p1)]; the same holds for p1 - s'/2.  The argument only fails when a difference
//     overflowed or an input is non-finite, and then the face state itself is non-finite; that is what the slow
//     path tests.  Selecting the upwind state before clamping is exact because both states get the same clamp.
// Every other expression keeps the reference's operand order, and the build has no FMA contraction.
//
// Only faces whose four fluxes are all of the MUSCL kind (2 <= ii <= nx-1, 2 <= jj <= ny-1 in raw indices, valid
// for u and v) are handled here; the first-order ring and the rows/columns next to it stay with k_rhs_simple.
#pragma once
#include "kernels_step.cuh"

#define MARCH_WARPS 4
#ifndef MARCH_UNROLL
#define MARCH_UNROLL 3 // B200, 8192^2: unroll 1/2/3/4 -> stage 1 0.555/0.558/0.542/0.614 ms (3 = period of the q rotation)
#endif
#define MARCH_STR2(x) #x
#define MARCH_STR(x) MARCH_STR2(x)
#ifndef MARCH_MINBLOCKS
#define MARCH_MINBLOCKS 4 // resident blocks per SM the register allocator must allow (tuned on B200, see profiles/)
#endif
#define MARCH_COLS 60 // output columns per warp (lanes 1..30, 2 columns each)

// 0.5 * minmod(a, b)   (advection.cpp:10-12, then the 0.5 of :61-62)
__device__ __forceinline__ double half_minmod(double a, double b) {
    double s = (fabs(a) < fabs(b)) ? a : b;
    s = (a * b <= 0.0) ? 0.0 : s;
    return 0.5 * s;
}

// Flux through the face between c and p1 given the half slopes at c and p1.  No clip here: a face state that the
// clip could change is non-finite (header), a non-finite state makes the flux, the flux difference and finally the
// derivative non-finite, and the caller recomputes every non-finite derivative with the reference formula (rhs_point).
__device__ __forceinline__ double flux_shared(double c, double hc, double p1, double hp, double w) {
    double qL = c + hc;
    double qR = p1 - hp;
    return ((w > 0.0) ? qL : qR) * w;
}

// Out-of-line slow path: the reference formula for one interior face, read straight from memory (keeps the hot
// loop's register budget free of the clip logic).
__device__ __noinline__ double rhs_point_slow(const double *q, const double *u, const double *v, int S, double idx,
                                              double idy, double idx2, double idy2, double invRe) {
    RhsConst k{idx, idy, idx2, idy2, invRe};
    return rhs_point(q, u, v, S, 0, 0, k);
}

struct MarchCol { // y-state of one column of one field while marching upwards
    double qm1, q0, q1; // q(jj-1), q(jj), q(jj+1)
    double dy0;         // q(jj+1) - q(jj)
    double hs0;         // half slope at jj
    double fym;         // y-flux through the face below jj
};

__device__ __forceinline__ double shfl_dn(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_up(double x) { return __shfl_up_sync(0xffffffffu, x, 1); }

// MODE as in k_rhs_simple (uacc / vacc: the accumulator read by MODE 2).  Outputs: ii in [XA, XB], jj in [YA, YB]; each block
// row-strip covers RY rows.
//
// MODE 4 = MODE 2 + the divergence of what it produces (project.cpp:7-42 with the scaling of time_integrator.cpp:158-182):
// a thread that has just formed u*, v* of row jj holds its own faces of rows jj-1 and jj and gets the u face right of its
// pair from the lane above, so rhs = div(u*, v*) / dt of cell row jj-1 is written from registers and k_divergence<PRE> never
// re-reads u*, v* (48 + 24 -> 56 bytes per cell).  Lane 31, which otherwise only feeds its neighbours, forms u* of its first
// column for that purpose (one extra load per row: the u1 column right of the warp's 64).  Cells this kernel closes:
// columns [XA, XB-1], rows [YA, YB-1] except the top row of every strip; k_div_fix_pre does the others -- the tile strips
// around the marching region, whose faces come from k_rhs_simple and the boundary pass, and those strip-top rows.
#ifndef MARCH_MINBLOCKS_DIV
#define MARCH_MINBLOCKS_DIV 4
#endif
template <int MODE>
__global__ void __launch_bounds__(32 * MARCH_WARPS, MODE == 4 ? MARCH_MINBLOCKS_DIV : MARCH_MINBLOCKS)
    k_rhs_march(Geom g, RhsConst k, const double *__restrict__ uin, const double *__restrict__ vin, const double *uacc,
                const double *vacc, double *uout, double *vout, Scalars *sc, int XA, int XB, int YA, int YB, int RY, WaitArgs hw,
                double *rhs, double *partials) { pdl_begin();
    constexpr bool HEUN = MODE == 2 || MODE == 4, DIV = MODE == 4;
    double dacc[1] = {0.0};
    // the first strip of rows goes to the LAST blocks of the grid (see k_rbgs_stream: strips that wait for a neighbour's rows
    // should be scheduled when the rows are there)
    const int by = blockIdx.y + 1 == gridDim.y ? 0 : (int)blockIdx.y + 1;
    {   // row slabs: only the strips that read the neighbours' rows wait for them (rows y0-2 .. y1+3 are touched)
        const int yb0 = YA + by * RY, yb1 = min(yb0 + RY - 1, YB);
        p2p_wait_block(hw, yb0 - 2 < 1, yb1 + 3 > g.ny, sc);
    }
    const int lane = threadIdx.x & 31;
    const int wx = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    const int X0 = XA - 2 + MARCH_COLS * wx; // raw column of lane 0's first column (odd => aligned double2)
    if (X0 + 2 > XB) return;                 // whole warp outside (before any shuffle: the rest is convergent code)
    const int a = X0 + 2 * lane;             // this thread's columns: a, a+1
    const int y0 = YA + by * RY;
    const int y1 = min(y0 + RY - 1, YB);
    const bool loadable = (a + 1 <= g.P - CFD_XOFF - 1);
    const size_t P = g.P;
    const double *pu = uin + gidx(g, a, 0);
    const double *pv = vin + gidx(g, a, 0);
    auto ld2 = [&](const double *base, int jj) -> double2 {
        return loadable ? *reinterpret_cast<const double2 *>(base + (size_t)jj * P) : make_double2(0.0, 0.0);
    };
    const double dt = (MODE == 0) ? 0.0 : sc->dt;
    const bool st_a = lane >= 1 && lane <= 30 && a >= XA && a <= XB;
    const bool st_b = lane >= 1 && lane <= 30 && a + 1 >= XA && a + 1 <= XB;
    int flag = 0;
    // MODE 4: lane 31 forms u* of its first column (the face right of lane 30's pair); cells this thread closes
    const bool ext = DIV && lane == 31 && a <= XB;
    const bool cA = DIV && st_a && st_b, cB = DIV && st_b && a + 2 <= XB;
    const double inv_dt = DIV ? sc->inv_dt : 0.0;
    double2 pun = make_double2(0.0, 0.0), pvn = make_double2(0.0, 0.0); // u*, v* of the row below
    double urn = 0.0;                                                     // u*(a+2) of the row below
    double xq = ext ? pu[(size_t)y0 * P + 2] : 0.0;                       // u1(a+2, jj) for lane 31

    MarchCol U[2], V[2];
    // ---- prologue: rows y0-2 .. y0+1 -> state at jj = y0 (fym = flux through the face between y0-1 and y0)
    {
        double2 u0 = ld2(pu, y0 - 2), u1 = ld2(pu, y0 - 1), u2 = ld2(pu, y0), u3 = ld2(pu, y0 + 1);
        double2 v0 = ld2(pv, y0 - 2), v1 = ld2(pv, y0 - 1), v2 = ld2(pv, y0), v3 = ld2(pv, y0 + 1);
        auto init = [&](MarchCol &c, double q_2, double q_1, double q0, double q1, double w_below) {
            double dm2 = q_1 - q_2, dm1 = q0 - q_1, d0 = q1 - q0;
            double hsm1 = half_minmod(dm2, dm1);
            c.hs0 = half_minmod(dm1, d0);
            c.fym = flux_shared(q_1, hsm1, q0, c.hs0, w_below);
            c.qm1 = q_1;
            c.q0 = q0;
            c.q1 = q1;
            c.dy0 = d0;
        };
        // y-faces are carried by v(ii, jj) of the row below (advection.cpp:113, :275)
        init(U[0], u0.x, u1.x, u2.x, u3.x, v1.x);
        init(U[1], u0.y, u1.y, u2.y, u3.y, v1.y);
        init(V[0], v0.x, v1.x, v2.x, v3.x, v1.x);
        init(V[1], v0.y, v1.y, v2.y, v3.y, v1.y);
    }

    // rows are requested one iteration ahead of their use (ncu: the loop was stalled on its own loads)
    double2 pre_u = ld2(pu, y0 + 2), pre_v = ld2(pv, y0 + 2);
_Pragma(MARCH_STR(unroll MARCH_UNROLL))
    for (int jj = y0; jj <= y1; ++jj) {
        const double2 nu = pre_u, nv = pre_v;
        if (jj < y1) {
            pre_u = ld2(pu, jj + 3);
            pre_v = ld2(pv, jj + 3);
        }
        const double xq_now = xq;
        if (ext && jj < y1) xq = pu[(size_t)(jj + 1) * P + 2];
        const size_t o = gidx(g, a, jj);
        double2 au = make_double2(0.0, 0.0), av = make_double2(0.0, 0.0);
        if (HEUN && (st_a || st_b || ext)) { // s.u / s.v (uacc may be uout: only this thread touches these elements)
            au = *reinterpret_cast<const double2 *>(uacc + o);
            av = *reinterpret_cast<const double2 *>(vacc + o);
        }
        const double ua = U[0].q0, ub = U[1].q0, va = V[0].q0, vb = V[1].q0; // face velocities of this row

        double du[2], dv[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) { // f = 0: q = u, f = 1: q = v
            MarchCol *C = f ? V : U;
            const double q2a = f ? nv.x : nu.x, q2b = f ? nv.y : nu.y;
            // ---- y direction: one new slope and one new flux per column (advection.cpp:85-114, :247-276)
            double dy1a = q2a - C[0].q1, dy1b = q2b - C[1].q1;
            double hs1a = half_minmod(C[0].dy0, dy1a), hs1b = half_minmod(C[1].dy0, dy1b);
            double fya = flux_shared(C[0].q0, C[0].hs0, C[0].q1, hs1a, va);
            double fyb = flux_shared(C[1].q0, C[1].hs0, C[1].q1, hs1b, vb);
            // ---- x direction on row jj, shared across the warp (advection.cpp:52-81, :280-309)
            const double qa = C[0].q0, qb = C[1].q0;
            double qa_n = shfl_dn(qa);       // q(a+2)
            if (DIV && f == 0 && lane == 31) qa_n = xq_now;
            const double qb_p = shfl_up(qb); // q(a-1)
            double dxp = qa - qb_p, dxa = qb - qa, dxb = qa_n - qb;
            double hsa = half_minmod(dxp, dxa), hsb = half_minmod(dxa, dxb);
            const double hsa_n = shfl_dn(hsa);
            double fxa = flux_shared(qa, hsa, qb, hsb, ua);     // face a+1/2, carried by u(a, jj)
            double fxb = flux_shared(qb, hsb, qa_n, hsa_n, ub);   // face b+1/2, carried by u(b, jj)
            const double fxb_p = shfl_up(fxb);                        // face a-1/2
            // ---- flux divergence + diffusion (advection.cpp:117-118, diffusion.cpp:16-18)
            double conva = (fxa - fxb_p) * k.idx + (fya - C[0].fym) * k.idy;
            double convb = (fxb - fxa) * k.idx + (fyb - C[1].fym) * k.idy;
            double lapa = (qb - 2.0 * qa + qb_p) * k.idx2 + (C[0].q1 - 2.0 * qa + C[0].qm1) * k.idy2;
            double lapb = (qa_n - 2.0 * qb + qa) * k.idx2 + (C[1].q1 - 2.0 * qb + C[1].qm1) * k.idy2;
            double da = -conva, db = -convb;
            da = da + k.invRe * lapa;
            db = db + k.invRe * lapb;
            // slow path (overflow / non-finite input only): the reference formula, clip included, straight from memory
            const double *qg = (f ? vin : uin) + o;
            const bool need_a = st_a || (ext && f == 0);
            if (need_a && !finite_d(da)) da = rhs_point_slow(qg, uin + o, vin + o, (int)P, k.idx, k.idy, k.idx2, k.idy2, k.invRe);
            if (st_b && !finite_d(db)) db = rhs_point_slow(qg + 1, uin + o + 1, vin + o + 1, (int)P, k.idx, k.idy, k.idx2, k.idy2, k.invRe);
            double *d = f ? dv : du;
            d[0] = need_a ? sanitize_val(da, flag) : 0.0; // lanes that only feed their neighbours hold no result
            d[1] = st_b ? sanitize_val(db, flag) : 0.0;
            // ---- rotate the y-state
            C[0].qm1 = C[0].q0; C[0].q0 = C[0].q1; C[0].q1 = q2a; C[0].dy0 = dy1a; C[0].hs0 = hs1a; C[0].fym = fya;
            C[1].qm1 = C[1].q0; C[1].q0 = C[1].q1; C[1].q1 = q2b; C[1].dy0 = dy1b; C[1].hs0 = hs1b; C[1].fym = fyb;
        }
        // ---- Runge-Kutta update + store (time_integrator.cpp:84-99, :122-143); qm1 now holds q(jj)
        double2 ou, ov;
        if (MODE == 0) {
            ou = make_double2(du[0], du[1]);
            ov = make_double2(dv[0], dv[1]);
        } else if (MODE == 1) {
            ou = make_double2(sanitize_val(ua + dt * du[0], flag), sanitize_val(ub + dt * du[1], flag));
            ov = make_double2(sanitize_val(va + dt * dv[0], flag), sanitize_val(vb + dt * dv[1], flag));
        } else {
            ou = make_double2(sanitize_val(0.5 * (au.x + ua + dt * du[0]), flag),
                              sanitize_val(0.5 * (au.y + ub + dt * du[1]), flag));
            ov = make_double2(sanitize_val(0.5 * (av.x + va + dt * dv[0]), flag),
                              sanitize_val(0.5 * (av.y + vb + dt * dv[1]), flag));
        }
        if (st_a && st_b) {
            *reinterpret_cast<double2 *>(uout + o) = ou;
            *reinterpret_cast<double2 *>(vout + o) = ov;
        } else if (st_a) {
            uout[o] = ou.x;
            vout[o] = ov.x;
        } else if (st_b) {
            uout[o + 1] = ou.y;
            vout[o + 1] = ov.y;
        }
        if (DIV) { // divergence of cell row jj-1 (its per-value guards are identities: u*, v* were sanitised above)
            const double uright = shfl_dn(ou.x); // u*(a+2, jj): an output of the lane above, or lane 31's extra column
            if (jj > y0) {
                const double dA = fin0((pun.y - pun.x) * k.idx + (ov.x - pvn.x) * k.idy);
                const double dB = fin0((urn - pun.y) * k.idx + (ov.y - pvn.y) * k.idy);
                double *pr = rhs + o - P;
                if (cA && cB) {
                    *reinterpret_cast<double2 *>(pr) = make_double2(fin0(dA * inv_dt), fin0(dB * inv_dt));
                    dacc[0] += dA * dA;
                    dacc[0] += dB * dB;
                } else if (cA) {
                    pr[0] = fin0(dA * inv_dt);
                    dacc[0] += dA * dA;
                } else if (cB) {
                    pr[1] = fin0(dB * inv_dt);
                    dacc[0] += dB * dB;
                }
            }
            pun = ou, pvn = ov, urn = uright;
        }
    }
    if (flag) sc->nonfinite_any = 1;
    if (DIV) {
        // Sum of squares over the grid, one partial per WARP (the warps of a block are independent of each other here and
        // the ones outside the region have left): shuffle tree -> slot of this warp -> the last warp to arrive (ticket) adds
        // the slots in a fixed order.  Deterministic like grid_sum; k_div_fix_pre adds its part and finishes div_before.
        const int nwx = (XB - XA) / MARCH_COLS + 1, nlive = nwx * (int)gridDim.y;
        double x = dacc[0];
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) x += __shfl_down_sync(0xffffffffu, x, o2);
        int last = 0;
        if (lane == 0) {
            partials[wx + nwx * (int)blockIdx.y] = x;
            __threadfence();
            last = atomicAdd(&sc->ticket[0], 1u) == (unsigned)(nlive - 1);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            double t = 0.0;
            for (int i = lane; i < nlive; i += 32) t += __ldcg(&partials[i]);
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) t += __shfl_down_sync(0xffffffffu, t, o2);
            if (lane == 0) sc->ticket[0] = 0u, sc->div_before = t;
        }
    }
}

// The cells k_rhs_march<4> leaves open, behind the boundary pass on u*, v*: the rows below YA and from YB up, the top row of
// every strip of RY rows, and in the other rows the columns left of XA and from XB on.  Exactly k_divergence<PRE> on those
// cells; adds its sum of squares to the marching kernel's, finishes div_before (time_integrator.cpp:162-171) and sets
// s.p(0,0) = 0 (:186).  The blocks of the full rows come last: on a slab below another one the top cell row reads the
// neighbour's v* row and waits for it there.
__device__ __forceinline__ bool dfp_full_row(int r, int YA, int YB, int RY) { return r < YA || r >= YB || (r - YA + 1) % RY == 0; }
__global__ void __launch_bounds__(256)
    k_div_fix_pre(Geom g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ rhs, double *p,
                  Scalars *sc, double *partials, int XA, int XB, int YA, int YB, int RY, WaitArgs hw) { pdl_begin();
    const int nside = (XA - 1) + (g.nx - XB + 1);                // columns 1 .. XA-1 and XB .. nx
    const int ntop = (YB - YA) / RY;                             // strip tops below YB: YA + m RY - 1, m = 1 .. ntop
    const int nfull = (YA - 1) + ntop + (g.ny - YB + 1);         // full rows
    const long long n_side = (long long)g.ny * nside, ntot = n_side + (long long)nfull * g.nx;
    const long long chunk = (long long)blockDim.x * gridDim.x;
    // row ny is the last full row: the blocks that can reach its cells wait for the neighbour's v* row (row slabs)
    {
        const long long first = (long long)blockIdx.x * blockDim.x;
        bool reach = false;
        for (long long b = first; b < ntot; b += chunk)
            if (b + blockDim.x > ntot - g.nx) reach = true;
        p2p_wait_block(hw, false, reach, sc);
    }
    const double inv_dt = sc->inv_dt;
    double acc[1] = {0.0};
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < ntot; k += chunk) {
        int ii, jj;
        if (k < n_side) {
            jj = 1 + (int)(k / nside);
            const int c = (int)(k - (long long)(jj - 1) * nside);
            ii = c < XA - 1 ? 1 + c : XB + (c - (XA - 1));
            if (dfp_full_row(jj, YA, YB, RY)) continue; // that cell goes with its row
        } else {
            const long long q = k - n_side;
            const int rr = (int)(q / g.nx);
            ii = 1 + (int)(q - (long long)rr * g.nx);
            jj = rr < YA - 1 ? 1 + rr : (rr < YA - 1 + ntop ? YA + (rr - (YA - 1) + 1) * RY - 1 : YB + (rr - (YA - 1) - ntop));
        }
        const size_t o = gidx(g, ii, jj);
        const double uR = fin0(u[o + 1]), uL = fin0(u[o]), vT = fin0(v[o + g.P]), vB = fin0(v[o]);
        double d = fin0((uR - uL) * g.idx + (vT - vB) * g.idy);
        acc[0] += d * d;
        rhs[o] = fin0(d * inv_dt);
    }
    if (p && blockIdx.x == 0 && threadIdx.x == 0 && g.j0 == 0) p[gidx(g, 1, 1)] = 0.0; // s.p(0,0) = 0 before the solve
    const double ncell = (double)(g.nx * g.ny_glob);
    grid_sum<1>(acc, partials, &sc->ticket[0], [=](double(&t)[1]) {
        const double total = sc->div_before + t[0];
        sc->div_before = (g.ny == g.ny_glob) ? sqrt(total / ncell) : total; // slabs finish after the allreduce
    });
}
