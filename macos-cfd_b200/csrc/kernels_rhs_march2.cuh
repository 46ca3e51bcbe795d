// kernels_rhs_march2.cuh -- round-2 rewrite of the marching advection + diffusion + Runge-Kutta kernel (kernels_rhs_march.cuh
// has the algorithm, the exactness proof of the clip-free fast path and the meaning of MODE; the arithmetic here is the same,
// expression by expression).  ncu on the first version: issue-bound (69 % of the issue slots, 344 warp instructions per
// thread-row) with only ~190 of them arithmetic.  What this version removes:
//   * register moves: the y-state of a column is a chain of 4 rows (jj-1 .. jj+2; the slot of row jj-1 takes the prefetch of
//     row jj+3 as soon as the y-Laplacian has read it) and three 2-chains (dy, half slope, y-flux).  The row body is
//     instantiated for the four phases of that rotation (S = 0..3) with compile-time slot indices, so no value ever changes
//     its register; the loop is the four phases back to back (the first version rotated through named variables and unrolled
//     by 3: 17 - 27 moves per row at the back edge);
//   * the slow path inside the loop (4 finiteness tests + 4 call sites + 8 sanitising selects per row): the hot loop stores
//     what it computes and only folds the high words of its results into two running maxima (4 integer instructions per
//     row).  A result that is not finite is the ONLY sign that a clip or a sanitize_field replacement was due (a non-finite
//     derivative makes dt * d and the sums after it non-finite; a finite result proves that every flux it was made from
//     was finite, which is the case the proof in kernels_rhs_march.cuh covers).  A warp that saw one redoes exactly those
//     elements with the reference formula after the loop (march_fixup, out of line; inputs and outputs never alias here);
//   * address arithmetic: every access is a field pointer (a kernel parameter) plus ONE 32-bit running element offset (one
//     IMAD.WIDE per access, no per-field pointer registers); loads are never guarded (lanes right of the pitch use a clamped
//     column: their values feed nothing that is stored);
//   * store variants: the region starts at an odd and ends at an even raw column (checked on the host), so a thread stores
//     both of its columns or none.
//   * memory latency (what ncu showed once the instructions were gone: 4.7 long-scoreboard stall cycles per issue, DRAM at
//     56 %; with rows prefetched ONE row ahead into registers a warp has 1 KB in flight, 16 KB per SM): rows now arrive
//     through a private shared-memory ring per warp, one 16-byte cp.async per lane and field M2_D rows ahead of the register
//     stage (stage 2: also the accumulator rows and lane 31's extra column).  A lane reads back exactly the bytes it copied,
//     so cp.async.wait_group is the only synchronisation.
// In-place stage 2 (uacc == uout, `variant = 2`) cannot defer its slow path and stays with the first version.
#pragma once
#include "kernels_rhs_march.cuh"
#include <type_traits>

#ifndef MARCH2_MINBLOCKS
#define MARCH2_MINBLOCKS 4
#endif
#define M2_D 4 // ring slots per warp = rows in flight ahead of the register stage (= the period of the row phases)

__device__ __forceinline__ void m2_cp16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void m2_cp8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void m2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void m2_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Repairs what the hot loop of one warp left non-finite (all 32 lanes call it together).  o0 = element offset of
// (column a, row y0) in every field; flags: bit 0 = this thread stores, bit 1 = lane 31 forming u* of its first column,
// bit 2 / 3 = this thread closes the cell of its first / second column.  Returns the warp lane's sum of squared divergences
// (MODE 4), recomputed from the repaired fields in the hot loop's order.
template <int MODE>
__device__ __noinline__ double march_fixup(const double *uin, const double *vin, const double *uacc, const double *vacc, double *uout,
                                           double *vout, double *rhs, size_t o0, int P, int nrows, double idx, double idy,
                                           double idx2, double idy2, double invRe, double dt, double inv_dt, int flags,
                                           Scalars *sc) {
    const RhsConst k{idx, idy, idx2, idy2, invRe};
    const bool st = flags & 1, ext = flags & 2, cA = flags & 4, cB = flags & 8;
    int flag = 0;
    auto redo = [&](int f, size_t o) -> double { // one face by the reference formula (clip and sanitising included)
        const double *q = (f ? vin : uin) + o;
        const double d = sanitize_val(rhs_point(q, uin + o, vin + o, P, 0, 0, k), flag);
        if (MODE == 0) return d;
        if (MODE == 1) return sanitize_val(q[0] + dt * d, flag);
        return sanitize_val(0.5 * ((f ? vacc : uacc)[o] + q[0] + dt * d), flag);
    };
    if (st)
        for (int r = 0; r < nrows; ++r)
            for (int e = 0; e < 2; ++e) {
                const size_t o = o0 + (size_t)r * P + e;
                if (!finite_d(uout[o])) uout[o] = redo(0, o);
                if (!finite_d(vout[o])) vout[o] = redo(1, o);
            }
    double acc = 0.0;
    if (MODE == 4) {
        __syncwarp();
        double2 pun = make_double2(0.0, 0.0), pvn = pun;
        double urn = 0.0;
        for (int r = 0; r < nrows; ++r) {
            const size_t o = o0 + (size_t)r * P;
            double2 ou = make_double2(0.0, 0.0), ov = ou;
            if (st) ou = *reinterpret_cast<const double2 *>(uout + o), ov = *reinterpret_cast<const double2 *>(vout + o);
            if (ext) ou.x = redo(0, o);
            const double uright = __shfl_down_sync(0xffffffffu, ou.x, 1);
            if (r > 0) { // project.cpp:7-42 with the scaling of time_integrator.cpp:158-182, as k_divergence<PRE>
                const double dA = fin0((pun.y - pun.x) * idx + (ov.x - pvn.x) * idy);
                const double dB = fin0((urn - pun.y) * idx + (ov.y - pvn.y) * idy);
                double *pr = rhs + o - P;
                if (cA) pr[0] = fin0(dA * inv_dt), acc += dA * dA;
                if (cB) pr[1] = fin0(dB * inv_dt), acc += dB * dB;
            }
            pun = ou, pvn = ov, urn = uright;
        }
    }
    if (flag) sc->nonfinite_any = 1;
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(32 * MARCH_WARPS, MARCH2_MINBLOCKS)
    k_rhs_march2(Geom g, RhsConst k, const double *__restrict__ uin, const double *__restrict__ vin, const double *__restrict__ uacc,
                 const double *__restrict__ vacc, double *__restrict__ uout, double *__restrict__ vout, Scalars *sc, int XA, int XB,
                 int YA, int YB, int RY, WaitArgs hw, double *__restrict__ rhs, double *partials) { pdl_begin();
    constexpr bool HEUN = MODE == 2 || MODE == 4, DIV = MODE == 4;
    const int by = blockIdx.y + 1 == gridDim.y ? 0 : (int)blockIdx.y + 1; // edge strips last (see the first version)
    {
        const int yb0 = YA + by * RY, yb1 = min(yb0 + RY - 1, YB);
        p2p_wait_block(hw, yb0 - 2 < 1, yb1 + 3 > g.ny, sc);
    }
    const int lane = threadIdx.x & 31;
    const int wx = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    const int X0 = XA - 2 + MARCH_COLS * wx;
    if (X0 + 2 > XB) return;
    const int a = X0 + 2 * lane;
    const int y0 = YA + by * RY;
    const int y1 = min(y0 + RY - 1, YB);
    const unsigned P = (unsigned)g.P;
    const int ac = min(a, g.P - CFD_XOFF - 2); // lanes right of the pitch use its last pair (they store nothing)
    const bool st = lane >= 1 && lane <= 30 && a >= XA && a + 1 <= XB;
    const bool ext = DIV && lane == 31 && a <= XB;
    const bool cB = DIV && st && a + 2 <= XB; // cA == st
    const double dt = (MODE == 0) ? 0.0 : sc->dt;
    const double inv_dt = DIV ? sc->inv_dt : 0.0;

    // Every access is <field pointer>[32-bit element offset]: `rel` = offset of (column ac, row jj) in any field (the host
    // checks that a field has fewer than 2^32 elements), advanced by one pitch per row.
    const size_t oa = gidx(g, ac, y0);
    unsigned rel = (unsigned)oa;
    auto ld2 = [](const double *p) { return *reinterpret_cast<const double2 *>(p); };
    // The ring: slot s of a warp holds row jj+3 of u_in, v_in (and row jj of the accumulators, and u_in(a+2, jj) for lane 31)
    // for the row phase S = s; it is refilled M2_D rows ahead as soon as it has been read.
    constexpr int NF = HEUN ? 4 : 2;
    __shared__ __align__(16) double2 ring_all[MARCH_WARPS][M2_D][NF][32];
    __shared__ double xring_all[MARCH_WARPS][M2_D];
    double2(*ring)[NF][32] = ring_all[threadIdx.x >> 5];
    double *xring = xring_all[threadIdx.x >> 5];
    auto fill = [&](int s, unsigned r) { // r = element offset of (ac, jj) of the row phase that will read slot s
        m2_cp16(&ring[s][0][lane], uin + (r + 3u * P));
        m2_cp16(&ring[s][1][lane], vin + (r + 3u * P));
        if (HEUN) m2_cp16(&ring[s][2][lane], uacc + r), m2_cp16(&ring[s][3][lane], vacc + r);
        if (ext) m2_cp8(&xring[s], uin + (r + 2u));
        m2_commit();
    };
#pragma unroll
    for (int s = 0; s < M2_D; ++s) fill(s, rel + (unsigned)s * P);
    // R[f][slot]: rows jj-1 .. jj+2 of field f (0 = u, 1 = v) in rotating slots; DY / HS / FY: q(jj+1) - q(jj), half slope at
    // jj and y-flux through the face below jj, ping-pong
    double2 R[2][4], DY[2][2], HS[2][2], FY[2][2];
    {
        const double *bu = uin + oa, *bv = vin + oa;
        const double2 u0 = ld2(bu - 2 * (size_t)P), v0 = ld2(bv - 2 * (size_t)P);
        R[0][0] = ld2(bu - (size_t)P), R[1][0] = ld2(bv - (size_t)P);
        R[0][1] = ld2(bu), R[1][1] = ld2(bv);
        R[0][2] = ld2(bu + (size_t)P), R[1][2] = ld2(bv + (size_t)P);
        R[0][3] = ld2(bu + 2 * (size_t)P), R[1][3] = ld2(bv + 2 * (size_t)P);
        auto init = [&](double q_2, double q_1, double q0, double q1, double w_below, double &dy, double &hs, double &fy) {
            const double dm2 = q_1 - q_2, dm1 = q0 - q_1, d0 = q1 - q0;
            const double hsm1 = half_minmod(dm2, dm1);
            hs = half_minmod(dm1, d0);
            fy = flux_shared(q_1, hsm1, q0, hs, w_below);
            dy = d0;
        };
        // y-faces are carried by v of the row below (advection.cpp:113, :275)
        init(u0.x, R[0][0].x, R[0][1].x, R[0][2].x, R[1][0].x, DY[0][0].x, HS[0][0].x, FY[0][0].x);
        init(u0.y, R[0][0].y, R[0][1].y, R[0][2].y, R[1][0].y, DY[0][0].y, HS[0][0].y, FY[0][0].y);
        init(v0.x, R[1][0].x, R[1][1].x, R[1][2].x, R[1][0].x, DY[1][0].x, HS[1][0].x, FY[1][0].x);
        init(v0.y, R[1][0].y, R[1][1].y, R[1][2].y, R[1][0].y, DY[1][0].y, HS[1][0].y, FY[1][0].y);
    }
    int mxs = 0;        // running maxima of the results' high words, signed and unsigned: a result is non-finite iff its high
    unsigned mxu = 0u;  // word is >= 0x7ff00000 as a signed or >= 0xfff00000 as an unsigned number
    double dacc = 0.0;
    double2 pun = make_double2(0.0, 0.0), pvn = pun; // u*, v* of the row below
    double urn = 0.0;                                // u*(a + 2) of the row below

    auto row = [&](auto S_, auto first_) {
        constexpr int S = decltype(S_)::value;
        constexpr bool FIRSTROW = decltype(first_)::value;
        constexpr int im = S & 3, i0 = (S + 1) & 3, i1 = (S + 2) & 3, i2 = (S + 3) & 3, pi = S & 1, po = pi ^ 1;
        // ---- y-Laplacian first: the last reader of row jj-1, whose slot then takes row jj+3
        double two[2][2], ly[2][2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            two[f][0] = 2.0 * R[f][i0].x, two[f][1] = 2.0 * R[f][i0].y;
            ly[f][0] = (R[f][i1].x - two[f][0] + R[f][im].x) * k.idy2;
            ly[f][1] = (R[f][i1].y - two[f][1] + R[f][im].y) * k.idy2;
        }
        // ---- the ring slot of this phase: the copies issued M2_D rows ago have landed; read it, then refill it
        m2_wait<M2_D - 1>();
        R[0][im] = ring[S][0][lane];
        R[1][im] = ring[S][1][lane];
        double2 au = make_double2(0.0, 0.0), av = au;
        if (HEUN) au = ring[S][2][lane], av = ring[S][3][lane];
        double xq = 0.0;
        if (ext) xq = xring[S];
        fill(S, rel + (unsigned)M2_D * P);
        const double ua = R[0][i0].x, ub = R[0][i0].y, va = R[1][i0].x, vb = R[1][i0].y; // face velocities of this row
        double du[2], dv[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const double2 q0 = R[f][i0], q1 = R[f][i1], q2 = R[f][i2];
            // ---- y direction: one new slope and one new flux per column (advection.cpp:85-114, :247-276)
            const double dy1a = q2.x - q1.x, dy1b = q2.y - q1.y;
            const double hs1a = half_minmod(DY[f][pi].x, dy1a), hs1b = half_minmod(DY[f][pi].y, dy1b);
            const double fya = flux_shared(q0.x, HS[f][pi].x, q1.x, hs1a, va);
            const double fyb = flux_shared(q0.y, HS[f][pi].y, q1.y, hs1b, vb);
            // ---- x direction on row jj, shared across the warp (advection.cpp:52-81, :280-309)
            const double qa = q0.x, qb = q0.y;
            double qa_n = shfl_dn(qa);       // q(a+2)
            if (DIV && f == 0 && lane == 31) qa_n = xq;
            const double qb_p = shfl_up(qb); // q(a-1)
            const double dxp = qa - qb_p, dxa = qb - qa, dxb = qa_n - qb;
            const double hsa = half_minmod(dxp, dxa), hsb = half_minmod(dxa, dxb);
            const double hsa_n = shfl_dn(hsa);
            const double fxa = flux_shared(qa, hsa, qb, hsb, ua);     // face a+1/2, carried by u(a, jj)
            const double fxb = flux_shared(qb, hsb, qa_n, hsa_n, ub); // face b+1/2, carried by u(b, jj)
            const double fxb_p = shfl_up(fxb);                        // face a-1/2
            // ---- flux divergence + diffusion (advection.cpp:117-118, diffusion.cpp:16-18)
            const double conva = (fxa - fxb_p) * k.idx + (fya - FY[f][pi].x) * k.idy;
            const double convb = (fxb - fxa) * k.idx + (fyb - FY[f][pi].y) * k.idy;
            const double lapa = (qb - two[f][0] + qb_p) * k.idx2 + ly[f][0];
            const double lapb = (qa_n - two[f][1] + qa) * k.idx2 + ly[f][1];
            double da = -conva, db = -convb;
            da = da + k.invRe * lapa;
            db = db + k.invRe * lapb;
            (f ? dv : du)[0] = da;
            (f ? dv : du)[1] = db;
            DY[f][po] = make_double2(dy1a, dy1b);
            HS[f][po] = make_double2(hs1a, hs1b);
            FY[f][po] = make_double2(fya, fyb);
        }
        // ---- Runge-Kutta update (time_integrator.cpp:84-99, :122-143), unsanitised: see the header
        double2 ou, ov;
        if (MODE == 0) {
            ou = make_double2(du[0], du[1]);
            ov = make_double2(dv[0], dv[1]);
        } else if (MODE == 1) {
            ou = make_double2(ua + dt * du[0], ub + dt * du[1]);
            ov = make_double2(va + dt * dv[0], vb + dt * dv[1]);
        } else {
            ou = make_double2(0.5 * (au.x + ua + dt * du[0]), 0.5 * (au.y + ub + dt * du[1]));
            ov = make_double2(0.5 * (av.x + va + dt * dv[0]), 0.5 * (av.y + vb + dt * dv[1]));
        }
        mxs = max(mxs, max(__double2hiint(ou.x), __double2hiint(ou.y)));
        mxu = max(mxu, max((unsigned)__double2hiint(ou.x), (unsigned)__double2hiint(ou.y)));
        mxs = max(mxs, max(__double2hiint(ov.x), __double2hiint(ov.y)));
        mxu = max(mxu, max((unsigned)__double2hiint(ov.x), (unsigned)__double2hiint(ov.y)));
        if (st) {
            *reinterpret_cast<double2 *>(uout + rel) = ou;
            *reinterpret_cast<double2 *>(vout + rel) = ov;
        }
        if (DIV) { // divergence of cell row jj-1 from registers
            const double uright = shfl_dn(ou.x); // u*(a+2, jj): an output of the lane above, or lane 31's extra column
            if (!FIRSTROW) {
                const double dA = (pun.y - pun.x) * k.idx + (ov.x - pvn.x) * k.idy;
                const double dB = (urn - pun.y) * k.idx + (ov.y - pvn.y) * k.idy;
                const double rA = dA * inv_dt, rB = dB * inv_dt; // finite rA proves dA finite: inv_dt is finite and positive
                mxs = max(mxs, max(__double2hiint(rA), __double2hiint(rB)));
                mxu = max(mxu, max((unsigned)__double2hiint(rA), (unsigned)__double2hiint(rB)));
                double *pr = rhs + (rel - P); // cell row jj - 1
                if (cB) {
                    *reinterpret_cast<double2 *>(pr) = make_double2(rA, rB);
                    dacc += dA * dA;
                    dacc += dB * dB;
                } else if (st) {
                    pr[0] = rA;
                    dacc += dA * dA;
                }
            }
            pun = ou, pvn = ov, urn = uright;
        }
        rel += P;
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    row(I0{}, std::true_type{}); // row y0: no cell row below it in this strip
    for (int n = y1 - y0;;) {    // rows left
        if (n < 1) break;
        row(I1{}, std::false_type{});
        if (n < 2) break;
        row(I2{}, std::false_type{});
        if (n < 3) break;
        row(I3{}, std::false_type{});
        if (n < 4) break;
        row(I0{}, std::false_type{});
        n -= 4;
    }
    // ---- slow path, once per warp and only if a result was not finite
    const bool bad = mxs >= 0x7ff00000 || mxu >= 0xfff00000u;
    if (__any_sync(0xffffffffu, bad)) {
        const double fixed = march_fixup<MODE>(uin, vin, uacc, vacc, uout, vout, rhs, oa, (int)P, y1 - y0 + 1, k.idx, k.idy, k.idx2,
                                               k.idy2, k.invRe, dt, inv_dt, (st ? 1 : 0) | (ext ? 2 : 0) | (st && DIV ? 4 : 0) | (cB ? 8 : 0), sc);
        if (DIV) dacc = fixed;
    }
    if (DIV) { // sum of squares: one partial per warp, the last warp to arrive adds them in a fixed order (first version)
        const int nwx = (XB - XA) / MARCH_COLS + 1, nlive = nwx * (int)gridDim.y;
        double x = dacc;
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
