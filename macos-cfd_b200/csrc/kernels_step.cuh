// kernels_step.cuh -- boundary conditions, CFL/dt, the fused advection+diffusion+RK kernels, divergence and
// pressure-gradient projection.  Every arithmetic expression keeps the reference's operand order; the
// comment on each device function names the reference lines it restates.
#pragma once
#include "common.cuh"

// =========================================================================================================
// Boundary conditions  (bc.cpp:31-380).  The reference applies each field's BC in two ordered passes:
// left/right over rows, then bottom/top over columns (which therefore owns the corner faces), with the
// Outflow copies of the second pass reading values the first pass wrote.  Two launches keep that order:
// k_bc_lr (one thread per row: u, v and p of that row) and k_bc_bt (one thread per column).
// `sanitize` additionally applies sanitize_field (time_integrator.cpp:8-27) to the ghost cells, which no
// producing kernel writes (interior values are sanitised where they are produced).
// =========================================================================================================

// bc.cpp:7-29
__device__ __forceinline__ double jet_profile(const BcD &bc, double y) {
    double mask = (fabs(y - bc.jet_center) <= 0.5 * bc.jet_width) ? 1.0 : 0.0;
    return bc.left.inflow_u * mask;
}

// One side of one line: Wall -> 0, Moving -> mov, Inflow -> inf, Outflow -> copy of `inner`, Periodic -> nothing.
__device__ __forceinline__ bool side_value(int type, double mov, double inf, double inner, double &val) {
    switch (type) {
    case BC_WALL: val = 0.0; return true;
    case BC_MOVING: val = mov; return true;
    case BC_INFLOW: val = inf; return true;
    case BC_OUTFLOW: val = inner; return true;
    default: return false;
    }
}

// left/right pass on one row; row points at raw column 0; the right boundary line is raw column cR.
__device__ __forceinline__ void bc_row(double *row, int cR, bool both, int ltype, double lmov, double linf,
                                       int rtype, double rmov, double rinf, bool sanitize, int &flag) {
    if (sanitize) {
        row[0] = sanitize_val(row[0], flag);
        row[cR + 1] = sanitize_val(row[cR + 1], flag);
    }
    double val;
    if (!both && side_value(ltype, lmov, linf, row[2], val)) {
        row[1] = val;
        row[0] = val;
    }
    if (side_value(rtype, rmov, rinf, row[cR - 1], val)) {
        row[cR] = val;
        row[cR + 1] = val;
    }
    if (both) { // bc.cpp:95-107 / :252-264 -- the boundary columns swap on every call
        double lb = row[1], li = row[2], rb = row[cR], ri = row[cR - 1];
        row[0] = ri;
        row[1] = rb;
        row[cR] = lb;
        row[cR + 1] = li;
    }
}

// Row slabs: the pass also covers the 2 halo rows towards a neighbouring slab.  It is row-local, so this reproduces
// exactly what the neighbour does to its own rows and the halo rows stay current without an exchange.
// blockIdx.y selects the field (0 = u, 1 = v, 2 = p): one thread per row AND field keeps the chain of dependent global
// accesses of a thread short -- the pass is pure latency (7.5 us with one thread per row, all three fields, at 1024^2).
// halo = 0: own rows only -- for fields whose halo rows are about to arrive from the neighbours with the pass already
// applied (a one-way send may land before this kernel runs; working on such a row a second time would, with periodic
// sides, swap its boundary columns back).
__global__ void k_bc_lr(Geom g, BcD bc, double *u, double *v, double *p, int sanitize, Scalars *sc, int halo) { pdl_begin();
    const int jlo = (g.bottom || !halo) ? 1 : -1;
    int jj = blockIdx.x * blockDim.x + threadIdx.x + jlo;
    const int f = blockIdx.y;
    int flag = 0;
    bool both = bc.left.type == BC_PERIODIC && bc.right.type == BC_PERIODIC;
    if (f == 0 && u && jj <= ((g.top || !halo) ? g.ny : g.ny + 2)) { // apply_bc_u rows, bc.cpp:38-107: u is the normal component here
        double y = ((g.j0 + jj - 1) + 0.5) * g.dy;
        bc_row(u + gidx(g, 0, jj), g.nx + 1, both, bc.left.type, 0.0, jet_profile(bc, y), bc.right.type, 0.0,
               bc.right.inflow_u, sanitize, flag);
    }
    if (f == 1 && v && jj <= (g.top ? g.ny + 1 : (halo ? g.ny + 2 : g.ny))) { // apply_bc_v rows, bc.cpp:192-264: tangential; left Inflow gives 0
        bc_row(v + gidx(g, 0, jj), g.nx, both, bc.left.type, bc.left.moving, 0.0, bc.right.type,
               bc.right.moving, bc.right.inflow_v, sanitize, flag);
    }
    if (f == 2 && p && jj >= 1 && jj <= g.ny) { // apply_bc_p rows, bc.cpp:340-358
        double *row = p + gidx(g, 0, jj);
        if (sanitize) {
            row[0] = sanitize_val(row[0], flag);
            row[g.nx + 1] = sanitize_val(row[g.nx + 1], flag);
        }
        double first = row[1], last = row[g.nx];
        row[0] = both ? last : first;
        row[g.nx + 1] = both ? first : last;
    }
    if (flag) sc->nonfinite_any = 1;
}

// bottom/top pass on one column; col points at raw row 0 of that column; rT = top boundary line.
// Both sides Periodic (bc.cpp:171-183 / :318-330) swaps the boundary rows on every call: ghost and boundary row of either
// end take the pre-pass values of the other end.  On row slabs the two ends live on the first and the last slab: `far`
// then points at this column of the two rows the OTHER end holds (far[0], far[farP]: its boundary-side pair in the order
// inner, boundary for the top end / boundary, inner for the bottom end), exchanged in front of this kernel.
__device__ __forceinline__ void bc_col(double *col, size_t P, int rT, bool both, bool do_b, bool do_t, int btype,
                                       double bmov, double binf, int ttype, double tmov, double tinf,
                                       const double *far = nullptr, size_t farP = 0) {
    double val;
    if (both) {
        if (do_b && do_t) {
            double bb = col[P], bi = col[2 * P], tb = col[(size_t)rT * P], ti = col[(size_t)(rT - 1) * P];
            col[0] = ti;
            col[P] = tb;
            col[(size_t)rT * P] = bb;
            col[(size_t)(rT + 1) * P] = bi;
        } else if (do_b) { // first slab: far = the last slab's rows rT-1, rT
            col[0] = far[0];
            col[P] = far[farP];
        } else if (do_t) { // last slab: far = the first slab's rows 1, 2
            col[(size_t)rT * P] = far[0];
            col[(size_t)(rT + 1) * P] = far[farP];
        }
        return;
    }
    if (do_b && side_value(btype, bmov, binf, col[2 * P], val)) {
        col[P] = val;
        col[0] = val;
    }
    if (do_t && side_value(ttype, tmov, tinf, col[(size_t)(rT - 1) * P], val)) {
        col[(size_t)rT * P] = val;
        col[(size_t)(rT + 1) * P] = val;
    }
}

// far (row slabs with both bottom/top Periodic, else NULL): 5 rows of pitch g.P received from the slab at the other end of
// the domain -- u (2 rows), v (2 rows), p (1 row), see bc_col.
__global__ void k_bc_bt(Geom g, BcD bc, double *u, double *v, double *p, int sanitize, Scalars *sc, const double *far) { pdl_begin();
    int ii = blockIdx.x * blockDim.x + threadIdx.x; // raw column 0 .. nx+2
    const int f = blockIdx.y;                        // field, as in k_bc_lr
    int flag = 0;
    bool both = bc.bottom.type == BC_PERIODIC && bc.top.type == BC_PERIODIC;
    const size_t P = g.P;
    const double *farc = far ? far + CFD_XOFF + ii : nullptr; // this column in the received rows
    if (f == 0 && u && ii <= g.nx + 2) {
        double *col = u + gidx(g, ii, 0);
        if (sanitize) { // ghost rows (true domain ghosts only)
            if (g.bottom) col[0] = sanitize_val(col[0], flag);
            if (g.top) col[(size_t)(g.ny + 1) * P] = sanitize_val(col[(size_t)(g.ny + 1) * P], flag);
        }
        if (ii >= 1 && ii <= g.nx + 1) // apply_bc_u columns, bc.cpp:111-183: tangential
            bc_col(col, P, g.ny, both, g.bottom, g.top, bc.bottom.type, bc.bottom.moving, bc.bottom.inflow_u,
                   bc.top.type, bc.top.moving, bc.top.inflow_u, farc, P);
    }
    if (f == 1 && v && ii <= g.nx + 1) {
        double *col = v + gidx(g, ii, 0);
        if (sanitize) {
            if (g.bottom) col[0] = sanitize_val(col[0], flag);
            if (g.top) col[(size_t)(g.ny + 2) * P] = sanitize_val(col[(size_t)(g.ny + 2) * P], flag);
        }
        if (ii >= 1 && ii <= g.nx) // apply_bc_v columns, bc.cpp:268-330: normal (Wall and Moving give 0)
            bc_col(col, P, g.ny + 1, both, g.bottom, g.top, bc.bottom.type, 0.0, bc.bottom.inflow_v, bc.top.type,
                   0.0, bc.top.inflow_v, farc ? farc + 2 * P : nullptr, P);
    }
    if (f == 2 && p && ii <= g.nx + 1) {
        double *col = p + gidx(g, ii, 0);
        if (sanitize) {
            if (g.bottom) col[0] = sanitize_val(col[0], flag);
            if (g.top) col[(size_t)(g.ny + 1) * P] = sanitize_val(col[(size_t)(g.ny + 1) * P], flag);
        }
        if (ii >= 1 && ii <= g.nx) { // apply_bc_p columns, bc.cpp:361-379
            const bool whole = g.bottom && g.top;
            double first = (both && !whole && g.top) ? farc[4 * P] : col[P];                    // p of the domain's first row
            double last = (both && !whole && g.bottom) ? farc[4 * P] : col[(size_t)g.ny * P];   // ... and of its last row
            if (g.bottom) col[0] = both ? last : first;
            if (g.top) col[(size_t)(g.ny + 1) * P] = both ? first : last;
        }
    }
    if (flag) sc->nonfinite_any = 1;
}

// =========================================================================================================
// compute_cfl (metrics.cpp:6-50) + the dt logic of TimeIntegrator::step (time_integrator.cpp:55-66)
// =========================================================================================================

// max |q| over the faces ii in [1,IX], jj in [1,JY]; block = 256 threads = 512 columns as double2
__global__ void k_absmax(Geom g, const double *__restrict__ u, const double *__restrict__ v, Scalars *sc) { pdl_begin();
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x; // double2 index along the pitch
    const int ii = 2 * c2 - CFD_XOFF;                     // raw column of .x; .y is ii+1
    double mu = 0.0, mv = 0.0;
    if (2 * c2 + 1 < g.P) {
        const int JYv = v_rows(g);
        for (int jj = 1 + blockIdx.y; jj <= JYv; jj += gridDim.y) {
            const size_t off = (size_t)(jj + CFD_YOFF) * g.P + 2 * c2;
            if (jj <= g.ny) {
                double2 a = *reinterpret_cast<const double2 *>(u + off);
                if (ii >= 1 && ii <= g.nx + 1) mu = max2(mu, fabs(a.x));
                if (ii + 1 >= 1 && ii + 1 <= g.nx + 1) mu = max2(mu, fabs(a.y));
            }
            double2 b = *reinterpret_cast<const double2 *>(v + off);
            if (ii >= 1 && ii <= g.nx) mv = max2(mv, fabs(b.x));
            if (ii + 1 >= 1 && ii + 1 <= g.nx) mv = max2(mv, fabs(b.y));
        }
    }
    block_max_to_global(mu, &sc->umax_bits);
    block_max_to_global(mv, &sc->vmax_bits);
}

// One thread: dt from umax/vmax exactly as the host code does, and reset of the per-step accumulators.
__global__ void k_dt(Geom g, double Re, double CFL, double dt_override, Scalars *sc, int cfl_only) { pdl_begin();
    double umax = __longlong_as_double((long long)sc->umax_bits);
    double vmax = __longlong_as_double((long long)sc->vmax_bits);
    sc->umax_bits = 0ull;
    sc->vmax_bits = 0ull;
    // metrics.cpp:26-49
    double dt_diff = 0.5 * Re * min2(g.dx * g.dx, g.dy * g.dy);
    double dt_adv = 1e30;
    if (umax > 1e-12 && umax < 1e6) dt_adv = min2(dt_adv, g.dx / umax);
    if (vmax > 1e-12 && vmax < 1e6) dt_adv = min2(dt_adv, g.dy / vmax);
    if (dt_adv >= 1e30) dt_adv = dt_diff;
    double dt_min = 1e-6;
    double dt = max2(min2(dt_adv, dt_diff), dt_min);
    dt = CFL * dt;
    dt = clamp3(dt, dt_min, 1e-3);
    if (!cfl_only) { // time_integrator.cpp:55-66
        double dt_cfl = dt;
        if (dt_override > 0.0) dt = min2(dt_cfl, dt_override);
        double dt_cap = (dt_override > 0.0) ? dt_override : 1e-3;
        if (!finite_d(dt) || dt <= 0.0) dt = min2(1e-3, dt_diff);
        dt = clamp3(dt, 1e-8, dt_cap);
    }
    sc->dt = dt;
    sc->inv_dt = 1.0 / dt; // time_integrator.cpp:179 "(1.0 / dt)"
    sc->div_before = 0.0;
    sc->div_after = 0.0;
    sc->nonfinite_any = 0;
    sc->nonfinite_p = 0;
    sc->nonfinite_res = 0;
    sc->breakdown_iter = -1;
    sc->iters = 0;
    sc->done = 0;  // k_solve_begin's part (pressure.cpp:117): "double res = 1e9" survives vcycles == 0
    sc->res = 1e9;
}

// =========================================================================================================
// Advection + diffusion + Runge-Kutta update, fused (advection.cpp:33-421, diffusion.cpp:5-39,
// time_integrator.cpp:71-143).  du_dt / dv_dt are never materialised in the step.
// =========================================================================================================

// advection.cpp:10-12
__device__ __forceinline__ double minmod(double a, double b) {
    return (a * b <= 0.0) ? 0.0 : (fabs(a) < fabs(b) ? a : b);
}
// MUSCL/minmod face flux between samples c and p1 of the stencil (m1, c, p1, p2), carried by w:
// advection.cpp:59-69 for the "+" face, :72-81 for the "-" face (same expression one sample to the left),
// clip_face_states = :15-31.
__device__ __forceinline__ double muscl_flux(double m1, double c, double p1, double p2, double w) {
    double sc = minmod(c - m1, p1 - c);
    double sp = minmod(p1 - c, p2 - p1);
    double qL = c + 0.5 * sc;
    double qR = p1 - 0.5 * sp;
    double lo = min2(m1, min2(c, p1));
    double hi = max2(m1, max2(c, p1));
    qL = clamp3(qL, lo, hi);
    qR = clamp3(qR, lo, hi);
    return (w > 0.0 ? qL : qR) * w;
}
// first-order upwind flux of the outer ring, e.g. advection.cpp:133
__device__ __forceinline__ double upwind1(double a, double b, double w) { return (w > 0.0 ? a : b) * w; }

struct RhsConst {
    double idx, idy, idx2, idy2, invRe;
};

// diffuse_u / diffuse_v as free functions (diffusion.cpp:5-39): out += invRe * laplacian(q) on the advection
// interior set ii in [2, IX-1], jj in [2, JY-1]; nothing else is touched.
__global__ void k_diffuse_add(Geom g, const double *__restrict__ q, double *out, double invRe, int IX, int JY) { pdl_begin();
    const int ii = 2 + blockIdx.x * blockDim.x + threadIdx.x, jj = 2 + blockIdx.y;
    if (ii > IX - 1 || jj > JY - 1) return;
    const size_t o = gidx(g, ii, jj);
    double lap = (q[o + 1] - 2.0 * q[o] + q[o - 1]) * g.idx2 + (q[o + g.P] - 2.0 * q[o] + q[o - g.P]) * g.idy2;
    out[o] += invRe * lap;
}

// d(q)/dt at one face.  qc/uc/vc point at (ii,jj) of the q/u/v tiles (row stride S).  In raw indices both
// advect_u and advect_v carry q with x-face velocities u(ii,jj), u(ii-1,jj) and y-face velocities v(ii,jj),
// v(ii,jj-1) (advection.cpp:68,80,101,113 and :263,275,296,308).  ex/ey = -1/0/+1 mark the first-order ring
// (advection.cpp:124-225, :319-419); corner faces stay 0 and get no diffusion (diffusion.cpp:12-13).
template <bool DIFFUSE = true>
__device__ __forceinline__ double rhs_point(const double *qc, const double *uc, const double *vc, int S, int ex,
                                            int ey, const RhsConst &k) {
    if (ex != 0 && ey != 0) return 0.0;
    const double c = qc[0];
    const double wxp = uc[0], wxm = uc[-1], wyp = vc[0], wym = vc[-S];
    double d;
    if (ex == 0 && ey == 0) {
        double fxp = muscl_flux(qc[-1], c, qc[1], qc[2], wxp);
        double fxm = muscl_flux(qc[-2], qc[-1], c, qc[1], wxm);
        double fyp = muscl_flux(qc[-S], c, qc[S], qc[2 * S], wyp);
        double fym = muscl_flux(qc[-2 * S], qc[-S], c, qc[S], wym);
        double conv = (fxp - fxm) * k.idx + (fyp - fym) * k.idy;
        double lap = (qc[1] - 2.0 * c + qc[-1]) * k.idx2 + (qc[S] - 2.0 * c + qc[-S]) * k.idy2;
        d = -conv;
        if (DIFFUSE) d = d + k.invRe * lap; // du_dt += invRe * lap
    } else {
        double fxp = (ex == 1) ? 0.0 : upwind1(c, qc[1], wxp);
        double fxm = (ex == -1) ? 0.0 : upwind1(qc[-1], c, wxm);
        double fyp = (ey == 1) ? 0.0 : upwind1(c, qc[S], wyp);
        double fym = (ey == -1) ? 0.0 : upwind1(qc[-S], c, wym);
        double conv = (fxp - fxm) * k.idx + (fyp - fym) * k.idy;
        d = -conv;
    }
    return d;
}

// MODE 0: write the sanitised derivatives (stage-level parity entry point cfd_b200_rhs_terms)
// MODE 1: RK stage 1   out = sanitize(in + dt * sanitize(d))                 (time_integrator.cpp:81-103)
// MODE 2: RK stage 2   out = sanitize(0.5 * (acc + in + dt * sanitize(d)))   (time_integrator.cpp:119-147)
//         (in = u1/v1 is the field the derivative is taken of; acc = s.u/s.v; out may be acc itself or another array)
// MODE 3: advect_u / advect_v as free functions: the advective derivative only, not sanitised
#define RHS_TX 32
#define RHS_TY 8
// Up to four rectangles of tiles handled by ONE launch (the strips around the marching kernel's region): a linear block
// index is mapped to (rectangle, tile) through the running tile counts.
struct RhsRects {
    int n;
    int tx0[4], ty0[4], ntx[4], end[4]; // end[k] = number of tiles in rectangles 0..k
};
template <int MODE>
__global__ void __launch_bounds__(RHS_TX *RHS_TY)
    k_rhs_simple(Geom g, RhsConst k, const double *__restrict__ uin, const double *__restrict__ vin, const double *uacc,
                 const double *vacc, double *uout, double *vout, Scalars *sc, RhsRects rc, WaitArgs hw) { pdl_begin();
    constexpr int S = RHS_TX + 4 + 1; // +1: odd stride
    __shared__ double us[(RHS_TY + 4) * S], vs[(RHS_TY + 4) * S];
    int rk = 0, first = 0;
    while (rk < rc.n - 1 && (int)blockIdx.x >= rc.end[rk]) first = rc.end[rk], ++rk;
    const int t = blockIdx.x - first, nty = (rc.end[rk] - first) / rc.ntx[rk];
    const int tyq = t / rc.ntx[rk], txr = t - tyq * rc.ntx[rk];
    const int tyr = tyq + 1 == nty ? 0 : tyq + 1; // the first tile row of a rectangle goes last (see k_rbgs_stream)
    const int x0 = 1 + (rc.tx0[rk] + txr) * RHS_TX, y0 = 1 + (rc.ty0[rk] + tyr) * RHS_TY;
    p2p_wait_block(hw, y0 - 2 < 1, y0 + RHS_TY + 1 > g.ny, sc); // row slabs: tiles that read the neighbours' rows
    const int tid = threadIdx.y * RHS_TX + threadIdx.x;
    const int max_col = g.P - CFD_XOFF - 1, max_row = g.rows - CFD_YOFF - 1;
    for (int t = tid; t < (RHS_TY + 4) * (RHS_TX + 4); t += RHS_TX * RHS_TY) {
        int ly = t / (RHS_TX + 4), lx = t - ly * (RHS_TX + 4);
        int ii = x0 - 2 + lx, jj = y0 - 2 + ly;
        double a = 0.0, b = 0.0;
        if (ii <= max_col && jj <= max_row) {
            size_t o = gidx(g, ii, jj);
            a = uin[o];
            b = vin[o];
        }
        us[ly * S + lx] = a;
        vs[ly * S + lx] = b;
    }
    __syncthreads();
    const int ii = x0 + threadIdx.x, jj = y0 + threadIdx.y;
    const int c = (threadIdx.y + 2) * S + threadIdx.x + 2;
    const double dt = (MODE == 0 || MODE == 3) ? 0.0 : sc->dt;
    int flag = 0;
    const int JYv = v_rows(g);
    if (ii <= g.nx + 1 && jj <= g.ny) { // u faces: ii in [1, nx+1], jj in [1, ny]
        int ex = (ii == 1) ? -1 : (ii == g.nx + 1) ? 1 : 0;
        int ey = (g.bottom && jj == 1) ? -1 : (g.top && jj == g.ny) ? 1 : 0;
        double d = (MODE == 3) ? rhs_point<false>(us + c, us + c, vs + c, S, ex, ey, k)
                               : sanitize_val(rhs_point(us + c, us + c, vs + c, S, ex, ey, k), flag);
        size_t o = gidx(g, ii, jj);
        if (MODE == 0 || MODE == 3) uout[o] = d;
        if (MODE == 1) uout[o] = sanitize_val(us[c] + dt * d, flag);
        if (MODE == 2) uout[o] = sanitize_val(0.5 * (uacc[o] + us[c] + dt * d), flag);
    }
    if (ii <= g.nx && jj <= JYv) { // v faces: ii in [1, nx], jj in [1, ny+1]
        int ex = (ii == 1) ? -1 : (ii == g.nx) ? 1 : 0;
        int ey = (g.bottom && jj == 1) ? -1 : (g.top && jj == g.ny + 1) ? 1 : 0;
        double d = (MODE == 3) ? rhs_point<false>(vs + c, us + c, vs + c, S, ex, ey, k)
                               : sanitize_val(rhs_point(vs + c, us + c, vs + c, S, ex, ey, k), flag);
        size_t o = gidx(g, ii, jj);
        if (MODE == 0 || MODE == 3) vout[o] = d;
        if (MODE == 1) vout[o] = sanitize_val(vs[c] + dt * d, flag);
        if (MODE == 2) vout[o] = sanitize_val(0.5 * (vacc[o] + vs[c] + dt * d), flag);
    }
    if (flag) sc->nonfinite_any = 1;
}

// =========================================================================================================
// divergence (project.cpp:7-42), fused with the serial sum-of-squares loop of TimeIntegrator::step
// (time_integrator.cpp:162-171 / :211-220) and -- PRE only -- the 1/dt scaling (:174-182) and the p(0,0) pin
// (:186).  POST leaves the unscaled divergence in rhs (:210), which callers can observe.
// =========================================================================================================
template <int PRE>
__global__ void __launch_bounds__(256)
    k_divergence(Geom g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ rhs,
                 double *p, Scalars *sc, double *partials, int want_sum, int rpb, int want_max, WaitArgs hw) { pdl_begin();
    p2p_wait_block(hw, false, 1 + (int)(blockIdx.y + 1) * rpb > g.ny, sc); // row slabs: the block of the top cell row reads v(ny+1)
    const int ii = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    double mu = 0.0, mv = 0.0; // max |u|, max |v| over all faces (compute_cfl of the NEXT step, metrics.cpp:8-24)
    if (ii <= g.nx) {
        const int jbase = 1 + blockIdx.y * rpb; // contiguous chunk of rpb rows per block
        const double inv_dt = PRE ? sc->inv_dt : 1.0;
#pragma unroll 4
        for (int r = 0; r < rpb; ++r) {
            const int jj = jbase + r;
            if (jj > g.ny) break;
            const size_t o = gidx(g, ii, jj);
            const double uRr = u[o + 1], uLr = u[o], vTr = v[o + g.P], vBr = v[o];
            if (!PRE && want_max) {
                mu = max2(mu, fabs(uLr));
                if (ii == g.nx) mu = max2(mu, fabs(uRr));
                mv = max2(mv, fabs(vBr));
                if (jj == g.ny && g.top) mv = max2(mv, fabs(vTr));
            }
            double uR = fin0(uRr), uL = fin0(uLr);
            double vT = fin0(vTr), vB = fin0(vBr);
            double divx = (uR - uL) * g.idx;
            double divy = (vT - vB) * g.idy;
            double d = fin0(divx + divy);
            acc[0] += d * d;
            if (PRE) d = fin0(d * inv_dt);
            rhs[o] = d;
        }
    }
    if (PRE && p && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && g.j0 == 0)
        p[gidx(g, 1, 1)] = 0.0; // s.p(0,0) = 0 before the solve, time_integrator.cpp:186
    if (!PRE && want_max) {
        block_max_to_global(mu, &sc->umax_bits);
        block_max_to_global(mv, &sc->vmax_bits);
    }
    if (want_sum) {
        const double ncell = (double)(g.nx * g.ny_glob); // "(g.nx * g.ny)" is an int product in the reference
        grid_sum<1>(acc, partials, &sc->ticket[0], [=](double(&t)[1]) {
            // single-GPU: finish here; slabs finish after the allreduce (host side)
            if (PRE) sc->div_before = (g.ny == g.ny_glob) ? sqrt(t[0] / ncell) : t[0];
            else sc->div_after = (g.ny == g.ny_glob) ? sqrt(t[0] / ncell) : t[0];
        });
    }
}

// =========================================================================================================
// subtract_grad_p (project.cpp:44-84) with the two pins of time_integrator.cpp:186-188 folded in: p(0,0) is
// read as 0 and written as 0 here.  The inflow re-imposition of :194-202 is exactly what the apply_bc_u call
// that follows writes for a left Inflow side (bc.cpp:53-57), so it needs no kernel of its own.  Outputs are
// sanitised here, equivalent to the final sanitize_field calls (:234-235) because the BC pass in between only
// copies or overwrites values.
// =========================================================================================================
__global__ void __launch_bounds__(256)
    k_project(Geom g, double *__restrict__ u, double *__restrict__ v, double *p, const Scalars *sc, double dt_arg,
              Scalars *flags, int rpb, int pin, int fin_solve) { pdl_begin();
    if (fin_solve && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const double res = flags->res; // last_residual_ = isfinite(res) ? res : 1e9 (pressure.cpp:123, :232)
        if (!finite_d(res)) {
            flags->nonfinite_res = 1;
            flags->res = 1e9;
        }
    }
    // One thread = two adjacent columns (aligned double2) marching up rpb rows; the p pair of the row below is
    // carried in registers, p(ii-1) comes from the left lane.
    const int ii = 1 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const bool active = ii <= g.nx + 1;
    const double dt = sc ? sc->dt : dt_arg;
    const int jbase = 1 + blockIdx.y * rpb;
    const int JYv = v_rows(g);
    const bool pin_here = pin && (g.j0 == 0); // the pins belong to TimeIntegrator::step, not to subtract_grad_p
    const bool u_b = ii + 1 <= g.nx + 1, v_a = ii <= g.nx, v_b = ii + 1 <= g.nx; // which faces of the pair exist
    int flag = 0;
    double2 pb = make_double2(0.0, 0.0);
    if (active) pb = *reinterpret_cast<const double2 *>(p + gidx(g, ii, jbase - 1));
    if (pin_here && jbase - 1 == 1 && ii == 1) pb.x = 0.0;
#pragma unroll 2
    for (int r = 0; r < rpb; ++r) {
        const int jj = jbase + r;
        if (jj > JYv) break;
        const size_t o = gidx(g, ii, jj);
        double2 pc = make_double2(0.0, 0.0);
        if (active) pc = *reinterpret_cast<const double2 *>(p + o);
        if (pin_here && jj == 1 && ii == 1) pc.x = 0.0; // s.p(0,0) = 0 after the solve, time_integrator.cpp:188
        double pl = __shfl_up_sync(0xffffffffu, pc.y, 1);
        if ((threadIdx.x & 31) == 0 && active) {
            pl = p[o - 1];
            if (pin_here && jj == 1 && ii == 2) pl = 0.0; // (never true: ii is odd) kept for clarity of the pin rule
        }
        if (active) {
            if (jj <= g.ny) { // u faces ii, ii+1 of row jj (project.cpp:51-65)
                double2 uu = *reinterpret_cast<const double2 *>(u + o);
                double ga = fin0((pc.x - pl) * g.idx);
                double gb = fin0((pc.y - pc.x) * g.idx);
                uu.x = sanitize_val(uu.x - dt * ga, flag);
                if (u_b) {
                    uu.y = sanitize_val(uu.y - dt * gb, flag);
                    *reinterpret_cast<double2 *>(u + o) = uu;
                } else {
                    u[o] = uu.x;
                }
            }
            if (v_a) { // v faces of row jj (project.cpp:69-83); rows up to ny+1 on the top slab
                double2 vv = *reinterpret_cast<const double2 *>(v + o);
                double ga = fin0((pc.x - pb.x) * g.idy);
                double gb = fin0((pc.y - pb.y) * g.idy);
                vv.x = sanitize_val(vv.x - dt * ga, flag);
                if (v_b) {
                    vv.y = sanitize_val(vv.y - dt * gb, flag);
                    *reinterpret_cast<double2 *>(v + o) = vv;
                } else {
                    v[o] = vv.x;
                }
            }
            if (pin_here && jj == 1 && ii == 1) p[o] = 0.0;
        }
        pb = pc;
    }
    if (flag) flags->nonfinite_any = 1;
}

// =========================================================================================================
// k_project + k_divergence<POST> in ONE pass, OUT OF PLACE (time_integrator.cpp:188-220).  RK stage 2 leaves the
// unprojected velocity in two scratch fields (us, vs); this kernel reads them and p and writes the State's u, v and --
// from the same registers -- the divergence of the projected velocity into s.rhs: a thread that has just projected row
// jj holds its own faces of rows jj-1 and jj, gets the u face right of its pair from the lane above (the last lane of a
// warp recomputes that one face from us, p: nobody overwrites the inputs) and closes cell row jj-1.  Every block also
// projects, without storing, the v row above its chunk of rows to close its top cell row.  40 + 24 -> 48 bytes per cell,
// the final u, v are never re-read.  What the boundary pass behind this kernel changes (time_integrator.cpp:205-207
// rewrites the ring faces) is redone afterwards by k_div_fix: the cell rows at the domain bottom / top and the cell
// columns 1 / nx.  The sum of squares and the CFL maxima (compute_cfl of the next step) are split the same way:
// interior here, ring there.
// =========================================================================================================
#ifndef PD_MINB
#define PD_MINB 4
#endif
#ifndef PD_UNROLL
#define PD_UNROLL 2
#endif
#define PD_STR2(x) #x
#define PD_STR(x) PD_STR2(x)
__global__ void __launch_bounds__(256, PD_MINB)
    k_project_div(Geom g, const double *__restrict__ us, const double *__restrict__ vs, double *p, double *__restrict__ u,
                  double *__restrict__ v, double *__restrict__ rhs, Scalars *sc, double *partials, int rpb, WaitArgs hw) { pdl_begin();
    // chunk of rows of this block; the first chunk goes to the LAST blocks of the grid (see k_rbgs_stream)
    const int by = blockIdx.y + 1 == gridDim.y ? 0 : (int)blockIdx.y + 1;
    // row slabs: the chunks at the slab's ends read the neighbours' p row (and the top one the v* row above the slab)
    p2p_wait_block(hw, by == 0, 1 + (by + 1) * rpb > g.ny, sc);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const double res = sc->res; // last_residual_ = isfinite(res) ? res : 1e9 (pressure.cpp:123, :232)
        if (!finite_d(res)) {
            sc->nonfinite_res = 1;
            sc->res = 1e9;
        }
    }
    const int ii = 1 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31;
    const bool active = ii <= g.nx + 1;
    const double dt = sc->dt, idx = g.idx, idy = g.idy;
    const int jbase = 1 + by * rpb;
    // v rows: up to ny+1 on every slab.  Below another slab that row belongs to the neighbour; it is projected here as well
    // (its vs and p rows were delivered) and stored into the halo row, so that the top cell row closes without waiting
    // for the neighbour's projection -- the same expression on the same operands, hence the same bits.
    const int JYv = g.ny + 1;
    const bool pin_t = g.j0 == 0 && ii == 1;     // the thread that holds p(0,0)
    const bool u_b = ii + 1 <= g.nx + 1, v_a = active && ii <= g.nx, v_b = ii + 1 <= g.nx; // which faces of the pair exist
    const bool ur_own = lane == 31 && ii + 2 <= g.nx + 1; // this lane recomputes the face right of its pair
    // cells this kernel closes (the ring: k_div_fix) and faces the boundary pass cannot change (CFL maxima)
    const bool cA = ii >= 2 && ii <= g.nx - 1, cB = ii + 1 <= g.nx - 1; // cells ii, ii+1 = v faces ii, ii+1
    const bool fuA = ii >= 2 && ii <= g.nx, fuB = ii + 1 <= g.nx;        // u faces ii, ii+1
    const int jf_lo = g.bottom ? 2 : 1, ju_hi = g.top ? g.ny - 1 : g.ny, jv_hi = g.ny, jc_hi = g.top ? g.ny - 1 : g.ny;
    int flag = 0;
    double acc[1] = {0.0};
    double mu = 0.0, mv = 0.0;
    const size_t P = g.P;
    size_t o = gidx(g, ii, jbase);
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 pb = active ? *reinterpret_cast<const double2 *>(p + o - P) : zero2;
    if (pin_t && jbase - 1 == 1) pb.x = 0.0;
    double2 un = zero2, vn = zero2; // projected u, v of the row below
    double ur = 0.0;                // projected u(ii+2) of the row below
    const int jlast = min(jbase + rpb, JYv); // jbase + rpb: the row above the chunk (its v faces only, nothing stored)
_Pragma(PD_STR(unroll PD_UNROLL))
    for (int jj = jbase; jj <= jlast; ++jj, o += P) {
        const bool extra = jj == jbase + rpb;
        const bool urow = !extra && jj <= g.ny;
        double2 pc = zero2, su = zero2, sv = zero2;
        double pl = 0.0, pr = 0.0, sur = 0.0;
        if (active) {
            pc = *reinterpret_cast<const double2 *>(p + o);
            if (urow) su = *reinterpret_cast<const double2 *>(us + o);
            if (v_a) sv = *reinterpret_cast<const double2 *>(vs + o);
            if (lane == 0 && urow) pl = p[o - 1];
            if (ur_own && urow) pr = p[o + 2], sur = us[o + 2];
        }
        if (pin_t && jj == 1) pc.x = 0.0; // s.p(0,0) = 0 after the solve, time_integrator.cpp:188
        const double pls = __shfl_up_sync(0xffffffffu, pc.y, 1);
        if (lane != 0) pl = pls;
        double2 uu = zero2, vv = zero2;
        if (urow && active) { // u faces ii, ii+1 of row jj (project.cpp:51-65)
            uu.x = sanitize_val(su.x - dt * fin0((pc.x - pl) * idx), flag);
            if (u_b) uu.y = sanitize_val(su.y - dt * fin0((pc.y - pc.x) * idx), flag);
            if (u_b) *reinterpret_cast<double2 *>(u + o) = uu;
            else u[o] = uu.x;
            if (jj >= jf_lo && jj <= ju_hi) {
                if (fuA) mu = max2(mu, fabs(uu.x));
                if (fuB) mu = max2(mu, fabs(uu.y));
            }
        }
        double uright = __shfl_down_sync(0xffffffffu, uu.x, 1); // projected u(ii+2, jj)
        if (ur_own) uright = sanitize_val(sur - dt * fin0((pr - pc.y) * idx), flag); // the same expression, recomputed
        if (v_a) { // v faces of row jj (project.cpp:69-83); rows up to ny+1 on the top slab
            vv.x = sanitize_val(sv.x - dt * fin0((pc.x - pb.x) * idy), flag);
            if (v_b) vv.y = sanitize_val(sv.y - dt * fin0((pc.y - pb.y) * idy), flag);
            if (!extra) {
                if (v_b) *reinterpret_cast<double2 *>(v + o) = vv;
                else v[o] = vv.x;
                if (jj >= jf_lo && jj <= jv_hi) {
                    if (cA) mv = max2(mv, fabs(vv.x));
                    if (cB) mv = max2(mv, fabs(vv.y));
                }
            }
        }
        if (pin_t && jj == 1) p[o] = 0.0;
        // divergence of cell row jj-1 (project.cpp:7-42: its per-value guards are identities here, the operands were
        // sanitised above; the guard on the result stays)
        const int jc = jj - 1;
        if (jj > jbase && jc >= jf_lo && jc <= jc_hi) {
            const double dA = fin0((un.y - un.x) * idx + (vv.x - vn.x) * idy);
            const double dB = fin0((ur - un.y) * idx + (vv.y - vn.y) * idy);
            double *pr_ = rhs + o - P;
            if (cA && cB) {
                *reinterpret_cast<double2 *>(pr_) = make_double2(dA, dB);
                acc[0] += dA * dA;
                acc[0] += dB * dB;
            } else if (cA) {
                pr_[0] = dA;
                acc[0] += dA * dA;
            } else if (cB) {
                pr_[1] = dB;
                acc[0] += dB * dB;
            }
        }
        un = uu, vn = vv, ur = uright;
        pb = pc;
    }
    if (flag) sc->nonfinite_any = 1;
    block_max_to_global(mu, &sc->umax_bits);
    block_max_to_global(mv, &sc->vmax_bits);
    grid_sum<1>(acc, partials, &sc->ticket[0], [=](double(&t)[1]) { sc->div_after = t[0]; }); // k_div_fix finishes
}

// The cells k_project_div leaves open, after the boundary pass: cell row 1 on the bottom slab, cell row ny on the top slab,
// then cell columns 1 and nx without the cells of those rows.  Adds its sum of squares
// to the interior part, finishes div_after (time_integrator.cpp:211-220) and takes the CFL maxima over the faces of its
// cells.
__global__ void __launch_bounds__(256)
    k_div_fix(Geom g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ rhs, Scalars *sc,
              double *partials) { pdl_begin();
    const int nrow = (g.bottom ? 1 : 0) + (g.top ? 1 : 0);
    const int jlo = g.bottom ? 2 : 1, jhi = g.top ? g.ny - 1 : g.ny, ncolrows = jhi - jlo + 1; // column cells: rows jlo .. jhi
    const int ntot = nrow * g.nx + 2 * ncolrows;
    double acc[1] = {0.0};
    double mu = 0.0, mv = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ntot; k += gridDim.x * blockDim.x) {
        int ii, jj;
        if (k < nrow * g.nx) {
            const int rr = k / g.nx;
            ii = 1 + k - rr * g.nx;
            jj = (rr == 0 && g.bottom) ? 1 : g.ny;
        } else {
            const int c = k - nrow * g.nx;
            jj = jlo + (c >> 1);
            ii = (c & 1) ? g.nx : 1;
        }
        const size_t o = gidx(g, ii, jj);
        const double uRr = u[o + 1], uLr = u[o], vTr = v[o + g.P], vBr = v[o];
        mu = max2(mu, max2(fabs(uLr), fabs(uRr)));
        mv = max2(mv, max2(fabs(vBr), fabs(vTr)));
        const double d = fin0((fin0(uRr) - fin0(uLr)) * g.idx + (fin0(vTr) - fin0(vBr)) * g.idy);
        acc[0] += d * d;
        rhs[o] = d;
    }
    block_max_to_global(mu, &sc->umax_bits);
    block_max_to_global(mv, &sc->vmax_bits);
    const double ncell = (double)(g.nx * g.ny_glob);
    grid_sum<1>(acc, partials, &sc->ticket[0], [=](double(&t)[1]) {
        const double total = sc->div_after + t[0];
        sc->div_after = (g.ny == g.ny_glob) ? sqrt(total / ncell) : total; // slabs finish after the allreduce
    });
}

// Row slabs, end of a step: one all-gather carries what compute_cfl of the NEXT step and the divergence log need.
__global__ void k_slab_pack(Scalars *sc) { pdl_begin();
    sc->red[0] = __longlong_as_double((long long)sc->umax_bits);
    sc->red[1] = __longlong_as_double((long long)sc->vmax_bits);
    sc->red[2] = sc->div_before; // partial sums of squares
    sc->red[3] = sc->div_after;
}
__global__ void k_slab_unpack(Scalars *sc, const double *gathered, int nranks) { pdl_begin();
    double um = 0.0, vm = 0.0, db = 0.0, da = 0.0;
    for (int r = 0; r < nranks; ++r) { // fixed rank order: identical on every rank
        um = max2(um, gathered[4 * r]);
        vm = max2(vm, gathered[4 * r + 1]);
        db += gathered[4 * r + 2];
        da += gathered[4 * r + 3];
    }
    sc->umax_bits = (unsigned long long)__double_as_longlong(um);
    sc->vmax_bits = (unsigned long long)__double_as_longlong(vm);
    sc->div_before = db;
    sc->div_after = da;
}

// sanitize_field(s.p) (time_integrator.cpp:236); only does work when the solver flagged a non-finite p
__global__ void k_sanitize_if(Geom g, double *a, Scalars *sc) { pdl_begin();
    if (!sc->nonfinite_p) return;
    size_t n = (size_t)g.P * g.rows;
    int flag = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a[i] = sanitize_val(a[i], flag);
    if (flag) sc->nonfinite_any = 1;
}

// =========================================================================================================
// Driver diagnostics: divergence_l2 (metrics.cpp:68-86) and max_velocity (metrics.cpp:52-66)
// =========================================================================================================
__global__ void __launch_bounds__(256)
    k_diag(Geom g, const double *__restrict__ u, const double *__restrict__ v, Scalars *sc, double *partials) { pdl_begin();
    const int ii = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    double m = 0.0;
    if (ii <= g.nx) {
        for (int jj = 1 + blockIdx.y; jj <= g.ny; jj += gridDim.y) {
            const size_t o = gidx(g, ii, jj);
            double ul = u[o], ur = u[o + 1], vb = v[o], vt = v[o + g.P];
            double d = (ur - ul) * g.idx + (vt - vb) * g.idy; // divx*idx + divy*idy
            acc[0] += d * d;
            double uc = 0.5 * (ul + ur), vc = 0.5 * (vb + vt);
            m = max2(m, sqrt(uc * uc + vc * vc));
        }
    }
    block_max_to_global(m, &sc->diag_max_bits);
    grid_sum<1>(acc, partials, &sc->ticket[1], [=](double(&t)[1]) { sc->diag_sum = t[0]; });
}
