/*
 * cfd_oracle.c -- plain-C restatement of the per-timestep hot path of the reference's 2D MAC-grid
 * solver (src/solver/ of archhie/MacOS-CFD).
 *
 * TEST INFRASTRUCTURE ONLY (see cfdo_api.h): this is the checker that travels with the repo to the GPU
 * box, where /root/reference does not exist.  It is never linked into or called by the product path.
 *
 * PARITY PIN: tests/test_oracle.py checks this file BITWISE (OMP_NUM_THREADS=1) against
 * oracle/_ref/libcfdo_ref.so (the reference's own sources compiled here) stage by stage and over whole
 * multi-step runs for every preset x solver, and against the committed fixtures in tests/golden/ that were
 * generated from oracle/_ref (tests/golden/make_golden.py).  The reference itself ships no tests or golden
 * vectors, so oracle/_ref as built on x86-64 without FMA contraction IS the pin.
 *
 * The code is organised differently from the reference (one generic advection routine, table-driven
 * boundary conditions), but every arithmetic expression keeps the reference's operand order; each
 * function cites the file:line it restates.  Build: -O3 -ffp-contract=off (no FMA), see oracle/Makefile.
 */
#include "cfdo_api.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

/* std::min / std::max / std::clamp exactly as libstdc++ evaluates them (argument order matters for
 * +-0 and NaN): min(a,b) = b<a ? b : a ; max(a,b) = a<b ? b : a ; clamp(x,lo,hi) = x<lo ? lo : hi<x ? hi : x */
static inline double min2(double a, double b) { return (b < a) ? b : a; }
static inline double max2(double a, double b) { return (a < b) ? b : a; }
static inline double clamp3(double x, double lo, double hi) { return (x < lo) ? lo : ((hi < x) ? hi : x); }
static inline double fin0(double x) { return isfinite(x) ? x : 0.0; }

typedef struct {
    int nx, ny;
    double Lx, Ly, dx, dy;
} grid_t;

/* grid.cpp:3-12 */
static grid_t grid_make(int nx, int ny, double Lx, double Ly) {
    grid_t g;
    g.nx = nx;
    g.ny = ny;
    g.Lx = Lx;
    g.Ly = Ly;
    g.dx = Lx / (double)nx;
    g.dy = Ly / (double)ny;
    return g;
}
/* pitches, grid.hpp:11-23 with ng = 1 */
static inline int pitch_u(const grid_t *g) { return g->nx + 3; }
static inline int pitch_c(const grid_t *g) { return g->nx + 2; } /* v, p and every cell-centred array */
static inline size_t count_u(const grid_t *g) { return (size_t)pitch_u(g) * (g->ny + 2); }
static inline size_t count_v(const grid_t *g) { return (size_t)pitch_c(g) * (g->ny + 3); }
static inline size_t count_c(const grid_t *g) { return (size_t)pitch_c(g) * (g->ny + 2); }

/* ------------------------------------------------------------------ boundary conditions */

/* bc.cpp:7-29: top-hat jet profile (the perturbation term is commented out in the reference). */
static double jet_profile(const cfdo_bc *bc, double y) {
    double mask = (fabs(y - bc->jet_center) <= 0.5 * bc->jet_width) ? 1.0 : 0.0;
    return bc->left.inflow_u * mask;
}

enum { ACT_NONE, ACT_SET, ACT_COPY };
/* What one side does to its boundary line: Wall -> 0, Moving -> `moving`, Inflow -> `inflow`,
 * Outflow -> copy the inner neighbour, Periodic -> nothing (bc.cpp switch statements). */
static int side_action(int type, double moving, double inflow, double *val) {
    switch (type) {
    case CFDO_WALL: *val = 0.0; return ACT_SET;
    case CFDO_MOVING: *val = moving; return ACT_SET;
    case CFDO_INFLOW: *val = inflow; return ACT_SET;
    case CFDO_OUTFLOW: return ACT_COPY;
    default: return ACT_NONE;
    }
}

/* Generic two-phase BC for one velocity component, restating apply_bc_u (bc.cpp:31-184) and
 * apply_bc_v (bc.cpp:186-331).  `a` has pitch P.  Left/right act on rows 1..rows_lr: the left side writes
 * columns (1,0) and the right side (cR, cR+1); bottom/top act on columns 1..cols_bt: rows (1,0) and
 * (rT, rT+1).  lr_* / bt_* are the Moving / Inflow values of each side (0 where the reference writes 0). */
typedef struct {
    double l_mov, l_inf, r_mov, r_inf, b_mov, b_inf, t_mov, t_inf;
    int left_inflow_is_jet;
} bc_vals;

static void bc_velocity(const grid_t *g, double *a, int P, int rows_lr, int cR, int cols_bt, int rT,
                        const cfdo_bc *bc, const bc_vals *bv) {
    int both_lr = bc->left.type == CFDO_PERIODIC && bc->right.type == CFDO_PERIODIC;
    int both_bt = bc->bottom.type == CFDO_PERIODIC && bc->top.type == CFDO_PERIODIC;
    for (int jj = 1; jj <= rows_lr; ++jj) {
        double *row = a + (size_t)jj * P;
        double val;
        if (!both_lr) {
            double inflow = bv->left_inflow_is_jet ? jet_profile(bc, ((jj - 1) + 0.5) * g->dy) : bv->l_inf;
            int act = side_action(bc->left.type, bv->l_mov, inflow, &val);
            if (act == ACT_COPY) val = row[2];
            if (act != ACT_NONE) { row[1] = val; row[0] = val; }
        }
        int act = side_action(bc->right.type, bv->r_mov, bv->r_inf, &val);
        if (act == ACT_COPY) val = row[cR - 1];
        if (act != ACT_NONE) { row[cR] = val; row[cR + 1] = val; }
    }
    if (both_lr) { /* bc.cpp:95-107 / :252-264: the two boundary columns swap */
        for (int jj = 1; jj <= rows_lr; ++jj) {
            double *row = a + (size_t)jj * P;
            double lb = row[1], li = row[2], rb = row[cR], ri = row[cR - 1];
            row[0] = ri; row[1] = rb; row[cR] = lb; row[cR + 1] = li;
        }
    }
    for (int ii = 1; ii <= cols_bt; ++ii) {
        double val;
        int act = side_action(bc->bottom.type, bv->b_mov, bv->b_inf, &val);
        if (act == ACT_COPY) val = a[(size_t)2 * P + ii];
        if (act != ACT_NONE) { a[(size_t)1 * P + ii] = val; a[ii] = val; }
        act = side_action(bc->top.type, bv->t_mov, bv->t_inf, &val);
        if (act == ACT_COPY) val = a[(size_t)(rT - 1) * P + ii];
        if (act != ACT_NONE) { a[(size_t)rT * P + ii] = val; a[(size_t)(rT + 1) * P + ii] = val; }
    }
    if (both_bt) { /* bc.cpp:171-183 / :318-330 */
        for (int ii = 1; ii <= cols_bt; ++ii) {
            double bb = a[(size_t)1 * P + ii], bi = a[(size_t)2 * P + ii];
            double tb = a[(size_t)rT * P + ii], ti = a[(size_t)(rT - 1) * P + ii];
            a[ii] = ti; a[(size_t)1 * P + ii] = tb;
            a[(size_t)rT * P + ii] = bb; a[(size_t)(rT + 1) * P + ii] = bi;
        }
    }
}

/* apply_bc_u, bc.cpp:31-184.  u is the NORMAL component on left/right (Wall and Moving both give 0,
 * left Inflow is the jet, right Inflow is right.inflow_u) and TANGENTIAL on bottom/top (Moving gives
 * `moving`, Inflow gives inflow_u; the first interior row is overwritten together with the ghost row). */
static void bc_u(const grid_t *g, double *u, const cfdo_bc *bc) {
    bc_vals bv = {0.0, 0.0, 0.0, bc->right.inflow_u, bc->bottom.moving, bc->bottom.inflow_u,
                  bc->top.moving, bc->top.inflow_u, 1};
    bc_velocity(g, u, pitch_u(g), g->ny, g->nx + 1, g->nx + 1, g->ny, bc, &bv);
}
/* apply_bc_v, bc.cpp:186-331.  v is TANGENTIAL on left/right (left Inflow gives 0, right Inflow gives
 * right.inflow_v; first interior column overwritten) and NORMAL on bottom/top (Wall, Moving -> 0). */
static void bc_v(const grid_t *g, double *v, const cfdo_bc *bc) {
    bc_vals bv = {bc->left.moving, 0.0, bc->right.moving, bc->right.inflow_v, 0.0, bc->bottom.inflow_v,
                  0.0, bc->top.inflow_v, 0};
    bc_velocity(g, v, pitch_c(g), g->ny + 1, g->nx, g->nx, g->ny + 1, bc, &bv);
}
/* apply_bc_p, bc.cpp:333-380: ghost = adjacent interior cell, or the opposite interior cell when both
 * sides of that direction are Periodic.  Corner ghosts are never written. */
static void bc_p(const grid_t *g, double *p, const cfdo_bc *bc) {
    int P = pitch_c(g), nx = g->nx, ny = g->ny;
    int both_lr = bc->left.type == CFDO_PERIODIC && bc->right.type == CFDO_PERIODIC;
    int both_bt = bc->bottom.type == CFDO_PERIODIC && bc->top.type == CFDO_PERIODIC;
    for (int jj = 1; jj <= ny; ++jj) {
        double *row = p + (size_t)jj * P;
        double first = row[1], last = row[nx];
        row[0] = both_lr ? last : first;
        row[nx + 1] = both_lr ? first : last;
    }
    for (int ii = 1; ii <= nx; ++ii) {
        double first = p[(size_t)1 * P + ii], last = p[(size_t)ny * P + ii];
        p[ii] = both_bt ? last : first;
        p[(size_t)(ny + 1) * P + ii] = both_bt ? first : last;
    }
}

/* ------------------------------------------------------------------ advection + diffusion */

/* advection.cpp:10-12 */
static inline double minmod(double a, double b) {
    return (a * b <= 0.0) ? 0.0 : (fabs(a) < fabs(b) ? a : b);
}
/* MUSCL/minmod flux through the face between samples c and p1 of the stencil (m1, c, p1, p2), carried by
 * velocity w: reconstruct the two face states, clip them to [min,max] of (m1,c,p1), upwind on w > 0.
 * Restates advection.cpp:59-69 ("+" face: (m1,c,p1,p2) = (q[i-1],q[i],q[i+1],q[i+2])) and :72-81 ("-" face:
 * the same expression shifted by one sample); clip_face_states is advection.cpp:15-31 without the
 * tvd_debug logging branch. */
static inline double muscl_flux(double m1, double c, double p1, double p2, double w) {
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
/* first-order upwind flux used on the outermost ring, e.g. advection.cpp:133 */
static inline double upwind1(double a, double b, double w) { return (w > 0.0 ? a : b) * w; }

/* One routine for advect_u (advection.cpp:33-226) and advect_v (:228-421).  In raw indices BOTH
 * functions carry the field q with the same four face velocities: x-faces U(ii,jj), U(ii-1,jj) and y-faces
 * V(ii,jj), V(ii,jj-1) -- for q = u that is "u_face = ui, u_face_m = uim1, v_face = v(ii,jj),
 * v_face_m = v(ii,jj-1)" (:68,:80,:101,:113), for q = v it is "v_face = vj, v_face_m = vjm1,
 * u_face = u(ii,jj), u_face_m = u(ii-1,jj)" (:263,:275,:296,:308).  Faces live on ii in [1,IX], jj in [1,JY]
 * (u: IX = nx+1, JY = ny; v: IX = nx, JY = ny+1).  Interior = MUSCL; the outer ring is first order with a
 * zero flux on the outside; the four corner faces stay 0; the output is zero-filled first. */
static void advect(const grid_t *g, const double *q, int Pq, int IX, int JY, const double *u, const double *v,
                   double *out, size_t out_count) {
    double idx = 1.0 / g->dx, idy = 1.0 / g->dy;
    int Pu = pitch_u(g), Pv = pitch_c(g);
    memset(out, 0, out_count * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= JY; ++jj) {
        for (int ii = 1; ii <= IX; ++ii) {
            int edge_x = (ii == 1) ? -1 : (ii == IX) ? 1 : 0;
            int edge_y = (jj == 1) ? -1 : (jj == JY) ? 1 : 0;
            if (edge_x && edge_y) continue; /* corner faces keep 0 */
            const double *c = q + (size_t)jj * Pq + ii;
            double wxp = u[(size_t)jj * Pu + ii], wxm = u[(size_t)jj * Pu + ii - 1];
            double wyp = v[(size_t)jj * Pv + ii], wym = v[(size_t)(jj - 1) * Pv + ii];
            double fxp, fxm, fyp, fym;
            if (!edge_x && !edge_y) { /* interior, advection.cpp:44-120 / :239-315 */
                fxp = muscl_flux(c[-1], c[0], c[1], c[2], wxp);
                fxm = muscl_flux(c[-2], c[-1], c[0], c[1], wxm);
                fyp = muscl_flux(c[-Pq], c[0], c[Pq], c[2 * Pq], wyp);
                fym = muscl_flux(c[-2 * Pq], c[-Pq], c[0], c[Pq], wym);
            } else { /* ring, advection.cpp:124-225 / :319-419 */
                fxp = (edge_x == 1) ? 0.0 : upwind1(c[0], c[1], wxp);
                fxm = (edge_x == -1) ? 0.0 : upwind1(c[-1], c[0], wxm);
                fyp = (edge_y == 1) ? 0.0 : upwind1(c[0], c[Pq], wyp);
                fym = (edge_y == -1) ? 0.0 : upwind1(c[-Pq], c[0], wym);
            }
            double conv = (fxp - fxm) * idx + (fyp - fym) * idy;
            out[(size_t)jj * Pq + ii] = -conv;
        }
    }
}

/* diffuse_u / diffuse_v, diffusion.cpp:5-39: out += invRe * lap on the advection INTERIOR set only. */
static void diffuse(const grid_t *g, const double *q, int Pq, int IX, int JY, double *out, double invRe) {
    double idx2 = 1.0 / (g->dx * g->dx), idy2 = 1.0 / (g->dy * g->dy);
#pragma omp parallel for schedule(static)
    for (int jj = 2; jj <= JY - 1; ++jj)
        for (int ii = 2; ii <= IX - 1; ++ii) {
            const double *c = q + (size_t)jj * Pq + ii;
            double lap = (c[1] - 2.0 * c[0] + c[-1]) * idx2 + (c[Pq] - 2.0 * c[0] + c[-Pq]) * idy2;
            out[(size_t)jj * Pq + ii] += invRe * lap;
        }
}

/* ------------------------------------------------------------------ metrics */

static double absmax_faces(const double *q, int Pq, int IX, int JY) {
    double m = 0.0;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int jj = 1; jj <= JY; ++jj)
        for (int ii = 1; ii <= IX; ++ii) m = max2(m, fabs(q[(size_t)jj * Pq + ii]));
    return m;
}

/* compute_cfl, metrics.cpp:6-50 */
static double cfl_dt(const grid_t *g, const double *u, const double *v, double Re, double CFL) {
    double umax = absmax_faces(u, pitch_u(g), g->nx + 1, g->ny);
    double vmax = absmax_faces(v, pitch_c(g), g->nx, g->ny + 1);
    double dt_diff = 0.5 * Re * min2(g->dx * g->dx, g->dy * g->dy);
    double dt_adv = 1e30;
    if (umax > 1e-12 && umax < 1e6) dt_adv = min2(dt_adv, g->dx / umax);
    if (vmax > 1e-12 && vmax < 1e6) dt_adv = min2(dt_adv, g->dy / vmax);
    if (dt_adv >= 1e30) dt_adv = dt_diff;
    double dt_min = 1e-6;
    double dt = max2(min2(dt_adv, dt_diff), dt_min);
    dt = CFL * dt;
    return clamp3(dt, dt_min, 1e-3);
}

/* divergence, project.cpp:7-42 */
static void div_field(const grid_t *g, const double *u, const double *v, double *rhs) {
    double idx = 1.0 / g->dx, idy = 1.0 / g->dy;
    int Pu = pitch_u(g), P = pitch_c(g);
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            double uR = fin0(u[(size_t)jj * Pu + ii + 1]), uL = fin0(u[(size_t)jj * Pu + ii]);
            double vT = fin0(v[(size_t)(jj + 1) * P + ii]), vB = fin0(v[(size_t)jj * P + ii]);
            double divx = (uR - uL) * idx;
            double divy = (vT - vB) * idy;
            rhs[(size_t)jj * P + ii] = fin0(divx + divy);
        }
}

/* subtract_grad_p, project.cpp:44-84 (boundary faces included; they read p's ghosts) */
static void grad_p_correct(const grid_t *g, double *u, double *v, const double *p, double dt) {
    double idx = 1.0 / g->dx, idy = 1.0 / g->dy;
    int Pu = pitch_u(g), P = pitch_c(g);
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx + 1; ++ii) {
            double gradp = fin0((p[(size_t)jj * P + ii] - p[(size_t)jj * P + ii - 1]) * idx);
            u[(size_t)jj * Pu + ii] -= dt * gradp;
        }
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny + 1; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            double gradp = fin0((p[(size_t)jj * P + ii] - p[(size_t)(jj - 1) * P + ii]) * idy);
            v[(size_t)jj * P + ii] -= dt * gradp;
        }
}

/* divergence_l2, metrics.cpp:68-86 (no finite guards, different association from div_field) */
static double div_l2(const grid_t *g, const double *u, const double *v) {
    double idx = 1.0 / g->dx, idy = 1.0 / g->dy, sum = 0.0;
    int Pu = pitch_u(g), P = pitch_c(g);
#pragma omp parallel for reduction(+ : sum) schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            double divx = u[(size_t)jj * Pu + ii + 1] - u[(size_t)jj * Pu + ii];
            double divy = v[(size_t)(jj + 1) * P + ii] - v[(size_t)jj * P + ii];
            double d = divx * idx + divy * idy;
            sum += d * d;
        }
    double denom = (double)g->nx * g->ny;
    return sqrt(sum / denom);
}

/* max_velocity, metrics.cpp:52-66 */
static double max_speed(const grid_t *g, const double *u, const double *v) {
    double m = 0.0;
    int Pu = pitch_u(g), P = pitch_c(g);
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            double uc = 0.5 * (u[(size_t)jj * Pu + ii] + u[(size_t)jj * Pu + ii + 1]);
            double vc = 0.5 * (v[(size_t)jj * P + ii] + v[(size_t)(jj + 1) * P + ii]);
            m = max2(m, sqrt(uc * uc + vc * vc));
        }
    return m;
}

/* sanitize_field, time_integrator.cpp:8-27 (without the warn-once message) */
static void sanitize(double *a, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i)
        if (!isfinite(a[i])) a[i] = 0.0;
}

/* ------------------------------------------------------------------ pressure solver */

typedef struct {
    grid_t g;
    double *r, *z, *s, *Ap; /* pressure.cpp:8-18, zero-filled incl. ghosts */
    double last_residual;
    int last_iters;
} psolver_t;

static void psolver_init(psolver_t *ps, const grid_t *g) {
    size_t n = count_c(g);
    ps->g = *g;
    ps->r = calloc(n, sizeof(double));
    ps->z = calloc(n, sizeof(double));
    ps->s = calloc(n, sizeof(double));
    ps->Ap = calloc(n, sizeof(double));
    ps->last_residual = 0.0;
    ps->last_iters = 0;
}
static void psolver_free(psolver_t *ps) {
    free(ps->r); free(ps->z); free(ps->s); free(ps->Ap);
}

/* apply_laplacian, pressure.cpp:20-44: 5-point stencil that reads the ghosts of `in` as data */
static void laplacian(const grid_t *g, const double *in, double *out) {
    double idx2 = 1.0 / (g->dx * g->dx), idy2 = 1.0 / (g->dy * g->dy);
    int P = pitch_c(g);
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            const double *q = in + (size_t)jj * P + ii;
            double c = fin0(q[0]), l = fin0(q[-1]), rg = fin0(q[1]), b = fin0(q[-P]), t = fin0(q[P]);
            out[(size_t)jj * P + ii] = (rg - 2.0 * c + l) * idx2 + (t - 2.0 * c + b) * idy2;
        }
}
/* dot, pressure.cpp:46-62 */
static double dotp(const grid_t *g, const double *a, const double *b) {
    double sum = 0.0;
    int P = pitch_c(g);
#pragma omp parallel for reduction(+ : sum) schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) sum += fin0(a[(size_t)jj * P + ii]) * fin0(b[(size_t)jj * P + ii]);
    return sum;
}
/* r = rhs - Ap with finite guards, pressure.cpp:98-109 and :146-157 */
static void residual_field(const grid_t *g, const double *rhs, const double *Ap, double *r) {
    int P = pitch_c(g);
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii)
            r[(size_t)jj * P + ii] = fin0(rhs[(size_t)jj * P + ii]) - fin0(Ap[(size_t)jj * P + ii]);
}

/* smooth, pressure.cpp:64-111: `sweeps` red-black Gauss-Seidel sweeps (colour (i+j)%2 == 0 first, on
 * 0-based interior indices) of  p = (rhs + (pr+pl)*idx2 + (pt+pb)*idy2) / (2*(idx2+idy2))  -- note the sign:
 * this relaxes lap(p) = -rhs while the residual below measures lap(p) = +rhs; that is the reference. */
static double rbgs_smooth(psolver_t *ps, double *p, const double *rhs, int sweeps) {
    const grid_t *g = &ps->g;
    double idx2 = 1.0 / (g->dx * g->dx), idy2 = 1.0 / (g->dy * g->dy);
    double denom = 2.0 * (idx2 + idy2);
    int P = pitch_c(g);
    for (int sw = 0; sw < sweeps; ++sw)
        for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
            for (int jj = 1; jj <= g->ny; ++jj)
                for (int ii = 1 + ((jj - 1 + colour) & 1); ii <= g->nx; ii += 2) {
                    double *q = p + (size_t)jj * P + ii;
                    double rv = fin0(rhs[(size_t)jj * P + ii]);
                    double pl = fin0(q[-1]), pr = fin0(q[1]), pb = fin0(q[-P]), pt = fin0(q[P]);
                    double sum = (pr + pl) * idx2 + (pt + pb) * idy2;
                    q[0] = (rv + sum) / denom;
                }
        }
    laplacian(g, p, ps->Ap);
    residual_field(g, rhs, ps->Ap, ps->r);
    return sqrt(dotp(g, ps->r, ps->r));
}

/* PressureSolver::solve, pressure.cpp:113-242 */
static double psolve(psolver_t *ps, double *p, const double *rhs, const cfdo_params *pp) {
    const grid_t *g = &ps->g;
    int P = pitch_c(g);
    ps->last_iters = 0;
    if (pp->solver == CFDO_SOLVER_MG) { /* :115-129 -- fine-grid RBGS only; mg_levels/rbgs_sweeps unused */
        double res = 1e9;
        for (int vc = 0; vc < pp->vcycles; ++vc) {
            res = rbgs_smooth(ps, p, rhs, 2);
            ps->last_iters = vc + 1;
            if (res < pp->tol) break;
        }
        ps->last_residual = isfinite(res) ? res : 1e9;
        return ps->last_residual;
    }
    /* Jacobi-PCG, cold start (:133-134 clears p including ghosts) */
    memset(p, 0, count_c(g) * sizeof(double));
    p[(size_t)P + 1] = 0.0;
    laplacian(g, p, ps->Ap);
    residual_field(g, rhs, ps->Ap, ps->r);
    double idx2 = 1.0 / (g->dx * g->dx), idy2 = 1.0 / (g->dy * g->dy);
    double diag = 2.0 * (idx2 + idy2);
    double inv_diag = 1.0 / max2(diag, 1e-12);
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            size_t k = (size_t)jj * P + ii;
            ps->z[k] = inv_diag * ps->r[k];
            ps->s[k] = ps->z[k];
        }
    double rho = dotp(g, ps->r, ps->z);
    double res = sqrt(max2(0.0, dotp(g, ps->r, ps->r)));
    int iter;
    for (iter = 0; iter < pp->pcg_max_iters && res > pp->tol; ++iter) { /* :179 */
        laplacian(g, ps->s, ps->Ap);
        double denom = dotp(g, ps->s, ps->Ap);
        if (!isfinite(denom) || fabs(denom) < 1e-20) break; /* "PCG breakdown", :183-186 */
        double alpha = rho / denom;
#pragma omp parallel for schedule(static)
        for (int jj = 1; jj <= g->ny; ++jj)
            for (int ii = 1; ii <= g->nx; ++ii) {
                size_t k = (size_t)jj * P + ii;
                p[k] += alpha * ps->s[k];
                ps->r[k] -= alpha * ps->Ap[k];
            }
        p[(size_t)P + 1] = 0.0; /* re-pin, :200 */
        res = sqrt(max2(0.0, dotp(g, ps->r, ps->r)));
        if (!isfinite(res) || res < pp->tol) { ++iter; break; } /* :203 -- this iteration did its update */
#pragma omp parallel for schedule(static)
        for (int jj = 1; jj <= g->ny; ++jj)
            for (int ii = 1; ii <= g->nx; ++ii) {
                size_t k = (size_t)jj * P + ii;
                ps->z[k] = inv_diag * ps->r[k];
            }
        double rho_new = dotp(g, ps->r, ps->z);
        if (!isfinite(rho_new)) { ++iter; break; }
        double beta = rho_new / rho;
#pragma omp parallel for schedule(static)
        for (int jj = 1; jj <= g->ny; ++jj)
            for (int ii = 1; ii <= g->nx; ++ii) {
                size_t k = (size_t)jj * P + ii;
                ps->s[k] = ps->z[k] + beta * ps->s[k];
            }
        rho = rho_new;
    }
    ps->last_iters = iter; /* number of p/r updates performed */
    ps->last_residual = isfinite(res) ? res : 1e9;
    p[(size_t)P + 1] = 0.0;
    return ps->last_residual;
}

/* ------------------------------------------------------------------ time integrator */

typedef struct {
    grid_t g;
    double *u, *v, *p, *rhs;        /* State, state.hpp:5-9 (tmp/scalar are never touched by the solver) */
    double *du, *dv, *u1, *v1;      /* TimeIntegrator work arrays, time_integrator.cpp:29-36 */
    psolver_t ps;
    double div_before, div_after;
} sim_t;

static void rhs_terms(const grid_t *g, const double *u, const double *v, double *du, double *dv,
                      double invRe) {
    advect(g, u, pitch_u(g), g->nx + 1, g->ny, u, v, du, count_u(g));
    diffuse(g, u, pitch_u(g), g->nx + 1, g->ny, du, invRe);
    advect(g, v, pitch_c(g), g->nx, g->ny + 1, u, v, dv, count_v(g));
    diffuse(g, v, pitch_c(g), g->nx, g->ny + 1, dv, invRe);
}

static double rms_cells(const grid_t *g, const double *a) { /* serial loops, time_integrator.cpp:162-171 */
    double acc = 0.0;
    int P = pitch_c(g);
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            double val = a[(size_t)jj * P + ii];
            acc += val * val;
        }
    return sqrt(acc / (g->nx * g->ny));
}

/* TimeIntegrator::step, time_integrator.cpp:47-239 */
static double sim_step(sim_t *s, const cfdo_bc *bc, double Re, double CFL, const cfdo_params *pp,
                       double dt_override, double *residual) {
    const grid_t *g = &s->g;
    int Pu = pitch_u(g), P = pitch_c(g);
    bc_u(g, s->u, bc); bc_v(g, s->v, bc); bc_p(g, s->p, bc); /* :51-53 */

    double dt_cfl = cfl_dt(g, s->u, s->v, Re, CFL); /* :55-66 */
    double dt = dt_cfl;
    if (dt_override > 0.0) dt = min2(dt_cfl, dt_override);
    double dt_diff = 0.5 * Re * min2(g->dx * g->dx, g->dy * g->dy);
    double dt_cap = (dt_override > 0.0) ? dt_override : 1e-3;
    if (!isfinite(dt) || dt <= 0.0) dt = min2(1e-3, dt_diff);
    dt = clamp3(dt, 1e-8, dt_cap);

    /* RK stage 1, :71-107 */
    rhs_terms(g, s->u, s->v, s->du, s->dv, 1.0 / Re);
    sanitize(s->du, count_u(g)); sanitize(s->dv, count_v(g));
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx + 1; ++ii) {
            size_t k = (size_t)jj * Pu + ii;
            s->u1[k] = s->u[k] + dt * s->du[k];
        }
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny + 1; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            size_t k = (size_t)jj * P + ii;
            s->v1[k] = s->v[k] + dt * s->dv[k];
        }
    sanitize(s->u1, count_u(g)); sanitize(s->v1, count_v(g));
    bc_u(g, s->u1, bc); bc_v(g, s->v1, bc);

    /* RK stage 2 (Heun), :109-155 */
    rhs_terms(g, s->u1, s->v1, s->du, s->dv, 1.0 / Re);
    sanitize(s->du, count_u(g)); sanitize(s->dv, count_v(g));
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx + 1; ++ii) {
            size_t k = (size_t)jj * Pu + ii;
            s->u[k] = 0.5 * (s->u[k] + s->u1[k] + dt * s->du[k]);
        }
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny + 1; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            size_t k = (size_t)jj * P + ii;
            s->v[k] = 0.5 * (s->v[k] + s->v1[k] + dt * s->dv[k]);
        }
    sanitize(s->u, count_u(g)); sanitize(s->v, count_v(g));
    bc_u(g, s->u, bc); bc_v(g, s->v, bc);
    sanitize(s->u, count_u(g)); sanitize(s->v, count_v(g));

    /* projection, :158-207 */
    div_field(g, s->u, s->v, s->rhs);
    sanitize(s->rhs, count_c(g));
    s->div_before = rms_cells(g, s->rhs);
    double inv_dt = 1.0 / dt;
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= g->ny; ++jj)
        for (int ii = 1; ii <= g->nx; ++ii) {
            size_t k = (size_t)jj * P + ii;
            s->rhs[k] = fin0(s->rhs[k] * inv_dt);
        }
    s->p[(size_t)P + 1] = 0.0;
    double res = psolve(&s->ps, s->p, s->rhs, pp);
    s->p[(size_t)P + 1] = 0.0;
    grad_p_correct(g, s->u, s->v, s->p, dt);
    if (bc->left.type == CFDO_INFLOW) /* :194-202 */
        for (int jj = 1; jj <= g->ny; ++jj) {
            double val = jet_profile(bc, ((jj - 1) + 0.5) * g->dy);
            s->u[(size_t)jj * Pu + 1] = val;
            s->u[(size_t)jj * Pu] = val;
        }
    bc_u(g, s->u, bc); bc_v(g, s->v, bc); bc_p(g, s->p, bc);

    /* post-projection divergence stays in s.rhs, unscaled, :210-220 */
    div_field(g, s->u, s->v, s->rhs);
    s->div_after = rms_cells(g, s->rhs);
    sanitize(s->u, count_u(g)); sanitize(s->v, count_v(g)); sanitize(s->p, count_c(g));
    if (residual) *residual = res;
    return dt;
}

/* ------------------------------------------------------------------ cfdo_api.h */

const char *cfdo_kind(void) { return "port"; }

void *cfdo_create(int nx, int ny, double Lx, double Ly) {
    sim_t *s = calloc(1, sizeof(sim_t));
    s->g = grid_make(nx, ny, Lx, Ly);
    const grid_t *g = &s->g;
    s->u = calloc(count_u(g), sizeof(double));
    s->du = calloc(count_u(g), sizeof(double));
    s->u1 = calloc(count_u(g), sizeof(double));
    s->v = calloc(count_v(g), sizeof(double));
    s->dv = calloc(count_v(g), sizeof(double));
    s->v1 = calloc(count_v(g), sizeof(double));
    s->p = calloc(count_c(g), sizeof(double));
    s->rhs = calloc(count_c(g), sizeof(double));
    psolver_init(&s->ps, g);
    return s;
}
void cfdo_destroy(void *sim) {
    sim_t *s = sim;
    if (!s) return;
    free(s->u); free(s->du); free(s->u1); free(s->v); free(s->dv); free(s->v1); free(s->p); free(s->rhs);
    psolver_free(&s->ps);
    free(s);
}
double *cfdo_field(void *sim, int which, int *pitch, int *rows) {
    sim_t *s = sim;
    const grid_t *g = &s->g;
    double *tab[8] = {s->u, s->v, s->p, s->rhs, s->du, s->dv, s->u1, s->v1};
    if (which < 0 || which > 7) return 0;
    int is_u = (which == CFDO_F_U || which == CFDO_F_DU || which == CFDO_F_U1);
    int is_v = (which == CFDO_F_V || which == CFDO_F_DV || which == CFDO_F_V1);
    if (pitch) *pitch = is_u ? pitch_u(g) : pitch_c(g);
    if (rows) *rows = is_v ? g->ny + 3 : g->ny + 2;
    return tab[which];
}
void cfdo_clear_work(void *sim) { /* time_integrator.cpp:38-45 */
    sim_t *s = sim;
    memset(s->du, 0, count_u(&s->g) * sizeof(double));
    memset(s->u1, 0, count_u(&s->g) * sizeof(double));
    memset(s->dv, 0, count_v(&s->g) * sizeof(double));
    memset(s->v1, 0, count_v(&s->g) * sizeof(double));
}
double cfdo_step(void *sim, const cfdo_bc *bc, double Re, double CFL, const cfdo_params *pp,
                 double dt_override, double *residual) {
    return sim_step(sim, bc, Re, CFL, pp, dt_override, residual);
}
int cfdo_last_iters(void *sim) { return ((sim_t *)sim)->ps.last_iters; }
/* port-only extra: the two RMS divergences the reference logs to stderr (time_integrator.cpp:227-228) */
void cfdo_last_divs(void *sim, double *before, double *after) {
    *before = ((sim_t *)sim)->div_before;
    *after = ((sim_t *)sim)->div_after;
}

void cfdo_apply_bc_u(int nx, int ny, double Lx, double Ly, double *u, const cfdo_bc *bc) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    bc_u(&g, u, bc);
}
void cfdo_apply_bc_v(int nx, int ny, double Lx, double Ly, double *v, const cfdo_bc *bc) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    bc_v(&g, v, bc);
}
void cfdo_apply_bc_p(int nx, int ny, double Lx, double Ly, double *p, const cfdo_bc *bc) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    bc_p(&g, p, bc);
}
void cfdo_advect_u(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *du) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    advect(&g, u, pitch_u(&g), nx + 1, ny, u, v, du, count_u(&g));
}
void cfdo_advect_v(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *dv) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    advect(&g, v, pitch_c(&g), nx, ny + 1, u, v, dv, count_v(&g));
}
void cfdo_diffuse_u(int nx, int ny, double Lx, double Ly, const double *u, double *du, double invRe) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    diffuse(&g, u, pitch_u(&g), nx + 1, ny, du, invRe);
}
void cfdo_diffuse_v(int nx, int ny, double Lx, double Ly, const double *v, double *dv, double invRe) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    diffuse(&g, v, pitch_c(&g), nx, ny + 1, dv, invRe);
}
double cfdo_compute_cfl(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double Re,
                        double CFL) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    return cfl_dt(&g, u, v, Re, CFL);
}
void cfdo_divergence(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *rhs) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    div_field(&g, u, v, rhs);
}
void cfdo_subtract_grad_p(int nx, int ny, double Lx, double Ly, double *u, double *v, const double *p,
                          double dt) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    grad_p_correct(&g, u, v, p, dt);
}
double cfdo_pressure_solve(int nx, int ny, double Lx, double Ly, double *p, const double *rhs,
                           const cfdo_params *pp, int *iters) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    psolver_t ps;
    psolver_init(&ps, &g);
    double res = psolve(&ps, p, rhs, pp);
    if (iters) *iters = ps.last_iters;
    psolver_free(&ps);
    return res;
}
double cfdo_divergence_l2(int nx, int ny, double Lx, double Ly, const double *u, const double *v) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    return div_l2(&g, u, v);
}
double cfdo_max_velocity(int nx, int ny, double Lx, double Ly, const double *u, const double *v) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    return max_speed(&g, u, v);
}
void cfdo_sanitize(double *a, long n) { sanitize(a, (size_t)n); }

/* ================================================================== readers / mutators around the step
 * (SURVEY.md §8f rank 4): the visualisation scalar (viz.cpp:6-65), its statistics (gui.cpp:1452-1474), the colour
 * lookup (colormaps.cpp), the texture pixel loop (gui.cpp:1523-1551) and the obstacle masks (shapes.cpp:7-74). */

/* one cell of compute_scalar, viz.cpp:20-55; (ii, jj) raw indices */
static double viz_value(const grid_t *g, const double *u, const double *v, const double *p, int field, int ii, int jj) {
    const int Pu = pitch_u(g), Pc = pitch_c(g);
    double val = 0.0;
    if (field == CFDO_FIELD_U || field == CFDO_FIELD_SPEED) {
        double uc = 0.5 * (u[(size_t)jj * Pu + ii] + u[(size_t)jj * Pu + ii + 1]);
        if (field == CFDO_FIELD_U) return isfinite(uc) ? uc : 0.0;
        double vc = 0.5 * (v[(size_t)jj * Pc + ii] + v[(size_t)(jj + 1) * Pc + ii]);
        val = sqrt(uc * uc + vc * vc);
    } else if (field == CFDO_FIELD_V) {
        val = 0.5 * (v[(size_t)jj * Pc + ii] + v[(size_t)(jj + 1) * Pc + ii]);
    } else if (field == CFDO_FIELD_PRESSURE || field == CFDO_FIELD_TEMPERATURE) {
        val = p[(size_t)jj * Pc + ii]; /* Temperature is a placeholder for pressure, viz.cpp:50-53 */
    } else if (field == CFDO_FIELD_VORTICITY) {
        double dv_dx = (v[(size_t)jj * Pc + ii + 1] - v[(size_t)jj * Pc + ii - 1]) / (2.0 * g->dx);
        double du_dy = (u[(size_t)(jj + 1) * Pu + ii] - u[(size_t)(jj - 1) * Pu + ii]) / (2.0 * g->dy);
        val = dv_dx - du_dy;
    } /* Streamlines: placeholder 0.0, viz.cpp:45-49 */
    return isfinite(val) ? val : 0.0;
}

void cfdo_compute_scalar(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p,
                         int field, double *out, double *vmin, double *vmax) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    double lo = 1e30, hi = -1e30;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            double val = viz_value(&g, u, v, p, field, i + 1, j + 1);
            out[(size_t)j * nx + i] = val;
            lo = min2(lo, val);
            hi = max2(hi, val);
        }
    *vmin = lo;
    *vmax = lo + max2(1e-6, hi - lo); /* viz.cpp:63-64 */
}

void cfdo_field_stats(const double *out, long n, double *mean, double *stddev) {
    double sum = 0.0, sum_sq = 0.0;
    long count = 0;
    for (long k = 0; k < n; ++k)
        if (isfinite(out[k])) {
            sum += out[k];
            sum_sq += out[k] * out[k];
            ++count;
        }
    *mean = 0.0;
    *stddev = 0.0;
    if (count > 0) {
        *mean = sum / (double)count;
        *stddev = sqrt(max2(0.0, sum_sq / (double)count - *mean * *mean));
    }
}

static int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

void cfdo_apply_shapes(int nx, int ny, double Lx, double Ly, double *u, double *v, const cfdo_shape *shapes, int n) {
    grid_t g = grid_make(nx, ny, Lx, Ly);
    const int Pu = pitch_u(&g), Pc = pitch_c(&g);
    for (int k = 0; k < n; ++k) {
        const cfdo_shape *sh = &shapes[k];
        if (!sh->enabled) continue;
        /* bounding box in cells (float -> int truncation), clamped to the grid: shapes.cpp:12-30 */
        float ex = sh->type == 0 ? sh->pos_x + sh->size_x : sh->pos_x + sh->size_x; /* circle: pos + "radius" = size.x */
        float ey = sh->type == 0 ? sh->pos_y + sh->size_y : sh->pos_y + sh->size_x;
        int i0 = clampi((int)sh->pos_x, 0, nx - 1), j0 = clampi((int)sh->pos_y, 0, ny - 1);
        int i1 = clampi((int)ex, 0, nx - 1), j1 = clampi((int)ey, 0, ny - 1);
        float cx = sh->pos_x + sh->size_x / 2.0f, cy = sh->pos_y + sh->size_y / 2.0f, rad = sh->size_x / 2.0f;
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) {
                if (sh->type != 0) { /* circle: cell centre within the radius, all in float (shapes.cpp:41-47) */
                    float dx = ((float)i + 0.5f) - cx, dy = ((float)j + 0.5f) - cy;
                    if (!(dx * dx + dy * dy <= rad * rad)) continue;
                }
                double *pu = &u[(size_t)(j + 1) * Pu + i + 1], *pv = &v[(size_t)(j + 1) * Pc + i + 1];
                if (sh->bc_type == 0) { /* solid: the cell's left and bottom face */
                    *pu = 0.0;
                    *pv = 0.0;
                } else { /* porous: 90 % reduction */
                    *pu *= 0.1;
                    *pv *= 0.1;
                }
            }
    }
}

/* colormaps.cpp: degree-6 Horner fits (viridis, plasma), piecewise ramps (jet, rdylbu), quartic in power form (turbo) */
static const float CM_HORNER[2][7][3] = {
    {{0.2777273272234177f, 0.005407344544966578f, 0.3340998053353061f},
     {0.1050930431085774f, 1.404613529898575f, 1.384590162594685f},
     {-0.3308618287255563f, 0.214847559468213f, 0.09509516302823659f},
     {-4.634230498983486f, -5.799100973351585f, -19.33244095627987f},
     {6.228269936347081f, 14.17993336680509f, 56.69055260068105f},
     {4.776384997670288f, -13.74514537774601f, -65.35303263337234f},
     {-5.435455855934631f, 12.86216349749025f, 26.312873249486f}},
    {{0.05873234392399702f, 0.02333670892565664f, 0.5433401826748754f},
     {2.176514634195958f, 0.2383834171260182f, 0.7539604599784036f},
     {-2.689460476458034f, -7.455851135738909f, 3.110799939717086f},
     {6.130348345893603f, 42.3461881477227f, -28.51885465332158f},
     {-11.10743619062271f, -82.66631104428044f, 60.13984767418263f},
     {10.02306557647065f, 71.41361770095349f, -54.07218655560067f},
     {-3.658713842777788f, -22.93153465461149f, 18.19190778539828f}}};
static const float CM_TURBO[5][3] = {{0.18995f, 0.07176f, 0.23217f}, {0.30315f, 0.50427f, 0.94791f},
                                     {0.38343f, 0.87344f, 0.54467f}, {0.98853f, 0.64411f, 0.09395f},
                                     {0.95839f, 0.82344f, 0.08425f}};
static float clampf01(float x) { return (x < 0.0f) ? 0.0f : ((1.0f < x) ? 1.0f : x); }

void cfdo_apply_colormap(int colormap, float t, unsigned char *rgb) {
    float c[3];
    if (colormap == 2) { /* jet, colormaps.cpp:45-71 */
        if (t < 0.125f) { c[0] = 0.0f; c[1] = 0.0f; c[2] = 0.5f + 4.0f * t; }
        else if (t < 0.375f) { c[0] = 0.0f; c[1] = 4.0f * (t - 0.125f); c[2] = 1.0f; }
        else if (t < 0.625f) { c[0] = 4.0f * (t - 0.375f); c[1] = 1.0f; c[2] = 1.0f - 4.0f * (t - 0.375f); }
        else if (t < 0.875f) { c[0] = 1.0f; c[1] = 1.0f - 4.0f * (t - 0.625f); c[2] = 0.0f; }
        else { c[0] = 1.0f - 4.0f * (t - 0.875f); c[1] = 0.0f; c[2] = 0.0f; }
    } else if (colormap == 3) { /* grayscale */
        c[0] = c[1] = c[2] = t;
    } else if (colormap == 4) { /* turbo, :77-96 */
        float t2 = t * t, t3 = t2 * t, t4 = t3 * t;
        for (int k = 0; k < 3; ++k)
            c[k] = clampf01(CM_TURBO[0][k] + CM_TURBO[1][k] * t + CM_TURBO[2][k] * t2 + CM_TURBO[3][k] * t3 + CM_TURBO[4][k] * t4);
    } else if (colormap == 5) { /* rdylbu, :98-112 */
        if (t < 0.5f) { c[0] = 1.0f; c[1] = 2.0f * t; c[2] = 0.0f; }
        else { float t2 = 2.0f * (t - 0.5f); c[0] = 1.0f - t2; c[1] = 1.0f - t2; c[2] = t2; }
    } else { /* viridis (0 and every unknown id), plasma (1) */
        const float(*q)[3] = CM_HORNER[colormap == 1 ? 1 : 0];
        for (int k = 0; k < 3; ++k) {
            float acc = q[6][k];
            for (int d = 5; d >= 1; --d) acc = q[d][k] + t * acc;
            c[k] = clampf01(q[0][k] + t * acc);
        }
    }
    for (int k = 0; k < 3; ++k) rgb[k] = (unsigned char)(c[k] * 255.0f);
}

void cfdo_render_rgb(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p, int field,
                     int colormap, float scale_min, float scale_max, int normalize, float gamma, unsigned char *rgb) {
    size_t n = (size_t)nx * ny;
    double *tmp = (double *)malloc(n * sizeof(double));
    double vmin, vmax;
    cfdo_compute_scalar(nx, ny, Lx, Ly, u, v, p, field, tmp, &vmin, &vmax);
    if (scale_min != 0.0f || scale_max != 1.0f) {
        vmin = scale_min;
        vmax = scale_max;
    }
    double scale = 1.0 / max2(1e-6, vmax - vmin);
    for (size_t k = 0; k < n; ++k) {
        double val = normalize ? fabs(tmp[k]) : tmp[k];
        double norm = clamp3((val - vmin) * scale, 0.0, 1.0);
        norm = pow(norm, 1.0 / gamma);
        cfdo_apply_colormap(colormap, (float)norm, rgb + 3 * k);
    }
    free(tmp);
}
