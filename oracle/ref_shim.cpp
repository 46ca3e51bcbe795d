// ref_shim.cpp -- extern "C" wrapper (cfdo_api.h) around the UNMODIFIED reference solver.
//
// TEST INFRASTRUCTURE ONLY (see cfdo_api.h).  This file contains no solver arithmetic of its own: it
// includes the reference's public headers from /root/reference/src (never copied into this repo) and
// forwards to the reference's functions (two exceptions, marked RESTATED below: loops that live in the unbuildable gui.cpp).  oracle/Makefile compiles it together with the reference's
// src/solver/{grid,bc,advection,diffusion,metrics,pressure,project,time_integrator}.cpp into
// oracle/_ref/libcfdo_ref.so.  It mirrors the object set-up of the headless driver, main.cpp:104-130.
#include "cfdo_api.h"

#include <cstdio>
#include <cstring>

#include <algorithm>
#include <cmath>
#include <vector>

#include "colormaps.hpp"
#include "gui/gui.hpp" // struct Shape (type-only stubs for the GUI toolkit headers: oracle/stubs/)
#include "shapes.hpp"
#include "solver/time_integrator.hpp"
#include "viz.hpp"

namespace {

struct Sim {
    Grid grid;
    State state;
    PressureSolver *pressure = nullptr;
    TimeIntegrator *integ = nullptr;
};

BC to_bc(const cfdo_bc *b) {
    BC bc;
    auto side = [](const cfdo_bc_side &s) {
        BCSide o;
        o.type = static_cast<BCType>(s.type);
        o.moving = s.moving;
        o.inflow_u = s.inflow_u;
        o.inflow_v = s.inflow_v;
        return o;
    };
    bc.left = side(b->left);
    bc.right = side(b->right);
    bc.bottom = side(b->bottom);
    bc.top = side(b->top);
    bc.jet_center = b->jet_center;
    bc.jet_width = b->jet_width;
    return bc;
}

PressureParams to_pp(const cfdo_params *p) {
    PressureParams pp;
    pp.type = p->solver == CFDO_SOLVER_PCG ? PressureSolverType::PCG : PressureSolverType::Multigrid;
    pp.mg_levels = p->mg_levels;
    pp.vcycles = p->vcycles;
    pp.rbgs_sweeps = p->rbgs_sweeps;
    pp.tol = p->tol;
    pp.pcg_max_iters = p->pcg_max_iters;
    return pp;
}

Grid make_grid(int nx, int ny, double Lx, double Ly) {
    Grid g;
    g.init(nx, ny, Lx, Ly, 1);
    return g;
}

// A Field2D view over caller memory; released (not freed) on scope exit.
struct View {
    Field2D<double> f;
    View(const double *data, int NX, int NY, int pitch) {
        f.nx = NX;
        f.ny = NY;
        f.pitch = pitch;
        f.ngx = 1;
        f.ngy = 1;
        f.data = const_cast<double *>(data);
    }
    ~View() { f.data = nullptr; }
};
struct UView : View { UView(const Grid &g, const double *d) : View(d, g.u_nx(), g.ny, g.u_pitch()) {} };
struct VView : View { VView(const Grid &g, const double *d) : View(d, g.nx, g.v_ny(), g.v_pitch()) {} };
struct PView : View { PView(const Grid &g, const double *d) : View(d, g.nx, g.ny, g.p_pitch()) {} };

// A State whose u, v, p are views over caller memory (released, not freed, on scope exit).
struct StateView {
    State s;
    static void view(Field2D<double> &f, const double *data, int NX, int NY, int pitch) {
        f.nx = NX, f.ny = NY, f.pitch = pitch, f.ngx = 1, f.ngy = 1;
        f.data = const_cast<double *>(data);
    }
    StateView(const Grid &g, const double *u, const double *v, const double *p) {
        view(s.u, u, g.u_nx(), g.ny, g.u_pitch());
        view(s.v, v, g.nx, g.v_ny(), g.v_pitch());
        if (p) view(s.p, p, g.nx, g.ny, g.p_pitch());
    }
    ~StateView() { s.u.data = s.v.data = s.p.data = nullptr; }
};

} // namespace

extern "C" {

const char *cfdo_kind(void) { return "reference"; }

void *cfdo_create(int nx, int ny, double Lx, double Ly) {
    Sim *s = new Sim;
    s->grid.init(nx, ny, Lx, Ly, 1);
    const Grid &g = s->grid;
    s->state.u.allocate(g.u_nx(), g.ny, g.u_pitch(), g.ngx, g.ngy);
    s->state.v.allocate(g.nx, g.v_ny(), g.v_pitch(), g.ngx, g.ngy);
    s->state.p.allocate(g.nx, g.ny, g.p_pitch(), g.ngx, g.ngy);
    s->state.rhs.allocate(g.nx, g.ny, g.p_pitch(), g.ngx, g.ngy);
    s->state.tmp.allocate(g.nx, g.ny, g.p_pitch(), g.ngx, g.ngy);
    s->pressure = new PressureSolver(s->grid);
    s->integ = new TimeIntegrator(s->grid);
    s->integ->clear_work();
    return s;
}

void cfdo_destroy(void *sim) {
    Sim *s = static_cast<Sim *>(sim);
    if (!s) return;
    delete s->integ;
    delete s->pressure;
    delete s;
}

double *cfdo_field(void *sim, int which, int *pitch, int *rows) {
    Sim *s = static_cast<Sim *>(sim);
    Field2D<double> *f = nullptr;
    switch (which) {
    case CFDO_F_U: f = &s->state.u; break;
    case CFDO_F_V: f = &s->state.v; break;
    case CFDO_F_P: f = &s->state.p; break;
    case CFDO_F_RHS: f = &s->state.rhs; break;
    case CFDO_F_DU: f = &s->integ->du_dt; break;
    case CFDO_F_DV: f = &s->integ->dv_dt; break;
    case CFDO_F_U1: f = &s->integ->u1; break;
    case CFDO_F_V1: f = &s->integ->v1; break;
    default: return nullptr;
    }
    if (pitch) *pitch = f->pitch;
    if (rows) *rows = f->ny + 2 * f->ngy;
    return f->data;
}

void cfdo_clear_work(void *sim) { static_cast<Sim *>(sim)->integ->clear_work(); }

double cfdo_step(void *sim, const cfdo_bc *bc, double Re, double CFL, const cfdo_params *pp,
                 double dt_override, double *residual) {
    Sim *s = static_cast<Sim *>(sim);
    BC b = to_bc(bc);
    PressureParams p = to_pp(pp);
    double res = 0.0;
    double dt = s->integ->step(s->state, b, Re, CFL, *s->pressure, p, res, dt_override);
    if (residual) *residual = res;
    return dt;
}

int cfdo_last_iters(void *) { return -1; }

void cfdo_apply_bc_u(int nx, int ny, double Lx, double Ly, double *u, const cfdo_bc *bc) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    apply_bc_u(g, U.f, to_bc(bc));
}
void cfdo_apply_bc_v(int nx, int ny, double Lx, double Ly, double *v, const cfdo_bc *bc) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    VView V(g, v);
    apply_bc_v(g, V.f, to_bc(bc));
}
void cfdo_apply_bc_p(int nx, int ny, double Lx, double Ly, double *p, const cfdo_bc *bc) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    PView P(g, p);
    apply_bc_p(g, P.f, to_bc(bc));
}
void cfdo_advect_u(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *du) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u), D(g, du);
    VView V(g, v);
    advect_u(g, U.f, V.f, D.f);
}
void cfdo_advect_v(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *dv) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v), D(g, dv);
    advect_v(g, U.f, V.f, D.f);
}
void cfdo_diffuse_u(int nx, int ny, double Lx, double Ly, const double *u, double *du, double invRe) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u), D(g, du);
    diffuse_u(g, U.f, D.f, invRe);
}
void cfdo_diffuse_v(int nx, int ny, double Lx, double Ly, const double *v, double *dv, double invRe) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    VView V(g, v), D(g, dv);
    diffuse_v(g, V.f, D.f, invRe);
}
double cfdo_compute_cfl(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double Re,
                        double CFL) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v);
    return compute_cfl(g, U.f, V.f, Re, CFL);
}
void cfdo_divergence(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *rhs) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v);
    PView R(g, rhs);
    divergence(g, U.f, V.f, R.f);
}
void cfdo_subtract_grad_p(int nx, int ny, double Lx, double Ly, double *u, double *v, const double *p,
                          double dt) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v);
    PView P(g, p);
    subtract_grad_p(g, U.f, V.f, P.f, dt);
}
double cfdo_pressure_solve(int nx, int ny, double Lx, double Ly, double *p, const double *rhs,
                           const cfdo_params *pp, int *iters) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    PView P(g, p), R(g, rhs);
    PressureSolver solver(g);
    if (iters) *iters = -1;
    return solver.solve(P.f, R.f, to_pp(pp));
}
double cfdo_divergence_l2(int nx, int ny, double Lx, double Ly, const double *u, const double *v) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v);
    return divergence_l2(g, U.f, V.f);
}
double cfdo_max_velocity(int nx, int ny, double Lx, double Ly, const double *u, const double *v) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    UView U(g, u);
    VView V(g, v);
    return max_velocity(g, U.f, V.f);
}
void cfdo_sanitize(double *a, long n) {
    // sanitize_field works on a whole Field2D allocation; present the buffer as one row.
    Field2D<double> f;
    f.nx = static_cast<int>(n);
    f.ny = 1;
    f.pitch = static_cast<int>(n);
    f.ngx = 0;
    f.ngy = 0;
    f.data = a;
    sanitize_field(f);
    f.data = nullptr;
}

// ---- readers / mutators around the step -------------------------------------------------
void cfdo_compute_scalar(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p,
                         int field, double *out, double *vmin, double *vmax) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    StateView S(g, u, v, p);
    std::vector<double> tmp;
    compute_scalar(g, S.s, static_cast<ScalarField>(field), tmp, *vmin, *vmax); // viz.cpp:6-65
    std::copy(tmp.begin(), tmp.end(), out);
}
// RESTATED from Gui::update_field_statistics, gui.cpp:1452-1474 (gui.cpp needs the GUI toolkit and does not build here)
void cfdo_field_stats(const double *out, long n, double *mean, double *stddev) {
    double sum = 0.0, sum_sq = 0.0;
    int count = 0;
    for (long k = 0; k < n; ++k) {
        const double val = out[k];
        if (std::isfinite(val)) {
            sum += val;
            sum_sq += val * val;
            count++;
        }
    }
    if (count > 0) {
        *mean = sum / count;
        double variance = (sum_sq / count) - (*mean * *mean);
        *stddev = std::sqrt(std::max(0.0, variance));
    } else {
        *mean = 0.0;
        *stddev = 0.0;
    }
}
void cfdo_apply_shapes(int nx, int ny, double Lx, double Ly, double *u, double *v, const cfdo_shape *shapes, int n) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    StateView S(g, u, v, nullptr);
    std::vector<Shape> list;
    for (int k = 0; k < n; ++k) {
        Shape sh;
        sh.type = shapes[k].type;
        sh.pos = ImVec2(shapes[k].pos_x, shapes[k].pos_y);
        sh.size = ImVec2(shapes[k].size_x, shapes[k].size_y);
        sh.bcType = shapes[k].bc_type;
        sh.enabled = shapes[k].enabled != 0;
        list.push_back(sh);
    }
    apply_shape_boundary_conditions(g, S.s, list); // shapes.cpp:7-74
}
void cfdo_apply_colormap(int colormap, float t, unsigned char *rgb) {
    Colormaps::apply_colormap(colormap, t, rgb[0], rgb[1], rgb[2]); // colormaps.cpp:114-127
}
// The pixel loop is RESTATED from update_field_texture, gui.cpp:1523-1551 (gui.cpp does not build here); the scalar field
// and the colour come from the reference's own compute_scalar and Colormaps::apply_colormap.
void cfdo_render_rgb(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p, int field,
                     int colormap, float scale_min, float scale_max, int normalize, float gamma, unsigned char *rgb) {
    Grid g = make_grid(nx, ny, Lx, Ly);
    StateView S(g, u, v, p);
    std::vector<double> tmp;
    double vmin = 0.0, vmax = 0.0;
    compute_scalar(g, S.s, static_cast<ScalarField>(field), tmp, vmin, vmax);
    if (scale_min != 0.0f || scale_max != 1.0f) {
        vmin = scale_min;
        vmax = scale_max;
    }
    double scale = 1.0 / std::max(1e-6, vmax - vmin);
    for (int idx = 0; idx < nx * ny; ++idx) {
        double val = tmp[idx];
        if (normalize) val = std::abs(val);
        double norm = (val - vmin) * scale;
        norm = std::clamp(norm, 0.0, 1.0);
        norm = std::pow(norm, 1.0 / gamma);
        Colormaps::apply_colormap(colormap, static_cast<float>(norm), rgb[3 * idx], rgb[3 * idx + 1], rgb[3 * idx + 2]);
    }
}

} // extern "C"
