/*
 * cfdo_api.h -- the C interface shared by BOTH CPU checkers of this repo:
 *
 *   oracle/_ref/libcfdo_ref.so   the reference's own src/solver/ *.cpp, compiled unmodified from
 *                                /root/reference by oracle/Makefile and wrapped by oracle/ref_shim.cpp
 *   oracle/_build/libcfdo_port.so  oracle/cfd_oracle.c, a plain-C restatement of the same algorithm
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these libraries,
 * and only as the checker / the CPU baseline.  The CUDA path never calls into them.
 *
 * All arrays use the REFERENCE host layout (ng = 1, row-major, ghosts included):
 *   u   : pitch nx+3, rows ny+2      (grid.hpp:15-18, main.cpp:107)
 *   v   : pitch nx+2, rows ny+3      (grid.hpp:20-23, main.cpp:108)
 *   p/rhs/r/s/Ap : pitch nx+2, rows ny+2   (grid.hpp:11-13, main.cpp:109-110)
 */
#ifndef CFDO_API_H
#define CFDO_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* BCType order of bc.hpp:6 */
enum { CFDO_WALL = 0, CFDO_MOVING = 1, CFDO_INFLOW = 2, CFDO_OUTFLOW = 3, CFDO_PERIODIC = 4 };
/* PressureSolverType order of pressure.hpp:6 */
enum { CFDO_SOLVER_MG = 0, CFDO_SOLVER_PCG = 1 };

typedef struct cfdo_bc_side {
    int type;
    double moving, inflow_u, inflow_v; /* bc.hpp:8-13 */
} cfdo_bc_side;

typedef struct cfdo_bc {
    cfdo_bc_side left, right, bottom, top; /* bc.hpp:15-19 */
    double jet_center, jet_width;          /* bc.hpp:21-22; the other jet_* members are dead (bc.cpp:21-26) */
} cfdo_bc;

typedef struct cfdo_params {
    int solver;        /* CFDO_SOLVER_* */
    int mg_levels;     /* parsed, never read by the reference (pressure.cpp:113-129) */
    int vcycles;
    int rbgs_sweeps;   /* never read by the reference: the sweep count is the literal 2 (pressure.cpp:119) */
    double tol;
    int pcg_max_iters;
} cfdo_params;

enum { CFDO_F_U = 0, CFDO_F_V, CFDO_F_P, CFDO_F_RHS, CFDO_F_DU, CFDO_F_DV, CFDO_F_U1, CFDO_F_V1 };

const char *cfdo_kind(void); /* "reference" or "port" */

/* A whole simulation: Grid + State + PressureSolver + TimeIntegrator (main.cpp:104-130). */
void *cfdo_create(int nx, int ny, double Lx, double Ly);
void cfdo_destroy(void *sim);
double *cfdo_field(void *sim, int which, int *pitch, int *rows);
void cfdo_clear_work(void *sim);
/* TimeIntegrator::step (time_integrator.cpp:47-239); returns dt. */
double cfdo_step(void *sim, const cfdo_bc *bc, double Re, double CFL, const cfdo_params *pp,
                 double dt_override, double *residual);
/* PCG iterations of the last solve; -1 when the implementation cannot know (the unmodified reference). */
int cfdo_last_iters(void *sim);

/* Stage-level entry points on caller-owned arrays. */
void cfdo_apply_bc_u(int nx, int ny, double Lx, double Ly, double *u, const cfdo_bc *bc);
void cfdo_apply_bc_v(int nx, int ny, double Lx, double Ly, double *v, const cfdo_bc *bc);
void cfdo_apply_bc_p(int nx, int ny, double Lx, double Ly, double *p, const cfdo_bc *bc);
void cfdo_advect_u(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *du);
void cfdo_advect_v(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *dv);
void cfdo_diffuse_u(int nx, int ny, double Lx, double Ly, const double *u, double *du, double invRe);
void cfdo_diffuse_v(int nx, int ny, double Lx, double Ly, const double *v, double *dv, double invRe);
double cfdo_compute_cfl(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double Re,
                        double CFL);
void cfdo_divergence(int nx, int ny, double Lx, double Ly, const double *u, const double *v, double *rhs);
void cfdo_subtract_grad_p(int nx, int ny, double Lx, double Ly, double *u, double *v, const double *p,
                          double dt);
double cfdo_pressure_solve(int nx, int ny, double Lx, double Ly, double *p, const double *rhs,
                           const cfdo_params *pp, int *iters);
double cfdo_divergence_l2(int nx, int ny, double Lx, double Ly, const double *u, const double *v);
double cfdo_max_velocity(int nx, int ny, double Lx, double Ly, const double *u, const double *v);
void cfdo_sanitize(double *a, long n);

/* ---- host-side readers / mutators around the step ---------------------------------- */
/* ScalarField order of viz.hpp:8 */
enum { CFDO_FIELD_U = 0, CFDO_FIELD_V, CFDO_FIELD_SPEED, CFDO_FIELD_PRESSURE, CFDO_FIELD_VORTICITY,
       CFDO_FIELD_STREAMLINES, CFDO_FIELD_TEMPERATURE };
/* compute_scalar (viz.cpp:6-65): out = nx*ny doubles (index i + j*nx); vmin / vmax after the span rule. */
void cfdo_compute_scalar(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p,
                         int field, double *out, double *vmin, double *vmax);
/* mean / standard deviation of Gui::update_field_statistics (gui.cpp:1433-1481) over `out` */
void cfdo_field_stats(const double *out, long n, double *mean, double *stddev);
/* Shape of gui.hpp:14-23 (pos / size in cell units, float) */
typedef struct cfdo_shape {
    int type;     /* 0 = rectangle, 1 = circle */
    float pos_x, pos_y, size_x, size_y;
    int bc_type;  /* 0 = solid wall, 1 = porous */
    int enabled;
} cfdo_shape;
/* apply_shape_boundary_conditions (shapes.cpp:7-74) on u, v */
void cfdo_apply_shapes(int nx, int ny, double Lx, double Ly, double *u, double *v, const cfdo_shape *shapes, int n);
/* Colormaps::apply_colormap (colormaps.cpp:114-127) */
void cfdo_apply_colormap(int colormap, float t, unsigned char *rgb);
/* The pixel loop of update_field_texture (gui.cpp:1505-1561) without the GL upload: rgb = nx*ny*3 bytes. */
void cfdo_render_rgb(int nx, int ny, double Lx, double Ly, const double *u, const double *v, const double *p, int field,
                     int colormap, float scale_min, float scale_max, int normalize, float gamma, unsigned char *rgb);

#ifdef __cplusplus
}
#endif
#endif
