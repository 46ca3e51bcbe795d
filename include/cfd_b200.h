/*
 * cfd_b200.h -- C ABI of the B200-native per-timestep solver path (libcfd_b200.so).
 *
 * This is the drop-in boundary for the hot path of archhie/MacOS-CFD's 2D incompressible MAC-grid solver
 * (src/solver/): everything TimeIntegrator::step does (time_integrator.cpp:47-239) runs as hand-written
 * sm_100a CUDA kernels behind these entry points.  Plain pointers and sizes only -- no C++ or torch types.
 * The reference-shaped C++ classes in macos-cfd_b200/cpp/ (Grid, Field2D, State, BC, PressureSolver,
 * TimeIntegrator with the reference's signatures) forward to this ABI; INTEGRATION.md shows the binding.
 *
 * Host arrays always use the REFERENCE layout (ng = 1, row-major, ghosts included):
 *   u            : pitch nx+3, rows ny+2      grid.hpp:15-18 / main.cpp:107
 *   v            : pitch nx+2, rows ny+3      grid.hpp:20-23 / main.cpp:108
 *   p, rhs       : pitch nx+2, rows ny+2      grid.hpp:11-13 / main.cpp:109-110
 * The device layout (padded, 128-byte aligned rows) is private to the library.
 *
 * Every function returns 0 on success or a negative CFD_B200_E* code; cfd_b200_last_error() then holds a
 * message.  No exceptions cross this boundary.  There is NO CPU fallback: without a CUDA device
 * cfd_b200_create fails with CFD_B200_ENODEVICE.  One host thread drives a handle (like the reference,
 * SURVEY.md §8b "Threading").
 */
#ifndef CFD_B200_H
#define CFD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFD_B200_ABI_VERSION 1

enum {
    CFD_B200_OK = 0,
    CFD_B200_EINVAL = -1,    /* bad argument */
    CFD_B200_ENODEVICE = -2, /* no usable CUDA device (the product path has no CPU fallback) */
    CFD_B200_ECUDA = -3,     /* CUDA runtime error */
    CFD_B200_ENOMEM = -4,    /* device/host allocation failed (reference: std::bad_alloc, field.hpp:46,52) */
    CFD_B200_ECOMM = -5,     /* NCCL / multi-GPU error */
    CFD_B200_EUNSUPPORTED = -6
};

/* BCType, bc.hpp:6 -- same order */
enum { CFD_B200_WALL = 0, CFD_B200_MOVING = 1, CFD_B200_INFLOW = 2, CFD_B200_OUTFLOW = 3, CFD_B200_PERIODIC = 4 };
/* PressureSolverType, pressure.hpp:6 -- same order */
enum { CFD_B200_SOLVER_MG = 0, CFD_B200_SOLVER_PCG = 1 };
/* How --solver=mg is executed.
 *   REFERENCE : what the reference really does (pressure.cpp:113-129): `vcycles` x (2 red-black Gauss-Seidel
 *               sweeps on the fine grid + residual); mg_levels / rbgs_sweeps ignored; sign quirk kept.  Parity mode.
 *   VCYCLE    : a true geometric V-cycle honouring mg_levels / rbgs_sweeps.  Extension named by the project
 *               goal; the reference has no such code, so its parity is UNPINNED (own CPU restatement only). */
enum { CFD_B200_MG_REFERENCE = 0, CFD_B200_MG_VCYCLE = 1 };

/* BCSide, bc.hpp:8-13 */
typedef struct cfd_b200_bc_side {
    int type;
    double moving;   /* tangential speed for Moving */
    double inflow_u; /* imposed inflow velocities */
    double inflow_v;
} cfd_b200_bc_side;

/* BC, bc.hpp:15-27.  Only jet_center / jet_width / left.inflow_u shape the jet: the reference's
 * jet_thickness / jet_eps / jet_k / jet_phase are dead (perturbation commented out, bc.cpp:21-26). */
typedef struct cfd_b200_bc {
    cfd_b200_bc_side left, right, bottom, top;
    double jet_center;
    double jet_width;
} cfd_b200_bc;

/* PressureParams, pressure.hpp:8-15 (+ mg_mode, see above; 0 = reference behaviour) */
typedef struct cfd_b200_params {
    int solver;        /* CFD_B200_SOLVER_* */
    int mg_levels;
    int vcycles;
    int rbgs_sweeps;
    double tol;
    int pcg_max_iters;
    int mg_mode;       /* CFD_B200_MG_* */
} cfd_b200_params;

/* What one step reports (everything TimeIntegrator::step returns, writes or logs). */
typedef struct cfd_b200_step_info {
    double dt;            /* return value of TimeIntegrator::step */
    double residual;      /* pressure_residual out-parameter (1e9 when non-finite, pressure.cpp:123,232) */
    int iters;            /* PCG: number of p/r updates; mg: smooth() calls performed */
    int pcg_breakdown;    /* iteration at which "PCG breakdown" hit (pressure.cpp:183-186), -1 if none */
    double div_before;    /* RMS divergence before projection (stderr log, time_integrator.cpp:162-171) */
    double div_after;     /* ... and after (:211-220) */
    int nonfinite;        /* bit 0: a non-finite value was replaced by 0 (sanitize_field warning, :20-25);
                             bit 1: non-finite pressure residual (pressure.cpp:124-127, :233-236) */
} cfd_b200_step_info;

typedef struct cfd_b200_solver cfd_b200_solver; /* opaque: Grid + device State + PressureSolver + TimeIntegrator */

int cfd_b200_abi_version(void);
/* Host memory for the caller-owned State: pinned when CUDA is reachable (fast copies), 64-byte aligned otherwise
 * (Field2D::allocate, field.hpp:35-55).  NULL on failure. */
void *cfd_b200_host_alloc(size_t bytes);
void cfd_b200_host_free(void *p);
const char *cfd_b200_last_error(void);
/* Number of CUDA devices visible, or a negative error code. */
int cfd_b200_device_count(void);

/* Grid::init (grid.cpp:3-12) + State allocation (main.cpp:106-112) + PressureSolver / TimeIntegrator
 * constructors (pressure.cpp:8-18, time_integrator.cpp:29-36), all zero-filled, on CUDA device `device`. */
int cfd_b200_create(int nx, int ny, double Lx, double Ly, int device, cfd_b200_solver **out);

/* Row-slab member of a multi-GPU run (one process per GPU).  The global grid is nx x ny; this rank owns cell rows
 * [ny*rank/nranks, ny*(rank+1)/nranks).  nccl_id is the 128-byte ncclUniqueId produced by
 * cfd_b200_comm_id() on rank 0 and distributed by the caller (torch.distributed, MPI, a file ...).
 * upload/download then take/return only this rank's rows (see cfd_b200_slab_rows). */
int cfd_b200_comm_id(void *id128);
int cfd_b200_create_slab(int nx, int ny, double Lx, double Ly, int device, int rank, int nranks,
                         const void *nccl_id128, cfd_b200_solver **out);
/* How this handle talks to the other slabs: 0 = single GPU, 1 = NCCL send/recv + allreduce, 2 = peer memory over
 * NVLink (CUDA IPC: halo rows stored straight into the neighbour's arrays, reductions fused with their scalar tails).
 * 2 is chosen when every rank could map every peer; CFD_B200_P2P=0 in the environment forces 1. */
int cfd_b200_comm_mode(const cfd_b200_solver *s);
/* Row slabs: every wait for a peer GPU is bounded (CFD_B200_P2P_TIMEOUT_S seconds, default 20).  A slab whose peer does not
 * arrive in time -- or whose peer reports that one of its own waits ran out -- finishes the step with CFD_B200_ECOMM from
 * cfd_b200_sync / cfd_b200_step; the State is then invalid on every slab.  cfd_b200_comm_reset (collective: every slab
 * must call it) clears that condition and restarts the flag protocol; upload or reset the State afterwards. */
int cfd_b200_comm_reset(cfd_b200_solver *s);
/* Test hook: keeps this handle's stream busy for `microseconds`, to make one slab late on purpose. */
int cfd_b200_debug_delay(cfd_b200_solver *s, double microseconds);
/* Measurement hook: enqueue `reps` slab-communication operations of one kind back to back (0 = reduction fused with its
 * tail, 1 = one-way send of 2 rows of u, v + wait, 2 = the handshake exchange of the same rows, 3 = reduction with a row of
 * p riding on it); the caller times them with CUDA events on cfd_b200_stream(). */
int cfd_b200_debug_comm(cfd_b200_solver *s, int kind, int reps);
/* Measurement: microseconds spent polling for a peer GPU (us4[k]) and number of polls that found their data not yet there
 * (count4[k]) since creation / the last call with reset != 0; k = 0 stand-alone wait kernels, 1 edge blocks of consumer
 * kernels, 2 reductions, 3 handshake exchanges.  Edge blocks poll in parallel, so their sum is not elapsed time. */
int cfd_b200_wait_stats(cfd_b200_solver *s, double *us4, long long *count4, int reset);
/* Tuning / test hook: rows > 0 forces the streaming smoother with `rows` rows per strip on any grid size (the default,
 * 0, picks the strip height from the grid and uses the tile smoother on grids too small to fill the GPU). */
int cfd_b200_set_smoother_strip(cfd_b200_solver *s, int rows);
/* Host-only (needs no device): the strip height the streaming smoother would use for a single-GPU nx x ny grid on a GPU
 * with sm_count SMs (257, 129, 65 or 33 rows), or 0 when the grid is too small for it and the tile smoother runs. */
int cfd_b200_smoother_strip_for(int nx, int ny, int sm_count);
/* Which kernel ran the last smooth() of --solver=mg (reference mode): 1 = streaming smoother (large grids), 2 = tile
 * smoother (small grids, and the guarded slow path of 1), 3 = one launch per colour pass (variant 0), 0 = none yet. */
int cfd_b200_last_smoother(const cfd_b200_solver *s);
/* Which kernels ran the last PCG solve: 1 = k_pcg_coop (the whole solve in one cooperative persistent launch; grids whose
 * p, r, s fit into the SMs' shared memory), 2 = the streaming pair k_pcg_phase1_march / _phase2_march, 3 = the tile pair
 * k_pcg_phase1 / _phase2, 4 = k_pcg_cluster (grids of at most 512 x 256 cells: the whole solve inside one 16-CTA thread-block
 * cluster, hardware cluster barriers, sums and halo rows through distributed shared memory; k_pcg_coop behind it as its
 * guarded slow path), 0 = none yet. */
int cfd_b200_last_pcg(const cfd_b200_solver *s);
/* First owned cell row and number of owned cell rows of this handle (0, ny for a single-GPU handle). */
int cfd_b200_slab_rows(const cfd_b200_solver *s, int *row0, int *nrows);

void cfd_b200_destroy(cfd_b200_solver *s);

/* Host State -> device.  NULL pointers are skipped.  For a slab handle the arrays hold the local rows only:
 * u (nrows+2) x (nx+3), v (nrows+3) x (nx+2), p (nrows+2) x (nx+2), first row = the ghost/halo row below. */
int cfd_b200_upload(cfd_b200_solver *s, const double *u, const double *v, const double *p);
int cfd_b200_upload_rhs(cfd_b200_solver *s, const double *rhs);
/* Device State -> host (blocks until the device is idle).  NULL pointers are skipped. */
int cfd_b200_download(cfd_b200_solver *s, double *u, double *v, double *p, double *rhs);
/* TimeIntegrator work arrays for tests: u1, v1 (time_integrator.hpp:18-19). */
int cfd_b200_download_work(cfd_b200_solver *s, double *u1, double *v1);

/* TimeIntegrator::clear_work (time_integrator.cpp:38-45) */
int cfd_b200_clear_work(cfd_b200_solver *s);
/* The drivers' reset (main.cpp:113-126, :338-346): u, v, p, rhs and the work arrays zeroed on the device, asynchronously. */
int cfd_b200_reset(cfd_b200_solver *s);

/* TimeIntegrator::step (time_integrator.hpp:27-29, time_integrator.cpp:47-239) on the device-resident State.
 * Synchronous: returns when the step is complete; *info (may be NULL) receives dt, residual, ... */
int cfd_b200_step(cfd_b200_solver *s, const cfd_b200_bc *bc, double Re, double CFL, const cfd_b200_params *pp,
                  double dt_override, cfd_b200_step_info *info);
/* Same step, enqueued without waiting (no host round trip inside the step: dt, alpha, beta, the convergence
 * test all live on the device).  cfd_b200_sync waits and returns the info of the LAST enqueued step. */
int cfd_b200_step_async(cfd_b200_solver *s, const cfd_b200_bc *bc, double Re, double CFL,
                        const cfd_b200_params *pp, double dt_override);
int cfd_b200_sync(cfd_b200_solver *s, cfd_b200_step_info *info);

/* Diagnostics the drivers call after every step (main.cpp:151, :255-256): metrics.cpp:68-86 / :52-66. */
int cfd_b200_divergence_l2(cfd_b200_solver *s, double *out);
int cfd_b200_max_velocity(cfd_b200_solver *s, double *out);

/* ---- stage-level entry points (the reference's public free functions, for kernel-level parity tests) ---- */
/* apply_bc_u / _v / _p (bc.hpp:32-34) on the device u, v, p; `which` is a bit mask 1=u, 2=v, 4=p. */
int cfd_b200_apply_bc(cfd_b200_solver *s, const cfd_b200_bc *bc, int which);
/* compute_cfl (metrics.hpp:6) */
int cfd_b200_compute_cfl(cfd_b200_solver *s, double Re, double CFL, double *dt_out);
/* The fused right-hand side TimeIntegrator::step uses: sanitize(advect + invRe * diffuse) of the device u, v
 * (time_integrator.cpp:75-82), returned in the reference host layout (du: like u, dv: like v). */
int cfd_b200_rhs_terms(cfd_b200_solver *s, double invRe, double *du, double *dv);
/* advect_u / advect_v alone (advection.hpp:9-10): zero-filled output, advective derivative, not sanitised.
 * Either pointer may be NULL. */
int cfd_b200_advect(cfd_b200_solver *s, double *du, double *dv);
/* diffuse_u (which = 0) / diffuse_v (which = 1) (diffusion.hpp:6-7): d += invRe * laplacian of the device u / v on the
 * interior set; d is a host array in the layout of u / v, updated in place. */
int cfd_b200_diffuse(cfd_b200_solver *s, int which, double invRe, double *d);
/* divergence (project.hpp:8): device rhs = div(u, v) */
int cfd_b200_divergence(cfd_b200_solver *s);
/* PressureSolver::solve (pressure.hpp:22) on the device p and rhs */
int cfd_b200_pressure_solve(cfd_b200_solver *s, const cfd_b200_params *pp, double *residual, int *iters);
/* subtract_grad_p (project.hpp:12) on the device u, v, p.  (Inside cfd_b200_step the same kernel also applies the
 * two p(0,0) = 0 pins of time_integrator.cpp:186-188; this entry point, like the reference function, does not.) */
int cfd_b200_subtract_grad_p(cfd_b200_solver *s, double dt);

/* ---- measurement ---- */
/* ---- Readers / mutators around the step ------------------------------------------------------
 * An interactive caller can colour a frame and apply obstacles without the State ever crossing PCIe.  On a row-slab handle
 * the calls are collective: `out` / `rgb` receive this slab's cell rows (cfd_b200_slab_rows), the statistics and the colour
 * scale are those of the whole grid, shapes are given in cells of the whole grid. */
/* ScalarField, viz.hpp:8 -- same order */
enum { CFD_B200_FIELD_U = 0, CFD_B200_FIELD_V, CFD_B200_FIELD_SPEED, CFD_B200_FIELD_PRESSURE, CFD_B200_FIELD_VORTICITY,
       CFD_B200_FIELD_STREAMLINES, CFD_B200_FIELD_TEMPERATURE };
typedef struct cfd_b200_field_stats {
    double vmin, vmax;   /* compute_scalar's out-parameters, span rule included (viz.cpp:63-64) */
    double mean, stddev; /* Gui::field_mean / field_std (gui.cpp:1452-1474) */
} cfd_b200_field_stats;
/* compute_scalar (viz.hpp:11-12, viz.cpp:6-65) on the device State: out (may be NULL) receives nx*ny doubles, index
 * i + j*nx; st (may be NULL) the statistics of Gui::update_field_statistics (gui.cpp:1433-1481). */
int cfd_b200_compute_scalar(cfd_b200_solver *s, int field, double *out, cfd_b200_field_stats *st);
/* update_field_texture (gui.hpp / gui.cpp:1505-1561) without the GL upload: rgb receives nx*ny*3 bytes.  colormap =
 * Gui::Colormap (gui.hpp:32-34; ids above 5 render as viridis, colormaps.cpp:124). */
int cfd_b200_render_rgb(cfd_b200_solver *s, int field, int colormap, float scale_min, float scale_max, int normalize,
                        float gamma, unsigned char *rgb);
/* Shape, gui.hpp:14-23 (pos / size in cell units) */
typedef struct cfd_b200_shape {
    int type;    /* 0 = rectangle, 1 = circle */
    float pos_x, pos_y, size_x, size_y;
    int bc_type; /* 0 = solid wall, 1 = porous */
    int enabled;
} cfd_b200_shape;
/* apply_shape_boundary_conditions (shapes.hpp:13, shapes.cpp:7-74) on the device u, v; asynchronous. */
int cfd_b200_apply_shapes(cfd_b200_solver *s, const cfd_b200_shape *shapes, int n);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches claim). */
long long cfd_b200_launch_count(const cfd_b200_solver *s);
/* CUDA-event time of the stages of the most recent synchronous step, in ms.  Fills up to `n` entries of
 * names[]/ms[]; returns the number of stages.  Enable with cfd_b200_set_profiling(s, 1) (adds event records). */
int cfd_b200_set_profiling(cfd_b200_solver *s, int on);
int cfd_b200_stage_times(cfd_b200_solver *s, int n, const char **names, float *ms);
/* Stream the handle launches on (a cudaStream_t), for callers that time with their own CUDA events. */
void *cfd_b200_stream(cfd_b200_solver *s);
/* Tuning switch for A/B measurements: 0 = simple kernels, 1 = optimised kernels (default). */
int cfd_b200_set_variant(cfd_b200_solver *s, int variant);
/* Which marching advection + diffusion + Runge-Kutta kernel the optimised path launches: 2 = k_rhs_march2 (default; slow path
 * deferred behind the loop), 1 = the first version (slow path inside the loop).  Same results bit for bit; for A/B timing.
 * The environment variable CFD_B200_MARCH sets the default of new handles. */
int cfd_b200_set_march_kernel(cfd_b200_solver *s, int version);

#ifdef __cplusplus
}
#endif
#endif /* CFD_B200_H */
