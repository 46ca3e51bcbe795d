// capi.cu -- implementation of include/cfd_b200.h: device State, launch sequences, host<->device copies.
// The launch sequence of one step follows TimeIntegrator::step (time_integrator.cpp:47-239) line by line; see
// enqueue_step().  No CPU fallback exists: every entry point needs a CUDA device.
#include "../../include/cfd_b200.h"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "comm.cuh"
#include "common.cuh"
#include "kernels_pressure.cuh"
#include "kernels_step.cuh"
#include "kernels_rhs_march.cuh"
#include "kernels_rhs_march2.cuh"
#include "kernels_pcg_cluster.cuh"
#include "kernels_rbgs_fused.cuh"
#include "kernels_rbgs_stream.cuh"
#include "kernels_vcycle.cuh"
#include "kernels_pcg_march.cuh"
#include "kernels_pcg_coop.cuh"
#include "p2p.cuh"
#include "kernels_viz.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                                \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return fail(e_ == cudaErrorMemoryAllocation ? CFD_B200_ENOMEM : CFD_B200_ECUDA, "%s:%d %s: %s",     \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                                     \
    } while (0)

enum { ST_BC0 = 0, ST_CFL, ST_RHS1, ST_BC1, ST_RHS2, ST_BC2, ST_DIV, ST_SOLVE, ST_PROJECT, ST_BC3, ST_DIV2, ST_END, ST_COUNT };
const char *const kStageNames[ST_COUNT - 1] = {"bc_pre", "cfl_dt", "rhs_stage1", "bc_stage1", "rhs_stage2", "bc_stage2",
                                               "divergence", "pressure_solve", "project", "bc_post", "divergence_post"};

} // namespace

struct cfd_b200_solver {
    Geom g{};
    int device = 0;
    int rank = 0, nranks = 1;
    double Lx = 1.0, Ly = 1.0;
    size_t n = 0; // doubles per field
    double *arena = nullptr; // the NFIELDS fields below + this rank's P2PBlock in ONE allocation (one IPC handle per rank)
    double *u = nullptr, *v = nullptr, *p = nullptr, *rhs = nullptr;
    double *p_alt = nullptr; // ping-pong partner of p for the fused red-black smoother
    double *u1 = nullptr, *v1 = nullptr;           // TimeIntegrator work arrays (time_integrator.hpp:18-19)
    double *ustar = nullptr, *vstar = nullptr;     // the unprojected velocity of a step (RK stage 2 -> out-of-place projection)
    double *r = nullptr, *sA = nullptr, *sB = nullptr; // PressureSolver work arrays (pressure.hpp:29); z, Ap never stored
    double *scratch_u = nullptr, *scratch_v = nullptr; // stage-level entry points only (lazy)
    double *viz_out = nullptr;                         // nx*ny visualisation scalar (lazy)
    unsigned char *viz_rgb = nullptr;                  // nx*ny*3 bytes (lazy)
    Scalars *d_sc = nullptr;
    Scalars *h_sc = nullptr; // pinned
    double *partials = nullptr;
    size_t partials_n = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr; // row slabs: the one-way sends run here, next to the solver's kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // The tile kernel of the ring strips (5 % of the faces, 35 us of small blocks at 8192^2) runs here, next to the marching
    // kernel of the same stage instead of behind it; CFD_B200_AUX=0 keeps both in the solver's stream.
    cudaStream_t aux_stream = nullptr;
    cudaStream_t launch_stream = nullptr; // non-null: where launch_kp puts the next kernels (set and cleared around a forked launch)
    cudaEvent_t ev_afork = nullptr, ev_ajoin = nullptr;
    // Captured CUDA graph of a steady-state step (single GPU, reference-mode mg): replayed while the step's arguments, the
    // launch sequence (cached CFL maxima, same BC) and the roles of p / p_alt stay what they were at capture time.  Two slots:
    // an odd number of smooth() calls per step swaps p and p_alt, so consecutive steps alternate between two graphs.
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        cfd_b200_bc bc{};
        cfd_b200_params pp{};
        double Re = 0, CFL = 0, dt_override = 0;
        int variant = 0, rb_strip = 0, march_kernel = 0;
        double *p = nullptr;
        long long launches = 0;
        bool swaps_p = false;
        int last_smoother = 0;
    } step_graph[2];
    int graph_mode = -1;                // -1 unread, 0 off (CFD_B200_GRAPH=0 or a failed capture), 1 on
    int graph_seen = 0;                 // consecutive eligible steps seen with the same key (capture on the second)
    bool comm_pending = false;          // a send was enqueued on comm_stream since the last join
    bool defer_rbgs = false;            // the residual sum of the step's last smooth() is completed by the reduction that ends the step
    double defer_tol = 0.0;
    long long launches = 0;
    int variant = 1;
    int march_kernel = -1; // marching advection kernel: 2 = k_rhs_march2 (default), 1 = first version; -1 = CFD_B200_MARCH not read yet
    int rb_strip = 0;      // > 0: force the streaming smoother with this many rows per strip (tests, tuning)
    int last_smoother = 0; // which kernel ran the last smooth(): 1 streaming, 2 tile, 3 per-colour passes
    int last_pcg = 0;      // which kernels ran the last PCG solve: 1 cooperative persistent, 2 streaming, 3 tiles, 4 one cluster
    int pdl = 3; // programmatic dependent launch between consecutive kernels (bit mask, see cfd_b200_create_slab; CFD_B200_PDL=0 turns it off)
    bool cfl_cached = false;   // umax/vmax of the current device State are already in d_sc (from the last step)
    cfd_b200_bc last_bc{};
    cfd_b200_bc last_bc_step{}; // BC of the previous step (the p halo rows sent at its end carry that BC's ghost columns)
    int profiling = 0;
    cudaEvent_t ev[ST_COUNT] = {};
    bool ev_valid = false;
    int sm_count = 148;
    ncclComm_t comm = nullptr; // row slabs only
    // peer-memory path (p2p.cuh): every rank's arena mapped through CUDA IPC
    bool p2p_on = false;
    void *peer_arena[P2P_MAXR] = {};
    size_t peer_n[P2P_MAXR] = {};
    int peer_ny[P2P_MAXR] = {};
    P2PBlock *blk[P2P_MAXR] = {};
    unsigned long long halo_epoch = 0, red_epoch = 0;
    unsigned long long p2p_timeout_ns = 20000000000ull;
    double *gathered = nullptr; // 4 doubles per rank (k_slab_pack / k_slab_unpack)
    double *bt_far = nullptr;   // both bottom/top Periodic: 5 rows received from the slab at the other end of the domain
    unsigned sent_since_barrier = 0;      // bit fi: a one-way send wrote the neighbours' halo rows of arena field fi since the last reduction
    unsigned long long ep_uv = 0;         // one-way send that delivers the u, v (and p) halo rows for the next step; 0 = none pending
    bool p_halo_valid = false;            // the 5 halo rows of p the smoother reads are current (sent at the end of the last step)
    bool halo_uv_valid = false; // the 2 halo rows of u, v are current (set by the exchange that ends a step)
    // V-cycle extension: coarse levels (level 0 aliases p / rhs), allocated on first use
    int mg_nlev = 0;
    MgLevel mg[MG_CTA_MAXLEV + 4] = {};
    MgLevel mg_gather = {}; // row slabs: the whole first one-block level, gathered onto every slab
    int mg_cta_smem_max = 0;
    // cooperative persistent PCG (small grids)
    double *coop_xrows = nullptr, *coop_slots = nullptr;
    int coop_ok = -1, coop_rows = 0, coop_blocks = 0, coop_smem = 0; // -1 = not probed yet
    int cluster_ok = 0, cluster_rows = 0, cluster_smem = 0;          // k_pcg_cluster (one 16-CTA cluster) can run this grid
};

namespace {

// Every kernel launch of the path: programmatic stream serialization (see pdl_begin(), common.cuh) unless disabled.
template <class... KArgs, class... Args>
inline void launch_kp(cfd_b200_solver *s, int pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s->launch_stream ? s->launch_stream : s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    s->launches++;
}
template <class... KArgs, class... Args>
inline void launch_k(cfd_b200_solver *s, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    launch_kp(s, s->pdl & 1, kernel, grid, block, smem, static_cast<Args &&>(args)...);
}
#define LAUNCH(s, kernel, grid, block, ...) launch_k(s, kernel, dim3(grid), dim3(block), 0, __VA_ARGS__)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int alloc_field(cfd_b200_solver *s, double **ptr) {
    CU(cudaMalloc(ptr, s->n * sizeof(double)));
    CU(cudaMemsetAsync(*ptr, 0, s->n * sizeof(double), s->stream));
    return 0;
}

void fill_geom(Geom &g, int nx, int ny_glob, double Lx, double Ly, int j0, int ny_loc, int bottom, int top) {
    g.nx = nx;
    g.ny = ny_loc;
    g.ny_glob = ny_glob;
    g.j0 = j0;
    g.bottom = bottom;
    g.top = top;
    g.P = ((nx + 3 + CFD_XOFF + 2) + 15) / 16 * 16;
    g.rows = ny_loc + 3 + 2 * CFD_YOFF;
    // Grid::init, grid.cpp:10-11 -- the same expressions on the host, so every constant is the reference's double
    g.dx = Lx / static_cast<double>(nx);
    g.dy = Ly / static_cast<double>(ny_glob);
    g.idx = 1.0 / g.dx;
    g.idy = 1.0 / g.dy;
    g.idx2 = 1.0 / (g.dx * g.dx);
    g.idy2 = 1.0 / (g.dy * g.dy);
    g.denom = 2.0 * (g.idx2 + g.idy2);
    g.inv_denom = 1.0 / g.denom;
    int e = 0;
    g.denom_pow2 = (std::frexp(g.denom, &e) == 0.5) ? 1 : 0;
    double diag = 2.0 * (g.idx2 + g.idy2);
    g.inv_diag = 1.0 / (diag < 1e-12 ? 1e-12 : diag); // 1.0 / std::max(diag, 1e-12), pressure.cpp:163
}

BcD to_bcd(const cfd_b200_bc *b) {
    BcD d;
    auto side = [](const cfd_b200_bc_side &s) { return BcSideD{s.type, s.moving, s.inflow_u, s.inflow_v}; };
    d.left = side(b->left);
    d.right = side(b->right);
    d.bottom = side(b->bottom);
    d.top = side(b->top);
    d.jet_center = b->jet_center;
    d.jet_width = b->jet_width;
    return d;
}

bool bc_equal(const cfd_b200_bc &a, const cfd_b200_bc &b) { // field by field: the structs carry padding bytes
    auto side = [](const cfd_b200_bc_side &x, const cfd_b200_bc_side &y) {
        return x.type == y.type && x.moving == y.moving && x.inflow_u == y.inflow_u && x.inflow_v == y.inflow_v;
    };
    return side(a.left, b.left) && side(a.right, b.right) && side(a.bottom, b.bottom) && side(a.top, b.top) &&
           a.jet_center == b.jet_center && a.jet_width == b.jet_width;
}

void stage_mark(cfd_b200_solver *s, int st) {
    if (s->profiling) cudaEventRecord(s->ev[st], s->stream);
}

// apply_bc_u / apply_bc_v / apply_bc_p in the reference's pass order (left/right, then bottom/top)
void join_comm(cfd_b200_solver *s);
void enqueue_bc(cfd_b200_solver *s, const BcD &bc, double *u, double *v, double *p, int sanitize, int halo = 1) {
    const Geom &g = s->g;
    const int nf = p ? 3 : 2; // blockIdx.y = field (u, v, p)
    LAUNCH(s, k_bc_lr, dim3(cdiv(g.ny + 5, 128), nf), 128, g, bc, u, v, p, sanitize, s->d_sc, halo);
    const double *far = nullptr;
    if (s->nranks > 1 && bc.bottom.type == BC_PERIODIC && bc.top.type == BC_PERIODIC && (g.bottom || g.top)) {
        // both bottom/top Periodic across slabs (bc.cpp:171-183, :318-330, :361-379): the pass swaps the boundary rows of the
        // two ends of the domain, which live on the first and the last slab.  They trade those rows (as left by the left/right
        // pass above) in front of the kernel: u 2 rows, v 2 rows, p 1 row.
        NcclApi &nc = nccl_api();
        join_comm(s);
        const int peer = g.bottom ? s->nranks - 1 : 0;
        const size_t P = g.P;
        auto row = [&](double *f, int jj) { return f + (size_t)(jj + CFD_YOFF) * P; };
        nc.GroupStart();
        if (u) nc.Send(row(u, g.bottom ? 1 : g.ny - 1), 2 * P, ncclDouble, peer, s->comm, s->stream), nc.Recv(s->bt_far, 2 * P, ncclDouble, peer, s->comm, s->stream);
        if (v) nc.Send(row(v, g.bottom ? 1 : g.ny), 2 * P, ncclDouble, peer, s->comm, s->stream), nc.Recv(s->bt_far + 2 * P, 2 * P, ncclDouble, peer, s->comm, s->stream);
        if (p) nc.Send(row(p, g.bottom ? 1 : g.ny), P, ncclDouble, peer, s->comm, s->stream), nc.Recv(s->bt_far + 4 * P, P, ncclDouble, peer, s->comm, s->stream);
        nc.GroupEnd();
        far = s->bt_far;
    }
    if (g.bottom || g.top) LAUNCH(s, k_bc_bt, dim3(cdiv(g.nx + 3, 128), nf), 128, g, bc, u, v, p, sanitize, s->d_sc, far);
}

int reduce_fin(cfd_b200_solver *s, int kind, double tol, int max_iters, double *halo = nullptr, int rows = 0, int dirs = 0);
WaitArgs wait_args(const cfd_b200_solver *s, unsigned long long epoch);
int halo_wait(cfd_b200_solver *s, unsigned long long epoch);
void join_comm(cfd_b200_solver *s);

void enqueue_cfl(cfd_b200_solver *s, double Re, double CFL, double dt_override, int cfl_only) {
    const Geom &g = s->g;
    if (!s->cfl_cached) {
        // max|u|, max|v| of compute_cfl.  Skipped when the previous step's post-projection divergence kernel already
        // left them in d_sc: the BC pass in between only rewrites ring faces with the values they already hold (or,
        // for periodic sides, swaps two columns), so the maxima over all faces are unchanged.
        cudaMemsetAsync(&s->d_sc->umax_bits, 0, 2 * sizeof(unsigned long long), s->stream);
        dim3 grid(cdiv(g.P / 2, 256), g.ny + 1 < 64 ? g.ny + 1 : 64);
        LAUNCH(s, k_absmax, grid, 256, g, s->u, s->v, s->d_sc);
        reduce_fin(s, RED_CFL, 0.0, 0);
    }
    s->cfl_cached = false;
    LAUNCH(s, k_dt, 1, 1, g, Re, CFL, dt_override, s->d_sc, cfl_only);
}

// What the marching kernel covered (MODE 4 needs it for the fix-up of the divergence): marched = false when the grid is too
// small for it and the tile kernel did everything.
struct MarchRegion {
    bool marched;
    int XA, XB, YA, YB, RY;
};
template <int MODE>
MarchRegion enqueue_rhs(cfd_b200_solver *s, double invRe, const double *uin, const double *vin, double *uout, double *vout,
                        const double *uacc = nullptr, const double *vacc = nullptr, unsigned long long wait_epoch = 0) {
    if (!uacc) uacc = uout, vacc = vout; // MODE 2 in place
    const WaitArgs hw = wait_args(s, wait_epoch); // row slabs: the blocks that read the neighbours' rows wait for this send
    const Geom &g = s->g;
    RhsConst k{g.idx, g.idy, g.idx2, g.idy2, invRe};
    const dim3 block(RHS_TX, RHS_TY);
    const int ntx = cdiv(g.nx + 1, RHS_TX), nty = cdiv(g.ny + 1, RHS_TY);
    RhsRects rc{};
    auto add = [&](int tx0, int ty0, int ntx_, int nty_) {
        if (ntx_ <= 0 || nty_ <= 0) return;
        const int k = rc.n++;
        rc.tx0[k] = tx0, rc.ty0[k] = ty0, rc.ntx[k] = ntx_;
        rc.end[k] = (k ? rc.end[k - 1] : 0) + ntx_ * nty_;
    };
    constexpr int SMODE = MODE == 4 ? 2 : MODE; // the tile kernel has no divergence of its own
    auto launch_rects = [&]() {
        if (rc.n > 0) LAUNCH(s, k_rhs_simple<SMODE>, rc.end[rc.n - 1], block, g, k, uin, vin, uacc, vacc, uout, vout, s->d_sc, rc, hw);
    };
    // Interior faces (all four fluxes MUSCL, for u and v alike): 2 <= ii <= nx-1, 2 <= jj <= ny-1.  The marching kernel
    // takes the part of it that is aligned to the simple kernel's tiles; the tile strips around it (ring included)
    // stay with the simple kernel.  A slab that does not touch the bottom / top has no ring there.
    const int txr = (g.nx - 1) / RHS_TX;                        // first tile column of the right strip
    const int tyb = g.bottom ? 1 : 0;                           // tile rows of the bottom strip
    const int tyt = g.top ? (g.ny - 1) / RHS_TY : nty;          // first tile row of the top strip
    const int XA = RHS_TX + 1, XB = txr * RHS_TX, YA = tyb * RHS_TY + 1, YB = g.top ? tyt * RHS_TY : g.ny;
    if (s->variant == 0 || XB < XA + 16 || YB < YA + 8) {
        add(0, 0, ntx, nty);
        launch_rects();
        return MarchRegion{false, 0, 0, 0, 0, 0};
    }
    static const int ry_env = getenv("CFD_B200_RHS_RY") ? atoi(getenv("CFD_B200_RHS_RY")) : 0;
    // Rows per strip: 64 on grids that fill the GPU; smaller grids get shorter strips until the launch has one full
    // wave of blocks (148 SMs x MARCH_MINBLOCKS) -- at 1024^2, 64-row strips are only 80 blocks (2 warps per SM).
    // 128 where that still gives two full waves (8192^2: 2.398 -> 2.377 ms per step: 4 of 132 rows read twice instead of 4 of 68).
    const int gxm = cdiv(cdiv(XB - XA + 1, MARCH_COLS), MARCH_WARPS);
    int RY = (long long)gxm * cdiv(YB - YA + 1, 128) >= 2LL * s->sm_count * MARCH_MINBLOCKS ? 128 : 64;
    while (RY > 8 && (long long)gxm * cdiv(YB - YA + 1, RY) < (long long)s->sm_count * MARCH_MINBLOCKS) RY /= 2;
    if (ry_env > 0) RY = ry_env;
    dim3 mgrid(gxm, cdiv(YB - YA + 1, RY));
    add(0, 0, ntx, tyb);                    // bottom strip
    add(0, tyt, ntx, nty - tyt);            // top strip
    add(0, tyb, 1, tyt - tyb);              // left strip
    add(txr, tyb, ntx - txr, tyt - tyb);    // right strip
    // One launch for all four strips.  It writes other faces than the marching kernel and reads the same inputs, so it runs
    // beside it on the auxiliary stream (forked here, joined behind the marching kernel's launch; in-place stage 2 of
    // `variant = 2` keeps the old order: there the tile kernel reads faces next to the ones the marching kernel overwrites).
    const bool fork = s->aux_stream && rc.n > 0 && uin != uout && vin != vout && (!(MODE == 2 || MODE == 4) || (uacc != uout && vacc != vout));
    if (fork) {
        cudaEventRecord(s->ev_afork, s->stream);
        cudaStreamWaitEvent(s->aux_stream, s->ev_afork, 0);
        s->launch_stream = s->aux_stream;
        const int pdl = s->pdl;
        s->pdl = 0; // its predecessor in that stream is an event, not a kernel
        launch_rects();
        s->pdl = pdl;
        s->launch_stream = nullptr;
    }
    auto join_strips = [&]() {
        if (fork) {
            cudaEventRecord(s->ev_ajoin, s->aux_stream);
            cudaStreamWaitEvent(s->stream, s->ev_ajoin, 0);
        } else {
            launch_rects();
        }
    };
    // k_rhs_march2 (kernels_rhs_march2.cuh) defers its slow path: inputs, accumulators and outputs must not alias (in-place
    // stage 2 of `variant = 2` does), a thread stores both of its columns or none (XA odd, XB even) and indexes a field with
    // 32 bits.  Anything else, and CFD_B200_MARCH=1 / cfd_b200_set_march_kernel(s, 1), runs the first version.
    if (s->march_kernel < 0) s->march_kernel = getenv("CFD_B200_MARCH") ? atoi(getenv("CFD_B200_MARCH")) : 2;
    const bool heun = MODE == 2 || MODE == 4; // the others never read uacc / vacc (set to the outputs above)
    const bool v2 = s->march_kernel != 1 && (!heun || (uacc != uout && vacc != vout)) && uin != uout && vin != vout && (XA & 1) && !(XB & 1) &&
                    (unsigned long long)g.rows * (unsigned long long)g.P < (1ULL << 32);
    if (v2)
        LAUNCH(s, k_rhs_march2<MODE>, mgrid, 32 * MARCH_WARPS, g, k, uin, vin, uacc, vacc, uout, vout, s->d_sc, XA, XB, YA, YB, RY,
               hw, MODE == 4 ? s->rhs : (double *)nullptr, s->partials);
    else
        LAUNCH(s, k_rhs_march<MODE>, mgrid, 32 * MARCH_WARPS, g, k, uin, vin, uacc, vacc, uout, vout, s->d_sc, XA, XB, YA, YB, RY, hw,
               MODE == 4 ? s->rhs : (double *)nullptr, s->partials);
    join_strips();
    return MarchRegion{true, XA, XB, YA, YB, RY};
}

dim3 row_grid(const cfd_b200_solver *s, int ncols, int nrows, int *rows_per_block) {
    // <= ~4096 blocks: one x-block per 256 columns, contiguous row chunks in y
    int gx = cdiv(ncols, 256);
    int gy_max = 4096 / gx;
    if (gy_max < 1) gy_max = 1;
    int rpb = cdiv(nrows, gy_max);
    if (rpb < 4) rpb = 4;
    *rows_per_block = rpb;
    return dim3(gx, cdiv(nrows, rpb));
}

void enqueue_divergence(cfd_b200_solver *s, int pre, int want_sum, int want_max = 0, const double *u = nullptr, const double *v = nullptr,
                        unsigned long long wait_epoch = 0) {
    const Geom &g = s->g;
    if (!u) u = s->u, v = s->v;
    const WaitArgs hw = wait_args(s, wait_epoch);
    int rpb = 4;
    dim3 grid = row_grid(s, g.nx, g.ny, &rpb);
    if (pre)
        LAUNCH(s, k_divergence<1>, grid, 256, g, u, v, s->rhs, s->p, s->d_sc, s->partials, want_sum, rpb, 0, hw);
    else
        LAUNCH(s, k_divergence<0>, grid, 256, g, u, v, s->rhs, (double *)nullptr, s->d_sc, s->partials, want_sum, rpb, want_max, hw);
}

int pcg_grid(const cfd_b200_solver *s, int ntiles) {
    int cap = s->sm_count * 4;
    return ntiles < cap ? ntiles : cap;
}

// ---- row-slab communication (no-ops on a single-GPU handle) -----------------------------------------------------
enum { HALO_UP = 1, HALO_DOWN = 2, HALO_BOTH = 3 }; // direction in which OWNED rows are sent

#define NCCLCHK(call)                                                                                           \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess) return fail(CFD_B200_ECOMM, "%s:%d %s: %s", __FILE__, __LINE__, #call,           \
                                           nccl_api().GetErrorString(r_));                                      \
    } while (0)

#define TRY(call)                                                                                               \
    do {                                                                                                        \
        int rc_ = (call);                                                                                       \
        if (rc_) return rc_;                                                                                    \
    } while (0)

#define NFIELDS 12 // u, v, p, p_alt, rhs, u1, v1, r, sA, sB, ustar, vstar: arena slots of n doubles each

// ---- peer-memory bootstrap: exchange CUDA IPC handles over NCCL, map every peer's arena, verify, agree ------------
struct P2PInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long n; // doubles per field on that rank (slabs may differ by one row)
    int ny;
    int ok;
};

// Wait for the stream, but never for more than `seconds` (a dead peer must not hang teardown).
bool bounded_sync(cfd_b200_solver *s, double seconds) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        cudaError_t e = cudaStreamQuery(s->stream);
        if (e == cudaSuccess) return true;
        if (e != cudaErrorNotReady) { cudaGetLastError(); return false; }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) return false;
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

void p2p_unmap(cfd_b200_solver *s) {
    for (int r = 0; r < s->nranks; ++r) {
        if (r != s->rank && s->peer_arena[r]) {
            cudaIpcCloseMemHandle(s->peer_arena[r]);
            s->peer_arena[r] = nullptr;
            s->blk[r] = nullptr;
        }
    }
    cudaGetLastError();
}

int p2p_setup_impl(cfd_b200_solver *s, P2PInfo *&d_all);
int p2p_setup(cfd_b200_solver *s) {
    P2PInfo *d_all = nullptr;
    const int rc = p2p_setup_impl(s, d_all);
    if (d_all) cudaFree(d_all);
    if (rc) p2p_unmap(s), s->p2p_on = false; // nothing stays mapped behind a failed bootstrap
    return rc;
}
int p2p_setup_impl(cfd_b200_solver *s, P2PInfo *&d_all) {
    static const int env = getenv("CFD_B200_P2P") ? atoi(getenv("CFD_B200_P2P")) : 1;
    NcclApi &nc = nccl_api();
    const int R = s->nranks;
    P2PInfo mine{};
    mine.n = s->n, mine.ny = s->g.ny;
    const unsigned long long magic = P2P_MAGIC + (unsigned long long)s->rank;
    mine.ok = env && cudaMemcpyAsync(&s->blk[s->rank]->magic, &magic, sizeof(magic), cudaMemcpyHostToDevice, s->stream) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.handle, s->arena) == cudaSuccess;
    cudaGetLastError();
    // all-gather of the descriptors (bytes) on the solver's stream: ordered after the arena memset and the magic
    CU(cudaMalloc(&d_all, (size_t)R * sizeof(P2PInfo)));
    CU(cudaMemcpyAsync(d_all + s->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream));
    NCCLCHK(nc.AllGather(d_all + s->rank, d_all, sizeof(P2PInfo), ncclInt8, s->comm, s->stream));
    P2PInfo all[P2P_MAXR];
    CU(cudaMemcpyAsync(all, d_all, (size_t)R * sizeof(P2PInfo), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    int ok = 1;
    for (int r = 0; r < R; ++r) ok &= all[r].ok;
    for (int r = 0; r < R && ok; ++r) {
        s->peer_n[r] = (size_t)all[r].n, s->peer_ny[r] = all[r].ny;
        if (r == s->rank) continue;
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        s->peer_arena[r] = ptr;
        s->blk[r] = reinterpret_cast<P2PBlock *>(static_cast<double *>(ptr) + (size_t)NFIELDS * s->peer_n[r]);
        unsigned long long seen = 0; // read the peer's magic THROUGH the mapping: proves base address and peer access
        if (cudaMemcpy(&seen, &s->blk[r]->magic, sizeof(seen), cudaMemcpyDeviceToHost) != cudaSuccess ||
            seen != P2P_MAGIC + (unsigned long long)r)
            ok = 0;
    }
    cudaGetLastError();
    // every rank must take the same path: min over ranks of `ok`
    double okd = ok;
    CU(cudaMemcpyAsync(s->gathered, &okd, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    NCCLCHK(nc.AllReduce(s->gathered, s->gathered, 1, ncclDouble, ncclMin, s->comm, s->stream));
    CU(cudaMemcpyAsync(&okd, s->gathered, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->p2p_on = okd > 0.5;
    if (!s->p2p_on) p2p_unmap(s);
    s->peer_arena[s->rank] = s->arena;
    return 0;
}

// Peers store into my arena and I into theirs: nobody may unmap or free before everybody has stopped (two barriers).
void p2p_teardown(cfd_b200_solver *s) {
    if (!s->p2p_on || !s->comm) return;
    NcclApi &nc = nccl_api();
    bool alive = nc.AllReduce(s->gathered, s->gathered, 1, ncclDouble, ncclMin, s->comm, s->stream) == ncclSuccess && bounded_sync(s, 10.0);
    p2p_unmap(s);
    if (alive) alive = nc.AllReduce(s->gathered, s->gathered, 1, ncclDouble, ncclMin, s->comm, s->stream) == ncclSuccess && bounded_sync(s, 10.0);
    s->p2p_on = false;
}

// Peer-memory version of the exchange below: one kernel, plain stores into the neighbours' halo rows (p2p.cuh).
int halo_exchange_p2p(cfd_b200_solver *s, double *const *fields, int nfields, int rows, int dirs) {
    const Geom &g = s->g;
    HaloArgs a{};
    a.count = (unsigned long long)rows * g.P;
    a.epoch = ++s->halo_epoch;
    const int lower = g.bottom ? -1 : s->rank - 1, upper = g.top ? -1 : s->rank + 1;
    auto rowptr = [&](double *f, int jj) { return f + (size_t)(jj + CFD_YOFF) * g.P; };
    auto peer_row = [&](int r, size_t fi, int jj) {
        return static_cast<double *>(s->peer_arena[r]) + fi * s->peer_n[r] + (size_t)(jj + CFD_YOFF) * g.P;
    };
    for (int k = 0; k < nfields; ++k) {
        double *f = fields[k];
        const size_t fi = (size_t)(f - s->arena) / s->n;
        if (f < s->arena || fi >= NFIELDS) return fail(CFD_B200_EINVAL, "halo exchange of a field outside the arena");
        if (upper >= 0 && (dirs & HALO_UP)) a.job[a.njobs++] = HaloJob{rowptr(f, g.ny - rows + 1), peer_row(upper, fi, 1 - rows), 1};
        if (lower >= 0 && (dirs & HALO_DOWN)) a.job[a.njobs++] = HaloJob{rowptr(f, 1), peer_row(lower, fi, s->peer_ny[lower] + 1), 0};
    }
    if (lower >= 0) a.peer_enter[0] = &s->blk[lower]->enter_flag[1], a.peer_data[0] = &s->blk[lower]->data_flag[1];
    if (upper >= 0) a.peer_enter[1] = &s->blk[upper]->enter_flag[0], a.peer_data[1] = &s->blk[upper]->data_flag[0];
    int grid = (int)((a.count / 2 + 255) / 256);
    grid = grid < 1 ? 1 : (grid > 32 ? 32 : grid);
    LAUNCH(s, k_halo_xchg, grid, 256, a, s->blk[s->rank], s->d_sc);
    return 0;
}

// Send the top `rows` owned rows of each field to the slab above (HALO_UP) and/or the bottom `rows` owned rows to the
// slab below (HALO_DOWN); receive the neighbours' rows into the halo rows.  Rows are contiguous: one message each.
int halo_exchange(cfd_b200_solver *s, double *const *fields, int nfields, int rows, int dirs) {
    if (s->nranks == 1) return 0;
    join_comm(s);
    if (s->p2p_on) return halo_exchange_p2p(s, fields, nfields, rows, dirs);
    const Geom &g = s->g;
    NcclApi &n = nccl_api();
    const size_t cnt = (size_t)rows * g.P;
    auto rowptr = [&](double *f, int jj) { return f + (size_t)(jj + CFD_YOFF) * g.P; };
    NCCLCHK(n.GroupStart());
    for (int k = 0; k < nfields; ++k) {
        double *f = fields[k];
        if (!g.top) { // neighbour above = rank + 1
            if (dirs & HALO_UP) NCCLCHK(n.Send(rowptr(f, g.ny - rows + 1), cnt, ncclDouble, s->rank + 1, s->comm, s->stream));
            if (dirs & HALO_DOWN) NCCLCHK(n.Recv(rowptr(f, g.ny + 1), cnt, ncclDouble, s->rank + 1, s->comm, s->stream));
        }
        if (!g.bottom) { // neighbour below = rank - 1
            if (dirs & HALO_DOWN) NCCLCHK(n.Send(rowptr(f, 1), cnt, ncclDouble, s->rank - 1, s->comm, s->stream));
            if (dirs & HALO_UP) NCCLCHK(n.Recv(rowptr(f, 1 - rows), cnt, ncclDouble, s->rank - 1, s->comm, s->stream));
        }
    }
    NCCLCHK(n.GroupEnd());
    return 0;
}
int halo_exchange2(cfd_b200_solver *s, double *a, double *b, int rows, int dirs) {
    double *f[2] = {a, b};
    return halo_exchange(s, f, b ? 2 : 1, rows, dirs);
}
// The solver's stream waits for everything enqueued on the communication stream so far (free when it has finished).
void join_comm(cfd_b200_solver *s) {
    if (!s->comm_pending) return;
    cudaEventRecord(s->ev_join, s->comm_stream);
    cudaStreamWaitEvent(s->stream, s->ev_join, 0);
    s->comm_pending = false;
}

// One-way version of the exchange (p2p.cuh: k_halo_send): my edge rows go into the neighbours' halo rows, nothing waits.
// fields[k] sends rows[k] rows in direction(s) dirs[k].  *epoch receives what halo_wait / the consumer's edge blocks wait
// for (0: the rows are already in place in stream order -- single GPU, the NCCL transport, or the handshake fallback).
// Without the handshake a send is only safe if the neighbours have stopped reading the rows it overwrites: true when a
// reduction (a barrier) has passed since the last send into the same field; otherwise the old exchange is used.
int halo_send(cfd_b200_solver *s, double *const *fields, const int *rows, const int *dirs, int nfields, unsigned long long *epoch) {
    *epoch = 0;
    if (s->nranks == 1) return 0;
    const Geom &g = s->g;
    unsigned mask = 0;
    for (int k = 0; k < nfields; ++k) {
        const size_t fi = (size_t)(fields[k] - s->arena) / s->n;
        if (fields[k] < s->arena || fi >= NFIELDS) return fail(CFD_B200_EINVAL, "halo exchange of a field outside the arena");
        mask |= 1u << fi;
    }
    if (!s->p2p_on || (s->sent_since_barrier & mask)) {
        for (int k = 0; k < nfields; ++k) TRY(halo_exchange(s, &fields[k], 1, rows[k], dirs[k]));
        return 0;
    }
    s->sent_since_barrier |= mask;
    SendArgs a{};
    a.epoch = ++s->halo_epoch;
    const int lower = g.bottom ? -1 : s->rank - 1, upper = g.top ? -1 : s->rank + 1;
    int anydir = 0;
    for (int k = 0; k < nfields; ++k) {
        double *f = fields[k];
        const size_t fi = (size_t)(f - s->arena) / s->n;
        const unsigned long long cnt = (unsigned long long)rows[k] * g.P;
        auto rowptr = [&](int jj) { return f + (size_t)(jj + CFD_YOFF) * g.P; };
        auto peer_row = [&](int r, int jj) { return static_cast<double *>(s->peer_arena[r]) + fi * s->peer_n[r] + (size_t)(jj + CFD_YOFF) * g.P; };
        if (upper >= 0 && (dirs[k] & HALO_UP)) a.job[a.njobs++] = SendJob{rowptr(g.ny - rows[k] + 1), peer_row(upper, 1 - rows[k]), cnt};
        if (lower >= 0 && (dirs[k] & HALO_DOWN)) a.job[a.njobs++] = SendJob{rowptr(1), peer_row(lower, s->peer_ny[lower] + 1), cnt};
        anydir |= dirs[k];
    }
    if (upper >= 0 && (anydir & HALO_UP)) a.peer_data[1] = &s->blk[upper]->data_flag[0];
    if (lower >= 0 && (anydir & HALO_DOWN)) a.peer_data[0] = &s->blk[lower]->data_flag[1];
    unsigned long long total = 0;
    for (int j = 0; j < a.njobs; ++j) total += a.job[j].count;
    int grid = (int)((total / 2 + 2047) / 2048);
    grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
    // On its own stream, behind everything the solver's stream holds so far: the next kernels do not wait for it (what
    // they need from the NEIGHBOURS they wait for themselves), so the store, the fence and the flag -- a chain of NVLink
    // latencies, ~10 us -- overlap with computation.  join_comm() puts the solver's stream behind it again.
    cudaEventRecord(s->ev_fork, s->stream);
    cudaStreamWaitEvent(s->comm_stream, s->ev_fork, 0);
    k_halo_send<<<grid, 256, 0, s->comm_stream>>>(a, s->blk[s->rank], s->d_sc);
    s->launches++;
    s->comm_pending = true;
    *epoch = a.epoch | ((unsigned long long)anydir << 62); // the directions travel with the epoch (epochs stay far below 2^62)
    return 0;
}
WaitArgs wait_args(const cfd_b200_solver *s, unsigned long long epoch) {
    WaitArgs w{};
    static const int nowait = getenv("CFD_B200_DEBUG_NOWAIT") ? atoi(getenv("CFD_B200_DEBUG_NOWAIT")) : 0; // timing experiments only
    if (nowait) return w;
    const int dirs = (int)(epoch >> 62);
    w.epoch = epoch & ((1ull << 62) - 1);
    if (!s->g.bottom && (dirs & HALO_UP)) w.f[0] = &s->blk[s->rank]->data_flag[0];  // rows sent UP by the slab below
    if (!s->g.top && (dirs & HALO_DOWN)) w.f[1] = &s->blk[s->rank]->data_flag[1];   // rows sent DOWN by the slab above
    return w;
}
int halo_wait(cfd_b200_solver *s, unsigned long long epoch) {
    if (!epoch) return 0;
    const WaitArgs w = wait_args(s, epoch);
    if (w.f[0] || w.f[1]) LAUNCH(s, k_halo_wait, 1, 32, w, s->d_sc);
    return 0;
}
int halo_send2(cfd_b200_solver *s, double *a, double *b, int rows, int dirs, unsigned long long *epoch) {
    double *f[2] = {a, b};
    const int r[2] = {rows, rows}, d[2] = {dirs, dirs};
    return halo_send(s, f, r, d, b ? 2 : 1, epoch);
}
// in-place allreduce of `count` doubles in device memory (NCCL path)
int allreduce(cfd_b200_solver *s, void *dev, int count, ncclRedOp_t op) {
    if (s->nranks == 1) return 0;
    NCCLCHK(nccl_api().AllReduce(dev, dev, (size_t)count, ncclDouble, op, s->comm, s->stream));
    return 0;
}
// Row slabs: complete a grid-level reduction across ranks and run the scalar tail that consumes it (RED_* of p2p.cuh).
// Peer-memory path: one fused kernel.  NCCL path: allreduce / all-gather + a one-thread tail kernel.
// `halo` != nullptr: `rows` rows of that field are exchanged (dirs as in halo_exchange) right behind the reduction --
// in the same kernel on the peer-memory path, as a separate exchange on the NCCL path.
int reduce_fin(cfd_b200_solver *s, int kind, double tol, int max_iters, double *halo, int rows, int dirs) {
    if (s->nranks == 1) return 0;
    join_comm(s); // flags are written in epoch order: nothing of the communication stream may still be on its way
    s->sent_since_barrier = 0; // every rank has left the kernels in front of this point once it has passed it
    if (s->p2p_on) {
        RedArgs a{};
        for (int r = 0; r < s->nranks; ++r) a.blk[r] = s->blk[r];
        a.nranks = s->nranks, a.rank = s->rank, a.kind = kind, a.epoch = ++s->red_epoch, a.tol = tol, a.max_iters = max_iters;
        if (kind == RED_STEP_END && s->defer_rbgs) a.defer_rbgs = 1, a.tol = s->defer_tol, s->defer_rbgs = false;
        if (halo) {
            const Geom &g = s->g;
            const int lower = g.bottom ? -1 : s->rank - 1, upper = g.top ? -1 : s->rank + 1;
            const size_t fi = (size_t)(halo - s->arena) / s->n;
            if (halo < s->arena || fi >= NFIELDS) return fail(CFD_B200_EINVAL, "halo exchange of a field outside the arena");
            auto rowptr = [&](double *f, int jj) { return f + (size_t)(jj + CFD_YOFF) * g.P; };
            auto peer_row = [&](int r, int jj) {
                return static_cast<double *>(s->peer_arena[r]) + fi * s->peer_n[r] + (size_t)(jj + CFD_YOFF) * g.P;
            };
            a.count = (unsigned long long)rows * g.P;
            a.hepoch = ++s->halo_epoch;
            if (upper >= 0 && (dirs & HALO_UP)) a.job[a.njobs++] = HaloJob{rowptr(halo, g.ny - rows + 1), peer_row(upper, 1 - rows), 1};
            if (lower >= 0 && (dirs & HALO_DOWN)) a.job[a.njobs++] = HaloJob{rowptr(halo, 1), peer_row(lower, s->peer_ny[lower] + 1), 0};
            // both neighbours exchange flags even when rows travel one way only (uniform epochs)
            if (lower >= 0) a.peer_data[0] = &s->blk[lower]->data_flag[1];
            if (upper >= 0) a.peer_data[1] = &s->blk[upper]->data_flag[0];
            a.halo = 1; // an end slab with nothing to send still signals and waits
        }
        LAUNCH(s, k_p2p_reduce, 1, P2P_RED_THREADS, a, s->d_sc);
        return 0;
    }
    if (halo) { // NCCL path: the reduction first, then the rows
        TRY(reduce_fin(s, kind, tol, max_iters, nullptr, 0, 0));
        return halo_exchange2(s, halo, nullptr, rows, dirs);
    }
    switch (kind) {
    case RED_PCG_INIT:
        TRY(allreduce(s, s->d_sc->red, 2, ncclSum));
        LAUNCH(s, k_pcg_fin_init, 1, 1, s->d_sc, tol, max_iters);
        break;
    case RED_PCG1:
        TRY(allreduce(s, s->d_sc->red, 1, ncclSum));
        LAUNCH(s, k_pcg_fin1, 1, 1, s->d_sc);
        break;
    case RED_PCG2:
        TRY(allreduce(s, s->d_sc->red, 2, ncclSum));
        LAUNCH(s, k_pcg_fin2, 1, 1, s->d_sc, tol, max_iters);
        break;
    case RED_RBGS:
        TRY(allreduce(s, s->d_sc->red, 1, ncclSum));
        LAUNCH(s, k_rbgs_fin, 1, 1, s->d_sc, tol);
        break;
    case RED_VC:
        TRY(allreduce(s, s->d_sc->red, 1, ncclSum));
        LAUNCH(s, k_mg_norm_fin, 1, 1, s->d_sc, tol);
        break;
    case RED_STEP_END: // one collective: CFL maxima of the next step + the two partial sums of squares
        LAUNCH(s, k_slab_pack, 1, 1, s->d_sc);
        NCCLCHK(nccl_api().AllGather(s->d_sc->red, s->gathered, 4, ncclDouble, s->comm, s->stream));
        LAUNCH(s, k_slab_unpack, 1, 1, s->d_sc, s->gathered, s->nranks);
        break;
    case RED_CFL: // non-negative doubles order like their bit patterns, so the maxima can be all-reduced as doubles
        TRY(allreduce(s, &s->d_sc->umax_bits, 2, ncclMax));
        break;
    case RED_DIAG:
        TRY(allreduce(s, &s->d_sc->diag_sum, 1, ncclSum));
        TRY(allreduce(s, &s->d_sc->diag_max_bits, 1, ncclMax));
        break;
    case RED_VIZ:
        TRY(allreduce(s, &s->d_sc->viz[0], 1, ncclMin));
        TRY(allreduce(s, &s->d_sc->viz[1], 1, ncclMax));
        TRY(allreduce(s, &s->d_sc->viz[2], 2, ncclSum));
        break;
    default:
        return fail(CFD_B200_EINVAL, "bad reduction kind %d", kind);
    }
    return 0;
}

template <int POW2>
void launch_fused(cfd_b200_solver *s, double tol, int ntx, int nty, int finish, int only_if_redo) {
    const Geom &g = s->g;
    const int ntiles = ntx * nty;
    // stand-alone: one block per tile.  As the slow path of the streaming smoother: two blocks per SM walking over the
    // tiles, so that the launch costs next to nothing when there is nothing to redo.
    const int grid = only_if_redo && ntiles > 2 * s->sm_count ? 2 * s->sm_count : ntiles;
    if ((1 + g.j0) & 1)
        launch_kp(s, only_if_redo ? (s->pdl >> 2) & 1 : s->pdl & 1, k_rbgs_fused<POW2, 1>, dim3(grid), dim3(RB_THREADS), RB_SMEM, g, s->p, s->p_alt, s->rhs, s->d_sc, s->partials, tol, ntx, ntiles, finish, only_if_redo);
    else
        launch_kp(s, only_if_redo ? (s->pdl >> 2) & 1 : s->pdl & 1, k_rbgs_fused<POW2, 0>, dim3(grid), dim3(RB_THREADS), RB_SMEM, g, s->p, s->p_alt, s->rhs, s->d_sc, s->partials, tol, ntx, ntiles, finish, only_if_redo);
}

// Strip height and grid of the streaming smoother for an nx-wide grid with `nrows` stored rows; false when the grid is
// too small to fill the GPU with it (the caller then uses its tile / per-colour kernels).
// Pure host arithmetic (no device needed): strip height for an nx-wide grid with `nrows` stored rows on a GPU with
// `sm_count` SMs, or 0 when the streaming smoother would leave SMs empty and the tile kernel is the better choice.
int stream_strip_rows(int nx, int nrows, int sm_count, int *gx_out, int *nstrips_out) {
    const int nstrips = cdiv(nx + 1, RS_OUT), gx = cdiv(nstrips, RS_WARPS);
    // Rows per strip.  Tall strips waste fewer steps on the pipeline fill, short ones give more blocks.  Cost model fitted on
    // B200 with the round-2 kernel (scripts/strip_sweep.py, 8192x1024 ... 16384x4096): a block costs its steps -- whole turns
    // of the 16-row ring, 16 * ceil((RY + 15) / 16) -- plus ~12 steps of fixed overhead; the blocks run in rounds of
    // RS_BLOCKS_PER_SM per SM, a full round at unit cost per step and a partly filled last round at 0.5 + 0.5 x fill (fewer
    // resident warps hide less latency, but each runs faster).  It picks the measured optimum in every case tried:
    // 8192^2 and 16384x4096 -> 257, 8192x4096, 8192x2048 and 4096^2 -> 193, 8192x1024 -> 97, 2048^2 -> 49.
    int RY = 257;
    const double slots = (double)RS_BLOCKS_PER_SM * sm_count;
    double best = 1e300;
    static const int cand[] = {257, 193, 129, 97, 65, 49, 33};
    for (int ry : cand) {
        const double blocks = (double)gx * cdiv(nrows, ry);
        const double full = std::floor(blocks / slots), rem = blocks - full * slots;
        const double steps = 16.0 * cdiv(ry + 15, 16) + 12.0;
        const double cost = steps * (full + (rem > 0 ? 0.5 + 0.5 * rem / slots : 0.0));
        if (cost < best) best = cost, RY = ry;
    }
    if (gx_out) *gx_out = gx;
    if (nstrips_out) *nstrips_out = nstrips;
    // too few warps to fill the GPU: the tile kernel is the better one (1024^2: 21 us per smooth against 25)
    return (long long)gx * cdiv(nrows, RY) * RS_WARPS >= 6LL * sm_count ? RY : 0;
}

// Strip height and grid of the streaming smoother for this handle; false when the grid is too small to fill the GPU
// with it (the caller then uses its tile / per-colour kernels).  A forced strip height (tests, tuning) always streams.
bool stream_plan(const cfd_b200_solver *s, int nx, int nrows, int *RY_out, int *gx_out, int *nstrips_out) {
    static const int ry_env = getenv("CFD_B200_RB_RY") ? atoi(getenv("CFD_B200_RB_RY")) : 0;
    const int forced = s->rb_strip > 0 ? s->rb_strip : ry_env;
    const int ry = stream_strip_rows(nx, nrows, s->sm_count, gx_out, nstrips_out);
    *RY_out = forced > 0 ? forced : (ry > 0 ? ry : 33);
    return forced > 0 || ry > 0;
}

// One smooth() (pressure.cpp:64-111), p -> p_alt.  variant 1: the streaming kernel with the guarded tile kernel behind it
// as its slow path (a no-op unless a non-finite value turned up); variant 2: the tile kernel alone.
void enqueue_smooth(cfd_b200_solver *s, double tol, int finish, unsigned long long wait_epoch = 0) {
    const WaitArgs hw = wait_args(s, wait_epoch); // row slabs: the neighbours' p / rhs rows this smooth() reads
    const Geom &g = s->g;
    const int ntx = cdiv(g.nx, RB_TX), nty = cdiv(g.ny, RB_TY);
    bool stream = s->variant == 1;
    if (stream) {
        const int nrows = (g.top ? g.ny + 1 : g.ny) - (g.bottom ? 0 : 1) + 1;
        int RY, gx, nstrips;
        stream = stream_plan(s, g.nx, nrows, &RY, &gx, &nstrips);
        const dim3 grid(gx, cdiv(nrows, RY));
        if (!stream) {
            // too small for the streaming kernel: the tile kernel below runs on its own
        } else if (g.denom_pow2)
            launch_kp(s, (s->pdl >> 1) & 1, k_rbgs_stream<1, 0>, grid, dim3(32 * RS_WARPS), RS_SMEM, g, s->p, s->p_alt, s->rhs, s->d_sc, s->partials, tol, RY, nstrips, finish, RsCoarse{}, hw);
        else
            launch_kp(s, (s->pdl >> 1) & 1, k_rbgs_stream<0, 0>, grid, dim3(32 * RS_WARPS), RS_SMEM, g, s->p, s->p_alt, s->rhs, s->d_sc, s->partials, tol, RY, nstrips, finish, RsCoarse{}, hw);
    }
    if (!stream) halo_wait(s, wait_epoch); // the tile kernel has no edge blocks of its own
    if (g.denom_pow2) launch_fused<1>(s, tol, ntx, nty, finish, stream);
    else launch_fused<0>(s, tol, ntx, nty, finish, stream);
    s->last_smoother = stream ? 1 : 2;
}

// ---- EXTENSION: true geometric V-cycle (kernels_vcycle.cuh; algorithm = the CPU restatement cfd_vcycle.c of the test tree; parity unpinned) ----
int vcycle_levels(int nx, int ny, int mg_levels) {
    int n = 1;
    while (n < mg_levels && n < MG_CTA_MAXLEV + 3 && nx % 2 == 0 && ny % 2 == 0 && nx / 2 >= 4 && ny / 2 >= 4) nx /= 2, ny /= 2, ++n;
    return n;
}
// The same V-cycle on row slabs.  Every level is split like the fine grid (its rows are the slab's rows halved l times, so
// a 2x2 block of children never straddles two slabs); the neighbours' row of p is exchanged after every kernel that changes
// p on a level (NCCL send/recv: the coarse levels live outside the IPC-mapped arena); the levels that fit into one block's
// shared memory are gathered onto EVERY slab (all-gather of the restricted right-hand side), solved there redundantly by
// the same one-block kernel, and each slab takes its rows (and the two rows next to them) of the result.  Every update
// sees the operands the single-GPU cycle sees, so p is bit-identical to it; only the residual norm is summed in another
// order.  One launch per colour pass on every level (the streaming sweeps of the single-GPU path are not slab-aware yet).
int enqueue_vcycle_slab(cfd_b200_solver *s, const cfd_b200_params *pp, double Lx, double Ly, bool in_step) {
    const Geom &g = s->g;
    NcclApi &nc = nccl_api();
    const int nlev = vcycle_levels(g.nx, g.ny_glob, pp->mg_levels), sweeps = pp->rbgs_sweeps;
    if (g.ny_glob % s->nranks != 0 || (g.ny & ((1 << (nlev - 1)) - 1)) != 0)
        return fail(CFD_B200_EUNSUPPORTED, "mg_mode=VCYCLE on %d slabs with %d levels needs ny divisible by %d (ny = %d)", s->nranks,
                    nlev, s->nranks << (nlev - 1), g.ny_glob);
    join_comm(s);
    for (int l = 0; l < nlev; ++l) {
        MgLevel &L = s->mg[l];
        const int nx = g.nx >> l, ny = g.ny >> l, nyg = g.ny_glob >> l;
        if (l > 0 && (L.nx != nx || L.ny != ny || !L.p)) {
            if (L.p) cudaFree(L.p), cudaFree(L.rhs), cudaFree(L.alt);
            L.P = ((nx + 3 + CFD_XOFF + 2) + 15) / 16 * 16;
            const size_t n = (size_t)L.P * (ny + 3 + 2 * CFD_YOFF);
            CU(cudaMalloc(&L.p, n * sizeof(double)));
            CU(cudaMalloc(&L.rhs, n * sizeof(double)));
            L.alt = nullptr;
            CU(cudaMemsetAsync(L.p, 0, n * sizeof(double), s->stream));
            CU(cudaMemsetAsync(L.rhs, 0, n * sizeof(double), s->stream));
        }
        L.nx = nx, L.ny = ny;
        const double dx = Lx / static_cast<double>(nx), dy = Ly / static_cast<double>(nyg);
        L.idx2 = 1.0 / (dx * dx);
        L.idy2 = 1.0 / (dy * dy);
        L.denom = 2.0 * (L.idx2 + L.idy2);
        if (l == 0) L.P = g.P, L.p = s->p, L.rhs = s->rhs, L.alt = s->p_alt;
        L.j0 = g.j0 >> l, L.bottom = g.bottom, L.top = g.top;
    }
    if (nlev > s->mg_nlev) s->mg_nlev = nlev;
    // first level that (with everything below it) fits into one block's shared memory -- sizes of the WHOLE level
    MgCtaPlan plan{};
    int lc = nlev;
    for (int l = 1; l < nlev; ++l) {
        size_t doubles = 0;
        for (int m = l; m < nlev; ++m) doubles += 2 * (size_t)((g.nx >> m) + 2) * ((g.ny_glob >> m) + 2);
        if (doubles * 8 <= 200 * 1024 && nlev - l <= MG_CTA_MAXLEV) {
            lc = l;
            break;
        }
    }
    int cta_smem = 0;
    MgLevel &G = s->mg_gather; // the whole level lc, on every slab
    if (lc < nlev) {
        plan.nlev = nlev - lc;
        plan.sweeps = sweeps;
        int off = 0;
        for (int m = lc; m < nlev; ++m) {
            const int k = m - lc, nx = g.nx >> m, nyg = g.ny_glob >> m;
            const double dx = Lx / static_cast<double>(nx), dy = Ly / static_cast<double>(nyg);
            plan.nx[k] = nx, plan.ny[k] = nyg, plan.off[k] = off;
            plan.idx2[k] = 1.0 / (dx * dx), plan.idy2[k] = 1.0 / (dy * dy), plan.denom[k] = 2.0 * (plan.idx2[k] + plan.idy2[k]);
            off += 2 * (nx + 2) * (nyg + 2);
        }
        cta_smem = off * 8;
        if (cta_smem > s->mg_cta_smem_max) {
            CU(cudaFuncSetAttribute(k_mg_coarse_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, cta_smem));
            s->mg_cta_smem_max = cta_smem;
        }
        const MgLevel &C = s->mg[lc];
        const int nyg = g.ny_glob >> lc;
        if (G.nx != C.nx || G.ny != nyg || !G.p) {
            if (G.p) cudaFree(G.p), cudaFree(G.rhs);
            const size_t n = (size_t)C.P * (nyg + 3 + 2 * CFD_YOFF);
            CU(cudaMalloc(&G.p, n * sizeof(double)));
            CU(cudaMalloc(&G.rhs, n * sizeof(double)));
            CU(cudaMemsetAsync(G.p, 0, n * sizeof(double), s->stream));
            CU(cudaMemsetAsync(G.rhs, 0, n * sizeof(double), s->stream));
        }
        G = MgLevel{C.nx, nyg, C.P, C.idx2, C.idy2, C.denom, G.p, G.rhs, nullptr, 0, 1, 1};
    }
    // the neighbours' row of L.p (1 row both ways)
    auto xchg = [&](const MgLevel &L) -> int {
        const size_t cnt = (size_t)L.P;
        auto row = [&](int jj) { return L.p + (size_t)(jj + CFD_YOFF) * L.P; };
        NCCLCHK(nc.GroupStart());
        if (!g.top) {
            NCCLCHK(nc.Send(row(L.ny), cnt, ncclDouble, s->rank + 1, s->comm, s->stream));
            NCCLCHK(nc.Recv(row(L.ny + 1), cnt, ncclDouble, s->rank + 1, s->comm, s->stream));
        }
        if (!g.bottom) {
            NCCLCHK(nc.Send(row(1), cnt, ncclDouble, s->rank - 1, s->comm, s->stream));
            NCCLCHK(nc.Recv(row(0), cnt, ncclDouble, s->rank - 1, s->comm, s->stream));
        }
        NCCLCHK(nc.GroupEnd());
        return 0;
    };
    auto smooth = [&](int l, int n) -> int {
        const MgLevel &L = s->mg[l];
        dim3 block(32, 8), cgrid(cdiv(cdiv(L.nx + 1, 2), 32), cdiv(L.ny, 8));
        for (int k = 0; k < n; ++k)
            for (int colour = 0; colour < 2; ++colour) {
                LAUNCH(s, k_mg_colour, cgrid, block, L, colour, s->d_sc);
                TRY(xchg(L));
            }
        return 0;
    };
    if (!in_step) LAUNCH(s, k_solve_begin, 1, 1, s->d_sc);
    LAUNCH(s, k_mg_zero_ghosts, cdiv((g.nx > g.ny ? g.nx : g.ny) + 2, 128), 128, s->mg[0], s->d_sc);
    TRY(xchg(s->mg[0]));
    const int top = lc < nlev ? lc : nlev - 1; // levels [0, top) run as global-memory kernels on the way down
    for (int vc = 0; vc < pp->vcycles; ++vc) {
        for (int l = 0; l < top; ++l) {
            TRY(smooth(l, sweeps));
            const MgLevel &F = s->mg[l], &C = s->mg[l + 1];
            LAUNCH(s, k_mg_restrict, dim3(cdiv(C.nx, 32), cdiv(C.ny, 8)), dim3(32, 8), F, C, s->d_sc);
            if (l + 1 < top || lc >= nlev) TRY(xchg(C)); // the zeroed guess (a level handed to the one-block kernel gets its rows back below)
        }
        if (lc < nlev) {
            const MgLevel &C = s->mg[lc];
            // my rows of the restricted right-hand side into the whole-level array, all-gather in place, the coarse part of
            // the V in one block, my rows of the result (+ the row on either side, for the prolongation) back
            const size_t cnt = (size_t)C.P * C.ny, r0 = (size_t)(1 + C.j0 + CFD_YOFF) * C.P;
            CU(cudaMemcpyAsync(G.rhs + r0, C.rhs + (size_t)(1 + CFD_YOFF) * C.P, cnt * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
            NCCLCHK(nc.AllGather(G.rhs + r0, G.rhs + (size_t)(1 + CFD_YOFF) * C.P, cnt, ncclDouble, s->comm, s->stream));
            launch_k(s, k_mg_coarse_cta, dim3(1), dim3(1024), (size_t)cta_smem, G, plan, s->d_sc);
            CU(cudaMemcpyAsync(C.p + (size_t)CFD_YOFF * C.P, G.p + (size_t)(C.j0 + CFD_YOFF) * C.P, (size_t)C.P * (C.ny + 2) * sizeof(double),
                               cudaMemcpyDeviceToDevice, s->stream));
        } else {
            TRY(smooth(nlev - 1, nlev == 1 ? sweeps : 16 * sweeps));
        }
        for (int l = top - 1; l >= 0; --l) {
            const MgLevel &F = s->mg[l];
            LAUNCH(s, k_mg_prolong, dim3(cdiv(F.nx, 32), cdiv(F.ny, 8)), dim3(32, 8), F, s->mg[l + 1], s->d_sc);
            TRY(xchg(F));
            TRY(smooth(l, sweeps));
        }
        const MgLevel &L0 = s->mg[0];
        dim3 rgrid(cdiv(L0.nx, 256), L0.ny < 256 ? L0.ny : 256);
        LAUNCH(s, k_mg_residual_norm, rgrid, 256, L0, s->d_sc, s->partials, pp->tol, 0);
        TRY(reduce_fin(s, RED_VC, pp->tol, 0));
    }
    if (!in_step) LAUNCH(s, k_solve_finish, 1, 1, s->d_sc);
    return 0;
}

int enqueue_vcycle(cfd_b200_solver *s, const cfd_b200_params *pp, double Lx, double Ly, bool in_step) {
    const Geom &g = s->g;
    if (s->nranks > 1) return enqueue_vcycle_slab(s, pp, Lx, Ly, in_step);
    const int nlev = vcycle_levels(g.nx, g.ny, pp->mg_levels);
    const int sweeps = pp->rbgs_sweeps;
    for (int l = 0; l < nlev; ++l) { // geometry of every level exactly as the restatement computes it
        MgLevel &L = s->mg[l];
        const int nx = g.nx >> l, ny = g.ny >> l;
        if (l > 0 && (L.nx != nx || L.ny != ny || !L.p)) {
            if (L.p) cudaFree(L.p), cudaFree(L.rhs), cudaFree(L.alt);
            L.P = ((nx + 3 + CFD_XOFF + 2) + 15) / 16 * 16;
            const size_t n = (size_t)L.P * (ny + 3 + 2 * CFD_YOFF);
            CU(cudaMalloc(&L.p, n * sizeof(double)));
            CU(cudaMalloc(&L.rhs, n * sizeof(double)));
            CU(cudaMalloc(&L.alt, n * sizeof(double))); // ping-pong partner of p for the streaming sweeps
            CU(cudaMemsetAsync(L.p, 0, n * sizeof(double), s->stream));
            CU(cudaMemsetAsync(L.rhs, 0, n * sizeof(double), s->stream));
            CU(cudaMemsetAsync(L.alt, 0, n * sizeof(double), s->stream));
        }
        L.nx = nx;
        L.ny = ny;
        const double dx = Lx / static_cast<double>(nx), dy = Ly / static_cast<double>(ny);
        L.idx2 = 1.0 / (dx * dx);
        L.idy2 = 1.0 / (dy * dy);
        L.denom = 2.0 * (L.idx2 + L.idy2);
        if (l == 0) L.P = g.P, L.p = s->p, L.rhs = s->rhs, L.alt = s->p_alt;
        L.j0 = 0, L.bottom = L.top = 1;
    }
    if (nlev > s->mg_nlev) s->mg_nlev = nlev;
    // first level that (with everything below it) fits into one block's shared memory
    MgCtaPlan plan{};
    int lc = nlev;
    for (int l = 1; l < nlev; ++l) {
        size_t doubles = 0;
        for (int m = l; m < nlev; ++m) doubles += 2 * (size_t)(s->mg[m].nx + 2) * (s->mg[m].ny + 2);
        if (doubles * 8 <= 200 * 1024 && nlev - l <= MG_CTA_MAXLEV) {
            lc = l;
            break;
        }
    }
    int cta_smem = 0;
    if (lc < nlev) {
        plan.nlev = nlev - lc;
        plan.sweeps = sweeps;
        int off = 0;
        for (int m = lc; m < nlev; ++m) {
            const MgLevel &L = s->mg[m];
            const int k = m - lc;
            plan.nx[k] = L.nx, plan.ny[k] = L.ny, plan.off[k] = off;
            plan.idx2[k] = L.idx2, plan.idy2[k] = L.idy2, plan.denom[k] = L.denom;
            off += 2 * (L.nx + 2) * (L.ny + 2);
        }
        cta_smem = off * 8;
        if (cta_smem > s->mg_cta_smem_max) {
            CU(cudaFuncSetAttribute(k_mg_coarse_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, cta_smem));
            s->mg_cta_smem_max = cta_smem;
        }
    }
    // n red-black sweeps on level l, followed by `then`: RS_VC_RESTRICT (residual restricted to level l+1, its p zeroed)
    // or RS_VC_NORM (residual norm of the level + the cycle loop's exit test).  Levels that fill the GPU take two sweeps
    // per launch with the streaming smoother (k_rbgs_stream<., 1>: the four colour passes of two sweeps in one pass over
    // the level, p ping-ponged) and the last such launch does `then` in its residual stage; an odd sweep goes first, one
    // launch per colour like every sweep of a small level, and `then` is a kernel of its own.
    auto smooth = [&](int l, int n, int then) {
        MgLevel &L = s->mg[l];
        dim3 block(32, 8), cgrid(cdiv(cdiv(L.nx + 1, 2), 32), cdiv(L.ny, 8));
        auto colour_sweeps = [&](int k) {
            for (; k > 0; --k)
                for (int colour = 0; colour < 2; ++colour) LAUNCH(s, k_mg_colour, cgrid, block, L, colour, s->d_sc);
        };
        int RY = 0, gx = 0, nstrips = 0;
        const bool stream = s->variant == 1 && n >= 2 && L.nx % 2 == 0 && L.ny % 2 == 0 &&
                            stream_plan(s, L.nx, L.ny + 2, &RY, &gx, &nstrips);
        if (stream) {
            colour_sweeps(n & 1);
            Geom lg{};
            lg.nx = L.nx, lg.ny = L.ny, lg.ny_glob = L.ny, lg.P = L.P, lg.j0 = 0, lg.bottom = lg.top = 1;
            lg.idx2 = L.idx2, lg.idy2 = L.idy2, lg.denom = L.denom, lg.inv_denom = 1.0 / L.denom;
            int e = 0;
            const bool pow2 = std::frexp(L.denom, &e) == 0.5;
            const int ry = RY - 1; // even: strips start at odd rows (see the kernel)
            const dim3 grid(gx, cdiv(L.ny + 1, ry));
            RsCoarse cz{};
            if (then == RS_VC_RESTRICT) cz = RsCoarse{s->mg[l + 1].rhs, s->mg[l + 1].p, s->mg[l + 1].P};
            for (int k = n / 2; k > 0; --k) {
                const int fin = k == 1 ? then : RS_VC_SMOOTH;
                if (pow2)
                    launch_kp(s, (s->pdl >> 1) & 1, k_rbgs_stream<1, 1>, grid, dim3(32 * RS_WARPS), RS_SMEM, lg, L.p, L.alt, L.rhs, s->d_sc, s->partials, pp->tol, ry, nstrips, fin, cz, WaitArgs{});
                else
                    launch_kp(s, (s->pdl >> 1) & 1, k_rbgs_stream<0, 1>, grid, dim3(32 * RS_WARPS), RS_SMEM, lg, L.p, L.alt, L.rhs, s->d_sc, s->partials, pp->tol, ry, nstrips, fin, cz, WaitArgs{});
                double *t = L.p;
                L.p = L.alt, L.alt = t;
                if (l == 0) s->p = L.p, s->p_alt = L.alt; // level 0 aliases the State's pressure
            }
        } else {
            colour_sweeps(n);
        }
        if (stream || then == RS_VC_SMOOTH) return;
        if (then == RS_VC_RESTRICT) {
            const MgLevel &C = s->mg[l + 1];
            LAUNCH(s, k_mg_restrict, dim3(cdiv(C.nx, 32), cdiv(C.ny, 8)), dim3(32, 8), L, C, s->d_sc);
        } else {
            dim3 rgrid(cdiv(L.nx, 256), L.ny < 256 ? L.ny : 256);
            LAUNCH(s, k_mg_residual_norm, rgrid, 256, L, s->d_sc, s->partials, pp->tol, 1);
        }
    };
    if (!in_step) LAUNCH(s, k_solve_begin, 1, 1, s->d_sc);
    LAUNCH(s, k_mg_zero_ghosts, cdiv((g.nx > g.ny ? g.nx : g.ny) + 2, 128), 128, s->mg[0], s->d_sc);
    const int top = lc < nlev ? lc : nlev - 1; // levels [0, top) run as global-memory kernels on the way down
    for (int vc = 0; vc < pp->vcycles; ++vc) {
        for (int l = 0; l < top; ++l) smooth(l, sweeps, RS_VC_RESTRICT);
        if (lc < nlev) {
            launch_k(s, k_mg_coarse_cta, dim3(1), dim3(1024), (size_t)cta_smem, s->mg[lc], plan, s->d_sc);
        } else {
            smooth(nlev - 1, nlev == 1 ? sweeps : 16 * sweeps, top == 0 ? RS_VC_NORM : RS_VC_SMOOTH);
        }
        for (int l = top - 1; l >= 0; --l) {
            const MgLevel &F = s->mg[l];
            LAUNCH(s, k_mg_prolong, dim3(cdiv(F.nx, 32), cdiv(F.ny, 8)), dim3(32, 8), F, s->mg[l + 1], s->d_sc);
            smooth(l, sweeps, l == 0 ? RS_VC_NORM : RS_VC_SMOOTH);
        }
    }
    if (!in_step) LAUNCH(s, k_solve_finish, 1, 1, s->d_sc);
    return 0;
}

// Can this handle run the cooperative persistent PCG kernel (working set in shared memory, all blocks co-resident)?
bool coop_probe(cfd_b200_solver *s) {
    if (s->coop_ok >= 0) return s->coop_ok != 0;
    s->coop_ok = 0;
    static const int env = getenv("CFD_B200_PCG_COOP") ? atoi(getenv("CFD_B200_PCG_COOP")) : 1;
    const Geom &g = s->g;
    int coop = 0;
    if (!env || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device) != cudaSuccess || !coop) return false;
    const int nblk = s->sm_count, R = cdiv(g.ny, nblk);
    const size_t smem = (size_t)(3 * R + 2) * (g.nx + 2) * sizeof(double);
    if (smem > 200 * 1024) return false;
    if (cudaFuncSetAttribute(k_pcg_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_coop, PC_THREADS, smem) != cudaSuccess || per_sm < 1) return false;
    if (cudaMalloc(&s->coop_xrows, (size_t)2 * nblk * 2 * (g.nx + 2) * sizeof(double)) != cudaSuccess) return false;
    if (cudaMalloc(&s->coop_slots, (size_t)3 * 2 * nblk * sizeof(double)) != cudaSuccess) return false;
    cudaMemsetAsync(s->coop_xrows, 0, (size_t)2 * nblk * 2 * (g.nx + 2) * sizeof(double), s->stream);
    cudaMemsetAsync(s->coop_slots, 0, (size_t)3 * 2 * nblk * sizeof(double), s->stream);
    s->coop_rows = R, s->coop_blocks = nblk, s->coop_smem = (int)smem;
    s->coop_ok = 1;
    // Grids of at most 512 x 256 cells: the whole solve inside ONE cluster of 16 CTAs (kernels_pcg_cluster.cuh), with the
    // cooperative kernel behind it as its guarded slow path.  Probed like a launch: opt-in cluster size, shared memory, occupancy.
    static const int cl_env = getenv("CFD_B200_PCG_CLUSTER") ? atoi(getenv("CFD_B200_PCG_CLUSTER")) : 1;
    if (cl_env && g.nx <= PK_COLS && g.ny <= PK_CTAS * PK_ROWS) {
        const int Rc = cdiv(g.ny, PK_CTAS);
        const size_t smem_c = (size_t)(3 * Rc + 4) * (g.nx + 2) * sizeof(double);
        if (cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c) == cudaSuccess &&
            cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(PK_CTAS), cfg.blockDim = dim3(PK_THREADS), cfg.dynamicSmemBytes = smem_c;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = PK_CTAS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
            cfg.attrs = at, cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_pcg_cluster, &cfg) == cudaSuccess && ncl >= 1)
                s->cluster_ok = 1, s->cluster_rows = Rc, s->cluster_smem = (int)smem_c;
        }
        cudaGetLastError(); // a refused attribute leaves the cooperative kernel in charge
    }
    return true;
}

// PressureSolver::solve, pressure.cpp:113-242.  The whole solve is enqueued; convergence is tested on the device.
// Row slabs: every grid-level sum is completed by an allreduce + a one-thread tail kernel (identical on all ranks),
// and the rows a neighbour's stencil reads are exchanged where the single-GPU path would simply read them.
// in_step: k_dt has already done k_solve_begin's resets and k_project will do k_solve_finish's check (two launches less)
int enqueue_solve(cfd_b200_solver *s, const cfd_b200_params *pp, bool in_step = false) {
    const Geom &g = s->g;
    const int single = s->nranks == 1;
    if (pp->solver == CFD_B200_SOLVER_PCG) {
        cudaMemsetAsync(s->p, 0, s->n * sizeof(double), s->stream); // pressure.cpp:133-134
        if (single && s->variant != 0 && coop_probe(s)) { // small grid: the whole solve in one persistent kernel
            if (s->cluster_ok) {
                PcgClusterArgs ca{g, s->rhs, s->p, s->d_sc, pp->tol, pp->pcg_max_iters, s->cluster_rows};
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(PK_CTAS), cfg.blockDim = dim3(PK_THREADS), cfg.dynamicSmemBytes = (size_t)s->cluster_smem;
                cfg.stream = s->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = PK_CTAS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
                cfg.attrs = at, cfg.numAttrs = 1;
                if (cudaLaunchKernelEx(&cfg, k_pcg_cluster, ca) == cudaSuccess) {
                    s->launches++;
                } else { // a cluster of this size cannot be placed after all: the cooperative kernel alone, from now on
                    cudaGetLastError();
                    s->cluster_ok = 0;
                }
            }
            // alone, or as the guarded slow path of k_pcg_cluster (returns at once unless that kernel met a non-finite value)
            PcgCoopArgs a{g, s->rhs, s->p, s->coop_xrows, s->coop_slots, s->d_sc, pp->tol, pp->pcg_max_iters, s->coop_rows, s->cluster_ok};
            void *args[] = {&a};
            CU(cudaLaunchCooperativeKernel((void *)k_pcg_coop, dim3(s->coop_blocks), dim3(PC_THREADS), args,
                                           (size_t)s->coop_smem, s->stream));
            s->launches++;
            s->last_pcg = s->cluster_ok ? 4 : 1;
            if (!in_step) LAUNCH(s, k_solve_finish, 1, 1, s->d_sc);
            return 0;
        }
        {
            dim3 grid(cdiv(g.nx, 512), g.ny < 256 ? g.ny : 256);
            LAUNCH(s, k_pcg_init, grid, PT_THREADS, g, s->rhs, s->r, s->d_sc, s->partials, pp->tol, pp->pcg_max_iters, single);
        }
        if (!single) {
            TRY(reduce_fin(s, RED_PCG_INIT, pp->tol, pp->pcg_max_iters, s->r, 1, HALO_BOTH));
        }
        const int ntx = cdiv(g.nx, PT_X), nty = cdiv(g.ny, PT_Y), ntiles = ntx * nty;
        const int grid = pcg_grid(s, ntiles);
        double *s_old = s->sA, *s_new = s->sB;
        // Streaming kernels for grids that fill the GPU, tile kernels below that (B200: 512x256 tile 3.1 ms vs streaming
        // 4.1 ms per step; 4096^2 tile 49.9 vs streaming 46.4; 8192^2 187 vs 162).  Rows per strip: a power of two
        // (4096^2: RY 32/58/64/128 -> 46.4/52.2/47.0/52.2 ms; 8192^2: 64/128/256 -> 164.7/161.6/174.7) that leaves
        // >= 8 blocks per SM in the grid.
        static const int ry_env = getenv("CFD_B200_PCG_RY") ? atoi(getenv("CFD_B200_PCG_RY")) : 0;
        const int gxm = cdiv(g.nx / PM_COLS + 1, PM_WARPS);
        int RY = 8;
        while (RY < 128 && (long long)g.ny * gxm / (2 * RY) >= 8LL * s->sm_count) RY *= 2;
        if (ry_env > 0) RY = ry_env;
        const dim3 mgrid(gxm, cdiv(g.ny, RY));
        const bool march = s->variant != 0 && (ry_env > 0 || (long long)g.nx * g.ny >= (2LL << 20));
        s->last_pcg = march ? 2 : 3;
        for (int it = 0; it < pp->pcg_max_iters; ++it) {
            if (march)
                LAUNCH(s, k_pcg_phase1_march, mgrid, 32 * PM_WARPS, g, s->r, s_old, s_new, s->d_sc, s->partials, RY, single);
            else
                LAUNCH(s, k_pcg_phase1, grid, PT_THREADS, g, s->r, s_old, s_new, s->d_sc, s->partials, ntx, ntiles, single);
            if (!single) TRY(reduce_fin(s, RED_PCG1, 0.0, 0));
            if (march)
                LAUNCH(s, k_pcg_phase2_march, mgrid, 32 * PM_WARPS, g, s->p, s->r, s_new, s->d_sc, s->partials, RY, pp->tol,
                       pp->pcg_max_iters, single);
            else
                LAUNCH(s, k_pcg_phase2, grid, PT_THREADS, g, s->p, s->r, s_new, s->d_sc, s->partials, ntx, ntiles, pp->tol,
                       pp->pcg_max_iters, single);
            if (!single) {
                // + r, 1 row both ways: the next direction update reads r one row out
                TRY(reduce_fin(s, RED_PCG2, pp->tol, pp->pcg_max_iters, s->r, 1, HALO_BOTH));
            }
            double *t = s_old;
            s_old = s_new;
            s_new = t;
        }
    } else {
        if (pp->mg_mode == CFD_B200_MG_VCYCLE) return enqueue_vcycle(s, pp, s->Lx, s->Ly, in_step);
        if (pp->mg_mode != CFD_B200_MG_REFERENCE) return fail(CFD_B200_EINVAL, "bad mg_mode %d", pp->mg_mode);
        if (!in_step) LAUNCH(s, k_solve_begin, 1, 1, s->d_sc);
        if (s->variant != 0) { // fused smooth(): 4 colour passes + residual per launch, p ping-ponged
            for (int vc = 0; vc < pp->vcycles; ++vc) {
                unsigned long long ep = 0;
                // 5 rows of p and, once, of rhs (it is constant during the solve) for the redundant compute.  One-way sends:
                // the reduction of the previous smooth() (or of the previous step) is the barrier in front of them; the p
                // rows of the first smooth() of a step usually arrived at the end of the previous step already.
                if (!single) {
                    double *f[2];
                    int n = 0;
                    if (vc > 0 || !s->p_halo_valid) f[n++] = s->p;
                    if (vc == 0) f[n++] = s->rhs;
                    const int rows[2] = {RB_HY, RB_HY}, dirs[2] = {HALO_BOTH, HALO_BOTH};
                    if (n) TRY(halo_send(s, f, rows, dirs, n, &ep));
                    s->p_halo_valid = false;
                }
                enqueue_smooth(s, pp->tol, single, ep); // its edge strips wait for those rows
                // the residual sum.  Peer-memory transport, inside a step: the last one rides on the reduction that ends the
                // step (one barrier less: nothing in between depends on it -- the vcycle loop is over either way)
                if (!single && in_step && s->p2p_on && vc == pp->vcycles - 1) s->defer_rbgs = true, s->defer_tol = pp->tol;
                else if (!single) TRY(reduce_fin(s, RED_RBGS, pp->tol, 0));
                double *t = s->p;
                s->p = s->p_alt;
                s->p_alt = t;
            }
            if (!in_step) LAUNCH(s, k_solve_finish, 1, 1, s->d_sc);
            return 0;
        }
        dim3 cblock(32, 8), cgrid(cdiv(cdiv(g.nx, 2), 32), cdiv(g.ny, 8));
        dim3 rgrid(cdiv(g.nx, 256), g.ny < 256 ? g.ny : 256);
        s->last_smoother = 3;
        for (int vc = 0; vc < pp->vcycles; ++vc) {
            for (int sweep = 0; sweep < 2; ++sweep) // the literal 2 of pressure.cpp:119
                for (int colour = 0; colour < 2; ++colour) {
                    if (!single) TRY(halo_exchange2(s, s->p, nullptr, 1, HALO_BOTH));
                    if (g.denom_pow2)
                        LAUNCH(s, k_rbgs_colour<1>, cgrid, cblock, g, s->p, s->rhs, colour, s->d_sc);
                    else
                        LAUNCH(s, k_rbgs_colour<0>, cgrid, cblock, g, s->p, s->rhs, colour, s->d_sc);
                }
            if (!single) TRY(halo_exchange2(s, s->p, nullptr, 1, HALO_BOTH));
            LAUNCH(s, k_rbgs_residual, rgrid, 256, g, s->p, s->rhs, s->d_sc, s->partials, pp->tol, single);
            if (!single) TRY(reduce_fin(s, RED_RBGS, pp->tol, 0));
        }
    }
    if (!in_step) LAUNCH(s, k_solve_finish, 1, 1, s->d_sc);
    return 0;
}

void enqueue_project(cfd_b200_solver *s, const Scalars *dt_src, double dt_arg, int pin = 1, int fin_solve = 0) {
    const Geom &g = s->g;
    int rpb = 4;
    dim3 grid = row_grid(s, cdiv(g.nx + 1, 2), g.ny + 1, &rpb);
    LAUNCH(s, k_project, grid, 256, g, s->u, s->v, s->p, dt_src, dt_arg, s->d_sc, rpb, pin, fin_solve);
}

// TimeIntegrator::step, time_integrator.cpp:47-239
int enqueue_step(cfd_b200_solver *s, const cfd_b200_bc *bc_in, double Re, double CFL, const cfd_b200_params *pp,
                 double dt_override) {
    const Geom &g = s->g;
    const BcD bc = to_bcd(bc_in);
    const double invRe = 1.0 / Re; // "1.0 / Re" at time_integrator.cpp:76
    if (!bc_equal(s->last_bc, *bc_in) || s->variant == 0) s->cfl_cached = false;
    s->last_bc = *bc_in;
    s->defer_rbgs = false;
    stage_mark(s, ST_BC0);
    // the advection stencil reads 2 rows beyond the slab: sent by the neighbours at the end of the previous step (the
    // boundary pass below also works on them: it is row-local and reproduces what the neighbour does to its own rows),
    // or exchanged here when the State was uploaded since
    if (s->halo_uv_valid) TRY(halo_wait(s, s->ep_uv));
    s->ep_uv = 0;
    if (!bc_equal(s->last_bc_step, *bc_in)) s->p_halo_valid = false;
    s->last_bc_step = *bc_in;
    enqueue_bc(s, bc, s->u, s->v, s->p, 0); // :51-53
    if (!s->halo_uv_valid) TRY(halo_exchange2(s, s->u, s->v, 2, HALO_BOTH));
    stage_mark(s, ST_CFL);
    enqueue_cfl(s, Re, CFL, dt_override, 0); // :55-66
    stage_mark(s, ST_RHS1);
    enqueue_rhs<1>(s, invRe, s->u, s->v, s->u1, s->v1); // :71-103
    stage_mark(s, ST_BC1);
    enqueue_bc(s, bc, s->u1, s->v1, nullptr, 1, 0); // :106-107 (own rows: the halo rows arrive with the pass applied)
    unsigned long long ep = 0;
    TRY(halo_send2(s, s->u1, s->v1, 2, HALO_BOTH, &ep)); // one-way; the edge strips of stage 2 wait for the neighbours' rows
    stage_mark(s, ST_RHS2);
    // variant 1: stage 2 leaves the unprojected velocity in ustar / vstar and the projection writes the State's u, v from
    // them, out of place, together with the divergence of the result (k_project_div); otherwise everything is in place.
    const bool fused_div = s->variant == 1;
    double *us = fused_div ? s->ustar : s->u, *vs = fused_div ? s->vstar : s->v;
    // :109-147; variant 1 also takes the divergence of u*, v* (:158-182) where the faces are final (k_rhs_march<4>)
    static const int fuse_pre = getenv("CFD_B200_FUSE_PRE") ? atoi(getenv("CFD_B200_FUSE_PRE")) : 1;
    MarchRegion mr{};
    if (fused_div && fuse_pre) mr = enqueue_rhs<4>(s, invRe, s->u1, s->v1, us, vs, s->u, s->v, ep);
    else mr = enqueue_rhs<2>(s, invRe, s->u1, s->v1, us, vs, s->u, s->v, ep), mr.marched = false;
    stage_mark(s, ST_BC2);
    enqueue_bc(s, bc, us, vs, nullptr, 1, 0); // :150-155
    TRY(halo_send2(s, vs, nullptr, 1, HALO_DOWN, &ep)); // the divergence of the top cell row reads the v face above it
    stage_mark(s, ST_DIV);
    if (mr.marched) { // the cells the marching kernel could not close; the blocks of the top cell row wait for the neighbour's row
        const long long nfix = (long long)g.ny * ((mr.XA - 1) + (g.nx - mr.XB + 1)) +
                               (long long)((mr.YA - 1) + (mr.YB - mr.YA) / mr.RY + (g.ny - mr.YB + 1)) * g.nx;
        const long long nblk = (nfix + 255) / 256, cap = 8LL * s->sm_count;
        LAUNCH(s, k_div_fix_pre, (unsigned)(nblk < cap ? nblk : cap), 256, g, us, vs, s->rhs, s->p, s->d_sc, s->partials, mr.XA, mr.XB,
               mr.YA, mr.YB, mr.RY, wait_args(s, ep));
    } else {
        enqueue_divergence(s, 1, 1, 0, us, vs, ep); // :158-186; the blocks of the top cell row wait for that row
    }
    stage_mark(s, ST_SOLVE);
    int rc = enqueue_solve(s, pp, true); // :187
    if (rc) return rc;
    // v faces of the bottom row read p one row below the slab; the fused projection also projects the v row above the slab
    // (the neighbour's faces, redundantly) to close its top cell row and needs p one row above it.  One-way, behind the
    // solver's last reduction; the chunks of rows at the slab's ends wait for the neighbours' rows.
    TRY(halo_send2(s, s->p, nullptr, 1, fused_div ? HALO_BOTH : HALO_UP, &ep));
    stage_mark(s, ST_PROJECT);
    // :188-202 (+ last_residual_ = isfinite(res) ? res : 1e9)
    if (fused_div) {
        int rpb = 4;
        const dim3 grid = row_grid(s, cdiv(g.nx + 1, 2), g.ny + 1, &rpb);
        LAUNCH(s, k_project_div, grid, 256, g, us, vs, s->p, s->u, s->v, s->rhs, s->d_sc, s->partials, rpb, wait_args(s, ep));
    } else {
        TRY(halo_wait(s, ep));
        enqueue_project(s, s->d_sc, 0.0, 1, 1);
    }
    stage_mark(s, ST_BC3);
    enqueue_bc(s, bc, s->u, s->v, s->p, 1); // :205-207
    if (!fused_div) TRY(halo_exchange2(s, s->v, nullptr, 1, HALO_DOWN)); // the v row above the slab for the divergence
    stage_mark(s, ST_DIV2);
    if (fused_div) { // :210-220 on the ring (+ its part of max|u|, max|v| for the next step's compute_cfl)
        const int nfix = ((g.bottom ? 1 : 0) + (g.top ? 1 : 0)) * g.nx + 2 * g.ny;
        LAUNCH(s, k_div_fix, cdiv(nfix, 256), 256, g, s->u, s->v, s->rhs, s->d_sc, s->partials);
    } else {
        enqueue_divergence(s, 0, 1, 1);
    }
    s->cfl_cached = true;
    TRY(reduce_fin(s, RED_STEP_END, 0.0, 0)); // CFL maxima of the next step + the two partial sums of squares
    if (s->nranks > 1) { // behind that barrier: the rows the neighbours read in the next step (nobody waits for them here)
        double *f[3] = {s->u, s->v, s->p};
        const int rows[3] = {2, 2, RB_HY}, dirs[3] = {HALO_BOTH, HALO_BOTH, HALO_BOTH};
        const bool with_p = pp->solver == CFD_B200_SOLVER_MG && pp->mg_mode == CFD_B200_MG_REFERENCE && s->variant != 0;
        TRY(halo_send(s, f, rows, dirs, with_p ? 3 : 2, &s->ep_uv));
        s->halo_uv_valid = true;
        s->p_halo_valid = with_p;
    }
    LAUNCH(s, k_sanitize_if, s->sm_count * 4, 256, g, s->p, s->d_sc); // :236
    stage_mark(s, ST_END);
    s->ev_valid = s->profiling != 0;
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaGetLastError());
    return 0;
}

void fill_info(const cfd_b200_solver *s, cfd_b200_step_info *info) {
    if (!info) return;
    const Scalars &h = *s->h_sc;
    info->dt = h.dt;
    info->residual = h.res;
    info->iters = h.iters;
    info->pcg_breakdown = h.breakdown_iter;
    info->div_before = h.div_before;
    info->div_after = h.div_after;
    if (s->nranks > 1) { // the kernels left global sums of squares (time_integrator.cpp:171, :220 finish them)
        const double ncell = (double)(s->g.nx * s->g.ny_glob);
        info->div_before = std::sqrt(h.div_before / ncell);
        info->div_after = std::sqrt(h.div_after / ncell);
    }
    info->nonfinite = (h.nonfinite_any ? 1 : 0) | (h.nonfinite_res ? 2 : 0);
}

// The one-way send that ends a step may still be arriving from the neighbours: wait for it before anything outside a step
// touches or invalidates the halo rows.
int settle_halos(cfd_b200_solver *s) {
    join_comm(s);
    if (s->ep_uv) TRY(halo_wait(s, s->ep_uv));
    s->ep_uv = 0;
    return 0;
}

// host (reference layout) <-> device (padded layout) copy of one field
int copy_field(cfd_b200_solver *s, double *dev, double *host, int host_pitch, int host_rows, bool to_device) {
    const Geom &g = s->g;
    double *d0 = dev + gidx(g, 0, 0);
    if (to_device)
        CU(cudaMemcpy2DAsync(d0, (size_t)g.P * 8, host, (size_t)host_pitch * 8, (size_t)host_pitch * 8, host_rows,
                             cudaMemcpyHostToDevice, s->stream));
    else
        CU(cudaMemcpy2DAsync(host, (size_t)host_pitch * 8, d0, (size_t)g.P * 8, (size_t)host_pitch * 8, host_rows,
                             cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int need_scratch(cfd_b200_solver *s) {
    if (!s->scratch_u) {
        int rc = alloc_field(s, &s->scratch_u);
        if (rc) return rc;
        rc = alloc_field(s, &s->scratch_v);
        if (rc) return rc;
    }
    return 0;
}

} // namespace

extern "C" {

// Pinned host memory for the caller-owned State (Field2D::allocate of the C++ layer): host<->device copies of the
// drop-in then run at full PCIe speed.  Falls back to aligned_alloc when no CUDA runtime/device is reachable.
void *cfd_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 64;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p; // page aligned
    cudaGetLastError();
    p = aligned_alloc(64, (bytes + 63) / 64 * 64);
    return p;
}
void cfd_b200_host_free(void *p) {
    if (!p) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost) {
        cudaFreeHost(p);
        return;
    }
    cudaGetLastError();
    free(p);
}

int cfd_b200_abi_version(void) { return CFD_B200_ABI_VERSION; }
const char *cfd_b200_last_error(void) { return g_err; }

int cfd_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(CFD_B200_ENODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int cfd_b200_comm_id(void *id128) {
    if (!id128) return fail(CFD_B200_EINVAL, "id128 is NULL");
    NcclApi &n = nccl_api();
    if (!n.load()) return fail(CFD_B200_ECOMM, "NCCL unavailable: %s", n.error);
    ncclUniqueId id;
    ncclResult_t r = n.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(CFD_B200_ECOMM, "ncclGetUniqueId: %s", n.GetErrorString(r));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int cfd_b200_create_slab(int nx, int ny, double Lx, double Ly, int device, int rank, int nranks, const void *nccl_id128,
                         cfd_b200_solver **out) {
    if (!out) return fail(CFD_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (nx < 4 || ny < 4 || !(Lx > 0.0) || !(Ly > 0.0)) return fail(CFD_B200_EINVAL, "need nx, ny >= 4 and Lx, Ly > 0");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(CFD_B200_EINVAL, "bad rank %d of %d", rank, nranks);
    if (nranks > 1 && !nccl_id128) return fail(CFD_B200_EINVAL, "nccl_id128 is NULL");
    if (nranks > 1 && ny / nranks < 8) return fail(CFD_B200_EINVAL, "need at least 8 cell rows per slab");
    if (nranks > P2P_MAXR) return fail(CFD_B200_EINVAL, "at most %d row slabs", P2P_MAXR);
    if (nranks > 1 && !nccl_api().load()) return fail(CFD_B200_ECOMM, "NCCL unavailable: %s", nccl_api().error);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(CFD_B200_ENODEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(CFD_B200_EINVAL, "device %d out of range [0,%d)", device, ndev);
    CU(cudaSetDevice(device));
    cfd_b200_solver *s = new (std::nothrow) cfd_b200_solver;
    if (!s) return fail(CFD_B200_ENOMEM, "host allocation failed");
    s->device = device;
    s->Lx = Lx, s->Ly = Ly;
    s->rank = rank;
    s->nranks = nranks;
    s->pdl = getenv("CFD_B200_PDL") ? atoi(getenv("CFD_B200_PDL")) : 3;
    // bit 0: all kernels, bit 1: k_rbgs_stream, bit 2: its slow path.  Bit 2 stays off: measured on B200 at 8192^2, the
    // (normally idle) slow-path blocks become resident while the streaming kernel is still running and take two 64 KB
    // shared-memory slots per SM from it: 3.49 ms per step against 3.27.
    int row0 = 0, nrows = ny;
    slab_partition(ny, rank, nranks, &row0, &nrows);
    fill_geom(s->g, nx, ny, Lx, Ly, row0, nrows, rank == 0, rank == nranks - 1);
    s->n = (size_t)s->g.P * s->g.rows;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(CFD_B200_ECUDA, "stream create failed"); break; }
        if (nranks > 1 && (cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
                           cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                           cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess)) { rc = fail(CFD_B200_ECUDA, "communication stream create failed"); break; }
        if (!(getenv("CFD_B200_AUX") && atoi(getenv("CFD_B200_AUX")) == 0)) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi); // hi = greatest priority: the strips' small blocks go first
            if (cudaStreamCreateWithPriority(&s->aux_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
                cudaEventCreateWithFlags(&s->ev_afork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&s->ev_ajoin, cudaEventDisableTiming) != cudaSuccess) { rc = fail(CFD_B200_ECUDA, "auxiliary stream create failed"); break; }
        }
        {   // one arena (a multiple of 2 MiB: its own mapping, one IPC handle): NFIELDS fields + this rank's P2PBlock
            size_t bytes = NFIELDS * s->n * sizeof(double) + sizeof(P2PBlock);
            bytes = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
            cudaError_t ea = cudaMalloc(&s->arena, bytes);
            if (ea != cudaSuccess) { rc = fail(ea == cudaErrorMemoryAllocation ? CFD_B200_ENOMEM : CFD_B200_ECUDA, "arena of %zu bytes: %s", bytes, cudaGetErrorString(ea)); break; }
            if (cudaMemsetAsync(s->arena, 0, bytes, s->stream) != cudaSuccess) { rc = fail(CFD_B200_ECUDA, "arena memset failed"); break; }
            double **fields[NFIELDS] = {&s->u, &s->v, &s->p, &s->p_alt, &s->rhs, &s->u1, &s->v1, &s->r, &s->sA, &s->sB, &s->ustar, &s->vstar};
            for (int k = 0; k < NFIELDS; ++k) *fields[k] = s->arena + (size_t)k * s->n;
            s->blk[rank] = reinterpret_cast<P2PBlock *>(s->arena + (size_t)NFIELDS * s->n);
        }
        if (cudaMalloc(&s->d_sc, sizeof(Scalars)) != cudaSuccess) { rc = fail(CFD_B200_ENOMEM, "scalars alloc failed"); break; }
        cudaMemsetAsync(s->d_sc, 0, sizeof(Scalars), s->stream);
        {   // bound of every wait for a peer GPU (row slabs): a rank that falls this far behind is reported as CFD_B200_ECOMM
            const char *e = getenv("CFD_B200_P2P_TIMEOUT_S");
            const double sec = e && atof(e) > 0.0 ? atof(e) : 20.0;
            s->p2p_timeout_ns = (unsigned long long)(sec * 1e9);
            cudaMemcpyAsync(&s->d_sc->p2p_timeout_ns, &s->p2p_timeout_ns, sizeof(unsigned long long), cudaMemcpyHostToDevice, s->stream);
        }
        if (cudaMallocHost(&s->h_sc, sizeof(Scalars)) != cudaSuccess) { rc = fail(CFD_B200_ENOMEM, "pinned alloc failed"); break; }
        memset(s->h_sc, 0, sizeof(Scalars));
        s->partials_n = 2 * 65536;
        {
            size_t tiles = (size_t)cdiv(nx, RB_TX) * cdiv(nrows, RB_TY);
            if (2 * tiles > s->partials_n) s->partials_n = 2 * tiles;
        }
        if (cudaFuncSetAttribute(k_rbgs_fused<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_rbgs_fused<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_rbgs_fused<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_rbgs_fused<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess) {
            rc = fail(CFD_B200_ECUDA, "cudaFuncSetAttribute(k_rbgs_fused) failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        {
            auto prep = [](auto kernel) {
                return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
                       cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM) == cudaSuccess;
            };
            if (!prep(k_rbgs_stream<0, 0>) || !prep(k_rbgs_stream<1, 0>) || !prep(k_rbgs_stream<0, 1>) || !prep(k_rbgs_stream<1, 1>)) {
                rc = fail(CFD_B200_ECUDA, "cudaFuncSetAttribute(k_rbgs_stream) failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
        {   // k_rhs_march2: 4 blocks per SM of 16 / 33 KB of (static) shared memory each
            auto prep = [](auto kernel) {
                return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess;
            };
            if (!prep(k_rhs_march2<0>) || !prep(k_rhs_march2<1>) || !prep(k_rhs_march2<2>) || !prep(k_rhs_march2<4>)) {
                rc = fail(CFD_B200_ECUDA, "cudaFuncSetAttribute(k_rhs_march2) failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
        if (cudaMalloc(&s->partials, s->partials_n * sizeof(double)) != cudaSuccess) { rc = fail(CFD_B200_ENOMEM, "partials alloc failed"); break; }
        for (int i = 0; i < ST_COUNT; ++i)
            if (cudaEventCreate(&s->ev[i]) != cudaSuccess) { rc = fail(CFD_B200_ECUDA, "event create failed"); break; }
        if (rc) break;
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) { rc = fail(CFD_B200_ECUDA, "init sync failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (nranks > 1) { // collective: every rank of the job must be here
            ncclUniqueId id;
            memcpy(&id, nccl_id128, sizeof(id));
            ncclResult_t r = nccl_api().CommInitRank(&s->comm, nranks, id, rank);
            if (r != ncclSuccess) { rc = fail(CFD_B200_ECOMM, "ncclCommInitRank: %s", nccl_api().GetErrorString(r)); break; }
            if (cudaMalloc(&s->gathered, 4 * nranks * sizeof(double)) != cudaSuccess) { rc = fail(CFD_B200_ENOMEM, "gather buffer"); break; }
            if (cudaMalloc(&s->bt_far, 5 * (size_t)s->g.P * sizeof(double)) != cudaSuccess) { rc = fail(CFD_B200_ENOMEM, "periodic rows buffer"); break; }
            if ((rc = p2p_setup(s)) != 0) break;
        }
    } while (0);
    if (rc) {
        cfd_b200_destroy(s);
        return rc;
    }
    *out = s;
    return 0;
}

int cfd_b200_create(int nx, int ny, double Lx, double Ly, int device, cfd_b200_solver **out) {
    return cfd_b200_create_slab(nx, ny, Lx, Ly, device, 0, 1, nullptr, out);
}

int cfd_b200_slab_rows(const cfd_b200_solver *s, int *row0, int *nrows) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    if (row0) *row0 = s->g.j0;
    if (nrows) *nrows = s->g.ny;
    return 0;
}

void cfd_b200_destroy(cfd_b200_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) settle_halos(s), cudaStreamSynchronize(s->stream);
    p2p_teardown(s);
    if (s->comm) nccl_api().CommDestroy(s->comm);
    if (s->gathered) cudaFree(s->gathered);
    if (s->bt_far) cudaFree(s->bt_far);
    if (s->coop_xrows) cudaFree(s->coop_xrows);
    if (s->coop_slots) cudaFree(s->coop_slots);
    for (int l = 1; l < s->mg_nlev; ++l) {
        if (s->mg[l].p) cudaFree(s->mg[l].p);
        if (s->mg[l].rhs) cudaFree(s->mg[l].rhs);
        if (s->mg[l].alt) cudaFree(s->mg[l].alt);
    }
    if (s->mg_gather.p) cudaFree(s->mg_gather.p), cudaFree(s->mg_gather.rhs);
    if (s->viz_rgb) cudaFree(s->viz_rgb);
    double *fields[] = {s->arena, s->scratch_u, s->scratch_v, s->partials, s->viz_out};
    for (double *f : fields)
        if (f) cudaFree(f);
    if (s->d_sc) cudaFree(s->d_sc);
    if (s->h_sc) cudaFreeHost(s->h_sc);
    for (int i = 0; i < ST_COUNT; ++i)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (auto &g : s->step_graph)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    if (s->ev_afork) cudaEventDestroy(s->ev_afork);
    if (s->ev_ajoin) cudaEventDestroy(s->ev_ajoin);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int cfd_b200_upload(cfd_b200_solver *s, const double *u, const double *v, const double *p) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    const Geom &g = s->g;
    int rc = 0;
    if ((rc = settle_halos(s))) return rc;
    if (u || v) s->cfl_cached = false, s->halo_uv_valid = false;
    if (p) s->p_halo_valid = false;
    if (u && (rc = copy_field(s, s->u, const_cast<double *>(u), g.nx + 3, g.ny + 2, true))) return rc;
    if (v && (rc = copy_field(s, s->v, const_cast<double *>(v), g.nx + 2, g.ny + 3, true))) return rc;
    if (p && (rc = copy_field(s, s->p, const_cast<double *>(p), g.nx + 2, g.ny + 2, true))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int cfd_b200_upload_rhs(cfd_b200_solver *s, const double *rhs) {
    if (!s || !rhs) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = copy_field(s, s->rhs, const_cast<double *>(rhs), s->g.nx + 2, s->g.ny + 2, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int cfd_b200_download(cfd_b200_solver *s, double *u, double *v, double *p, double *rhs) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    const Geom &g = s->g;
    int rc = 0;
    if ((rc = settle_halos(s))) return rc;
    if (u && (rc = copy_field(s, s->u, u, g.nx + 3, g.ny + 2, false))) return rc;
    if (v && (rc = copy_field(s, s->v, v, g.nx + 2, g.ny + 3, false))) return rc;
    if (p && (rc = copy_field(s, s->p, p, g.nx + 2, g.ny + 2, false))) return rc;
    if (rhs && (rc = copy_field(s, s->rhs, rhs, g.nx + 2, g.ny + 2, false))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int cfd_b200_download_work(cfd_b200_solver *s, double *u1, double *v1) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    const Geom &g = s->g;
    int rc = 0;
    if (u1 && (rc = copy_field(s, s->u1, u1, g.nx + 3, g.ny + 2, false))) return rc;
    if (v1 && (rc = copy_field(s, s->v1, v1, g.nx + 2, g.ny + 3, false))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// The GUI's reset (main.cpp:338-346): every State field and work array back to zero, on the device.
int cfd_b200_reset(cfd_b200_solver *s) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    TRY(settle_halos(s));
    double *fields[] = {s->u, s->v, s->p, s->p_alt, s->rhs, s->u1, s->v1, s->ustar, s->vstar};
    for (double *f : fields) CU(cudaMemsetAsync(f, 0, s->n * sizeof(double), s->stream));
    s->cfl_cached = false, s->halo_uv_valid = false, s->p_halo_valid = false;
    return 0;
}

int cfd_b200_clear_work(cfd_b200_solver *s) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->u1, 0, s->n * sizeof(double), s->stream));
    CU(cudaMemsetAsync(s->v1, 0, s->n * sizeof(double), s->stream));
    return 0;
}

int cfd_b200_step_async(cfd_b200_solver *s, const cfd_b200_bc *bc, double Re, double CFL, const cfd_b200_params *pp,
                        double dt_override) {
    if (!s || !bc || !pp) return fail(CFD_B200_EINVAL, "NULL argument");
    if (pp->solver != CFD_B200_SOLVER_PCG && pp->solver != CFD_B200_SOLVER_MG) return fail(CFD_B200_EINVAL, "bad solver type");
    CU(cudaSetDevice(s->device));
    // ---- steady-state steps of a single-GPU handle go through a captured graph (launch-bound small grids gain most)
    if (s->graph_mode < 0) s->graph_mode = getenv("CFD_B200_GRAPH") ? atoi(getenv("CFD_B200_GRAPH")) : 1;
    const bool eligible = s->graph_mode == 1 && s->nranks == 1 && !s->profiling && s->variant == 1 && s->cfl_cached &&
                          pp->solver == CFD_B200_SOLVER_MG && pp->mg_mode == CFD_B200_MG_REFERENCE && pp->vcycles >= 1 &&
                          bc_equal(s->last_bc, *bc);
    if (!eligible) {
        s->graph_seen = 0;
        return enqueue_step(s, bc, Re, CFL, pp, dt_override);
    }
    auto matches = [&](const cfd_b200_solver::StepGraph &g) {
        return g.exec && g.p == s->p && bc_equal(g.bc, *bc) && g.Re == Re && g.CFL == CFL && g.dt_override == dt_override &&
               g.variant == s->variant && g.march_kernel == s->march_kernel && g.rb_strip == s->rb_strip && g.pp.vcycles == pp->vcycles && g.pp.tol == pp->tol &&
               g.pp.mg_levels == pp->mg_levels && g.pp.rbgs_sweeps == pp->rbgs_sweeps;
    };
    for (auto &g : s->step_graph)
        if (matches(g)) {
            CU(cudaGraphLaunch(g.exec, s->stream));
            s->launches += g.launches;
            if (g.swaps_p) { double *t = s->p; s->p = s->p_alt; s->p_alt = t; }
            s->last_smoother = g.last_smoother;
            s->last_bc_step = *bc;
            s->ev_valid = false;
            return 0;
        }
    if (++s->graph_seen < 2) return enqueue_step(s, bc, Re, CFL, pp, dt_override); // first: run plainly (lazy initialisations)
    // capture this step; if anything in it cannot be captured, graphs stay off for this handle
    cfd_b200_solver::StepGraph *slot = s->step_graph[0].exec && s->step_graph[0].p != s->p ? &s->step_graph[1] : &s->step_graph[0];
    if (slot->exec) cudaGraphExecDestroy(slot->exec), slot->exec = nullptr;
    double *const p0 = s->p, *const palt0 = s->p_alt;
    const long long l0 = s->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        s->graph_mode = 0;
        return enqueue_step(s, bc, Re, CFL, pp, dt_override);
    }
    const int rc = enqueue_step(s, bc, Re, CFL, pp, dt_override);
    const cudaError_t ec = cudaStreamEndCapture(s->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == 0 && ec == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        slot->exec = exec, slot->bc = *bc, slot->pp = *pp, slot->Re = Re, slot->CFL = CFL, slot->dt_override = dt_override;
        slot->variant = s->variant, slot->march_kernel = s->march_kernel, slot->rb_strip = s->rb_strip, slot->p = p0, slot->launches = s->launches - l0;
        slot->swaps_p = s->p != p0, slot->last_smoother = s->last_smoother;
        CU(cudaGraphLaunch(exec, s->stream)); // the capture recorded the step, it did not run it
        return 0;
    }
    // not capturable: nothing was executed, put the host-side state back and run the step the plain way
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    s->p = p0, s->p_alt = palt0, s->launches = l0, s->graph_mode = 0;
    return enqueue_step(s, bc, Re, CFL, pp, dt_override);
}

int cfd_b200_sync(cfd_b200_solver *s, cfd_b200_step_info *info) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    fill_info(s, info);
    if (s->h_sc->comm_error)
        return fail(CFD_B200_ECOMM, "rank %d: a peer did not arrive within %.3g s, or reported that itself (peer-memory exchange); cfd_b200_comm_reset re-arms the handles",
                    s->rank, (double)s->p2p_timeout_ns * 1e-9);
    return 0;
}

// Collective over all slabs of a job: forget a communication failure (CFD_B200_ECOMM) and start the flag protocol afresh.
// Every rank must call it; the State itself is left as it is (upload or reset it: a failed step's fields are invalid).
int cfd_b200_comm_reset(cfd_b200_solver *s) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    if (s->nranks == 1) return 0;
    CU(cudaSetDevice(s->device));
    NcclApi &nc = nccl_api();
    if (!bounded_sync(s, 2.0 * (double)s->p2p_timeout_ns * 1e-9 + 10.0)) return fail(CFD_B200_ECOMM, "rank %d: the stream did not drain", s->rank);
    // nobody may still be storing into my flags when I clear them, nobody may signal before everybody has cleared
    NCCLCHK(nc.AllReduce(s->gathered, s->gathered, 1, ncclDouble, ncclMin, s->comm, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (s->p2p_on) {
        P2PBlock *b = s->blk[s->rank];
        CU(cudaMemsetAsync(b->enter_flag, 0, sizeof(b->enter_flag), s->stream));
        CU(cudaMemsetAsync(b->data_flag, 0, sizeof(b->data_flag), s->stream));
        CU(cudaMemsetAsync(b->red_flag, 0, sizeof(b->red_flag), s->stream));
        CU(cudaMemsetAsync(&b->ticket, 0, sizeof(b->ticket), s->stream));
    }
    CU(cudaMemsetAsync(&s->d_sc->comm_error, 0, sizeof(int), s->stream));
    CU(cudaMemsetAsync(s->d_sc->ticket, 0, sizeof(s->d_sc->ticket), s->stream));
    s->h_sc->comm_error = 0;
    s->halo_epoch = s->red_epoch = 0;
    s->sent_since_barrier = 0, s->ep_uv = 0;
    s->halo_uv_valid = s->p_halo_valid = s->cfl_cached = false;
    NCCLCHK(nc.AllReduce(s->gathered, s->gathered, 1, ncclDouble, ncclMin, s->comm, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// Test hook: keeps the handle's stream busy for `microseconds` (one spinning thread), so that a test can make one slab
// late and let its neighbours run ahead as far as the flag protocol allows.
int cfd_b200_debug_delay(cfd_b200_solver *s, double microseconds) {
    if (!s || !(microseconds >= 0.0)) return fail(CFD_B200_EINVAL, "bad argument");
    CU(cudaSetDevice(s->device));
    LAUNCH(s, k_debug_delay, 1, 1, (unsigned long long)(microseconds * 1e3));
    return 0;
}

// Measurement hook (scripts/comm_probe.py): enqueue `reps` communication operations of one kind back to back.
//   0 = reduction fused with its tail (k_p2p_reduce / allreduce), 1 = one-way send of 2 rows of u, v + stand-alone wait,
//   2 = the two-round-trip exchange of the same rows (k_halo_xchg / NCCL send-recv), 3 = reduction with one row of p riding on it
int cfd_b200_debug_comm(cfd_b200_solver *s, int kind, int reps) {
    if (!s || reps < 0) return fail(CFD_B200_EINVAL, "bad argument");
    if (s->nranks == 1) return 0;
    CU(cudaSetDevice(s->device));
    TRY(settle_halos(s));
    for (int i = 0; i < reps; ++i) {
        unsigned long long ep = 0;
        switch (kind) {
        case 0: TRY(reduce_fin(s, RED_DIAG, 0.0, 0)); break;
        case 1:
            s->sent_since_barrier = 0; // measurement only: the rows carry no meaning here
            TRY(halo_send2(s, s->u, s->v, 2, HALO_BOTH, &ep));
            TRY(halo_wait(s, ep));
            break;
        case 2: TRY(halo_exchange2(s, s->u, s->v, 2, HALO_BOTH)); break;
        case 3: TRY(reduce_fin(s, RED_DIAG, 0.0, 0, s->p, 1, HALO_BOTH)); break;
        default: return fail(CFD_B200_EINVAL, "bad kind %d", kind);
        }
    }
    s->halo_uv_valid = s->p_halo_valid = false;
    CU(cudaGetLastError());
    return 0;
}

// Measurement: microseconds spent polling for a peer and number of polls that had to wait since the handle was created
// (or since the last call with reset != 0), for the 4 kinds of waits of Scalars::wait_ns.  Synchronises the stream.
int cfd_b200_wait_stats(cfd_b200_solver *s, double *us4, long long *count4, int reset) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < 4; ++k) {
        if (us4) us4[k] = (double)s->h_sc->wait_ns[k] * 1e-3;
        if (count4) count4[k] = (long long)s->h_sc->wait_n[k];
    }
    if (reset) CU(cudaMemsetAsync(s->d_sc->wait_ns, 0, sizeof(s->d_sc->wait_ns) + sizeof(s->d_sc->wait_n), s->stream));
    return 0;
}

int cfd_b200_set_smoother_strip(cfd_b200_solver *s, int rows) {
    if (!s || rows < 0) return fail(CFD_B200_EINVAL, "bad argument");
    s->rb_strip = rows;
    return 0;
}

int cfd_b200_smoother_strip_for(int nx, int ny, int sm_count) {
    if (nx < 4 || ny < 4 || sm_count < 1) return fail(CFD_B200_EINVAL, "bad argument");
    return stream_strip_rows(nx, ny + 2, sm_count, nullptr, nullptr);
}

int cfd_b200_last_smoother(const cfd_b200_solver *s) { return s ? s->last_smoother : 0; }
int cfd_b200_last_pcg(const cfd_b200_solver *s) { return s ? s->last_pcg : 0; }

int cfd_b200_comm_mode(const cfd_b200_solver *s) { return !s || s->nranks == 1 ? 0 : (s->p2p_on ? 2 : 1); }

int cfd_b200_step(cfd_b200_solver *s, const cfd_b200_bc *bc, double Re, double CFL, const cfd_b200_params *pp,
                  double dt_override, cfd_b200_step_info *info) {
    int rc = cfd_b200_step_async(s, bc, Re, CFL, pp, dt_override);
    if (rc) return rc;
    return cfd_b200_sync(s, info);
}

int cfd_b200_divergence_l2(cfd_b200_solver *s, double *out) {
    if (!s || !out) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    const Geom &g = s->g;
    CU(cudaMemsetAsync(&s->d_sc->diag_max_bits, 0, sizeof(unsigned long long), s->stream));
    dim3 grid(cdiv(g.nx, 256), g.ny < 128 ? g.ny : 128);
    if (s->nranks > 1) { // the top cell row reads the v face owned by the slab above
        TRY(settle_halos(s));
        TRY(halo_exchange2(s, s->v, nullptr, 1, HALO_DOWN));
    }
    LAUNCH(s, k_diag, grid, 256, g, s->u, s->v, s->d_sc, s->partials);
    {
        int rc = reduce_fin(s, RED_DIAG, 0.0, 0);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    double denom = static_cast<double>(g.nx) * g.ny_glob; // metrics.cpp:84
    *out = std::sqrt(s->h_sc->diag_sum / denom);
    return 0;
}

int cfd_b200_max_velocity(cfd_b200_solver *s, double *out) {
    double dummy;
    int rc = cfd_b200_divergence_l2(s, &dummy);
    if (rc) return rc;
    if (!out) return fail(CFD_B200_EINVAL, "NULL argument");
    unsigned long long b = s->h_sc->diag_max_bits;
    memcpy(out, &b, sizeof(double));
    return 0;
}

int cfd_b200_apply_bc(cfd_b200_solver *s, const cfd_b200_bc *bc, int which) {
    if (!s || !bc) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    TRY(settle_halos(s));
    s->cfl_cached = false, s->halo_uv_valid = false, s->p_halo_valid = false;
    enqueue_bc(s, to_bcd(bc), (which & 1) ? s->u : nullptr, (which & 2) ? s->v : nullptr, (which & 4) ? s->p : nullptr, 0);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_compute_cfl(cfd_b200_solver *s, double Re, double CFL, double *dt_out) {
    if (!s || !dt_out) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    enqueue_cfl(s, Re, CFL, 0.0, 1);
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    *dt_out = s->h_sc->dt;
    return 0;
}

int cfd_b200_rhs_terms(cfd_b200_solver *s, double invRe, double *du, double *dv) {
    if (!s || !du || !dv) return fail(CFD_B200_EINVAL, "NULL argument");
    if (s->nranks > 1) return fail(CFD_B200_EUNSUPPORTED, "stage-level entry points are single-GPU");
    CU(cudaSetDevice(s->device));
    int rc = need_scratch(s);
    if (rc) return rc;
    const Geom &g = s->g;
    CU(cudaMemsetAsync(s->scratch_u, 0, s->n * sizeof(double), s->stream)); // advect_* zero-fill first (advection.cpp:41)
    CU(cudaMemsetAsync(s->scratch_v, 0, s->n * sizeof(double), s->stream));
    enqueue_rhs<0>(s, invRe, s->u, s->v, s->scratch_u, s->scratch_v);
    if ((rc = copy_field(s, s->scratch_u, du, g.nx + 3, g.ny + 2, false))) return rc;
    if ((rc = copy_field(s, s->scratch_v, dv, g.nx + 2, g.ny + 3, false))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_advect(cfd_b200_solver *s, double *du, double *dv) {
    if (!s || (!du && !dv)) return fail(CFD_B200_EINVAL, "NULL argument");
    if (s->nranks > 1) return fail(CFD_B200_EUNSUPPORTED, "stage-level entry points are single-GPU");
    CU(cudaSetDevice(s->device));
    int rc = need_scratch(s);
    if (rc) return rc;
    const Geom &g = s->g;
    CU(cudaMemsetAsync(s->scratch_u, 0, s->n * sizeof(double), s->stream)); // advect_* zero-fill first (advection.cpp:41)
    CU(cudaMemsetAsync(s->scratch_v, 0, s->n * sizeof(double), s->stream));
    RhsConst k{g.idx, g.idy, g.idx2, g.idy2, 0.0};
    dim3 block(RHS_TX, RHS_TY);
    RhsRects rects{};
    rects.n = 1, rects.ntx[0] = cdiv(g.nx + 1, RHS_TX), rects.end[0] = rects.ntx[0] * cdiv(g.ny + 1, RHS_TY);
    LAUNCH(s, k_rhs_simple<3>, rects.end[0], block, g, k, s->u, s->v, s->scratch_u, s->scratch_v, s->scratch_u, s->scratch_v, s->d_sc, rects, WaitArgs{});
    if (du && (rc = copy_field(s, s->scratch_u, du, g.nx + 3, g.ny + 2, false))) return rc;
    if (dv && (rc = copy_field(s, s->scratch_v, dv, g.nx + 2, g.ny + 3, false))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_diffuse(cfd_b200_solver *s, int which, double invRe, double *d) {
    if (!s || !d || (which != 0 && which != 1)) return fail(CFD_B200_EINVAL, "bad argument");
    if (s->nranks > 1) return fail(CFD_B200_EUNSUPPORTED, "stage-level entry points are single-GPU");
    CU(cudaSetDevice(s->device));
    int rc = need_scratch(s);
    if (rc) return rc;
    const Geom &g = s->g;
    double *scratch = which ? s->scratch_v : s->scratch_u;
    const int pitch = which ? g.nx + 2 : g.nx + 3, rows = which ? g.ny + 3 : g.ny + 2;
    const int IX = which ? g.nx : g.nx + 1, JY = which ? g.ny + 1 : g.ny;
    if ((rc = copy_field(s, scratch, d, pitch, rows, true))) return rc;
    if (IX > 2 && JY > 2)
        LAUNCH(s, k_diffuse_add, dim3(cdiv(IX - 2, 128), JY - 2), 128, g, which ? s->v : s->u, scratch, invRe, IX, JY);
    if ((rc = copy_field(s, scratch, d, pitch, rows, false))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_divergence(cfd_b200_solver *s) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    if (s->nranks > 1) { // the top cell row reads the v face owned by the slab above
        TRY(settle_halos(s));
        TRY(halo_exchange2(s, s->v, nullptr, 1, HALO_DOWN));
    }
    enqueue_divergence(s, 0, 0);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_pressure_solve(cfd_b200_solver *s, const cfd_b200_params *pp, double *residual, int *iters) {
    if (!s || !pp) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    TRY(settle_halos(s));
    s->p_halo_valid = false;
    int rc = enqueue_solve(s, pp);
    if (rc) return rc;
    s->p_halo_valid = false;
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (residual) *residual = s->h_sc->res;
    if (iters) *iters = s->h_sc->iters;
    return 0;
}

int cfd_b200_subtract_grad_p(cfd_b200_solver *s, double dt) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    TRY(settle_halos(s));
    s->cfl_cached = false, s->halo_uv_valid = false;
    enqueue_project(s, nullptr, dt, 0);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

// ---- readers / mutators around the step -------------------------------------------------
namespace {
// compute_scalar on the device; leaves the field in s->viz_out and {min, max, sum, sum^2} in h_sc->viz
int viz_scalar_device(cfd_b200_solver *s, int field) {
    if (field < 0 || field > VIZ_TEMPERATURE) return fail(CFD_B200_EINVAL, "bad scalar field %d", field);
    const Geom &g = s->g;
    if (s->nranks > 1) { // row slabs: the vorticity reads u one row below / above the slab, every field v one row above
        TRY(settle_halos(s));
        if (!s->halo_uv_valid) TRY(halo_exchange2(s, s->u, s->v, 2, HALO_BOTH));
        s->halo_uv_valid = true;
    }
    if (!s->viz_out) CU(cudaMalloc(&s->viz_out, (size_t)g.nx * g.ny * sizeof(double)));
    dim3 grid(cdiv(g.nx, 256), g.ny < 128 ? g.ny : 128);
    LAUNCH(s, k_viz_scalar, grid, 256, g, s->u, s->v, s->p, field, s->viz_out, s->d_sc, s->partials);
    TRY(reduce_fin(s, RED_VIZ, 0.0, 0)); // row slabs: minimum, maximum and the two sums over all slabs
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s->stream));
    return 0;
}
} // namespace

int cfd_b200_compute_scalar(cfd_b200_solver *s, int field, double *out, cfd_b200_field_stats *st) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    int rc = viz_scalar_device(s, field);
    if (rc) return rc;
    const size_t n = (size_t)s->g.nx * s->g.ny;
    if (out) CU(cudaMemcpyAsync(out, s->viz_out, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (st) {
        const double *z = s->h_sc->viz;
        const double span = (1e-6 < z[1] - z[0]) ? z[1] - z[0] : 1e-6; // std::max(1e-6, vmax - vmin), viz.cpp:63
        st->vmin = z[0];
        st->vmax = z[0] + span;
        // Gui::update_field_statistics, gui.cpp:1452-1474 (every value is finite after compute_scalar's own guard)
        const double count = static_cast<double>((size_t)s->g.nx * s->g.ny_glob);
        st->mean = z[2] / count;
        const double variance = (z[3] / count) - (st->mean * st->mean);
        st->stddev = std::sqrt(0.0 < variance ? variance : 0.0);
    }
    return 0;
}

int cfd_b200_render_rgb(cfd_b200_solver *s, int field, int colormap, float scale_min, float scale_max, int normalize,
                        float gamma, unsigned char *rgb) {
    if (!s || !rgb) return fail(CFD_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = viz_scalar_device(s, field);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream)); // vmin / vmax decide the scale (gui.cpp:1523-1532)
    const size_t n = (size_t)s->g.nx * s->g.ny;
    if (!s->viz_rgb) CU(cudaMalloc(&s->viz_rgb, 3 * n));
    double vmin = s->h_sc->viz[0], hi = s->h_sc->viz[1];
    double vmax = vmin + ((1e-6 < hi - vmin) ? hi - vmin : 1e-6);
    if (scale_min != 0.0f || scale_max != 1.0f) {
        vmin = scale_min;
        vmax = scale_max;
    }
    const double scale = 1.0 / ((1e-6 < vmax - vmin) ? vmax - vmin : 1e-6);
    const double inv_gamma = 1.0 / gamma; // "1.0 / gamma" with a float gamma, gui.cpp:1543
    LAUNCH(s, k_viz_rgb, s->sm_count * 8, 256, n, s->viz_out, vmin, scale, normalize, inv_gamma, colormap, s->viz_rgb);
    CU(cudaMemcpyAsync(rgb, s->viz_rgb, 3 * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}

int cfd_b200_apply_shapes(cfd_b200_solver *s, const cfd_b200_shape *shapes, int n) {
    if (!s || (n > 0 && !shapes) || n < 0) return fail(CFD_B200_EINVAL, "bad argument");
    CU(cudaSetDevice(s->device));
    const Geom &g = s->g;
    if (s->nranks > 1) TRY(settle_halos(s)); // row slabs: shapes are given in cells of the WHOLE grid; every slab masks its rows
    auto clampi = [](int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); };
    for (int k = 0; k < n; ++k) {
        const cfd_b200_shape &sh = shapes[k];
        if (!sh.enabled) continue;
        // bounding box in cells, float -> int truncation, clamped to the grid (shapes.cpp:12-30); a circle's box ends at
        // pos + size.x in both directions while its centre uses size.x/2 and size.y/2 (shapes.cpp:20-23, :42-44)
        const float ex = sh.pos_x + sh.size_x, ey = sh.type == 0 ? sh.pos_y + sh.size_y : sh.pos_y + sh.size_x;
        const int i0 = clampi(static_cast<int>(sh.pos_x), 0, g.nx - 1), i1 = clampi(static_cast<int>(ex), 0, g.nx - 1);
        int j0 = clampi(static_cast<int>(sh.pos_y), 0, g.ny_glob - 1), j1 = clampi(static_cast<int>(ey), 0, g.ny_glob - 1);
        if (i1 < i0 || j1 < j0) continue;
        j0 = j0 < g.j0 ? g.j0 : j0, j1 = j1 > g.j0 + g.ny - 1 ? g.j0 + g.ny - 1 : j1; // the rows of the box inside this slab
        if (j1 < j0) continue;
        const float cx = sh.pos_x + sh.size_x / 2.0f, cy = sh.pos_y + sh.size_y / 2.0f, rad = sh.size_x / 2.0f;
        LAUNCH(s, k_shape_apply, dim3(cdiv(i1 - i0 + 1, 256), j1 - j0 + 1), 256, g, s->u, s->v, i0, j0, i1, j1,
               sh.type != 0 ? 1 : 0, cx, cy, rad, sh.bc_type != 0 ? 1 : 0);
    }
    s->cfl_cached = false; // u, v changed behind the cached maxima
    s->halo_uv_valid = false; // ... and behind the neighbours' copies of the slab's edge rows
    CU(cudaGetLastError());
    return 0;
}

long long cfd_b200_launch_count(const cfd_b200_solver *s) { return s ? s->launches : 0; }

int cfd_b200_set_profiling(cfd_b200_solver *s, int on) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    s->profiling = on ? 1 : 0;
    s->ev_valid = false;
    return 0;
}

int cfd_b200_stage_times(cfd_b200_solver *s, int n, const char **names, float *ms) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    if (!s->ev_valid) return fail(CFD_B200_EINVAL, "no profiled step: call cfd_b200_set_profiling(s,1) and run a step");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    int count = ST_COUNT - 1;
    for (int i = 0; i < count && i < n; ++i) {
        if (names) names[i] = kStageNames[i];
        if (ms) CU(cudaEventElapsedTime(&ms[i], s->ev[i], s->ev[i + 1]));
    }
    return count;
}

void *cfd_b200_stream(cfd_b200_solver *s) { return s ? (void *)s->stream : nullptr; }

int cfd_b200_set_variant(cfd_b200_solver *s, int variant) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    s->variant = variant;
    return 0;
}

int cfd_b200_set_march_kernel(cfd_b200_solver *s, int version) {
    if (!s) return fail(CFD_B200_EINVAL, "solver is NULL");
    if (version != 1 && version != 2) return fail(CFD_B200_EINVAL, "march kernel version must be 1 or 2");
    s->march_kernel = version;
    return 0;
}

} // extern "C"
