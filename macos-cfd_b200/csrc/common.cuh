// common.cuh -- geometry, device layout, exact-arithmetic helpers and the grid-level reduction used by every
// kernel of the B200 path.  Compiled with -fmad=false: the oracle (x86-64, no FMA) never contracts a*b+c,
// so neither may we (SURVEY.md §7 hard part 1).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------------------------------------
// Device layout.  Every field (u, v, p, rhs, u1, v1, r, s, ...) uses ONE geometry so that one index function
// serves all of them: raw reference index (ii, jj) [ii = i + ngx, jj = j + ngy, ng = 1; field.hpp:69] lives at
//     element (jj + YOFF) * P + (ii + XOFF)
// XOFF = 15 puts raw column 1 (the first interior cell / face) on a 128-byte boundary, so a warp that starts
// at an even interior column issues aligned 16-byte (double2) accesses.  YOFF = 6 rows below and above hold the
// halo rows of a row slab (2 for the advection stencil, 5 for the temporally blocked smoother) and, like the >= 2
// columns of right padding, let stencil kernels read +-2 without bounds checks inside the allocation.  The pitch P is a
// multiple of 16 doubles (128 B).  Padding is zero-filled once and never written.
// ---------------------------------------------------------------------------------------------------------
#define CFD_XOFF 15
#define CFD_YOFF 6

struct Geom {
    int nx;       // cells per row
    int ny;       // LOCAL cell rows (the whole grid for a single-GPU handle)
    int P;        // pitch in doubles
    int rows;     // allocated rows = ny + 3 + 2*YOFF
    int j0;       // global index of local cell row 0 (slab decomposition)
    int ny_glob;  // global cell rows
    int bottom;   // this slab touches the domain bottom / top
    int top;
    double dx, dy;            // grid.cpp:10-11
    double idx, idy;          // 1.0/dx, 1.0/dy           (advection.cpp:34-35, project.cpp:9-10)
    double idx2, idy2;        // 1.0/(dx*dx), 1.0/(dy*dy) (diffusion.cpp:6-7, pressure.cpp:22-23)
    double denom;             // 2.0*(idx2+idy2)          (pressure.cpp:68)
    double inv_denom;         // 1.0/denom, only exact (== division) when denom is a power of two
    int denom_pow2;
    double inv_diag;          // 1.0/max(denom,1e-12)     (pressure.cpp:162-163)
};

__host__ __device__ __forceinline__ size_t gidx(const Geom &g, int ii, int jj) {
    return (size_t)(jj + CFD_YOFF) * (size_t)g.P + (size_t)(ii + CFD_XOFF);
}

// number of v-face rows this slab updates: the top slab also owns face j = ny (grid.hpp:21)
__host__ __device__ __forceinline__ int v_rows(const Geom &g) { return g.ny + (g.top ? 1 : 0); }

struct BcSideD {
    int type;
    double moving, inflow_u, inflow_v;
};
struct BcD {
    BcSideD left, right, bottom, top;
    double jet_center, jet_width;
};
enum { BC_WALL = 0, BC_MOVING = 1, BC_INFLOW = 2, BC_OUTFLOW = 3, BC_PERIODIC = 4 };

// Device-resident scalars of a step: nothing here ever needs a host round trip inside the step.
struct Scalars {
    // time step
    double dt, inv_dt;                    // inv_dt = 1.0/dt as the host would compute it (time_integrator.cpp:179)
    unsigned long long umax_bits, vmax_bits; // max |u|, max |v| as IEEE bit patterns (non-negative => ordered)
    double div_before, div_after;         // RMS divergences (time_integrator.cpp:162-171, :211-220)
    // pressure solve
    double rho, alpha, beta, res, denom;
    int iters, done, breakdown_iter, first;
    // flags
    int nonfinite_any;   // some value was sanitised to 0
    int nonfinite_p;     // p holds a non-finite value (triggers the p sanitise pass)
    int nonfinite_res;   // non-finite pressure residual
    // diagnostics
    double diag_sum;     // divergence_l2 accumulator
    unsigned long long diag_max_bits;
    unsigned int ticket[4]; // last-block tickets of the grid reductions
    double red[4];          // row slabs: local partial sums handed to the allreduce, then to the k_*_fin kernels
    double viz[4];          // k_viz_scalar: min, max, sum, sum of squares of the visualisation scalar
    int comm_error;         // a bounded peer wait ran out, or a peer reported that one of its waits did: results are invalid
    unsigned long long p2p_timeout_ns; // bound of every peer wait (CFD_B200_P2P_TIMEOUT_S, default 20 s)
    // measurement (cfd_b200_wait_stats): time spent polling for a peer and number of polls that had to wait, per kind of
    // consumer -- 0 stand-alone wait kernel, 1 a consumer kernel's edge blocks, 2 a reduction, 3 the handshake exchange
    unsigned long long wait_ns[4], wait_n[4];
    int rb_redo;            // k_rbgs_stream met a non-finite value: the guarded tile kernel redoes this smooth()
    int pcg_redo;           // k_pcg_cluster met a non-finite value: the guarded cooperative kernel behind it redoes the solve
};

// Programmatic dependent launch (sm_90+): every kernel of the path starts with pdl_begin().  launch_dependents lets
// the NEXT kernel of the stream be scheduled as soon as all blocks of this one are resident (its blocks then sit in
// griddepcontrol.wait), so launch latency and block ramp-up overlap this kernel's tail; wait returns once every
// predecessor grid has completed and its memory is visible -- nothing before it may touch global memory.
// Without the launch attribute (CFD_B200_PDL=0, or a non-kernel predecessor) both instructions are no-ops.
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Exact restatements of the C++ library calls the reference uses (argument order matters for +-0 / NaN):
//   std::min(a,b) = b<a ? b : a      std::max(a,b) = a<b ? b : a      std::clamp(x,lo,hi) = x<lo ? lo : hi<x ? hi : x
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ double min2(double a, double b) { return (b < a) ? b : a; }
__host__ __device__ __forceinline__ double max2(double a, double b) { return (a < b) ? b : a; }
__host__ __device__ __forceinline__ double clamp3(double x, double lo, double hi) {
    return (x < lo) ? lo : ((hi < x) ? hi : x);
}
// std::isfinite on the exponent bits (integer pipe, keeps the FP64 pipe free)
__device__ __forceinline__ bool finite_d(double x) {
    return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000;
}
__device__ __forceinline__ double fin0(double x) { return finite_d(x) ? x : 0.0; }
// sanitize_field semantics on a value about to be stored; records that a replacement happened
__device__ __forceinline__ double sanitize_val(double x, int &flag) {
    if (finite_d(x)) return x;
    flag = 1;
    return 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// Deterministic grid-level sum of NV per-thread values: warp shuffles -> shared memory -> one partial per
// block -> the LAST block to finish (ticket counter) adds the partials in a fixed order and calls
// fin(totals) from thread 0.  One grid-level pass, no host involvement, run-to-run reproducible.
// `partials` holds NV * gridDim.x * gridDim.y doubles.
// ---------------------------------------------------------------------------------------------------------
template <int NV> __device__ __forceinline__ void block_sum(double (&v)[NV], double *sm /* NV*32 */) {
    const int tid_ = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid_ & 31, warp = tid_ >> 5;
    const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[k * 32 + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = (lane < nwarp) ? sm[k * 32 + lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            v[k] = x; // valid in lane 0
        }
    }
    __syncthreads();
}

template <int NV, class Fin>
__device__ __forceinline__ void grid_sum(double (&v)[NV], double *partials, unsigned int *ticket, Fin fin) {
    __shared__ double sm[NV * 32];
    __shared__ int is_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthr = blockDim.x * blockDim.y;
    const unsigned int nblk = gridDim.x * gridDim.y;
    const unsigned int bid = blockIdx.x + blockIdx.y * gridDim.x;
    block_sum<NV>(v, sm);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[(size_t)k * nblk + bid] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        acc[k] = 0.0;
        for (unsigned int b = tid; b < nblk; b += nthr) acc[k] += __ldcg(&partials[(size_t)k * nblk + b]);
    }
    block_sum<NV>(acc, sm);
    if (tid == 0) {
        *ticket = 0u;
        fin(acc);
    }
}

// max over non-negative doubles through their bit patterns; NaN never wins: (m < a) is false for a NaN a,
// exactly like `umax = std::max(umax, std::abs(x))` (metrics.cpp:13).
__device__ __forceinline__ void block_max_to_global(double m, unsigned long long *dst) {
    __shared__ unsigned long long smx[32];
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x * blockDim.y + 31) >> 5;
    unsigned long long b = (unsigned long long)__double_as_longlong(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    if (lane == 0) smx[warp] = b;
    __syncthreads();
    if (warp == 0) {
        b = lane < nwarp ? smx[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_down_sync(0xffffffffu, b, o);
            b = other > b ? other : b;
        }
        if (lane == 0 && b != 0ull) atomicMax(dst, b);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Waiting for a peer GPU (row slabs over NVLink peer memory, p2p.cuh).  Flags are monotonically increasing 64-bit epochs
// with a single writer.  Every wait is bounded: a peer that never arrives sets Scalars::comm_error instead of hanging
// the GPU.  A rank in that state writes the POISON epoch wherever it would have signalled, so the failure reaches its
// peers with the next message (and everybody with the next reduction) instead of leaving them with stale rows.
// ---------------------------------------------------------------------------------------------------------
#define P2P_POISON 0xffffffffffffffffull
__device__ __forceinline__ unsigned long long p2p_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool p2p_wait(const unsigned long long *flag, unsigned long long epoch, Scalars *sc, int kind = 0) {
    const volatile unsigned long long *f = flag;
    bool ok = true;
    unsigned long long seen = *f;
    if (seen < epoch) {
        const unsigned long long t0 = p2p_now(), limit = sc->p2p_timeout_ns;
        while ((seen = *f) < epoch) {
            if (*reinterpret_cast<volatile int *>(&sc->comm_error) || p2p_now() - t0 > limit) {
                ok = false;
                break;
            }
        }
        atomicAdd(&sc->wait_ns[kind], p2p_now() - t0);
        atomicAdd(&sc->wait_n[kind], 1ull);
    }
    if (seen == P2P_POISON) ok = false; // the writer of this flag failed
    if (!ok) sc->comm_error = 1;
    __threadfence_system();
    return ok;
}
// What a consumer of halo rows waits for: the neighbours' one-way send number `epoch` (p2p.cuh: k_halo_send);
// f[0] = my data flag for the slab below, f[1] = for the slab above, NULL where nothing is expected.
struct WaitArgs {
    const unsigned long long *f[2];
    unsigned long long epoch;
};
// The wait at the top of a consumer block that reads halo rows: one thread polls, the block follows.  Blocks that only
// read the slab's own rows never call this, so their work overlaps the transfer and the neighbour's lateness.
__device__ __forceinline__ void p2p_wait_block(const WaitArgs &w, bool below, bool above, Scalars *sc) {
    if (!(below && w.f[0]) && !(above && w.f[1])) return; // block-uniform
    if (threadIdx.x + threadIdx.y * blockDim.x == 0) {
        if (below && w.f[0]) p2p_wait(w.f[0], w.epoch, sc, 1);
        if (above && w.f[1]) p2p_wait(w.f[1], w.epoch, sc, 1);
    }
    __syncthreads();
}
