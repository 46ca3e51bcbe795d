// p2p.cuh -- row-slab communication straight over NVLink 5 / NVSwitch peer memory (one process per GPU, CUDA IPC).
//
// Every message of the slab decomposition is tiny -- a few rows to the two neighbours, or 1-4 doubles
// to everybody -- so what matters is latency, not bandwidth.  An NCCL send/recv group or a 2-double allreduce costs
// ~25-35 us per call on 8 GPUs; the kernels here do the same work in one launch each with plain stores into the peer's
// memory and flag words for ordering (a few us):
//
//   k_halo_xchg   handshake ("I have entered exchange #e": everything that reads or writes my halo rows before this
//                 point of my stream is done) -> wait for the neighbour's handshake -> store my edge rows into the
//                 neighbour's halo rows -> fence -> data flag -> wait for the neighbour's data flag.
//   k_p2p_reduce  the allreduce / all-gather of the PCG dots, the smoother residual, the CFL maxima and the divergence
//                 sums FUSED with the scalar tail that consumes them: every rank stores its partial into slot [rank] of
//                 every peer, waits for all slots of its own block and adds them in rank order, so all ranks compute
//                 bit-identical totals (and therefore take identical decisions) without any library call.
//
// All flags are monotonically increasing 64-bit epochs with a single writer; nothing is ever reset, so there is no
// ABA window.  The reduce slots are double-buffered on the epoch parity: a rank can be at most one reduce ahead of the
// slowest one (it needs everybody's contribution to pass the next).  Waits are bounded (common.cuh: p2p_wait): a peer that
// never arrives sets Scalars::comm_error instead of hanging the GPU, the failure travels to the peers as a poison epoch,
// and cfd_b200_sync reports CFD_B200_ECOMM on every rank; cfd_b200_comm_reset re-arms the handles.
//
// NCCL (comm.cuh) stays the bootstrap (handle exchange, agreement that every rank mapped its peers) and the fallback
// when peer mapping is not possible.
#pragma once
#include "common.cuh"

#define P2P_MAXR 16
#define P2P_MAGIC 0xCFDB200000000000ull

struct P2PBlock {
    unsigned long long magic;               // P2P_MAGIC + rank: checked through the mapping before it is trusted
    unsigned long long enter_flag[2];       // [0] written by the slab below, [1] by the slab above
    unsigned long long data_flag[2];
    unsigned long long red_flag[2][P2P_MAXR]; // [epoch parity][source rank]
    double red_val[2][P2P_MAXR][6];
    unsigned int ticket;      // last-block ticket of k_halo_xchg (solver stream)
    unsigned int send_ticket; // ... of k_halo_send (communication stream; may run next to the kernel above)
};

__device__ __forceinline__ void p2p_signal(unsigned long long *flag, unsigned long long epoch, const Scalars *sc) {
    __threadfence_system(); // everything this thread has observed or written is visible before the flag
    if (*reinterpret_cast<const volatile int *>(&sc->comm_error)) epoch = P2P_POISON; // my data cannot be trusted: say so
    *reinterpret_cast<volatile unsigned long long *>(flag) = epoch;
}
struct HaloJob {
    const double *src; // my rows
    double *dst;       // the neighbour's halo rows (peer memory)
    int nb;            // 0 = slab below, 1 = slab above
};
struct HaloArgs {
    HaloJob job[4];
    int njobs;
    unsigned long long count; // doubles per job (whole rows: a multiple of 16)
    unsigned long long *peer_enter[2], *peer_data[2]; // the neighbours' flag words I write (NULL: no neighbour there)
    unsigned long long epoch;
};

__global__ void __launch_bounds__(256) k_halo_xchg(HaloArgs a, P2PBlock *me, Scalars *sc) { pdl_begin();
    __shared__ int s_go, s_last;
    const int tid = threadIdx.x;
    // 1. handshake: the neighbour may write into my halo rows from now on, and I into its rows once it says the same
    if (tid < 2) {
        if (blockIdx.x == 0 && a.peer_enter[tid]) p2p_signal(a.peer_enter[tid], a.epoch, sc);
        bool ok = true;
        if (a.peer_enter[tid]) ok = p2p_wait(&me->enter_flag[tid], a.epoch, sc, 3);
        if (tid == 0) s_go = 1;
        __syncwarp(0x3);
        if (!ok) s_go = 0;
    }
    __syncthreads();
    // 2. my edge rows -> the neighbours' halo rows (16-byte stores over NVLink)
    if (s_go) {
        const size_t n2 = a.count / 2;
        for (int j = 0; j < a.njobs; ++j) {
            const double2 *src = reinterpret_cast<const double2 *>(a.job[j].src);
            double2 *dst = reinterpret_cast<double2 *>(a.job[j].dst);
            for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&me->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // 3. the last block: all rows of all blocks are out -> data flags, then wait for the neighbours' rows
    if (tid < 2 && a.peer_data[tid]) {
        p2p_signal(a.peer_data[tid], a.epoch, sc);
        p2p_wait(&me->data_flag[tid], a.epoch, sc, 3);
    }
    if (tid == 0) me->ticket = 0u;
}

// ---- one-way sends: the exchange above without its two round trips ---------------------------------------------
// k_halo_send stores my edge rows into the neighbours' halo rows and raises the data flag; it waits for nothing.  The
// handshake can go because every use of it sits behind a reduction that is a barrier (k_p2p_reduce: a rank that has seen
// every contribution knows that every rank has finished the kernels in front of its own contribution, i.e. has stopped
// reading the halo rows of the previous round); capi.cu keeps track of that and falls back to k_halo_xchg otherwise.  The
// receiver waits where it needs the rows -- k_halo_wait in front of the consumer, or the consumer's own edge blocks
// (p2p_wait_block), so that the transfer and the neighbour's lateness hide behind the interior work.
struct SendJob {
    const double *src;
    double *dst;
    unsigned long long count; // doubles (whole rows: a multiple of 16)
};
struct SendArgs {
    SendJob job[8];
    int njobs;
    unsigned long long *peer_data[2]; // the neighbours' data flags I raise (NULL: no neighbour there / nothing sent that way)
    unsigned long long epoch;
};
__global__ void __launch_bounds__(256) k_halo_send(SendArgs a, P2PBlock *me, const Scalars *sc) {
    __shared__ int s_last;
    const int tid = threadIdx.x;
    for (int j = 0; j < a.njobs; ++j) {
        const double2 *src = reinterpret_cast<const double2 *>(a.job[j].src);
        double2 *dst = reinterpret_cast<double2 *>(a.job[j].dst);
        const size_t n2 = a.job[j].count / 2;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&me->send_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (tid < 2 && a.peer_data[tid]) p2p_signal(a.peer_data[tid], a.epoch, sc); // all rows of all blocks are out
    if (tid == 0) me->send_ticket = 0u;
}
// Stand-alone wait for the neighbours' send number `epoch` (consumers that have no edge blocks of their own).
__global__ void k_halo_wait(WaitArgs w, Scalars *sc) { pdl_begin();
    if (threadIdx.x < 2 && w.f[threadIdx.x]) p2p_wait(w.f[threadIdx.x], w.epoch, sc);
}
__global__ void k_debug_delay(unsigned long long ns) { pdl_begin();
    const unsigned long long t0 = p2p_now();
    while (p2p_now() - t0 < ns) {}
}

// ---- reductions fused with their scalar tails ------------------------------------------------------------------
enum {
    RED_PCG_INIT = 0, // 2 sums  -> pcg_tail_init
    RED_PCG1,         // 1 sum   -> pcg_tail1
    RED_PCG2,         // 2 sums  -> pcg_tail2
    RED_RBGS,         // 1 sum   -> rbgs_tail
    RED_STEP_END,     // max|u|, max|v|, sum div^2 before / after (end of a step); + the deferred residual sum of the last smooth()
    RED_CFL,          // max|u|, max|v| (compute_cfl after an upload)
    RED_DIAG,         // divergence_l2 sum + max_velocity max
    RED_VC,           // 1 sum   -> mg_norm_tail (true V-cycle on row slabs: residual norm after a cycle)
    RED_VIZ           // min, max, sum, sum of squares of the visualisation scalar (Scalars::viz)
};
struct RedArgs {
    P2PBlock *blk[P2P_MAXR]; // every rank's block (blk[rank] = mine)
    int nranks, rank, kind;
    unsigned long long epoch;
    double tol;
    int max_iters;
    int defer_rbgs; // RED_STEP_END also completes the last smooth() of the step (its local sum waits in Scalars::red[0])
    // optional: a halo exchange riding on the reduction (halo = 1).  The reduction is a barrier -- a rank that has seen
    // every contribution knows that every rank has finished the kernels in front of its own contribution, i.e. has
    // stopped reading its halo rows -- so the rows can be stored without the handshake of k_halo_xchg.
    HaloJob job[2];
    int halo, njobs; // halo = 1: flags are exchanged with both neighbours even if this rank has no rows to send
    unsigned long long count, hepoch;
    unsigned long long *peer_data[2];
};
#define P2P_RED_THREADS 256

__global__ void __launch_bounds__(P2P_RED_THREADS) k_p2p_reduce(RedArgs a, Scalars *sc) { pdl_begin();
    // a solver that has converged keeps enqueuing its iterations (no host round trip): their reductions, and the rows that
    // ride on them, are skipped -- by every rank alike, `done` comes from totals that are bit-identical everywhere
    if ((a.kind == RED_PCG1 || a.kind == RED_PCG2 || a.kind == RED_RBGS || a.kind == RED_VC) && sc->done) return;
    const int t = threadIdx.x, par = (int)(a.epoch & 1ull);
    P2PBlock *me = a.blk[a.rank];
    double v0, v1, v2 = 0.0, v3 = 0.0, v4 = 0.0;
    switch (a.kind) { // every lane reads the same few words; lanes < nranks forward them
    case RED_STEP_END:
        v2 = sc->div_before, v3 = sc->div_after; // partial sums of squares
        v4 = sc->red[0];
        // fallthrough
    case RED_CFL:
        v0 = __longlong_as_double((long long)sc->umax_bits), v1 = __longlong_as_double((long long)sc->vmax_bits);
        break;
    case RED_DIAG:
        v0 = sc->diag_sum, v1 = __longlong_as_double((long long)sc->diag_max_bits);
        break;
    case RED_VIZ:
        v0 = sc->viz[0], v1 = sc->viz[1], v2 = sc->viz[2], v3 = sc->viz[3];
        break;
    default:
        v0 = sc->red[0], v1 = sc->red[1];
    }
    if (t < a.nranks) {
        P2PBlock *dst = a.blk[t];
        volatile double *slot = dst->red_val[par][a.rank];
        slot[0] = v0, slot[1] = v1, slot[2] = v2, slot[3] = v3, slot[4] = v4;
        p2p_signal(&dst->red_flag[par][a.rank], a.epoch, sc);
        p2p_wait(&me->red_flag[par][t], a.epoch, sc, 2);
    }
    __syncthreads();
    if (t == 0) {
    const bool is_max0 = a.kind == RED_STEP_END || a.kind == RED_CFL, is_max1 = is_max0 || a.kind == RED_DIAG;
    const bool viz = a.kind == RED_VIZ; // slot 0: minimum, slot 1: maximum (both start from the first rank's values)
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0, r4 = 0.0;
    if (viz) r0 = me->red_val[par][0][0], r1 = me->red_val[par][0][1];
    for (int r = 0; r < a.nranks; ++r) { // fixed rank order: bit-identical on every rank
        const volatile double *slot = me->red_val[par][r];
        const double x0 = slot[0], x1 = slot[1];
        r0 = viz ? min2(r0, x0) : is_max0 ? max2(r0, x0) : r0 + x0;
        r1 = (viz || is_max1) ? max2(r1, x1) : r1 + x1;
        r2 += slot[2];
        r3 += slot[3];
        r4 += slot[4];
    }
    switch (a.kind) {
    case RED_PCG_INIT: pcg_tail_init(sc, r0, r1, a.tol, a.max_iters); break;
    case RED_PCG1: if (!sc->done) pcg_tail1(sc, r0); break;
    case RED_PCG2: if (!sc->done) pcg_tail2(sc, r0, r1, a.tol, a.max_iters); break;
    case RED_RBGS: if (!sc->done) rbgs_tail(sc, r0, a.tol); break;
    case RED_VC: if (!sc->done) mg_norm_tail(sc, r0, a.tol); break;
    case RED_STEP_END:
        sc->div_before = r2, sc->div_after = r3;
        if (a.defer_rbgs && !sc->done) { // the last smooth() of the step: no decision of the step depended on its residual
            rbgs_tail(sc, r4, a.tol);
            if (!finite_d(sc->res)) sc->nonfinite_res = 1, sc->res = 1e9; // last_residual_ (pressure.cpp:123)
        }
        // fallthrough
    case RED_CFL:
        sc->umax_bits = (unsigned long long)__double_as_longlong(r0), sc->vmax_bits = (unsigned long long)__double_as_longlong(r1);
        break;
    case RED_DIAG:
        sc->diag_sum = r0, sc->diag_max_bits = (unsigned long long)__double_as_longlong(r1);
        break;
    case RED_VIZ:
        sc->viz[0] = r0, sc->viz[1] = r1, sc->viz[2] = r2, sc->viz[3] = r3;
        break;
    }
    } // t == 0
    if (!a.halo) return;
    // ---- the halo rows that ride along: store, fence, data flag, wait for the neighbours' rows
    const size_t n2 = a.count / 2;
    for (int j = 0; j < a.njobs; ++j) {
        const double2 *src = reinterpret_cast<const double2 *>(a.job[j].src);
        double2 *dst = reinterpret_cast<double2 *>(a.job[j].dst);
#pragma unroll 4
        for (size_t i = t; i < n2; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (t < 2 && a.peer_data[t]) {
        p2p_signal(a.peer_data[t], a.hepoch, sc);
        p2p_wait(&me->data_flag[t], a.hepoch, sc, 2);
    }
}
