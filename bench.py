#!/usr/bin/env python
"""bench.py -- cell-updates/s of the full time step (incl. the Poisson solve) + HBM-roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A "step" is one TimeIntegrator::step over the whole synthetic grid.  ONE JSON line goes to stdout.

  value        whole-job cell-updates/s, State resident in HBM, K steps enqueued back to back, device-timed
  e2e          the same metric through the C ABI with HOST buffers: every step uploads u, v, p from pinned host memory,
               runs the step and downloads u, v, p, rhs (what a GUI-style caller of TimeIntegrator::step(State&) sees)
  roofline     the dominant stage of the step: algorithmic bytes / live CUDA-event time, against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json; traffic = ncu dram bytes from profiles/ncu_traffic.json
  cpu_baseline the reference's own OpenMP solver (oracle/_ref, else the C port) on this box's host cores, SAME grid, few steps
  parity       N > 1: after 2 steps from the reset state the N row slabs hold the same u, v, p as ONE GPU running the
               same global grid (bit for bit with --solver=mg, rel-L2 <= 1e-10 with PCG), checked on the box in this run
  also         BASELINE.json's other configs, measured in the same run (short): PCG at 8192^2 per GPU, the fixed 8192^2
               lid cavity split over the N GPUs (strong scaling), and at N = 1 the small grids

Workloads are BASELINE.json's configs on synthetic (deterministic preset) data; the default is the one its metric and
target are quoted on: jet plume, --solver=mg --vcycles=2, 8192^2 fp64 per GPU (weak scaling over row slabs).
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (preset, solver, nx, ny_per_gpu, Lx, Ly_per_gpu, vcycles, tol, pcg_max_iters, BASELINE config index)
WORKLOADS = {
    "jet_mg_8192": ("jet", "mg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 4),
    "jet_pcg_8192": ("jet", "pcg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 4),
    "lid_mg_8192": ("lid", "mg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 3),
    "shear_pcg_4096": ("shear", "pcg", 4096, 4096, 1.0, 1.0, 2, 1e-6, 200, 2),
    "jet_mg_1024": ("jet", "mg", 1024, 1024, 1.0, 1.0, 2, 1e-6, 200, 1),
    "lid_pcg_512x256": ("lid", "pcg", 512, 256, 2.0, 1.0, 2, 1e-8, 200, 0),
    # one slab of the 8-GPU strong-scaling case per GPU (tuning: what a 1024-row slab costs)
    "lid_mg_8192x1024": ("lid", "mg", 8192, 1024, 1.0, 0.125, 2, 1e-6, 200, 3),
    # STRONG scaling: the global grid is fixed and split into N row slabs (BASELINE configs[3], configs[2])
    "lid_mg_8192_strong": ("lid", "mg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 3),
    "shear_pcg_4096_strong": ("shear", "pcg", 4096, 4096, 1.0, 1.0, 2, 1e-6, 200, 2),
}
DEFAULT_WORKLOAD = "jet_mg_8192"

# Algorithmic HBM bytes per cell of each stage AS IMPLEMENTED (fp64).  SURVEY.md §8d's figures are the same
# except for the pressure solve, where the implemented kernels need less than the SURVEY's formulation:
#   mg  : the smoother moves 24 B/cell per smooth() launch (R p, R rhs, W p; SURVEY's per-pass accounting: 64)
#   PCG : phase 1 + phase 2 move 24 + 40 = 64 B/cell per iteration (Ap recomputed, z never stored; SURVEY: 80), + 24 init
# The CFL maxima (16 B/cell in SURVEY) ride on the post-projection divergence kernel and cost no pass of their own.
# RK stage 2 also writes rhs = div(u*, v*) / dt of the cells it can close (k_rhs_march2<4>: 48 + 24 -> 56 B/cell) and the
# projection the post-projection divergence (k_project_div: 40 + 24 -> 48 B/cell); "divergence" / "divergence_post" are what is
# left of the two: the fix-up kernels on a few per cent of the cells.
STAGE_BYTES = {"cfl_dt": 0, "rhs_stage1": 32, "rhs_stage2": 56, "divergence": 0, "project": 48, "divergence_post": 0}
SURVEY_STEP_BYTES = {"mg": lambda vcycles, iters: 184 + 64 * vcycles, "pcg": lambda vcycles, iters: 184 + 32 + 80 * iters}
_MARCH = "k_rhs_march" if os.environ.get("CFD_B200_MARCH") == "1" else "k_rhs_march2"  # capi.cu: enqueue_rhs picks the same way
STAGE_KERNEL = {"rhs_stage1": _MARCH + "<1>", "rhs_stage2": _MARCH + "<4>", "divergence": "k_div_fix_pre",
                "project": "k_project_div", "divergence_post": "k_div_fix"}
PCG_KERNEL = "k_pcg_phase1_march + k_pcg_phase2_march (one iteration)"


def ncu_traffic(kernel, nx, nrows):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by profiles/ncu_traffic.py from the .ncu-rep), or None off the profiled grid."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, None
    if t.get("grid") != f"{nx}x{nrows}":
        return None, None
    e = t.get("kernels", {}).get(kernel)
    return (e["dram_bytes"], e.get("source")) if e else (None, None)


def solve_bytes(solver, vcycles, iters):
    return 24 * vcycles if solver == "mg" else 24 + 64 * iters


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def grid_of(name, N):
    """(nx, ny_global, Lx, Ly_global, ny_per_gpu, strong) of a workload on N GPUs."""
    preset, solver, nx, ny_gpu, Lx, Ly_gpu, vcycles, tol, iters, cfg_idx = WORKLOADS[name]
    strong = name.endswith("_strong")
    if strong:
        return nx, ny_gpu, Lx, Ly_gpu, ny_gpu // N, True
    return nx, ny_gpu * N, Lx, Ly_gpu * N, ny_gpu, False


class CaptureStderr:
    """fd 2 into a temp file for the length of the block (the reference reports sanitised values on stderr only)."""

    def __enter__(self):
        import tempfile
        sys.stderr.flush()
        self.saved, self.tmp = os.dup(2), tempfile.TemporaryFile()
        os.dup2(self.tmp.fileno(), 2)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 2)
        os.close(self.saved)
        self.tmp.seek(0)
        self.text = self.tmp.read().decode(errors="replace")
        self.tmp.close()


def cpu_reference_run(name, N, steps, warmup):
    """The reference's OpenMP solver on the host cores: the workload's OWN preset, solver settings and grid (the global
    grid of the N-GPU run when it fits comfortably into host memory, else one GPU's slab of it -- stated in `sample`).
    Like the GPU arm, `mg` workloads are timed in chunks that start from the reset state and end before the first step
    in which the reference reports a non-finite value (its mg mode diverges by design).  Only the step count is bounded."""
    import numpy as np
    ncores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    from oracle import cfdo
    kind = "ref" if cfdo.have_ref() else "port"
    lib = cfdo.load(kind)
    preset, solver, _, ny_1, _, Ly_1, vcycles, tol, iters, _ = WORKLOADS[name]
    nx, ny, Lx, Ly, ny_gpu, strong = grid_of(name, N)
    note = ""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    need = 14 * 8 * (nx + 3) * (ny + 3)  # 13 fp64 fields of State / PressureSolver / TimeIntegrator
    if N > 1 and not strong and need > 0.4 * avail:
        ny, Ly = ny_1, Ly_1
        note = f" (one GPU's slab of the {N}-GPU grid: the global grid needs {need / 2**30:.0f} GiB of host memory)"
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(lib, nx, ny, Lx, Ly)

    def reset():  # the drivers' reset (main.cpp:113-126, :338-358)
        for f in (sim.u, sim.v, sim.p, sim.rhs):
            f[...] = 0.0
        lib.L.cfdo_clear_work(sim.h)
        if preset == "shear":
            cfdo.shear_init(sim.u, nx, ny, Ly)

    reset()
    t0 = time.perf_counter()
    sim.step(bc, Re, CFL, pp)
    t_step = time.perf_counter() - t0
    # bound the sample: about a minute and a half of CPU work (a PCG step at 8192^2 is ~10 s), at least 2 timed steps
    budget = float(os.environ.get("CFD_BENCH_CPU_SECONDS", "90"))
    steps_run = max(2, min(steps, int(budget / max(t_step, 1e-3))))
    chunk, probe = steps_run, None
    if solver == "mg" and preset != "shear":
        # finite-regime chunk length, probed from the reset state like the GPU arm does; the reference's own stderr is the
        # detector ("Non-finite values detected ..." time_integrator.cpp:23, "Non-finite pressure residual" pressure.cpp:125)
        reset()
        clean, said = 0, ""
        for _ in range(8):
            with CaptureStderr() as cap:
                sim.step(bc, Re, CFL, pp)
            if "on-finite" in cap.text or not math.isfinite(sim.last_residual) or sim.last_residual == 1e9:
                said = cap.text.strip().splitlines()[0] if cap.text.strip() else "residual not finite"
                break
            clean += 1
        chunk = max(2, clean)
        probe = f"probe from reset: {clean} clean steps, then the reference said '{said}'" if said else f"probe: {clean} clean steps"
    else:
        for _ in range(max(0, min(warmup, steps_run // 2) - 1)):  # the first step above was a warm-up step too
            sim.step(bc, Re, CFL, pp)
    total, left = 0.0, steps_run
    while left > 0:
        n = min(chunk, left)
        if probe:
            reset()
        t0 = time.perf_counter()
        for _ in range(n):
            sim.step(bc, Re, CFL, pp)
        total += time.perf_counter() - t0
        left -= n
    sim.close()
    return {"value": nx * ny * steps_run / total, "unit": "cell-updates/s", "cores": ncores,
            "kind": "reference" if kind == "ref" else "port", "grid": f"{nx}x{ny}", "steps_run": steps_run,
            "sample": f"{preset} {solver} {nx}x{ny}{note}, {steps_run} timed steps"
                      + (f" in chunks of {chunk} from the reset state ({probe})" if probe else "")
                      + f", OMP_NUM_THREADS={ncores}",
            "seconds": total}


class Ctx:
    """What every workload of one bench run shares: the ranks, torch.distributed, the package."""

    def __init__(self, N):
        import torch
        self.torch = torch
        self.N = N
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != N:
            raise SystemExit(f"--gpus {N} needs {N} ranks (torchrun --nproc-per-node {N}); WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.cfd = importlib.import_module("macos-cfd_b200")
        self.dist = None
        if N > 1:  # one process per GPU; torch.distributed (NCCL) is the plumbing: id broadcast, barriers, max over ranks
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def comm_id(self):
        if self.N == 1:
            return None
        ids = [self.cfd.comm_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.N > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.N == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())


def owned_rows(r, N, nrows):
    """Row ranges (local, inside a slab's host arrays u / v / p) that slab r of N owns: its cell rows, the shared v face
    row at its bottom, and the domain's ghost rows on the end slabs.  Same rule as tests/slab_worker.py."""
    b, t = r == 0, r == N - 1
    return {"u": (0 if b else 1, nrows + 2 if t else nrows + 1), "v": (0 if b else 1, nrows + 3 if t else nrows + 1),
            "p": (0 if b else 1, nrows + 2 if t else nrows + 1)}


def slab_parity(ctx, name, s, reset, bc, Re, CFL, pp, steps=2):
    """N row slabs against ONE GPU running the same global grid: `steps` steps from the reset state on both, then every
    rank compares the rows it owns (domain ghost rows included on the end slabs) with the same rows of the 1-GPU State,
    which rank 0 computes on its own GPU and sends over NCCL.  Bitwise for --solver=mg; PCG differs only through the
    summation order of its inner products (north_star: rel-L2 <= 1e-10, iterations +-1)."""
    import numpy as np
    torch, dist, cfd = ctx.torch, ctx.dist, ctx.cfd
    preset, solver = WORKLOADS[name][0], WORKLOADS[name][1]
    nx, ny, Lx, Ly, ny_gpu, strong = grid_of(name, ctx.N)
    reset()
    hist = [s.step(bc, Re, CFL, pp) for _ in range(steps)]
    mine = dict(zip(("u", "v", "p"), s.download()[:3]))
    n, r0 = s.nrows, s.row0
    bottom, top = ctx.rank == 0, ctx.rank == ctx.N - 1

    def owned(r, nrows):
        return owned_rows(r, ctx.N, nrows)

    ref_parts = {}
    ref_hist = None
    if ctx.rank == 0:
        g = cfd.Solver(nx, ny, Lx, Ly, device=ctx.local_rank)
        if preset == "shear":
            g.upload(u=cfd.shear_initial_u(nx, ny, Ly))
        ref_hist = [g.step(bc, Re, CFL, pp) for _ in range(steps)]
        full = dict(zip(("u", "v", "p"), g.download()[:3]))
        g.close()
        for r in range(ctx.N):
            a, nr = cfd.slab_rows(ny, r, ctx.N)
            rng = owned(r, nr)
            parts = {k: np.ascontiguousarray(full[k][a + rng[k][0]:a + rng[k][1]]) for k in full}
            if r == 0:
                ref_parts = parts
            else:
                for k in ("u", "v", "p"):
                    dist.send(torch.from_numpy(parts[k]).cuda(), dst=r)
        del full
    rng = owned(ctx.rank, n)
    bitwise, d2, n2 = True, {}, {}
    for k in ("u", "v", "p"):
        a = torch.from_numpy(np.ascontiguousarray(mine[k][rng[k][0]:rng[k][1]])).cuda()
        if ctx.rank == 0:
            b = torch.from_numpy(ref_parts[k]).cuda()
        else:
            b = torch.empty_like(a)
            dist.recv(b, src=0)
        bitwise = bitwise and bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))
        d2[k], n2[k] = float(((a - b) ** 2).sum().item()), float((b ** 2).sum().item())
        del a, b
    bitwise = ctx.reduce(1.0 if bitwise else 0.0, "min") > 0.5
    rel = {k: math.sqrt(ctx.reduce(d2[k], "sum") / max(ctx.reduce(n2[k], "sum"), 1e-300)) for k in ("u", "v", "p")}
    out = {"vs": "1gpu", "grid": f"{nx}x{ny}", "steps": steps, "bitwise": bitwise, "rel_l2": rel,
           "bar": "bitwise" if solver == "mg" else "rel-L2 <= 1e-10, iterations +-1"}
    if ctx.rank == 0:
        out["dt_equal"] = all(a.dt == b.dt for a, b in zip(hist, ref_hist)) if solver == "mg" else \
            all(abs(a.dt - b.dt) <= 1e-12 * b.dt for a, b in zip(hist, ref_hist))
        out["iters"] = [[a.iters, b.iters] for a, b in zip(hist, ref_hist)]
        ok_iters = all(abs(a.iters - b.iters) <= 1 for a, b in zip(hist, ref_hist))
        out["ok"] = bool(out["dt_equal"] and (bitwise if solver == "mg" else (max(rel.values()) <= 1e-10 and ok_iters)))
    return out


def run_workload(ctx, name, steps, warmup, variant=1, want_e2e=True, want_parity=True, want_solves=False):
    """Measure one workload; returns the fields of its JSON entry (identical on every rank up to rank-0-only details)."""
    import numpy as np
    torch, cfd, N = ctx.torch, ctx.cfd, ctx.N
    preset, solver, _, _, _, _, vcycles, tol, iters, cfg_idx = WORKLOADS[name]
    nx, ny, Lx, Ly, ny_gpu, strong = grid_of(name, N)
    config = {"workload": name, "baseline_config": cfg_idx, "preset": preset, "solver": solver,
              "grid": f"{nx}x{ny}", "domain": f"{Lx}x{Ly}", "vcycles": vcycles, "tol": tol, "pcg_max_iters": iters,
              "decomposition": f"{N} row slab(s) of {nx}x{ny_gpu}",
              "l2": "fields (>= 0.5 GB each) exceed the 126 MB L2" if nx * ny_gpu * 8 > 200e6 else
              "L2 flushed (256 MB write) before every timed step",
              "timing": "CUDA events on the solver's stream, max over ranks" +
                        ("; mg diverges by design, so the K steps are timed in chunks that start from the reset state and "
                         "stay in the regime where the reference's residual is finite" if solver == "mg" else "")}
    bc, Re, CFL = cfd.preset(preset, Ly)
    pp = cfd.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    s = cfd.Solver(nx, ny, Lx, Ly, device=ctx.local_rank, rank=ctx.rank, nranks=N, comm_id=ctx.comm_id())
    s.set_variant(variant)
    config["comm"] = {"single": "none (one GPU)", "nccl": "NCCL send/recv rows + allreduce",
                      "p2p": "peer memory over NVLink (CUDA IPC): halo rows stored into the neighbour's arrays, "
                             "reductions fused with their scalar tails"}[s.comm_mode]
    sh = s.shapes()
    stream = torch.cuda.ExternalStream(s.stream, device=ctx.local_rank)
    flush = None if nx * ny_gpu * 8 > 200e6 else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cells = nx * ny
    zero = {k: np.zeros(v) for k, v in sh.items()}
    u_init = (np.ascontiguousarray(cfd.shear_initial_u(nx, ny, Ly)[s.row0:s.row0 + s.nrows + 2]) if preset == "shear"
              else zero["u"])

    def reset():
        """GUI-style reset (main.cpp:338-358): all fields and work arrays back to the deterministic initial state.
        The reference's `mg` mode diverges to inf within ~15-30 steps, so every timed region
        starts from this state and stays in the regime where the reference itself is finite."""
        s.upload(u_init, zero["v"], zero["p"], rhs=zero["p"])
        s.clear_work()
        s.sync()

    def dev_reset():  # the same state without host traffic (shear needs its initial profile: host upload)
        if preset == "shear":
            reset()
        else:
            s.reset()

    for _ in range(warmup):
        s.step_async(bc, Re, CFL, pp)
    s.sync()
    reset()

    # ---- device-resident timing: K steps, CUDA events on the launching stream
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    if N > 1:
        s.wait_stats(reset=True)
    l0 = s.launch_count
    ctx.barrier()
    # The reference's `mg` mode diverges: at 8192^2 the pressure residual is non-finite by step ~6-8.
    # From then on the guarded slow paths run (on the GPU as in the reference) and the timing no longer describes the
    # solver.  mg workloads are therefore timed in chunks that start from the reset state and end before the first step
    # whose residual or divergence is non-finite (probed here, identical on every rank: both are global reductions);
    # the device-side resets sit outside the timed events.
    chunk = steps
    if solver == "mg" and preset != "shear":
        s.reset()
        clean, probe_ms = 0, []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            i = s.step(bc, Re, CFL, pp)
            e1.record(stream)
            e1.synchronize()
            probe_ms.append(round(e0.elapsed_time(e1), 4))
            bad = bool(i.nonfinite) or not math.isfinite(i.div_after) or not math.isfinite(i.div_before)
            # the sanitise flag is rank-local: every rank must leave the probe at the same step
            if ctx.reduce(1.0 if bad else 0.0) > 0.0:
                break
            clean += 1
        config["probe_ms_per_step_from_reset"] = probe_ms
        chunk = max(2, clean)
        config["timing"] += f" (chunk = {chunk} steps)"
        s.reset()
        s.sync()
    # Everything is enqueued without a host round trip in between: one CUDA-event pair per chunk (or, on L2-resident grids,
    # per step, with the L2 flush in front of it and outside it); the device-side resets between chunks sit outside the
    # pairs too.  barrier + synchronize bracket the whole region; the time is the sum of the pairs, max over ranks.
    pairs = []
    ctx.barrier()
    if flush is None:
        left = steps
        while left > 0:
            n = min(chunk, left)
            if chunk < steps:
                s.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                s.step_async(bc, Re, CFL, pp)
            e1.record(stream)
            pairs.append((e0, e1))
            left -= n
    else:
        for i in range(steps):
            if chunk < steps and i > 0 and i % chunk == 0:
                s.reset()
            with torch.cuda.stream(stream):
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.step_async(bc, Re, CFL, pp)
            e1.record(stream)
            pairs.append((e0, e1))
    info = s.sync()
    ctx.barrier()
    total_ms = ctx.reduce(sum(a.elapsed_time(b) for a, b in pairs))
    launches = s.launch_count - l0
    if N > 1:  # where this rank polled for its neighbours during the timed steps (+ the probe), per step
        config["peer_wait_us_per_step_rank0"] = {k: [round(v[0] / max(steps, 1), 2), v[1]] for k, v in s.wait_stats().items()}
    clocks = sampler.stop()
    ms_per_step = total_ms / steps
    value = cells * steps / (total_ms * 1e-3)

    # ---- per-stage CUDA-event times of a step (live, same run) -> roofline of the dominant stage
    dev_reset()
    s.set_profiling(True)
    acc = {}
    nprof = max(2, min(steps, chunk, 5))
    for _ in range(nprof):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        pinfo = s.step(bc, Re, CFL, pp)
        for k, v in s.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / nprof
    s.set_profiling(False)
    peak, peak_src = measured_peaks()
    stage_bytes = dict(STAGE_BYTES, pressure_solve=solve_bytes(solver, vcycles, pinfo.iters))
    stages = {}
    cells_rank = nx * s.nrows  # stage times are this rank's: its kernels move this rank's cells
    for k, ms in acc.items():
        b = stage_bytes.get(k, 0) * cells_rank
        stages[k] = {"ms": round(ms, 4), "gbs": round(b / (ms * 1e-3) / 1e9, 1) if ms > 0 and b else None}
    top = max((k for k in acc if stage_bytes.get(k)), key=lambda k: acc[k])
    if top == "pressure_solve":  # the stage is `launches` identical kernel launches (mg) / iterations (PCG)
        launches_top = max(pinfo.iters, 1)
        kernel = s.last_smoother if solver == "mg" else s.last_pcg
        bytes_launch = (24 if solver == "mg" else 64) * cells_rank
    else:
        launches_top, kernel, bytes_launch = 1, STAGE_KERNEL[top], stage_bytes[top] * cells_rank
    ms_launch = acc[top] / launches_top
    achieved = bytes_launch / (ms_launch * 1e-3) / 1e9
    step_bytes = sum(stage_bytes.values()) * cells_rank
    survey_bytes = SURVEY_STEP_BYTES[solver](vcycles, pinfo.iters) * cells_rank
    traffic, traffic_src = ncu_traffic(kernel, nx, s.nrows)
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_launch": bytes_launch, "ms_per_launch": ms_launch, "launches_per_step": launches_top,
                "note": "achieved = algorithmic bytes of the kernel as implemented / CUDA-event time of this run; traffic = "
                        "ncu dram bytes per launch (profiles/ncu_traffic.json), null off the profiled grid",
                "whole_step": {"bytes_per_cell": sum(stage_bytes.values()),
                               "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "survey_bytes_per_cell": SURVEY_STEP_BYTES[solver](vcycles, pinfo.iters),
                               "survey_frac": survey_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                "stages": stages}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if want_e2e:
        hu = torch.zeros(sh["u"], dtype=torch.float64).pin_memory()
        hv = torch.zeros(sh["v"], dtype=torch.float64).pin_memory()
        hp = torch.zeros(sh["p"], dtype=torch.float64).pin_memory()
        hr = torch.zeros(sh["p"], dtype=torch.float64).pin_memory()
        reset()
        s.download(hu, hv, hp, hr)
        ke = max(2, min(steps, 4))
        up_p = None if solver == "pcg" else hp  # PCG starts from p = 0 (pressure.cpp:133-134): the host p is not read
        for _ in range(1):
            s.upload(hu, hv, up_p); s.step(bc, Re, CFL, pp); s.download(hu, hv, hp, hr)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            s.upload(hu, hv, up_p)
            s.step(bc, Re, CFL, pp)
            s.download(hu, hv, hp, hr)
        torch.cuda.synchronize()
        t = ctx.reduce(time.perf_counter() - t0)
        h2d = (hu.numel() + hv.numel() + (0 if up_p is None else hp.numel())) * 8
        d2h = h2d + hr.numel() * 8
        e2e = {"value": cells * ke / t, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d * N, "d2h_bytes_per_step": d2h * N,
               "steps": ke, "ms_per_step": 1e3 * t / ke,
               "path": "cfd_b200_upload -> cfd_b200_step -> cfd_b200_download (pinned host State)"}
        del hu, hv, hp, hr

    parity = None
    if N > 1 and want_parity:
        parity = slab_parity(ctx, name, s, reset, bc, Re, CFL, pp)
    solves = None
    if want_solves:  # BASELINE configs[4]: "PCG vs MG pressure-solve time" -- the same right-hand side, from p = 0
        reset()
        for _ in range(2):
            s.step(bc, Re, CFL, pp)
        rhs = s.download()[3]
        solves = []
        for label, sp in (("mg as the reference runs it: 2 x smooth() on the fine grid", cfd.make_params("mg", vcycles=2, tol=0.0)),
                          ("Jacobi-PCG, 200 iterations", cfd.make_params("pcg", tol=0.0, pcg_max_iters=200)),
                          ("true V-cycle (extension, parity unpinned): 2 x V(2,2), 8 levels",
                           cfd.make_params("mg", vcycles=2, tol=0.0, mg_levels=8, rbgs_sweeps=2, mg_mode=cfd.MG_VCYCLE))):
            best, res, it = None, None, None
            for _ in range(2):
                s.upload(p=zero["p"], rhs=rhs)
                ctx.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                res, it = s.pressure_solve(sp)
                e1.record(stream)
                e1.synchronize()
                t = ctx.reduce(e0.elapsed_time(e1))
                best = t if best is None else min(best, t)
            solves.append({"solver": label, "ms": round(best, 3), "residual": res, "iterations": it})
    s.close()
    del flush
    torch.cuda.empty_cache()
    ctx.barrier()
    out = {"value": value, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup, "scaling": "strong" if strong else "weak",
           "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
           "last_step": {"dt": info.dt, "residual": info.residual, "iters": info.iters}}
    if parity is not None:
        out["parity"] = parity
    if solves is not None:
        out["pressure_solve_times"] = solves
    return out


def also_list(name, N):
    """BASELINE.json's other configs, measured (short) next to the default workload: (workload, steps)."""
    if name != DEFAULT_WORKLOAD:
        return []
    out = [("jet_pcg_8192", 3)]
    if N > 1:
        out.append(("lid_mg_8192_strong", 10))
    if N == 2:
        out.append(("shear_pcg_4096_strong", 3))
    if N == 1:
        out += [("lid_mg_8192", 10), ("shear_pcg_4096", 3), ("jet_mg_1024", 20), ("lid_pcg_512x256", 10)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--variant", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    N = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    nx, ny, Lx, Ly, ny_gpu, strong = grid_of(args.workload, N)
    base = {"metric": "cell-updates/s (full timestep incl. Poisson)", "unit": "cell-updates/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (deterministic preset initial/boundary data)"}

    if args.impl == "reference":
        if rank != 0:
            return
        preset, solver, _, _, _, _, vcycles, tol, iters, cfg_idx = WORKLOADS[args.workload]
        r = cpu_reference_run(args.workload, N, args.steps, args.warmup)
        config = {"workload": args.workload, "baseline_config": cfg_idx, "preset": preset, "solver": solver,
                  "grid": f"{nx}x{ny}", "grid_ran": r["grid"], "domain": f"{Lx}x{Ly}", "vcycles": vcycles, "tol": tol,
                  "pcg_max_iters": iters, "steps_ran": r["steps_run"],
                  "timing": "host wall clock around the reference's TimeIntegrator::step calls; " + r["sample"]}
        line = dict(base, impl="reference", config=config, value=r["value"], ms_per_step=1e3 * r["seconds"] / r["steps_run"],
                    cpu_baseline={k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    e2e={"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
        print(json.dumps(line))
        return

    ctx = Ctx(N)
    main_r = run_workload(ctx, args.workload, args.steps, args.warmup, args.variant, not args.no_e2e, not args.no_parity,
                          want_solves=not args.no_also and args.workload == DEFAULT_WORKLOAD)
    line = dict(base, **main_r)
    if not args.no_also:
        also = []
        for name, k in also_list(args.workload, N):
            r = run_workload(ctx, name, k, 3, args.variant, want_e2e=False, want_parity=not args.no_parity)
            rf = r["roofline"]
            e = {"workload": name, "grid": r["config"]["grid"], "decomposition": r["config"]["decomposition"], "scaling": r["scaling"],
                 "steps": k, "warmup": 3, "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": "cell-updates/s",
                 "gpu_launches": r["gpu_launches"], "last_step": r["last_step"],
                 "roofline": {k2: rf[k2] for k2 in ("kernel", "achieved", "peak", "frac", "traffic", "ms_per_launch", "launches_per_step")},
                 "whole_step_frac": rf["whole_step"]["frac"], "stages": rf["stages"],
                 "peer_wait_us_per_step_rank0": r["config"].get("peer_wait_us_per_step_rank0")}
            if "parity" in r:
                e["parity"] = r["parity"]
            also.append(e)
        line["also"] = also
    if N > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if rank != 0:
        return
    if not args.no_cpu_baseline and N == 1:
        r = cpu_reference_run(args.workload, 1, 2, 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
