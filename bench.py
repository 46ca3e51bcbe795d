#!/usr/bin/env python
"""bench.py -- cell-updates/s of the full time step (incl. the Poisson solve) + HBM-roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A "step" is one TimeIntegrator::step over the whole synthetic grid.  ONE JSON line goes to stdout.

  value        whole-job cell-updates/s, State resident in HBM, K steps enqueued back to back, device-timed
  e2e          the same metric through the C ABI with HOST buffers: every step uploads u, v, p from pinned host memory,
               runs the step and downloads u, v, p, rhs (what a GUI-style caller of TimeIntegrator::step(State&) sees)
  roofline     the dominant stage of the step: algorithmic bytes / live CUDA-event time, against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the reference's own OpenMP solver (oracle/_ref, else the C port) on this box's host cores, bounded sample

Workloads are BASELINE.json's configs on synthetic (deterministic preset) data; the default is the one its metric and
target are quoted on: jet plume, --solver=mg --vcycles=2, 8192^2 fp64 per GPU (weak scaling over row slabs).
"""
from __future__ import annotations

import argparse
import importlib
import json
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
    # BASELINE configs[3]: STRONG scaling -- the 8192^2 grid is fixed and split into N row slabs
    "lid_mg_8192_strong": ("lid", "mg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 3),
}
DEFAULT_WORKLOAD = "jet_mg_8192"

# Algorithmic HBM bytes per cell of each stage AS IMPLEMENTED (fp64).  SURVEY.md §8d's figures are the same
# except for the pressure solve, where the implemented kernels need less than the SURVEY's formulation:
#   mg  : k_rbgs_fused moves 24 B/cell per smooth() launch (R p, R rhs, W p; SURVEY's per-pass accounting: 64)
#   PCG : phase 1 + phase 2 move 24 + 40 = 64 B/cell per iteration (Ap recomputed, z never stored; SURVEY: 80), + 24 init
# The CFL maxima (16 B/cell in SURVEY) ride on the post-projection divergence kernel and cost no pass of their own.
STAGE_BYTES = {"cfl_dt": 0, "rhs_stage1": 32, "rhs_stage2": 48, "divergence": 24, "project": 40, "divergence_post": 24}
SURVEY_STEP_BYTES = {"mg": lambda vcycles, iters: 184 + 64 * vcycles, "pcg": lambda vcycles, iters: 184 + 32 + 80 * iters}
STAGE_KERNEL = {"rhs_stage1": "k_rhs_march<1>", "rhs_stage2": "k_rhs_march<2>", "divergence": "k_divergence<1>",
                "project": "k_project", "divergence_post": "k_divergence<0>"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/r01_*, r01d_* for k_rbgs_stream), 8192 x 8192 only
NCU_TRAFFIC_8192 = {"k_rbgs_fused": 1.0816e9 + 0.5221e9, "k_rbgs_stream": 1.1235e9 + 0.5104e9, "k_rhs_march<1>": 1.1385e9 + 1.0216e9,
                    "k_rhs_march<2>": 2.1998e9 + 1.0244e9, "k_project": 1.6301e9 + 1.0227e9,
                    "k_divergence<1>": 1.0831e9 + 0.5114e9, "k_divergence<0>": 1.0832e9 + 0.5101e9,
                    "k_pcg_phase1_march + k_pcg_phase2_march (one iteration)": 1.112e9 + 0.510e9 + 1.629e9 + 1.022e9}


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


def cpu_reference_run(wl, steps, warmup, sample_n):
    """The reference's OpenMP solver on the host cores, on a bounded sample (a sample_n^2 grid of the same preset
    and solver settings; cell-updates/s is size-independent once the fields leave the CPU caches)."""
    ncores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    from oracle import cfdo
    kind = "ref" if cfdo.have_ref() else "port"
    lib = cfdo.load(kind)
    preset, solver, _, _, Lx, Ly, vcycles, tol, iters, _ = wl
    nx = ny = sample_n
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(lib, nx, ny, Lx, Ly)
    if preset == "shear":
        cfdo.shear_init(sim.u, nx, ny, Ly)
    for _ in range(warmup):
        sim.step(bc, Re, CFL, pp)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step(bc, Re, CFL, pp)
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": nx * ny * steps / dt, "unit": "cell-updates/s", "cores": ncores,
            "kind": "reference" if kind == "ref" else "port",
            "sample": f"{preset} {solver} {nx}x{ny}, {steps} steps after {warmup} warm-up, OMP_NUM_THREADS={ncores}",
            "seconds": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    wl = WORKLOADS[args.workload]
    preset, solver, nx, ny_gpu, Lx, Ly_gpu, vcycles, tol, iters, cfg_idx = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    strong = args.workload.endswith("_strong")
    if strong:
        ny, Ly = ny_gpu, Ly_gpu          # strong scaling: the global grid is fixed
        ny_gpu = ny // N
    else:
        ny, Ly = ny_gpu * N, Ly_gpu * N  # weak scaling: row slabs stacked in y
    config = {"workload": args.workload, "baseline_config": cfg_idx, "preset": preset, "solver": solver,
              "grid": f"{nx}x{ny}", "domain": f"{Lx}x{Ly}", "vcycles": vcycles, "tol": tol, "pcg_max_iters": iters,
              "decomposition": f"{N} row slab(s) of {nx}x{ny_gpu}",
              "l2": "fields (>= 0.5 GB each) exceed the 126 MB L2" if nx * ny_gpu * 8 > 200e6 else
              "L2 flushed (256 MB write) before every timed step",
              "timing": "CUDA events on the solver's stream, max over ranks" +
                        ("; mg diverges by design, so the K steps are timed in chunks that start from the reset state and "
                         "stay in the regime where the reference's residual is finite" if solver == "mg" else "")}
    base = {"metric": "cell-updates/s (full timestep incl. Poisson)", "unit": "cell-updates/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (deterministic preset initial/boundary data)",
            "config": config}

    if args.impl == "reference":
        if rank != 0:
            return
        sample_n = 2048 if solver == "mg" else 768
        r = cpu_reference_run(wl, args.steps, args.warmup, sample_n)
        line = dict(base, impl="reference", value=r["value"], ms_per_step=1e3 * r["seconds"] / args.steps,
                    cpu_baseline={k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    e2e={"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if world != N:
        raise SystemExit(f"--gpus {N} needs {N} ranks (torchrun --nproc-per-node {N}); WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    cfd = importlib.import_module("macos-cfd_b200")
    comm = None
    if N > 1:  # one process per GPU; torch.distributed (NCCL) is the plumbing: id broadcast, barriers, max over ranks
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [cfd.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = ids[0]
    bc, Re, CFL = cfd.preset(preset, Ly)
    pp = cfd.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    s = cfd.Solver(nx, ny, Lx, Ly, device=local_rank, rank=rank, nranks=N, comm_id=comm)
    s.set_variant(args.variant)
    config["comm"] = {"single": "none (one GPU)", "nccl": "NCCL send/recv rows + allreduce",
                      "p2p": "peer memory over NVLink (CUDA IPC): halo rows stored into the neighbour's arrays, "
                             "reductions fused with their scalar tails"}[s.comm_mode]
    sh = s.shapes()

    def barrier():
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    stream = torch.cuda.ExternalStream(s.stream, device=local_rank)
    flush = None if nx * ny_gpu * 8 > 200e6 else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cells = nx * ny

    import numpy as np
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

    for _ in range(args.warmup):
        s.step_async(bc, Re, CFL, pp)
    s.sync()
    reset()

    # ---- device-resident timing: K steps, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = s.launch_count
    barrier()
    # The reference's `mg` mode diverges: at 8192^2 the pressure residual is non-finite by step ~6-8.
    # From then on the guarded slow paths run (on the GPU as in the reference) and the timing no longer describes the
    # solver.  mg workloads are therefore timed in chunks that start from the reset state and end before the first step
    # whose residual or divergence is non-finite (probed here, identical on every rank: both are global reductions);
    # the device-side resets sit outside the timed events.
    chunk = args.steps
    if solver == "mg" and preset != "shear":
        import math
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
            if max_over_ranks(1.0 if bad else 0.0) > 0.0:
                break
            clean += 1
        config["probe_ms_per_step_from_reset"] = probe_ms
        chunk = max(2, clean)
        config["timing"] += f" (chunk = {chunk} steps)"
        s.reset()
        s.sync()
    if flush is None:
        total_ms, left = 0.0, args.steps
        while left > 0:
            n = min(chunk, left)
            if chunk < args.steps:
                s.reset()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                s.step_async(bc, Re, CFL, pp)
            e1.record(stream)
            info = s.sync()
            barrier()
            total_ms += max_over_ranks(e0.elapsed_time(e1))
            left -= n
    else:
        total_ms = 0.0
        for i in range(args.steps):
            if chunk < args.steps and i > 0 and i % chunk == 0:
                s.reset()
            with torch.cuda.stream(stream):
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.step_async(bc, Re, CFL, pp)
            e1.record(stream)
            info = s.sync()
            barrier()
            total_ms += max_over_ranks(e0.elapsed_time(e1))
    launches = s.launch_count - l0
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = cells * args.steps / (total_ms * 1e-3)

    # ---- per-stage CUDA-event times of a step (live, same run) -> roofline of the dominant stage
    reset()
    s.set_profiling(True)
    acc = {}
    nprof = max(3, min(args.steps, 5))
    for _ in range(nprof):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        s.step(bc, Re, CFL, pp)
        for k, v in s.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / nprof
    s.set_profiling(False)
    peak, peak_src = measured_peaks()
    stage_bytes = dict(STAGE_BYTES, pressure_solve=solve_bytes(solver, vcycles, info.iters))
    stages = {}
    cells_rank = nx * s.nrows  # stage times are this rank's: its kernels move this rank's cells
    for k, ms in acc.items():
        b = stage_bytes.get(k, 0) * cells_rank
        stages[k] = {"ms": round(ms, 4), "gbs": round(b / (ms * 1e-3) / 1e9, 1) if ms > 0 and b else None}
    top = max((k for k in acc if stage_bytes.get(k)), key=lambda k: acc[k])
    if top == "pressure_solve":  # the stage is `launches` identical kernel launches (mg) / iterations (PCG)
        launches_top = max(info.iters, 1)
        kernel = s.last_smoother if solver == "mg" else "k_pcg_phase1_march + k_pcg_phase2_march (one iteration)"
        bytes_launch = (24 if solver == "mg" else 64) * cells_rank
    else:
        launches_top, kernel, bytes_launch = 1, STAGE_KERNEL[top], stage_bytes[top] * cells_rank
    ms_launch = acc[top] / launches_top
    achieved = bytes_launch / (ms_launch * 1e-3) / 1e9
    step_bytes = sum(stage_bytes.values()) * cells_rank
    survey_bytes = SURVEY_STEP_BYTES[solver](vcycles, info.iters) * cells_rank
    traffic = NCU_TRAFFIC_8192.get(kernel) if (nx, s.nrows) == (8192, 8192) else None
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": bytes_launch, "ms_per_launch": ms_launch, "launches_per_step": launches_top,
                "note": "achieved = algorithmic bytes of the kernel as implemented / CUDA-event time of this run; traffic = "
                        "ncu dram bytes per launch (profiles/), null off the profiled grid",
                "whole_step": {"bytes_per_cell": sum(stage_bytes.values()),
                               "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "survey_bytes_per_cell": SURVEY_STEP_BYTES[solver](vcycles, info.iters),
                               "survey_frac": survey_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                "stages": stages}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hu = torch.zeros(sh["u"], dtype=torch.float64).pin_memory()
        hv = torch.zeros(sh["v"], dtype=torch.float64).pin_memory()
        hp = torch.zeros(sh["p"], dtype=torch.float64).pin_memory()
        hr = torch.zeros(sh["p"], dtype=torch.float64).pin_memory()
        reset()
        s.download(hu, hv, hp, hr)
        ke = max(2, min(args.steps, 4))
        for _ in range(1):
            s.upload(hu, hv, hp); s.step(bc, Re, CFL, pp); s.download(hu, hv, hp, hr)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            s.upload(hu, hv, hp)
            s.step(bc, Re, CFL, pp)
            s.download(hu, hv, hp, hr)
        torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        h2d = (hu.numel() + hv.numel() + hp.numel()) * 8
        d2h = h2d + hr.numel() * 8
        e2e = {"value": cells * ke / t, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d * N, "d2h_bytes_per_step": d2h * N,
               "steps": ke, "ms_per_step": 1e3 * t / ke,
               "path": "cfd_b200_upload -> cfd_b200_step -> cfd_b200_download (pinned host State)"}
        del hu, hv, hp, hr

    line = dict(base, value=value, ms_per_step=ms_per_step, clocks=clocks, e2e=e2e, gpu_launches=launches,
                roofline=roofline, last_step={"dt": info.dt, "residual": info.residual, "iters": info.iters})
    s.close()
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if not args.no_cpu_baseline and N == 1:
        r = cpu_reference_run(wl, 2, 1, 4096 if solver == "mg" else 1024)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
