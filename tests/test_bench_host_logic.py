"""CPU coverage of bench.py's host logic: workload grids per GPU count, the also-list, the ownership rule of the in-bench
N-slab-vs-1-GPU parity check, the ncu traffic file, the stderr capture of the reference arm, and the reference arm itself
on the reference's own default grid (bounded to a few steps)."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bench = importlib.import_module("bench")


def test_workload_grids_weak_and_strong():
    assert bench.grid_of("jet_mg_8192", 1) == (8192, 8192, 1.0, 1.0, 8192, False)
    assert bench.grid_of("jet_mg_8192", 8) == (8192, 65536, 1.0, 8.0, 8192, False)   # weak: slabs stacked in y
    assert bench.grid_of("lid_mg_8192_strong", 8) == (8192, 8192, 1.0, 1.0, 1024, True)  # strong: the grid is fixed
    names = [n for n, _ in bench.also_list("jet_mg_8192", 1)]
    assert "jet_pcg_8192" in names and "lid_pcg_512x256" in names and "jet_mg_1024" in names
    names8 = [n for n, _ in bench.also_list("jet_mg_8192", 8)]
    assert names8 == ["jet_pcg_8192", "lid_mg_8192_strong"]
    assert bench.also_list("jet_pcg_8192", 4) == []  # only the default workload carries the list


def test_owned_rows_tile_the_global_arrays_exactly():
    cfd = importlib.import_module("macos-cfd_b200")
    nx, ny = 7, 203
    full = {"u": np.arange((ny + 2) * (nx + 3), dtype=np.float64).reshape(ny + 2, nx + 3),
            "v": np.arange((ny + 3) * (nx + 2), dtype=np.float64).reshape(ny + 3, nx + 2),
            "p": np.arange((ny + 2) * (nx + 2), dtype=np.float64).reshape(ny + 2, nx + 2)}
    for N in (2, 3, 4, 8):
        parts = {k: [] for k in full}
        for r in range(N):
            a, n = cfd.slab_rows(ny, r, N)
            local = {"u": full["u"][a:a + n + 2], "v": full["v"][a:a + n + 3], "p": full["p"][a:a + n + 2]}  # a slab's host arrays
            rng = bench.owned_rows(r, N, n)
            for k in full:
                parts[k].append(local[k][rng[k][0]:rng[k][1]])
        for k in full:
            assert np.array_equal(np.concatenate(parts[k]), full[k]), (N, k)  # every row exactly once, in order


def test_ncu_traffic_file_is_read_and_off_grid_is_null():
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    nx, ny = (int(x) for x in t["grid"].split("x"))
    some = next(iter(t["kernels"]))
    b, src = bench.ncu_traffic(some, nx, ny)
    assert b == t["kernels"][some]["dram_bytes"] and src.startswith("profiles/")
    assert os.path.exists(os.path.join(ROOT, src.split(" + ")[0]))  # the summary it cites is committed
    assert bench.ncu_traffic(some, nx, ny // 2) == (None, None)
    assert bench.ncu_traffic("no_such_kernel", nx, ny) == (None, None)


def test_capture_stderr_sees_what_c_code_writes():
    import ctypes
    libc = ctypes.CDLL(None)
    with bench.CaptureStderr() as cap:
        libc.dprintf(2, b"Non-finite values detected and replaced with 0\n")
    assert "Non-finite" in cap.text
    with bench.CaptureStderr() as cap:
        pass
    assert cap.text == ""


def test_reference_arm_runs_the_workloads_own_grid():
    env = dict(os.environ, CFD_BENCH_CPU_SECONDS="6", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "lid_pcg_512x256",
                        "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"]["grid"] == line["config"]["grid_ran"] == "512x256"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert "Non-finite" not in r.stderr
