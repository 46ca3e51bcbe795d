"""Row-slab (multi-GPU) runs against the single-GPU run of the same grid and against the oracle.

Needs >= 2 CUDA devices (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); on a 1-GPU
box every test here is skipped.  Bars: the stencil stages and the whole mg path are BIT-IDENTICAL to
one GPU; PCG differs only through the order of its inner-product sums (rel-L2 <= 1e-10, iterations +-1).
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpu():
    import torch
    return torch.cuda.device_count()


def run_slabs(n, tmp_path, **kw):
    out = str(tmp_path / f"slab_{n}.npz")
    args = sum([[f"--{k}", str(v)] for k, v in kw.items()], [])
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + os.getpid() % 500), os.path.join(ROOT, "tests", "slab_worker.py"),
           "--out", out] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return np.load(out)


def run_single(kw):
    cfd = importlib.import_module("macos-cfd_b200")
    bc, Re, CFL = cfd.preset(kw["preset"], kw.get("Ly", 1.0))
    if kw.get("bt_periodic"):
        bc.bottom.type = bc.top.type = cfd.PERIODIC
    s = cfd.Solver(kw["nx"], kw["ny"], kw.get("Lx", 2.0), kw.get("Ly", 1.0))
    s.set_smoother_strip(kw.get("strip", 0))
    if kw["preset"] == "shear":
        s.upload(u=cfd.shear_initial_u(kw["nx"], kw["ny"], kw.get("Ly", 1.0)))
    pp = cfd.make_params(kw["solver"], vcycles=kw.get("vcycles", 2), tol=kw.get("tol", 1e-8), pcg_max_iters=kw.get("iters", 200),
                         mg_levels=kw.get("mg_levels", 3), rbgs_sweeps=kw.get("sweeps", 2), mg_mode=kw.get("mg_mode", 0))
    hist = []
    for _ in range(kw["steps"]):
        i = s.step(bc, Re, CFL, pp)
        hist.append((i.dt, i.residual, i.iters, i.div_before, i.div_after))
    return np.array(hist), s.download(), s.divergence_l2(), s.max_velocity()


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


# (preset, solver, variant, p2p, strip): strip > 0 forces the streaming smoother (automatic only on grids that fill the GPU)
CASES = [(preset, solver, variant, 1, 0) for solver, variant in (("mg", 1), ("mg", 0), ("pcg", 1))
         for preset in ("jet", "lid", "shear")] + [("jet", "mg", 1, 0, 0), ("shear", "pcg", 1, 0, 0), ("lid", "mg", 0, 0, 0)] + [
         ("jet", "mg", 1, 1, 33), ("lid", "mg", 1, 1, 17), ("shear", "mg", 1, 0, 33)]


@pytest.mark.parametrize("preset,solver,variant,p2p,strip", CASES)
def test_slabs_match_single_gpu(tmp_path, preset, solver, variant, p2p, strip):
    """p2p=1: halo rows and reductions through peer memory (CUDA IPC over NVLink); p2p=0: the NCCL path."""
    n = min(ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(preset=preset, solver=solver, nx=260, ny=203, steps=4, iters=60)
    got = run_slabs(n, tmp_path, variant=variant, p2p=p2p, strip=strip, **kw)
    assert str(got["comm_mode"]) == ("p2p" if p2p else "nccl")  # no silent fallback
    hist, (u, v, p, rhs), div_l2, vmax = run_single(kw)
    if solver == "mg":
        assert np.array_equal(got["hist"][:, 0], hist[:, 0])  # dt
    else:  # dt = f(max|u|) inherits the last-bit differences of the PCG inner products (as in test_gpu_parity.compare_run)
        assert np.allclose(got["hist"][:, 0], hist[:, 0], rtol=1e-12, atol=0.0)
    assert abs(float(got["vmax"]) - vmax) <= (0.0 if solver == "mg" else 1e-10 * vmax)
    if solver == "mg":
        for name, a in (("u", u), ("v", v), ("p", p), ("rhs", rhs)):
            assert same(got[name], a), name
        assert np.allclose(got["hist"][:, 1], hist[:, 1], rtol=1e-12)
    else:
        assert np.all(np.abs(got["hist"][:, 2] - hist[:, 2]) <= 1)
        for name, a in (("u", u), ("v", v), ("p", p)):
            assert np.linalg.norm(got[name] - a) <= 1e-10 * np.linalg.norm(a), name
    assert abs(float(got["div_l2"]) - div_l2) <= 1e-9 * max(div_l2, 1e-300)
    assert np.allclose(got["hist"][:, 3:], hist[:, 3:], rtol=1e-9)


@pytest.mark.parametrize("preset,solver,delay_rank", [("shear", "mg", 0), ("jet", "mg", -1), ("lid", "pcg", 0)])
def test_slabs_with_a_late_rank_and_a_comm_reset(tmp_path, preset, solver, delay_rank):
    """One slab is made late before every step (its stream is kept busy for 3 ms, about ten steps of this grid), so its
    neighbours run ahead as far as the flag protocol allows: one-way sends that land before the late slab has even
    started the step, epochs of the next exchange arriving early.  Half-way through, every slab re-arms the protocol
    (cfd_b200_comm_reset).  The result must not notice any of it: mg bit-identical to one GPU, PCG within 1e-10."""
    n = min(ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(preset=preset, solver=solver, nx=260, ny=203, steps=6, iters=40)
    got = run_slabs(n, tmp_path, variant=1, p2p=1, strip=33, **{"delay-rank": delay_rank, "delay-us": 3000, "reset-after": 3}, **kw)
    assert str(got["comm_mode"]) == "p2p"
    hist, (u, v, p, rhs), div_l2, vmax = run_single(dict(kw, strip=33))
    for name, a in (("u", u), ("v", v), ("p", p)):
        if solver == "mg":
            assert same(got[name], a), name
        else:
            assert np.linalg.norm(got[name] - a) <= 1e-10 * np.linalg.norm(a), name


@pytest.mark.parametrize("preset,nx,ny,levels,sweeps,vcycles", [("lid", 256, 192, 5, 2, 3), ("jet", 512, 256, 7, 1, 2),
                                                                ("shear", 128, 128, 3, 3, 4)])
def test_true_vcycle_on_slabs_matches_single_gpu(tmp_path, preset, nx, ny, levels, sweeps, vcycles):
    """mg_mode = VCYCLE (the true geometric V-cycle extension) on row slabs: every level split like the fine grid, the
    one-block coarse part gathered onto every slab -- u, v, p bit-identical to one GPU after whole time steps."""
    n = min(ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(preset=preset, solver="mg", nx=nx, ny=ny, steps=3, vcycles=vcycles, tol=1e-9)
    got = run_slabs(n, tmp_path, variant=1, p2p=1, **{"mg-mode": 1, "mg-levels": levels, "sweeps": sweeps}, **kw)
    hist, (u, v, p, rhs), div_l2, vmax = run_single(dict(kw, mg_mode=1, mg_levels=levels, sweeps=sweeps))
    assert np.array_equal(got["hist"][:, 0], hist[:, 0]) and np.array_equal(got["hist"][:, 2], hist[:, 2])  # dt, cycles
    for name, a in (("u", u), ("v", v), ("p", p), ("rhs", rhs)):
        assert same(got[name], a), name
    assert np.allclose(got["hist"][:, 1], hist[:, 1], rtol=1e-10)


@pytest.mark.parametrize("preset,solver", [("shear", "mg"), ("jet", "mg"), ("shear", "pcg")])
def test_bottom_top_periodic_across_slabs(tmp_path, preset, solver):
    """Both bottom/top Periodic (bc.cpp:171-183, :318-330, :361-379): the pass swaps the boundary rows of the two ends of the
    domain, which live on the first and the last slab -- they trade those rows in front of the kernel.  Bit-identical to one
    GPU (mg) / within the PCG tolerance, with the peer-memory transport for everything else."""
    n = min(ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(preset=preset, solver=solver, nx=260, ny=203, steps=5, iters=50)
    got = run_slabs(n, tmp_path, variant=1, p2p=1, **{"bt-periodic": 1}, **kw)
    hist, (u, v, p, rhs), div_l2, vmax = run_single(dict(kw, bt_periodic=1))
    for name, a in (("u", u), ("v", v), ("p", p)):
        if solver == "mg":
            assert same(got[name], a), name
        else:
            assert np.linalg.norm(got[name] - a) <= 1e-10 * np.linalg.norm(a), name


def test_shapes_and_visualisation_on_slabs(tmp_path):
    """apply_shape_boundary_conditions / compute_scalar / the coloured frame on row slabs (collective calls; shapes in cells
    of the whole grid, statistics and colour scale of the whole grid): fields and bytes identical to one GPU, and a step on
    the masked State still is."""
    n = min(ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import slab_worker
    cfd = importlib.import_module("macos-cfd_b200")
    kw = dict(preset="lid", solver="mg", nx=260, ny=203, steps=4)
    got = run_slabs(n, tmp_path, variant=1, p2p=1, viz=1, **kw)
    bc, Re, CFL = cfd.preset("lid", 1.0)
    pp = cfd.make_params("mg", vcycles=2, tol=1e-8)
    s = cfd.Solver(kw["nx"], kw["ny"], 2.0, 1.0)
    for _ in range(kw["steps"]):
        s.step(bc, Re, CFL, pp)
    s.apply_shapes(slab_worker.viz_shapes(cfd, kw["nx"], kw["ny"]))
    for name in ("speed", "vorticity", "pressure"):
        out, lo, hi, mean, std = s.compute_scalar(cfd.FIELDS[name])
        assert same(got[name], out), name
        st = got[name + "_stats"]
        assert st[0] == lo and st[1] == hi, name
        assert abs(st[2] - mean) <= 1e-12 * abs(mean) + 1e-15 and abs(st[3] - std) <= 1e-10 * std, name
    assert np.array_equal(got["rgb"], s.render_rgb(cfd.FIELDS["speed"], 4))
    s.step(bc, Re, CFL, pp)
    u, v, p, rhs = s.download()
    for name, a in (("u", u), ("v", v), ("p", p)):
        assert same(got[name], a), name
