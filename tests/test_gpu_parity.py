"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star / SURVEY.md §8c):
  * BC, advection, diffusion, RK updates, divergence, projection, dt, and the whole `--solver=mg` path: BIT-EXACT.
  * PCG path: relative L2 <= 1e-10 on u, v, p after N steps, iteration count within +-1 (reduction order differs).
The checker is oracle/cfd_oracle.c (pinned bitwise on the reference's own sources by tests/test_oracle.py) and,
when its prebuilt .so travelled with the snapshot, oracle/_ref itself.
"""
import importlib
import itertools
import os

import numpy as np
import pytest

from oracle import cfdo

pytestmark = pytest.mark.gpu

PCG_TOL = 1e-10  # relative L2, stated by north_star

RICH = dict(left_inflow_u=0.8, jet_center=0.45, jet_width=0.3, left_moving=0.11, right_moving=-0.12,
            bottom_moving=0.13, top_moving=1.0, right_inflow_u=0.21, right_inflow_v=0.22, bottom_inflow_u=0.23,
            bottom_inflow_v=0.24, top_inflow_u=0.25, top_inflow_v=0.26)


@pytest.fixture(scope="module")
def cfd():
    return importlib.import_module("macos-cfd_b200")


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


def same_or_nan(a, b):
    """Bitwise equal, except that a NaN matches any NaN (sign / payload bits of a NaN are implementation-defined
    and differ between x86 and the GPU).  Only for UNSANITISED outputs; the solver path itself never stores a NaN."""
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | both_nan))


def nbad(a, b):
    return int(np.count_nonzero(a.view(np.uint64) != b.view(np.uint64)))


def where_bad(a, b, k=5):
    idx = np.argwhere(a.view(np.uint64) != b.view(np.uint64))
    return [(int(j), int(i), float(a[j, i]), float(b[j, i])) for j, i in idx[:k]]


def rel_l2(a, b):
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


def both_bc(cfd, combo, **kw):
    return cfd.make_bc(*combo, **kw), cfdo.make_bc(*combo, **kw)


def rand_fields(nx, ny, seed):
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(seed)
    return tuple(rng.uniform(-0.5, 0.5, sh[k]) for k in ("u", "v", "p"))


# ------------------------------------------------------------------------------------------------ stages

def test_upload_download_roundtrip(cfd):
    nx, ny = 37, 21
    u, v, p = rand_fields(nx, ny, 1)
    s = cfd.Solver(nx, ny, 2.0, 0.9)
    s.upload(u, v, p, rhs=p * 2)
    u2, v2, p2, r2 = s.download()
    assert same(u, u2) and same(v, v2) and same(p, p2) and same(p * 2, r2)


def test_bc_all_combinations_bit_exact(cfd, port):
    """5^4 side-type combinations x (u, v, p), sentinel arrays, two successive calls (periodic sides swap)."""
    nx, ny, Lx, Ly = 9, 7, 1.1, 0.6
    u, v, p = rand_fields(nx, ny, 7)
    s = cfd.Solver(nx, ny, Lx, Ly)
    bad = []
    for combo in itertools.product(range(5), repeat=4):
        gb, ob = both_bc(cfd, combo, **RICH)
        s.upload(u, v, p)
        ou, ov, op = u.copy(), v.copy(), p.copy()
        for call in range(2):
            s.apply_bc(gb, 7)
            port.apply_bc("u", nx, ny, Lx, Ly, ou, ob)
            port.apply_bc("v", nx, ny, Lx, Ly, ov, ob)
            port.apply_bc("p", nx, ny, Lx, Ly, op, ob)
            gu, gv, gp, _ = s.download()
            if not (same(gu, ou) and same(gv, ov) and same(gp, op)):
                bad.append((combo, call, nbad(gu, ou), nbad(gv, ov), nbad(gp, op)))
    assert not bad, bad[:10]


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("nx,ny,Lx,Ly", [(16, 16, 1.0, 1.0), (37, 21, 2.0, 0.9), (8, 50, 0.3, 1.9), (64, 32, 2.0, 1.0),
                                         (130, 67, 1.0, 1.0), (4, 4, 1.0, 1.0), (300, 211, 1.3, 0.7),
                                         (257, 130, 2.0, 1.0)])
def test_stage_kernels_bit_exact(cfd, port, nx, ny, Lx, Ly, variant):
    u, v, p = rand_fields(nx, ny, nx * 1000 + ny)
    u[3, 2] = 0.0
    v[2, 2] = -0.0
    if nx > 100 and ny > 100:  # zeros, signed zeros, plateaus and huge values inside the marching kernel's region
        u[40:44, 50:70] = 0.0
        v[40:44, 50:70] = -0.0
        u[60:64, 80:90] = 0.25
        v[70, 60:64] = 1e300
        u[80, 70:74] = -1e300
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_variant(variant)
    s.upload(u, v, p)
    # advection + diffusion
    for invRe in (0.0, 1.0 / 50.0):
        du, dv = np.zeros_like(u), np.zeros_like(v)
        port.advect_u(nx, ny, Lx, Ly, u, v, du)
        port.advect_v(nx, ny, Lx, Ly, u, v, dv)
        port.diffuse_u(nx, ny, Lx, Ly, u, du, invRe)  # with invRe = 0 too: -conv + 0*lap turns -0 into +0
        port.diffuse_v(nx, ny, Lx, Ly, v, dv, invRe)
        port.sanitize(du)  # TimeIntegrator::step sanitises the derivatives (time_integrator.cpp:81-82)
        port.sanitize(dv)
        gdu, gdv = s.rhs_terms(invRe)
        assert same(gdu, du), ("du", invRe, nbad(gdu, du), where_bad(gdu, du))
        assert same(gdv, dv), ("dv", invRe, nbad(gdv, dv), where_bad(gdv, dv))
    # compute_cfl
    for Re, CFL, scale in ((50.0, 0.5, 1.0), (1000.0, 0.1, 1e-3), (50.0, 0.01, 0.0)):
        s.upload(u * scale, v * scale)
        assert s.compute_cfl(Re, CFL) == port.compute_cfl(nx, ny, Lx, Ly, u * scale, v * scale, Re, CFL)
    s.upload(u, v, p)
    # divergence
    rhs = np.zeros_like(p)
    port.divergence(nx, ny, Lx, Ly, u, v, rhs)
    s.divergence()
    grhs = s.download()[3]
    assert same(grhs[1:-1, 1:-1], rhs[1:-1, 1:-1]), nbad(grhs, rhs)
    # diagnostics (reductions: tolerance)
    gd, od = s.divergence_l2(), port.divergence_l2(nx, ny, Lx, Ly, u, v)
    assert gd == od or abs(gd - od) <= 1e-13 * abs(od)  # (the 1e300 sentinels make both inf)
    assert s.max_velocity() == port.max_velocity(nx, ny, Lx, Ly, u, v)
    # subtract_grad_p
    u2, v2 = u.copy(), v.copy()
    port.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p, 3e-4)
    s.subtract_grad_p(3e-4)
    gu, gv, gp, _ = s.download()
    assert same(gu, u2) and same(gv, v2) and same(gp, p)
    # advect_* / diffuse_* as separate free functions (no sanitise, diffusion ADDS to its argument)
    s.upload(u, v, p)
    du, dv = np.full_like(u, 9.0), np.full_like(v, 9.0)
    port.advect_u(nx, ny, Lx, Ly, u, v, du)
    port.advect_v(nx, ny, Lx, Ly, u, v, dv)
    gdu, gdv = s.advect()
    assert same_or_nan(gdu, du) and same_or_nan(gdv, dv)
    port.diffuse_u(nx, ny, Lx, Ly, u, du, 0.02)
    port.diffuse_v(nx, ny, Lx, Ly, v, dv, 0.02)
    s.diffuse(0, 0.02, gdu)
    s.diffuse(1, 0.02, gdv)
    assert same_or_nan(gdu, du) and same_or_nan(gdv, dv)


# strip > 0: the streaming smoother with that many rows per strip (it is automatic only on grids that fill the GPU);
# strip = 0 on these sizes: the tile smoother
@pytest.mark.parametrize("strip", [0, 17, 33, 257])
@pytest.mark.parametrize("nx,ny,Lx,Ly", [(16, 16, 1.0, 1.0), (37, 21, 2.0, 0.9), (64, 32, 2.0, 1.0), (130, 67, 1.0, 1.0),
                                         (52, 300, 1.0, 1.0), (53, 40, 1.0, 1.0), (421, 533, 1.7, 0.9)])
def test_pressure_solve_mg_bit_exact(cfd, port, nx, ny, Lx, Ly, strip):
    _, _, p = rand_fields(nx, ny, 5)
    rhs = rand_fields(nx, ny, 6)[2]
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_smoother_strip(strip)
    for vcycles, tol in ((1, 1e-9), (3, 1e-9), (3, 1e3)):
        s.upload(p=p, rhs=rhs)
        op = p.copy()
        ores, oit = port.pressure_solve(nx, ny, Lx, Ly, op, rhs, cfdo.make_params("mg", vcycles=vcycles, tol=tol))
        gres, git = s.pressure_solve(cfd.make_params("mg", vcycles=vcycles, tol=tol))
        gp = s.download()[2]
        assert same(gp, op), (vcycles, nbad(gp, op), where_bad(gp, op))
        assert git == oit
        assert abs(gres - ores) <= 1e-12 * ores
        assert s.last_smoother == ("k_rbgs_stream" if strip else "k_rbgs_fused")


@pytest.mark.parametrize("strip", [0, 33])
def test_pressure_solve_mg_nonfinite_operands(cfd, port, strip):
    """Non-finite p / rhs entries: the reference reads them as 0 (pressure.cpp:84-88).  The streaming smoother detects
    them and hands the smooth() to the guarded tile kernel; the result must still be the reference's, bit for bit."""
    nx, ny, Lx, Ly = 200, 150, 1.0, 1.0
    _, _, p = rand_fields(nx, ny, 15)
    rhs = rand_fields(nx, ny, 16)[2]
    p[7, 9] = np.nan
    p[100, 150] = np.inf
    p[0, 30] = -np.inf      # a ghost cell
    rhs[60, 61] = np.inf
    rhs[149, 199] = np.nan
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_smoother_strip(strip)
    for vcycles in (1, 2):
        s.upload(p=p, rhs=rhs)
        op = p.copy()
        ores, oit = port.pressure_solve(nx, ny, Lx, Ly, op, rhs, cfdo.make_params("mg", vcycles=vcycles, tol=1e-9))
        gres, git = s.pressure_solve(cfd.make_params("mg", vcycles=vcycles, tol=1e-9))
        gp = s.download()[2]
        assert same_or_nan(gp, op), (vcycles, nbad(gp, op), where_bad(gp, op))
        assert git == oit and abs(gres - ores) <= 1e-12 * ores
    # and a clean solve afterwards takes the fast path again with the same answer
    _, _, p2 = rand_fields(nx, ny, 17)
    rhs2 = rand_fields(nx, ny, 18)[2]
    s.upload(p=p2, rhs=rhs2)
    op = p2.copy()
    port.pressure_solve(nx, ny, Lx, Ly, op, rhs2, cfdo.make_params("mg", vcycles=2, tol=1e-9))
    s.pressure_solve(cfd.make_params("mg", vcycles=2, tol=1e-9))
    assert same(s.download()[2], op)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("nx,ny,Lx,Ly", [(16, 16, 1.0, 1.0), (37, 21, 2.0, 0.9), (64, 32, 2.0, 1.0), (130, 67, 1.0, 1.0),
                                         (300, 211, 1.3, 0.7)])
def test_pressure_solve_pcg(cfd, port, nx, ny, Lx, Ly, variant):
    rhs = rand_fields(nx, ny, 8)[2]
    p_junk = rand_fields(nx, ny, 9)[2]  # PCG must cold-start: junk in p (ghosts included) is cleared
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_variant(variant)
    for tol, cap in ((1e-6, 500), (1e-8, 25), (1e3, 50), (1e-6, 0)):
        s.upload(p=p_junk, rhs=rhs)
        op = p_junk.copy()
        ores, oit = port.pressure_solve(nx, ny, Lx, Ly, op, rhs, cfdo.make_params("pcg", tol=tol, pcg_max_iters=cap))
        gres, git = s.pressure_solve(cfd.make_params("pcg", tol=tol, pcg_max_iters=cap))
        gp = s.download()[2]
        assert abs(git - oit) <= 1, (tol, cap, git, oit)
        if git == oit:
            assert rel_l2(gp, op) <= PCG_TOL, (tol, cap, rel_l2(gp, op))
            assert abs(gres - ores) <= 1e-8 * max(ores, 1e-300)
        assert same(gp[0, :], op[0, :]) and same(gp[:, 0], op[:, 0])  # ghosts are 0 after PCG


@pytest.mark.parametrize("nx,ny", [(64, 32), (300, 211), (512, 256)])
def test_pcg_nonfinite_rhs_takes_the_guarded_kernel(cfd, port, nx, ny):
    """Non-finite entries in the right-hand side: the reference's guards turn them into zeros (pressure.cpp:20-60).  The
    cluster-resident solve computes without guards, notices, gives up and the guarded cooperative kernel redoes the solve."""
    rhs = rand_fields(nx, ny, 8)[2]
    rhs[5, 7], rhs[9, 11], rhs[3, 3] = np.nan, np.inf, -np.inf
    s = cfd.Solver(nx, ny, 2.0, 1.0)
    for tol, cap in ((1e-8, 60), (1e-2, 500)):
        s.upload(p=np.zeros_like(rhs), rhs=rhs)
        op = np.zeros_like(rhs)
        ores, oit = port.pressure_solve(nx, ny, 2.0, 1.0, op, rhs, cfdo.make_params("pcg", tol=tol, pcg_max_iters=cap))
        gres, git = s.pressure_solve(cfd.make_params("pcg", tol=tol, pcg_max_iters=cap))
        gp = s.download()[2]
        assert "k_pcg_coop" in s.last_pcg or "k_pcg_cluster" in s.last_pcg
        assert abs(git - oit) <= 1, (tol, cap, git, oit)
        if git == oit:
            assert rel_l2(gp, op) <= PCG_TOL, (tol, cap, rel_l2(gp, op))
            assert abs(gres - ores) <= 1e-8 * max(ores, 1e-300)
    # and a clean right-hand side afterwards goes through the cluster kernel again
    rhs2 = rand_fields(nx, ny, 9)[2]
    s.upload(p=np.zeros_like(rhs2), rhs=rhs2)
    op = np.zeros_like(rhs2)
    ores, oit = port.pressure_solve(nx, ny, 2.0, 1.0, op, rhs2, cfdo.make_params("pcg", tol=1e-8, pcg_max_iters=40))
    gres, git = s.pressure_solve(cfd.make_params("pcg", tol=1e-8, pcg_max_iters=40))
    assert abs(git - oit) <= 1 and rel_l2(s.download()[2], op) <= PCG_TOL


def test_pcg_breakdown_and_zero_rhs(cfd, port):
    nx, ny = 16, 16
    s = cfd.Solver(nx, ny, 1.0, 1.0)
    rhs = np.zeros(cfdo.shapes(nx, ny)["p"])
    s.upload(p=rhs + 1.0, rhs=rhs)
    gres, git = s.pressure_solve(cfd.make_params("pcg", tol=1e-6, pcg_max_iters=10))
    assert gres == 0.0 and git == 0 and not s.download()[2].any()


# ------------------------------------------------------------------------------------------------ whole steps

def run_oracle(lib, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, vcycles, dt_override=0.0):
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(lib, nx, ny, Lx, Ly)
    if preset == "shear":
        cfdo.shear_init(sim.u, nx, ny, Ly)
    hist = []
    for _ in range(steps):
        dt = sim.step(bc, Re, CFL, pp, dt_override)
        hist.append((dt, sim.last_residual, sim.last_iters))
    out = {k: getattr(sim, k).copy() for k in ("u", "v", "p", "rhs")}
    sim.close()
    return hist, out


def run_gpu(cfd, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, vcycles, dt_override=0.0, strip=0):
    bc, Re, CFL = cfd.preset(preset, Ly)
    pp = cfd.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_smoother_strip(strip)
    if preset == "shear":
        s.upload(u=cfd.shear_initial_u(nx, ny, Ly))
    hist = []
    for _ in range(steps):
        info = s.step(bc, Re, CFL, pp, dt_override)
        hist.append((info.dt, info.residual, info.iters))
    u, v, p, rhs = s.download()
    s.close()
    return hist, dict(u=u, v=v, p=p, rhs=rhs)


def compare_run(solver, g, o, tag):
    (gh, gf), (oh, of) = g, o
    if solver == "mg":
        assert [h[0] for h in gh] == [h[0] for h in oh], (tag, "dt", gh, oh)
    else:  # dt = f(max|u|) inherits the last-bit differences of the PCG inner products
        assert all(abs(a[0] - b[0]) <= 1e-12 * b[0] for a, b in zip(gh, oh)), (tag, "dt", gh, oh)
    if solver == "mg":
        for k in ("u", "v", "p", "rhs"):
            assert same(gf[k], of[k]), (tag, k, nbad(gf[k], of[k]), where_bad(gf[k], of[k]))
        for a, b in zip(gh, oh):
            assert abs(a[1] - b[1]) <= 1e-12 * abs(b[1])
    else:
        for a, b in zip(gh, oh):
            assert abs(a[2] - b[2]) <= 1, (tag, "iters", gh, oh)
        for k in ("u", "v", "p"):
            assert rel_l2(gf[k], of[k]) <= PCG_TOL, (tag, k, rel_l2(gf[k], of[k]))


@pytest.mark.parametrize("preset", ["jet", "lid", "shear"])
@pytest.mark.parametrize("solver", ["mg", "pcg"])
@pytest.mark.parametrize("grid", [(64, 32, 2.0, 1.0), (48, 36, 1.5, 1.0), (130, 67, 1.0, 1.0)])
def test_whole_runs_vs_oracle(cfd, port, preset, solver, grid):
    nx, ny, Lx, Ly = grid
    args = (preset, solver, nx, ny, Lx, Ly, 6, 1e-8, 200, 2)
    compare_run(solver, run_gpu(cfd, *args), run_oracle(port, *args), (preset, solver, grid))


@pytest.mark.parametrize("preset", ["jet", "lid", "shear"])
@pytest.mark.parametrize("grid,strip", [((64, 32, 2.0, 1.0), 17), ((130, 67, 1.0, 1.0), 33), ((300, 140, 2.0, 1.0), 65)])
def test_whole_runs_mg_streaming_smoother_vs_oracle(cfd, port, preset, grid, strip):
    """The streaming smoother (automatic only on grids that fill the GPU) forced onto oracle-sized grids: bit-exact."""
    nx, ny, Lx, Ly = grid
    args = (preset, "mg", nx, ny, Lx, Ly, 6, 1e-8, 200, 2)
    compare_run("mg", run_gpu(cfd, *args, strip=strip), run_oracle(port, *args), (preset, grid, strip))


@pytest.mark.parametrize("key", [f"{p}_{s}_{g}" for p in ("jet", "lid", "shear") for s in ("pcg", "mg")
                                 for g in ("64x32", "40x24")])
def test_whole_runs_vs_golden_fixtures(cfd, golden_runs, golden_fields, key):
    """Against the committed vectors generated from the reference's own sources (tests/golden/make_golden.py)."""
    c = golden_runs[key]
    gh, gf = run_gpu(cfd, c["preset"], c["solver"], c["nx"], c["ny"], c["Lx"], c["Ly"], c["steps"], c["tol"],
                     c["pcg_max_iters"], c["vcycles"])
    if c["solver"] == "mg":
        assert [h[0] for h in gh] == [h[0] for h in c["history"]]
    else:
        assert all(abs(a[0] - b[0]) <= 1e-12 * b[0] for a, b in zip(gh, c["history"]))
    for k in ("u", "v", "p", "rhs"):
        want = golden_fields[f"{key}/{k}"]
        if c["solver"] == "mg":
            assert same(gf[k], want), (k, nbad(gf[k], want))
        elif k != "rhs":
            assert rel_l2(gf[k], want) <= PCG_TOL, (k, rel_l2(gf[k], want))


def test_dt_override_gui_style(cfd, port):
    args = ("lid", "pcg", 32, 32, 1.0, 1.0, 3, 1e-8, 50, 2, 2.5e-4)
    g, o = run_gpu(cfd, *args), run_oracle(port, *args)
    assert [h[0] for h in g[0]] == [2.5e-4] * 3
    compare_run("pcg", g, o, "dt_override")


def test_step_info_divergences(cfd, port):
    """div_before / div_after (the reference's stderr log) against the port's serial sums."""
    import ctypes as C
    nx, ny, Lx, Ly = 64, 32, 2.0, 1.0
    bc, Re, CFL = cfdo.preset("jet", Ly)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    sim.step(bc, Re, CFL, cfdo.make_params("mg"))
    b, a = C.c_double(), C.c_double()
    port.L.cfdo_last_divs.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    port.L.cfdo_last_divs(sim.h, C.byref(b), C.byref(a))
    gbc, _, _ = cfd.preset("jet", Ly)
    s = cfd.Solver(nx, ny, Lx, Ly)
    info = s.step(gbc, Re, CFL, cfd.make_params("mg"))
    assert abs(info.div_before - b.value) <= 1e-12 * b.value and abs(info.div_after - a.value) <= 1e-12 * a.value
    assert info.nonfinite == 0 and info.pcg_breakdown == -1


def test_nonfinite_input_is_sanitised_like_the_reference(cfd, port):
    nx, ny, Lx, Ly = 24, 20, 1.0, 1.0
    u, v, p = rand_fields(nx, ny, 3)
    u[4, 5] = np.nan
    v[6, 3] = np.inf
    bc, Re, CFL = cfdo.preset("lid", Ly)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    sim.u[:], sim.v[:], sim.p[:] = u, v, p
    dt = sim.step(bc, Re, CFL, cfdo.make_params("mg"))
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.upload(u, v, p)
    info = s.step(cfd.preset("lid", Ly)[0], Re, CFL, cfd.make_params("mg"))
    gu, gv, gp, grhs = s.download()
    assert info.dt == dt and (info.nonfinite & 1)
    assert same(gu, sim.u) and same(gv, sim.v) and same(gp, sim.p) and same(grhs[1:-1, 1:-1], sim.rhs[1:-1, 1:-1])


@pytest.mark.parametrize("march", [2, 1])
@pytest.mark.parametrize("preset,nx,ny,dt_override", [("lid", 300, 211, 0.0), ("jet", 257, 130, 2.5e-4), ("shear", 130, 67, 1e-3),
                                                       ("lid", 560, 300, 1e-3)])
def test_slow_path_of_the_marching_kernel_inside_a_step(cfd, port, preset, nx, ny, dt_override, march):
    """Overflowing differences, infinities and NaNs INSIDE the marching kernels' region, through whole steps: the values
    k_rhs_march2 repairs behind its loop (RK stage 1, stage 2 and the divergence closed from registers) and the ones the
    first version recomputes inside the loop must both be the reference's, bit for bit (rhs included)."""
    Lx, Ly = 2.0, 1.0
    u, v, p = rand_fields(nx, ny, 7 * nx + ny)
    u[40:44, 50:70] = 0.0
    v[40:44, 50:70] = -0.0
    v[52, 60:64] = 1e300          # products and differences overflow
    u[45, 70:74] = -1e300
    u[60, 81] = 8e307             # 2 * q overflows in the Laplacian
    u[48, 90] = np.inf
    v[55, 100] = np.nan
    u[ny // 2, nx - 40] = -np.inf
    v[ny - 12, 45:47] = 1.7e308
    bc, Re, CFL = cfdo.preset(preset, Ly)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    sim.u[:], sim.v[:], sim.p[:] = u, v, p
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.set_march_kernel(march)
    s.upload(u, v, p)
    gbc = cfd.preset(preset, Ly)[0]
    for step in range(2):  # the second step starts from the sanitised state and large, finite values
        dt = sim.step(bc, Re, CFL, cfdo.make_params("mg"), dt_override)
        info = s.step(gbc, Re, CFL, cfd.make_params("mg"), dt_override)
        gu, gv, gp, grhs = s.download()
        assert info.dt == dt, (step, info.dt, dt)
        assert same(gu, sim.u), ("u", step, nbad(gu, sim.u), where_bad(gu, sim.u))
        assert same(gv, sim.v), ("v", step, nbad(gv, sim.v), where_bad(gv, sim.v))
        assert same(gp, sim.p), ("p", step, nbad(gp, sim.p), where_bad(gp, sim.p))
        assert same(grhs[1:-1, 1:-1], sim.rhs[1:-1, 1:-1]), ("rhs", step, nbad(grhs[1:-1, 1:-1], sim.rhs[1:-1, 1:-1]))
        if step == 0:
            assert info.nonfinite & 1


def test_marching_kernels_agree_bitwise_2048(cfd):
    """k_rhs_march2 against the first version on a grid the oracle would be slow on: State and work arrays after 3 steps."""
    nx, ny = 2048, 1536
    u, v, p = rand_fields(nx, ny, 43)
    u[700:704, 900:940] = 1e300
    v[300, 1000:1004] = -np.inf
    outs = []
    for march in (1, 2):
        s = cfd.Solver(nx, ny, 2.0, 1.5)
        s.set_march_kernel(march)
        s.upload(u, v, p)
        bc, Re, CFL = cfd.preset("jet", 1.5)
        for _ in range(3):
            s.step(bc, Re, CFL, cfd.make_params("mg"))
        outs.append(s.download() + s.download_work())
    for a, b in zip(outs[0], outs[1]):
        assert same(a, b), (nbad(a, b), where_bad(a, b))


def test_kernel_variants_agree_bitwise_1024(cfd):
    """The optimised kernels (variant 1) against the simple ones (variant 0) on a grid the oracle would be slow on."""
    nx = ny = 1024
    u, v, p = rand_fields(nx, ny, 42)
    outs = []
    for variant in (0, 1, 2):
        s = cfd.Solver(nx, ny, 1.0, 1.0)
        s.set_variant(variant)
        s.upload(u, v, p)
        bc, Re, CFL = cfd.preset("jet", 1.0)
        for _ in range(2):
            s.step(bc, Re, CFL, cfd.make_params("mg"))
        outs.append(s.download() + s.download_work())
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert same(a, b), (nbad(a, b), where_bad(a, b))


def test_async_steps_match_sync_steps(cfd):
    nx, ny, Lx, Ly = 64, 64, 1.0, 1.0
    bc, Re, CFL = cfd.preset("jet", Ly)
    pp = cfd.make_params("pcg", tol=1e-8, pcg_max_iters=30)
    a, b = cfd.Solver(nx, ny, Lx, Ly), cfd.Solver(nx, ny, Lx, Ly)
    for _ in range(4):
        a.step(bc, Re, CFL, pp)
        b.step_async(bc, Re, CFL, pp)
    ib = b.sync()
    assert ib.iters == 30
    for x, y in zip(a.download(), b.download()):
        assert same(x, y)  # run-to-run deterministic, reductions included
    assert a.launch_count == b.launch_count > 0


# ------------------------------------------------------------------------------------------------ reference sizes

def test_default_grid_lid_pcg_vs_oracle(cfd, port):
    """BASELINE configs[0]: lid-driven cavity Re=1000 CFL=0.1, PCG tol=1e-8 on the default 512x256 / 2x1 grid."""
    args = ("lid", "pcg", 512, 256, 2.0, 1.0, 3, 1e-8, 200, 2)
    compare_run("pcg", run_gpu(cfd, *args), run_oracle(port, *args), "cfg0")


def test_jet_mg_1024_vs_oracle(cfd):
    """BASELINE configs[1]: jet plume, mg vcycles=2 at 1024^2 -- bit-exact against the multithreaded oracle
    (the mg path has no reduction feeding the fields, so the thread count does not matter)."""
    lib = cfdo.load("ref") if cfdo.have_ref() else cfdo.load("port")
    args = ("jet", "mg", 1024, 1024, 1.0, 1.0, 4, 1e-6, 200, 2)
    old = os.environ.get("OMP_NUM_THREADS")
    o = run_oracle(lib, *args)
    compare_run("mg", run_gpu(cfd, *args), o, "cfg1")
    compare_run("mg", run_gpu(cfd, *args, strip=257), o, "cfg1, streaming smoother, 257-row strips")


class omp_threads:
    """Thread count of the CPU checkers for the full-size cases (conftest pins OMP_NUM_THREADS=1 before libgomp loads;
    the mg path is bit-identical across thread counts, PCG differs by ~1e-15 -- SURVEY.md section 8c)."""

    def __init__(self, n=None):
        import ctypes
        self.n = n or (os.cpu_count() or 1)
        self.gomp = ctypes.CDLL("libgomp.so.1")

    def __enter__(self):
        self.gomp.omp_set_num_threads(self.n)

    def __exit__(self, *a):
        self.gomp.omp_set_num_threads(1)


def oracle_lib():
    return cfdo.load("ref") if cfdo.have_ref() else cfdo.load("port")


@pytest.mark.parametrize("preset", ["lid", "jet"])
def test_mg_8192_two_steps_vs_oracle_bitwise(cfd, preset):
    """BASELINE configs[3] / [4] at their named size: 8192^2 fp64, --solver=mg --vcycles=2, two full steps from the
    preset's initial state, u, v, p and rhs BIT-IDENTICAL to the reference's own sources (oracle/_ref, all host cores;
    the C port where the prebuilt .so did not travel)."""
    args = (preset, "mg", 8192, 8192, 1.0, 1.0, 2, 1e-6, 200, 2)
    with omp_threads():
        o = run_oracle(oracle_lib(), *args)
    g = run_gpu(cfd, *args)
    compare_run("mg", g, o, f"{preset} mg 8192^2")


def test_shear_pcg_4096_vs_oracle(cfd, port):
    """BASELINE configs[2] at its named size: periodic shear, PCG (200 iterations, tol 1e-6), 4096^2 fp64, two steps:
    rel-L2 <= 1e-10 on u, v, p, iteration counts within +-1 (the port counts iterations; the unmodified reference
    cannot)."""
    args = ("shear", "pcg", 4096, 4096, 1.0, 1.0, 2, 1e-6, 200, 2)
    with omp_threads():
        o = run_oracle(port, *args)
    compare_run("pcg", run_gpu(cfd, *args), o, "shear pcg 4096^2")


def test_default_grid_pcg_20_steps_vs_recorded_lines(cfd, port):
    """The reference's own CPU-runnable case (main.cpp:105: 512x256 on 2x1, PCG): 20+ steps against the stdout lines
    SURVEY.md section 8c recorded from the reference binary (printed with 6-7 significant digits), and field by field
    against the oracle."""
    recorded = {  # (preset, tol): {step: (format, dt, divergence_l2, residual)} exactly as printed (main.cpp:152-155)
        ("lid", 1e-8): {0: ("%.6e", "0.000390625", "2.000060e-02", "5.191889e-01"), 19: ("%.6e", None, "1.231257e-01", "4.603675e+01")},
        ("jet", 1e-6): {0: ("%g", "3.8147e-06", "4.56377", "1.53529e+06"), 20: ("%g", None, "2.60937", "1.68593e+06")},
        ("shear", 1e-6): {0: ("%.6e", "0.000390654", "4.730380e-03", "2.371961e+02"), 19: ("%.6e", None, "5.835455e-02", "3.044367e+02")},
    }
    for (preset, tol), lines in recorded.items():
        nx, ny, Lx, Ly, steps = 512, 256, 2.0, 1.0, max(lines) + 1
        bc, Re, CFL = cfd.preset(preset, Ly)
        pp = cfd.make_params("pcg", tol=tol, pcg_max_iters=200)
        s = cfd.Solver(nx, ny, Lx, Ly)
        if preset == "shear":
            s.upload(u=cfd.shear_initial_u(nx, ny, Ly))
        for k in range(steps):
            info = s.step(bc, Re, CFL, pp)
            if k in lines:
                fmt, dt, div, res = lines[k]
                got_div = s.divergence_l2()  # what the driver prints after the step (main.cpp:151)
                assert dt is None or "%g" % info.dt == dt, (preset, k, info.dt)
                assert fmt % got_div == div, (preset, k, got_div, div)
                assert fmt % info.residual == res, (preset, k, info.residual, res)
        if preset == "lid":
            assert "%.7g" % s.max_velocity() == "1.000203"  # "umax 1.000203" of step 19
        g = dict(zip(("u", "v", "p", "rhs"), s.download()))
        s.close()
        with omp_threads():
            oh, of = run_oracle(port, preset, "pcg", nx, ny, Lx, Ly, steps, tol, 200, 2)
        for k in ("u", "v", "p"):
            assert rel_l2(g[k], of[k]) <= PCG_TOL, (preset, k, rel_l2(g[k], of[k]))


def test_full_size_properties_8192(cfd):
    """Size-independent properties at BASELINE's full size (8192^2), next to the field comparison above: (1) the BC pass
    is idempotent for non-periodic presets, (2) a projection step leaves s.rhs equal to the divergence of the final
    (u, v) and divergence_l2 agrees with info.div_after, (3) the top lid row carries the imposed value, (4) the
    streaming and the tile smoother agree bitwise."""
    nx = ny = 8192
    bc, Re, CFL = cfd.preset("lid", 1.0)
    pp = cfd.make_params("mg", vcycles=2)
    s = cfd.Solver(nx, ny, 1.0, 1.0)
    infos = [s.step(bc, Re, CFL, pp) for _ in range(2)]
    div_l2 = s.divergence_l2()
    assert abs(div_l2 - infos[-1].div_after) <= 1e-10 * div_l2
    u = np.empty((ny + 2, nx + 3))
    s.download(u=u)
    assert (u[ny, 1:nx + 2] == 1.0).all() and (u[ny + 1, 1:nx + 2] == 1.0).all() and (u[1:ny, 1] == 0.0).all()  # corner faces take the top value (bc.cpp:111-168 runs last)
    s.apply_bc(bc, 3)
    u2 = np.empty_like(u)
    s.download(u=u2)
    assert same(u, u2)
    assert s.last_smoother == "k_rbgs_stream"
    p1 = np.empty((ny + 2, nx + 2))
    s.download(p=p1)
    s.close()
    t = cfd.Solver(nx, ny, 1.0, 1.0)
    t.set_variant(2)
    for _ in range(2):
        t.step(bc, Re, CFL, pp)
    assert t.last_smoother == "k_rbgs_fused"
    u3, p3 = np.empty_like(u), np.empty_like(p1)
    t.download(u=u3, p=p3)
    assert same(u, u3) and same(p1, p3)


def test_largest_single_gpu_grid_16384(cfd):
    """16384 x 16384 fp64 (2.1 GB per field, element offsets beyond 2^31 bytes): one mg step and 3 PCG iterations run,
    stay finite, are reproducible, and the lid row carries the imposed value."""
    nx = ny = 16384
    bc, Re, CFL = cfd.preset("lid", 1.0)
    s = cfd.Solver(nx, ny, 1.0, 1.0)
    a = s.step(bc, Re, CFL, cfd.make_params("mg", vcycles=2))
    b = s.step(bc, Re, CFL, cfd.make_params("pcg", pcg_max_iters=3))
    assert a.nonfinite == 0 and b.nonfinite == 0 and b.iters == 3 and np.isfinite(b.residual)
    assert a.dt == 1e-6  # pinned to the floor of compute_cfl at this resolution
    u = np.empty((ny + 2, nx + 3))
    s.download(u=u)
    assert np.isfinite(u).all() and (u[ny, 1:nx + 2] == 1.0).all() and abs(u[1:ny, 2:nx + 1]).max() > 0
    d1 = s.divergence_l2()
    s.close()
    t = cfd.Solver(nx, ny, 1.0, 1.0)
    t.step(bc, Re, CFL, cfd.make_params("mg", vcycles=2))
    t.step(bc, Re, CFL, cfd.make_params("pcg", pcg_max_iters=3))
    assert t.divergence_l2() == d1
