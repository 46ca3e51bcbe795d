"""The CPU checkers themselves: oracle/cfd_oracle.c (port) is pinned BITWISE against oracle/_ref (the
reference's own sources) and against the committed golden vectors generated from oracle/_ref."""
import hashlib
import itertools
import os
import subprocess

import numpy as np
import pytest

from oracle import cfdo

RICH = dict(left_inflow_u=0.8, jet_center=0.45, jet_width=0.3, left_moving=0.11, right_moving=-0.12,
            bottom_moving=0.13, top_moving=1.0, right_inflow_u=0.21, right_inflow_v=0.22, bottom_inflow_u=0.23,
            bottom_inflow_v=0.24, top_inflow_u=0.25, top_inflow_v=0.26)


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


def run(lib, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, vcycles):
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(lib, nx, ny, Lx, Ly)
    if preset == "shear":
        cfdo.shear_init(sim.u, nx, ny, Ly)
    hist = []
    for _ in range(steps):
        dt = sim.step(bc, Re, CFL, pp)
        hist.append([dt, lib.divergence_l2(nx, ny, Lx, Ly, sim.u, sim.v), sim.last_residual])
    out = {k: getattr(sim, k).copy() for k in ("u", "v", "p", "rhs")}
    iters_done = sim.last_iters
    sim.close()
    return hist, out, iters_done


def test_survey_known_answers(golden_runs):
    """The committed fixtures carry the known answers recorded in SURVEY.md §8c (sha256 prefixes)."""
    want = {"jet_pcg": "9fad705549df0f7b", "jet_mg": "d84f4997d99c64c1", "lid_pcg": "dd1d17c008cabbf4",
            "lid_mg": "3688db8f5e5e372b", "shear_pcg": "41dcb17bb366aa71", "shear_mg": "253bdba309616df4"}
    for k, sha in want.items():
        assert golden_runs[k + "_64x32"]["sha256_uvp"].startswith(sha)
    assert abs(golden_runs["jet_pcg_64x32"]["norm_p"] - 365.1580683168618) < 1e-12


@pytest.mark.parametrize("key", [f"{p}_{s}_{g}" for p in ("jet", "lid", "shear") for s in ("pcg", "mg")
                                 for g in ("64x32", "40x24")])
def test_port_matches_golden_runs(port, golden_runs, golden_fields, key):
    c = golden_runs[key]
    hist, out, _ = run(port, c["preset"], c["solver"], c["nx"], c["ny"], c["Lx"], c["Ly"], c["steps"], c["tol"],
                       c["pcg_max_iters"], c["vcycles"])
    assert hist == c["history"]  # dt, div, res: exact doubles survive the JSON round trip
    for k in ("u", "v", "p", "rhs"):
        assert same(out[k], golden_fields[f"{key}/{k}"]), k
    sha = hashlib.sha256(out["u"].tobytes() + out["v"].tobytes() + out["p"].tobytes()).hexdigest()
    assert sha == c["sha256_uvp"]


def test_port_matches_golden_stages(port, golden_stages):
    g = golden_stages
    nx, ny, Lx, Ly = int(g["grid"][0]), int(g["grid"][1]), float(g["grid"][2]), float(g["grid"][3])
    u, v, p = g["u"], g["v"], g["p"]
    du, dv = np.zeros_like(u), np.zeros_like(v)
    port.advect_u(nx, ny, Lx, Ly, u, v, du); port.advect_v(nx, ny, Lx, Ly, u, v, dv)
    assert same(du, g["advect_u"]) and same(dv, g["advect_v"])
    port.diffuse_u(nx, ny, Lx, Ly, u, du, 1.0 / 50.0); port.diffuse_v(nx, ny, Lx, Ly, v, dv, 1.0 / 50.0)
    assert same(du, g["rhs_u"]) and same(dv, g["rhs_v"])
    assert port.compute_cfl(nx, ny, Lx, Ly, u, v, 50.0, 0.5) == g["cfl"][0]
    assert port.compute_cfl(nx, ny, Lx, Ly, u * 1e-3, v * 1e-3, 1000.0, 0.1) == g["cfl"][1]
    rhs = np.zeros_like(p); port.divergence(nx, ny, Lx, Ly, u, v, rhs)
    assert same(rhs, g["divergence"])
    u2, v2 = u.copy(), v.copy(); port.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p, 1e-3)
    assert same(u2, g["gradp_u"]) and same(v2, g["gradp_v"])
    assert port.divergence_l2(nx, ny, Lx, Ly, u, v) == g["div_l2"][0]
    assert port.max_velocity(nx, ny, Lx, Ly, u, v) == g["max_vel"][0]
    for solver in ("pcg", "mg"):
        ps = g[f"solve_{solver}_p0"].copy()
        res, _ = port.pressure_solve(nx, ny, Lx, Ly, ps, rhs, cfdo.make_params(solver, vcycles=3, tol=1e-9,
                                                                               pcg_max_iters=25))
        assert same(ps, g[f"solve_{solver}_p"]) and res == g[f"solve_{solver}_res"][0]
    for n, c in enumerate(g["bc_combos"]):
        bc = cfdo.make_bc(*[int(x) for x in c], **RICH)
        for name, a in (("u", u), ("v", v), ("p", p)):
            b = a.copy(); port.apply_bc(name, nx, ny, Lx, Ly, b, bc)
            assert same(b, g[f"bc{n}_{name}"]), (n, name)


def test_port_vs_ref_all_bc_combinations(port, ref):
    """All 5^4 side-type combinations, sentinel (random) arrays, bit-exact, u/v/p."""
    nx, ny, Lx, Ly = 9, 7, 1.1, 0.6
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(7)
    base = {k: rng.uniform(-1, 1, s) for k, s in sh.items()}
    for combo in itertools.product(range(5), repeat=4):
        bc = cfdo.make_bc(*combo, **RICH)
        for name in ("u", "v", "p"):
            a, b = base[name].copy(), base[name].copy()
            port.apply_bc(name, nx, ny, Lx, Ly, a, bc)
            ref.apply_bc(name, nx, ny, Lx, Ly, b, bc)
            assert same(a, b), (combo, name)
            # applied twice (periodic sides swap the boundary columns on every call)
            port.apply_bc(name, nx, ny, Lx, Ly, a, bc)
            ref.apply_bc(name, nx, ny, Lx, Ly, b, bc)
            assert same(a, b), (combo, name, "second call")


@pytest.mark.parametrize("nx,ny,Lx,Ly", [(16, 16, 1.0, 1.0), (37, 21, 2.0, 0.9), (8, 50, 0.3, 1.9)])
def test_port_vs_ref_stages_random(port, ref, nx, ny, Lx, Ly):
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(nx * 1000 + ny)
    u, v, p = (rng.uniform(-0.5, 0.5, sh[k]) for k in ("u", "v", "p"))
    u[3, 4] = 0.0; v[2, 2] = -0.0  # exercise the strict `> 0` upwind test and minmod's `<= 0`
    outs = []
    for lib in (port, ref):
        du, dv = np.full(sh["u"], 9.0), np.full(sh["v"], 9.0)  # advect_* must zero-fill first
        lib.advect_u(nx, ny, Lx, Ly, u, v, du); lib.advect_v(nx, ny, Lx, Ly, u, v, dv)
        lib.diffuse_u(nx, ny, Lx, Ly, u, du, 1 / 50.0); lib.diffuse_v(nx, ny, Lx, Ly, v, dv, 1 / 50.0)
        rhs = np.full(sh["p"], 9.0); lib.divergence(nx, ny, Lx, Ly, u, v, rhs)
        u2, v2 = u.copy(), v.copy(); lib.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p, 3e-4)
        sc = [lib.compute_cfl(nx, ny, Lx, Ly, u, v, 50.0, 0.5), lib.divergence_l2(nx, ny, Lx, Ly, u, v),
              lib.max_velocity(nx, ny, Lx, Ly, u, v)]
        sol = []
        for solver in ("pcg", "mg"):
            ps = p.copy()
            res, _ = lib.pressure_solve(nx, ny, Lx, Ly, ps, rhs, cfdo.make_params(solver, vcycles=3, tol=1e-7,
                                                                                  pcg_max_iters=40))
            sol += [ps, np.array([res])]
        outs.append([du, dv, rhs, u2, v2, np.array(sc)] + sol)
    for a, b in zip(*outs):
        assert same(a, b)


def test_port_vs_ref_nonfinite_inputs(port, ref):
    """NaN/Inf handling: guards in divergence/gradient/Laplacian and sanitize (SURVEY §5 failure detection)."""
    nx, ny, Lx, Ly = 12, 10, 1.0, 1.0
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(3)
    u, v, p = (rng.uniform(-0.5, 0.5, sh[k]) for k in ("u", "v", "p"))
    u[4, 5] = np.nan; v[6, 3] = np.inf; p[5, 5] = -np.inf
    res = []
    for lib in (port, ref):
        rhs = np.zeros(sh["p"]); lib.divergence(nx, ny, Lx, Ly, u, v, rhs)
        u2, v2 = u.copy(), v.copy(); lib.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p, 1e-3)
        pm = p.copy(); r1, _ = lib.pressure_solve(nx, ny, Lx, Ly, pm, rhs, cfdo.make_params("mg", vcycles=2))
        w = u.copy(); lib.sanitize(w)
        res.append([rhs, u2, v2, pm, np.array([r1]), w])
    for a, b in zip(*res):
        assert same(a, b)


@pytest.mark.parametrize("preset,solver", [(p, s) for p in ("jet", "lid", "shear") for s in ("pcg", "mg")])
def test_port_vs_ref_whole_runs(port, ref, preset, solver):
    """Whole multi-step runs on a grid that is NOT in the golden set; also with a GUI-style dt_override."""
    nx, ny, Lx, Ly = 48, 36, 1.5, 1.0
    a = run(port, preset, solver, nx, ny, Lx, Ly, 6, 1e-7, 80, 2)
    b = run(ref, preset, solver, nx, ny, Lx, Ly, 6, 1e-7, 80, 2)
    assert a[0] == b[0]
    for k in ("u", "v", "p", "rhs"):
        assert same(a[1][k], b[1][k]), k


def test_dt_override(port, ref):
    nx, ny, Lx, Ly = 32, 32, 1.0, 1.0
    bc, Re, CFL = cfdo.preset("lid", Ly)
    pp = cfdo.make_params("pcg", tol=1e-8, pcg_max_iters=50)
    out = []
    for lib in (port, ref):
        sim = cfdo.Sim(lib, nx, ny, Lx, Ly)
        dts = [sim.step(bc, Re, CFL, pp, dt_override=2.5e-4) for _ in range(3)]
        out.append((dts, sim.u.copy(), sim.p.copy()))
    assert out[0][0] == out[1][0] == [2.5e-4] * 3
    assert same(out[0][1], out[1][1]) and same(out[0][2], out[1][2])


def test_pcg_iteration_counter_by_cap_bisection(port, ref):
    """The reference does not expose its PCG iteration count; the port does.  Pin the port's counter on the
    reference: with cap = n the reference must give the converged answer, with cap = n-1 a different one."""
    nx, ny, Lx, Ly = 20, 20, 1.0, 1.0
    rng = np.random.default_rng(11)
    rhs = rng.uniform(-0.5, 0.5, cfdo.shapes(nx, ny)["p"])
    p0 = np.zeros_like(rhs)
    _, n = port.pressure_solve(nx, ny, Lx, Ly, p0.copy(), rhs, cfdo.make_params("pcg", tol=1e-6, pcg_max_iters=500))
    assert 5 < n < 500
    sols = {}
    for cap in (n - 1, n, 500):
        pr = p0.copy()
        ref.pressure_solve(nx, ny, Lx, Ly, pr, rhs, cfdo.make_params("pcg", tol=1e-6, pcg_max_iters=cap))
        sols[cap] = pr
    assert same(sols[n], sols[500]) and not same(sols[n - 1], sols[n])


def test_no_fma_in_cpu_checkers(port):
    """SURVEY §7 hard part 1: the pin is an x86-64 build WITHOUT fused multiply-add."""
    libs = [port.path]
    if cfdo.have_ref():
        libs.append(cfdo.load("ref").path)
    for path in libs:
        dis = subprocess.run(["objdump", "-d", path], capture_output=True, text=True).stdout
        assert "vfmadd" not in dis and "vfnmadd" not in dis and "vfmsub" not in dis, path


def test_mg_is_fine_grid_rbgs_only(port):
    """mg_levels and rbgs_sweeps are dead parameters in the reference-faithful mode."""
    nx, ny, Lx, Ly = 16, 12, 1.0, 1.0
    rng = np.random.default_rng(5)
    sh = cfdo.shapes(nx, ny)["p"]
    rhs, p0 = rng.uniform(-1, 1, sh), rng.uniform(-1, 1, sh)
    a, b = p0.copy(), p0.copy()
    port.pressure_solve(nx, ny, Lx, Ly, a, rhs, cfdo.make_params("mg", vcycles=2, mg_levels=3, rbgs_sweeps=2))
    port.pressure_solve(nx, ny, Lx, Ly, b, rhs, cfdo.make_params("mg", vcycles=2, mg_levels=7, rbgs_sweeps=9))
    assert same(a, b)


def test_tvd_debug_log_of_the_reference_is_unreachable_for_finite_fields(ref, tmp_path):
    """advection.cpp:19-27 prints "TVD max overshoot before clipping" when a limited face state leaves the interval of its
    three samples by more than 1e-10.  For finite samples it cannot: c + minmod(c - m1, p1 - c) / 2 lies between c and p1,
    overflowing differences included (minmod then returns the finite one, or 0 when the signs differ).  TimeIntegrator::step
    only ever advects sanitised, i.e. finite, fields, so inside a step the line is dead -- which is why the B200 path
    evaluates the clip only on its slow path and has no overshoot log.  Evidence: the reference's own code with
    tvd_debug = true on fields mixing ordinary, tiny, huge and overflowing values.  (Non-finite caller arrays handed to the
    free functions advect_u / advect_v do print "inf"; checked last, to show the capture works.)"""
    import ctypes
    flag = ctypes.c_bool.in_dll(ref.L, "tvd_debug")
    nx, ny, Lx, Ly = 24, 16, 1.5, 1.0
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(5)
    pool = np.array([0.0, -0.0, 1.0, -1.0, 1e-300, 3.0, 1e300, -1e300, 1.0e308, 1.7e308, -1.7e308])

    def run(u, v):
        log = tmp_path / "stderr.txt"
        fd = os.open(str(log), os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        saved = os.dup(2)
        os.dup2(fd, 2)
        flag.value = True
        try:
            du, dv = np.zeros(sh["u"]), np.zeros(sh["v"])
            ref.advect_u(nx, ny, Lx, Ly, u, v, du)
            ref.advect_v(nx, ny, Lx, Ly, u, v, dv)
        finally:
            flag.value = False
            os.dup2(saved, 2)
            os.close(saved)
            os.close(fd)
        return log.read_text()

    for trial in range(60):
        u, v = rng.choice(pool, size=sh["u"]), rng.choice(pool, size=sh["v"])
        if trial % 2:
            u, v = u * rng.uniform(0.5, 1.0, sh["u"]), v * rng.uniform(0.5, 1.0, sh["v"])
        assert run(u, v) == "", trial
    wild = np.r_[pool, [np.inf, -np.inf, np.inf, -np.inf]]
    assert "TVD max overshoot before clipping: inf" in run(rng.choice(wild, size=sh["u"]), rng.choice(wild, size=sh["v"]))
