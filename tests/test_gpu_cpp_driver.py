"""The reference-shaped C++ layer (macos-cfd_b200/cpp) and the headless driver, end to end on the GPU: stdout lines,
raw dump and out/final.vtk against the oracle run of the same configuration (main.cpp:147-157 is the model)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cfdo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "macos-cfd_b200", "lib", "cfd2d_b200")


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "macos-cfd_b200", "cpp")], check=True)
    return EXE


def oracle_run(port, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, every):
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=2, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    if preset == "shear":
        cfdo.shear_init(sim.u, nx, ny, Ly)
    t, lines = 0.0, []
    for step in range(steps):
        dt = sim.step(bc, Re, CFL, pp)
        t += dt
        div = port.divergence_l2(nx, ny, Lx, Ly, sim.u, sim.v)
        if step % every == 0:
            lines.append((step, t, dt, div, sim.last_residual))
    return lines, {k: getattr(sim, k).copy() for k in ("u", "v", "p", "rhs")}


def vtk_text(nx, ny, Lx, Ly, f):
    """save_vtk of the reference (main.cpp:23-50); '%g' is the default ostream formatting."""
    dx, dy = Lx / nx, Ly / ny
    out = ["# vtk DataFile Version 3.0", "CFD2D snapshot", "ASCII", "DATASET STRUCTURED_POINTS",
           f"DIMENSIONS {nx} {ny} 1", "ORIGIN 0 0 0", f"SPACING {dx:g} {dy:g} 1", f"POINT_DATA {nx * ny}",
           "SCALARS pressure double", "LOOKUP_TABLE default"]
    p, u, v = f["p"], f["u"], f["v"]
    out += [f"{p[j + 1, i + 1]:g}" for j in range(ny) for i in range(nx)]
    out.append("VECTORS velocity double")
    for j in range(ny):
        for i in range(nx):
            uc = 0.5 * (u[j + 1, i + 1] + u[j + 1, i + 2])
            vc = 0.5 * (v[j + 1, i + 1] + v[j + 2, i + 1])
            out.append(f"{uc:g} {vc:g} 0")
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("preset", ["jet", "lid", "shear"])
@pytest.mark.parametrize("solver", ["mg", "pcg"])
def test_headless_driver_matches_oracle(exe, port, tmp_path, preset, solver):
    nx, ny, Lx, Ly, steps, tol, iters = 96, 48, 2.0, 1.0, 12, 1e-8, 80
    vtk, dump = tmp_path / "out" / "final.vtk", tmp_path / "dump.bin"
    r = subprocess.run([exe, "--no-gui", f"--preset={preset}", f"--solver={solver}", f"--nx={nx}", f"--ny={ny}",
                        f"--Lx={Lx}", f"--Ly={Ly}", f"--steps={steps}", f"--tol={tol}", f"--pcg-iters={iters}",
                        "--vcycles=2", "--mg-levels=3", "--print-every=5", f"--vtk={vtk}", f"--dump={dump}"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines, f = oracle_run(port, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, 5)
    got = r.stdout.strip().splitlines()
    assert len(got) == len(lines)
    for g, (step, t, dt, div, res) in zip(got, lines):
        want = "step %d time %g dt %g div %g res %g" % (step, t, dt, div, res)
        if solver == "mg":
            assert g == want
        else:  # PCG: same text up to the last printed digit of the reductions
            gt, wt = g.split(), want.split()
            assert gt[:6] == wt[:6]
            assert abs(float(gt[7]) - float(wt[7])) <= 2e-5 * abs(float(wt[7])) + 1e-300
            assert abs(float(gt[9]) - float(wt[9])) <= 2e-5 * abs(float(wt[9])) + 1e-300
    raw = np.fromfile(dump)
    sizes = [f[k].size for k in ("u", "v", "p", "rhs")]
    assert raw.size == sum(sizes)
    parts = np.split(raw, np.cumsum(sizes)[:-1])
    for k, a in zip(("u", "v", "p", "rhs"), parts):
        a = a.reshape(f[k].shape)
        if solver == "mg":
            assert np.array_equal(a.view(np.uint64), f[k].view(np.uint64)), k
        elif k != "rhs":
            assert np.linalg.norm(a - f[k]) <= 1e-10 * np.linalg.norm(f[k]), k
    if solver == "mg":
        assert vtk.read_text() == vtk_text(nx, ny, Lx, Ly, f)
    if preset == "jet":
        assert "Divergence: before=" in r.stderr  # the reference's stderr log (time_integrator.cpp:227)


def test_eager_coherence_is_a_true_drop_in(exe, port, tmp_path):
    """CFD_B200_COHERENCE has no effect on results: the driver forces lazy mode, so run it and compare with the same
    configuration through the Python binding in upload/step/download (eager-style) order."""
    import importlib
    cfd = importlib.import_module("macos-cfd_b200")
    nx, ny = 64, 48
    bc, Re, CFL = cfd.preset("lid", 1.0)
    pp = cfd.make_params("mg")
    s = cfd.Solver(nx, ny, 1.0, 1.0)
    sh = s.shapes()
    u, v, p, rhs = np.zeros(sh["u"]), np.zeros(sh["v"]), np.zeros(sh["p"]), np.zeros(sh["p"])
    for _ in range(5):
        s.upload(u, v, p)
        s.step(bc, Re, CFL, pp)
        s.download(u, v, p, rhs)
    dump = tmp_path / "d.bin"
    r = subprocess.run([exe, "--preset=lid", "--solver=mg", f"--nx={nx}", f"--ny={ny}", "--Lx=1", "--Ly=1", "--steps=5",
                        "--vtk=none", f"--dump={dump}"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(dump)
    assert np.array_equal(raw[:u.size].view(np.uint64), u.ravel().view(np.uint64))


def test_eager_drop_in_like_the_gui_loop(exe, port, tmp_path):
    """cpp/test_dropin.cpp: TimeIntegrator::step in eager mode with dt_override and a host-side edit of State between
    steps (the GUI's shapes), then every public free function on caller arrays -- all against the oracle, bit for bit
    (mg path / stencil stages) or within the PCG tolerance."""
    import ctypes as C
    prog = os.path.join(ROOT, "macos-cfd_b200", "lib", "test_dropin")
    r = subprocess.run([prog, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    nx, ny, Lx, Ly = 72, 40, 1.8, 1.0
    bc, Re, CFL = cfdo.preset("lid", Ly)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    want = []
    for _ in range(6):
        dt = sim.step(bc, 1000.0, 0.1, cfdo.make_params("mg", vcycles=2), dt_override=0.0005)
        want.append((dt, sim.last_residual, port.divergence_l2(nx, ny, Lx, Ly, sim.u, sim.v),
                     port.max_velocity(nx, ny, Lx, Ly, sim.u, sim.v)))
        sim.u[10:14, 20:26] *= 0.1
    lines = (tmp_path / "log.txt").read_text().splitlines()
    for line, w in zip(lines, want):
        got = [float(x) for x in line.split()]
        assert got[0] == w[0] and got[3] == w[3]
        assert abs(got[1] - w[1]) <= 1e-12 * w[1] and abs(got[2] - w[2]) <= 1e-12 * w[2]
    sh = cfdo.shapes(nx, ny)

    def load(name, key):
        return np.fromfile(tmp_path / name).reshape(sh[key])

    def same(a, b):
        return np.array_equal(a.view(np.uint64), b.view(np.uint64))

    assert same(load("u.bin", "u"), sim.u) and same(load("v.bin", "v"), sim.v) and same(load("p.bin", "p"), sim.p)
    assert same(load("rhs.bin", "p")[1:-1, 1:-1], sim.rhs[1:-1, 1:-1])
    # free functions
    u2, v2, p2 = sim.u.copy(), sim.v.copy(), sim.p.copy()
    jet = cfdo.make_bc(cfdo.INFLOW, cfdo.OUTFLOW, cfdo.WALL, cfdo.WALL, left_inflow_u=1.0, jet_center=0.5, jet_width=0.18)
    port.apply_bc("u", nx, ny, Lx, Ly, u2, jet)
    port.apply_bc("v", nx, ny, Lx, Ly, v2, jet)
    port.apply_bc("p", nx, ny, Lx, Ly, p2, jet)
    assert float(lines[6]) == port.compute_cfl(nx, ny, Lx, Ly, u2, v2, 50.0, 0.01)
    du = np.zeros(sh["u"])
    port.advect_u(nx, ny, Lx, Ly, u2, v2, du)
    port.diffuse_u(nx, ny, Lx, Ly, u2, du, 1.0 / 50.0)
    rhs2 = np.zeros(sh["p"])
    port.divergence(nx, ny, Lx, Ly, u2, v2, rhs2)
    port.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p2, 1e-3)
    res, iters = port.pressure_solve(nx, ny, Lx, Ly, p2, rhs2, cfdo.make_params("pcg", tol=1e-9, pcg_max_iters=30))
    assert same(load("f_du.bin", "u"), du) and same(load("f_rhs.bin", "p")[1:-1, 1:-1], rhs2[1:-1, 1:-1])
    assert same(load("f_u.bin", "u"), u2) and same(load("f_v.bin", "v"), v2)
    gp = load("f_p.bin", "p")
    assert np.linalg.norm(gp - p2) <= 1e-10 * np.linalg.norm(p2)
    gres, giters = lines[7].split()
    assert abs(float(gres) - res) <= 1e-8 * res and abs(int(giters) - iters) <= 1
    # viz.hpp / shapes.hpp through the C++ layer (reference signatures)
    su, sv = sim.u.copy(), sim.v.copy()
    port.apply_shapes(nx, ny, Lx, Ly, su, sv, [cfdo.Shape(0, 30.0, 8.0, 9.0, 5.0, 0, 1), cfdo.Shape(1, 50.5, 20.25, 11.0, 11.0, 1, 1)])
    assert same(load("s_u.bin", "u"), su) and same(load("s_v.bin", "v"), sv)
    speed = port.compute_scalar(nx, ny, Lx, Ly, su, sv, sim.p, cfdo.FIELDS["speed"])
    assert same(np.fromfile(tmp_path / "speed.bin").reshape(ny, nx), speed[0])
    vort = port.compute_scalar(nx, ny, Lx, Ly, su, sv, sim.p, cfdo.FIELDS["vorticity"])
    got = [float(x) for x in lines[8].split()]
    assert got[0] == speed[1] and got[1] == speed[2] and got[2] == vort[1] and got[3] == vort[2]
    assert abs(got[4] - vort[3]) <= 1e-12 * abs(vort[3]) + 1e-15 and abs(got[5] - vort[4]) <= 1e-10 * vort[4]
    rgb = port.render_rgb(nx, ny, Lx, Ly, su, sv, sim.p, cfdo.FIELDS["speed"], 4)
    assert np.array_equal(np.fromfile(tmp_path / "rgb.bin", dtype=np.uint8).reshape(ny, nx, 3), rgb)
    # lazy coherence: 6 steps with a stand-alone PressureSolver::solve on other arrays in the middle, twice (the second
    # State reuses the first one's freed addresses) -- both must equal 6 undisturbed steps from the zero state
    sim2 = cfdo.Sim(port, nx, ny, Lx, Ly)
    for _ in range(6):
        sim2.step(bc, 1000.0, 0.1, cfdo.make_params("mg", vcycles=2), dt_override=0.0005)
    for tag in ("lazy", "lazy2"):
        assert same(load(f"{tag}_u.bin", "u"), sim2.u) and same(load(f"{tag}_p.bin", "p"), sim2.p), tag
