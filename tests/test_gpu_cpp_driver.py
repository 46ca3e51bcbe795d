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
