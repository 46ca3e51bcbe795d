"""GPU versions of the readers / mutators around the step against the CPU checker: the
visualisation scalar, its statistics, the obstacle masks (bit-exact) and the coloured frame (bit-exact for gamma = 1;
with gamma != 1 the device pow and glibc's may differ in the last place, which can move a byte by one)."""
import importlib

import numpy as np
import pytest

from oracle import cfdo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfd():
    return importlib.import_module("macos-cfd_b200")


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


def fields(nx, ny, seed, bad=True):
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(seed)
    u, v, p = (rng.uniform(-0.5, 0.5, sh[k]) for k in ("u", "v", "p"))
    if bad:
        u[3, 4] = np.nan
        v[5, 2] = np.inf
        p[2, 2] = -np.inf
    return u, v, p


SHAPES = [(0, 3.0, 2.0, 10.0, 6.0, 0, 1), (1, 12.5, 4.25, 9.0, 9.0, 1, 1), (1, 20.0, 10.0, 7.5, 3.0, 0, 1),
          (0, 1.0, 1.0, 50.0, 50.0, 1, 0), (0, -4.0, 14.0, 9.0, 100.0, 1, 1), (1, 30.2, 1.7, 0.8, 0.8, 0, 1)]


@pytest.mark.parametrize("nx,ny,Lx,Ly", [(37, 21, 2.0, 0.9), (64, 40, 1.3, 0.7), (700, 333, 2.0, 1.0)])
def test_compute_scalar_bit_exact(cfd, port, nx, ny, Lx, Ly):
    u, v, p = fields(nx, ny, 21)
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.upload(u, v, p)
    for name, f in cfd.FIELDS.items():
        g = s.compute_scalar(f)
        o = port.compute_scalar(nx, ny, Lx, Ly, u, v, p, f)
        assert same(g[0], o[0]), name
        assert g[1] == o[1] and g[2] == o[2], (name, g[1:3], o[1:3])          # vmin, vmax: order-independent
        assert abs(g[3] - o[3]) <= 1e-12 * max(abs(o[3]), 1e-300) + 1e-15      # mean, std: the sums are reordered
        assert abs(g[4] - o[4]) <= 1e-10 * max(abs(o[4]), 1e-300) + 1e-15
    assert s.compute_scalar(2, want_field=False)[0] is None


def test_compute_scalar_after_steps_matches_oracle(cfd, port):
    """The scalar of a stepped State: the device State never crosses PCIe for it."""
    nx, ny, Lx, Ly = 96, 64, 2.0, 1.0
    gbc, Re, CFL = cfd.preset("lid", Ly)
    obc, _, _ = cfdo.preset("lid", Ly)
    s, sim = cfd.Solver(nx, ny, Lx, Ly), cfdo.Sim(port, nx, ny, Lx, Ly)
    for _ in range(3):
        s.step(gbc, Re, CFL, cfd.make_params("mg"))
        sim.step(obc, Re, CFL, cfdo.make_params("mg"))
    for f in (2, 4):
        assert same(s.compute_scalar(f)[0], port.compute_scalar(nx, ny, Lx, Ly, sim.u, sim.v, sim.p, f)[0])


def test_apply_shapes_bit_exact(cfd, port):
    nx, ny, Lx, Ly = 40, 24, 2.0, 1.0
    u, v, p = fields(nx, ny, 22, bad=False)
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.upload(u, v, p)
    s.apply_shapes([cfd.Shape(*a) for a in SHAPES])
    s.apply_shapes([])
    gu, gv, _, _ = s.download()
    ou, ov = u.copy(), v.copy()
    port.apply_shapes(nx, ny, Lx, Ly, ou, ov, [cfdo.Shape(*a) for a in SHAPES])
    assert same(gu, ou) and same(gv, ov) and not same(gu, u)
    # a step after the masks still matches the oracle (the cached CFL maxima are dropped)
    gbc, Re, CFL = cfd.preset("jet", Ly)
    obc, _, _ = cfdo.preset("jet", Ly)
    sim = cfdo.Sim(port, nx, ny, Lx, Ly)
    sim.u[:], sim.v[:], sim.p[:] = ou, ov, p
    info = s.step(gbc, Re, CFL, cfd.make_params("mg"))
    dt = sim.step(obc, Re, CFL, cfdo.make_params("mg"))
    gu, gv, gp, _ = s.download()
    assert info.dt == dt and same(gu, sim.u) and same(gv, sim.v) and same(gp, sim.p)


@pytest.mark.parametrize("field", ["speed", "vorticity", "pressure"])
@pytest.mark.parametrize("cm,smin,smax,normalize", [(0, 0.0, 1.0, False), (1, 0.0, 1.0, True), (2, -0.3, 0.4, True),
                                                     (3, 0.0, 1.0, False), (4, 0.0, 1.0, True), (5, 0.1, 0.1, False),
                                                     (7, 0.0, 1.0, False)])
def test_render_rgb_gamma_one_is_bit_exact(cfd, port, field, cm, smin, smax, normalize):
    nx, ny, Lx, Ly = 130, 67, 2.0, 1.0
    u, v, p = fields(nx, ny, 23)
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.upload(u, v, p)
    g = s.render_rgb(cfd.FIELDS[field], cm, smin, smax, normalize, 1.0)
    o = port.render_rgb(nx, ny, Lx, Ly, u, v, p, cfdo.FIELDS[field], cm, smin, smax, normalize, 1.0)
    assert np.array_equal(g, o)


@pytest.mark.parametrize("gamma", [2.2, 0.7])
def test_render_rgb_with_gamma(cfd, port, gamma):
    nx, ny, Lx, Ly = 130, 67, 2.0, 1.0
    u, v, p = fields(nx, ny, 24)
    s = cfd.Solver(nx, ny, Lx, Ly)
    s.upload(u, v, p)
    for cm in (0, 2, 4):
        g = s.render_rgb(2, cm, 0.0, 1.0, False, gamma).astype(int)
        o = port.render_rgb(nx, ny, Lx, Ly, u, v, p, 2, cm, 0.0, 1.0, False, gamma).astype(int)
        d = np.abs(g - o)
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3, (cm, d.max(), (d != 0).mean())  # pow: last-place differences only
