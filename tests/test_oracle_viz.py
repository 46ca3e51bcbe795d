"""The readers / mutators around the step in the CPU checkers: the plain-C restatement
(oracle/cfd_oracle.c) is pinned BITWISE against the reference's own viz.cpp / colormaps.cpp / shapes.cpp compiled into
oracle/_ref, on random fields with non-finite entries, every scalar field, every colormap and a set of shapes."""
import numpy as np
import pytest

from oracle import cfdo


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


SHAPES = [cfdo.Shape(0, 3.0, 2.0, 10.0, 6.0, 0, 1),      # solid rectangle
          cfdo.Shape(1, 12.5, 4.25, 9.0, 9.0, 1, 1),     # porous circle
          cfdo.Shape(1, 20.0, 10.0, 7.5, 3.0, 0, 1),     # circle with size.y != size.x (centre uses size.y, shapes.cpp:42-43)
          cfdo.Shape(0, 1.0, 1.0, 50.0, 50.0, 1, 0),     # disabled
          cfdo.Shape(0, -4.0, 14.0, 9.0, 100.0, 1, 1),   # clipped by the grid, overlaps the circle: porous twice
          cfdo.Shape(1, 30.2, 1.7, 0.8, 0.8, 0, 1)]      # smaller than a cell


@pytest.mark.parametrize("nx,ny,Lx,Ly", [(37, 21, 2.0, 0.9), (16, 16, 1.0, 1.0), (64, 40, 1.3, 0.7)])
def test_compute_scalar_and_stats_port_vs_ref(port, ref, nx, ny, Lx, Ly):
    u, v, p = fields(nx, ny, 11)
    for name, f in cfdo.FIELDS.items():
        a, b = port.compute_scalar(nx, ny, Lx, Ly, u, v, p, f), ref.compute_scalar(nx, ny, Lx, Ly, u, v, p, f)
        assert same(a[0], b[0]), name
        assert a[1:] == b[1:], (name, a[1:], b[1:])  # vmin, vmax, mean, std
    z = np.zeros_like(u), np.zeros_like(v), np.zeros_like(p)
    a, b = port.compute_scalar(nx, ny, Lx, Ly, *z, 2), ref.compute_scalar(nx, ny, Lx, Ly, *z, 2)
    assert a[1:] == b[1:] == (0.0, 1e-6, 0.0, 0.0)  # the span rule of viz.cpp:63-64


def test_apply_shapes_port_vs_ref(port, ref):
    nx, ny, Lx, Ly = 40, 24, 2.0, 1.0
    u, v, _ = fields(nx, ny, 12, bad=False)
    ua, va, ub, vb = u.copy(), v.copy(), u.copy(), v.copy()
    port.apply_shapes(nx, ny, Lx, Ly, ua, va, SHAPES)
    ref.apply_shapes(nx, ny, Lx, Ly, ub, vb, SHAPES)
    assert same(ua, ub) and same(va, vb)
    assert not same(ua, u) and (ua == 0.0).sum() > 40  # something was masked


def test_colormaps_port_vs_ref(port, ref):
    ts = np.concatenate([np.linspace(0.0, 1.0, 4001, dtype=np.float32),
                         np.float32([0.125, 0.375, 0.5, 0.625, 0.875, np.nextafter(np.float32(0.5), np.float32(0))])])
    for cm in range(-1, 9):  # unknown ids fall back to viridis (colormaps.cpp:124)
        for t in ts:
            assert port.apply_colormap(cm, float(t)) == ref.apply_colormap(cm, float(t)), (cm, t)


@pytest.mark.parametrize("field", ["speed", "vorticity", "pressure"])
@pytest.mark.parametrize("cm,smin,smax,normalize,gamma", [(0, 0.0, 1.0, False, 1.0), (2, -0.3, 0.4, True, 2.2),
                                                            (4, 0.0, 1.0, True, 0.7), (5, 0.1, 0.1, False, 1.0)])
def test_render_rgb_port_vs_ref(port, ref, field, cm, smin, smax, normalize, gamma):
    nx, ny, Lx, Ly = 48, 30, 2.0, 1.0
    u, v, p = fields(nx, ny, 13)
    a = port.render_rgb(nx, ny, Lx, Ly, u, v, p, cfdo.FIELDS[field], cm, smin, smax, normalize, gamma)
    b = ref.render_rgb(nx, ny, Lx, Ly, u, v, p, cfdo.FIELDS[field], cm, smin, smax, normalize, gamma)
    assert np.array_equal(a, b)  # same libm pow on both sides
