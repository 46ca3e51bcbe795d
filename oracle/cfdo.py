"""ctypes loader for the two CPU checkers behind oracle/cfdo_api.h.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (``macos-cfd_b200``) never imports this module.

    lib = load("ref")    # oracle/_ref/libcfdo_ref.so   -- the reference's own sources (main oracle)
    lib = load("port")   # oracle/_build/libcfdo_port.so -- plain-C restatement (travels as source)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

WALL, MOVING, INFLOW, OUTFLOW, PERIODIC = range(5)  # bc.hpp:6
SOLVER_MG, SOLVER_PCG = 0, 1  # pressure.hpp:6
F_U, F_V, F_P, F_RHS, F_DU, F_DV, F_U1, F_V1 = range(8)


class BCSide(C.Structure):
    _fields_ = [("type", C.c_int), ("moving", C.c_double), ("inflow_u", C.c_double), ("inflow_v", C.c_double)]


class BC(C.Structure):
    _fields_ = [("left", BCSide), ("right", BCSide), ("bottom", BCSide), ("top", BCSide),
                ("jet_center", C.c_double), ("jet_width", C.c_double)]


class Shape(C.Structure):
    """gui.hpp:14-23 (pos / size in cell units)."""
    _fields_ = [("type", C.c_int), ("pos_x", C.c_float), ("pos_y", C.c_float), ("size_x", C.c_float),
                ("size_y", C.c_float), ("bc_type", C.c_int), ("enabled", C.c_int)]


FIELDS = {"u": 0, "v": 1, "speed": 2, "pressure": 3, "vorticity": 4, "streamlines": 5, "temperature": 6}  # viz.hpp:8
COLORMAPS = {"viridis": 0, "plasma": 1, "jet": 2, "grayscale": 3, "turbo": 4, "rdylbu": 5, "rdbu": 6, "coolwarm": 7}


class Params(C.Structure):
    _fields_ = [("solver", C.c_int), ("mg_levels", C.c_int), ("vcycles", C.c_int), ("rbgs_sweeps", C.c_int),
                ("tol", C.c_double), ("pcg_max_iters", C.c_int)]


def make_bc(left=WALL, right=WALL, bottom=WALL, top=WALL, *, left_inflow_u=1.0, jet_center=0.5, jet_width=0.2,
            **side_vals) -> BC:
    """BC with the reference's defaults (bc.hpp:15-22: left.inflow_u = 1.0, everything else 0)."""
    bc = BC()
    for name, t in (("left", left), ("right", right), ("bottom", bottom), ("top", top)):
        getattr(bc, name).type = t
    bc.left.inflow_u = left_inflow_u
    bc.jet_center = jet_center
    bc.jet_width = jet_width
    for k, val in side_vals.items():  # e.g. top_moving=1.0, right_inflow_v=0.3
        side, field = k.split("_", 1)
        setattr(getattr(bc, side), field, val)
    return bc


def preset(name: str, Ly: float = 1.0):
    """(bc, Re, CFL) of the reference's presets: main.cpp:131-142 / :273-304."""
    if name == "jet":
        return make_bc(INFLOW, OUTFLOW, WALL, WALL, left_inflow_u=1.0, jet_center=0.5 * Ly,
                       jet_width=0.18 * Ly), 50.0, 0.01
    if name == "lid":
        return make_bc(WALL, WALL, WALL, MOVING, top_moving=1.0), 1000.0, 0.1
    if name == "shear":
        return make_bc(PERIODIC, PERIODIC, WALL, WALL), 1000.0, 0.1
    raise ValueError(name)


def make_params(solver="pcg", vcycles=2, tol=1e-6, pcg_max_iters=200, mg_levels=3, rbgs_sweeps=2) -> Params:
    return Params(SOLVER_PCG if solver == "pcg" else SOLVER_MG, mg_levels, vcycles, rbgs_sweeps, tol,
                  pcg_max_iters)


def shear_init(u: np.ndarray, nx: int, ny: int, Ly: float, amp: float = 1.0) -> None:
    """Periodic-shear initial condition, main.cpp:348-358: u(all faces, row j) = amp*sin(2*pi*y/Ly)."""
    dy = Ly / ny
    for j in range(ny):
        y = (j + 0.5) * dy
        u[j + 1, 1:nx + 2] = amp * np.sin(2.0 * 3.141592653589793 * y / Ly)


def shapes(nx: int, ny: int):
    """(rows, pitch) of the reference host layout for u, v, p."""
    return {"u": (ny + 2, nx + 3), "v": (ny + 3, nx + 2), "p": (ny + 2, nx + 2)}


_dp = C.POINTER(C.c_double)


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def build(kind: str) -> str:
    """Build (if needed) and return the path of the requested checker library."""
    path = os.path.join(HERE, "_ref/libcfdo_ref.so" if kind == "ref" else "_build/libcfdo_port.so")
    if kind == "ref" and not os.path.isdir("/root/reference"):
        if os.path.exists(path):
            return path  # prebuilt, travelled with the snapshot
        raise FileNotFoundError("oracle/_ref is not prebuilt and /root/reference is absent")
    subprocess.run(["make", "-s", "-C", HERE, "ref" if kind == "ref" else "port"], check=True)
    return path


class Lib:
    def __init__(self, kind: str):
        self.kind = kind
        self.path = build(kind)
        L = self.L = C.CDLL(self.path)
        L.cfdo_kind.restype = C.c_char_p
        L.cfdo_create.restype = C.c_void_p
        L.cfdo_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.cfdo_destroy.argtypes = [C.c_void_p]
        L.cfdo_field.restype = _dp
        L.cfdo_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cfdo_clear_work.argtypes = [C.c_void_p]
        L.cfdo_step.restype = C.c_double
        L.cfdo_step.argtypes = [C.c_void_p, C.POINTER(BC), C.c_double, C.c_double, C.POINTER(Params),
                                C.c_double, _dp]
        L.cfdo_last_iters.restype = C.c_int
        L.cfdo_last_iters.argtypes = [C.c_void_p]
        g = [C.c_int, C.c_int, C.c_double, C.c_double]
        for n in ("u", "v", "p"):
            getattr(L, "cfdo_apply_bc_" + n).argtypes = g + [_dp, C.POINTER(BC)]
        L.cfdo_advect_u.argtypes = L.cfdo_advect_v.argtypes = g + [_dp, _dp, _dp]
        L.cfdo_diffuse_u.argtypes = L.cfdo_diffuse_v.argtypes = g + [_dp, _dp, C.c_double]
        L.cfdo_compute_cfl.restype = C.c_double
        L.cfdo_compute_cfl.argtypes = g + [_dp, _dp, C.c_double, C.c_double]
        L.cfdo_divergence.argtypes = g + [_dp, _dp, _dp]
        L.cfdo_subtract_grad_p.argtypes = g + [_dp, _dp, _dp, C.c_double]
        L.cfdo_pressure_solve.restype = C.c_double
        L.cfdo_pressure_solve.argtypes = g + [_dp, _dp, C.POINTER(Params), C.POINTER(C.c_int)]
        L.cfdo_divergence_l2.restype = L.cfdo_max_velocity.restype = C.c_double
        L.cfdo_divergence_l2.argtypes = L.cfdo_max_velocity.argtypes = g + [_dp, _dp]
        L.cfdo_sanitize.argtypes = [_dp, C.c_long]
        L.cfdo_compute_scalar.argtypes = g + [_dp, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.cfdo_field_stats.argtypes = [_dp, C.c_long, _dp, _dp]
        L.cfdo_apply_shapes.argtypes = g + [_dp, _dp, C.POINTER(Shape), C.c_int]
        L.cfdo_apply_colormap.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_ubyte)]
        L.cfdo_render_rgb.argtypes = g + [_dp, _dp, _dp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                          C.POINTER(C.c_ubyte)]

    # ---- stage-level wrappers (arrays are modified in place) ----
    def apply_bc(self, which, nx, ny, Lx, Ly, a, bc):
        getattr(self.L, "cfdo_apply_bc_" + which)(nx, ny, Lx, Ly, _ptr(a), C.byref(bc))

    def advect_u(self, nx, ny, Lx, Ly, u, v, du):
        self.L.cfdo_advect_u(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(du))

    def advect_v(self, nx, ny, Lx, Ly, u, v, dv):
        self.L.cfdo_advect_v(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(dv))

    def diffuse_u(self, nx, ny, Lx, Ly, u, du, invRe):
        self.L.cfdo_diffuse_u(nx, ny, Lx, Ly, _ptr(u), _ptr(du), invRe)

    def diffuse_v(self, nx, ny, Lx, Ly, v, dv, invRe):
        self.L.cfdo_diffuse_v(nx, ny, Lx, Ly, _ptr(v), _ptr(dv), invRe)

    def compute_cfl(self, nx, ny, Lx, Ly, u, v, Re, CFL):
        return self.L.cfdo_compute_cfl(nx, ny, Lx, Ly, _ptr(u), _ptr(v), Re, CFL)

    def divergence(self, nx, ny, Lx, Ly, u, v, rhs):
        self.L.cfdo_divergence(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(rhs))

    def subtract_grad_p(self, nx, ny, Lx, Ly, u, v, p, dt):
        self.L.cfdo_subtract_grad_p(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(p), dt)

    def pressure_solve(self, nx, ny, Lx, Ly, p, rhs, pp):
        it = C.c_int(-1)
        res = self.L.cfdo_pressure_solve(nx, ny, Lx, Ly, _ptr(p), _ptr(rhs), C.byref(pp), C.byref(it))
        return res, it.value

    def divergence_l2(self, nx, ny, Lx, Ly, u, v):
        return self.L.cfdo_divergence_l2(nx, ny, Lx, Ly, _ptr(u), _ptr(v))

    def max_velocity(self, nx, ny, Lx, Ly, u, v):
        return self.L.cfdo_max_velocity(nx, ny, Lx, Ly, _ptr(u), _ptr(v))

    def sanitize(self, a):
        self.L.cfdo_sanitize(_ptr(a), a.size)

    # ---- readers / mutators around the step (viz.cpp, shapes.cpp, colormaps.cpp, gui.cpp) ----
    def compute_scalar(self, nx, ny, Lx, Ly, u, v, p, field):
        """(out[ny, nx], vmin, vmax, mean, std): compute_scalar + the statistics of Gui::update_field_statistics."""
        out = np.empty((ny, nx))
        lo, hi, mean, std = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self.L.cfdo_compute_scalar(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(p), field, _ptr(out), C.byref(lo), C.byref(hi))
        self.L.cfdo_field_stats(_ptr(out), out.size, C.byref(mean), C.byref(std))
        return out, lo.value, hi.value, mean.value, std.value

    def apply_shapes(self, nx, ny, Lx, Ly, u, v, shapes):
        arr = (Shape * len(shapes))(*shapes)
        self.L.cfdo_apply_shapes(nx, ny, Lx, Ly, _ptr(u), _ptr(v), arr, len(shapes))

    def apply_colormap(self, colormap, t):
        rgb = (C.c_ubyte * 3)()
        self.L.cfdo_apply_colormap(colormap, t, rgb)
        return tuple(rgb)

    def render_rgb(self, nx, ny, Lx, Ly, u, v, p, field, colormap, scale_min=0.0, scale_max=1.0, normalize=False,
                   gamma=1.0):
        rgb = np.empty((ny, nx, 3), dtype=np.uint8)
        self.L.cfdo_render_rgb(nx, ny, Lx, Ly, _ptr(u), _ptr(v), _ptr(p), field, colormap, scale_min, scale_max,
                               int(normalize), gamma, rgb.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return rgb

    def vcycle_solve(self, nx, ny, Lx, Ly, p, rhs, mg_levels=3, vcycles=2, rbgs_sweeps=2, tol=1e-6):
        """True geometric V-cycle (oracle/cfd_vcycle.c; port only -- the reference has no such code)."""
        f = self.L.cfdo_vcycle_solve
        f.restype = C.c_double
        f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double,
                      C.POINTER(C.c_int)]
        it = C.c_int()
        res = f(nx, ny, Lx, Ly, _ptr(p), _ptr(rhs), mg_levels, vcycles, rbgs_sweeps, tol, C.byref(it))
        return res, it.value


class Sim:
    """A whole CPU simulation (Grid + State + PressureSolver + TimeIntegrator), main.cpp:104-130."""

    def __init__(self, lib: Lib, nx, ny, Lx, Ly):
        self.lib, self.nx, self.ny, self.Lx, self.Ly = lib, nx, ny, Lx, Ly
        self.h = lib.L.cfdo_create(nx, ny, Lx, Ly)
        self.last_residual = 0.0

    def field(self, which) -> np.ndarray:
        """numpy VIEW of the simulation's own host array (rows x pitch)."""
        pitch, rows = C.c_int(), C.c_int()
        ptr = self.lib.L.cfdo_field(self.h, which, C.byref(pitch), C.byref(rows))
        return np.ctypeslib.as_array(ptr, shape=(rows.value, pitch.value))

    u = property(lambda self: self.field(F_U))
    v = property(lambda self: self.field(F_V))
    p = property(lambda self: self.field(F_P))
    rhs = property(lambda self: self.field(F_RHS))

    def step(self, bc, Re, CFL, pp, dt_override=0.0) -> float:
        res = C.c_double()
        dt = self.lib.L.cfdo_step(self.h, C.byref(bc), Re, CFL, C.byref(pp), dt_override, C.byref(res))
        self.last_residual = res.value
        return dt

    @property
    def last_iters(self) -> int:
        return self.lib.L.cfdo_last_iters(self.h)

    def close(self):
        if self.h:
            self.lib.L.cfdo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


_cache: dict[str, Lib] = {}


def load(kind: str) -> Lib:
    if kind not in _cache:
        _cache[kind] = Lib(kind)
    return _cache[kind]


def have_ref() -> bool:
    return os.path.isdir("/root/reference") or os.path.exists(os.path.join(HERE, "_ref/libcfdo_ref.so"))
