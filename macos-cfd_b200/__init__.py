"""macos-cfd_b200 -- B200-native per-timestep solver path of archhie/MacOS-CFD (src/solver/).

The product is ``lib/libcfd_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/cfd_b200.h``).  This module is the thin Python view of that ABI used by the tests and bench.py; the
reference-shaped C++ classes live in ``cpp/``.  There is no CPU fallback: creating a ``Solver`` without a CUDA
device raises ``CfdError``.

Import with ``importlib.import_module("macos-cfd_b200")`` (the directory name carries a hyphen).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

# BCType (bc.hpp:6) and PressureSolverType (pressure.hpp:6)
WALL, MOVING, INFLOW, OUTFLOW, PERIODIC = range(5)
SOLVER_MG, SOLVER_PCG = 0, 1
MG_REFERENCE, MG_VCYCLE = 0, 1


class CfdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cfd_b200 error {code}: {msg}")
        self.code = code


class BCSide(C.Structure):
    _fields_ = [("type", C.c_int), ("moving", C.c_double), ("inflow_u", C.c_double), ("inflow_v", C.c_double)]


class BC(C.Structure):
    """bc.hpp:15-27 (live members only)."""
    _fields_ = [("left", BCSide), ("right", BCSide), ("bottom", BCSide), ("top", BCSide),
                ("jet_center", C.c_double), ("jet_width", C.c_double)]


class Params(C.Structure):
    """PressureParams, pressure.hpp:8-15 (+ mg_mode)."""
    _fields_ = [("solver", C.c_int), ("mg_levels", C.c_int), ("vcycles", C.c_int), ("rbgs_sweeps", C.c_int),
                ("tol", C.c_double), ("pcg_max_iters", C.c_int), ("mg_mode", C.c_int)]


class Shape(C.Structure):
    """cfd_b200_shape = Shape of gui.hpp:14-23 (pos / size in cell units)."""
    _fields_ = [("type", C.c_int), ("pos_x", C.c_float), ("pos_y", C.c_float), ("size_x", C.c_float),
                ("size_y", C.c_float), ("bc_type", C.c_int), ("enabled", C.c_int)]


class FieldStats(C.Structure):
    _fields_ = [("vmin", C.c_double), ("vmax", C.c_double), ("mean", C.c_double), ("stddev", C.c_double)]


FIELDS = {"u": 0, "v": 1, "speed": 2, "pressure": 3, "vorticity": 4, "streamlines": 5, "temperature": 6}  # viz.hpp:8


class StepInfo(C.Structure):
    _fields_ = [("dt", C.c_double), ("residual", C.c_double), ("iters", C.c_int), ("pcg_breakdown", C.c_int),
                ("div_before", C.c_double), ("div_after", C.c_double), ("nonfinite", C.c_int)]


def make_bc(left=WALL, right=WALL, bottom=WALL, top=WALL, *, left_inflow_u=1.0, jet_center=0.5, jet_width=0.2,
            **side_vals) -> BC:
    """BC with the reference's defaults (bc.hpp:15-22: left.inflow_u = 1.0, jet 0.5 / 0.2, the rest 0)."""
    bc = BC()
    for name, t in (("left", left), ("right", right), ("bottom", bottom), ("top", top)):
        getattr(bc, name).type = t
    bc.left.inflow_u = left_inflow_u
    bc.jet_center, bc.jet_width = jet_center, jet_width
    for k, val in side_vals.items():
        side, field = k.split("_", 1)
        setattr(getattr(bc, side), field, val)
    return bc


def preset(name: str, Ly: float = 1.0):
    """(bc, Re, CFL) of the reference's presets: Jet Plume main.cpp:131-142/:273-287, Lid-Driven Cavity
    :288-296, Periodic Shear :297-304."""
    if name == "jet":
        return make_bc(INFLOW, OUTFLOW, WALL, WALL, left_inflow_u=1.0, jet_center=0.5 * Ly,
                       jet_width=0.18 * Ly), 50.0, 0.01
    if name == "lid":
        return make_bc(WALL, WALL, WALL, MOVING, top_moving=1.0), 1000.0, 0.1
    if name == "shear":
        return make_bc(PERIODIC, PERIODIC, WALL, WALL), 1000.0, 0.1
    raise ValueError(name)


def make_params(solver="pcg", vcycles=2, tol=1e-6, pcg_max_iters=200, mg_levels=3, rbgs_sweeps=2,
                mg_mode=MG_REFERENCE) -> Params:
    """Defaults of the headless driver (main.cpp:56-59): PCG, 200 iterations, tol 1e-6."""
    return Params(SOLVER_PCG if solver == "pcg" else SOLVER_MG, mg_levels, vcycles, rbgs_sweeps, tol, pcg_max_iters,
                  mg_mode)


def shear_initial_u(nx: int, ny: int, Ly: float, amp: float = 1.0) -> np.ndarray:
    """Periodic-shear initial condition (main.cpp:348-358) in the reference host layout of u."""
    u = np.zeros((ny + 2, nx + 3))
    dy = Ly / ny
    for j in range(ny):
        y = (j + 0.5) * dy
        u[j + 1, 1:nx + 2] = amp * np.sin(2.0 * 3.141592653589793 * y / Ly)
    return u


_dp = C.POINTER(C.c_double)
_lib = None


def lib_path() -> str:
    return _build.LIB


def load_library(build: bool = True):
    """Load libcfd_b200.so (building it with nvcc first when sources are newer).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CFD_B200_LIB")  # tuning experiments: an alternative build of the same sources
    if not path:
        if build and (os.path.exists("/usr/local/cuda/bin/nvcc") or not os.path.exists(_build.LIB)):
            _build.build()
        path = _build.LIB
    if not os.path.exists(path):
        raise CfdError(-2, f"{path} is missing and cannot be built: there is no CPU fallback")
    L = C.CDLL(path)
    S = C.c_void_p
    L.cfd_b200_last_error.restype = C.c_char_p
    L.cfd_b200_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(S)]
    L.cfd_b200_create_slab.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.POINTER(S)]
    L.cfd_b200_comm_id.argtypes = [C.c_void_p]
    L.cfd_b200_slab_rows.argtypes = [S, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.cfd_b200_comm_mode.argtypes = [S]
    L.cfd_b200_last_smoother.argtypes = [S]
    L.cfd_b200_last_pcg.argtypes = [S]
    L.cfd_b200_compute_scalar.argtypes = [S, C.c_int, C.c_void_p, C.POINTER(FieldStats)]
    L.cfd_b200_render_rgb.argtypes = [S, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
    L.cfd_b200_apply_shapes.argtypes = [S, C.POINTER(Shape), C.c_int]
    L.cfd_b200_set_smoother_strip.argtypes = [S, C.c_int]
    L.cfd_b200_comm_reset.argtypes = [S]
    L.cfd_b200_wait_stats.argtypes = [S, _dp, C.POINTER(C.c_longlong), C.c_int]
    L.cfd_b200_debug_comm.argtypes = [S, C.c_int, C.c_int]
    L.cfd_b200_debug_delay.argtypes = [S, C.c_double]
    L.cfd_b200_smoother_strip_for.argtypes = [C.c_int, C.c_int, C.c_int]
    L.cfd_b200_destroy.argtypes = [S]
    L.cfd_b200_destroy.restype = None
    L.cfd_b200_upload.argtypes = [S, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cfd_b200_upload_rhs.argtypes = [S, C.c_void_p]
    L.cfd_b200_download.argtypes = [S, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cfd_b200_download_work.argtypes = [S, C.c_void_p, C.c_void_p]
    L.cfd_b200_clear_work.argtypes = [S]
    L.cfd_b200_reset.argtypes = [S]
    L.cfd_b200_step.argtypes = [S, C.POINTER(BC), C.c_double, C.c_double, C.POINTER(Params), C.c_double,
                                C.POINTER(StepInfo)]
    L.cfd_b200_step_async.argtypes = [S, C.POINTER(BC), C.c_double, C.c_double, C.POINTER(Params), C.c_double]
    L.cfd_b200_sync.argtypes = [S, C.POINTER(StepInfo)]
    L.cfd_b200_divergence_l2.argtypes = [S, _dp]
    L.cfd_b200_max_velocity.argtypes = [S, _dp]
    L.cfd_b200_apply_bc.argtypes = [S, C.POINTER(BC), C.c_int]
    L.cfd_b200_compute_cfl.argtypes = [S, C.c_double, C.c_double, _dp]
    L.cfd_b200_rhs_terms.argtypes = [S, C.c_double, C.c_void_p, C.c_void_p]
    L.cfd_b200_advect.argtypes = [S, C.c_void_p, C.c_void_p]
    L.cfd_b200_diffuse.argtypes = [S, C.c_int, C.c_double, C.c_void_p]
    L.cfd_b200_divergence.argtypes = [S]
    L.cfd_b200_pressure_solve.argtypes = [S, C.POINTER(Params), _dp, C.POINTER(C.c_int)]
    L.cfd_b200_subtract_grad_p.argtypes = [S, C.c_double]
    L.cfd_b200_launch_count.argtypes = [S]
    L.cfd_b200_launch_count.restype = C.c_longlong
    L.cfd_b200_set_profiling.argtypes = [S, C.c_int]
    L.cfd_b200_stage_times.argtypes = [S, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.cfd_b200_stream.argtypes = [S]
    L.cfd_b200_stream.restype = C.c_void_p
    L.cfd_b200_set_variant.argtypes = [S, C.c_int]
    L.cfd_b200_set_march_kernel.argtypes = [S, C.c_int]
    _lib = L
    return L


def _addr(a):
    """Host address of a numpy array, a torch CPU tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (pinned host memory in bench.py)


def comm_id() -> bytes:
    """A fresh 128-byte NCCL unique id (call on rank 0, hand the bytes to every rank's Solver)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.cfd_b200_comm_id(buf)
    if rc != 0:
        raise CfdError(rc, L.cfd_b200_last_error().decode())
    return buf.raw


def slab_rows(ny: int, rank: int, nranks: int):
    """(row0, nrows) of rank's slab -- the same partition the library uses (csrc/comm.cuh slab_partition)."""
    a, b = ny * rank // nranks, ny * (rank + 1) // nranks
    return a, b - a


class Solver:
    """Grid + device State + PressureSolver + TimeIntegrator of the reference (main.cpp:104-130) on one B200."""

    def __init__(self, nx, ny, Lx=1.0, Ly=1.0, device=0, rank=0, nranks=1, comm_id: bytes | None = None):
        """Single GPU by default.  With nranks > 1 this is one row slab of the nx x ny grid (one process per GPU);
        comm_id is the 128-byte id from ``comm_id()`` on rank 0, distributed by the caller."""
        self.L = load_library()
        self.nx, self.ny, self.Lx, self.Ly = nx, ny, Lx, Ly
        self.rank, self.nranks = rank, nranks
        self.h = C.c_void_p()
        if nranks == 1:
            self._check(self.L.cfd_b200_create(nx, ny, Lx, Ly, device, C.byref(self.h)))
        else:
            buf = C.create_string_buffer(comm_id, 128)
            self._check(self.L.cfd_b200_create_slab(nx, ny, Lx, Ly, device, rank, nranks, buf, C.byref(self.h)))
        r0, nr = C.c_int(), C.c_int()
        self._check(self.L.cfd_b200_slab_rows(self.h, C.byref(r0), C.byref(nr)))
        self.row0, self.nrows = r0.value, nr.value  # owned cell rows [row0, row0 + nrows)

    @property
    def comm_mode(self) -> str:
        """How the slabs talk: "single", "nccl" (send/recv + allreduce) or "p2p" (peer memory over NVLink, CUDA IPC)."""
        return ("single", "nccl", "p2p")[self.L.cfd_b200_comm_mode(self.h)]

    def comm_reset(self):
        """Collective over all slabs: clear a communication failure (CfdError -5) and restart the flag protocol."""
        self._check(self.L.cfd_b200_comm_reset(self.h))

    def wait_stats(self, reset=True):
        """{kind: (microseconds spent polling for a peer, polls that had to wait)} since the last reset (row slabs)."""
        us, n = (C.c_double * 4)(), (C.c_longlong * 4)()
        self._check(self.L.cfd_b200_wait_stats(self.h, us, n, int(reset)))
        return {k: (us[i], n[i]) for i, k in enumerate(("wait_kernel", "edge_blocks", "reduce", "handshake"))}

    def debug_comm(self, kind: int, reps: int):
        """Measurement hook: enqueue `reps` slab-communication operations (see cfd_b200_debug_comm)."""
        self._check(self.L.cfd_b200_debug_comm(self.h, kind, reps))

    def debug_delay(self, microseconds: float):
        """Test hook: keep this slab's stream busy for that long (makes it late; its neighbours run ahead)."""
        self._check(self.L.cfd_b200_debug_delay(self.h, microseconds))

    def set_smoother_strip(self, rows: int):
        """rows > 0: force the streaming smoother with that strip height (tests / tuning); 0: automatic."""
        self._check(self.L.cfd_b200_set_smoother_strip(self.h, rows))

    @property
    def last_smoother(self) -> str:
        """Kernel that ran the last smooth() of the reference-mode mg solver."""
        return ("none", "k_rbgs_stream", "k_rbgs_fused", "k_rbgs_colour x4 + k_rbgs_residual")[self.L.cfd_b200_last_smoother(self.h)]

    @property
    def last_pcg(self) -> str:
        """Kernels that ran the last PCG solve."""
        return ("none", "k_pcg_coop (whole solve, one cooperative launch)", "k_pcg_phase1_march + k_pcg_phase2_march (one iteration)",
                "k_pcg_phase1 + k_pcg_phase2 (one iteration)",
                "k_pcg_cluster (whole solve, one 16-CTA cluster)")[self.L.cfd_b200_last_pcg(self.h)]

    def _check(self, rc):
        if rc != 0:
            raise CfdError(rc, self.L.cfd_b200_last_error().decode())

    # reference host layout
    def shapes(self):
        """Host arrays of THIS handle: the whole grid, or the slab's rows plus one ghost/halo row on either side."""
        n = self.nrows
        return {"u": (n + 2, self.nx + 3), "v": (n + 3, self.nx + 2), "p": (n + 2, self.nx + 2)}

    def upload(self, u=None, v=None, p=None, rhs=None):
        self._check(self.L.cfd_b200_upload(self.h, _addr(u), _addr(v), _addr(p)))
        if rhs is not None:
            self._check(self.L.cfd_b200_upload_rhs(self.h, _addr(rhs)))

    def download(self, u=None, v=None, p=None, rhs=None):
        """Fill the given host arrays; with no arguments return fresh (u, v, p, rhs)."""
        if u is None and v is None and p is None and rhs is None:
            sh = self.shapes()
            u, v, p, rhs = np.empty(sh["u"]), np.empty(sh["v"]), np.empty(sh["p"]), np.empty(sh["p"])
            self._check(self.L.cfd_b200_download(self.h, _addr(u), _addr(v), _addr(p), _addr(rhs)))
            return u, v, p, rhs
        self._check(self.L.cfd_b200_download(self.h, _addr(u), _addr(v), _addr(p), _addr(rhs)))

    def download_work(self):
        sh = self.shapes()
        u1, v1 = np.empty(sh["u"]), np.empty(sh["v"])
        self._check(self.L.cfd_b200_download_work(self.h, _addr(u1), _addr(v1)))
        return u1, v1

    def clear_work(self):
        self._check(self.L.cfd_b200_clear_work(self.h))

    def reset(self):
        """Device-side zero of the State and work arrays (the GUI's reset, main.cpp:338-346); asynchronous."""
        self._check(self.L.cfd_b200_reset(self.h))

    def step(self, bc, Re, CFL, pp, dt_override=0.0) -> StepInfo:
        info = StepInfo()
        self._check(self.L.cfd_b200_step(self.h, C.byref(bc), Re, CFL, C.byref(pp), dt_override, C.byref(info)))
        return info

    def step_async(self, bc, Re, CFL, pp, dt_override=0.0):
        self._check(self.L.cfd_b200_step_async(self.h, C.byref(bc), Re, CFL, C.byref(pp), dt_override))

    def sync(self) -> StepInfo:
        info = StepInfo()
        self._check(self.L.cfd_b200_sync(self.h, C.byref(info)))
        return info

    def divergence_l2(self) -> float:
        out = C.c_double()
        self._check(self.L.cfd_b200_divergence_l2(self.h, C.byref(out)))
        return out.value

    def max_velocity(self) -> float:
        out = C.c_double()
        self._check(self.L.cfd_b200_max_velocity(self.h, C.byref(out)))
        return out.value

    # readers / mutators around the step (viz.cpp, gui.cpp, shapes.cpp)
    def compute_scalar(self, field: int, want_field=True):
        """(out[ny, nx] or None, vmin, vmax, mean, std) of the visualisation scalar, computed on the device."""
        out = np.empty((self.nrows, self.nx)) if want_field else None
        st = FieldStats()
        self._check(self.L.cfd_b200_compute_scalar(self.h, field, _addr(out), C.byref(st)))
        return out, st.vmin, st.vmax, st.mean, st.stddev

    def render_rgb(self, field: int, colormap: int, scale_min=0.0, scale_max=1.0, normalize=False, gamma=1.0):
        rgb = np.empty((self.nrows, self.nx, 3), dtype=np.uint8)
        self._check(self.L.cfd_b200_render_rgb(self.h, field, colormap, scale_min, scale_max, int(normalize), gamma,
                                               rgb.ctypes.data))
        return rgb

    def apply_shapes(self, shapes):
        arr = (Shape * max(len(shapes), 1))(*shapes)
        self._check(self.L.cfd_b200_apply_shapes(self.h, arr, len(shapes)))

    # stage-level
    def apply_bc(self, bc, which=7):
        self._check(self.L.cfd_b200_apply_bc(self.h, C.byref(bc), which))

    def compute_cfl(self, Re, CFL) -> float:
        out = C.c_double()
        self._check(self.L.cfd_b200_compute_cfl(self.h, Re, CFL, C.byref(out)))
        return out.value

    def rhs_terms(self, invRe):
        sh = self.shapes()
        du, dv = np.empty(sh["u"]), np.empty(sh["v"])
        self._check(self.L.cfd_b200_rhs_terms(self.h, invRe, _addr(du), _addr(dv)))
        return du, dv

    def advect(self):
        sh = self.shapes()
        du, dv = np.empty(sh["u"]), np.empty(sh["v"])
        self._check(self.L.cfd_b200_advect(self.h, _addr(du), _addr(dv)))
        return du, dv

    def diffuse(self, which, invRe, d):
        self._check(self.L.cfd_b200_diffuse(self.h, which, invRe, _addr(d)))

    def divergence(self):
        self._check(self.L.cfd_b200_divergence(self.h))

    def pressure_solve(self, pp):
        res, it = C.c_double(), C.c_int()
        self._check(self.L.cfd_b200_pressure_solve(self.h, C.byref(pp), C.byref(res), C.byref(it)))
        return res.value, it.value

    def subtract_grad_p(self, dt):
        self._check(self.L.cfd_b200_subtract_grad_p(self.h, dt))

    # measurement
    @property
    def launch_count(self) -> int:
        return self.L.cfd_b200_launch_count(self.h)

    @property
    def stream(self) -> int:
        return self.L.cfd_b200_stream(self.h) or 0

    def set_profiling(self, on=True):
        self._check(self.L.cfd_b200_set_profiling(self.h, int(on)))

    def set_variant(self, v):
        self._check(self.L.cfd_b200_set_variant(self.h, int(v)))

    def set_march_kernel(self, version):
        """2 = k_rhs_march2 (default), 1 = the first marching kernel (A/B timing; identical results)."""
        self._check(self.L.cfd_b200_set_march_kernel(self.h, int(version)))

    def stage_times(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        n = self.L.cfd_b200_stage_times(self.h, 16, names, ms)
        if n < 0:
            self._check(n)
        return {names[i].decode(): ms[i] for i in range(n)}

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.cfd_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
