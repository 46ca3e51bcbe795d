"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/cfd_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cfd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfd_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    m = importlib.import_module("macos-cfd_b200")
    L = m.load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(L, name), f"{name} declared in include/cfd_b200.h but not exported"
    assert L.cfd_b200_abi_version() == 1


def test_struct_layouts_match_header():
    m = importlib.import_module("macos-cfd_b200")
    assert ctypes.sizeof(m.BCSide) == 32 and ctypes.sizeof(m.BC) == 4 * 32 + 16
    assert ctypes.sizeof(m.Params) == 32 and ctypes.sizeof(m.StepInfo) == 48


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = importlib.import_module("macos-cfd_b200")
    with pytest.raises(m.CfdError) as e:
        m.Solver(16, 16)
    assert e.value.code == -2  # CFD_B200_ENODEVICE


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing in the product package may reference it."""
    pkg = os.path.join(ROOT, "macos-cfd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cfdo_" not in txt and "oracle/" not in txt and "import oracle" not in txt, f


def test_no_fma_in_device_code():
    """-fmad=false: the SASS of the library holds no DFMA (SURVEY §7 hard part 1)."""
    import shutil
    import subprocess
    m = importlib.import_module("macos-cfd_b200")
    m.load_library()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", m.lib_path()], capture_output=True, text=True).stdout
    assert "DMUL" in sass and "DADD" in sass
    # IEEE division and sqrt expand to FMA-based Newton iterations inside the CUDA math library (correctly rounded,
    # so allowed); kernels whose last block finishes a reduction with sqrt or a quotient therefore hold a few DFMA.
    # The kernels without any division must hold none.
    for kern in ("k_rhs_simple", "k_project", "k_absmax", "k_bc_lr", "k_bc_bt", "k_rbgs_colourILi1"):
        blocks = [b for b in sass.split("Function : ")[1:] if kern in b.split("\n", 1)[0]]
        assert blocks, kern
        for b in blocks:
            assert "DFMA" not in b, kern


def test_argument_validation_happens_before_the_device_is_touched():
    """Bad arguments are CFD_B200_EINVAL (-1) even on a box without a GPU; errors carry a message."""
    import ctypes as C
    m = importlib.import_module("macos-cfd_b200")
    L = m.load_library()
    h = C.c_void_p()
    for nx, ny, Lx, Ly in ((3, 16, 1.0, 1.0), (16, 2, 1.0, 1.0), (16, 16, 0.0, 1.0), (16, 16, 1.0, -1.0)):
        assert L.cfd_b200_create(nx, ny, Lx, Ly, 0, C.byref(h)) == -1
        assert L.cfd_b200_last_error()
    assert L.cfd_b200_create(16, 16, 1.0, 1.0, 0, None) == -1
    assert L.cfd_b200_create_slab(64, 64, 1.0, 1.0, 0, 3, 2, None, C.byref(h)) == -1   # rank out of range
    assert L.cfd_b200_step(None, None, 1.0, 1.0, None, 0.0, None) == -1
    assert L.cfd_b200_upload(None, None, None, None) == -1
    assert L.cfd_b200_launch_count(None) == 0


def test_smoother_strip_model_picks_the_measured_optima():
    """Host-only arithmetic: the strip height of the streaming smoother against the B200 sweep in
    profiles/r02_strip_sweep.txt (148 SMs, round-2 kernel); 0 = grid too small, the tile smoother runs."""
    import importlib
    L = importlib.import_module("macos-cfd_b200").load_library()
    want = {(8192, 8192): 257, (8192, 4096): 193, (8192, 2048): 193, (4096, 4096): 193, (8192, 1024): 97,
            (2048, 2048): 49, (16384, 4096): 257, (1024, 1024): 0, (512, 256): 0}
    for (nx, ny), ry in want.items():
        assert L.cfd_b200_smoother_strip_for(nx, ny, 148) == ry, (nx, ny)
    assert L.cfd_b200_smoother_strip_for(2, 2, 148) < 0  # argument validation
