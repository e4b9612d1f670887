"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/rangelib_b200.h declares (and the ctypes table mirrors the header one to one), fails loudly instead
of falling back to a CPU path, and the host-side helpers (map handles, PNG ingest mirror) behave."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

import range_libc_b200 as rl
from range_libc_b200 import cabi, mapio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "rangelib_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "librangelib_b200.so does not export %s" % s
    assert sorted(cabi.SIGNATURES.keys()) == syms, "cabi.SIGNATURES and the header disagree"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = rl.PyOMap(np.zeros((16, 16), bool))
    for ctor in (lambda: rl.PyRayMarchingGPU(m, 100.0), lambda: rl.PyBresenhamsLine(m, 100.0),
                 lambda: rl.PyCDDTCast(m, 100.0, 16)):
        with pytest.raises(rl.RangeLibError) as ei:
            ctor()
        assert ei.value.code == cabi.RL_E_NO_DEVICE


def test_map_handle_host_side():
    arr = np.zeros((6, 5), bool)  # 6 rows (y), 5 cols (x)
    arr[2, 3] = True
    m = rl.PyOMap(arr)
    assert (m.width(), m.height()) == (5, 6)
    assert m.isOccupied(3, 2) and not m.isOccupied(2, 3)
    assert not m.isOccupied(-1, 0) and not m.isOccupied(5, 0)  # out of bounds -> False (RangeLib.h:204-210)
    m.update(np.ones((2, 2), np.uint8), 0, 0)
    assert m.grid()[:2, :2].all() and m.grid().sum() == 5
    with pytest.raises(rl.RangeLibError):
        m.update(np.ones((2, 2), np.uint8), 4, 5)  # patch sticks out of the map


def test_png_ingest_formula():
    """gray = (int)(float)(0.229*byte2 + 0.587*byte1 + 0.114*byte0), occupied iff gray < threshold."""
    rgba = np.zeros((2, 3, 4), np.uint8)
    rgba[0, 0] = (255, 255, 255, 255)  # 0.93*255 = 237.15 -> 237 : free
    rgba[0, 1] = (0, 0, 0, 255)        # 0 : occupied
    rgba[0, 2] = (0, 0, 255, 255)      # byte2 = 255 -> 0.229*255 = 58.395 -> 58 : occupied at 128, free at 58
    rgba[1, 0] = (255, 0, 0, 255)      # byte0 = 255 -> 0.114*255 = 29.07 -> 29
    rgba[1, 1] = (0, 218, 0, 255)      # 0.587*218 = 127.966 -> 127 : occupied at 128
    rgba[1, 2] = (0, 219, 0, 255)      # 128.553 -> 128 : free at 128
    occ = mapio.occupancy_from_rgba(rgba, 128)
    assert occ.shape == (3, 2)
    assert occ.T.tolist() == [[0, 1, 1], [1, 1, 0]]
    assert mapio.occupancy_from_rgba(rgba, 58).T.tolist() == [[0, 1, 0], [1, 0, 0]]


@pytest.mark.skipif(not os.path.isdir("/root/reference/maps"), reason="reference maps not present")
def test_png_ingest_equals_reference_loader():
    from oracle import ref
    if not ref.available("strict"):
        pytest.skip("oracle/_ref not built")
    for f in sorted(glob.glob("/root/reference/maps/*.png")):
        if "gigantic" in f or "huge" in f:
            continue
        thr = 1 if "synthetic.map" in f else 128
        assert np.array_equal(ref.RefMap(png=f, threshold=thr).occ(), mapio.load_png(f, thr)), f


def test_cpp_header_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    """include/rangelib_b200.hpp (C++ mirror of ranges::RangeMethod & co.) builds against the library; without a
    GPU constructing a method throws (exit code 3), with one it casts rays (exit code 0)."""
    import shutil
    import subprocess
    import torch
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "hpp_smoke")
    libdir = os.path.dirname(cabi.LIB_PATH)
    subprocess.check_call([gxx, "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "hpp_smoke.cpp"), "-o", exe, "-L", libdir,
                           "-lrangelib_b200", "-Wl,-rpath," + libdir])
    rc = subprocess.run([exe], capture_output=True, text=True)
    assert rc.returncode == (0 if torch.cuda.is_available() else 3), rc.stdout + rc.stderr


def test_cython_drop_in_exposes_the_reference_module_surface():
    """`range_libc` (pywrapper/RangeLibc.pyx) must import without a GPU and carry every class, method and module
    constant a caller of the reference's module can name (RangeLibc.pyx:94-352 of the reference)."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
    try:
        import range_libc as m
    except ImportError:
        import pytest
        pytest.skip("Cython extension not built")
    for const in ("USE_CACHED_TRIG", "USE_ALTERNATE_MOD", "USE_CACHED_CONSTANTS", "USE_FAST_ROUND", "NO_INLINE", "USE_LRU_CACHE",
                  "LRU_CACHE_SIZE", "SHOULD_USE_CUDA"):
        assert hasattr(m, const), const
    common = ["calc_range", "calc_range_many", "calc_range_repeat_angles", "calc_range_repeat_angles_eval_sensor_model",
              "eval_sensor_model", "set_sensor_model"]
    surface = {"PyOMap": ["save", "isOccupied", "error", "width", "height"],
               "PyBresenhamsLine": common + ["saveTrace"], "PyRayMarching": common + ["saveTrace"],
               "PyRayMarchingGPU": common, "PyCDDTCast": common + ["prune", "calc_range_many_radial_optimized"],
               "PyGiantLUTCast": common, "PyNull": ["calc_range", "calc_range_many"]}
    for cls, methods in surface.items():
        assert hasattr(m, cls), cls
        for name in methods:
            assert hasattr(getattr(m, cls), name), "%s.%s" % (cls, name)
    omap = m.PyOMap(np.zeros((3, 4), dtype=bool))  # arr[row = y, col = x]
    assert (omap.width(), omap.height()) == (4, 3) and not omap.isOccupied(1, 1) and not omap.error()
