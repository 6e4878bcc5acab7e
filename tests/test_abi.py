"""The C-ABI library loads and exports every symbol include/gsr.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from gaustudio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gsr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", src))
    names.discard("gsr_alloc_fn")
    return names


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert {"gsr_forward", "gsr_backward", "gsr_mark_visible", "gsr_depth2normal", "gsr_last_error"} <= names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/gsr.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= names


def test_host_only_entry_points():
    L = _lib.lib()
    assert L.gsr_abi_version() == 1
    g1, g2 = L.gsr_geometry_bytes(1000), L.gsr_geometry_bytes(2000)
    assert 0 < g1 < g2 and g2 < 2 * g1 + 8192
    assert L.gsr_geometry_bytes(1_000_000) < 144 * 1_000_000  # ~141 B per Gaussian of forward+backward state
    i1 = L.gsr_image_bytes(1920, 1080)
    assert 8 * 1920 * 1080 <= i1 < 9 * 1920 * 1080 + 1_000_000
    assert L.gsr_binning_bytes(0) >= 0 and 20 * 2**20 <= L.gsr_binning_bytes(10**6) < 21 * 2**20  # 20 B per instance, 1 Mi granules
    assert L.gsr_binning_bytes(2**20 + 1) > L.gsr_binning_bytes(2**20) == L.gsr_binning_bytes(2**20 - 5)
    assert isinstance(_lib.last_error(), str)


def test_speculation_switch_round_trip():
    """gsr_set_speculation / gsr_speculation_stats are host-only: the switch returns the previous setting (None =
    default), the counters start at zero in a process that has rendered nothing, and the default build carries no
    experimental cluster-scan kernel."""
    from gaustudio_b200 import _C
    prev = _C.set_speculation(False)
    try:
        assert _C.set_speculation(True) is False
        assert _C.set_speculation(None) is True
        assert _C.set_speculation(None) is None
        hits, redos = _C.speculation_stats()
        assert hits >= 0 and redos >= 0
    finally:
        _C.set_speculation(prev)
    blob = open(_lib.LIB_PATH, "rb").read()
    assert b"k_tile_scan_cluster" not in blob


def test_no_oracle_import_in_product_path():
    """The product package must not reference oracle/ (a CPU fallback would void the parity claims)."""
    pkg = os.path.join(ROOT, "gaustudio_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "gsr_oracle" not in txt, f


def test_header_is_plain_c():
    """include/gsr.h is a C header: it must pass a C compiler on its own."""
    import subprocess
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", os.path.join(ROOT, "include", "gsr.h")],
                   check=True)


def test_reference_side_binding_compiles_against_the_header():
    """integration/rasterize_points_gsr.cpp -- what the reference's rasterize_points.cu becomes on top of the C ABI
     -- type-checks against include/gsr.h, the torch headers and the CUDA runtime headers."""
    import subprocess
    import sysconfig
    from torch.utils import cpp_extension
    inc = [os.path.join(ROOT, "include"), sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    inc += cpp_extension.include_paths()
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    for i in inc:
        cmd += ["-I", i]
    cmd.append(os.path.join(ROOT, "integration", "rasterize_points_gsr.cpp"))
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
