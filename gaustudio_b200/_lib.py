"""ctypes binding of libgsr_b200.so -- the C ABI declared in include/gsr.h.

There is NO CPU or PyTorch fallback: if the CUDA library is missing or does not export the ABI, importing
this module raises.  (oracle/ is test infrastructure and is never imported from here.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSR_LIB selects another build of the same sources (the ring-stress variant of the tests); default: the in-tree library
LIB_PATH = os.environ.get("GSR_LIB") or os.path.join(_HERE, "libgsr_b200.so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64

SIGNATURES = {
    "gsr_abi_version": (_I, []),
    "gsr_last_error": (C.c_char_p, []),
    "gsr_geometry_bytes": (C.c_size_t, [_I]),
    "gsr_image_bytes": (C.c_size_t, [_I, _I]),
    "gsr_binning_bytes": (C.c_size_t, [_L]),
    "gsr_forward": (_L, [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _F, _P,
                         _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _I, _L, _P, _P]),
    "gsr_backward": (_I, [_I, _I, _I, _L, _P, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P,
                          _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "gsr_forward_fused": (_L, [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _F, _P,
                               _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _I, _L, _P, _P]),
    "gsr_backward_fused": (_I, [_I, _I, _I, _L, _P, _I, _I, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _F, _F, _P, _P, _P,
                                _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "gsr_mark_visible": (_I, [_I, _P, _P, _P, _P, _P]),
    "gsr_depth2normal": (_I, [_P, _I, _I, _F, _F, _F, _F, _F, _F, _P, _P, _P]),
    "gsr_depth2point": (_I, [_P, _I, _I, _F, _F, _F, _F, _P, _P, _P]),
    "gsr_masked_bilateral": (_I, [_P, _P, _I, _I, _I, _F, _F, _P, _P, _P, _P]),
    "gsr_extract_normals": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _P, _F, _F, _P, _P, _P, _P]),
    "gsr_normal_fusion_pass": (_I, [_L, _P, _P, _P, _I, _P, _F, _F, _F, _P, _F, _P, _P, _P, _P]),
    "gsr_normal_fusion_mean": (_I, [_I, _P, _P, _P, _P]),
    "gsr_knn_grid": (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "gsr_adam_step": (_I, [_I, _P, C.c_double, C.c_double, C.c_double, _L, _I, _F, _I, _P]),
    "gsr_set_tile_order": (_I, [_I]),
    "gsr_set_speculation": (_I, [_I]),
    "gsr_speculation_stats": (_I, [_P, _P]),
    "gsr_profile_enable": (_I, [_I]),
    "gsr_profile_read": (_I, [_P, _P]),
    "gsr_debug_export": (_I, [_I, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def ensure_built():
    """Build the library in-tree if the sources are newer (needs nvcc; a no-op on a box that got the .so)."""
    from . import build
    try:
        return build.build_library()
    except Exception:
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            ensure_built()
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                              "gaustudio_b200 has no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if l.gsr_abi_version() != 1:
            raise ImportError("libgsr_b200.so ABI version mismatch")
        _lib = l
    return _lib


def last_error():
    return lib().gsr_last_error().decode()
