"""ctypes binding of libfov360.so - one Python prototype per entry point of include/fov360.h.

This is plumbing for the tests and the benchmark: the product is the CUDA library and its C ABI.
Loading fails loudly when the library is missing and cannot be built; nothing in this package
falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Kept in the order of include/fov360.h.
PROTOTYPES = {
    "fov_ctx_create": (_vp, [_i, C.POINTER(_i)]),
    "fov_ctx_destroy": (None, [_vp]),
    "fov_last_error_string": (C.c_char_p, [_vp]),
    "fov_device_count": (_i, []),
    "fov_ctx_device": (_i, [_vp]),
    "fov_ctx_stream": (_vp, [_vp]),
    "fov_sync": (_i, [_vp]),
    "fov_ctx_launch_count": (C.c_uint64, [_vp]),
    "fov_ctx_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "fov_ctx_get_option": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "fov_profile_enable": (_i, [_vp, _i]),
    "fov_profile_reset": (_i, [_vp]),
    "fov_profile_count": (_i, [_vp]),
    "fov_profile_get": (_i, [_vp, _i, C.c_char_p, _sz, C.POINTER(C.c_double),
                             C.POINTER(C.c_uint64)]),
    "fov_malloc": (_i, [_vp, C.POINTER(_vp), _sz]),
    "fov_free": (_i, [_vp, _vp]),
    "fov_memset": (_i, [_vp, _vp, _i, _sz]),
    "fov_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "fov_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "fov_memcpy_h2d_async": (_i, [_vp, _vp, _vp, _sz]),
    "fov_memcpy_d2h_async": (_i, [_vp, _vp, _vp, _sz]),
    "fov_host_alloc": (_i, [_vp, C.POINTER(_vp), _sz]),
    "fov_host_free": (_i, [_vp, _vp]),
    "fov_sat_encode": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "fov_sat_encode_batched": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _i, _i, _i]),
    "fov_sat_grid_init": (_i, [_vp, _i, _i, _i, _i]),
    "fov_sat_grid_export": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "fov_sat_sample_rect": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _f, _f]),
    "fov_sat_sample_rect_batched": (_i, [_vp, _i, _vp, _sz, _i, _i, _i, _vp, _sz, _i, _i, _fp]),
    "fov_sat_interpolate_rect": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _f]),
    "fov_sat_interpolate_rect_batched": (_i, [_vp, _i, _vp, _sz, _i, _i, _vp, _sz, _i, _i, _fp]),
    "fov_sat_decode": (_i, [_vp, _vp, _i, _vp, _i, _i]),
    "fov_sat_foveate_batched": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz,
                                     _i, _i, _i, _i, _i, _fp]),
    "fov_sat_encode_sample_batched": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _vp, _sz,
                                           _i, _i, _i, _i, _i, _fp]),
    "fov_sat_encode_sample_batched_dev": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _vp, _sz,
                                               _i, _i, _i, _i, _i, _vp]),
    "fov_sat_foveate_batched_dev": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz,
                                         _i, _i, _i, _i, _i, _vp]),
    "fov_graph_begin_capture": (_i, [_vp]),
    "fov_graph_end_capture": (_i, [_vp, C.POINTER(_vp)]),
    "fov_graph_launch": (_i, [_vp, _vp]),
    "fov_graph_destroy": (None, [_vp, _vp]),
    "fov_img_grid_init": (_i, [_vp, _i, _i, _i, _i]),
    "fov_img_grid_export": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "fov_img_sample_rect": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _f]),
    "fov_img_logpolar_grid_init": (_i, [_vp, _i, _i, _i, _i]),
    "fov_img_logpolar_grid_export": (_i, [_vp, _vp, _i, _i]),
    "fov_img_sample_logpolar": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _f]),
    "fov_img_interpolate_logpolar": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _f]),
    "fov_img_logpolar_blur": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "fov_gnomonic": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _f]),
    "fov_sat_interpolate_gnomonic": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _f, _f, _f, _f]),
    "fov_rgb0_to_yuv420p": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i]),
    "fov_rgb0_to_nv12": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i]),
    "fov_rgb0_to_yuv420p_batched": (_i, [_vp, _i, _vp, _sz, _i, _vp, _vp, _sz, _i, _vp, _sz, _i,
                                         _i, _i]),
    "fov_rgb0_to_nv12_batched": (_i, [_vp, _i, _vp, _sz, _i, _vp, _sz, _i, _vp, _sz, _i, _i, _i]),
    "fov_yuv420p_to_rgb0": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i]),
    "fov_nv12_to_rgb0": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i]),
    "fov_yuv420p_to_rgb0_batched": (_i, [_vp, _i, _vp, _sz, _i, _vp, _sz, _i, _vp, _vp, _sz, _i,
                                         _i, _i]),
    "fov_nv12_to_rgb0_batched": (_i, [_vp, _i, _vp, _sz, _i, _vp, _sz, _i, _vp, _sz, _i, _i, _i]),
    "fov_debug_bounds_violations": (_i, [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "fov_reduced_dim": (_i, [_i]),
}

_LIB = None


def header_symbols() -> list[str]:
    """Every function name declared in include/fov360.h (used by the CPU-side ABI test)."""
    with open(os.path.join(_build.INCLUDE, "fov360.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fov_[a-z0-9_]+)\s*\(", text)))


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_stale: bool = True) -> C.CDLL:
    """Loads libfov360.so, (re)building it in-tree first when nvcc is present."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    override = os.environ.get("FOV360_LIB")  # development only: A/B runs of kernel variants
    if override:
        path, build_if_stale = override, False
    if build_if_stale:
        try:
            path = _build.build()
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(
            "libfov360.so is missing (%s): build it with `python __graft_entry__.py build`; "
            "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
