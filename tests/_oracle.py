"""ctypes bindings for the parity checkers under oracle/ (test infrastructure).

``Oracle("port")`` wraps oracle/libfovoracle.so (the C restatement), ``Oracle("ref")``
wraps oracle/_ref/libfovref.so (the reference's own .cl sources compiled through the
g++ shim).  Both expose the same method names over numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i, _f = C.c_int, C.c_float


def reduced_size(dim: int) -> int:
    """16*ceil(dim/1.8/16), run_satlogrectilinear.cc:113-114."""
    import math

    return 16 * math.ceil(dim / 1.8 / 16)


def build_oracles(force: bool = False) -> None:
    args = [sys.executable, os.path.join(ORACLE_DIR, "build_oracle.py")]
    if force:
        args.append("--force")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)


def ref_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libfovref.so")) or os.path.isdir(
        os.environ.get("FOV_REF_DIR", "/root/reference/src")
    )


class Oracle:
    def __init__(self, kind: str = "port"):
        build_oracles()
        self.kind = kind
        if kind == "port":
            path, p = os.path.join(ORACLE_DIR, "libfovoracle.so"), "orc_"
        elif kind == "ref":
            path, p = os.path.join(ORACLE_DIR, "_ref", "libfovref.so"), "ref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        self.p = p

        def sig(name, *argtypes, res=None):
            fn = getattr(L, p + name)
            fn.argtypes = list(argtypes)
            fn.restype = res
            return fn

        self._set_threads = sig("set_threads", _i)
        self._get_threads = sig("get_threads", res=_i)
        self._sat_encode = sig("sat_encode", _u32p, _u8p, _i, _i, _i)
        if kind == "ref" and hasattr(L, "ref_sat_encode_cpu"):
            self._sat_encode_cpu = sig("sat_encode_cpu", _u32p, _u8p, _i, _i, _i)
        self._sat_create_grid = sig("sat_create_grid", _i16p, _i, _i, _i, _i)
        self._sat_sample_rect = sig("sat_sample_rect", _u8p, _i, _i, _i, _u32p, _i, _i, _i16p, _f, _f)
        self._sat_interpolate_rect = sig("sat_interpolate_rect", _u8p, _i, _i, _u8p, _i, _i, _f, _f)
        self._sat_decode = sig("sat_decode", _u8p, _i, _u32p, _i, _i)
        self._img_create_grid = sig("img_create_grid", _i16p, _i, _i, _i, _i)
        self._img_sample_rect = sig(
            "img_sample_rect", _u8p, _i, _i, _i, _u8p, _i, _i, _i, _i16p, _f, _f
        )
        self._img_create_logpolar_grid = sig("img_create_logpolar_grid", _i16p, _i, _i, _i, _i)
        self._img_sample_logpolar = sig(
            "img_sample_logpolar", _u8p, _i, _i, _i, _u8p, _i, _i, _i, _i16p, _f, _f
        )
        self._img_interpolate_logpolar = sig(
            "img_interpolate_logpolar", _u8p, _i, _i, _u8p, _i, _i, _f, _f
        )
        self._img_logpolar_blur = sig("img_logpolar_blur", _u8p, _i, _i, _i, _u8p)
        self._gnomonic = sig("gnomonic", _u8p, _i, _i, _u8p, _i, _i, _f, _f)
        if kind == "port":
            self._rgb0_to_yuv420p = sig("rgb0_to_yuv420p", _u8p, _i, _u8p, _i, _u8p, _i, _u8p, _i,
                                        _i, _i, res=_i)
            self._yuv420p_to_rgb0 = sig("yuv420p_to_rgb0", _u8p, _i, _u8p, _i, _u8p, _i, _u8p, _i,
                                        _i, _i, res=_i)
            self._sat_grid_edges = sig("sat_grid_edges", _i16p, _i16p, _i, _i, _i, _i)
            self._fnv = sig("fnv1a64", C.c_void_p, C.c_size_t, res=C.c_uint64)
            self._fill = sig("fill_frame_lcg", _u8p, C.c_size_t, C.c_uint32)

    # -- threads ------------------------------------------------------------
    def set_threads(self, n: int) -> None:
        self._set_threads(int(n))

    def get_threads(self) -> int:
        return int(self._get_threads())

    # -- SAT path -------------------------------------------------------------
    def sat_encode(self, frame: np.ndarray) -> np.ndarray:
        """frame: u8[H][W][4] (RGB0, linesize 4W) -> u32[H][W][3]."""
        H, W, bpp = frame.shape
        sat = np.empty((H, W, 3), np.uint32)
        self._sat_encode(sat, np.ascontiguousarray(frame), W, H, W * bpp)
        return sat

    def sat_encode_cpu(self, frame: np.ndarray, out=None) -> np.ndarray:
        """SATEncoder::EncodeFrameCPU (sat_encoder.cc:137-185): the reference's host twin ("ref" only)."""
        H, W, bpp = frame.shape
        sat = np.empty((H, W, 3), np.uint32) if out is None else out
        self._sat_encode_cpu(sat, np.ascontiguousarray(frame), W, H, W * bpp)
        return sat

    def sat_create_grid(self, ow: int, oh: int, W: int, H: int) -> np.ndarray:
        grid = np.zeros((oh + 1, ow + 1, 2), np.int16)
        self._sat_create_grid(grid, ow, oh, W, H)
        return grid

    def sat_sample_rect(self, sat, ow, oh, cx, cy, grid=None, out=None) -> np.ndarray:
        H, W, _ = sat.shape
        if grid is None:
            grid = self.sat_create_grid(ow, oh, W, H)
        if out is None:
            out = np.zeros((oh, ow, 4), np.uint8)
        self._sat_sample_rect(out, ow, oh, out.shape[1] * 4, sat, W, H, grid, cx, cy)
        return out

    def sat_interpolate_rect(self, reduced, W, H, cx, cy, out=None) -> np.ndarray:
        oh, ow, _ = reduced.shape
        if out is None:
            out = np.zeros((H, W, 4), np.uint8)
        self._sat_interpolate_rect(out, W, H, np.ascontiguousarray(reduced), ow, oh, cx, cy)
        return out

    def sat_decode(self, sat, bpp: int = 4, out=None) -> np.ndarray:
        H, W, _ = sat.shape
        if out is None:
            out = np.zeros((H, W, bpp), np.uint8)
        self._sat_decode(out, W * bpp, sat, W, H)
        return out

    # -- ImageSampler path ------------------------------------------------------
    def img_create_grid(self, ow, oh, W, H) -> np.ndarray:
        grid = np.zeros((oh, ow, 2), np.int16)
        self._img_create_grid(grid, ow, oh, W, H)
        return grid

    def img_sample_rect(self, frame, ow, oh, cx, cy, grid=None, out=None) -> np.ndarray:
        H, W, bpp = frame.shape
        if grid is None:
            grid = self.img_create_grid(ow, oh, W, H)
        if out is None:
            out = np.zeros((oh, ow, 4), np.uint8)
        self._img_sample_rect(
            out, ow, oh, out.shape[1] * out.shape[2], np.ascontiguousarray(frame), W, H, W * bpp,
            grid, cx, cy,
        )
        return out

    def img_create_logpolar_grid(self, ow, oh, W, H) -> np.ndarray:
        grid = np.zeros((oh, ow, 2), np.int16)
        self._img_create_logpolar_grid(grid, ow, oh, W, H)
        return grid

    def img_sample_logpolar(self, frame, ow, oh, cx, cy, grid=None, out=None) -> np.ndarray:
        H, W, bpp = frame.shape
        if grid is None:
            grid = self.img_create_logpolar_grid(ow, oh, W, H)
        if out is None:
            out = np.zeros((oh, ow, 4), np.uint8)
        self._img_sample_logpolar(
            out, ow, oh, out.shape[1] * out.shape[2], np.ascontiguousarray(frame), W, H, W * bpp,
            grid, cx, cy,
        )
        return out

    def img_interpolate_logpolar(self, reduced, W, H, cx, cy, out=None) -> np.ndarray:
        oh, ow, _ = reduced.shape
        if out is None:
            out = np.zeros((H, W, 4), np.uint8)
        self._img_interpolate_logpolar(out, W, H, np.ascontiguousarray(reduced), ow, oh, cx, cy)
        return out

    def gnomonic(self, frame, tw, th, cx, cy, out=None) -> np.ndarray:
        """Viewport tw x th out of a dense 4-byte-pixel equirect frame (projections.cc:51-86)."""
        H, W, _ = frame.shape
        if out is None:
            out = np.zeros((th, tw, 4), np.uint8)
        self._gnomonic(out, tw, th, np.ascontiguousarray(frame), W, H, cx, cy)
        return out

    def img_logpolar_blur(self, reduced, out=None) -> np.ndarray:
        oh, ow, _ = reduced.shape
        if out is None:
            out = np.zeros((oh, ow, 4), np.uint8)
        self._img_logpolar_blur(out, ow, oh, ow * 4, np.ascontiguousarray(reduced))
        return out

    # -- colour conversion (port only: the pin is libswscale itself, tests/golden/swscale_*) ----
    def rgb0_to_yuv420p(self, rgb0: np.ndarray):
        """u8[H][W][4] -> (Y u8[H][W], U u8[H/2][W/2], V u8[H/2][W/2]); video_encoder.cc:380-398."""
        H, W, _ = rgb0.shape
        y = np.zeros((H, W), np.uint8)
        u = np.zeros((H // 2, W // 2), np.uint8)
        v = np.zeros((H // 2, W // 2), np.uint8)
        rc = self._rgb0_to_yuv420p(y, W, u, W // 2, v, W // 2, np.ascontiguousarray(rgb0), W * 4,
                                   W, H)
        if rc != 0:
            raise ValueError("rgb0_to_yuv420p: unsupported size %dx%d" % (W, H))
        return y, u, v

    def yuv420p_to_rgb0(self, y: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
        """(Y u8[H][W], U, V u8[H/2][W/2]) -> RGB0 u8[H][W][4]; video_decoder.cc:165-222."""
        H, W = y.shape
        out = np.zeros((H, W, 4), np.uint8)
        rc = self._yuv420p_to_rgb0(out.reshape(H, W * 4), W * 4, np.ascontiguousarray(y), W,
                                   np.ascontiguousarray(u), W // 2, np.ascontiguousarray(v),
                                   W // 2, W, H)
        if rc != 0:
            raise ValueError("yuv420p_to_rgb0: unsupported size %dx%d" % (W, H))
        return out

    # -- helpers (port only) ------------------------------------------------------
    def sat_grid_edges(self, ow, oh, W, H):
        xe = np.zeros(ow + 1, np.int16)
        ye = np.zeros(oh + 1, np.int16)
        self._sat_grid_edges(xe, ye, ow, oh, W, H)
        return xe, ye


_PORT = None


def port() -> Oracle:
    global _PORT
    if _PORT is None:
        _PORT = Oracle("port")
    return _PORT


def fnv1a64(arr: np.ndarray) -> str:
    a = np.ascontiguousarray(arr)
    return "%016x" % port()._fnv(a.ctypes.data, a.nbytes)


def lcg_frame(W: int, H: int, seed: int = 12345) -> np.ndarray:
    """RGB0 u8[H][W][4]: LCG noise, padding byte 0 (SURVEY.md 8(c) generator)."""
    buf = np.empty(H * W * 4, np.uint8)
    port()._fill(buf, buf.size, seed & 0xFFFFFFFF)
    return buf.reshape(H, W, 4)


def lcg_planes(W: int, H: int, seed: int):
    """Y/U/V planes cut from one LCG frame (as tests/golden/make_golden_swscale.py does)."""
    f = lcg_frame(W, H, seed)
    return (np.ascontiguousarray(f[..., 0]), np.ascontiguousarray(f[0::2, 0::2, 1]),
            np.ascontiguousarray(f[0::2, 0::2, 2]))


def smooth_frame(W: int, H: int, seed: int = 7) -> np.ndarray:
    """Natural-image-like RGB0 frame: low-passed noise plus gradients (deterministic)."""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, size=(H // 16 + 2, W // 16 + 2, 3)).astype(np.float32)
    ys = np.linspace(0, small.shape[0] - 1.001, H)
    xs = np.linspace(0, small.shape[1] - 1.001, W)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    a = small[y0][:, x0] * (1 - fx) + small[y0][:, x0 + 1] * fx
    b = small[y0 + 1][:, x0] * (1 - fx) + small[y0 + 1][:, x0 + 1] * fx
    img = a * (1 - fy) + b * fy
    out = np.zeros((H, W, 4), np.uint8)
    out[..., :3] = np.clip(img, 0, 255).astype(np.uint8)
    return out
