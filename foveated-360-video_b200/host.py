"""Host-side mirror of the reference's foveation classes over the C ABI (tests / bench harness).

Class and method names follow the reference (src/opencl_manager.h:8-22, src/sat_encoder.h:21-43,
src/sat_decoder.h:20-83, src/image_sampler.h:29-102); argument order and meaning are the
reference's: sizes in pixels, linesizes in bytes, gaze as two floats in [0,1].  ``cl_mem``
handles become :class:`DeviceBuffer` objects (a device pointer + size).  The shipped drop-in for
the reference's C++ call sites is include/fov360/*.h; this module exists so that the parity tests
read like the reference's call sequences (video_server.cc:296-345, run_satlogrectilinear.cc:915-949).

Every method ends in a launch on the manager's stream (the in-order command queue); results are
visible after ``Finish()`` or a blocking ``copy_to_host``.  Errors raise :class:`FovError` - the
reference prints to stderr and returns; a Python caller is better served by an exception.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


class FovError(RuntimeError):
    pass


class DeviceBuffer:
    """cl::Buffer(context, CL_MEM_READ_WRITE, nbytes) (video_server.cc:224-232)."""

    def __init__(self, manager: "OpenCLManager", nbytes: int):
        self.manager = manager
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        manager._check(manager.lib.fov_malloc(manager.ctx, C.byref(p), self.nbytes))
        self.ptr = p.value

    def __call__(self):  # cl::Buffer::operator() -> cl_mem
        return self.ptr

    def at(self, byte_offset: int) -> int:
        return self.ptr + int(byte_offset)

    def free(self) -> None:
        if self.ptr:
            self.manager.lib.fov_free(self.manager.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.manager.ctx:
                self.free()
        except Exception:
            pass


def _ptr(buf) -> int:
    if isinstance(buf, DeviceBuffer):
        return buf.ptr
    return int(buf)


class OpenCLManager:
    """Device / stream holder; replaces OpenCLManager::InitializeContext (opencl_manager.cc:7-67)."""

    def __init__(self, device: int = 0):
        self.lib = _capi.load()
        self.ctx = None
        self.device = device

    def InitializeContext(self) -> None:
        err = C.c_int(0)
        ctx = self.lib.fov_ctx_create(self.device, C.byref(err))
        if not ctx:
            raise FovError("fov_ctx_create(%d) failed (%d): %s" % (
                self.device, err.value, self.lib.fov_last_error_string(None).decode()))
        self.ctx = ctx

    # -- helpers ------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.fov_last_error_string(self.ctx if rc != -1 else None)
            raise FovError("fov360 error %d: %s" % (rc, (msg or b"").decode()))

    @property
    def stream(self) -> int:
        return self.lib.fov_ctx_stream(self.ctx) or 0

    @property
    def launch_count(self) -> int:
        return int(self.lib.fov_ctx_launch_count(self.ctx))

    OPT_REDUCED_PAD_ZERO = 1

    def set_option(self, option: int, value: bool) -> None:
        """fov_ctx_set_option: promises the reference interface cannot state (include/fov360.h)."""
        self._check(self.lib.fov_ctx_set_option(self.ctx, int(option), int(bool(value))))

    def get_option(self, option: int) -> bool:
        v = C.c_int(0)
        self._check(self.lib.fov_ctx_get_option(self.ctx, int(option), C.byref(v)))
        return bool(v.value)

    def profile(self, on: bool) -> None:
        self._check(self.lib.fov_profile_enable(self.ctx, int(on)))

    def profile_reset(self) -> None:
        self._check(self.lib.fov_profile_reset(self.ctx))

    def profile_totals(self) -> dict:
        """{kernel name: (total ms, launches)} since the last reset (synchronises the stream)."""
        out = {}
        name = C.create_string_buffer(64)
        ms, cnt = C.c_double(0), C.c_uint64(0)
        for i in range(self.lib.fov_profile_count(self.ctx)):
            self._check(self.lib.fov_profile_get(self.ctx, i, name, 64, C.byref(ms), C.byref(cnt)))
            out[name.value.decode()] = (ms.value, int(cnt.value))
        return out

    def Buffer(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def copy_to_device(self, dst, host: np.ndarray, dst_offset: int = 0) -> None:
        """cl::copy(queue, begin, end, buffer): blocking (video_server.cc:297-299)."""
        host = np.ascontiguousarray(host)
        self._check(self.lib.fov_memcpy_h2d(self.ctx, _ptr(dst) + dst_offset, host.ctypes.data,
                                            host.nbytes))

    def copy_to_host(self, host: np.ndarray, src, src_offset: int = 0) -> np.ndarray:
        """cl::copy(queue, buffer, begin, end): blocking (video_server.cc:342-345)."""
        assert host.flags["C_CONTIGUOUS"]
        self._check(self.lib.fov_memcpy_d2h(self.ctx, host.ctypes.data, _ptr(src) + src_offset,
                                            host.nbytes))
        return host

    def upload(self, host: np.ndarray) -> DeviceBuffer:
        buf = self.Buffer(host.nbytes)
        self.copy_to_device(buf, host)
        return buf

    def memset(self, dst, byte: int, nbytes: int, offset: int = 0) -> None:
        self._check(self.lib.fov_memset(self.ctx, _ptr(dst) + offset, byte, nbytes))

    def Finish(self) -> None:
        """clFlush + clFinish (video_server.cc:302-303)."""
        self._check(self.lib.fov_sync(self.ctx))

    # -- CUDA-graph capture of a call sequence (no reference counterpart) ------------------
    def BeginCapture(self) -> None:
        self._check(self.lib.fov_graph_begin_capture(self.ctx))

    def EndCapture(self) -> int:
        g = C.c_void_p()
        self._check(self.lib.fov_graph_end_capture(self.ctx, C.byref(g)))
        return g.value

    def LaunchGraph(self, graph: int) -> None:
        self._check(self.lib.fov_graph_launch(self.ctx, graph))

    def DestroyGraph(self, graph: int) -> None:
        self.lib.fov_graph_destroy(self.ctx, graph)

    def copy_to_device_async(self, dst, host: np.ndarray, dst_offset: int = 0) -> None:
        """Stream-ordered host -> device copy; `host` must stay alive (and unchanged) until it ran."""
        self._check(self.lib.fov_memcpy_h2d_async(self.ctx, _ptr(dst) + dst_offset, host.ctypes.data,
                                                  host.nbytes))

    def close(self) -> None:
        if self.ctx:
            self.lib.fov_ctx_destroy(self.ctx)
            self.ctx = None


def _gaze_array(gaze) -> np.ndarray:
    g = np.ascontiguousarray(np.asarray(gaze, dtype=np.float32).reshape(-1, 2))
    return g


class SATEncoder:
    """sat_encoder.h:21-43."""

    def __init__(self, cl_manager: OpenCLManager | None = None):
        self.m = cl_manager

    def _need(self):
        if self.m is None or not self.m.ctx:
            raise FovError("Not initialized with OpenCL")  # sat_encoder.cc:70-74

    def EncodeFrameGPU(self, target_buffer, source_buffer, width, height, source_linesize):
        self._need()
        self.m._check(self.m.lib.fov_sat_encode(self.m.ctx, _ptr(target_buffer),
                                                _ptr(source_buffer), width, height,
                                                source_linesize))

    def EncodeFramesGPU(self, n, target_buffer, target_stride, source_buffer, source_stride, width,
                        height, source_linesize):
        """Batched form: n independent frames, strides in bytes (no reference counterpart)."""
        self._need()
        self.m._check(self.m.lib.fov_sat_encode_batched(
            self.m.ctx, n, _ptr(target_buffer), target_stride, _ptr(source_buffer), source_stride,
            width, height, source_linesize))


class SATDecoder:
    """sat_decoder.h:20-83 (log-rectilinear sampler / inverse warp / exact decode)."""

    def __init__(self, cl_manager: OpenCLManager | None = None):
        self.m = cl_manager

    def _need(self):
        if self.m is None or not self.m.ctx:
            raise FovError("Not initialized with OpenCL")  # sat_decoder.cc:179-183

    def InitializeGrid(self, target_width, target_height, source_width, source_height):
        self._need()
        self.m._check(self.m.lib.fov_sat_grid_init(self.m.ctx, target_width, target_height,
                                                   source_width, source_height))

    def ExportGrid(self, target_width, target_height, source_width, source_height) -> np.ndarray:
        """The reference's int16 [(oh+1)][(ow+1)][2] grid buffer, for parity checks."""
        self._need()
        g = np.zeros((target_height + 1, target_width + 1, 2), np.int16)
        self.m._check(self.m.lib.fov_sat_grid_export(self.m.ctx, g.ctypes.data, target_width,
                                                     target_height, source_width, source_height))
        return g

    def SampleFrameRectGPU(self, target_buffer, target_width, target_height, target_linesize,
                           source_buffer, source_width, source_height, center_x, center_y):
        """`source_width/height` stand in for the AVCodecContext* (sat_decoder.cc:328-329)."""
        self._need()
        self.m._check(self.m.lib.fov_sat_sample_rect(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, center_x, center_y))

    def SampleFramesRectGPU(self, n, target_buffer, target_stride, target_width, target_height,
                            target_linesize, source_buffer, source_stride, source_width,
                            source_height, gaze):
        self._need()
        g = _gaze_array(gaze)
        assert g.shape[0] == n
        self.m._check(self.m.lib.fov_sat_sample_rect_batched(
            self.m.ctx, n, _ptr(target_buffer), target_stride, target_width, target_height,
            target_linesize, _ptr(source_buffer), source_stride, source_width, source_height,
            g.ctypes.data_as(C.POINTER(C.c_float))))

    def InterpolateFrameRectGPU(self, target_buffer, target_width, target_height, target_linesize,
                                source_buffer, source_width, source_height, source_linesize,
                                center_x, center_y):
        self._need()
        self.m._check(self.m.lib.fov_sat_interpolate_rect(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, source_linesize, center_x, center_y))

    def InterpolateFramesRectGPU(self, n, target_buffer, target_stride, target_width, target_height,
                                 source_buffer, source_stride, source_width, source_height, gaze):
        self._need()
        g = _gaze_array(gaze)
        assert g.shape[0] == n
        self.m._check(self.m.lib.fov_sat_interpolate_rect_batched(
            self.m.ctx, n, _ptr(target_buffer), target_stride, target_width, target_height,
            _ptr(source_buffer), source_stride, source_width, source_height,
            g.ctypes.data_as(C.POINTER(C.c_float))))

    def DecodeFrameGPU(self, target_buffer, target_linesize, source_buffer, width, height):
        self._need()
        self.m._check(self.m.lib.fov_sat_decode(self.m.ctx, _ptr(target_buffer), target_linesize,
                                                _ptr(source_buffer), width, height))


def FoveateFramesGPU(m: OpenCLManager, n, full_out, full_stride, reduced, reduced_stride, sat,
                     sat_stride, source, source_stride, width, height, source_linesize, ow, oh,
                     gaze):
    """run_satlogrectilinear.cc:926-943 (encode -> sample -> interpolate) for n frames."""
    g = _gaze_array(gaze)
    assert g.shape[0] == n
    m._check(m.lib.fov_sat_foveate_batched(
        m.ctx, n, _ptr(full_out), full_stride, _ptr(reduced), reduced_stride, _ptr(sat),
        sat_stride, _ptr(source), source_stride, width, height, source_linesize, ow, oh,
        g.ctypes.data_as(C.POINTER(C.c_float))))


def FoveateFramesDeviceGazeGPU(m: OpenCLManager, n, full_out, full_stride, reduced, reduced_stride,
                               sat, sat_stride, source, source_stride, width, height,
                               source_linesize, ow, oh, gaze_dev):
    """FoveateFramesGPU with the 2n gaze floats in a device buffer: capturable as a CUDA graph."""
    m._check(m.lib.fov_sat_foveate_batched_dev(
        m.ctx, n, _ptr(full_out), full_stride, _ptr(reduced), reduced_stride, _ptr(sat),
        sat_stride, _ptr(source), source_stride, width, height, source_linesize, ow, oh,
        _ptr(gaze_dev)))


def EncodeSampleFramesGPU(m: OpenCLManager, n, reduced, reduced_stride, sat, sat_stride, source,
                          source_stride, width, height, source_linesize, ow, oh, gaze):
    """video_server.cc:300-338 (encode -> sample; the client un-warps) for n frames / streams."""
    g = _gaze_array(gaze)
    assert g.shape[0] == n
    m._check(m.lib.fov_sat_encode_sample_batched(
        m.ctx, n, _ptr(reduced), reduced_stride, _ptr(sat), sat_stride, _ptr(source),
        source_stride, width, height, source_linesize, ow, oh,
        g.ctypes.data_as(C.POINTER(C.c_float))))


class ImageSampler:
    """image_sampler.h:29-102 (no-SAT baseline: log-rect point sampling and log-polar)."""

    def __init__(self, cl_manager: OpenCLManager | None = None):
        self.m = cl_manager

    def _need(self):
        if self.m is None or not self.m.ctx:
            raise FovError("Not initialized with OpenCL")

    def InitializeGrid(self, target_width, target_height, source_width, source_height):
        self._need()
        self.m._check(self.m.lib.fov_img_grid_init(self.m.ctx, target_width, target_height,
                                                   source_width, source_height))

    def ExportGrid(self, target_width, target_height, source_width, source_height) -> np.ndarray:
        self._need()
        g = np.zeros((target_height, target_width, 2), np.int16)
        self.m._check(self.m.lib.fov_img_grid_export(self.m.ctx, g.ctypes.data, target_width,
                                                     target_height, source_width, source_height))
        return g

    def InitializeLogpolarGrid(self, target_width, target_height, source_width, source_height):
        self._need()
        self.m._check(self.m.lib.fov_img_logpolar_grid_init(
            self.m.ctx, target_width, target_height, source_width, source_height))

    def ExportLogpolarGrid(self, target_width, target_height) -> np.ndarray:
        self._need()
        g = np.zeros((target_height, target_width, 2), np.int16)
        self.m._check(self.m.lib.fov_img_logpolar_grid_export(self.m.ctx, g.ctypes.data,
                                                              target_width, target_height))
        return g

    def SampleFrameRectGPU(self, target_buffer, target_width, target_height, target_linesize,
                           source_buffer, source_width, source_height, source_linesize, center_x,
                           center_y):
        self._need()
        self.m._check(self.m.lib.fov_img_sample_rect(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, source_linesize, center_x, center_y))

    def SampleFrameLogPolarGPU(self, target_buffer, target_width, target_height, target_linesize,
                               source_buffer, source_width, source_height, source_linesize,
                               center_x, center_y):
        self._need()
        self.m._check(self.m.lib.fov_img_sample_logpolar(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, source_linesize, center_x, center_y))

    def InterpolateFrameLogPolarGPU(self, target_buffer, target_width, target_height,
                                    target_linesize, source_buffer, source_width, source_height,
                                    source_linesize, center_x, center_y):
        self._need()
        self.m._check(self.m.lib.fov_img_interpolate_logpolar(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, source_linesize, center_x, center_y))

    def ApplyLogPolarGaussianBlur(self, target_buffer, width, height, linesize, source_buffer):
        self._need()
        self.m._check(self.m.lib.fov_img_logpolar_blur(self.m.ctx, _ptr(target_buffer), width,
                                                       height, linesize, _ptr(source_buffer)))


class Projections:
    """projections.h:20-35: viewport rendering (inverse gnomonic projection)."""

    def __init__(self, cl_manager: OpenCLManager | None = None):
        self.m = cl_manager

    def _need(self):
        if self.m is None or not self.m.ctx:
            raise FovError("Not initialized with OpenCL")

    def GnomonicProjection(self, target_buffer, target_width, target_height, target_linesize,
                           source_buffer, source_width, source_height, source_linesize, center_x,
                           center_y):
        self._need()
        self.m._check(self.m.lib.fov_gnomonic(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, target_linesize,
            _ptr(source_buffer), source_width, source_height, source_linesize, center_x, center_y))

    def InterpolateGnomonicGPU(self, target_buffer, target_width, target_height, reduced_buffer,
                               reduced_width, reduced_height, full_width, full_height, gaze_x,
                               gaze_y, view_x, view_y):
        """interpolate_rect + gnomonic in one kernel (no reference counterpart)."""
        self._need()
        self.m._check(self.m.lib.fov_sat_interpolate_gnomonic(
            self.m.ctx, _ptr(target_buffer), target_width, target_height, _ptr(reduced_buffer),
            reduced_width, reduced_height, full_width, full_height, gaze_x, gaze_y, view_x,
            view_y))


class VideoFrameConverter:
    """The swscale steps either side of the foveation path, on the device: RGB0 -> YUV420P / NV12
    of VideoEncoder::EncodeFrame (video_encoder.cc:380-398) and YUV420P / NV12 -> RGB0 of
    VideoDecoder::GetFrame (video_decoder.cc:165-170, :222), with libswscale's arithmetic."""

    def __init__(self, cl_manager: OpenCLManager | None = None):
        self.m = cl_manager

    def _need(self):
        if self.m is None or not self.m.ctx:
            raise FovError("Not initialized with OpenCL")

    def RGB0ToYUV420P(self, y, y_linesize, u, u_linesize, v, v_linesize, source_buffer,
                      source_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_rgb0_to_yuv420p(
            self.m.ctx, _ptr(y), y_linesize, _ptr(u), u_linesize, _ptr(v), v_linesize,
            _ptr(source_buffer), source_linesize, width, height))

    def RGB0ToNV12(self, y, y_linesize, uv, uv_linesize, source_buffer, source_linesize, width,
                   height):
        self._need()
        self.m._check(self.m.lib.fov_rgb0_to_nv12(
            self.m.ctx, _ptr(y), y_linesize, _ptr(uv), uv_linesize, _ptr(source_buffer),
            source_linesize, width, height))

    # -- decoder side: video_decoder.cc:165-170, :222 ----------------------------------------
    def YUV420PToRGB0(self, target_buffer, target_linesize, y, y_linesize, u, u_linesize, v,
                      v_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_yuv420p_to_rgb0(
            self.m.ctx, _ptr(target_buffer), target_linesize, _ptr(y), y_linesize, _ptr(u),
            u_linesize, _ptr(v), v_linesize, width, height))

    def NV12ToRGB0(self, target_buffer, target_linesize, y, y_linesize, uv, uv_linesize, width,
                   height):
        self._need()
        self.m._check(self.m.lib.fov_nv12_to_rgb0(
            self.m.ctx, _ptr(target_buffer), target_linesize, _ptr(y), y_linesize, _ptr(uv),
            uv_linesize, width, height))

    def YUV420PToRGB0Frames(self, n, target_buffer, target_stride, target_linesize, y, y_stride,
                            y_linesize, u, v, chroma_stride, chroma_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_yuv420p_to_rgb0_batched(
            self.m.ctx, n, _ptr(target_buffer), target_stride, target_linesize, _ptr(y), y_stride,
            y_linesize, _ptr(u), _ptr(v), chroma_stride, chroma_linesize, width, height))

    def NV12ToRGB0Frames(self, n, target_buffer, target_stride, target_linesize, y, y_stride,
                         y_linesize, uv, uv_stride, uv_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_nv12_to_rgb0_batched(
            self.m.ctx, n, _ptr(target_buffer), target_stride, target_linesize, _ptr(y), y_stride,
            y_linesize, _ptr(uv), uv_stride, uv_linesize, width, height))

    def RGB0ToYUV420PFrames(self, n, y, y_stride, y_linesize, u, v, chroma_stride, chroma_linesize,
                            source_buffer, source_stride, source_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_rgb0_to_yuv420p_batched(
            self.m.ctx, n, _ptr(y), y_stride, y_linesize, _ptr(u), _ptr(v), chroma_stride,
            chroma_linesize, _ptr(source_buffer), source_stride, source_linesize, width, height))

    def RGB0ToNV12Frames(self, n, y, y_stride, y_linesize, uv, uv_stride, uv_linesize,
                         source_buffer, source_stride, source_linesize, width, height):
        self._need()
        self.m._check(self.m.lib.fov_rgb0_to_nv12_batched(
            self.m.ctx, n, _ptr(y), y_stride, y_linesize, _ptr(uv), uv_stride, uv_linesize,
            _ptr(source_buffer), source_stride, source_linesize, width, height))


class GazeViewPoint:
    """gaze_view_points.h:11-17."""
    __slots__ = ("frame", "view_point", "gaze_point", "pred_view_point", "pred_gaze_point")

    def __init__(self, frame, view_point, gaze_point):
        self.frame = frame
        self.view_point = view_point
        self.gaze_point = gaze_point
        self.pred_view_point = view_point
        self.pred_gaze_point = gaze_point


class GazeViewPoints:
    """gaze_view_points.cc:3-37: one record per line containing
    ``frame,<n>,forward,<x>,<y>,eye,<x>,<y>``; pred_* = the previous record's measured points."""

    _FLOAT = r"([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)"

    def __init__(self, file_path: str | None = None):
        import re

        self.points: list[GazeViewPoint] = []
        if file_path is None:
            return
        rx = re.compile(r"frame,(\d+),forward," + self._FLOAT + "," + self._FLOAT + ",eye," +
                        self._FLOAT + "," + self._FLOAT, re.ASCII)
        try:
            fh = open(file_path, "r", errors="replace")
        except OSError:
            import sys

            print("Cannot open file: " + file_path, file=sys.stderr)  # gaze_view_points.cc:35
            return
        with fh:
            for line in fh:
                m = rx.search(line)
                if not m:
                    continue
                f = [float(np.float32(m.group(k))) for k in range(2, 6)]
                p = GazeViewPoint(int(m.group(1)), (f[0], f[1]), (f[2], f[3]))
                if self.points:
                    p.pred_view_point = self.points[-1].view_point
                    p.pred_gaze_point = self.points[-1].gaze_point
                self.points.append(p)

    def gaze_array(self) -> np.ndarray:
        """float32 [n][2] of gaze_point, the (center_x, center_y) of frame f
        (run_satlogrectilinear.cc:519-521)."""
        return np.asarray([p.gaze_point for p in self.points], np.float32).reshape(-1, 2)


def reduced_dim(full_dim: int) -> int:
    """16*ceil(dim/1.8/16), run_satlogrectilinear.cc:113-114."""
    return int(_capi.load().fov_reduced_dim(int(full_dim)))
