"""Out-of-bounds WRITE detection with guard bands (compute-sanitizer is closed on this GPU pool, so
SURVEY section 5's memcheck pass is replaced by checks of our own): every kernel writes into a
payload that sits between two 4 KiB guard bands filled with a pattern; after the call the bands must
be untouched and - for kernels that define every output byte - no pattern byte may remain inside.
Geometries cover the aligned fast paths and the ragged / 3-byte-pixel generic ones."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
GUARD, PAT = 4096, 0x5A


class Arena:
    def __init__(self, m, nbytes):
        self.m, self.n = m, int(nbytes)
        self.buf = m.upload(np.full(self.n + 2 * GUARD, PAT, np.uint8))

    @property
    def ptr(self):
        return self.buf.at(GUARD)

    def check(self, what, fully_written=False):
        host = self.m.copy_to_host(np.empty(self.n + 2 * GUARD, np.uint8), self.buf)
        assert (host[:GUARD] == PAT).all(), what + ": wrote below its buffer"
        assert (host[GUARD + self.n:] == PAT).all(), what + ": wrote past its buffer"
        payload = host[GUARD:GUARD + self.n]
        if fully_written:
            # a random frame may hold the pattern byte here and there; long runs mean unwritten rows
            run = np.flatnonzero(np.diff(np.concatenate(([0], (payload == PAT).view(np.int8), [0]))))
            longest = int((run[1::2] - run[0::2]).max()) if len(run) else 0
            assert longest < 64, what + ": %d consecutive bytes were never written" % longest
        self.buf.free()
        return payload


@pytest.mark.parametrize("W,H,bpp", [(256, 128, 4), (512, 192, 4), (250, 130, 4), (96, 64, 3),
                                     (1000, 36, 4)])
def test_no_kernel_writes_outside_its_buffers(fov, mgr, W, H, bpp):
    m = mgr
    enc, dec, img = fov.SATEncoder(m), fov.SATDecoder(m), fov.ImageSampler(m)
    proj, conv = fov.Projections(m), fov.VideoFrameConverter(m)
    ow, oh = fov.reduced_dim(W), fov.reduced_dim(H)
    rng = np.random.default_rng(W * 7 + H)
    frame = rng.integers(0, 256, size=(H, W, bpp), dtype=np.uint8)
    src = m.upload(frame)
    sat = Arena(m, 12 * W * H)
    enc.EncodeFrameGPU(sat.ptr, src, W, H, W * bpp)
    for cx, cy in [(0.5, 0.5), (0.02, 0.97), (1.0, 0.0)]:
        red, full, back = Arena(m, 4 * ow * oh), Arena(m, 4 * W * H), Arena(m, 4 * W * H)
        view, view2 = Arena(m, 4 * 64 * 32), Arena(m, 4 * 64 * 32)
        dec.SampleFrameRectGPU(red.ptr, ow, oh, 4 * ow, sat.ptr, W, H, cx, cy)
        dec.InterpolateFrameRectGPU(full.ptr, W, H, 4 * W, red.ptr, ow, oh, 4 * ow, cx, cy)
        dec.DecodeFrameGPU(back.ptr, 4 * W, sat.ptr, W, H)
        proj.GnomonicProjection(view.ptr, 64, 32, 4 * 64, full.ptr, W, H, 4 * W, cx, cy)
        proj.InterpolateGnomonicGPU(view2.ptr, 64, 32, red.ptr, ow, oh, W, H, cx, cy, 0.4, 0.6)
        if bpp == 4:
            lp, bl, lpfull = Arena(m, 4 * ow * oh), Arena(m, 4 * ow * oh), Arena(m, 4 * W * H)
            rect = Arena(m, 4 * ow * oh)
            img.SampleFrameRectGPU(rect.ptr, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
            img.SampleFrameLogPolarGPU(lp.ptr, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
            img.ApplyLogPolarGaussianBlur(bl.ptr, ow, oh, 4 * ow, lp.ptr)
            img.InterpolateFrameLogPolarGPU(lpfull.ptr, W, H, 4 * W, bl.ptr, ow, oh, 4 * ow, cx, cy)
            y, uv, rgb = Arena(m, ow * oh), Arena(m, ow * oh // 2), Arena(m, 4 * ow * oh)
            conv.RGB0ToNV12(y.ptr, ow, uv.ptr, ow, red.ptr, 4 * ow, ow, oh)
            conv.NV12ToRGB0(rgb.ptr, 4 * ow, y.ptr, ow, uv.ptr, ow, ow, oh)
            u, v, y2 = Arena(m, ow * oh // 4), Arena(m, ow * oh // 4), Arena(m, ow * oh)
            conv.RGB0ToYUV420P(y2.ptr, ow, u.ptr, ow // 2, v.ptr, ow // 2, red.ptr, 4 * ow, ow, oh)
            rect.check("img sample_rect")
            lp.check("sample_logpolar")
            bl.check("logpolar blur", fully_written=True)
            lpfull.check("interpolate_logpolar", fully_written=True)
            y.check("rgb0_to_nv12 luma", fully_written=True)
            uv.check("rgb0_to_nv12 chroma", fully_written=True)
            rgb.check("nv12_to_rgb0", fully_written=True)
            for a, name in ((u, "yuv420p U"), (v, "yuv420p V"), (y2, "yuv420p Y")):
                a.check(name, fully_written=True)
        red.check("sample_rect")
        full.check("interpolate_rect", fully_written=True)
        back.check("decode")
        view.check("gnomonic", fully_written=True)
        view2.check("interpolate_gnomonic", fully_written=True)
    got = sat.check("SAT encode", fully_written=False).view(np.uint32).reshape(H, W, 3)
    assert np.array_equal(got, O.port().sat_encode(frame))
    src.free()


def test_batched_pipeline_stays_inside_its_buffers(fov, mgr):
    m = mgr
    W, H, B = 256, 128, 3
    ow, oh = fov.reduced_dim(W), fov.reduced_dim(H)
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, size=(B, H, W, 4), dtype=np.uint8)
    src = m.upload(frames)
    sat, red, full = Arena(m, B * 12 * W * H), Arena(m, B * 4 * ow * oh), Arena(m, B * 4 * W * H)
    gaze = rng.random((B, 2)).astype(np.float32)
    for _ in range(2):
        fov.FoveateFramesGPU(m, B, full.ptr, 4 * W * H, red.ptr, 4 * ow * oh, sat.ptr, 12 * W * H, src,
                             4 * W * H, W, H, 4 * W, ow, oh, gaze)
    sat.check("batched SAT")
    red.check("batched sample_rect")
    full.check("batched interpolate_rect", fully_written=True)
    src.free()
