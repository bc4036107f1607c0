#!/usr/bin/env python3
"""One small pass through EVERY kernel of libfov360.so, for `compute-sanitizer`:

    compute-sanitizer --tool memcheck  python tools/sanitize_pass.py
    compute-sanitizer --tool racecheck python tools/sanitize_pass.py

Small frames (the sanitizer slows kernels 10-100x), aligned and ragged geometries so that both the
single-pass and the three-kernel SAT builds, the vector and the generic blur, and the 3-byte-pixel
gathers run.  No torch import: the process holds nothing but numpy, ctypes and the library.
Results are compared with nothing here (the parity tests do that); a non-zero exit means a CUDA
error surfaced through the C ABI."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fov = importlib.import_module("foveated-360-video_b200")


def red(d):
    return fov.reduced_dim(d)


m = fov.OpenCLManager(0)
m.InitializeContext()
enc, dec, img = fov.SATEncoder(m), fov.SATDecoder(m), fov.ImageSampler(m)
proj, conv = fov.Projections(m), fov.VideoFrameConverter(m)
rng = np.random.default_rng(0)
kernels = 0
for (W, H, bpp) in [(256, 128, 4), (512, 192, 4), (250, 130, 4), (96, 64, 3)]:
    ow, oh = red(W), red(H)
    frame = rng.integers(0, 256, size=(H, W, bpp), dtype=np.uint8)
    src, sat = m.upload(frame), m.Buffer(12 * W * H)
    enc.EncodeFrameGPU(sat, src, W, H, W * bpp)
    for cx, cy in [(0.5, 0.5), (0.02, 0.97), (1.0, 0.0)]:
        r, full, back = m.Buffer(4 * ow * oh), m.Buffer(4 * W * H), m.Buffer(4 * W * H)
        m.memset(r, 0, 4 * ow * oh)
        dec.SampleFrameRectGPU(r, ow, oh, 4 * ow, sat, W, H, cx, cy)
        dec.InterpolateFrameRectGPU(full, W, H, 4 * W, r, ow, oh, 4 * ow, cx, cy)
        dec.DecodeFrameGPU(back, 4 * W, sat, W, H)
        proj.GnomonicProjection(back, 64, 32, 4 * 64, full, W, H, 4 * W, cx, cy)
        proj.InterpolateGnomonicGPU(back, 64, 32, r, ow, oh, W, H, cx, cy, 0.4, 0.6)
        if bpp == 4:
            lp, bl = m.Buffer(4 * ow * oh), m.Buffer(4 * ow * oh)
            m.memset(lp, 0, 4 * ow * oh)
            img.SampleFrameRectGPU(lp, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
            img.SampleFrameLogPolarGPU(lp, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
            img.ApplyLogPolarGaussianBlur(bl, ow, oh, 4 * ow, lp)
            img.InterpolateFrameLogPolarGPU(full, W, H, 4 * W, bl, ow, oh, 4 * ow, cx, cy)
            if W % 2 == 0 and H % 2 == 0:
                y, uv = m.Buffer(ow * oh), m.Buffer(ow * oh // 2)
                conv.RGB0ToNV12(y, ow, uv, ow, r, 4 * ow, ow, oh)
                conv.NV12ToRGB0(lp, 4 * ow, y, ow, uv, ow, ow, oh)
                u, v = m.Buffer(ow * oh // 4), m.Buffer(ow * oh // 4)
                conv.RGB0ToYUV420P(y, ow, u, ow // 2, v, ow // 2, r, 4 * ow, ow, oh)
                conv.YUV420PToRGB0(lp, 4 * ow, y, ow, u, ow // 2, v, ow // 2, ow, oh)
                for b in (y, uv, u, v):
                    b.free()
            for b in (lp, bl):
                b.free()
        for b in (r, full, back):
            b.free()
    m.Finish()
    src.free()
    sat.free()
# generic blur (width not a multiple of 4) and a batched foveation call
lp, bl = m.Buffer(4 * 50 * 20), m.Buffer(4 * 50 * 20)
m.memset(lp, 7, 4 * 50 * 20)
img.ApplyLogPolarGaussianBlur(bl, 50, 20, 200, lp)
W, H, B = 256, 128, 3
ow, oh = red(W), red(H)
frames = rng.integers(0, 256, size=(B, H, W, 4), dtype=np.uint8)
src, sat = m.upload(frames), m.Buffer(B * 12 * W * H)
r, full = m.Buffer(B * 4 * ow * oh), m.Buffer(B * 4 * W * H)
m.memset(r, 0, B * 4 * ow * oh)
gaze = rng.random((B, 2)).astype(np.float32)
for _ in range(2):
    fov.FoveateFramesGPU(m, B, full, 4 * W * H, r, 4 * ow * oh, sat, 12 * W * H, src, 4 * W * H, W, H,
                         4 * W, ow, oh, gaze)
m.Finish()
print("sanitize pass complete: %d kernel launches" % m.launch_count)
m.close()
