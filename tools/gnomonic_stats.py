#!/usr/bin/env python3
"""Timing of the viewport kernels: interpolate_rect + gnomonic against the fused kernel, and a check
that the two produce the same viewport.  (The comparison with the CPU restatement lives in
tests/test_gpu_parity.py::test_gnomonic_vs_oracle, which prints its mismatch counts with -s.)

    python tools/gnomonic_stats.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

fov = importlib.import_module("foveated-360-video_b200")
m = fov.OpenCLManager(0)
m.InitializeContext()
proj, enc, dec = fov.Projections(m), fov.SATEncoder(m), fov.SATDecoder(m)
for W, H, tw, th in [(3840, 1920, 1920, 1080), (7680, 3840, 1920, 1080)]:
    ow, oh = fov.reduced_dim(W), fov.reduced_dim(H)
    frame = bench.synth_frame(W, H, 1)
    src, sat = m.upload(frame), m.Buffer(W * H * 12)
    red = m.upload(np.zeros((oh, ow, 4), np.uint8))
    full, view, view2 = m.Buffer(W * H * 4), m.Buffer(tw * th * 4), m.Buffer(tw * th * 4)
    enc.EncodeFrameGPU(sat, src, W, H, 4 * W)
    for (gx, gy), (vx, vy) in [((0.5, 0.5), (0.5, 0.5)), ((0.3, 0.6), (0.33, 0.58)), ((0.9, 0.2), (0.05, 0.9))]:
        dec.SampleFrameRectGPU(red, ow, oh, 4 * ow, sat, W, H, gx, gy)
        dec.InterpolateFrameRectGPU(full, W, H, 4 * W, red, ow, oh, 4 * ow, gx, gy)
        proj.GnomonicProjection(view, tw, th, 4 * tw, full, W, H, 4 * W, vx, vy)
        proj.InterpolateGnomonicGPU(view2, tw, th, red, ow, oh, W, H, gx, gy, vx, vy)
        a = m.copy_to_host(np.empty((th, tw, 4), np.uint8), view)
        b = m.copy_to_host(np.empty((th, tw, 4), np.uint8), view2)
        assert np.array_equal(a, b)
    m.profile_reset()
    m.profile(True)
    for _ in range(20):
        dec.InterpolateFrameRectGPU(full, W, H, 4 * W, red, ow, oh, 4 * ow, 0.3, 0.6)
        proj.GnomonicProjection(view, tw, th, 4 * tw, full, W, H, 4 * W, 0.33, 0.58)
        proj.InterpolateGnomonicGPU(view2, tw, th, red, ow, oh, W, H, 0.3, 0.6, 0.33, 0.58)
    t = m.profile_totals()
    m.profile(False)
    print("%dx%d -> %dx%d viewport (fused == two kernels): " % (W, H, tw, th)
          + "  ".join("%s %.4f ms" % (k, v[0] / v[1]) for k, v in sorted(t.items())))
m.close()
