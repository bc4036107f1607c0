#!/usr/bin/env python3
"""Parity statistics of the ImageSampler log-polar path against the oracle (per-pixel difference
histogram of interpolate_logpolar, exactness of sample_logpolar / blur) and per-kernel times.

    python tools/logpolar_stats.py [--sizes 1080p,4k] [--time 4k,8k] [--ref]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import _oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1080p,4k")
ap.add_argument("--time", default="4k,8k")
ap.add_argument("--ref", action="store_true", help="compare with oracle/_ref (the reference's .cl)")
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
fov = importlib.import_module("foveated-360-video_b200")
m = fov.OpenCLManager(0)
m.InitializeContext()
img = fov.ImageSampler(m)
orc = O.Oracle("ref") if args.ref else O.port()
orc.set_threads(len(os.sched_getaffinity(0)))
GAZES = [(0.5, 0.5), (0.65, 0.75), (0.02, 0.3), (0.98, 0.9), (0.0, 0.0), (1.0, 1.0), (0.3337, 0.6113)]

for name in [s for s in args.sizes.split(",") if s]:
    W, H = bench.WORKLOADS[name]
    ow, oh = bench.reduced(W), bench.reduced(H)
    for kind, frame in (("noise", O.lcg_frame(W, H, 7)), ("smooth", bench.synth_frame(W, H, 3))):
        src = m.upload(frame)
        for cx, cy in GAZES:
            red = m.upload(np.full((oh, ow, 4), 0xAB, np.uint8))
            img.SampleFrameLogPolarGPU(red, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
            lp = m.copy_to_host(np.empty((oh, ow, 4), np.uint8), red)
            want_lp = orc.img_sample_logpolar(frame, ow, oh, cx, cy, out=np.full((oh, ow, 4), 0xAB, np.uint8))
            bl = m.Buffer(4 * ow * oh)
            img.ApplyLogPolarGaussianBlur(bl, ow, oh, 4 * ow, red)
            got_bl = m.copy_to_host(np.empty((oh, ow, 4), np.uint8), bl)
            want_bl = orc.img_logpolar_blur(lp)
            full = m.Buffer(4 * W * H)
            img.InterpolateFrameLogPolarGPU(full, W, H, 4 * W, red, ow, oh, 4 * ow, cx, cy)
            got = m.copy_to_host(np.empty((H, W, 4), np.uint8), full)
            t0 = time.perf_counter()
            want = orc.img_interpolate_logpolar(lp, W, H, cx, cy)
            dt = time.perf_counter() - t0
            d = np.abs(got[..., :3].astype(np.int16) - want[..., :3].astype(np.int16)).max(axis=2)
            n = d.size
            print("%s %s gaze=(%g,%g): sample exact=%s blur maxdiff=%d (4 bytes equal=%s) | interpolate: "
                  "equal %.5f%%  1 LSB %d  >1 LSB %d  max %d  4th-byte mismatches %d  (oracle %.2f s)" % (
                      name, kind, cx, cy, np.array_equal(lp, want_lp),
                      int(np.abs(got_bl[..., :3].astype(np.int16) - want_bl[..., :3].astype(np.int16)).max()),
                      np.array_equal(got_bl, want_bl), 100.0 * (d == 0).sum() / n, int((d == 1).sum()),
                      int((d > 1).sum()), int(d.max()), int((got[..., 3] != want[..., 3]).sum()), dt),
                  flush=True)
            for b in (red, bl, full):
                b.free()
        src.free()

for name in [s for s in args.time.split(",") if s]:
    W, H = bench.WORKLOADS[name]
    ow, oh = bench.reduced(W), bench.reduced(H)
    src = m.upload(bench.synth_frame(W, H, 0))
    red, blur, full = m.Buffer(4 * ow * oh), m.Buffer(4 * ow * oh), m.Buffer(4 * W * H)
    m.memset(red, 0, 4 * ow * oh)
    gaze = bench.gaze_trace(args.steps + 3, 1, seed=1)[:, 0]
    flush = m.Buffer(256 << 20)

    def step(i):
        cx, cy = float(gaze[i, 0]), float(gaze[i, 1])
        m.memset(flush, i & 0xff, 256 << 20)  # L2 flush between frames (a single frame fits L2)
        img.SampleFrameLogPolarGPU(red, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
        img.ApplyLogPolarGaussianBlur(blur, ow, oh, 4 * ow, red)
        img.InterpolateFrameLogPolarGPU(full, W, H, 4 * W, blur, ow, oh, 4 * ow, cx, cy)
        img.SampleFrameRectGPU(red, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)

    for i in range(3):
        step(i)
    m.profile_reset()
    m.profile(True)
    for i in range(args.steps):
        step(3 + i)
    tot = m.profile_totals()
    m.profile(False)
    peak, _ = bench.peak_hbm_gbs()
    algo = {"img_sample_logpolar": 7 * ow * oh, "img_sample_rect": 7 * ow * oh,
            "img_logpolar_blur": 8 * ow * oh, "img_interpolate_logpolar": 4 * ow * oh + 4 * W * H}
    for k, v in sorted(tot.items()):
        ms = v[0] / v[1]
        gbs = algo[k] / (ms * 1e-3) / 1e9
        print("%s %-26s %.4f ms  %7.1f GB/s algorithmic  %.3f of %.0f" % (name, k, ms, gbs, gbs / peak, peak))
    for b in (src, red, blur, full, flush):
        b.free()
m.close()
