#!/usr/bin/env python3
"""Device timing of RGB0 -> YUV420P / NV12 on the reduced buffers of a batch (CUDA events inside
the library).

    python tools/color_convert_bench.py [--workload 8k] [--batch 16] [--steps 50]
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="8k")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()

fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(0)
m.InitializeContext()
conv = fov.VideoFrameConverter(m)
B = args.batch
rng = np.random.default_rng(0)
red = m.upload(rng.integers(0, 256, (B, oh, ow, 4), dtype=np.uint8))
y, u, v, uv = m.Buffer(B * ow * oh), m.Buffer(B * ow * oh // 4), m.Buffer(B * ow * oh // 4), \
    m.Buffer(B * ow * oh // 2)


def planar():
    conv.RGB0ToYUV420PFrames(B, y, ow * oh, ow, u, v, ow * oh // 4, ow // 2, red, 4 * ow * oh, 4 * ow,
                             ow, oh)


def nv12():
    conv.RGB0ToNV12Frames(B, y, ow * oh, ow, uv, ow * oh // 2, ow, red, 4 * ow * oh, 4 * ow, ow, oh)


W2, H2 = W, H  # the decoder direction works on full frames
dy, duv = m.upload(rng.integers(0, 256, (B, H2, W2), dtype=np.uint8)), \
    m.upload(rng.integers(0, 256, (B, H2 // 2, W2), dtype=np.uint8))
rgb = m.Buffer(B * W2 * H2 * 4)


def decode_nv12():
    conv.NV12ToRGB0Frames(B, rgb, W2 * H2 * 4, 4 * W2, dy, W2 * H2, W2, duv, W2 * H2 // 2, W2, W2, H2)


def decode_planar():
    conv.YUV420PToRGB0Frames(B, rgb, W2 * H2 * 4, 4 * W2, dy, W2 * H2, W2, duv, duv.at(B * W2 * H2 // 4),
                             W2 * H2 // 4, W2 // 2, W2, H2)


for _ in range(3):
    planar()
    nv12()
    decode_nv12()
    decode_planar()
m.profile_reset()
m.profile(True)
for _ in range(args.steps):
    planar()
    nv12()
    decode_nv12()
    decode_planar()
tot = m.profile_totals()
m.profile(False)
peak, _ = bench.peak_hbm_gbs()
for k, (ms, cnt) in sorted(tot.items()):
    t = ms / cnt
    w, h = (W2, H2) if k.endswith("to_rgb0") else (ow, oh)
    bytes_per_launch = B * w * h * 5.5  # 4 B/px on the RGB0 side + 1.5 B/px on the YUV side
    print("%s %dx%d x%d: %.4f ms/launch, %.0f GB/s algorithmic (%.1f%% of the measured copy peak)" % (
        k, w, h, B, t, bytes_per_launch / t / 1e6, 100 * bytes_per_launch / t / 1e6 / peak))
m.close()
