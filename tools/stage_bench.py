#!/usr/bin/env python3
"""Per-kernel timing of the batched pipeline (CUDA events inside the library), no e2e / CPU legs.

    python tools/stage_bench.py [--workload 8k] [--batch 8] [--steps 20]
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
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--tag", default="")
ap.add_argument("--pad-zero", action="store_true",
                help="FOV_OPT_REDUCED_PAD_ZERO: sample_rect writes whole pixels (buffer cleared below)")
args = ap.parse_args()

fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(0)
m.InitializeContext()
B = args.batch
fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
base = bench.synth_frame(W, H, 0)
frames = np.stack([np.roll(base, 97 * f, axis=1) for f in range(B)])
src, sat, red, full = m.upload(frames), m.Buffer(B * sb), m.Buffer(B * rb), m.Buffer(B * fb)
m.memset(red, 0, B * rb)
m.set_option(m.OPT_REDUCED_PAD_ZERO, args.pad_zero)
gaze = bench.gaze_trace(args.steps + 3, B, seed=1)
for i in range(3):
    fov.FoveateFramesGPU(m, B, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh, gaze[i])
m.profile_reset()
m.profile(True)
for i in range(args.steps):
    fov.FoveateFramesGPU(m, B, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh, gaze[3 + i])
tot = m.profile_totals()
m.profile(False)
step = sum(v[0] for v in tot.values()) / args.steps
ab = bench.algorithmic_bytes(W, H, ow, oh)
peak, _ = bench.peak_hbm_gbs()
print("%s %s B=%d: %.4f ms/step  %.0f fps  %.1f%% of pipeline roofline" % (
    args.tag, args.workload, B, step, B / step * 1e3, 100 * ab["total"] * B / step / 1e6 / peak))
for k, (ms, cnt) in sorted(tot.items()):
    print("   %-24s %.4f ms/launch x %d" % (k, ms / cnt, cnt // args.steps))
m.close()
