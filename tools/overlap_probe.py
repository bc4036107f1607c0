#!/usr/bin/env python3
"""Does the foveation step gain from two in-order queues?  The SAT build is DRAM-bound, the inverse
warp issue-bound: with the 16 frames of a step split into groups that alternate between two
contexts, one group's sample + interpolate can overlap the other's SAT build.

    python tools/overlap_probe.py [--workload 8k] [--batch 16] [--groups 4] [--steps 30]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="8k")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--groups", type=int, default=4)
ap.add_argument("--queues", type=int, default=2)
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
B, G, Q = args.batch, args.groups, args.queues
per = B // G
fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
base = bench.synth_frame(W, H, 0)
ms = [fov.OpenCLManager(0) for _ in range(Q)]
for m in ms:
    m.InitializeContext()
    fov.SATDecoder(m).InitializeGrid(ow, oh, W, H)
groups = []
for g in range(G):
    m = ms[g % Q]
    frames = np.stack([np.roll(base, 97 * (g * per + f), axis=1) for f in range(per)])
    grp = {"m": m, "src": m.upload(frames), "sat": m.Buffer(per * sb), "red": m.Buffer(per * rb),
           "full": m.Buffer(per * fb)}
    m.memset(grp["red"], 0, per * rb)
    groups.append(grp)
gaze = bench.gaze_trace(args.steps + 3, B, seed=1)


def step(i):
    for g, grp in enumerate(groups):
        fov.FoveateFramesGPU(grp["m"], per, grp["full"], fb, grp["red"], rb, grp["sat"], sb, grp["src"],
                             fb, W, H, 4 * W, ow, oh, gaze[i, g * per:(g + 1) * per])


def sync():
    for m in ms:
        m.Finish()


for i in range(3):
    step(i)
sync()
t0 = time.perf_counter()
for i in range(args.steps):
    step(3 + i)
sync()
dt = time.perf_counter() - t0
ab = bench.algorithmic_bytes(W, H, ow, oh)
peak, _ = bench.peak_hbm_gbs()
fps = B * args.steps / dt
print("%s B=%d in %d group(s) over %d queue(s): %.4f ms/step  %.0f fps  %.1f%% of the pipeline roofline (wall clock)" % (
    args.workload, B, G, Q, 1e3 * dt / args.steps, fps, 100 * ab["total"] * fps / 1e9 / peak))
for m in ms:
    m.close()
