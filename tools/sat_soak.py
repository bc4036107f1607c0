#!/usr/bin/env python3
"""Soak of the one-pass SAT build's inter-CTA protocol (tickets, epoch-tagged 16-byte carry units,
decoupled look-back with no fences): the evidence `compute-sanitizer --tool racecheck` would give
if it were open on this pool.

    python tools/sat_soak.py [--launches 2000] [--seed 0] [--no-noise]

Random geometries and batch sizes (aligned ones take the one-pass kernel, the others the
three-kernel build that shares its scratch), queued in bursts with NO synchronisation inside a
burst, so launch k + 1 becomes resident (programmatic dependent launch) while launch k drains and
every launch finds the carry units of the previous ones in the scratch.  A second context keeps
the GPU busy with large SAT builds of its own on another stream ("noise"), which moves CTA timing
around from launch to launch.  Every table of every launch is compared with numpy's cumulative
sums (uint32, wrapping) - all of it, not a checksum.
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def numpy_sat(frame):
    """Inclusive 2-D prefix sums of channels 0..2, uint32 with wrap-around (sat_encoder.cc:67-135)."""
    v = frame[..., :3].astype(np.uint32)
    return np.cumsum(np.cumsum(v, axis=0, dtype=np.uint32), axis=1, dtype=np.uint32)


def soak(fov, launches, seed=0, noise=True, burst=8, max_px=1 << 20, log=None):
    rng = np.random.default_rng(seed)
    m = fov.OpenCLManager(0)
    m.InitializeContext()
    enc = fov.SATEncoder(m)
    other = enc2 = None
    if noise:
        other = fov.OpenCLManager(0)
        other.InitializeContext()
        enc2 = fov.SATEncoder(other)
        NW, NH, NN = 3840, 1920, 2
        nsrc = other.upload(rng.integers(0, 256, (NN, NH, NW, 4), dtype=np.uint8))
        nsat = other.Buffer(NN * NW * NH * 12)
    done = mismatches = onepass = 0
    t0 = time.perf_counter()
    while done < launches:
        work = []
        for _ in range(min(burst, launches - done)):
            aligned = rng.random() < 0.8
            W = int(rng.integers(32, 1025)) * 4 if aligned else int(rng.integers(33, 2000))
            H = int(rng.integers(1, max(2, min(2200, max_px // W))))
            n = int(rng.integers(1, 6))
            n = max(1, min(n, max_px // (W * H)))
            kind = rng.integers(0, 3)
            if kind == 0:
                frames = np.full((n, H, W, 4), 255, np.uint8)  # fastest-growing sums
            else:
                frames = rng.integers(0, 256, (n, H, W, 4), dtype=np.uint8)
            work.append((W, H, n, frames, m.upload(frames), m.Buffer(n * W * H * 12)))
            onepass += int(aligned)
        m.Finish()
        for W, H, n, frames, src, sat in work:  # nothing waits inside this loop
            if noise:
                enc2.EncodeFramesGPU(NN, nsat, NW * NH * 12, nsrc, NW * NH * 4, NW, NH, 4 * NW)
            enc.EncodeFramesGPU(n, sat, W * H * 12, src, W * H * 4, W, H, 4 * W)
        m.Finish()
        for W, H, n, frames, src, sat in work:
            got = m.copy_to_host(np.empty((n, H, W, 3), np.uint32), sat)
            for f in range(n):
                if not np.array_equal(got[f], numpy_sat(frames[f])):
                    mismatches += 1
                    if log:
                        log("MISMATCH launch %d: %dx%d n=%d frame %d" % (done, W, H, n, f))
            src.free()
            sat.free()
            done += 1
    if noise:
        other.Finish()
        nsrc.free()
        nsat.free()
        other.close()
    m.close()
    return {"launches": done, "aligned_one_pass": onepass, "mismatching_tables": mismatches,
            "seconds": round(time.perf_counter() - t0, 1), "noise": bool(noise), "seed": seed}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-noise", action="store_true")
    args = ap.parse_args()
    fov = importlib.import_module("foveated-360-video_b200")
    res = soak(fov, args.launches, args.seed, not args.no_noise, log=print)
    print(res)
    sys.exit(1 if res["mismatching_tables"] else 0)
