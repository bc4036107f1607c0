#!/usr/bin/env python3
"""Generates the committed golden fixtures FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It loads oracle/_ref/libfovref.so - the reference's own .cl kernel sources compiled by g++ through
oracle/ref_shim (see oracle/build_oracle.py) - and never the restatement in oracle/fov_oracle.c, so
the fixtures pin the restatement (tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py) to
the reference's arithmetic.  Outputs:

* golden.json  - FNV-1a-64 hashes / counts / probe pixels of every hot-path stage on seeded
                 synthetic frames at 1080p and 4K plus grid hashes up to 8K;
* small.npz    - complete input/output arrays of a 96x64 case (every stage), small enough to diff
                 element-wise;
* gnomonic_small.npz, gnomonic.json - Projections::GnomonicProjection viewports (complete arrays of
                 a 192x96 frame, hashes at 1080p).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

GAZES = [(0.5, 0.5), (0.65, 0.75), (0.02, 0.3), (0.98, 0.9), (0.0, 0.0), (1.0, 1.0), (0.0, 1.0)]


def h(a) -> str:
    return O.fnv1a64(a)


def sat_case(ref, W, H, ow, oh, gazes, seed=12345):
    frame = O.lcg_frame(W, H, seed)
    sat = ref.sat_encode(frame)
    grid = ref.sat_create_grid(ow, oh, W, H)
    case = {
        "W": W, "H": H, "ow": ow, "oh": oh, "seed": seed,
        "frame": h(frame), "sat": h(sat), "sat_last": [int(v) for v in sat[-1, -1]],
        "grid": h(grid), "gaze": [],
    }
    for cx, cy in gazes:
        red = ref.sat_sample_rect(sat, ow, oh, cx, cy, grid=grid, out=np.full((oh, ow, 4), 0xAB, np.uint8))
        written = int((red[..., 0:3] != 0xAB).any(axis=2).sum())
        red0 = ref.sat_sample_rect(sat, ow, oh, cx, cy, grid=grid)
        full = ref.sat_interpolate_rect(red0, W, H, cx, cy)
        case["gaze"].append({
            "cx": cx, "cy": cy, "reduced_ab": h(red), "reduced_zero": h(red0),
            "written_min": written, "centre_px": [int(v) for v in red0[oh // 2, ow // 2, :3]],
            "interp": h(full),
        })
    return case


def logpolar_case(ref, W, H, ow, oh, gazes, seed=12345):
    frame = O.lcg_frame(W, H, seed)
    grid = ref.img_create_logpolar_grid(ow, oh, W, H)
    rgrid = ref.img_create_grid(ow, oh, W, H)
    case = {"W": W, "H": H, "ow": ow, "oh": oh, "seed": seed, "lp_grid": h(grid),
            "rect_grid": h(rgrid), "gaze": []}
    for cx, cy in gazes:
        lp = ref.img_sample_logpolar(frame, ow, oh, cx, cy, grid=grid)
        rs = ref.img_sample_rect(frame, ow, oh, cx, cy, grid=rgrid)
        bl = ref.img_logpolar_blur(lp)
        it = ref.img_interpolate_logpolar(lp, W, H, cx, cy)
        case["gaze"].append({"cx": cx, "cy": cy, "logpolar": h(lp), "rect": h(rs), "blur": h(bl),
                             "interp_logpolar": h(it)})
    return case


def main() -> None:
    ref = O.Oracle("ref")
    gold = {"generator": "oracle/_ref/libfovref.so (reference .cl sources, g++ shim)",
            "hash": "FNV-1a-64 over the raw buffer", "sat": [], "logpolar": [], "grids": []}
    gold["sat"].append(sat_case(ref, 1920, 1080, 1072, 608, GAZES))
    gold["sat"].append(sat_case(ref, 3840, 1920, 2144, 1072, GAZES[:3]))
    gold["sat"].append(sat_case(ref, 640, 360, O.reduced_size(640), O.reduced_size(360), GAZES, seed=99))
    gold["logpolar"].append(logpolar_case(ref, 1920, 1080, 1072, 608, GAZES[:4]))
    gold["logpolar"].append(logpolar_case(ref, 640, 360, 368, 208, GAZES, seed=99))
    for W, H in [(1920, 1080), (3840, 1920), (7680, 3840), (1280, 720), (640, 360)]:
        ow, oh = O.reduced_size(W), O.reduced_size(H)
        g = ref.sat_create_grid(ow, oh, W, H)
        gi = ref.img_create_grid(ow, oh, W, H)
        gold["grids"].append({"W": W, "H": H, "ow": ow, "oh": oh, "sat_grid": h(g),
                              "img_grid": h(gi),
                              "x_head": [int(v) for v in g[0, :4, 0]],
                              "x_tail": [int(v) for v in g[0, -2:, 0]],
                              "y_first": int(g[0, 0, 1]), "y_last": int(g[-1, 0, 1])})
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(gold, fh, indent=1)

    # small complete vectors
    W, H, ow, oh = 96, 64, 64, 48
    frame = O.lcg_frame(W, H, 4242)
    sat = ref.sat_encode(frame)
    grid = ref.sat_create_grid(ow, oh, W, H)
    arrays = {"frame": frame, "sat": sat, "sat_grid": grid,
              "decode": ref.sat_decode(sat), "img_grid": ref.img_create_grid(ow, oh, W, H),
              "lp_grid": ref.img_create_logpolar_grid(ow, oh, W, H),
              "gazes": np.array(GAZES, np.float32)}
    for k, (cx, cy) in enumerate(GAZES):
        red = ref.sat_sample_rect(sat, ow, oh, cx, cy, grid=grid, out=np.full((oh, ow, 4), 0xAB, np.uint8))
        arrays["reduced_%d" % k] = red
        arrays["interp_%d" % k] = ref.sat_interpolate_rect(red, W, H, cx, cy)
        lp = ref.img_sample_logpolar(frame, ow, oh, cx, cy, out=np.full((oh, ow, 4), 0xAB, np.uint8))
        arrays["logpolar_%d" % k] = lp
        arrays["rect_%d" % k] = ref.img_sample_rect(frame, ow, oh, cx, cy, out=np.full((oh, ow, 4), 0xAB, np.uint8))
        arrays["blur_%d" % k] = ref.img_logpolar_blur(lp)
        arrays["interp_logpolar_%d" % k] = ref.img_interpolate_logpolar(lp, W, H, cx, cy)
    np.savez_compressed(os.path.join(HERE, "small.npz"), **arrays)
    print("wrote golden.json and small.npz")
    gnomonic_vectors(ref)


GNOMONIC_VIEWS = [(0.5, 0.5), (0.1, 0.9), (0.0, 0.0), (1.0, 1.0), (0.73, 0.31), (0.98, 0.5)]


def gnomonic_vectors(ref) -> None:
    """Projections::GnomonicProjection: complete viewports of a small frame + hashes at 1080p."""
    W, H, tw, th = 192, 96, 80, 48
    frame = O.lcg_frame(W, H, 31337)
    frame[..., 3] = (np.arange(W * H, dtype=np.uint32) % 251).reshape(H, W).astype(np.uint8)  # 4th byte travels
    arrays = {"frame": frame, "views": np.array(GNOMONIC_VIEWS, np.float32)}
    for k, (cx, cy) in enumerate(GNOMONIC_VIEWS):
        arrays["view_%d" % k] = ref.gnomonic(frame, tw, th, cx, cy)
    np.savez_compressed(os.path.join(HERE, "gnomonic_small.npz"), **arrays)
    big = O.lcg_frame(1920, 1080, 12345)
    gold = {"generator": "oracle/_ref/libfovref.so (projections_program.cl, g++ shim)",
            "W": 1920, "H": 1080, "seed": 12345, "tw": 960, "th": 540,
            "views": [{"cx": cx, "cy": cy, "hash": h(ref.gnomonic(big, 960, 540, cx, cy))}
                      for cx, cy in GNOMONIC_VIEWS]}
    with open(os.path.join(HERE, "gnomonic.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote gnomonic_small.npz and gnomonic.json")


if __name__ == "__main__":
    main()
