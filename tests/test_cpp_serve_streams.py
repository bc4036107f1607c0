"""Concurrent connections, one thread and one context each, through the C++ drop-in headers
(tests/cpp/serve_streams.cc): the library is thread-compatible like the reference's per-connection
objects (video_server.cc:62-66, 85), and every connection's encoder surface equals the oracle's."""
import json
import os
import subprocess

import numpy as np
import pytest

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "serve_streams.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "serve_streams.bin")
TRACE = os.path.join(ROOT, "tests", "golden", "gaze_trace.txt")


def build_exe(fov):
    fov.load()
    libdir = os.path.dirname(fov.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                           SRC, "-o", EXE, "-L", libdir, "-lfov360", "-Wl,-rpath," + libdir])
    return EXE


def test_serve_streams_compiles_and_links(fov):
    assert os.path.exists(build_exe(fov))


@pytest.mark.gpu
def test_concurrent_connections_match_oracle(fov, oracle):
    exe = build_exe(fov)
    streams, frames, W, H = 6, 12, 512, 256
    res = subprocess.run([exe, TRACE, str(streams), str(frames), str(W), str(H)], capture_output=True,
                         text=True, check=True, timeout=600)
    got = json.loads(res.stdout)
    assert got["streams"] == streams and len(got["results"]) == streams
    trace = fov.GazeViewPoints(TRACE).gaze_array()
    ow, oh = O.reduced_size(W), O.reduced_size(H)
    for s, r in enumerate(got["results"]):
        assert 0 <= r["device"] < got["devices"]
        rec = (frames - 1 + 7 * s) % len(trace)
        assert r["record"] == rec
        # the connection's NV12 surface: LCG bytes, seed 1000 + s, padding rule not applied
        buf = np.empty(W * H * 3 // 2, np.uint8)
        state = np.uint32(1000 + s)
        with np.errstate(over="ignore"):
            for i in range(buf.size):
                state = state * np.uint32(1664525) + np.uint32(1013904223)
                buf[i] = state >> np.uint32(24)
        y = buf[:W * H].reshape(H, W)
        uv = buf[W * H:].reshape(H // 2, W)
        rgb = oracle.yuv420p_to_rgb0(y, np.ascontiguousarray(uv[:, 0::2]), np.ascontiguousarray(uv[:, 1::2]))
        # pixels whose box misses the frame keep what an earlier frame wrote: replay the sequence
        sat = oracle.sat_encode(rgb)
        red = np.zeros((oh, ow, 4), np.uint8)
        for f in range(frames):
            k = (f + 7 * s) % len(trace)
            red = oracle.sat_sample_rect(sat, ow, oh, float(trace[k, 0]), float(trace[k, 1]), out=red)
        wy, wu, wv = oracle.rgb0_to_yuv420p(red)
        wuv = np.empty((oh // 2, ow), np.uint8)
        wuv[:, 0::2], wuv[:, 1::2] = wu, wv
        assert r["hash"] == O.fnv1a64(np.concatenate([wy.ravel(), wuv.ravel()])), s
