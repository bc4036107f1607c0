"""Gaze-trace ingestion and session placement: the drop-in
GazeViewPoints parser (include/fov360/gaze_view_points.h) and its Python mirror against
tests/golden/gaze_trace.json, which was produced by the REFERENCE's own parser
(src/gaze_view_points.cc compiled with g++, see tests/cpp/gaze_dump.cc) on tests/golden/gaze_trace.txt."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SRC = os.path.join(ROOT, "tests", "cpp", "gaze_dump.cc")
REF_SRC = "/root/reference/src"


def golden_records():
    with open(os.path.join(GOLD, "gaze_trace.json")) as fh:
        return json.load(fh)


def run_dump(tmp_path, extra, name, args=()):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", *extra, SRC, "-o", exe])
    res = subprocess.run([exe, os.path.join(GOLD, "gaze_trace.txt"), *args], capture_output=True,
                         text=True)
    assert res.returncode == 0, res
    return json.loads(res.stdout)


def test_cpp_parser_matches_reference_parser_output(tmp_path):
    # the second argument also runs SessionPlacement's self-check (non-zero exit on failure)
    got = run_dump(tmp_path, ["-I", os.path.join(ROOT, "include")], "gaze_ours.bin", args=("placement",))
    want = golden_records()
    assert len(got) == len(want) == 65
    assert got == want


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not present")
def test_golden_trace_is_what_the_reference_parser_returns(tmp_path):
    got = run_dump(tmp_path, ["-DUSE_REFERENCE", "-I", REF_SRC,
                              os.path.join(REF_SRC, "gaze_view_points.cc")], "gaze_ref.bin")
    assert got == golden_records()


def test_python_mirror_matches(fov):
    gv = fov.GazeViewPoints(os.path.join(GOLD, "gaze_trace.txt"))
    want = golden_records()
    assert len(gv.points) == len(want)
    for p, w in zip(gv.points, want):
        got = [p.frame, *p.view_point, *p.gaze_point, *p.pred_view_point, *p.pred_gaze_point]
        assert got[0] == w[0]
        assert np.array_equal(np.asarray(got[1:], np.float32), np.asarray(w[1:], np.float32)), w
    # the runner's use: gaze of frame f drives center_x / center_y (run_satlogrectilinear.cc:519-525)
    g = gv.gaze_array()
    assert g.dtype == np.float32 and g.shape == (len(want), 2)
    assert np.array_equal(g[3], np.asarray(want[3][3:5], np.float32))
    assert fov.GazeViewPoints(os.path.join(GOLD, "missing.txt")).points == []


def test_session_placement_python(fov):
    pl = fov.sharding.SessionPlacement(8)
    dev = [pl.acquire() for _ in range(19)]
    assert dev == [s % 8 for s in range(19)]  # == the benchmark's static stream -> GPU rule
    for d in (dev[0], dev[8], dev[5]):
        pl.release(d)
    assert [pl.acquire(), pl.acquire(), pl.acquire()] == [0, 5, 0]
    assert pl.sessions(0) == 3 and sum(pl.sessions(d) for d in range(8)) == 19
