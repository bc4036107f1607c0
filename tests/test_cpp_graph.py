"""include/fov360.h from plain C: it must compile with `gcc -std=c11` (CPU check: the header is C,
not C++), and on a GPU the captured-graph replay of the offline runner's sequence must equal the
three eager calls and the golden hashes generated from the reference's kernels."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "graph_replay.c")
EXE = os.path.join(ROOT, "tests", "cpp", "graph_replay.bin")


def build_exe(fov):
    fov.load()
    libdir = os.path.dirname(fov.library_path())
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           SRC, "-o", EXE, "-L", libdir, "-lfov360", "-Wl,-rpath," + libdir])
    return EXE


def test_c_abi_header_is_plain_c(fov):
    exe = build_exe(fov)
    out = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    assert "fov_graph_launch" in out and "fov_sat_foveate_batched_dev" in out


@pytest.mark.gpu
def test_graph_replay_from_c_matches_eager_and_golden(fov, golden):
    exe = build_exe(fov)
    c = golden["sat"][0]
    args = [exe, str(c["W"]), str(c["H"]), str(c["seed"])]
    for g in c["gaze"]:
        args += [repr(g["cx"]), repr(g["cy"])]
    res = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    assert got["graph"] == got["eager"]
    assert got["graph"] == [g["interp"] for g in c["gaze"]]
    assert got["launches"] == 3 + 3 * len(c["gaze"]) + 3 * len(c["gaze"])  # warm-up, replays, eager
