"""The C++ drop-in layer (include/fov360/*.h): it must compile with plain g++ against the C ABI
(CPU check) and, on a GPU, reproduce the golden hashes through the reference's call sequence."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dropin_pipeline.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_pipeline.bin")


def build_exe(fov):
    lib = fov.load()  # builds libfov360.so when stale
    del lib
    libdir = os.path.dirname(fov.library_path())
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", libdir, "-lfov360", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return EXE


def test_dropin_headers_compile_and_link(fov):
    exe = build_exe(fov)
    assert os.path.exists(exe)
    out = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    assert "fov_sat_encode" in out and "fov_memcpy_h2d" in out


@pytest.mark.gpu
def test_dropin_call_sequence_matches_golden(fov, golden, oracle):
    exe = build_exe(fov)
    c = golden["sat"][0]
    g = c["gaze"][1]
    res = subprocess.run([exe, str(c["W"]), str(c["H"]), str(c["seed"]), repr(g["cx"]), repr(g["cy"])],
                         capture_output=True, text=True, check=True)
    got = json.loads(res.stdout)
    assert (got["ow"], got["oh"]) == (c["ow"], c["oh"])
    assert got["sat"] == c["sat"]
    assert got["reduced_zero"] == g["reduced_zero"]
    assert got["interp"] == g["interp"]
    lp = [x for x in golden["logpolar"][0]["gaze"] if (x["cx"], x["cy"]) == (g["cx"], g["cy"])][0]
    assert got["logpolar"] == lp["logpolar"]
    assert got["view_equal"] == 1  # Projections: fused viewport == interpolate + gnomonic
    # VideoFrameConverter: the encoder surface equals libswscale's conversion of the reduced buffer
    import numpy as np

    import _oracle as O
    red = oracle.sat_sample_rect(oracle.sat_encode(O.lcg_frame(c["W"], c["H"], c["seed"])),
                                 c["ow"], c["oh"], g["cx"], g["cy"])
    assert O.fnv1a64(red) == g["reduced_zero"]
    y, u, v = oracle.rgb0_to_yuv420p(red)
    assert got["yuv420p"] == O.fnv1a64(np.concatenate([y.ravel(), u.ravel(), v.ravel()]))
    assert got["launches"] >= 7
