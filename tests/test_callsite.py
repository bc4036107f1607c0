"""Boundary proof on the reference's own call-site text.

oracle/build_callsite_check.py extracts FoveateLogCartesianVideo (run_satlogrectilinear.cc) and the
server connection loop body (video_server.cc) from the reference tree, and compiles that text
against include/fov360/*.h with stubs only for FFmpeg / VideoDecoder / VideoEncoder.  The CPU test
is the compile (it fails when a signature, type or cl:: name drifts from what the reference
writes); the GPU test runs the binary and holds its output to the golden hashes generated from the
reference's kernels."""
import importlib.util
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "callsite_check")
REF_DIR = os.environ.get("FOV_REF_DIR", "/root/reference/src")


def _builder():
    spec = importlib.util.spec_from_file_location(
        "build_callsite_check", os.path.join(ROOT, "oracle", "build_callsite_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not present")
def test_reference_call_sites_compile_against_dropin_headers(fov):
    fov.build_module.build()
    b = _builder()
    runner = b.extract_runner()
    setup, body = b.extract_server()
    # the text really is the reference's call sequence, with its own argument lists
    assert "sat_encoder.EncodeFrameGPU(cl_sat_buffer(), cl_source_frame()," in runner
    assert "video_decoder.source_codec_ctx, center_x, center_y);" in runner
    assert "sat_decoder.InterpolateFrameRectGPU(" in runner
    assert "cl::Buffer cl_sat_buffer(cl_manager->context, CL_MEM_READ_WRITE," in setup
    assert "video_decoder->source_codec_ctx, center_x, center_y);" in body
    assert "OpenCLManager::GetCLErrorString(ret)" in body
    out = b.build(force=True)
    assert out and os.path.exists(out)
    # nothing of the reference leaks into tracked files
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True,
                             text=True).stdout.strip()
    assert tracked == ""


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/callsite_check not built")
@pytest.mark.parametrize("k", [0, 1])
def test_reference_call_sites_reproduce_golden_hashes(golden, k):
    c = golden["sat"][0]
    assert (c["W"], c["H"], c["ow"], c["oh"]) == (1920, 1080, 1072, 608)
    gaze = os.path.join(ROOT, "tests", "golden", "callsite_gaze%d.txt" % k)
    res = subprocess.run([BIN, "%dx%d:%d:1" % (c["W"], c["H"], c["seed"]), gaze],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    g = c["gaze"][k]
    assert out["runner_interp"] == [g["interp"]]        # run_satlogrectilinear.cc:915-949
    assert out["server_reduced"] == [g["reduced_zero"]]  # video_server.cc:291-345
