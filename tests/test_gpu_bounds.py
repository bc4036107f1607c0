"""Index checks of our own, in place of compute-sanitizer memcheck (closed on this GPU pool).

`foveated-360-video_b200/libfov360_check.so` is the library built with -DFOV360_BOUNDS_CHECK
(__graft_entry__.build() makes it): every table index and gathered coordinate of the sampling and
warp kernels is compared with its limit on the device and violations are counted.  The whole
geometry matrix of tools/sanitize_pass.py plus the gaze corners runs through it in a separate
process (FOV360_LIB selects the library) and must report zero.  This is the test that would have
caught the one out-of-bounds READ of round 2 (a table index formed from log(0) at the gaze pixel),
which the guard bands - they see writes only - could not."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "foveated-360-video_b200", "libfov360_check.so")

SCRIPT = r"""
import ctypes as C, importlib, sys
import numpy as np
sys.path.insert(0, %(root)r)
fov = importlib.import_module("foveated-360-video_b200")
m = fov.OpenCLManager(0); m.InitializeContext()
cnt, site = C.c_uint(0), C.c_uint(0)
assert m.lib.fov_debug_bounds_violations(m.ctx, C.byref(cnt), C.byref(site)) == 1, "not a checking build"
enc, dec, img = fov.SATEncoder(m), fov.SATDecoder(m), fov.ImageSampler(m)
rng = np.random.default_rng(0)
GAZES = [(0.5, 0.5), (0.0, 0.0), (1.0, 1.0), (0.0, 1.0), (1.0, 0.0), (0.02, 0.97), (0.98, 0.31),
         (0.25, 0.999), (0.5001, 0.0), (0.3337, 0.6113)]
for (W, H) in [(256, 128), (512, 192), (250, 130), (1000, 36), (31, 500), (333, 77), (1920, 1080)]:
    ow, oh = fov.reduced_dim(W), fov.reduced_dim(H)
    frame = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    src, sat = m.upload(frame), m.Buffer(12 * W * H)
    enc.EncodeFrameGPU(sat, src, W, H, 4 * W)
    r, full = m.Buffer(4 * ow * oh), m.Buffer(4 * W * H)
    lp, bl = m.Buffer(4 * ow * oh), m.Buffer(4 * ow * oh)
    for b in (r, lp):
        m.memset(b, 0, 4 * ow * oh)
    for cx, cy in GAZES:
        dec.SampleFrameRectGPU(r, ow, oh, 4 * ow, sat, W, H, cx, cy)
        dec.InterpolateFrameRectGPU(full, W, H, 4 * W, r, ow, oh, 4 * ow, cx, cy)
        fov.FoveateFramesGPU(m, 1, full, 4 * W * H, r, 4 * ow * oh, sat, 12 * W * H, src, 4 * W * H,
                             W, H, 4 * W, ow, oh, np.asarray([(cx, cy)], np.float32))
        img.SampleFrameRectGPU(lp, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
        img.SampleFrameLogPolarGPU(lp, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
        img.ApplyLogPolarGaussianBlur(bl, ow, oh, 4 * ow, lp)
        img.InterpolateFrameLogPolarGPU(full, W, H, 4 * W, bl, ow, oh, 4 * ow, cx, cy)
    m.Finish()
    for b in (src, sat, r, full, lp, bl):
        b.free()
m.lib.fov_debug_bounds_violations(m.ctx, C.byref(cnt), C.byref(site))
print("violations=%%d first_site=%%d launches=%%d" %% (cnt.value, site.value, m.launch_count))
m.close()
"""


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(CHECK_LIB), reason="libfov360_check.so not built")
def test_no_index_leaves_its_table_or_buffer():
    env = dict(os.environ, FOV360_LIB=CHECK_LIB)
    res = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True,
                         text=True, env=env, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    line = res.stdout.strip().splitlines()[-1]
    assert line.startswith("violations=0 "), line


def test_checking_library_is_built_with_the_product(fov):
    """CPU side: build() produces the checking variant next to the product library and it exports
    the same C ABI."""
    import ctypes as C

    path = fov.build_module.build_check()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    missing = [s for s in fov.header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
