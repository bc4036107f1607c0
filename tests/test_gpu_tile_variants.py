"""interpolate_rect picks its tile height (8, 16 or 32 rows per warp) from the amount of work in
the call, so the small frames of the parity tests would only ever exercise the 8-row kernel.  This
module re-runs the interpolate parity tests in child processes with the height forced
(FOV360_INTERP_ROWS, read once per process by the launcher)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = ("small_golden or golden_hashes_sat_path or sample_and_interpolate or ragged_geometries "
          "or gaze_sweep or batched_pipeline or roundtrip_identity_near_gaze")


@pytest.mark.parametrize("rows", [16, 32])
def test_interpolate_parity_with_forced_tile_height(rows):
    env = dict(os.environ, FOV360_INTERP_ROWS=str(rows))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"),
                          "-x", "-q", "-k", SELECT, "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout
