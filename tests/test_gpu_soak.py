"""Soak of the SAT build's fence-free inter-CTA protocol, in place of `compute-sanitizer --tool
racecheck` (closed on this GPU pool): tools/sat_soak.py queues bursts of SAT builds of random
geometry with no synchronisation in between, under load from a second context, and compares every
table with numpy's wrapping cumulative sums.  The test runs a short soak; the long one
(`python tools/sat_soak.py`, 2000 launches) is recorded under profiles/."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_tool():
    spec = importlib.util.spec_from_file_location("sat_soak", os.path.join(ROOT, "tools", "sat_soak.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_numpy_sat_wraps_like_the_kernels():
    """CPU side: the checker itself - wrapping uint32 sums, last entry = channel totals mod 2^32."""
    import numpy as np

    tool = load_tool()
    small = np.full((300, 1000, 4), 255, np.uint8)
    sat = tool.numpy_sat(small)
    assert sat.dtype == np.uint32 and sat.shape == (300, 1000, 3)
    assert int(sat[-1, -1, 0]) == 255 * 300 * 1000
    assert int(sat[9, 4, 1]) == 255 * 10 * 5
    big = np.full((4400, 4096, 4), 255, np.uint8)  # 255 * 18 M pixels = 4.6e9 > 2^32: wraps
    assert int(tool.numpy_sat(big)[-1, -1, 2]) == (255 * 4400 * 4096) % (1 << 32)


@pytest.mark.gpu
@pytest.mark.parametrize("noise", [False, True])
def test_sat_protocol_soak(fov, noise):
    tool = load_tool()
    res = tool.soak(fov, 160, seed=3 + int(noise), noise=noise)
    assert res["launches"] == 160
    assert res["mismatching_tables"] == 0, res
