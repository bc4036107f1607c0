"""The benchmark's JSON contract: one line, the keys the driver reads, on both arms.  The reference
arm runs the reference's own kernels on the host (CPU test); the product arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                         text=True, cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--workload", "1080p", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_product_arm_line():
    d = run_bench("--workload", "1080p", "--batch", "4", "--steps", "5", "--warmup", "3")
    assert BASE_KEYS <= set(d) and "impl" not in d or d.get("impl") != "reference"
    assert d["metric"].startswith("frames/s") and d["unit"] == "frames/s" and d["value"] > 0
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "u32"
    assert d["gpu_launches"] == 3 * 5  # three kernels per step, none hidden in a library
    fb = 4 * 1920 * 1080 * 4
    assert d["e2e"]["h2d_bytes_per_step"] == fb and d["e2e"]["d2h_bytes_per_step"] == fb
    assert 0 < d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1
    assert d["value"] / cb["value"] > 10
    # round 2: the traffic entry says where it comes from and whether the kernel changed since
    assert {"traffic", "traffic_source", "traffic_state"} <= set(r)
    assert all("frac" in k and "alg_gbs" in k for k in r["kernels"].values())
    # the server-shaped lane moves NV12 up and the reduced NV12 buffer down
    es = d["e2e_server"]
    assert es["h2d_bytes_per_step"] == 4 * 1920 * 1080 * 3 // 2
    assert es["d2h_bytes_per_step"] == 4 * 1072 * 608 * 3 // 2 and es["value"] > 0
    # every other BASELINE configuration is in the same line
    cfg = d["configs"]
    assert {"4k_gaze_sweep_single", "4k_gaze_sweep_batch8", "8k_single_frame",
            "4k_logpolar_vs_logrect", "serving_4k_streams", "1080p_centre_gaze"} <= set(cfg)
    for name in ("4k_gaze_sweep_single", "4k_gaze_sweep_batch8", "8k_single_frame"):
        assert cfg[name]["frames_per_s"] > 0 and 0 < cfg[name]["pipeline_frac"] < 1.2
        assert set(cfg[name]["kernels"]) == {"sat_onepass", "sat_sample_rect", "sat_interpolate_rect"}
    lp = cfg["4k_logpolar_vs_logrect"]
    assert set(lp["kernels"]) == {"img_sample_logpolar", "img_logpolar_blur", "img_interpolate_logpolar"}
    sv = cfg["serving_4k_streams"]
    assert sv["streams_per_gpu"] == 8 and sv["resident"]["frames_per_s"] > 0
    assert "perf_counter" in sv["resident"]["timing"] and sv["server_lane"]["value"] > 0
    assert cfg["1080p_centre_gaze"]["cpu_reference"]["value"] > 0


def test_committed_traffic_capture_matches_the_kernel_sources():
    """`roofline.traffic` comes from an `ncu --set full` capture under profiles/; every entry carries
    the hash of the kernel's source files at capture time.  This fails when a kernel was edited
    without a new capture (tools/refresh_profiles.sh), instead of the number going silently stale."""
    import sys

    sys.path.insert(0, ROOT)
    import bench

    for kernel in ("sat_onepass", "sat_sample_rect", "sat_interpolate_rect"):
        t = bench.measured_traffic("8k", 16, kernel)
        assert t["traffic"] and t["traffic_state"] == "current", (kernel, t)
    t = bench.measured_traffic("8k", 16, "no_such_kernel")
    assert t["traffic"] is None and t["traffic_state"] == "no capture"


def test_numa_binding_is_a_no_op_without_topology():
    import sys

    sys.path.insert(0, ROOT)
    import bench

    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_numa(0)  # no NVML / no GPU here: must not raise, must not bind
    assert info["bound"] is False and os.sched_getaffinity(0) == before
