#!/usr/bin/env python3
"""profiles/rNN_traffic.json from an `ncu --set full` capture: DRAM bytes per launch of every kernel
in the report, stamped with the hash of the source files that define the kernel AT CAPTURE TIME
(run this right after the capture, on the tree that was profiled).  bench.py re-hashes those files
when it fills `roofline.traffic` and reports the entry as stale if any of them changed.

    python tools/make_traffic_json.py gpurun_out/prof.ncu-rep 8k 16 profiles/r02_traffic.json \
        [capture description]
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = "foveated-360-video_b200/csrc/"
SOURCES = {  # kernel (profiler name in fov_profile_*) -> files whose change invalidates a capture
    "sat_onepass": [CSRC + "sat_onepass.cu", CSRC + "sat_common.cuh"],
    "sat_sample_rect": [CSRC + "sat_decode.cu", CSRC + "pixel_math.cuh"],
    "sat_interpolate_rect": [CSRC + "sat_decode.cu", CSRC + "pixel_math.cuh"],
    "img_sample_logpolar": [CSRC + "image_sampler.cu", CSRC + "pixel_math.cuh"],
    "img_sample_rect": [CSRC + "image_sampler.cu", CSRC + "pixel_math.cuh"],
    "img_logpolar_blur": [CSRC + "image_sampler.cu", CSRC + "pixel_math.cuh"],
    "img_interpolate_logpolar": [CSRC + "image_sampler.cu", CSRC + "pixel_math.cuh"],
}


def sha16(rel):
    with open(os.path.join(ROOT, rel), "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()[:16]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return int(round(float(value.replace(",", "")) * scale))


rep, workload, batch, out = sys.argv[1:5]
note = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                     check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
try:
    with open(out) as fh:
        doc = json.load(fh)
except OSError:
    doc = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` "
                       "captures (cold cache, one launch), with the sha256[:16] of the kernel's "
                       "source files at capture time; bench.py copies the entry matching its "
                       "workload/batch into roofline.traffic and flags it stale when a hash differs"}
entry = doc.setdefault(workload, {}).setdefault(str(batch), {})
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = next((k for k in SOURCES if k + "_kernel" in name or (k == "img_logpolar_blur" and "blur4" in name)), None)
    if not key or key in entry and entry[key].get("capture") == note:
        continue
    entry[key] = {
        "dram_read": to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]),
        "dram_write": to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]),
        "duration_us": round(float(r[idx["gpu__time_duration.sum"]].replace(",", "")) *
                             {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                              "msecond": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1), 1),
        "capture": note,
        "sources": {p: sha16(p) for p in SOURCES[key]},
    }
with open(out, "w") as fh:
    json.dump(doc, fh, indent=1)
print(json.dumps(entry, indent=1))
