#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libfov360.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt

Lists every kernel with its instruction count and opcode mix, and flags the mnemonics that prove
what the kernels are built from: UBLKCP / UTMA* (TMA bulk copies), FADD2 / FMUL2 / FFMA2 (Blackwell
packed fp32), IDP (dp2a/dp4a), D* (fp64 pipe), MUFU, and - there must be none - HMMA / UTC*MMA."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "foveated-360-video_b200", "libfov360.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("library: %s\ncubin architectures: %s" % (os.path.relpath(lib, ROOT), ", ".join(archs)))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|fov::", "", cur).split("(")[0]
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
FLAG = ("UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "FADD2", "FMUL2", "FFMA2", "IDP", "DFMA",
        "DADD", "DMUL", "MUFU", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "LDGSTS", "REDUX",
        "ATOMG", "ACQBULK", "SYNCS")
for name, c in kern.items():
    tot = sum(c.values())
    print("\n%s: %d instructions" % (name, tot))
    print("  " + "  ".join("%s %d" % kv for kv in c.most_common(14)))
    marks = ["%s x%d" % (k, c[k]) for k in FLAG if c.get(k)]
    if marks:
        print("  notable: " + ", ".join(marks))
all_ops = collections.Counter()
for c in kern.values():
    all_ops.update(c)
print("\ntensor-core instructions in the library (HMMA / UTC*MMA / LDTM / STTM): %d - no stage is a "
      "contraction" % sum(v for k, v in all_ops.items() if k.startswith(("HMMA", "UTC", "LDTM", "STTM"))))
