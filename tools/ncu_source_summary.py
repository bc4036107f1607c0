#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: opcode mix and per-basic-block execution counts.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_source_summary.py src.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) > 8:
        cur["rows"].append(r)
for k in kernels[:1] if len(sys.argv) < 3 else kernels:
    idx = {h: i for i, h in enumerate(k["hdr"])}
    tot, byop, data = 0, collections.Counter(), []
    for r in k["rows"]:
        n = int(r[idx["Instructions Executed"]])
        src = r[idx["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        byop[op.split(".")[0]] += n
        tot += n
        data.append((n, src, r[idx["Avg. Threads Executed"]], int(r[idx["# Samples"]])))
    print(k["name"][:90], "total warp-instr", tot)
    for op, v in byop.most_common(22):
        print("  %-10s %12d %5.1f%%" % (op, v, 100.0 * v / tot))
    cnts = collections.Counter(d[0] for d in data)
    print("  blocks by execution count:")
    for c, m in sorted(cnts.items(), key=lambda x: -x[0] * x[1])[:10]:
        print("    exec %10d x %4d instrs = %5.1f%%" % (c, m, 100.0 * c * m / tot))
    ns = sum(d[3] for d in data)
    print("  top stall samples:")
    for d in sorted(data, key=lambda d: -d[3])[:12]:
        print("    %5.1f%%  thr=%s  %s" % (100.0 * d[3] / max(ns, 1), d[2], d[1][:80]))
