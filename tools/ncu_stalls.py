#!/usr/bin/env python3
"""Stall-reason totals and the top instructions per reason from `ncu --page source --csv`.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_stalls.py src.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:  # first kernel instance only
    if r and r[0] == "Kernel Name":
        break
    if len(r) > 8:
        data.append(r)
tot = collections.Counter()
for r in data:
    for c in cols:
        tot[c] += int(r[idx[c]] or 0)
T = sum(tot.values()) or 1
print(rows[0][1][:100])
for c, v in tot.most_common(9):
    print("  %-24s %8d %5.1f%%" % (c, v, 100.0 * v / T))
for key, _ in tot.most_common(5):
    print("  == " + key)
    for r in sorted(data, key=lambda r: -int(r[idx[key]] or 0))[:4]:
        print("     %6s x%-8s %s" % (r[idx[key]], r[idx["Instructions Executed"]],
                                       r[idx["Source"]].strip()[:72]))
