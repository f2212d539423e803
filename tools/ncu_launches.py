#!/usr/bin/env python3
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`)
per kernel: launches, total device time, share.  usage: ncu_launches.py X.csv "<command that was profiled>" > summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], encoding="latin-1")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", r[kn])
    name = re.sub(r"\(Simb.*$", "", name).replace("void ", "").replace("simb::", "")
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("ncu launch list of `%s`" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("(--metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: the SHARE carries over)")
print("simb_step_kernel<RNG, INSTR, SHAPE, VARIANT, RARE, BASIC>: SHAPE -1 = generic interpreter, 0 = mm1, 1 = mm1_plain; VARIANT =")
print("streaming launch (baked) / narrow tile (generic)\n")
print("%-70s %8s %14s %8s %10s" % ("kernel", "launches", "total us", "share", "avg us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %14.1f %7.2f%% %10.1f" % (k[:70], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
print("total: %d launches, %.1f ms of device time" % (sum(v[0] for v in agg.values()), tot / 1e3))
