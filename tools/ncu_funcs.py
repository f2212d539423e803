#!/usr/bin/env python3
"""Dynamic warp instructions per source function (core.cuh / step_kernel.cuh) from an ncu source-page CSV.
Usage: ncu_funcs.py page.csv [iterations]  (iterations: divide by this count to get instructions per iteration)"""
import csv, sys, re, bisect, collections, os
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sim_b200", "csrc")
funcs = {}
for fn in ("core.cuh", "step_kernel.cuh"):
    fs = []
    for i, l in enumerate(open(os.path.join(root, fn)).read().split("\n")):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?SIMB_HD\s+[\w:<>&\* ]+?\s+(\w+)\(", l) or re.match(r"\s*__device__ __forceinline__ \w+ (\w+)\(", l) or re.match(r"(simb_step_kernel)\(", l)
        if m:
            fs.append((i + 1, m.group(1)))
    funcs[fn] = fs
by = collections.Counter()
byt = collections.Counter()
f = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if len(r) >= 9 and r[0].isdigit():
        try:
            e, t = int(r[7]), int(r[8])
        except ValueError:
            continue
        name = f
        if f in funcs:
            k = bisect.bisect_right([x[0] for x in funcs[f]], int(r[0])) - 1
            name = funcs[f][k][1] if k >= 0 else f
        by[name] += e
        byt[name] += t
tot = sum(by.values())
print("total warp instructions %d (%.0f per iteration)" % (tot, tot / div))
for k, v in by.most_common(30):
    print("%7.2f%%  %8.1f/iter  %5.1f lanes  %s" % (100.0 * v / tot, v / div, byt[k] / max(1, v), k))
