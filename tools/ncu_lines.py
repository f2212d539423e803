#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line: warp
instructions executed, active lanes per instruction and stall samples, sorted by instruction count.
Usage: ncu_lines.py page.csv [top] [--by-lanes]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
cur_file = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) >= 9 and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[7]), int(r[4]) if r[4].isdigit() else 0, int(r[8])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
tots = sum(l[4] for l in lines)
tott = sum(l[5] for l in lines)
print("total warp instructions", tot, "thread instructions", tott, "lanes/inst %.2f" % (tott / max(1, tot)), "samples", tots)
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%6.2f%% inst %5.1f lanes %6.2f%% stall  %s:%d  %s" % (100.0 * l[3] / tot, l[5] / max(1, l[3]), 100.0 * l[4] / max(1, tots), l[0], l[1], l[2][:80]))
# histogram of warp instructions by active lanes
hist = collections.Counter()
for l in lines:
    if l[3]:
        hist[min(32, int(l[5] / l[3] / 4) * 4)] += l[3]
print("warp instructions by active lanes (bucket of 4):", " ".join("%d:%.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(hist.items())))
