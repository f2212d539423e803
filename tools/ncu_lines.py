#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line:
instructions executed and stall samples, sorted."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) >= 8 and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[7]), int(r[4]) if r[4].isdigit() else 0))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
tots = sum(l[4] for l in lines)
print("total instructions", tot, "samples", tots)
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%6.2f%% inst %6.2f%% stall  %s:%d  %s" % (100.0 * l[3] / tot, 100.0 * l[4] / max(1, tots), l[0], l[1], l[2][:90]))
