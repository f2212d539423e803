#!/usr/bin/env python3
"""Extract the Student-t critical-value table (100 df x 7 alphas) that the reference
hard-codes at sim/src/output_analysis/t_scores.rs:41-140 into a plain data include.

Run once in the build container (needs /root/reference); the output files are committed,
so nothing reads the reference at test or run time.
"""
import re, sys, pathlib
src = pathlib.Path("/root/reference/sim/src/output_analysis/t_scores.rs").read_text()
body = src[src.index("fn t_lookup"):]
rows = re.findall(r"\[\s*([0-9.]+),\s*([0-9.]+),\s*([0-9.]+),\s*([0-9.]+),\s*([0-9.]+),\s*([0-9.]+),\s*([0-9.]+)\s*\]", body)
assert len(rows) == 100, len(rows)
out = ["/* Student-t upper critical values, rows df=1..100, columns alpha = .1 .05 .025 .01 .005 .001 .0005",
       " * data restated from sim/src/output_analysis/t_scores.rs:41-140 by tools/extract_t_table.py */"]
for r in rows:
    out.append("{" + ", ".join(r) + "},")
text = "\n".join(out) + "\n"
for p in sys.argv[1:]:
    pathlib.Path(p).write_text(text)
print("rows", len(rows))
