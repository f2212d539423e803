#!/usr/bin/env python3
"""Instruction budget of a step-kernel capture by PHASE: the ncu source page (`ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass`) aggregated per enclosing function of core.cuh / region of step_kernel.cuh, and grouped
into the phases of a step (RNG, samplers, model handlers, routing, time advance, tile load / store ...).

usage: ncu_phases.py source.csv [warp_steps]   (warp_steps: replications / 32 x steps per launch, for per-step figures)"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sim_b200", "csrc")

PHASE_OF = [  # first match wins
    (r"philox|cache_push|RC$|key[01]|rep_lo|rep_hi", "Philox (block function, draw cache)"),
    (r"next_u64|standard_f64|open01", "uniform draws (next_u64, float conversion)"),
    (r"sample_exp1|sample_standard_normal|sample_continuous|sample_uniform_int|sample_index|sample_bernoulli|sample_discrete|log_gamma|powi", "samplers (ziggurat, inversion)"),
    (r"flush_draws|defer_draw", "deferred-draw bookkeeping"),
    (r"events_ext|ext_model|ext2|ext_special|ring_find", "events_ext handlers"),
    (r"events_int|int2|release_n|release_one", "events_int handlers"),
    (r"emit|mail_word", "routing / mailbox"),
    (r"ring_push|ring_pop|ring_slot|RW$|RD$", "queue rings (global memory)"),
    (r"wait_sample", "waiting-time statistic"),
    (r"record", "instrumentation"),
    (r"^step$|^step2$|micro|apply_u|prog_init", "step driver: votes, min / advance walk, dispatch"),
    (r"lane_load|lane_store", "lane context load / store"),
    (r"^D$|^W$|phase$|set_phase", "state tile accessors"),
    (r"warp_|lowest_bit|bits_f64|mulhi64", "helpers"),
]


def function_ranges(path):
    """(first line, name) of every function-like definition, in order."""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:SIMB_HD|__device__|__global__|inline|static)[^;(]*?\b(\w+)\s*\(")
    for i, line in enumerate(open(path, errors="ignore"), 1):
        m = pat.match(line)
        if m and not line.strip().startswith("//"):
            out.append((i, m.group(1)))
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1], encoding="latin-1")))
    warp_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    ranges = {f: function_ranges(os.path.join(CSRC, f)) for f in ("core.cuh", "step_kernel.cuh")}
    per_fn = collections.Counter()
    per_fn_thr = collections.Counter()
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if len(r) > 8 and r[0].isdigit():
            try:
                n, t = int(r[7]), int(r[8])
            except ValueError:
                continue
            ln = int(r[0])
            name = cur
            if cur in ranges:
                name = cur + ":?"
                for first, fn in ranges[cur]:
                    if first <= ln:
                        name = fn
                    else:
                        break
                if cur == "step_kernel.cuh":
                    name = "kernel body (tile load / store, step loop)"
            per_fn[name] += n
            per_fn_thr[name] += t
    total = sum(per_fn.values())
    phases = collections.Counter()
    phases_thr = collections.Counter()
    for fn, n in per_fn.items():
        ph = "other (" + fn + ")"
        if fn.startswith("kernel body"):
            ph = "kernel body: tile load / store, step loop"
        else:
            for pat, label in PHASE_OF:
                if re.search(pat, fn):
                    ph = label
                    break
        phases[ph] += n
        phases_thr[ph] += per_fn_thr[fn]
    print("warp instructions: %d, lanes / instruction %.2f" % (total, sum(per_fn_thr.values()) / max(1, total)))
    if warp_steps:
        print("warp-steps: %d -> %.1f warp instructions per warp-step" % (warp_steps, total / warp_steps))
    print("\n%-58s %7s %12s %7s" % ("phase", "share", "inst/w-step" if warp_steps else "inst", "lanes"))
    for ph, n in phases.most_common():
        print("%-58s %6.2f%% %12.1f %7.1f" % (ph[:58], 100.0 * n / total, n / warp_steps if warp_steps else n, phases_thr[ph] / max(1, n)))
    print("\n%-40s %7s %12s %7s" % ("function", "share", "inst/w-step" if warp_steps else "inst", "lanes"))
    for fn, n in per_fn.most_common(25):
        print("%-40s %6.2f%% %12.1f %7.1f" % (fn[:40], 100.0 * n / total, n / warp_steps if warp_steps else n, per_fn_thr[fn] / max(1, n)))


if __name__ == "__main__":
    main()
