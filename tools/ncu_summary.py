#!/usr/bin/env python3
"""Summarise ncu captures into profiles/ncu_summary.json (read by bench.py for roofline.traffic / .ncu).

usage: ncu_summary.py tag=path.ncu-rep|path_raw.csv [tag=path ...]
       (tags used by bench.py: fused, k1, k1_2p22, cfg_tandem, cfg_gateway_network, cfg_large_mixed; a .csv is the
       `ncu -i X.ncu-rep --page raw --csv` export made on the GPU box, so that the reports need not travel)
For each capture the LAST profiled launch of simb_step_kernel is summarised (steady state)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_mb",
    "dram__bytes_write.sum": "dram_write_mb",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving_per_issue",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}


def summarise(path):
    if path.endswith(".csv"):
        out = open(path, encoding="latin-1").read()
    else:
        out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    data = [r for r in rows[2:] if "simb_step_kernel" in r[hdr.index("Kernel Name")]]
    def complete(row):  # ncu leaves metrics of a launch it could not replay fully as nan / empty
        try:
            v = float(row[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            return v == v
        except (ValueError, IndexError):
            return False
    r = ([x for x in data if complete(x)] or data)[-1]
    res = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "launches_profiled": len(data), "source": os.path.basename(path)}
    for k, name in WANT.items():
        if k in hdr:
            try:
                v = float(r[hdr.index(k)].replace(",", ""))
            except ValueError:
                continue
            if v != v:  # ncu prints nan for a metric it could not collect: leave it out (NaN is not JSON)
                continue
            u = units[hdr.index(k)]
            if name == "duration_us" and u == "ms":
                v *= 1e3
            if name.endswith("_mb") and u == "Gbyte":
                v *= 1e3
            res[name] = v
    res["dram_bytes_per_launch"] = (res.get("dram_read_mb", 0) + res.get("dram_write_mb", 0)) * 1e6
    return res


def main():
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    summary = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        tag, rep = arg.split("=", 1)
        summary[tag] = summarise(rep)
    with open(path, "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    print(json.dumps(summary, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
