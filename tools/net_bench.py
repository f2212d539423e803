#!/usr/bin/env python3
"""Throughput of the step kernel on the test networks (tests/networks.py): events/s from the handle's
own counters (device time of the launches).  Usage: net_bench.py [--networks a,b] [--replications N] [--k K]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

p = argparse.ArgumentParser()
p.add_argument("--networks", default="mm1,tandem,gateway_network,large_mixed")
p.add_argument("--replications", type=int, default=1 << 18)
p.add_argument("--k", type=int, default=0)
p.add_argument("--warm-until", type=float, default=50.0)
p.add_argument("--window", type=float, default=200.0)
a = p.parse_args()

import networks
from sim_b200 import Simulation

for name in a.networks.split(","):
    models, conns = networks.ALL[name]()
    sim = Simulation.post(models, conns, replications=a.replications, seed=42, steps_per_launch=a.k)
    sim.step_until(a.warm_until)
    c0 = sim.get_counters()
    sim.reset_counters()
    sim.step_until(a.warm_until + a.window)
    c1 = sim.get_counters()
    ev, st = c1["events"] - c0["events"], c1["steps"] - c0["steps"]
    d = sim.describe()
    print(json.dumps({"network": name, "models": len(models), "replications": a.replications,
                      "events_per_s": ev / (c1["kernel_ms"] * 1e-3), "steps_per_s": st / (c1["kernel_ms"] * 1e-3),
                      "kernel_ms": c1["kernel_ms"], "launches": c1["launches"], "tile": d.get("tile"),
                      "kernel_shape": d.get("kernel_shape"), "K": d.get("steps_per_launch")}))
    sim.close()
