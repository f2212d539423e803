#!/usr/bin/env python3
"""Profiling driver: brings a batch to steady state, then runs a few step-kernel launches inside a
cudaProfilerStart/Stop window (use `ncu --profile-from-start off`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

p = argparse.ArgumentParser()
p.add_argument("--network", default="mm1")
p.add_argument("--replications", type=int, default=1 << 20)
p.add_argument("--k", type=int, default=1)
p.add_argument("--warm-until", type=float, default=300.0)
p.add_argument("--launches", type=int, default=3)
p.add_argument("--instrument", action="store_true")
p.add_argument("--until-window", type=float, default=0.0,
               help="profile a step_until(warm_until + window) call instead of step_n (use ncu -s to skip launches)")
a = p.parse_args()

import torch
import networks
from sim_b200 import Simulation

models, conns = networks.ALL[a.network]()
stat = {"mm1": "processor-01"}.get(a.network)
sim = Simulation.post(models, conns, replications=a.replications, seed=42, steps_per_launch=a.k, stat_model_id=stat,
                      instrument=a.instrument)
sim.step_until(a.warm_until)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if a.until_window > 0:
    sim.step_until(a.warm_until + a.until_window)
else:
    sim.step_n(a.k * a.launches)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
c = sim.get_counters()
print("ok", sim.describe(), c)
