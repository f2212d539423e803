#!/usr/bin/env python3
"""Differential fuzz on the device: seeded random networks (tests/random_networks.py) through the C ABI against the
oracle, the long-running sibling of tests/test_random_networks.py.  Usage (from the repo root, on a GPU box):
    python tools/fuzz_device.py [first_seed] [last_seed]
Steps per launch cycle through 1..9, 32, 128 and the engine's own choice.  Prints every disagreement and a summary line; see DESIGN.md section 5 for the two allowances."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_random_networks as T
from random_networks import random_network, random_coupled_network, random_injections
import hostsim_api as H
from sim_b200 import Simulation, Message
bad = cmp_ = cap = 0
t0 = time.time()
FIRST = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
LAST = int(sys.argv[2]) if len(sys.argv) > 2 else FIRST + 1500
for seed in range(FIRST, LAST):
    mode = seed % 3
    models, conns = random_coupled_network(seed) if mode == 0 else random_network(seed, acyclic=(mode == 1))
    inject = random_injections(seed, models) if mode else []
    room = H.HostSim(models, conns, n=1).layout()["max_pending"] + len(inject)
    sim = Simulation.post(models, conns, replications=96, seed=seed, instrument=True, steps_per_launch=(1 + seed % 9, 32, 128, 0)[(seed // 9) % 4], max_pending=room)
    for tid, port, content in inject:
        sim.inject_input(Message("manual", "manual", tid, port, 0.0, content))
    try: sim.step_n(300)
    except Exception: pass
    errs, steps, events, draws = sim.get_errors(), sim.get_steps(), sim.get_events(), sim.get_draws()
    times, hashes = sim.get_global_time(), sim.get_trace_hash()
    sw = T._holds_stopwatch(models)
    for r in (0, 37, 95):
        if int(errs[r]) == 32: cap += 1; continue
        e, o = T._oracle(models, conns, seed, r, steps=300, inject=inject)
        if sw and o.trace_hash() != int(hashes[r]): continue
        c = o.counters(); cmp_ += 1
        ok = e == int(errs[r]) and (c["steps"], c["events"], o.rng_draws()) == (int(steps[r]), int(events[r]), int(draws[r])) and T._same_time(o.get_global_time(), float(times[r])) and o.trace_hash() == int(hashes[r])
        if not ok:
            bad += 1; print("MISMATCH", seed, r, e, errs[r], c, steps[r], events[r], o.rng_draws(), draws[r], o.get_global_time(), times[r], flush=True)
    sim.close()
    if bad > 5: break
print("bad", bad, "compared", cmp_, "capacity", cap, "%.1fs" % (time.time() - t0))
