#!/usr/bin/env python3
"""Generate tests/golden/traces.json with the CPU oracle: for every test network, replications 0..5 of the
Philox(seed 42) streams, step_n(1200) -> trace hash, step / event / draw counts and the final clock (hex).
The fixtures decouple the GPU parity tests from the oracle build and guard the oracle itself against
regressions (tests/test_golden.py checks both)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import networks  # noqa: E402
import oracle_api  # noqa: E402

SEED, REPS, STEPS = 42, 6, 1200
out = {"seed": SEED, "steps": STEPS, "rng": "philox4x32-10, key=(replication, seed)", "networks": {}}
for name in sorted(networks.ALL):
    models, conns = networks.ALL[name]()
    rows = []
    for r in range(REPS):
        o = oracle_api.OracleSimulation(models, conns)
        o.set_rng_philox(SEED, r)
        o.trace_enable(False)
        e, _ = o.step_n(STEPS, collect=False)
        c = o.counters()
        rows.append({"replication": r, "error": e, "hash": "%016x" % o.trace_hash(), "steps": c["steps"],
                     "events": c["events"], "draws": o.rng_draws(), "time": float(o.get_global_time()).hex()})
    out["networks"][name] = rows
path = os.path.join(ROOT, "tests", "golden", "traces.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
