"""Differential test on seeded random networks (tests/random_networks.py): the engine -- its CPU build here, the
device under `-m gpu` -- against the oracle, replication by replication: error code (InvalidMessage and
InvalidModelState arise on their own in random wiring), steps, events, draws consumed, clock, and the hash over every
record and message.

Two allowances, both in DESIGN.md's ledger: a replication the engine stops with CapacityError (its containers are
bounded, the reference's Vecs are not -- random wiring often grows a queue without bound) is not compared; and the
trace hash is not compared when the network holds a Stopwatch, whose minimum / maximum job is exact only while job names
are not reused (two generators both send "job 1")."""
import numpy as np
import pytest

import hostsim_api as H
import oracle_api as O
from random_networks import random_network

STEPS = 400


def _oracle(models, conns, seed, rep, steps=STEPS, until=None):
    o = O.OracleSimulation(models, conns)
    o.set_rng_philox(seed, rep)
    o.trace_enable(False)
    e, _ = o.step_n(steps, collect=False) if until is None else o.step_until(until, collect=False)
    return e, o


def _same_time(a, b):
    return a == b or (a != a and b != b) or abs(a - b) <= 1e-12 * max(1.0, abs(a), abs(b))


@pytest.mark.parametrize("acyclic", [False, True])
@pytest.mark.parametrize("block", range(8))
def test_engine_matches_oracle_on_random_networks(block, acyclic):
    compared = 0
    for seed in range(5000 + 25 * block, 5000 + 25 * (block + 1)):
        models, conns = random_network(seed, acyclic=acyclic)
        hs = H.HostSim(models, conns, n=2, seed=seed)
        hs.set_reload_every(1 + seed % 7)   # state stored and reloaded as consecutive launches do
        hs.step_n(STEPS)
        stopwatch = any(m.inner.type_name == "Stopwatch" for m in models)
        for r in range(2):
            he = int(hs.errors()[r])
            if he == 32:
                continue
            e, o = _oracle(models, conns, seed, r)
            c = o.counters()
            what = "seed %d acyclic %s replication %d" % (seed, acyclic, r)
            assert e == he, what
            assert (c["steps"], c["events"], o.rng_draws()) == (int(hs.steps()[r]), int(hs.events()[r]), int(hs.draws()[r])), what
            assert _same_time(o.get_global_time(), float(hs.time()[r])), what
            assert stopwatch or o.trace_hash() == int(hs.hash()[r]), what
            compared += 1
    assert compared >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("acyclic", [False, True])
def test_device_matches_oracle_on_random_networks(acyclic):
    from sim_b200 import Simulation
    compared = 0
    for seed in range(7000, 7030):
        models, conns = random_network(seed, acyclic=acyclic)
        sim = Simulation.post(models, conns, replications=64, seed=seed, instrument=True, steps_per_launch=1 + seed % 5)
        try:
            sim.step_n(STEPS)
        except Exception:  # noqa: BLE001 - per-replication error codes are compared below
            pass
        errs, steps, events, draws = sim.get_errors(), sim.get_steps(), sim.get_events(), sim.get_draws()
        times, hashes = sim.get_global_time(), sim.get_trace_hash()
        stopwatch = any(m.inner.type_name == "Stopwatch" for m in models)
        for r in (0, 63):
            if int(errs[r]) == 32:
                continue
            e, o = _oracle(models, conns, seed, r)
            c = o.counters()
            what = "seed %d acyclic %s replication %d" % (seed, acyclic, r)
            assert e == int(errs[r]), what
            assert (c["steps"], c["events"], o.rng_draws()) == (int(steps[r]), int(events[r]), int(draws[r])), what
            assert _same_time(o.get_global_time(), float(times[r])), what
            assert stopwatch or o.trace_hash() == int(hashes[r]), what
            compared += 1
    assert compared >= 20
