"""Committed golden vectors (tests/golden/traces.json, made by tools/gen_golden.py with the oracle):
the oracle must keep reproducing them, the interpreter's host build must match them, and so must the
CUDA kernels (GPU tier) -- independently of the oracle library being present on the GPU box."""
import json
import os

import numpy as np
import pytest

import networks

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traces.json")))
NAMES = sorted(GOLD["networks"])


def _rows(name):
    rows = GOLD["networks"][name]
    return (np.array([int(r["hash"], 16) for r in rows], np.uint64), np.array([r["steps"] for r in rows]),
            np.array([r["events"] for r in rows]), np.array([r["draws"] for r in rows]),
            np.array([float.fromhex(r["time"]) for r in rows]))


def test_golden_file_covers_every_network():
    assert NAMES == sorted(networks.ALL) and GOLD["steps"] == 1200 and GOLD["seed"] == 42


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    import oracle_api
    models, conns = networks.ALL[name]()
    h, steps, events, draws, t = _rows(name)
    ref = oracle_api.OracleSimulation(models, conns).run_batch(GOLD["seed"], 0, len(h), threads=2, nsteps=GOLD["steps"])
    assert (ref["hash"] == h).all() and (ref["steps"] == steps).all() and (ref["events"] == events).all()
    assert (ref["time"] == t).all()


@pytest.mark.parametrize("name", NAMES)
def test_interpreter_host_build_matches_golden(name):
    import hostsim_api
    models, conns = networks.ALL[name]()
    h, steps, events, draws, t = _rows(name)
    hs = hostsim_api.HostSim(models, conns, n=len(h), seed=GOLD["seed"])
    hs.step_n(GOLD["steps"])
    assert (hs.hash() == h).all() and (hs.steps() == steps).all() and (hs.events() == events).all()
    assert (hs.draws() == draws).all() and (hs.time() == t).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(name):
    from sim_b200 import Simulation
    models, conns = networks.ALL[name]()
    h, steps, events, draws, t = _rows(name)
    sim = Simulation.post(models, conns, replications=len(h), seed=GOLD["seed"], instrument=True, steps_per_launch=9)
    sim.step_n(GOLD["steps"])
    assert (sim.get_trace_hash() == h).all() and (sim.get_steps() == steps).all() and (sim.get_events() == events).all()
    assert (sim.get_draws() == draws).all()
    got = sim.get_global_time()
    assert (np.abs(got - t) <= 1e-12 * np.maximum(1.0, np.abs(t)))[np.isfinite(t)].all()
    assert (np.isfinite(got) == np.isfinite(t)).all()
