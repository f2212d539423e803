"""Coupled models (sim/src/models/coupled.rs:188-275): the oracle's restatement with the reference's
parked-message semantics, sim/tests/coupled.rs replayed on it, and what distinguishes a Coupled model from the
flattened network (late delivery of internal messages)."""
import numpy as np
import pytest

import networks
import oracle_api as O
from sim_b200 import output_analysis as oa


def _response_times(msgs, arrive, depart):
    arrivals = {m["content"].split()[-1]: m["time"] for m in msgs if arrive(m)}
    return [m["time"] - arrivals[m["content"].split()[-1]] for m in msgs if depart(m)]


def test_closure_under_coupling():
    """sim/tests/coupled.rs:14-207 with the reference's default stream (Pcg64Mcg::new(42), a fresh Simulation each):
    the response-time confidence intervals (SteadyStateOutput, alpha = 0.001) of the atomic and the coupled form overlap."""
    cis = []
    for idx, (models, conns) in enumerate((networks.coupled_atomic(), networks.coupled_closure())):
        sim = O.OracleSimulation(models, conns)
        e, msgs = sim.step_n(1000)
        assert e == 0
        if idx == 0:
            rt = _response_times(msgs, lambda m: m["target_id"] == "processor-01", lambda m: m["target_id"] == "storage-01")
        else:
            rt = _response_times(msgs, lambda m: m["target_id"] == "storage-02" and m["source_port"] == "start",
                                 lambda m: m["target_id"] == "storage-02" and m["source_port"] == "stop")
        assert len(rt) > 100
        ss = oa.SteadyStateOutput.post(rt)
        cis.append(ss.confidence_interval_mean(0.001))
    assert cis[0].lower() < cis[1].upper() and cis[1].lower() < cis[0].upper()


def test_internal_messages_are_parked_until_the_next_firing():
    """coupled.rs:186-255: the generator's job reaches the processor inside the coupled model only when the coupled
    model fires again (the next generation or completion), not at the next step: the first job's service starts at
    the SECOND generation time, and no step of the coupled simulation is a zero-time step."""
    models, conns = networks.coupled_closure(lam_gen=0.5, lam_proc=2.0, store_records=True)
    sim = O.OracleSimulation(models, conns)
    sim.set_rng_philox(7, 0)
    sim.trace_enable(True)
    e, msgs = sim.step_n(400)
    assert e == 0
    tr = sim.trace()
    gen = [ev["time"] for ev in tr if ev["kind"] == 1 and ev["action"] == 1]      # Generation records
    arr = [ev["time"] for ev in tr if ev["kind"] == 1 and ev["action"] == 2 and ev["index"] == 1]   # Arrival at the processor
    assert len(gen) > 50 and len(arr) > 50
    assert arr[0] > gen[0] and arr[0] in gen[1:] + [m["time"] for m in msgs]     # delivered late, at a later firing
    # the atomic form delivers it at the generation time itself (a zero-time step)
    am, ac = networks.coupled_atomic(lam_gen=0.5, lam_proc=2.0)
    a = O.OracleSimulation(am, ac)
    a.set_rng_philox(7, 0)
    e, amsgs = a.step_n(400)
    first_job = [m for m in amsgs if m["target_id"] == "processor-01"][0]
    assert first_job["time"] == gen[0] or abs(first_job["time"] - gen[0]) < 1e-9
    # every message the coupled model emits leaves through an external output coupling under ITS id and port names
    assert {(m["source_id"], m["source_port"]) for m in msgs} == {("coupled-01", "start"), ("coupled-01", "stop")}
    assert sim.get_status("coupled-01")[1] in ("Processing no messages", "Processing 0 messages")


def test_external_input_couplings_and_nested_models():
    """External input couplings deliver at once (coupled.rs:261-271) to every component they name, in coupling order;
    a chain of two Coupled models conserves jobs: everything generated is seen, and what the batcher releases
    arrives in threes (or on its timer)."""
    models, conns = networks.coupled_pipeline(store_records=True)
    sim = O.OracleSimulation(models, conns)
    sim.set_rng_philox(3, 0)
    e, msgs = sim.step_n(3000)
    assert e == 0
    per = dict(zip([c.id for c in conns], sim.connector_messages()))
    assert per["c1"] > 200
    assert 0 <= per["c1"] // 2 - per["c4"] <= 2          # the balancer's `a` path takes every other job
    assert per["c2"] <= per["c1"] and per["c2"] >= per["c1"] - 20   # jobs in service / parked at the end
    assert per["c3"] <= per["c2"] and per["c3"] >= per["c2"] - 8
    c = sim.counters()
    assert c["events"] > c["steps"]


def test_coupled_batches_run_and_count_component_events():
    models, conns = networks.coupled_closure()
    o = O.OracleSimulation(models, conns)
    r = o.run_batch(42, 0, 8, threads=4, nsteps=500)
    assert r["err"] == 0 and (r["steps"] == 500).all() and np.unique(r["hash"]).size == 8
    assert (r["events"] >= r["steps"]).all()
