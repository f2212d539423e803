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


# ---- the device: Coupled models lowered into the flat model list (sim_b200/csrc/lower.cpp, core.cuh micro) --------
@pytest.mark.gpu
def test_closure_under_coupling_on_the_device():
    """sim/tests/coupled.rs:14-207 through the string API on the GPU (the reference's default stream): the messages
    step_n(1000) returns give response-time intervals that overlap between the atomic and the coupled form -- and they
    are the oracle's messages, one for one."""
    import json
    from sim_b200 import WebSimulation
    from sim_b200.simulator import RNG_PCG64MCG, connectors_to_json, models_to_json
    cis = []
    for idx, (models, conns) in enumerate((networks.coupled_atomic(), networks.coupled_closure())):
        web = WebSimulation.post_json(models_to_json(models), connectors_to_json(conns), 1, seed=42, rng_kind=RNG_PCG64MCG,
                                      trace_capacity=1 << 16)
        msgs = json.loads(web.step_n_json(1000))
        o = O.OracleSimulation(models, conns)
        e, ref = o.step_n(1000)
        assert e == 0 and len(msgs) == len(ref)
        for got, want in zip(msgs, ref):
            assert (got["sourceId"], got["sourcePort"], got["targetId"], got["targetPort"], got["content"]) == \
                   (want["source_id"], want["source_port"], want["target_id"], want["target_port"], want["content"])
            assert abs(got["time"] - want["time"]) <= 1e-12 * max(1.0, abs(want["time"]))
        key = {"source_id": "sourceId", "source_port": "sourcePort", "target_id": "targetId"}
        norm = [{k: m[v] for k, v in key.items()} | {"time": m["time"], "content": m["content"]} for m in msgs]
        if idx == 0:
            rt = _response_times(norm, lambda m: m["target_id"] == "processor-01", lambda m: m["target_id"] == "storage-01")
        else:
            rt = _response_times(norm, lambda m: m["target_id"] == "storage-02" and m["source_port"] == "start",
                                 lambda m: m["target_id"] == "storage-02" and m["source_port"] == "stop")
        cis.append(oa.SteadyStateOutput.post(rt).confidence_interval_mean(0.001))
    assert cis[0].lower() < cis[1].upper() and cis[1].lower() < cis[0].upper()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["coupled_closure", "coupled_pipeline"])
@pytest.mark.parametrize("k", [1, 16])
def test_coupled_networks_match_the_oracle_on_the_production_kernel(name, k):
    """Uninstrumented handles (the kernel a benchmark would run) against the oracle: steps, events (component
    transitions), clocks, every atomic model's until_next_event, errors -- on step_until, across launch boundaries."""
    from sim_b200 import Simulation
    models, conns = networks.COUPLED[name]()
    n, until = 200, 120.0 if name == "coupled_pipeline" else 4000.0
    sim = Simulation.post(models, conns, replications=n, seed=8, instrument=False, steps_per_launch=k)
    assert sim.describe()["kernel_shape"] == "generic"
    sim.step_until(until)
    ref = O.OracleSimulation(models, conns).run_batch(8, 0, n, threads=8, until=until)
    assert ref["err"] == 0 and (sim.get_errors() == 0).all()
    assert (sim.get_steps() == ref["steps"]).all() and (sim.get_events() == ref["events"]).all()
    t = sim.get_global_time()
    assert np.all(np.abs(t - ref["time"]) <= 1e-12 * np.abs(ref["time"]))
    inst = Simulation.post(models, conns, replications=n, seed=8, instrument=True, steps_per_launch=k)
    inst.step_until(until)
    assert (inst.get_trace_hash() == ref["hash"]).all()
    assert np.array_equal(inst.get_message_counts(), np.asarray(ref["conn"], np.uint64))


@pytest.mark.gpu
def test_messages_to_a_coupled_model_are_one_message_each():
    """A message to a Coupled model whose external input couplings fan it out to two components is ONE pending
    message (get_messages / inject_input), delivered to both components in coupling order."""
    from sim_b200 import Message, Simulation
    models, conns = networks.coupled_pipeline(store_records=True)
    sim = Simulation.post(models, conns, replications=3, seed=4, instrument=True, trace_replications=3, trace_capacity=1 << 14)
    o = O.OracleSimulation(models, conns)
    o.set_rng_philox(4, 1)
    o.trace_enable(True)
    for _ in range(40):  # walk until the generator's job is in flight towards the coupled model
        sim.step()
        o.step()
        want = [(m["target_id"], m["target_port"], m["content"]) for m in o.get_messages()]
        got = [(m.target_id, m.target_port, m.content) for m in sim.get_messages(1)]
        assert got == want
    sim.inject_input(Message("manual", "manual", "stage", "in", 0.0, "extra 1"))
    o.trace_intern("extra 1")
    o.inject_input("stage", "in", "extra 1")
    assert [(m.target_id, m.target_port, m.content) for m in sim.get_messages(1)][-1] == ("stage", "in", "extra 1")
    sim.step_n(300)
    e, _ = o.step_n(300, collect=False)
    assert e == 0 and sim.get_trace_hash()[1] == o.trace_hash()
    assert sim.get_status("generator", 1) == "Generating jobs"
    from sim_b200 import SimulationError
    with pytest.raises(SimulationError) as ei:   # components are not models of the Simulation (mod.rs:106-113)
        sim.get_status("balancer", 1)
    assert ei.value.name == "ModelNotFound"
