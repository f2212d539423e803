"""CPU tier for the device interpreter: sim_b200/csrc/core.cuh + lower.cpp compiled as plain C++
(tests/hostsim, a unit-test harness -- never part of the product) against the oracle, on identical
per-replication streams.  The same comparisons run against the real CUDA kernels in
test_gpu_parity.py."""
import numpy as np
import pytest

import hostsim_api as H
import networks
import oracle_api as O


def _oracle(models, conns, seed, rep, steps=None, until=None, store=False):
    o = O.OracleSimulation(models, conns)
    o.set_rng_philox(seed, rep)
    o.trace_enable(store)
    e, _ = (o.step_n(steps, collect=False) if steps is not None else o.step_until(until, collect=False))
    return e, o


@pytest.mark.parametrize("name", sorted(networks.ALL))
def test_step_n_hash_counts_time(name):
    models, conns = networks.ALL[name]()
    n, steps = 6, 2500
    hs = H.HostSim(models, conns, n=n, seed=42)
    hs.step_n(steps)
    for r in range(n):
        e, o = _oracle(models, conns, 42, r, steps=steps)
        c = o.counters()
        assert e == 0 and hs.errors()[r] == 0
        assert o.trace_hash() == int(hs.hash()[r])
        assert (c["steps"], c["events"], o.rng_draws()) == (int(hs.steps()[r]), int(hs.events()[r]), int(hs.draws()[r]))
        assert o.get_global_time() == hs.time()[r]  # same libm on both sides here: bit-equal


@pytest.mark.parametrize("k", [1, 2, 5])
def test_results_do_not_depend_on_steps_per_launch(k):
    """Reloading the lane state every k steps (what consecutive kernel launches do) must not change
    anything: deferred draws persist through the state, the Philox prefetch restarts cleanly."""
    for name in ("mm1", "gateway_network", "stochastic_gate"):
        models, conns = networks.ALL[name]()
        a = H.HostSim(models, conns, n=3, seed=11)
        b = H.HostSim(models, conns, n=3, seed=11)
        b.set_reload_every(k)
        a.step_n(1200)
        b.step_n(1200)
        assert (a.hash() == b.hash()).all() and (a.time() == b.time()).all() and (a.draws() == b.draws()).all()


def test_full_trace_and_counts_gateway_network():
    models, conns = networks.gateway_network()
    hs = H.HostSim(models, conns, n=2, seed=3, trace_reps=2, trace_cap=100000)
    hs.step_n(4000)
    tot_conn = np.zeros(len(conns), np.uint64)
    for r in range(2):
        e, o = _oracle(models, conns, 3, r, steps=4000, store=True)
        assert o.trace() == hs.trace(r)
        tot_conn += np.array(o.connector_messages(), np.uint64)
    assert (hs.conn_counts() == tot_conn).all()


def test_replication_offset_is_a_pure_relabelling():
    """Sharding invariance: local replication r of a shard with offset k is global replication k + r."""
    models, conns = networks.tandem()
    a = H.HostSim(models, conns, n=8, seed=9)
    b = H.HostSim(models, conns, n=3, seed=9, rep_offset=5)
    a.step_n(1500)
    b.step_n(1500)
    assert (a.hash()[5:8] == b.hash()).all() and (a.time()[5:8] == b.time()).all()


def test_step_until_semantics_and_wait_statistic():
    models, conns = networks.mm1()
    rec_models, _ = networks.mm1(store_records=True)
    hs = H.HostSim(models, conns, n=5, seed=42, stat_model_id="processor-01")
    hs.step_until(300.0)
    s, cnt = hs.wait()
    for r in range(5):
        e, o = _oracle(rec_models, conns, 42, r, until=300.0)
        assert o.trace_hash() == int(hs.hash()[r]) and o.get_global_time() == hs.time()[r]
        ws, wn = o.mean_wait("processor-01")
        assert wn == cnt[r] and abs(ws - s[r]) <= 1e-12 * max(1.0, abs(ws))
    before = hs.steps().copy()
    hs.step_until(300.0)  # already past `until`: exactly one more step each (mod.rs:279-286)
    assert (hs.steps() == before + 1).all()


def test_pcg_mode_is_the_reference_default_stream():
    models, conns = networks.mm1()
    hs = H.HostSim(models, conns, n=3, seed=42, rng_kind=1)
    hs.step_n(2000)
    for r in range(3):
        o = O.OracleSimulation(models, conns)
        o.set_rng_pcg(42 + r)  # r = 0: Pcg64Mcg::new(42), dynamic_rng.rs:7-9
        o.trace_enable(False)
        o.step_n(2000, collect=False)
        assert o.trace_hash() == int(hs.hash()[r]) and o.get_global_time() == hs.time()[r]


def test_external_stream_mode():
    models, conns = networks.exclusive_gateway()
    n, per = 4, 3000
    draws = np.stack([O.pcg_fill(77 + r, per) for r in range(n)], axis=1)
    hs = H.HostSim(models, conns, n=n, rng_kind=2)
    hs.set_stream(draws)
    hs.step_n(900)
    for r in range(n):
        o = O.OracleSimulation(models, conns)
        o.set_rng_external(draws[:, r])
        o.trace_enable(False)
        o.step_n(900, collect=False)
        assert o.trace_hash() == int(hs.hash()[r]) and o.rng_draws() == int(hs.draws()[r])
    short = H.HostSim(models, conns, n=1, rng_kind=2)
    short.set_stream(draws[:5, :1])
    short.step_n(900)
    assert short.errors()[0] == 37  # SIMB_ERR_STREAM_EXHAUSTED


def test_injection_and_storage_exchange():
    models, conns = networks.storage_exchange()
    hs = H.HostSim(models, conns, n=2, max_pending=4)
    assert hs.inject_input("storage-01", "store", "42") == 0
    hs.step_n(1)
    hs.inject_input("storage-01", "read", "")
    hs.step_n(1)
    assert [p[0] for p in hs.pending(0)] == ["42"]
    hs.inject_input("storage-02", "read", "")
    hs.step_n(1)
    assert hs.pending(1)[0][0] == "42" and (hs.time() == 0.0).all()


def test_gate_strands_jobs_when_activation_arrives_in_the_same_step():
    """gate.rs quirk: an activation arriving while Pass with queued jobs resets
    until_next_event to +INF and strands them until the next job arrives."""
    from sim_b200 import Connector, ContinuousRandomVariable, Gate, Generator, Model, Storage
    models = [Model.new("gen", Generator.new(ContinuousRandomVariable.Exp(1.0), None, "job", False, None)),
              Model.new("gate", Gate.new("job", "activation", "deactivation", "job", False)),
              Model.new("store", Storage.new("store", "read", "stored", False))]
    conns = [Connector.new("c1", "gen", "gate", "job", "job"), Connector.new("c2", "gen", "gate", "job", "activation"),
             Connector.new("c3", "gate", "store", "job", "store")]
    hs = H.HostSim(models, conns, n=1, seed=1, trace_reps=1, trace_cap=10000, queue_capacity=64)
    hs.step_n(40)
    e, o = _oracle(models, conns, 1, 0, steps=40, store=True)
    assert o.trace() == hs.trace(0)
    assert hs.conn_counts()[2] == 0  # every job is stranded: nothing ever reaches the storage
    assert hs.errors()[0] == 0
    hs.step_n(400)                   # ... until the bounded device queue overflows (no reference analogue)
    assert hs.errors()[0] == 32      # SIMB_ERR_CAPACITY


def test_runtime_errors_match_the_oracle():
    for kw, want in ((dict(capacity=0), 5), (dict(lam_gen=0.0), 15), (dict(lam_proc=-1.0), 15)):
        models, conns = networks.mm1(**kw)
        hs = H.HostSim(models, conns, n=1)
        hs.step_n(5)
        o = O.OracleSimulation(models, conns)
        o.set_rng_philox(42, 0)
        assert o.step_n(5)[0] == want and hs.errors()[0] == want
    models, conns = networks.mm1()
    hs = H.HostSim(models, conns, n=1)
    hs.inject_input("processor-01", "bogus", "x")
    hs.step_n(3)
    assert hs.errors()[0] == 7 and hs.steps()[0] == 0


def test_unbounded_reference_queue_overflows_with_capacity_error():
    """Processor default capacity is usize::MAX (processor.rs:40-42); the device ring is bounded."""
    models, conns = networks.mm1(capacity=None, lam_gen=5.0, lam_proc=0.5)  # rho = 10: the queue explodes
    hs = H.HostSim(models, conns, n=1, queue_capacity=8)
    hs.step_n(400)
    assert hs.errors()[0] == 32
    big = H.HostSim(models, conns, n=1, queue_capacity=4096)
    big.step_n(400)
    assert big.errors()[0] == 0


def test_all_passive_network():
    models, conns = networks.storage_exchange()
    hs = H.HostSim(models, conns, n=1)
    hs.step_n(1)
    assert hs.time()[0] == float("inf") and np.isnan(hs.until(0)[0])


BAKED = [("mm1", "processor-01"), ("mm1", None)]  # tools/gen_shapes.py BASELINE (network, statistic model)


@pytest.mark.parametrize("shape,name,stat", [(i, n, s) for i, (n, s) in enumerate(BAKED)])
def test_baked_shape_code_path_equals_generic_and_oracle(shape, name, stat):
    """The literal-index ("baked shape") expansion of the interpreter -- what the specialised CUDA
    kernels compile -- must behave exactly like the generic interpreter and the oracle."""
    models, conns = networks.ALL[name]()
    a = H.HostSim(models, conns, n=12, seed=5, instrument=False, stat_model_id=stat)
    b = H.HostSim(models, conns, n=12, seed=5, instrument=False, stat_model_id=stat)
    assert b.use_shape(shape)
    assert not b.use_shape(shape + 7)  # no such baked shape
    assert b.use_shape(shape)
    b.set_reload_every(3)
    # the baked path prefetches Philox draws into a shared-memory column: 8 slots on fused launches, one block (2)
    # on streaming launches; the stream consumed must not depend on it
    c = H.HostSim(models, conns, n=12, seed=5, instrument=False, stat_model_id=stat)
    assert c.use_shape(shape)
    c.set_rng_slots(2)
    c.set_reload_every(1)
    a.step_until(200.0)
    b.step_until(200.0)
    c.step_until(200.0)
    assert (a.steps() == b.steps()).all() and (a.events() == b.events()).all() and (a.draws() == b.draws()).all()
    assert (a.time() == b.time()).all() and (b.errors() == 0).all()
    assert (c.steps() == b.steps()).all() and (c.draws() == b.draws()).all() and (c.time() == b.time()).all()
    for m in range(len(models)):
        assert np.array_equal(a.until(m), b.until(m), equal_nan=True)
    for r in (0, 11):
        e, o = _oracle(models, conns, 5, r, until=200.0)
        c = o.counters()
        assert (c["steps"], c["events"], o.rng_draws(), o.get_global_time()) == \
               (int(b.steps()[r]), int(b.events()[r]), int(b.draws()[r]), b.time()[r])


def test_shape_match_is_structural():
    models, conns = networks.mm1(lam_gen=0.9, lam_proc=2.0)       # other rates: same shape
    assert H.HostSim(models, conns, n=1, instrument=False, stat_model_id="processor-01").use_shape(0)
    models, conns = networks.mm1(capacity=9)                       # other capacity: another structure
    assert not H.HostSim(models, conns, n=1, instrument=False, stat_model_id="processor-01").use_shape(0)
    models, conns = networks.mm1()
    assert not H.HostSim(models, conns, n=1, instrument=False).use_shape(0)  # no statistic rows: another layout


def test_baked_shapes_are_up_to_date():
    """sim_b200/csrc/shapes.inc must be what tools/gen_shapes.py produces from sim_b200/shapes/*.json."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "sim_b200", "csrc", "shapes.inc")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_shapes.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == before


def _batch_means_from_waits(waits, warmup, size, max_batches=30):
    """streaming batch means exactly as the device accumulates them (sim_b200/csrc/core.cuh wait_sample)"""
    means = []
    w = waits[warmup:]
    for b in range(min(len(w) // size, max_batches)):
        acc = 0.0
        for x in w[b * size:(b + 1) * size]:
            acc += x
        means.append(acc / size)
    return means


def _oracle_waits(models, conns, seed, rep, until, model_id):
    o = O.OracleSimulation(models, conns)
    o.set_rng_philox(seed, rep)
    assert o.step_until(until, collect=False)[0] == 0
    arrival, waits = {}, []
    for t, action, subject in o.get_records(model_id)[1]:
        if action == "Arrival":
            arrival[subject] = t
        elif action == "Processing Start":
            waits.append(t - arrival.pop(subject))
    return waits


def test_streaming_batch_means_match_the_oracle_series():
    """BASELINE config 3 statistic: per-replication batch means with a fixed warm-up and batch size,
    accumulated on the fly, equal the batch means of the oracle's waiting-time series."""
    models, conns = networks.tandem()
    rec = networks.tandem()[0]
    for m in rec:
        m.inner.store_records = True
    warm, size = 40, 25
    hs = H.HostSim(models, conns, n=4, seed=13, stat_model_id="processor-3", batch_warmup=warm, batch_size=size)
    hs.set_reload_every(7)
    hs.step_until(1500.0)
    sm, sq, open_sum = hs.batch()
    _, cnt = hs.wait()
    for r in range(4):
        waits = _oracle_waits(rec, conns, 13, r, 1500.0, "processor-3")
        assert len(waits) == cnt[r]
        means = _batch_means_from_waits(waits, warm, size)
        assert len(means) == 30  # the horizon is long enough to fill Schmeiser's 30 batches
        assert abs(sm[r] - sum(means)) <= 1e-9 * abs(sum(means))
        assert abs(sq[r] - sum(x * x for x in means)) <= 1e-9 * sum(x * x for x in means)
    short = H.HostSim(models, conns, n=2, seed=13, stat_model_id="processor-3", batch_warmup=warm, batch_size=size)
    short.step_until(60.0)  # fewer than warm-up + one batch
    sm, sq, open_sum = short.batch()
    assert (sm == 0).all() and (sq == 0).all()


def test_models_sharing_an_id_all_receive_the_message():
    """mod.rs:203-221: `step` filters the messages by id for EVERY model in turn, so two posted models with one id both
    run events_ext on a message addressed to it -- over a connector and through inject_input -- while the reference
    still holds ONE Message (one trace entry, one count on the connector)."""
    models, conns = networks.shared_ids()
    hs = H.HostSim(models, conns, n=2, seed=9, trace_reps=2, trace_cap=20000)
    assert hs.inject_input("twin", "job", "manual 1") == 0
    hs.step_n(600)
    for r in range(2):
        o = O.OracleSimulation(models, conns)
        o.set_rng_philox(9, r)
        o.trace_enable(True)
        o.trace_intern("manual 1")
        o.inject_input("twin", "job", "manual 1")
        e, _ = o.step_n(600, collect=False)
        assert e == 0 and hs.errors()[r] == 0
        assert o.trace() == hs.trace(r) and o.trace_hash() == int(hs.hash()[r])
        assert o.counters()["events"] == int(hs.events()[r])
    sent = int(hs.conn_counts()[0])                      # Messages on c-in, both replications
    rec = hs.record_counts()
    # every job sent over c-in, plus the injected one of each replication, reached BOTH twins (Arrival or Drop; the
    # jobs of the last step are still in the mailbox)
    got = [int(rec[m][2] + rec[m][4]) for m in (1, 3)]
    in_flight = sum(1 for r in range(2) for p in hs.pending(r) if p[1] == 1)
    assert got[0] == got[1] == sent + 2 - in_flight and sent > 100
