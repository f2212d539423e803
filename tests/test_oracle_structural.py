"""The reference's own integration tests (sim/tests/simulations.rs, web.rs, custom.rs), replayed on
the CPU oracle with the reference's default stream Pcg64Mcg::new(42).  The structural assertions
hold for any stream, so they pin the oracle's engine/model semantics independently of sampler
fidelity; the statistical ones are the reference's own brackets."""
import numpy as np

import networks
import oracle_api as O


def _count(msgs, **kw):
    return sum(1 for m in msgs if all(m[k] == v for k, v in kw.items()))


def _num(content):
    return content.split()[-1]


def test_poisson_generator_processor_with_capacity():
    # simulations.rs:20-129 (BASELINE config 1 shape)
    sim = O.OracleSimulation(*networks.mm1())
    e, msgs = sim.step_n(3000)
    assert e == 0
    departures = [(m["time"], m["content"]) for m in msgs if m["target_id"] == "storage-01"]
    arrivals = [(m["time"], m["content"]) for m in msgs if m["target_id"] == "processor-01"]
    arr_by_num = {}
    for t, c in arrivals:
        arr_by_num.setdefault(_num(c), t)
    response = [t - arr_by_num[_num(c)] for t, c in departures]
    e, ci = O.steady_state(response, 0.001)
    expected = (172285188.0 / 14316139.0) / (4766600.0 / 14316169.0)
    assert e == 0 and ci["lower"] < expected < ci["upper"]
    last = _num(departures[-1][1])
    generated = [_num(c) for _, c in arrivals].index(last) + 1
    rate = 0.5 * len(departures) / generated
    assert abs(rate - 4766600.0 / 14316169.0) / (4766600.0 / 14316169.0) < 0.34


def test_step_until_activities():
    # simulations.rs:132-178: reset + put per replication, one shared stream, CI brackets 50
    sim = O.OracleSimulation(*networks.generator_storage())
    counts = []
    for _ in range(10):
        sim.reset()
        sim.put()
        e, msgs = sim.step_until(100.0)
        assert e == 0 and sim.get_global_time() >= 100.0
        assert all(m["time"] < 100.0 for m in msgs)  # the crossing step is executed but not returned
        counts.append(float(len(msgs)))
    e, ci = O.independent_sample(counts, 0.001)
    assert ci["lower"] < 50.0 < ci["upper"]


def _arrival_counts_oracle(reps=10, until=480.0):
    models, conns = networks.mm1(lam_gen=0.0957, lam_proc=0.1659)
    sim = O.OracleSimulation(models, conns)
    counts = []
    for _ in range(reps):
        sim.reset()
        sim.put()
        e, msgs = sim.step_until(until)
        assert e == 0
        counts.append(float(sum(1 for m in msgs if m["target_id"] == "processor-01")))
    return counts


def test_non_stationary_generation():
    # simulations.rs:181-255 (no thinning despite the name): arrivals in 480 time units over 10 replications on
    # one shared stream; the CI must overlap the empirical one the reference hard-codes
    counts = _arrival_counts_oracle()
    e, ci = O.independent_sample(counts, 0.05)
    e2, emp = O.independent_sample([47.0, 42.0, 45.0, 34.0, 37.0], 0.05)
    assert e == 0 and e2 == 0
    assert ci["lower"] < emp["upper"] and ci["upper"] > emp["lower"]


def test_exclusive_gateway_601_steps_200_jobs():
    # simulations.rs:258-379
    sim = O.OracleSimulation(*networks.exclusive_gateway())
    msgs = []
    for _ in range(601):
        e, m = sim.step()
        assert e == 0
        msgs += m
    outs = [_count(msgs, target_id="storage-0%d" % i) for i in (1, 2, 3)]
    assert sum(outs) == 200
    chi2 = sum((o - x) ** 2 / x for o, x in zip(outs, (120, 60, 20)))
    assert chi2 < 9.21


def test_gate_blocking_proportions():
    # simulations.rs:382-496
    sim = O.OracleSimulation(*networks.gate_blocking())
    passed = []
    for _ in range(10):
        sim.reset()
        sim.put()
        msgs = []
        for _ in range(1000):
            e, m = sim.step()
            assert e == 0
            msgs += m
        arrivals = _count(msgs, source_id="generator-03", target_id="gate-01")
        departures = _count(msgs, target_id="storage-01")
        if arrivals:
            passed.append(departures / arrivals)
    e, ci = O.independent_sample(passed, 0.01)
    assert ci["lower"] < 0.5 < ci["upper"]


def test_load_balancer_round_robin_outputs():
    # simulations.rs:499-605: 28 steps = 1 + 9 x 3; first job leaves on flow_paths[1]
    sim = O.OracleSimulation(*networks.load_balancer())
    e, msgs = sim.step_n(28)
    assert e == 0
    assert [_count(msgs, target_id="storage-0%d" % i) for i in (1, 2, 3)] == [3, 3, 3]
    order = [m["target_id"] for m in msgs if m["target_id"].startswith("storage")]
    assert order[:3] == ["storage-02", "storage-03", "storage-01"]


def test_injection_initiated_stored_value_exchange():
    # simulations.rs:608-678
    sim = O.OracleSimulation(*networks.storage_exchange())
    sim.inject_input("storage-01", "store", "42")
    assert sim.step()[0] == 0
    sim.inject_input("storage-01", "read", "")
    e, m = sim.step()
    assert [x["content"] for x in m] == ["42"] and m[0]["target_id"] == "storage-02"
    sim.inject_input("storage-02", "read", "")  # appended AFTER the routed message (mod.rs:190)
    e, m = sim.step()
    assert m[0]["content"] == "42"
    assert sim.get_global_time() == 0.0


def test_parallel_gateway_splits_and_joins():
    # simulations.rs:681-787
    sim = O.OracleSimulation(*networks.parallel_gateway())
    e, msgs = sim.step_n(101)
    a, b, d, s = (_count(msgs, target_port=p) for p in ("alpha", "beta", "delta", "store"))
    assert a == b == d == s and s > 0


def test_match_status_reporting():
    # simulations.rs:790-823
    sim = O.OracleSimulation(*networks.status_pair())
    assert sim.get_status("generator-01") == (0, "Generating jobs")
    assert sim.get_status("load-balancer-01") == (0, "Listening for requests")
    assert sim.get_status("nope")[0] == 2  # ModelNotFound (mod.rs:111)
    assert sim.get_records("nope")[0] == 2


def test_stochastic_gate_blocking():
    # simulations.rs:826-893
    sim = O.OracleSimulation(*networks.stochastic_gate())
    e, msgs = sim.step_n(101)
    results = [1.0] * _count(msgs, target_id="storage-01")
    passes = len(results)
    results += [0.0 for i, m in enumerate(msgs) if m["target_id"] == "stochastic-gate-01" and i > passes]
    e, ci = O.independent_sample(results, 0.01)
    assert ci["lower"] < 0.2 < ci["upper"]


def test_batch_sizing():
    # simulations.rs:896-964
    sim = O.OracleSimulation(*networks.batcher())
    sizes = []
    for _ in range(10000):
        e, m = sim.step()
        assert e == 0
        sizes.append(_count(m, target_id="storage-01"))
    assert any(0 < s < 10 for s in sizes) and any(s == 10 for s in sizes) and not any(s > 10 for s in sizes)


def test_min_and_max_stopwatch():
    # simulations.rs:967-1104
    sim = O.OracleSimulation(*networks.stopwatch())
    assert sim.step_n(12)[0] == 0
    sim.inject_input("stopwatch-01", "min", "42")
    sim.inject_input("stopwatch-02", "max", "42")
    e, responses = sim.step_n(2)
    assert e == 0 and len(responses) >= 2
    assert responses[0]["content"] != responses[1]["content"]


def test_processor_from_queue_cadence():
    # web.rs:13-126: odd steps emit nothing, step 2k emits exactly "job k"; mean service time ~ 1/3
    sim = O.OracleSimulation(*networks.processor_from_queue())
    times = []
    for step in range(1, 201):
        e, m = sim.step()
        assert e == 0
        if step % 2:
            assert m == []
        else:
            assert m[0]["content"] == "job %d" % (step // 2)
            times.append(m[0]["time"])
    assert sim.rng_draws() >= 100
    gaps = np.diff([0.0] + times)
    e, ci = O.independent_sample(gaps, 0.01)
    assert ci["lower"] < 1.0 / 3.0 < ci["upper"]


def test_processor_network_no_job_loss():
    # web.rs:130-317: fan-out of one out-port over two connectors; 720 steps -> 30 storage arrivals
    sim = O.OracleSimulation(*networks.processor_network())
    msgs = []
    for _ in range(720):
        e, m = sim.step()
        assert e == 0
        msgs += m
    assert _count(msgs, target_id="storage-0") == 30


def test_step_n_with_custom_passive_model():
    # custom.rs:89-120
    sim = O.OracleSimulation(*networks.generator_passive())
    e, msgs = sim.step_n(9)
    assert e == 0 and len(msgs) == 4


def test_mm1_first_steps_records():
    # SURVEY.md Appendix G7, hand-traced from the sources
    sim = O.OracleSimulation(*networks.mm1(store_records=True))
    sim.step()
    assert sim.get_global_time() == 0.0 and sim.get_records("generator-01")[1][0][1] == "Initialization"
    d1 = sim.until_next_event("generator-01")
    e, m = sim.step()
    assert sim.get_global_time() == d1 and [x["content"] for x in m] == ["job 1"]
    e, m = sim.step()
    assert m == [] and sim.get_global_time() == d1
    recs = sim.get_records("processor-01")[1]
    assert [(r[1], r[2]) for r in recs] == [("Arrival", "job 1"), ("Processing Start", "job 1")]
    assert sim.rng_draws() >= 3


def test_error_paths():
    sim = O.OracleSimulation(*networks.mm1())
    sim.inject_input("processor-01", "bogus", "x")
    assert sim.step()[0] == 7      # InvalidMessage (processor.rs:225)
    sim = O.OracleSimulation(*networks.mm1(capacity=0))
    assert sim.step_n(3)[0] == 5   # InvalidModelState: capacity 0 (processor.rs:221)
    sim = O.OracleSimulation(*networks.mm1(lam_gen=0.0))
    assert sim.step()[0] == 15     # ExpError at the first draw
    sim = O.OracleSimulation(*networks.storage_exchange())
    sim.inject_input("storage-01", "bogus", "x")
    assert sim.step()[0] == 7      # storage.rs:159


def test_all_passive_network_goes_to_infinity():
    # mod.rs:225-236 with every model at +INF and no messages: dt = +INF, U = INF - INF = NaN
    sim = O.OracleSimulation(*networks.storage_exchange())
    assert sim.step()[0] == 0
    assert sim.get_global_time() == float("inf")
    assert np.isnan(sim.until_next_event("storage-01"))


def test_gateway_network_join_fires_once_per_departure():
    """BASELINE config 4 as benchmarked: the ParallelGateway JOIN upstream of the batcher completes
    one collection per proc-p departure (parallel_gateway.rs:100-143), and the batcher releases in tens."""
    models, conns = networks.gateway_network()
    o = O.OracleSimulation(models, conns)
    o.set_rng_philox(42, 0)
    e, _ = o.step_n(20000, collect=False)
    assert e == 0
    per = dict(zip([c.id for c in conns], o.connector_messages()))
    assert per["c10"] > 100 and per["c11"] > 100
    assert per["c10"] == per["c03"]            # every job of the split reaches the join early ...
    assert per["c11"] == per["c06"]            # ... and late, when proc-p releases it
    assert per["c12"] == per["c11"]            # one full collection per departure -> one job to the batcher
    assert 0 < per["c21"] <= per["c12"]        # released by the batcher in tens (or on its timer)
