"""GPU parity tests: the CUDA engine (through the C ABI, via sim_b200.Simulation) against the CPU
oracle on identical per-replication uniform streams.  Discrete results (event order, message and
record counts, step / event / draw counts) must match bit-exactly -- they are compared through the
per-replication trace hash and, for traced replications, event by event -- and f64 clocks within
1e-12 relative (north_star)."""
import os

import numpy as np
import pytest

import networks
import oracle_api
from sim_b200 import ContinuousRandomVariable, Message

Exp = ContinuousRandomVariable.Exp

pytestmark = pytest.mark.gpu

REL = 1e-12


def _sim(models, conns, n, **kw):
    from sim_b200 import Simulation
    kw.setdefault("instrument", True)
    return Simulation.post(models, conns, replications=n, **kw)


def _close(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | both_inf | (np.abs(a - b) <= REL * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))
    return bool(ok.all())


@pytest.mark.parametrize("name", sorted(networks.ALL))
@pytest.mark.parametrize("k", [1, 7])
def test_step_n_matches_oracle(name, k):
    models, conns = networks.ALL[name]()
    n, steps = 96, 1500
    sim = _sim(models, conns, n, seed=42, steps_per_launch=k)
    sim.step_n(steps)
    ref = oracle_api.OracleSimulation(models, conns).run_batch(42, 0, n, threads=8, nsteps=steps)
    assert ref["err"] == 0
    assert (sim.get_errors() == 0).all()
    assert (sim.get_trace_hash() == ref["hash"]).all()
    assert (sim.get_steps() == ref["steps"]).all()
    assert (sim.get_events() == ref["events"]).all()
    assert _close(sim.get_global_time(), ref["time"])
    c = sim.get_counters()
    assert c["steps"] == ref["totals"]["steps"]
    assert c["events"] == ref["totals"]["ext"] + ref["totals"]["int"]


def test_trace_event_by_event_mm1():
    models, conns = networks.mm1(store_records=True)
    sim = _sim(models, conns, 8, seed=7, trace_replications=2, trace_capacity=20000, steps_per_launch=5)
    sim.step_n(3000)
    for r in range(2):
        o = oracle_api.OracleSimulation(models, conns)
        o.set_rng_philox(7, r)
        o.trace_enable(True)
        e, _ = o.step_n(3000, collect=False)
        assert e == 0
        a, b = o.trace(), sim.get_trace(r)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert {k: x[k] for k in ("step", "kind", "index", "action", "aux", "content")} == \
                   {k: y[k] for k in ("step", "kind", "index", "action", "aux", "content")}
            assert _close(x["time"], y["time"])
        # get_records parity (time, action, subject) for every model
        for m in models:
            e, recs = o.get_records(m.id)
            got = sim.get_records(m.id, r)
            assert [(x[1], x[2]) for x in recs] == [(x[1], x[2]) for x in got]
            assert _close([x[0] for x in recs], [x[0] for x in got])


def test_step_until_and_wait_statistic():
    """BASELINE config 1/2 shape: step_until + IID mean waiting time over replications."""
    n, until = 512, 400.0
    models, conns = networks.mm1()
    sim = _sim(models, conns, n, seed=42, stat_model_id="processor-01", steps_per_launch=16)
    sim.step_until(until)
    rec_models, _ = networks.mm1(store_records=True)  # the oracle derives waits from records (web.rs:571-597)
    ref = oracle_api.OracleSimulation(rec_models, conns).run_batch(42, 0, n, threads=8, until=until,
                                                                   stat_model_id="processor-01")
    assert (sim.get_trace_hash() == ref["hash"]).all()
    assert (sim.get_steps() == ref["steps"]).all()
    t = sim.get_global_time()
    assert (t >= until).all() and _close(t, ref["time"])
    s, cnt = sim.get_wait_stats()
    assert (cnt == ref["wait_n"]).all()
    assert _close(s, ref["wait_sum"])
    means = ref["wait_sum"][ref["wait_n"] > 0] / ref["wait_n"][ref["wait_n"] > 0]
    e, want = oracle_api.independent_sample(means, 0.05)
    got = sim.independent_sample_ci(0.05)
    assert got["n"] == means.size
    for key in ("mean", "variance", "lower", "upper"):
        assert abs(got[key] - want[key]) <= 1e-10 * max(1.0, abs(want[key])), key
    # a second step_until continues every replication (at least one more step each, mod.rs:279-286)
    before = sim.get_steps().copy()
    sim.step_until(until)
    assert (sim.get_steps() == before + 1).all()


def test_reference_default_stream_pcg():
    """Replication 0 under SIMB_RNG_PCG64MCG seed 42 consumes the reference's default stream
    Pcg64Mcg::new(42) (dynamic_rng.rs:7-9)."""
    from sim_b200.simulator import RNG_PCG64MCG
    models, conns = networks.mm1()
    sim = _sim(models, conns, 4, seed=42, rng_kind=RNG_PCG64MCG, steps_per_launch=3)
    sim.step_n(3000)
    for r in range(4):
        o = oracle_api.OracleSimulation(models, conns)
        o.set_rng_pcg(42 + r)
        o.trace_enable(False)
        o.step_n(3000, collect=False)
        assert o.trace_hash() == int(sim.get_trace_hash()[r])
        assert _close(o.get_global_time(), sim.get_global_time()[r])


def test_external_uniform_stream():
    """Fed an explicit u64 draw stream, the engine consumes it exactly like the oracle does."""
    from sim_b200.simulator import RNG_EXTERNAL
    models, conns = networks.tandem()
    n, per_rep = 32, 4000
    draws = np.stack([oracle_api.pcg_fill(1000 + r, per_rep) for r in range(n)], axis=1)
    sim = _sim(models, conns, n, rng_kind=RNG_EXTERNAL, steps_per_launch=4)
    sim.set_rng_stream(draws)
    sim.step_n(1200)
    assert (sim.get_errors() == 0).all()
    used = sim.get_draws()
    for r in range(0, n, 5):
        o = oracle_api.OracleSimulation(models, conns)
        o.set_rng_external(draws[:, r])
        o.trace_enable(False)
        e, _ = o.step_n(1200, collect=False)
        assert e == 0
        assert o.trace_hash() == int(sim.get_trace_hash()[r])
        assert o.rng_draws() == int(used[r])


def test_structural_fixtures_on_device():
    # load_balancer_round_robin_outputs (simulations.rs:586-603): 28 steps -> 3 arrivals per storage
    models, conns = networks.load_balancer()
    sim = _sim(models, conns, 33, seed=1)
    sim.step_n(28)
    assert (sim.get_message_counts() == np.array([9, 3, 3, 3]) * 33).all()
    # exclusive gateway (simulations.rs:348-374): 601 steps -> exactly 200 routed jobs
    models, conns = networks.exclusive_gateway()
    sim = _sim(models, conns, 10, seed=1)
    sim.step_n(601)
    assert int(sim.get_message_counts()[1:].sum()) == 200 * 10
    # processor network (web.rs:303-316): 720 steps -> 30 storage arrivals
    models, conns = networks.processor_network()
    sim = _sim(models, conns, 5, seed=3)
    sim.step_n(720)
    mc = sim.get_message_counts()
    assert int(mc[2] + mc[7]) == 30 * 5
    # custom passive (custom.rs:113-118): 9 steps -> 4 messages
    models, conns = networks.generator_passive()
    sim = _sim(models, conns, 3, seed=9)
    sim.step_n(9)
    assert int(sim.get_message_counts()[0]) == 4 * 3


def test_injection_exchange_on_device():
    """injection_initiated_stored_value_exchange (simulations.rs:608-678)."""
    models, conns = networks.storage_exchange()
    sim = _sim(models, conns, 5, max_pending=4)
    sim.inject_input(Message.new("manual", "manual", "storage-01", "store", 0.0, "42"))
    sim.step()
    sim.inject_input(Message.new("manual", "manual", "storage-01", "read", 0.0, ""))
    sim.step()
    sim.inject_input(Message.new("manual", "manual", "storage-02", "read", 0.0, ""))
    sim.step()
    for r in (0, 4):
        msgs = sim.get_messages(r)
        assert msgs[0].content == "42"
        assert msgs[0].target_id == "storage-01"
    assert (sim.get_global_time() == 0.0).all()
    assert sim.get_status("storage-02", 0) == "Storing 42"


def test_processor_from_queue_cadence():
    """processor_from_queue_response_time_is_correct (web.rs:71-95): message exactly on even steps."""
    models, conns = networks.processor_from_queue()
    sim = _sim(models, conns, 4, seed=5)
    for step in range(1, 41):
        sim.step()
        msgs = sim.get_messages(1)
        if step % 2 == 0:
            assert len(msgs) == 1 and msgs[0].content == "job %d" % (step // 2)
        else:
            assert msgs == []


def test_errors_are_sticky_and_named():
    from sim_b200 import SimulationError
    models, conns = networks.mm1()
    sim = _sim(models, conns, 6)
    mask = np.array([0, 1, 0, 0, 1, 0], np.uint8)
    sim.inject_input(Message.new("manual", "manual", "processor-01", "no-such-port", 0.0, "x"), replica_mask=mask)
    with pytest.raises(SimulationError) as ei:
        sim.step_n(10)
    assert ei.value.name == "InvalidMessage"  # processor.rs:225
    assert (sim.get_errors() == np.array([0, 7, 0, 0, 7, 0])).all()
    steps = sim.get_steps()
    assert steps[1] == 0 and steps[0] == 10


def test_status_strings():
    models, conns = networks.status_pair()
    sim = _sim(models, conns, 2)
    assert sim.get_status("generator-01") == "Generating jobs"
    assert sim.get_status("load-balancer-01") == "Listening for requests"


@pytest.mark.parametrize("name,stat,shape", [("mm1", "processor-01", "mm1"), ("mm1", None, "mm1_plain"), ("tandem", None, "generic"),
                                             ("gateway_network", None, "generic"), ("large_mixed", None, "generic")])
@pytest.mark.parametrize("k", [1, 16])
def test_production_kernels_match_oracle(name, stat, shape, k):
    """The production configuration (Philox, not instrumented): M/M/1 runs on the step kernel
    specialised for its baked shape (sim_b200/csrc/shapes.inc), the other BASELINE networks on the
    generic interpreter.  Both must agree with the oracle in everything observable without
    instrumentation: steps, events, draws, clocks, every model's until_next_event, the waiting-time
    accumulators -- and with the instrumented kernel's trace hash taken on the same replications."""
    models, conns = networks.ALL[name]()
    n, until = 300, 150.0
    sim = _sim(models, conns, n, seed=5, instrument=False, stat_model_id=stat, steps_per_launch=k)
    assert sim.describe()["kernel_shape"] == shape
    sim.step_until(until)
    ref_models = networks.mm1(store_records=True)[0] if name == "mm1" else models
    ref = oracle_api.OracleSimulation(ref_models, conns).run_batch(5, 0, n, threads=8, until=until, stat_model_id=stat)
    assert ref["err"] == 0 and (sim.get_errors() == 0).all()
    assert (sim.get_steps() == ref["steps"]).all()
    assert (sim.get_events() == ref["events"]).all()
    assert _close(sim.get_global_time(), ref["time"])
    if stat:
        s, cnt = sim.get_wait_stats()
        assert (cnt == ref["wait_n"]).all() and _close(s, ref["wait_sum"])
    # the instrumented handle of a baked shape runs the SAME specialised code path with the trace compiled in
    gen = _sim(models, conns, n, seed=5, instrument=True, stat_model_id=stat, steps_per_launch=k)
    assert gen.describe()["kernel_shape"] == (shape + "_instr" if shape != "generic" else "generic")
    gen.step_until(until)
    assert (gen.get_trace_hash() == ref["hash"]).all()
    if shape != "generic":  # ... and agrees with the generic interpreter's trace, message and record counts
        os.environ["SIMB_NO_SHAPES"] = "1"
        try:
            itp = _sim(models, conns, n, seed=5, instrument=True, stat_model_id=stat, steps_per_launch=k)
        finally:
            del os.environ["SIMB_NO_SHAPES"]
        assert itp.describe()["kernel_shape"] == "generic"
        itp.step_until(until)
        assert (itp.get_trace_hash() == gen.get_trace_hash()).all()
        assert (itp.get_message_counts() == gen.get_message_counts()).all()
        assert (itp.get_record_counts() == gen.get_record_counts()).all()
    assert (gen.get_draws() == sim.get_draws()).all()
    for m in models:
        assert _close(gen.get_until_next_event(m.id), sim.get_until_next_event(m.id)), m.id
    # a second call continues both the same way
    sim.step_n(77)
    gen.step_n(77)
    assert (gen.get_steps() == sim.get_steps()).all() and _close(gen.get_global_time(), sim.get_global_time())


def test_shape_match_ignores_distribution_values():
    """A baked shape is the network STRUCTURE; rates are launch parameters."""
    models, conns = networks.mm1(lam_gen=0.7, lam_proc=1.1)
    sim = _sim(models, conns, 256, seed=2, instrument=False, stat_model_id="processor-01")
    assert sim.describe()["kernel_shape"] == "mm1"
    sim.step_n(500)
    ref = oracle_api.OracleSimulation(models, conns).run_batch(2, 0, 256, threads=4, nsteps=500)
    assert (sim.get_events() == ref["events"]).all() and _close(sim.get_global_time(), ref["time"])
    other, conns2 = networks.mm1(capacity=9)
    assert _sim(other, conns2, 256, instrument=False, stat_model_id="processor-01").describe()["kernel_shape"] == "generic"
    # baked-shape kernels assume the full 256-wide tile: small batches run on the generic interpreter
    small = _sim(*networks.mm1(), 64, seed=2, instrument=False, stat_model_id="processor-01")
    assert small.describe()["kernel_shape"] == "generic"
    small.step_n(500)
    ref = oracle_api.OracleSimulation(*networks.mm1()).run_batch(2, 0, 64, threads=4, nsteps=500)
    assert (small.get_events() == ref["events"]).all() and _close(small.get_global_time(), ref["time"])


# ---- BASELINE.json full size: 2^20 replications x step_until(10 000) ---------------------------------------
FULL_N = 1 << 20
FULL_UNTIL = 10000.0


@pytest.fixture(scope="module")
def full_mm1():
    models, conns = networks.mm1()
    sim = _sim(models, conns, FULL_N, seed=42, instrument=False, stat_model_id="processor-01")
    assert sim.describe()["kernel_shape"] == "mm1"
    sim.step_until(FULL_UNTIL)
    return sim


def test_full_size_properties(full_mm1):
    sim = full_mm1
    t = sim.get_global_time()
    steps, events, draws = sim.get_steps(), sim.get_events(), sim.get_draws()
    s, cnt = sim.get_wait_stats()
    assert (sim.get_errors() == 0).all()
    assert np.isfinite(t).all() and (t >= FULL_UNTIL).all() and (t < FULL_UNTIL + 60.0).all()
    # M/M/1 cadence: every step runs at least one event, at most three
    assert (events >= steps).all() and (events <= 3 * steps.astype(np.uint64)).all()
    # one draw per generator firing / service start, plus rare ziggurat retries: draws ~ events / 2.4
    assert (draws > events // 3).all() and (draws < events).all()
    assert (cnt > 0).all() and np.isfinite(s).all() and (s >= 0).all()
    # replication-to-replication variability is real (not one replication copied 2^20 times)
    assert np.unique(steps).size > 200 and np.unique(t).size > FULL_N * 0.999
    # expected number of arrivals 5000 +- a few hundred; steps ~ 3.33 per arrival
    assert 15000 < steps.mean() < 18500
    c = sim.get_counters()
    assert c["steps"] == int(steps.astype(np.uint64).sum()) and c["events"] == int(events.astype(np.uint64).sum())


def test_full_size_sampled_replications_match_oracle(full_mm1):
    """Replications picked across the whole 2^20 batch, each compared with the oracle over the full horizon."""
    sim = full_mm1
    t, steps, events = sim.get_global_time(), sim.get_steps(), sim.get_events()
    s, cnt = sim.get_wait_stats()
    rec_models, conns = networks.mm1(store_records=True)
    o = oracle_api.OracleSimulation(rec_models, conns)
    for r0 in (0, 255, 256, 65535, 300001, 524288, 777777, FULL_N - 4):
        ref = o.run_batch(42, r0, 4, threads=4, until=FULL_UNTIL, stat_model_id="processor-01", trace_hash=False)
        sl = slice(r0, r0 + 4)
        assert (steps[sl] == ref["steps"]).all() and (events[sl] == ref["events"]).all()
        assert (cnt[sl] == ref["wait_n"]).all()
        assert _close(t[sl], ref["time"]) and _close(s[sl], ref["wait_sum"])


@pytest.mark.parametrize("stat", ["processor-01", None])
def test_full_size_baked_kernel_trace(full_mm1, stat):
    """north_star test 1 for the kernel that produces the headline number: the baked M/M/1 code path with the trace
    compiled in (shapes mm1_instr / mm1_plain_instr: the same template instantiation plus record<> / emit<>
    instrumentation) at 2^20 replications over the full horizon.  Event order (per-replication FNV trace hash),
    per-connector message counts and per-(model, action) record counts against the oracle on replications sampled
    across the batch; totals against the identities of the M/M/1 cadence; and the production (uninstrumented)
    kernel agrees with it in everything it can observe on all 2^20 replications."""
    models, conns = networks.mm1()
    sim = _sim(models, conns, FULL_N, seed=42, instrument=True, stat_model_id=stat, trace_replications=0)
    assert sim.describe()["kernel_shape"] == ("mm1_instr" if stat else "mm1_plain_instr")
    sim.step_until(FULL_UNTIL)
    assert (sim.get_errors() == 0).all()
    h, steps, events = sim.get_trace_hash(), sim.get_steps(), sim.get_events()
    rec_models, _ = networks.mm1(store_records=True)
    o = oracle_api.OracleSimulation(rec_models, conns)
    for r0 in (0, 255, 256, 65535, 300001, 524288, 777777, FULL_N - 4):
        ref = o.run_batch(42, r0, 4, threads=4, until=FULL_UNTIL, stat_model_id=stat)
        sl = slice(r0, r0 + 4)
        assert (h[sl] == ref["hash"]).all() and (steps[sl] == ref["steps"]).all() and (events[sl] == ref["events"]).all()
    # message / record totals over the batch: exact identities of this network (generator.rs:98-123,
    # processor.rs:119-196, storage.rs:106-113)
    mc = sim.get_message_counts()
    rc = sim.get_record_counts().reshape(len(models), -1)
    A = {n: i for i, n in enumerate(("Initialization", "Generation", "Arrival", "Processing Start", "Drop", "Departure"))}
    assert rc[0, A["Initialization"]] == FULL_N and rc[0, A["Generation"]] == mc[0]        # one message per generated job
    # every job arrives or is dropped, processed jobs reach the storage -- up to the one message a replication
    # may hold in flight when it crosses the horizon
    assert 0 <= int(mc[0]) - int(rc[1, A["Arrival"]] + rc[1, A["Drop"]]) <= FULL_N
    assert rc[1, A["Departure"]] == mc[1] and 0 <= int(mc[1]) - int(rc[2, A["Arrival"]]) <= FULL_N
    assert 0 <= rc[1, A["Processing Start"]] - rc[1, A["Departure"]] <= FULL_N             # at most one job in service each
    # the oracle's totals on a contiguous block agree with a second instrumented handle posted on that block
    blk = _sim(models, conns, 512, seed=42, instrument=True, stat_model_id=stat, replication_offset=4096)
    blk.step_until(FULL_UNTIL)
    ref = o.run_batch(42, 4096, 512, threads=8, until=FULL_UNTIL, stat_model_id=stat)
    assert (blk.get_trace_hash() == ref["hash"]).all() and (blk.get_trace_hash() == h[4096:4608]).all()
    assert (blk.get_message_counts() == np.array(ref["conn"], np.uint64)).all()
    assert (np.asarray(blk.get_record_counts()).reshape(-1) == np.array(ref["records"], np.uint64).reshape(-1)).all()
    if stat:  # the production kernel of the headline bench (fixture) saw the same replications
        assert np.array_equal(full_mm1.get_steps(), steps) and np.array_equal(full_mm1.get_events(), events)
        assert np.array_equal(full_mm1.get_global_time(), sim.get_global_time())
        assert np.array_equal(full_mm1.get_draws(), sim.get_draws())


def test_full_size_statistic_and_sharding_invariance(full_mm1):
    """IID mean waiting time CI over 2^20 replications (BASELINE config 2) equals IndependentSample on the
    host, brackets the stationary M/M/1/14 waiting time up to the warm-up bias, and a shard posted with a
    replication offset reproduces the same replications bit for bit."""
    sim = full_mm1
    s, cnt = sim.get_wait_stats()
    means = s / cnt
    from sim_b200 import output_analysis as oa
    host = oa.IndependentSample.post(means)
    ci = sim.independent_sample_ci(0.05)
    hci = host.confidence_interval_mean(0.05)
    assert ci["n"] == FULL_N
    assert abs(ci["mean"] - host.point_estimate_mean()) <= 1e-10 * ci["mean"]
    assert abs(ci["lower"] - hci.lower()) <= 1e-9 and abs(ci["upper"] - hci.upper()) <= 1e-9
    # stationary M/M/1/14 (lambda .5, mu .333333): W = L / lambda_eff - 1/mu, closed form of simulations.rs:105-108
    w_q = (172285188.0 / 14316139.0) / (4766600.0 / 14316169.0) - 1.0 / 0.333333
    assert abs(ci["mean"] - w_q) / w_q < 0.02
    models, conns = networks.mm1()
    shard = _sim(models, conns, 512, seed=42, instrument=False, stat_model_id="processor-01", replication_offset=700000)
    shard.step_until(FULL_UNTIL)
    assert np.array_equal(shard.get_global_time(), sim.get_global_time()[700000:700512])
    assert np.array_equal(shard.get_steps(), sim.get_steps()[700000:700512])
    s2, c2 = shard.get_wait_stats()
    assert np.array_equal(s2, s[700000:700512]) and np.array_equal(c2, cnt[700000:700512])


def test_steady_state_output_on_device_records():
    """BASELINE config 3 statistic: per-replication response-time series of the last tandem stage from the
    captured device records -> SteadyStateOutput (MSER-style deletion, <= 30 batch means).  The series and the
    interval must equal what the oracle's records give."""
    models, conns = networks.tandem()
    for m in models:
        m.inner.store_records = True
    sim = _sim(models, conns, 16, seed=8, instrument=True, trace_replications=4, trace_capacity=200000, steps_per_launch=32)
    sim.step_until(800.0)
    from sim_b200 import output_analysis as oa

    def series(records):
        arrival, out = {}, []
        for time, action, subject in records:
            if action == "Arrival":
                arrival[subject] = time
            elif action == "Departure" and subject in arrival:
                out.append(time - arrival.pop(subject))
        return out

    for r in range(4):
        o = oracle_api.OracleSimulation(models, conns)
        o.set_rng_philox(8, r)
        assert o.step_until(800.0, collect=False)[0] == 0
        want = series(o.get_records("processor-3")[1])
        got = series(sim.get_records("processor-3", r))
        assert len(want) == len(got) > 100 and _close(want, got)
        e, ref = oracle_api.steady_state(want, 0.05)
        ss = oa.SteadyStateOutput.post(got)
        ci = ss.confidence_interval_mean(0.05)
        assert (ss.deletion_point, ss.batch_size, ss.batch_count) == (ref["deletion_point"], ref["batch_size"], ref["batch_count"])
        assert abs(ci.lower() - ref["lower"]) < 1e-9 and abs(ci.upper() - ref["upper"]) < 1e-9


def test_streaming_batch_means_on_device():
    """BASELINE config 3: tandem network, steady-state batch means (fixed warm-up, fixed batch size, <= 30
    batches) accumulated per replication on the device; checked against the oracle's record series and the
    SteadyStateOutput-style interval of the host analysis."""
    from sim_b200 import output_analysis as oa
    models, conns = networks.tandem()
    rec = networks.tandem()[0]
    for m in rec:
        m.inner.store_records = True
    warm, size, n, until = 40, 25, 2048, 1500.0
    sim = _sim(models, conns, n, seed=13, instrument=False, stat_model_id="processor-3", batch_warmup=warm,
               batch_size=size)
    sim.step_until(until)
    mean, var, batches = sim.get_batch_means()
    assert (batches == 30).mean() > 0.95 and (batches <= 30).all()
    for r in (0, 7, 1000, 2047):
        o = oracle_api.OracleSimulation(rec, conns)
        o.set_rng_philox(13, r)
        assert o.step_until(until, collect=False)[0] == 0
        arrival, waits = {}, []
        for t, action, subject in o.get_records("processor-3")[1]:
            if action == "Arrival":
                arrival[subject] = t
            elif action == "Processing Start":
                waits.append(t - arrival.pop(subject))
        w = waits[warm:]
        bm = [sum(w[b * size:(b + 1) * size]) / size for b in range(min(len(w) // size, 30))]
        assert batches[r] == len(bm)
        m = sum(bm) / len(bm)
        v = sum((x - m) ** 2 for x in bm) / len(bm)
        assert abs(mean[r] - m) <= 1e-9 * max(1.0, abs(m)) and abs(var[r] - v) <= 1e-7 * max(1.0, v)
        ci = oa.batch_means_confidence_interval(mean[r], var[r], int(batches[r]), 0.05)
        assert ci.lower() < mean[r] < ci.upper()
    # replications are independent: the grand mean of the replication estimates has a tight IID interval
    ok = batches == 30
    grand = oa.IndependentSample.post(mean[ok]).confidence_interval_mean(0.05)
    assert grand.half_width() < 0.05 * abs(mean[ok].mean())


# ---- BASELINE.json full sizes of the other networks (configs[2..4]) ---------------------------------------
def _sampled_equal(sim, models, conns, seed, picks, until=None, nsteps=None, stat=None):
    t, steps, events = sim.get_global_time(), sim.get_steps(), sim.get_events()
    if stat:  # the oracle derives the waiting times from the records of that model
        for m in models:
            m.inner.store_records = True
    o = oracle_api.OracleSimulation(models, conns)
    for r0 in picks:
        ref = o.run_batch(seed, r0, 2, threads=2, until=until, nsteps=nsteps, stat_model_id=stat, trace_hash=False)
        sl = slice(r0, r0 + 2)
        assert ref["err"] == 0
        assert (steps[sl] == ref["steps"]).all() and (events[sl] == ref["events"]).all()
        assert _close(t[sl], ref["time"])
        if stat:
            s, cnt = sim.get_wait_stats()
            assert (cnt[sl] == ref["wait_n"]).all() and _close(s[sl], ref["wait_sum"])


def test_full_size_tandem_batch_means():
    """configs[2]: tandem queue, 2^22 replications, steady-state batch means per replication on the device."""
    n, until, warm, size = 1 << 22, 260.0, 20, 6
    models, conns = networks.tandem()
    sim = _sim(models, conns, n, seed=21, instrument=False, stat_model_id="processor-3", batch_warmup=warm, batch_size=size)
    sim.step_until(until)
    assert (sim.get_errors() == 0).all()
    t = sim.get_global_time()
    assert (t >= until).all() and (t < until + 30.0).all()
    mean, var, batches = sim.get_batch_means()
    assert (batches <= 30).all() and (batches >= 10).mean() > 0.99
    ok = batches >= 10
    assert np.isfinite(mean[ok]).all() and (mean[ok] >= 0).all() and (var[ok] >= 0).all()
    # whole-batch accounting: every job that started service at the observed stage is in the wait count
    s, cnt = sim.get_wait_stats()
    done = np.minimum((np.maximum(cnt.astype(np.int64) - warm, 0)) // size, 30)
    assert np.array_equal(done, batches.astype(np.int64))
    _sampled_equal(sim, models, conns, 21, (0, 4095, 1 << 21, n - 2), until=until, stat="processor-3")
    from sim_b200 import output_analysis as oa
    grand = oa.IndependentSample.post(mean[ok]).confidence_interval_mean(0.05)
    assert grand.half_width() < 0.01 * abs(grand.lower() + grand.upper()) / 2.0


def test_full_size_gateway_network():
    """configs[3]: 16-model gateway DAG (exclusive / parallel / stochastic gates, batcher), 2^20 replications."""
    n, until = 1 << 20, 40.0
    models, conns = networks.gateway_network()
    sim = _sim(models, conns, n, seed=33, instrument=False)
    sim.step_until(until)
    assert (sim.get_errors() == 0).all()
    t, steps, events = sim.get_global_time(), sim.get_steps(), sim.get_events()
    assert (t >= until).all() and np.isfinite(t).all()
    assert (events >= steps).all() and np.unique(steps).size > 100
    c = sim.get_counters()
    assert c["steps"] == int(steps.astype(np.uint64).sum()) and c["events"] == int(events.astype(np.uint64).sum())
    _sampled_equal(sim, models, conns, 33, (0, 511, 500000, n - 2), until=until)


def test_full_size_large_mixed_shard():
    """configs[4]: 32-model network; one GPU's shard of the 2^26-replication job at 8 GPUs (2^23 replications,
    44 GB of state and rings), posted with the replication offset of rank 5."""
    n, nsteps, offset = 1 << 23, 48, 5 << 23
    models, conns = networks.large_mixed()
    sim = _sim(models, conns, n, seed=7, instrument=False, replication_offset=offset, steps_per_launch=16)
    sim.step_n(nsteps)
    assert (sim.get_errors() == 0).all()
    steps, events, t = sim.get_steps(), sim.get_events(), sim.get_global_time()
    assert (steps == nsteps).all() and (events >= steps).all() and np.isfinite(t).all()
    o = oracle_api.OracleSimulation(models, conns)
    for r0 in (0, 1 << 22, n - 2):  # global replication index = offset + local index
        ref = o.run_batch(7, offset + r0, 2, threads=2, nsteps=nsteps, trace_hash=False)
        sl = slice(r0, r0 + 2)
        assert (events[sl] == ref["events"]).all() and _close(t[sl], ref["time"])
    sim.close()


def test_device_steady_state_output_matches_host_and_oracle():
    """SteadyStateOutput (deletion point of set_to_fixed_budget, <= 30 batch means, asymmetric
    degrees of freedom; output_analysis/mod.rs:205-333) per replication ON THE DEVICE over the captured waiting
    time series.  Bit-equal to the host restatement over the same series; the series itself and the resulting
    interval equal what the oracle's records give."""
    from sim_b200 import SimulationError
    from sim_b200 import output_analysis as oa
    models, conns = networks.tandem()
    n, until, cap = 4096, 700.0, 1024
    sim = _sim(models, conns, n, seed=29, instrument=False, stat_model_id="processor-3", series_capacity=cap)
    sim.step_until(until)
    for alpha in (0.05, 0.001):
        res = sim.steady_state(alpha)
        assert (res["status"] == 0).all()
        assert (res["batch_count"] <= 30).all() and (res["batch_count"] >= 1).all()
        assert (res["lower"] <= res["mean"]).all() and (res["mean"] <= res["upper"]).all()
        _, cnt = sim.get_wait_stats()
        assert ((res["deletion_point"] + res["batch_size"] * res["batch_count"]) == cnt).all()
        for r in (0, 1, 63, 64, 2047, n - 1):
            series = sim.get_series(r)
            assert series.size == cnt[r]
            ss = oa.SteadyStateOutput.post(series)
            ci = ss.confidence_interval_mean(alpha)
            got = res[r]
            assert (got["deletion_point"], got["batch_size"], got["batch_count"]) == (ss.deletion_point, ss.batch_size, ss.batch_count)
            assert got["lower"] == ci.lower() and got["upper"] == ci.upper()  # same operation order: bit-equal
    # against the oracle: the series from its records, its SteadyStateOutput
    rec = networks.tandem()[0]
    for m in rec:
        m.inner.store_records = True
    res = sim.steady_state(0.05)
    for r in (5, 1000):
        o = oracle_api.OracleSimulation(rec, conns)
        o.set_rng_philox(29, r)
        assert o.step_until(until, collect=False)[0] == 0
        arrival, waits = {}, []
        for t, action, subject in o.get_records("processor-3")[1]:
            if action == "Arrival":
                arrival[subject] = t
            elif action == "Processing Start":
                waits.append(t - arrival.pop(subject))
        assert _close(sim.get_series(r), waits)
        e, ref = oracle_api.steady_state(waits, 0.05)
        assert e == 0
        assert (res[r]["deletion_point"], res[r]["batch_size"], res[r]["batch_count"]) == \
            (ref["deletion_point"], ref["batch_size"], ref["batch_count"])
        assert abs(res[r]["lower"] - ref["lower"]) < 1e-9 and abs(res[r]["upper"] - ref["upper"]) < 1e-9
    # the grand mean over replications has a tight IID interval around the replication estimates
    grand = oa.IndependentSample.post(res["mean"]).confidence_interval_mean(0.05)
    assert grand.half_width() < 0.05 * abs(res["mean"].mean())
    # an unlisted alpha panics in the reference (t_scores.rs:19-22); a too-small capacity is reported per replication
    with pytest.raises(SimulationError) as ei:
        sim.steady_state(0.3)
    assert ei.value.name == "Panic"
    small = _sim(models, conns, 64, seed=29, instrument=False, stat_model_id="processor-3", series_capacity=16)
    small.step_until(until)
    assert (small.steady_state(0.05)["status"] == 32).all()
    none = _sim(models, conns, 64, seed=29, instrument=False, stat_model_id="processor-3")
    with pytest.raises(SimulationError):
        none.steady_state(0.05)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_streaming_passes_at_full_width(k):
    """One / two steps per launch over 2^20 replications: every pass is two concurrent half-batch grids, consecutive
    passes sweep in opposite directions, the state tile copies carry L2 hints and (k <= 2) the 6-CTA instantiation
    with a one-block Philox cache runs.  None of that may change a single replication."""
    n, nsteps = 1 << 20, 51
    models, conns = networks.mm1()
    sim = _sim(models, conns, n, seed=77, instrument=False, stat_model_id="processor-01", steps_per_launch=k)
    assert sim.describe()["kernel_shape"] == "mm1" and sim.describe()["grids_per_pass"] == 2
    sim.step_n(nsteps)
    steps, events, t = sim.get_steps(), sim.get_events(), sim.get_global_time()
    s, cnt = sim.get_wait_stats()
    assert (steps == nsteps).all() and (sim.get_errors() == 0).all()
    rec = networks.mm1(store_records=True)[0]
    o = oracle_api.OracleSimulation(rec, conns)
    for r0 in (0, 254, 256, 524286, 524288, 786432, n - 2):  # tile and half-batch boundaries
        ref = o.run_batch(77, r0, 2, threads=2, nsteps=nsteps, stat_model_id="processor-01", trace_hash=False)
        sl = slice(r0, r0 + 2)
        assert (events[sl] == ref["events"]).all() and _close(t[sl], ref["time"])
        assert (cnt[sl] == ref["wait_n"]).all() and _close(s[sl], ref["wait_sum"])
    # the same batch stepped with fused launches ends in the same state
    fused = _sim(models, conns, n, seed=77, instrument=False, stat_model_id="processor-01")
    fused.step_n(nsteps)
    assert np.array_equal(fused.get_global_time(), t) and np.array_equal(fused.get_events(), events)
    assert np.array_equal(fused.get_draws(), sim.get_draws())
    for m in models:
        assert np.array_equal(fused.get_until_next_event(m.id), sim.get_until_next_event(m.id), equal_nan=True)


# ---- Simulation::put with messages in flight (mod.rs:82-85: `messages` survive, matched by id / port strings) --------
def _put_case(new_models, new_conns, steps=60, seed=17):
    """Post M/M/1, inject two messages, put the new network, step; the oracle posts the new network directly and gets
    the same two messages injected (put keeps them as strings)."""
    from sim_b200 import Simulation
    models, conns = networks.mm1()
    sim = Simulation.post(models, conns, replications=64, seed=seed, instrument=True, trace_replications=2, trace_capacity=4096)
    sim.inject_input(Message("manual", "manual", "processor-01", "job", 0.0, "x 1"))
    sim.inject_input(Message("manual", "manual", "storage-01", "store", 0.0, "y 2"))
    sim.put(new_models, new_conns)
    pend = [(m.target_id, m.target_port, m.content) for m in sim.get_messages(1)]
    err = None
    try:
        sim.step_n(steps)
    except Exception as e:  # noqa: BLE001 - the status is compared below
        err = e.status
    ref = []
    for r in range(2):
        o = oracle_api.OracleSimulation(new_models, new_conns)
        o.set_rng_philox(seed, r)
        o.trace_enable(False)
        o.trace_intern("x 1")   # content codes are handed out in the order the strings are first seen: the engine saw
        o.trace_intern("y 2")   # them at inject_input
        o.inject_input("processor-01", "job", "x 1")
        o.inject_input("storage-01", "store", "y 2")
        e, _ = o.step_n(steps, collect=False)
        ref.append((e, o.trace_hash(), o.counters()["steps"], o.get_global_time()))
    return sim, pend, err, ref


def test_put_keeps_pending_messages_when_models_are_reordered():
    models, conns = networks.mm1()
    sim, pend, err, ref = _put_case([models[2], models[0], models[1]], conns)
    assert pend == [("processor-01", "job", "x 1"), ("storage-01", "store", "y 2")] and err is None
    h, st, t = sim.get_trace_hash(), sim.get_steps(), sim.get_global_time()
    for r in range(2):
        assert ref[r][0] == 0 and h[r] == ref[r][1] and st[r] == ref[r][2] and _close(t[r], ref[r][3])


def test_put_hands_pending_messages_to_every_model_of_their_id():
    """The messages `put` keeps are strings (mod.rs:82-85); if the new model list has two models of the target id, the
    next step delivers the kept message to both (mod.rs:203-221) -- and a later `put` back to one model undoes that."""
    from sim_b200 import Model, Processor, Simulation
    base, conns = networks.mm1()
    models = list(base) + [Model.new("processor-01", Processor.new(Exp(2.0), 3, "job", "processed", False, None))]
    sim, pend, err, ref = _put_case(models, conns)
    assert pend == [("processor-01", "job", "x 1"), ("storage-01", "store", "y 2")] and err is None
    h, st, t = sim.get_trace_hash(), sim.get_steps(), sim.get_global_time()
    for r in range(2):
        assert ref[r][0] == 0 and h[r] == ref[r][1] and st[r] == ref[r][2] and _close(t[r], ref[r][3])
    # put twice before stepping: two models of the id, then one again -- one delivery per kept message is left
    sim = Simulation.post(base, conns, replications=32, seed=17, instrument=True)
    sim.inject_input(Message("manual", "manual", "processor-01", "job", 0.0, "x 1"))
    sim.put(models, conns)
    assert len(sim.get_messages(0)) == 1
    sim.put(base, conns)
    sim.step()
    plain = Simulation.post(base, conns, replications=32, seed=17, instrument=True)
    plain.inject_input(Message("manual", "manual", "processor-01", "job", 0.0, "x 1"))
    plain.step()
    assert (sim.get_events() == plain.get_events()).all() and (sim.get_trace_hash() == plain.get_trace_hash()).all()
    twins = Simulation.post(models, conns, replications=32, seed=17, instrument=True)
    twins.inject_input(Message("manual", "manual", "processor-01", "job", 0.0, "x 1"))
    twins.step()
    assert (twins.get_events() == plain.get_events() + 1).all()   # one more events_ext: the second processor-01


def test_put_re_resolves_ports_and_layout_changes():
    from sim_b200 import Generator, Model, Processor, Storage
    # (a) the processor's input port is renamed: the pending "job" message now names a port it does not have
    models = [Model.new("generator-01", Generator.new(Exp(0.5), None, "job", False, None)),
              Model.new("processor-01", Processor.new(Exp(0.333333), 14, "task", "processed", False, None)),
              Model.new("storage-01", Storage.new("store", "read", "stored", False))]
    conns = [networks.Connector.new("connector-01", "generator-01", "processor-01", "job", "task"),
             networks.Connector.new("connector-02", "processor-01", "storage-01", "processed", "store")]
    sim, pend, err, ref = _put_case(models, conns)
    assert ref[0][0] == 7 and err == 7          # InvalidMessage, as the reference's processor raises (processor.rs:225)
    # (b) a model is added in front (different state layout): the messages follow their targets
    base, bconns = networks.mm1()
    models = [Model.new("storage-00", Storage.new("store", "read", "stored", False))] + list(base)
    sim, pend, err, ref = _put_case(models, bconns)
    assert pend == [("processor-01", "job", "x 1"), ("storage-01", "store", "y 2")] and err is None
    h, st = sim.get_trace_hash(), sim.get_steps()
    for r in range(2):
        assert ref[r][0] == 0 and h[r] == ref[r][1] and st[r] == ref[r][2]
    # (c) the target of a pending message is gone: it is never delivered, and the step still is a zero-time one
    models = [base[0], base[1]]
    conns = [bconns[0]]
    sim, pend, err, ref = _put_case(models, conns)
    assert err is None and ref[0][0] == 0
    h, st, t = sim.get_trace_hash(), sim.get_steps(), sim.get_global_time()
    for r in range(2):
        assert h[r] == ref[r][1] and st[r] == ref[r][2] and _close(t[r], ref[r][3])


def test_empty_model_list_on_device():
    """Simulation::post(vec![], vec![]) (mod.rs:49-56): nothing is ever scheduled, each step adds +INF to the clock."""
    from sim_b200 import Simulation
    sim = Simulation.post([], [], replications=100)
    sim.step()
    sim.step_n(2)
    assert np.isinf(sim.get_global_time()).all() and (sim.get_steps() == 3).all() and (sim.get_events() == 0).all()
    sim.inject_input(Message("manual", "manual", "nobody", "in", 0.0, "x 1"))
    sim.step()
    assert (sim.get_errors() == 0).all() and len(sim.get_messages(0)) == 0


def test_c_abi_rejects_bad_arguments_without_touching_memory():
    """Every entry point takes plain pointers and sizes; a null handle, a null or short output buffer, a replication
    index past the batch, a model the network does not have -- each is an error return (InvalidArgument /
    ModelNotFound), never a write past the buffer or a crash."""
    import ctypes as C
    from sim_b200 import Simulation
    from sim_b200.simulator import load_library
    L = load_library()
    models, conns = networks.mm1()
    sim = Simulation.post(models, conns, replications=8, seed=1, instrument=True, trace_replications=1, trace_capacity=64,
                          stat_model_id="processor-01")
    sim.step_n(20)
    h = sim._h
    NULL = None
    INVALID, NOT_FOUND = 36, 2   # SIMB_ERR_INVALID_ARGUMENT, SIMB_ERR_MODEL_NOT_FOUND (include/simb.h)
    guard = np.full(16, -7.0)
    p = guard.ctypes.data
    vp, u64, cp, dbl = C.c_void_p, C.c_uint64, C.c_char_p, C.c_double
    def call(fn, types, *a):
        fn.argtypes = types
        fn.restype = C.c_int
        return fn(*a)
    assert call(L.simb_get_global_time, [vp, vp, u64], NULL, p, 16) == INVALID
    assert call(L.simb_get_global_time, [vp, vp, u64], h, NULL, 16) == INVALID
    assert call(L.simb_get_global_time, [vp, vp, u64], h, p, 7) == INVALID and (guard == -7.0).all()   # 8 replications
    assert call(L.simb_get_global_time, [vp, vp, u64], h, p, 8) == 0 and (guard[8:] == -7.0).all()
    assert call(L.simb_get_steps, [vp, vp, u64], h, p, 3) == INVALID
    assert call(L.simb_get_until_next_event, [vp, cp, vp, u64], h, b"nobody", p, 16) == NOT_FOUND
    assert call(L.simb_get_until_next_event, [vp, cp, vp, u64], h, NULL, p, 16) == INVALID
    n = C.c_uint64(99)
    assert call(L.simb_get_messages, [vp, u64, vp, u64, vp], h, 8, NULL, 0, C.byref(n)) == INVALID       # replication 8 of 8
    assert call(L.simb_get_messages, [vp, u64, vp, u64, vp], h, 0, NULL, 0, NULL) == INVALID
    assert call(L.simb_get_messages, [vp, u64, vp, u64, vp], h, 0, NULL, 0, C.byref(n)) == 0 and n.value <= 2
    assert call(L.simb_get_records, [vp, cp, u64, vp, u64, vp], h, b"nobody", 0, NULL, 0, C.byref(n)) == NOT_FOUND
    assert call(L.simb_get_records, [vp, cp, u64, vp, u64, vp], h, b"processor-01", 1, NULL, 0, C.byref(n)) != 0  # not a traced one
    assert call(L.simb_get_trace, [vp, u64, vp, u64, vp], h, 5, NULL, 0, C.byref(n)) != 0
    buf = C.create_string_buffer(4)
    assert call(L.simb_get_status, [vp, cp, u64, vp, u64], h, b"nobody", 0, buf, 4) == NOT_FOUND
    assert call(L.simb_get_status, [vp, cp, u64, vp, u64], h, b"processor-01", 0, buf, 4) == 0 and len(buf.value) <= 3  # truncated, terminated
    out = C.c_char_p()
    assert call(L.simb_get_json, [vp, u64, vp], h, 8, C.byref(out)) == INVALID
    assert call(L.simb_get_json, [vp, u64, vp], h, 0, NULL) == INVALID
    assert call(L.simb_step_n, [vp, u64], NULL, 1) == INVALID and call(L.simb_step_until, [vp, dbl], NULL, 1.0) == INVALID
    assert call(L.simb_inject_input, [vp, cp, cp, cp, vp], h, NULL, b"job", b"x", NULL) == INVALID
    assert call(L.simb_put_json, [vp, cp, cp], h, b"not json", b"[]") == 13
    size = C.c_uint64()
    assert call(L.simb_state_size, [vp, vp], h, C.byref(size)) == 0 and size.value > 0
    blob = C.create_string_buffer(int(size.value))
    written = C.c_uint64()
    assert call(L.simb_get_state, [vp, vp, u64, vp], h, blob, size.value - 1, C.byref(written)) == INVALID
    assert call(L.simb_get_state, [vp, vp, u64, vp], h, blob, size.value, C.byref(written)) == 0 and written.value == size.value
    assert call(L.simb_put_state, [vp, vp, u64], h, blob, size.value - 8) != 0        # truncated checkpoint
    assert call(L.simb_put_state, [vp, vp, u64], h, b"\x00" * int(size.value), size.value) != 0   # not a checkpoint
    assert call(L.simb_put_state, [vp, vp, u64], h, blob, size.value) == 0
    assert call(L.simb_get_wait_stats, [vp, vp, vp, u64], h, p, p, 2) == INVALID
    # the handle still works
    sim.step_n(5)
    assert (sim.get_steps() == 25).all() and (sim.get_errors() == 0).all()
