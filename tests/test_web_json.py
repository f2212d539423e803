"""The string API of WebSimulation (sim/src/simulator/web.rs:87-215) over the C ABI.

CPU tier: the f64 text form (serde_json / ryu layout) of simb_format_f64.
GPU tier: step_json / step_n_json / step_until_json / get_messages_json / get_records_json /
inject_input_json against the oracle's returned Vec<Message> / records on the same uniform stream,
and the reference's own web.rs scenarios (tests/web.rs:13-125, 303-316, 565-597)."""
import json
import math

import numpy as np
import pytest

import networks
import oracle_api

REL = 1e-12


def test_f64_text_form_is_serde_json():
    from sim_b200.simulator import format_f64
    # ryu `format64` layout rules as used by serde_json: always a fraction or an exponent
    cases = {
        0.0: "0.0", 1.0: "1.0", -1.5: "-1.5", 0.5: "0.5", 0.3: "0.3", 10.0: "10.0",
        123456.0: "123456.0", 1e15: "1000000000000000.0", 1e16: "1e16", 1.5e16: "1.5e16",
        1234567890123456.0: "1234567890123456.0", 1.2345678901234568e17: "1.2345678901234568e17",
        0.0001: "0.0001", 0.00001: "0.00001", 0.000001: "1e-6", 1.5e-7: "1.5e-7", 1e21: "1e21",
        1e-300: "1e-300", 5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e308",
        2.0 / 3.0: "0.6666666666666666", 100.0 / 3.0: "33.333333333333336", 12.34: "12.34",
    }
    for v, text in cases.items():
        assert format_f64(v) == text, (v, format_f64(v), text)
    assert format_f64(-0.0) == "-0.0"
    assert format_f64(float("inf")) == "null" and format_f64(float("nan")) == "null"
    rng = np.random.default_rng(11)
    for v in np.concatenate([rng.exponential(size=200) * 10.0 ** rng.integers(-12, 18, 200), rng.normal(size=50)]):
        t = format_f64(float(v))
        assert float(t) == float(v) and ("." in t or "e" in t)
        assert len(t) <= len(repr(float(v))) + 2


def _web(models, conns, n=3, replication=1, **kw):
    from sim_b200 import WebSimulation
    from sim_b200.simulator import connectors_to_json, models_to_json
    return WebSimulation.post_json(models_to_json(models), connectors_to_json(conns), n, replication=replication, **kw)


def _oracle(models, conns, seed, rep):
    o = oracle_api.OracleSimulation(models, conns)
    o.set_rng_philox(seed, rep)
    return o


def _same(got, want):
    assert len(got) == len(want), (got, want)
    for g, w in zip(got, want):
        assert list(g) == ["sourceId", "sourcePort", "targetId", "targetPort", "time", "content"]
        assert (g["sourceId"], g["sourcePort"], g["targetId"], g["targetPort"], g["content"]) == \
            (w["source_id"], w["source_port"], w["target_id"], w["target_port"], w["content"])
        assert abs(g["time"] - w["time"]) <= REL * max(1.0, abs(w["time"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mm1", "exclusive_gateway", "parallel_gateway", "batcher", "load_balancer"])
def test_step_json_forms_match_oracle(name):
    models, conns = getattr(networks, name)()
    web = _web(models, conns, n=3, replication=1, seed=7)
    o = _oracle(models, conns, 7, 1)
    for _ in range(25):
        e, want = o.step()
        assert e == 0
        _same(json.loads(web.step_json()), want)
        _same(json.loads(web.get_messages_json()), o.get_messages())
    e, want = o.step_n(40)
    _same(json.loads(web.step_n_json(40)), want)
    until = o.get_global_time() + 20.0
    e, want = o.step_until(until)
    _same(json.loads(web.step_until_json(until)), want)
    assert abs(web.get_global_time() - o.get_global_time()) <= REL * o.get_global_time()
    # the crossing step ran, its messages were not returned but are the pending set (mod.rs:277-288)
    _same(json.loads(web.get_messages_json()), o.get_messages())


@pytest.mark.gpu
def test_records_json_matches_oracle():
    models, conns = networks.mm1(store_records=True)
    web = _web(models, conns, n=2, replication=0, seed=3)
    o = _oracle(models, conns, 3, 0)
    web.step_until_json(60.0)
    o.step_until(60.0, collect=False)
    for mid in ("generator-01", "processor-01", "storage-01"):
        got = json.loads(web.get_records_json(mid))
        e, want = o.get_records(mid)
        assert e == 0 and len(got) == len(want) and len(got) > 0
        for g, w in zip(got, want):
            assert list(g) == ["time", "action", "subject"]
            assert (g["action"], g["subject"]) == (w[1], w[2])
            assert abs(g["time"] - w[0]) <= REL * max(1.0, abs(w[0]))
    from sim_b200 import SimulationError
    with pytest.raises(SimulationError) as ei:
        web.get_records_json("no-such-model")
    assert ei.value.name == "ModelNotFound"


@pytest.mark.gpu
def test_inject_input_json_exchange():
    """injection_initiated_stored_value_exchange (simulations.rs:608-678) through the string API."""
    models, conns = networks.storage_exchange()
    web = _web(models, conns, n=4, replication=2, max_pending=4)

    def msg(target, port, content):
        return json.dumps({"sourceId": "manual", "sourcePort": "manual", "targetId": target, "targetPort": port,
                           "time": 0.0, "content": content})
    web.inject_input_json(msg("storage-01", "store", "42"))
    pending = json.loads(web.get_messages_json())
    assert pending == [{"sourceId": "manual", "sourcePort": "manual", "targetId": "storage-01", "targetPort": "store",
                        "time": 0.0, "content": "42"}]
    assert json.loads(web.step_json()) == []
    web.inject_input_json(msg("storage-01", "read", ""))
    out = json.loads(web.step_json())
    assert [m["content"] for m in out] == ["42"] and out[0]["targetId"] == "storage-02"
    web.inject_input_json(msg("storage-02", "read", ""))
    out = json.loads(web.step_json())
    assert [m["content"] for m in out] == ["42"] and out[0]["targetId"] == "storage-01"
    from sim_b200 import SimulationError
    with pytest.raises(SimulationError):
        web.inject_input_json('{"targetId": "storage-01"}')  # serde: missing fields (web.rs:136-139 unwraps)


@pytest.mark.gpu
def test_web_processor_from_queue_response_time():
    """processor_from_queue_response_time_is_correct (tests/web.rs:13-125), JSON in, JSON out."""
    models = json.dumps([
        {"type": "Processor", "id": "processor-01",
         "portsIn": {"job": "job"}, "portsOut": {"job": "processed job"},
         "serviceTime": {"exp": {"lambda": 3.0}},
         "state": {"phase": "Passive", "untilNextEvent": 0.0,
                   "queue": ["job %d" % i for i in range(1, 101)], "records": []}},
        {"type": "Storage", "id": "storage-01",
         "portsIn": {"put": "store", "get": "read"}, "portsOut": {"stored": "stored"}}])
    conns = json.dumps([{"id": "connector-01", "sourceID": "processor-01", "targetID": "storage-01",
                         "sourcePort": "processed job", "targetPort": "store"}])
    from sim_b200 import WebSimulation
    web = WebSimulation.post_json(models, conns, 8, replication=5, seed=2024, queue_capacity=128)
    times = []
    for step in range(200):
        out = json.loads(web.step_json())
        if (step + 1) % 2 == 0:
            assert out[0]["content"] == "job %d" % ((step + 1) // 2)
            times.append(out[0]["time"])
        else:
            assert out == []
    batches = [times[i] - (times[i - 2] if i >= 2 else 0.0) for i in range(1, 100, 2)]
    mean = sum(batches) / 50.0
    assert abs(mean - 2.0 / 3.0) / (2.0 / 3.0) < 0.34  # the reference's own bracket (web.rs:123-125)


@pytest.mark.gpu
def test_web_processor_network_storage_arrivals():
    """processor_network_no_job_loss (tests/web.rs:129-316): 720 x step_json -> 30 arrivals at storage-0."""
    models, conns = networks.processor_network()
    web = _web(models, conns, n=2, replication=1, seed=99)
    arrivals = 0
    for _ in range(720):
        arrivals += sum(1 for m in json.loads(web.step_json()) if m["targetId"] == "storage-0")
    assert arrivals == 30


@pytest.mark.gpu
def test_web_ci_half_width_replication_loop():
    """ci_half_width_for_average_waiting_time (tests/web.rs:417-616): the replication loop
    reset / put / step_until_json(480) / get_records_json -> waiting times per processor."""
    from sim_b200 import output_analysis as oa
    from sim_b200.simulator import connectors_to_json, models_to_json
    models, conns = networks.mm1(store_records=True, capacity=None)
    mj, cj = models_to_json(models), connectors_to_json(conns)
    web = _web(models, conns, n=4, replication=0, seed=17, queue_capacity=512)
    means = []
    for _ in range(12):
        web.reset()
        web.put_json(mj, cj)
        web.step_until_json(480.0)
        stamps = {}
        for r in json.loads(web.get_records_json("processor-01")):
            stamps.setdefault(r["subject"], {})[r["action"]] = r["time"]
        waits = [s["Processing Start"] - s["Arrival"] for s in stamps.values() if "Arrival" in s and "Processing Start" in s]
        m = oa.IndependentSample.post(waits).point_estimate_mean()
        if not math.isnan(m):
            means.append(m)
    assert len(means) == 12 and len(set(means)) == 12  # the uniform stream continues across replications
    ci = oa.IndependentSample.post(means).confidence_interval_mean(0.05)
    assert ci.lower() < ci.upper() and ci.half_width() > 0.0


@pytest.mark.gpu
def test_yaml_forms_round_trip():
    """post_yaml / put_yaml / step_n_yaml / get_records_yaml (web.rs:49-70, 115-117, 213-215): the YAML documents
    of sim/tests/web.rs:419-555 style describe the same network as their JSON form."""
    import yaml
    from sim_b200 import WebSimulation
    from sim_b200.simulator import connectors_to_json, models_to_json
    models, conns = networks.mm1(store_records=True)
    my = yaml.safe_dump(json.loads(models_to_json(models)), sort_keys=False)
    cy = yaml.safe_dump(json.loads(connectors_to_json(conns)), sort_keys=False)
    a = WebSimulation.post_yaml(my, cy, 2, replication=1, seed=5)
    b = _web(models, conns, n=2, replication=1, seed=5)
    assert yaml.safe_load(a.step_n_yaml(60)) == json.loads(b.step_n_json(60))
    assert yaml.safe_load(a.get_records_yaml("processor-01")) == json.loads(b.get_records_json("processor-01"))
    a.put_yaml(my, cy)
    assert json.loads(a.get_records_json("processor-01")) == []
    a.inject_input_yaml("sourceId: manual\nsourcePort: manual\ntargetId: storage-01\ntargetPort: store\ntime: 0.0\ncontent: x\n")
    assert yaml.safe_load(a.get_messages_yaml())[-1]["content"] == "x"


@pytest.mark.gpu
def test_non_stationary_generation_replication_loop():
    """non_stationary_generation (simulations.rs:181-255) on the device with the reference's default stream
    (Pcg64Mcg::new(42)): reset + put per replication, the stream carries on; the arrivals counted from the returned
    messages of step_until(480) equal the oracle's replication by replication, and the CI overlaps the reference's."""
    from test_oracle_structural import _arrival_counts_oracle
    from sim_b200 import output_analysis as oa
    from sim_b200.simulator import RNG_PCG64MCG, connectors_to_json, models_to_json
    models, conns = networks.mm1(lam_gen=0.0957, lam_proc=0.1659)
    mj, cj = models_to_json(models), connectors_to_json(conns)
    web = _web(models, conns, n=1, replication=0, seed=42, rng_kind=RNG_PCG64MCG)
    counts = []
    for _ in range(10):
        web.reset()
        web.put_json(mj, cj)
        msgs = json.loads(web.step_until_json(480.0))
        counts.append(float(sum(1 for m in msgs if m["targetId"] == "processor-01")))
    assert counts == _arrival_counts_oracle()
    ci = oa.IndependentSample.post(counts).confidence_interval_mean(0.05)
    emp = oa.IndependentSample.post([47.0, 42.0, 45.0, 34.0, 37.0]).confidence_interval_mean(0.05)
    assert ci.lower() < emp.upper() and ci.upper() > emp.lower()


@pytest.mark.gpu
def test_string_api_on_random_networks():
    """step_json / get_messages_json / step_n_json / get_records_json of seeded random networks (tests/random_networks.py;
    Coupled models and shared ids among them) against the oracle's own Vec<Message> and records: what a caller of
    WebSimulation sees, message by message."""
    from random_networks import random_coupled_network, random_network
    from sim_b200 import SimulationError
    compared = 0
    for seed in range(16000, 16060):
        models, conns = random_coupled_network(seed) if seed % 3 == 0 else random_network(seed, acyclic=(seed % 2 == 0))
        def strip(ms):  # (oracle_api binds a model's own generator by id; records on, to read them back)
            for m in ms:
                if getattr(m.inner, "rng", None) is not None:
                    m.inner.rng = None
                if m.inner.type_name == "Coupled":
                    strip(m.inner.components)
                else:
                    m.inner.store_records = True
        strip(models)
        if any(m.inner.type_name == "Stopwatch" for m in models) or \
                any(m.inner.type_name == "Coupled" and any(c.inner.type_name == "Stopwatch" for c in m.inner.components) for m in models):
            continue  # (job names reused across generators: DESIGN.md ledger)
        web = _web(models, conns, n=1, replication=0, seed=seed, trace_capacity=1 << 14)   # (one replication: an error
        o = _oracle(models, conns, seed, 0)                                                 # of another one would surface too)
        try:
            failed = False
            for _ in range(12):
                e, want = o.step()
                if e:
                    with pytest.raises(SimulationError) as ei:
                        web.step_json()
                    assert ei.value.status == e, seed
                    failed = True
                    break
                _same(json.loads(web.step_json()), want)
                _same(json.loads(web.get_messages_json()), o.get_messages())
            if failed:
                compared += 1
                continue
            e, want = o.step_n(60)
            if e:
                with pytest.raises(SimulationError) as ei:
                    web.step_n_json(60)
                assert ei.value.status == e, seed
                compared += 1
                continue
            _same(json.loads(web.step_n_json(60)), want)
        except SimulationError as err:
            assert err.status == 32, (seed, str(err))   # a bounded container filled: not comparable
            continue
        for m in models:
            if m.inner.type_name == "Coupled":
                continue
            if sum(1 for x in models if x.id == m.id) > 1:
                continue
            e, want = o.get_records(m.id)
            got = json.loads(web.get_records_json(m.id))
            assert e == 0 and [(g["action"], g["subject"]) for g in got] == [(w[1], w[2]) for w in want], (seed, m.id)
        compared += 1
    assert compared >= 25
