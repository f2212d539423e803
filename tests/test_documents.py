"""Whole-simulation documents and checkpoints: WebSimulation::{get_json, get_yaml, post_yaml, put_yaml}
(sim/src/simulator/web.rs:43-71) and the exact checkpoint simb_get_state / simb_put_state.

CPU tier: the library's YAML reader / writer (csrc/yaml.hpp) on the reference's own YAML fixtures
(sim/tests/web.rs:322-407) and against PyYAML.  GPU tier: get_json's schema, the post-again round trip of a
document on an identical uniform stream, and a bit-exact resume of a half-run batch."""
import json

import numpy as np
import pytest

import networks
import oracle_api

# sim/tests/web.rs:322-347 (field order does not matter) and :369-397
MODELS_YAML_SHUFFLED = """
- id: "generator-01"
  portsIn: {}
  portsOut:
    job: "job"
  messageInterdepartureTime:
    exp:
      lambda: 0.5
  type: "Generator"
- type: "Processor"
  portsIn:
    job: "job"
  portsOut:
    job: "processed job"
  serviceTime:
    exp:
      lambda: 0.333333
  queueCapacity: 14
  id: "processor-01"
- portsIn:
    put: "store"
    get: "read"
  portsOut:
    stored: "stored"
  type: "Storage"
  id: "storage-01"
"""
CONNECTORS_YAML = """
- sourcePort: "job"
  targetPort: "job"
  id: "connector-01"
  sourceID: "generator-01"
  targetID: "processor-01"
- id: "connector-02"
  targetPort: "store"
  targetID: "storage-01"
  sourcePort: "processed job"
  sourceID: "processor-01"
"""


def test_yaml_reader_on_the_reference_fixtures():
    import yaml
    from sim_b200.simulator import _yaml_to_json
    for text in (MODELS_YAML_SHUFFLED, CONNECTORS_YAML):
        assert json.loads(_yaml_to_json(text)) == yaml.safe_load(text)
    m = json.loads(_yaml_to_json(MODELS_YAML_SHUFFLED))
    assert [x["type"] for x in m] == ["Generator", "Processor", "Storage"] and m[1]["queueCapacity"] == 14
    assert m[0]["portsIn"] == {} and m[0]["messageInterdepartureTime"]["exp"]["lambda"] == 0.5


def test_yaml_writer_reader_round_trip_against_pyyaml():
    import yaml
    from sim_b200.simulator import _json_to_yaml, _yaml_to_json, connectors_to_json, models_to_json
    docs = []
    for name in sorted(networks.ALL):
        models, conns = networks.ALL[name]()
        docs += [json.loads(models_to_json(models)), json.loads(connectors_to_json(conns))]
    docs.append({"a": [], "b": {}, "c": None, "d": [1, 2.5, -3, 1.5e-7, "x: y", "- z", "", " lead", "true", "12", "#c", "q\"uote", "it's"],
                 "nested": [[1, [2, 3]], {"k": [{"p": 1}, {"q": [True, False]}]}], "e": [{"only": {}}], "f": 0.1})
    for doc in docs:
        text = _json_to_yaml(json.dumps(doc))
        assert text.startswith("---\n")
        assert yaml.safe_load(text) == doc, text                 # PyYAML reads what the writer wrote
        assert json.loads(_yaml_to_json(text)) == doc             # ... and so does the library's reader
        assert json.loads(_yaml_to_json(yaml.safe_dump(doc, sort_keys=False))) == doc   # PyYAML's block style is read too
    # serde_yaml 0.8 layout: sequences indented under their key, `~` for null
    assert _json_to_yaml('{"models":[{"id":"a","state":{"job":null,"jobs":[]}}]}') == \
        "---\nmodels:\n  - id: a\n    state:\n      job: ~\n      jobs: []\n"


def test_yaml_documents_of_random_networks():
    """Model and connector documents of seeded random networks (Coupled ones included) through the writer and the
    reader, and through PyYAML in both directions -- its block style and its flow style, which wraps long collections
    over several lines."""
    import yaml
    from random_networks import random_coupled_network, random_network
    from sim_b200.simulator import _json_to_yaml, _yaml_to_json, connectors_to_json, models_to_json
    for seed in range(14000, 14060):
        models, conns = random_network(seed) if seed % 2 else random_coupled_network(seed)
        for text in (models_to_json(models), connectors_to_json(conns)):
            doc = json.loads(text)
            y = _json_to_yaml(text)
            assert json.loads(_yaml_to_json(y)) == doc and yaml.safe_load(y) == doc, seed
            for flow in (False, True):
                assert json.loads(_yaml_to_json(yaml.safe_dump(doc, default_flow_style=flow, sort_keys=False))) == doc, (seed, flow)


def test_yaml_errors_are_reported():
    from sim_b200 import SimulationError
    from sim_b200.simulator import _yaml_to_json
    for bad in ("a: &x 1", "a:\n\t- 1", "k: [1, 2", "- a\n---\n- b"):
        with pytest.raises(SimulationError):
            _yaml_to_json(bad)


@pytest.mark.gpu
def test_get_json_document_schema_and_values():
    """get_json (web.rs:43-45): Simulation {models, connectors, messages, services} with every model's live state in the
    field order of the reference's State structs; values against the oracle's view of the same replication."""
    from sim_b200 import Simulation
    models, conns = networks.mm1(store_records=True)
    sim = Simulation.post(models, conns, replications=4, seed=9, instrument=True, trace_replications=4, trace_capacity=4096)
    sim.step_n(101)
    doc = json.loads(sim.get_json(2))
    assert list(doc) == ["models", "connectors", "messages", "services"]
    assert [m["id"] for m in doc["models"]] == [m.id for m in models]
    g, p, st = doc["models"]
    assert list(g)[:2] == ["id", "type"] and list(g["state"]) == ["phase", "untilNextEvent", "untilJob", "lastJob", "records"]
    assert list(p["state"]) == ["phase", "untilNextEvent", "queue", "records"]
    assert list(st["state"]) == ["phase", "untilNextEvent", "job", "records"]
    o = oracle_api.OracleSimulation(models, conns)
    o.set_rng_philox(9, 2)
    o.step_n(101, collect=False)
    assert doc["services"]["globalTime"] == pytest.approx(o.get_global_time(), rel=1e-12)
    assert p["state"]["untilNextEvent"] == pytest.approx(o.until_next_event("processor-01"), rel=1e-12)
    assert [(r["action"], r["subject"]) for r in p["state"]["records"]] == [(a, s) for _, a, s in o.get_records("processor-01")[1]]
    assert [(m["targetId"], m["content"]) for m in doc["messages"]] == [(m["target_id"], m["content"]) for m in o.get_messages()]
    assert p["state"]["phase"] == ("Active" if o.get_status("processor-01")[1] == "Processing" else "Passive")
    assert doc["connectors"] == json.loads(__import__("sim_b200").simulator.connectors_to_json(conns))
    # the YAML form is the same document
    import yaml
    def no_inf(v):  # JSON has no infinity (serde_json writes null); YAML has `.inf`
        if isinstance(v, dict):
            return {k: no_inf(x) for k, x in v.items()}
        if isinstance(v, list):
            return [no_inf(x) for x in v]
        return None if isinstance(v, float) and v == float("inf") else v
    assert no_inf(yaml.safe_load(sim.get_yaml(2))) == doc


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mm1", "state_injection", "state_injection_tables", "gateway_network", "stopwatch",
                                  "coupled_closure", "coupled_pipeline"])
def test_posting_a_document_again_continues_the_same_way(name):
    """The models (with `state`) and connectors of get_json posted again -- pending messages re-injected, the uniform
    stream positioned where the first run left it -- continue exactly like the first run: `state` blocks of every
    model type round-trip through the lowering."""
    models, conns = networks.ALL[name]()
    assert _repost_and_compare(models, conns, 157, 400)


def _repost_and_compare(models, conns, first, more, n=3):
    """False when the run is not comparable (a replication failed or filled a bounded container)."""
    from sim_b200 import Message, Simulation
    from sim_b200.simulator import RNG_EXTERNAL
    draws = np.stack([oracle_api.philox_fill(21, r, 6000) for r in range(n)], axis=1)   # draws[i, r]
    a = Simulation.post(models, conns, replications=n, rng_kind=RNG_EXTERNAL, instrument=True, trace_replications=n, trace_capacity=1 << 15)
    a.set_rng_stream(draws)
    try:
        a.step_n(first)
        used = a.get_draws()
        docs = [json.loads(a.get_json(r)) for r in range(n)]
        t0 = a.get_global_time()
        a.step_n(more)
    except RuntimeError:
        return False
    if a.get_errors().any():
        return False
    for r in range(n):
        b = Simulation.post_json(json.dumps(docs[r]["models"]), json.dumps(docs[r]["connectors"]), 1, rng_kind=RNG_EXTERNAL,
                                 instrument=True, trace_replications=1, trace_capacity=1 << 15)
        b.set_rng_stream(draws[used[r]:, r:r + 1])
        for m in docs[r]["messages"]:
            b.inject_input(Message(m["sourceId"], m["sourcePort"], m["targetId"], m["targetPort"], m["time"], m["content"]))
        b.step_n(more)
        # clocks restart at 0 in the new simulation (services are not posted): compare elapsed time and the model states
        assert b.get_global_time()[0] == pytest.approx(a.get_global_time()[r] - t0[r], rel=1e-9, abs=1e-9)
        da, db = json.loads(a.get_json(r)), json.loads(b.get_json(0))
        def atomic(models):  # the components of a Coupled model in its place, and its own state (the parked messages)
            out = []
            for m in models:
                if m["type"] == "Coupled":
                    out.append({"id": m["id"], "type": "CoupledState", "state": m["state"]})
                    out += atomic(m["components"])
                else:
                    out.append(m)
            return out
        for ma, mb in zip(atomic(da["models"]), atomic(db["models"])):
            sa, sb = dict(ma.get("state", {})), dict(mb.get("state", {}))
            sa.pop("records", None), sb.pop("records", None)
            if ma["type"] == "Stopwatch":
                # the new simulation's clock restarts at 0 (post takes models and connectors, not services), so the
                # durations of jobs that straddle the restart differ: compare the jobs still missing a side
                open_a = [(j["name"], j["start"] is None) for j in sa["jobs"] if (j["start"] is None) != (j["stop"] is None)]
                open_b = [(j["name"], j["start"] is None) for j in sb["jobs"] if (j["start"] is None) != (j["stop"] is None)]
                assert open_a == open_b
                sa.pop("jobs"), sb.pop("jobs")
            for k in sa:
                if isinstance(sa[k], float) and sa[k] is not None:
                    assert sb[k] == pytest.approx(sa[k], rel=1e-9, abs=1e-9), (ma["id"], k)
                else:
                    assert sa[k] == sb[k], (ma["id"], k)
        assert [(m["targetId"], m["content"]) for m in da["messages"]] == [(m["targetId"], m["content"]) for m in db["messages"]]
        assert b.get_steps()[0] == more
    return True


@pytest.mark.gpu
def test_documents_of_random_networks_continue_the_same_way():
    """The same round trip on seeded random networks (tests/random_networks.py), Coupled ones included: whatever state
    the wiring drives the models into -- half-filled ParallelGateway collections, parked messages, blocked gates --
    is written by get_json and read back by the lowering."""
    from random_networks import random_coupled_network, random_network
    done = 0
    for seed in range(13000, 13040):
        models, conns = random_network(seed, acyclic=True) if seed % 2 else random_coupled_network(seed)
        def strip(ms):
            for m in ms:
                if getattr(m.inner, "rng", None) is not None:
                    m.inner.rng = None       # (a model's own generator needs the Philox streams)
                if m.inner.type_name == "Coupled":
                    strip(m.inner.components)
        strip(models)
        done += bool(_repost_and_compare(models, conns, 61 + seed % 40, 200, n=2))
    assert done >= 15


@pytest.mark.gpu
@pytest.mark.parametrize("name,stat", [("mm1", "processor-01"), ("gateway_network", None), ("large_mixed", None)])
def test_checkpoint_resumes_bit_exactly(name, stat):
    """simb_get_state / simb_put_state: a batch stopped half way, saved, restored into a freshly posted handle and
    continued ends in exactly the state of the uninterrupted run (clocks, counters, draws, trace hashes, statistics,
    every model's until_next_event, pending messages)."""
    from sim_b200 import Simulation
    models, conns = networks.ALL[name]()
    n, half, full = 777, 60.0, 150.0
    kw = dict(replications=n, seed=31, replication_offset=1000, instrument=True, stat_model_id=stat, steps_per_launch=16)
    ref = Simulation.post(models, conns, **kw)
    ref.step_until(full)
    a = Simulation.post(models, conns, **kw)
    a.step_until(half)
    blob = a.get_state()
    a.close()
    b = Simulation.post(models, conns, **kw)
    b.step_n(5)          # any state at all: the checkpoint replaces it
    b.put_state(blob)
    b.step_until(full)
    assert np.array_equal(b.get_global_time(), ref.get_global_time())
    for get in ("get_steps", "get_events", "get_draws", "get_trace_hash", "get_errors"):
        assert np.array_equal(getattr(b, get)(), getattr(ref, get)()), get
    for m in models:
        assert np.array_equal(b.get_until_next_event(m.id), ref.get_until_next_event(m.id), equal_nan=True)
    assert np.array_equal(b.get_message_counts(), ref.get_message_counts())
    assert np.array_equal(b.get_record_counts(), ref.get_record_counts())
    if stat:
        (s1, c1), (s2, c2) = b.get_wait_stats(), ref.get_wait_stats()
        assert np.array_equal(s1, s2) and np.array_equal(c1, c2)
    assert b.get_json(5) == ref.get_json(5)
    # a blob from another network is refused
    other = Simulation.post(*networks.tandem(), replications=n, instrument=True)
    from sim_b200 import SimulationError
    with pytest.raises(SimulationError):
        other.put_state(blob)
