"""Descriptor lowering (sim_b200/csrc/lower.cpp): the JSON wire schema of the reference's serde
attributes, error codes, and the layout that the roofline accounting uses."""
import json

import pytest

import hostsim_api as H
import networks
import oracle_api as O


def _lower(models_json, connectors_json="[]", **kw):
    return H.HostSim(None, None, raw_json=(models_json, connectors_json), **kw)


def test_mm1_layout_is_frozen():
    """DESIGN.md "state layout": M/M/1 = 4 f64 rows + 13 u32 rows = 84 B per replication
    (96 B with the waiting-time accumulators)."""
    models, conns = networks.mm1()
    lay = H.HostSim(models, conns, n=1, instrument=False).layout()
    assert (lay["n_drows"], lay["n_wrows"], lay["max_pending"], lay["ring_w_slots"]) == (4, 13, 2, 14)
    lay = H.HostSim(models, conns, n=1, instrument=False, stat_model_id="processor-01").layout()
    assert (lay["n_drows"], lay["n_wrows"], lay["ring_d_slots"]) == (5, 14, 14)


def test_field_order_is_irrelevant_and_unknown_fields_are_ignored():
    # tests/web.rs:321-414 (YAML field order) / web.rs:150 (`untilJobCompletion` is not a field)
    a = [{"type": "Processor", "id": "p", "portsIn": {"job": "job"}, "portsOut": {"job": "out"},
          "serviceTime": {"exp": {"lambda": 1.0}}, "queueCapacity": 3}]
    b = [{"queueCapacity": 3, "serviceTime": {"exp": {"lambda": 1.0}}, "portsOut": {"job": "out"},
          "untilJobCompletion": 0.0, "portsIn": {"job": "job"}, "id": "p", "type": "Processor"}]
    assert _lower(json.dumps(a)).layout() == _lower(json.dumps(b)).layout()


@pytest.mark.parametrize("doc,status", [
    ("[{\"id\": \"x\", \"type\": \"MyCustomModel\"}]", 33),     # custom models: SIMB_ERR_UNSUPPORTED_MODEL
    ("[{\"id\": \"x\", \"type\": \"Coupled\"}]", 13),           # a Coupled model needs its ports, components and couplings
    ("[{\"id\": \"x\"}]", 13),                                   # missing field `type`
    ("[{\"id\": \"g\", \"type\": \"Generator\", \"portsIn\": {}, \"portsOut\": {\"job\": \"job\"}}]", 13),
    ("[{\"id\": \"g\", \"type\": \"Generator\", \"portsIn\": {}, \"portsOut\": {\"job\": \"job\"}, "
     "\"messageInterdepartureTime\": {\"cauchy\": {\"median\": 1, \"scale\": 1}}}]", 13),  # unknown variant
    ("not json", 13),
])
def test_rejections(doc, status):
    with pytest.raises(RuntimeError) as ei:
        _lower(doc)
    assert "(%d)" % status in str(ei.value)


def test_an_empty_model_list_is_a_simulation():
    # Simulation::post(vec![], vec![]) is accepted (mod.rs:49-56); each step adds the +INF minimum to the clock (:225-236)
    hs = _lower("[]", n=2)
    assert hs.layout()["n_models"] == 0
    hs.step_n(2)
    assert (hs.time() == float("inf")).all() and (hs.steps() == 2).all() and (hs.events() == 0).all() and (hs.errors() == 0).all()
    o = O.OracleSimulation([], [])
    o.set_rng_philox(42, 0)
    assert o.step_n(2, collect=False)[0] == 0 and o.get_global_time() == float("inf")


def test_stat_model_must_be_a_processor():
    models, conns = networks.mm1()
    with pytest.raises(RuntimeError, match=r"\(2\)"):
        H.HostSim(models, conns, stat_model_id="nope")
    with pytest.raises(RuntimeError, match=r"\(1\)"):
        H.HostSim(models, conns, stat_model_id="storage-01")


def test_state_block_preloads_processor_queue():
    # tests/web.rs:30-46
    models, conns = networks.processor_from_queue(7)
    hs = H.HostSim(models, conns, n=1)
    assert hs.layout()["ring_w_slots"] >= 7
    hs.step_n(2)
    assert [p[0] for p in hs.pending(0)] == ["job 1"]


def test_every_network_fits_the_descriptor():
    for name, f in networks.ALL.items():
        models, conns = f()
        lay = H.HostSim(models, conns, n=1).layout()
        # (the components of a Coupled model take its place in the device's model list)
        atomic = sum(len(m.inner.components) if m.inner.type_name == "Coupled" else 1 for m in models)
        assert lay["n_models"] == atomic and lay["n_connectors"] == len(conns), name


def test_yaml_documents_of_the_reference_lower_like_their_json_form():
    """simulation_serialization_deserialization_field_ordering / _round_trip (sim/tests/web.rs:321-414): the YAML
    documents, in scrambled field order, describe the M/M/1 network; and a Model -> document -> Model round trip
    is the identity."""
    from sim_b200.simulator import _yaml_to_json, connectors_to_json, models_to_json
    models_yaml = """
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
    connectors_yaml = """
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
    from_yaml = _lower(_yaml_to_json(models_yaml), _yaml_to_json(connectors_yaml), instrument=False)
    models, conns = networks.mm1()
    conns[1].source_port = "processed job"
    models[1].inner.processed_job_port = "processed job"
    canonical = H.HostSim(models, conns, n=1, instrument=False)
    assert from_yaml.layout() == canonical.layout()
    # the scrambled document and the canonical models simulate the same replication
    from_yaml.step_n(400)
    canonical.step_n(400)
    assert from_yaml.hash()[0] == canonical.hash()[0] and from_yaml.time()[0] == canonical.time()[0]
    # round trip: document -> Model objects (to_dict) -> document
    assert json.loads(models_to_json(models)) == json.loads(json.dumps([m.to_dict() for m in models]))
    again = _lower(models_to_json(models), connectors_to_json(conns), instrument=False)
    assert again.layout() == canonical.layout()
