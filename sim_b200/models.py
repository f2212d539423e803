"""Model descriptors mirroring ``sim::models`` (sim/src/models/*.rs): same constructor argument
order as the reference's ``new`` functions, and ``to_dict`` produces the reference's serde wire form
(``Model::serialize``, sim/src/models/model.rs:28-41: ``id``, ``type`` and the model's fields
flattened, camelCase).  The behaviour itself runs on the device (sim_b200/csrc/core.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

from .input_modeling import _Variant


def _own_rng(rng):
    """``rng: Option<DynRng>`` of Generator / Processor / ExclusiveGateway / StochasticGate (e.g. generator.rs:39,
    85-95): None, or the SEED of the model's own stream -- on the batched path a model's generator is the Philox stream
    (seed, replication index), one per replication, written ``"rng": {"seed": S}`` in the wire form (an extension: the
    reference's field is #[serde(skip)], only its Rust constructors can set it)."""
    if rng is None:
        return None
    if isinstance(rng, bool) or not isinstance(rng, int) or rng < 0:
        raise ValueError("a per-model rng is given as the non-negative integer seed of its Philox stream")
    return int(rng)


class _ModelBody:
    type_name = "?"
    # optional ``state`` block of the wire form (the reference's models carry `#[serde(default)] state`, e.g.
    # processor.rs:36-37, gate.rs:26-27): set it as an attribute to inject model state on input (web.rs:30-46)
    state = None

    def fields(self) -> dict:
        raise NotImplementedError


@dataclass
class Generator(_ModelBody):
    """generator.rs:24-40; ``Generator::new(message_interdeparture_time, thinning, job_port,
    store_records, rng)`` (:76-96).  ``thinning`` is stored but never evaluated by the reference
    (SURVEY.md App. F #2); ``rng`` (a per-model DynRng) has no wire form and must be None."""
    message_interdeparture_time: _Variant
    thinning: Optional[dict]
    job_port: str
    store_records: bool = False
    rng: Optional[int] = None
    type_name = "Generator"

    @classmethod
    def new(cls, message_interdeparture_time, thinning, job_port, store_records, rng=None):
        rng = _own_rng(rng)
        return cls(message_interdeparture_time, thinning, job_port, store_records, rng)

    def fields(self):
        d = {"messageInterdepartureTime": self.message_interdeparture_time.to_dict(),
             "portsIn": {}, "portsOut": {"job": self.job_port}, "storeRecords": bool(self.store_records)}
        if self.thinning is not None:
            d["thinning"] = self.thinning
        return d


@dataclass
class Processor(_ModelBody):
    """processor.rs:24-38; ``Processor::new(service_time, queue_capacity, job_port,
    processed_job_port, store_records, rng)`` (:98-117)."""
    service_time: _Variant
    queue_capacity: Optional[int]
    job_port: str
    processed_job_port: str
    store_records: bool = False
    rng: Optional[int] = None
    state: Optional[dict] = None
    type_name = "Processor"

    @classmethod
    def new(cls, service_time, queue_capacity, job_port, processed_job_port, store_records, rng=None):
        rng = _own_rng(rng)
        return cls(service_time, queue_capacity, job_port, processed_job_port, store_records, rng)

    def fields(self):
        d = {"serviceTime": self.service_time.to_dict(), "portsIn": {"job": self.job_port},
             "portsOut": {"job": self.processed_job_port}, "storeRecords": bool(self.store_records)}
        if self.queue_capacity is not None:
            d["queueCapacity"] = int(self.queue_capacity)
        if self.state is not None:
            d["state"] = self.state
        return d


@dataclass
class Storage(_ModelBody):
    """storage.rs:15-24; ``Storage::new(put_port, get_port, stored_port, store_records)`` (:80-98)."""
    put_port: str
    get_port: str
    stored_port: str
    store_records: bool = False
    type_name = "Storage"

    @classmethod
    def new(cls, put_port, get_port, stored_port, store_records):
        return cls(put_port, get_port, stored_port, store_records)

    def fields(self):
        return {"portsIn": {"put": self.put_port, "get": self.get_port},
                "portsOut": {"stored": self.stored_port}, "storeRecords": bool(self.store_records)}


@dataclass
class LoadBalancer(_ModelBody):
    """load_balancer.rs:15-24; ``LoadBalancer::new(job_port, flow_path_ports, store_records)``."""
    job_port: str
    flow_path_ports: List[str]
    store_records: bool = False
    type_name = "LoadBalancer"

    @classmethod
    def new(cls, job_port, flow_path_ports, store_records):
        return cls(job_port, list(flow_path_ports), store_records)

    def fields(self):
        return {"portsIn": {"job": self.job_port}, "portsOut": {"flowPaths": list(self.flow_path_ports)},
                "storeRecords": bool(self.store_records)}


@dataclass
class Gate(_ModelBody):
    """gate.rs:19-28; ``Gate::new(job_in_port, activation_port, deactivation_port, job_out_port,
    store_records)``."""
    job_in_port: str
    activation_port: str
    deactivation_port: str
    job_out_port: str
    store_records: bool = False
    type_name = "Gate"

    @classmethod
    def new(cls, job_in_port, activation_port, deactivation_port, job_out_port, store_records):
        return cls(job_in_port, activation_port, deactivation_port, job_out_port, store_records)

    def fields(self):
        return {"portsIn": {"job": self.job_in_port, "activation": self.activation_port,
                            "deactivation": self.deactivation_port},
                "portsOut": {"job": self.job_out_port}, "storeRecords": bool(self.store_records)}


@dataclass
class ExclusiveGateway(_ModelBody):
    """exclusive_gateway.rs:20-32; ``ExclusiveGateway::new(flow_paths_in, flow_paths_out,
    port_weights, store_records, rng)``."""
    flow_paths_in: List[str]
    flow_paths_out: List[str]
    port_weights: _Variant
    store_records: bool = False
    rng: Optional[int] = None
    type_name = "ExclusiveGateway"

    @classmethod
    def new(cls, flow_paths_in, flow_paths_out, port_weights, store_records, rng=None):
        rng = _own_rng(rng)
        return cls(list(flow_paths_in), list(flow_paths_out), port_weights, store_records, rng)

    def fields(self):
        return {"portsIn": {"flowPaths": list(self.flow_paths_in)},
                "portsOut": {"flowPaths": list(self.flow_paths_out)},
                "portWeights": self.port_weights.to_dict(), "storeRecords": bool(self.store_records)}


@dataclass
class ParallelGateway(_ModelBody):
    """parallel_gateway.rs:19-28; ``ParallelGateway::new(flow_paths_in, flow_paths_out, store_records)``."""
    flow_paths_in: List[str]
    flow_paths_out: List[str]
    store_records: bool = False
    type_name = "ParallelGateway"

    @classmethod
    def new(cls, flow_paths_in, flow_paths_out, store_records):
        return cls(list(flow_paths_in), list(flow_paths_out), store_records)

    def fields(self):
        return {"portsIn": {"flowPaths": list(self.flow_paths_in)},
                "portsOut": {"flowPaths": list(self.flow_paths_out)}, "storeRecords": bool(self.store_records)}


@dataclass
class StochasticGate(_ModelBody):
    """stochastic_gate.rs:19-31; ``StochasticGate::new(pass_distribution, job_in_port,
    job_out_port, store_records, rng)``."""
    pass_distribution: _Variant
    job_in_port: str
    job_out_port: str
    store_records: bool = False
    rng: Optional[int] = None
    type_name = "StochasticGate"

    @classmethod
    def new(cls, pass_distribution, job_in_port, job_out_port, store_records, rng=None):
        rng = _own_rng(rng)
        return cls(pass_distribution, job_in_port, job_out_port, store_records, rng)

    def fields(self):
        return {"passDistribution": self.pass_distribution.to_dict(), "portsIn": {"job": self.job_in_port},
                "portsOut": {"job": self.job_out_port}, "storeRecords": bool(self.store_records)}


class Metric:
    """stopwatch.rs:60-65"""
    Minimum = "Minimum"
    Maximum = "Maximum"


@dataclass
class Stopwatch(_ModelBody):
    """stopwatch.rs:21-32; ``Stopwatch::new(start_port, stop_port, metric_port, job_port, metric,
    store_records)``."""
    start_port: str
    stop_port: str
    metric_port: str
    job_port: str
    metric: str = Metric.Minimum
    store_records: bool = False
    type_name = "Stopwatch"

    @classmethod
    def new(cls, start_port, stop_port, metric_port, job_port, metric, store_records):
        return cls(start_port, stop_port, metric_port, job_port, metric, store_records)

    def fields(self):
        return {"portsIn": {"start": self.start_port, "stop": self.stop_port, "metric": self.metric_port},
                "portsOut": {"job": self.job_port}, "metric": self.metric, "storeRecords": bool(self.store_records)}


@dataclass
class Batcher(_ModelBody):
    """batcher.rs:22-33; ``Batcher::new(job_in_port, job_out_port, max_batch_time, max_batch_size,
    store_records)``."""
    job_in_port: str
    job_out_port: str
    max_batch_time: float
    max_batch_size: int
    store_records: bool = False
    type_name = "Batcher"

    @classmethod
    def new(cls, job_in_port, job_out_port, max_batch_time, max_batch_size, store_records):
        return cls(job_in_port, job_out_port, float(max_batch_time), int(max_batch_size), store_records)

    def fields(self):
        return {"portsIn": {"job": self.job_in_port}, "portsOut": {"job": self.job_out_port},
                "maxBatchTime": float(self.max_batch_time), "maxBatchSize": int(self.max_batch_size),
                "storeRecords": bool(self.store_records)}


@dataclass
class Passive(_ModelBody):
    """The zero-behaviour custom model of sim/tests/custom.rs:18-86 (``Passive::new(job_port)``),
    supported as an eleventh built-in descriptor so the plugin-contract test can be mirrored."""
    job_port: str
    type_name = "Passive"

    @classmethod
    def new(cls, job_port):
        return cls(job_port)

    def fields(self):
        return {"portsIn": {"job": self.job_port}}


@dataclass
class ExternalInputCoupling:
    """coupled.rs:40-47"""
    target_id: str
    source_port: str
    target_port: str

    def to_dict(self):
        return {"targetID": self.target_id, "sourcePort": self.source_port, "targetPort": self.target_port}


@dataclass
class ExternalOutputCoupling:
    """coupled.rs:49-56"""
    source_id: str
    source_port: str
    target_port: str

    def to_dict(self):
        return {"sourceID": self.source_id, "sourcePort": self.source_port, "targetPort": self.target_port}


@dataclass
class InternalCoupling:
    """coupled.rs:58-67"""
    source_id: str
    target_id: str
    source_port: str
    target_port: str

    def to_dict(self):
        return {"sourceID": self.source_id, "targetID": self.target_id, "sourcePort": self.source_port,
                "targetPort": self.target_port}


@dataclass
class Coupled(_ModelBody):
    """coupled.rs:14-25; ``Coupled::new(ports_in, ports_out, components, external_input_couplings,
    external_output_couplings, internal_couplings)`` (:87-106).  Messages between components travel over the internal
    couplings and are delivered when the coupled model next fires (parked messages, :186-255)."""
    ports_in: List[str]
    ports_out: List[str]
    components: List["Model"]
    external_input_couplings: List[ExternalInputCoupling]
    external_output_couplings: List[ExternalOutputCoupling]
    internal_couplings: List[InternalCoupling]
    type_name = "Coupled"

    @classmethod
    def new(cls, ports_in, ports_out, components, external_input_couplings, external_output_couplings, internal_couplings):
        return cls(list(ports_in), list(ports_out), list(components), list(external_input_couplings),
                   list(external_output_couplings), list(internal_couplings))

    def fields(self):
        return {"portsIn": {"flowPaths": list(self.ports_in)}, "portsOut": {"flowPaths": list(self.ports_out)},
                "components": [c.to_dict() for c in self.components],
                "externalInputCouplings": [c.to_dict() for c in self.external_input_couplings],
                "externalOutputCouplings": [c.to_dict() for c in self.external_output_couplings],
                "internalCouplings": [c.to_dict() for c in self.internal_couplings]}


@dataclass
class Model:
    """``Model`` (sim/src/models/model.rs:13-26): an id plus a boxed model."""
    id: str
    inner: _ModelBody

    @classmethod
    def new(cls, id, inner):  # noqa: A002 - reference argument name
        return cls(id, inner)

    def to_dict(self) -> dict:
        d = {"id": self.id, "type": self.inner.type_name}
        d.update(self.inner.fields())
        if getattr(self.inner, "rng", None) is not None:
            d["rng"] = {"seed": int(self.inner.rng)}
        if "state" not in d and getattr(self.inner, "state", None) is not None:
            d["state"] = self.inner.state
        return d
