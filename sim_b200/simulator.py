"""``Simulation`` -- host-side mirror of ``sim::simulator::Simulation``
(sim/src/simulator/mod.rs:39-304) for N lockstep replications, over the C ABI of
``libsimb200.so`` (include/simb.h).

Differences from the reference that follow from batching: ``step*`` return
nothing (a ``Vec<Message>`` per replication cannot be materialised for 10^6 replications); results
are read through ``get_global_time`` / ``get_messages(replication)`` / counters / statistics, and
the full event trace is available for a chosen number of leading replications.

There is no CPU fallback: if the CUDA library is missing or no device is present this raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

from .coupling import Connector, Message
from .models import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("SIMB_LIBRARY", os.path.join(_HERE, "libsimb200.so"))  # override: kernel experiments

RNG_PHILOX, RNG_PCG64MCG, RNG_EXTERNAL = 0, 1, 2
FLAG_INSTRUMENT = 1

ACTION_NAMES = ["Initialization", "Generation", "Arrival", "Processing Start", "Drop", "Departure", "Activation",
                "Deactivation", "Pass", "Block", "Start", "Stop", "Minimum Fetch", "Maximum Fetch"]


class SimbConfig(C.Structure):
    """simb_config (include/simb.h)"""
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32), ("replications", C.c_uint64),
        ("replication_offset", C.c_uint64), ("seed", C.c_uint64), ("rng_kind", C.c_int32), ("device", C.c_int32),
        ("steps_per_launch", C.c_uint32), ("queue_capacity", C.c_uint32), ("max_pending", C.c_uint32),
        ("trace_replications", C.c_uint32), ("trace_capacity", C.c_uint32), ("reserved0", C.c_uint32),
        ("stat_model_id", C.c_char_p), ("batch_warmup", C.c_uint32), ("batch_size", C.c_uint32),
        ("series_capacity", C.c_uint32), ("reserved1", C.c_uint32),
    ]


class _Message(C.Structure):
    _fields_ = [("source_id", C.c_char * 48), ("source_port", C.c_char * 48), ("target_id", C.c_char * 48),
                ("target_port", C.c_char * 48), ("content", C.c_char * 48), ("time", C.c_double)]


class _Record(C.Structure):
    _fields_ = [("time", C.c_double), ("action", C.c_char * 24), ("subject", C.c_char * 104)]


class _TraceEvent(C.Structure):
    _fields_ = [("step", C.c_uint32), ("kind", C.c_uint8), ("action", C.c_uint8), ("aux", C.c_uint8),
                ("reserved", C.c_uint8), ("index", C.c_uint16), ("reserved1", C.c_uint16), ("content", C.c_uint32),
                ("time", C.c_double)]


class _Counters(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("events", C.c_uint64), ("launches", C.c_uint64), ("kernel_ms", C.c_double),
                ("state_bytes_per_replication", C.c_uint64), ("failed", C.c_uint64), ("ring_bytes", C.c_uint64)]


class _SteadyState(C.Structure):
    _fields_ = [("mean", C.c_double), ("variance", C.c_double), ("lower", C.c_double), ("upper", C.c_double),
                ("deletion_point", C.c_uint32), ("batch_size", C.c_uint32), ("batch_count", C.c_uint32),
                ("status", C.c_int32)]


class _CI(C.Structure):
    _fields_ = [("mean", C.c_double), ("variance", C.c_double), ("lower", C.c_double), ("upper", C.c_double),
                ("n", C.c_uint64)]


class SimulationError(RuntimeError):
    """``SimulationError`` (sim/src/utils/errors.rs:5-97) plus the engine's own codes (simb_status)."""

    def __init__(self, status: int, name: str, message: str):
        super().__init__("%s (%d): %s" % (name, status, message))
        self.status = status
        self.name = name


_lib = None


def load_library() -> C.CDLL:
    """Load libsimb200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError("sim_b200: %s is missing -- build it with `make -C sim_b200/csrc` or "
                          "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback" % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    L.simb_last_error_message.restype = C.c_char_p
    L.simb_status_name.restype = C.c_char_p
    L.simb_status_name.argtypes = [C.c_int]
    L.simb_usize_sqrt.restype = C.c_uint64
    L.simb_usize_sqrt.argtypes = [C.c_uint64]
    L.simb_post_json.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(SimbConfig), C.POINTER(C.c_void_p)]
    L.simb_post_yaml.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(SimbConfig), C.POINTER(C.c_void_p)]
    L.simb_put_yaml.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.simb_get_json.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_get_yaml.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_yaml_to_json.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_json_to_yaml.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_state_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.simb_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_put_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    for name in ("simb_put_json",):
        getattr(L, name).argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    for name in ("simb_reset", "simb_reset_messages", "simb_reset_global_time", "simb_destroy", "simb_step",
                 "simb_synchronize", "simb_reset_counters"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.simb_set_rng.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.simb_set_rng_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.simb_inject_input.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p]
    L.simb_step_n.argtypes = [C.c_void_p, C.c_uint64]
    L.simb_step_until.argtypes = [C.c_void_p, C.c_double]
    for name in ("simb_get_global_time", "simb_get_errors", "simb_get_steps", "simb_get_events", "simb_get_draws",
                 "simb_get_trace_hash", "simb_get_message_counts", "simb_get_record_counts"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.simb_get_until_next_event.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64]
    L.simb_get_messages.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_get_status.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    L.simb_get_records.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_get_counters.argtypes = [C.c_void_p, C.POINTER(_Counters)]
    L.simb_get_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_render_content.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64]
    L.simb_intern_content.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32)]
    L.simb_get_wait_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.simb_get_batch_means.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.simb_steady_state.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_uint64]
    L.simb_get_series.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.simb_oa_batch_means_ci.argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_double, C.POINTER(_CI)]
    L.simb_stat_partial.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.simb_stat_partial_sqdev.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    L.simb_independent_sample_ci.argtypes = [C.c_void_p, C.c_double, C.POINTER(_CI)]
    L.simb_oa_confidence_interval.argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_double, C.POINTER(_CI)]
    L.simb_oa_independent_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.POINTER(_CI)]
    L.simb_oa_steady_state.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.POINTER(_CI), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.simb_oa_t_score.argtypes = [C.c_double, C.c_uint64, C.POINTER(C.c_double)]
    L.simb_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    L.simb_device_count.argtypes = [C.POINTER(C.c_int)]
    L.simb_get_messages_json.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_get_records_json.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_inject_input_json.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    L.simb_step_json.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_step_n_json.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_step_until_json.argtypes = [C.c_void_p, C.c_double, C.c_uint64, C.POINTER(C.c_char_p)]
    L.simb_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_uint64]
    _lib = L
    return L


def _check(status: int):
    if status != 0:
        L = load_library()
        raise SimulationError(status, L.simb_status_name(status).decode(), L.simb_last_error_message().decode())


def models_to_json(models: Iterable[Model]) -> str:
    return json.dumps([m.to_dict() for m in models])


def connectors_to_json(connectors: Iterable[Connector]) -> str:
    return json.dumps([c.to_dict() for c in connectors])


class Simulation:
    """N lockstep replications of one Model/Connector network on one GPU."""

    def __init__(self, handle, n, models_json, connectors_json):
        self._h = handle
        self.replications = n
        self._models_json = models_json
        self._connectors_json = connectors_json

    # ---- constructors ---------------------------------------------------------------------------------
    @classmethod
    def post(cls, models: Sequence[Model], connectors: Sequence[Connector], replications: int = 1, **kw) -> "Simulation":
        """``Simulation::post(models, connectors)`` (mod.rs:49-56) for ``replications`` replications.

        Keyword options (simb_config): seed=42, rng_kind=RNG_PHILOX, replication_offset=0, device=0,
        steps_per_launch=0, queue_capacity=0, max_pending=0, instrument=False, trace_replications=0,
        trace_capacity=0, stat_model_id=None."""
        return cls.post_json(models_to_json(models), connectors_to_json(connectors), replications, **kw)

    @classmethod
    def post_with_rng(cls, models, connectors, rng_kind: int, seed: int, replications: int = 1, **kw) -> "Simulation":
        """``Simulation::post_with_rng`` (mod.rs:60-75): the uniform source is a (kind, seed) pair."""
        return cls.post(models, connectors, replications, rng_kind=rng_kind, seed=seed, **kw)

    @classmethod
    def post_json(cls, models_json: str, connectors_json: str, replications: int = 1, *, seed: int = 42,
                  rng_kind: int = RNG_PHILOX, replication_offset: int = 0, device: int = 0, steps_per_launch: int = 0,
                  queue_capacity: int = 0, max_pending: int = 0, instrument: bool = False, trace_replications: int = 0,
                  trace_capacity: int = 0, stat_model_id: Optional[str] = None, batch_warmup: int = 0,
                  batch_size: int = 0, series_capacity: int = 0, _yaml: bool = False) -> "Simulation":
        """``WebSimulation::post_json`` (web.rs:23-31) form of ``post``."""
        L = load_library()
        cfg = SimbConfig()
        cfg.struct_size = C.sizeof(SimbConfig)
        cfg.flags = FLAG_INSTRUMENT if instrument else 0
        cfg.replications = replications
        cfg.replication_offset = replication_offset
        cfg.seed = seed
        cfg.rng_kind = rng_kind
        cfg.device = device
        cfg.steps_per_launch = steps_per_launch
        cfg.queue_capacity = queue_capacity
        cfg.max_pending = max_pending
        cfg.trace_replications = trace_replications
        cfg.trace_capacity = trace_capacity
        cfg.stat_model_id = stat_model_id.encode() if stat_model_id else None
        cfg.batch_warmup = batch_warmup
        cfg.batch_size = batch_size
        cfg.series_capacity = series_capacity
        h = C.c_void_p()
        post = L.simb_post_yaml if _yaml else L.simb_post_json
        _check(post(models_json.encode(), connectors_json.encode(), C.byref(cfg), C.byref(h)))
        return cls(h, replications, models_json, connectors_json)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            load_library().simb_destroy(self._h)
            self._h = None

    # ---- lifecycle --------------------------------------------------------------------------------------
    def put(self, models: Sequence[Model], connectors: Sequence[Connector]):
        """``Simulation::put`` (mod.rs:82-85): fresh models, same clock / messages / RNG."""
        _check(load_library().simb_put_json(self._h, models_to_json(models).encode(), connectors_to_json(connectors).encode()))

    def reset(self):
        """``Simulation::reset`` (mod.rs:131-134)."""
        _check(load_library().simb_reset(self._h))

    def reset_messages(self):
        _check(load_library().simb_reset_messages(self._h))

    def reset_global_time(self):
        _check(load_library().simb_reset_global_time(self._h))

    def set_rng(self, rng_kind: int, seed: int):
        """``Simulation::set_rng`` (mod.rs:77-79)."""
        _check(load_library().simb_set_rng(self._h, rng_kind, seed))

    def set_rng_stream(self, draws: np.ndarray):
        """RNG_EXTERNAL: ``draws[i, r]`` is the i-th u64 consumed by replication r."""
        a = np.ascontiguousarray(draws, dtype=np.uint64)
        assert a.ndim == 2 and a.shape[1] == self.replications
        _check(load_library().simb_set_rng_stream(self._h, a.ctypes.data, a.shape[0]))

    # ---- stepping ---------------------------------------------------------------------------------------
    def inject_input(self, message: Message, replica_mask: Optional[np.ndarray] = None):
        """``Simulation::inject_input`` (mod.rs:189-191) on all (or the masked) replications."""
        m = None
        if replica_mask is not None:
            m = np.ascontiguousarray(replica_mask, dtype=np.uint8)
            assert m.size == self.replications
        _check(load_library().simb_inject_input(self._h, message.target_id.encode(), message.target_port.encode(),
                                                message.content.encode(), m.ctypes.data if m is not None else None))

    def step(self):
        """``Simulation::step`` (mod.rs:198-272) once on every replication."""
        _check(load_library().simb_step(self._h))

    def step_n(self, n: int):
        """``Simulation::step_n`` (mod.rs:293-303)."""
        _check(load_library().simb_step_n(self._h, n))

    def step_until(self, until: float):
        """``Simulation::step_until`` (mod.rs:277-288), per replication."""
        _check(load_library().simb_step_until(self._h, float(until)))

    # ---- observation ------------------------------------------------------------------------------------
    def _vec(self, fn, dtype):
        out = np.zeros(self.replications, dtype)
        _check(fn(self._h, out.ctypes.data, self.replications))
        return out

    def get_global_time(self) -> np.ndarray:
        """``get_global_time`` (mod.rs:99) of every replication."""
        return self._vec(load_library().simb_get_global_time, np.float64)

    def get_messages(self, replication: int = 0) -> List[Message]:
        """``get_messages`` (mod.rs:94) of one replication."""
        L = load_library()
        cnt = C.c_uint64()
        buf = (_Message * 256)()
        _check(L.simb_get_messages(self._h, replication, buf, 256, C.byref(cnt)))
        return [Message(b.source_id.decode(), b.source_port.decode(), b.target_id.decode(), b.target_port.decode(),
                        b.time, b.content.decode()) for b in buf[:min(cnt.value, 256)]]

    def get_status(self, model_id: str, replication: int = 0) -> str:
        """``get_status`` (mod.rs:106-113)."""
        buf = C.create_string_buffer(256)
        _check(load_library().simb_get_status(self._h, model_id.encode(), replication, buf, 256))
        return buf.value.decode()

    def get_records(self, model_id: str, replication: int = 0):
        """``get_records`` (mod.rs:118-125) as (time, action, subject) tuples; traced replications only."""
        L = load_library()
        cnt = C.c_uint64()
        _check(L.simb_get_records(self._h, model_id.encode(), replication, None, 0, C.byref(cnt)))
        buf = (_Record * max(1, cnt.value))()
        _check(L.simb_get_records(self._h, model_id.encode(), replication, buf, cnt.value, C.byref(cnt)))
        return [(b.time, b.action.decode(), b.subject.decode()) for b in buf[:cnt.value]]

    def get_until_next_event(self, model_id: str) -> np.ndarray:
        out = np.zeros(self.replications, np.float64)
        _check(load_library().simb_get_until_next_event(self._h, model_id.encode(), out.ctypes.data, self.replications))
        return out

    def get_errors(self) -> np.ndarray:
        return self._vec(load_library().simb_get_errors, np.int32)

    def get_steps(self) -> np.ndarray:
        return self._vec(load_library().simb_get_steps, np.uint32)

    def get_events(self) -> np.ndarray:
        return self._vec(load_library().simb_get_events, np.uint32)

    def get_draws(self) -> np.ndarray:
        return self._vec(load_library().simb_get_draws, np.uint32)

    def get_trace_hash(self) -> np.ndarray:
        return self._vec(load_library().simb_get_trace_hash, np.uint64)

    def get_json(self, replication: int = 0) -> str:
        """``WebSimulation::get_json`` (web.rs:43-45) of one replication: models with their live ``state``,
        connectors, messages, services."""
        out = C.c_char_p()
        _check(load_library().simb_get_json(self._h, replication, C.byref(out)))
        return out.value.decode()

    def get_yaml(self, replication: int = 0) -> str:
        out = C.c_char_p()
        _check(load_library().simb_get_yaml(self._h, replication, C.byref(out)))
        return out.value.decode()

    def get_state(self) -> np.ndarray:
        """Checkpoint of the whole batch (simb_get_state): an opaque byte blob."""
        n = C.c_uint64()
        _check(load_library().simb_state_size(self._h, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        _check(load_library().simb_get_state(self._h, buf.ctypes.data, n.value, None))
        return buf

    def put_state(self, blob: np.ndarray):
        """Resume from a checkpoint taken on a handle posted with the same network and configuration."""
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        _check(load_library().simb_put_state(self._h, b.ctypes.data, b.size))

    def get_counters(self) -> dict:
        c = _Counters()
        _check(load_library().simb_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in _Counters._fields_}

    def reset_counters(self):
        _check(load_library().simb_reset_counters(self._h))

    def get_message_counts(self) -> np.ndarray:
        n = len(json.loads(self._connectors_json))
        out = np.zeros(max(1, n), np.uint64)
        _check(load_library().simb_get_message_counts(self._h, out.ctypes.data, out.size))
        return out[:n]

    def get_record_counts(self) -> np.ndarray:
        n = len(json.loads(self._models_json))
        out = np.zeros(n * len(ACTION_NAMES), np.uint64)
        _check(load_library().simb_get_record_counts(self._h, out.ctypes.data, out.size))
        return out.reshape(n, len(ACTION_NAMES))

    def render_content(self, code: int) -> str:
        buf = C.create_string_buffer(256)
        _check(load_library().simb_render_content(self._h, code, buf, 256))
        return buf.value.decode()

    def get_trace(self, replication: int = 0) -> List[dict]:
        L = load_library()
        cnt = C.c_uint64()
        _check(L.simb_get_trace(self._h, replication, None, 0, C.byref(cnt)))
        buf = (_TraceEvent * max(1, cnt.value))()
        _check(L.simb_get_trace(self._h, replication, buf, cnt.value, C.byref(cnt)))
        return [dict(step=b.step, kind=b.kind, index=b.index, action=b.action, aux=b.aux,
                     content=self.render_content(b.content), time=b.time) for b in buf[:cnt.value]]

    # ---- statistics -------------------------------------------------------------------------------------
    def get_wait_stats(self):
        """Per-replication (sum, count) of Processing Start - Arrival on ``stat_model_id`` (web.rs:571-597)."""
        s = np.zeros(self.replications, np.float64)
        n = np.zeros(self.replications, np.uint32)
        _check(load_library().simb_get_wait_stats(self._h, s.ctypes.data, n.ctypes.data, self.replications))
        return s, n

    def get_batch_means(self):
        """Per-replication streaming batch means of the waiting times (batch_size > 0): (mean of the batch
        means, their population variance, number of completed batches <= 30)."""
        m = np.zeros(self.replications, np.float64)
        v = np.zeros(self.replications, np.float64)
        b = np.zeros(self.replications, np.uint32)
        _check(load_library().simb_get_batch_means(self._h, m.ctypes.data, v.ctypes.data, b.ctypes.data, self.replications))
        return m, v, b

    def get_series(self, replication: int = 0) -> np.ndarray:
        """The captured waiting-time series of one replication (``series_capacity`` > 0)."""
        L = load_library()
        cnt = C.c_uint64()
        _check(L.simb_get_series(self._h, replication, None, 0, C.byref(cnt)))
        out = np.zeros(cnt.value, np.float64)
        _check(L.simb_get_series(self._h, replication, out.ctypes.data, out.size, C.byref(cnt)))
        return out

    def steady_state(self, alpha: float) -> np.ndarray:
        """``SteadyStateOutput::post`` + ``confidence_interval_mean`` (output_analysis/mod.rs:205-333) of every
        replication's captured series, computed on the device.  Structured array with fields mean, variance,
        lower, upper, deletion_point, batch_size, batch_count, status."""
        dt = np.dtype([("mean", "f8"), ("variance", "f8"), ("lower", "f8"), ("upper", "f8"), ("deletion_point", "u4"),
                       ("batch_size", "u4"), ("batch_count", "u4"), ("status", "i4")])
        assert dt.itemsize == C.sizeof(_SteadyState)
        out = np.zeros(self.replications, dt)
        _check(load_library().simb_steady_state(self._h, float(alpha), out.ctypes.data, self.replications))
        return out

    def stat_partial(self):
        s = C.c_double()
        n = C.c_uint64()
        _check(load_library().simb_stat_partial(self._h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def stat_partial_sqdev(self, mean: float) -> float:
        s = C.c_double()
        _check(load_library().simb_stat_partial_sqdev(self._h, float(mean), C.byref(s)))
        return s.value

    def independent_sample_ci(self, alpha: float) -> dict:
        """``IndependentSample::post(replication means).confidence_interval_mean(alpha)`` (mod.rs:94-125)."""
        ci = _CI()
        _check(load_library().simb_independent_sample_ci(self._h, float(alpha), C.byref(ci)))
        return {k: getattr(ci, k) for k, _ in _CI._fields_}

    def describe(self) -> dict:
        buf = C.create_string_buffer(2048)
        _check(load_library().simb_describe(self._h, buf, 2048))
        return json.loads(buf.value.decode())


class WebSimulation:
    """String API of the reference's ``WebSimulation`` (sim/src/simulator/web.rs:16-215) over a batch.

    Every call acts on all ``replications``; the JSON that comes back describes the observed
    ``replication`` (default 0), which must lie inside the traced prefix of the batch."""

    def __init__(self, simulation: Simulation, replication: int = 0):
        self.simulation = simulation
        self.replication = replication

    @classmethod
    def post_json(cls, models: str, connectors: str, replications: int = 1, *, replication: int = 0,
                  trace_capacity: int = 1 << 16, **kw) -> "WebSimulation":
        """``WebSimulation::post_json`` (web.rs:23-31)."""
        kw.setdefault("trace_replications", replication + 1)
        sim = Simulation.post_json(models, connectors, replications, instrument=True, trace_capacity=trace_capacity, **kw)
        return cls(sim, replication)

    def put_json(self, models: str, connectors: str):
        """``put_json`` (web.rs:35-40)."""
        _check(load_library().simb_put_json(self.simulation._h, models.encode(), connectors.encode()))

    def _text(self, fn, *args) -> str:
        out = C.c_char_p()
        _check(fn(self.simulation._h, *args, C.byref(out)))
        return out.value.decode()

    def get_messages_json(self) -> str:
        """``get_messages_json`` (web.rs:87-89)."""
        return self._text(load_library().simb_get_messages_json, self.replication)

    def get_global_time(self) -> float:
        """``get_global_time`` (web.rs:98-100)."""
        return float(self.simulation.get_global_time()[self.replication])

    def get_status(self, model_id: str) -> str:
        """``get_status`` (web.rs:103-105)."""
        return self.simulation.get_status(model_id, self.replication)

    def get_records_json(self, model_id: str) -> str:
        """``get_records_json`` (web.rs:109-111)."""
        return self._text(load_library().simb_get_records_json, model_id.encode(), self.replication)

    def reset(self):
        self.simulation.reset()

    def reset_messages(self):
        self.simulation.reset_messages()

    def reset_global_time(self):
        self.simulation.reset_global_time()

    def inject_input_json(self, message: str):
        """``inject_input_json`` (web.rs:136-139)."""
        _check(load_library().simb_inject_input_json(self.simulation._h, message.encode(), None))

    def step_json(self) -> str:
        """``step_json`` (web.rs:161-163)."""
        return self._text(load_library().simb_step_json, self.replication)

    def step_until_json(self, until: float) -> str:
        """``step_until_json`` (web.rs:184-186)."""
        return self._text(load_library().simb_step_until_json, C.c_double(until), self.replication)

    def step_n_json(self, n: int) -> str:
        """``step_n_json`` (web.rs:207-209)."""
        return self._text(load_library().simb_step_n_json, C.c_uint64(n), self.replication)


    def get_json(self) -> str:
        """``get_json`` (web.rs:43-45): the whole Simulation of the observed replication, pretty-printed."""
        return self._text(load_library().simb_get_json, self.replication)

    def get_yaml(self) -> str:
        """``get_yaml`` (web.rs:69-71)."""
        return self._text(load_library().simb_get_yaml, self.replication)

    # ---- YAML forms (web.rs:49-70, 93-95, 115-117, 143-146, 167-169, 190-192, 213-215): the library's own YAML
    # reader / writer (csrc/yaml.hpp) behind simb_post_yaml / simb_put_yaml / simb_get_yaml and the two converters.
    @classmethod
    def post_yaml(cls, models: str, connectors: str, replications: int = 1, *, replication: int = 0,
                  trace_capacity: int = 1 << 16, **kw) -> "WebSimulation":
        """``post_yaml`` (web.rs:56-64) through simb_post_yaml."""
        kw.setdefault("trace_replications", replication + 1)
        sim = Simulation.post_json(models, connectors, replications, instrument=True, trace_capacity=trace_capacity,
                                   _yaml=True, **kw)
        return cls(sim, replication)

    def put_yaml(self, models: str, connectors: str):
        """``put_yaml`` (web.rs:66-67)."""
        _check(load_library().simb_put_yaml(self.simulation._h, models.encode(), connectors.encode()))

    def get_messages_yaml(self) -> str:
        return _json_to_yaml(self.get_messages_json())

    def get_records_yaml(self, model_id: str) -> str:
        return _json_to_yaml(self.get_records_json(model_id))

    def inject_input_yaml(self, message: str):
        self.inject_input_json(_yaml_to_json(message))

    def step_yaml(self) -> str:
        return _json_to_yaml(self.step_json())

    def step_until_yaml(self, until: float) -> str:
        return _json_to_yaml(self.step_until_json(until))

    def step_n_yaml(self, n: int) -> str:
        return _json_to_yaml(self.step_n_json(n))


def _convert(fn, text: str) -> str:
    need = C.c_uint64()
    _check(fn(text.encode(), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(fn(text.encode(), buf, need.value, None))
    return buf.value.decode()


def _yaml_to_json(text: str) -> str:
    """YAML text -> compact JSON text (simb_yaml_to_json)."""
    return _convert(load_library().simb_yaml_to_json, text)


def _json_to_yaml(text: str) -> str:
    """JSON text -> YAML text in serde_yaml 0.8 layout (simb_json_to_yaml)."""
    return _convert(load_library().simb_json_to_yaml, text)


def format_f64(value: float) -> str:
    """serde_json's text form of an f64 (the formatter behind the ``*_json`` calls)."""
    buf = C.create_string_buffer(64)
    _check(load_library().simb_format_f64(C.c_double(value), buf, 64))
    return buf.value.decode()


def device_count() -> int:
    n = C.c_int()
    st = load_library().simb_device_count(C.byref(n))
    return n.value if st == 0 else 0
