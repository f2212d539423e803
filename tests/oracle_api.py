"""ctypes wrapper over oracle/liboracle.so (the CPU restatement of the reference; test
infrastructure).  Accepts the same sim_b200 model/connector descriptor objects the product takes."""
import ctypes as C
import os

import numpy as np

from sim_b200 import models as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

DIST_KIND = {"exp": 0, "normal": 1, "triangular": 2, "uniform": 3, "gamma": 4, "logNormal": 5, "weibull": 6, "beta": 7}


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.orc_new.restype = C.c_void_p
        L.orc_global_time.restype = C.c_double
        L.orc_until_next_event.restype = C.c_double
        L.orc_records_count.restype = C.c_int64
        for f in ("orc_rng_draws", "orc_returned_count", "orc_messages_count", "orc_trace_hash", "orc_trace_count",
                  "orc_connector_messages", "orc_usize_sqrt"):
            getattr(L, f).restype = C.c_uint64
        L.orc_trace_intern.restype = C.c_uint32
        L.orc_action_name.restype = C.c_char_p
        _lib = L
    return _lib


def _b(s):
    return s.encode("utf-8")


def _dist(v):
    p = v.params
    k = v.variant
    if k == "exp":
        return DIST_KIND[k], p["lambda"], 0.0, 0.0
    if k == "normal":
        return DIST_KIND[k], p["mean"], p["std_dev"], 0.0
    if k == "triangular":
        return DIST_KIND[k], p["min"], p["max"], p["mode"]
    if k == "uniform":
        return DIST_KIND[k], p["min"], p["max"], 0.0
    if k == "gamma":
        return DIST_KIND[k], p["shape"], p["scale"], 0.0
    if k == "logNormal":
        return DIST_KIND[k], p["mu"], p["sigma"], 0.0
    if k == "weibull":
        return DIST_KIND[k], p["shape"], p["scale"], 0.0
    if k == "beta":
        return DIST_KIND[k], p["alpha"], p["beta"], 0.0
    raise ValueError(k)


class OracleSimulation:
    """Mirrors Simulation::post/put/reset/inject_input/step/step_n/step_until on the oracle."""

    def __init__(self, models, connectors):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_new())
        self.models = list(models)
        self.connectors = list(connectors)
        for m in models:
            self._add(m)
        for c in connectors:
            self.L.orc_add_connector(self.h, _b(c.id), _b(c.source_id), _b(c.target_id), _b(c.source_port),
                                     _b(c.target_port))
        self.L.orc_post(self.h)

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def _add(self, model):
        L, h, i, b = self.L, self.h, model.inner, _b(model.id)
        d = C.c_double
        if isinstance(i, M.Generator):
            k, a, bb, c = _dist(i.message_interdeparture_time)
            L.orc_add_generator(h, b, k, d(a), d(bb), d(c), _b(i.job_port), int(i.store_records))
        elif isinstance(i, M.Processor):
            k, a, bb, c = _dist(i.service_time)
            cap = -1 if i.queue_capacity is None else int(i.queue_capacity)
            L.orc_add_processor(h, b, k, d(a), d(bb), d(c), C.c_int64(cap), _b(i.job_port), _b(i.processed_job_port),
                                int(i.store_records))
            if i.state is not None:
                st = i.state
                until = st.get("untilNextEvent", float("inf"))
                until = float("inf") if until is None else float(until)
                L.orc_preload_processor(h, b, 1 if st.get("phase") == "Active" else 0, d(until),
                                        _b("\n".join(st.get("queue", []))))
        elif isinstance(i, M.Storage):
            L.orc_add_storage(h, b, _b(i.put_port), _b(i.get_port), _b(i.stored_port), int(i.store_records))
        elif isinstance(i, M.LoadBalancer):
            L.orc_add_load_balancer(h, b, _b(i.job_port), _b("\n".join(i.flow_path_ports)), int(i.store_records))
        elif isinstance(i, M.Gate):
            L.orc_add_gate(h, b, _b(i.job_in_port), _b(i.activation_port), _b(i.deactivation_port), _b(i.job_out_port),
                           int(i.store_records))
        elif isinstance(i, M.ExclusiveGateway):
            pw = i.port_weights
            if pw.variant == "weightedIndex":
                w = (C.c_uint64 * len(pw.params["weights"]))(*pw.params["weights"])
                L.orc_add_exclusive_gateway(h, b, _b("\n".join(i.flow_paths_in)), _b("\n".join(i.flow_paths_out)), 1,
                                            C.c_uint64(0), C.c_uint64(0), w, len(pw.params["weights"]),
                                            int(i.store_records))
            else:
                L.orc_add_exclusive_gateway(h, b, _b("\n".join(i.flow_paths_in)), _b("\n".join(i.flow_paths_out)), 0,
                                            C.c_uint64(pw.params["min"]), C.c_uint64(pw.params["max"]), None, 0,
                                            int(i.store_records))
        elif isinstance(i, M.ParallelGateway):
            L.orc_add_parallel_gateway(h, b, _b("\n".join(i.flow_paths_in)), _b("\n".join(i.flow_paths_out)),
                                       int(i.store_records))
        elif isinstance(i, M.StochasticGate):
            L.orc_add_stochastic_gate(h, b, d(i.pass_distribution.params["p"]), _b(i.job_in_port), _b(i.job_out_port),
                                      int(i.store_records))
        elif isinstance(i, M.Stopwatch):
            L.orc_add_stopwatch(h, b, _b(i.start_port), _b(i.stop_port), _b(i.metric_port), _b(i.job_port),
                                1 if i.metric == M.Metric.Maximum else 0, int(i.store_records))
        elif isinstance(i, M.Batcher):
            L.orc_add_batcher(h, b, _b(i.job_in_port), _b(i.job_out_port), d(i.max_batch_time),
                              C.c_uint64(i.max_batch_size), int(i.store_records))
        elif isinstance(i, M.Passive):
            L.orc_add_passive(h, b, _b(i.job_port))
        elif isinstance(i, M.Coupled):
            L.orc_begin_coupled(h, b, _b("\n".join(i.ports_in)), _b("\n".join(i.ports_out)))
            for comp in i.components:
                self._add(comp)
            for c in i.external_input_couplings:
                L.orc_coupled_eic(h, _b(c.target_id), _b(c.source_port), _b(c.target_port))
            for c in i.external_output_couplings:
                L.orc_coupled_eoc(h, _b(c.source_id), _b(c.source_port), _b(c.target_port))
            for c in i.internal_couplings:
                L.orc_coupled_ic(h, _b(c.source_id), _b(c.target_id), _b(c.source_port), _b(c.target_port))
            assert L.orc_end_coupled(h) == 0
        else:
            raise TypeError(type(i))
        if getattr(i, "rng", None) is not None:
            assert L.orc_set_model_rng(h, b, C.c_uint64(int(i.rng))) == 0
        st = getattr(i, "state", None)
        if st is not None and isinstance(i, (M.ParallelGateway, M.StochasticGate, M.Stopwatch)):
            until = st.get("untilNextEvent", float("inf"))
            until = float("inf") if until is None else float(until)
            if isinstance(i, M.ParallelGateway):  # parallel_gateway.rs:47-53
                col = st.get("collections", {})
                cnt = (C.c_uint64 * max(1, len(col)))(*col.values())
                e = L.orc_preload_parallel_gateway(h, b, d(until), _b("\n".join(col.keys())), cnt)
            elif isinstance(i, M.StochasticGate):  # stochastic_gate.rs:50-64
                jobs = st.get("jobs", [])
                flags = (C.c_uint8 * max(1, len(jobs)))(*[1 if j["pass"] else 0 for j in jobs])
                e = L.orc_preload_stochastic_gate(h, b, d(until), _b("\n".join(j["content"] for j in jobs)), flags)
            else:  # stopwatch.rs:66-93
                jobs = st.get("jobs", [])
                nan = float("nan")
                a = (C.c_double * max(1, len(jobs)))(*[nan if j.get("start") is None else j["start"] for j in jobs])
                z = (C.c_double * max(1, len(jobs)))(*[nan if j.get("stop") is None else j["stop"] for j in jobs])
                e = L.orc_preload_stopwatch(h, b, 1 if st.get("phase") == "JobFetch" else 0, d(until),
                                            _b("\n".join(j["name"] for j in jobs)), a, z)
            assert e == 0, e
        elif st is not None and not isinstance(i, M.Processor):
            phases = {M.Gate: ["Open", "Closed", "Pass"], M.Batcher: ["Passive", "Batching", "Release"],
                      M.ExclusiveGateway: ["Passive", "Pass"], M.LoadBalancer: ["Passive", "LoadBalancing"]}[type(i)]
            until = st.get("untilNextEvent", float("inf"))
            until = float("inf") if until is None else float(until)
            e = L.orc_preload_state(h, b, phases.index(st.get("phase", phases[0])), d(until), _b("\n".join(st.get("jobs", []))))
            assert e == 0, e

    # ---- rng ----
    def set_rng_pcg(self, seed=42):
        self.L.orc_set_rng_pcg(self.h, C.c_uint64(seed & (2**64 - 1)), C.c_uint64(seed >> 64))

    def set_rng_philox(self, seed, replication):
        self.L.orc_set_rng_philox(self.h, C.c_uint64(seed), C.c_uint64(replication))

    def set_rng_external(self, values):
        a = np.ascontiguousarray(values, dtype=np.uint64)
        self.L.orc_set_rng_external(self.h, a.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(a.size))

    def rng_draws(self):
        return int(self.L.orc_rng_draws(self.h))

    # ---- lifecycle / stepping ----
    def put(self):
        self.L.orc_post(self.h)

    def reset(self):
        self.L.orc_reset(self.h)

    def inject_input(self, target_id, target_port, content):
        self.L.orc_inject(self.h, _b(target_id), _b(target_port), _b(content))

    def _msgs(self, count_fn, get_fn):
        n = int(count_fn(self.h))
        out = []
        bufs = [C.create_string_buffer(128) for _ in range(5)]
        t = C.c_double()
        for i in range(n):
            get_fn(self.h, C.c_uint64(i), *bufs, 128, C.byref(t))
            out.append(dict(source_id=bufs[0].value.decode(), source_port=bufs[1].value.decode(),
                            target_id=bufs[2].value.decode(), target_port=bufs[3].value.decode(),
                            content=bufs[4].value.decode(), time=t.value))
        return out

    def step(self):
        e = self.L.orc_step(self.h)
        return e, self._msgs(self.L.orc_returned_count, self.L.orc_returned_get)

    def step_n(self, n, collect=True):
        e = self.L.orc_step_n(self.h, C.c_uint64(n), int(collect))
        return e, (self._msgs(self.L.orc_returned_count, self.L.orc_returned_get) if collect else [])

    def step_until(self, until, collect=True):
        e = self.L.orc_step_until(self.h, C.c_double(until), int(collect))
        return e, (self._msgs(self.L.orc_returned_count, self.L.orc_returned_get) if collect else [])

    def get_messages(self):
        return self._msgs(self.L.orc_messages_count, self.L.orc_messages_get)

    def get_global_time(self):
        return float(self.L.orc_global_time(self.h))

    def until_next_event(self, model_id):
        return float(self.L.orc_until_next_event(self.h, _b(model_id)))

    def get_status(self, model_id):
        buf = C.create_string_buffer(256)
        e = self.L.orc_status(self.h, _b(model_id), buf, 256)
        return e, buf.value.decode()

    def get_records(self, model_id):
        n = int(self.L.orc_records_count(self.h, _b(model_id)))
        if n < 0:
            return -n, []
        out = []
        t = C.c_double()
        a = C.create_string_buffer(64)
        s = C.create_string_buffer(256)
        for i in range(n):
            self.L.orc_record_get(self.h, _b(model_id), C.c_uint64(i), C.byref(t), a, s, 256)
            out.append((t.value, a.value.decode(), s.value.decode()))
        return 0, out

    # ---- instrumentation ----
    def trace_enable(self, store=True):
        self.L.orc_trace_enable(self.h, int(store))

    def trace_intern(self, content):
        return int(self.L.orc_trace_intern(self.h, _b(content)))

    def trace_hash(self):
        return int(self.L.orc_trace_hash(self.h))

    def trace(self):
        n = int(self.L.orc_trace_count(self.h))
        out = []
        f = (C.c_uint32 * 6)()
        t = C.c_double()
        buf = C.create_string_buffer(256)
        for i in range(n):
            if self.L.orc_trace_get(self.h, C.c_uint64(i), f, C.byref(t)) != 0:
                break
            self.L.orc_trace_render(self.h, C.c_uint32(f[5]), buf, 256)
            out.append(dict(step=f[0], kind=f[1], index=f[2], action=f[3], aux=f[4], content=buf.value.decode(),
                            time=t.value))
        return out

    def counters(self):
        c = (C.c_uint64 * 3)()
        self.L.orc_counters(self.h, c)
        return dict(steps=c[0], ext=c[1], int=c[2], events=c[1] + c[2])

    def connector_messages(self):
        return [int(self.L.orc_connector_messages(self.h, C.c_uint64(i))) for i in range(len(self.connectors))]

    def mean_wait(self, model_id):
        s = C.c_double()
        n = C.c_uint64()
        self.L.orc_mean_wait(self.h, _b(model_id), C.byref(s), C.byref(n))
        return s.value, n.value

    def run_batch(self, seed, rep0, nreps, threads=1, until=None, nsteps=None, stat_model_id=None, trace_hash=True):
        """Replications rep0..rep0+nreps on Philox(seed, r): returns dict of per-replication arrays."""
        t = np.zeros(nreps, np.float64)
        hsh = np.zeros(nreps, np.uint64)
        st = np.zeros(nreps, np.uint64)
        ev = np.zeros(nreps, np.uint64)
        ws = np.zeros(nreps, np.float64)
        wn = np.zeros(nreps, np.uint64)
        tot = (C.c_uint64 * 3)()
        conn = np.zeros(max(1, len(self.connectors)), np.uint64)
        recs = np.zeros((len(self.models), 14), np.uint64)
        p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
        e = self.L.orc_run_batch_counts(self.h, C.c_uint64(seed), C.c_uint64(rep0), C.c_uint64(nreps), int(threads),
                                 0 if until is not None else 1, C.c_double(until or 0.0), C.c_uint64(nsteps or 0),
                                 _b(stat_model_id) if stat_model_id else None, int(trace_hash), tot,
                                 p(t, C.c_double), p(hsh, C.c_uint64), p(st, C.c_uint64), p(ev, C.c_uint64),
                                 p(ws, C.c_double), p(wn, C.c_uint64), p(conn, C.c_uint64), p(recs, C.c_uint64))
        return dict(err=e, time=t, hash=hsh, steps=st, events=ev, wait_sum=ws, wait_n=wn,
                    conn=conn[:len(self.connectors)], records=recs,
                    totals=dict(steps=tot[0], ext=tot[1], int=tot[2]))


# ---- output_analysis / sampler probes ----
def independent_sample(points, alpha):
    a = np.ascontiguousarray(points, dtype=np.float64)
    out = (C.c_double * 4)()
    e = lib().orc_independent_sample(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(a.size), C.c_double(alpha), out)
    return e, dict(mean=out[0], variance=out[1], lower=out[2], upper=out[3])


def steady_state(series, alpha):
    a = np.ascontiguousarray(series, dtype=np.float64)
    out = (C.c_double * 4)()
    outn = (C.c_uint64 * 3)()
    e = lib().orc_steady_state(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(a.size), C.c_double(alpha), out, outn)
    return e, dict(mean=out[0], variance=out[1], lower=out[2], upper=out[3], deletion_point=outn[0],
                   batch_size=outn[1], batch_count=outn[2])


def t_score(alpha, df):
    out = C.c_double()
    e = lib().orc_t_score(C.c_double(alpha), C.c_uint64(df), C.byref(out))
    return e, out.value


def usize_sqrt(n):
    return int(lib().orc_usize_sqrt(C.c_uint64(n)))


def pcg_fill(seed, n):
    out = np.zeros(n, np.uint64)
    lib().orc_pcg_fill(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(seed >> 64), out.ctypes.data_as(C.POINTER(C.c_uint64)),
                       C.c_uint64(n))
    return out


def philox_block(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox_block(c, k, o)
    return list(o)


def philox_fill(seed, rep, n):
    out = np.zeros(n, np.uint64)
    lib().orc_philox_fill(C.c_uint64(seed), C.c_uint64(rep), out.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(n))
    return out


def sample_continuous(variant, n, rng_kind=0, seed=42, rep=0):
    k, a, b, c = _dist(variant)
    out = np.zeros(n, np.float64)
    used = C.c_uint64()
    e = lib().orc_sample_continuous(k, C.c_double(a), C.c_double(b), C.c_double(c), rng_kind, C.c_uint64(seed),
                                    C.c_uint64(rep), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(n),
                                    C.byref(used))
    return e, out, used.value


def sample_bernoulli(p, n, rng_kind=0, seed=42, rep=0):
    out = np.zeros(n, np.uint8)
    used = C.c_uint64()
    e = lib().orc_sample_bernoulli(C.c_double(p), rng_kind, C.c_uint64(seed), C.c_uint64(rep),
                                   out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(n), C.byref(used))
    return e, out, used.value


def sample_index(variant, n, rng_kind=0, seed=42, rep=0):
    out = np.zeros(n, np.uint64)
    used = C.c_uint64()
    if variant.variant == "weightedIndex":
        w = (C.c_uint64 * len(variant.params["weights"]))(*variant.params["weights"])
        e = lib().orc_sample_index(1, C.c_uint64(0), C.c_uint64(0), w, len(variant.params["weights"]), rng_kind,
                                   C.c_uint64(seed), C.c_uint64(rep), out.ctypes.data_as(C.POINTER(C.c_uint64)),
                                   C.c_uint64(n), C.byref(used))
    else:
        e = lib().orc_sample_index(0, C.c_uint64(variant.params["min"]), C.c_uint64(variant.params["max"]), None, 0,
                                   rng_kind, C.c_uint64(seed), C.c_uint64(rep),
                                   out.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(n), C.byref(used))
    return e, out, used.value


def zig_table(which):
    out = np.zeros(257, np.float64)
    lib().orc_zig_table(which, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def sample_discrete(kind, n, p=0.0, mn=0, mx=1, rng_kind=0, seed=42, rep=0):
    """kind: 0 Geometric(p), 1 Poisson(lambda = p), 2 Uniform[mn, mx)"""
    out = np.zeros(n, np.uint64)
    used = C.c_uint64()
    e = lib().orc_sample_discrete(kind, C.c_double(p), C.c_uint64(mn), C.c_uint64(mx), rng_kind, C.c_uint64(seed),
                                  C.c_uint64(rep), out.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(n), C.byref(used))
    return e, out, used.value


def thinning_evaluate(coefficients, point):
    a = (C.c_double * max(1, len(coefficients)))(*coefficients)
    out = C.c_double()
    lib().orc_thinning_evaluate(a, C.c_uint64(len(coefficients)), C.c_double(point), C.byref(out))
    return out.value
