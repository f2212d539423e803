"""Test networks.  The first group restates the networks of the reference's own tests
(sim/tests/simulations.rs, sim/tests/web.rs, sim/tests/custom.rs; cited per function); the second
group are the BASELINE.json configurations with the concrete shapes of SURVEY.md 8(d)."""
from sim_b200 import (Batcher, BooleanRandomVariable, Connector, ContinuousRandomVariable, Coupled, ExclusiveGateway,
                      ExternalInputCoupling, ExternalOutputCoupling, Gate, Generator, IndexRandomVariable, InternalCoupling,
                      LoadBalancer, Metric, Model, ParallelGateway, Passive, Processor, StochasticGate, Stopwatch, Storage)

Exp = ContinuousRandomVariable.Exp


def _storage(i, store=False):
    return Model.new(i, Storage.new("store", "read", "stored", store))


def mm1(store_records=False, capacity=14, lam_gen=0.5, lam_proc=0.333333):
    """simulations.rs:20-72 (BASELINE config 1/2): Generator Exp(0.5) -> Processor Exp(0.333333) cap 14 -> Storage."""
    models = [
        Model.new("generator-01", Generator.new(Exp(lam_gen), None, "job", store_records, None)),
        Model.new("processor-01", Processor.new(Exp(lam_proc), capacity, "job", "processed", store_records, None)),
        _storage("storage-01", store_records),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "processor-01", "job", "job"),
        Connector.new("connector-02", "processor-01", "storage-01", "processed", "store"),
    ]
    return models, connectors


def generator_storage():
    """simulations.rs:132-160 step_until_activities"""
    models = [Model.new("generator-01", Generator.new(Exp(0.5), None, "job", False, None)), _storage("storage-01")]
    connectors = [Connector.new("connector-01", "generator-01", "storage-01", "job", "store")]
    return models, connectors


def exclusive_gateway(store_records=False):
    """simulations.rs:258-345"""
    models = [
        Model.new("generator-01", Generator.new(Exp(5.0), None, "job", store_records, None)),
        Model.new("exclusive-01", ExclusiveGateway.new(["in"], ["s01", "s02", "s03"],
                                                        IndexRandomVariable.WeightedIndex([6, 3, 1]), store_records, None)),
        _storage("storage-01", store_records), _storage("storage-02", store_records), _storage("storage-03", store_records),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "exclusive-01", "job", "in"),
        Connector.new("connector-02", "exclusive-01", "storage-01", "s01", "store"),
        Connector.new("connector-03", "exclusive-01", "storage-02", "s02", "store"),
        Connector.new("connector-04", "exclusive-01", "storage-03", "s03", "store"),
    ]
    return models, connectors


def gate_blocking(store_records=False):
    """simulations.rs:382-463"""
    models = [
        Model.new("generator-01", Generator.new(Exp(10.0), None, "job", False, None)),
        Model.new("generator-02", Generator.new(Exp(10.0), None, "job", False, None)),
        Model.new("generator-03", Generator.new(Exp(1.0), None, "job", False, None)),
        Model.new("gate-01", Gate.new("job", "activation", "deactivation", "job", store_records)),
        _storage("storage-01"),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "gate-01", "job", "activation"),
        Connector.new("connector-02", "generator-02", "gate-01", "job", "deactivation"),
        Connector.new("connector-03", "generator-03", "gate-01", "job", "job"),
        Connector.new("connector-04", "gate-01", "storage-01", "job", "store"),
    ]
    return models, connectors


def load_balancer(store_records=False):
    """simulations.rs:499-584"""
    models = [
        Model.new("generator-01", Generator.new(Exp(0.01), None, "job", False, None)),
        Model.new("load-balancer-01", LoadBalancer.new("request", ["server-1", "server-2", "server-3"], store_records)),
        _storage("storage-01"), _storage("storage-02"), _storage("storage-03"),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "load-balancer-01", "job", "request"),
        Connector.new("connector-02", "load-balancer-01", "storage-01", "server-1", "store"),
        Connector.new("connector-03", "load-balancer-01", "storage-02", "server-2", "store"),
        Connector.new("connector-04", "load-balancer-01", "storage-03", "server-3", "store"),
    ]
    return models, connectors


def storage_exchange():
    """simulations.rs:608-638 injection_initiated_stored_value_exchange"""
    models = [_storage("storage-01"), _storage("storage-02")]
    connectors = [
        Connector.new("connector-01", "storage-02", "storage-01", "stored", "store"),
        Connector.new("connector-02", "storage-01", "storage-02", "stored", "store"),
    ]
    return models, connectors


def parallel_gateway(store_records=False):
    """simulations.rs:681-763"""
    models = [
        Model.new("generator-01", Generator.new(Exp(5.0), None, "job", False, None)),
        Model.new("parallel-01", ParallelGateway.new(["in"], ["alpha", "beta", "delta"], store_records)),
        Model.new("parallel-02", ParallelGateway.new(["alpha", "beta", "delta"], ["out"], store_records)),
        _storage("storage-01"),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "parallel-01", "job", "in"),
        Connector.new("connector-02", "parallel-01", "parallel-02", "alpha", "alpha"),
        Connector.new("connector-03", "parallel-01", "parallel-02", "beta", "beta"),
        Connector.new("connector-04", "parallel-01", "parallel-02", "delta", "delta"),
        Connector.new("connector-05", "parallel-02", "storage-01", "out", "store"),
    ]
    return models, connectors


def status_pair():
    """simulations.rs:790-823 match_status_reporting"""
    models = [
        Model.new("generator-01", Generator.new(Exp(5.0), None, "job", False, None)),
        Model.new("load-balancer-01", LoadBalancer.new("request", ["alpha", "beta", "delta"], False)),
    ]
    return models, []


def stochastic_gate(store_records=False, p=0.2):
    """simulations.rs:826-872"""
    models = [
        Model.new("generator-01", Generator.new(Exp(5.0), None, "job", False, None)),
        Model.new("stochastic-gate-01", StochasticGate.new(BooleanRandomVariable.Bernoulli(p), "job", "job", store_records, None)),
        _storage("storage-01"),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "stochastic-gate-01", "job", "job"),
        Connector.new("connector-02", "stochastic-gate-01", "storage-01", "job", "store"),
    ]
    return models, connectors


def batcher(store_records=False, max_time=10.0, max_size=10, lam=1.0):
    """simulations.rs:896-939"""
    models = [
        Model.new("generator-01", Generator.new(Exp(lam), None, "job", False, None)),
        Model.new("batcher-01", Batcher.new("job", "job", max_time, max_size, store_records)),
        _storage("storage-01"),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "batcher-01", "job", "job"),
        Connector.new("connector-02", "batcher-01", "storage-01", "job", "store"),
    ]
    return models, connectors


def stopwatch(store_records=False):
    """simulations.rs:967-1070 (note connector-04 names a port the processor does not have: it never fires)"""
    models = [
        Model.new("generator-01", Generator.new(Exp(0.01), None, "job", False, None)),
        Model.new("processor-01", Processor.new(Exp(0.333333), 14, "job", "processed", False, None)),
        _storage("storage-01"),
        Model.new("stopwatch-01", Stopwatch.new("start", "stop", "min", "min", Metric.Minimum, store_records)),
        Model.new("stopwatch-02", Stopwatch.new("start", "stop", "max", "max", Metric.Maximum, store_records)),
    ]
    connectors = [
        Connector.new("connector-01", "generator-01", "processor-01", "job", "job"),
        Connector.new("connector-02", "generator-01", "stopwatch-01", "job", "start"),
        Connector.new("connector-03", "generator-01", "stopwatch-02", "job", "start"),
        Connector.new("connector-04", "processor-01", "storage-01", "job", "store"),
        Connector.new("connector-05", "processor-01", "stopwatch-01", "processed", "stop"),
        Connector.new("connector-06", "processor-01", "stopwatch-02", "processed", "stop"),
        Connector.new("connector-07", "stopwatch-01", "storage-01", "min", "store"),
        Connector.new("connector-06", "stopwatch-02", "storage-01", "max", "store"),
    ]
    return models, connectors


def processor_from_queue(n_jobs=100):
    """web.rs:13-69: Processor preloaded through the JSON `state` block -> Storage"""
    state = {"phase": "Passive", "untilNextEvent": 0.0, "queue": ["job %d" % i for i in range(1, n_jobs + 1)],
             "records": []}
    proc = Processor.new(Exp(3.0), None, "job", "processed job", False, None)
    proc.state = state
    models = [Model.new("processor-01", proc), _storage("storage-01")]
    connectors = [Connector.new("connector-01", "processor-01", "storage-01", "processed job", "store")]
    return models, connectors


def processor_network():
    """web.rs:130-302: six processors, fan-out of one out-port over two connectors"""
    lams = [1.0, 2.0, 3.0, 5.0, 7.0, 11.0]
    models = []
    for i, lam in enumerate(lams):
        p = Processor.new(Exp(lam), None, "job", "processed job", False, None)
        if i == 0:
            p.state = {"phase": "Passive", "untilNextEvent": 0.0, "untilJobCompletion": 0.0,
                       "queue": ["job %d" % k for k in range(10)], "records": []}
        models.append(Model.new("processor-%d" % i, p))
    models.append(Model.new("storage-0", Storage.new("store", "read", "stored", False)))
    C = Connector.new
    connectors = [
        C("connector-01", "processor-0", "processor-1", "processed job", "job"),
        C("connector-02", "processor-1", "processor-4", "processed job", "job"),
        C("connector-03", "processor-4", "storage-0", "processed job", "store"),
        C("connector-04", "processor-1", "processor-3", "processed job", "job"),
        C("connector-05", "processor-0", "processor-2", "processed job", "job"),
        C("connector-06", "processor-2", "processor-5", "processed job", "job"),
        C("connector-07", "processor-3", "processor-5", "processed job", "job"),
        C("connector-08", "processor-5", "storage-0", "processed job", "store"),
    ]
    return models, connectors


def generator_passive():
    """custom.rs:89-112"""
    models = [Model.new("generator-01", Generator.new(Exp(0.5), None, "job", False, None)),
              Model.new("passive-01", Passive.new("job"))]
    connectors = [Connector.new("connector-01", "generator-01", "passive-01", "job", "job")]
    return models, connectors


def more_variables():
    """the remaining continuous variables of random_variable.rs:19-27 that the engine
    accelerates -- Gamma (large and small shape, and shape == 1 which rand_distr turns into an Exp),
    LogNormal and Weibull (with the reference's swapped-parameter quirk)."""
    T = ContinuousRandomVariable
    models = [
        Model.new("gen", Generator.new(T.Gamma(2.5, 0.8), None, "job", False, None)),
        Model.new("p-lognormal", Processor.new(T.LogNormal(-0.2, 0.5), 12, "job", "done", False, None)),
        Model.new("p-weibull", Processor.new(T.Weibull(1.2, 2.0), 12, "job", "done", False, None)),
        Model.new("p-gamma-small", Processor.new(T.Gamma(0.6, 1.5), 12, "job", "done", False, None)),
        Model.new("p-gamma-one", Processor.new(T.Gamma(1.0, 0.7), 12, "job", "done", False, None)),
        _storage("sink"),
    ]
    C = Connector.new
    connectors = [
        C("c1", "gen", "p-lognormal", "job", "job"), C("c2", "p-lognormal", "p-weibull", "done", "job"),
        C("c3", "p-weibull", "p-gamma-small", "done", "job"), C("c4", "p-gamma-small", "p-gamma-one", "done", "job"),
        C("c5", "p-gamma-one", "sink", "done", "store"),
    ]
    return models, connectors


def beta_variables():
    """Continuous::Beta (random_variable.rs:20,72) through every branch of rand_distr 0.4.3's sampler:
    Cheng's BB (min(alpha, beta) > 1) and BC, each with and without switched parameters."""
    T = ContinuousRandomVariable
    models = [
        Model.new("gen", Generator.new(T.Beta(2.0, 5.0), None, "job", False, None)),
        Model.new("p-bb-switched", Processor.new(T.Beta(5.0, 2.0), 12, "job", "done", False, None)),
        Model.new("p-bc", Processor.new(T.Beta(0.5, 0.7), 12, "job", "done", False, None)),
        Model.new("p-bc-switched", Processor.new(T.Beta(3.0, 0.8), 12, "job", "done", False, None)),
        Model.new("p-uniform", Processor.new(T.Beta(1.0, 1.0), 12, "job", "done", False, None)),
        _storage("sink"),
    ]
    C = Connector.new
    connectors = [
        C("c1", "gen", "p-bb-switched", "job", "job"), C("c2", "p-bb-switched", "p-bc", "done", "job"),
        C("c3", "p-bc", "p-bc-switched", "done", "job"), C("c4", "p-bc-switched", "p-uniform", "done", "job"),
        C("c5", "p-uniform", "sink", "done", "store"),
    ]
    return models, connectors


def state_injection(store_records=False):
    """`state` blocks injected on input (the models' `#[serde(default)] state`, as tests/web.rs:30-46 does for a
    Processor) for the other queueing models: a Gate, an ExclusiveGateway and a LoadBalancer that start with
    jobs to pass, a Batcher that starts with a half-built batch and a running timer."""
    def with_state(body, state):
        body.state = state
        return body
    models = [
        Model.new("generator", Generator.new(Exp(1.0), None, "job", store_records, None)),
        Model.new("gate", with_state(Gate.new("job", "activation", "deactivation", "job", store_records),
                                     {"phase": "Pass", "untilNextEvent": 0.0, "jobs": ["pre 1", "pre 2"], "records": []})),
        Model.new("exclusive", with_state(ExclusiveGateway.new(["in"], ["a", "b"], IndexRandomVariable.WeightedIndex([1, 1]),
                                                               store_records, None),
                                          {"phase": "Pass", "untilNextEvent": 0.0, "jobs": ["x 1", "x 2", "x 3"], "records": []})),
        Model.new("balancer", with_state(LoadBalancer.new("job", ["p", "q"], store_records),
                                         {"phase": "LoadBalancing", "untilNextEvent": 0.0, "jobs": ["l 1", "l 2"], "records": []})),
        Model.new("batcher-p", with_state(Batcher.new("job", "job", 3.0, 3, store_records),
                                          {"phase": "Batching", "untilNextEvent": 2.5, "jobs": ["b 1", "b 2"], "records": []})),
        Model.new("batcher-q", Batcher.new("job", "job", 3.0, 3, store_records)),
        _storage("storage-p", store_records), _storage("storage-q", store_records),
    ]
    C = Connector.new
    connectors = [
        C("c1", "generator", "gate", "job", "job"), C("c2", "gate", "exclusive", "job", "in"),
        C("c3", "exclusive", "balancer", "a", "job"), C("c4", "exclusive", "balancer", "b", "job"),
        C("c5", "balancer", "batcher-p", "p", "job"), C("c6", "balancer", "batcher-q", "q", "job"),
        C("c7", "batcher-p", "storage-p", "job", "store"), C("c8", "batcher-q", "storage-q", "job", "store"),
    ]
    return models, connectors


def batch_bursts(store_records=False):
    """Bursts larger than the default 16-slot transient queues (the reference's Vecs are unbounded): a Batcher of 20
    feeds a Gate and, through an ExclusiveGateway, a LoadBalancer; each must take a whole batch in one step."""
    models = [
        Model.new("generator", Generator.new(Exp(4.0), None, "job", store_records, None)),
        Model.new("batcher", Batcher.new("job", "job", 50.0, 20, store_records)),
        Model.new("gate", Gate.new("job", "activation", "deactivation", "job", store_records)),
        Model.new("exclusive", ExclusiveGateway.new(["in"], ["a", "b"], IndexRandomVariable.WeightedIndex([3, 1]),
                                                    store_records, None)),
        Model.new("balancer", LoadBalancer.new("job", ["p", "q"], store_records)),
        Model.new("sgate", StochasticGate.new(BooleanRandomVariable.Bernoulli(0.5), "job", "job", store_records, None)),
        _storage("storage-gate", store_records), _storage("storage-p", store_records), _storage("storage-q", store_records),
        _storage("storage-s", store_records),
    ]
    C = Connector.new
    connectors = [
        C("c1", "generator", "batcher", "job", "job"), C("c2", "batcher", "gate", "job", "job"),
        C("c3", "batcher", "exclusive", "job", "in"), C("c4", "gate", "storage-gate", "job", "store"),
        C("c5", "exclusive", "balancer", "a", "job"), C("c6", "exclusive", "sgate", "b", "job"),
        C("c7", "balancer", "storage-p", "p", "store"), C("c8", "balancer", "storage-q", "q", "store"),
        C("c9", "sgate", "storage-s", "job", "store"),
    ]
    return models, connectors


def state_injection_tables(store_records=False):
    """`state` blocks of the three table-keeping models (parallel_gateway.rs:47-53, stochastic_gate.rs:50-73,
    stopwatch.rs:66-93): a join that starts with one collection half built and one complete, a stochastic gate
    that starts with decided jobs queued, a stopwatch that starts with a completed job, one started only, one
    stopped only and one with neither side, and is asked for its metric at once."""
    def with_state(body, state):
        body.state = state
        return body
    inf = float("inf")
    models = [
        Model.new("generator", Generator.new(Exp(1.0), None, "job", store_records, None)),
        Model.new("split", ParallelGateway.new(["in"], ["a", "b"], store_records)),
        Model.new("proc-a", Processor.new(Exp(2.0), None, "job", "done", store_records, None)),
        Model.new("sgate-b", with_state(StochasticGate.new(BooleanRandomVariable.Bernoulli(1.0), "job", "job", store_records, None),
                                        {"untilNextEvent": 0.0, "records": [],
                                         "jobs": [{"content": "pre 1", "pass": True}, {"content": "pre 2", "pass": False},
                                                  {"content": "job 900", "pass": True}]})),
        Model.new("join", with_state(ParallelGateway.new(["a", "b"], ["out"], store_records),
                                     {"untilNextEvent": 0.0, "records": [],
                                      "collections": {"job 900": 1, "whole 1": 2, "pre 1": 1}})),
        Model.new("watch", with_state(Stopwatch.new("start", "stop", "metric", "job", Metric.Minimum, store_records),
                                      {"phase": "JobFetch", "untilNextEvent": 0.0, "records": [],
                                       "jobs": [{"name": "old 1", "start": 0.25, "stop": 2.0},
                                                {"name": "old 2", "start": 1.0, "stop": 1.5},
                                                {"name": "old 3", "start": 3.0, "stop": 3.5},
                                                {"name": "job 1", "start": 0.125, "stop": None},
                                                {"name": "job 2", "start": None, "stop": 0.5},
                                                {"name": "job 3", "start": None, "stop": None}]})),
        Model.new("watch-max", with_state(Stopwatch.new("start", "stop", "metric", "job", Metric.Maximum, store_records),
                                          {"phase": "Passive", "records": [],
                                           "jobs": [{"name": "old 1", "start": 0.0, "stop": 1.0},
                                                    {"name": "old 2", "start": 0.5, "stop": 1.5}]})),
        _storage("sink", store_records), _storage("best", store_records),
    ]
    C = Connector.new
    connectors = [
        C("c1", "generator", "split", "job", "in"), C("c2", "split", "proc-a", "a", "job"),
        C("c3", "split", "sgate-b", "b", "job"), C("c4", "proc-a", "join", "done", "a"),
        C("c5", "sgate-b", "join", "job", "b"), C("c6", "join", "sink", "out", "store"),
        C("c7", "generator", "watch", "job", "start"), C("c8", "proc-a", "watch", "done", "stop"),
        C("c9", "watch", "best", "job", "store"), C("c10", "generator", "watch-max", "job", "start"),
        C("c11", "join", "watch-max", "out", "stop"), C("c12", "sink", "watch-max", "stored", "metric"),
    ]
    return models, connectors


# ---- BASELINE.json configurations (shapes fixed in SURVEY.md 8(d)) -----------------------------------
def tandem():
    """config 3: generator Exp(1.0) -> load balancer (3 paths) -> 3x stage-1 processor Exp(0.4) cap 8 ->
    processor-2 Triangular(0.2,1.4,0.6) cap 8 -> processor-3 Uniform(0.3,0.9) cap 8 -> storage  (9 models)."""
    T = ContinuousRandomVariable
    models = [
        Model.new("generator", Generator.new(Exp(1.0), None, "job", False, None)),
        Model.new("load-balancer", LoadBalancer.new("job", ["a", "b", "c"], False)),
        Model.new("stage1-a", Processor.new(Exp(0.4), 8, "job", "done", False, None)),
        Model.new("stage1-b", Processor.new(Exp(0.4), 8, "job", "done", False, None)),
        Model.new("stage1-c", Processor.new(Exp(0.4), 8, "job", "done", False, None)),
        Model.new("processor-2", Processor.new(T.Triangular(0.2, 1.4, 0.6), 8, "job", "done", False, None)),
        Model.new("processor-3", Processor.new(T.Uniform(0.3, 0.9), 8, "job", "done", False, None)),
        _storage("storage"),
    ]
    C = Connector.new
    connectors = [
        C("c1", "generator", "load-balancer", "job", "job"),
        C("c2", "load-balancer", "stage1-a", "a", "job"),
        C("c3", "load-balancer", "stage1-b", "b", "job"),
        C("c4", "load-balancer", "stage1-c", "c", "job"),
        C("c5", "stage1-a", "processor-2", "done", "job"),
        C("c6", "stage1-b", "processor-2", "done", "job"),
        C("c7", "stage1-c", "processor-2", "done", "job"),
        C("c8", "processor-2", "processor-3", "done", "job"),
        C("c9", "processor-3", "storage", "done", "store"),
    ]
    return models, connectors


def gateway_network():
    """config 4 (16 models): generator Exp(2) -> parallel split (1 -> 4) -> {processor Exp;
    stochastic gate p=.9 -> processor Normal; exclusive [6,3,1] -> 3 processors, with a gate on that branch} and a
    parallel JOIN (2 -> 1) of the split's fourth output with the first branch's departures, upstream of the
    batcher(10, 10) -> storage; a stopwatch across the first branch.  The join collects every job twice (at the
    split and when proc-p releases it), so exactly one collection fills per departure (App. F)."""
    T = ContinuousRandomVariable
    models = [
        Model.new("generator", Generator.new(Exp(2.0), None, "job", False, None)),
        Model.new("split", ParallelGateway.new(["in"], ["p", "q", "r", "j"], False)),
        Model.new("proc-p", Processor.new(Exp(4.0), None, "job", "done", False, None)),
        Model.new("sgate-q", StochasticGate.new(BooleanRandomVariable.Bernoulli(0.9), "job", "job", False, None)),
        Model.new("proc-q", Processor.new(T.Normal(0.25, 0.02), 16, "job", "done", False, None)),
        Model.new("xgw-r", ExclusiveGateway.new(["in"], ["x", "y", "z"], IndexRandomVariable.WeightedIndex([6, 3, 1]),
                                                False, None)),
        Model.new("proc-x", Processor.new(Exp(3.0), 16, "job", "done", False, None)),
        Model.new("proc-y", Processor.new(T.Triangular(0.1, 0.5, 0.2), 16, "job", "done", False, None)),
        Model.new("proc-z", Processor.new(T.Uniform(0.1, 0.4), 16, "job", "done", False, None)),
        Model.new("gate-r", Gate.new("job", "activation", "deactivation", "job", False)),
        Model.new("join", ParallelGateway.new(["early", "late"], ["out"], False)),
        Model.new("batcher", Batcher.new("job", "job", 10.0, 10, False)),
        Model.new("stopwatch", Stopwatch.new("start", "stop", "metric", "job", Metric.Minimum, False)),
        _storage("storage-p"), _storage("storage-q"), _storage("storage-r"),
    ]
    C = Connector.new
    connectors = [
        C("c01", "generator", "split", "job", "in"),
        C("c02", "generator", "stopwatch", "job", "start"),
        C("c03", "split", "proc-p", "p", "job"),
        C("c04", "split", "sgate-q", "q", "job"),
        C("c05", "split", "xgw-r", "r", "in"),
        C("c06", "proc-p", "storage-p", "done", "store"),
        C("c07", "proc-p", "stopwatch", "done", "stop"),
        C("c08", "sgate-q", "proc-q", "job", "job"),
        C("c09", "proc-q", "storage-q", "done", "store"),
        C("c10", "split", "join", "j", "early"),
        C("c11", "proc-p", "join", "done", "late"),
        C("c12", "join", "batcher", "out", "job"),
        C("c13", "xgw-r", "proc-x", "x", "job"),
        C("c14", "xgw-r", "proc-y", "y", "job"),
        C("c15", "xgw-r", "proc-z", "z", "job"),
        C("c16", "proc-x", "gate-r", "done", "job"),
        C("c17", "proc-y", "gate-r", "done", "activation"),
        C("c18", "proc-z", "gate-r", "done", "deactivation"),
        C("c19", "proc-y", "storage-r", "done", "store"),
        C("c20", "gate-r", "storage-r", "job", "store"),
        C("c21", "batcher", "storage-q", "job", "store"),
    ]
    return models, connectors


def large_mixed():
    """config 5 (32 models): two gateway sub-networks and the tandem share one generator through a split."""
    models, connectors = [], []
    models.append(Model.new("source", Generator.new(Exp(2.0), None, "job", False, None)))
    models.append(Model.new("fanout", ParallelGateway.new(["in"], ["g1", "g2", "t"], False)))
    connectors.append(Connector.new("s0", "source", "fanout", "job", "in"))

    def sub(prefix, ms, cs, entry_port, entry_model, entry_target_port):
        for m in ms:
            if m.inner.type_name == "Generator":
                continue
            models.append(Model.new(prefix + m.id, m.inner))
        for c in cs:
            src = c.source_id
            if src == "generator":
                if c.target_id == entry_model and c.target_port == entry_target_port:
                    connectors.append(Connector.new(prefix + c.id, "fanout", prefix + c.target_id, entry_port, c.target_port))
                else:
                    connectors.append(Connector.new(prefix + c.id, "fanout", prefix + c.target_id, entry_port, c.target_port))
                continue
            connectors.append(Connector.new(prefix + c.id, prefix + c.source_id, prefix + c.target_id, c.source_port,
                                            c.target_port))

    g_models, g_conn = gateway_network()
    # trim each gateway sub-network to 11 models: drop the three storages and the stopwatch (and their connectors);
    # split -> ... -> join -> batcher stays whole
    drop = {"storage-p", "storage-q", "storage-r", "stopwatch"}
    gm = [m for m in g_models if m.id not in drop]
    gc = [c for c in g_conn if c.source_id not in drop and c.target_id not in drop]
    sub("a/", gm, gc, "g1", "split", "in")
    sub("b/", gm, gc, "g2", "split", "in")
    t_models, t_conn = tandem()
    tm = [m for m in t_models if m.id not in ("storage",)]
    tc = [c for c in t_conn if c.target_id != "storage"]
    sub("t/", tm, tc, "t", "load-balancer", "job")
    models.append(_storage("sink"))
    connectors.append(Connector.new("t/sink", "t/processor-3", "sink", "done", "store"))
    models.append(Model.new("monitor", Stopwatch.new("start", "stop", "metric", "job", Metric.Maximum, False)))
    connectors.append(Connector.new("m/start", "fanout", "monitor", "g1", "start"))
    connectors.append(Connector.new("m/stop", "a/proc-p", "monitor", "done", "stop"))
    assert len(models) <= 32, len(models)
    return models, connectors


def own_rngs(store_records=False, proc_lambda=1.5):
    """Models constructed with their own generator (`rng: Some(..)`: generator.rs:39,102-109, processor.rs,
    exclusive_gateway.rs, stochastic_gate.rs): each draws from its own stream, the second processor from the
    simulation's."""
    models = [
        Model.new("generator", Generator.new(Exp(1.0), None, "job", store_records, 11)),
        Model.new("proc-own", Processor.new(Exp(proc_lambda), 10, "job", "done", store_records, 12)),
        Model.new("exclusive", ExclusiveGateway.new(["in"], ["a", "b"], IndexRandomVariable.WeightedIndex([2, 1]), store_records, 13)),
        Model.new("sgate", StochasticGate.new(BooleanRandomVariable.Bernoulli(0.6), "job", "job", store_records, 14)),
        Model.new("proc-global", Processor.new(ContinuousRandomVariable.Normal(0.5, 0.05), 10, "job", "done", store_records, None)),
        _storage("sink-a", store_records), _storage("sink-b", store_records),
    ]
    C = Connector.new
    connectors = [C("c1", "generator", "proc-own", "job", "job"), C("c2", "proc-own", "exclusive", "done", "in"),
                  C("c3", "exclusive", "sgate", "a", "job"), C("c4", "exclusive", "proc-global", "b", "job"),
                  C("c5", "sgate", "sink-a", "job", "store"), C("c6", "proc-global", "sink-b", "done", "store")]
    return models, connectors


# ---- Coupled models (sim/src/models/coupled.rs, sim/tests/coupled.rs) --------------------------------------
def coupled_atomic(lam_gen=0.007, lam_proc=0.011):
    """coupled.rs:15-59: the atomic form of closure_under_coupling"""
    models = [Model.new("generator-01", Generator.new(Exp(lam_gen), None, "job", False, None)),
              Model.new("processor-01", Processor.new(Exp(lam_proc), 14, "job", "processed", False, None)),
              _storage("storage-01")]
    connectors = [Connector.new("connector-01", "generator-01", "processor-01", "job", "job"),
                  Connector.new("connector-02", "processor-01", "storage-01", "processed", "store")]
    return models, connectors


def coupled_closure(lam_gen=0.007, lam_proc=0.011, store_records=False):
    """coupled.rs:60-142: generator and processor inside one Coupled model; the generator's jobs leave on `start`
    and reach the processor over an internal coupling, the processed jobs leave on `stop`."""
    comp = [Model.new("generator-01", Generator.new(Exp(lam_gen), None, "job", store_records, None)),
            Model.new("processor-01", Processor.new(Exp(lam_proc), 14, "job", "processed", store_records, None))]
    models = [
        Model.new("coupled-01", Coupled.new([], ["start", "stop"], comp, [],
                                            [ExternalOutputCoupling("generator-01", "job", "start"),
                                             ExternalOutputCoupling("processor-01", "processed", "stop")],
                                            [InternalCoupling("generator-01", "processor-01", "job", "job")])),
        _storage("storage-02", store_records),
    ]
    connectors = [Connector.new("connector-01", "coupled-01", "storage-02", "start", "store"),
                  Connector.new("connector-02", "coupled-01", "storage-02", "stop", "store")]
    return models, connectors


def coupled_pipeline(store_records=False):
    """A Coupled model with external INPUT couplings (one in-port fanned out to two components), a zero-time
    component (load balancer) whose internal message is parked until the next firing, and a second Coupled model
    downstream: generator -> [balancer -> 2 processors] -> [gate -> batcher] -> storage."""
    stage = Coupled.new(
        ["in"], ["out", "seen"],
        [Model.new("balancer", LoadBalancer.new("job", ["a", "b"], store_records)),
         Model.new("proc-a", Processor.new(Exp(1.5), 6, "job", "done", store_records, None)),
         Model.new("proc-b", Processor.new(ContinuousRandomVariable.Uniform(0.2, 0.9), 6, "job", "done", store_records, None)),
         Model.new("monitor", _storage("x", store_records).inner)],
        [ExternalInputCoupling("balancer", "in", "job"), ExternalInputCoupling("monitor", "in", "store")],
        [ExternalOutputCoupling("proc-a", "done", "out"), ExternalOutputCoupling("proc-b", "done", "out"),
         ExternalOutputCoupling("balancer", "a", "seen")],
        [InternalCoupling("balancer", "proc-a", "a", "job"), InternalCoupling("balancer", "proc-b", "b", "job")])
    tail = Coupled.new(
        ["in"], ["out"],
        [Model.new("gate", Gate.new("job", "activation", "deactivation", "job", store_records)),
         Model.new("batcher", Batcher.new("job", "job", 4.0, 3, store_records))],
        [ExternalInputCoupling("gate", "in", "job")],
        [ExternalOutputCoupling("batcher", "job", "out")],
        [InternalCoupling("gate", "batcher", "job", "job")])
    models = [Model.new("generator", Generator.new(Exp(1.0), None, "job", store_records, None)),
              Model.new("stage", stage), Model.new("tail", tail), _storage("sink", store_records),
              _storage("seen", store_records)]
    C = Connector.new
    connectors = [C("c1", "generator", "stage", "job", "in"), C("c2", "stage", "tail", "out", "in"),
                  C("c3", "tail", "sink", "out", "store"), C("c4", "stage", "seen", "seen", "store")]
    return models, connectors


def shared_ids():
    """Two posted models may share an id: `Simulation::post` does not look (mod.rs:49-56), and `step` hands a message
    to EVERY model whose id matches (mod.rs:203-221), while `get_message_target_ids` finds the connectors of both by
    their common source id (mod.rs:155-182).  One connector therefore feeds both processors, and one collects from both."""
    T = ContinuousRandomVariable
    models = [
        Model.new("gen", Generator.new(T.Exp(0.8), None, "job", False, None)),
        Model.new("twin", Processor.new(T.Exp(1.5), 6, "job", "done", False, None)),
        Model.new("gate", Gate.new("job", "activation", "deactivation", "job", False)),
        Model.new("twin", Processor.new(T.Uniform(0.4, 1.1), 3, "job", "done", False, None)),
        Model.new("store", Storage.new("store", "read", "stored", False)),
    ]
    C = Connector.new
    connectors = [
        C("c-in", "gen", "twin", "job", "job"),
        C("c-out", "twin", "gate", "done", "job"),
        C("c-store", "gate", "store", "job", "store"),
    ]
    return models, connectors


COUPLED = {"coupled_closure": coupled_closure, "coupled_pipeline": coupled_pipeline}

ALL = {
    "mm1": mm1, "generator_storage": generator_storage, "exclusive_gateway": exclusive_gateway,
    "gate_blocking": gate_blocking, "load_balancer": load_balancer, "parallel_gateway": parallel_gateway,
    "stochastic_gate": stochastic_gate, "batcher": batcher, "stopwatch": stopwatch,
    "processor_from_queue": processor_from_queue, "processor_network": processor_network,
    "generator_passive": generator_passive, "more_variables": more_variables, "beta_variables": beta_variables, "state_injection": state_injection,
    "state_injection_tables": state_injection_tables, "batch_bursts": batch_bursts, "tandem": tandem, "gateway_network": gateway_network,
    "large_mixed": large_mixed, "own_rngs": own_rngs, "coupled_closure": coupled_closure, "coupled_pipeline": coupled_pipeline,
    "shared_ids": shared_ids,
}
