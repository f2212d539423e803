"""Seeded random DEVS networks for differential testing (tests/test_random_networks.py): arbitrary wiring of the ten
built-in models -- cycles, fan-out of one port over several connectors, fan-in, self-loops, connectors that name a
port or a model that does not exist -- so that the engine and the oracle are compared on shapes nobody wrote by hand."""
import random

from sim_b200 import (Batcher, BooleanRandomVariable, Connector, ContinuousRandomVariable, ExclusiveGateway, Gate,
                      Generator, IndexRandomVariable, LoadBalancer, Metric, Model, ParallelGateway, Processor,
                      StochasticGate, Stopwatch, Storage)

T = ContinuousRandomVariable


def _continuous(rnd):
    k = rnd.randrange(6)
    if k == 0:
        return T.Exp(rnd.choice([0.3, 0.7, 1.0, 2.5]))
    if k == 1:
        return T.Uniform(0.1, rnd.choice([0.5, 1.5, 3.0]))
    if k == 2:
        return T.Triangular(0.1, 2.0, rnd.choice([0.3, 1.0, 1.9]))
    if k == 3:
        return T.Normal(rnd.choice([1.0, 2.0]), 0.1)
    if k == 4:
        return T.Gamma(rnd.choice([0.5, 1.0, 2.5]), 0.8)
    return T.Weibull(1.5, rnd.choice([0.7, 1.2]))


def random_network(seed, allow_bad_ports=True, acyclic=False):
    """(models, connectors).  Every model has an id `m<i>`.  `acyclic`: connectors only run from a model to a later one
    and every queue is finite -- networks that settle instead of growing without bound."""
    rnd = random.Random(seed)
    n = rnd.randint(2, 10)
    models, ins, outs = [], [], []   # ins[i] / outs[i]: port names of model i
    rec = rnd.random() < 0.5
    for i in range(n):
        mid = "m%d" % i
        k = rnd.randrange(11) if i else 0   # the first model is always a generator: something must happen
        if k == 0 or k == 10:
            body, pi, po = Generator.new(_continuous(rnd), None, "job", rec, None), [], ["job"]
        elif k == 1:
            cap = rnd.choice([None, 1, 2, 5, 0]) if rnd.random() < 0.9 else 14
            if acyclic and cap is None:
                cap = 3
            body, pi, po = Processor.new(_continuous(rnd), cap, "job", "done", rec, None), ["job"], ["done"]
        elif k == 2:
            body, pi, po = Storage.new("store", "read", "stored", rec), ["store", "read"], ["stored"]
        elif k == 3:
            paths = ["p%d" % j for j in range(rnd.randint(1, 3))]
            body, pi, po = LoadBalancer.new("job", paths, rec), ["job"], paths
        elif k == 4:
            body, pi, po = Gate.new("job", "open", "close", "out", rec), ["job", "open", "close"], ["out"]
        elif k == 5:
            pin = ["a%d" % j for j in range(rnd.randint(1, 2))]
            pout = ["b%d" % j for j in range(rnd.randint(1, 3))]
            w = IndexRandomVariable.WeightedIndex([rnd.randint(1, 5) for _ in pout]) if rnd.random() < 0.7 else \
                IndexRandomVariable.Uniform(0, len(pout))
            body, pi, po = ExclusiveGateway.new(pin, pout, w, rec, None), pin, pout
        elif k == 6:
            pin = ["a%d" % j for j in range(rnd.randint(1, 2))]
            pout = ["b%d" % j for j in range(rnd.randint(1, 2))]
            body, pi, po = ParallelGateway.new(pin, pout, rec), pin, pout
        elif k == 7:
            p = rnd.choice([0.0, 0.3, 0.8, 1.0])
            body, pi, po = StochasticGate.new(BooleanRandomVariable.Bernoulli(p), "job", "out", rec, None), ["job"], ["out"]
        elif k == 8:
            body = Stopwatch.new("start", "stop", "metric", "job", rnd.choice([Metric.Minimum, Metric.Maximum]), rec)
            pi, po = ["start", "stop", "metric"], ["job"]
        else:
            body = Batcher.new("job", "out", rnd.choice([0.5, 2.0, 10.0]), rnd.choice([1, 2, 4]), rec)
            pi, po = ["job"], ["out"]
        models.append(Model.new(mid, body))
        ins.append(pi)
        outs.append(po)
    connectors = []
    n_conn = rnd.randint(n - 1, 2 * n + 2)
    for c in range(n_conn):
        s = rnd.randrange(n)
        if not outs[s]:
            continue
        t = rnd.randrange(n)
        if acyclic:
            if s == n - 1:
                continue
            t = rnd.randrange(s + 1, n)
        if not ins[t]:          # a generator has no input port: a message to it is an error of the network, keep it rare
            if rnd.random() < 0.9:
                continue
            tp = "job"
        else:
            tp = rnd.choice(ins[t])
            # a message on a Storage `read` or a Stopwatch `metric` port is legal and exercises those branches
        sp = rnd.choice(outs[s])
        tid = "m%d" % t
        if allow_bad_ports and rnd.random() < 0.03:
            tp = "no-such-port"
        if allow_bad_ports and rnd.random() < 0.03:
            tid = "no-such-model"
        connectors.append(Connector.new("c%d" % c, "m%d" % s, tid, sp, tp))
    return models, connectors
