"""Seeded random DEVS networks for differential testing (tests/test_random_networks.py): arbitrary wiring of the ten
built-in models -- cycles, fan-out of one port over several connectors, fan-in, self-loops, connectors that name a
port or a model that does not exist -- so that the engine and the oracle are compared on shapes nobody wrote by hand."""
import random

from sim_b200 import (Batcher, BooleanRandomVariable, Connector, ContinuousRandomVariable, ExclusiveGateway, Gate,
                      Generator, IndexRandomVariable, LoadBalancer, Metric, Model, ParallelGateway, Processor,
                      StochasticGate, Stopwatch, Storage)

T = ContinuousRandomVariable


def _continuous(rnd):
    k = rnd.randrange(8)
    if k == 6:
        return T.LogNormal(-0.3, 0.4)
    if k == 7:
        return T.Beta(rnd.choice([0.6, 2.0]), rnd.choice([0.8, 3.0]))
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


def _random_model(rnd, mid, rec, own, generator=False, finite=False):
    """(Model, input port names, output port names)"""
    k = 0 if generator else rnd.randrange(11)
    if k == 0 or k == 10:
        body, pi, po = Generator.new(_continuous(rnd), None, "job", rec, own()), [], ["job"]
    elif k == 1:
        cap = rnd.choice([None, 1, 2, 5, 0]) if rnd.random() < 0.9 else 14
        if finite and cap is None:
            cap = 3
        body, pi, po = Processor.new(_continuous(rnd), cap, "job", "done", rec, own()), ["job"], ["done"]
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
        body, pi, po = ExclusiveGateway.new(pin, pout, w, rec, own()), pin, pout
    elif k == 6:
        pin = ["a%d" % j for j in range(rnd.randint(1, 2))]
        pout = ["b%d" % j for j in range(rnd.randint(1, 2))]
        body, pi, po = ParallelGateway.new(pin, pout, rec), pin, pout
    elif k == 7:
        p = rnd.choice([0.0, 0.3, 0.8, 1.0])
        body, pi, po = StochasticGate.new(BooleanRandomVariable.Bernoulli(p), "job", "out", rec, own()), ["job"], ["out"]
    elif k == 8:
        body = Stopwatch.new("start", "stop", "metric", "job", rnd.choice([Metric.Minimum, Metric.Maximum]), rec)
        pi, po = ["start", "stop", "metric"], ["job"]
    else:
        body = Batcher.new("job", "out", rnd.choice([0.5, 2.0, 10.0]), rnd.choice([1, 2, 4]), rec)
        pi, po = ["job"], ["out"]
    return Model.new(mid, body), pi, po


def random_network(seed, allow_bad_ports=True, acyclic=False):
    """(models, connectors).  Every model has an id `m<i>`.  `acyclic`: connectors only run from a model to a later one
    and every queue is finite -- networks that settle instead of growing without bound."""
    rnd = random.Random(seed)
    n = rnd.randint(2, 10)
    models, ins, outs = [], [], []   # ins[i] / outs[i]: port names of model i
    rec = rnd.random() < 0.5

    def own():  # some drawing models bring their own generator (generator.rs:85-95 and alike)
        return rnd.randrange(1, 1000) if rnd.random() < 0.15 else None

    for i in range(n):
        # the first model is always a generator: something must happen
        m, pi, po = _random_model(rnd, "m%d" % i, rec, own, generator=(i == 0), finite=acyclic)
        models.append(m)
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


def random_injections(seed, models):
    """A few (target id, target port, content) triples for inject_input: real ports mostly, now and then a port or a
    model that does not exist."""
    rnd = random.Random(seed * 7919 + 1)
    out = []
    for k in range(rnd.randrange(0, 4)):
        m = rnd.choice(models)
        ports = list(m.inner.fields().get("portsIn", {}).values())
        flat = []
        for p in ports:
            flat.extend(p if isinstance(p, list) else [p])
        port = rnd.choice(flat) if flat and rnd.random() < 0.9 else "nowhere"
        tid = m.id if rnd.random() < 0.95 else "nobody"
        out.append((tid, port, "manual %d" % (k + 1)))
    return out


def random_coupled_network(seed):
    """(models, connectors) with one to three Coupled models (coupled.rs:14-25) among the posted models: random
    components, external input couplings (one port may reach several components, or none), external output couplings,
    and internal couplings in any direction -- the parked-message rules of coupled.rs:186-255 on wiring nobody wrote by
    hand."""
    from sim_b200 import Coupled, ExternalInputCoupling, ExternalOutputCoupling, InternalCoupling
    rnd = random.Random(seed * 31 + 7)
    rec = rnd.random() < 0.5

    def own():
        return rnd.randrange(1, 1000) if rnd.random() < 0.1 else None

    n = rnd.randint(2, 5)
    models, ins, outs = [], [], []
    n_coupled = 0
    for i in range(n):
        mid = "m%d" % i
        if i == 0:
            m, pi, po = _random_model(rnd, mid, rec, own, generator=True)
        elif (i == 1 or rnd.random() < 0.3) and n_coupled < 3:
            n_coupled += 1
            k = rnd.randint(1, 4)
            comps, cin, cout = [], [], []
            for j in range(k):
                cm, ci, co = _random_model(rnd, "%s.k%d" % (mid, j), rec, own, generator=(rnd.random() < 0.1), finite=True)
                comps.append(cm)
                cin.append(ci)
                cout.append(co)
            pi = ["in%d" % q for q in range(rnd.randint(1, 2))]
            po = ["out%d" % q for q in range(rnd.randint(1, 2))]
            with_in = [j for j in range(k) if cin[j]]
            with_out = [j for j in range(k) if cout[j]]
            eic, eoc, ic = [], [], []
            for port in pi:
                for _ in range(rnd.randint(0 if rnd.random() < 0.15 else 1, 2)):
                    if with_in:
                        j = rnd.choice(with_in)
                        eic.append(ExternalInputCoupling(comps[j].id, port, rnd.choice(cin[j])))
            for _ in range(rnd.randint(1, 3)):
                if with_out:
                    j = rnd.choice(with_out)
                    eoc.append(ExternalOutputCoupling(comps[j].id, rnd.choice(cout[j]), rnd.choice(po)))
            for _ in range(rnd.randint(0, k + 1)):
                if with_out and with_in:
                    a, b = rnd.choice(with_out), rnd.choice(with_in)
                    ic.append(InternalCoupling(comps[a].id, comps[b].id, rnd.choice(cout[a]), rnd.choice(cin[b])))
            m = Model.new(mid, Coupled.new(pi, po, comps, eic, eoc, ic))
        else:
            m, pi, po = _random_model(rnd, mid, rec, own, finite=True)
        models.append(m)
        ins.append(pi)
        outs.append(po)
    connectors = []
    for c in range(rnd.randint(n - 1, 2 * n + 1)):
        s_, t = rnd.randrange(n), rnd.randrange(n)
        if not outs[s_] or not ins[t]:
            continue
        tp = rnd.choice(ins[t]) if rnd.random() < 0.97 else "no-such-port"
        connectors.append(Connector.new("c%d" % c, "m%d" % s_, "m%d" % t, rnd.choice(outs[s_]), tp))
    return models, connectors
