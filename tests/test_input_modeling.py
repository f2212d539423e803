"""input_modeling: the Discrete variables (random_variable.rs:36-50, 108-115), Thinning::evaluate (thinning.rs:27-34) and
the device sampler simb_sample_random_variable.

CPU tier: the reference's own statistical tests of the Discrete variables (random_variable.rs:381-416: Geometric and
Poisson means within 2.5 %, chi-square of Discrete::Uniform at alpha 0.01) on the oracle with the reference's default
stream, the error variants, the polynomial doctest of utils/mod.rs:23-33.
GPU tier: every family's variates from the device equal the oracle's on the same per-replication Philox streams
(integers and booleans bit for bit, f64 within 1e-12), draw counts included."""
import numpy as np
import pytest

import oracle_api as O
from sim_b200 import (BooleanRandomVariable, ContinuousRandomVariable, DiscreteRandomVariable, IndexRandomVariable,
                      Thinning)

PCG = 1  # the reference's default_rng(): Pcg64Mcg::new(42)


def test_geometric_samples_match_expectation():  # random_variable.rs:381-386
    e, v, _ = O.sample_discrete(0, 10000, p=0.2, rng_kind=PCG)
    assert e == 0
    expected = (1.0 - 0.2) / 0.2
    assert abs(v.mean() - expected) / expected < 0.025


def test_poisson_samples_match_expectation():  # random_variable.rs:389-394
    e, v, _ = O.sample_discrete(1, 10000, p=7.0, rng_kind=PCG)
    assert e == 0 and abs(v.mean() - 7.0) / 7.0 < 0.025
    # the rejection branch (lambda >= 12): mean and variance of a Poisson
    e, v, _ = O.sample_discrete(1, 20000, p=40.0, rng_kind=PCG)
    assert e == 0 and abs(v.mean() - 40.0) / 40.0 < 0.025 and abs(v.var() - 40.0) / 40.0 < 0.06


def test_discrete_uniform_samples_chi_square():  # random_variable.rs:397-416
    e, v, used = O.sample_discrete(2, 10000, mn=7, mx=11, rng_kind=PCG)
    assert e == 0 and v.min() >= 7 and v.max() <= 10 and used == 10000
    counts = np.bincount((v - 7).astype(np.int64), minlength=4)
    chi = float((((counts - 2500.0) ** 2) / 2500.0).sum())
    assert chi < 13.277


def test_geometric_branches_and_errors():
    # p >= 2/3: the trivial algorithm; p == 0: u64::MAX without a draw; p == 1: always 0 failures
    e, v, _ = O.sample_discrete(0, 20000, p=0.8, rng_kind=PCG)
    assert e == 0 and abs(v.mean() - 0.25) < 0.02
    e, v, used = O.sample_discrete(0, 3, p=0.0)
    assert e == 0 and (v == np.uint64(2**64 - 1)).all() and used == 0
    e, v, _ = O.sample_discrete(0, 100, p=1.0)
    assert e == 0 and (v == 0).all()
    assert O.sample_discrete(0, 1, p=1.5)[0] == 21 and O.sample_discrete(0, 1, p=float("nan"))[0] == 21   # GeoError
    assert O.sample_discrete(1, 1, p=0.0)[0] == 22 and O.sample_discrete(1, 1, p=-1.0)[0] == 22           # PoissonError
    assert O.sample_discrete(2, 1, mn=5, mx=5)[0] == 34                                                   # Uniform::new panics


def test_thinning_polynomial():  # utils/mod.rs:23-33 doctest; thinning.rs:27-34
    co, x = [2.0, -3.0, 1.0, -2.0, 3.0], 44.0
    want = 2.0 * x ** 4 + -3.0 * x ** 3 + 1.0 * x ** 2 + -2.0 * x + 3.0
    assert O.thinning_evaluate(co, x) == want
    assert Thinning(co).evaluate(x) == want
    assert Thinning([]).evaluate(3.0) == 0.0 and Thinning([0.5]).evaluate(1e9) == 0.5
    assert Thinning(co).to_dict() == {"function": {"polynomial": {"coefficients": co}}}


CASES = [
    ("continuous", ContinuousRandomVariable.Exp(0.5)), ("continuous", ContinuousRandomVariable.Normal(3.0, 0.5)),
    ("continuous", ContinuousRandomVariable.Triangular(0.2, 1.4, 0.6)), ("continuous", ContinuousRandomVariable.Uniform(0.3, 0.9)),
    ("continuous", ContinuousRandomVariable.Gamma(2.5, 0.8)), ("continuous", ContinuousRandomVariable.Gamma(0.6, 1.5)),
    ("continuous", ContinuousRandomVariable.LogNormal(-0.2, 0.5)), ("continuous", ContinuousRandomVariable.Weibull(1.2, 2.0)),
    ("continuous", ContinuousRandomVariable.Beta(2.0, 5.0)), ("continuous", ContinuousRandomVariable.Beta(0.5, 0.7)),
    ("boolean", BooleanRandomVariable.Bernoulli(0.3)), ("boolean", BooleanRandomVariable.Bernoulli(1.0)),
    ("discrete", DiscreteRandomVariable.Geometric(0.2)), ("discrete", DiscreteRandomVariable.Geometric(0.8)),
    ("discrete", DiscreteRandomVariable.Geometric(0.01)), ("discrete", DiscreteRandomVariable.Poisson(7.0)),
    ("discrete", DiscreteRandomVariable.Poisson(40.0)), ("discrete", DiscreteRandomVariable.Uniform(7, 11)),
    ("index", IndexRandomVariable.Uniform(7, 11)), ("index", IndexRandomVariable.WeightedIndex([1, 2, 3, 4])),
]


def _oracle(family, var, n, seed, rep):
    if family == "continuous":
        return O.sample_continuous(var, n, 0, seed, rep)
    if family == "boolean":
        return O.sample_bernoulli(var.params["p"], n, 0, seed, rep)
    if family == "discrete":
        k = {"geometric": 0, "poisson": 1, "uniform": 2}[var.variant]
        return O.sample_discrete(k, n, p=var.params.get("p", var.params.get("lambda", 0.0)), mn=var.params.get("min", 0),
                                 mx=var.params.get("max", 1), seed=seed, rep=rep)
    return O.sample_index(var, n, 0, seed, rep)


@pytest.mark.gpu
@pytest.mark.parametrize("family,var", CASES, ids=["%s-%s-%s" % (f, v.variant, "-".join(str(x) for x in v.params.values())) for f, v in CASES])
def test_device_variates_equal_the_oracle(family, var):
    from sim_b200.input_modeling import sample
    reps, count, seed, off = 70, 400, 123, 5
    got, draws = sample(family, var, reps, count, seed=seed, replication_offset=off, with_draws=True)
    assert got.shape == (count, reps)
    for r in (0, 1, 33, 69):
        e, want, used = _oracle(family, var, count, seed, off + r)
        assert e == 0 and used == draws[r]
        if family == "continuous":
            assert np.allclose(got[:, r], want, rtol=1e-12, atol=0.0)
        else:
            assert np.array_equal(got[:, r], want.astype(np.uint64))


@pytest.mark.gpu
def test_device_sampler_reference_statistics_and_errors():
    """The reference's own tests on the device streams (replication 0..) and its error variants at draw time."""
    from sim_b200 import SimulationError
    from sim_b200.input_modeling import sample
    v = sample("discrete", DiscreteRandomVariable.Geometric(0.2), 100, 1000).astype(np.float64)
    assert abs(v.mean() - 4.0) / 4.0 < 0.025
    v = sample("discrete", DiscreteRandomVariable.Poisson(7.0), 100, 1000).astype(np.float64)
    assert abs(v.mean() - 7.0) / 7.0 < 0.025
    v = sample("discrete", DiscreteRandomVariable.Uniform(7, 11), 100, 1000)
    counts = np.bincount((v.reshape(-1) - 7).astype(np.int64), minlength=4)
    assert float((((counts - 25000.0) ** 2) / 25000.0).sum()) < 13.277
    for fam, var, name in (("discrete", DiscreteRandomVariable.Geometric(1.5), "GeoError"),
                           ("discrete", DiscreteRandomVariable.Poisson(0.0), "PoissonError"),
                           ("discrete", DiscreteRandomVariable.Uniform(5, 5), "Panic"),
                           ("continuous", ContinuousRandomVariable.Exp(-1.0), "ExpError"),
                           ("boolean", BooleanRandomVariable.Bernoulli(1.5), "BernoulliError"),
                           ("index", IndexRandomVariable.WeightedIndex([0, 0]), "WeightedError")):
        with pytest.raises(SimulationError) as ei:
            sample(fam, var, 4, 2)
        assert ei.value.name == name


def test_a_model_with_its_own_rng_is_independent_of_the_other_draws():
    """generator.rs:102-109 (`match &self.rng`): a Generator constructed with its own generator produces the same
    arrival times whatever the rest of the network draws; without one its stream is shared."""
    import networks

    def generation_times(models, conns):
        o = O.OracleSimulation(models, conns)
        o.set_rng_philox(5, 3)
        o.trace_enable(True)
        e, _ = o.step_n(600, collect=False)
        assert e == 0
        return [ev["time"] for ev in o.trace() if ev["kind"] == 1 and ev["action"] == 1 and ev["index"] == 0]

    a = generation_times(*networks.own_rngs(proc_lambda=1.5))
    b = generation_times(*networks.own_rngs(proc_lambda=4.0))
    n = min(len(a), len(b))
    # (the clock is a running sum over different step sequences: equal to rounding, not to the bit)
    assert n > 50 and np.allclose(a[:n], b[:n], rtol=1e-12, atol=0.0)
    # the same networks with the generator on the shared stream diverge
    def shared(lam):
        m, c = networks.own_rngs(proc_lambda=lam)
        m[0].inner.rng = None
        m[1].inner.rng = None
        return generation_times(m, c)
    a, b = shared(1.5), shared(4.0)
    n = min(len(a), len(b))
    assert not np.allclose(a[:n], b[:n], rtol=1e-6, atol=0.0)


@pytest.mark.gpu
def test_device_variates_equal_the_oracle_for_random_parameters():
    """Seeded random parameters of every family -- small and large shapes, both Geometric and both Poisson algorithms,
    wide integer ranges, lop-sided weights, and parameters the reference rejects at draw time: the device's variates
    and draw counts equal the oracle's on the same stream, and a rejected parameter raises the same error."""
    import random
    from sim_b200 import SimulationError
    from sim_b200.input_modeling import sample
    T, D, I = ContinuousRandomVariable, DiscreteRandomVariable, IndexRandomVariable
    rnd = random.Random(2024)
    def logu(lo, hi):
        return float(np.exp(rnd.uniform(np.log(lo), np.log(hi))))
    makers = [
        lambda: ("continuous", T.Exp(logu(1e-3, 1e3) if rnd.random() < 0.9 else -1.0)),
        lambda: ("continuous", T.Normal(rnd.uniform(-50, 50), logu(1e-3, 10) if rnd.random() < 0.9 else -1.0)),
        lambda: ("continuous", T.Uniform(rnd.uniform(-5, 5), rnd.uniform(5.1, 50))),
        lambda: ("continuous", (lambda a, b: T.Triangular(a, b, rnd.uniform(a, b)))(rnd.uniform(-3, 3), rnd.uniform(3.5, 9))),
        lambda: ("continuous", T.Gamma(logu(0.05, 40), logu(0.01, 20))),
        lambda: ("continuous", T.LogNormal(rnd.uniform(-2, 2), logu(0.01, 1.5))),
        lambda: ("continuous", T.Weibull(logu(0.2, 8), logu(0.1, 10))),
        lambda: ("continuous", T.Beta(logu(0.05, 30), logu(0.05, 30))),
        lambda: ("boolean", BooleanRandomVariable.Bernoulli(rnd.choice([0.0, 1.0, rnd.random(), 1.5]))),
        lambda: ("discrete", D.Geometric(rnd.choice([logu(1e-4, 0.66), rnd.uniform(0.67, 1.0), 1.0, 0.0, -0.1]))),
        lambda: ("discrete", D.Poisson(rnd.choice([logu(0.01, 11.9), logu(12.0, 5000.0), 0.0]))),
        lambda: ("discrete", (lambda a: D.Uniform(a, a + rnd.choice([0, 1, 2, 1000, 2**40])))(rnd.randrange(0, 10**6))),
        lambda: ("index", (lambda a: I.Uniform(a, a + rnd.choice([1, 3, 2**33])))(rnd.randrange(0, 1000))),
        lambda: ("index", I.WeightedIndex([rnd.choice([0, 1, 1, 7, 10**6]) for _ in range(rnd.randint(1, 12))])),
    ]
    for case in range(140):
        family, var = makers[case % len(makers)]()
        reps, count, seed = 5, 300, 1000 + case
        e, want, used = _oracle(family, var, count, seed, 3)
        what = "%s %s %s" % (family, var.variant, var.params)
        if e:
            with pytest.raises(SimulationError) as ei:
                sample(family, var, reps, count, seed=seed)
            assert ei.value.status == e, what
            continue
        got, draws = sample(family, var, reps, count, seed=seed, with_draws=True)
        assert used == draws[3], what
        if family == "continuous":
            assert np.allclose(got[:, 3], want, rtol=1e-11, atol=0.0, equal_nan=True), what
        else:
            assert np.array_equal(got[:, 3], want.astype(np.uint64)), what
