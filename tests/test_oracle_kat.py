"""Pins the CPU oracle (oracle/) against every exact fixture the reference holds for this path:
the output_analysis known-answer test, usize_sqrt, the t table, and the upstream RNG vectors."""
import math

import numpy as np
import pytest

import oracle_api as O
from sim_b200 import BooleanRandomVariable, ContinuousRandomVariable, IndexRandomVariable


def test_independent_sample_kat():
    # sim/src/output_analysis/mod.rs:356-364
    e, r = O.independent_sample([1.02, 0.73, 3.20, 0.23, 1.76, 0.47, 1.89, 1.45, 0.44, 0.23], 0.1)
    assert e == 0
    assert abs(r["lower"] - 0.7492630635369267) < 1e-12
    assert abs(r["upper"] - 1.534736936463073) < 1e-12


def test_usize_sqrt_kat():
    # sim/src/utils/mod.rs:98-106
    assert [O.usize_sqrt(n) for n in (1, 3, 4, 8, 9, 15)] == [1, 1, 2, 2, 3, 3]
    for n in list(range(1, 2000)) + [10**6, 10**9 + 7]:
        assert O.usize_sqrt(n) == math.isqrt(n)


def test_t_score_table_and_panics():
    # t_scores.rs:9-30: alpha is one-sided; df > 100 -> z row; unknown alpha / df 0 -> the reference panics
    assert O.t_score(0.1, 9) == (0, 1.383)
    assert O.t_score(0.05, 1) == (0, 6.314)
    assert O.t_score(0.0005, 1) == (0, 636.578)
    assert O.t_score(0.1, 100) == (0, 1.290)
    assert O.t_score(0.025, 101) == (0, 1.9600)
    assert O.t_score(0.001, 10**6) == (0, 3.0902)
    assert O.t_score(0.2, 5)[0] == 34
    assert O.t_score(0.05, 0)[0] == 34


def test_single_point_sample_is_degenerate():
    # output_analysis/mod.rs:110-115
    e, r = O.independent_sample([2.5], 0.05)
    assert e == 0 and r["lower"] == r["upper"] == 2.5


def test_steady_state_quirks():
    """SteadyStateOutput (mod.rs:224-333): deletion point from the MSER-style scan, batch count
    min(isqrt(n-d), 30), leftovers dropped from the front, asymmetric df in the interval."""
    rng = np.random.default_rng(1)
    x = np.concatenate([10 + rng.normal(size=50), rng.normal(size=950)])
    e, r = O.steady_state(x, 0.05)
    assert e == 0
    n = x.size
    # recompute in plain Python following the source line by line
    s = q = 0.0
    mser = [0.0] * (n - 1)
    for d in range(n - 2, -1, -1):
        s += x[d + 1]
        q += x[d + 1] ** 2
        mser[d] = q - s ** 2 / (n - d) ** 3
    m = min(mser[: (n - 1) // 2])
    d0 = mser.index(m)
    bc = min(math.isqrt(n - d0), 30)
    bs = (n - d0) // bc
    d1 = n - bc * bs
    assert (r["deletion_point"], r["batch_size"], r["batch_count"]) == (d1, bs, bc)
    means = [sum(x[d1 + bs * b: d1 + bs * (b + 1)]) / bs for b in range(bc)]
    bm = sum(means) / bc
    bv = sum((v - bm) ** 2 for v in means) / bc
    assert abs(r["mean"] - bm) < 1e-12 and abs(r["variance"] - bv) < 1e-12
    t_lo, t_hi = O.t_score(0.05, bc)[1], O.t_score(0.05, bc - 1)[1]
    assert abs(r["lower"] - (bm - t_lo * math.sqrt(bv) / math.sqrt(bc))) < 1e-12
    assert abs(r["upper"] - (bm + t_hi * math.sqrt(bv) / math.sqrt(bc))) < 1e-12
    assert O.steady_state([1.0], 0.05)[0] == 34  # `len() - 2` underflow panics in the reference


def test_pcg64mcg_upstream_vector():
    # rand_pcg 0.3 test vector for Mcg128Xsl64::new(42)
    want = [0x63b4a3a813ce700a, 0x382954200617ab24, 0xa7fd85ae3fe950ce, 0xd715286aa2887737, 0x60c92fee2e59f32c,
            0x84c4e96beff30017]
    assert [int(v) for v in O.pcg_fill(42, 6)] == want


def test_philox4x32_10_random123_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert O.philox_block([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert O.philox_block([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox_block([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_stream_layout():
    # u64 number i of replication r = words (2*(i&1), 2*(i&1)+1) of the block with counter (i>>1, 0, r_lo, r_hi) under
    # the key (seed_lo, seed_hi): the whole 64-bit seed and the whole 64-bit replication index enter, nothing aliases
    for seed, rep in ((42, 7), ((5 << 32) | 42, (3 << 32) | 7)):
        s = O.philox_fill(seed, rep, 6)
        for i in range(6):
            o = O.philox_block([i >> 1, 0, rep & 0xffffffff, rep >> 32], [seed & 0xffffffff, seed >> 32])
            w = o[0] | (o[1] << 32) if i % 2 == 0 else o[2] | (o[3] << 32)
            assert int(s[i]) == w
    assert not np.array_equal(O.philox_fill(42, 1 << 32, 4), O.philox_fill(43, 0, 4))


def test_ziggurat_tables_match_generated_include_and_upstream_literals():
    """The oracle builds the tables itself; the CUDA engine uses sim_b200/csrc/zig_tables.inc
    (tools/gen_ziggurat_tables.py).  Both must agree bit for bit, and with the leading literals of
    rand_distr's ziggurat_tables.rs."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "sim_b200", "csrc", "zig_tables.inc")).read()
    tabs = re.findall(r"double (ZIG_\w+)\[258\] = \{[^\n]*\n(.*?)\};", text, re.S)
    assert [t[0] for t in tabs] == ["ZIG_EXP_X", "ZIG_EXP_F", "ZIG_NORM_X", "ZIG_NORM_F"]
    for which, (_, body) in enumerate(tabs):
        vals = np.array([float.fromhex(v) for v in re.findall(r"(-?0x[0-9a-f.]+p[-+]?\d+)", body)])
        assert vals.size == 258 and vals[257] == 0.0  # 257 entries + one pad for the 16-byte bulk copy
        assert (vals[:257].view(np.uint64) == O.zig_table(which).view(np.uint64)).all()
    ex, nx = O.zig_table(0), O.zig_table(2)
    assert ex[0] == 8.697117470131052741 and ex[1] == 7.697117470131049720 and ex[2] == 6.941033629377212577
    assert ex[256] == 0.0 and abs(ex[255] - 0.06385216381500157) < 1e-15
    assert nx[0] == 3.910757959537090045 and nx[1] == 3.654152885361008796 and nx[2] == 3.449278298560964462
    assert nx[256] == 0.0 and abs(nx[255] - 0.2152418959132738) < 1e-15


# ---- the reference's own statistical tests of the samplers (random_variable.rs:213-461), run on the
# ---- restated algorithms with the reference's default PCG stream ------------------------------------
def _chi2(observed, expected):
    return float(sum((o - e) ** 2 / e for o, e in zip(observed, expected)))


def test_exp_mean_within_2_5_percent():  # random_variable.rs:228-239
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Exp(7.0), 10000, rng_kind=1)
    assert e == 0 and abs(x.mean() - 1 / 7.0) / (1 / 7.0) < 0.025 and (x > 0).all()


def test_normal_chi_square():  # :255-295 (mean 11, std 3, 8 equiprobable-ish bins, alpha .01)
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Normal(11.0, 3.0), 10000, rng_kind=1)
    edges = [-math.inf, 5, 8, 10, 11, 12, 14, 17, math.inf]
    cdf = lambda v: 0.5 * (1 + math.erf((v - 11.0) / (3.0 * math.sqrt(2))))
    obs = [int(((x >= a) & (x < b)).sum()) for a, b in zip(edges[:-1], edges[1:])]
    exp = [10000 * (cdf(b) - cdf(a)) for a, b in zip(edges[:-1], edges[1:])]
    assert _chi2(obs, exp) < 18.475  # 7 dof, alpha .01


def test_triangular_chi_square():  # :298-320
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Triangular(5.0, 25.0, 15.0), 10000, rng_kind=1)
    cdf = lambda v: (v - 5) ** 2 / 200.0 if v <= 15 else 1 - (25 - v) ** 2 / 200.0
    edges = [5, 10, 15, 20, 25]
    obs = [int(((x >= a) & (x < b)).sum()) for a, b in zip(edges[:-1], edges[1:])]
    exp = [10000 * (cdf(b) - cdf(a)) for a, b in zip(edges[:-1], edges[1:])]
    assert (x >= 5).all() and (x <= 25).all() and _chi2(obs, exp) < 11.345  # 3 dof


def test_uniform_chi_square():  # :323-346
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Uniform(7.0, 11.0), 10000, rng_kind=1)
    obs = [int(((x >= a) & (x < a + 1)).sum()) for a in (7, 8, 9, 10)]
    assert (x >= 7).all() and (x < 11).all() and _chi2(obs, [2500] * 4) < 11.345


def test_bernoulli_chi_square_and_p1_draws_nothing():  # :360-378
    e, b, used = O.sample_bernoulli(0.3, 10000, rng_kind=1)
    assert e == 0 and used == 10000
    assert _chi2([int(b.sum()), int(10000 - b.sum())], [3000, 7000]) < 6.635
    e, b, used = O.sample_bernoulli(1.0, 100, rng_kind=1)
    assert e == 0 and b.all() and used == 0  # p == 1.0 never draws (rand 0.8 Bernoulli ALWAYS_TRUE)
    assert O.sample_bernoulli(1.5, 1)[0] == 20  # BernoulliError


def test_weighted_index_chi_square():  # :419-439
    e, x, _ = O.sample_index(IndexRandomVariable.WeightedIndex([1, 2, 3, 4]), 10000, rng_kind=1)
    obs = [int((x == i).sum()) for i in range(4)]
    assert e == 0 and _chi2(obs, [1000, 2000, 3000, 4000]) < 11.345
    assert O.sample_index(IndexRandomVariable.WeightedIndex([0, 0]), 1)[0] == 23  # AllWeightsZero
    assert O.sample_index(IndexRandomVariable.WeightedIndex([]), 1)[0] == 23      # NoItem


def test_index_uniform_chi_square():  # :442-461
    e, x, _ = O.sample_index(IndexRandomVariable.Uniform(7, 11), 10000, rng_kind=1)
    obs = [int((x == i).sum()) for i in (7, 8, 9, 10)]
    assert e == 0 and _chi2(obs, [2500] * 4) < 11.345
    assert O.sample_index(IndexRandomVariable.Uniform(5, 5), 1)[0] == 34  # Uniform::new(low >= high) panics


def test_distribution_parameter_errors_surface_per_draw():
    assert O.sample_continuous(ContinuousRandomVariable.Exp(0.0), 1)[0] == 15
    assert O.sample_continuous(ContinuousRandomVariable.Exp(-1.0), 1)[0] == 15
    assert O.sample_continuous(ContinuousRandomVariable.Normal(0.0, float("inf")), 1)[0] == 17
    assert O.sample_continuous(ContinuousRandomVariable.Triangular(0.0, 1.0, 2.0), 1)[0] == 18
    assert O.sample_continuous(ContinuousRandomVariable.Triangular(1.0, 1.0, 1.0), 1)[0] == 18
    assert O.sample_continuous(ContinuousRandomVariable.Uniform(2.0, 1.0), 1)[0] == 34


# ---- SURVEY.md 8(f)-3 variables: the reference's own mean tests (random_variable.rs:232-252, 348-357) ----------
def test_gamma_mean():
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Gamma(7.0, 11.0), 10000, rng_kind=1)
    assert e == 0 and abs(x.mean() - 77.0) / 77.0 < 0.025 and (x > 0).all()
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Gamma(0.4, 2.0), 20000, rng_kind=1)  # small-shape path
    assert e == 0 and abs(x.mean() - 0.8) / 0.8 < 0.05
    e, x, used = O.sample_continuous(ContinuousRandomVariable.Gamma(1.0, 3.0), 10000, rng_kind=1)  # Exp(1/scale)
    assert e == 0 and abs(x.mean() - 3.0) / 3.0 < 0.025
    assert O.sample_continuous(ContinuousRandomVariable.Gamma(0.0, 1.0), 1)[0] == 16
    assert O.sample_continuous(ContinuousRandomVariable.Gamma(1.0, -1.0), 1)[0] == 16


def test_beta_mean_and_variance():
    """random_variable.rs:214-222 (alpha 7, beta 11 -> mean 7/18 within 2.5 %), plus every branch of
    rand_distr 0.4.3's sampler: BB (min > 1) with and without switched parameters, BC (min <= 1) likewise."""
    for alpha, beta in ((7.0, 11.0), (11.0, 7.0), (2.0, 2.0), (0.5, 0.7), (3.0, 0.8), (0.8, 3.0), (1.0, 1.0), (0.3, 0.3)):
        e, x, _ = O.sample_continuous(ContinuousRandomVariable.Beta(alpha, beta), 40000, rng_kind=1)
        mean = alpha / (alpha + beta)
        var = alpha * beta / ((alpha + beta) ** 2 * (alpha + beta + 1.0))
        assert e == 0 and (x >= 0).all() and (x <= 1).all()
        assert abs(x.mean() - mean) / mean < 0.025, (alpha, beta, x.mean(), mean)
        assert abs(x.var() - var) / var < 0.06, (alpha, beta, x.var(), var)
    # a symmetric Beta is symmetric: the third central moment vanishes
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Beta(2.0, 2.0), 40000, rng_kind=1)
    assert abs(((x - 0.5) ** 3).mean()) < 8e-4
    assert O.sample_continuous(ContinuousRandomVariable.Beta(0.0, 1.0), 1)[0] == 14
    assert O.sample_continuous(ContinuousRandomVariable.Beta(1.0, -2.0), 1)[0] == 14


def test_lognormal_mean():
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.LogNormal(11.0, 1.0), 10000, rng_kind=1)
    expected = math.exp(11.0 + 0.5)
    assert e == 0 and abs(x.mean() - expected) / expected < 0.06  # heavy tail: the reference's 2.5 % holds for ITS seed


def test_weibull_mean_pins_the_swapped_parameters():
    """random_variable.rs:348-357 expects mean 14 for {shape: 7, scale: 0.5}: that is rand_distr's
    Weibull(scale = 7, shape = 0.5), 7 * Gamma(3) -- the reference passes the two parameters swapped."""
    e, x, _ = O.sample_continuous(ContinuousRandomVariable.Weibull(7.0, 0.5), 40000, rng_kind=1)
    assert e == 0 and abs(x.mean() - 14.0) / 14.0 < 0.04
    assert O.sample_continuous(ContinuousRandomVariable.Weibull(0.0, 1.0), 1)[0] == 19
