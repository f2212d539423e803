"""Random variables of the accelerated path, mirroring ``sim::input_modeling``
(sim/src/input_modeling/random_variable.rs:17-63).

The objects only carry parameters and know their wire form (the reference's serde layout:
camelCase variant name, Rust field names inside); sampling happens on the
device (sim_b200/csrc/core.cuh) from per-replication streams.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass(frozen=True)
class _Variant:
    variant: str
    params: dict

    def to_dict(self) -> dict:
        return {self.variant: dict(self.params)}


class ContinuousRandomVariable:
    """``Continuous`` (random_variable.rs:17-28): Beta, Exp, Normal, Triangular, Uniform, Gamma, LogNormal, Weibull
    ."""

    @staticmethod
    def Exp(lambda_: float) -> _Variant:
        return _Variant("exp", {"lambda": float(lambda_)})

    @staticmethod
    def Normal(mean: float, std_dev: float) -> _Variant:
        return _Variant("normal", {"mean": float(mean), "std_dev": float(std_dev)})

    @staticmethod
    def Triangular(min: float, max: float, mode: float) -> _Variant:  # noqa: A002 - reference field names
        return _Variant("triangular", {"min": float(min), "max": float(max), "mode": float(mode)})

    @staticmethod
    def Uniform(min: float, max: float) -> _Variant:  # noqa: A002
        return _Variant("uniform", {"min": float(min), "max": float(max)})


    @staticmethod
    def Beta(alpha: float, beta: float) -> _Variant:
        return _Variant("beta", {"alpha": float(alpha), "beta": float(beta)})

    @staticmethod
    def Gamma(shape: float, scale: float) -> _Variant:
        return _Variant("gamma", {"shape": float(shape), "scale": float(scale)})

    @staticmethod
    def LogNormal(mu: float, sigma: float) -> _Variant:
        return _Variant("logNormal", {"mu": float(mu), "sigma": float(sigma)})

    @staticmethod
    def Weibull(shape: float, scale: float) -> _Variant:
        """NOTE: the reference hands (shape, scale) to rand_distr's ``Weibull::new(scale, shape)``
        (random_variable.rs:85-87), so the two parameters act with swapped roles; reproduced."""
        return _Variant("weibull", {"shape": float(shape), "scale": float(scale)})


class BooleanRandomVariable:
    """``Boolean`` (random_variable.rs:30-34)."""

    @staticmethod
    def Bernoulli(p: float) -> _Variant:
        return _Variant("bernoulli", {"p": float(p)})


class IndexRandomVariable:
    """``Index`` (random_variable.rs:52-63); Uniform is [min, max)."""

    @staticmethod
    def Uniform(min: int, max: int) -> _Variant:  # noqa: A002
        return _Variant("uniform", {"min": int(min), "max": int(max)})

    @staticmethod
    def WeightedIndex(weights: List[int]) -> _Variant:
        return _Variant("weightedIndex", {"weights": [int(w) for w in weights]})


class DiscreteRandomVariable:
    """``Discrete`` (random_variable.rs:36-50): Geometric, Poisson, Uniform [min, max).  No built-in model draws
    them; ``sample`` below does, on the device."""

    @staticmethod
    def Geometric(p: float) -> _Variant:
        return _Variant("geometric", {"p": float(p)})

    @staticmethod
    def Poisson(lambda_: float) -> _Variant:
        return _Variant("poisson", {"lambda": float(lambda_)})

    @staticmethod
    def Uniform(min: int, max: int) -> _Variant:  # noqa: A002
        return _Variant("uniform", {"min": int(min), "max": int(max)})


def sample(family: str, variable: _Variant, replications: int, count: int, *, seed: int = 42, replication_offset: int = 0,
           device: int = 0, with_draws: bool = False):
    """``random_variate`` (random_variable.rs:65-131) on the device: ``count`` variates per replication of one random
    variable from the per-replication Philox streams the engine uses.  Returns an array [count, replications]
    (float64 for ``family == "continuous"``, uint64 otherwise), and the u64 draws each replication consumed when
    ``with_draws``."""
    import ctypes as C
    import json

    import numpy as np

    from .simulator import _check, load_library
    L = load_library()
    L.simb_sample_random_variable.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros((count, replications), np.float64 if family == "continuous" else np.uint64)
    draws = np.zeros(replications, np.uint32)
    _check(L.simb_sample_random_variable(family.encode(), json.dumps(variable.to_dict()).encode(), replications, count, seed,
                                         replication_offset, device, out.ctypes.data if family == "continuous" else None,
                                         None if family == "continuous" else out.ctypes.data, draws.ctypes.data))
    return (out, draws) if with_draws else out


class Thinning:
    """``Thinning`` (thinning.rs:6-34): a normalised polynomial thinning function, coefficients from the highest order
    down.  The reference's Generator stores it and never evaluates it (generator.rs:98-146); so does the engine."""

    def __init__(self, coefficients):
        self.coefficients = [float(c) for c in coefficients]

    def to_dict(self) -> dict:
        return {"function": {"polynomial": {"coefficients": list(self.coefficients)}}}

    def evaluate(self, point: float) -> float:
        import ctypes as C

        from .simulator import _check, load_library
        L = load_library()
        L.simb_thinning_evaluate.argtypes = [C.POINTER(C.c_double), C.c_uint64, C.c_double, C.POINTER(C.c_double)]
        a = (C.c_double * max(1, len(self.coefficients)))(*self.coefficients)
        out = C.c_double()
        _check(L.simb_thinning_evaluate(a, len(self.coefficients), float(point), C.byref(out)))
        return out.value
