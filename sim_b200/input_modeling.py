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
