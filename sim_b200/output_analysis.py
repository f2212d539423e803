"""``sim::output_analysis`` (sim/src/output_analysis/mod.rs) over the C ABI's host restatement
(sim_b200/csrc/output_analysis.cpp): IndependentSample, SteadyStateOutput, t_score, usize_sqrt."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .simulator import _CI, _check, load_library


class ConfidenceInterval:
    """output_analysis/mod.rs:52-72"""

    def __init__(self, lower: float, upper: float):
        self._lower, self._upper = lower, upper

    def lower(self) -> float:
        return self._lower

    def upper(self) -> float:
        return self._upper

    def half_width(self) -> float:
        return (self._upper - self._lower) / 2.0


class IndependentSample:
    """output_analysis/mod.rs:76-135"""

    def __init__(self, points):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        ci = _CI()
        # alpha is irrelevant for the moments; 0.05 is always a valid table column
        _check(load_library().simb_oa_independent_sample(self.points.ctypes.data, self.points.size, 0.05, C.byref(ci)))
        self.mean, self._variance = ci.mean, ci.variance

    @classmethod
    def post(cls, points) -> "IndependentSample":
        return cls(points)

    def confidence_interval_mean(self, alpha: float) -> ConfidenceInterval:
        ci = _CI()
        _check(load_library().simb_oa_independent_sample(self.points.ctypes.data, self.points.size, float(alpha), C.byref(ci)))
        return ConfidenceInterval(ci.lower, ci.upper)

    def point_estimate_mean(self) -> float:
        return self.mean

    def variance(self) -> float:
        return self._variance


class SteadyStateOutput:
    """output_analysis/mod.rs:175-345 (MSER-style deletion point, at most 30 batches)."""

    def __init__(self, time_series):
        self.time_series = np.ascontiguousarray(time_series, dtype=np.float64)
        self.deletion_point = self.batch_size = self.batch_count = None

    @classmethod
    def post(cls, time_series) -> "SteadyStateOutput":
        return cls(time_series)

    def _run(self, alpha):
        ci = _CI()
        d, bs, bc = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(load_library().simb_oa_steady_state(self.time_series.ctypes.data, self.time_series.size, float(alpha),
                                                   C.byref(ci), C.byref(d), C.byref(bs), C.byref(bc)))
        self.deletion_point, self.batch_size, self.batch_count = d.value, bs.value, bc.value
        self.batches_mean, self.batches_variance = ci.mean, ci.variance
        return ci

    def confidence_interval_mean(self, alpha: float) -> ConfidenceInterval:
        ci = self._run(alpha)
        return ConfidenceInterval(ci.lower, ci.upper)

    def point_estimate_mean(self) -> float:
        return self._run(0.05).mean


def batch_means_confidence_interval(mean: float, variance: float, batch_count: int, alpha: float) -> ConfidenceInterval:
    """``SteadyStateOutput::confidence_interval_mean`` (mod.rs:302-333) from batch statistics."""
    ci = _CI()
    _check(load_library().simb_oa_batch_means_ci(float(mean), float(variance), int(batch_count), float(alpha), C.byref(ci)))
    return ConfidenceInterval(ci.lower, ci.upper)


def t_score(alpha: float, df: int) -> float:
    """t_scores.rs:9-30"""
    out = C.c_double()
    _check(load_library().simb_oa_t_score(float(alpha), int(df), C.byref(out)))
    return out.value


def usize_sqrt(n: int) -> int:
    """utils/mod.rs:84-92"""
    return int(load_library().simb_usize_sqrt(int(n)))
