"""sim_b200 -- B200-native batched-replication engine for ndebuhr/sim discrete event simulations.

Host-side mirror of the reference's Rust API for the accelerated path
(``Simulation::post/put/reset/inject_input/step/step_n/step_until``, the ten built-in models, the
``input_modeling`` random variables and ``output_analysis``) over the C ABI of ``libsimb200.so``
(include/simb.h).  There is no CPU fallback: without the CUDA library the engine raises.
"""
from .coupling import Connector, Message  # noqa: F401
from .input_modeling import (BooleanRandomVariable, ContinuousRandomVariable, DiscreteRandomVariable,  # noqa: F401
                             IndexRandomVariable, Thinning)
from .models import (Batcher, Coupled, ExclusiveGateway, ExternalInputCoupling, ExternalOutputCoupling, Gate,  # noqa: F401
                     Generator, InternalCoupling, LoadBalancer, Metric, Model, ParallelGateway, Passive, Processor,
                     StochasticGate, Stopwatch, Storage)

__all__ = [
    "Connector", "Message", "BooleanRandomVariable", "ContinuousRandomVariable", "DiscreteRandomVariable", "IndexRandomVariable", "Thinning",
    "Batcher", "Coupled", "ExternalInputCoupling", "ExternalOutputCoupling", "InternalCoupling", "ExclusiveGateway", "Gate", "Generator", "LoadBalancer", "Metric", "Model", "ParallelGateway",
    "Passive", "Processor", "StochasticGate", "Stopwatch", "Storage",
]


def __getattr__(name):
    # Simulation / output_analysis load the CUDA library lazily so that describing models works
    # on machines without it.
    import importlib
    if name in ("Simulation", "WebSimulation", "SimulationError", "SimbConfig"):
        return getattr(importlib.import_module(".simulator", __name__), name)
    if name in ("output_analysis", "simulator"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
