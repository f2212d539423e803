"""The C-ABI library loads on a machine without a GPU, exports every symbol include/simb.h declares,
answers the host-only entry points, and refuses loudly (no CPU fallback) when asked to simulate."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from sim_b200 import simulator
    return simulator.load_library()


def test_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "simb.h")).read()
    names = set(re.findall(r"\b(simb_[a-z0-9_]+)\s*\(", header))
    names -= {"simb_sim"}
    assert len(names) >= 40
    lib = C.CDLL(os.path.join(ROOT, "sim_b200", "libsimb200.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.simb_abi_version() == 2


def test_struct_layout_matches_header():
    from sim_b200.simulator import SimbConfig, _CI, _Counters, _Message, _Record, _TraceEvent
    assert C.sizeof(SimbConfig) == 88 and C.sizeof(_Message) == 248 and C.sizeof(_Record) == 136
    assert C.sizeof(_TraceEvent) == 24 and C.sizeof(_Counters) == 56 and C.sizeof(_CI) == 40


def test_host_only_entry_points():
    from sim_b200 import output_analysis as oa
    ci = oa.IndependentSample.post([1.02, 0.73, 3.20, 0.23, 1.76, 0.47, 1.89, 1.45, 0.44, 0.23]).confidence_interval_mean(0.1)
    assert abs(ci.lower() - 0.7492630635369267) < 1e-12 and abs(ci.upper() - 1.534736936463073) < 1e-12
    assert [oa.usize_sqrt(n) for n in (1, 3, 4, 8, 9, 15)] == [1, 1, 2, 2, 3, 3]
    assert oa.t_score(0.1, 9) == 1.383
    from sim_b200 import SimulationError
    with pytest.raises(SimulationError) as ei:
        oa.t_score(0.3, 9)
    assert ei.value.name == "Panic"


def test_product_output_analysis_matches_oracle():
    import numpy as np
    import oracle_api as O
    from sim_b200 import output_analysis as oa
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 100, 101, 102, 5000):
        x = rng.exponential(size=n) + (10.0 if n > 50 else 0.0) * (np.arange(n) < n // 10)
        for alpha in (0.1, 0.05, 0.001):
            e, want = O.independent_sample(x, alpha)
            got = oa.IndependentSample.post(x).confidence_interval_mean(alpha)
            assert (got.lower(), got.upper()) == (want["lower"], want["upper"])
            e, want = O.steady_state(x, alpha)
            ss = oa.SteadyStateOutput.post(x)
            if e:  # n == 2: the first-half minimum is taken over an empty range -> PrerequisiteCalcError
                from sim_b200 import SimulationError
                with pytest.raises(SimulationError) as ei:
                    ss.confidence_interval_mean(alpha)
                assert ei.value.status == e == 10 and n == 2
                continue
            got = ss.confidence_interval_mean(alpha)
            assert (got.lower(), got.upper()) == (want["lower"], want["upper"])
            assert (ss.deletion_point, ss.batch_size, ss.batch_count) == (want["deletion_point"], want["batch_size"], want["batch_count"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import networks
    from sim_b200 import Simulation, SimulationError
    with pytest.raises(SimulationError) as ei:
        Simulation.post(*networks.mm1(), replications=4)
    assert ei.value.name == "CudaError" and "no CPU fallback" in str(ei.value)


def test_lowering_errors_come_before_the_device_check():
    from sim_b200 import Simulation, SimulationError
    with pytest.raises(SimulationError) as ei:
        Simulation.post_json("[{\"id\": \"c\", \"type\": \"MyCustomModel\"}]", "[]", 1)
    assert ei.value.name == "UnsupportedModel"   # custom (sim_derive) models cannot run on the device
    nested = {"id": "outer", "type": "Coupled", "portsIn": {"flowPaths": []}, "portsOut": {"flowPaths": []},
              "components": [{"id": "inner", "type": "Coupled"}], "externalInputCouplings": [],
              "externalOutputCouplings": [], "internalCouplings": []}
    import json
    with pytest.raises(SimulationError) as ei:
        Simulation.post_json(json.dumps([nested]), "[]", 1)
    assert ei.value.name == "UnsupportedModel"   # one level of coupling is lowered; nesting is refused explicitly


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or execute anything under oracle/ or tests/hostsim."""
    pkg = os.path.join(ROOT, "sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_api" not in text and "hostsim" not in text.replace(
                    "tests/hostsim", "").replace("hostsim:", "").replace("hostsim (", ""), f
