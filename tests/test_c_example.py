"""examples/mm1_ci.c: the C ABI used from plain C99 (no Python, no torch types in any signature).
CPU tier: it compiles with -pedantic against include/simb.h, links libsimb200.so and -- without a GPU -- fails
loudly instead of falling back.  GPU tier: its result equals the Python host path bit for bit."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "mm1_ci")
    lib = os.path.join(ROOT, "sim_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "mm1_ci.c"), "-o", exe, "-L", lib, "-lsimb200",
                           "-Wl,-rpath," + lib])
    return exe


def test_c_example_builds_and_refuses_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "16", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and "CudaError" in r.stderr


@pytest.mark.gpu
def test_c_example_matches_the_python_host_path(tmp_path):
    import networks
    from sim_b200 import Simulation
    exe = _build(tmp_path)
    r = subprocess.run([exe, "4096", "2000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = re.search(r"steps (\d+) events (\d+)", r.stdout)
    ci = re.search(r"mean_wait (\S+) lower (\S+) upper (\S+) n (\d+)", r.stdout)
    models, conns = networks.mm1()
    sim = Simulation.post(models, conns, replications=4096, seed=42, stat_model_id="processor-01")
    sim.step_until(2000.0)
    c = sim.get_counters()
    want = sim.independent_sample_ci(0.05)
    assert (int(m.group(1)), int(m.group(2))) == (c["steps"], c["events"])
    assert float(ci.group(1)) == want["mean"] and float(ci.group(2)) == want["lower"] and float(ci.group(3)) == want["upper"]
    assert int(ci.group(4)) == 4096 and np.isfinite(want["mean"])
