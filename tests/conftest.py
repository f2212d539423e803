import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """Build the CPU oracle and the hostsim unit-test harness (both are test infrastructure)."""
    _build(["make", "-s"], os.path.join(ROOT, "oracle"))
    _build(["make", "-s"], os.path.join(ROOT, "tests", "hostsim"))
    yield
