import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(autouse=True)
def _clean_path_selection(monkeypatch):
    """Which kernel path a test exercises is chosen explicitly (`force_generic=` of the trainers /
    SVIEngine), never through the process environment: no state leaks between tests."""
    monkeypatch.delenv("PVB_FORCE_GENERIC", raising=False)
    yield


def record_margin(test, what, value, bound):
    """Append a measured error and the bound it was checked against to gpurun_out/margins.tsv
    (the tolerances in the GPU tests are justified by these measurements)."""
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "margins.tsv"), "a") as f:
            f.write("{}\t{}\t{:.3e}\t{:.1e}\n".format(test, what, value, bound))
    except OSError:
        pass
