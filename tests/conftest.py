import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Parity numbers recorded by the tests (helpers.report) are printed after the run, so `pytest -q` shows them."""
    import helpers
    if helpers.REPORT:
        terminalreporter.write_line("measured parity numbers (max-abs error vs the oracle / the reference's goldens):")
        for name, value in helpers.REPORT:
            terminalreporter.write_line(f"  {name}: {value}")
