import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Measured parity errors of the GPU tests (tests/parity_log.py), printed even with -q and saved as JSON."""
    try:
        import parity_log
    except Exception:  # pragma: no cover
        return
    lines = parity_log.summary_lines()
    if not lines:
        return
    path = parity_log.dump()
    terminalreporter.write_sep("-", f"measured parity errors ({len(lines)} records -> {path})")
    for ln in lines:
        terminalreporter.write_line(ln)
