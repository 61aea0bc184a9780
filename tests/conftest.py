"""pytest configuration: registers the `gpu` marker and puts the product package and the oracle on
sys.path.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 through the C-ABI library."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "efficient-llm-inference_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); runs the CUDA path through the C-ABI")
    config.addinivalue_line("markers", "slow: exhaustive sweeps (still CPU-only, minutes)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import kvq_oracle

    kvq_oracle.build()
    return kvq_oracle
