import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def tiny_params():
    """reduced-width networks with the yaml's topology (channel_mult, attention levels, heads>1)"""
    return dict(model_channels=64, num_heads=4, context_dim=64)
