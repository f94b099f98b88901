import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


def load_golden(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rt") as handle:
        return json.load(handle)


@pytest.fixture(scope="session")
def golden():
    return load_golden
