import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C++ restatement of the reference path) — test infrastructure only."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def kz():
    """The product package (directory name contains a dash, hence importlib)."""
    return importlib.import_module("kanzi-go_b200")


@pytest.fixture(scope="session")
def gpu(kz):
    """Engine context on cuda:0. Fails (never skips silently to a CPU path) when the CUDA library cannot start."""
    return kz.Context(0)


@pytest.fixture(scope="session")
def synth(kz):
    return importlib.import_module("kanzi-go_b200.synth")
