import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def fov_package():
    return importlib.import_module("foveated-360-video_b200")


@pytest.fixture(scope="session")
def fov():
    return fov_package()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def small():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "small.npz")))


@pytest.fixture(scope="session")
def oracle():
    import _oracle

    return _oracle.port()


@pytest.fixture(scope="session")
def mgr(fov):
    """One context for the whole GPU session = the reference's per-connection OpenCLManager."""
    m = fov.OpenCLManager(0)
    m.InitializeContext()  # raises FovError without a GPU: the gpu tests must not silently pass
    yield m
    m.close()
