import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_pkg():
    return importlib.import_module("voxel-hashing-sdf_b200")


def load_synth():
    return importlib.import_module("voxel-hashing-sdf_b200.synth")


@pytest.fixture(scope="session")
def vh():
    return load_pkg()


@pytest.fixture(scope="session")
def synth():
    return load_synth()


@pytest.fixture(scope="session")
def ob():
    from oracle import binding
    binding.build()
    return binding
