import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)["data"]


@pytest.fixture(scope="session")
def rules_goldens():
    return load_golden("rules_goldens.json")


@pytest.fixture(scope="session")
def search_goldens():
    return load_golden("search_goldens.json")


@pytest.fixture(scope="session")
def selfplay_goldens():
    return load_golden("selfplay_goldens.json")


@pytest.fixture(scope="session")
def nets_goldens():
    return load_golden("nets_goldens.json")


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (test infrastructure), built on demand with gcc."""
    from oracle import c4oracle

    c4oracle.build()
    return c4oracle


def golden_uniforms(run):
    """The uniforms np.random.choice consumed in a golden self-play run: one per (step, slot)."""
    import numpy as np

    np.random.seed(run["seed"])
    return np.random.random_sample(run["n_draws"] + run["E"] * 2)
