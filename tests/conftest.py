import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def gold_json(name):
    return json.load(open(os.path.join(GOLD, name)))


def gold_npy(name):
    return np.load(os.path.join(GOLD, name))


def schwarz(name):
    return np.unpackbits(gold_npy(name + "_bits.npy")).reshape(128, 128, 128)


@pytest.fixture(scope="session")
def K():
    """The product package, bound to cuda:0.  Raises (does not skip) when the CUDA library is absent."""
    import keff_cfd_b200 as K
    K.init(0)
    return K


@pytest.fixture(scope="session")
def R():
    """The checker: oracle restatement (+ unmodified reference when oracle/_ref is present)."""
    from oracle import refbind
    refbind.build()
    return refbind
