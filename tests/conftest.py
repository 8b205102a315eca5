import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("small_ckg.npz")


@pytest.fixture(scope="session")
def golden_g12():
    return load_golden("g1_g2.npz")


@pytest.fixture(scope="session")
def golden_eval():
    return load_golden("eval_small.npz")


@pytest.fixture(scope="session")
def golden_g3():
    return load_golden("g3_metric.npz")


def params_from_golden(g, dtype=None):
    import torch as th
    p = {k[len("param."):]: th.from_numpy(v.copy()) for k, v in g.items() if k.startswith("param.")}
    if dtype is not None:
        p = {k: v.to(dtype) for k, v in p.items()}
    return p
