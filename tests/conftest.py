import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu via gpurun)")
    config.addinivalue_line("markers", "slow: larger CPU case")


@pytest.fixture(scope="session")
def model_cache():
    """Construct each (kind, kwargs, seed) product model once per session."""
    import torch
    import moleculediffusiontransformer_b200 as mdt

    cache = {}

    def get(kind, kw, seed):
        key = (kind, tuple(sorted(kw.items())), seed)
        if key not in cache:
            torch.manual_seed(seed)
            cls = mdt.QMDiffusion if kind == "inverse" else mdt.QMDiffusionForward
            cache[key] = cls(**kw).eval()
        return cache[key]

    return get


def golden(name):
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
