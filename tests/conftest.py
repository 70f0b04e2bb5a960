import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu via gpurun)")
    config.addinivalue_line("markers", "slow: larger CPU case")
    # the C-ABI library is a build artefact (git-ignored): compile it for sm_100a if this checkout has not done so yet
    lib = os.path.join(ROOT, "moleculediffusiontransformer_b200", "libmdt_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "moleculediffusiontransformer_b200", "csrc"), "-j4"], check=True,
                       stdout=subprocess.DEVNULL)


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
            cls = {"inverse": mdt.QMDiffusion, "forward": mdt.QMDiffusionForward, "analog_sparse": mdt.AnalogDiffusionSparse,
                   "analog_full": mdt.AnalogDiffusionFull}[kind]
            cache[key] = cls(**kw).eval()
        return cache[key]

    return get


def golden(name):
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
