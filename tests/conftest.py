import glob
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before CUDA is initialised: see gpnet_b200/__init__.py

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(path):
    g = np.load(path, allow_pickle=True)
    return {k: (g[k].item() if g[k].shape == () else g[k]) for k in g.files}


def nrel(a, b):
    """norm-wise relative error max|a-b| / max|b| (SURVEY 8c tolerance definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="session")
def gpu_ctx():
    """Library context sharing one non-default torch stream with the test's torch ops."""
    import torch
    from gpnet_b200._lib import Context
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(0, stream.cuda_stream)
    yield ctx
    ctx.close()
