"""gpnet_b200 -- B200-native (sm_100a) implementation of the dense GP hot path of filippocastelli/GPnet.

    from gpnet_b200 import GPnetRegressor, GPnetClassifier      # same constructor / methods as the reference
    import gpnet_b200; gpnet_b200.install_dropin()              # makes `from GPnetRegressor import GPnetRegressor` work

Importing the package never touches the GPU; the first numerical call loads ``libgpnet_b200.so`` and creates
the device context, and raises if either is unavailable (there is no CPU fallback).
"""
import os
import sys

# One evaluation of the default route uses up to 14 CUDA streams (one or two per part of the dissection, one for the other-node
# separator), two lanes twice that.  The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams
# that share a queue serialise, so that an independent GEMM waits behind another part's 0.7 ms cooperative factorisation
# (measured at C3: 8 -> 32 connections: 402 -> 569 evals/s with one lane, 479 -> 753 with two).  The variable is read when the CUDA
# context is created: it takes effect if this package (or the variable) is set before the process first touches the GPU.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .GPnet import COEFS, LAMBDAS, GPnetBase
from .GPnetClassifier import GPnetClassifier
from .GPnetRegressor import GPnetRegressor

__all__ = ["GPnetBase", "GPnetRegressor", "GPnetClassifier", "LAMBDAS", "COEFS", "install_dropin"]


def install_dropin():
    """Register the reference's flat module names (graph_test.py:6-7 imports) as aliases of this package's modules."""
    for name in ("GPnet", "GPnetRegressor", "GPnetClassifier"):
        sys.modules.setdefault(name, sys.modules[__name__ + "." + name])
