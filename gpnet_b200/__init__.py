"""gpnet_b200 -- B200-native (sm_100a) implementation of the dense GP hot path of filippocastelli/GPnet.

    from gpnet_b200 import GPnetRegressor, GPnetClassifier      # same constructor / methods as the reference
    import gpnet_b200; gpnet_b200.install_dropin()              # makes `from GPnetRegressor import GPnetRegressor` work

Importing the package never touches the GPU; the first numerical call loads ``libgpnet_b200.so`` and creates
the device context, and raises if either is unavailable (there is no CPU fallback).
"""
import sys

from .GPnet import COEFS, LAMBDAS, GPnetBase
from .GPnetClassifier import GPnetClassifier
from .GPnetRegressor import GPnetRegressor

__all__ = ["GPnetBase", "GPnetRegressor", "GPnetClassifier", "LAMBDAS", "COEFS", "install_dropin"]


def install_dropin():
    """Register the reference's flat module names (graph_test.py:6-7 imports) as aliases of this package's modules."""
    for name in ("GPnet", "GPnetRegressor", "GPnetClassifier"):
        sys.modules.setdefault(name, sys.modules[__name__ + "." + name])
