"""Run the UNMODIFIED reference classes (filippocastelli/GPnet, mounted read-only at /root/reference in
the build container) under the three shims of SURVEY.md Appendix B.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py); used by oracle/make_golden.py to produce tests/golden/*.npz and by
``bench.py --impl reference`` to time the literal reference on the GPU box's host cores.

/root/reference does not exist on the GPU box.  ``oracle/make_ref.py`` (run by ``__graft_entry__.build()`` in the
build container) compiles the three reference modules of the hot path, unmodified and from where they lie, to
bytecode under ``oracle/_ref/`` -- git-ignored build output, not gpurun-ignored, so it travels to the box like the
built ``.so``.  ``find_reference()`` returns whichever of the two locations exists; the -m gpu tests and ``smoke()``
never import this module.

Shims (none edits a reference file):
 1. a stub ``matplotlib`` package (the reference imports pyplot/pylab/lines at module level,
    GPnet.py:4,8; matplotlib is not installed here);
 2. ``GPnet.sl`` -> proxy of scipy.linalg whose ``expm`` accepts the sparse matrix the reference passes
    (GPnet.py:340, delegated to scipy.sparse.linalg.expm as 2018 scipy did) and whose ``inv`` returns
    an ndarray subclass with ``.toarray()`` (GPnet.py:342);
 3. the eager all-pairs shortest paths + Kamada-Kawai layout of ``__init__`` (GPnet.py:204-207) are
    patched out for speed (unused by the hot path).
"""
import importlib
import sys
import types

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

import os

REFERENCE_PATH = os.environ.get("GPNET_REFERENCE_PATH", "/root/reference")
VENDORED_EXT = ".bytecode"      # CPython bytecode under a neutral extension (snapshot tools drop *.pyc)
VENDORED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_MODULES = ("GPnet.py", "GPnetRegressor.py", "GPnetClassifier.py")


def find_reference():
    """Directory the reference modules import from: the read-only source mount (build container) or the compiled copy
    oracle/_ref (GPU box); None when neither exists."""
    if all(os.path.exists(os.path.join(REFERENCE_PATH, f)) for f in REFERENCE_MODULES):
        return REFERENCE_PATH
    if all(os.path.exists(os.path.join(VENDORED_PATH, f[:-3] + VENDORED_EXT)) for f in REFERENCE_MODULES):
        return VENDORED_PATH
    return None


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()

    def __iter__(self):
        return iter(())


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_gpnet_stub", False):
        return
    root = types.ModuleType("matplotlib")
    root._gpnet_stub = True
    root.__path__ = []
    for sub in ("pyplot", "pylab", "lines", "patches", "cm", "colors"):
        m = types.ModuleType("matplotlib." + sub)
        m.__getattr__ = lambda name, _m=sub: _Inert()
        setattr(root, sub, m)
        sys.modules["matplotlib." + sub] = m
    sys.modules["matplotlib"] = root


class _ArrayWithToarray(np.ndarray):
    def toarray(self):
        return np.asarray(self)


class _ScipyLinalgProxy:
    def __getattr__(self, name):
        return getattr(scipy.linalg, name)

    @staticmethod
    def expm(A):
        if scipy.sparse.issparse(A):
            return scipy.sparse.linalg.expm(scipy.sparse.csc_matrix(A))
        return scipy.linalg.expm(A)

    @staticmethod
    def inv(A):
        return np.asarray(scipy.linalg.inv(np.asarray(A))).view(_ArrayWithToarray)


def load_reference(path=None):
    """Returns (GPnet module, GPnetRegressor class, GPnetClassifier class) of the shimmed reference."""
    if path is None:
        path = find_reference()
    if path is None:
        raise ImportError("the reference modules are neither at /root/reference nor under oracle/_ref")
    _install_matplotlib_stub()
    if path not in sys.path:
        sys.path.insert(0, path)
    for name in ("GPnet", "GPnetRegressor", "GPnetClassifier"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(path):
            del sys.modules[name]
    if path == VENDORED_PATH:      # sourceless import of the compiled copies, in dependency order
        import importlib.machinery
        import importlib.util
        for name in ("GPnet", "GPnetRegressor", "GPnetClassifier"):
            if name in sys.modules:
                continue
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(path, name + VENDORED_EXT))
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            if name == "GPnet":       # the shims below must be in place before the subclasses import it
                pass
            loader.exec_module(mod)
    gp = importlib.import_module("GPnet")
    gp.sl = _ScipyLinalgProxy()
    gp.GPnetBase.calc_shortest_paths = lambda self: setattr(self, "dist", None)
    gp.nx.kamada_kawai_layout = lambda G, *a, **k: {}
    reg = importlib.import_module("GPnetRegressor").GPnetRegressor
    cls = importlib.import_module("GPnetClassifier").GPnetClassifier
    return gp, reg, cls
