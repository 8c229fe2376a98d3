"""ctypes binding of libgpnet_b200.so (the C-ABI declared in include/gpnet_b200.h).

There is NO CPU fallback: if the shared library is missing or no sm_100a GPU is visible, loading /
context creation raises.  torch is used only as plumbing by callers (device tensors for the unit-level
primitives, torch.distributed for the sharded landscape); the drop-in classes work without importing it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpnet_b200.so")

KERNEL_IDS = {"diffusion": 0, "regularized_laplacian": 1, "pstep_walk": 2}

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)

# name -> (restype, argtypes); every symbol declared in include/gpnet_b200.h
SIGNATURES = {
    "gpn_create": (C.c_void_p, [C.c_int, C.c_void_p]),
    "gpn_destroy": (None, [C.c_void_p]),
    "gpn_last_error": (C.c_char_p, []),
    "gpn_sync": (C.c_int, [C.c_void_p]),
    "gpn_launch_count": (C.c_longlong, [C.c_void_p]),
    "gpn_flop_count": (C.c_double, [C.c_void_p]),
    "gpn_reset_counters": (None, [C.c_void_p]),
    "gpn_fp64_peak": (C.c_int, [C.c_void_p, C.c_double, c_dbl_p]),
    "gpn_set_profile": (None, [C.c_void_p, C.c_int]),
    "gpn_profile_summary": (C.c_int, [C.c_void_p, c_dbl_p]),
    "gpn_spectral_prepare": (C.c_int, [C.c_void_p, c_int_p, c_dbl_p]),
    "gpn_spectral_enable": (None, [C.c_void_p, C.c_int]),
    "gpn_spectral_get": (C.c_int, [C.c_void_p, c_dbl_p, c_dbl_p]),
    "gpn_gpr_logp_grad_batch": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, C.c_int, c_dbl_p, C.c_int, c_dbl_p, c_int_p]),
    "gpn_gpc_newton_batch": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, C.c_int, c_dbl_p, C.c_double, C.c_double, C.c_double, c_dbl_p,
                                      c_int_p, c_int_p]),
    "gpn_profile_intervals": (C.c_int, [C.c_void_p, C.c_void_p, c_dbl_p, C.c_int, c_int_p]),
    "gpn_debug_leaf_prof": (None, [C.c_void_p, C.c_void_p]),
    "gpn_debug_scalars": (C.c_int, [C.c_void_p, c_dbl_p]),
    "gpn_debug_d6_layout": (C.c_int, [C.c_void_p, c_int_p]),
    "gpn_debug_d6_buffer": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]),
    "gpn_debug_df_trace": (None, [C.c_void_p, C.c_void_p]),
    "gpn_debug_leaf_bench": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpn_set_graph": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_dbl_p]),
    "gpn_set_nodes": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.c_int, c_int_p]),
    "gpn_set_targets": (C.c_int, [C.c_void_p, c_dbl_p, C.c_int]),
    "gpn_laplacian_dense": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_long]),
    "gpn_gemm_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_long,
                               C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_long]),
    "gpn_potrf_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, c_int_p]),
    "gpn_spd_inverse_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, c_int_p]),
    "gpn_syrk_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_long]),
    "gpn_posv_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long,
                               C.POINTER(C.c_int)]),
    "gpn_trsm_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int]),
    "gpn_potri_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long]),
    "gpn_expm_sym_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, c_int_p, c_int_p]),
    "gpn_matpow_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long]),
    "gpn_last_route": (C.c_int, [C.c_void_p, c_int_p]),
    "gpn_set_dissection_parts": (None, [C.c_void_p, C.c_int]),
    "gpn_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpn_gather_block_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                       C.c_void_p, C.c_long]),
    "gpn_kernel_build": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, C.c_int]),
    "gpn_kernel_ptr": (C.c_void_p, [C.c_void_p, c_long_p]),
    "gpn_expm_info": (None, [C.c_void_p, c_int_p, c_int_p]),
    "gpn_kernel_block_h": (C.c_int, [C.c_void_p, c_int_p, C.c_int, c_int_p, C.c_int, C.c_double, c_dbl_p]),
    "gpn_gpr_logp_grad": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, c_int_p]),
    "gpn_gpr_logp_grad_begin": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, C.c_int]),
    "gpn_gpr_logp_grad_end": (C.c_int, [C.c_void_p, c_dbl_p, c_int_p]),
    "gpn_set_sm_budget": (None, [C.c_void_p, C.c_int]),
    "gpn_num_sms": (C.c_int, [C.c_void_p]),
    "gpn_set_fast_paths": (None, [C.c_void_p, C.c_int]),
    "gpn_gpr_predict": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p]),
    "gpn_gpr_fetch": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p]),
    "gpn_gpc_newton": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, C.c_int, C.c_void_p, C.c_long, C.c_double,
                                 C.c_double, C.c_double, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, c_int_p]),
    "gpn_gpc_predict": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p,
                                  c_dbl_p, c_dbl_p, c_int_p]),
    "gpn_gpc_grad": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p, c_dbl_p, c_int_p, c_int_p]),
    "gpn_landscape": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_int, c_dbl_p, C.c_int, c_dbl_p,
                                C.c_int, c_dbl_p, C.c_long, C.c_long, C.c_long, C.c_int, c_dbl_p]),
    "gpn_landscape_points": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_int, c_dbl_p, C.c_int, c_dbl_p,
                                       C.c_int, c_dbl_p, C.POINTER(C.c_long), C.c_long, C.c_int, c_dbl_p]),
}

_lib = None


class GpnetError(RuntimeError):
    """A C-ABI call returned a negative status (bad argument or CUDA error)."""


def load_library(path: str = LIB_PATH):
    """dlopen the in-tree shared library and attach the prototypes.  Raises if it is missing (build it
    with ``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C gpnet_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: the CUDA library has not been built (no CPU fallback exists)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def workspace_bytes(N: int, n: int, m: int, kerneltype: str) -> int:
    """Device memory a context needs for a model of this size (host arithmetic in the library, no GPU touched)."""
    return int(load_library().gpn_workspace_bytes(int(N), int(n), int(m), KERNEL_IDS[kerneltype]))


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_dbl_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int_p)


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))


def _eth(theta):
    """``np.exp(theta)`` -- the reference's own expression (GPnet.py:303) evaluated on the HOST.  The C-ABI takes the
    exponentiated parameters so that ``int(np.exp(theta)[1])`` (the p-step order, quirk A-4) and the domain checks see
    exactly numpy's values: numpy's exp and libm's exp differ by an ulp on some inputs (e.g. log 146, log 179)."""
    return np.ascontiguousarray(np.exp(np.asarray(theta, dtype=np.float64).reshape(-1)))


class Context:
    """One library context (one GPU, one stream).  Thin, explicit wrapper: every method is one C-ABI call."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        self.h = self.lib.gpn_create(int(device), C.c_void_p(stream) if stream else None)
        if not self.h:
            raise GpnetError("gpn_create failed: " + self.lib.gpn_last_error().decode())
        self.device = int(device)
        self.N = 0
        self.n_train = 0
        self.n_test = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- status handling ---------------------------------------------------------------------------
    def _check(self, st: int, what: str):
        if st < 0:
            raise GpnetError(f"{what} failed with status {st}: {self.lib.gpn_last_error().decode()}")
        return st

    # -- graph / node sets ---------------------------------------------------------------------------
    def set_graph(self, indptr, indices, weights=None):
        indptr, indices = _i32(indptr), _i32(indices)
        N = len(indptr) - 1
        w = None if weights is None else _f64(weights)
        if w is not None and np.all(w == 1.0):
            w = None
        self._check(self.lib.gpn_set_graph(self.h, N, _ip(indptr), _ip(indices), None if w is None else _dp(w)), "gpn_set_graph")
        self.N = N
        self.n_train = self.n_test = 0

    def set_nodes(self, train, test=()):
        tr, te = _i32(sorted(int(x) for x in train)), _i32(sorted(int(x) for x in test))
        self._check(self.lib.gpn_set_nodes(self.h, len(tr), _ip(tr), len(te), _ip(te)), "gpn_set_nodes")
        self.n_train, self.n_test = len(tr), len(te)

    def set_targets(self, t):
        tt = _f64(t)
        self._check(self.lib.gpn_set_targets(self.h, _dp(tt), len(tt)), "gpn_set_targets")

    # -- counters -----------------------------------------------------------------------------------
    def sync(self):
        self._check(self.lib.gpn_sync(self.h), "gpn_sync")

    def launch_count(self) -> int:
        return int(self.lib.gpn_launch_count(self.h))

    def flop_count(self) -> float:
        return float(self.lib.gpn_flop_count(self.h))

    def reset_counters(self):
        self.lib.gpn_reset_counters(self.h)

    def fp64_peak(self, seconds=0.25) -> float:
        out = C.c_double(0)
        self._check(self.lib.gpn_fp64_peak(self.h, float(seconds), C.byref(out)), "gpn_fp64_peak")
        return out.value

    def set_profile(self, on: bool):
        self.lib.gpn_set_profile(self.h, int(bool(on)))

    def profile_summary(self):
        out = np.zeros(11)
        self._check(self.lib.gpn_profile_summary(self.h, _dp(out)), "gpn_profile_summary")
        return dict(gemm_ms=out[0], gemm_flops=out[1], gemm_launches=int(out[2]), big_ms=out[3], big_flops=out[4],
                    big_launches=int(out[5]), potrf_ms=out[6], potrf_flops=out[7], potrf_launches=int(out[8]),
                    potrf_union_ms=out[9], gemm_union_ms=out[10])

    def profile_intervals(self, ref=None, max_rows=4096) -> np.ndarray:
        """(start ms, end ms, executed flops, kind) of every launch recorded since set_profile(True), relative to the first
        recorded launch of the context ``ref`` (lanes evaluating concurrently are merged by the caller)."""
        out = np.zeros((max_rows, 4))
        rows = C.c_int(0)
        self._check(self.lib.gpn_profile_intervals(self.h, ref.h if ref is not None else None, _dp(out), max_rows, C.byref(rows)),
                    "gpn_profile_intervals")
        return out[: rows.value].copy()

    # -- device-pointer primitives (pointers are ints, e.g. torch.Tensor.data_ptr()) --------------------
    def laplacian_dense(self, alpha, gamma, out_ptr, ld):
        self._check(self.lib.gpn_laplacian_dense(self.h, alpha, gamma, C.c_void_p(out_ptr), ld), "gpn_laplacian_dense")

    def gemm(self, transa, transb, M, N, K, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
        self._check(self.lib.gpn_gemm_f64(self.h, int(transa), int(transb), M, N, K, alpha, C.c_void_p(a_ptr), lda, C.c_void_p(b_ptr),
                                          ldb, beta, C.c_void_p(c_ptr), ldc), "gpn_gemm_f64")

    def syrk(self, n, k, alpha, a_ptr, lda, beta, c_ptr, ldc):
        self._check(self.lib.gpn_syrk_f64(self.h, n, k, alpha, C.c_void_p(a_ptr), lda, beta, C.c_void_p(c_ptr), ldc), "gpn_syrk_f64")

    def posv(self, n, a_ptr, lda, b_ptr=None, ldb=0, brows=0, wt_ptr=None, ldw=0) -> int:
        """Cholesky of A in place with B <- B L^-T and Winv_t <- L^-T in the same kernel; returns the LAPACK-style info."""
        info = C.c_int(0)
        self._check(self.lib.gpn_posv_f64(self.h, n, C.c_void_p(a_ptr), lda, C.c_void_p(b_ptr) if b_ptr else None, ldb, brows,
                                          C.c_void_p(wt_ptr) if wt_ptr else None, ldw, C.byref(info)), "gpn_posv_f64")
        return info.value

    def trsm(self, n, l_ptr, ldl, b_ptr, ldb, brows):
        """B <- B L^-T with the factor of the immediately preceding potrf()/posv()."""
        self._check(self.lib.gpn_trsm_f64(self.h, n, C.c_void_p(l_ptr), ldl, C.c_void_p(b_ptr), ldb, brows), "gpn_trsm_f64")

    def potri(self, n, l_ptr, ldl, out_ptr, ldo):
        """out <- (L L^T)^-1 from the factor of the immediately preceding potrf()/posv()."""
        self._check(self.lib.gpn_potri_f64(self.h, n, C.c_void_p(l_ptr), ldl, C.c_void_p(out_ptr), ldo), "gpn_potri_f64")

    def expm_sym(self, N, a_ptr, lda, out_ptr, ldo):
        """out <- expm(A), A symmetric negative semi-definite; returns (Pade order, squarings)."""
        m, s = C.c_int(0), C.c_int(0)
        self._kernel_key = None
        self._check(self.lib.gpn_expm_sym_f64(self.h, N, C.c_void_p(a_ptr), lda, C.c_void_p(out_ptr), ldo, C.byref(m), C.byref(s)),
                    "gpn_expm_sym_f64")
        return m.value, s.value

    def matpow(self, N, b_ptr, ldb, p, out_ptr, ldo, out_pm1_ptr=None, ldo1=0):
        self._kernel_key = None
        self._check(self.lib.gpn_matpow_f64(self.h, N, C.c_void_p(b_ptr), ldb, int(p), C.c_void_p(out_ptr), ldo,
                                            C.c_void_p(out_pm1_ptr) if out_pm1_ptr else None, ldo1), "gpn_matpow_f64")

    def last_route(self):
        """(level, parts) of the last regressor likelihood evaluation: parts = [number of parts, separator size, sizes...]."""
        parts = (C.c_int * 16)()
        level = int(self.lib.gpn_last_route(self.h, parts))
        k = parts[0]
        return level, [int(x) for x in parts[: 2 + k]]

    def potrf(self, n, a_ptr, lda) -> int:
        info = C.c_int(0)
        self._check(self.lib.gpn_potrf_f64(self.h, n, C.c_void_p(a_ptr), lda, C.byref(info)), "gpn_potrf_f64")
        return info.value

    def spd_inverse(self, n, a_ptr, lda, out_ptr, ldo, work_ptr) -> int:
        info = C.c_int(0)
        self._check(self.lib.gpn_spd_inverse_f64(self.h, n, C.c_void_p(a_ptr), lda, C.c_void_p(out_ptr), ldo, C.c_void_p(work_ptr),
                                                 C.byref(info)), "gpn_spd_inverse_f64")
        return info.value

    def gather_block_add(self, k_ptr, ld, rows_ptr, nr, cols_ptr, nc, addc, out_ptr, ldo):
        self._check(self.lib.gpn_gather_block_add(self.h, C.c_void_p(k_ptr), ld, C.c_void_p(rows_ptr), nr, C.c_void_p(cols_ptr), nc, addc,
                                                  C.c_void_p(out_ptr), ldo), "gpn_gather_block_add")

    # -- full kernel ---------------------------------------------------------------------------------
    def kernel_build(self, kind: int, theta, want_grad=False) -> int:
        th = _eth(theta)
        st = self.lib.gpn_kernel_build(self.h, kind, _dp(th), len(th), int(want_grad))
        if st in (-100, -101):
            return st
        return self._check(st, "gpn_kernel_build")

    def kernel_ptr(self):
        ld = C.c_long(0)
        p = self.lib.gpn_kernel_ptr(self.h, C.byref(ld))
        return p, ld.value

    def expm_info(self):
        m, s = C.c_int(0), C.c_int(0)
        self.lib.gpn_expm_info(self.h, C.byref(m), C.byref(s))
        return m.value, s.value

    def kernel_block(self, rows, cols, addc) -> np.ndarray:
        r, c = _i32(rows), _i32(cols)
        out = np.empty((len(r), len(c)), dtype=np.float64)
        if out.size:
            st = self.lib.gpn_kernel_block_h(self.h, _ip(r), len(r), _ip(c), len(c), float(addc), _dp(out))
            if st in (-2, -4):      # the reference's np.delete / fancy indexing raises IndexError for such labels
                raise IndexError(f"node position out of bounds for a graph with {self.N} nodes")
            self._check(st, "gpn_kernel_block_h")
        return out

    # -- regression ----------------------------------------------------------------------------------
    def gpr_logp_grad(self, kind, theta, t, want_grad):
        """t=None: use the resident targets of set_targets()."""
        th = _eth(theta)
        tt = None if t is None else _f64(t)
        if tt is not None and len(tt) != self.n_train:
            raise ValueError("len(t) != number of training nodes")
        out = np.zeros(1 + len(th), dtype=np.float64)
        info = C.c_int(0)
        st = self.lib.gpn_gpr_logp_grad(self.h, kind, _dp(th), len(th), None if tt is None else _dp(tt), int(want_grad), _dp(out),
                                        C.byref(info))
        if st in (-100, -101):
            return st, out, 0
        self._check(st, "gpn_gpr_logp_grad")
        return 0, out, info.value

    def gpr_logp_grad_begin(self, kind, theta, t, want_grad):
        """Enqueue one evaluation (asynchronous on the level-6 route); pair with gpr_logp_grad_end.  Returns the status
        (-100 / -101 for parameter-domain violations, like gpr_logp_grad)."""
        th = _eth(theta)
        tt = None if t is None else _f64(t)
        if tt is not None and len(tt) != self.n_train:
            raise ValueError("len(t) != number of training nodes")
        self._pending_P = len(th)
        st = self.lib.gpn_gpr_logp_grad_begin(self.h, kind, _dp(th), len(th), None if tt is None else _dp(tt), int(want_grad))
        if st in (-100, -101):
            return st
        return self._check(st, "gpn_gpr_logp_grad_begin")

    def gpr_logp_grad_end(self):
        out = np.zeros(1 + self._pending_P, dtype=np.float64)
        info = C.c_int(0)
        self._check(self.lib.gpn_gpr_logp_grad_end(self.h, _dp(out), C.byref(info)), "gpn_gpr_logp_grad_end")
        return out, info.value

    # -- spectral fast path ----------------------------------------------------------------------------
    SPECTRAL_BATCH_MAX_N = 112

    def spectral_prepare(self):
        """Eigendecomposition of the bound graph's normalized Laplacian (once per graph); returns (sweeps, offdiag_rel)."""
        sweeps, off = C.c_int(0), C.c_double(0)
        self._check(self.lib.gpn_spectral_prepare(self.h, C.byref(sweeps), C.byref(off)), "gpn_spectral_prepare")
        return sweeps.value, off.value

    def spectral_enable(self, on=True):
        self.lib.gpn_spectral_enable(self.h, int(bool(on)))

    def spectral_get(self):
        lam, VT = np.zeros(self.N), np.zeros((self.N, self.N))
        self._check(self.lib.gpn_spectral_get(self.h, _dp(lam), _dp(VT)), "gpn_spectral_get")
        return lam, VT

    def gpr_logp_grad_batch(self, kind, thetas, t, want_grad=True):
        """(-logp, -grad) rows for B log-parameter vectors in one launch (one CTA per theta); t=None: resident targets."""
        th = np.ascontiguousarray(np.exp(np.asarray(thetas, dtype=np.float64)))
        if th.ndim != 2:
            raise ValueError("thetas must be B x P")
        B, P = th.shape
        tt = None if t is None else _f64(t)
        if tt is not None and len(tt) != self.n_train:
            raise ValueError("len(t) != number of training nodes")
        out = np.zeros((B, 1 + P))
        info = np.zeros(B, dtype=np.int32)
        st = self.lib.gpn_gpr_logp_grad_batch(self.h, kind, _dp(th), B, P, None if tt is None else _dp(tt), int(want_grad), _dp(out),
                                              info.ctypes.data_as(c_int_p))
        if st in (-100, -101, -20):
            return st, out, info
        self._check(st, "gpn_gpr_logp_grad_batch")
        return 0, out, info

    def gpc_newton_batch(self, kind, thetas, y, tol=0.1, phif=1e100, scale=1.0):
        """logq, Newton iteration counts and Cholesky infos of the Laplace loop for B log-parameter vectors in one launch."""
        th = np.ascontiguousarray(np.exp(np.asarray(thetas, dtype=np.float64)))
        if th.ndim != 2:
            raise ValueError("thetas must be B x P")
        B, P = th.shape
        yy = _f64(y)
        if len(yy) != self.n_train:
            raise ValueError("len(y) != number of training nodes")
        logq = np.zeros(B)
        iters = np.zeros(B, dtype=np.int32)
        info = np.zeros(B, dtype=np.int32)
        st = self.lib.gpn_gpc_newton_batch(self.h, kind, _dp(th), B, P, _dp(yy), tol, phif, scale, _dp(logq), iters.ctypes.data_as(c_int_p),
                                           info.ctypes.data_as(c_int_p))
        if st in (-100, -101, -20):
            return st, logq, iters, info
        self._check(st, "gpn_gpc_newton_batch")
        return 0, logq, iters, info

    def debug_scalars(self):
        out = np.zeros(32)
        self.lib.gpn_debug_scalars(self.h, _dp(out))
        return out

    def debug_d6_layout(self):
        out = (C.c_int * 40)()
        self.lib.gpn_debug_d6_layout(self.h, out)
        v = [int(x) for x in out]
        K = v[0]
        return dict(K=K, s=v[1], so=v[2], ND=v[3], off=v[4:13][: K + 1], nk=v[13:21][:K], ok=v[21:29][:K], c0=v[29:37][:K])

    def debug_d6_buffer(self, which):
        n = C.c_longlong(0)
        p = self.lib.gpn_debug_d6_buffer(self.h, int(which), C.byref(n))
        return p, n.value

    def set_sm_budget(self, sms=0):
        self.lib.gpn_set_sm_budget(self.h, int(sms))

    def num_sms(self) -> int:
        return int(self.lib.gpn_num_sms(self.h))

    def set_fast_paths(self, level=True):
        """Route of the regularized-Laplacian regressor likelihood: 0/False full kernel, 1 block route, 2 Schur route,
        3 Schur route with sparse coupling, 4 the same with the triangular work appended to the Cholesky kernels, 5 the
        same with M_oo^-1 from two independent halves of a dissection of the other nodes, 6/True (default) no dense Schur
        complement at all: K-way dissection of the whole graph, one augmented factorisation per part (falls back to 5, 4)."""
        self.lib.gpn_set_fast_paths(self.h, 6 if level is True else int(level))

    def set_dissection_parts(self, parts=0):
        """Parts of the whole-graph dissection of route level 6 (0: from the graph size; 2..8: forced, small graphs allowed)."""
        self.lib.gpn_set_dissection_parts(self.h, int(parts))

    def gpr_predict(self, kind, theta, t):
        th, tt = _eth(theta), _f64(t)
        if len(tt) != self.n_train:
            raise ValueError("len(t) != number of training nodes")
        n, m = self.n_train, self.n_test
        alpha, fstar, s, kss = (np.zeros(n), np.zeros(m), np.zeros(m), np.zeros(m))
        info = (C.c_int * 2)(0, 0)
        st = self.lib.gpn_gpr_predict(self.h, kind, _dp(th), len(th), _dp(tt), _dp(alpha), _dp(fstar), _dp(s), _dp(kss), info)
        if st in (-100, -101):
            return st, None
        self._check(st, "gpn_gpr_predict")
        return 0, dict(alpha=alpha, fstar=fstar, s=s, kstarstar_diag=kss, info_k=info[0], info_kss=info[1])

    def gpr_fetch(self, which: str) -> np.ndarray:
        n, m = self.n_train, self.n_test
        shape = {"k": (n, n), "kstar": (m, n), "L": (n, n), "v": (n, m), "kstarstar": (m, m)}[which]
        idx = {"k": 0, "kstar": 1, "L": 2, "v": 3, "kstarstar": 4}[which]
        out = np.empty(shape, dtype=np.float64)
        if out.size:
            self._check(self.lib.gpn_gpr_fetch(self.h, idx, _dp(out)), "gpn_gpr_fetch")
        return out

    # -- classification ------------------------------------------------------------------------------
    def gpc_newton(self, kind, theta, y, tol=0.1, phif=1e100, scale=1.0, variant=0, K_ptr=None, ldk=0):
        th, yy = _eth(theta), _f64(y)
        n = len(yy)
        f, a = np.zeros(n), np.zeros(n)
        logq, iters, info = C.c_double(0), C.c_int(0), C.c_int(0)
        st = self.lib.gpn_gpc_newton(self.h, kind, _dp(th), len(th), _dp(yy), n, C.c_void_p(K_ptr) if K_ptr else None, ldk, tol, phif,
                                     scale, variant, _dp(f), _dp(a), C.byref(logq), C.byref(iters), C.byref(info))
        if st in (-100, -101):
            return st, None
        self._check(st, "gpn_gpc_newton")
        return 0, dict(f=f, a=a, logq=logq.value, iters=iters.value, info=info.value)

    def gpc_predict(self, kind, theta, y):
        th, yy = _eth(theta), _f64(y)
        n, m = self.n_train, self.n_test
        if len(yy) != n:
            raise ValueError("len(y) != number of training nodes")
        f, a, fstar, V, pi = np.zeros(n), np.zeros(n), np.zeros(m), np.zeros(m), np.zeros(m)
        logq, iters, info = C.c_double(0), C.c_int(0), C.c_int(0)
        st = self.lib.gpn_gpc_predict(self.h, kind, _dp(th), len(th), _dp(yy), _dp(f), _dp(a), C.byref(logq), C.byref(iters), _dp(fstar),
                                      _dp(V), _dp(pi), C.byref(info))
        if st in (-100, -101):
            return st, None
        self._check(st, "gpn_gpc_predict")
        return 0, dict(f=f, a=a, logq=logq.value, iters=iters.value, info=info.value, fstar=fstar, V=V, pi_star=pi)

    def gpc_grad(self, kind, theta, y):
        th, yy = _eth(theta), _f64(y)
        if len(yy) != self.n_train:
            raise ValueError("len(y) != number of training nodes")
        out = np.zeros(1 + len(th))
        iters, info = C.c_int(0), C.c_int(0)
        st = self.lib.gpn_gpc_grad(self.h, kind, _dp(th), len(th), _dp(yy), _dp(out), C.byref(iters), C.byref(info))
        if st in (-100, -101):
            return st, out, 0, 0
        self._check(st, "gpn_gpc_grad")
        return 0, out, iters.value, info.value

    # -- landscape slice ------------------------------------------------------------------------------
    def landscape(self, kind, classifier, theta, axidx, ax1, ax2, t, begin, end, stride, want_grad=False):
        th, a1, a2, tt = _eth(theta), _eth(ax1), _eth(ax2), _f64(t)
        total = len(a1) * len(a2)
        end = min(end, total)
        count = max(0, (end - begin + stride - 1) // stride)
        out = np.zeros((count, 1 + len(th)), dtype=np.float64)
        if count:
            st = self.lib.gpn_landscape(self.h, kind, int(classifier), _dp(th), len(th), int(axidx[0]), int(axidx[1]), _dp(a1), len(a1),
                                        _dp(a2), len(a2), _dp(tt), begin, end, stride, int(want_grad), _dp(out))
            if st in (-100, -101):
                return st, out
            self._check(st, "gpn_landscape")
        return 0, out

    def landscape_points(self, kind, classifier, theta, axidx, ax1, ax2, t, indices, want_grad=False):
        """(status, (len(indices), 1+P) array) for an explicit list of flat grid indices (i1*len(ax2) + i2)."""
        th, a1, a2, tt = _eth(theta), _eth(ax1), _eth(ax2), _f64(t)
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.zeros((len(idx), 1 + len(th)), dtype=np.float64)
        if len(idx):
            st = self.lib.gpn_landscape_points(self.h, kind, int(classifier), _dp(th), len(th), int(axidx[0]), int(axidx[1]), _dp(a1),
                                               len(a1), _dp(a2), len(a2), _dp(tt), idx.ctypes.data_as(C.POINTER(C.c_long)), len(idx),
                                               int(want_grad), _dp(out))
            if st in (-100, -101):
                return st, out
            self._check(st, "gpn_landscape_points")
        return 0, out
