"""Held-out scoring on the GPU -- the step right after the hot path in the reference CLI
(`measure_cv_performance`, prmf/__init__.py:768-798, called at script/prmf_runner.py:1074-1079), SURVEY.md
section 8(f) rank 4.  One batched non-negative least-squares solve for all held-out samples
(prmf_b200/csrc/cv.cu) instead of a Python loop over scipy.optimize.nnls.
"""
import ctypes

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def nnls_rows(V, X, device=None):
    """u_i = argmin_{u >= 0} ||X[i] - V u|| for every row of X.  V: (n, k) with k <= 128, X: (mt, n).
    Returns (U (mt, k), rnorm (mt,), xnorm (mt,)).  Raises when a sample hits scipy's iteration limit (3 k),
    as scipy.optimize.nnls does."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.PrmfLibraryError("nnls_rows needs a CUDA device (no CPU fallback)")
    lib = _lib.load()
    V = np.ascontiguousarray(V, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    if V.ndim != 2 or X.ndim != 2 or X.shape[1] != V.shape[0]:
        raise ValueError("Incompatible dimensions: V %s, X %s" % (V.shape, X.shape))
    if not np.isfinite(V).all() or not np.isfinite(X).all():
        raise ValueError("array must not contain infs or NaNs")            # np.asarray_chkfinite in scipy's nnls
    if X.strides[1] != 8 or X.strides[0] % 8 != 0 or X.strides[0] < X.shape[1] * 8:
        X = np.ascontiguousarray(X)
    n, k = V.shape
    mt = X.shape[0]
    U = np.zeros((mt, k)); rnorm = np.zeros(mt); xx = np.zeros(mt); status = np.zeros(mt, dtype=np.int32)
    dev = torch.cuda.current_device() if device is None else int(device)
    rc = lib.prmf_nnls_rows(dev, _ptr(V), n, k, _ptr(X), mt, X.strides[0] // 8 if mt else n, _ptr(U), _ptr(rnorm),
                            _ptr(xx), _ptr(status))
    if rc != 0:
        raise _lib.PrmfLibraryError("prmf_nnls_rows failed (%d): %s" % (rc, lib.prmf_cv_last_error().decode()))
    if (status == -1).any():
        raise RuntimeError("Maximum number of iterations reached.")          # scipy's message
    if (status == -2).any():          # (dependent columns are handled inside the solver; this is a numerical breakdown)
        raise np.linalg.LinAlgError("V^T V lost positive definiteness on a passive set")
    return U, rnorm, np.sqrt(xx)


def measure_cv_performance(gene_by_latent_train, data_test):
    """Drop-in for prmf.measure_cv_performance: normalised reconstruction error ||x - V u*|| / ||x|| per
    held-out sample."""
    _, rnorm, xnorm = nnls_rows(np.asarray(gene_by_latent_train), np.asarray(data_test))
    return rnorm / xnorm
