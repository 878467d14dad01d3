"""GPU quantile normalisation -- the step right before the hot path in the reference CLI
(`X = quantile_transform(X)`, script/prmf_runner.py:1019-1020), SURVEY.md section 8(f) rank 1.

`quantile_transform(X)` mirrors `sklearn.preprocessing.quantile_transform(X)` with its defaults
(axis=0, n_quantiles=1000, output_distribution='uniform', subsample=100000): everything numpy derives on the
host (reference levels, order-statistic indices and interpolation weights, the row subsample drawn from the
global RNG when m > subsample) is computed here with the same numpy / sklearn calls and handed to the device
kernels (prmf_b200/csrc/preprocess.cu), which evaluate numpy's formulas without FMA contraction.
"""
import ctypes

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def quantile_transform(X, n_quantiles=1000, subsample=int(1e5), random_state=None, return_device=False,
                       return_quantiles=False, device=None):
    """X: (m, n) float64 numpy array or CUDA torch tensor.  Returns the transformed matrix as a numpy array
    (or as a CUDA tensor with `return_device=True`, ready for `nmf_pathway` without a round trip)."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.PrmfLibraryError("quantile_transform needs a CUDA device (no CPU fallback)")
    lib = _lib.load()
    if hasattr(X, "is_cuda"):
        Xd = X if X.is_cuda else X.cuda(device)
        if Xd.dtype != torch.float64:
            Xd = Xd.double()
    else:
        Xh = np.ascontiguousarray(X, dtype=np.float64)
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        Xd = torch.from_numpy(Xh).to(dev)
    if Xd.dim() != 2:
        raise ValueError("Expected 2D array, got %dD" % Xd.dim())
    Xd = Xd.contiguous()
    m, n = Xd.shape
    nq = max(1, min(int(n_quantiles), m))                              # sklearn: n_quantiles_
    refs = np.linspace(0, 1, nq, endpoint=True)                        # references_
    rows = None
    ms = m
    if subsample is not None and subsample < m:
        # the same call sklearn makes on X, applied to the row indices: identical rows, identical RNG use
        from sklearn.utils import check_random_state, resample
        rows = resample(np.arange(m, dtype=np.int64), replace=False, n_samples=subsample,
                        random_state=check_random_state(random_state))
        ms = int(subsample)
    # np.nanpercentile(col, refs*100): q = (refs*100)/100, virtual index (ms-1)*q, linear method
    q = np.true_divide(refs * 100, np.float64(100))
    vi = (ms - 1) * q
    lo = np.floor(vi).astype(np.int64)
    hi = lo + 1
    top = vi >= ms - 1
    lo[top] = ms - 1
    hi[top] = ms - 1
    g = vi - np.floor(vi)
    g[top] = 0.0
    out = torch.empty_like(Xd)
    rows_d = torch.from_numpy(np.ascontiguousarray(rows)).to(Xd.device) if rows is not None else None
    Q = np.empty((n, nq)) if return_quantiles else None
    stream = torch.cuda.current_stream(Xd.device)
    stream.synchronize()
    rc = lib.prmf_quantile_transform(
        Xd.device.index, ctypes.c_void_p(stream.cuda_stream), ctypes.c_void_p(Xd.data_ptr()), m, n, Xd.stride(0),
        ctypes.c_void_p(rows_d.data_ptr()) if rows_d is not None else None, ms, nq, _ptr(refs), _ptr(lo), _ptr(hi),
        _ptr(g), ctypes.c_void_p(out.data_ptr()), out.stride(0), _ptr(Q))
    if rc != 0:
        msg = lib.prmf_preprocess_last_error().decode()
        if "NaN" in msg:
            raise ValueError(msg)
        raise _lib.PrmfLibraryError("prmf_quantile_transform failed (%d): %s" % (rc, msg))
    res = out if return_device else out.cpu().numpy()
    return (res, Q.T.copy()) if return_quantiles else res
