"""Held-out scoring: the batched GPU NNLS against scipy.optimize.nnls, the routine the
reference's measure_cv_performance calls per sample (prmf/__init__.py:795).  u relative 1e-8 (absolute 1e-10 for
coefficients on the boundary), normalised error relative 1e-9."""
import numpy as np
import pytest


def _scipy_cv(V, X):
    import scipy.optimize
    U = np.zeros((X.shape[0], V.shape[1])); err = np.zeros(X.shape[0])
    for i in range(X.shape[0]):
        U[i], e = scipy.optimize.nnls(V, X[i])
        err[i] = e / np.linalg.norm(X[i])
    return U, err


@pytest.mark.gpu
@pytest.mark.parametrize("mt,n,k,seed", [(40, 300, 6, 0), (257, 1000, 10, 1), (5, 64, 1, 2), (33, 500, 33, 3),
                                          (64, 900, 64, 4), (40, 1200, 128, 5)])
def test_batched_nnls_matches_scipy(mt, n, k, seed):
    from prmf_b200 import measure_cv_performance, nnls_rows
    rng = np.random.Generator(np.random.PCG64(seed))
    V = rng.random((n, k)) * (rng.random((n, k)) < 0.7)              # sparse-ish non-negative factors
    Utrue = rng.random((mt, k)) * (rng.random((mt, k)) < 0.5)         # many coefficients exactly on the boundary
    X = Utrue @ V.T + 0.05 * rng.standard_normal((mt, n))             # noise makes some unconstrained optima negative
    Uo, erro = _scipy_cv(V, X)
    Ug, rnorm, xnorm = nnls_rows(V, X)
    assert (Ug >= 0).all()
    np.testing.assert_allclose(Ug, Uo, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(rnorm / xnorm, erro, rtol=1e-9)
    np.testing.assert_allclose(measure_cv_performance(V, X), erro, rtol=1e-9)
    assert ((Uo == 0) == (Ug == 0)).all(), "different active sets"


@pytest.mark.gpu
def test_batched_nnls_edge_cases():
    from prmf_b200 import nnls_rows
    rng = np.random.Generator(np.random.PCG64(9))
    V = rng.random((50, 4))
    X = np.vstack([-rng.random(50), V @ np.array([1.0, 0.0, 2.0, 0.0]), np.zeros(50)])   # all-negative, exact, zero
    U, rnorm, xnorm = nnls_rows(V, X)
    np.testing.assert_array_equal(U[0], 0.0)
    np.testing.assert_allclose(U[1], [1.0, 0.0, 2.0, 0.0], atol=1e-10)
    np.testing.assert_allclose(rnorm[1], 0.0, atol=1e-10)
    np.testing.assert_array_equal(U[2], 0.0)
    assert xnorm[2] == 0.0
    # strided rows (a column slice of a wider matrix) and an empty batch
    wide = rng.random((7, 80))
    U2, _, _ = nnls_rows(V, wide[:, :50])
    U3, _, _ = nnls_rows(V, np.ascontiguousarray(wide[:, :50]))
    np.testing.assert_array_equal(U2, U3)
    assert nnls_rows(V, np.zeros((0, 50)))[0].shape == (0, 4)
    with pytest.raises(ValueError):
        nnls_rows(V, rng.random((3, 49)))
    with pytest.raises(ValueError):
        nnls_rows(V, np.full((2, 50), np.nan))


@pytest.mark.gpu
def test_batched_nnls_dependent_columns():
    """V with duplicated / linearly dependent columns (V^T V singular): scipy's nnls copes; the batched solver must too
    (the minimiser is not unique there, the residual is)."""
    import scipy.optimize
    from prmf_b200 import nnls_rows
    rng = np.random.Generator(np.random.PCG64(21))
    base = rng.random((200, 5))
    V = np.hstack([base, base[:, :2], (base[:, 2] + base[:, 3])[:, None]])     # columns 5, 6 duplicate 0, 1; 7 = 2 + 3
    X = rng.random((30, 5)) @ base.T + 0.02 * rng.standard_normal((30, 200))
    U, rnorm, xnorm = nnls_rows(V, X)
    assert (U >= 0).all()
    for i in range(X.shape[0]):
        _, e = scipy.optimize.nnls(V, X[i])
        np.testing.assert_allclose(rnorm[i], e, rtol=1e-7)
        np.testing.assert_allclose(np.linalg.norm(X[i] - V @ U[i]), e, rtol=1e-7)
