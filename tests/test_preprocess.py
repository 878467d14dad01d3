"""GPU quantile_transform against the installed sklearn (the routine the reference CLI calls, :1019-1020)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sk(X, **kw):
    import warnings
    from sklearn.preprocessing import quantile_transform
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return quantile_transform(X, **kw)


@pytest.mark.parametrize("m,n,kind", [
    (100, 1000, "gamma"),        # the reference test instance's shape: n_quantiles clipped to m
    (2500, 301, "uniform"),      # m > 1000 quantiles
    (1200, 64, "ties"),          # heavy ties and repeated quantiles
    (1500, 40, "const"),         # constant and two-valued columns
    (37, 5, "gamma"),
])
def test_matches_sklearn(m, n, kind):
    from prmf_b200.preprocess import quantile_transform
    rng = np.random.Generator(np.random.PCG64(m + n))
    if kind == "gamma":
        X = rng.gamma(5.0, size=(m, n)) @ np.diag(rng.gamma(2.0, size=n))
    elif kind == "uniform":
        X = rng.random((m, n))
    elif kind == "ties":
        X = np.round(rng.gamma(2.0, size=(m, n)), 1)
        X[:, 0] = np.round(X[:, 0])
    else:
        X = rng.random((m, n))
        X[:, 0] = 3.25
        X[:, 1] = (rng.random(m) > 0.7).astype(float)
        X[:, 2] = 0.0
    ref = _sk(X.copy())
    got, Q = quantile_transform(X, return_quantiles=True)
    from sklearn.preprocessing import QuantileTransformer
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        qt = QuantileTransformer(subsample=int(1e5)).fit(X)
    np.testing.assert_allclose(Q, qt.quantiles_, rtol=1e-15, atol=0)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)
    assert np.max(np.abs(got - ref)) <= 4 * np.finfo(float).eps


def test_subsample_uses_the_same_rows_and_rng():
    from prmf_b200.preprocess import quantile_transform
    rng = np.random.Generator(np.random.PCG64(9))
    X = rng.gamma(3.0, size=(3000, 20))
    np.random.seed(5)
    ref = _sk(X.copy(), subsample=1000, n_quantiles=200)
    tail_ref = np.random.rand()
    np.random.seed(5)
    got = quantile_transform(X, subsample=1000, n_quantiles=200)
    tail_got = np.random.rand()
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)
    assert tail_ref == tail_got


def test_nan_is_rejected_and_device_output():
    import torch
    from prmf_b200.preprocess import quantile_transform
    X = np.random.Generator(np.random.PCG64(1)).random((50, 8))
    out = quantile_transform(X, return_device=True)
    assert out.is_cuda and out.shape == (50, 8)
    np.testing.assert_allclose(out.cpu().numpy(), _sk(X.copy()), rtol=1e-12, atol=1e-15)
    X[3, 2] = np.nan
    with pytest.raises(ValueError):
        quantile_transform(X)


def test_full_shape_speed_and_properties():
    """recount2 shape: monotone per gene (order preserved), range [0, 1], a column sample equal to sklearn."""
    import time
    from prmf_b200.preprocess import quantile_transform
    rng = np.random.Generator(np.random.PCG64(3))
    X = rng.gamma(2.0, size=(37032, 6750))
    t0 = time.perf_counter()
    out = quantile_transform(X)
    dt = time.perf_counter() - t0
    assert out.min() >= 0.0 and out.max() <= 1.0
    cols = [0, 17, 6749]
    np.testing.assert_allclose(out[:, cols], _sk(X[:, cols].copy()), rtol=1e-12, atol=1e-15)
    order = np.argsort(X[:, 5], kind="stable")
    assert np.all(np.diff(out[order, 5]) >= 0)
    print("GPU quantile_transform 37032x6750 incl. H2D/D2H: %.2f s" % dt)
    assert dt < 30
