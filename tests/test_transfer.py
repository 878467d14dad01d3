"""Ridge transfer of a fitted V to new samples (reference script/transfer_learning.R:98-110) against numpy."""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k", [(50, 300, 6), (700, 1031, 10), (33, 2100, 24)])
def test_ridge_transfer_matches_closed_form(m, n, k):
    from prmf_b200.transfer import default_l2, ridge_transfer
    rng = np.random.Generator(np.random.PCG64(m + n + k))
    X = rng.random((m, n))
    Z = rng.gamma(2.0, size=(n, k))
    B = ridge_transfer(X, Z)
    l2 = default_l2(n, m, k)
    want = np.linalg.solve(Z.T @ Z + l2 * np.eye(k), Z.T @ X.T).T          # t(solve(Z'Z + L2 I) Z' Y), :108-110
    np.testing.assert_allclose(B, want, rtol=1e-11, atol=1e-14)
    B2 = ridge_transfer(X, Z, l2=3.5)
    np.testing.assert_allclose(B2, np.linalg.solve(Z.T @ Z + 3.5 * np.eye(k), Z.T @ X.T).T, rtol=1e-11, atol=1e-14)


def test_default_l2_is_the_reference_ratio():
    from prmf_b200.transfer import default_l2
    assert default_l2(6750, 37032, 10) == 6750 / 10 * 100                  # (m*n)/(k*n) * 100, :93-96
