"""CPU test: the library's host-side dense eigen-solvers (used by LOBPCG for the Rayleigh-Ritz step,
reference src/lobpcg.jl:615,622 -> LAPACK syevd/sygvd) pinned against LAPACK (numpy/scipy)."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla


@pytest.fixture(scope="module")
def L():
    import iterativesolvers_jl_b200 as isb
    return isb.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 32, 48])
def test_sym_eig_and_generalized(L, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = np.asfortranarray(M + M.T)
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    assert L.b200_dense_sygv_host(n, _p(A), None, _p(w), _p(Z)) == 0
    wr = np.linalg.eigvalsh(A)
    np.testing.assert_allclose(w, wr, rtol=1e-11, atol=1e-12 * max(1.0, np.abs(wr).max()))
    assert np.linalg.norm(A @ Z - Z * w[None, :]) <= 1e-11 * max(1.0, np.linalg.norm(A))
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= 1e-12 * n
    # generalized: B SPD
    Q = rng.standard_normal((n, n))
    B = np.asfortranarray(Q @ Q.T + n * np.eye(n))
    assert L.b200_dense_sygv_host(n, _p(A), _p(B), _p(w), _p(Z)) == 0
    wr = sla.eigh(A, B, eigvals_only=True)
    np.testing.assert_allclose(w, wr, rtol=1e-10, atol=1e-12)
    assert np.linalg.norm(A @ Z - (B @ Z) * w[None, :]) <= 1e-10 * np.linalg.norm(A)
    assert np.linalg.norm(Z.T @ B @ Z - np.eye(n)) <= 1e-11 * n          # LAPACK sygvd normalisation


def test_clustered_and_rayleigh_ritz_like(L):
    """the shape LOBPCG produces: gramA = [diag(lambda) small; small ...], gramB = I + small."""
    rng = np.random.default_rng(5)
    n = 48
    E = 1e-3 * rng.standard_normal((n, n))
    A = np.asfortranarray(np.diag(np.repeat(np.linspace(0.1, 3.0, n // 3), 3)) + E + E.T)   # triple clusters
    F = 1e-4 * rng.standard_normal((n, n))
    B = np.asfortranarray(np.eye(n) + F + F.T)
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    assert L.b200_dense_sygv_host(n, _p(A), _p(B), _p(w), _p(Z)) == 0
    np.testing.assert_allclose(w, sla.eigh(A, B, eigvals_only=True), rtol=1e-10, atol=1e-13)
    assert np.all(np.diff(w) >= 0)


def test_not_positive_definite_is_reported(L):
    A = np.asfortranarray(np.eye(3))
    B = np.asfortranarray(np.diag([1.0, -1.0, 1.0]))
    w = np.zeros(3)
    Z = np.zeros((3, 3), order="F")
    assert L.b200_dense_sygv_host(3, _p(A), _p(B), _p(w), _p(Z)) == -5      # B200_ERR_BREAKDOWN
