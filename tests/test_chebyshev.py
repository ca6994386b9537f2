"""Chebyshev iteration (reference src/chebyshev.jl): oracle pinned by the
reference's tests (test/chebyshev.jl:29-62, termination :64-89), and GPU parity against the oracle."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

SEED = 1234321


def bounds(D):
    lam = np.linalg.eigvalsh(D)
    d = (lam[-1] - lam[0]) / 100                                      # approx_eigenvalue_bounds  test/chebyshev.jl:13-18
    return lam[0] - d, lam[-1] + d


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_chebyshev_reference_properties(oracle, dtype):
    rng = np.random.default_rng(SEED)
    n = 10
    M = rng.random((n, n)).astype(dtype) + n * np.eye(n, dtype=dtype)
    A = M.T @ M                                                       # randSPD  :8-11
    b = rng.random(n).astype(dtype)
    reltol = math.sqrt(np.finfo(dtype).eps)
    lo, hi = bounds(A.astype(np.float64))
    x, h = oracle.chebyshev(A, b, lo, hi, reltol=reltol, maxiter=10 * n, log=True)
    assert h.isconverged and np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= reltol
    x0 = rng.random(n).astype(dtype)
    r0 = np.linalg.norm(A @ x0 - b)
    x, h = oracle.chebyshev_(x0, A, b, lo, hi, reltol=reltol, maxiter=10 * n, log=True)
    assert h.isconverged and x is x0 and np.linalg.norm(A @ x - b) <= reltol * r0 * 1.01
    Pl = oracle.JacobiPrec(np.diag(A).copy())
    A64 = A.astype(np.float64)
    Dh = np.diag(1.0 / np.sqrt(np.diag(A64)))                         # D^-1 A is similar to D^-1/2 A D^-1/2
    lo2, hi2 = bounds(Dh @ A64 @ Dh)
    x, h = oracle.chebyshev(A, b, lo2, hi2, Pl=Pl, reltol=reltol, maxiter=10 * n, log=True)
    assert h.isconverged and np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 10 * reltol


def test_oracle_chebyshev_termination(oracle):
    A = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    b = np.ones(3)
    x0 = np.linalg.solve(A, b)
    pert = 10 * math.sqrt(np.finfo(np.float64).eps) * np.array([-1.0, 1.0, -1.0])
    lo, hi = bounds(A)
    x = x0 + pert
    r0 = np.linalg.norm(A @ x - b)
    x, ch = oracle.chebyshev_(x, A, b, lo, hi, abstol=2 * r0, reltol=0.0, log=True)
    assert ch.niters == 0


@pytest.mark.gpu
def test_gpu_chebyshev_vs_oracle(oracle):
    import iterativesolvers_jl_b200 as isb
    rng = np.random.default_rng(SEED)
    N, shift = 16, 6.0
    # The reference's recurrence (u .= c .+ beta .* c, src/chebyshev.jl:45) only converges for small condition
    # numbers (its own tests use randSPD with kappa ~ 1.3), so the operator is the 3-D Laplacian + 6 I (kappa ~ 3).
    O = oracle.laplace_matrix(np.float64, N, 3, base=1)
    cols = np.repeat(np.arange(O.n), np.diff(O.colptr))
    O.nzval[(O.rowval - 1) == cols] += shift
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    lo, hi = (3 * lam1[0] + shift) * 0.99, (3 * lam1[-1] + shift) * 1.01   # analytic spectrum
    b = rng.standard_normal(O.n)
    for kw_d, kw_o in [({}, {}), ({"Pl": isb.JacobiPrec(A.diag())}, {"Pl": oracle.JacobiPrec(O.diagonal())})]:
        sc = 1.0 if not kw_d else 1.0 / (6.0 + shift)                 # Jacobi scales the spectrum by 1/diag
        xo, ho = oracle.chebyshev(O, b, lo * sc, hi * sc, log=True, maxiter=400, reltol=1e-8, **kw_o)
        x, h = isb.chebyshev(A, b, lo * sc, hi * sc, log=True, maxiter=400, reltol=1e-8, **kw_d)
        assert h.isconverged and h.niters == ho.niters and h.mvps == ho.mvps
        assert np.max(np.abs(h["resnorm"] - ho["resnorm"]) / ho["resnorm"]) <= 1e-10
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10
    # initial guess, in place
    x0 = rng.standard_normal(O.n)
    xo, ho = oracle.chebyshev_(x0.copy(), O, b, lo, hi, log=True, maxiter=20, reltol=0.0)
    x1, h1 = isb.chebyshev_(x0, A, b, lo, hi, log=True, maxiter=20, reltol=0.0)
    assert x1 is x0 and h1.niters == ho.niters == 20 and h1.mvps == ho.mvps == 21 and not h1.isconverged
    np.testing.assert_allclose(h1["resnorm"], ho["resnorm"], rtol=1e-10)
    np.testing.assert_allclose(x1, xo, rtol=0, atol=1e-10 * np.linalg.norm(xo))
    # ill-conditioned operator: the reference recurrence diverges; the device path must diverge the same way
    L = oracle.laplace_matrix(np.float64, N, 3, base=1)
    AL = isb.B200CSR.from_csc_arrays(L.colptr, L.rowval, L.nzval, L.shape, base=1)
    _, ho = oracle.chebyshev(L, b, 3 * lam1[0], 3 * lam1[-1], log=True, maxiter=12, reltol=1e-8)
    _, h2 = isb.chebyshev(AL, b, 3 * lam1[0], 3 * lam1[-1], log=True, maxiter=12, reltol=1e-8)
    assert not h2.isconverged and h2.niters == 12 and ho["resnorm"][-1] > ho["resnorm"][0]
    np.testing.assert_allclose(h2["resnorm"], ho["resnorm"], rtol=1e-9)
