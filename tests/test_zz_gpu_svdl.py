"""GPU tests for svdl (reference src/svdl.jl, method = :ritz) through the C ABI: the engine-vs-oracle case shared with
the serial backend (tests/widening_cases.py), a sparse rectangular operator against scipy's svds, and a callback
operator.  (Written after the round's GPU budget was spent: first executed by the round-end GPU run.)"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import widening_cases as cases

pytestmark = pytest.mark.gpu
SEED = 1234321


@pytest.fixture(scope="module")
def isb():
    import iterativesolvers_jl_b200 as m
    m.default_context()
    return m


def run_gpu(isb):
    def run(A, v0, **kw):
        op = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(v0.dtype))
        X, L, h = isb.svdl(op, v0=v0, vecs="both", log=True, **kw)
        return dict(sigma=np.asarray(X.S, dtype=np.float64), U=X.U, V=X.Vt.T, iters=h.iters, mvps=h.mvps, mtvps=h.mtvps,
                    converged=h.isconverged, ritz=h["ritz"], resnorm=h["resnorm"], conv=h["conv"], betas=h["betas"])
    return run


@pytest.mark.parametrize("method", ["ritz", "harmonic"])
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 2e-5)])
def test_svdl_matches_oracle(isb, oracle, dtype, tol, method):
    cases.case_svdl_matches_oracle(oracle, run_gpu(isb), dtype, tol, method)


def test_svdl_sparse_rectangular_against_scipy_and_callback_operator(isb):
    rng = np.random.default_rng(SEED)
    m, n, nsv = 4000, 1500, 6
    M = sp.random(m, n, density=0.004, random_state=9, format="csc")
    A = isb.B200CSR.from_scipy(M)
    v0 = rng.standard_normal(n)
    v0 /= np.linalg.norm(v0)
    X, L, h = isb.svdl(A, nsv=nsv, k=24, v0=v0, tol=1e-8, reltol=1e-10, maxiter=200, vecs="both", log=True)
    ref = np.sort(spl.svds(M, k=nsv, tol=1e-12, return_singular_vectors=False))[::-1]
    assert h.isconverged and np.abs(X.S - ref).max() <= 1e-7 * ref[0]
    assert np.linalg.norm(M @ X.Vt.T - X.U * X.S[None, :]) <= 1e-5 * ref[0]
    assert np.abs(X.U.T @ X.U - np.eye(nsv)).max() <= 1e-6 and np.abs(X.Vt @ X.Vt.T - np.eye(nsv)).max() <= 1e-6
    assert h.mtvps > h.mvps > 0 and h["ritz"].shape == (h.iters, 24) and L.B.shape == (24, 24)
    # the same operator through callbacks
    op = isb.B200LinearOperator((m, n), np.float64, lambda y, x: A.mul_(y, x), lambda y, x: A.adjoint().mul_(y, x))
    s2, L2 = isb.svdl(op, nsv=nsv, k=24, v0=v0, tol=1e-8, reltol=1e-10, maxiter=200)
    assert np.abs(s2 - X.S).max() <= 1e-10 * ref[0]
