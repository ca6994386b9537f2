"""
CPU tests that pin the oracle (oracle/oracle.py + oracle/oracle.c) BEFORE it is trusted as the
checker for the CUDA path.  They are ports of the reference's own property tests
(reference test/cg.jl, test/gmres.jl, test/minres.jl, test/bicgstabl.jl, test/orthogonalize.jl,
test/hessenberg.jl -- line numbers cited per test) plus analytic known answers and scipy
cross-checks.  numpy's default_rng replaces Julia's Random.seed!(1234321) stream (unavailable).
"""
import json
import math
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SEED = 1234321
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ generators
@pytest.mark.parametrize("n,dims", [(10, 1), (10, 2), (7, 3), (4, 4)])
def test_laplace_direct_equals_kron_recursion(oracle, n, dims):
    """direct C construction == literal kron recursion of reference test/laplace_matrix.jl:1-10."""
    A = oracle.laplace_matrix(np.float64, n, dims)
    K = oracle.laplace_matrix_scipy(np.float64, n, dims)
    assert A.nnz == K.nnz == (dims * 2 + 1) * n ** dims - 2 * dims * n ** (dims - 1)
    assert np.array_equal(A.colptr, K.indptr)
    assert np.array_equal(A.rowval, K.indices)
    assert np.array_equal(A.nzval, K.data)


def test_laplace_one_based_and_f32(oracle):
    A0 = oracle.laplace_matrix(np.float32, 6, 3, base=0)
    A1 = oracle.laplace_matrix(np.float32, 6, 3, base=1)
    assert A1.nzval.dtype == np.float32
    assert np.array_equal(A0.colptr + 1, A1.colptr) and np.array_equal(A0.rowval + 1, A1.rowval)
    x = np.random.default_rng(SEED).standard_normal(A0.n).astype(np.float32)
    assert np.array_equal(oracle.csc_spmv(A0, x), oracle.csc_spmv(A1, x))


def test_nnz_formula_of_survey(oracle):
    # 7N^3 - 6N^2 (3-D), 5N^2 - 4N (2-D)
    assert oracle.lib().oracle_laplace_nnz(24, 3) == 7 * 24 ** 3 - 6 * 24 ** 2 == 93312
    assert oracle.lib().oracle_laplace_nnz(128, 2) == 5 * 128 ** 2 - 4 * 128 == 81408


def test_advection_dominated_structure(oracle):
    """reference benchmark/advection_diffusion.jl:3-30."""
    N, beta = 6, 1000.0
    A, b = oracle.advection_dominated(N, beta)
    h = 1.0 / (N + 1)
    assert A.nnz == 7 * N ** 3 - 6 * N ** 2
    D = A.toarray()
    assert D[0, 0] == 6.0 / -(h ** 2)
    assert D[1, 0] == -1.0 / -(h ** 2) + (-beta / (2 * h))       # sub-diagonal on the fastest index
    assert D[0, 1] == -1.0 / -(h ** 2) + (beta / (2 * h))        # super-diagonal
    assert D[N, 0] == -1.0 / -(h ** 2) and D[N * N, 0] == -1.0 / -(h ** 2)
    x, y, z = 2 * h, 3 * h, 1 * h                                # b index = x + N*y + N^2*z (x fastest)
    want = math.exp(x * y * z) * math.sin(math.pi * x) * math.sin(math.pi * y) * math.sin(math.pi * z)
    assert b[1 + N * 2 + N * N * 0] == pytest.approx(want, rel=1e-14)


def test_csc_spmv_matches_scipy(oracle):
    rng = np.random.default_rng(SEED)
    M = sp.random(57, 43, density=0.2, random_state=7, format="csc", dtype=np.float64)
    A = oracle.CSC.from_scipy(M, base=1)
    x = rng.standard_normal(43)
    np.testing.assert_allclose(oracle.csc_spmv(A, x), M @ x, rtol=1e-14, atol=1e-15)
    X = np.asfortranarray(rng.standard_normal((43, 5)))
    np.testing.assert_allclose(oracle.csc_spmm(A, X), M @ X, rtol=1e-14, atol=1e-15)


# ------------------------------------------------------------------ Hessenberg (golden fixtures)
def _fixtures():
    with open(os.path.join(GOLDEN, "hessenberg_fixtures.json")) as f:
        g = json.load(f)
    H1 = np.array(g["H1"], dtype=np.float64)
    H2 = np.array(g["H2_re"], dtype=np.float64) + 1j * np.array(g["H2_im"], dtype=np.float64)
    return H1, H2


@pytest.mark.parametrize("which", [0, 1])
def test_hessenberg_fixtures(oracle, which):
    """reference test/hessenberg.jl:28-44 on its literal H1 (7x6 real) and H2 (5x4 complex)."""
    H = _fixtures()[which]
    rhs = np.zeros(H.shape[0], dtype=H.dtype)
    rhs[0] = 1
    sol_res = oracle.hessenberg_ldiv(H.copy(), rhs.copy())
    solution = np.linalg.lstsq(H, rhs, rcond=None)[0]
    np.testing.assert_allclose(sol_res[: H.shape[1]], solution, rtol=1e-10)
    assert abs(sol_res[-1]) == pytest.approx(np.linalg.norm(H @ solution - rhs), rel=1e-10)


def test_givens_convention(oracle):
    for f, g in [(3.0, 4.0), (-3.0, 4.0), (4.0, -3.0), (-4.0, 3.0), (0.0, 2.0), (2.0, 0.0), (1 + 2j, 3 - 1j)]:
        c, s, r = oracle.givens_algorithm(f, g)
        assert abs(c * f + s * g - r) < 1e-14 and abs(-np.conj(s) * f + c * g) < 1e-14
        assert abs(c * c + abs(s) ** 2 - 1) < 1e-14


# ------------------------------------------------------------------ orthogonalize (test/orthogonalize.jl:14-62)
@pytest.mark.parametrize("dtype", [np.complex64, np.float64])
@pytest.mark.parametrize("method", ["dgks", "cgs", "mgs"])
def test_orthogonalize(oracle, dtype, method):
    rng = np.random.default_rng(SEED)
    n, m = 10, 3
    R = rng.random((n, m)) + (1j * rng.random((n, m)) if np.iscomplexobj(np.zeros(1, dtype)) else 0)
    V = np.asfortranarray(np.linalg.qr(R.astype(dtype))[0])
    w_original = (rng.random(n) + (1j * rng.random(n) if np.iscomplexobj(np.zeros(1, dtype)) else 0)).astype(dtype)
    h = np.zeros(m, dtype=dtype)
    w = w_original.copy()
    nrm = oracle.orthogonalize_and_normalize_(V, w, h, method)
    eps = np.finfo(dtype).eps
    assert np.linalg.norm(w) == pytest.approx(1.0, rel=10 * eps ** 0.5)
    assert np.linalg.norm(V.conj().T @ w) <= 10 * eps
    np.testing.assert_allclose(nrm * w + V @ h, w_original, rtol=np.sqrt(eps))


# ------------------------------------------------------------------ CG (test/cg.jl)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cg_small_full_system(oracle, dtype):
    """test/cg.jl:24-53 (real types; complex is outside the hot path's configs)."""
    rng = np.random.default_rng(SEED)
    n = 10
    A = rng.random((n, n)).astype(dtype)
    A = A.T @ A + np.eye(n, dtype=dtype)
    b = rng.random(n).astype(dtype)
    reltol = math.sqrt(np.finfo(dtype).eps)
    x, ch = oracle.cg(A, b, reltol=reltol, maxiter=2 * n, log=True)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= reltol
    assert ch.isconverged
    x0 = np.linalg.solve(A.astype(np.float64), b.astype(np.float64)).astype(dtype)
    x, ch = oracle.cg_(x0, A, b, abstol=2 * n * np.finfo(dtype).eps, reltol=0.0, log=True)
    assert ch.niters <= 1 and ch.nprods <= 2
    x, ch = oracle.cg(A, b, Pl=oracle.MatrixPrec(A), log=True)        # exact factorization as Pl
    assert ch.niters <= 2 and ch.nprods <= 2
    x0 = oracle.cg(A, np.zeros(n, dtype=dtype))
    assert np.array_equal(x0, np.zeros(n, dtype=dtype))


def test_cg_sparse_laplacian(oracle):
    """test/cg.jl:55-87: 2-D Laplacian 10^2, CG vs Jacobi-PCG, LinearMap, starting guess."""
    rng = np.random.default_rng(SEED)
    A = oracle.laplace_matrix(np.float64, 10, 2, base=1)
    S = A.to_scipy()
    P = oracle.JacobiPrec(A.diagonal())
    rhs = rng.standard_normal(A.n)
    rhs *= 1.0 / np.linalg.norm(rhs)
    abstol = reltol = 1e-5
    linear_map = type("LM", (), {"shape": A.shape, "__call__": lambda self, v: S @ v})()
    for o in (A, linear_map):
        xCG = oracle.cg(o, rhs, reltol=reltol, maxiter=100)
        xJAC = oracle.cg(o, rhs, Pl=P, reltol=reltol, maxiter=100)
        assert np.linalg.norm(S @ xCG - rhs) <= reltol
        assert np.linalg.norm(S @ xJAC - rhs) <= reltol
    x0 = rng.standard_normal(A.n)
    xCG, hCG = oracle.cg_(x0.copy(), A, rhs, abstol=abstol, reltol=0.0, maxiter=100, log=True)
    xJAC, hJAC = oracle.cg_(x0.copy(), A, rhs, Pl=P, abstol=abstol, reltol=0.0, maxiter=100, log=True)
    assert np.linalg.norm(S @ xCG - rhs) <= reltol and np.linalg.norm(S @ xJAC - rhs) <= reltol
    assert hJAC.niters == hCG.niters                                  # constant diagonal => same iterates


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cg_termination(oracle, dtype):
    """test/cg.jl:98-122."""
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=dtype)
    n = 3
    b = np.ones(n, dtype=dtype)
    x0 = np.linalg.solve(A, b)
    pert = (10 * math.sqrt(np.finfo(dtype).eps) * np.array([(-1) ** i for i in range(1, n + 1)])).astype(dtype)
    x, ch = oracle.cg_(x0 + pert, A, b, log=True)
    assert 2 <= ch.niters <= n
    x = x0 + pert
    r0 = np.linalg.norm(A @ x - b)
    x, ch = oracle.cg_(x, A, b, abstol=2 * r0, reltol=0.0, log=True)
    assert ch.niters == 0


def test_cg_c_and_numpy_restatements_agree(oracle):
    """the all-C CG and the numpy CG are the same algorithm; config #1 shape (2-D Poisson)."""
    rng = np.random.default_rng(SEED)
    A = oracle.laplace_matrix(np.float64, 32, 2, base=1)
    b = rng.standard_normal(A.n)
    b /= np.linalg.norm(b)
    x1, h1 = oracle.cg_(np.zeros(A.n), A, b, initially_zero=True, log=True)
    x2, h2 = oracle.cg_csc_c(np.zeros(A.n), A, b, initially_zero=True)
    assert h1.niters == h2.niters and h1.mvps == h2.mvps and h1.isconverged and h2.isconverged
    np.testing.assert_allclose(h1["resnorm"], h2["resnorm"], rtol=1e-10)
    np.testing.assert_allclose(x1, x2, rtol=1e-10, atol=1e-14)
    d = A.diagonal()
    x3, h3 = oracle.cg_(np.zeros(A.n), A, b, initially_zero=True, log=True, Pl=oracle.JacobiPrec(d))
    x4, h4 = oracle.cg_csc_c(np.zeros(A.n), A, b, initially_zero=True, Pl_diag=d)
    assert h3.niters == h4.niters == h1.niters
    np.testing.assert_allclose(x3, x4, rtol=1e-10, atol=1e-14)


def test_cg_config1_iteration_count_and_scipy(oracle):
    """5-pt 2-D Poisson n=128^2, reltol=sqrt(eps).  Cross-check vs scipy."""
    rng = np.random.default_rng(SEED)
    A = oracle.laplace_matrix(np.float64, 128, 2, base=1)
    b = rng.standard_normal(A.n)
    b /= np.linalg.norm(b)
    x, h = oracle.cg_csc_c(np.zeros(A.n), A, b, initially_zero=True)
    assert h.isconverged and 300 < h.niters < 450
    S = A.to_scipy().tocsr()
    assert np.linalg.norm(S @ x - b) <= 2 * math.sqrt(np.finfo(np.float64).eps)
    xs, info = spla.cg(S, b, rtol=1e-12, atol=0.0, maxiter=5000)
    assert info == 0
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-5       # both are approximations of A\b


# ------------------------------------------------------------------ GMRES (test/gmres.jl)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gmres_dense(oracle, dtype):
    """test/gmres.jl:16-36."""
    rng = np.random.default_rng(SEED)
    n = 10
    A = (rng.random((n, n)) + np.eye(n)).astype(dtype)
    b = rng.random(n).astype(dtype)
    F = oracle.MatrixPrec(A)
    reltol = math.sqrt(np.finfo(dtype).eps)
    x, hist = oracle.gmres(A, b, log=True, restart=3, maxiter=10, reltol=reltol)
    assert np.all(np.diff(hist["resnorm"]) <= 0.0)
    x, hist = oracle.gmres(A, b, Pl=F, maxiter=1, restart=1, reltol=reltol, log=True)
    assert hist.isconverged
    assert np.linalg.norm(np.linalg.solve(A, A @ x - b)) / np.linalg.norm(b) <= reltol
    x, hist = oracle.gmres(A, b, Pl=oracle.Identity(), Pr=F, maxiter=1, restart=1, reltol=reltol, log=True)
    assert hist.isconverged
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= reltol


@pytest.mark.parametrize("orth", ["mgs", "cgs", "dgks"])
def test_gmres_sparse_and_orth_methods(oracle, orth):
    """test/gmres.jl:38-57 shape + all three orth_meth; null-vector residual == true residual."""
    rng = np.random.default_rng(SEED)
    n = 10
    M = sp.random(n, n, density=0.5, random_state=3, format="csc") + sp.identity(n, format="csc")
    A = oracle.CSC.from_scipy(M, base=1)
    b = rng.random(n)
    x, hist = oracle.gmres(A, b, log=True, restart=3, maxiter=10, orth_meth=orth)
    assert np.all(np.diff(hist["resnorm"]) <= 1e-15)
    x, hist = oracle.gmres(A, b, log=True, restart=n, maxiter=n, orth_meth=orth, reltol=1e-12)
    assert hist.isconverged
    assert np.linalg.norm(M @ x - b) / np.linalg.norm(b) < 1e-10


def test_gmres_linear_operator_cumsum(oracle):
    """test/gmres.jl:59-66: LinearMap(cumsum!, 100)."""
    rng = np.random.default_rng(SEED)
    op = type("LM", (), {"shape": (100, 100), "__call__": lambda self, v: np.cumsum(v)})()
    b = rng.random(100)
    x = oracle.gmres(op, b, reltol=1e-5, maxiter=2000)
    assert np.linalg.norm(np.cumsum(x) - b) / np.linalg.norm(b) <= 1e-5


def test_gmres_lucky_breakdown(oracle):
    """test/gmres.jl:68-73: identity matrix => x .== b exactly."""
    A = np.eye(2)
    b = np.array([1.0, 2.2])
    x = oracle.gmres(A, b)
    assert np.all(x == b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gmres_termination(oracle, dtype):
    """test/gmres.jl:75-99."""
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=dtype)
    n = 3
    b = np.ones(n, dtype=dtype)
    x0 = np.linalg.solve(A, b)
    pert = (10 * math.sqrt(np.finfo(dtype).eps) * np.array([(-1) ** i for i in range(1, n + 1)])).astype(dtype)
    x, ch = oracle.gmres_(x0 + pert, A, b, log=True)
    assert 2 <= ch.niters <= n
    x = x0 + pert
    r0 = np.linalg.norm(A @ x - b)
    x, ch = oracle.gmres_(x, A, b, abstol=2 * r0, reltol=0.0, log=True)
    assert ch.niters == 0


def test_gmres_advection_nullvec_residual_is_true_residual(oracle):
    """SURVEY.md section 8c probe: GMRES(30)+CGS on advection_dominated(N=12): implicit residual
    equals the true residual at the end, monotone history; cross-check the solution vs scipy spsolve."""
    M, b = oracle.advection_dominated(12, 1000.0)
    A = oracle.CSC.from_scipy(M, base=1)
    x, h = oracle.gmres(A, b, restart=30, orth_meth="cgs", log=True, maxiter=600, reltol=1e-10)
    assert h.isconverged
    assert np.all(np.diff(h["resnorm"]) <= 1e-12 * h["resnorm"][0])
    true = np.linalg.norm(b - M @ x)
    assert true == pytest.approx(h["resnorm"][-1], rel=1e-4)
    xs = spla.spsolve(M.tocsc(), b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-7


# ------------------------------------------------------------------ MINRES (test/minres.jl)
def test_minres_hermitian_and_inplace(oracle):
    rng = np.random.default_rng(SEED)
    n = 15
    B = rng.random((n, n)) + n * np.eye(n)                            # test/minres.jl:12-18
    A = B + B.T
    b = B @ np.ones(n)
    x, hist = oracle.minres(A, b, maxiter=10 * n, reltol=1e-10, log=True)
    assert hist.isconverged
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) < 1e-9
    x0 = rng.standard_normal(n)
    x2, hist = oracle.minres_(x0, A, b, maxiter=10 * n, reltol=1e-10, log=True)
    assert x2 is x0 and hist.isconverged                              # test/minres.jl:44
    assert np.linalg.norm(b - A @ x2) / np.linalg.norm(b) < 1e-9


def test_minres_skew_hermitian(oracle):
    rng = np.random.default_rng(SEED)
    n = 15
    B = rng.random((n, n)) + n * np.eye(n)                            # test/minres.jl:20-26
    A = B - B.T
    b = A @ np.ones(n)
    reltol = math.sqrt(np.finfo(np.float64).eps)
    x, hist = oracle.minres(A, b, skew_hermitian=True, maxiter=10 * n, reltol=reltol, log=True)
    assert hist.isconverged
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= reltol


def test_minres_sparse_laplacian_and_scipy(oracle):
    rng = np.random.default_rng(SEED)
    A = oracle.laplace_matrix(np.float64, 12, 2, base=1)
    S = A.to_scipy()
    b = rng.standard_normal(A.n)
    x, hist = oracle.minres(A, b, reltol=1e-10, log=True)
    assert hist.isconverged and np.linalg.norm(S @ x - b) / np.linalg.norm(b) < 1e-9
    # implicit residual == true residual (MINRES property)
    assert hist["resnorm"][-1] == pytest.approx(np.linalg.norm(S @ x - b), rel=1e-5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_minres_termination(oracle, dtype):
    """test/minres.jl:72-96."""
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=dtype)
    n = 3
    b = np.ones(n, dtype=dtype)
    x0 = np.linalg.solve(A, b)
    pert = (10 * math.sqrt(np.finfo(dtype).eps) * np.array([(-1) ** i for i in range(1, n + 1)])).astype(dtype)
    x, ch = oracle.minres_(x0 + pert, A, b, log=True)
    assert 2 <= ch.niters <= n
    x = x0 + pert
    r0 = np.linalg.norm(A @ x - b)
    x, ch = oracle.minres_(x, A, b, abstol=2 * r0, reltol=0.0, log=True)
    assert ch.niters == 0


# ------------------------------------------------------------------ BiCGStab(l) (test/bicgstabl.jl)
@pytest.mark.parametrize("l", [2, 4])
def test_bicgstabl(oracle, l):
    """test/bicgstabl.jl:13-44."""
    rng = np.random.default_rng(SEED)
    n = 20
    A = rng.random((n, n)) + 15 * np.eye(n)
    b = rng.random(n)
    reltol = math.sqrt(np.finfo(np.float64).eps)
    x1, h1 = oracle.bicgstabl(A, b, l, max_mv_products=100, log=True, reltol=reltol, rng=np.random.default_rng(1))
    assert h1.isconverged and np.linalg.norm(A @ x1 - b) / np.linalg.norm(b) <= reltol
    x0 = np.zeros(n)
    x2, h2 = oracle.bicgstabl_(x0, A, b, l, max_mv_products=100, log=True, reltol=reltol, rng=np.random.default_rng(1))
    assert x2 is x0 and np.allclose(x2, x1)
    Pl = oracle.MatrixPrec(A + 1e-2 * rng.random((n, n)))             # "LU of nearby matrix" :36-42
    x3, h3 = oracle.bicgstabl(A, b, l, Pl=Pl, max_mv_products=100, log=True, reltol=reltol)
    assert h3.isconverged and np.linalg.norm(A @ x3 - b) / np.linalg.norm(b) <= 10 * reltol


def test_bicgstabl_termination(oracle):
    """test/bicgstabl.jl:46-69 (1 <= iters <= n/2 with l=2... on the 3x3 tridiagonal)."""
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=np.float64)
    b = np.ones(3)
    x0 = np.linalg.solve(A, b)
    pert = 10 * math.sqrt(np.finfo(np.float64).eps) * np.array([-1.0, 1.0, -1.0])
    x, ch = oracle.bicgstabl_(x0 + pert, A, b, 1, log=True, max_mv_products=20)
    assert 1 <= ch.niters <= 3
    x = x0 + pert
    r0 = np.linalg.norm(A @ x - b)
    x, ch = oracle.bicgstabl_(x, A, b, 1, abstol=2 * r0, reltol=0.0, log=True)
    assert ch.niters == 0


def test_threaded_bench_variant_equals_serial_steps(oracle):
    """bench.py's `threaded_variant_not_reference_behaviour` (oracle_cg_steps_f64_omp) performs the same CG
    steps as the serial restatement; only the summation order of the reductions differs."""
    import ctypes as C
    A = oracle.laplace_matrix(np.float64, 14, 3, base=1)
    n = A.n
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n)
    out = []
    for fn in ("oracle_cg_steps_f64", "oracle_cg_steps_f64_omp"):
        x, u, c = np.zeros(n), np.zeros(n), np.zeros(n)
        r = b.copy()
        res, prev = C.c_double(float(np.linalg.norm(r))), C.c_double(1.0)
        getattr(oracle.lib(), fn)(C.c_int64(n), oracle._p(A.colptr), oracle._p(A.rowval), oracle._p(A.nzval),
                                  C.c_int64(1), oracle._p(x), oracle._p(r), oracle._p(u), oracle._p(c),
                                  C.byref(res), C.byref(prev), C.c_int64(25))
        out.append((x, res.value))
    assert abs(out[0][1] - out[1][1]) <= 1e-12 * out[0][1]
    assert np.linalg.norm(out[0][0] - out[1][0]) <= 1e-12 * np.linalg.norm(out[0][0])
