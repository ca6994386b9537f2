"""GPU tests for matrix-free operators and callback preconditioners (b200_linop, csrc/cg_core.h, the *_op entry
points): the reference's duck-typed `mul!` / `ldiv!` contract (docs/src/getting_started.md:25-30,
docs/src/preconditioning.md:5-15; LinearMaps in test/cg.jl:71-77, test/lsqr.jl:36) through the C ABI.

The engine-vs-oracle cases are shared with the CPU run on the serial backend (tests/widening_cases.py).
(Written after the round's GPU budget was spent: first executed by the round-end GPU run.)
"""
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp

import widening_cases as cases

pytestmark = pytest.mark.gpu
SEED = 1234321


@pytest.fixture(scope="module")
def isb():
    import iterativesolvers_jl_b200 as m
    m.default_context()
    return m


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


class GpuCgRunner:
    """general-operator cg!: the operator always goes through the callback interface."""

    def __init__(self, isb):
        self.isb = isb

    def cg(self, x, A, b, mode, d, **kw):
        isb = self.isb
        csr = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        if mode == "callback":
            jac = isb.JacobiPrec(np.asarray(d, dtype=x.dtype))
            op, Pl = csr, isb.FunctionPrec(csr.m_local, x.dtype, lambda y, v: jac.ldiv_(y, v))   # CSR A + callback Pl
        else:
            op = isb.B200LinearOperator.from_csr(csr)                                              # callback A
            Pl = isb.JacobiPrec(np.asarray(d, dtype=x.dtype)) if mode == "jacobi" else None
        x, h = isb.cg_(x, op, np.asarray(b, dtype=x.dtype), Pl=Pl, log=True, **kw)
        res = isb.cg_.last_result
        return x, SimpleNamespace(iters=h.iters, converged=h.isconverged, mvps=h.mvps, hist=h["resnorm"],
                                  breakdown=res.status != 0)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 2e-5)])
def test_cg_general_operator_matches_oracle(isb, oracle, dtype, tol):
    cases.case_cg_general_operator(oracle, [GpuCgRunner(isb)], dtype, tol)


def test_general_cg_equals_the_specialised_engine(isb, oracle):
    """same operator, same right-hand side: the pass-based engine behind a callback and the fused CSR engine of
    cg.cu follow the same recurrence (history 1e-10, x 1e-10)."""
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float64, 24, 3, base=1)
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    b = rng.standard_normal(O.n)
    b /= np.linalg.norm(b)
    x1, h1 = isb.cg(A, b, log=True)
    x2, h2 = isb.cg(isb.B200LinearOperator.from_csr(A), b, log=True)
    assert h1.iters == h2.iters and h1.mvps == h2.mvps and h1.isconverged and h2.isconverged
    assert np.max(np.abs(h1["resnorm"] - h2["resnorm"]) / h1["resnorm"]) <= 1e-10 and relerr(x2, x1) <= 1e-10
    Pj = isb.JacobiPrec(A.diag(), A.ctx)
    x3, h3 = isb.cg(A, b, Pl=Pj, log=True)
    x4, h4 = isb.cg(A, b, Pl=isb.FunctionPrec(A.m_local, np.float64, lambda y, v: Pj.ldiv_(y, v)), log=True)
    assert h3.iters == h4.iters and relerr(x4, x3) <= 1e-10


def test_shifted_operator_by_callback_all_solvers(isb, oracle):
    """A matrix-free operator built from library calls inside the callback: B = A + sigma I (mul: y = A x, then
    y += sigma x), solved by cg / qmr / idrs / lsqr / lsmr and compared with the assembled matrix."""
    rng = np.random.default_rng(SEED)
    n, sigma = 2000, 0.75
    M = sp.random(n, n, 0.004, random_state=5, format="csc")
    M = (M + M.T + 8 * sp.eye(n)).tocsc()                     # symmetric positive definite, so cg applies too
    A = isb.B200CSR.from_scipy(M)
    L = isb.lib()
    calls = {"mul": 0}

    def mul(y, x):
        calls["mul"] += 1
        A.mul_(y, x)
        assert L.b200_axpby(A.ctx._h, n, sigma, x._p, 1.0, y._p, 0) == 0

    B = isb.B200LinearOperator((n, n), np.float64, mul, adjoint_mul=mul)     # symmetric: A' = A
    Bs = isb.B200CSR.from_scipy((M + sigma * sp.eye(n)).tocsc())
    b = rng.standard_normal(n)
    for name, kw in (("cg", {}), ("qmr", {}), ("idrs", dict(s=4, rng=np.random.default_rng(1))),
                     ("lsqr", dict(atol=1e-12, btol=1e-12)), ("lsmr", dict(atol=1e-12, btol=1e-12))):
        fn = getattr(isb, name)
        before = calls["mul"]
        kw2 = dict(kw)
        if name == "idrs":
            kw2["rng"] = np.random.default_rng(1)
        x_cb, h_cb = fn(B, b, log=True, **kw)
        x_as, h_as = fn(Bs, b, log=True, **kw2)
        assert calls["mul"] > before
        assert h_cb.isconverged and h_as.isconverged and abs(h_cb.iters - h_as.iters) <= 1, name
        assert relerr(x_cb, x_as) <= (1e-8 if h_cb.iters == h_as.iters else 1e-5), name
        assert relerr((M + sigma * sp.eye(n)) @ x_cb, b) <= 1e-6, name


def test_callback_errors_surface_as_python_exceptions(isb):
    n = 64
    A = isb.B200CSR.from_scipy((sp.eye(n) * 2.0).tocsc())

    def bad(y, x):
        raise RuntimeError("operator failed")

    op = isb.B200LinearOperator((n, n), np.float64, bad)
    with pytest.raises(RuntimeError, match="operator failed"):
        isb.cg(op, np.ones(n))
    with pytest.raises(TypeError):
        isb.qmr(isb.B200LinearOperator((n, n), np.float64, lambda y, x: A.mul_(y, x)), np.ones(n))   # no adjoint
    # the library still works afterwards
    x = isb.cg(A, np.ones(n))
    assert relerr(x, 0.5 * np.ones(n)) <= 1e-12


def test_idrs_callback_preconditioner(isb, oracle):
    from test_zy_gpu_widening import GpuRunner
    cases.case_idrs_callback_preconditioner(oracle, [GpuRunner(isb)])
    # an unknown preconditioner kind is refused, not ignored
    A = isb.B200CSR.from_scipy((sp.eye(32) * 2.0).tocsc())
    L = isb.lib()
    import ctypes as C
    from iterativesolvers_jl_b200 import _lib
    opts = _lib.CgOpts(0.0, 1e-8, 10, 1, 0, _lib.Precond(9, 0, None), 0, 0)
    x, b = isb.DeviceArray.zeros(A.ctx, 32), isb.DeviceArray.from_numpy(A.ctx, np.ones(32))
    assert L.b200_cg_solve(A.ctx._h, A._h, x._p, b._p, C.byref(opts), None, None, 0) != 0
    assert b"preconditioner" in L.b200_last_error()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 5e-5)])
def test_gmres_general_operator_and_callback_preconditioners(isb, oracle, dtype, tol):
    """gmres! with `mul!` / `ldiv!` callbacks (b200_gmres_solve_op; b200_gmres_solve forwarding a callback Pl / Pr to the
    same engine) against the oracle -- the case the serial backend runs (tests/widening_cases.py)."""
    def run(x, A, b, d, pl, pr, restart, maxiter, meth, **kw):
        csr = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        jac = isb.JacobiPrec(np.asarray(d, dtype=x.dtype))
        mk = lambda kind: None if kind is None else (jac if kind == "jac" else
                                                     isb.FunctionPrec(csr.m_local, x.dtype, lambda y, v: jac.ldiv_(y, v)))
        op = csr if "cb" in (pl, pr) else isb.B200LinearOperator.from_csr(csr)
        x, h = isb.gmres_(x, op, b, Pl=mk(pl), Pr=mk(pr), restart=restart, maxiter=maxiter, orth_meth=meth, log=True, **kw)
        return x, SimpleNamespace(iters=h.iters, mvps=h.mvps, converged=h.isconverged, hist=h["resnorm"])
    cases.case_gmres_general(oracle, run, dtype, tol)


def test_general_gmres_equals_the_specialised_engine(isb, oracle):
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float64, 16, 3, base=1)
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    b = rng.standard_normal(O.n)
    for meth in ("mgs", "cgs", "dgks"):
        x1, h1 = isb.gmres(A, b, restart=30, maxiter=90, orth_meth=meth, log=True)
        x2, h2 = isb.gmres(isb.B200LinearOperator.from_csr(A), b, restart=30, maxiter=90, orth_meth=meth, log=True)
        assert h1.iters == h2.iters and h1.mvps == h2.mvps and h1.isconverged == h2.isconverged
        assert np.max(np.abs(h1["resnorm"] - h2["resnorm"])) <= 1e-9 * h1["resnorm"][0] and relerr(x2, x1) <= 1e-8


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 5e-5)])
def test_minres_general_operator(isb, oracle, dtype, tol):
    """minres! with a `mul!` callback (b200_minres_solve_op) against the oracle -- the case of the serial backend."""
    def run(x, A, b, **kw):
        csr = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        x, h = isb.minres_(x, isb.B200LinearOperator.from_csr(csr), b, log=True, **kw)
        return x, SimpleNamespace(iters=h.iters, mvps=h.mvps, converged=h.isconverged, hist=h["resnorm"])
    cases.case_minres_general(oracle, run, dtype, tol)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 2e-4)])
def test_bicgstabl_general_operator_and_callback_preconditioner(isb, oracle, dtype, tol):
    """bicgstabl! with `mul!` / `ldiv!` callbacks (b200_bicgstabl_solve_op; b200_bicgstabl_solve forwarding a callback Pl)
    against the oracle with the same shadow residual -- the case of the serial backend."""
    def run(x, A, b, l, shadow, d, pk, **kw):
        csr = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        jac = isb.JacobiPrec(np.asarray(d, dtype=x.dtype))
        Pl = None if pk is None else (jac if pk == "jac" else
                                      isb.FunctionPrec(csr.m_local, x.dtype, lambda y, v: jac.ldiv_(y, v)))
        op = csr if pk == "cb" else isb.B200LinearOperator.from_csr(csr)
        try:
            x, h = isb.bicgstabl_(x, op, b, l, Pl=Pl, r_shadow=shadow, log=True, **kw)
        except np.linalg.LinAlgError:
            return x, SimpleNamespace(singular=True)
        return x, SimpleNamespace(iters=h.iters, mvps=h.mvps, converged=h.isconverged, hist=h["resnorm"], singular=False)
    cases.case_bicgstabl_general(oracle, run, dtype, tol)


def test_general_minres_and_bicgstabl_equal_the_specialised_engines(isb, oracle):
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float64, 16, 3, base=1)
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    op = isb.B200LinearOperator.from_csr(A)
    b = rng.standard_normal(O.n)
    x1, h1 = isb.minres(A, b, log=True)
    x2, h2 = isb.minres(op, b, log=True)
    assert h1.iters == h2.iters and h1.isconverged and h2.isconverged and relerr(x2, x1) <= 1e-8
    k = min(30, h1.iters)
    assert np.max(np.abs(h1["resnorm"][:k] - h2["resnorm"][:k])) <= 1e-9 * h1["resnorm"][0]
    sh = rng.random(O.n)
    x1, h1 = isb.bicgstabl(A, b, 2, r_shadow=sh, log=True)
    x2, h2 = isb.bicgstabl(op, b, 2, r_shadow=sh, log=True)
    assert abs(h1.iters - h2.iters) <= 1 and h1.isconverged and h2.isconverged and relerr(x2, x1) <= 1e-6
    k = min(10, h1.iters, h2.iters)
    assert np.max(np.abs(h1["resnorm"][:k] - h2["resnorm"][:k])) <= 1e-6 * h1["resnorm"][0]


def test_nested_solve_inside_a_callback_is_refused(isb):
    """the context's scratch belongs to the running solve: starting another solve on the same context from inside an
    operator callback fails loudly instead of overwriting it."""
    n = 64
    A = isb.B200CSR.from_scipy((sp.eye(n) * 2.0).tocsc())

    def inner(y, x):
        isb.cg_(y, A, x)

    with pytest.raises(isb.B200Error, match="inside an operator"):
        isb.cg(isb.B200LinearOperator((n, n), np.float64, inner), np.ones(n))
    assert relerr(isb.cg(A, np.ones(n)), 0.5 * np.ones(n)) <= 1e-12


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 5e-6)])
def test_chebyshev_general_operator_and_callback_preconditioner(isb, oracle, dtype, tol):
    """chebyshev! with `mul!` / `ldiv!` callbacks (b200_chebyshev_solve_op; b200_chebyshev_solve forwarding a callback Pl)
    against the oracle -- the case of the serial backend."""
    def run(x, A, b, lmin, lmax, d, pk, **kw):
        csr = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        jac = isb.JacobiPrec(np.asarray(d, dtype=x.dtype))
        Pl = None if pk is None else (jac if pk == "jac" else
                                      isb.FunctionPrec(csr.m_local, x.dtype, lambda y, v: jac.ldiv_(y, v)))
        op = csr if pk == "cb" else isb.B200LinearOperator.from_csr(csr)
        x, h = isb.chebyshev_(x, op, b, lmin, lmax, Pl=Pl, log=True, **kw)
        return x, SimpleNamespace(iters=h.iters, mvps=h.mvps, converged=h.isconverged, hist=h["resnorm"])
    cases.case_chebyshev_general(oracle, run, dtype, tol)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_powm_and_invpowm(isb, oracle, dtype):
    """powm! / invpowm! (reference src/simple.jl, test/simple_eigensolvers.jl:14-50) through b200_powm on a CSR operator --
    the case of the serial backend (inverse iteration with the explicit inverse standing for the reference's LU LinearMap) --
    and the dominant eigenvalue of a 2-D Laplacian."""
    def run(A, x0, tol, maxiter):
        lam, x, h = isb.powm_(isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x0.dtype)), x0, tol=tol, maxiter=maxiter, log=True)
        return float(lam), x, SimpleNamespace(iters=h.iters, converged=h.isconverged, hist=h["resnorm"])
    cases.case_powm(oracle, run, dtype)
    # a larger sparse operator: dominant eigenvalue of the 2-D Laplacian = 8 - (smallest eigenvalue)
    N = 24
    O = oracle.laplace_matrix(np.float64, N, 2, base=1)
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    lam, x, h = isb.powm(A, tol=1e-3, maxiter=20000, log=True, rng=np.random.default_rng(SEED))
    exact = 4 + 4 * np.cos(np.pi / (N + 1))
    assert h.isconverged and abs(lam - exact) <= 1e-3 and abs(np.linalg.norm(x) - 1) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_stationary_methods(isb, oracle, dtype):
    """jacobi! / gauss_seidel! / sor! / ssor! (reference src/stationary_sparse.jl) through b200_stationary: the level-scheduled
    sweeps against the oracle's sequential column sweeps and the reference's own tests -- the case of the serial backend."""
    def run(name, x, A, b, w, mi):
        op = isb.B200CSR.from_scipy(sp.csc_matrix(A).astype(x.dtype))
        fn = getattr(isb, name + "_")
        return fn(x, op, b, w, maxiter=mi) if name in ("sor", "ssor") else fn(x, op, b, maxiter=mi)
    cases.case_stationary(oracle, run, dtype, exact=False)
    O = oracle.laplace_matrix(np.float64, 16, 3, base=1)                 # 46 wavefront levels, 4096 rows
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    b = np.ones(O.n)
    x = isb.ssor(A, b, 1.2, maxiter=8)
    assert relerr(x, oracle.ssor_(np.zeros(O.n), O, b, 1.2, maxiter=8)) <= 1e-13
    # a dense matrix selects the arithmetic of the reference's dense methods (src/stationary.jl), dense SSOR included
    rng = np.random.default_rng(7)
    n = 12
    D = (rng.random((n, n)) + 2 * n * np.eye(n)).astype(dtype)
    bd, x0 = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
    eps = float(np.finfo(dtype).eps)
    for kind in ("jacobi", "gauss_seidel", "sor", "ssor"):
        fn = getattr(isb, kind + "_")
        xs = fn(x0.copy(), D, bd, 1.2, maxiter=5) if kind in ("sor", "ssor") else fn(x0.copy(), D, bd, maxiter=5)
        xo = oracle.stationary_dense_(kind, x0, D, bd, 1.2, maxiter=5)
        assert np.abs(xs - xo).max() <= 40 * eps * np.abs(xo).max(), kind


def test_cg_solve_forwards_a_callback_preconditioner(isb):
    """b200_cg_solve on a CSR operator with B200_PREC_CALLBACK in the option block runs the general CG engine (what a C / Julia
    caller gets; the Python cg_ reaches the same engine through b200_cg_solve_op)."""
    import ctypes as C
    from iterativesolvers_jl_b200 import _lib
    rng = np.random.default_rng(SEED)
    n = 500
    M = sp.random(n, n, 0.01, random_state=5, format="csc")
    M = (M + M.T + 10 * sp.eye(n)).tocsc()
    A = isb.B200CSR.from_scipy(M)
    jac = isb.JacobiPrec(M.diagonal())
    Pl = isb.FunctionPrec(n, np.float64, lambda y, v: jac.ldiv_(y, v))
    b = rng.standard_normal(n)
    x_ref, h_ref = isb.cg(A, b, Pl=jac, log=True)                                 # tuned engine, Jacobi
    xd, bd = isb.DeviceArray.zeros(A.ctx, n), isb.DeviceArray.from_numpy(A.ctx, b)
    opts = _lib.CgOpts(0.0, float(np.sqrt(np.finfo(np.float64).eps)), n, 1, 0, isb.operators.precond_to_c(Pl, A), 0, 0)
    res = _lib.Result()
    hist = np.zeros(n + 1)
    assert isb.lib().b200_cg_solve(A.ctx._h, A._h, xd._p, bd._p, C.byref(opts), C.byref(res),
                                   hist.ctypes.data_as(C.c_void_p), n + 1) == 0
    Pl.op.raise_pending()
    assert res.isconverged and abs(res.iters - h_ref.iters) <= 1 and relerr(xd.numpy(), x_ref) <= 1e-7
