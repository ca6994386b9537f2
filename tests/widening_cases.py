"""Engine-vs-oracle cases for qmr!/lsqr!/lsmr!/idrs!, shared by
  * tests/test_oracle_widening.py  -- the engines on the serial test backend (tests/hostsim), CPU, and
  * tests/test_zy_gpu_widening.py  -- the same engines through the C ABI on the GPU (-m gpu).
A `runner` hides which of the two executes the engine:
    runner.qmr(x0, A, b, **kw)            -> x, R   (R.iters, R.converged, R.breakdown, R.hist, R.tol, R.nprods)
    runner.lsqr / runner.lsmr(x0, A, b, **kw) -> x, R   (R.iters, R.istop, R.converged, R.mvps, R.mtvps, R.hist{...},
                                                        R.ctol, R.early, R.bad_x)
    runner.idrs(x0, A, b, P, diag=None, **kw) -> x, R
    runner.cg(x0, A, b, mode, d, **kw)    -> x, R   general-operator cg!: mode "cg" (Identity), "jacobi" (diagonal d
                                                    through the fused division), "callback" (ldiv! by callback)
A is a scipy sparse matrix; x0 is updated in place and returned.
"""
import math

import pytest

import numpy as np
import scipy.sparse as sp

SEED = 1234321


def sol_matrix(m, n):
    """reference test/lsqr.jl:25-29 / test/lsmr.jl:59-63."""
    mn = min(m, n)
    I = np.r_[np.arange(1, mn), np.arange(mn)]
    J = np.r_[np.arange(mn - 1), np.arange(mn)]
    V = np.r_[np.arange(1.0, mn), np.arange(1.0, mn + 1)]
    return sp.coo_matrix((V, (I, J)), shape=(m, n)).tocsc()


def tridiag(T):
    return np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=T)


def _systems():
    yield 10, sp.random(10, 10, 0.5, random_state=3, format="csc") + 10 * sp.eye(10, format="csc")
    yield 200, sp.random(200, 200, 0.05, random_state=4, format="csc") + 8 * sp.eye(200, format="csc")
    yield 1000, sp.random(1000, 1000, 0.01, random_state=5, format="csc") + 6 * sp.eye(1000, format="csc")


def case_qmr_matches_oracle(oracle, runners, dtype, tol):
    rng = np.random.default_rng(7)
    for n, M in _systems():
        A = M.astype(dtype)
        O = oracle.CSC.from_scipy(A.tocsc(), base=1)
        b, x0 = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
        for init_zero in (False, True):
            start = np.zeros(n, dtype) if init_zero else x0
            xo, ho = oracle.qmr_(start.copy(), O, b, log=True, initially_zero=init_zero)
            for run in runners:
                xs, hs = run.qmr(start.copy(), A, b, initially_zero=init_zero, check_every=3)
                assert hs.iters == ho.iters and hs.converged == ho.isconverged and not hs.breakdown
                assert hs.nprods == 2 * ho.iters + (0 if init_zero else 1)
                assert np.max(np.abs(hs.hist - ho["resnorm"])) <= tol * ho["resnorm"][0]
                assert np.linalg.norm(xs - xo) <= tol * np.linalg.norm(xo)
                assert abs(hs.tol - ho["tol"]) <= 1e-6 * ho["tol"]


def case_qmr_advection_maxiter_breakdown(oracle, runners):
    """a non-symmetric problem that needs ~100 Lanczos steps; maxiter; zero iterations; exact breakdown."""
    M, b = oracle.advection_dominated(8, 50.0)
    O = oracle.CSC.from_scipy(M.tocsc(), base=1)
    xo, ho = oracle.qmr(O, b, log=True, reltol=1e-8)
    assert ho.isconverged and ho.iters > 40
    rng = np.random.default_rng(1)
    A5, b5 = sp.csr_matrix(rng.random((5, 5))), rng.random(5)
    # documented deviation: at delta == 0 the reference updates x with the wrong Lanczos vector (x stays 0 for A = I)
    xi, hi = oracle.qmr(np.eye(4), np.ones(4), log=True)
    assert np.all(xi == 0) and hi.iters == 1 and hi["resnorm"][0] == 0
    for run in runners:
        xs, hs = run.qmr(np.zeros_like(b), M, b, initially_zero=True, reltol=1e-8)
        assert abs(hs.iters - ho.iters) <= 2 and hs.converged
        k = min(20, ho.iters)
        assert np.max(np.abs(hs.hist[:k] - ho["resnorm"][:k]) / ho["resnorm"][:k]) <= 1e-9
        # resnorm is the quasi-residual; the true residuals of engine and oracle must agree with each other
        assert np.linalg.norm(M @ xs - b) <= 3 * np.linalg.norm(M @ xo - b) <= 1e-5 * np.linalg.norm(b)
        assert np.linalg.norm(xs - xo) <= 1e-6 * np.linalg.norm(xo)
        xs, hs = run.qmr(np.zeros(5), A5, b5, initially_zero=True, maxiter=2)                    # test/qmr.jl:38-42
        assert hs.iters == 2 and len(hs.hist) == 2
        xs, hs = run.qmr(np.zeros(4), sp.eye(4, format="csr"), np.zeros(4), initially_zero=True)  # zero rhs
        assert hs.iters == 0 and np.all(xs == 0)
        xs, hs = run.qmr(np.zeros(4), sp.eye(4, format="csr"), np.ones(4), initially_zero=True)
        assert hs.breakdown and hs.iters == 1 and hs.hist[0] == 0 and np.allclose(xs, 1.0)
        for T in (np.float32, np.float64):                                                        # test/qmr.jl:44-66
            A = sp.csr_matrix(tridiag(T))
            bb = np.ones(3, dtype=T)
            x0 = np.linalg.solve(A.toarray().astype(np.float64), bb.astype(np.float64)).astype(T)
            pert = (10 * math.sqrt(np.finfo(T).eps) * np.array([-1, 1, -1])).astype(T)
            xs, hs = run.qmr(x0 + pert, A, bb)
            assert 2 <= hs.iters <= 3
            r0 = np.linalg.norm(A @ (x0 + pert) - bb)
            xs, hs = run.qmr(x0 + pert, A, bb, abstol=2 * r0, reltol=0.0)
            assert hs.iters == 0


LS_CASES = [
    ("tall", lambda: sp.random(300, 120, 0.05, random_state=1, format="csc"), {}),
    ("wide", lambda: sp.random(120, 300, 0.05, random_state=2, format="csc"), {}),
    ("sol20x10", lambda: sol_matrix(20, 10), dict(atol=1e-7, btol=1e-7, conlim=1e10, maxiter=100)),
    ("sol10x10", lambda: sol_matrix(10, 10), dict(atol=1e-7, btol=1e-7, conlim=1e10, maxiter=100)),
]


def case_lsqr_lsmr_match_oracle(oracle, runners, solver, dtype):
    """The Golub-Kahan recurrences amplify rounding differences exponentially once the process has converged
    (measured: 1e-16 -> 1e-2 between iterations 10 and 18 on `tall`, identically between two summation orders of
    the SAME code), so histories are compared over the first 8 iterations (1e-9 of their scale in fp64) and the
    solutions at the accuracy the stopping rule delivers; counters, stopping rule and flags must agree (+-2 iterations)."""
    rng = np.random.default_rng(5)
    orc = getattr(oracle, solver + "_")
    damp_kw = "damp" if solver == "lsqr" else "lam"
    first = "resnorm" if solver == "lsqr" else None
    htol, xtol = (1e-9, 2e-6) if dtype == np.float64 else (2e-3, 2e-2)
    for name, mk, kw in LS_CASES:
        A = mk().astype(dtype)
        m, n = A.shape
        O = oracle.CSC.from_scipy(A.tocsc(), base=1)
        b, x0 = rng.random(m).astype(dtype), rng.random(n).astype(dtype)
        for extra in ({}, {damp_kw: 0.1}):
            xo, ho = orc(x0.copy(), O, b, log=True, **kw, **extra)
            for run in runners:
                xs, hs = getattr(run, solver)(x0.copy(), A, b, check_every=5, **kw, **extra)
                assert abs(hs.iters - ho.iters) <= 2, (name, extra)
                assert hs.converged == ho.isconverged
                if hs.iters == ho.iters:
                    assert hs.istop == ho["istop"], (name, extra)
                assert hs.mvps - hs.iters == ho.mvps - ho.iters and hs.mtvps - hs.iters == ho.mtvps - ho.iters
                k = min(8 if dtype == np.float64 else 5, ho.iters, hs.iters)
                for key in ([first] if first else []) + ["anorm", "rnorm", "cnorm"]:
                    ref = ho[key][:k]
                    assert np.max(np.abs(hs.hist[key][:k] - ref)) <= htol * np.max(np.abs(ref)), (name, key)
                assert np.linalg.norm(xs - xo) <= xtol * np.linalg.norm(xo), (name, extra)
                assert abs(hs.ctol - ho["ctol"]) <= 1e-6 * ho["ctol"]


def case_lsqr_lsmr_edge_cases(oracle, runners):
    A = sp.csr_matrix(np.eye(3))
    x0 = np.array([1.0, 2.0, 3.0])
    xo, ho = oracle.lsqr_(x0.copy(), np.eye(3), x0.copy(), log=True)
    assert ho.iters == 0 and not ho.isconverged and ho.mtvps == 0
    rng = np.random.default_rng(2)
    M = sp.random(40, 30, 0.3, random_state=3, format="csr")
    b = rng.random(40)
    for run in runners:
        # initial guess not finite (src/lsqr.jl:102-104)
        xs, hs = run.lsqr(np.array([0.0, np.inf, 0.0]), A, np.ones(3))
        assert hs.bad_x and hs.iters == 0
        # b - A x == 0: lsqr returns at once (src/lsqr.jl:141-144), history empty, not converged
        xs, hs = run.lsqr(x0.copy(), A, x0.copy())
        assert hs.iters == 0 and not hs.converged and np.array_equal(xs, x0) and hs.mtvps == 0
        # lsmr: the same input divides by zero in the reference (NaN iterates); the engine takes the announced exit
        xs, hs = run.lsmr(x0.copy(), A, x0.copy())
        assert hs.iters == 0 and hs.converged and np.array_equal(xs, x0) and (hs.mvps, hs.mtvps) == (1, 1)
        # maxiter
        for solver in ("lsqr", "lsmr"):
            xs, hs = getattr(run, solver)(np.zeros(30), M, b, maxiter=3, atol=0.0, btol=0.0, conlim=0.0)
            xo, ho = getattr(oracle, solver + "_")(np.zeros(30), oracle.CSC.from_scipy(M.tocsc()), b, maxiter=3,
                                                    atol=0.0, btol=0.0, conlim=0.0, log=True)
            assert hs.iters == ho.iters == 3 and hs.istop == ho["istop"] == 7 and hs.converged == ho.isconverged
            assert np.linalg.norm(xs - xo) <= 1e-12 * np.linalg.norm(xo)


def case_idrs_matches_oracle(oracle, runners, dtype, tol):
    rng = np.random.default_rng(11)
    for n, M in _systems():
        A = M.astype(dtype)
        O = oracle.CSC.from_scipy(A.tocsc(), base=1)
        b, x0 = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
        for s in (1, 2, 4, 8):
            P = np.asfortranarray(rng.random((n, s)).astype(dtype))
            for smoothing in (False, True):
                for jac in (False, True):
                    d = A.diagonal().astype(dtype)
                    xo, ho = oracle.idrs_(x0.copy(), O, b, s=s, P=[P[:, j].copy() for j in range(s)], log=True,
                                          smoothing=smoothing, Pl=oracle.JacobiPrec(d) if jac else None)
                    for run in runners:
                        xs, hs = run.idrs(x0.copy(), A, b, P, smoothing=smoothing, diag=d if jac else None,
                                          check_every=3)
                        assert hs.iters == ho.iters and hs.converged == ho.isconverged and not hs.breakdown
                        # fp32: IDR(s)'s intermediate residual peaks differ by up to ~1 % between fp32 and fp64 scalars
                        htol = tol if dtype == np.float64 else 2e-2
                        assert np.max(np.abs(hs.hist - ho["resnorm"])) <= htol * ho["resnorm"][0], (n, s, smoothing, jac)
                        assert np.linalg.norm(xs - xo) <= tol * np.linalg.norm(xo), (n, s, smoothing, jac)


def case_idrs_s16_maxiter_zero_iterations(oracle, runners):
    rng = np.random.default_rng(3)
    M, b = oracle.advection_dominated(8, 50.0)
    n = M.shape[0]
    O = oracle.CSC.from_scipy(M.tocsc(), base=1)
    P = np.asfortranarray(rng.random((n, 16)))
    xo, ho = oracle.idrs(O, b, s=16, P=[P[:, j].copy() for j in range(16)], log=True, reltol=1e-8)
    assert ho.isconverged
    A5, b5, P5 = sp.csr_matrix(rng.random((5, 5))), rng.random(5), rng.random((5, 8))
    A = sp.csr_matrix(tridiag(np.float64))
    bb = np.ones(3)
    x0 = np.linalg.solve(A.toarray(), bb) + 10 * math.sqrt(np.finfo(np.float64).eps) * np.array([-1, 1, -1])
    r0 = np.linalg.norm(A @ x0 - bb)
    P3 = rng.random((3, 8))
    for run in runners:
        xs, hs = run.idrs(np.zeros(n), M, b, P, reltol=1e-8)
        assert hs.converged and abs(hs.iters - ho.iters) <= 2
        k = min(34, ho.iters, hs.iters)                      # two full cycles of 17 steps
        assert np.max(np.abs(hs.hist[:k] - ho["resnorm"][:k]) / ho["resnorm"][:k]) <= 1e-7
        assert np.linalg.norm(M @ xs - b) <= 1e-7 * np.linalg.norm(b)
        xs, hs = run.idrs(np.zeros(5), A5, b5, P5, maxiter=2)                                     # test/idrs.jl:65-69
        assert hs.iters == 2 and len(hs.hist) == 2
        xs, hs = run.idrs(x0.copy(), A, bb, P3, abstol=2 * r0, reltol=0.0)                        # :100-104
        assert hs.iters == 0 and hs.converged and np.array_equal(xs, x0)


def case_cg_general_operator(oracle, runners, dtype, tol):
    """cg! / pcg through the general-operator engine (csrc/cg_core.h): operator and preconditioner callbacks."""
    rng = np.random.default_rng(3)
    for N, dims in ((16, 2), (10, 3)):
        L = oracle.laplace_matrix(dtype, N, dims, base=1).to_scipy().tocsr()
        n = L.shape[0]
        D = sp.diags(1.0 + rng.random(n)).astype(dtype)          # a non-constant diagonal, so that Jacobi matters
        M = (D @ L @ D).tocsr().astype(dtype)
        O = oracle.CSC.from_scipy(M.tocsc(), base=1)
        b, x0 = rng.standard_normal(n).astype(dtype), rng.standard_normal(n).astype(dtype)
        d = M.diagonal().astype(dtype)
        for init_zero in (False, True):
            start = np.zeros(n, dtype) if init_zero else x0
            for mode in ("cg", "jacobi", "callback"):
                Pl = oracle.JacobiPrec(d.copy()) if mode != "cg" else None
                xo, ho = oracle.cg_(start.copy(), O, b, log=True, Pl=Pl, initially_zero=init_zero)
                for run in runners:
                    xs, hs = run.cg(start.copy(), M, b, mode, d, initially_zero=init_zero, check_every=5)
                    assert hs.iters == ho.iters and hs.converged == ho.isconverged and hs.mvps == ho.mvps
                    # unpreconditioned CG on this badly scaled matrix is the sensitive one (1e-10 between two orders)
                    t = tol if mode != "cg" or dtype == np.float32 else max(tol, 2e-9)
                    assert np.max(np.abs(hs.hist - ho["resnorm"])) <= t * ho["resnorm"][0], (N, mode)
                    assert np.linalg.norm(xs - xo) <= t * np.linalg.norm(xo), (N, mode)
    # reference termination tests (test/cg.jl:98-122) through the general engine
    for T in (np.float32, np.float64):
        A = sp.csr_matrix(tridiag(T))
        bb = np.ones(3, dtype=T)
        x0 = np.linalg.solve(A.toarray().astype(np.float64), bb.astype(np.float64)).astype(T)
        pert = (10 * math.sqrt(np.finfo(T).eps) * np.array([-1, 1, -1])).astype(T)
        r0 = np.linalg.norm(A @ (x0 + pert) - bb)
        for run in runners:
            xs, hs = run.cg(x0 + pert, A, bb, "cg", None)
            assert 2 <= hs.iters <= 3
            xs, hs = run.cg(x0 + pert, A, bb, "cg", None, abstol=2 * r0, reltol=0.0)
            assert hs.iters == 0
            xs, hs = run.cg(np.zeros(3, dtype=T), A, np.zeros(3, dtype=T), "cg", None, initially_zero=True)   # :50-51
            assert hs.iters == 0 and np.all(xs == 0)


def case_constraint_apply(oracle, apply_fns, dtype, tol):
    """Constraint (reference src/lobpcg.jl:144-224): X <- X - Y (chol(Y'Y) \\ Y'X) against the oracle's Constraint,
    for narrow / full / wider-than-16 bases, and with update!'s identity-extended factor (:188-206).
    apply_fns: callables (X, Y, appended) -> X (in place)."""
    rng = np.random.default_rng(2)
    for n, nc, bs in ((50, 1, 1), (200, 5, 3), (300, 16, 16), (400, 37, 16), (64, 20, 7)):
        Y = rng.standard_normal((n, nc)).astype(dtype)
        X = rng.standard_normal((n, bs)).astype(dtype)
        c = oracle.Constraint(Y.copy())
        Xo = np.asfortranarray(X.copy())
        c.apply(Xo)
        assert np.abs(Y.T.astype(np.float64) @ Xo).max() <= 100 * tol * np.linalg.norm(Y) * np.linalg.norm(Xo)
        # appended columns: orthonormal, orthogonal to Y (what a constrained lobpcg batch delivers)
        Q = rng.standard_normal((n, 3))
        oracle.Constraint(Y.astype(np.float64)).apply(Q)
        Q = np.linalg.qr(Q)[0].astype(dtype)
        c3 = oracle.Constraint(Y.copy())
        c3.update(Q.copy(), Q.copy())
        Xo3 = np.asfortranarray(X.copy())
        c3.apply(Xo3)
        for fn in apply_fns:
            Xs = fn(np.asfortranarray(X.copy()), Y, 0)
            assert np.linalg.norm(Xs - Xo) <= tol * np.linalg.norm(Xo), (n, nc, bs)
            Xs = fn(np.asfortranarray(X.copy()), np.hstack([Y, Q]), 3)
            assert np.linalg.norm(Xs - Xo3) <= tol * np.linalg.norm(Xo3), (n, nc, bs, "appended")


def separated_spectrum_matrix(n, seed=7):
    """dense symmetric test matrix with a prescribed, well separated spectrum at both ends (LOBPCG converges in tens
    of iterations, no multiple eigenvalues: the 2-D Laplacian's double eigenvalues make a block miss a copy)."""
    rng = np.random.default_rng(seed)
    Q = np.linalg.qr(rng.standard_normal((n, n)))[0]
    d = np.r_[np.array([1, 2, 4, 7, 11, 16, 22, 29, 37, 46.0]), np.linspace(60, 100, n - 20),
              np.array([120, 135, 150, 170, 190, 215, 240, 270, 300, 340.0])]
    return (Q * d) @ Q.T, np.sort(d)


def case_svdl_matches_oracle(oracle, run_svdl, dtype, tol, method="ritz"):
    """svdl (reference src/svdl.jl): engine vs oracle -- identical iteration counts and product counts, Ritz values,
    error bounds (:resnorm), convergence flags, betas and singular values; and the reference's own tests
    (test/svdl.jl:16-69: diagonal matrix incl. the +-1 structure of the vectors, rectangular random matrix).
    run_svdl(A_scipy, v0, **kw) -> dict(sigma, U, V, iters, mvps, mtvps, converged, ritz, resnorm, conv, betas).
    method: "ritz" (thickrestart! :376-404) or "harmonic" (harmonicrestart! :424-493; the reference's tests run both,
    test/svdl.jl:13; dolock exists for :ritz only)."""
    mk = dict(method=method)
    n, ns, t = 30, 5, 1e-5
    A = np.diag(np.arange(1.0, n + 1)).astype(dtype)
    q = (np.ones(n) / np.sqrt(n)).astype(dtype)
    (Uo, So, Vto), L, h = oracle.svdl(A, nsv=ns, v0=q, tol=t, reltol=t, maxiter=n, vecs="both", log=True, **mk)
    r = run_svdl(sp.csr_matrix(A), q, nsv=ns, tol=t, reltol=t, maxiter=n, **mk)
    assert r["iters"] == h.iters and r["converged"] and h.isconverged and (r["mvps"], r["mtvps"]) == (h.mvps, h.mtvps)
    assert np.linalg.norm(r["sigma"] - np.arange(n, n - 5, -1.0)) < 5 ** 2 * 1e-5            # test/svdl.jl:27
    assert np.abs(r["sigma"] - So).max() <= tol * So[0]
    assert np.abs(np.array(h["ritz"]) - r["ritz"]).max() <= tol * So[0]
    assert np.abs(np.array(h["resnorm"]) - r["resnorm"]).max() <= tol * So[0]
    assert np.abs(np.array(h["betas"]) - r["betas"]).max() <= tol * So[0]
    assert np.array_equal(np.array(h["conv"]), r["conv"])
    U, V = r["U"].copy(), r["V"].copy()                                                      # test/svdl.jl:33-49
    for i in range(5):
        U[n - 1 - i, i] -= np.sign(U[n - 1 - i, i])
        V[n - 1 - i, i] -= np.sign(V[n - 1 - i, i])
    assert np.linalg.norm(U) < So[0] * math.sqrt(t)
    if method == "ritz":                                                 # (the reference checks U twice, test/svdl.jl:43, :47)
        assert np.linalg.norm(V) < So[0] * math.sqrt(t)
    # rectangular
    rng = np.random.default_rng(1)
    m, n2, k, l = 300, 200, 5, 10
    A = rng.standard_normal((m, n2)).astype(dtype)
    q = rng.standard_normal(n2).astype(dtype)
    q /= np.linalg.norm(q)
    so, L, h = oracle.svdl(A, nsv=k, k=l, v0=q, tol=1e-5, maxiter=30, log=True, **mk)
    r = run_svdl(sp.csr_matrix(A), q, nsv=k, k=l, tol=1e-5, maxiter=30, **mk)
    exact = np.linalg.svd(A.astype(np.float64), compute_uv=False)[:k]
    assert np.linalg.norm(r["sigma"] - exact) < k ** 2 * 1e-5                                # test/svdl.jl:66
    assert abs(r["iters"] - h.iters) <= (0 if dtype == np.float64 else 1)
    if r["iters"] == h.iters:
        assert np.abs(r["sigma"] - so).max() <= tol * so[0]
        assert np.abs(np.array(h["resnorm"]) - r["resnorm"]).max() <= 10 * tol * so[0]
    # the singular vectors really are singular vectors: A v = sigma u, to the accuracy asked for
    if method == "ritz":
        assert np.linalg.norm(A.astype(np.float64) @ r["V"] - r["U"] * r["sigma"][None, :]) <= 1e-3 * exact[0]
    # maxiter cut and dolock
    r1 = run_svdl(sp.csr_matrix(A), q, nsv=k, k=l, tol=1e-12, reltol=1e-14, maxiter=3, **mk)
    s1, L1, h1 = oracle.svdl(A, nsv=k, k=l, v0=q, tol=1e-12, reltol=1e-14, maxiter=3, log=True, **mk)
    assert r1["iters"] == h1.iters == 3 and not r1["converged"] and not h1.isconverged
    assert np.abs(r1["sigma"] - s1).max() <= 10 * tol * s1[0]
    if method != "ritz":
        return
    r2 = run_svdl(sp.csr_matrix(A), q, nsv=k, k=l, tol=1e-5, maxiter=30, dolock=True)
    s2, L2, h2 = oracle.svdl(A, nsv=k, k=l, v0=q, tol=1e-5, maxiter=30, dolock=True, log=True)
    assert abs(r2["iters"] - h2.iters) <= 1 and np.linalg.norm(r2["sigma"] - exact) < k ** 2 * 1e-5


def case_idrs_callback_preconditioner(oracle, runners):
    """idrs! with ldiv!(Pl, V) by callback (src/idrs.jl:199, :246): the fused direction pass is split around the
    callback; against the oracle with the same diagonal preconditioner.  runner.idrs(..., cb_diag=d)."""
    rng = np.random.default_rng(11)
    n = 200
    A = (sp.random(n, n, 0.05, random_state=4, format="csc") + sp.diags(4 + 8 * rng.random(n))).tocsc()
    O = oracle.CSC.from_scipy(A, base=1)
    b, x0, d = rng.random(n), rng.random(n), A.diagonal()
    for s in (1, 4, 8):
        P = np.asfortranarray(rng.random((n, s)))
        for smoothing in (False, True):
            xo, ho = oracle.idrs_(x0.copy(), O, b, s=s, P=[P[:, j].copy() for j in range(s)], log=True,
                                  smoothing=smoothing, Pl=oracle.JacobiPrec(d.copy()))
            for run in runners:
                xs, hs = run.idrs(x0.copy(), A, b, P, smoothing=smoothing, cb_diag=d)
                assert hs.iters == ho.iters and hs.converged
                assert np.max(np.abs(hs.hist - ho["resnorm"])) <= 1e-9 * ho["resnorm"][0]
                assert np.linalg.norm(xs - xo) <= 1e-9 * np.linalg.norm(xo)


def case_lobpcg_general(oracle, run, dtype, tol, ltol, same_arithmetic=True):
    """the general LOBPCG engine (csrc/lobpcg_general_core.h) against the oracle's lobpcg: standard and generalized
    problem, block sizes 1 and 3, Jacobi and callback preconditioner, constraint in the (B-)inner product -- identical
    iteration counts in fp64 (the two follow the same recurrence to rounding), Ritz values and residual norms.
    run(A, largest, X0, B=None, jac=None, cb_diag=None, C=None, tol, maxiter) -> dict(lam, X, resnorm, iterations,
    converged, status) with status != 0 for a PosDefException.
    same_arithmetic=False (the GPU: other summation order in every reduction) allows the iteration counts to differ by
    max(3, 10 %) and does not require an fp32 PosDefException -- which hangs on the last bits of a Gram matrix -- to
    strike in the same configuration, only that most configurations run to the end on both sides."""
    rng = np.random.default_rng(5)
    compared = 0

    def same_count(a, b):
        return a == b if (exact64 and same_arithmetic) else abs(a - b) <= max(3 if exact64 else 5, b // 10)

    n = 60
    M, d = separated_spectrum_matrix(n)
    Bm = rng.standard_normal((n, n))
    Bm = Bm @ Bm.T / n + 2 * np.eye(n)                       # symmetric positive definite
    dg = np.abs(M.diagonal()) + 1.0
    Md = M.astype(dtype)
    exact64 = dtype == np.float64
    for gen in (False, True):
        Bd = Bm.astype(dtype) if gen else None
        Bq = Bm if gen else np.eye(n)
        for largest in (False, True):
            for bs in (1, 3):
                X0 = rng.random((n, bs)).astype(dtype)
                r = run(Md, largest, X0, B=Bd, tol=tol, maxiter=300)
                try:
                    ro = oracle.lobpcg(Md, largest, X0, B=Bd, tol=tol, maxiter=300, not_zeros=True)
                except np.linalg.LinAlgError:
                    # PosDefException in the reference's algorithm (fp32, clustered Ritz values): the engine must break
                    # down the same way (status != 0), not return something
                    assert not exact64 and (r["status"] != 0 or not same_arithmetic), (gen, largest, bs)
                    continue
                if r["status"] != 0 and not exact64 and not same_arithmetic:
                    continue
                compared += 1
                assert r["status"] == 0 and ro.converged and r["converged"]
                assert same_count(r["iterations"], ro.iterations), (r["iterations"], ro.iterations)
                assert np.abs(np.sort(ro.lam) - np.sort(r["lam"])).max() <= ltol * np.abs(ro.lam).max()
                if "trace" in r and exact64 and same_arithmetic:            # log = true: the LOBPCGState of every iteration
                    rt = oracle.lobpcg(Md, largest, X0, B=Bd, tol=tol, maxiter=300, not_zeros=True, log=True).trace
                    assert len(r["trace"]) == len(rt) == ro.iterations
                    for (i1, rn1, l1), (i2, rn2, l2) in zip(r["trace"], rt):
                        assert i1 == i2 and np.abs(l1 - l2).max() <= 1e-8 * np.abs(ro.lam).max()
                        assert np.abs(rn1 - rn2).max() <= 1e-4 * np.abs(rn2).max() + 1e-2 * tol   # (residuals near tol carry rounding noise)
                X = np.asarray(r["X"], dtype=np.float64)
                assert np.max(np.linalg.norm(M @ X - Bq @ X * r["lam"][None, :], axis=0)) <= 4 * tol
                assert np.abs(X.T @ Bq @ X - np.eye(bs)).max() <= 2 * n * tol          # test/lobpcg.jl:62-69
        if not exact64:
            assert compared >= (gen + 1) * 2, compared
            continue           # fp32 + preconditioner / constraint: the reference itself runs into PosDefException here
        X0 = rng.random((n, 3)).astype(dtype)
        ro = oracle.lobpcg(Md, False, X0, B=Bd, P=oracle.JacobiPrec(dg.astype(dtype)), tol=tol, maxiter=300, not_zeros=True)
        for kw in (dict(jac=dg), dict(cb_diag=dg)):
            r = run(Md, False, X0, B=Bd, tol=tol, maxiter=300, **kw)
            assert r["converged"] and same_count(r["iterations"], ro.iterations)
            assert np.abs(np.sort(ro.lam) - np.sort(r["lam"])).max() <= ltol * np.abs(ro.lam).max()
        rc = oracle.lobpcg(Md, False, rng.random((n, 2)).astype(dtype), B=Bd, tol=tol, maxiter=300, not_zeros=True)
        X1 = rng.random((n, 3)).astype(dtype)
        ro = oracle.lobpcg(Md, False, X1, B=Bd, C=rc.X.copy(), tol=tol, maxiter=300, not_zeros=True)
        r = run(Md, False, X1, B=Bd, C=rc.X.copy(), tol=tol, maxiter=300)
        assert r["converged"] and same_count(r["iterations"], ro.iterations)
        assert np.abs(np.sort(ro.lam) - np.sort(r["lam"])).max() <= ltol * np.abs(ro.lam).max()
        assert np.abs(rc.X.astype(np.float64).T @ Bq @ np.asarray(r["X"], dtype=np.float64)).max() <= 2 * n * tol


def case_nev_driver(lobpcg, make_A, block_size, nev):
    """lobpcg(A, largest, X0, nev) (reference test/lobpcg.jl:291-306, :324-342): batches with deflation, with and without
    the tail batch (`cutoff` branch, src/lobpcg.jl:945-952); residuals, orthonormality, the prescribed eigenvalues; an
    initial constraint on top; lobpcg(A, largest, nev::Int).  lobpcg: the package's function; make_A(M) its operator."""
    rng = np.random.default_rng(SEED)
    n = 60
    M, d = separated_spectrum_matrix(n)
    A = make_A(M)
    tol = 1e-6
    for largest in (False, True):
        ex = d[::-1] if largest else d
        r = lobpcg(A, largest, rng.random((n, block_size)), nev, tol=tol, maxiter=2000, rng=rng, log=True)
        assert r.X.shape == (n, nev) and len(r.iterations) == -(-nev // block_size) and np.all(r.converged)
        assert len(r.trace) == len(r.iterations)                    # one LOBPCGTrace per batch (src/lobpcg.jl:74, :88)
        for tr, its in zip(r.trace, r.iterations):
            assert len(tr) == its and [t[0] for t in tr] == list(range(1, its + 1))
            assert np.all(tr[-1][1] <= tol) and tr[-1][1].shape == tr[-1][2].shape
        assert np.max(np.linalg.norm(M @ r.X - r.X * r.lam[None, :], axis=0)) <= tol
        assert np.allclose(r.X.T @ r.X, np.eye(nev), atol=2 * n * tol)
        assert np.allclose(np.sort(r.lam), np.sort(ex[:nev]), atol=1e-5)
    # with an initial constraint: the eigenpairs after the first two
    r1 = lobpcg(A, False, rng.random((n, 2)), tol=tol, maxiter=2000)
    k = max(3, 2 * block_size) if block_size > 1 else 3
    r2 = lobpcg(A, False, rng.random((n, block_size)), k, C=r1.X.copy(), tol=tol, maxiter=2000, rng=rng)
    assert np.allclose(np.sort(r2.lam), d[2:2 + k], atol=1e-5)
    assert np.max(np.abs(r1.X.T @ r2.X)) <= 2 * n * tol
    if block_size == 1:
        r3 = lobpcg(A, False, 3, tol=tol, maxiter=2000, rng=rng)                           # lobpcg(A, largest, nev::Int)
        assert np.allclose(np.sort(r3.lam), d[:3], atol=1e-5)


def case_gmres_general(oracle, run, dtype, tol):
    """the general gmres! engine (csrc/gmres_core.h: callback operator, Pl / Pr as Jacobi diagonals or callbacks, the three
    orthogonalisation methods) against the oracle's gmres_: iteration and product counts, the residual history, x.
    run(x0, A, b, d, pl, pr, restart, maxiter, orth_meth) -> (x, outcome with iters, mvps, converged, hist); pl / pr in
    {None, "jac", "cb"} select Identity, the Jacobi diagonal d, or d applied through a callback."""
    rng = np.random.default_rng(3)
    n = 300
    A = (sp.random(n, n, 0.03, random_state=1, format="csc") + 4 * sp.eye(n)).tocsc()
    b = rng.standard_normal(n)
    d = A.diagonal()
    Ad, bd = A.astype(dtype), b.astype(dtype)
    mk = lambda kind: None if kind is None else oracle.JacobiPrec(d.astype(dtype))
    for meth in ("mgs", "cgs", "dgks"):
        for pl, pr in ((None, None), ("jac", None), (None, "jac"), ("cb", "cb"), ("jac", "cb"), ("cb", None)):
            for restart in (5, 20):
                x0 = rng.standard_normal(n).astype(dtype)
                xo, ho = oracle.gmres_(x0.copy(), Ad, bd, Pl=mk(pl), Pr=mk(pr), restart=restart, maxiter=60, log=True,
                                       orth_meth=meth)
                xs, hs = run(x0.copy(), Ad, bd, d.astype(dtype), pl, pr, restart, 60, meth)
                ro = np.asarray(ho["resnorm"])
                assert hs.converged == ho.isconverged and abs(hs.iters - ho.iters) <= (0 if dtype == np.float64 else 1)
                if hs.iters == ho.iters:
                    assert hs.mvps == ho.mvps
                k = min(hs.iters, ho.iters)
                assert np.max(np.abs(ro[:k] - np.asarray(hs.hist)[:k])) <= tol * ro[0], (meth, pl, pr, restart)
                assert np.linalg.norm(xs - xo) <= 20 * tol * np.linalg.norm(xo), (meth, pl, pr, restart)
    # restart > 16 columns in one cycle (two dot / update passes), maxiter reached inside a cycle, zero iterations
    x0 = np.zeros(n, dtype=dtype)
    xo, ho = oracle.gmres_(x0.copy(), Ad, bd, restart=40, maxiter=23, reltol=1e-30, log=True, orth_meth="dgks", initially_zero=True)
    xs, hs = run(x0.copy(), Ad, bd, d.astype(dtype), None, None, 40, 23, "dgks", reltol=1e-30, initially_zero=True)
    assert hs.iters == ho.iters == 23 and hs.mvps == ho.mvps and not hs.converged
    assert np.linalg.norm(xs - xo) <= 20 * tol * np.linalg.norm(xo)
    xe = rng.standard_normal(n).astype(dtype)
    be = (Ad @ xe).astype(dtype)
    xs, hs = run(xe.copy(), Ad, be, d.astype(dtype), None, None, 20, 60, "mgs", abstol=1e-3)
    assert hs.iters == 0 and hs.converged and np.array_equal(xs, xe)


def case_minres_general(oracle, run, dtype, tol):
    """the general minres! engine (csrc/minres_core.h) against the oracle's minres_ on a symmetric indefinite and on a
    skew-symmetric operator: counts, residual history, x; maxiter; zero iterations.
    run(x0, A, b, **kw) -> (x, outcome with iters, mvps, converged, hist)."""
    rng = np.random.default_rng(3)
    n = 400
    R = sp.random(n, n, 0.02, random_state=2, format="csc")
    S = (0.1 * (R + R.T) + sp.diags(np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * np.linspace(1, 3, n))).tocsc()
    K = sp.random(n, n, 0.02, random_state=4, format="csc")
    K = (K - K.T).tocsc()
    b = rng.standard_normal(n).astype(dtype)
    for M, skew in ((S, False), (K, True)):
        Md = M.astype(dtype)
        for iz in (True, False):
            x0 = np.zeros(n, dtype) if iz else rng.standard_normal(n).astype(dtype)
            kw = dict(maxiter=80, initially_zero=iz, skew_hermitian=skew)
            xo, ho = oracle.minres_(x0.copy(), Md, b, log=True, **kw)
            xs, hs = run(x0.copy(), Md, b, **kw)
            ro = np.asarray(ho["resnorm"])
            assert abs(hs.iters - ho.iters) <= (0 if dtype == np.float64 else 1) and hs.converged == ho.isconverged
            if hs.iters == ho.iters:
                assert hs.mvps == ho.mvps
            k = min(hs.iters, ho.iters, 40)
            assert np.max(np.abs(ro[:k] - np.asarray(hs.hist)[:k])) <= tol * ro[0], (skew, iz)
            if ho.isconverged:                                  # (a run cut by maxiter is compared through its history)
                assert np.linalg.norm(xs - xo) <= 50 * tol * np.linalg.norm(xo), (skew, iz)
    xo, ho = oracle.minres_(np.zeros(n, dtype), S.astype(dtype), b, log=True, maxiter=7, reltol=1e-30, initially_zero=True)
    xs, hs = run(np.zeros(n, dtype), S.astype(dtype), b, maxiter=7, reltol=1e-30, initially_zero=True)
    assert hs.iters == ho.iters == 7 and not hs.converged and np.linalg.norm(xs - xo) <= 50 * tol * np.linalg.norm(xo)
    xe = rng.standard_normal(n).astype(dtype)
    xs, hs = run(xe.copy(), S.astype(dtype), (S.astype(dtype) @ xe).astype(dtype), abstol=1e-3)
    assert hs.iters == 0 and hs.converged and np.array_equal(xs, xe)


def case_bicgstabl_general(oracle, run, dtype, tol):
    """the general bicgstabl! engine (csrc/bicgstabl_core.h: callback operator, Pl as Jacobi diagonal or callback) against
    the oracle's bicgstabl_ with the same shadow residual: l = 1, 2, 4; counts, residual history, x; the product budget;
    the SingularException of the MR step (reference test/bicgstabl.jl:34-38).
    run(x0, A, b, l, shadow, d, pk, **kw) -> (x, outcome with iters, mvps, converged, hist, singular)."""
    rng = np.random.default_rng(3)
    n = 400
    A = (sp.random(n, n, 0.02, random_state=1, format="csc") + 4 * sp.eye(n)).tocsc()
    d = A.diagonal().astype(dtype)
    b = rng.standard_normal(n).astype(dtype)
    Ad = A.astype(dtype)
    for l in (1, 2, 4):
        for pk in (None, "jac", "cb"):
            for iz in (True, False):
                sh = rng.random(n).astype(dtype)
                x0 = np.zeros(n, dtype) if iz else rng.standard_normal(n).astype(dtype)
                Pl = None if pk is None else oracle.JacobiPrec(d)
                xo, ho = oracle.bicgstabl_(x0.copy(), Ad, b, l, Pl=Pl, max_mv_products=200, log=True, initial_zero=iz,
                                           r_shadow=sh)
                xs, hs = run(x0.copy(), Ad, b, l, sh, d, pk, max_mv_products=200, initial_zero=iz)
                ro = np.asarray(ho["resnorm"])
                assert not hs.singular and hs.converged == ho.isconverged
                assert abs(hs.iters - ho.iters) <= (0 if dtype == np.float64 else 1)
                if hs.iters == ho.iters:
                    assert hs.mvps == ho.mvps
                k = min(hs.iters, ho.iters)
                scale = 10.0 ** (l - 1)                          # the MR step's normal equations lose digits with l
                assert np.max(np.abs(ro[:k] - np.asarray(hs.hist)[:k])) <= scale * tol * ro[0], (l, pk, iz)
                assert np.linalg.norm(xs - xo) <= 50 * tol * np.linalg.norm(xo), (l, pk, iz)
    sh = rng.random(n).astype(dtype)
    xo, ho = oracle.bicgstabl_(np.zeros(n, dtype), Ad, b, 2, max_mv_products=9, reltol=1e-30, log=True, initial_zero=True,
                               r_shadow=sh)
    xs, hs = run(np.zeros(n, dtype), Ad, b, 2, sh, d, None, max_mv_products=9, reltol=1e-30, initial_zero=True)
    assert hs.iters == ho.iters == 3 and hs.mvps == ho.mvps == 12 and not hs.converged
    assert np.linalg.norm(xs - xo) <= 50 * tol * np.linalg.norm(xo)
    # exact solution after the BiCG part: rs[:, 2:end] = 0, the MR Gram matrix is singular -> SingularException
    I = sp.identity(16, format="csc").astype(dtype)
    xs, hs = run(np.zeros(16, dtype), I, np.ones(16, dtype), 2, np.ones(16, dtype), np.ones(16, dtype), None,
                 initial_zero=True)
    assert hs.singular


def case_chebyshev_general(oracle, run, dtype, tol):
    """the general chebyshev! engine (csrc/chebyshev_core.h: callback operator, Pl as Jacobi diagonal or callback) against
    the oracle's chebyshev_ (bounds from the exact spectrum, widened by 10 %): counts, history, x.
    run(x0, A, b, lmin, lmax, d, pk, **kw) -> (x, outcome with iters, mvps, converged, hist)."""
    rng = np.random.default_rng(5)
    n = 300
    R = sp.random(n, n, 0.03, random_state=3, format="csc")
    S = (R + R.T).tocsc()
    S = (S + sp.diags(np.asarray(np.abs(S).sum(axis=1)).ravel() + 1.0)).tocsc()
    d = S.diagonal()
    Ds = sp.diags(1 / np.sqrt(d))
    ev, evp = np.linalg.eigvalsh(S.toarray()), np.linalg.eigvalsh((Ds @ S @ Ds).toarray())
    b = rng.standard_normal(n).astype(dtype)
    Sd = S.astype(dtype)
    for pk in (None, "jac", "cb"):
        lo, hi = (ev[0] * 0.9, ev[-1] * 1.1) if pk is None else (evp[0] * 0.9, evp[-1] * 1.1)
        for iz in (True, False):
            x0 = np.zeros(n, dtype) if iz else rng.standard_normal(n).astype(dtype)
            Pl = None if pk is None else oracle.JacobiPrec(d.astype(dtype))
            xo, ho = oracle.chebyshev_(x0.copy(), Sd, b, lo, hi, Pl=Pl, maxiter=200, log=True, initially_zero=iz)
            xs, hs = run(x0.copy(), Sd, b, lo, hi, d.astype(dtype), pk, maxiter=200, initially_zero=iz)
            ro = np.asarray(ho["resnorm"])
            assert ho.isconverged and hs.converged and abs(hs.iters - ho.iters) <= (0 if dtype == np.float64 else 1)
            if hs.iters == ho.iters:
                assert hs.mvps == ho.mvps
            k = min(hs.iters, ho.iters)
            assert np.max(np.abs(ro[:k] - np.asarray(hs.hist)[:k])) <= tol * ro[0], (pk, iz)
            assert np.linalg.norm(xs - xo) <= 20 * tol * np.linalg.norm(xo), (pk, iz)
    xo, ho = oracle.chebyshev_(np.zeros(n, dtype), Sd, b, ev[0], ev[-1], maxiter=5, reltol=1e-30, log=True, initially_zero=True)
    xs, hs = run(np.zeros(n, dtype), Sd, b, ev[0], ev[-1], d.astype(dtype), None, maxiter=5, reltol=1e-30, initially_zero=True)
    assert hs.iters == ho.iters == 5 and not hs.converged and np.linalg.norm(xs - xo) <= 20 * tol * np.linalg.norm(xo)


def case_nev_driver_generalized(lobpcg, make_A):
    """lobpcg(A, B, largest, X0, nev) (reference src/lobpcg.jl:925-962 with B given): batches of 2, five pairs, against the
    dense generalized eigenproblem; residuals A x - lambda B x, B-orthonormality across the batches (the deflation basis is
    kept B-orthogonal through update!(constraint, X, BX), :188-206)."""
    import scipy.linalg as sla
    rng = np.random.default_rng(SEED)
    n = 60
    M, d = separated_spectrum_matrix(n)
    Bm = rng.standard_normal((n, n))
    Bm = Bm @ Bm.T / n + 2 * np.eye(n)
    ex = sla.eigh(M, Bm, eigvals_only=True)
    A, B = make_A(M), make_A(Bm)
    tol = 1e-6
    for largest in (False, True):
        want = ex[::-1][:5] if largest else ex[:5]
        r = lobpcg(A, largest, rng.random((n, 2)), 5, B=B, tol=tol, maxiter=2000, rng=rng)
        assert r.X.shape == (n, 5) and len(r.iterations) == 3 and np.all(r.converged)
        assert np.max(np.linalg.norm(M @ r.X - Bm @ r.X * r.lam[None, :], axis=0)) <= 4 * tol
        assert np.allclose(r.X.T @ Bm @ r.X, np.eye(5), atol=2 * n * tol)
        assert np.allclose(np.sort(r.lam), np.sort(want), atol=1e-5 * np.abs(ex).max())


def case_powm(oracle, run, dtype):
    """powm! / invpowm! (reference src/simple.jl; test/simple_eigensolvers.jl:14-50, real element types): the engine against
    the oracle's restatement (iterations, Rayleigh quotient, residual history, x) and the reference's own assertions.
    run(A, x0, tol, maxiter) -> (theta, x, outcome with iters, converged, hist); A: scipy matrix the engine applies
    (for inverse iteration the explicit inverse of A - sigma I, standing for the reference's LU LinearMap)."""
    rng = np.random.default_rng(SEED)
    n = 10
    A = rng.random((n, n)).astype(dtype) + np.eye(n, dtype=dtype)
    A = (A.T @ A).astype(dtype)
    ls = np.linalg.eigvalsh(A.astype(np.float64))
    tol = n ** 2 * np.linalg.cond(A.astype(np.float64)) * float(np.finfo(dtype).eps)
    x0 = rng.random(n).astype(dtype)
    x0 /= np.linalg.norm(x0)
    lam, x, h = oracle.powm_(A, x0.copy(), tol=tol, maxiter=10 * n, log=True)
    assert h.isconverged and abs(lam - ls[-1]) <= 1e-5 * ls[-1] and np.linalg.norm(A @ x - lam * x) <= tol   # :27-28
    th, xs, hs = run(sp.csr_matrix(A), x0.copy(), tol, 10 * n)
    eps = float(np.finfo(dtype).eps)
    assert hs.iters == h.iters and hs.converged and abs(th - lam) <= 50 * eps * abs(lam)
    assert np.abs(xs - x).max() <= 50 * eps and np.max(np.abs(hs.hist[: h.iters] - h["resnorm"])) <= 200 * eps * ls[-1]
    idx = n // 2                                                         # inverse iteration near a middle eigenvalue :32-48
    sigma = dtype(0.75 * ls[idx - 1] + 0.25 * ls[idx])
    Finv = np.linalg.inv(A.astype(np.float64) - float(sigma) * np.eye(n))
    lam2, x2, h2 = oracle.invpowm_(lambda v: (Finv @ v.astype(np.float64)).astype(dtype), x0.copy(), shift=float(sigma),
                                   tol=tol, maxiter=10 * n, log=True)
    assert h2.isconverged and abs(lam2 - ls[idx - 1]) <= 1e-4 * ls[idx - 1]
    th2, xs2, hs2 = run(sp.csr_matrix(Finv.astype(dtype)), x0.copy(), tol, 10 * n)
    lam_e = float(sigma) + 1.0 / th2
    assert abs(hs2.iters - h2.iters) <= 1 and hs2.converged and abs(lam_e - ls[idx - 1]) <= 1e-4 * ls[idx - 1]
    assert np.linalg.norm(A.astype(np.float64) @ xs2 - lam_e * xs2) <= 10 * tol * max(1.0, abs(lam_e))
    # maxiter: done() tests `iteration > maxiter` -> maxiter + 1 steps
    lam3, x3, h3 = oracle.powm_(A, x0.copy(), tol=0.0, maxiter=3, log=True)
    th3, xs3, hs3 = run(sp.csr_matrix(A), x0.copy(), 0.0, 3)
    assert h3.iters == hs3.iters == 4 and not hs3.converged and np.abs(xs3 - x3).max() <= 50 * eps


def case_stationary(oracle, run, dtype, exact):
    """jacobi! / gauss_seidel! / sor! / ssor! for sparse matrices (reference src/stationary_sparse.jl, test/stationary.jl:17-95)
    against the oracle's column-by-column restatement: the reference's assertions (residual after 2n sweeps of a diagonally
    dominant system, SOR(1) == Gauss-Seidel, SingularException for a zero diagonal) and the iterates themselves -- bit for
    bit when `exact` (serial backend: same arithmetic in the same order), to a few ulps otherwise (the GPU contracts a*b+c).
    run(method, x0, A_scipy, b, omega, maxiter) -> x (raises np.linalg.LinAlgError for a singular diagonal)."""
    rng = np.random.default_rng(1234322)
    n, omega = 10, 1.2
    eps = float(np.finfo(dtype).eps)
    A = (sp.random(n, n, 4 / n, random_state=7, format="csc") + 2 * n * sp.eye(n)).tocsc().astype(dtype)
    b, x0 = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
    tol = math.sqrt(eps)
    fo = dict(jacobi=oracle.jacobi_, gauss_seidel=oracle.gauss_seidel_, sor=oracle.sor_, ssor=oracle.ssor_)

    def same(a, c, sweeps):
        return np.array_equal(a, c) if exact else np.abs(a - c).max() <= 8 * sweeps * eps * np.abs(c).max()

    for name in fo:
        args = (omega,) if name in ("sor", "ssor") else ()
        for start in (np.zeros(n, dtype), x0):                           # jacobi(A, b) and jacobi!(copy(x0), A, b)
            for mi in (1, 2, 7, 2 * n):
                xo = fo[name](start.copy(), A, b, *args, maxiter=mi)
                xs = run(name, start.copy(), A, b, omega, mi)
                assert same(xs, xo, mi), (name, mi)
            assert np.linalg.norm(b - A @ xs) / np.linalg.norm(b) <= tol, name                 # test/stationary.jl:33-54
    A2 = (sp.random(10, 10, 0.4, random_state=11, format="csc") + 4 * sp.eye(10)).tocsc().astype(dtype)   # :60-73
    b2 = (A2 @ np.ones(10)).astype(dtype)
    for mi in range(1, 6):
        g = run("gauss_seidel", np.zeros(10, dtype), A2, b2, 1.0, mi)
        r = run("sor", np.zeros(10, dtype), A2, b2, 1.0, mi)
        assert np.allclose(g, r, rtol=50 * eps, atol=0)
    Z = sp.csc_matrix(np.array([[0.0, 1.0], [1.0, 0.0]], dtype=dtype))                          # :75-90
    for name in fo:
        with pytest.raises(np.linalg.LinAlgError):
            run(name, np.zeros(2, dtype), Z, np.ones(2, dtype), omega, 3)
    # larger patterns: a 3-D Laplacian (wavefront levels) and a random non-symmetric pattern
    O = oracle.laplace_matrix(dtype, 6, 3)
    L3 = O.to_scipy()
    R = (sp.random(200, 200, 0.03, random_state=3, format="csc") + 10 * sp.eye(200)).tocsc().astype(dtype)
    for M, w in ((L3, 1.5), (R, 0.8)):
        m = M.shape[0]
        bb, xx = rng.random(m).astype(dtype), rng.random(m).astype(dtype)
        for name in fo:
            args = (w,) if name in ("sor", "ssor") else ()
            assert same(run(name, xx.copy(), M, bb, w, 5), fo[name](xx.copy(), M, bb, *args, maxiter=5), 5), name
