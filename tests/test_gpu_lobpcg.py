"""GPU parity tests for lobpcg (reference src/lobpcg.jl) through the C ABI against the CPU oracle.

Eigenvectors are defined up to sign (and up to rotation inside degenerate clusters of the Laplacian),
so parity is stated on: Ritz values per iteration horizon, residual norms, ||A X - X L||, X'X = I and
the analytic spectrum.  Tolerances: fp64 Ritz values 1e-9 relative at a fixed short horizon (before
rounding differences are amplified by the Rayleigh-Ritz selection), 1e-6 at convergence; fp32 1e-4.
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
SEED = 1234321


@pytest.fixture(scope="module")
def isb():
    import iterativesolvers_jl_b200 as m
    m.default_context()
    return m


def lap_eigs(N, dims, k, largest=False):
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    lam = lam1
    for _ in range(dims - 1):
        lam = (lam[:, None] + lam1[None, :]).ravel()
    lam = np.sort(lam)
    return lam[::-1][:k] if largest else lam[:k]


def _op(isb, O):
    return isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=O.base)


@pytest.mark.parametrize("bs", [1, 2, 5, 16])
@pytest.mark.parametrize("largest", [False, True])
def test_lobpcg_fixed_horizon_matches_oracle_fp64(isb, oracle, bs, largest):
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float64, 14, 3)
    A = _op(isb, O)
    X0 = rng.random((O.n, bs))
    ro = oracle.lobpcg(O, largest, X0, maxiter=6, fixed_iterations=True)
    r = isb.lobpcg(A, largest, X0, maxiter=6, _fixed_iterations=True)
    assert r.iterations == ro.iterations == 7
    np.testing.assert_allclose(r.lam, ro.lam, rtol=1e-9)
    np.testing.assert_allclose(r.residual_norms, ro.residual_norms, rtol=1e-6)
    S = O.to_scipy()
    R = S @ r.X - r.X * r.lam[None, :]
    np.testing.assert_allclose(np.linalg.norm(R, axis=0), r.residual_norms, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(r.X.T @ r.X, np.eye(bs), atol=1e-10)
    # same subspace as the oracle: principal angles
    sv = np.linalg.svd(ro.X.T @ r.X, compute_uv=False)
    assert sv.min() > 1 - 1e-8


def test_lobpcg_converges_to_analytic_spectrum_fp64(isb, oracle):
    """reference test/lobpcg.jl:72-84 shape: sparse Laplacian, ||A X - X L|| <= tol, analytic eigenvalues."""
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float64, 20, 2)
    A = _op(isb, O)
    S = O.to_scipy()
    tol = np.finfo(np.float64).eps ** 0.3
    for largest in (False, True):
        X0 = rng.random((O.n, 4))
        r = isb.lobpcg(A, largest, X0, tol=tol, maxiter=600)
        ro = oracle.lobpcg(O, largest, X0, tol=tol, maxiter=600)
        assert r.converged and ro.converged
        assert abs(r.iterations - ro.iterations) <= max(3, ro.iterations // 10)
        assert np.all(np.linalg.norm(S @ r.X - r.X * r.lam[None, :], axis=0) <= tol * 1.001)
        np.testing.assert_allclose(r.lam, lap_eigs(20, 2, 4, largest), rtol=1e-6)
        np.testing.assert_allclose(r.lam, ro.lam, rtol=1e-6)
        assert r.tolerance == pytest.approx(tol) and r.maxiter == 600


def test_lobpcg_soft_locking_jacobi_and_exact_start(isb, oracle):
    rng = np.random.default_rng(SEED)
    L = oracle.laplace_matrix_scipy(np.float64, 12, 2)
    M = (L + sp.diags(np.linspace(0.0, 5.0, L.shape[0]))).tocsc()
    O = oracle.CSC.from_scipy(M)
    A = isb.B200CSR.from_scipy(M)
    X0 = rng.random((O.n, 3))
    tol = 1e-6
    r0 = isb.lobpcg(A, False, X0, tol=tol, maxiter=400)
    r1 = isb.lobpcg(A, False, X0, P=isb.JacobiPrec(A.diag()), tol=tol, maxiter=400)     # test/lobpcg.jl:182-212
    ro = oracle.lobpcg(O, False, X0, P=oracle.JacobiPrec(O.diagonal()), tol=tol, maxiter=400)
    assert r0.converged and r1.converged
    np.testing.assert_allclose(r0.lam, r1.lam, rtol=1e-6)
    np.testing.assert_allclose(r1.lam, ro.lam, rtol=1e-6)
    assert r1.iterations <= r0.iterations and abs(r1.iterations - ro.iterations) <= 3
    w, V = np.linalg.eigh(M.toarray())
    re = isb.lobpcg(A, False, V[:, :2].copy(), tol=1e-8)                                # test/lobpcg.jl:47-48
    assert re.converged and re.iterations == 1
    np.testing.assert_allclose(re.lam, w[:2], rtol=1e-12)


def test_lobpcg_errors(isb):
    A = isb.B200CSR.from_scipy(sp.identity(5, format="csc"))
    with pytest.raises(isb.B200Error):                                                  # src/lobpcg.jl:834
        isb.lobpcg(A, False, np.ones((5, 2)))


def test_lobpcg_fp32_config5_shape(isb, oracle):
    """BASELINE.json configs[4] at oracle size: laplace_matrix(Float32, 24, 3), block 16, fp32."""
    rng = np.random.default_rng(SEED)
    O = oracle.laplace_matrix(np.float32, 24, 3)
    A = _op(isb, O)
    X0 = rng.random((O.n, 16)).astype(np.float32)
    ro = oracle.lobpcg(O, False, X0, maxiter=8, fixed_iterations=True)
    r = isb.lobpcg(A, False, X0, maxiter=8, _fixed_iterations=True)
    assert r.iterations == ro.iterations == 9 and r.lam.dtype == np.float32
    np.testing.assert_allclose(r.lam, ro.lam, rtol=1e-4)
    np.testing.assert_allclose(r.residual_norms, ro.residual_norms, rtol=5e-3)
    ex = lap_eigs(24, 3, 16)
    assert np.all(r.lam >= ex * (1 - 1e-4))
    # natural run (soft locking, tol = eps32^0.3): converges, eigenvalues to 1e-3 of analytic
    rn = isb.lobpcg(A, False, X0, maxiter=300)
    assert rn.converged
    np.testing.assert_allclose(rn.lam, ex, rtol=2e-3)


@pytest.mark.parametrize("bs", [16, 5])
def test_lobpcg_fp32_tensor_pipe_vs_simt(isb, oracle, bs):
    """fp32 blocks run the update and the Rayleigh-Ritz Gram products as 3xTF32 MMAs by default; option
    lobpcg_mma = 0 selects the SIMT kernels.  Both must agree with each other and with the oracle to fp32
    accuracy.  n = 24^3 is a multiple of the 128-row chunk and of the 16-row MMA tile, n = 23^3 is neither."""
    import ctypes as C
    L = isb.lib()
    ctx = isb.default_context()
    rng = np.random.default_rng(SEED + bs)
    for N in (24, 23):
        O = oracle.laplace_matrix(np.float32, N, 3)
        A = _op(isb, O)
        X0 = rng.random((O.n, bs)).astype(np.float32)
        ro = oracle.lobpcg(O, False, X0, maxiter=6, fixed_iterations=True)
        out = {}
        try:
            for mode in (1, 2, 0):       # 1: tcgen05 Gram (TMEM accumulators), 2: legacy mma.sync Gram, 0: SIMT
                assert L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", mode) == 0
                out[mode] = isb.lobpcg(A, False, X0, maxiter=6, _fixed_iterations=True)
        finally:
            L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", 1)
        for mode in (1, 2, 0):
            np.testing.assert_allclose(out[mode].lam, ro.lam, rtol=1e-4)
            np.testing.assert_allclose(out[mode].residual_norms, ro.residual_norms, rtol=5e-3)
        np.testing.assert_allclose(out[1].lam, out[0].lam, rtol=2e-5)
        np.testing.assert_allclose(out[1].lam, out[2].lam, rtol=2e-5)
        np.testing.assert_allclose(out[1].X.T @ out[1].X, np.eye(bs), atol=2e-5)


PRODUCTS = [(0, 2), (0, 1), (1, 2), (0, 4), (0, 3), (1, 3), (2, 3), (3, 4)]     # X'AR X'R R'AR X'AP X'P R'P AR'P P'AP


def _gram_rr(isb, blocks, variant):
    """the engine's Rayleigh-Ritz Gram kernel on five row-major n x 16 fp32 blocks (C-ABI test hook)."""
    import ctypes as C
    ctx = isb.default_context()
    n = blocks[0].shape[0]
    devs = [isb.DeviceArray.from_numpy(ctx, np.ascontiguousarray(b, dtype=np.float32).reshape(-1)) for b in blocks]
    ptrs = (C.c_void_p * 5)(*[d.ptr for d in devs])
    out = np.zeros(8 * 256)
    isb._lib.check(isb.lib().b200_debug_lobpcg_gram_rr(ctx._h, ptrs, n, variant, out.ctypes.data_as(C.c_void_p)))
    return out.reshape(8, 16, 16)


@pytest.mark.parametrize("n", [8, 64, 200, 64 * 148 * 9 + 37])
def test_gram_rr_tcgen05_kernel_vs_fp64(isb, n):
    """k_gram_umma (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators) against fp64 numpy Gram products and against
    the legacy mma.sync kernel: one MMA (n = 8), one stage, a ragged tail, and several accumulator hand-overs per CTA.
    fp32-level accuracy: 2e-6 of ||a_i|| ||b_j||."""
    rng = np.random.default_rng(SEED + n)
    blocks = [rng.standard_normal((n, 16)).astype(np.float32) * (1.0 + b) for b in range(5)]
    got = {v: _gram_rr(isb, blocks, v) for v in (1, 2)}
    B64 = [b.astype(np.float64) for b in blocks]
    for p, (l, r) in enumerate(PRODUCTS):
        ref = B64[l].T @ B64[r]
        scale = np.linalg.norm(B64[l], axis=0)[:, None] * np.linalg.norm(B64[r], axis=0)[None, :]
        for v in (1, 2):
            err = float(np.max(np.abs(got[v][p] - ref) / scale))
            assert err <= 2e-6, (n, p, v, err, got[v][p][:2, :4], ref[:2, :4])


def test_lobpcg_fp32_tcgen05_gram_many_stages(isb, oracle):
    """n = 64^3: every CTA of the tcgen05 Gram kernel runs ~28 stages of 64 rows, i.e. several accumulator hand-overs
    between the two TMEM accumulators (lobpcg_gram_umma.cuh: kUmDrain = 8 stages per group).  Against the oracle and
    against the legacy mma.sync Gram kernel, 5 steps."""
    L = isb.lib()
    ctx = isb.default_context()
    N, bs = 64, 16
    O = oracle.laplace_matrix(np.float32, N, 3)
    A = _op(isb, O)
    X0 = np.random.default_rng(SEED).random((O.n, bs)).astype(np.float32)
    ro = oracle.lobpcg(O, False, X0, maxiter=5, fixed_iterations=True)
    out = {}
    try:
        for mode in (1, 2):
            assert L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", mode) == 0
            out[mode] = isb.lobpcg(A, False, X0, maxiter=5, _fixed_iterations=True)
    finally:
        L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", 1)
    for mode in (1, 2):
        np.testing.assert_allclose(out[mode].lam, ro.lam, rtol=1e-4)
    np.testing.assert_allclose(out[1].lam, out[2].lam, rtol=2e-5)


def test_lobpcg_256cubed_fp32_properties(isb):
    """configs[4] full size: lobpcg block=16 on laplace_matrix(Float32, 256, 3), 4 steps: Ritz values
    decrease monotonically (Rayleigh-Ritz optimality), stay above the analytic lambda_min, X'X = I."""
    import ctypes as C
    ctx = isb.default_context()
    N, bs = 256, 16
    n = N ** 3
    A = isb.B200CSR.laplacian(N, 3, np.float32)
    rng = np.random.default_rng(SEED)
    X0 = rng.random((n, bs), dtype=np.float32)
    prev = None
    for steps in (2, 4):
        Xd = isb.DeviceArray.from_numpy(ctx, X0)
        r = isb.lobpcg(A, False, Xd, maxiter=steps, _fixed_iterations=True)
        assert r.iterations == steps + 1
        lam = np.sort(r.lam.astype(np.float64))
        assert lam[0] >= lap_eigs(N, 3, 1)[0] * (1 - 1e-3)
        if prev is not None:
            assert np.all(lam <= prev * (1 + 1e-5))
        prev = lam
        X = Xd.numpy()
        G = X[:: 64].T.astype(np.float64) @ X[:: 64].astype(np.float64)   # sampled rows: cheap sanity on scale
        assert np.all(np.isfinite(G))
        Xd.free()
