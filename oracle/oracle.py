"""
oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle).

numpy restatement of the reference's CPU algorithms for the Krylov hot path.  Every function cites
the reference lines (relative to /root/reference) that it follows one-to-one.  SpMV on
SparseMatrixCSC goes through the plain-C column-scatter loop in oracle.c (what Julia's SparseArrays
stdlib executes); dense BLAS-2/3 pieces go through numpy (OpenBLAS, the same backend family Julia
calls through libblastrampoline).

PARITY UNPINNED for raw SpMV/dot/norm values: the reference holds no golden vectors for them
 and cannot be executed in this image (no Julia).  What IS pinned:
  * the literal Hessenberg fixtures of reference test/hessenberg.jl:10-26 (tests/golden/),
  * the reference's own property tests, ported in tests/test_oracle_*.py,
  * analytic known answers (Laplacian spectrum, manufactured solutions), and scipy cross-checks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_laplace_nnz.restype = ctypes.c_int64
        _LIB.oracle_laplace_nnz.argtypes = [ctypes.c_int64, ctypes.c_int32]
        _LIB.oracle_laplace_csc_f64.restype = ctypes.c_int64
        _LIB.oracle_num_threads.restype = ctypes.c_int32
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------
# SparseMatrixCSC stand-in
# --------------------------------------------------------------------------------------------
@dataclass
class CSC:
    """SparseMatrixCSC{Tv,Int64}: colptr (n+1), rowval (nnz), nzval (nnz); `base` = 1 for Julia
    indexing, 0 for C indexing.  Field names as in reference src/stationary_sparse.jl:14-18."""
    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray
    base: int = 0

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def dtype(self):
        return self.nzval.dtype

    @property
    def nnz(self):
        return int(self.colptr[-1] - self.base)

    @staticmethod
    def from_scipy(A, base=0):
        A = A.tocsc()
        A.sort_indices()
        return CSC(A.shape[0], A.shape[1], A.indptr.astype(np.int64) + base,
                   A.indices.astype(np.int64) + base, np.ascontiguousarray(A.data), base)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - self.base, self.colptr - self.base),
                             shape=(self.m, self.n))

    def diagonal(self):
        return self.to_scipy().diagonal()


def csc_spmv(A: CSC, x: np.ndarray, y: np.ndarray | None = None) -> np.ndarray:
    """mul!(y, A::SparseMatrixCSC, x) -- SparseArrays stdlib, serial column scatter."""
    x = np.ascontiguousarray(x, dtype=A.dtype)
    if y is None:
        y = np.empty(A.m, dtype=A.dtype)
    fn = lib().oracle_csc_spmv_f64 if A.dtype == np.float64 else lib().oracle_csc_spmv_f32
    fn(ctypes.c_int64(A.m), ctypes.c_int64(A.n), _p(A.colptr), _p(A.rowval), _p(A.nzval),
       ctypes.c_int64(A.base), _p(x), _p(y))
    return y


def csc_spmv_adjoint(A: CSC, x: np.ndarray, y: np.ndarray | None = None) -> np.ndarray:
    """mul!(y, adjoint(A), x) for A::SparseMatrixCSC -- SparseArrays stdlib, one gather dot per column
    (call sites: reference src/qmr.jl:76, src/lsqr.jl:132,172, src/lsmr.jl:118,172)."""
    x = np.ascontiguousarray(x, dtype=A.dtype)
    if y is None:
        y = np.empty(A.n, dtype=A.dtype)
    fn = lib().oracle_csc_spmv_adj_f64 if A.dtype == np.float64 else lib().oracle_csc_spmv_adj_f32
    fn(ctypes.c_int64(A.m), ctypes.c_int64(A.n), _p(A.colptr), _p(A.rowval), _p(A.nzval),
       ctypes.c_int64(A.base), _p(x), _p(y))
    return y


def csc_spmm(A: CSC, X: np.ndarray, Y: np.ndarray | None = None) -> np.ndarray:
    """mul!(Y, A, X) on column-major n x bs blocks (Fortran-ordered numpy arrays)."""
    X = np.asfortranarray(X, dtype=A.dtype)
    if Y is None:
        Y = np.empty((A.m, X.shape[1]), dtype=A.dtype, order="F")
    fn = lib().oracle_csc_spmm_f64 if A.dtype == np.float64 else lib().oracle_csc_spmm_f32
    fn(ctypes.c_int64(A.m), ctypes.c_int64(A.n), _p(A.colptr), _p(A.rowval), _p(A.nzval),
       ctypes.c_int64(A.base), _p(X), ctypes.c_int64(X.shape[0]), _p(Y), ctypes.c_int64(Y.shape[0]),
       ctypes.c_int64(X.shape[1]))
    return Y


def mul(A, x):
    """A*x for the operator kinds the reference tests use: SparseMatrixCSC, dense Matrix,
    LinearMap-like callables (test/cg.jl:71-77)."""
    if isinstance(A, CSC):
        return csc_spmm(A, x) if x.ndim == 2 else csc_spmv(A, x)
    if callable(A):
        return A(x)
    return A @ x


def mul_adjoint(A, x):
    """adjoint(A)*x for SparseMatrixCSC and dense Matrix operators."""
    if isinstance(A, CSC):
        return csc_spmv_adjoint(A, x)
    return A.conj().T @ x


def opsize(A, d=None):
    s = A.shape
    return s if d is None else s[d]


# --------------------------------------------------------------------------------------------
# Generators
# --------------------------------------------------------------------------------------------
def second_order_central_diff(T, dim):
    """reference test/laplace_matrix.jl:12 -- SymTridiagonal(fill(2,dim), fill(-1,dim-1))."""
    import scipy.sparse as sp
    return sp.diags([np.full(dim - 1, -1, dtype=T), np.full(dim, 2, dtype=T), np.full(dim - 1, -1, dtype=T)],
                    [-1, 0, 1], format="csc", dtype=T)


def laplace_matrix_scipy(T, n, dims):
    """reference test/laplace_matrix.jl:1-10, literal kron recursion (small sizes)."""
    import scipy.sparse as sp
    D = second_order_central_diff(T, n)
    A = D.copy()
    for _ in range(2, dims + 1):
        A = sp.kron(A, sp.identity(n, dtype=T, format="csc"), format="csc") + \
            sp.kron(sp.identity(A.shape[0], dtype=T, format="csc"), D, format="csc")
    A = A.tocsc()
    A.sort_indices()
    return A


def laplace_matrix(T, n, dims, base=0) -> CSC:
    """laplace_matrix(T, n, dims) as a CSC{T,Int64}; direct C construction (same matrix as the
    kron recursion, checked against laplace_matrix_scipy in tests)."""
    T = np.dtype(T)
    N = int(n) ** dims
    nnz = lib().oracle_laplace_nnz(n, dims)
    colptr = np.empty(N + 1, dtype=np.int64)
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz, dtype=np.float64)
    got = lib().oracle_laplace_csc_f64(ctypes.c_int64(n), ctypes.c_int32(dims), ctypes.c_int64(base),
                                       _p(colptr), _p(rowval), _p(nzval))
    assert got == nnz
    return CSC(N, N, colptr, rowval, nzval.astype(T, copy=False), base)


def advection_dominated(N=50, beta=1000.0):
    """reference benchmark/advection_diffusion.jl:3-30.  A = laplace/(-h^2) + kron(I_{N^2}, dx),
    dx = tridiag(-beta/2h, 0, +beta/2h), b = f(x,y,z) on interior points, x fastest.
    Returns (scipy csc, b)."""
    import scipy.sparse as sp
    n = N ** 3
    h = 1.0 / (N + 1)
    # xs = range(0, stop=1, length=N+2)[2:N+1]; Julia's twice-precision range gives i/(N+1)
    xs = np.arange(1, N + 1, dtype=np.float64) / (N + 1)
    lap = laplace_matrix_scipy(np.float64, N, 3)
    lap = lap.copy()
    lap.data = lap.data / -(h ** 2)                     # ./ -h^2
    lo = np.full(N - 1, -beta / (2 * h))                # -1 => fill(-beta / 2h, N-1)
    up = np.full(N - 1, beta / (2 * h))                 # +1 => fill( beta / 2h, N-1)
    dx1 = sp.diags([lo, up], [-1, 1], format="csc")
    dx = sp.kron(sp.identity(N * N, format="csc"), dx1, format="csc")
    A = (lap + dx).tocsc()
    A.sort_indices()
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")   # [f(x,y,z) for x, y, z] column-major: x fastest
    F = np.exp(X * Y * Z) * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    b = F.reshape(n, order="F")
    assert A.shape == (n, n)
    return A, b


# --------------------------------------------------------------------------------------------
# ConvergenceHistory (reference src/history.jl:54-66, 127-252) -- counts only
# --------------------------------------------------------------------------------------------
@dataclass
class ConvergenceHistory:
    mvps: int = 0
    mtvps: int = 0
    iters: int = 0
    restart: int | None = None
    isconverged: bool = False
    data: dict = field(default_factory=dict)

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, k, v):
        self.data[k] = v

    @property
    def niters(self):
        return self.iters

    @property
    def nprods(self):
        return self.mvps + self.mtvps

    @property
    def nrests(self):
        """src/history.jl:252: ceil(iters/restart)."""
        return int(math.ceil(self.iters / self.restart))


class Identity:
    """reference src/common.jl:28-32."""

    def ldiv(self, x):           # ldiv!(::Identity, x) = x
        return x

    def ldiv3(self, y, x):       # ldiv!(y, ::Identity, x) = copyto!(y, x)
        y[...] = x
        return y


class JacobiPrec:
    """reference test/cg.jl:10-18: ldiv!(y, P, x) = y .= x ./ P.diagonal."""

    def __init__(self, diagonal):
        self.diagonal = np.asarray(diagonal)

    def ldiv(self, x):
        x /= self.diagonal
        return x

    def ldiv3(self, y, x):
        np.divide(x, self.diagonal, out=y)
        return y


class MatrixPrec:
    """Factorization-like preconditioner P \\ x with a dense matrix (test/gmres.jl:19,28,33 use lu(A))."""

    def __init__(self, M):
        import scipy.linalg as sla
        self.lu = sla.lu_factor(np.asarray(M))

    def ldiv(self, x):
        import scipy.linalg as sla
        x[...] = sla.lu_solve(self.lu, x)
        return x

    def ldiv3(self, y, x):
        import scipy.linalg as sla
        y[...] = sla.lu_solve(self.lu, x)
        return y


def _eps(dtype):
    return np.finfo(np.dtype(dtype)).eps


def _real_dtype(dtype):
    return np.zeros(1, dtype=dtype).real.dtype


# --------------------------------------------------------------------------------------------
# CG (reference src/cg.jl)
# --------------------------------------------------------------------------------------------
def cg_(x, A, b, *, abstol=0.0, reltol=None, maxiter=None, log=False, Pl=None, initially_zero=False):
    """cg!(x, A, b; ...) -- reference src/cg.jl:209-242 driving cg_iterator! :120-155 and
    iterate :43-66 (Identity) / :72-100 (PCG)."""
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))
    if maxiter is None:
        maxiter = opsize(A, 1)
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    resnorm = []
    # cg_iterator!
    u = np.zeros_like(x)                                   # :129
    r = b.astype(x.dtype, copy=True)                       # :130
    if initially_zero:
        mv_products = 0
    else:
        mv_products = 1
        c = mul(A, x)                                      # :137
        r -= c                                             # :138
    c = np.empty_like(x)
    residual = float(np.linalg.norm(r))                    # :140
    tol = max(reltol * residual, abstol)                   # :141
    identity = Pl is None or isinstance(Pl, Identity)
    prev_residual = 1.0                                    # :146
    rho = 1.0                                              # :151
    if log:
        history.mvps = mv_products                         # :226-228
    iteration = 0
    while True:
        if iteration >= maxiter or residual <= tol:        # :36
            break
        if identity:
            beta = residual ** 2 / prev_residual ** 2      # :50
            u[...] = r + beta * u                          # :51
            c = mul(A, u)                                  # :54
            alpha = residual ** 2 / float(np.dot(u, c))    # :55
            x += alpha * u                                 # :58
            r -= alpha * c                                 # :59
            prev_residual = residual                       # :61
            residual = float(np.linalg.norm(r))            # :62
        else:
            c = Pl.ldiv3(np.empty_like(r), r)              # :79
            rho_prev = rho
            rho = float(np.dot(c, r))                      # :82
            beta = rho / rho_prev                          # :85
            u[...] = c + beta * u                          # :86
            c = mul(A, u)                                  # :89
            alpha = rho / float(np.dot(u, c))              # :90
            x += alpha * u                                 # :93
            r -= alpha * c                                 # :94
            residual = float(np.linalg.norm(r))            # :96
        iteration += 1
        if log:
            history.iters += 1                             # nextiter!(history, mvps=1) :231
            history.mvps += 1
            resnorm.append(residual)                       # :232
    if log:
        history.isconverged = residual <= tol              # :238
        history["resnorm"] = np.array(resnorm)
        history["tol"] = tol
        return x, history
    return x


def cg(A, b, **kw):
    """cg(A, b; kw...) = cg!(zerox(A, b), A, b; initially_zero = true, kw...)  src/cg.jl:162."""
    x = np.zeros(opsize(A, 1), dtype=np.result_type(b.dtype, np.float32))
    return cg_(x, A, b, initially_zero=True, **kw)


def cg_csc_c(x, A: CSC, b, *, abstol=0.0, reltol=None, maxiter=None, Pl_diag=None, initially_zero=False):
    """Same algorithm, entirely inside oracle.c (fast path for bigger parity cases)."""
    if reltol is None:
        reltol = math.sqrt(_eps(np.float64))
    if maxiter is None:
        maxiter = A.n

    class Res(ctypes.Structure):
        _fields_ = [("iters", ctypes.c_int64), ("mvps", ctypes.c_int64), ("isconverged", ctypes.c_int32),
                    ("pad", ctypes.c_int32), ("tol", ctypes.c_double), ("residual", ctypes.c_double)]

    res = Res()
    resnorm = np.zeros(maxiter + 1, dtype=np.float64)      # reserve!(history, :resnorm, maxiter+1) :221
    b = np.ascontiguousarray(b, dtype=np.float64)
    diag = None if Pl_diag is None else np.ascontiguousarray(Pl_diag, dtype=np.float64)
    lib().oracle_cg_f64(ctypes.c_int64(A.n), _p(A.colptr), _p(A.rowval), _p(A.nzval), ctypes.c_int64(A.base),
                        _p(x), _p(b), ctypes.c_double(abstol), ctypes.c_double(reltol), ctypes.c_int64(maxiter),
                        ctypes.c_int32(1 if initially_zero else 0),
                        _p(diag) if diag is not None else ctypes.c_void_p(0), _p(resnorm), ctypes.byref(res))
    h = ConvergenceHistory(mvps=res.mvps, iters=res.iters, isconverged=bool(res.isconverged))
    h["abstol"], h["reltol"], h["tol"] = abstol, reltol, res.tol
    h["resnorm"] = resnorm[: res.iters].copy()             # shrink! :239
    return x, h


# --------------------------------------------------------------------------------------------
# Chebyshev iteration (reference src/chebyshev.jl) -- SURVEY.md section 8(f) item 2
# --------------------------------------------------------------------------------------------
def chebyshev_(x, A, b, lmin, lmax, *, abstol=0.0, reltol=None, Pl=None, maxiter=None, log=False,
               initially_zero=False):
    """chebyshev!(x, A, b, λmin, λmax; ...) -- reference src/chebyshev.jl:131-160, iterate :29-57,
    chebyshev_iterable! :59-92.  Restated literally, including `iteration == 1` being the SECOND call
    (start = 0, :26) and `u .= c .+ β .* c` (:45)."""
    T = x.dtype
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))
    if maxiter is None:
        maxiter = opsize(A, 1)
    Pl = Pl or Identity()
    l_avg = (lmax + lmin) / 2                                           # :65
    l_diff = (lmax - lmin) / 2                                          # :66
    r = b.astype(T, copy=True)                                          # :70
    u = np.zeros_like(x)
    if initially_zero:
        mv_products = 0
    else:
        mv_products = 1
        r -= mul(A, x)                                                  # :80-81
    resnorm = float(np.linalg.norm(r))                                  # :83
    tol = max(reltol * resnorm, abstol)                                 # :84
    alpha = _real_dtype(T).type(0)                                      # :89
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    history.mvps = mv_products                                          # :148
    resnorms = []
    iteration = 0                                                       # start :26
    RT = _real_dtype(T).type
    while True:
        if iteration >= maxiter or resnorm <= tol:                      # :27
            break
        c = Pl.ldiv3(np.empty_like(r), r)                               # :37
        if iteration == 1:                                              # :39
            alpha = RT(2) / RT(l_avg)
            u[...] = c
        else:
            beta = (RT(l_diff) * alpha / 2) ** 2                        # :43
            alpha = RT(1) / (RT(l_avg) - beta)                          # :44
            u[...] = c + T.type(beta) * c                               # :45 (sic)
        c = mul(A, u)                                                   # :48
        mv_products += 1
        x += T.type(alpha) * u                                          # :51
        r -= T.type(alpha) * c                                          # :52
        resnorm = float(np.linalg.norm(r))                              # :54
        iteration += 1
        history.iters += 1                                              # :150-152
        history.mvps = mv_products
        resnorms.append(resnorm)
    history.isconverged = resnorm <= tol                                # :157
    history["resnorm"] = np.array(resnorms)
    history["tol"] = tol
    return (x, history) if log else x


def chebyshev(A, b, lmin, lmax, **kw):
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return chebyshev_(x, A, b, lmin, lmax, initially_zero=True, **kw)


# --------------------------------------------------------------------------------------------
# Givens (LinearAlgebra.givensAlgorithm, external stdlib) + FastHessenberg ldiv!
# --------------------------------------------------------------------------------------------
def givens_algorithm(f, g):
    """LinearAlgebra.givensAlgorithm(f, g) -> (c real, s, r) with [c s; -conj(s) c][f; g] = [r; 0].
    Restates LAPACK xLARTG as Julia's stdlib does (sign convention: c >= 0 when |f| > |g| for
    reals; version-dependent, see SURVEY.md section 8c -- all users below are invariant to it)."""
    if np.iscomplexobj(f) or np.iscomplexobj(g):
        f = complex(f)
        g = complex(g)
        if g == 0:
            return 1.0, 0j, f
        if f == 0:
            ag = abs(g)
            return 0.0, np.conj(g) / ag, ag
        f1, g1 = abs(f), abs(g)
        d = math.hypot(f1, g1)
        ph = f / f1
        return f1 / d, ph * np.conj(g) / d, ph * d
    f = float(f)
    g = float(g)
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, 1.0, g
    r = math.hypot(f, g)
    c, s = f / r, g / r
    if abs(f) > abs(g) and c < 0:
        c, s, r = -c, -s, -r
    return c, s, r


def hessenberg_ldiv(H, rhs):
    """ldiv!(H::FastHessenberg, rhs) -- reference src/hessenberg.jl:15-46.  H is (m+1) x m,
    mutated to upper triangular; rhs (m+1) mutated: rhs[:m] = LS solution, rhs[m] = signed residual."""
    width = H.shape[1]
    for i in range(width):                                              # :24
        c, s, _ = givens_algorithm(H[i, i], H[i + 1, i])                # :25
        H[i, i] = c * H[i, i] + s * H[i + 1, i]                         # :28
        for j in range(i + 1, width):                                   # :31
            tmp = -np.conj(s) * H[i, j] + c * H[i + 1, j]               # :32
            H[i, j] = c * H[i, j] + s * H[i + 1, j]                     # :33
            H[i + 1, j] = tmp                                           # :34
        tmp = -np.conj(s) * rhs[i] + c * rhs[i + 1]                     # :38
        rhs[i] = c * rhs[i] + s * rhs[i + 1]                            # :39
        rhs[i + 1] = tmp                                                # :40
    # UpperTriangular back-substitution  :44-45
    for i in range(width - 1, -1, -1):
        acc = rhs[i]
        for j in range(i + 1, width):
            acc -= H[i, j] * rhs[j]
        rhs[i] = acc / H[i, i]
    return rhs


# --------------------------------------------------------------------------------------------
# orthogonalize_and_normalize! (reference src/orthogonalize.jl)
# --------------------------------------------------------------------------------------------
def orthogonalize_and_normalize_(V, w, h, method="mgs"):
    """V: n x k (column views), w: n, h: k; all mutated in place.  Returns nrm.
    method in {"dgks" :13-39, "cgs" :41-51, "mgs" :67-79}."""
    if method in ("dgks", "cgs"):
        h[...] = V.conj().T @ w                                         # :15 / :43
        w -= V @ h                                                      # :16 / :44
        nrm = float(np.linalg.norm(w))                                  # :17 / :45
        if method == "dgks":
            eta = 1.0 / math.sqrt(2.0)                                  # :20
            projection_size = float(np.linalg.norm(h))                  # :22
            while nrm < eta * projection_size:                          # :26
                correction = V.conj().T @ w                             # :27
                projection_size = float(np.linalg.norm(correction))     # :28
                w -= V @ correction                                     # :30
                h += correction                                         # :31
                nrm = float(np.linalg.norm(w))                          # :32
    elif method == "mgs":
        for i in range(V.shape[1]):                                     # :69
            column = V[:, i]
            h[i] = np.vdot(column, w)                                   # :71
            w -= h[i] * column                                          # :72
        nrm = float(np.linalg.norm(w))                                  # :75
    else:
        raise ValueError(method)
    with np.errstate(divide="ignore", invalid="ignore"):
        w *= w.dtype.type(1.0) / w.dtype.type(nrm) if nrm != 0 else np.inf  # w .*= inv(nrm)
    return nrm


# --------------------------------------------------------------------------------------------
# Stationary methods for sparse matrices (reference src/stationary_sparse.jl), restated column by column as the reference
# sweeps its SparseMatrixCSC.  A: CSC (this module's class) or anything CSC.from_scipy accepts.
# --------------------------------------------------------------------------------------------
def _st_prepare(A):
    if not isinstance(A, CSC):
        import scipy.sparse as sp
        A = CSC.from_scipy(sp.csc_matrix(A))
    cp = A.colptr - A.base
    rv = A.rowval - A.base
    n = A.n
    diag = np.zeros(n, dtype=np.int64)                                  # DiagonalIndices :6-27
    for col in range(n):
        r1, r2 = cp[col], cp[col + 1]
        k = r1 + np.searchsorted(rv[r1:r2], col)
        if k >= r2 or rv[k] != col or A.nzval[k] == 0:
            raise np.linalg.LinAlgError(f"SingularException({col + 1})")   # :19
        diag[col] = k
    return A, cp, rv, A.nzval, diag, n


def _gs_multiply_upper(cp, rv, nz, diag, n, alpha, x, beta, y, z):     # gauss_seidel_multiply!(alpha, U, x, beta, y, z) :179-191
    for col in range(n):
        ax = alpha * x[col]
        for j in range(cp[col], diag[col]):
            z[rv[j]] += nz[j] * ax
        z[col] = beta * y[col]


def _gs_multiply_lower(cp, rv, nz, diag, n, alpha, x, beta, y, z):     # gauss_seidel_multiply!(alpha, L, x, beta, y, z) :197-209
    for col in range(n - 1, -1, -1):
        ax = alpha * x[col]
        z[col] = beta * y[col]
        for j in range(diag[col] + 1, cp[col + 1]):
            z[rv[j]] += nz[j] * ax


def _forward_sub(cp, rv, nz, diag, n, x, alpha=None, beta=None, y=None):   # forward_sub! :64-79 / :84-102
    for col in range(n):
        idx = diag[col]
        x[col] = x[col] / nz[idx] if alpha is None else alpha * x[col] / nz[idx] + beta * y[col]
        for i in range(idx + 1, cp[col + 1]):
            x[rv[i]] -= nz[i] * x[col]


def _backward_sub(cp, rv, nz, diag, n, x, alpha, beta, y):              # backward_sub!(alpha, U, x, beta, y) :126-143
    for col in range(n - 1, -1, -1):
        idx = diag[col]
        x[col] = alpha * x[col] / nz[idx] + beta * y[col]
        for i in range(cp[col], idx):
            x[rv[i]] -= nz[i] * x[col]


def jacobi_(x, A, b, *, maxiter=10):
    """jacobi!(x, A::SparseMatrixCSC, b; maxiter = 10) -- reference src/stationary_sparse.jl:203-237."""
    A, cp, rv, nz, diag, n = _st_prepare(A)
    T = x.dtype.type
    for _ in range(maxiter):
        nxt = b.astype(x.dtype, copy=True)                              # copyto!(j.next, j.b) :214
        for col in range(n):                                            # mul!(-one(T), O, x, one(T), next) :148-173
            ax = T(-1) * x[col]
            for j in range(cp[col], diag[col]):
                nxt[rv[j]] += nz[j] * ax
            for j in range(diag[col] + 1, cp[col + 1]):
                nxt[rv[j]] += nz[j] * ax
        x[...] = nxt / nz[diag]                                         # ldiv!(j.x, j.O.diag, j.next) :217
    return x


def gauss_seidel_(x, A, b, *, maxiter=10):
    """gauss_seidel!(x, A::SparseMatrixCSC, b; maxiter = 10) -- reference src/stationary_sparse.jl:247-284."""
    A, cp, rv, nz, diag, n = _st_prepare(A)
    T = x.dtype.type
    for _ in range(maxiter):
        _gs_multiply_upper(cp, rv, nz, diag, n, T(-1), x, T(1), b, x)   # :264
        _forward_sub(cp, rv, nz, diag, n, x)                            # :265
    return x


def sor_(x, A, b, omega, *, maxiter=10):
    """sor!(x, A::SparseMatrixCSC, b, omega; maxiter = 10) -- reference src/stationary_sparse.jl:293-348.  Returns the
    iterate (the reference returns iterable.x, which after an odd number of pointer swaps is not the caller's array)."""
    A, cp, rv, nz, diag, n = _st_prepare(A)
    T = x.dtype.type
    nxt = np.zeros_like(x)
    for _ in range(maxiter):
        _gs_multiply_upper(cp, rv, nz, diag, n, T(-1), x, T(1), b, nxt)                     # next = b - U x :312
        _forward_sub(cp, rv, nz, diag, n, nxt, T(omega), T(1) - T(omega), x)                # :315
        x, nxt = nxt, x                                                                     # :318
    return x


def ssor_(x, A, b, omega, *, maxiter=10):
    """ssor!(x, A::SparseMatrixCSC, b, omega; maxiter = 10) -- reference src/stationary_sparse.jl:358-424."""
    A, cp, rv, nz, diag, n = _st_prepare(A)
    T = x.dtype.type
    tmp = np.zeros_like(x)
    for _ in range(maxiter):
        _gs_multiply_upper(cp, rv, nz, diag, n, T(-1), x, T(1), b, tmp)                     # :394
        _forward_sub(cp, rv, nz, diag, n, tmp, T(omega), T(1) - T(omega), x)                # :397-400
        _gs_multiply_lower(cp, rv, nz, diag, n, T(-1), tmp, T(1), b, x)                     # :403
        _backward_sub(cp, rv, nz, diag, n, x, T(omega), T(1) - T(omega), tmp)               # :406
    return x


def stationary_dense_(kind, x, A, b, omega=1.0, *, maxiter=10):
    """jacobi! / gauss_seidel! / sor! / ssor! for an AbstractMatrix -- reference src/stationary.jl: DenseJacobiIterable :48-70,
    DenseGaussSeidelIterable :108-127, DenseSORIterable :167-186, DenseSSORIterable :227-258, restated loop by loop.
    kind in {"jacobi", "gauss_seidel", "sor", "ssor"}; A dense (n x n array); returns the iterate."""
    A = np.asarray(A)
    n = A.shape[0]
    T = x.dtype.type
    w = T(omega)
    if np.any(np.diag(A) == 0):
        raise np.linalg.LinAlgError("SingularException")                 # check_diag :6-16
    x = x.copy()
    tmp = np.zeros_like(x)
    for _ in range(maxiter):
        if kind == "jacobi":
            nxt = b.astype(x.dtype, copy=True)                           # :52
            for col in range(n):                                         # :55-63
                for row in range(col):
                    nxt[row] -= A[row, col] * x[col]
                for row in range(col + 1, n):
                    nxt[row] -= A[row, col] * x[col]
            for col in range(n):                                         # :66-68
                x[col] = nxt[col] / A[col, col]
            continue
        if kind == "gauss_seidel":
            for col in range(n):                                         # :113-119
                for row in range(col):
                    x[row] -= A[row, col] * x[col]
                x[col] = b[col]
            for col in range(n):                                         # :121-126
                x[col] /= A[col, col]
                for row in range(col + 1, n):
                    x[row] -= A[row, col] * x[col]
            continue
        for col in range(n):                                             # SOR :172-178 / SSOR :232-238
            for row in range(col):
                tmp[row] -= A[row, col] * x[col]
            tmp[col] = b[col]
        for col in range(n):                                             # :180-185 / :240-245
            x[col] += w * (tmp[col] / A[col, col] - x[col])
            for row in range(col + 1, n):
                tmp[row] -= A[row, col] * x[col]
        if kind == "ssor":
            for col in range(n - 1, -1, -1):                             # :247-252
                tmp[col] = b[col]
                for row in range(col + 1, n):
                    tmp[row] -= A[row, col] * x[col]
            for col in range(n - 1, -1, -1):                             # :254-259
                for row in range(col):
                    tmp[row] -= A[row, col] * x[col]
                x[col] += w * (tmp[col] / A[col, col] - x[col])
    return x


# --------------------------------------------------------------------------------------------
# Power method / inverse iteration (reference src/simple.jl)
# --------------------------------------------------------------------------------------------
def powm_(B, x, *, tol=None, maxiter=None, shift=0.0, inverse=False, log=False):
    """powm!(B, x; shift, inverse, tol, maxiter, log) -- reference src/simple.jl:118-151 with PowerMethodIterable :6-15,
    iterate :29-48 (done uses `iteration > maxiter`, :27: up to maxiter + 1 steps).  B: anything mul() accepts, or a
    callable y = B(x) (the LinearMap of shift-and-invert, :83-88).  Returns (lambda, x[, history]); x is updated in place."""
    T = x.dtype
    n1, n2 = (x.shape[0], x.shape[0]) if callable(B) else (opsize(B, 0), opsize(B, 1))
    if tol is None:
        tol = float(_eps(_real_dtype(T))) * n2 ** 3                      # :119
    if maxiter is None:
        maxiter = n1                                                    # :120
    apply = B if callable(B) else (lambda v: mul(B, v))
    history = ConvergenceHistory()
    history["tol"] = tol
    resnorms = []
    theta = T.type(0)
    residual = float(np.finfo(_real_dtype(T)).max)                      # floatmax :55
    iteration = 0
    while not (iteration > maxiter or residual <= tol):                 # done :27
        Ax = np.asarray(apply(np.ascontiguousarray(x)), dtype=T)        # :32
        theta = np.vdot(x, Ax)                                          # :35
        r = Ax - theta * x                                              # :38-39
        residual = float(np.linalg.norm(r))                             # :42
        x[...] = Ax                                                     # :45
        x *= T.type(1) / T.type(np.linalg.norm(x))                      # :46
        iteration += 1
        history.iters += 1                                              # nextiter!(history, mvps = 1) :133
        history.mvps += 1
        resnorms.append(residual)
    history.isconverged = residual <= tol                               # :137
    lam = shift + (1 / theta if inverse else theta)                     # transform_eigenvalue :51
    if log:
        history["resnorm"] = np.array(resnorms)
        return lam, x, history
    return lam, x


def invpowm_(B, x, **kw):
    """invpowm!(B, x0; shift, kwargs...) = powm!(B, x0; inverse = true, kwargs...) -- src/simple.jl:186."""
    return powm_(B, x, inverse=True, **kw)


# --------------------------------------------------------------------------------------------
# GMRES (reference src/gmres.jl)
# --------------------------------------------------------------------------------------------
def gmres_(x, A, b, *, Pl=None, Pr=None, abstol=0.0, reltol=None, restart=None, maxiter=None, log=False,
           initially_zero=False, orth_meth="mgs"):
    """gmres!(x, A, b; ...) -- reference src/gmres.jl:184-222 with gmres_iterable! :108-136,
    iterate :57-106, update_residual! :224-233, init! :235-255, solve_least_squares! :262-271,
    update_solution! :273-283, expand! :285-304."""
    T = x.dtype
    n = opsize(A, 0)
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))
    if restart is None:
        restart = min(20, opsize(A, 1))
    if maxiter is None:
        maxiter = opsize(A, 1)
    Pl = Pl or Identity()
    Pr = Pr or Identity()
    history = ConvergenceHistory(restart=restart)
    history["abstol"], history["reltol"] = abstol, reltol
    resnorm = []

    V = np.zeros((n, restart + 1), dtype=T, order="F")                  # ArnoldiDecomp :11-15
    H = np.zeros((restart + 1, restart), dtype=T, order="F")
    nullvec = np.ones(restart + 1, dtype=T)                             # Residual :24-29
    state = {"accumulator": 1.0, "current": 1.0, "beta": 1.0}
    mv_products = 1 if initially_zero else 0                            # :122 (sic)

    def init(initially_zero=False):                                     # init! :235-255
        first_col = V[:, 0]
        first_col[...] = b                                              # :241
        if not initially_zero:
            first_col -= mul(A, x)                                      # :245-246
        Pl.ldiv(first_col)                                              # :249
        beta = float(np.linalg.norm(first_col))                         # :252
        first_col *= T.type(1.0) / T.type(beta) if beta != 0 else np.inf  # :253
        return beta

    state["current"] = init(initially_zero)                             # :126
    state["accumulator"], state["beta"] = 1.0, state["current"]         # init_residual! :257-260
    tol = max(reltol * state["current"], abstol)                        # :129
    beta = state["current"]                                             # g.β  :133
    k = 1
    iteration = 0

    def done(it):
        return it >= maxiter or state["current"] <= tol                 # :55

    while True:
        if done(iteration):                                             # :59
            break
        # expand! :285-304
        if isinstance(Pr, Identity):
            V[:, k] = mul(A, np.ascontiguousarray(V[:, k - 1]))
            if not isinstance(Pl, Identity):
                Pl.ldiv(V[:, k])
        else:
            nextV = V[:, k]
            Pr.ldiv3(nextV, V[:, k - 1])
            nextV[...] = mul(A, np.ascontiguousarray(nextV))
            Pl.ldiv(nextV)
        mv_products += 1                                                # :65
        H[k, k - 1] = orthogonalize_and_normalize_(V[:, :k], V[:, k], H[:k, k - 1], orth_meth)  # :68-73
        # update_residual! :224-233
        if H[k, k - 1] == 0:
            state["current"] = 0.0
        else:
            nullvec[k] = -np.conj(np.vdot(nullvec[:k], H[:k, k - 1]) / H[k, k - 1])
            state["accumulator"] += abs(nullvec[k]) ** 2
            state["current"] = state["beta"] / math.sqrt(state["accumulator"])
        k += 1                                                          # :78
        if k == restart + 1 or done(iteration + 1):                     # :82
            rhs = np.zeros(k, dtype=T)                                  # solve_least_squares! :262-271
            rhs[0] = beta
            hessenberg_ldiv(H[:k, : k - 1], rhs)
            y = rhs[: k - 1]
            if isinstance(Pr, Identity):                                # update_solution! :273-283
                x += V[:, : k - 1] @ y
            else:
                Ax = V[:, : k - 1] @ y
                Pr.ldiv(Ax)
                x += Ax
            k = 1                                                       # :90
            if not done(iteration):                                     # :93 (sic: iteration, not +1)
                beta = init()                                           # :96
                state["accumulator"], state["beta"] = 1.0, beta         # :99 (current NOT reset)
                mv_products += 1                                        # :101
        iteration += 1
        if log:
            history.iters += 1                                          # nextiter! :209
            history.mvps = mv_products                                  # :210
            resnorm.append(state["current"])                            # :211
    history.isconverged = state["current"] <= tol                       # setconv :218 (always)
    if log:
        history["resnorm"] = np.array(resnorm)
        history["tol"] = tol
        return x, history
    return x


def gmres(A, b, **kw):
    """gmres(A, b; kw...) = gmres!(zerox(A, b), A, b; initially_zero = true, kw...)  src/gmres.jl:143."""
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return gmres_(x, A, b, initially_zero=True, **kw)


# --------------------------------------------------------------------------------------------
# MINRES (reference src/minres.jl)
# --------------------------------------------------------------------------------------------
def minres_(x, A, b, *, skew_hermitian=False, log=False, abstol=0.0, reltol=None, maxiter=None,
            initially_zero=False):
    """minres!(x, A, b; ...) -- reference src/minres.jl:200-237, minres_iterable! :39-89,
    iterate :97-159."""
    T = x.dtype
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))
    if maxiter is None:
        maxiter = opsize(A, 1)
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    resnorms = []
    HT = T if skew_hermitian else _real_dtype(T)
    v_prev = np.empty_like(x)
    v_curr = b.astype(T, copy=True)                                     # :49
    v_next = np.empty_like(x)
    w_prev = np.zeros_like(x)   # similar(x): contents unspecified, never read before being written
    w_curr = np.zeros_like(x)
    w_next = np.zeros_like(x)
    mv_products = 0
    if not initially_zero:                                              # :58-63
        v_next[...] = mul(A, x)
        v_curr -= v_next
        mv_products = 1
    resnorm = float(np.linalg.norm(v_curr))                             # :65
    tol = max(reltol * resnorm, abstol)                                 # :66
    H = np.zeros(4, dtype=HT)                                           # :70
    rhs = np.array([resnorm, 0], dtype=HT)                              # :71
    v_curr *= T.type(1.0 / resnorm) if resnorm != 0 else np.inf         # :74
    c_prev, s_prev, c_curr, s_curr = 1.0, 0.0, 1.0, 0.0                 # :77-78
    if log:
        history.mvps = mv_products
    iteration = 1                                                       # start = 1 :93
    while True:
        if iteration > maxiter or resnorm <= tol:                       # :95
            break
        v_next[...] = mul(A, v_curr)                                    # :104
        if iteration > 1:
            v_next -= H[1] * v_prev                                     # :106
        proj = np.vdot(v_curr, v_next)                                  # :109
        H[2] = proj if skew_hermitian else np.real(proj)                # :110
        v_next -= proj * v_curr                                         # :111
        H[3] = np.linalg.norm(v_next)                                   # :114
        with np.errstate(divide="ignore", invalid="ignore"):
            v_next *= T.type(1.0) / H[3]                                # :115
        if iteration > 2:                                               # :118-121
            H[0] = s_prev * H[1]
            H[1] = c_prev * H[1]
        if iteration > 1:                                               # :124-128
            tmp = -np.conj(s_curr) * H[1] + c_curr * H[2]
            H[1] = c_curr * H[1] + s_curr * H[2]
            H[2] = tmp
        c, s, H[2] = givens_algorithm(H[2], H[3])                       # :131
        rhs[1] = -np.conj(s) * rhs[0]                                   # :134
        rhs[0] = c * rhs[0]                                             # :135
        w_next[...] = v_curr                                            # :138
        if iteration > 1:
            w_next -= H[1] * w_curr                                     # :139
        if iteration > 2:
            w_next -= H[0] * w_prev                                     # :140
        with np.errstate(divide="ignore", invalid="ignore"):
            w_next *= T.type(1.0) / H[2]                                # :141
        x += rhs[0] * w_next                                            # :144
        v_prev, v_curr, v_next = v_curr, v_next, v_prev                 # :147
        w_prev, w_curr, w_next = w_curr, w_next, w_prev                 # :148
        c_prev, s_prev, c_curr, s_curr = c_curr, s_curr, c, s           # :149
        rhs[0] = rhs[1]                                                 # :150
        H[1] = -H[3] if skew_hermitian else H[3]                        # :153
        resnorm = float(abs(rhs[1]))                                    # :156
        iteration += 1
        if log:
            history.iters += 1
            history.mvps += 1
            resnorms.append(resnorm)
    if log:
        history.isconverged = resnorm <= tol
        history["resnorm"] = np.array(resnorms)
        history["tol"] = tol
        return x, history
    return x


def minres(A, b, **kw):
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return minres_(x, A, b, initially_zero=True, **kw)


# --------------------------------------------------------------------------------------------
# BiCGStab(l) (reference src/bicgstabl.jl)
# --------------------------------------------------------------------------------------------
def bicgstabl_(x, A, b, l=2, *, abstol=0.0, reltol=None, max_mv_products=None, log=False, Pl=None,
               initial_zero=False, r_shadow=None, rng=None):
    """bicgstabl!(x, A, b, l; ...) -- reference src/bicgstabl.jl:181-219, bicgstabl_iterator!
    :27-73, iterate :79-134.  The reference draws r_shadow = rand(T, n) (:38); parity runs inject
    the same `r_shadow` into both implementations."""
    import scipy.linalg as sla
    T = x.dtype
    n = opsize(A, 0)
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))
    if max_mv_products is None:
        max_mv_products = opsize(A, 1)
    Pl = Pl or Identity()
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    resnorms = []
    mv_products = 0
    if r_shadow is None:
        rng = rng or np.random.default_rng(0)
        r_shadow = rng.random(n).astype(T)                              # :38
    rs = np.zeros((n, l + 1), dtype=T, order="F")                       # :39 (undef)
    us = np.zeros((n, l + 1), dtype=T, order="F")                       # :40
    residual = rs[:, 0]
    if initial_zero:
        residual[...] = b                                               # :47
    else:
        residual[...] = mul(A, x)                                       # :49
        residual[...] = b - residual                                    # :50
        mv_products += 1
    Pl.ldiv(residual)                                                   # :55
    gamma = np.zeros(l, dtype=T)
    omega = sigma = T.type(1)                                           # :58
    nrm = float(np.linalg.norm(residual))                               # :60
    M = np.zeros((l + 1, l + 1), dtype=T, order="F")
    tol = max(reltol * nrm, abstol)                                     # :66
    if log:
        history.mvps = mv_products
    res = nrm
    while True:
        if mv_products >= max_mv_products or res <= tol:                # :77
            break
        sigma = -omega * sigma                                          # :85
        for j in range(1, l + 1):                                       # :88
            rho = np.vdot(r_shadow, rs[:, j - 1])                       # :89
            beta = rho / sigma                                          # :90
            us[:, :j] = rs[:, :j] - beta * us[:, :j]                    # :93
            us[:, j] = mul(A, np.ascontiguousarray(us[:, j - 1]))       # :97
            Pl.ldiv(us[:, j])                                           # :98
            sigma = np.vdot(r_shadow, us[:, j])                         # :100
            alpha = rho / sigma                                         # :101
            rs[:, :j] -= alpha * us[:, 1: j + 1]                        # :103
            rs[:, j] = mul(A, np.ascontiguousarray(rs[:, j - 1]))       # :107
            Pl.ldiv(rs[:, j])                                           # :108
            x += alpha * us[:, 0]                                       # :111
        mv_products += 2 * l                                            # :115
        M[...] = rs.conj().T @ rs                                       # :120
        lu = sla.lu_factor(M[1:, 1:])                                   # :123 (raises on singular, like lu!)
        gamma[...] = sla.lu_solve(lu, M[1:, 0])                         # :124
        us[:, 0] -= us[:, 1:] @ gamma                                   # :126
        x += rs[:, :l] @ gamma                                          # :127
        rs[:, 0] -= rs[:, 1:] @ gamma                                   # :128
        omega = gamma[l - 1]                                            # :130
        res = float(np.linalg.norm(rs[:, 0]))                           # :131
        if log:
            history.iters += 1                                          # :206
            history.mvps = mv_products                                  # :207
            resnorms.append(res)
    if log:
        history.isconverged = res <= tol
        history["resnorm"] = np.array(resnorms)
        history["tol"] = tol
        return x, history
    return x


def bicgstabl(A, b, l=2, **kw):
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return bicgstabl_(x, A, b, l, initial_zero=True, **kw)


# --------------------------------------------------------------------------------------------
# LOBPCG (reference src/lobpcg.jl) -- standard and generalized problem, optional preconditioner,
# no constraint (C = nothing), the path of BASELINE.json configs[4].
# --------------------------------------------------------------------------------------------
@dataclass
class LOBPCGResults:
    lam: np.ndarray
    X: np.ndarray
    tolerance: float
    residual_norms: np.ndarray
    iterations: int
    maxiter: int
    converged: bool
    trace: list


def _rdiv_upper(Ablk, R):
    """rdiv!(A, B::UpperTriangular) -- reference src/lobpcg.jl:345-355 (column sweeps)."""
    s = Ablk.shape[1]
    Ablk[:, 0] = Ablk[:, 0] / R[0, 0]
    for i in range(1, s):
        for j in range(i):
            Ablk[:, i] = Ablk[:, i] - Ablk[:, j] * R[j, i]
        Ablk[:, i] = Ablk[:, i] / R[i, i]
    return Ablk


def _cholqr(X, BX, AX=None, update_AX=False, update_BX=False, generalized=False):
    """CholQR functor -- reference src/lobpcg.jl:365-393."""
    gram = X.conj().T @ BX                                              # :376-379
    R = np.linalg.cholesky(gram.astype(gram.dtype)).conj().T            # cholesky!(Hermitian(gram)) :381 (upper factor)
    _rdiv_upper(X, R)                                                   # :384/:388
    if update_AX:
        _rdiv_upper(AX, R)
    if generalized and update_BX:
        _rdiv_upper(BX, R)


class Constraint:
    """Constraint(Y, B, X) -- reference src/lobpcg.jl:144-224: deflation of a block against span(Y) in the
    B inner product.  `update` mirrors update! (:188-206): the Cholesky factor is EXTENDED BY AN IDENTITY BLOCK
    for the appended columns (they are B-orthonormal Ritz vectors, B-orthogonal to the old Y), not recomputed."""

    def __init__(self, Y, B=None):
        self.B = B
        self.Y = np.array(Y, order="F", copy=True)
        self.BY = self.Y if B is None else np.asfortranarray(mul(B, self.Y))      # :163-168
        g = self.Y.conj().T @ self.BY                                               # :178
        g[np.diag_indices_from(g)] = g.diagonal().real                              # realdiag! :181
        self.U = np.linalg.cholesky(g).conj().T if g.size else g                   # cholesky!(Hermitian(.)) :182 (upper)

    def update(self, X, BX):                                                        # update! :188-206
        k, m = self.Y.shape[1], X.shape[1]
        self.Y = np.asfortranarray(np.hstack([self.Y, X]))
        self.BY = self.Y if self.B is None else np.asfortranarray(np.hstack([self.BY, BX]))
        U = np.eye(k + m, dtype=self.U.dtype if self.U.size else X.dtype)
        U[:k, :k] = self.U
        self.U = U

    def apply(self, X):                                                             # (constr!::Constraint)(X, X_temp) :212-224
        if self.Y.shape[1] > 0:
            import scipy.linalg as sla
            g = self.BY.conj().T @ X                                                # :217
            t = sla.solve_triangular(self.U, sla.solve_triangular(self.U, g, trans="C"))   # ldiv!(tmp, gram_chol, g) :219
            X -= (self.Y @ t).astype(X.dtype)                                       # :220-221


def lobpcg(A, largest, X0, *, B=None, P=None, C=None, tol=None, maxiter=200, log=False, not_zeros=False, rng=None,
           fixed_iterations=False, _constraint=None, _copy=True):
    """lobpcg(A, [B,] largest, X0; P, C, tol, maxiter, log, not_zeros) -- reference src/lobpcg.jl:827-839 +
    lobpcg!(iterator) :865-893 + the step functor :692-749.  `fixed_iterations` (not in the
    reference) disables soft-locking and the early exit so that throughput runs do constant work.
    `_constraint` / `_copy=False` are used by lobpcg_nev (the iterator keeps X and its Constraint between batches)."""
    import scipy.linalg as sla
    T = X0.dtype
    if tol is None:
        tol = float(np.finfo(T).eps) ** 0.3                             # default_tolerance :751
    X = np.array(X0, dtype=T, order="F", copy=True) if _copy else X0    # X = copy(X0) :830
    n, sizeX = X.shape
    if sizeX > n:
        raise ValueError("X column dimension exceeds the row dimension")            # :833
    if 3 * sizeX > n:
        raise ValueError("The LOBPCG algorithms is not stable to use when the matrix size is less than 3 times "
                         "the block size. Please use a dense solver instead.")       # :834
    generalized = B is not None
    constr = _constraint if _constraint is not None else (Constraint(C, B) if C is not None else None)   # :452
    if constr is not None:
        constr.apply(X)                                                 # :868
    if not not_zeros:                                                   # :869-876
        rng = rng or np.random.default_rng(0)
        for j in range(sizeX):
            if np.all(X[:, j] == 0):
                X[:, j] = rng.random(n).astype(T)
        if constr is not None:
            constr.apply(X)                                             # :875
    AXb = np.zeros_like(X)
    BXb = np.zeros_like(X) if generalized else X
    R = np.zeros_like(X); AR = np.zeros_like(X); BR = np.zeros_like(X) if generalized else None
    Pb = np.zeros_like(X); AP = np.zeros_like(X); BP = np.zeros_like(X) if generalized else None
    ritz = np.zeros(3 * sizeX, dtype=T)
    residuals = np.full(sizeX, np.nan, dtype=_real_dtype(T))            # :477
    mask = np.ones(sizeX, dtype=bool)
    trace = []
    iteration = 1
    bs = sizeX

    def precond(blk):                                                   # RPreconditioner :236-242
        if P is not None and not isinstance(P, Identity):
            for j in range(blk.shape[1]):
                col = np.ascontiguousarray(blk[:, j])
                blk[:, j] = P.ldiv3(np.empty_like(col), col)

    def eig_select(gA, gB, subdim):                                     # sub_problem! :607-627
        if gB is None:
            vals, vecs = sla.eigh(gA[:subdim, :subdim])
        else:
            vals, vecs = sla.eigh(gA[:subdim, :subdim], gB[:subdim, :subdim])
        perm = np.argsort(-vals if largest else vals, kind="stable")[:sizeX]   # partialsortperm!(...; rev=largest)
        ritz[:sizeX] = vals[perm]
        return vecs[:, perm].astype(T)

    def residuals_():                                                   # residuals! :533-547
        nonlocal R
        R = AXb - BXb * ritz[:sizeX][None, :]
        for j in range(sizeX):
            residuals[j] = np.sqrt(np.sum((R[:, j] * R[:, j].conj()).real, dtype=R.dtype))

    while iteration <= maxiter:                                         # :880
        if iteration == 1:                                              # :695-703
            if generalized:
                BXb = np.asfortranarray(mul(B, X))
            _cholqr(X, BXb, update_BX=True, generalized=generalized)
            if not generalized:
                BXb = X
            AXb = np.asfortranarray(mul(A, X))
            XAX = X.conj().T @ AXb
            V = eig_select(XAX, None, sizeX)
            X[...] = X @ V                                              # update_X_P!(0, 0) :629-690
            AXb = AXb @ V
            if generalized:
                BXb = BXb @ V
            else:
                BXb = X
            residuals_()
        else:
            aR = np.array(R[:, mask], order="F")                        # update_active! :557-562
            if iteration > 2:
                aP = np.array(Pb[:, mask], order="F")
                aAP = np.array(AP[:, mask], order="F")
                aBP = np.array(BP[:, mask], order="F") if generalized else aP
            precond(aR)                                                 # precond_constr! :564-569
            if constr is not None:
                constr.apply(aR)                                        # :567
            aBR = np.asfortranarray(mul(B, aR)) if generalized else aR  # ortho_AB_mul_X! :524-532
            _cholqr(aR, aBR, update_BX=True, generalized=generalized)
            aAR = np.asfortranarray(mul(A, aR))
            if iteration > 2:
                _cholqr(aP, aBP, AX=aAP, update_AX=True, update_BX=True, generalized=generalized)   # :733
            n1, n2 = sizeX, bs
            n3 = bs if iteration > 2 else 0
            sub = n1 + n2 + n3
            gA = np.zeros((sub, sub), dtype=T)
            gB = np.zeros((sub, sub), dtype=T)
            xr, rr, pr = slice(0, n1), slice(n1, n1 + n2), slice(n1 + n2, sub)
            gA[xr, xr] = np.diag(ritz[:n1])                             # BlockGram functor :282-307
            gA[rr, rr] = aR.conj().T @ aAR                              # RAR :266
            gA[xr, rr] = X.conj().T @ aAR                               # XAR :265
            gA[rr, xr] = gA[xr, rr].conj().T
            gB[xr, xr] = np.eye(n1)                                     # normalized :308-338
            gB[rr, rr] = np.eye(n2)
            gB[xr, rr] = X.conj().T @ aBR                               # XBR :270
            gB[rr, xr] = gB[xr, rr].conj().T
            if n3:
                gA[pr, pr] = aP.conj().T @ aAP                          # PAP :268
                gA[rr, pr] = aAR.conj().T @ aP                          # RAP :267
                gA[xr, pr] = X.conj().T @ aAP                           # XAP :264
                gA[pr, rr] = gA[rr, pr].conj().T
                gA[pr, xr] = gA[xr, pr].conj().T
                gB[pr, pr] = np.eye(n3)
                gB[rr, pr] = aBR.conj().T @ aP                          # RBP :271
                gB[xr, pr] = X.conj().T @ aBP                           # XBP :269
                gB[pr, rr] = gB[rr, pr].conj().T
                gB[pr, xr] = gB[xr, pr].conj().T
            V = eig_select(gA, gB, sub)
            Vx, Vr, Vp = V[xr, :], V[rr, :], V[pr, :]
            Pb = aR @ Vr                                                # update_X_P! :645-651
            AP = aAR @ Vr
            if generalized:
                BP = aBR @ Vr
            if n3:
                Pb = Pb + aP @ Vp                                       # :652-663
                AP = AP + aAP @ Vp
                if generalized:
                    BP = BP + aBP @ Vp
            X[...] = X @ Vx + Pb                                        # :664-689
            AXb = AXb @ Vx + AP
            if generalized:
                BXb = BXb @ Vx + BP
            else:
                BXb = X
            residuals_()
        mask = residuals > tol                                          # update_mask! :549-555
        if fixed_iterations:
            mask[:] = True
        bs = int(mask.sum())
        if log:
            trace.append((iteration, residuals.copy(), ritz[:sizeX].copy()))
        if bs == 0:                                                     # :885
            break
        iteration += 1                                                  # :886
    lam = ritz[:sizeX].copy()
    res = LOBPCGResults(lam, X, tol, residuals.copy(), iteration, maxiter, bool(np.all(residuals <= tol)), trace)
    res.BX = BXb
    return res


def lobpcg_nev(A, largest, X0, nev, *, B=None, P=None, C=None, tol=None, maxiter=200, log=False, not_zeros=False,
               rng=None):
    """lobpcg(A, [B,] largest, X0, nev; ...) -- reference src/lobpcg.jl:925-962: batches of size(X0, 2) Ritz pairs,
    every converged batch appended to the constraint (update!, :947/:954) and the block refilled with rand!."""
    T = X0.dtype
    rng = rng or np.random.default_rng(0)
    n, sizeX = X0.shape
    if nev > n:
        raise ValueError("Number of eigenvectors desired exceeds the row dimension.")          # :933
    if 3 * sizeX > n:
        raise ValueError("The LOBPCG algorithms is not stable to use when the matrix size is less than 3 times "
                         "the block size. Please use a dense solver instead.")                # :934
    if tol is None:
        tol = float(np.finfo(T).eps) ** 0.3
    sizeX = min(nev, sizeX)                                             # :936
    X = np.array(X0[:, :sizeX], dtype=T, order="F", copy=True)         # :937
    constr = Constraint(np.zeros((n, 0), dtype=T) if C is None else C, B)   # LOBPCGIterator(..., nev, P, C) :497-522
    lam = np.zeros(nev, dtype=T)
    resn = np.zeros(nev, dtype=_real_dtype(T))
    Xall = np.zeros((n, nev), dtype=T, order="F")
    iterations, conv = [], np.zeros(nev, dtype=bool)

    def run(nz):
        return lobpcg(A, largest, X, B=B, P=P, tol=tol, maxiter=maxiter, log=log, not_zeros=nz, rng=rng,
                      _constraint=constr, _copy=False)

    def append(r, n1, n2):                                              # append! :79-91
        lam[n1:n1 + n2] = r.lam[-n2:]
        resn[n1:n1 + n2] = r.residual_norms[-n2:]
        Xall[:, n1:n1 + n2] = r.X[:, -n2:]
        iterations.append(r.iterations)
        conv[n1:n1 + n2] = r.converged

    r = run(not_zeros)                                                  # :941
    append(r, 0, sizeX)
    converged_x = sizeX
    while converged_x < nev:                                            # :944
        BX = r.BX if B is not None else X
        if nev - converged_x < sizeX:                                   # :945-952
            cutoff = sizeX - (nev - converged_x)
            constr.update(X[:, :cutoff].copy(), BX[:, :cutoff].copy())
            X[:, :sizeX - cutoff] = X[:, cutoff:sizeX].copy()
            X[:, cutoff:sizeX] = rng.random((n, sizeX - cutoff)).astype(T)
            r = run(True)
            append(r, converged_x, sizeX - cutoff)
            converged_x += sizeX - cutoff
        else:                                                           # :953-959
            constr.update(X.copy(), BX.copy())
            X[...] = rng.random((n, sizeX)).astype(T)
            r = run(True)
            append(r, converged_x, sizeX)
            converged_x += sizeX
    return LOBPCGResults(lam, Xall, tol, resn, iterations, maxiter, conv, [])


# ============================================================================================
# the solvers that need A' (QMR, LSQR, LSMR) and IDR(s)
# ============================================================================================
# --------------------------------------------------------------------------------------------
# QMR (reference src/qmr.jl)
# --------------------------------------------------------------------------------------------
def qmr_(x, A, b, *, abstol=0.0, reltol=None, maxiter=None, log=False, initially_zero=False):
    """qmr!(x, A, b; ...) -- reference src/qmr.jl:262-297; LanczosDecomp constructor :24-60 and
    iterate :63-100; qmr_iterable! :119-151; QMRIterable iterate :157-215.  Restated literally,
    including the early return of the Lanczos step at delta == 0 (:84-86) BEFORE the vector rotation."""
    T = x.dtype
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(b.dtype)))                   # :267
    if maxiter is None:
        maxiter = opsize(A, 1)                                          # :268
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    resnorms = []
    # LanczosDecomp(x, A, b) :24-60
    v_prev = np.zeros_like(x)                                           # :32
    v_curr = b.astype(T, copy=True)                                     # :33
    v_next = np.empty_like(x)                                           # :34
    if not initially_zero:                                              # :35-39
        v_next[...] = mul(A, x)
        v_curr -= v_next
    resnorm = float(np.linalg.norm(v_curr))                             # :40
    with np.errstate(divide="ignore", invalid="ignore"):
        v_curr *= T.type(1.0) / T.type(resnorm)                         # :41
    w_prev = np.zeros_like(x)                                           # :43
    w_curr = v_curr.copy()                                              # :44
    w_next = np.empty_like(x)                                           # :45
    alpha = beta_prev = beta_curr = delta = T.type(0)                   # :47-50
    # qmr_iterable! :119-151
    g = np.array([resnorm, 0], dtype=T)                                 # :131
    H = np.zeros(4, dtype=T)                                            # :132
    c_prev, s_prev, c_curr, s_curr = T.type(1), T.type(0), T.type(1), T.type(0)   # :135-136
    p_prev = np.zeros_like(x)                                           # :139
    p_curr = np.zeros_like(x)                                           # :140
    tol = max(reltol * resnorm, abstol)                                 # :142
    iteration = 1                                                       # start :154
    while True:
        if iteration > maxiter or resnorm <= tol:                       # done :155
            break
        # iterate(::LanczosDecomp, iteration) :63-100
        v_next[...] = mul(A, v_curr)                                    # :68
        alpha = np.vdot(v_next, w_curr)                                 # dot(v_next, w_curr) :70 (conj on the 1st)
        v_next -= np.conj(alpha) * v_curr                               # :71
        if iteration > 1:
            v_next -= np.conj(beta_curr) * v_prev                       # :72-74
        w_next[...] = mul_adjoint(A, w_curr)                            # :76
        w_next -= alpha * w_curr                                        # :77
        if iteration > 1:
            w_next -= delta * w_prev                                    # :78-80
        vw = np.vdot(v_next, w_next)                                    # :82
        delta = math.sqrt(abs(vw))                                      # :83
        if delta != 0:                                                  # :84-86 (early return otherwise)
            beta_prev = beta_curr                                       # :88
            beta_curr = vw / delta                                      # :89
            v_next *= T.type(1.0) / T.type(delta)                       # :91
            w_next *= T.type(1.0) / beta_curr                           # :92
            w_next, w_curr, w_prev = w_prev, w_next, w_curr             # :94
            v_next, v_curr, v_prev = v_prev, v_next, v_curr             # :95
        # QMRIterable iterate :165-212
        H[1] = np.conj(beta_prev)                                       # :168
        H[2] = np.conj(alpha)                                           # :169
        H[3] = delta                                                    # :170
        if iteration > 2:                                               # :173-176
            H[0] = s_prev * H[1]
            H[1] = c_prev * H[1]
        if iteration > 1:                                               # :179-183
            tmp = -np.conj(s_curr) * H[1] + c_curr * H[2]
            H[1] = c_curr * H[1] + s_curr * H[2]
            H[2] = tmp
        c, s, H[2] = givens_algorithm(H[2], H[3])                       # :187
        g[1] = -np.conj(s) * g[0]                                       # :190
        g[0] = c * g[0]                                                 # :191
        v_next[...] = v_prev                                            # :196
        if iteration > 1:
            v_next -= H[1] * p_curr                                     # :197
        if iteration > 2:
            v_next -= H[0] * p_prev                                     # :198
        with np.errstate(divide="ignore", invalid="ignore"):
            v_next *= T.type(1.0) / H[2]                                # :199
        x += g[0] * v_next                                              # :202
        c_prev, s_prev, c_curr, s_curr = c_curr, s_curr, c, s           # :205
        p_prev[...] = p_curr                                            # :206
        p_curr[...] = v_next                                            # :207
        g[0] = g[1]                                                     # :208
        resnorm = float(abs(g[1]))                                      # :212
        iteration += 1
        if log:
            history.iters += 1                                          # nextiter!(history) :285 (mvps not counted)
            resnorms.append(resnorm)
    if log:
        history.isconverged = resnorm <= tol                            # :293
        history["resnorm"] = np.array(resnorms)
        history["tol"] = tol
        return x, history
    return x


def qmr(A, b, **kw):
    """qmr(A, b; kwargs...) = qmr!(zerox(A, b), A, b; initially_zero = true, kwargs...) -- src/qmr.jl:222."""
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return qmr_(x, A, b, initially_zero=True, **kw)


# --------------------------------------------------------------------------------------------
# LSQR (reference src/lsqr.jl)
# --------------------------------------------------------------------------------------------
def lsqr_(x, A, b, *, damp=0.0, atol=None, btol=None, conlim=None, maxiter=None, log=False):
    """lsqr!(x, A, b; ...) -- reference src/lsqr.jl:66-77 and lsqr_method! :90-275, restated literally
    (including `ddnorm += norm(wrho)` :206 and isconverged = istop > 0 :271, which is true for istop = 7)."""
    T = np.result_type(x.dtype, b.dtype).type                           # Adivtype(A, b) :108
    Tr = _real_dtype(np.dtype(T)).type
    m, n = opsize(A, 0), opsize(A, 1)
    eps = _eps(np.dtype(Tr))
    if atol is None:
        atol = math.sqrt(eps)                                           # :91
    if btol is None:
        btol = math.sqrt(eps)
    if conlim is None:
        conlim = 1.0 / math.sqrt(eps)                                   # :92
    if maxiter is None:
        maxiter = max(m, n)                                             # :67
    history = ConvergenceHistory()
    hist = {"resnorm": [], "anorm": [], "rnorm": [], "cnorm": []}
    if len(x) != n:
        raise ValueError(f"x should be of length {n}")                  # :99
    if len(b) != m:
        raise ValueError(f"b should be of length {m}")                  # :100
    if not np.all(np.isfinite(x)):
        raise ValueError("Initial guess for x must be finite")          # :102-104
    itn = istop = 0                                                     # :110
    ctol = Tr(1 / conlim) if conlim > 0 else Tr(0)                      # :111
    Anorm = Acond = ddnorm = res2 = xnorm = xxnorm = z = sn2 = Tr(0)    # :112
    cs2 = Tr(-1)                                                        # :113
    damp = Tr(damp)
    dampsq = damp * damp                                                # :114
    history["atol"], history["btol"], history["ctol"] = atol, btol, ctol

    def finish():
        for k, v_ in hist.items():
            history[k] = np.array(v_, dtype=np.float64)
        return (x, history) if log else x

    u = (b - mul(A, x)).astype(T)                                       # :124
    v = x.astype(T, copy=True)                                          # :125
    beta = Tr(np.linalg.norm(u))                                        # :126
    alpha = Tr(0)                                                       # :127
    if beta > 0:                                                        # :129-134
        history.mtvps = 1
        u *= Tr(1) / beta
        v[...] = mul_adjoint(A, u)
        alpha = Tr(np.linalg.norm(v))
    if alpha > 0:                                                       # :135-137
        v *= Tr(1) / alpha
    w = v.copy()                                                        # :138
    Arnorm = alpha * beta                                               # :141
    if Arnorm == 0:                                                     # :142-144
        return finish()
    rhobar = alpha                                                      # :146
    phibar = bnorm = rnorm = beta                                       # :147
    while itn < maxiter and not history.isconverged:                    # :152
        history.iters += 1                                              # nextiter!(log, mvps=1) :153
        history.mvps += 1
        itn += 1                                                        # :154
        tmpm = mul(A, v)                                                # :163
        u[...] = -alpha * u + tmpm                                      # :164
        beta = Tr(np.linalg.norm(u))                                    # :165
        if beta > 0:                                                    # :166-178
            history.mtvps += 1
            u *= Tr(1) / beta
            Anorm = Tr(math.sqrt(Anorm * Anorm + alpha * alpha + beta * beta + dampsq))   # :169
            tmpn = mul_adjoint(A, u)                                    # :172
            v[...] = -beta * v + tmpn                                   # :173
            alpha = Tr(np.linalg.norm(v))                               # :174
            if alpha > 0:
                v *= Tr(1) / alpha                                      # :176
        rhobar1 = Tr(math.sqrt(rhobar * rhobar + dampsq))               # :182
        cs1 = rhobar / rhobar1                                          # :183
        sn1 = damp / rhobar1                                            # :184
        psi = sn1 * phibar                                              # :185
        phibar = cs1 * phibar                                           # :186
        rho = Tr(math.sqrt(rhobar1 * rhobar1 + beta * beta))            # :190
        cs = rhobar1 / rho                                              # :191
        sn = beta / rho                                                 # :192
        theta = sn * alpha                                              # :193
        rhobar = -cs * alpha                                            # :194
        phi = cs * phibar                                               # :195
        phibar = sn * phibar                                            # :196
        tau = sn * phi                                                  # :197
        t1 = phi / rho                                                  # :200
        t2 = -theta / rho                                               # :201
        x += T(t1) * w                                                  # :203
        w = T(t2) * w + v                                               # :204
        wrho = w * (Tr(1) / rho)                                        # :205
        ddnorm = ddnorm + Tr(np.linalg.norm(wrho))                      # :206
        delta = sn2 * rho                                               # :211
        gambar = -cs2 * rho                                             # :212
        rhs = phi - delta * z                                           # :213
        zbar = rhs / gambar                                             # :214
        xnorm = Tr(math.sqrt(xxnorm + zbar * zbar))                     # :215
        gamma = Tr(math.sqrt(gambar * gambar + theta * theta))          # :216
        cs2 = gambar / gamma                                            # :217
        sn2 = theta / gamma                                             # :218
        z = rhs / gamma                                                 # :219
        xxnorm = xxnorm + z * z                                         # :220
        Acond = Anorm * Tr(math.sqrt(ddnorm))                           # :225
        res1 = phibar * phibar                                          # :226
        res2 = res2 + psi * psi                                         # :227
        rnorm = Tr(math.sqrt(res1 + res2))                              # :228
        Arnorm = alpha * abs(tau)                                       # :229
        r1sq = rnorm * rnorm - dampsq * xxnorm                          # :239
        r1norm = Tr(math.sqrt(abs(r1sq)))                               # :240
        if r1sq < 0:
            r1norm = -r1norm
        hist["resnorm"].append(float(r1norm))                           # :242
        with np.errstate(divide="ignore", invalid="ignore"):
            test1 = rnorm / bnorm                                       # :246
            test2 = Arnorm / (Anorm * rnorm)                            # :247
            test3 = Tr(1) / Acond                                       # :248
            t1 = test1 / (1 + Anorm * xnorm / bnorm)                    # :249
            rtol = btol + atol * Anorm * xnorm / bnorm                  # :250
        hist["cnorm"].append(float(test3))                              # :251
        hist["anorm"].append(float(test2))                              # :252
        hist["rnorm"].append(float(test1))                              # :253
        if itn >= maxiter:
            istop = 7                                                   # :261
        if Tr(1) + test3 <= 1:
            istop = 6                                                   # :262
        if Tr(1) + test2 <= 1:
            istop = 5                                                   # :263
        if Tr(1) + t1 <= 1:
            istop = 4                                                   # :264
        if test3 <= ctol:
            istop = 3                                                   # :267
        if test2 <= atol:
            istop = 2                                                   # :268
        if test1 <= rtol:
            istop = 1                                                   # :269
        history.isconverged = istop > 0                                 # :271
    history["istop"] = istop
    return finish()


def lsqr(A, b, **kw):
    """lsqr(A, b; kwargs...) = lsqr!(zerox(A, b), A, b; kwargs...) -- src/lsqr.jl:8."""
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return lsqr_(x, A, b, **kw)


# --------------------------------------------------------------------------------------------
# LSMR (reference src/lsmr.jl)
# --------------------------------------------------------------------------------------------
def lsmr_(x, A, b, *, atol=1e-6, btol=1e-6, conlim=1e8, maxiter=None, lam=0.0, log=False):
    """lsmr!(x, A, b; atol, btol, conlim, maxiter, λ) -- reference src/lsmr.jl:67-82 and lsmr_method! :88-287,
    restated literally (first-match stopping tests with `break` :274-281, isconverged = istop ∉ (3, 6, 7) :285,
    minrbar = 1e100 :150, unguarded divisions by the norms :116, :120, :175)."""
    T = np.result_type(x.dtype, b.dtype).type                           # Adivtype(A, b) :104
    Tr = _real_dtype(np.dtype(T)).type
    m, n = opsize(A, 0), opsize(A, 1)
    if maxiter is None:
        maxiter = max(m, n)                                             # :68
    if len(x) != n:
        raise ValueError(f"x has length {len(x)} but should have length {n}")   # :98
    if len(b) != m:
        raise ValueError(f"b has length {len(b)} but should have length {m}")   # :102
    history = ConvergenceHistory()
    hist = {"anorm": [], "rnorm": [], "cnorm": []}
    ctol = Tr(1 / conlim) if conlim > 0 else Tr(0)                      # :108
    lam = Tr(lam)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = (b.astype(T) - mul(A, x)).astype(T)                         # :112-114 (on the copy btmp :76-77)
        beta = Tr(np.linalg.norm(u))                                    # :115
        u *= Tr(1) / beta                                               # :116
        v = mul_adjoint(A, u).astype(T)                                 # :118
        alpha = Tr(np.linalg.norm(v))                                   # :119
        v *= Tr(1) / alpha                                              # :120
    history["atol"], history["btol"], history["ctol"] = atol, btol, ctol
    zetabar = alpha * beta                                              # :127
    alphabar = alpha                                                    # :128
    rho = rhobar = cbar = Tr(1)                                         # :129-131
    sbar = Tr(0)                                                        # :132
    h = v.copy()                                                        # :134
    hbar = np.zeros_like(v)                                             # :135
    betadd = beta                                                       # :138
    betad = Tr(0)
    rhodold = Tr(1)
    tautildeold = thetatilde = zeta = d = Tr(0)                         # :141-144
    normA2 = alpha * alpha                                              # :148
    maxrbar = Tr(0)                                                     # :149
    minrbar = np.float64(1e100)                                         # :150 (a Float64 literal: promotes, as in Julia)
    normb = beta                                                        # :153
    istop = 0
    normAr = alpha * beta                                               # :156
    it = 0
    history.mvps = 1                                                    # :160
    history.mtvps = 1                                                   # :161
    if normAr != 0:                                                     # :162
        while it < maxiter:                                             # :163
            history.iters += 1                                          # nextiter!(log, mvps=1) :164
            history.mvps += 1
            it += 1
            with np.errstate(divide="ignore", invalid="ignore"):
                tmp_u = mul(A, v)                                       # :166
                u[...] = tmp_u + u * (-alpha)                           # :167
                beta = Tr(np.linalg.norm(u))                            # :168
                if beta > 0:                                            # :169-176
                    history.mtvps += 1
                    u *= Tr(1) / beta
                    tmp_v = mul_adjoint(A, u)                           # :172
                    v[...] = tmp_v + v * (-beta)                        # :173
                    alpha = Tr(np.linalg.norm(v))                       # :174
                    v *= Tr(1) / alpha                                  # :175
                alphahat = Tr(math.hypot(alphabar, lam))                # :179
                chat = alphabar / alphahat                              # :180
                shat = lam / alphahat                                   # :181
                rhoold = rho                                            # :184
                rho = Tr(math.hypot(alphahat, beta))                    # :185
                c = alphahat / rho                                      # :186
                s = beta / rho                                          # :187
                thetanew = s * alpha                                    # :188
                alphabar = c * alpha                                    # :189
                rhobarold = rhobar                                      # :192
                zetaold = zeta                                          # :193
                thetabar = sbar * rho                                   # :194
                rhotemp = cbar * rho                                    # :195
                rhobar = Tr(math.hypot(cbar * rho, thetanew))           # :196
                cbar = cbar * rho / rhobar                              # :197
                sbar = thetanew / rhobar                                # :198
                zeta = cbar * zetabar                                   # :199
                zetabar = -sbar * zetabar                               # :200
                hbar[...] = hbar * T(-thetabar * rho / (rhoold * rhobarold)) + h   # :203
                x += T(zeta / (rho * rhobar)) * hbar                    # :204
                h[...] = h * T(-thetanew / rho) + v                     # :205
                betaacute = chat * betadd                               # :214
                betacheck = -shat * betadd                              # :215
                betahat = c * betaacute                                 # :218
                betadd = -s * betaacute                                 # :219
                thetatildeold = thetatilde                              # :222
                rhotildeold = Tr(math.hypot(rhodold, thetabar))         # :223
                ctildeold = rhodold / rhotildeold                       # :224
                stildeold = thetabar / rhotildeold                      # :225
                thetatilde = stildeold * rhobar                         # :226
                rhodold = ctildeold * rhobar                            # :227
                betad = -stildeold * betad + ctildeold * betahat        # :228
                tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold   # :230
                taud = (zeta - thetatilde * tautildeold) / rhodold      # :231
                d = d + betacheck * betacheck                           # :232
                normr = Tr(math.sqrt(d + (betad - taud) ** 2 + betadd * betadd))   # :233
                normA2 = normA2 + beta * beta                           # :236
                normA = Tr(math.sqrt(normA2))                           # :237
                normA2 = normA2 + alpha * alpha                         # :238
                maxrbar = max(maxrbar, rhobarold)                       # :241
                if it > 1:
                    minrbar = min(minrbar, rhobarold)                   # :242-244
                condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)   # :245
                normAr = abs(zetabar)                                   # :254
                normx = Tr(np.linalg.norm(x))                           # :255
                test1 = normr / normb                                   # :259
                test2 = normAr / (normA * normr)                        # :260
                test3 = 1 / condA                                       # :261
                hist["cnorm"].append(float(test3))                      # :262
                hist["anorm"].append(float(test2))                      # :263
                hist["rnorm"].append(float(test1))                      # :264
                t1 = test1 / (Tr(1) + normA * normx / normb)            # :267
                rtol = btol + atol * normA * normx / normb              # :268
            if it >= maxiter:
                istop = 7                                               # :274
                break
            if Tr(1) + Tr(test3) <= 1:
                istop = 6                                               # :275
                break
            if Tr(1) + Tr(test2) <= 1:
                istop = 5                                               # :276
                break
            if Tr(1) + Tr(t1) <= 1:
                istop = 4                                               # :277
                break
            if test3 <= ctol:
                istop = 3                                               # :279
                break
            if test2 <= atol:
                istop = 2                                               # :280
                break
            if test1 <= rtol:
                istop = 1                                               # :281
                break
    history.isconverged = istop not in (3, 6, 7)                        # :285
    history["istop"] = istop
    for k, v_ in hist.items():
        history[k] = np.array(v_, dtype=np.float64)
    return (x, history) if log else x


def lsmr(A, b, **kw):
    """lsmr(A, b; kwargs...) = lsmr!(zerox(A, b), A, b; kwargs...) -- src/lsmr.jl:10."""
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return lsmr_(x, A, b, **kw)


# --------------------------------------------------------------------------------------------
# IDR(s) (reference src/idrs.jl)
# --------------------------------------------------------------------------------------------
def _idrs_omega(t, s):
    """omega(t, s) -- reference src/idrs.jl:70-81."""
    angle = math.sqrt(2.0) / 2                                          # :71
    ns = np.linalg.norm(s)                                              # :72
    nt = np.linalg.norm(t)                                              # :73
    ts = np.vdot(t, s)                                                  # :74
    rho = abs(ts / (nt * ns))                                           # :75
    om = ts / (nt * nt)                                                 # :76
    if rho < angle:                                                     # :77-79
        om = om * type(om)(angle) / rho
    return om


def idrs_(X, A, C, *, s=8, Pl=None, abstol=0.0, reltol=None, maxiter=None, log=False, smoothing=False, P=None,
          rng=None):
    """idrs!(x, A, b; s, Pl, abstol, reltol, maxiter, smoothing) -- reference src/idrs.jl:49-64,
    idrs_iterable! :112-145, iterate :163-272.  `P` (list of s vectors) replaces the reference's
    `rand!(copy(C))` draws (:132) so that runs are reproducible; otherwise they come from `rng`."""
    T = C.dtype.type
    Pl = Pl or Identity()
    if reltol is None:
        reltol = math.sqrt(_eps(_real_dtype(C.dtype)))                   # :53
    if maxiter is None:
        maxiter = opsize(A, 1)                                          # :54
    history = ConvergenceHistory()
    history["abstol"], history["reltol"] = abstol, reltol
    resnorms = []
    R = C - mul(A, X)                                                   # :115
    normR = float(np.linalg.norm(R))                                    # :116
    tol = max(reltol * normR, abstol)                                   # :117
    if smoothing:                                                       # :119-122
        X_s, R_s, T_s = X.copy(), R.copy(), np.zeros_like(R)
    if P is None:
        rng = rng or np.random.default_rng()
        P = [rng.random(len(C)).astype(C.dtype) for _ in range(s)]      # :132
    U = [np.zeros_like(C) for _ in range(s)]                            # :133
    G = [np.zeros_like(C) for _ in range(s)]                            # :134
    Q = np.zeros_like(C)                                                # :135
    V = np.zeros_like(C)                                                # :136
    M = np.eye(s, dtype=C.dtype)                                        # :138
    f = np.zeros(s, dtype=C.dtype)                                      # :139
    omega = T(1)                                                        # :142
    it, step = 1, 1                                                     # :163
    with np.errstate(divide="ignore", invalid="ignore"):
        while True:
            if normR < tol or it > maxiter:                             # :167
                history.isconverged = 0 <= normR < tol                  # :168
                if smoothing:
                    X[...] = X_s                                        # :170-172
                break
            if 1 <= step <= s:                                          # :176
                if step == 1:
                    for i in range(s):
                        f[i] = np.vdot(P[i], R)                         # :177-181
                k = step - 1                                            # 0-based
                L = np.tril(M[k:, k:])
                c = np.zeros(s - k, dtype=C.dtype)                      # LowerTriangular(M[k:s,k:s]) \ f[k:s] :186
                for j in range(s - k):
                    c[j] = (f[k + j] - L[j, :j] @ c[:j]) / L[j, j]
                V[...] = c[0] * G[k]                                    # :187
                Q[...] = c[0] * U[k]                                    # :188
                for i in range(k + 1, s):                               # :190-193
                    V += c[i - k] * G[i]
                    Q += c[i - k] * U[i]
                V[...] = R - V                                          # :196
                Pl.ldiv(V)                                              # :199
                U[k][...] = Q + omega * V                               # :201
                G[k][...] = mul(A, U[k])                                # :202
                for i in range(k):                                      # :206-210
                    alpha = np.vdot(P[i], G[k]) / M[i, i]
                    G[k] -= alpha * G[i]
                    U[k] -= alpha * U[i]
                for i in range(k, s):                                   # :214-216
                    M[i, k] = np.vdot(P[i], G[k])
                beta = f[k] / M[k, k]                                   # :220
                R -= beta * G[k]                                        # :221
                X += beta * U[k]                                        # :222
                normR = float(np.linalg.norm(R))                        # :224
                if smoothing:                                           # :225-234
                    T_s[...] = R_s - R
                    gamma = np.vdot(R_s, T_s) / np.vdot(T_s, T_s)
                    R_s -= gamma * T_s
                    X_s -= gamma * (X_s - X)
                    normR = float(np.linalg.norm(R_s))
                if k + 1 < s:                                           # :235-237
                    f[k + 1:] -= beta * M[k + 1:, k]
                nextstep = step + 1
            else:                                                       # step == s + 1 :239
                V[...] = R                                              # :243
                Pl.ldiv(V)                                              # :246
                Q[...] = mul(A, V)                                      # :248
                omega = T(_idrs_omega(Q, R))                            # :249
                R -= omega * Q                                          # :250
                X += omega * V                                          # :251
                normR = float(np.linalg.norm(R))                        # :253
                if smoothing:                                           # :254-263
                    T_s[...] = R_s - R
                    gamma = np.vdot(R_s, T_s) / np.vdot(T_s, T_s)
                    R_s -= gamma * T_s
                    X_s -= gamma * (X_s - X)
                    normR = float(np.linalg.norm(R_s))
                nextstep = 1
            history.iters += 1                                          # nextiter!(it.log, mvps=1) :267
            history.mvps += 1
            resnorms.append(normR)                                      # :268
            it, step = it + 1, nextstep                                 # :271
    if log:
        history["resnorm"] = np.array(resnorms)
        history["tol"] = tol
        return X, history
    return X


def idrs(A, b, **kw):
    """idrs(A, b; kwargs...) = idrs!(zerox(A, b), A, b; kwargs...) -- src/idrs.jl:11."""
    x = np.zeros(opsize(A, 1), dtype=b.dtype)
    return idrs_(x, A, b, **kw)


# --------------------------------------------------------------------------------------------
# svdl (reference src/svdl.jl) -- Golub-Kahan-Lanczos bidiagonalisation with thick restart (method = :ritz)
# --------------------------------------------------------------------------------------------
@dataclass
class PartialFactorization:
    """A ~ P * [B 0; 0 beta] * Q  -- reference src/svdl.jl:76-84.  B is kept dense here (the reference stores a
    Bidiagonal / BrokenArrowBidiagonal, :19-66: same entries)."""
    P: np.ndarray
    Q: np.ndarray
    B: np.ndarray
    beta: float


def _svdl_extend(history, A, L, k, orthleft=False, orthright=True, alpha_thr=1 / math.sqrt(2)):
    """extend!(log, A, L, k) -- reference src/svdl.jl:542-609."""
    l = L.B.shape[1] - 1                                                # :547
    p = L.P[:, l].copy()                                                # :548
    Tr = L.B.dtype.type
    B = np.zeros((k, k), dtype=L.B.dtype)
    B[: L.B.shape[0], : L.B.shape[1]] = L.B
    beta = L.beta                                                       # :561
    for j in range(l + 1, k + 1):                                       # :563 (1-based j)
        history.mtvps += 1                                              # :564
        q = mul_adjoint(A, p)                                           # :565
        if orthright:                                                   # :567-574
            oldqnorm = np.linalg.norm(q)
            q = q - L.Q @ (L.Q.conj().T @ q)
            if np.linalg.norm(q) <= alpha_thr * oldqnorm:
                q = q - L.Q @ (L.Q.conj().T @ q)
        beta = Tr(np.linalg.norm(q))                                    # :576
        q = q * (Tr(1) / beta)                                          # :577
        L.Q = np.column_stack([L.Q, q])                                 # :579
        if j == k:                                                      # :580
            break
        history.mvps += 1                                               # :582
        p = mul(A, q)                                                   # :584
        p = p - beta * L.P[:, j - 1]                                    # :585
        if orthleft:                                                    # :587-594
            oldpnorm = np.linalg.norm(p)
            p = p - L.P @ (L.P.conj().T @ p)
            if np.linalg.norm(p) <= alpha_thr * oldpnorm:
                p = p - L.P @ (L.P.conj().T @ p)
        alpha = Tr(np.linalg.norm(p))                                   # :596
        p = p * (Tr(1) / alpha)                                         # :597
        B[j, j] = alpha                                                 # push!(L.B.dv, alpha) :599 / :602
        B[j - 1, j] = beta                                              # push!(L.B.ev, beta)  :600 / :603
        L.P = np.column_stack([L.P, p])                                 # :605
    L.B = B
    L.beta = beta                                                       # :607
    return L


def _svdl_thickrestart(A, L, U, S, V, l):
    """thickrestart!(A, L, F, l) -- reference src/svdl.jl:376-404."""
    k = V.shape[0]                                                      # :379
    Tr = L.B.dtype.type
    Q = L.Q[:, :k] @ V[:, :l]                                           # :384
    L.Q = np.column_stack([Q, L.Q[:, k]])                               # :385
    f = mul(A, L.Q[:, l])                                               # :390
    rho = L.beta * U[-1, :l]                                            # :391
    L.P = L.P[:, :k] @ U[:, :l]                                         # :392
    f = f - L.P @ rho                                                   # :395
    alpha = Tr(np.linalg.norm(f))                                       # :396
    f = f * (Tr(1) / alpha)                                             # :397
    L.P = np.column_stack([L.P, f])                                     # :398
    g = mul_adjoint(A, f) - alpha * L.Q[:, -1]                          # :400
    L.beta = Tr(np.linalg.norm(g))                                      # :401
    B = np.zeros((l + 1, l + 1), dtype=L.B.dtype)                       # BrokenArrowBidiagonal([S[1:l]; alpha], rho, []) :402
    B[np.arange(l), np.arange(l)] = S[:l]
    B[l, l] = alpha
    B[:l, l] = rho
    L.B = B
    return L


def _svdl_harmonicrestart(A, L, U0, S0, V0, k):
    """harmonicrestart!(A, L, F, k) -- reference src/svdl.jl:424-490 (thick restart with harmonic Ritz values)."""
    import scipy.linalg as sla
    m = L.B.shape[0]                                                    # :427
    dt = L.B.dtype
    Tr = dt.type
    rho = L.beta * U0[-1, :]                                            # residuals of the singular values :431
    BA = np.column_stack([np.diag(S0), rho]).astype(dt)                 # broken arrow matrix :435
    U2, S2, V2t = np.linalg.svd(BA, full_matrices=True)                 # :436
    V2 = V2t.conj().T
    Sigma = S2[:k]                                                      # k largest triplets :439
    U = U0 @ U2[:, :k]                                                  # :440
    M = np.eye(m + 1, dtype=dt)                                         # :441-443
    M[:m, :m] = V0
    M = M @ V2
    Mend = M[-1, :k].copy()                                             # :444
    r0 = np.zeros(m, dtype=dt)                                          # scaled residual of the harmonic Ritz problem :446-447
    r0[-1] = 1
    Bd = np.asarray(L.B, dtype=dt)
    if np.any(np.diag(Bd) == 0):                                        # B \ r singular -> pinv(Matrix(L.B)) * r0 :458
        r = np.linalg.pinv(Bd) @ r0
    else:
        r = sla.solve_triangular(Bd, r0, lower=False)                   # ldiv!(L.B, r0), L.B upper :454
    r = (r * L.beta).astype(dt)                                         # :462
    M = M[:m, :] + np.outer(r, M[m, :])                                 # :463
    M2 = np.zeros((m + 1, k + 1), dtype=dt)                             # :465-468
    M2[:m, :k] = M[:, :k]
    M2[:m, k] = -r
    M2[m, k] = 1
    Qf, R = np.linalg.qr(M2)                                            # :469-470
    Q = L.Q @ Qf[:, : k + 1]                                            # :472
    P = L.P @ U[:, :k]                                                  # :473
    R = R[: k + 1, :k] + np.outer(R[:, k], Mend)                        # :475
    f = mul(A, np.ascontiguousarray(Q[:, k]))                           # :477
    f = f - P @ (P.conj().T @ f)                                        # :478
    alpha = Tr(np.linalg.norm(f))                                       # :479
    f = f * (Tr(1) / alpha)                                             # :480
    P = np.column_stack([P, f])                                         # :481
    B = np.zeros((k + 1, k + 1), dtype=dt)                              # UpperTriangular([Diagonal(Sigma) * triu(R'); 0 ... alpha]) :482
    B[:k, :] = np.diag(Sigma) @ np.triu(R.conj().T)
    B[k, k] = alpha
    g = mul_adjoint(A, f)                                               # :484
    q = Q[:, k]
    g = g - np.vdot(q, g) * q                                           # :487
    L.beta = Tr(np.linalg.norm(g))                                      # :488
    L.P, L.Q, L.B = P, Q, B                                             # :491-493
    return L


def _svdl_isconverged(L, U, S, k, tol, reltol):
    """isconverged(L, F, k, tol, reltol, log) -- reference src/svdl.jl:290-350; returns (conv, delta_sigma)."""
    sigma = S[:k]                                                       # :296
    dsig = L.beta * np.abs(U[-1, :k])                                   # :297
    delta = dsig.copy()                                                 # :300
    if k > 1:                                                           # :307
        d = np.inf
        for i in range(k):
            for j in range(i):
                d = min(d, abs(sigma[i] - sigma[j]))                    # :308-311
        for i in range(k):                                              # :316-340
            a = dsig[i]
            if 2 * a <= d:
                y = a * a / d                                           # :326
                delta[i] = min(delta[i], y)                             # :328
    return delta[:k] < max(tol, reltol * sigma[0]), delta[:k]           # :349


def svdl(A, *, nsv=6, k=None, tol=None, maxiter=None, method="ritz", log=False, v0=None, j=None, reltol=None,
         vecs="none", dolock=False, rng=None):
    """svdl(A; nsv, k, tol, maxiter, method, log, v0, j, reltol, vecs, dolock) -- reference src/svdl.jl:157-175 and
    svdl_method! :177-247 (method = :ritz or :harmonic)."""
    m, n = opsize(A, 0), opsize(A, 1)
    sq = math.sqrt(np.finfo(np.float64).eps)                            # tol::Real = sqrt(eps()) :158, :179 (Float64 literal)
    tol = sq if tol is None else tol
    reltol = sq if reltol is None else reltol
    k = 2 * nsv if k is None else k                                     # :158
    j = nsv if j is None else j                                         # :178
    maxiter = min(m, n) if maxiter is None else maxiter                 # :159
    if method not in ("ritz", "harmonic"):
        raise ValueError(f"Unknown restart method {method}")            # :199 ArgumentError
    if v0 is None:
        rng = rng or np.random.default_rng()
        v0 = rng.standard_normal(n)
        v0 = v0 / np.linalg.norm(v0)                                    # :178
    history = ConvergenceHistory()
    history["tol"] = tol
    conv_h, ritz_h, res_h, beta_h = [], [], [], []
    assert k > 1                                                        # :183
    T = v0.dtype.type
    # build(log, A, v0, k) :353-363
    q = v0.astype(v0.dtype, copy=True)
    beta = np.linalg.norm(q)                                            # :356
    q = q * (T(1) / T(beta))                                            # :357
    p = mul(A, q)                                                       # :358
    alpha = T(np.linalg.norm(p))                                        # :359
    p = p * (T(1) / alpha)                                              # :360
    L = PartialFactorization(p.reshape(m, 1), q.reshape(n, 1), np.array([[alpha]], dtype=v0.dtype), T(beta))
    L = _svdl_extend(history, A, L, k)                                  # :362
    U = S = Vt = None
    l = nsv
    for it in range(1, maxiter + 1):                                    # :188
        history.iters += 1                                              # nextiter!(log) :189
        U, S, Vt = np.linalg.svd(np.asarray(L.B, dtype=v0.dtype))       # F = svd(L.B) :192
        if method == "ritz":
            L = _svdl_thickrestart(A, L, U, S, Vt.conj().T, j)          # :195
        else:
            L = _svdl_harmonicrestart(A, L, U, S, Vt.conj().T, j)       # :197
        L = _svdl_extend(history, A, L, k)                              # :201
        conv, dsig = _svdl_isconverged(L, U, S, l, tol, reltol)         # :207
        conv_h.append(conv.copy()); ritz_h.append(S[:k].copy()); res_h.append(dsig.copy()); beta_h.append(float(L.beta))
        if dolock:                                                      # :214-221
            for i in range(len(conv)):
                if conv[i]:
                    L.B[i, j] = 0                                       # L.B.av[i] = 0
        if np.all(conv):                                                # :222
            history.isconverged = True
            break
    values = S[:l].copy()                                               # :227
    V = Vt.conj().T
    res = values
    if vecs != "none":
        leftvecs = L.P @ U[:, :l] if vecs in ("left", "both") else np.zeros((m, 0), dtype=v0.dtype)       # :230-241
        rightvecs = (L.Q[:, :-1] @ V[:, :l]).conj().T if vecs in ("right", "both") else np.zeros((0, n), dtype=v0.dtype)
        res = (leftvecs, values, rightvecs)                             # LinearAlgebra.SVD(leftvecs, values, rightvecs)
    history["conv"], history["ritz"], history["resnorm"], history["betas"] = conv_h, ritz_h, res_h, beta_h
    return (res, L, history) if log else (res, L)
