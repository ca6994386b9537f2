"""tests/hostsim/sim.py -- ctypes loader of the serial CPU backend for the fused-pass engines
(TEST INFRASTRUCTURE ONLY: see hostsim.cpp).  Builds libhostsim.so with the committed Makefile (g++ only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Csr(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("rowptr", C.c_void_p), ("colind", C.c_void_p),
                ("vals", C.c_void_p)]


class _Out(C.Structure):
    _fields_ = [("iters", C.c_int64), ("mvps", C.c_int64), ("mtvps", C.c_int64), ("n_hist", C.c_int64),
                ("resnorm", C.c_double), ("tol", C.c_double), ("converged", C.c_int32), ("breakdown", C.c_int32),
                ("passes", C.c_int64), ("applies", C.c_int64)]


def lib():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _LIB = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
    return _LIB


class Csr:
    """0-based CSR copy (int64 row pointers, int32 columns) of a scipy matrix, kept alive for the C call."""

    def __init__(self, A, dtype):
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.colind = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.vals = np.ascontiguousarray(A.data, dtype=dtype)
        self.c = _Csr(A.shape[0], A.shape[1], self.rowptr.ctypes.data, self.colind.ctypes.data, self.vals.ctypes.data)


@dataclass
class Outcome:
    iters: int
    mvps: int
    mtvps: int
    resnorm: float
    tol: float
    converged: bool
    breakdown: bool
    hist: np.ndarray
    passes: int
    applies: int


def _outcome(o: _Out, hist):
    return Outcome(o.iters, o.mvps, o.mtvps, o.resnorm, o.tol, bool(o.converged), bool(o.breakdown),
                   hist[: o.n_hist].copy(), o.passes, o.applies)


def qmr_(x, A, b, *, abstol=0.0, reltol=-1.0, maxiter=-1, initially_zero=False, check_every=0, order=0, split=0,
         At=None):
    """the qmr engine (csrc/qmr_core.h) on the serial backend; x updated in place.  At: the adjoint operator's own
    rows (row-partitioned runs); default: the transpose of A."""
    dt = x.dtype
    Ac, Atc = Csr(A, dt), Csr(sp.csr_matrix(A).T if At is None else At, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    cap = (maxiter if maxiter >= 0 else A.shape[1]) + 1
    hist = np.zeros(cap)
    out = _Out()
    st = lib().hostsim_qmr(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Atc.c), C.c_void_p(x.ctypes.data),
                           C.c_void_p(b.ctypes.data), C.c_double(abstol), C.c_double(reltol), C.c_int64(maxiter),
                           C.c_int(initially_zero), C.c_int(check_every), C.c_int64(cap),
                           hist.ctypes.data_as(C.c_void_p), C.c_int(order), C.c_int(split), C.byref(out))
    assert st == 0, st
    return x, _outcome(out, hist)


class _LsOut(C.Structure):
    _fields_ = [("iters", C.c_int64), ("mvps", C.c_int64), ("mtvps", C.c_int64), ("n_hist", C.c_int64),
                ("hist_stride", C.c_int64), ("istop", C.c_int32), ("converged", C.c_int32), ("bad_x", C.c_int32),
                ("early", C.c_int32), ("atol", C.c_double), ("btol", C.c_double), ("ctol", C.c_double),
                ("est", C.c_double * 5), ("passes", C.c_int64), ("applies", C.c_int64)]


@dataclass
class LsOutcome:
    iters: int
    mvps: int
    mtvps: int
    istop: int
    converged: bool
    bad_x: bool
    early: bool
    atol: float
    btol: float
    ctol: float
    est: tuple
    hist: dict          # resnorm (lsqr) / normr (lsmr), anorm, rnorm, cnorm
    passes: int
    applies: int


def _ls(fn, first_row, x, A, b, p0, atol, btol, conlim, maxiter, check_every, order, split, At=None):
    dt = x.dtype
    Ac, Atc = Csr(A, dt), Csr(sp.csr_matrix(A).T if At is None else At, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    cap = maxiter if maxiter >= 0 else max(A.shape)
    hist = np.zeros(4 * max(cap, 1))
    out = _LsOut()
    st = fn(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Atc.c), C.c_void_p(x.ctypes.data),
            C.c_void_p(b.ctypes.data), C.c_double(p0), C.c_double(atol), C.c_double(btol), C.c_double(conlim),
            C.c_int64(maxiter), C.c_int(check_every), C.c_int64(cap), hist.ctypes.data_as(C.c_void_p), C.c_int(order),
            C.c_int(split), C.byref(out))
    assert st == 0, st
    sd, k = out.hist_stride, out.n_hist
    rows = {name: hist[i * sd: i * sd + k].copy() for i, name in enumerate((first_row, "anorm", "rnorm", "cnorm"))}
    return x, LsOutcome(out.iters, out.mvps, out.mtvps, out.istop, bool(out.converged), bool(out.bad_x),
                        bool(out.early), out.atol, out.btol, out.ctol, tuple(out.est), rows, out.passes, out.applies)


def lsqr_(x, A, b, *, damp=0.0, atol=-1.0, btol=-1.0, conlim=-1.0, maxiter=-1, check_every=0, order=0, split=0,
          At=None):
    """the lsqr engine (csrc/lsqr_core.h) on the serial backend; x updated in place."""
    return _ls(lib().hostsim_lsqr, "resnorm", x, A, b, damp, atol, btol, conlim, maxiter, check_every, order, split, At)


def lsmr_(x, A, b, *, lam=0.0, atol=-1.0, btol=-1.0, conlim=-1.0, maxiter=-1, check_every=0, order=0, split=0,
          At=None):
    """the lsmr engine (csrc/lsmr_core.h) on the serial backend; x updated in place."""
    return _ls(lib().hostsim_lsmr, "normr", x, A, b, lam, atol, btol, conlim, maxiter, check_every, order, split, At)


def idrs_(x, A, b, P, *, diag=None, abstol=0.0, reltol=-1.0, maxiter=-1, smoothing=False, check_every=0, order=0,
          split=0, Pl=None):
    """the idrs engine (csrc/idrs_core.h) on the serial backend; P: n x s (Fortran order), x updated in place."""
    dt = x.dtype
    Ac = Csr(A, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    P = np.asfortranarray(P, dtype=dt)
    d = None if diag is None else np.ascontiguousarray(diag, dtype=dt)
    Pc = Csr(Pl, dt) if Pl is not None else None        # callback preconditioner: a matrix whose product is Pl \\ x
    cap = (maxiter if maxiter >= 0 else A.shape[1]) + 1
    hist = np.zeros(cap)
    out = _Out()
    st = lib().hostsim_idrs(C.c_int(dt == np.float64), C.byref(Ac.c), C.c_void_p(x.ctypes.data),
                            C.c_void_p(b.ctypes.data), C.c_int(P.shape[1]), C.c_void_p(P.ctypes.data),
                            C.c_int64(P.shape[0]), C.c_void_p(d.ctypes.data if d is not None else None),
                            C.c_double(abstol), C.c_double(reltol), C.c_int64(maxiter), C.c_int(smoothing),
                            C.c_int(check_every), C.c_int64(cap), hist.ctypes.data_as(C.c_void_p), C.c_int(order),
                            C.c_int(split), C.byref(out), C.byref(Pc.c) if Pc else None)
    assert st == 0, st
    return x, _outcome(out, hist)


def cg_(x, A, b, *, Pl=None, diag=None, abstol=0.0, reltol=-1.0, maxiter=-1, initially_zero=False, check_every=0,
        order=0, split=0):
    """the general-operator cg engine (csrc/cg_core.h) on the serial backend; Pl: a scipy matrix whose product is the
    application of the preconditioner (y = Pl_inverse @ x), diag: Jacobi diagonal; x updated in place."""
    dt = x.dtype
    Ac = Csr(A, dt)
    Pc = Csr(Pl, dt) if Pl is not None else None
    b = np.ascontiguousarray(b, dtype=dt)
    d = None if diag is None else np.ascontiguousarray(diag, dtype=dt)
    cap = (maxiter if maxiter >= 0 else A.shape[1]) + 1
    hist = np.zeros(cap)
    out = _Out()
    st = lib().hostsim_cg(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Pc.c) if Pc else None,
                          C.c_void_p(d.ctypes.data if d is not None else None), C.c_void_p(x.ctypes.data),
                          C.c_void_p(b.ctypes.data), C.c_double(abstol), C.c_double(reltol), C.c_int64(maxiter),
                          C.c_int(initially_zero), C.c_int(check_every), C.c_int64(cap),
                          hist.ctypes.data_as(C.c_void_p), C.c_int(order), C.c_int(split), C.byref(out))
    assert st == 0, st
    return x, _outcome(out, hist)


def gmres_(x, A, b, *, Pl=None, Pr=None, pl_diag=None, pr_diag=None, abstol=0.0, reltol=-1.0, restart=-1, maxiter=-1,
           initially_zero=False, orth_meth="mgs", order=0, split=0):
    """the general-operator gmres engine (csrc/gmres_core.h) on the serial backend; Pl / Pr: scipy matrices whose
    product applies the preconditioner (y = P_inverse @ x), pl_diag / pr_diag: Jacobi diagonals; x updated in place."""
    dt = x.dtype
    Ac = Csr(A, dt)
    Plc = Csr(Pl, dt) if Pl is not None else None
    Prc = Csr(Pr, dt) if Pr is not None else None
    b = np.ascontiguousarray(b, dtype=dt)
    dl = None if pl_diag is None else np.ascontiguousarray(pl_diag, dtype=dt)
    dr = None if pr_diag is None else np.ascontiguousarray(pr_diag, dtype=dt)
    cap = maxiter if maxiter >= 0 else A.shape[1]
    hist = np.zeros(max(cap, 1))
    out = _Out()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_gmres(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Plc.c) if Plc else None,
                             C.byref(Prc.c) if Prc else None, vp(dl), vp(dr), vp(x), vp(b), C.c_double(abstol),
                             C.c_double(reltol), C.c_int(restart), C.c_int64(maxiter), C.c_int(initially_zero),
                             C.c_int({"mgs": 0, "cgs": 1, "dgks": 2}[orth_meth]), C.c_int64(cap), vp(hist), C.c_int(order),
                             C.c_int(split), C.byref(out))
    assert st == 0, st
    return x, _outcome(out, hist)


def chebyshev_(x, A, b, lmin, lmax, *, Pl=None, diag=None, abstol=0.0, reltol=-1.0, maxiter=-1, initially_zero=False,
               check_every=0, order=0, split=0):
    """the general-operator chebyshev engine (csrc/chebyshev_core.h) on the serial backend."""
    dt = x.dtype
    Ac = Csr(A, dt)
    Pc = Csr(Pl, dt) if Pl is not None else None
    b = np.ascontiguousarray(b, dtype=dt)
    d = None if diag is None else np.ascontiguousarray(diag, dtype=dt)
    cap = maxiter if maxiter >= 0 else A.shape[1]
    hist = np.zeros(max(cap, 1))
    out = _Out()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_chebyshev(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Pc.c) if Pc else None, vp(d), vp(x), vp(b),
                                 C.c_double(lmin), C.c_double(lmax), C.c_double(abstol), C.c_double(reltol),
                                 C.c_int64(maxiter), C.c_int(initially_zero), C.c_int(check_every), C.c_int64(cap), vp(hist),
                                 C.c_int(order), C.c_int(split), C.byref(out))
    assert st == 0, st
    return x, _outcome(out, hist)


def powm_(A, x, *, tol=-1.0, maxiter=-1, check_every=0, order=0, split=0):
    """the power-method engine (csrc/powm_core.h) on the serial backend -> (theta, x, outcome); x updated in place."""
    dt = x.dtype
    Ac = Csr(A, dt)
    cap = (maxiter if maxiter >= 0 else A.shape[0]) + 1
    hist = np.zeros(max(cap, 1))
    out, theta = _Out(), C.c_double()
    st = lib().hostsim_powm(C.c_int(dt == np.float64), C.byref(Ac.c), C.c_void_p(x.ctypes.data), C.c_double(tol),
                            C.c_int64(maxiter), C.c_int(check_every), C.c_int64(cap), hist.ctypes.data_as(C.c_void_p),
                            C.c_int(order), C.c_int(split), C.byref(out), C.byref(theta))
    assert st == 0, st
    return theta.value, x, _outcome(out, hist)


def stationary_(method, x, A, b, omega=1.0, *, maxiter=10, order=0, dense=False):
    """jacobi! / gauss_seidel! / sor! / ssor! (csrc/stationary_core.h: one pass per dependency level) on the serial backend.
    Returns (x, info) with info.levels_f / levels_b / passes; raises LinAlgError for a zero or missing diagonal entry."""
    dt = x.dtype
    Ac = Csr(A, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    lf, lb, passes = C.c_int(), C.c_int(), C.c_long()
    fn = lib().hostsim_stationary
    fn.restype = C.c_int64
    r = fn(C.c_int(dt == np.float64), C.byref(Ac.c), C.c_void_p(x.ctypes.data), C.c_void_p(b.ctypes.data),
           C.c_int({"jacobi": 0, "gauss_seidel": 1, "sor": 2, "ssor": 3}[method] | (16 if dense else 0)), C.c_double(omega),
           C.c_int64(maxiter),
           C.c_int(order), C.byref(lf), C.byref(lb), C.byref(passes))
    if r > 0:
        raise np.linalg.LinAlgError(f"SingularException({r})")
    assert r == 0, r
    return x, type("Info", (), dict(levels_f=lf.value, levels_b=lb.value, passes=passes.value))


def minres_(x, A, b, *, abstol=0.0, reltol=-1.0, maxiter=-1, initially_zero=False, skew_hermitian=False, check_every=0,
            order=0, split=0):
    """the general-operator minres engine (csrc/minres_core.h) on the serial backend; x updated in place."""
    dt = x.dtype
    Ac = Csr(A, dt)
    b = np.ascontiguousarray(b, dtype=dt)
    cap = maxiter if maxiter >= 0 else A.shape[1]
    hist = np.zeros(max(cap, 1))
    out = _Out()
    st = lib().hostsim_minres(C.c_int(dt == np.float64), C.byref(Ac.c), C.c_void_p(x.ctypes.data), C.c_void_p(b.ctypes.data),
                              C.c_double(abstol), C.c_double(reltol), C.c_int64(maxiter), C.c_int(initially_zero),
                              C.c_int(skew_hermitian), C.c_int(check_every), C.c_int64(cap), hist.ctypes.data_as(C.c_void_p),
                              C.c_int(order), C.c_int(split), C.byref(out))
    assert st == 0, st
    return x, _outcome(out, hist)


def bicgstabl_(x, A, b, l, shadow, *, Pl=None, diag=None, abstol=0.0, reltol=-1.0, max_mv_products=-1, initial_zero=False,
               check_every=0, order=0, split=0):
    """the general-operator bicgstabl engine (csrc/bicgstabl_core.h) on the serial backend; Pl: a scipy matrix whose
    product applies the preconditioner, diag: Jacobi diagonal; breakdown bit 1 = SingularException in the MR step."""
    dt = x.dtype
    Ac = Csr(A, dt)
    Pc = Csr(Pl, dt) if Pl is not None else None
    b = np.ascontiguousarray(b, dtype=dt)
    sh = np.ascontiguousarray(shadow, dtype=dt)
    d = None if diag is None else np.ascontiguousarray(diag, dtype=dt)
    cap = max_mv_products if max_mv_products >= 0 else A.shape[1]
    hist = np.zeros(max(cap, 1))
    out = _Out()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_bicgstabl(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Pc.c) if Pc else None, vp(d), vp(x), vp(b),
                                 vp(sh), C.c_int(l), C.c_double(abstol), C.c_double(reltol), C.c_int64(max_mv_products),
                                 C.c_int(initial_zero), C.c_int(check_every), C.c_int64(cap), vp(hist), C.c_int(order),
                                 C.c_int(split), C.byref(out))
    assert st == 0, st
    o = _outcome(out, hist)
    o.singular = bool(out.breakdown & 2)
    return x, o


def chunked(kind, x, A, b, chunk, *, Pl=None, Pr=None, pl_diag=None, pr_diag=None, shadow=None, l=2, restart=-1,
            orth_meth="mgs", skew_hermitian=False, abstol=0.0, reltol=-1.0, maxiter=-1, initially_zero=False, order=0, split=0):
    """the resumable form of gmres / minres / bicgstabl on the serial backend: setup, then `chunk` iterations per call
    (fresh history window per call) until done -> (x, outcome, number of calls).  maxiter is max_mv_products for
    bicgstabl."""
    dt = x.dtype
    Ac = Csr(A, dt)
    Plc = Csr(Pl, dt) if Pl is not None else None
    Prc = Csr(Pr, dt) if Pr is not None else None
    b = np.ascontiguousarray(b, dtype=dt)
    arr = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dt)
    dl, dr, sh = arr(pl_diag), arr(pr_diag), arr(shadow)
    cap = (maxiter if maxiter >= 0 else A.shape[1]) + 1
    hist = np.zeros(cap)
    out, calls = _Out(), C.c_int()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_chunked(C.c_int({"gmres": 1, "minres": 2, "bicgstabl": 3, "cg": 4}[kind]), C.c_int(dt == np.float64),
                               C.byref(Ac.c), C.byref(Plc.c) if Plc else None, C.byref(Prc.c) if Prc else None, vp(dl),
                               vp(dr), vp(x), vp(b), vp(sh), C.c_int(l), C.c_int(restart),
                               C.c_int({"mgs": 0, "cgs": 1, "dgks": 2}[orth_meth]), C.c_int(skew_hermitian),
                               C.c_double(abstol), C.c_double(reltol), C.c_int64(maxiter), C.c_int(initially_zero),
                               C.c_int64(chunk), C.c_int64(cap), vp(hist), C.c_int(order), C.c_int(split), C.byref(out),
                               C.byref(calls))
    assert st == 0, st
    o = _outcome(out, hist)
    o.singular = bool(out.breakdown & 2)
    return x, o, calls.value


def constraint_apply_(X, Y, *, appended=0, row_major=False, order=0, split=0):
    """the Constraint passes (csrc/lobpcg_constraint_core.h) on the serial backend: X <- X - Y (chol(Y'Y) \\ Y'X).
    X: n x bs (bs <= 16); row_major: X is laid out like the LOBPCG engine's internal n x 16 blocks."""
    dt = X.dtype
    Y = np.asfortranarray(Y, dtype=dt)
    n, bs = X.shape
    if row_major:
        buf = np.zeros((n, 16), dtype=dt, order="C")
        buf[:, :bs] = X
        rs, cs = 16, 1
    else:
        buf = np.asfortranarray(X)
        rs, cs = 1, n
    st = lib().hostsim_constraint_apply(C.c_int(dt == np.float64), C.c_int64(n), C.c_void_p(Y.ctypes.data), C.c_int64(n),
                                        C.c_int(Y.shape[1]), C.c_int(appended), C.c_void_p(buf.ctypes.data),
                                        C.c_int64(rs), C.c_int64(cs), C.c_int(bs), C.c_int(order), C.c_int(split))
    assert st == 0, st
    X[...] = buf[:, :bs]
    return X


_APPLY_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
_ALLREDUCE_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int)
_dist_keep = []


def set_dist(apply=None, allreduce=None):
    """row-partitioned runs: `apply(op_id, x_ptr, y_ptr) -> 0` replaces every operator application (op_id = address of
    the operator's row-pointer array: Csr(...).rowptr.ctypes.data), `allreduce(buf, count)` sums the pass totals over
    the ranks in place.  set_dist() without arguments restores the serial behaviour."""
    _dist_keep.clear()
    if apply is None:
        lib().hostsim_set_dist(None, None)
        return
    a, r = _APPLY_CB(apply), _ALLREDUCE_CB(allreduce)
    _dist_keep.extend([a, r])
    lib().hostsim_set_dist(a, r)


class _SvdlOut(C.Structure):
    _fields_ = [("iters", C.c_int64), ("mvps", C.c_int64), ("mtvps", C.c_int64), ("converged", C.c_int32),
                ("kdim", C.c_int32), ("beta", C.c_double)]


def svdl(A, v0, *, nsv=6, k=None, j=None, tol=None, reltol=None, maxiter=None, dolock=False, vecs=True, order=0, split=0,
         At=None, method="ritz"):
    """the svdl engine (csrc/svdl_core.h) on the serial backend -> dict(sigma, U, V, iters, mvps, mtvps, converged,
    ritz, resnorm, conv, betas, B)."""
    dt = np.dtype(v0.dtype)
    Ac, Atc = Csr(A, dt), Csr(sp.csr_matrix(A).T if At is None else At, dt)
    m, n = Ac.shape[0], Atc.shape[0]
    k = 2 * nsv if k is None else k
    j = nsv if j is None else j
    sq = float(np.sqrt(np.finfo(np.float64).eps))
    tol = sq if tol is None else tol
    reltol = sq if reltol is None else reltol
    maxiter = min(A.shape) if maxiter is None else maxiter
    v0 = np.ascontiguousarray(v0, dtype=dt)
    sigma = np.zeros(nsv)
    U = np.zeros((m, nsv), dtype=dt, order="F") if vecs else None
    V = np.zeros((n, nsv), dtype=dt, order="F") if vecs else None
    ritz, resn = np.zeros((maxiter, k)), np.zeros((maxiter, nsv))
    conv, betas, Bk = np.zeros((maxiter, nsv), dtype=np.int32), np.zeros(maxiter), np.zeros((k, k), order="F")
    out = _SvdlOut()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_svdl(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Atc.c), vp(v0), C.c_int(nsv), C.c_int(k),
                            C.c_int(j), C.c_double(tol), C.c_double(reltol), C.c_int64(maxiter), C.c_int(dolock),
                            vp(sigma), vp(U), vp(V), vp(ritz), vp(resn), vp(conv), vp(betas), vp(Bk), C.c_int(order),
                            C.c_int(split), C.byref(out), C.c_int({"ritz": 0, "harmonic": 1}[method]))
    assert st == 0, st
    it = out.iters
    return dict(sigma=sigma, U=U, V=V, iters=it, mvps=out.mvps, mtvps=out.mtvps, converged=bool(out.converged),
                ritz=ritz[:it], resnorm=resn[:it], conv=conv[:it].astype(bool), betas=betas[:it], B=Bk, beta=out.beta)


def dense_qr(A):
    """the host thin QR of the svdl harmonic restart (Householder, csrc/svdl_core.h)."""
    A = np.asfortranarray(A, dtype=np.float64)
    rows, cols = A.shape
    Q, R = np.zeros((rows, cols), order="F"), np.zeros((cols, cols), order="F")
    lib().hostsim_dense_qr(C.c_int(rows), C.c_int(cols), C.c_void_p(A.ctypes.data), C.c_void_p(Q.ctypes.data),
                           C.c_void_p(R.ctypes.data))
    return Q, R


def dense_svd(A):
    """the host SVD the svdl engine applies to its projected matrix (one-sided Jacobi, csrc/svdl_core.h)."""
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    U, S, V = np.zeros((n, n), order="F"), np.zeros(n), np.zeros((n, n), order="F")
    lib().hostsim_dense_svd(C.c_int(n), C.c_void_p(A.ctypes.data), C.c_void_p(U.ctypes.data), C.c_void_p(S.ctypes.data),
                            C.c_void_p(V.ctypes.data))
    return U, S, V


def lobpcg_general(A, largest, X0, *, B=None, Pm=None, jac=None, C_=None, tol=-1.0, maxiter=200, fixed=False, order=0,
                   split=0):
    """the general LOBPCG engine (csrc/lobpcg_general_core.h) on the serial backend.  B: second operator (generalized
    problem), Pm: a matrix whose product applies the preconditioner (M \\ x), jac: Jacobi diagonal, C_: constraint basis.
    Returns dict(lam, X, resnorm, iterations, converged, status)."""
    dt = np.dtype(X0.dtype)
    Ac = Csr(A, dt)
    Bc = Csr(B, dt) if B is not None else None
    Pc = Csr(Pm, dt) if Pm is not None else None
    X = np.array(X0, dtype=dt, order="F", copy=True)
    n, sizeX = X.shape
    d = None if jac is None else np.ascontiguousarray(jac, dtype=dt)
    Y = None if C_ is None else np.asfortranarray(C_, dtype=dt)
    lam, rn = np.zeros(sizeX), np.zeros(sizeX)
    tr_r, tr_l = np.zeros((max(maxiter, 1), sizeX)), np.zeros((max(maxiter, 1), sizeX))   # LOBPCGState per iteration (log = true)
    it, conv, status = C.c_int64(), C.c_int(), C.c_int()
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    st = lib().hostsim_lobpcg_general(C.c_int(dt == np.float64), C.byref(Ac.c), C.byref(Bc.c) if Bc else None,
                                      C.byref(Pc.c) if Pc else None, vp(d), vp(Y), C.c_int(0 if Y is None else Y.shape[1]),
                                      vp(X), C.c_int(sizeX), C.c_int(bool(largest)), C.c_double(tol), C.c_int64(maxiter),
                                      C.c_int(fixed), vp(lam), vp(rn), C.c_int(order), C.c_int(split), C.byref(it),
                                      C.byref(conv), C.byref(status), vp(tr_r), vp(tr_l), C.c_int64(maxiter))
    assert st == 0, st
    k = min(it.value, maxiter)
    return dict(lam=lam, X=X, resnorm=rn, iterations=it.value, converged=bool(conv.value), status=status.value,
                trace=[(i + 1, tr_r[i].copy(), tr_l[i].copy()) for i in range(k)])
