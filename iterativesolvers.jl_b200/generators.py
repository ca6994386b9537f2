"""Host generators of the reference's test matrices, used as inputs by tests and bench.py.
(C++/OpenMP inside libb200krylov.so; see csrc/gen.cu.)"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import B200Error, check, lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def laplace_matrix(T, N: int, dims: int, base: int = 0, empty=np.empty):
    """laplace_matrix(T, N, dims) of reference test/laplace_matrix.jl:1-12 as the three arrays of a
    SparseMatrixCSC{T,Int64}: returns (colptr, rowval, nzval, (n, n)).  `empty(shape, dtype=...)` allocates
    the arrays (pass device.pinned_empty for page-locked host memory)."""
    n = int(N) ** dims
    nnz = lib().b200_gen_laplace_nnz(N, dims, 0, n)
    if nnz < 0:
        raise B200Error("bad laplace_matrix arguments")
    colptr = empty(n + 1, dtype=np.int64)
    rowval = empty(nnz, dtype=np.int64)
    nzval = empty(nnz, dtype=np.float64)
    got = lib().b200_gen_laplace_csc_i64(N, dims, base, _vp(colptr), _vp(rowval), _vp(nzval))
    assert got == nnz
    return colptr, rowval, nzval.astype(np.dtype(T), copy=False), (n, n)


def laplace_csr_slab(T, N: int, dims: int, row_begin: int, m_local: int, empty=np.empty):
    """rows [row_begin, row_begin+m_local) of laplace_matrix(T, N, dims) as CSR (int32, global columns)."""
    nnz = lib().b200_gen_laplace_nnz(N, dims, row_begin, m_local)
    if nnz < 0:
        raise B200Error("bad laplace slab arguments")
    rowptr = empty(m_local + 1, dtype=np.int32)
    colind = empty(nnz, dtype=np.int32)
    vals = empty(nnz, dtype=np.float64)
    got = lib().b200_gen_laplace_csr_slab_i32(N, dims, row_begin, m_local, _vp(rowptr), _vp(colind), _vp(vals))
    assert got == nnz
    return rowptr, colind, vals.astype(np.dtype(T), copy=False)


def advection_dominated(N: int = 50, beta: float = 1000.0, base: int = 0):
    """advection_dominated(N, beta) of reference benchmark/advection_diffusion.jl:3-30:
    returns (colptr, rowval, nzval, (n, n), b) -- SparseMatrixCSC{Float64,Int64} arrays and the rhs."""
    n = int(N) ** 3
    nnz = lib().b200_gen_laplace_nnz(N, 3, 0, n)
    colptr = np.empty(n + 1, dtype=np.int64)
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz, dtype=np.float64)
    b = np.empty(n, dtype=np.float64)
    got = lib().b200_gen_advection_csc_i64(N, float(beta), base, _vp(colptr), _vp(rowval), _vp(nzval), _vp(b))
    assert got == nnz
    return colptr, rowval, nzval, (n, n), b


def mmread(path, base: int = 0):
    """MatrixMarket.mmread for the sparse (coordinate) format -- what the reference's benchmark scripts use to load
    their operators (benchmark/matrixmarket.jl:2,9-10): returns (colptr, rowval, nzval, shape) of the
    SparseMatrixCSC{Float64,Int64} (symmetric storage expanded, duplicates summed, rows sorted; base = 1 for Julia's
    indexing).  Feed it to B200CSR.from_csc_arrays."""
    p = os.fsencode(path)
    m, n, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    field, sym = C.c_int(), C.c_int()
    check(lib().b200_mm_info(p, C.byref(m), C.byref(n), C.byref(nnz), C.byref(field), C.byref(sym)))
    colptr = np.empty(n.value + 1, dtype=np.int64)
    rowval = np.empty(max(nnz.value, 1), dtype=np.int64)
    nzval = np.empty(max(nnz.value, 1), dtype=np.float64)
    check(lib().b200_mm_read_csc_i64(p, base, nnz.value, _vp(colptr), _vp(rowval), _vp(nzval)))
    return colptr, rowval[: nnz.value], nzval[: nnz.value], (m.value, n.value)


def matread(path, problem="Problem", base: int = 0):
    """MAT.matread(file)["Problem"] of the reference's benchmark/matrixcollection.jl:4-12 -- the MATLAB files of the
    SuiteSparse (University of Florida) collection: returns (colptr, rowval, nzval, shape, b) with the operator
    `Problem.A` as SparseMatrixCSC{Float64,Int64} arrays (rows sorted, base = 1 for Julia's indexing; feed them to
    B200CSR.from_csc_arrays) and the right-hand side `Problem.b[:]` (None when the file has none).  Host-side ingestion,
    read with scipy.io.loadmat (MATLAB v5 / v7 files, zlib-compressed elements included); v7.3 (HDF5) files are rejected."""
    import scipy.io as sio
    import scipy.sparse as sp
    try:
        vars_ = sio.loadmat(os.fspath(path), squeeze_me=False, struct_as_record=False)
    except NotImplementedError as e:            # scipy: "Please use HDF reader for matlab v7.3 files"
        raise B200Error(f"{path}: MATLAB v7.3 (HDF5) files are not supported: {e}")
    if problem not in vars_:
        raise B200Error(f"{path}: no variable `{problem}` (found {sorted(k for k in vars_ if not k.startswith('__'))})")
    P = vars_[problem]
    P = P[0, 0] if isinstance(P, np.ndarray) and P.dtype == object else P
    A = getattr(P, "A", None) if not sp.issparse(P) else P
    if A is None or not sp.issparse(A):
        raise B200Error(f"{path}: `{problem}.A` is not a sparse matrix")
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sum_duplicates()
    A.sort_indices()
    b = getattr(P, "b", None) if not sp.issparse(P) else None
    if b is not None:
        b = np.asarray(b.todense() if sp.issparse(b) else b, dtype=np.float64).reshape(-1, order="F")
    colptr = A.indptr.astype(np.int64) + base
    rowval = A.indices.astype(np.int64) + base
    return colptr, rowval, A.data.astype(np.float64), A.shape, b
