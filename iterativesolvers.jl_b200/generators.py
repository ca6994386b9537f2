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
