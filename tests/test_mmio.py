"""Matrix Market ingestion (the reference's benchmark/matrixmarket.jl loads its operators
with MatrixMarket.jl): the host reader in libb200krylov.so against scipy.io.mmread -- general, symmetric,
skew-symmetric, pattern and integer fields, duplicate entries, comments, malformed files."""
import os

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp


@pytest.fixture(scope="module")
def isb():
    import iterativesolvers_jl_b200 as m
    return m


def as_csc(cp, rv, nz, shape, base):
    return sp.csc_matrix((nz, rv - base, cp - base), shape=shape)


@pytest.mark.parametrize("base", [0, 1])
def test_mmread_matches_scipy(isb, tmp_path, base):
    G = sp.random(30, 20, 0.2, random_state=1, format="coo")
    S = sp.random(25, 25, 0.2, random_state=2)
    S = (S + S.T).tocoo()
    K = sp.random(25, 25, 0.2, random_state=3)
    K = (K - K.T).tocoo()
    I = sp.coo_matrix(np.round(10 * sp.random(12, 9, 0.3, random_state=4).toarray()).astype(np.int64))
    cases = [("g", G, {}), ("s", S, dict(symmetry="symmetric")), ("k", K, dict(symmetry="skew-symmetric")),
             ("p", G, dict(field="pattern")), ("i", I, dict(field="integer"))]
    for name, M, kw in cases:
        path = os.path.join(tmp_path, name + ".mtx")
        sio.mmwrite(path, M, comment="written by the test-suite\nsecond comment line", **kw)
        cp, rv, nz, shape = isb.mmread(path, base=base)
        ref = sp.csc_matrix(sio.mmread(path, spmatrix=True))
        ref.sort_indices()
        assert shape == ref.shape and cp.dtype == np.int64 and rv.dtype == np.int64 and nz.dtype == np.float64
        assert np.array_equal(cp - base, ref.indptr) and np.array_equal(rv - base, ref.indices)
        assert np.array_equal(nz, ref.data.astype(np.float64))
        assert abs(as_csc(cp, rv, nz, shape, base) - ref).max() == 0


def test_mmread_sums_duplicates_and_handles_edge_cases(isb, tmp_path):
    p = os.path.join(tmp_path, "dup.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n% comment\n\n3 4 5\n1 1 1.5\n3 2 -2\n1 1 2.5\n"
                       "2 4 1e-3\n3 2 4\n")
    cp, rv, nz, shape = isb.mmread(p)
    A = as_csc(cp, rv, nz, shape, 0).toarray()
    assert shape == (3, 4) and A[0, 0] == 4.0 and A[2, 1] == 2.0 and A[1, 3] == 1e-3 and np.count_nonzero(A) == 3
    p = os.path.join(tmp_path, "empty.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n5 5 0\n")
    cp, rv, nz, shape = isb.mmread(p, base=1)
    assert shape == (5, 5) and len(nz) == 0 and np.all(cp == 1)


@pytest.mark.parametrize("text,frag", [
    ("%%NotMatrixMarket matrix coordinate real general\n1 1 0\n", "banner"),
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "coordinate"),
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "complex"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n", "entries announced"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n", "outside"),
    ("%%MatrixMarket matrix coordinate real symmetric\n2 3 1\n1 1 1.0\n", "square"),
])
def test_mmread_rejects_malformed_files(isb, tmp_path, text, frag):
    p = os.path.join(tmp_path, "bad.mtx")
    open(p, "w").write(text)
    with pytest.raises(isb.B200Error) as e:
        isb.mmread(p)
    assert frag in str(e.value)
    with pytest.raises(isb.B200Error):
        isb.mmread(os.path.join(tmp_path, "does_not_exist.mtx"))


@pytest.mark.gpu
def test_solve_an_operator_loaded_from_a_matrix_market_file(isb, oracle, tmp_path):
    """the flow of the reference's benchmark/matrixmarket.jl:9-30: read A, b = A * ones, cg (A symmetric positive
    definite), compare with the known solution."""
    O = oracle.laplace_matrix(np.float64, 12, 3, base=1)
    path = os.path.join(tmp_path, "lap.mtx")
    sio.mmwrite(path, sp.tril(O.to_scipy()).tocoo(), symmetry="symmetric")
    cp, rv, nz, shape = isb.mmread(path, base=1)
    assert np.array_equal(cp, O.colptr) and np.array_equal(rv, O.rowval) and np.array_equal(nz, O.nzval)
    A = isb.B200CSR.from_csc_arrays(cp, rv, nz, shape, base=1)
    xs = np.ones(O.n)
    b = A @ xs
    x, h = isb.cg(A, b, log=True, reltol=1e-12)
    assert h.isconverged and np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(xs)


def test_matread_suitesparse_layout(isb, tmp_path):
    """benchmark/matrixcollection.jl:4-12: vars = matread(file); A, b = vars["Problem"]["A"], vars["Problem"]["b"][:].
    A file written the way the SuiteSparse collection's .mat files are laid out (struct `Problem` with a sparse `A`, a dense
    `b` and metadata), compressed and uncompressed, comes back as the CSC{Float64,Int64} arrays of A and the vector b."""
    A = sp.random(40, 40, 0.1, random_state=5, format="csc") + sp.identity(40, format="csc")
    A = (A + A.T).tocsc()
    A.sort_indices()
    b = np.arange(1.0, 41.0).reshape(-1, 1)
    for compress in (False, True):
        path = os.path.join(tmp_path, f"ACUSIM_Pres_Poisson_{int(compress)}.mat")
        sio.savemat(path, {"Problem": {"A": A, "b": b, "name": "ACUSIM/Pres_Poisson", "id": 1}}, do_compression=compress)
        for base in (0, 1):
            cp, rv, nz, shape, bb = isb.matread(path, base=base)
            assert shape == (40, 40) and cp.dtype == rv.dtype == np.int64 and nz.dtype == np.float64
            assert np.array_equal(cp - base, A.indptr) and np.array_equal(rv - base, A.indices) and np.array_equal(nz, A.data)
            assert np.array_equal(bb, b[:, 0])
    path = os.path.join(tmp_path, "no_problem.mat")
    sio.savemat(path, {"X": np.eye(2)})
    with pytest.raises(isb.B200Error):
        isb.matread(path)
    sio.savemat(path, {"Problem": {"A": np.eye(3)}})
    with pytest.raises(isb.B200Error):
        isb.matread(path)
