"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/b200krylov.h declares,
compute entry points fail loudly without a device (no CPU fallback), host generators match the
oracle, and the multi-GPU halo plan (pure host code) is exercised with a world_size-2 gloo group."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 1234321


@pytest.fixture(scope="module")
def isb():
    import iterativesolvers_jl_b200 as m
    return m


def test_library_exports_every_symbol_of_the_header(isb):
    header = open(os.path.join(ROOT, "include", "b200krylov.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", header))
    assert len(declared) >= 50
    from importlib import import_module
    _lib = import_module("iterativesolvers_jl_b200._lib")
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    L = isb.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.b200_version() >= 100


def test_no_cpu_fallback_without_a_device(isb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(isb.B200Error) as e:
        isb.Context(0)
    assert "CUDA" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iterativesolvers.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_host_generators_match_oracle(isb, oracle):
    for N, dims in [(7, 1), (9, 2), (6, 3)]:
        cp, rv, nz, shape = isb.laplace_matrix(np.float64, N, dims, base=1)
        O = oracle.laplace_matrix(np.float64, N, dims, base=1)
        assert shape == O.shape
        assert np.array_equal(cp, O.colptr) and np.array_equal(rv, O.rowval) and np.array_equal(nz, O.nzval)
        n = N ** dims
        lo, m = n // 3, n - n // 3 - 1
        rp, ci, va = isb.laplace_csr_slab(np.float64, N, dims, lo, m)
        S = O.to_scipy().tocsr()[lo:lo + m]
        S.sort_indices()
        assert np.array_equal(rp, S.indptr) and np.array_equal(ci, S.indices) and np.array_equal(va, S.data)


def test_advection_generator_matches_oracle(isb, oracle):
    """reference benchmark/advection_diffusion.jl:3-30: library generator == scipy restatement (matrix bitwise)."""
    for N in (5, 12):
        cp, rv, nz, shape, b = isb.advection_dominated(N, 1000.0, base=1)
        M, bo = oracle.advection_dominated(N, 1000.0)
        O = oracle.CSC.from_scipy(M, base=1)
        assert shape == O.shape and np.array_equal(cp, O.colptr) and np.array_equal(rv, O.rowval)
        assert np.array_equal(nz, O.nzval)
        np.testing.assert_allclose(b, bo, rtol=4e-16, atol=1e-300)       # libm vs numpy: 1 ulp


def test_halo_plan_single_process(isb, oracle):
    """plan of rank 1 of 3 for a 2-D Laplacian slab, scanned from the CSR columns and analytically."""
    N, dims = 8, 2
    n = N ** dims
    offs = np.array([0, 24, 40, 64], dtype=np.int64)
    for rank in range(3):
        lo, hi = int(offs[rank]), int(offs[rank + 1])
        rp, ci, va = isb.laplace_csr_slab(np.float64, N, dims, lo, hi - lo)
        p1 = isb.HaloPlan(rank, 3, offs).scan_csr(rp, ci)
        p2 = isb.HaloPlan(rank, 3, offs).scan_laplacian(N, dims)
        want = np.unique(ci[(ci < lo) | (ci >= hi)])
        got1 = np.concatenate([p1.recv_cols(o) for o in range(3)])
        got2 = np.concatenate([p2.recv_cols(o) for o in range(3)])
        assert np.array_equal(got1, want) and np.array_equal(got2, want)
        assert p1.n_halo == want.size
        for o in range(3):
            c = p1.recv_cols(o)
            assert np.all((c >= offs[o]) & (c < offs[o + 1]))
        assert p1.local_index(lo) == 0 and p1.local_index(hi - 1) == hi - lo - 1
        if want.size:
            assert p1.local_index(int(want[0])) == hi - lo
        far = (lo + n // 2) % n
        if not (lo <= far < hi) and far not in want:
            assert p1.local_index(far) == -1                      # neither owned nor in the halo
        with pytest.raises(isb.B200Error):                        # a peer may only ask for rows this rank owns
            p1.set_send((rank + 1) % 3, np.array([hi % n], dtype=np.int64))


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
import iterativesolvers_jl_b200 as isb
from oracle import oracle

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), 2
N, dims = 10, 3
n = N ** dims
offs = np.array([0, 4 * N * N, n], dtype=np.int64)               # uneven z-slabs
lo, hi = int(offs[rank]), int(offs[rank + 1])
rp, ci, va = isb.laplace_csr_slab(np.float64, N, dims, lo, hi - lo)
plan = isb.HaloPlan(rank, world, offs).scan_csr(rp, ci).exchange()
other = 1 - rank
# what I send is what the peer asked for
objs = [None, None]
dist.all_gather_object(objs, plan.recv_cols(other))
assert plan.send_count(other) == len(objs[other])
# z-slabs of a stencil: the rows the peer needs are one contiguous plane -> the fused push of the CG update kernel
want_lo = (hi - lo) - N * N if rank == 0 else 0
assert plan.send_range(other) == want_lo and plan.send_count(other) == N * N and plan.send_range(rank) is None
scat = isb.HaloPlan(rank, world, offs)
scat.set_send(other, np.array([lo, lo + 2, lo + 3], dtype=np.int64))     # scattered rows: packed push, no range
assert scat.send_range(other) is None
scat.set_send(other, np.array([lo + 5, lo + 6, lo + 7], dtype=np.int64))
assert scat.send_range(other) == 5
# emulate the halo SpMV on the host: x slab + received halo values, local extended indices
rng = np.random.default_rng(1234321)
x_global = rng.standard_normal(n)
x_loc = x_global[lo:hi]
need = plan.recv_cols(other)
got = [None, None]
dist.all_gather_object(got, (need, None))
send_vals = x_loc[got[other][0] - lo]                               # pack what the peer needs
recv = [None, None]
dist.all_gather_object(recv, send_vals)
halo = recv[other]
assert halo.shape[0] == plan.n_halo
ext = np.concatenate([x_loc, halo])
loc = np.array([plan.local_index(int(c)) for c in ci])
assert loc.min() >= 0
y = np.zeros(hi - lo)
for i in range(hi - lo):
    s = 0.0
    for k in range(rp[i], rp[i + 1]):
        s += va[k] * ext[loc[k]]
    y[i] = s
O = oracle.laplace_matrix(np.float64, N, dims)
y_ref = oracle.csc_spmv(O, x_global)[lo:hi]
assert np.array_equal(y, y_ref), np.abs(y - y_ref).max()
# distributed CG on the host with the same partition (gloo allreduce for the two dots): the N>1 control flow
b_global = rng.standard_normal(n); b_global /= np.linalg.norm(b_global)
import torch
def spmv(v_loc):
    g = [None, None]
    dist.all_gather_object(g, v_loc[got[other][0] - lo])
    e = np.concatenate([v_loc, g[other]])
    out = np.zeros(hi - lo)
    np.add.at(out, np.repeat(np.arange(hi - lo), np.diff(rp)), va * e[loc])
    return out
def gdot(a, b):
    t = torch.tensor([float(a @ b)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])
x = np.zeros(hi - lo); r = b_global[lo:hi].copy(); u = np.zeros(hi - lo)
res = np.sqrt(gdot(r, r)); prev = 1.0; tol = 1e-8 * res; it = 0
while it < 200 and res > tol:
    u = r + (res ** 2 / prev ** 2) * u
    c = spmv(u)
    alpha = res ** 2 / gdot(u, c)
    x += alpha * u; r -= alpha * c
    prev, res = res, np.sqrt(gdot(r, r)); it += 1
xo, ho = oracle.cg_csc_c(np.zeros(n), O, b_global, initially_zero=True, reltol=1e-8)
assert it == ho.niters, (it, ho.niters)
assert np.linalg.norm(x - xo[lo:hi]) <= 1e-10 * np.linalg.norm(xo)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_halo_plan_and_partitioned_cg_world2_gloo(tmp_path):
    """world_size-2 gloo group on CPU: the real plan code (C++ in the .so) + torch.distributed exchange;
    a host emulation of the halo SpMV must reproduce the oracle's global SpMV bit for bit, and the
    partitioned CG control flow (2 allreduces per iteration) must match the oracle's CG."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o


def test_julia_binding_source_uses_only_declared_symbols_and_matching_structs():
    """integration/B200Krylov.jl (the reference-side binding; Julia is absent here, so the file is source only):
    every ccall names a symbol the header declares, and its struct definitions list the header's fields in
    order (names as in the ctypes mirror, which test_signature_table_matches_header pins to the header)."""
    import re
    src = open(os.path.join(ROOT, "integration", "B200Krylov.jl")).read()
    hdr = open(os.path.join(ROOT, "include", "b200krylov.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", hdr))
    used = set(re.findall(r"\(:(b200_\w+),\s*LIB\)", src))
    assert used and used <= declared, sorted(used - declared)
    from importlib import import_module
    L = import_module("iterativesolvers_jl_b200")._lib

    def julia_fields(name):
        body = re.search(r"struct\s+" + name + r"\b(.*?)\n\s*end", src, re.S).group(1)
        body = re.sub(r"#.*", "", body)
        return [f for f in re.findall(r"\b(\w+)::", body)]

    for jl, ct in (("Precond", L.Precond), ("Result", L.Result), ("CgOpts", L.CgOpts), ("GmresOpts", L.GmresOpts),
                   ("MinresOpts", L.MinresOpts), ("BicgstablOpts", L.BicgstablOpts), ("LobpcgOpts", L.LobpcgOpts),
                   ("LobpcgResult", L.LobpcgResult), ("QmrOpts", L.QmrOpts), ("LsqOpts", L.LsqOpts),
                   ("LsqResult", L.LsqResult), ("IdrsOpts", L.IdrsOpts), ("LinOp", L.LinOp),
                   ("SvdlOpts", L.SvdlOpts), ("SvdlResult", L.SvdlResult), ("PowmOpts", L.PowmOpts)):
        assert julia_fields(jl) == [f[0] for f in ct._fields_], jl


def test_committed_ncu_traffic_matches_the_algorithmic_bytes():
    """profiles/k2_traffic.json (echoed by bench.py as roofline.traffic) is the DRAM traffic of one launch of the
    dominant kernel from an `ncu --set full` capture; it must be within 2 % of SURVEY 8(d)'s SpMV bytes."""
    import json
    t = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
    N = t["grid"]
    n, nnz = N ** 3, 7 * N ** 3 - 6 * N ** 2
    algorithmic = nnz * 12 + (n + 1) * 4 + 2 * n * 8
    assert t["dram_bytes_per_launch"] == t["dram_bytes_read"] + t["dram_bytes_write"]
    assert abs(t["dram_bytes_per_launch"] - algorithmic) <= 0.02 * algorithmic
    assert os.path.exists(os.path.join(ROOT, t["source"].split(" ")[0]))


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every struct the host mirror passes by pointer has, field by field, the offset and the total size that gcc
    gives the declaration in include/b200krylov.h (the header is compiled, not parsed)."""
    from importlib import import_module
    L = import_module("iterativesolvers_jl_b200")._lib
    pairs = [("b200_precond", L.Precond), ("b200_result", L.Result), ("b200_cg_opts", L.CgOpts),
             ("b200_gmres_opts", L.GmresOpts), ("b200_minres_opts", L.MinresOpts),
             ("b200_bicgstabl_opts", L.BicgstablOpts), ("b200_lobpcg_opts", L.LobpcgOpts),
             ("b200_lobpcg_result", L.LobpcgResult), ("b200_qmr_opts", L.QmrOpts), ("b200_lsq_opts", L.LsqOpts),
             ("b200_lsq_result", L.LsqResult), ("b200_idrs_opts", L.IdrsOpts), ("b200_linop", L.LinOp),
             ("b200_svdl_opts", L.SvdlOpts), ("b200_svdl_result", L.SvdlResult), ("b200_powm_opts", L.PowmOpts)]
    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{os.path.join(ROOT, "include", "b200krylov.h")}"',
             'int main(void) {']
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, ct in pairs:
        assert got[(cname, "size")] == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert got[(cname, fname)] == getattr(ct, fname).offset, (cname, fname)
