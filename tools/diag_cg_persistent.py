#!/usr/bin/env python
"""persistent vs streaming cg! against the oracle, case by case (run under gpurun)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb
from oracle import oracle
L = isb.lib(); ctx = isb.default_context()
rng = np.random.default_rng(1234321 + 7)
def rel(a, b):
    k = min(len(a), len(b)); return float(np.max(np.abs(a[:k] - b[:k]) / b[:k])) if k else 0.0
for N, dims in ((128, 2), (23, 3)):
    O = oracle.laplace_matrix(np.float64, N, dims, base=1)
    A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=1)
    b = rng.standard_normal(O.n); b /= np.linalg.norm(b)
    d = O.diagonal() * (1.0 + 0.5 * rng.random(O.n))
    for name, kw_d, kw_o in (("cg", {}, {}), ("pcg_rand_diag", {"Pl": isb.JacobiPrec(d)}, {"Pl": oracle.JacobiPrec(d)}), ("cg_maxiter37", {"maxiter": 37}, {"maxiter": 37})):
        xo, ho = oracle.cg(O, b, log=True, **kw_o)
        out = {}
        for mode in (1, 0):
            L.b200_ctx_set_option(ctx._h, b"cg_persistent", mode)
            out[mode] = isb.cg(A, b, log=True, **kw_d)
        L.b200_ctx_set_option(ctx._h, b"cg_persistent", 1)
        (x1, h1), (x0, h0) = out[1], out[0]
        print(f"N={N}^{dims} {name}: iters oracle/persist/stream {ho.niters}/{h1.niters}/{h0.niters}  hist err vs oracle: persist {rel(h1['resnorm'], ho['resnorm']):.2e} stream {rel(h0['resnorm'], ho['resnorm']):.2e}  persist vs stream {rel(h1['resnorm'], h0['resnorm']):.2e} (first 100: {rel(h1['resnorm'][:100], h0['resnorm'][:100]):.2e})  x err persist {np.linalg.norm(x1-xo)/np.linalg.norm(xo):.2e} stream {np.linalg.norm(x0-xo)/np.linalg.norm(xo):.2e}")
