#!/usr/bin/env python
"""lobpcg fp32 with the tcgen05 Gram (lobpcg_mma = 1) vs the legacy mma.sync Gram (2): natural run at 24^3 and a
fixed-horizon run at 256^3 (run under gpurun)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb
from oracle import oracle
ctx = isb.default_context()
L = isb.lib()
O = oracle.laplace_matrix(np.float32, 24, 3)
A = isb.B200CSR.from_csc_arrays(O.colptr, O.rowval, O.nzval, O.shape, base=O.base)
X0 = np.random.default_rng(1234321).random((O.n, 16)).astype(np.float32)
for mode in (2, 1):
    L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", mode)
    try:
        r = isb.lobpcg(A, False, X0, maxiter=300, log=True)
        print(f"24^3 natural mode={mode}: converged={r.converged} iterations={r.iterations} lam[:3]={r.lam[:3]}")
    except Exception as e:
        print(f"24^3 natural mode={mode}: EXCEPTION {e}")
    for steps in (20, 40, 60, 80):
        try:
            r = isb.lobpcg(A, False, X0, maxiter=steps, _fixed_iterations=True)
            print(f"  fixed {steps}: lam0={r.lam[0]:.6f} max resnorm={np.max(r.residual_norms):.3e} min={np.min(r.residual_norms):.3e}")
        except Exception as e:
            print(f"  fixed {steps}: EXCEPTION {e}")
A2 = isb.B200CSR.laplacian(256, 3, np.float32)
X2 = np.random.default_rng(1234321).random((256 ** 3, 16), dtype=np.float32)
for mode in (2, 1):
    L.b200_ctx_set_option(ctx._h, b"lobpcg_mma", mode)
    for steps in (10, 30, 100):
        Xd = isb.DeviceArray.from_numpy(ctx, X2)
        ctx.sync(); t0 = time.perf_counter()
        try:
            r = isb.lobpcg(A2, False, Xd, maxiter=steps, _fixed_iterations=True)
            ctx.sync(); dt = time.perf_counter() - t0
            print(f"256^3 mode={mode} steps={steps}: {steps / dt:.1f} steps/s lam0={r.lam[0]:.6e} resnorm max={np.max(r.residual_norms):.3e}")
        except Exception as e:
            print(f"256^3 mode={mode} steps={steps}: EXCEPTION {e}")
        Xd.free()
