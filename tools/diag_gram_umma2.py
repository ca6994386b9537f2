#!/usr/bin/env python
"""k_gram_umma accuracy vs stages per CTA and accumulator hand-over period (run under gpurun)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb
PRODUCTS = [(0, 2), (0, 1), (1, 2), (0, 4), (0, 3), (1, 3), (2, 3), (3, 4)]
ctx = isb.default_context()
def gram(devs, n, variant):
    ptrs = (C.c_void_p * 5)(*[d.ptr for d in devs])
    out = np.zeros(8 * 256)
    isb._lib.check(isb.lib().b200_debug_lobpcg_gram_rr(ctx._h, ptrs, n, variant, out.ctypes.data_as(C.c_void_p)))
    return out.reshape(8, 16, 16)
rng = np.random.default_rng(11)
for n in (64 * 148 * 9 + 37, 64 * 148 * 20, 64 * 148 * 40):
    blocks = [rng.standard_normal((n, 16)).astype(np.float32) for _ in range(5)]
    devs = [isb.DeviceArray.from_numpy(ctx, b.reshape(-1)) for b in blocks]
    B = [b.astype(np.float64) for b in blocks]
    refs = [B[l].T @ B[r] for l, r in PRODUCTS]
    for variant in (2, 101, 102, 108, 132):
        g = gram(devs, n, variant)
        errs = [float(np.max(np.abs(g[p] - refs[p])) / n) for p in range(8)]
        print(f"n={n} stages/CTA~{n / 64 / 148:.1f} variant={variant}: max|err|/n per product " + " ".join(f"{e:.1e}" for e in errs), flush=True)
