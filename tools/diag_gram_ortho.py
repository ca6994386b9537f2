#!/usr/bin/env python
"""k_gram_umma vs legacy on LOBPCG-like data: orthonormal blocks, A-images, nearly dependent P (run under gpurun)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb
from oracle import oracle
ctx = isb.default_context()
NAMES = ["X'AR", "X'R", "R'AR", "X'AP", "X'P", "R'P", "AR'P", "P'AP"]
PRODUCTS = [(0, 2), (0, 1), (1, 2), (0, 4), (0, 3), (1, 3), (2, 3), (3, 4)]
def gram(blocks, variant):
    n = blocks[0].shape[0]
    devs = [isb.DeviceArray.from_numpy(ctx, np.ascontiguousarray(b, dtype=np.float32).reshape(-1)) for b in blocks]
    ptrs = (C.c_void_p * 5)(*[d.ptr for d in devs])
    out = np.zeros(8 * 256)
    isb._lib.check(isb.lib().b200_debug_lobpcg_gram_rr(ctx._h, ptrs, n, variant, out.ctypes.data_as(C.c_void_p)))
    return out.reshape(8, 16, 16)
rng = np.random.default_rng(9)
for N in (24, 64):
    S = oracle.laplace_matrix(np.float64, N, 3).to_scipy()
    n = S.shape[0]
    Q, _ = np.linalg.qr(rng.standard_normal((n, 32)))
    X, R = Q[:, :16], Q[:, 16:]
    P, _ = np.linalg.qr(X @ rng.standard_normal((16, 16)) * 0.7 + R @ rng.standard_normal((16, 16)) * 0.7 + 1e-3 * rng.standard_normal((n, 16)))
    blocks = [b.astype(np.float32) for b in (X, R, S @ R, P, S @ P)]
    B = [b.astype(np.float64) for b in blocks]
    for variant in (2, 1, 101):
        g = gram(blocks, variant)
        line = []
        for p, (l, r) in enumerate(PRODUCTS):
            ref = B[l].T @ B[r]
            line.append(f"{NAMES[p]}:{np.max(np.abs(g[p] - ref)):.1e}/{np.max(np.abs(ref)):.1e} asym")
        # signed mean error of the positive diagonals (bias) for R'AR and P'AP
        b1 = np.mean(np.diag(g[2]) - np.diag(B[1].T @ B[2])) / np.mean(np.diag(B[1].T @ B[2]))
        b2 = np.mean(np.diag(g[7]) - np.diag(B[3].T @ B[4])) / np.mean(np.diag(B[3].T @ B[4]))
        print(f"N={N} variant={variant}: " + " ".join(s.replace(' asym', '') for s in line) + f" | rel bias diag R'AR {b1:+.2e} P'AP {b2:+.2e}", flush=True)
