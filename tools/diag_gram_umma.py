#!/usr/bin/env python
"""Diagnostic for k_gram_umma (run under gpurun): structured inputs whose Gram blocks identify layout / descriptor
mistakes, printed compactly.  Not a test."""
import ctypes as C
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb

PRODUCTS = [(0, 2), (0, 1), (1, 2), (0, 4), (0, 3), (1, 3), (2, 3), (3, 4)]
NAMES = ["X'AR", "X'R", "R'AR", "X'AP", "X'P", "R'P", "AR'P", "P'AP"]
ctx = isb.default_context()


def gram(blocks, variant):
    n = blocks[0].shape[0]
    devs = [isb.DeviceArray.from_numpy(ctx, np.ascontiguousarray(b, dtype=np.float32).reshape(-1)) for b in blocks]
    ptrs = (C.c_void_p * 5)(*[d.ptr for d in devs])
    out = np.zeros(8 * 256)
    isb._lib.check(isb.lib().b200_debug_lobpcg_gram_rr(ctx._h, ptrs, n, variant, out.ctypes.data_as(C.c_void_p)))
    return out.reshape(8, 16, 16)


def report(tag, blocks):
    B = [b.astype(np.float64) for b in blocks]
    for v in (1, 2):
        try:
            g = gram(blocks, v)
        except Exception as e:
            print(tag, "variant", v, "ERROR", e)
            continue
        line = []
        for p, (l, r) in enumerate(PRODUCTS):
            ref = B[l].T @ B[r]
            err = np.max(np.abs(g[p] - ref))
            line.append(f"{NAMES[p]}:{err:.1e}/{np.max(np.abs(ref)):.1e}")
        print(f"{tag} v{v}  " + " ".join(line))
        if v == 1:
            for p, (l, r) in enumerate(PRODUCTS):
                ref = B[l].T @ B[r]
                if np.max(np.abs(g[p] - ref)) > 1e-3 * max(1.0, np.max(np.abs(ref))):
                    np.set_printoptions(linewidth=250, precision=3, suppress=True)
                    print(f"  {NAMES[p]} got (first 4 rows):\n{g[p][:4]}\n  ref:\n{ref[:4]}")
                    nz = np.argwhere(np.abs(g[p]) > 1e-6)
                    print("  nonzeros at", nz[:12].tolist(), "count", len(nz))
                    break


for n in (8, 16, 64):
    z = lambda: np.zeros((n, 16), dtype=np.float32)
    # case A: single row 0: X[0,i] = i+1, R[0,j] = 100(j+1)
    X, R, AR, P, AP = z(), z(), z(), z(), z()
    X[0] = np.arange(1, 17); R[0] = 100 * np.arange(1, 17)
    report(f"n={n} A(row0 X,R)", [X, R, AR, P, AP])
    # case B: row 3
    X, R, AR, P, AP = z(), z(), z(), z(), z()
    X[3] = np.arange(1, 17); AR[3] = 10 * np.arange(1, 17)
    report(f"n={n} B(row3 X,AR)", [X, R, AR, P, AP])
    # case C: P, AP on row n-1
    X, R, AR, P, AP = z(), z(), z(), z(), z()
    P[n - 1] = np.arange(1, 17); AP[n - 1] = 7 * np.arange(1, 17)
    report(f"n={n} C(last row P,AP)", [X, R, AR, P, AP])
rng = np.random.default_rng(5)
for n in (8, 64, 200, 64 * 148 * 3 + 5):
    blocks = [rng.standard_normal((n, 16)).astype(np.float32) for _ in range(5)]
    report(f"n={n} random", blocks)
# non-representable-in-tf32 values: checks the hi/lo split
blocks = [(rng.standard_normal((64, 16)) * 1.2345678).astype(np.float32) + np.float32(1e-3) for _ in range(5)]
report("n=64 split", blocks)
