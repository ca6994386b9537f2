#!/usr/bin/env python
"""k_gram_umma determinism: the same inputs many times; any run-to-run difference is a race (run under gpurun)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iterativesolvers_jl_b200 as isb
ctx = isb.default_context()
def gram(devs, n, variant):
    ptrs = (C.c_void_p * 5)(*[d.ptr for d in devs])
    out = np.zeros(8 * 256)
    isb._lib.check(isb.lib().b200_debug_lobpcg_gram_rr(ctx._h, ptrs, n, variant, out.ctypes.data_as(C.c_void_p)))
    return out
rng = np.random.default_rng(3)
for n, reps in ((13824, 300), (85285, 200), (64 ** 3, 100), (256 ** 3, 20)):
    blocks = [rng.standard_normal((n, 16)).astype(np.float32) for _ in range(5)]
    devs = [isb.DeviceArray.from_numpy(ctx, b.reshape(-1)) for b in blocks]
    for variant in (1, 2):
        ref = gram(devs, n, variant)
        bad, worst = 0, 0.0
        for _ in range(reps):
            g = gram(devs, n, variant)
            d = float(np.max(np.abs(g - ref)))
            if d != 0.0:
                bad += 1
                worst = max(worst, d)
        print(f"n={n} variant={variant}: {bad}/{reps} runs differ from the first, worst abs diff {worst:.3e} (|entries| ~ {np.sqrt(n):.0f})", flush=True)
    for d in devs:
        d.free()
