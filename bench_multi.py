#!/usr/bin/env python
"""bench_multi.py -- row-partitioned gmres!(restart = 30) and lobpcg (block 16, fp32) on N GPUs (allreduce of
h / of the Gram blocks, halo exchange of the SpMV / SpMM operand), same global 256^3 problem at every N (strong scaling):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_multi.py

Not the driver's bench (that is bench.py): one JSON line with gmres iterations/s and lobpcg steps/s, max over ranks, and a
clock sample of rank 0.  The operator is laplace_matrix(T, 256, 3) for both (config #3's advection matrix has the same
sparsity; its generator is a host-global one)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, rhs_slab
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import iterativesolvers_jl_b200 as isb
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = isb.Context.distributed(local)
    else:
        ctx = isb.Context(local)
    N = 256
    n = N ** 3
    planes = [N * r // world for r in range(world + 1)]
    offs = np.array([p * N * N for p in planes], dtype=np.int64)
    lo, m = int(offs[rank]), int(offs[rank + 1] - offs[rank])
    out = {"n_gpus": world, "grid": N, "n": n}
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    # ---- gmres!(restart = 30, CGS), fixed horizon
    plan = isb.HaloPlan(rank, world, offs).scan_laplacian(N, 3).exchange() if world > 1 else None
    A = isb.B200CSR.laplacian(N, 3, np.float64, lo, m, plan, ctx)
    b = rhs_slab(lo, m)
    bd = isb.DeviceArray.from_numpy(ctx, b)
    xd = isb.DeviceArray.zeros(ctx, m)
    iters = 300
    for rep in range(2):
        isb.lib().b200_fill(ctx._h, m, 0.0, xd._p, 0)
        ctx.barrier(); ctx.sync()
        t0 = time.perf_counter()
        _, h = isb.gmres_(xd, A, bd, restart=30, maxiter=iters, orth_meth="cgs", initially_zero=True, log=True, reltol=0.0)
        ctx.sync()
        dt = ctx.allreduce([time.perf_counter() - t0], op="max")[0]
    out["gmres30_cgs_iters_per_s"] = h.niters / dt
    out["gmres_resnorm_first_last"] = [float(h["resnorm"][0]), float(h["resnorm"][-1])]
    A.close(); bd.free(); xd.free()
    # ---- lobpcg block 16 fp32, 10 steps x 5 solves
    plan = isb.HaloPlan(rank, world, offs).scan_laplacian(N, 3).exchange() if world > 1 else None
    A = isb.B200CSR.laplacian(N, 3, np.float32, lo, m, plan, ctx)
    X0 = np.random.default_rng(1234321).random((n, 16), dtype=np.float32)[lo:lo + m].copy()
    X0d = isb.DeviceArray.from_numpy(ctx, X0)
    Xd = isb.DeviceArray(ctx, X0.shape, np.float32)
    steps, solves, dt = 10, 5, 0.0
    for rep in range(solves + 1):
        isb._lib.check(isb.lib().b200_copy(ctx._h, m * 16, X0d._p, Xd._p, 1))
        ctx.barrier(); ctx.sync()
        t0 = time.perf_counter()
        r = isb.lobpcg(A, False, Xd, maxiter=steps, _fixed_iterations=True)
        ctx.sync()
        if rep > 0:
            dt += ctx.allreduce([time.perf_counter() - t0], op="max")[0]
    out["lobpcg_bs16_f32_steps_per_s"] = steps * solves / dt
    out["lobpcg_lambda_min"] = float(np.min(r.lam))
    if rank == 0:
        out["clocks"] = clocks.stop()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
