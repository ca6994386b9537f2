"""Worker for the multi-GPU parity test; launched by tests/test_gpu_multi.py through torchrun
(one process per GPU, NCCL).  Every rank owns a z-slab of laplace_matrix(Float64, N, 3); the
row-partitioned cg! must reproduce the single-GPU cg! (run on rank 0 on the global matrix):
same iteration count, residual history and solution within 1e-10 relative (only the summation order
of the two global dots and of the halo rows differs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import iterativesolvers_jl_b200 as isb
    ctx = isb.Context.distributed(local)
    n = N ** 3
    # uneven slabs on purpose
    cuts = [0] + [int(round(N * (r + 1) / world + (0.3 if r % 2 == 0 and r + 1 < world else 0))) for r in range(world)]
    cuts[-1] = N
    offs = np.array([c * N * N for c in cuts], dtype=np.int64)
    lo, m = int(offs[rank]), int(offs[rank + 1] - offs[rank])
    rng = np.random.default_rng(1234321)
    b_global = rng.standard_normal(n)
    b_global /= np.linalg.norm(b_global)

    # (1) operator from a host CSR slab with global columns (general path), (2) on-device generator
    rp, ci, va = isb.laplace_csr_slab(np.float64, N, 3, lo, m)
    plan = isb.HaloPlan(rank, world, offs).scan_csr(rp, ci).exchange()
    A1 = isb.B200CSR.from_csr_slab(rp, ci, va, n, lo, 0, plan, ctx)
    plan2 = isb.HaloPlan(rank, world, offs).scan_laplacian(N, 3).exchange()
    A2 = isb.B200CSR.laplacian(N, 3, np.float64, lo, m, plan2, ctx)
    assert A1.n_halo == A2.n_halo and A1.nnz == A2.nnz
    r1, c1, v1 = A1.download()
    r2, c2, v2 = A2.download()
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)

    # distributed SpMV == slab of the global SpMV (bitwise: same per-row order)
    xg = rng.standard_normal(n)
    y_loc = A1 @ xg[lo:lo + m]
    # both transports of the collectives: NCCL (comm=1) and NVLink peer memory fused into the kernels (comm=2)
    import ctypes as C
    L = isb.lib()
    peer_ok = C.c_int64()
    assert L.b200_ctx_get_option(ctx._h, b"peer_ok", C.byref(peer_ok)) == 0
    results = {}
    for name, A, comm in (("slab", A1, 1), ("generated", A2, 2 if peer_ok.value else 1)):
        assert L.b200_ctx_set_option(ctx._h, b"comm", comm) == 0
        x_loc = np.zeros(m)
        x_loc, h = isb.cg_(x_loc, A, b_global[lo:lo + m].copy(), initially_zero=True, log=True, reltol=1e-9)
        results[name] = (x_loc, h)
        # Jacobi-PCG and a nonzero initial guess on the same transport
        xp = np.full(m, 0.25)
        xp, hp = isb.cg_(xp, A, b_global[lo:lo + m].copy(), Pl=isb.JacobiPrec(A.diag(), ctx), log=True, reltol=1e-9)
        results[name + "_pcg"] = (xp, hp)
    L.b200_ctx_set_option(ctx._h, b"comm", 0)
    if rank == 0:
        print(f"peer_ok={peer_ok.value}")
    # the other solvers on the same row partition (halo exchange + NCCL allreduce of their reductions)
    b_loc = b_global[lo:lo + m].copy()
    xm, hm = isb.minres_(np.zeros(m), A1, b_loc, initially_zero=True, log=True, reltol=1e-9)
    xgm, hg = isb.gmres_(np.zeros(m), A1, b_loc, initially_zero=True, log=True, restart=20, maxiter=60, orth_meth="dgks")
    rsh = np.random.default_rng(7).random(n)
    xb, hb = isb.bicgstabl_(np.zeros(m), A1, b_loc, 2, initial_zero=True, log=True, max_mv_products=120,
                            r_shadow=rsh[lo:lo + m].copy(), reltol=1e-9)
    # lobpcg on the row partition: block halo exchange in the SpMM, allreduce of the Gram blocks and norms
    X0 = np.random.default_rng(11).random((n, 4))
    rl = isb.lobpcg(A1, False, X0[lo:lo + m].copy(), maxiter=6, _fixed_iterations=True)
    # chebyshev: one global reduction per step; 12 steps with the analytic spectral bounds (fixed horizon)
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    xc, hc = isb.chebyshev_(np.zeros(m), A1, b_loc, 3 * lam1[0], 3 * lam1[-1], initially_zero=True, log=True,
                            maxiter=12, reltol=1e-12)
    others = [None] * world
    dist.all_gather_object(others, (xm, xgm, xb, rl.X, xc))

    gathered = [None] * world
    dist.all_gather_object(gathered, (y_loc, results["slab"][0], results["generated"][0],
                                      results["slab_pcg"][0], results["generated_pcg"][0]))
    if rank == 0:
        ctx1 = isb.Context(local)
        cp, rv, nz, shape = isb.laplace_matrix(np.float64, N, 3, base=1)
        Ag = isb.B200CSR.from_csc_arrays(cp, rv, nz, shape, base=1, ctx=ctx1)
        y_ref = Ag @ xg
        y_dist = np.concatenate([g[0] for g in gathered])
        assert np.array_equal(y_dist, y_ref), np.abs(y_dist - y_ref).max()
        xs, hs = isb.cg(Ag, b_global, log=True, reltol=1e-9)
        for idx, name in ((1, "slab"), (2, "generated")):
            xd = np.concatenate([g[idx] for g in gathered])
            h = results[name][1]
            assert h.isconverged and h.niters == hs.niters and h.mvps == hs.mvps, (h.niters, hs.niters)
            hist_err = float(np.max(np.abs(h["resnorm"] - hs["resnorm"]) / hs["resnorm"]))
            x_err = float(np.linalg.norm(xd - xs) / np.linalg.norm(xs))
            assert hist_err <= 1e-10 and x_err <= 1e-10, (name, hist_err, x_err)
            print(f"{name}: world={world} N={N} iters={h.niters} hist_err={hist_err:.2e} x_err={x_err:.2e}")
        xs, hs = isb.cg_(np.full(n, 0.25), Ag, b_global, Pl=isb.JacobiPrec(Ag.diag(), ctx1), log=True, reltol=1e-9)
        for idx, name in ((3, "slab_pcg"), (4, "generated_pcg")):
            xd = np.concatenate([g[idx] for g in gathered])
            h = results[name][1]
            assert h.niters == hs.niters and h.mvps == hs.mvps, (name, h.niters, hs.niters)
            hist_err = float(np.max(np.abs(h["resnorm"] - hs["resnorm"]) / hs["resnorm"]))
            x_err = float(np.linalg.norm(xd - xs) / np.linalg.norm(xs))
            assert hist_err <= 1e-10 and x_err <= 1e-10, (name, hist_err, x_err)
            print(f"{name}: iters={h.niters} hist_err={hist_err:.2e} x_err={x_err:.2e}")
        # minres / gmres / bicgstabl: partitioned == single GPU
        xs, hs = isb.minres_(np.zeros(n), Ag, b_global, initially_zero=True, log=True, reltol=1e-9)
        xd = np.concatenate([o[0] for o in others])
        assert hm.niters == hs.niters and np.max(np.abs(hm["resnorm"] - hs["resnorm"]) / hs["resnorm"]) <= 1e-8
        assert np.linalg.norm(xd - xs) <= 1e-8 * np.linalg.norm(xs)
        xs, hs = isb.gmres_(np.zeros(n), Ag, b_global, initially_zero=True, log=True, restart=20, maxiter=60, orth_meth="dgks")
        xd = np.concatenate([o[1] for o in others])
        assert hg.niters == hs.niters and hg.mvps == hs.mvps
        assert np.max(np.abs(hg["resnorm"] - hs["resnorm"]) / hs["resnorm"]) <= 1e-8
        assert np.linalg.norm(xd - xs) <= 1e-8 * np.linalg.norm(xs)
        xs, hs = isb.bicgstabl_(np.zeros(n), Ag, b_global, 2, initial_zero=True, log=True, max_mv_products=120,
                                r_shadow=rsh.copy(), reltol=1e-9)
        xd = np.concatenate([o[2] for o in others])
        assert hb.niters == hs.niters and hb.mvps == hs.mvps
        k = min(5, hs.niters)
        assert np.max(np.abs(hb["resnorm"][:k] - hs["resnorm"][:k]) / hs["resnorm"][:k]) <= 1e-7
        xs, hs = isb.chebyshev_(np.zeros(n), Ag, b_global, 3 * lam1[0], 3 * lam1[-1], initially_zero=True, log=True,
                                maxiter=12, reltol=1e-12)
        xd = np.concatenate([o[4] for o in others])
        assert hc.niters == hs.niters == 12 and np.max(np.abs(hc["resnorm"] - hs["resnorm"]) / hs["resnorm"]) <= 1e-10
        assert np.linalg.norm(xd - xs) <= 1e-10 * np.linalg.norm(xs)
        rs = isb.lobpcg(Ag, False, X0.copy(), maxiter=6, _fixed_iterations=True)
        Xd = np.concatenate([o[3] for o in others], axis=0)
        assert rl.iterations == rs.iterations
        assert np.max(np.abs(rl.lam - rs.lam) / np.abs(rs.lam)) <= 1e-9
        assert np.max(np.abs(rl.residual_norms - rs.residual_norms) / rs.residual_norms) <= 1e-6
        assert np.max(np.abs(Xd.T @ Xd - np.eye(4))) <= 1e-10
        sgn = np.sign(np.sum(Xd * rs.X, axis=0))
        assert np.max(np.abs(Xd * sgn[None, :] - rs.X)) <= 1e-6
        print(f"minres/gmres/bicgstabl/lobpcg partitioned == single: iters {hm.niters}/{hg.niters}/{hb.niters}/{rl.iterations}")
        print("DIST_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
