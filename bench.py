#!/usr/bin/env python
"""bench.py -- CG iterations/s and effective SpMV GB/s on the 7-point 3-D Laplacian (BASELINE.json).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...          # the reference's CPU algorithm (oracle) on host cores

One "step" = one cg!(x, A, b) call that runs exactly --iters iterations (convergence test disabled so
every step does the same work) of laplace_matrix(Float64, GRID, 3), GRID=512 by default (the size the
metric is quoted on; 17 GB, fits one B200).  With N GPUs the SAME global problem is row-partitioned
into N z-slabs (strong scaling).
  value : iterations/s with A, b, x resident in HBM (timed with CUDA events on the library's stream,
          barrier + synchronize on both sides, max over ranks).
  e2e   : the same metric through the public host-buffer API: every step uploads the operator from
          host arrays (SparseMatrixCSC{Float64,Int64} arrays at N=1, the rank's CSR slab at N>1) and
          b, x from host memory, solves, and copies x back.
  roofline : the dominant kernel (fused SpMV+dot) from CUDA-event brackets recorded inside the timed
          region (b200_ctx_profile_*), algorithmic bytes nnz*12 + (n+1)*4 + 2*n*8 per launch.
  cpu_baseline : the oracle's single-threaded restatement of the reference CPU path on a bounded
          sample (rank 0, N=1 only).
Inputs are larger than L2 (126 MB) by two orders of magnitude, so no explicit L2 flush is needed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cg_iterations_per_second_3d_laplacian_fp64"
UNIT = "iterations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="N of laplace_matrix(Float64, N, 3)")
    ap.add_argument("--iters", type=int, default=200, help="CG iterations per step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-maxiter", type=int, default=20000)
    ap.add_argument("--comm", type=int, default=0, help="multi-GPU collectives: 0 auto, 1 NCCL, 2 NVLink peer memory")
    ap.add_argument("--cpu-iters", type=int, default=4, help="iterations of the CPU baseline sample")
    return ap.parse_args()


def rhs_slab(row_begin, m_local):
    """b[i] for global rows [row_begin, row_begin+m_local): a counter-based pseudo-random value per
    GLOBAL index, so 1/2/4/8-GPU runs see identical data."""
    i = np.arange(row_begin, row_begin + m_local, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B97F4A7C15)) ^ (i >> np.uint64(29))
    h = (h * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(32)
    return (h >> np.uint64(11)).astype(np.float64) / float(1 << 53) - 0.5


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(n, nnz):
    V, I = 8, 4
    spmv = nnz * (V + I) + (n + 1) * I + 2 * n * V
    cg = nnz * (V + I) + (n + 1) * I + 11 * n * V
    return spmv, cg


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(grid, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K2 launch, from the committed `ncu --set full` capture
    (profiles/k2_traffic.json names the .ncu-rep).  Only valid for the workload it was captured on."""
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            t = json.load(f)
        if int(t["grid"]) == grid and int(t["n_gpus"]) == world:
            return int(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path (oracle restatement of
    SparseArrays' CSC SpMV + src/cg.jl iterate; Julia itself is not available in this image).
    Single-threaded, as the reference is.  Each step is a bounded sample: --cpu-iters iterations."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import iterativesolvers_jl_b200 as isb
    from oracle import oracle
    N = args.grid
    t0 = time.time()
    colptr, rowval, nzval, shape = isb.laplace_matrix(np.float64, N, 3, base=1)
    n = shape[0]
    b = rhs_slab(0, n)
    b /= np.linalg.norm(b)
    x, u, c = np.zeros(n), np.zeros(n), np.zeros(n)
    r = b.copy()
    res, prev = C.c_double(float(np.linalg.norm(r))), C.c_double(1.0)
    L = oracle.lib()
    its = max(1, args.cpu_iters // 2)

    def step():
        L.oracle_cg_steps_f64(C.c_int64(n), oracle._p(colptr), oracle._p(rowval), oracle._p(nzval), C.c_int64(1),
                              oracle._p(x), oracle._p(r), oracle._p(u), oracle._p(c), C.byref(res), C.byref(prev),
                              C.c_int64(its))
    for _ in range(args.warmup):
        step()
    t1 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t1
    value = args.steps * its / dt
    nnz = int(colptr[-1] - 1)
    # NOT reference behaviour (the reference's SpMV and broadcasts are serial): the same steps with every loop
    # spread over all host threads, valid because this A is symmetric.  Reported beside the reference number so
    # that the GPU is also compared with what the host could do at best; never used as `value`.
    thr_its = 4 * its
    t2 = time.time()
    L.oracle_cg_steps_f64_omp(C.c_int64(n), oracle._p(colptr), oracle._p(rowval), oracle._p(nzval), C.c_int64(1),
                              oracle._p(x), oracle._p(r), oracle._p(u), oracle._p(c), C.byref(res), C.byref(prev),
                              C.c_int64(thr_its))
    thr = thr_its / (time.time() - t2)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cg! 7-pt 3-D Laplacian n={N}^3 fp64 (SparseMatrixCSC{{Float64,Int64}} CPU path)",
                   "grid": N, "n": n, "nnz": nnz, "iters_per_step": its},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{its} CG iterations per step on the full n={N}^3 matrix, single thread "
                                   f"(reference behaviour); host has {os.cpu_count()} cores; setup {t1 - t0:.1f}s",
                         "threaded_variant_not_reference_behaviour": {
                             "value": thr, "unit": UNIT, "cores": os.cpu_count(),
                             "what": "OpenMP row-parallel gather SpMV (CSC read as CSR: symmetric A only) + "
                                     "parallel fused vector updates"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local_rank)
    import iterativesolvers_jl_b200 as isb
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx = isb.Context.distributed(local_rank)
    else:
        ctx = isb.Context(local_rank)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    N = args.grid
    n = N ** 3
    planes = [N * r // world for r in range(world + 1)]            # z-slabs (contiguous rows)
    row_offsets = np.array([p * N * N for p in planes], dtype=np.int64)
    row_begin, m_local = int(row_offsets[rank]), int(row_offsets[rank + 1] - row_offsets[rank])
    plan = None
    if world > 1:
        plan = isb.HaloPlan(rank, world, row_offsets).scan_laplacian(N, 3).exchange()
    A = isb.B200CSR.laplacian(N, 3, np.float64, row_begin, m_local, plan, ctx)
    nnz_global = 7 * N ** 3 - 6 * N ** 2
    b_host = rhs_slab(row_begin, m_local)
    nrm2 = ctx.allreduce([float(np.dot(b_host, b_host))])[0]
    b_host /= np.sqrt(nrm2)
    b = isb.DeviceArray.from_numpy(ctx, b_host)
    x = isb.DeviceArray.zeros(ctx, m_local)
    L = isb.lib()
    if world > 1:
        isb._lib.check(L.b200_ctx_set_option(ctx._h, b"comm", args.comm))
    peer_ok = C.c_int64(0)
    L.b200_ctx_get_option(ctx._h, b"peer_ok", C.byref(peer_ok))
    comm_name = "none" if world == 1 else ("nccl" if args.comm == 1 or not peer_ok.value else "nvlink-peer-memory")

    def step():
        L.b200_fill(ctx._h, m_local, 0.0, x._p, 0)
        isb.cg_(x, A, b, initially_zero=True, maxiter=args.iters, reltol=0.0, _fixed_iterations=True)

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    L.b200_ctx_profile_enable(ctx._h, 1)
    for s in range(4):
        L.b200_ctx_profile_read(ctx._h, s, None, None, 1)
    ctx.barrier()
    torch.cuda.synchronize()
    if clocks:
        clocks.start()
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    ctx.barrier()
    torch.cuda.synchronize()
    launches = ctx.launch_count() - launches0
    clk = clocks.stop() if clocks else None
    prof = []
    for s in range(4):
        t, c = C.c_double(), C.c_int64()
        L.b200_ctx_profile_read(ctx._h, s, C.byref(t), C.byref(c), 1)
        prof.append((t.value, c.value))
    L.b200_ctx_profile_enable(ctx._h, 0)
    ms = ctx.allreduce([ms], op="max")[0]
    total_iters = args.steps * args.iters
    value = total_iters / (ms / 1e3)
    k2_ms = prof[0][0] / max(prof[0][1], 1)
    k2_ms = ctx.allreduce([k2_ms], op="max")[0]

    # ---- end-to-end through the host-buffer API ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # host inputs live in page-locked memory (allocated through the C ABI), as the bench contract asks
        if world == 1:
            colptr, rowval, nzval, shape = isb.laplace_matrix(np.float64, N, 3, base=1, empty=isb.pinned_empty)
            h2d = colptr.nbytes + rowval.nbytes + nzval.nbytes + 2 * b_host.nbytes
        else:
            rp, ci, va = isb.laplace_csr_slab(np.float64, N, 3, row_begin, m_local, empty=isb.pinned_empty)
            h2d = rp.nbytes + ci.nbytes + va.nbytes + 2 * b_host.nbytes
        b_pinned = isb.pinned_empty(m_local)
        b_pinned[:] = b_host
        b_host = b_pinned
        xh = isb.pinned_empty(m_local)

        e2e_iters = []

        def e2e_step():
            # the call a user makes: operator from host arrays, then cg!(x, A, b) with the reference's default
            # tolerances (reltol = sqrt(eps), maxiter = n) -- it runs to convergence
            if world == 1:
                Ah = isb.B200CSR.from_csc_arrays(colptr, rowval, nzval, shape, base=1, ctx=ctx)
            else:
                Ah = isb.B200CSR.from_csr_slab(rp, ci, va, n, row_begin, 0, plan, ctx)
            xh[:] = 0.0
            isb.cg_(xh, Ah, b_host, initially_zero=True, maxiter=args.e2e_maxiter)
            e2e_iters.append(int(isb.cg_.last_result.iters))
            Ah.close()
        e2e_step()                                                  # warm-up
        e2e_iters.clear()
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        ctx.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = ctx.allreduce([dt], op="max")[0]
        e2e = {"value": sum(e2e_iters) / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(ctx.allreduce([float(h2d)])[0]),
               "d2h_bytes_per_step": int(ctx.allreduce([float(xh.nbytes)])[0]),
               "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "iterations_per_step": e2e_iters, "converged": bool(isb.cg_.last_result.isconverged),
               "includes": "operator upload (+CSC->CSR conversion at N=1), b/x H2D, cg! to reltol=sqrt(eps), "
                           "x D2H; host arrays in page-locked memory"}

    # ---- CPU baseline: oracle restatement, single thread, bounded sample (rank 0, N=1) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        if args.no_e2e:
            colptr, rowval, nzval, shape = isb.laplace_matrix(np.float64, N, 3, base=1)
        xo, uo, co = np.zeros(n), np.zeros(n), np.zeros(n)
        ro = b_host.copy()
        res, prev = C.c_double(float(np.linalg.norm(ro))), C.c_double(1.0)
        Lo = oracle.lib()
        t0 = time.perf_counter()
        Lo.oracle_cg_steps_f64(C.c_int64(n), oracle._p(colptr), oracle._p(rowval), oracle._p(nzval), C.c_int64(1),
                               oracle._p(xo), oracle._p(ro), oracle._p(uo), oracle._p(co), C.byref(res),
                               C.byref(prev), C.c_int64(args.cpu_iters))
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()                      # all host threads; NOT reference behaviour (see run_reference)
        Lo.oracle_cg_steps_f64_omp(C.c_int64(n), oracle._p(colptr), oracle._p(rowval), oracle._p(nzval), C.c_int64(1),
                                   oracle._p(xo), oracle._p(ro), oracle._p(uo), oracle._p(co), C.byref(res),
                                   C.byref(prev), C.c_int64(4 * args.cpu_iters))
        thr = 4 * args.cpu_iters / (time.perf_counter() - t0)
        cpu = {"value": args.cpu_iters / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "threaded_variant_not_reference_behaviour": {"value": thr, "unit": UNIT, "cores": os.cpu_count()},
               "sample": f"{args.cpu_iters} CG iterations on the full n={N}^3 SparseMatrixCSC{{Float64,Int64}} "
                         f"(single thread = reference behaviour; host has {os.cpu_count()} cores)",
               "spmv_gbs_csc_int64_accounting": None}

    if rank == 0:
        peak, peak_src = peaks()
        m_max = int(max(np.diff(row_offsets)))
        nnz_loc_max = 7 * m_max                                   # per-GPU share (upper bound, boundary planes have fewer)
        spmv_bytes, cg_bytes = algorithmic_bytes(m_max, A.nnz)
        spmv_gbs = spmv_bytes / (k2_ms / 1e3) / 1e9 if k2_ms > 0 else None
        it_ms = ms / total_iters
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cg! 7-pt 3-D Laplacian n={N}^3 fp64, row-partitioned over {world} GPU(s)",
                       "grid": N, "n": n, "nnz": nnz_global, "iters_per_step": args.iters,
                       "l2": "inputs (>=1.4 GB per GPU) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"row-slabs x{world}" if world > 1 else "single GPU",
                       "collectives": comm_name},
            "gpu_launches": int(launches),
            "effective_spmv_gbs": spmv_gbs,
            "ms_per_iteration": it_ms,
            "cg_step_algorithmic_gbs_per_gpu": cg_bytes / (it_ms / 1e3) / 1e9,
            "cg_step_frac_of_peak": cg_bytes / (it_ms / 1e3) / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": "k_cg_spmv_dot (c = A*u fused with dot(u,c))",
                         "achieved": spmv_gbs, "peak": peak, "unit": "GB/s",
                         "frac": (spmv_gbs / peak) if spmv_gbs else None, "traffic": ncu_traffic(N, world),
                         "peak_source": peak_src, "avg_launch_ms": k2_ms, "launches": prof[0][1],
                         "algorithmic_bytes_per_launch": spmv_bytes,
                         "other_kernels_ms": {"k3_r_update_nrm2": prof[1][0] / max(prof[1][1], 1),
                                              "k1_x_u_update": prof[2][0] / max(prof[2][1], 1)}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "clocks": clk,
        }
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
