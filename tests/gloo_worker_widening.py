"""World-size-2 gloo worker (CPU) for the multi-GPU form of the fused-pass engines (qmr!, lsqr!, lsmr!, idrs!, and the
general cg!, gmres!, minres!, bicgstabl!, chebyshev!, powm!, lobpcg): every rank owns a row slab of A (and of A'), the operator application is "gather the operand, multiply the
slab" (what the halo exchange + SpMV do on the GPUs) and every pass total goes through an allreduce before its scalar
section runs -- the control flow of the CUDA backend on a multi-GPU context (csrc/pass.cuh), with torch.distributed
in the place of NCCL.  Each rank compares its slab of the solution and the whole history with the single-process run."""
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, port = int(sys.argv[1]), sys.argv[2]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)
    from hostsim import sim
    rng = np.random.default_rng(99)
    n = 150
    M = (sp.random(n, n, 0.05, random_state=8, format="csr") + 6 * sp.eye(n, format="csr")).tocsr()
    Mt = M.T.tocsr()
    S = (M + Mt).tocsr()                                       # symmetric positive definite, for cg
    b = rng.random(n)
    P = np.asfortranarray(rng.random((n, 4)))
    dM = M.diagonal()
    Dinv = sp.diags(1.0 / dM).tocsr()                          # a preconditioner applied through the operator interface
    Sm = (S - 9.0 * sp.eye(n, format="csr")).tocsr()           # symmetric indefinite, for minres
    shadow = rng.random(n)
    DinvS = sp.diags(1.0 / S.diagonal()).tocsr()
    ds = 1.0 / np.sqrt(S.diagonal())
    evp = np.linalg.eigvalsh((sp.diags(ds) @ S @ sp.diags(ds)).toarray())
    lo, hi = 0.9 * evp[0], 1.1 * evp[-1]                       # bounds of the Jacobi-preconditioned spectrum, for chebyshev
    cuts = [0, 83, n]
    lo_, hi_ = cuts[rank], cuts[rank + 1]
    m = hi_ - lo_

    # ---- single-process references (every rank computes them; cheap)
    ref = {
        "qmr": sim.qmr_(np.zeros(n), M, b, initially_zero=True),
        "idrs": sim.idrs_(np.zeros(n), M, b, P),
        "lsqr": sim.lsqr_(np.zeros(n), M, b, maxiter=12, atol=0.0, btol=0.0, conlim=0.0),
        "lsmr": sim.lsmr_(np.zeros(n), M, b, maxiter=12, atol=0.0, btol=0.0, conlim=0.0),
        "cg": sim.cg_(np.zeros(n), S, b, initially_zero=True, diag=S.diagonal()),
        "gmres": sim.gmres_(np.zeros(n), M, b, pl_diag=dM, Pr=Dinv, restart=7, maxiter=40, orth_meth="dgks",
                            initially_zero=True),
        "minres": sim.minres_(np.zeros(n), Sm, b, maxiter=60, initially_zero=True),
        "bicgstabl": sim.bicgstabl_(np.zeros(n), M, b, 2, shadow, Pl=Dinv, max_mv_products=80, initial_zero=True),
        "chebyshev": sim.chebyshev_(np.zeros(n), S, b, lo, hi, Pl=DinvS, maxiter=60, initially_zero=True),
    }
    x0p = rng.random(n)
    x0p /= np.linalg.norm(x0p)
    th_ref, xp_ref, hp_ref = sim.powm_(S, x0p.copy(), tol=1e-9, maxiter=400)
    X0 = rng.random((n, 3))
    Yc = rng.random((n, 2))
    lob_ref = sim.lobpcg_general(S, False, X0, jac=S.diagonal(), C_=Yc, tol=1e-12, maxiter=25)

    # ---- row-partitioned runs
    slabs = {}                                                  # op_id -> the slab whose product the engine asks for

    def register(mat):
        c = sim.Csr(mat, np.float64)
        return c

    def apply(op_id, x_ptr, y_ptr):
        slab = slabs[op_id]
        x_loc = np.ctypeslib.as_array(C.cast(x_ptr, C.POINTER(C.c_double)), shape=(m,))
        mx = max(cuts[r + 1] - cuts[r] for r in range(2))               # gloo's all_gather wants equal sizes: pad
        mine = torch.zeros(mx, dtype=torch.float64)
        mine[:m] = torch.from_numpy(x_loc.copy())
        parts = [torch.zeros(mx, dtype=torch.float64) for _ in range(2)]
        dist.all_gather(parts, mine)                                    # the "halo exchange": here the whole operand
        x_full = np.concatenate([parts[r].numpy()[: cuts[r + 1] - cuts[r]] for r in range(2)])
        y = slab @ x_full
        np.ctypeslib.as_array(C.cast(y_ptr, C.POINTER(C.c_double)), shape=(m,))[:] = y
        return 0

    def allreduce(buf, count):
        a = np.ctypeslib.as_array(buf, shape=(count,))
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t)
        a[:] = t.numpy()

    # sim builds its Csr objects inside each call, so the slabs are looked up by CONTENT: patch Csr to register itself
    orig_csr = sim.Csr

    class RegCsr(orig_csr):
        def __init__(self, A, dtype):
            super().__init__(A, dtype)
            slabs[self.rowptr.ctypes.data] = sp.csr_matrix((self.vals, self.colind, self.rowptr), shape=self.shape)

    sim.Csr = RegCsr
    sim.set_dist(apply, allreduce)
    A_loc, At_loc, S_loc = M[lo_:hi_], Mt[lo_:hi_], S[lo_:hi_]
    out = {
        "qmr": sim.qmr_(np.zeros(m), A_loc, b[lo_:hi_], initially_zero=True, At=At_loc),
        "idrs": sim.idrs_(np.zeros(m), A_loc, b[lo_:hi_], P[lo_:hi_]),
        "lsqr": sim.lsqr_(np.zeros(m), A_loc, b[lo_:hi_], maxiter=12, atol=0.0, btol=0.0, conlim=0.0, At=At_loc),
        "lsmr": sim.lsmr_(np.zeros(m), A_loc, b[lo_:hi_], maxiter=12, atol=0.0, btol=0.0, conlim=0.0, At=At_loc),
        "cg": sim.cg_(np.zeros(m), S_loc, b[lo_:hi_], initially_zero=True, diag=S.diagonal()[lo_:hi_]),
        "gmres": sim.gmres_(np.zeros(m), A_loc, b[lo_:hi_], pl_diag=dM[lo_:hi_], Pr=Dinv[lo_:hi_], restart=7, maxiter=40,
                            orth_meth="dgks", initially_zero=True),
        "minres": sim.minres_(np.zeros(m), Sm[lo_:hi_], b[lo_:hi_], maxiter=60, initially_zero=True),
        "bicgstabl": sim.bicgstabl_(np.zeros(m), A_loc, b[lo_:hi_], 2, shadow[lo_:hi_], Pl=Dinv[lo_:hi_], max_mv_products=80,
                                    initial_zero=True),
        "chebyshev": sim.chebyshev_(np.zeros(m), S_loc, b[lo_:hi_], lo, hi, Pl=DinvS[lo_:hi_], maxiter=60, initially_zero=True),
    }
    th_d, xp_d, hp_d = sim.powm_(S_loc, x0p[lo_:hi_].copy(), tol=1e-9, maxiter=400)
    lob_d = sim.lobpcg_general(S_loc, False, X0[lo_:hi_], jac=S.diagonal()[lo_:hi_], C_=Yc[lo_:hi_], tol=1e-12, maxiter=25)
    sim.set_dist()
    sim.Csr = orig_csr

    # general LOBPCG (Jacobi preconditioner, constraint): every Gram entry is a pass total that goes through the allreduce
    assert lob_d["status"] == 0 and lob_ref["status"] == 0 and lob_d["iterations"] == lob_ref["iterations"] == 26
    assert np.abs(lob_d["lam"] - lob_ref["lam"]).max() <= 1e-9 * np.abs(lob_ref["lam"]).max()
    assert np.abs(lob_d["resnorm"] - lob_ref["resnorm"]).max() <= 1e-6 * np.abs(lob_ref["resnorm"]).max()
    ov = torch.from_numpy(np.sum(lob_d["X"] * lob_ref["X"][lo_:hi_], axis=0))
    dist.all_reduce(ov)                                        # <x_dist, x_ref> over the whole vectors: +-1 for the same Ritz vectors
    assert np.abs(np.abs(ov.numpy()) - 1).max() <= 1e-6
    assert hp_d.iters == hp_ref.iters and hp_d.converged and abs(th_d - th_ref) <= 1e-12 * abs(th_ref)
    assert np.linalg.norm(xp_d - xp_ref[lo_:hi_]) <= 1e-10
    for name in ("qmr", "idrs", "cg", "gmres", "minres", "bicgstabl", "chebyshev"):
        (xr, hr), (xd, hd) = ref[name], out[name]
        assert hd.iters == hr.iters and hd.converged == hr.converged and hd.mvps == hr.mvps, (name, hd.iters, hr.iters)
        # (minres on an indefinite operator amplifies the last-bit differences of another summation order by ~10x per
        # few iterations -- the same between two orders of the single-process run: compare the first 15 iterations)
        k = min(len(hr.hist), 15 if name == "minres" else 30)
        assert k > 3 and np.max(np.abs(hd.hist[:k] - hr.hist[:k])) <= 1e-10 * hr.hist[0], name
        assert np.linalg.norm(xd - xr[lo_:hi_]) <= (1e-2 if name == "minres" else 1e-10) * np.linalg.norm(xr), name
    for name in ("lsqr", "lsmr"):
        (xr, hr), (xd, hd) = ref[name], out[name]
        assert hd.iters == hr.iters == 12 and hd.istop == hr.istop == 7 and (hd.mvps, hd.mtvps) == (hr.mvps, hr.mtvps)
        for key in ("anorm", "rnorm", "cnorm"):
            assert np.max(np.abs(hd.hist[key][:8] - hr.hist[key][:8])) <= 1e-10 * np.max(np.abs(hr.hist[key][:8])), (name, key)
        assert np.linalg.norm(xd - xr[lo_:hi_]) <= 1e-8 * np.linalg.norm(xr), name
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")


if __name__ == "__main__":
    main()
