// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libb200krylov.so, never imported by the
// product package).
//
// A serial CPU backend for the fused-pass engines (csrc/pass_core.h): it runs the SAME pass functors, scalar
// sections and driver loops that libb200krylov.so instantiates with the CUDA backend (csrc/pass.cuh), so that the
// algorithmic content of those engines -- everything except the generic k_pass kernel and the SpMV kernels, which the
// GPU tests cover -- is checked against the oracle on machines without a GPU.
//
//   order:  0 = rows ascending, 1 = rows descending.  A pass whose element update of row i touched any other
//           row would give order-dependent vectors; the tests run both orders and compare.
//   split:  0 = finish(tot) right after the reduction (single-GPU form), 1 = totals stored to sums() and
//           finish(sums()) called separately (the multi-GPU form: allreduce between the two).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../iterativesolvers.jl_b200/csrc/pass_core.h"
#include "../../iterativesolvers.jl_b200/csrc/qmr_core.h"
#include "../../iterativesolvers.jl_b200/csrc/lsqr_core.h"
#include "../../iterativesolvers.jl_b200/csrc/lsmr_core.h"
#include "../../iterativesolvers.jl_b200/csrc/idrs_core.h"
#include "../../iterativesolvers.jl_b200/csrc/cg_core.h"
#include "../../iterativesolvers.jl_b200/csrc/gmres_core.h"
#include "../../iterativesolvers.jl_b200/csrc/chebyshev_core.h"
#include "../../iterativesolvers.jl_b200/csrc/powm_core.h"
#include "../../iterativesolvers.jl_b200/csrc/stationary_core.h"
#include "../../iterativesolvers.jl_b200/csrc/minres_core.h"
#include "../../iterativesolvers.jl_b200/csrc/bicgstabl_core.h"
#include "../../iterativesolvers.jl_b200/csrc/lobpcg_constraint_core.h"
#include "../../iterativesolvers.jl_b200/csrc/svdl_core.h"
#include "../../iterativesolvers.jl_b200/csrc/lobpcg_general_core.h"

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {

struct HostCsr {          // 0-based CSR, int64 row pointers, int32 columns
  int64_t m, n;
  const int64_t *rowptr;
  const int32_t *colind;
  const void *vals;
  int is_f64;
};

// Row-partitioned runs (one process per "GPU", torch.distributed/gloo in the test): the operator application and the
// reduction of the pass sums are delegated to the host program, exactly where the CUDA backend calls the halo
// exchange + SpMV and ncclAllReduce.
typedef int (*hs_apply_cb)(const void *op_id, const void *x, void *y);
typedef void (*hs_allreduce_cb)(double *buf, int count);
static hs_apply_cb g_apply = nullptr;
static hs_allreduce_cb g_allreduce = nullptr;

struct HostBackend {
  typedef HostCsr Op;
  int order = 0, split = 0;
  std::vector<char> ws;
  long passes = 0, applies = 0;

  bool single() const { return !split; }

  int apply(const Op *A, const void *x, void *y) {
    ++applies;
    if (g_apply) return g_apply((const void *)A->rowptr, x, y);
    if (A->is_f64) {
      const double *v = (const double *)A->vals, *xx = (const double *)x;
      double *yy = (double *)y;
      for (int64_t i = 0; i < A->m; ++i) {
        double t = 0.0;
        for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) t += v[k] * xx[A->colind[k]];
        yy[i] = t;
      }
    } else {
      const float *v = (const float *)A->vals, *xx = (const float *)x;
      float *yy = (float *)y;
      for (int64_t i = 0; i < A->m; ++i) {
        float t = 0.0f;
        for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) t += v[k] * xx[A->colind[k]];
        yy[i] = t;
      }
    }
    return 0;
  }

  template <typename P>
  int pass(const P &p_in, int64_t n) {
    ++passes;
    if (p_in.skip()) return 0;
    P p = p_in;
    p.load();
    double acc[P::NRED > 0 ? P::NRED : 1];
    for (int j = 0; j < (P::NRED > 0 ? P::NRED : 1); ++j) acc[j] = 0.0;
    if (order == 0) for (int64_t i = 0; i < n; ++i) p.elem(i, acc);
    else for (int64_t i = n - 1; i >= 0; --i) p.elem(i, acc);
    if (P::NRED > 0) {
      if (!split && !g_allreduce) p.finish(acc);
      else {
        double *out = p.sums();
        for (int j = 0; j < P::NRED; ++j) out[j] = acc[j];
        if (g_allreduce) g_allreduce(out, P::NRED);
        if (!p.skip()) p.finish(p.sums());
      }
    }
    return 0;
  }
  template <typename P>
  int scalar(const P &p) {
    if (p.skip()) return 0;
    p.finish(p.sums());
    return 0;
  }
  int zero(void *x, size_t bytes) { memset(x, 0, bytes); return 0; }
  int copy(void *dst, const void *src, size_t bytes) { if (dst != src) memmove(dst, src, bytes); return 0; }
  int to_device(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
  int to_host(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
  int read_flag(const int *flag, int *out) { *out = *flag; return 0; }
  int workspace(size_t bytes, void **out) {
    // poison the scratch so that reads of never-written storage show up as NaNs in the results
    ws.assign(bytes + 256, (char)0xff);
    *out = (void *)(((uintptr_t)ws.data() + 255) / 256 * 256);
    return 0;
  }
};

}  // namespace

struct hostsim_csr {
  int64_t m, n;
  const int64_t *rowptr;
  const int32_t *colind;
  const void *vals;
};
struct hostsim_out {
  int64_t iters, mvps, mtvps, n_hist;
  double resnorm, tol;
  int32_t converged, breakdown;
  int64_t passes, applies;
};

EXPORT void hostsim_set_dist(hs_apply_cb a, hs_allreduce_cb r) {
  g_apply = a;
  g_allreduce = r;
}

static HostCsr mk(const hostsim_csr *a, int is_f64) { return HostCsr{a->m, a->n, a->rowptr, a->colind, a->vals, is_f64}; }

EXPORT int hostsim_qmr(int is_f64, const hostsim_csr *A, const hostsim_csr *At, void *x, const void *b, double abstol,
                       double reltol, int64_t maxiter, int initially_zero, int check_every, int64_t hist_cap,
                       double *hist, int order, int split, hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), at = mk(At, is_f64);
  b200::QmrOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::qmr_run<double>(be, &a, &at, A->m, A->n, (double *)x, (const double *)b, abstol, reltol,
                                          maxiter, initially_zero, check_every, hist_cap, hist, &o)
                  : b200::qmr_run<float>(be, &a, &at, A->m, A->n, (float *)x, (const float *)b, abstol, reltol,
                                         maxiter, initially_zero, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = o.mtvps; out->n_hist = o.n_hist;
  out->resnorm = o.resnorm; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

struct hostsim_ls_out {
  int64_t iters, mvps, mtvps, n_hist, hist_stride;
  int32_t istop, converged, bad_x, early;
  double atol, btol, ctol;
  double est[5];   // lsqr: Anorm, Acond, rnorm, Arnorm, xnorm ; lsmr: normr, normAr, normA, condA, normx
  int64_t passes, applies;
};

EXPORT int hostsim_lsqr(int is_f64, const hostsim_csr *A, const hostsim_csr *At, void *x, const void *b, double damp,
                        double atol, double btol, double conlim, int64_t maxiter, int check_every, int64_t hist_cap,
                        double *hist, int order, int split, hostsim_ls_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), at = mk(At, is_f64);
  b200::LsqrOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::lsqr_run<double>(be, &a, &at, A->m, At->m, (double *)x, (const double *)b, damp, atol, btol,
                                           conlim, maxiter, check_every, hist_cap, hist, &o)
                  : b200::lsqr_run<float>(be, &a, &at, A->m, At->m, (float *)x, (const float *)b, damp, atol, btol,
                                          conlim, maxiter, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = o.mtvps; out->n_hist = o.n_hist; out->hist_stride = o.hist_stride;
  out->istop = o.istop; out->converged = o.converged; out->bad_x = o.bad_x; out->early = o.early;
  out->atol = o.atol; out->btol = o.btol; out->ctol = o.ctol;
  out->est[0] = o.anorm; out->est[1] = o.acond; out->est[2] = o.rnorm; out->est[3] = o.arnorm; out->est[4] = o.xnorm;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

EXPORT int hostsim_lsmr(int is_f64, const hostsim_csr *A, const hostsim_csr *At, void *x, const void *b, double lambda,
                        double atol, double btol, double conlim, int64_t maxiter, int check_every, int64_t hist_cap,
                        double *hist, int order, int split, hostsim_ls_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), at = mk(At, is_f64);
  b200::LsmrOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::lsmr_run<double>(be, &a, &at, A->m, At->m, (double *)x, (const double *)b, lambda, atol, btol,
                                           conlim, maxiter, check_every, hist_cap, hist, &o)
                  : b200::lsmr_run<float>(be, &a, &at, A->m, At->m, (float *)x, (const float *)b, lambda, atol, btol,
                                          conlim, maxiter, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = o.mtvps; out->n_hist = o.n_hist; out->hist_stride = o.hist_stride;
  out->istop = o.istop; out->converged = o.converged; out->bad_x = 0; out->early = o.early;
  out->atol = o.atol; out->btol = o.btol; out->ctol = o.ctol;
  out->est[0] = o.normr; out->est[1] = o.normAr; out->est[2] = o.normA; out->est[3] = o.condA; out->est[4] = o.normx;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

EXPORT int hostsim_idrs(int is_f64, const hostsim_csr *A, void *x, const void *b, int s_dim, const void *P, int64_t ldp,
                        const void *diag, double abstol, double reltol, int64_t maxiter, int smoothing, int check_every,
                        int64_t hist_cap, double *hist, int order, int split, hostsim_out *out, const hostsim_csr *Pl) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), pl;
  if (Pl) pl = mk(Pl, is_f64);
  const HostCsr *plp = Pl ? &pl : nullptr;
  b200::IdrsOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::idrs_run<double>(be, &a, A->m, A->n, (double *)x, (const double *)b, s_dim, (const double *)P,
                                           ldp, (const double *)diag, abstol, reltol, maxiter, smoothing, check_every,
                                           hist_cap, hist, &o, plp)
                  : b200::idrs_run<float>(be, &a, A->m, A->n, (float *)x, (const float *)b, s_dim, (const float *)P, ldp,
                                          (const float *)diag, abstol, reltol, maxiter, smoothing, check_every,
                                          hist_cap, hist, &o, plp);
  out->iters = o.iters; out->mvps = o.iters; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.normR; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// cg! / pcg on a general operator: Pl (may be NULL) is a second "operator" whose application is y = Pl \ x
EXPORT int hostsim_cg(int is_f64, const hostsim_csr *A, const hostsim_csr *Pl, const void *diag, void *x, const void *b,
                      double abstol, double reltol, int64_t maxiter, int initially_zero, int check_every,
                      int64_t hist_cap, double *hist, int order, int split, hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), p;
  if (Pl) p = mk(Pl, is_f64);
  b200::CgpOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::cgp_run<double>(be, &a, Pl ? &p : nullptr, (const double *)diag, A->m, A->n, (double *)x,
                                          (const double *)b, abstol, reltol, maxiter, initially_zero, check_every,
                                          hist_cap, hist, &o)
                  : b200::cgp_run<float>(be, &a, Pl ? &p : nullptr, (const float *)diag, A->m, A->n, (float *)x,
                                         (const float *)b, abstol, reltol, maxiter, initially_zero, check_every,
                                         hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// gmres! on general operators: Pl / Pr (may be NULL) are "operators" whose application is y = P \ x; pl_diag / pr_diag
// Jacobi diagonals
EXPORT int hostsim_gmres(int is_f64, const hostsim_csr *A, const hostsim_csr *Pl, const hostsim_csr *Pr, const void *pl_diag,
                         const void *pr_diag, void *x, const void *b, double abstol, double reltol, int restart,
                         int64_t maxiter, int initially_zero, int orth_meth, int64_t hist_cap, double *hist, int order,
                         int split, hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), pl, pr;
  if (Pl) pl = mk(Pl, is_f64);
  if (Pr) pr = mk(Pr, is_f64);
  b200::GmresOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::gmres_run<double>(be, &a, Pl ? &pl : nullptr, Pr ? &pr : nullptr, (const double *)pl_diag,
                                            (const double *)pr_diag, A->m, A->n, (double *)x, (const double *)b, abstol,
                                            reltol, restart, maxiter, initially_zero, orth_meth, hist_cap, hist, &o)
                  : b200::gmres_run<float>(be, &a, Pl ? &pl : nullptr, Pr ? &pr : nullptr, (const float *)pl_diag,
                                           (const float *)pr_diag, A->m, A->n, (float *)x, (const float *)b, abstol,
                                           reltol, restart, maxiter, initially_zero, orth_meth, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// chebyshev! on general operators
EXPORT int hostsim_chebyshev(int is_f64, const hostsim_csr *A, const hostsim_csr *Pl, const void *diag, void *x,
                             const void *b, double lmin, double lmax, double abstol, double reltol, int64_t maxiter,
                             int initially_zero, int check_every, int64_t hist_cap, double *hist, int order, int split,
                             hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), p;
  if (Pl) p = mk(Pl, is_f64);
  b200::ChebOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::chebyshev_run<double>(be, &a, Pl ? &p : nullptr, (const double *)diag, A->m, A->n, (double *)x,
                                                (const double *)b, lmin, lmax, abstol, reltol, maxiter, initially_zero,
                                                check_every, hist_cap, hist, &o)
                  : b200::chebyshev_run<float>(be, &a, Pl ? &p : nullptr, (const float *)diag, A->m, A->n, (float *)x,
                                               (const float *)b, lmin, lmax, abstol, reltol, maxiter, initially_zero,
                                               check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// powm! on a general operator (for inverse iteration the "operator" applies inv(A - shift I))
EXPORT int hostsim_powm(int is_f64, const hostsim_csr *A, void *x, double tol, int64_t maxiter, int check_every,
                        int64_t hist_cap, double *hist, int order, int split, hostsim_out *out, double *theta) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64);
  b200::PowmOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::powm_run<double>(be, &a, A->m, A->n, (double *)x, tol, maxiter, check_every, hist_cap, hist, &o)
                  : b200::powm_run<float>(be, &a, A->m, A->n, (float *)x, tol, maxiter, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.iters; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  *theta = o.theta;
  return st;
}

// jacobi! / gauss_seidel! / sor! / ssor! on a CSR matrix with sorted rows (method 0..3); returns 0, or i+1 for the
// SingularException of DiagonalIndices; *levels_f / *levels_b: number of dependency levels of the sweeps
EXPORT int64_t hostsim_stationary(int is_f64, const hostsim_csr *A, void *x, const void *b, int method, double omega,
                                  int64_t maxiter, int order, int *levels_f, int *levels_b, long *passes) {
  HostBackend be;
  be.order = order;
  HostCsr a = mk(A, is_f64);
  auto body = [&](auto tag) -> int64_t {
    typedef decltype(tag) T;
    b200::StLevels lv;
    const int base = method & ~b200::ST_DENSE_ARITHMETIC;
    const bool fwd = base != b200::ST_JACOBI, bwd = base == b200::ST_SSOR && !(method & b200::ST_DENSE_ARITHMETIC);
    const int64_t sing = b200::stationary_analyse<T, int64_t>(a.m, a.rowptr, a.colind, (const T *)a.vals, fwd, bwd, &lv);
    if (sing) return sing;
    *levels_f = fwd ? (int)lv.lptr_f.size() - 1 : 0;
    *levels_b = bwd ? (int)lv.lptr_b.size() - 1 : 0;
    b200::CsrView<T, int64_t> view{a.m, a.rowptr, a.colind, (const T *)a.vals};
    const int st = b200::stationary_run<T, int64_t>(be, view, lv, lv.dpos.data(), lv.rows_f.data(), lv.rows_b.data(), (T *)x,
                                                    (const T *)b, method, omega, maxiter);
    *passes = be.passes;
    return st ? -1 : 0;
  };
  return is_f64 ? body(double()) : body(float());
}

// minres! on a general operator
EXPORT int hostsim_minres(int is_f64, const hostsim_csr *A, void *x, const void *b, double abstol, double reltol,
                          int64_t maxiter, int initially_zero, int skew, int check_every, int64_t hist_cap, double *hist,
                          int order, int split, hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64);
  b200::MinresOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::minres_run<double>(be, &a, A->m, A->n, (double *)x, (const double *)b, abstol, reltol, maxiter,
                                             initially_zero, skew, check_every, hist_cap, hist, &o)
                  : b200::minres_run<float>(be, &a, A->m, A->n, (float *)x, (const float *)b, abstol, reltol, maxiter,
                                            initially_zero, skew, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown;
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// bicgstabl! on general operators: Pl (may be NULL) is an "operator" whose application is y = Pl \ x; diag: Jacobi
EXPORT int hostsim_bicgstabl(int is_f64, const hostsim_csr *A, const hostsim_csr *Pl, const void *diag, void *x,
                             const void *b, const void *shadow, int l, double abstol, double reltol, int64_t max_mv,
                             int initial_zero, int check_every, int64_t hist_cap, double *hist, int order, int split,
                             hostsim_out *out) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), p;
  if (Pl) p = mk(Pl, is_f64);
  b200::BcgOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::bicgstabl_run<double>(be, &a, Pl ? &p : nullptr, (const double *)diag, A->m, A->n, (double *)x,
                                                (const double *)b, (const double *)shadow, l, abstol, reltol, max_mv,
                                                initial_zero, check_every, hist_cap, hist, &o)
                  : b200::bicgstabl_run<float>(be, &a, Pl ? &p : nullptr, (const float *)diag, A->m, A->n, (float *)x,
                                               (const float *)b, (const float *)shadow, l, abstol, reltol, max_mv,
                                               initial_zero, check_every, hist_cap, hist, &o);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = 0; out->n_hist = o.n_hist;
  out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged; out->breakdown = o.breakdown | (o.singular << 1);
  out->passes = be.passes; out->applies = be.applies;
  return st;
}

// The resumable forms (what csrc/iterables.cu does with the CUDA backend): setup once, then advance `chunk` iterations
// per call with a fresh history window each time, until done.  kind: 1 gmres, 2 minres, 3 bicgstabl, 4 cg.  The histories of
// the calls are concatenated into hist; calls = number of advance calls made.
template <typename T>
static int chunked_impl(int kind, HostBackend &be, const HostCsr *a, const HostCsr *pl, const HostCsr *pr, const T *pl_diag,
                        const T *pr_diag, int64_t n, int64_t n_global, T *x, const T *b, const T *shadow, int l, int restart,
                        int orth, int skew, double abstol, double reltol, int64_t maxiter, int zero, int64_t chunk,
                        int64_t hist_cap, double *hist, hostsim_out *out, int *calls) {
  const int64_t W = 7;                                    // a small window, so that the reset is exercised
  int st;
  void *ws = nullptr;
  int64_t nh = 0;
  std::vector<double> win((size_t)W);
  *calls = 0;
  if (kind == 1) {
    if (restart <= 0) restart = (int)(n_global < 20 ? n_global : 20);
    if ((st = be.workspace(b200::gmres_ws_bytes<T>(n, restart, W), &ws))) return st;
    const b200::GmresLayout<T> L = b200::gmres_layout<T>(ws, n, restart, W);
    const b200::GmresOps<T, HostBackend> op{a, pl, pr, pl_diag, pr_diag};
    int64_t mv = 0;
    if ((st = b200::gmres_setup<T, HostBackend>(be, op, L, n, n_global, x, b, abstol, reltol, restart, maxiter, zero, &mv))) return st;
    for (;;) {
      b200::GmresOutcome o;
      memset(&o, 0, sizeof(o));
      if ((st = b200::gmres_reset_window(be, L.s))) return st;
      if ((st = b200::gmres_advance<T, HostBackend>(be, op, L, n, x, b, orth, chunk, &mv))) return st;
      if ((st = b200::gmres_collect<T, HostBackend>(be, L, mv, win.data(), &o))) return st;
      *calls += 1;
      for (int64_t i = 0; i < o.n_hist && nh < hist_cap; ++i) hist[nh++] = win[(size_t)i];
      out->iters = o.iters; out->mvps = o.mvps; out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged;
      out->breakdown = o.breakdown;
      if (o.done || o.breakdown) break;
    }
  } else if (kind == 2) {
    if ((st = be.workspace(b200::minres_ws_bytes<T>(n, W), &ws))) return st;
    const b200::MinresLayout<T> L = b200::minres_layout<T>(ws, n, W);
    int64_t mv0 = 0;
    if ((st = b200::minres_setup<T, HostBackend>(be, a, L, n, n_global, x, b, abstol, reltol, maxiter, zero, skew, &mv0))) return st;
    for (;;) {
      b200::MinresOutcome o;
      memset(&o, 0, sizeof(o));
      if ((st = b200::minres_reset_window(be, L.s))) return st;
      if ((st = b200::minres_advance<T, HostBackend>(be, a, L, n, x, chunk, 3))) return st;
      if ((st = b200::minres_collect<T, HostBackend>(be, L, mv0, win.data(), &o))) return st;
      *calls += 1;
      for (int64_t i = 0; i < o.n_hist && nh < hist_cap; ++i) hist[nh++] = win[(size_t)i];
      out->iters = o.iters; out->mvps = o.mvps; out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged;
      out->breakdown = o.breakdown;
      if (o.done || o.breakdown) break;
    }
  } else if (kind == 4) {
    if ((st = be.workspace(b200::cgp_ws_bytes<T>(n, W), &ws))) return st;
    const b200::CgpLayout<T> L = b200::cgp_layout<T>(ws, n, W);
    int64_t mv0 = 0;
    if ((st = b200::cgp_setup<T, HostBackend>(be, a, pl != nullptr || pl_diag != nullptr, L, n, n_global, x, b, abstol, reltol,
                                              maxiter, zero, &mv0)))
      return st;
    for (;;) {
      b200::CgpOutcome o;
      memset(&o, 0, sizeof(o));
      if ((st = b200::cgp_reset_window(be, L.s))) return st;
      if ((st = b200::cgp_advance<T, HostBackend>(be, a, pl, pl_diag, L, n, x, chunk, 3))) return st;
      if ((st = b200::cgp_collect<T, HostBackend>(be, L, mv0, win.data(), &o))) return st;
      *calls += 1;
      for (int64_t i = 0; i < o.n_hist && nh < hist_cap; ++i) hist[nh++] = win[(size_t)i];
      out->iters = o.iters; out->mvps = o.mvps; out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged;
      out->breakdown = o.breakdown;
      if (o.done || o.breakdown) break;
    }
  } else {
    if ((st = be.workspace(b200::bicgstabl_ws_bytes<T>(n, l, W), &ws))) return st;
    const b200::BcgLayout<T> L = b200::bicgstabl_layout<T>(ws, n, l, W);
    if ((st = b200::bicgstabl_setup<T, HostBackend>(be, a, pl, pl_diag, L, n, n_global, x, b, l, abstol, reltol, maxiter, zero))) return st;
    for (;;) {
      b200::BcgOutcome o;
      memset(&o, 0, sizeof(o));
      if ((st = b200::bicgstabl_reset_window(be, L.s))) return st;
      if ((st = b200::bicgstabl_advance<T, HostBackend>(be, a, pl, pl_diag, L, n, x, shadow, l, chunk, 2))) return st;
      if ((st = b200::bicgstabl_collect<T, HostBackend>(be, L, win.data(), &o))) return st;
      *calls += 1;
      for (int64_t i = 0; i < o.n_hist && nh < hist_cap; ++i) hist[nh++] = win[(size_t)i];
      out->iters = o.iters; out->mvps = o.mvps; out->resnorm = o.residual; out->tol = o.tol; out->converged = o.converged;
      out->breakdown = o.breakdown | (o.singular << 1);
      if (o.done || o.breakdown || o.singular) break;
    }
  }
  out->mtvps = 0;
  out->n_hist = nh;
  out->passes = be.passes; out->applies = be.applies;
  return 0;
}

EXPORT int hostsim_chunked(int kind, int is_f64, const hostsim_csr *A, const hostsim_csr *Pl, const hostsim_csr *Pr,
                           const void *pl_diag, const void *pr_diag, void *x, const void *b, const void *shadow, int l,
                           int restart, int orth, int skew, double abstol, double reltol, int64_t maxiter, int zero,
                           int64_t chunk, int64_t hist_cap, double *hist, int order, int split, hostsim_out *out, int *calls) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), pl, pr;
  if (Pl) pl = mk(Pl, is_f64);
  if (Pr) pr = mk(Pr, is_f64);
  return is_f64 ? chunked_impl<double>(kind, be, &a, Pl ? &pl : nullptr, Pr ? &pr : nullptr, (const double *)pl_diag,
                                       (const double *)pr_diag, A->m, A->n, (double *)x, (const double *)b,
                                       (const double *)shadow, l, restart, orth, skew, abstol, reltol, maxiter, zero, chunk,
                                       hist_cap, hist, out, calls)
                : chunked_impl<float>(kind, be, &a, Pl ? &pl : nullptr, Pr ? &pr : nullptr, (const float *)pl_diag,
                                      (const float *)pr_diag, A->m, A->n, (float *)x, (const float *)b, (const float *)shadow,
                                      l, restart, orth, skew, abstol, reltol, maxiter, zero, chunk, hist_cap, hist, out, calls);
}

// Constraint (reference src/lobpcg.jl:144-224): factor Y'Y, optionally extend the factor by an identity block for the
// last `appended` columns (update! :188-206), and deflate the block X (strides rs, cs; bs <= 16 columns).
EXPORT int hostsim_constraint_apply(int is_f64, int64_t n, const void *Y, int64_t ldy, int nc, int appended, void *X,
                                    int64_t rs, int64_t cs, int bs, int order, int split) {
  HostBackend be;
  be.order = order;
  be.split = split;
  const int nc0 = nc - appended;
  std::vector<double> g_dev((size_t)(nc > 0 ? nc : 1) * b200::kConBlock), g_host(g_dev.size()), U0((size_t)nc0 * nc0);
  int st = is_f64 ? b200::constraint_factor<double>(be, (const double *)Y, ldy, nc0, n, g_dev.data(), g_host.data(), U0.data())
                  : b200::constraint_factor<float>(be, (const float *)Y, ldy, nc0, n, g_dev.data(), g_host.data(), U0.data());
  if (st) return st;
  std::vector<double> U((size_t)nc * nc, 0.0);
  for (int j = 0; j < nc0; ++j)
    for (int i = 0; i <= j; ++i) U[i + (size_t)j * nc] = U0[i + (size_t)j * nc0];
  for (int j = nc0; j < nc; ++j) U[j + (size_t)j * nc] = 1.0;
  return is_f64 ? b200::constraint_apply<double>(be, (const double *)Y, ldy, nc, U.data(), (double *)X, rs, cs, bs, n,
                                                 g_dev.data(), g_host.data())
                : b200::constraint_apply<float>(be, (const float *)Y, ldy, nc, U.data(), (float *)X, rs, cs, bs, n,
                                                g_dev.data(), g_host.data());
}

struct hostsim_svdl_out {
  int64_t iters, mvps, mtvps;
  int32_t converged, kdim;
  double beta;
};
EXPORT int hostsim_svdl(int is_f64, const hostsim_csr *A, const hostsim_csr *At, const void *v0, int nsv, int k, int jkeep,
                        double tol, double reltol, int64_t maxiter, int dolock, double *sigma, void *U, void *V,
                        double *hist_ritz, double *hist_resnorm, int *hist_conv, double *hist_betas, double *Bk,
                        int order, int split, hostsim_svdl_out *out, int method) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), at = mk(At, is_f64);
  b200::SvdlOutcome o;
  memset(&o, 0, sizeof(o));
  int st = is_f64 ? b200::svdl_run<double>(be, &a, &at, A->m, At->m, (const double *)v0, nsv, k, jkeep, tol, reltol, maxiter,
                                           dolock, sigma, (double *)U, A->m, (double *)V, At->m, hist_ritz, hist_resnorm,
                                           hist_conv, hist_betas, Bk, &o, method)
                  : b200::svdl_run<float>(be, &a, &at, A->m, At->m, (const float *)v0, nsv, k, jkeep, tol, reltol, maxiter,
                                          dolock, sigma, (float *)U, A->m, (float *)V, At->m, hist_ritz, hist_resnorm,
                                          hist_conv, hist_betas, Bk, &o, method);
  out->iters = o.iters; out->mvps = o.mvps; out->mtvps = o.mtvps; out->converged = o.converged; out->kdim = o.kdim;
  out->beta = o.beta;
  return st;
}

EXPORT void hostsim_dense_qr(int rows, int cols, const double *A, double *Q, double *R) {
  std::vector<double> a(A, A + (size_t)rows * cols), q, r;
  b200::dense_qr_thin(a, rows, cols, q, r);
  memcpy(Q, q.data(), sizeof(double) * q.size());
  memcpy(R, r.data(), sizeof(double) * r.size());
}

EXPORT void hostsim_dense_svd(int n, const double *A, double *U, double *S, double *V) {
  std::vector<double> a(A, A + (size_t)n * n), u, s, v;
  b200::dense_svd(a, n, u, s, v);
  memcpy(U, u.data(), sizeof(double) * (size_t)n * n);
  memcpy(S, s.data(), sizeof(double) * n);
  memcpy(V, v.data(), sizeof(double) * (size_t)n * n);
}

// general LOBPCG: A, optional B, optional callback preconditioner (as a matrix whose product is M \ x) or Jacobi diagonal,
// optional constraint Y (BY and the factor of Y'BY are formed here, as Constraint(Y, B, X) does, src/lobpcg.jl:161-186)
EXPORT int hostsim_lobpcg_general(int is_f64, const hostsim_csr *A, const hostsim_csr *Bm, const hostsim_csr *Pm,
                                  const void *jac, const void *Y, int nc, void *X, int sizeX, int largest, double tol,
                                  int64_t maxiter, int fixed, double *lambda, double *resnorm, int order, int split,
                                  int64_t *iterations, int *converged, int *status, double *trace_resnorm,
                                  double *trace_ritz, int64_t trace_cap) {
  HostBackend be;
  be.order = order;
  be.split = split;
  HostCsr a = mk(A, is_f64), b, pm;
  if (Bm) b = mk(Bm, is_f64);
  if (Pm) pm = mk(Pm, is_f64);
  const int64_t n = A->m;
  b200::LobpcgGenOutcome o;
  memset(&o, 0, sizeof(o));
  int st;
  auto body = [&](auto tag) -> int {
    typedef decltype(tag) T;
    std::vector<T> BY;
    std::vector<double> U((size_t)nc * nc), g_dev((size_t)(nc > 0 ? nc : 1) * b200::kConBlock), g_host(g_dev.size());
    const T *Yp = (const T *)Y, *BYp = Yp;
    if (nc > 0) {
      if (Bm) {                                   // BY = B * Y :166-167
        BY.resize((size_t)n * nc);
        for (int j = 0; j < nc; ++j) be.apply(&b, Yp + (size_t)j * n, BY.data() + (size_t)j * n);
        BYp = BY.data();
      }
      // gram = Y' BY ; cholesky :178-182
      for (int c0 = 0; c0 < nc; c0 += b200::kConBlock) {
        const int bs = nc - c0 < b200::kConBlock ? nc - c0 : b200::kConBlock;
        int s2 = b200::constraint_gram<T>(be, Yp, n, nc, BYp + (size_t)c0 * n, 1, n, bs, n, g_dev.data(), g_host.data());
        if (s2) return s2;
        for (int k = 0; k < nc; ++k)
          for (int j = 0; j < bs; ++j) U[k + (size_t)(c0 + j) * nc] = g_host[(size_t)k * b200::kConBlock + j];
      }
      if (b200::con_cholesky_upper(U.data(), nc)) return -5;
    }
    return b200::lobpcg_general_run<T>(be, &a, Bm ? &b : nullptr, Pm ? &pm : nullptr, (const T *)jac, Yp, BYp, n, nc,
                                       U.data(), (T *)X, n, sizeX, n, largest, tol, maxiter, fixed, lambda, resnorm, &o,
                                       trace_resnorm, trace_ritz, trace_cap);
  };
  st = is_f64 ? body(double()) : body(float());
  *iterations = o.iterations; *converged = o.converged; *status = o.status;
  return st;
}
