// chebyshev_core.h -- chebyshev!(x, A, b, lmin, lmax; abstol, reltol, Pl, maxiter, initially_zero) of reference
// src/chebyshev.jl:131-160 (chebyshev_iterable! :59-92, iterate :29-57) written as fused passes (pass_core.h) for GENERAL
// operators and preconditioners (`b200_linop` callbacks: `mul!(y, A, x)`, `ldiv!(y, Pl, x)`).  For a `b200_csr` with
// Identity / Jacobi the engine of chebyshev.cu is the fast path.  The method has no inner products besides the residual
// norm: alpha and beta follow a data-independent recurrence that is advanced in the scalar section of the residual pass.
//
//   L   c = Pl \ r                         callback (Jacobi: fused into U)                             :37
//   U   u = c (second call) or c + beta c  (sic, :45)                                                  :39-46
//   S   c = A u                                                                                        :48
//   X   x += alpha u ; r -= alpha c ; ||r|| ; next alpha, beta ; done                                  :51-54
// Restated literally, including `iteration == 1` being the SECOND call (start = 0, :26).
#pragma once
#include "pass_core.h"

namespace b200 {

struct ChebScal {
  double alpha, beta, l_avg, l_diff;
  double resnorm, tol, abstol, reltol;
  double sum[2];
  double *hist;
  long long hist_cap, n_hist, iteration, maxiter;
  int done, breakdown, copy_mode, pad;
};

// alpha / beta of the iteration about to run (:39-46), in the arithmetic of real(T) as the reference
template <typename T>
B200_HD void cheb_prepare(ChebScal *q) {
  if (q->iteration == 1) {                                   // :39
    q->alpha = (double)((T)2 / (T)q->l_avg);                 // :40
    q->beta = 0.0;
    q->copy_mode = 1;                                        // copyto!(u, c) :41
  } else {
    const T h = (T)q->l_diff * (T)q->alpha / (T)2;
    const T beta = h * h;                                    // :43
    q->alpha = (double)((T)1 / ((T)q->l_avg - beta));        // :44
    q->beta = (double)beta;
    q->copy_mode = 0;
  }
}

template <typename T>
struct ChebInit {
  static constexpr int NRED = 1;
  const T *b, *ax;             // ax = A*x or nullptr (initially_zero)
  T *r, *u;
  ChebScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    T v = b[i];                                              // copyto!(r, b) :70
    if (ax) v = v - ax[i];                                   // r .-= c :81
    r[i] = v;
    u[i] = (T)0;                                             // zero(x) :71
    acc[0] += (double)v * (double)v;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    ChebScal *q = s;
    q->resnorm = sqrt(tot[0]);                               // :83
    q->tol = fmax(q->reltol * q->resnorm, q->abstol);        // :84
    q->alpha = 0.0;                                          // zero(real(T)) :89
    q->iteration = 0;                                        // start :26
    q->breakdown = !(q->resnorm == q->resnorm);
    q->done = (q->iteration >= q->maxiter) || (q->resnorm <= q->tol) || q->breakdown;   // done :27
    cheb_prepare<T>(q);
  }
};

template <typename T>
struct ChebU {
  static constexpr int NRED = 0;
  const T *z, *diag;           // z = r (Identity / Jacobi, with diag) or c (callback preconditioner)
  T *u;
  const ChebScal *s;
  T beta;
  int copy_mode;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    beta = (T)s->beta;
    copy_mode = s->copy_mode;
  }
  B200_HD void elem(int64_t i, double *) const {
    T c = z[i];
    if (diag) c = c / diag[i];                               // ldiv!(c, Pl, r) :37
    u[i] = copy_mode ? c : c + beta * c;                     // :41 / :45 (sic)
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

template <typename T>
struct ChebX {
  static constexpr int NRED = 1;
  T *x, *r;
  const T *u, *c;
  ChebScal *s;
  T alpha;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { alpha = (T)s->alpha; }
  B200_HD void elem(int64_t i, double *acc) const {
    x[i] = x[i] + alpha * u[i];                              // axpy!(alpha, u, x) :51
    const T ri = r[i] - alpha * c[i];                        // axpy!(-alpha, c, r) :52
    r[i] = ri;
    acc[0] += (double)ri * (double)ri;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    ChebScal *q = s;
    q->resnorm = sqrt(tot[0]);                               // :54
    if (!(q->resnorm == q->resnorm)) q->breakdown = 1;
    if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = q->resnorm;
    q->n_hist += 1;
    q->iteration += 1;
    q->done = (q->iteration >= q->maxiter) || (q->resnorm <= q->tol) || q->breakdown;
    cheb_prepare<T>(q);
  }
};

struct ChebOutcome {
  int64_t iters, mvps, n_hist;
  double residual, tol;
  int converged, breakdown;
};

template <typename T, typename B>
int chebyshev_run(B &be, const typename B::Op *A, const typename B::Op *Pl, const T *diag, int64_t n, int64_t n_global,
                  T *x, const T *b, double lmin, double lmax, double abstol, double reltol, int64_t maxiter,
                  int initially_zero, int check_every, int64_t hist_cap, double *hist_host, ChebOutcome *out) {
  if (reltol < 0) reltol = sqrt(eps_of<T>());                               // :137
  if (maxiter < 0) maxiter = n_global;                                      // :139
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;                               // reserve!(history, :resnorm, maxiter) :144
  const size_t vb = ((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256;
  const size_t hb = ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(3 * vb + 512 + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *r = (T *)p; p += vb;
  T *u = (T *)p; p += vb;
  T *c = (T *)p; p += vb;
  ChebScal *s = (ChebScal *)p; p += 512;
  static_assert(sizeof(ChebScal) <= 512, "ChebScal outgrew its slot");
  double *hist = hist_cap ? (double *)p : nullptr;

  ChebScal h;
  memset(&h, 0, sizeof(h));
  h.l_avg = (lmax + lmin) / 2;                                              // :65
  h.l_diff = (lmax - lmin) / 2;                                             // :66
  h.abstol = abstol;
  h.reltol = reltol;
  h.maxiter = maxiter;
  h.hist = hist;
  h.hist_cap = hist_cap;
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;
  int64_t mvps0 = 0;
  if (!initially_zero) {                                                    // :78-82
    if ((st = be.apply(A, x, c))) return st;
    mvps0 = 1;
  }
  if ((st = be.pass(ChebInit<T>{b, initially_zero ? nullptr : c, r, u, s}, n))) return st;

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= maxiter) break;
    const int64_t batch = check_every < maxiter - enqueued ? check_every : maxiter - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      if (Pl) {
        if ((st = be.apply(Pl, r, c))) return st;                                                // L :37
        if ((st = be.pass(ChebU<T>{c, nullptr, u, s, (T)0, 0}, n))) return st;                    // U
      } else {
        if ((st = be.pass(ChebU<T>{r, diag, u, s, (T)0, 0}, n))) return st;
      }
      if ((st = be.apply(A, u, c))) return st;                                                   // S :48
      if ((st = be.pass(ChebX<T>{x, r, u, c, s, (T)0}, n))) return st;                            // X
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  out->iters = h.iteration;
  out->mvps = mvps0 + h.iteration;
  out->residual = h.resnorm;
  out->tol = h.tol;
  out->converged = h.resnorm <= h.tol;                                      // converged :23
  out->breakdown = h.breakdown;
  out->n_hist = h.iteration < hist_cap ? h.iteration : hist_cap;
  if (out->n_hist > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}

}  // namespace b200
