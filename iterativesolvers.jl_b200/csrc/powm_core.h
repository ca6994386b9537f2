// powm_core.h -- powm!(B, x; shift, inverse, tol, maxiter) / invpowm!(B, x; shift, ...) of reference src/simple.jl:118-151,
// :186 (PowerMethodIterable :6-15, iterate :29-48, powm_iterable! :53-56) as fused passes (pass_core.h).  B is a device CSR
// operator or a callback (for inverse iteration the caller's B applies inv(A - shift I), src/simple.jl:83-88); the Rayleigh
// quotient, the residual norm and the stopping test stay in device memory.
//
//   S    Ax = B x                                                   :32
//   P1   theta = <x, Ax> ; ||Ax||^2                                 :35, :46
//   P2   r = Ax - theta x ; ||r|| ; x = Ax / ||Ax|| ; done          :38-46
// (the residual vector itself is not kept: the reference stores it in the iterable but returns only its norm)
#pragma once
#include "pass_core.h"

namespace b200 {

struct PowmScal {
  double theta, residual, tol, inv;
  double sum[2];
  double *hist;
  long long hist_cap, n_hist, iteration, maxiter;
  int done, breakdown;
};

template <typename T>
struct PowmDots {
  static constexpr int NRED = 2;
  const T *x, *Ax;
  PowmScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const double a = (double)Ax[i];
    acc[0] += (double)x[i] * a;                              // dot(p.x, p.Ax) :35
    acc[1] += a * a;                                         // norm(p.x) after copyto!(p.x, p.Ax) :45-46
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->theta = tot[0];
    s->inv = (double)((T)1 / (T)sqrt(tot[1]));               // one(eltype(p.x)) / norm(p.x) :46
  }
};

template <typename T>
struct PowmUpdate {
  static constexpr int NRED = 1;
  T *x;
  const T *Ax;
  PowmScal *s;
  T theta, inv;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    theta = (T)s->theta;
    inv = (T)s->inv;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    const T a = Ax[i];
    const T r = a - theta * x[i];                            // copyto!(p.r, p.Ax) ; axpy!(-theta, p.x, p.r) :38-39
    acc[0] += (double)r * (double)r;
    x[i] = a * inv;                                          // copyto!(p.x, p.Ax) ; rmul!(p.x, 1 / norm(p.x)) :45-46
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    PowmScal *q = s;
    q->residual = sqrt(tot[0]);                              // :42
    if (!(q->residual == q->residual)) q->breakdown = 1;
    if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = q->residual;
    q->n_hist += 1;
    q->iteration += 1;                                       // :48
    q->done = (q->iteration > q->maxiter) || (q->residual <= q->tol) || q->breakdown;   // done :27 (sic: >, not >=)
  }
};

struct PowmOutcome {
  int64_t iters, n_hist;
  double theta, residual, tol;
  int converged, breakdown;
};

template <typename T, typename B>
int powm_run(B &be, const typename B::Op *A, int64_t n, int64_t n_global, T *x, double tol, int64_t maxiter,
             int check_every, int64_t hist_cap, double *hist_host, PowmOutcome *out) {
  if (tol < 0) tol = eps_of<T>() * (double)n_global * (double)n_global * (double)n_global;   // eps(real(eltype(B))) * size(B, 2)^3 :119
  if (maxiter < 0) maxiter = n_global;                                                        // size(B, 1) :120
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter + 1) hist_cap = maxiter + 1;
  const size_t vb = ((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256;
  const size_t hb = ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(vb + 256 + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *Ax = (T *)p; p += vb;
  PowmScal *s = (PowmScal *)p; p += 256;
  static_assert(sizeof(PowmScal) <= 256, "PowmScal outgrew its slot");
  double *hist = hist_cap ? (double *)p : nullptr;

  PowmScal h;
  memset(&h, 0, sizeof(h));
  h.tol = tol;
  h.maxiter = maxiter;
  h.residual = sizeof(T) == 8 ? 1.7976931348623157e308 : 3.4028234663852886e38;   // floatmax(real(T)) :55
  h.hist = hist;
  h.hist_cap = hist_cap;
  h.done = (0 > maxiter) || (h.residual <= tol);                                  // done(p, start(p)) :27
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= maxiter + 1) break;
    const int64_t batch = check_every < maxiter + 1 - enqueued ? check_every : maxiter + 1 - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      if ((st = be.apply(A, x, Ax))) return st;                                   // S :32
      if ((st = be.pass(PowmDots<T>{x, Ax, s}, n))) return st;                     // P1
      if ((st = be.pass(PowmUpdate<T>{x, Ax, s, (T)0, (T)0}, n))) return st;       // P2
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  out->iters = h.iteration;
  out->theta = h.theta;
  out->residual = h.residual;
  out->tol = h.tol;
  out->converged = h.residual <= h.tol;                                           // converged :23
  out->breakdown = h.breakdown;
  out->n_hist = h.n_hist < hist_cap ? h.n_hist : hist_cap;
  if (out->n_hist > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}

}  // namespace b200
