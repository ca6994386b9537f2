// qmr_core.h -- qmr!(x, A, b; abstol, reltol, maxiter, initially_zero) of reference src/qmr.jl:262-297
// (LanczosDecomp :5-100, QMRIterable :102-215) written as fused passes (pass_core.h).
//
// Per iteration (reference order of operations kept inside every element update):
//   S1  v_next = A v_curr                                                      :68   (SpMV)
//   P1  alpha = dot(v_next, w_curr)                                            :70   (2 reads)
//   P2  v_next -= alpha v_curr (+ beta_curr v_prev when iteration > 1)         :71-74 (3 reads, 1 write)
//   S2  w_next = A' w_curr                                                     :76   (SpMV with the adjoint operator)
//   P3  w_next -= alpha w_curr (+ delta w_prev) ; vw = dot(v_next, w_next)     :77-82 (4 reads, 1 write)
//       scalar section: delta, beta (:83-89), the two plane rotations and the new one (:173-187),
//       rhs update (:190-191), coefficients of the update, rotation/rhs shuffle (:205-208), residual (:212)
//   P4  v_next /= delta ; w_next /= beta ; p_m = (v_m - H2 p_curr - H1 p_prev)/H3 ; x += g1 p_m
//                                                                              :91-92, :196-202 (6 reads, 4 writes)
//   then the unconditional pointer rotations of :94-95 and :206-207 on the host.
// Algorithmic bytes per iteration: 2*(nnz*(V+4) + (n+1)*4 + 2*n*V) + 21*n*V.
//
// Deviation (documented in DESIGN.md): at an exact Lanczos breakdown (delta == 0) the reference returns
// from the Lanczos step before rotating its vectors (:84-86) and QMR then updates x with v_{m-1} instead
// of v_m (for A = I it reports resnorm 0 with x unchanged).  This engine performs that last update with
// v_m, stops, and reports status = B200_ERR_BREAKDOWN.
#pragma once
#include "pass_core.h"

namespace b200 {

struct QmrScal {
  double alpha, beta_prev, beta_curr, delta;          // LanczosDecomp scalars :16-20
  double resnorm, tol, abstol, reltol;
  double g1, g2;                                      // q.g :108
  double H1, H2, H3, H4;                              // q.H :109
  double c_prev, s_prev, c_curr, s_curr;              // :111-114
  double inv_res, inv_delta, inv_beta, p_h2, p_h1, inv_h3, x_coef;   // coefficients handed to the vector passes
  double sum[2];
  double *hist;
  long long iteration, maxiter, hist_cap, n_hist;     // iteration: the reference's counter, starts at 1 (:154)
  int use_p1, use_p2;                                 // iteration > 1 / > 2 at the time of the update (:197-198)
  int done, breakdown;
};

// ---- initialisation (LanczosDecomp constructor :24-60, qmr_iterable! :119-151)
template <typename T>
struct QmrInit {
  static constexpr int NRED = 1;
  const T *b, *ax;   // ax = A*x (nullptr when initially_zero)
  T *v_curr, *v_prev, *w_prev, *p_prev, *p_curr;
  QmrScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    T v = b[i];
    if (ax) v = v - ax[i];                            // axpy!(-one(T), v_next, v_curr) :38
    v_curr[i] = v;
    v_prev[i] = (T)0;
    w_prev[i] = (T)0;
    p_prev[i] = (T)0;
    p_curr[i] = (T)0;
    acc[0] += (double)v * (double)v;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    const double res = sqrt(tot[0]);                  // :40
    s->resnorm = res;
    s->inv_res = 1.0 / res;                           // :41
    s->tol = fmax(s->reltol * res, s->abstol);        // :142
    s->g1 = res; s->g2 = 0.0;                         // :131
    s->H1 = s->H2 = s->H3 = s->H4 = 0.0;
    s->c_prev = 1.0; s->s_prev = 0.0; s->c_curr = 1.0; s->s_curr = 0.0;   // :135-136
    s->alpha = s->beta_prev = s->beta_curr = s->delta = 0.0;              // :47-50
    s->iteration = 1;
    s->n_hist = 0;
    s->breakdown = !(res == res);
    s->done = (1 > s->maxiter) || (res <= s->tol) || s->breakdown;        // done() :155
  }
};

template <typename T>
struct QmrScaleInit {       // rmul!(v_curr, inv(resnorm)) :41 ; w_curr = copy(v_curr) :44
  static constexpr int NRED = 0;
  T *v_curr, *w_curr;
  const QmrScal *s;
  T inv;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { inv = (T)s->inv_res; }
  B200_HD void elem(int64_t i, double *) const {
    const T v = v_curr[i] * inv;
    v_curr[i] = v;
    w_curr[i] = v;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- P1
template <typename T>
struct QmrAlpha {
  static constexpr int NRED = 1;
  const T *v_next, *w_curr;
  QmrScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const { acc[0] += (double)v_next[i] * (double)w_curr[i]; }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const { s->alpha = tot[0]; }    // :70
};

// ---- P2
template <typename T>
struct QmrVNext {
  static constexpr int NRED = 0;
  T *v_next;
  const T *v_curr, *v_prev;
  const QmrScal *s;
  T alpha, beta;
  bool use_prev;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    alpha = (T)s->alpha;
    beta = (T)s->beta_curr;
    use_prev = s->iteration > 1;
  }
  B200_HD void elem(int64_t i, double *) const {
    T t = v_next[i] - alpha * v_curr[i];                                   // :71
    if (use_prev) t = t - beta * v_prev[i];                                // :72-74
    v_next[i] = t;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- P3 and the scalar section of the iteration
template <typename T>
struct QmrWNext {
  static constexpr int NRED = 1;
  T *w_next;
  const T *w_curr, *w_prev, *v_next;
  QmrScal *s;
  T alpha, delta;
  bool use_prev;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    alpha = (T)s->alpha;
    delta = (T)s->delta;       // the delta of the PREVIOUS iteration (finish() overwrites it afterwards)
    use_prev = s->iteration > 1;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T t = w_next[i] - alpha * w_curr[i];                                   // :77
    if (use_prev) t = t - delta * w_prev[i];                               // :78-80
    w_next[i] = t;
    acc[0] += (double)v_next[i] * (double)t;                               // :82
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    QmrScal *q = s;
    const long long it = q->iteration;
    const double vw = tot[0];
    q->delta = sqrt(fabs(vw));                                             // :83
    q->beta_prev = q->beta_curr;                                           // :88
    if (q->delta == 0.0 || !(vw == vw)) {                                  // :84-86 (see the header: deviation)
      q->breakdown = 1;
      q->beta_curr = 0.0;
      q->inv_delta = 1.0;
      q->inv_beta = 1.0;
    } else {
      q->beta_curr = vw / q->delta;                                        // :89
      q->inv_delta = 1.0 / q->delta;                                       // :91
      q->inv_beta = 1.0 / q->beta_curr;                                    // :92
    }
    // QMRIterable iterate :157-215
    q->H2 = q->beta_prev;                                                  // :168
    q->H3 = q->alpha;                                                      // :169
    q->H4 = q->delta;                                                      // :170
    if (it > 2) {                                                          // :173-176
      q->H1 = q->s_prev * q->H2;
      q->H2 = q->c_prev * q->H2;
    }
    if (it > 1) {                                                          // :179-183
      const double tmp = -q->s_curr * q->H2 + q->c_curr * q->H3;
      q->H2 = q->c_curr * q->H2 + q->s_curr * q->H3;
      q->H3 = tmp;
    }
    double c, sn, r;
    givens_real(q->H3, q->H4, c, sn, r);                                   // :187
    q->H3 = r;
    q->g2 = -sn * q->g1;                                                   // :190
    q->g1 = c * q->g1;                                                     // :191
    q->use_p1 = it > 1;                                                    // :197
    q->use_p2 = it > 2;                                                    // :198
    q->p_h2 = q->H2;
    q->p_h1 = q->H1;
    q->inv_h3 = 1.0 / q->H3;                                               // :199
    q->x_coef = q->g1;                                                     // :202
    q->c_prev = q->c_curr; q->s_prev = q->s_curr; q->c_curr = c; q->s_curr = sn;   // :205
    q->g1 = q->g2;                                                         // :208
    q->resnorm = fabs(q->g2);                                              // :212
    if (!(q->resnorm == q->resnorm)) q->breakdown = 1;
    if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = q->resnorm;
    q->n_hist += 1;
    q->iteration = it + 1;
    // P4 of THIS iteration must still run: `done` is published by the scalar step after it
  }
};

// ---- P4
template <typename T>
struct QmrUpdate {
  static constexpr int NRED = 0;
  T *v_next, *w_next;
  const T *v_m;        // the Lanczos vector of this iteration (v_curr before the rotation)
  const T *p_curr;
  T *p_prev_new;       // reads p_prev, writes p_m into the same storage (same index, same thread)
  T *x;
  const QmrScal *s;
  T inv_delta, inv_beta, h2, h1, inv_h3, xc;
  bool use_p1, use_p2;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    inv_delta = (T)s->inv_delta; inv_beta = (T)s->inv_beta;
    h2 = (T)s->p_h2; h1 = (T)s->p_h1; inv_h3 = (T)s->inv_h3; xc = (T)s->x_coef;
    use_p1 = s->use_p1 != 0; use_p2 = s->use_p2 != 0;
  }
  B200_HD void elem(int64_t i, double *) const {
    v_next[i] = v_next[i] * inv_delta;                                     // :91
    w_next[i] = w_next[i] * inv_beta;                                      // :92
    T p = v_m[i];                                                          // :196
    if (use_p1) p = p - h2 * p_curr[i];                                    // :197
    if (use_p2) p = p - h1 * p_prev_new[i];                                // :198
    p = p * inv_h3;                                                        // :199
    p_prev_new[i] = p;
    x[i] = x[i] + xc * p;                                                  // :202
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

B200_HD void qmr_publish_done(QmrScal *q) {        // done() :155, evaluated for the NEXT iteration
  if (q->done) return;
  q->done = (q->iteration > q->maxiter) || (q->resnorm <= q->tol) || q->breakdown;
}
typedef ScalarStep<QmrScal, qmr_publish_done> QmrDone;

struct QmrOutcome {
  int64_t iters, mvps, mtvps, n_hist;
  double resnorm, tol;
  int converged, breakdown;
};

// The driver: identical for the CUDA backend and for the serial test backend.
//   x, b: n values; A, At: the operator and its adjoint (same row partition); hist_dev: hist_cap doubles or NULL.
template <typename T, typename B>
int qmr_run(B &be, const typename B::Op *A, const typename B::Op *At, int64_t n, int64_t n_global, T *x, const T *b,
            double abstol, double reltol, int64_t maxiter, int initially_zero, int check_every, int64_t hist_cap,
            double *hist_host, QmrOutcome *out) {
  if (reltol < 0) reltol = sqrt(eps_of<T>());                               // :122
  if (maxiter < 0) maxiter = n_global;                                      // :123
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;
  const size_t vb = ((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256;
  const size_t hb = ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(8 * vb + 512 + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *v[3], *w[3], *pp[2];
  for (int i = 0; i < 3; ++i) { v[i] = (T *)p; p += vb; }
  for (int i = 0; i < 3; ++i) { w[i] = (T *)p; p += vb; }
  for (int i = 0; i < 2; ++i) { pp[i] = (T *)p; p += vb; }
  QmrScal *s = (QmrScal *)p; p += 512;
  double *hist = hist_cap ? (double *)p : nullptr;
  static_assert(sizeof(QmrScal) <= 512, "QmrScal outgrew its slot");
  T *v_prev = v[0], *v_curr = v[1], *v_next = v[2];
  T *w_prev = w[0], *w_curr = w[1], *w_next = w[2];
  T *p_prev = pp[0], *p_curr = pp[1];

  QmrScal h;
  memset(&h, 0, sizeof(h));
  h.abstol = abstol;
  h.reltol = reltol;
  h.maxiter = maxiter;
  h.hist_cap = hist_cap;
  h.hist = hist;
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;

  int64_t mvps = 0;
  if (!initially_zero) {                                                    // :35-39
    if ((st = be.apply(A, x, v_next))) return st;
    mvps = 1;
  }
  if ((st = be.pass(QmrInit<T>{b, initially_zero ? nullptr : v_next, v_curr, v_prev, w_prev, p_prev, p_curr, s}, n)))
    return st;
  if ((st = be.pass(QmrScaleInit<T>{v_curr, w_curr, s}, n))) return st;

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= maxiter) break;
    const int64_t batch = check_every < maxiter - enqueued ? check_every : maxiter - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      if ((st = be.apply(A, v_curr, v_next))) return st;                                          // S1
      if ((st = be.pass(QmrAlpha<T>{v_next, w_curr, s}, n))) return st;                            // P1
      if ((st = be.pass(QmrVNext<T>{v_next, v_curr, v_prev, s}, n))) return st;                    // P2
      if ((st = be.apply(At, w_curr, w_next))) return st;                                         // S2
      if ((st = be.pass(QmrWNext<T>{w_next, w_curr, w_prev, v_next, s}, n))) return st;            // P3
      if ((st = be.pass(QmrUpdate<T>{v_next, w_next, v_curr, p_curr, p_prev, x, s}, n))) return st; // P4
      if ((st = be.scalar(QmrDone{s}))) return st;
      T *t = v_prev; v_prev = v_curr; v_curr = v_next; v_next = t;          // :95
      t = w_prev; w_prev = w_curr; w_curr = w_next; w_next = t;             // :94
      t = p_prev; p_prev = p_curr; p_curr = t;                              // :206-207 (p_m was written over p_prev)
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  const int64_t iters = h.iteration - 1;
  out->iters = iters;
  out->mvps = mvps + iters;            // one A product per iteration
  out->mtvps = iters;                  // one A' product per iteration
  out->resnorm = h.resnorm;
  out->tol = h.tol;
  out->converged = h.resnorm <= h.tol;                                      // converged() :153
  out->breakdown = h.breakdown;
  out->n_hist = iters < hist_cap ? iters : hist_cap;
  if (out->n_hist > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}

}  // namespace b200
