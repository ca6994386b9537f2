// minres_core.h -- minres!(x, A, b; skew_hermitian, abstol, reltol, maxiter, initially_zero) of reference
// src/minres.jl:200-237 (minres_iterable! :39-89, iterate :97-159) written as fused passes (pass_core.h) for a GENERAL
// operator: A may be a device callback (`b200_linop`: the reference's duck-typed `mul!(y, A, x)` contract,
// docs/src/getting_started.md:25-30).  For a `b200_csr` operator the specialised engine of minres.cu (Lanczos update
// fused into the SpMV epilogue) is the fast path; this one keeps what matters for a matrix-free operator: the Lanczos
// coefficients, both Givens rotations, the right-hand side and the stopping test stay in device memory, the host polls
// the done flag every `check_every` iterations.
//
//   S    v_next = A v_curr                                                                           :104
//   P1   v_next -= H[2] v_prev (iteration > 1) ; proj = <v_curr, v_next>                             :106-109
//   P2   v_next -= proj v_curr ; H[4] = ||v_next|| ; then the scalar section: previous rotations,    :111-135, :149-156
//        new rotation, rhs, |rhs[2]| = resnorm, done
//   P3   v_next *= 1/H[4] ; w_next = (v_curr - H[2] w_curr - H[1] w_prev) / H[3] ; x += rhs[1] w_next :115, :138-144
//   (vectors rotate by pointer on the host :147-148)
// Algorithmic bytes per iteration besides the operator: 3 + 3 + 7 = 13 n V (P3 reads v_next, v_curr, w_curr, w_prev, x
// and writes v_next, w_next, x: 8 on iterations > 2).
#pragma once
#include "pass_core.h"

namespace b200 {

struct MinresScal {
  double H[4];                 // the active column of the Hessenberg matrix, H[0..3] = reference H[1..4] :70
  double rhs[2];               // :71
  double c_prev, s_prev, c_curr, s_curr;                     // :77-78
  double resnorm, tol, abstol, reltol;
  double proj;
  // what P3 of the current iteration applies (captured before the state moves on to the next iteration)
  double w_h1, w_h0, w_div, x_coef, v_inv;
  double sum[2];
  double *hist;
  long long hist_cap, n_hist;
  long long iteration, maxiter;                              // iteration starts at 1 (start(::MINRESIterable) = 1 :93)
  long long stamp;                                           // host index of the last iteration whose P2 ran
  int done, breakdown, skew, pad;
};

// ---- setup :49-74: v_curr = b - A x ; resnorm ; tol ; rhs
template <typename T>
struct MinresInit {
  static constexpr int NRED = 1;
  const T *b, *ax;             // ax = A*x (nullptr when initially_zero)
  T *v;
  MinresScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    T r = b[i];                                              // copyto!(v_curr, b) :49
    if (ax) r = r - ax[i];                                   // axpy!(-one(T), v_next, v_curr) :62
    v[i] = r;
    acc[0] += (double)r * (double)r;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    MinresScal *q = s;
    q->resnorm = sqrt(tot[0]);                               // :65
    q->tol = fmax(q->reltol * q->resnorm, q->abstol);        // :66
    q->rhs[0] = q->resnorm;                                  // :71
    q->rhs[1] = 0.0;
    q->v_inv = 1.0 / q->resnorm;                             // :74
    q->breakdown = !(q->resnorm == q->resnorm);
    q->done = (q->iteration > q->maxiter) || (q->resnorm <= q->tol) || q->breakdown;   // done :95
  }
};

template <typename T>
struct MinresScaleV {          // v_curr .*= inv(resnorm) :74
  static constexpr int NRED = 0;
  T *v;
  const MinresScal *s;
  T inv;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { inv = (T)s->v_inv; }
  B200_HD void elem(int64_t i, double *) const { v[i] = v[i] * inv; }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- P1
template <typename T>
struct MinresP1 {
  static constexpr int NRED = 1;
  T *vnext;
  const T *vprev, *vcurr;
  MinresScal *s;
  T h1;
  int first;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    first = s->iteration <= 1;
    h1 = (T)s->H[1];
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T vn = vnext[i];
    if (!first) {
      vn = vn - h1 * vprev[i];                               // axpy!(-H[2], v_prev, v_next) :106
      vnext[i] = vn;
    }
    acc[0] += (double)vcurr[i] * (double)vn;                 // dot(v_curr, v_next) :109
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->proj = tot[0];
    s->H[2] = tot[0];                                        // H[3] = real(proj) (skew: proj) :110
  }
};

// ---- P2
template <typename T>
struct MinresP2 {
  static constexpr int NRED = 1;
  T *vnext;
  const T *vcurr;
  MinresScal *s;
  long long my_iter;           // host index of this iteration
  T proj;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { proj = (T)s->proj; }
  B200_HD void elem(int64_t i, double *acc) const {
    const T vn = vnext[i] - proj * vcurr[i];                 // axpy!(-proj, v_curr, v_next) :111
    vnext[i] = vn;
    acc[0] += (double)vn * (double)vn;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    MinresScal *q = s;
    double *H = q->H;
    H[3] = sqrt(tot[0]);                                     // H[4] = norm(v_next) :114
    q->v_inv = 1.0 / H[3];                                   // :115
    if (q->iteration > 2) {                                  // rotation on H[1] and H[2] :118-121
      H[0] = q->s_prev * H[1];
      H[1] = q->c_prev * H[1];
    }
    if (q->iteration > 1) {                                  // rotation on H[2] and H[3] :124-128
      const double tmp = -q->s_curr * H[1] + q->c_curr * H[2];
      H[1] = q->c_curr * H[1] + q->s_curr * H[2];
      H[2] = tmp;
    }
    double c, sn, r;
    givens_real(H[2], H[3], c, sn, r);                       // :131
    H[2] = r;
    q->rhs[1] = -sn * q->rhs[0];                             // :134
    q->rhs[0] = c * q->rhs[0];                               // :135
    // what P3 applies :138-144
    q->w_h1 = H[1];
    q->w_h0 = H[0];
    q->w_div = H[2];
    q->x_coef = q->rhs[0];
    // the state of the next iteration :149-156
    q->c_prev = q->c_curr;
    q->s_prev = q->s_curr;
    q->c_curr = c;
    q->s_curr = sn;
    q->rhs[0] = q->rhs[1];                                   // :150
    H[1] = q->skew ? -H[3] : H[3];                           // :153
    q->resnorm = fabs(q->rhs[1]);                            // :156
    if (!(q->resnorm == q->resnorm)) q->breakdown = 1;
    if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = q->resnorm;
    q->n_hist += 1;
    q->iteration += 1;
    q->stamp = my_iter;
    q->done = (q->iteration > q->maxiter) || (q->resnorm <= q->tol) || q->breakdown;
  }
};

// ---- P3 (runs for exactly the iterations whose P2 ran: the one that sets `done` still updates x)
template <typename T>
struct MinresP3 {
  static constexpr int NRED = 0;
  T *vnext, *wnext, *x;
  const T *vcurr, *wcurr, *wprev;
  const MinresScal *s;
  long long my_iter;
  T vinv, h1, h0, winv, xc;
  int use1, use0;
  B200_HD bool skip() const { return s->stamp != my_iter; }
  B200_HD void load() {
    vinv = (T)s->v_inv;
    h1 = (T)s->w_h1;
    h0 = (T)s->w_h0;
    winv = (T)1 / (T)s->w_div;
    xc = (T)s->x_coef;
    const long long it = s->iteration - 1;                   // the iteration P2 just finished
    use1 = it > 1;
    use0 = it > 2;
  }
  B200_HD void elem(int64_t i, double *) const {
    vnext[i] = vnext[i] * vinv;                              // v_next .*= inv(H[4]) :115
    T w = vcurr[i];                                          // copyto!(w_next, v_curr) :138
    if (use1) w = w - h1 * wcurr[i];                         // :139
    if (use0) w = w - h0 * wprev[i];                         // :140
    w = w * winv;                                            // w_next .*= inv(H[3]) :141
    wnext[i] = w;
    x[i] = x[i] + xc * w;                                    // axpy!(rhs[1], w_next, x) :144
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

struct MinresOutcome {
  int64_t iters, mvps, n_hist;
  double residual, tol;
  int converged, breakdown, done, pad;
};

// ---- the driver, in resumable pieces: layout of the scratch, setup (minres_iterable! :39-89), advance (up to k calls of
// iterate :97-159), collect.  minres_run is the one-shot form; the iterator of the C ABI keeps the scratch between calls.
template <typename T>
struct MinresLayout {
  T *v[3], *w[3];              // roles at iteration 1: v[0] = v_prev, v[1] = v_curr, v[2] = v_next (same for w)
  MinresScal *s;
  double *hist;
  int64_t hist_cap;
};

inline size_t minres_vec_bytes(size_t elem, int64_t n) { return ((elem * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256; }
template <typename T>
size_t minres_ws_bytes(int64_t n, int64_t hist_cap) {
  return 6 * minres_vec_bytes(sizeof(T), n) + 512 + ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
}
template <typename T>
MinresLayout<T> minres_layout(void *ws, int64_t n, int64_t hist_cap) {
  static_assert(sizeof(MinresScal) <= 512, "MinresScal outgrew its slot");
  const size_t vb = minres_vec_bytes(sizeof(T), n);
  MinresLayout<T> L;
  char *p = (char *)ws;
  for (int i = 0; i < 3; ++i) { L.v[i] = (T *)p; p += vb; }
  for (int i = 0; i < 3; ++i) { L.w[i] = (T *)p; p += vb; }
  L.s = (MinresScal *)p; p += 512;
  L.hist = hist_cap > 0 ? (double *)p : nullptr;
  L.hist_cap = hist_cap > 0 ? hist_cap : 0;
  return L;
}

template <typename T, typename B>
int minres_setup(B &be, const typename B::Op *A, const MinresLayout<T> &L, int64_t n, int64_t n_global, T *x, const T *b,
                 double abstol, double reltol, int64_t maxiter, int initially_zero, int skew_hermitian, int64_t *mvps0) {
  if (reltol < 0) reltol = sqrt(eps_of<T>());                               // :205
  if (maxiter < 0) maxiter = n_global;                                      // :206
  MinresScal h;
  memset(&h, 0, sizeof(h));
  h.abstol = abstol;
  h.reltol = reltol;
  h.maxiter = maxiter;
  h.iteration = 1;
  h.stamp = -1;
  h.c_prev = h.c_curr = 1.0;                                                // :77-78
  h.hist = L.hist;
  h.hist_cap = L.hist_cap;
  h.skew = skew_hermitian != 0;
  int st;
  if ((st = be.to_device(L.s, &h, sizeof(h)))) return st;
  *mvps0 = 0;
  if (!initially_zero) {                                                    // :58-63
    if ((st = be.apply(A, x, L.v[2]))) return st;
    *mvps0 = 1;
  }
  if ((st = be.pass(MinresInit<T>{b, initially_zero ? nullptr : L.v[2], L.v[1], L.s}, n))) return st;
  return be.pass(MinresScaleV<T>{L.v[1], L.s, (T)0}, n);
}

// up to k more iterations (k < 0: until done).  Everything the loop needs besides the scratch is read back from the device.
template <typename T, typename B>
int minres_advance(B &be, const typename B::Op *A, const MinresLayout<T> &L, int64_t n, T *x, int64_t k, int check_every) {
  int st;
  MinresScal h;
  if ((st = be.to_host(&h, L.s, sizeof(h)))) return st;
  if (h.done) return 0;
  const int64_t it0 = h.iteration - 1;                                      // iterations performed so far
  const int64_t left = h.maxiter - it0;
  const int64_t todo = (k < 0 || k > left) ? left : k;
  T *v[3], *w[3];
  for (int i = 0; i < 3; ++i) {                                             // the pointer rotation :147-148, it0 times
    v[i] = L.v[(i + it0) % 3];
    w[i] = L.w[(i + it0) % 3];
  }
  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&L.s->done, &done))) return st;
    if (done || enqueued >= todo) break;
    const int64_t batch = check_every < todo - enqueued ? check_every : todo - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      const long long me = (long long)(it0 + enqueued + q);
      if ((st = be.apply(A, v[1], v[2]))) return st;                                             // S  :104
      if ((st = be.pass(MinresP1<T>{v[2], v[0], v[1], L.s, (T)0, 0}, n))) return st;              // P1
      if ((st = be.pass(MinresP2<T>{v[2], v[1], L.s, me, (T)0}, n))) return st;                   // P2
      MinresP3<T> p3{v[2], w[2], x, v[1], w[1], w[0], L.s, me, (T)0, (T)0, (T)0, (T)0, (T)0, 0, 0};
      if ((st = be.pass(p3, n))) return st;                                                      // P3
      T *t = v[0]; v[0] = v[1]; v[1] = v[2]; v[2] = t;                                           // :147
      t = w[0]; w[0] = w[1]; w[1] = w[2]; w[2] = t;                                              // :148
    }
    enqueued += batch;
  }
  return 0;
}

// state of the iteration; the history entries recorded since hist_from (an iteration count) go to hist_host
template <typename T, typename B>
int minres_collect(B &be, const MinresLayout<T> &L, int64_t mvps0, double *hist_host, MinresOutcome *out) {
  int st;
  MinresScal h;
  if ((st = be.to_host(&h, L.s, sizeof(h)))) return st;
  out->iters = h.iteration - 1;
  out->mvps = mvps0 + (h.iteration - 1);                                    // nextiter!(history, mvps = 1) :227
  out->residual = h.resnorm;
  out->tol = h.tol;
  out->converged = h.resnorm <= h.tol;                                      // converged :91
  out->breakdown = h.breakdown;
  out->done = h.done;
  out->n_hist = h.n_hist < L.hist_cap ? h.n_hist : L.hist_cap;
  if (hist_host && out->n_hist > 0 && (st = be.to_host(hist_host, L.hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}

// start a new history window (the iterator returns the residuals of each call separately)
template <typename B>
int minres_reset_window(B &be, MinresScal *s) {
  const long long zero = 0;
  return be.to_device(&s->n_hist, &zero, sizeof(zero));
}

template <typename T, typename B>
int minres_run(B &be, const typename B::Op *A, int64_t n, int64_t n_global, T *x, const T *b, double abstol,
               double reltol, int64_t maxiter, int initially_zero, int skew_hermitian, int check_every, int64_t hist_cap,
               double *hist_host, MinresOutcome *out) {
  if (maxiter < 0) maxiter = n_global;
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;                               // reserve!(history, :resnorm, maxiter) :212
  void *ws = nullptr;
  int st = be.workspace(minres_ws_bytes<T>(n, hist_cap), &ws);
  if (st) return st;
  const MinresLayout<T> L = minres_layout<T>(ws, n, hist_cap);
  int64_t mvps0 = 0;
  if ((st = minres_setup<T>(be, A, L, n, n_global, x, b, abstol, reltol, maxiter, initially_zero, skew_hermitian, &mvps0)))
    return st;
  if ((st = minres_advance<T>(be, A, L, n, x, -1, check_every))) return st;
  return minres_collect<T>(be, L, mvps0, hist_host, out);
}

}  // namespace b200
