// bicgstabl_core.h -- bicgstabl!(x, A, b, l; abstol, reltol, max_mv_products, Pl, initial_zero) of reference
// src/bicgstabl.jl:181-219 (bicgstabl_iterator! :27-73, iterate :79-134) written as fused passes (pass_core.h) for
// GENERAL operators and preconditioners: A and Pl may be device callbacks (`b200_linop`: the reference's duck-typed
// `mul!(y, A, x)` / `ldiv!(y, P, x)` contract, docs/src/getting_started.md:25-30, docs/src/preconditioning.md:5-15).
// For a `b200_csr` operator with Identity / Jacobi the engine of bicgstabl.cu is the fast path; this one has the same
// structure -- rs, us as n x (l+1) column-major blocks, every scalar (rho, sigma, alpha, beta, omega, the Gram matrix M,
// its LU solve, the residual, the stopping test) in device memory -- and polls the done flag every `check_every` outer
// iterations.
//
//   per j = 1..l   D1  rho = <r_shadow, rs[:, j]> ; beta = rho / sigma                                       :89-90
//                  U1  us[:, 1:j] = rs[:, 1:j] - beta us[:, 1:j]                                             :93
//                  us[:, j+1] = Pl \ (A us[:, j])                                                            :97-98
//                  D2  sigma = <r_shadow, us[:, j+1]> ; alpha = rho / sigma                                  :100-101
//                  U2  rs[:, 1:j] -= alpha us[:, 2:j+1] ; x += alpha us[:, 1]                                :103, :111
//                  rs[:, j+1] = Pl \ (A rs[:, j])                                                            :107-108
//   MR part        G   M = rs' rs, one pass per row of the upper triangle (l+1-i sums) ; after the last row  :120-124, :130
//                      gamma = M[2:end, 2:end] \ M[2:end, 1] by LU with partial pivoting (lu!), omega
//                  MR  us[:, 1] -= us[:, 2:end] gamma ; x += rs[:, 1:l] gamma ; rs[:, 1] -= rs[:, 2:end] gamma ;  :126-131
//                      ||rs[:, 1]|| ; done
#pragma once
#include "pass_core.h"

namespace b200 {

constexpr int kBcMaxL = 8;

struct BcgScal {
  double rho, sigma, alpha, beta, omega;
  double residual, tol, abstol, reltol;
  double M[(kBcMaxL + 1) * (kBcMaxL + 1)];     // rs' rs, row-major (l+1) x (l+1) with leading dimension kBcMaxL + 1
  double gamma[kBcMaxL];
  double sum[kPassMaxRed];
  double *hist;
  long long hist_cap, n_hist, iters, mv_products, max_mv;
  int l, done, singular, breakdown;
};

B200_HD bool bcg_done(const BcgScal *q) { return q->mv_products >= q->max_mv || q->residual <= q->tol; }   // :77

B200_HD void bcg_set_initial(BcgScal *q, double sumsq) {
  q->residual = sqrt(sumsq);                               // nrm = norm(residual) :60
  q->tol = fmax(q->reltol * q->residual, q->abstol);       // :66
  q->omega = 1.0;                                          // :58
  q->sigma = 1.0;
  q->breakdown = !(q->residual == q->residual);
  q->done = bcg_done(q) || q->breakdown;
}

// ---- setup :45-66: rs[:, 1] = Pl \ (b - A x) ; ||.|| ; tol
template <typename T>
struct BcgInit {
  static constexpr int NRED = 1;
  const T *b, *ax, *diag;      // ax = A*x or nullptr (initial_zero) ; diag: Jacobi or nullptr
  T *r;
  BcgScal *s;
  int is_final;                // no callback preconditioner follows: this norm is the initial residual
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    T v = b[i];                                            // copyto!(residual, b) :47
    if (ax) v = v - ax[i];                                 // residual .= b .- residual :50
    if (diag) v = v / diag[i];                             // ldiv!(Pl, residual) :55
    r[i] = v;
    acc[0] += (double)v * (double)v;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    if (is_final) bcg_set_initial(s, tot[0]);
  }
};
template <typename T>
struct BcgNorm0 {
  static constexpr int NRED = 1;
  const T *r;
  BcgScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const { acc[0] += (double)r[i] * (double)r[i]; }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const { bcg_set_initial(s, tot[0]); }
};

// ---- D1 / D2: dot with the shadow residual
template <typename T, bool IS_RHO>
struct BcgShadowDot {
  static constexpr int NRED = 1;
  const T *shadow, *v;
  BcgScal *s;
  int first;                   // D1 of j = 1 also performs sigma = -omega sigma :85
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const { acc[0] += (double)shadow[i] * (double)v[i]; }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    BcgScal *q = s;
    if (IS_RHO) {
      if (first) q->sigma = -q->omega * q->sigma;          // :85
      q->rho = tot[0];                                     // :89
      q->beta = q->rho / q->sigma;                         // :90
    } else {
      q->sigma = tot[0];                                   // :100
      q->alpha = q->rho / q->sigma;                        // :101
    }
  }
};

// ---- U1: us[:, 0:j) = rs[:, 0:j) - beta us[:, 0:j)
template <typename T>
struct BcgUpdateU {
  static constexpr int NRED = 0;
  const T *rs;
  T *us;
  int64_t ld;
  int j;
  const BcgScal *s;
  T beta;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { beta = (T)s->beta; }
  B200_HD void elem(int64_t i, double *) const {
    for (int c = 0; c < j; ++c) us[i + c * ld] = rs[i + c * ld] - beta * us[i + c * ld];   // :93
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- U2: rs[:, 0:j) -= alpha us[:, 1:j+1) ; x += alpha us[:, 0]
template <typename T>
struct BcgUpdateR {
  static constexpr int NRED = 0;
  T *rs;
  const T *us;
  T *x;
  int64_t ld;
  int j;
  const BcgScal *s;
  T alpha;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { alpha = (T)s->alpha; }
  B200_HD void elem(int64_t i, double *) const {
    x[i] = x[i] + alpha * us[i];                                                          // :111
    for (int c = 0; c < j; ++c) rs[i + c * ld] = rs[i + c * ld] - alpha * us[i + (c + 1) * ld];   // :103
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

template <typename T>
struct BcgJacobi {             // ldiv!(Pl, v) with a diagonal, in place :98 / :108
  static constexpr int NRED = 0;
  T *v;
  const T *d;
  const BcgScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *) const { v[i] = v[i] / d[i]; }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// gamma = M[1:, 1:] \ M[1:, 0] with partial pivoting as lu! does (:123-124); omega = gamma[l] (:130)
B200_HD void bcg_mr_solve(BcgScal *q) {
  const int l = q->l, LD = kBcMaxL + 1;
  double a[kBcMaxL][kBcMaxL], rhs[kBcMaxL];
  for (int i = 0; i < l; ++i) {
    rhs[i] = q->M[(i + 1) * LD + 0];
    for (int j = 0; j < l; ++j) a[i][j] = q->M[(i + 1) * LD + (j + 1)];
  }
  int singular = 0;
  for (int k = 0; k < l && !singular; ++k) {
    int piv = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < l; ++i)
      if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); piv = i; }
    if (best == 0.0 || !(best == best)) { singular = 1; break; }                           // SingularException :123
    if (piv != k) {
      for (int j = 0; j < l; ++j) { const double t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
      const double t = rhs[k]; rhs[k] = rhs[piv]; rhs[piv] = t;
    }
    for (int i = k + 1; i < l; ++i) {
      const double f = a[i][k] / a[k][k];
      for (int j = k; j < l; ++j) a[i][j] -= f * a[k][j];
      rhs[i] -= f * rhs[k];
    }
  }
  if (!singular) {
    for (int i = l - 1; i >= 0; --i) {
      double acc = rhs[i];
      for (int j = i + 1; j < l; ++j) acc -= a[i][j] * q->gamma[j];
      q->gamma[i] = acc / a[i][i];
    }
    q->omega = q->gamma[l - 1];
  } else {
    q->singular = 1;
    q->done = 1;
  }
}

// ---- G: row `row` of the upper triangle of M = rs' rs
template <typename T>
struct BcgGramRow {
  static constexpr int NRED = kBcMaxL + 1;
  const T *rs;
  int64_t ld;
  int row, L;                  // L = l + 1 columns
  BcgScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const double a = (double)rs[i + row * ld];
    B200_UNROLL
    for (int c = 0; c < kBcMaxL + 1; ++c)
      if (c >= row && c < L) acc[c] += a * (double)rs[i + c * ld];
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    const int LD = kBcMaxL + 1;
    for (int c = row; c < L; ++c) {
      s->M[row * LD + c] = tot[c];                          // mul!(M, adjoint(rs), rs) :120 (symmetric for real T)
      s->M[c * LD + row] = tot[c];
    }
    if (row == L - 1) {
      s->mv_products += 2 * s->l;                           // :115
      bcg_mr_solve(s);
    }
  }
};

// ---- MR update
template <typename T>
struct BcgMr {
  static constexpr int NRED = 1;
  T *rs, *us, *x;
  int64_t ld;
  int l;
  BcgScal *s;
  T g[kBcMaxL];
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    B200_UNROLL
    for (int c = 0; c < kBcMaxL; ++c) g[c] = c < l ? (T)s->gamma[c] : (T)0;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T u0 = us[i], xv = x[i], r0 = rs[i];
    T su = (T)0, sx = (T)0, sr = (T)0;
    B200_UNROLL
    for (int c = 0; c < kBcMaxL; ++c)                       // (compile-time bound: gamma stays in registers)
      if (c < l) {
        su = su + us[i + (c + 1) * ld] * g[c];              // us[:, 2:end] * gamma :126
        sx = sx + rs[i + c * ld] * g[c];                    // rs[:, 1:l] * gamma :127
        sr = sr + rs[i + (c + 1) * ld] * g[c];              // rs[:, 2:end] * gamma :128
      }
    us[i] = u0 - su;
    x[i] = xv + sx;
    r0 = r0 - sr;
    rs[i] = r0;
    acc[0] += (double)r0 * (double)r0;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    BcgScal *q = s;
    q->residual = sqrt(tot[0]);                             // :131
    if (!(q->residual == q->residual)) q->breakdown = 1;
    if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = q->residual;   // push!(history, :resnorm, ...) :208
    q->n_hist += 1;
    q->iters += 1;
    q->done = bcg_done(q) || q->breakdown;
  }
};

struct BcgOutcome {
  int64_t iters, mvps, n_hist;
  double residual, tol;
  int converged, breakdown, singular, done;
};

// ---- the driver in resumable pieces: scratch layout, setup (bicgstabl_iterator! :27-73), advance (up to k outer
// iterations :79-134), collect.  bicgstabl_run is the one-shot form; the iterator of the C ABI keeps the scratch.
template <typename T>
struct BcgLayout {
  T *rs, *us, *tmp;
  int64_t ld;
  BcgScal *s;
  double *hist;
  int64_t hist_cap;
};
inline size_t bcg_vec_bytes(size_t elem, int64_t n) { return ((elem * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256; }
template <typename T>
size_t bicgstabl_ws_bytes(int64_t n, int l, int64_t hist_cap) {
  return bcg_vec_bytes(sizeof(T), n) * (size_t)(2 * (l + 1) + 1) + (sizeof(BcgScal) + 255) / 256 * 256 +
         ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
}
template <typename T>
BcgLayout<T> bicgstabl_layout(void *ws, int64_t n, int l, int64_t hist_cap) {
  const size_t vb = bcg_vec_bytes(sizeof(T), n);
  BcgLayout<T> L;
  char *p = (char *)ws;
  L.rs = (T *)p; p += vb * (size_t)(l + 1);
  L.us = (T *)p; p += vb * (size_t)(l + 1);
  L.tmp = (T *)p; p += vb;
  L.ld = (int64_t)(vb / sizeof(T));
  L.s = (BcgScal *)p; p += (sizeof(BcgScal) + 255) / 256 * 256;
  L.hist = hist_cap > 0 ? (double *)p : nullptr;
  L.hist_cap = hist_cap > 0 ? hist_cap : 0;
  return L;
}

// A: the operator; Pl: preconditioner callback or nullptr; diag: Jacobi diagonal or nullptr.
template <typename T, typename B>
int bicgstabl_setup(B &be, const typename B::Op *A, const typename B::Op *Pl, const T *diag, const BcgLayout<T> &L,
                    int64_t n, int64_t n_global, T *x, const T *b, int l, double abstol, double reltol, int64_t max_mv,
                    int initial_zero) {
  if (l < 1 || l > kBcMaxL) return -1;                                      // B200_ERR_INVALID (checked by the callers)
  if (reltol < 0) reltol = sqrt(eps_of<T>());                               // :183
  if (max_mv < 0) max_mv = n_global;                                        // :184
  BcgScal h;
  memset(&h, 0, sizeof(h));
  h.abstol = abstol;
  h.reltol = reltol;
  h.max_mv = max_mv;
  h.l = l;
  h.hist = L.hist;
  h.hist_cap = L.hist_cap;
  h.mv_products = initial_zero ? 0 : 1;                                     // :45-53
  int st;
  T *rs = L.rs, *us = L.us, *tmp = L.tmp;
  BcgScal *s = L.s;
  const size_t vb = bcg_vec_bytes(sizeof(T), n);
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;
  if ((st = be.zero(us, vb * (size_t)(l + 1)))) return st;                  // zeros(T, n, l + 1) :40
  if (!initial_zero && (st = be.apply(A, x, tmp))) return st;               // mul!(residual, A, x) :49
  if (Pl) {
    if ((st = be.pass(BcgInit<T>{b, initial_zero ? nullptr : tmp, nullptr, us, s, 0}, n))) return st;   // us[:, 1] as scratch
    if ((st = be.apply(Pl, us, rs))) return st;                             // ldiv!(Pl, residual) :55
    if ((st = be.zero(us, sizeof(T) * (size_t)n))) return st;
    return be.pass(BcgNorm0<T>{rs, s}, n);
  }
  return be.pass(BcgInit<T>{b, initial_zero ? nullptr : tmp, diag, rs, s, 1}, n);
}

// up to k more outer iterations (k < 0: until done)
template <typename T, typename B>
int bicgstabl_advance(B &be, const typename B::Op *A, const typename B::Op *Pl, const T *diag, const BcgLayout<T> &L,
                      int64_t n, T *x, const T *shadow, int l, int64_t k, int check_every) {
  int st;
  T *rs = L.rs, *us = L.us, *tmp = L.tmp;
  const int64_t ld = L.ld;
  BcgScal *s = L.s;
  BcgScal h;
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  if (h.done) return 0;
  // y = Pl \ (A v): callback through tmp, Jacobi in place, Identity nothing :97-98 / :107-108
  auto apply_prec_A = [&](const T *v, T *y) -> int {
    int s2;
    if (Pl) {
      if ((s2 = be.apply(A, v, tmp))) return s2;
      return be.apply(Pl, tmp, y);
    }
    if ((s2 = be.apply(A, v, y))) return s2;
    if (diag) return be.pass(BcgJacobi<T>{y, diag, s}, n);
    return 0;
  };
  const int64_t left_mv = h.max_mv - h.mv_products;
  const int64_t left = left_mv > 0 ? (left_mv + 2 * l - 1) / (2 * l) : 0;   // outer iterations until the product budget ends
  const int64_t todo = (k < 0 || k > left) ? left : k;
  if (check_every <= 0) check_every = 4;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= todo) break;
    const int64_t batch = check_every < todo - enqueued ? check_every : todo - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      for (int j = 1; j <= l; ++j) {                                                                       // :88
        T *rj = rs + (int64_t)j * ld, *rjm = rs + (int64_t)(j - 1) * ld;
        T *uj = us + (int64_t)j * ld, *ujm = us + (int64_t)(j - 1) * ld;
        if ((st = be.pass(BcgShadowDot<T, true>{shadow, rjm, s, j == 1}, n))) return st;                   // D1
        if ((st = be.pass(BcgUpdateU<T>{rs, us, ld, j, s, (T)0}, n))) return st;                           // U1
        if ((st = apply_prec_A(ujm, uj))) return st;                                                       // :97-98
        if ((st = be.pass(BcgShadowDot<T, false>{shadow, uj, s, 0}, n))) return st;                        // D2
        if ((st = be.pass(BcgUpdateR<T>{rs, us, x, ld, j, s, (T)0}, n))) return st;                        // U2
        if ((st = apply_prec_A(rjm, rj))) return st;                                                       // :107-108
      }
      for (int row = 0; row <= l; ++row)
        if ((st = be.pass(BcgGramRow<T>{rs, ld, row, l + 1, s}, n))) return st;                            // G
      BcgMr<T> mr{rs, us, x, ld, l, s, {}};
      if ((st = be.pass(mr, n))) return st;                                                                // MR
    }
    enqueued += batch;
  }
  return 0;
}

template <typename T, typename B>
int bicgstabl_collect(B &be, const BcgLayout<T> &L, double *hist_host, BcgOutcome *out) {
  int st;
  BcgScal h;
  if ((st = be.to_host(&h, L.s, sizeof(h)))) return st;
  out->iters = h.iters;
  out->mvps = h.mv_products;                                                // history.mvps = iterable.mv_products :207
  out->residual = h.residual;
  out->tol = h.tol;
  out->converged = h.residual <= h.tol;                                     // converged :75
  out->breakdown = h.breakdown;
  out->singular = h.singular;
  out->done = h.done;
  out->n_hist = h.n_hist < L.hist_cap ? h.n_hist : L.hist_cap;
  if (hist_host && out->n_hist > 0 && (st = be.to_host(hist_host, L.hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}
template <typename B>
int bicgstabl_reset_window(B &be, BcgScal *s) {
  const long long zero = 0;
  return be.to_device(&s->n_hist, &zero, sizeof(zero));
}

// shadow: r_shadow (:38, drawn by the caller).
template <typename T, typename B>
int bicgstabl_run(B &be, const typename B::Op *A, const typename B::Op *Pl, const T *diag, int64_t n, int64_t n_global,
                  T *x, const T *b, const T *shadow, int l, double abstol, double reltol, int64_t max_mv, int initial_zero,
                  int check_every, int64_t hist_cap, double *hist_host, BcgOutcome *out) {
  if (l < 1 || l > kBcMaxL) return -1;                                      // B200_ERR_INVALID (checked by the callers)
  if (max_mv < 0) max_mv = n_global;
  if (!hist_host) hist_cap = 0;
  if (hist_cap > max_mv) hist_cap = max_mv;                                 // reserve!(history, :resnorm, max_mv_products) :194
  void *ws = nullptr;
  int st = be.workspace(bicgstabl_ws_bytes<T>(n, l, hist_cap), &ws);
  if (st) return st;
  const BcgLayout<T> L = bicgstabl_layout<T>(ws, n, l, hist_cap);
  if ((st = bicgstabl_setup<T>(be, A, Pl, diag, L, n, n_global, x, b, l, abstol, reltol, max_mv, initial_zero))) return st;
  if ((st = bicgstabl_advance<T>(be, A, Pl, diag, L, n, x, shadow, l, -1, check_every))) return st;
  return bicgstabl_collect<T>(be, L, hist_host, out);
}

}  // namespace b200
