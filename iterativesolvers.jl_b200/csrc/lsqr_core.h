// lsqr_core.h -- lsqr!(x, A, b; damp, atol, btol, conlim, maxiter) of reference src/lsqr.jl:66-77 (lsqr_method!
// :90-275) written as fused passes (pass_core.h).  A is m x n (rectangular allowed), At its adjoint.
//
// Per iteration of the Golub-Kahan bidiagonalisation (reference order of operations inside every element update):
//   S1  tmpm = A v                                                            :163  (SpMV)
//   L1  u = -alpha u + tmpm ; beta = ||u||                                    :164-165   (2 reads, 1 write over m)
//       scalar: itn, mvps ; beta > 0: mtvps, Anorm                            :153-154, :166-169
//   L2  u *= 1/beta                                                           :168       (1 read, 1 write over m)
//   S2  tmpn = A' u                                                           :172  (SpMV with the adjoint operator)
//   L3  v = -beta v + tmpn ; alpha = ||v||                                    :173-174   (2 reads, 1 write over n)
//       scalar: the two plane rotations, t1, t2                               :182-201
//   L4  v *= 1/alpha ; x += t1 w ; w = t2 w + v ; ||w/rho||                   :176, :203-206 (3 reads, 3 writes over n)
//       scalar: norm/condition estimates, the four history pushes, the stopping tests :206-271
// (wrho of :205 is not materialised: only its norm is used.)  When beta == 0 the reference skips :167-177; here L2/L3
// become no-ops through a device flag and S2 writes a scratch vector that nobody reads.
// Algorithmic bytes per iteration: 2*nnz*(V+4) + (m+n+2)*4 + 2*(m+n)*V [the two SpMVs] + (5*m + 9*n)*V.
//
// The scalars are carried in fp64 also for Float32 vectors (the reference keeps them in real(T)); the
// `1 + test <= 1` guards of :261-264 are evaluated in the precision of T.
#pragma once
#include "pass_core.h"

namespace b200 {

struct LsqrScal {
  double alpha, beta, Anorm, Acond, ddnorm, res2, xnorm, xxnorm, z, sn2, cs2, dampsq, damp;   // :110-114
  double rhobar, phibar, bnorm, rnorm, r1norm, r2norm, Arnorm;                               // :141-147
  double atol, btol, ctol;
  double rho, phi, theta, tau, psi;                   // carried from the rotation section to the estimate section
  double inv_beta, inv_alpha, t1, t2, inv_rho;        // coefficients handed to the vector passes
  double sum[2];
  double *hist;                                       // 4 rows of hist_cap: resnorm | anorm | rnorm | cnorm
  long long hist_cap, n_hist;
  long long itn, maxiter, mvps, mtvps;
  int istop, done, beta_pos, alpha_pos, is_f32, bad_x, early, pad;
};

B200_HD bool one_plus_le_one(double t, int is_f32) {     // `1 + t <= 1` in the precision of the vectors
  if (is_f32) {
    volatile float s = 1.0f + (float)t;
    return s <= 1.0f;
  }
  volatile double s = 1.0 + t;
  return s <= 1.0;
}

// ---- initialisation :97-147
template <typename T>
struct LsqrCheckX {          // all(isfinite.(x)) :102-104 ; v = copy(x) :125
  static constexpr int NRED = 1;
  const T *x;
  T *v;
  LsqrScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const T xi = x[i];
    v[i] = xi;
    const double d = (double)xi;
    if (!(d - d == 0.0)) acc[0] += 1.0;      // Inf - Inf and NaN - NaN are NaN
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->bad_x = tot[0] > 0.0;
    if (s->bad_x) s->done = 1;
  }
};

template <typename T>
struct LsqrInitU {           // u = b - A*x ; beta = norm(u) :124-131
  static constexpr int NRED = 1;
  const T *b, *ax;
  T *u;
  LsqrScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const T t = b[i] - ax[i];
    u[i] = t;
    acc[0] += (double)t * (double)t;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->beta = sqrt(tot[0]);                            // :126
    s->alpha = 0.0;                                    // :127
    s->beta_pos = s->beta > 0.0;                       // :129
    s->inv_beta = s->beta_pos ? 1.0 / s->beta : 1.0;
    s->mtvps = s->beta_pos ? 1 : 0;                    // :130
    s->mvps = 0;
  }
};

template <typename T>
struct LsqrScaleU {          // u .*= inv(beta) :131 / :168
  static constexpr int NRED = 0;
  T *u;
  const LsqrScal *s;
  int in_loop;               // inside the main loop the pass is skipped once `done` is set
  T inv;
  bool on;
  B200_HD bool skip() const { return in_loop && s->done != 0; }
  B200_HD void load() { inv = (T)s->inv_beta; on = s->beta_pos != 0; }
  B200_HD void elem(int64_t i, double *) const {
    if (on) u[i] = u[i] * inv;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

template <typename T>
struct LsqrInitV {           // mul!(v, adjointA, u) ; alpha = norm(v) :132-133 ; then :141-147
  static constexpr int NRED = 1;
  const T *atu;
  T *v;
  LsqrScal *s;
  bool on;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { on = s->beta_pos != 0; }
  B200_HD void elem(int64_t i, double *acc) const {
    if (on) {
      const T t = atu[i];
      v[i] = t;
      acc[0] += (double)t * (double)t;
    }
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsqrScal *q = s;
    if (q->beta_pos) q->alpha = sqrt(tot[0]);          // :133
    q->alpha_pos = q->alpha > 0.0;                     // :135
    q->inv_alpha = q->alpha_pos ? 1.0 / q->alpha : 1.0;
    q->Arnorm = q->alpha * q->beta;                    // :141
    q->itn = 0;
    q->istop = 0;
    q->n_hist = 0;
    q->Anorm = q->Acond = q->ddnorm = q->res2 = q->xnorm = q->xxnorm = q->z = q->sn2 = 0.0;   // :112
    q->cs2 = -1.0;                                     // :113
    q->rhobar = q->alpha;                              // :146
    q->phibar = q->bnorm = q->rnorm = q->r1norm = q->r2norm = q->beta;   // :147
    if (q->Arnorm == 0.0 || !(q->Arnorm == q->Arnorm)) {               // :142-144 (NaN: stop as well)
      q->early = 1;
      q->done = 1;
    }
    if (!(q->itn < q->maxiter)) q->done = 1;           // loop condition :152
  }
};

template <typename T>
struct LsqrInitW {           // v .*= inv(alpha) :135-137 ; w = copy(v) :138
  static constexpr int NRED = 0;
  T *v, *w;
  const LsqrScal *s;
  T inv;
  bool on;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { inv = (T)s->inv_alpha; on = s->alpha_pos != 0; }
  B200_HD void elem(int64_t i, double *) const {
    T t = v[i];
    if (on) t = t * inv;
    v[i] = t;
    w[i] = t;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- L1
template <typename T>
struct LsqrU {
  static constexpr int NRED = 1;
  T *u;
  const T *av;
  LsqrScal *s;
  T nalpha;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { nalpha = (T)(-s->alpha); }
  B200_HD void elem(int64_t i, double *acc) const {
    const T t = nalpha * u[i] + av[i];                 // u .= -alpha .* u .+ tmpm :164
    u[i] = t;
    acc[0] += (double)t * (double)t;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsqrScal *q = s;
    q->itn += 1;                                       // :154
    q->mvps += 1;                                      // nextiter!(log, mvps=1) :153
    q->beta = sqrt(tot[0]);                            // :165
    q->beta_pos = q->beta > 0.0;                       // :166
    if (q->beta_pos) {
      q->mtvps += 1;                                   // :167
      q->inv_beta = 1.0 / q->beta;                     // :168
      q->Anorm = sqrt(q->Anorm * q->Anorm + q->alpha * q->alpha + q->beta * q->beta + q->dampsq);   // :169
    } else {
      q->inv_beta = 1.0;
    }
  }
};

// ---- L3 and the rotation section
template <typename T>
struct LsqrV {
  static constexpr int NRED = 1;
  T *v;
  const T *atu;
  LsqrScal *s;
  T nbeta;
  bool on;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { nbeta = (T)(-s->beta); on = s->beta_pos != 0; }
  B200_HD void elem(int64_t i, double *acc) const {
    if (on) {
      const T t = nbeta * v[i] + atu[i];               // v .= -beta .* v .+ tmpn :173
      v[i] = t;
      acc[0] += (double)t * (double)t;
    }
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsqrScal *q = s;
    q->alpha_pos = 0;
    q->inv_alpha = 1.0;
    if (q->beta_pos) {
      q->alpha = sqrt(tot[0]);                         // :174
      if (q->alpha > 0.0) {                            // :175-177
        q->alpha_pos = 1;
        q->inv_alpha = 1.0 / q->alpha;
      }
    }
    const double rhobar1 = sqrt(q->rhobar * q->rhobar + q->dampsq);   // :182
    const double cs1 = q->rhobar / rhobar1;            // :183
    const double sn1 = q->damp / rhobar1;              // :184
    q->psi = sn1 * q->phibar;                          // :185
    q->phibar = cs1 * q->phibar;                       // :186
    q->rho = sqrt(rhobar1 * rhobar1 + q->beta * q->beta);   // :190
    const double cs = rhobar1 / q->rho;                // :191
    const double sn = q->beta / q->rho;                // :192
    q->theta = sn * q->alpha;                          // :193
    q->rhobar = -cs * q->alpha;                        // :194
    q->phi = cs * q->phibar;                           // :195
    q->phibar = sn * q->phibar;                        // :196
    q->tau = sn * q->phi;                              // :197
    q->t1 = q->phi / q->rho;                           // :200
    q->t2 = -q->theta / q->rho;                        // :201
    q->inv_rho = 1.0 / q->rho;                         // :205
  }
};

// ---- L4 and the estimate / stopping section
template <typename T>
struct LsqrXW {
  static constexpr int NRED = 1;
  T *v, *x, *w;
  LsqrScal *s;
  T inv_alpha, t1, t2, inv_rho;
  bool scale;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    inv_alpha = (T)s->inv_alpha; t1 = (T)s->t1; t2 = (T)s->t2; inv_rho = (T)s->inv_rho;
    scale = s->alpha_pos != 0;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T vi = v[i];
    if (scale) {                                       // v .*= inv(alpha) :176
      vi = vi * inv_alpha;
      v[i] = vi;
    }
    const T wi = w[i];
    x[i] = x[i] + t1 * wi;                             // :203
    const T wn = t2 * wi + vi;                         // :204
    w[i] = wn;
    const T wr = wn * inv_rho;                         // :205
    acc[0] += (double)wr * (double)wr;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsqrScal *q = s;
    q->ddnorm += sqrt(tot[0]);                         // ddnorm += norm(wrho) :206 (sic: the norm, not its square)
    const double delta = q->sn2 * q->rho;              // :211
    const double gambar = -q->cs2 * q->rho;            // :212
    const double rhs = q->phi - delta * q->z;          // :213
    const double zbar = rhs / gambar;                  // :214
    q->xnorm = sqrt(q->xxnorm + zbar * zbar);          // :215
    const double gamma = sqrt(gambar * gambar + q->theta * q->theta);   // :216
    q->cs2 = gambar / gamma;                           // :217
    q->sn2 = q->theta / gamma;                         // :218
    q->z = rhs / gamma;                                // :219
    q->xxnorm += q->z * q->z;                          // :220
    q->Acond = q->Anorm * sqrt(q->ddnorm);             // :225
    const double res1 = q->phibar * q->phibar;         // :226
    q->res2 = q->res2 + q->psi * q->psi;               // :227
    q->rnorm = sqrt(res1 + q->res2);                   // :228
    q->Arnorm = q->alpha * fabs(q->tau);               // :229
    const double r1sq = q->rnorm * q->rnorm - q->dampsq * q->xxnorm;   // :239
    q->r1norm = sqrt(fabs(r1sq));                      // :240
    if (r1sq < 0) q->r1norm = -q->r1norm;
    q->r2norm = q->rnorm;                              // :241
    const double test1 = q->rnorm / q->bnorm;          // :246
    const double test2 = q->Arnorm / (q->Anorm * q->rnorm);   // :247
    const double test3 = 1.0 / q->Acond;               // :248
    const double t1 = test1 / (1 + q->Anorm * q->xnorm / q->bnorm);   // :249
    const double rtol = q->btol + q->atol * q->Anorm * q->xnorm / q->bnorm;   // :250
    if (q->hist && q->n_hist < q->hist_cap) {          // push! :242, :251-253
      q->hist[q->n_hist] = q->r1norm;
      q->hist[q->hist_cap + q->n_hist] = test2;        // :anorm
      q->hist[2 * q->hist_cap + q->n_hist] = test1;    // :rnorm
      q->hist[3 * q->hist_cap + q->n_hist] = test3;    // :cnorm
    }
    q->n_hist += 1;
    if (q->itn >= q->maxiter) q->istop = 7;            // :261
    if (one_plus_le_one(test3, q->is_f32)) q->istop = 6;   // :262
    if (one_plus_le_one(test2, q->is_f32)) q->istop = 5;   // :263
    if (one_plus_le_one(t1, q->is_f32)) q->istop = 4;      // :264
    if (test3 <= q->ctol) q->istop = 3;                // :267
    if (test2 <= q->atol) q->istop = 2;                // :268
    if (test1 <= rtol) q->istop = 1;                   // :269
    // while (itn < maxiter) & !log.isconverged :152 ; setconv(log, istop > 0) :271
    if (q->istop > 0 || q->itn >= q->maxiter || !(q->rnorm == q->rnorm)) q->done = 1;
  }
};

struct LsqrOutcome {
  int64_t iters, mvps, mtvps, n_hist, hist_stride;
  int istop, converged, bad_x, early;
  double atol, btol, ctol, anorm, acond, rnorm, arnorm, xnorm;
};

// A: m x n, At: n x m; x: n values (initial guess, updated in place); b: m values.
// hist_host: 4 rows (resnorm | anorm | rnorm | cnorm) of out->hist_stride = min(hist_cap, maxiter) doubles each, or NULL.
template <typename T, typename B>
int lsqr_run(B &be, const typename B::Op *A, const typename B::Op *At, int64_t m, int64_t n, T *x, const T *b,
             double damp, double atol, double btol, double conlim, int64_t maxiter, int check_every, int64_t hist_cap,
             double *hist_host, LsqrOutcome *out) {
  const double eps = eps_of<T>();
  if (atol < 0) atol = sqrt(eps);                                           // :91
  if (btol < 0) btol = sqrt(eps);
  if (conlim < 0) conlim = 1.0 / sqrt(eps);                                 // :92
  if (maxiter < 0) maxiter = m > n ? m : n;                                 // maximum(size(A)) :67
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;
  const size_t mb = ((sizeof(T) * (size_t)(m > 0 ? m : 1)) + 255) / 256 * 256;
  const size_t nb = ((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256;
  const size_t hb = ((sizeof(double) * 4 * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(2 * mb + 3 * nb + 512 + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *u = (T *)p; p += mb;
  T *tmpm = (T *)p; p += mb;
  T *v = (T *)p; p += nb;
  T *w = (T *)p; p += nb;
  T *tmpn = (T *)p; p += nb;
  LsqrScal *s = (LsqrScal *)p; p += 512;
  double *hist = hist_cap ? (double *)p : nullptr;
  static_assert(sizeof(LsqrScal) <= 512, "LsqrScal outgrew its slot");

  LsqrScal h;
  memset(&h, 0, sizeof(h));
  h.damp = damp;
  h.dampsq = damp * damp;                                                   // :114
  h.atol = atol;
  h.btol = btol;
  h.ctol = conlim > 0 ? 1.0 / conlim : 0.0;                                 // :111
  if (sizeof(T) == 4) h.ctol = (double)(float)h.ctol;                       // convert(Tr, 1/conlim)
  h.maxiter = maxiter;
  h.hist = hist;
  h.hist_cap = hist_cap;
  h.is_f32 = sizeof(T) == 4;
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;

  if ((st = be.pass(LsqrCheckX<T>{x, v, s}, n))) return st;                                   // :102-104, :125
  if ((st = be.apply(A, x, tmpm))) return st;                                                 // A*x :124
  if ((st = be.pass(LsqrInitU<T>{b, tmpm, u, s}, m))) return st;
  if ((st = be.pass(LsqrScaleU<T>{u, s, 0}, m))) return st;
  if ((st = be.apply(At, u, tmpn))) return st;                                                // :132
  if ((st = be.pass(LsqrInitV<T>{tmpn, v, s}, n))) return st;
  if ((st = be.pass(LsqrInitW<T>{v, w, s}, n))) return st;

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= maxiter) break;
    const int64_t batch = check_every < maxiter - enqueued ? check_every : maxiter - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      if ((st = be.apply(A, v, tmpm))) return st;                                             // S1
      if ((st = be.pass(LsqrU<T>{u, tmpm, s}, m))) return st;                                  // L1
      if ((st = be.pass(LsqrScaleU<T>{u, s, 1}, m))) return st;                                // L2
      if ((st = be.apply(At, u, tmpn))) return st;                                            // S2
      if ((st = be.pass(LsqrV<T>{v, tmpn, s}, n))) return st;                                  // L3
      if ((st = be.pass(LsqrXW<T>{v, x, w, s}, n))) return st;                                 // L4
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  out->iters = h.itn;
  out->mvps = h.mvps;
  out->mtvps = h.mtvps;
  out->istop = h.istop;
  out->converged = h.istop > 0;                                             // setconv(log, istop > 0) :271
  out->bad_x = h.bad_x;
  out->early = h.early;
  out->atol = h.atol; out->btol = h.btol; out->ctol = h.ctol;
  out->anorm = h.Anorm; out->acond = h.Acond; out->rnorm = h.rnorm; out->arnorm = h.Arnorm; out->xnorm = h.xnorm;
  out->n_hist = h.n_hist < hist_cap ? h.n_hist : hist_cap;
  out->hist_stride = hist_cap;
  if (hist_cap > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * 4 * (size_t)hist_cap))) return st;
  return 0;
}

}  // namespace b200
