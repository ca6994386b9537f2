// lsmr_core.h -- lsmr!(x, A, b; atol, btol, conlim, maxiter, λ) of reference src/lsmr.jl:67-82 (lsmr_method!
// :88-287) written as fused passes (pass_core.h).  A is m x n (rectangular allowed), At its adjoint.
//
// Per iteration (reference order of operations inside every element update):
//   S1  tmp_u = A v                                                           :166  (SpMV)
//   M1  u = tmp_u + u*(-alpha) ; beta = ||u||                                 :167-168   (2 reads, 1 write over m)
//   M2  u *= 1/beta                    (beta > 0)                             :171       (1 read, 1 write over m)
//   S2  tmp_v = A' u                                                          :172  (SpMV with the adjoint operator)
//   M3  v = tmp_v + v*(-beta) ; alpha = ||v||                                 :173-174   (2 reads, 1 write over n)
//       scalar: the rotations Qhat, Q, Qbar and the three update coefficients :179-205
//   M4  v *= 1/alpha ; hbar = hbar*c1 + h ; x += c2*hbar ; h = h*c3 + v ; ||x||   :175, :203-205, :255
//                                                                             (4 reads, 4 writes over n)
//       scalar: ||r|| / ||A|| / cond(A) estimates, history pushes, stopping tests :214-281
// Algorithmic bytes per iteration: the two SpMVs + (5*m + 11*n)*V.
//
// Deviation: for b - A*x == 0 (or A'u == 0) the reference divides by a zero norm (:116, :120) and then
// iterates on NaNs although its comment (:158) announces an early exit; this engine takes that exit (x unchanged,
// istop = 0).  Scalars are carried in fp64 also for Float32 vectors; the `1 + test <= 1` guards (:275-277) are
// evaluated in the precision of T.
#pragma once
#include "lsqr_core.h"   // one_plus_le_one

namespace b200 {

struct LsmrScal {
  double alpha, beta, lambda;
  double zetabar, alphabar, rho, rhobar, cbar, sbar;                    // :127-132
  double betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;      // :138-144
  double normA, condA, normx, normA2, maxrbar, minrbar;                 // :147-150
  double normb, normr, normAr;                                          // :153-156
  double atol, btol, ctol;
  // carried from the rotation section (M3) to the estimate section (M4)
  double chat, shat, c, s, rhoold, rhobarold, zetaold, thetabar, rhotemp;
  double inv_beta, inv_alpha, c1, c2, c3;                               // coefficients handed to the vector passes
  double sum[2];
  double *hist;                                                         // 4 rows of hist_cap: normr | anorm | rnorm | cnorm
  long long hist_cap, n_hist;
  long long iter, maxiter, mvps, mtvps;
  int istop, done, beta_pos, is_f32, early, pad;
};

// ---- initialisation :110-161
template <typename T>
struct LsmrInitU {           // b .-= A*x ; u = b ; beta = norm(u) :112-115
  static constexpr int NRED = 1;
  const T *b, *ax;
  T *u;
  LsmrScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const T t = b[i] - ax[i];
    u[i] = t;
    acc[0] += (double)t * (double)t;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->beta = sqrt(tot[0]);                            // :115
    s->beta_pos = s->beta > 0.0;
    s->inv_beta = s->beta_pos ? 1.0 / s->beta : 1.0;   // :116 (see the header for beta == 0)
  }
};

template <typename T>
struct LsmrScaleU {          // u .*= inv(beta) :116 / :171
  static constexpr int NRED = 0;
  T *u;
  const LsmrScal *s;
  int in_loop;
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
struct LsmrInitV {           // alpha = norm(v) :119 ; scalars :127-161
  static constexpr int NRED = 1;
  const T *v;
  LsmrScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const { acc[0] += (double)v[i] * (double)v[i]; }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsmrScal *q = s;
    q->alpha = q->beta_pos ? sqrt(tot[0]) : 0.0;       // :119
    q->inv_alpha = q->alpha > 0.0 ? 1.0 / q->alpha : 1.0;   // :120
    q->zetabar = q->alpha * q->beta;                   // :127
    q->alphabar = q->alpha;                            // :128
    q->rho = 1.0; q->rhobar = 1.0; q->cbar = 1.0; q->sbar = 0.0;   // :129-132
    q->betadd = q->beta;                               // :138
    q->betad = 0.0; q->rhodold = 1.0; q->tautildeold = 0.0; q->thetatilde = 0.0; q->zeta = 0.0; q->d = 0.0;   // :139-144
    q->normA = q->condA = q->normx = -1.0;             // :147
    q->normA2 = q->alpha * q->alpha;                   // :148
    q->maxrbar = 0.0;                                  // :149
    q->minrbar = 1e100;                                // :150
    q->normb = q->beta;                                // :153
    q->istop = 0;                                      // :154
    q->normr = q->beta;                                // :155
    q->normAr = q->alpha * q->beta;                    // :156
    q->iter = 0;                                       // :157
    q->mvps = 1;                                       // :160
    q->mtvps = 1;                                      // :161
    q->n_hist = 0;
    if (!(q->normAr != 0.0) || !(q->normAr == q->normAr)) {   // :162 (and the NaN case, see the header)
      q->early = 1;
      q->done = 1;
    }
    if (!(q->iter < q->maxiter)) q->done = 1;          // :163
  }
};

template <typename T>
struct LsmrInitH {           // v .*= inv(alpha) :120 ; copyto!(h, v) :134 ; fill!(hbar, 0) :135
  static constexpr int NRED = 0;
  T *v, *h, *hbar;
  const LsmrScal *s;
  T inv;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { inv = (T)s->inv_alpha; }
  B200_HD void elem(int64_t i, double *) const {
    const T t = v[i] * inv;
    v[i] = t;
    h[i] = t;
    hbar[i] = (T)0;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- M1
template <typename T>
struct LsmrU {
  static constexpr int NRED = 1;
  T *u;
  const T *av;
  LsmrScal *s;
  T nalpha;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { nalpha = (T)(-s->alpha); }
  B200_HD void elem(int64_t i, double *acc) const {
    const T t = av[i] + u[i] * nalpha;                 // u .= tmp_u .+ u .* -α :167
    u[i] = t;
    acc[0] += (double)t * (double)t;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsmrScal *q = s;
    q->iter += 1;                                      // :165
    q->mvps += 1;                                      // nextiter!(log, mvps=1) :164
    q->beta = sqrt(tot[0]);                            // :168
    q->beta_pos = q->beta > 0.0;                       // :169
    if (q->beta_pos) {
      q->mtvps += 1;                                   // :170
      q->inv_beta = 1.0 / q->beta;                     // :171
    } else {
      q->inv_beta = 1.0;
    }
  }
};

// ---- M3 and the rotation section
template <typename T>
struct LsmrV {
  static constexpr int NRED = 1;
  T *v;
  const T *atu;
  LsmrScal *s;
  T nbeta;
  bool on;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { nbeta = (T)(-s->beta); on = s->beta_pos != 0; }
  B200_HD void elem(int64_t i, double *acc) const {
    if (on) {
      const T t = atu[i] + v[i] * nbeta;               // v .= tmp_v .+ v .* -β :173
      v[i] = t;
      acc[0] += (double)t * (double)t;
    }
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsmrScal *q = s;
    q->inv_alpha = 1.0;
    if (q->beta_pos) {
      q->alpha = sqrt(tot[0]);                         // :174
      q->inv_alpha = 1.0 / q->alpha;                   // :175 (unguarded in the reference as well)
    }
    const double alphahat = hypot(q->alphabar, q->lambda);   // :179
    q->chat = q->alphabar / alphahat;                  // :180
    q->shat = q->lambda / alphahat;                    // :181
    q->rhoold = q->rho;                                // :184
    q->rho = hypot(alphahat, q->beta);                 // :185
    q->c = alphahat / q->rho;                          // :186
    q->s = q->beta / q->rho;                           // :187
    const double thetanew = q->s * q->alpha;           // :188
    q->alphabar = q->c * q->alpha;                     // :189
    q->rhobarold = q->rhobar;                          // :192
    q->zetaold = q->zeta;                              // :193
    q->thetabar = q->sbar * q->rho;                    // :194
    q->rhotemp = q->cbar * q->rho;                     // :195
    q->rhobar = hypot(q->cbar * q->rho, thetanew);     // :196
    q->cbar = q->cbar * q->rho / q->rhobar;            // :197
    q->sbar = thetanew / q->rhobar;                    // :198
    q->zeta = q->cbar * q->zetabar;                    // :199
    q->zetabar = -q->sbar * q->zetabar;                // :200
    q->c1 = -q->thetabar * q->rho / (q->rhoold * q->rhobarold);   // :203
    q->c2 = q->zeta / (q->rho * q->rhobar);            // :204
    q->c3 = -thetanew / q->rho;                        // :205
  }
};

// ---- M4 and the estimate / stopping section
template <typename T>
struct LsmrXH {
  static constexpr int NRED = 1;
  T *v, *h, *hbar, *x;
  LsmrScal *s;
  T inv_alpha, c1, c2, c3;
  bool scale;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    inv_alpha = (T)s->inv_alpha; c1 = (T)s->c1; c2 = (T)s->c2; c3 = (T)s->c3;
    scale = s->beta_pos != 0;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T vi = v[i];
    if (scale) {                                       // v .*= inv(α) :175
      vi = vi * inv_alpha;
      v[i] = vi;
    }
    const T hi = h[i];
    const T hb = hbar[i] * c1 + hi;                    // :203
    hbar[i] = hb;
    const T xi = x[i] + c2 * hb;                       // :204
    x[i] = xi;
    h[i] = hi * c3 + vi;                               // :205
    acc[0] += (double)xi * (double)xi;                 // norm(x) :255
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    LsmrScal *q = s;
    const double betaacute = q->chat * q->betadd;      // :214
    const double betacheck = -q->shat * q->betadd;     // :215
    const double betahat = q->c * betaacute;           // :218
    q->betadd = -q->s * betaacute;                     // :219
    const double thetatildeold = q->thetatilde;        // :222
    const double rhotildeold = hypot(q->rhodold, q->thetabar);   // :223
    const double ctildeold = q->rhodold / rhotildeold; // :224
    const double stildeold = q->thetabar / rhotildeold;   // :225
    q->thetatilde = stildeold * q->rhobar;             // :226
    q->rhodold = ctildeold * q->rhobar;                // :227
    q->betad = -stildeold * q->betad + ctildeold * betahat;   // :228
    q->tautildeold = (q->zetaold - thetatildeold * q->tautildeold) / rhotildeold;   // :230
    const double taud = (q->zeta - q->thetatilde * q->tautildeold) / q->rhodold;    // :231
    q->d += betacheck * betacheck;                     // :232
    q->normr = sqrt(q->d + (q->betad - taud) * (q->betad - taud) + q->betadd * q->betadd);   // :233
    q->normA2 += q->beta * q->beta;                    // :236
    q->normA = sqrt(q->normA2);                        // :237
    q->normA2 += q->alpha * q->alpha;                  // :238
    q->maxrbar = fmax(q->maxrbar, q->rhobarold);       // :241
    if (q->iter > 1) q->minrbar = fmin(q->minrbar, q->rhobarold);   // :242-244
    q->condA = fmax(q->maxrbar, q->rhotemp) / fmin(q->minrbar, q->rhotemp);   // :245
    q->normAr = fabs(q->zetabar);                      // :254
    q->normx = sqrt(tot[0]);                           // :255
    const double test1 = q->normr / q->normb;          // :259
    const double test2 = q->normAr / (q->normA * q->normr);   // :260
    const double test3 = 1.0 / q->condA;               // :261
    if (q->hist && q->n_hist < q->hist_cap) {          // push! :262-264
      q->hist[q->n_hist] = q->normr;                   // (not in the reference's history: the ||r|| estimate)
      q->hist[q->hist_cap + q->n_hist] = test2;        // :anorm
      q->hist[2 * q->hist_cap + q->n_hist] = test1;    // :rnorm
      q->hist[3 * q->hist_cap + q->n_hist] = test3;    // :cnorm
    }
    q->n_hist += 1;
    const double t1 = test1 / (1.0 + q->normA * q->normx / q->normb);   // :267
    const double rtol = q->btol + q->atol * q->normA * q->normx / q->normb;   // :268
    // first match wins (each test `break`s) :274-281
    if (q->iter >= q->maxiter) q->istop = 7;
    else if (one_plus_le_one(test3, q->is_f32)) q->istop = 6;
    else if (one_plus_le_one(test2, q->is_f32)) q->istop = 5;
    else if (one_plus_le_one(t1, q->is_f32)) q->istop = 4;
    else if (test3 <= q->ctol) q->istop = 3;
    else if (test2 <= q->atol) q->istop = 2;
    else if (test1 <= rtol) q->istop = 1;
    if (q->istop > 0 || !(q->normr == q->normr)) q->done = 1;
  }
};

struct LsmrOutcome {
  int64_t iters, mvps, mtvps, n_hist, hist_stride;
  int istop, converged, early;
  double atol, btol, ctol, normr, normAr, normA, condA, normx;
};

// A: m x n, At: n x m; x: n values (initial guess, updated in place); b: m values (not modified: the reference
// works on a copy, :76-77).  hist_host: 4 rows (normr | anorm | rnorm | cnorm) of out->hist_stride doubles, or NULL.
template <typename T, typename B>
int lsmr_run(B &be, const typename B::Op *A, const typename B::Op *At, int64_t m, int64_t n, T *x, const T *b,
             double lambda, double atol, double btol, double conlim, int64_t maxiter, int check_every,
             int64_t hist_cap, double *hist_host, LsmrOutcome *out) {
  if (atol < 0) atol = 1e-6;                                                // :89
  if (btol < 0) btol = 1e-6;
  if (conlim < 0) conlim = 1e8;
  if (maxiter < 0) maxiter = m > n ? m : n;                                 // maximum(size(A)) :68
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;
  const size_t mb = ((sizeof(T) * (size_t)(m > 0 ? m : 1)) + 255) / 256 * 256;
  const size_t nb = ((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256;
  const size_t hb = ((sizeof(double) * 4 * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(2 * mb + 4 * nb + 1024 + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *u = (T *)p; p += mb;
  T *tmp_u = (T *)p; p += mb;
  T *v = (T *)p; p += nb;
  T *h = (T *)p; p += nb;
  T *hbar = (T *)p; p += nb;
  T *tmp_v = (T *)p; p += nb;
  LsmrScal *s = (LsmrScal *)p; p += 1024;
  double *hist = hist_cap ? (double *)p : nullptr;
  static_assert(sizeof(LsmrScal) <= 1024, "LsmrScal outgrew its slot");

  LsmrScal hs;
  memset(&hs, 0, sizeof(hs));
  hs.lambda = lambda;
  hs.atol = atol;
  hs.btol = btol;
  hs.ctol = conlim > 0 ? 1.0 / conlim : 0.0;                                // :108
  if (sizeof(T) == 4) hs.ctol = (double)(float)hs.ctol;                     // convert(Tr, inv(conlim))
  hs.maxiter = maxiter;
  hs.hist = hist;
  hs.hist_cap = hist_cap;
  hs.is_f32 = sizeof(T) == 4;
  if ((st = be.to_device(s, &hs, sizeof(hs)))) return st;

  if ((st = be.apply(A, x, tmp_u))) return st;                                                // :112
  if ((st = be.pass(LsmrInitU<T>{b, tmp_u, u, s}, m))) return st;
  if ((st = be.pass(LsmrScaleU<T>{u, s, 0}, m))) return st;
  if ((st = be.apply(At, u, v))) return st;                                                   // :118
  if ((st = be.pass(LsmrInitV<T>{v, s}, n))) return st;
  if ((st = be.pass(LsmrInitH<T>{v, h, hbar, s}, n))) return st;

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&s->done, &done))) return st;
    if (done || enqueued >= maxiter) break;
    const int64_t batch = check_every < maxiter - enqueued ? check_every : maxiter - enqueued;
    for (int64_t q = 0; q < batch; ++q) {
      if ((st = be.apply(A, v, tmp_u))) return st;                                            // S1
      if ((st = be.pass(LsmrU<T>{u, tmp_u, s}, m))) return st;                                 // M1
      if ((st = be.pass(LsmrScaleU<T>{u, s, 1}, m))) return st;                                // M2
      if ((st = be.apply(At, u, tmp_v))) return st;                                           // S2
      if ((st = be.pass(LsmrV<T>{v, tmp_v, s}, n))) return st;                                 // M3
      if ((st = be.pass(LsmrXH<T>{v, h, hbar, x, s}, n))) return st;                           // M4
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&hs, s, sizeof(hs)))) return st;
  out->iters = hs.iter;
  out->mvps = hs.mvps;
  out->mtvps = hs.mtvps;
  out->istop = hs.istop;
  out->converged = !(hs.istop == 3 || hs.istop == 6 || hs.istop == 7);      // setconv(log, istop ∉ (3, 6, 7)) :285
  out->early = hs.early;
  out->atol = hs.atol; out->btol = hs.btol; out->ctol = hs.ctol;
  out->normr = hs.normr; out->normAr = hs.normAr; out->normA = hs.normA; out->condA = hs.condA; out->normx = hs.normx;
  out->n_hist = hs.n_hist < hist_cap ? hs.n_hist : hist_cap;
  out->hist_stride = hist_cap;
  if (hist_cap > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * 4 * (size_t)hist_cap))) return st;
  return 0;
}

}  // namespace b200
