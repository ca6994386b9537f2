// idrs_core.h -- idrs!(x, A, b; s, Pl, abstol, reltol, maxiter, smoothing) of reference src/idrs.jl:49-64
// (idrs_iterable! :112-145, iterate :163-272, omega :70-81) written as fused passes (pass_core.h).
//
// One call of the reference's iterate is one "step"; steps cycle k = 1..s (a new direction pair U_k, G_k in G_j) and
// s+1 (the polynomial step into G_{j+1}).  The small system (M, f, c, omega) lives in device memory (IdrsScal).
//
// step k in 1..s:
//   D   k == 1: f = P' R                                  :177-181   (s+1 reads, s sums in one pass)
//       c = LowerTriangular(M[k:s,k:s]) \ f[k:s]           :186       (scalar section)
//   V   U_k = sum c_i U_i + omega * (Pl \ (R - sum c_i G_i))   :187-201   (2(s-k+1)+1 reads, 1 write; V, Q not stored)
//   S   G_k = A U_k                                        :202       (SpMV)
//   E   for i < k: alpha = <P_i, G_k>/M[i,i]; G_k -= alpha G_i; U_k -= alpha U_i   :206-210
//       -- each update pass also produces the next dot, the last one the new column M[k:s, k] = P[k:s]' G_k  :214-216
//          (k passes instead of 2(k-1)+1)
//   X   beta = f_k/M[k,k]; R -= beta G_k; X += beta U_k; ||R||   :220-224   (4 reads, 2 writes)
//       f[k+1:s] -= beta M[k+1:s,k]                         :235-237   (scalar section)
// step s+1:
//   W   V = Pl \ R (Identity: V is R itself) ; Q = A V      :243-248   (SpMV)
//   O   ||R||, ||Q||, <Q,R> -> omega                        :249, :70-81   (2 reads, 3 sums)
//   X'  R -= omega Q; X += omega V; ||R||                   :250-253   (4 reads, 2 writes; Identity: 3 reads)
// residual smoothing (:225-234, :254-263): X/X' also produce <R_s,T_s> and <T_s,T_s> with T_s = R_s - R (not stored);
//   one more pass applies gamma to R_s, X_s and reduces ||R_s||.
//
// P (the shadow space, n x s) comes from the caller: the reference draws it with rand! (:132), the host binding does
// the draw so that runs are reproducible.  Scalars are fp64 also for Float32 vectors.
#pragma once
#include "pass_core.h"

namespace b200 {

constexpr int kIdrsMaxS = 16;

struct IdrsScal {
  double M[kIdrsMaxS * kIdrsMaxS];     // column-major with leading dimension kIdrsMaxS  :138
  double f[kIdrsMaxS], c[kIdrsMaxS];   // :139-140 ; c[j] multiplies direction k+j of the current step
  double omega, normR, tol, abstol, reltol, beta, gamma, alpha;
  double sum[kIdrsMaxS];
  double *hist;
  long long hist_cap, n_hist;
  long long iter, maxiter;             // iter: the reference's counter, starts at 1 (:163)
  int s, smoothing, done, breakdown, is_f32, pad;
};

B200_HD double &idrs_M(IdrsScal *q, int i, int j) { return q->M[i + kIdrsMaxS * j]; }   // 0-based (i, j)

// c = LowerTriangular(M[k:s,k:s]) \ f[k:s]  (k 0-based) :186
B200_HD void idrs_solve_c(IdrsScal *q, int k) {
  const int s = q->s;
  for (int j = k; j < s; ++j) {
    double t = q->f[j];
    for (int l = k; l < j; ++l) t -= idrs_M(q, j, l) * q->c[l - k];
    q->c[j - k] = t / idrs_M(q, j, j);
  }
}

// what every step does after its residual norm is known: history, counter, termination test of the NEXT call :167, :266-271
B200_HD void idrs_end_of_step(IdrsScal *q, double normR) {
  q->normR = normR;
  if (!(normR == normR)) q->breakdown = 1;
  if (q->hist && q->n_hist < q->hist_cap) q->hist[q->n_hist] = normR;   // :268
  q->n_hist += 1;
  q->iter += 1;                                                          // :271
  if (q->normR < q->tol || q->iter > q->maxiter || q->breakdown) q->done = 1;   // :167
}

// ---- initialisation :115-142
template <typename T>
struct IdrsInit {
  static constexpr int NRED = 1;
  const T *b, *ax, *x;
  T *R, *Xs, *Rs;            // Xs, Rs: nullptr without smoothing
  IdrsScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const T r = b[i] - ax[i];                          // :115
    R[i] = r;
    if (Rs) {                                          // :120-121
      Rs[i] = r;
      Xs[i] = x[i];
    }
    acc[0] += (double)r * (double)r;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    IdrsScal *q = s;
    q->normR = sqrt(tot[0]);                           // :116
    q->tol = fmax(q->reltol * q->normR, q->abstol);    // :117
    for (int j = 0; j < kIdrsMaxS; ++j) {
      for (int i = 0; i < kIdrsMaxS; ++i) idrs_M(q, i, j) = (i == j) ? 1.0 : 0.0;   // :138
      q->f[j] = 0.0;                                   // :139
      q->c[j] = 0.0;                                   // :140
    }
    q->omega = 1.0;                                    // :142
    q->iter = 1;
    q->n_hist = 0;
    q->breakdown = !(q->normR == q->normR);
    q->done = (q->normR < q->tol) || (q->iter > q->maxiter) || q->breakdown;   // :167
  }
};

// ---- D: f[i] = dot(P[i], R), then c for step 1 :177-186
template <typename T>
struct IdrsF {
  static constexpr int NRED = kIdrsMaxS;
  const T *P;                // n x s, column-major, leading dimension ld
  int64_t ld;
  const T *R;
  IdrsScal *s;
  int ns;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { ns = s->s; }
  B200_HD void elem(int64_t i, double *acc) const {
    const double r = (double)R[i];
    B200_UNROLL
    for (int j = 0; j < kIdrsMaxS; ++j)     // compile-time trip count: acc[] stays in registers
      if (j < ns) acc[j] += (double)P[i + j * ld] * r;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    for (int j = 0; j < s->s; ++j) s->f[j] = tot[j];
    idrs_solve_c(s, 0);
  }
};

// ---- V: the new direction U_k :187-201 (k 0-based)
template <typename T>
struct IdrsNewU {
  static constexpr int NRED = 0;
  const T *G, *R, *diag;     // diag: Jacobi Pl (nullptr = Identity)
  T *U;
  int64_t ld;
  int k;
  const IdrsScal *s;
  int ns;
  T c[kIdrsMaxS], omega;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    ns = s->s;
    for (int j = 0; j < kIdrsMaxS; ++j) c[j] = (T)s->c[j];
    omega = (T)s->omega;
  }
  B200_HD void elem(int64_t i, double *) const {
    T v = c[0] * G[i + k * ld];                        // :187
    T q = c[0] * U[i + k * ld];                        // :188
    B200_UNROLL
    for (int j = 1; j < kIdrsMaxS; ++j)                // :190-193 (compile-time trip count: c[] stays in registers)
      if (k + j < ns) {
        v = v + c[j] * G[i + (k + j) * ld];
        q = q + c[j] * U[i + (k + j) * ld];
      }
    v = R[i] - v;                                      // :196
    if (diag) v = v / diag[i];                         // ldiv!(Pl, V) :199
    U[i + k * ld] = q + omega * v;                     // :201
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- V with a callback preconditioner: the fused form above is split around ldiv!(Pl, V) (:199)
template <typename T>
struct IdrsVQ {               // V = R - sum c_i G_i ; Q = sum c_i U_i   :187-196
  static constexpr int NRED = 0;
  const T *G, *U, *R;
  T *V, *Q;
  int64_t ld;
  int k;
  const IdrsScal *s;
  int ns;
  T c[kIdrsMaxS];
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {
    ns = s->s;
    for (int j = 0; j < kIdrsMaxS; ++j) c[j] = (T)s->c[j];
  }
  B200_HD void elem(int64_t i, double *) const {
    T v = c[0] * G[i + k * ld];
    T q = c[0] * U[i + k * ld];
    B200_UNROLL
    for (int j = 1; j < kIdrsMaxS; ++j)
      if (k + j < ns) {
        v = v + c[j] * G[i + (k + j) * ld];
        q = q + c[j] * U[i + (k + j) * ld];
      }
    V[i] = R[i] - v;
    Q[i] = q;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};
template <typename T>
struct IdrsUfromVQ {          // U_k = Q + omega * (Pl \ V)   :201
  static constexpr int NRED = 0;
  const T *Q, *PV;
  T *Uk;
  const IdrsScal *s;
  T omega;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { omega = (T)s->omega; }
  B200_HD void elem(int64_t i, double *) const { Uk[i] = Q[i] + omega * PV[i]; }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- E: bi-orthogonalisation sweep :206-216.  One pass = [apply alpha_i] + [the dots the next scalar section needs].
//   upd >= 0: G_k -= alpha G_upd ; U_k -= alpha U_upd with alpha = s->alpha (set by the previous pass)
//   dot0, ndots: sums <P_{dot0+j}, G_k>, j < ndots, of the UPDATED G_k
//   kind 0: the single sum is the numerator of the next alpha (next = dot0)   :207
//   kind 1: the sums are the new column M[k:s, k]; then beta = f_k / M[k,k]    :214-220
template <typename T, int NR>     // NR = 1 for the single-dot passes (fewer registers), kIdrsMaxS for the column of M
struct IdrsOrth {
  static constexpr int NRED = NR;
  const T *P;
  T *G, *U;
  int64_t ld;
  int k, upd, dot0, ndots, kind;
  IdrsScal *s;
  T alpha;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { alpha = (T)s->alpha; }
  B200_HD void elem(int64_t i, double *acc) const {
    T g = G[i + k * ld];
    if (upd >= 0) {
      g = g - alpha * G[i + upd * ld];                 // :208
      G[i + k * ld] = g;
      U[i + k * ld] = U[i + k * ld] - alpha * U[i + upd * ld];   // :209
    }
    const double gd = (double)g;
    B200_UNROLL
    for (int j = 0; j < NR; ++j)
      if (j < ndots) acc[j] += (double)P[i + (dot0 + j) * ld] * gd;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    IdrsScal *q = s;
    if (kind == 0) {
      q->alpha = tot[0] / idrs_M(q, dot0, dot0);       // :207
    } else {
      for (int j = 0; j < ndots; ++j) idrs_M(q, k + j, k) = tot[j];   // :214-216
      q->beta = q->f[k] / idrs_M(q, k, k);             // :220
    }
  }
};

// ---- X / X': the residual and solution update of both kinds of step, with optional smoothing sums
//   step k <= s:  R -= beta G_k ; X += beta U_k                      :221-224
//   step s+1:     R -= omega Q  ; X += omega V   (V = Pl \ R_old)     :250-253
template <typename T>
struct IdrsUpdate {
  static constexpr int NRED = 3;
  T *R, *X;
  const T *dR, *dX;          // G_k, U_k  or  Q, V (V == nullptr: Identity, use the old R)
  const T *Rs;               // smoothing: R_s (nullptr otherwise)
  int k;                     // 0-based step index in [0, s) or -1 for the polynomial step
  IdrsScal *s;
  T coef;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { coef = (T)(k >= 0 ? s->beta : s->omega); }
  B200_HD void elem(int64_t i, double *acc) const {
    const T r_old = R[i];
    const T r = r_old - coef * dR[i];
    R[i] = r;
    X[i] = X[i] + coef * (dX ? dX[i] : r_old);
    acc[0] += (double)r * (double)r;
    if (Rs) {
      const T rs = Rs[i];
      const T t = rs - r;                              // T_s .= R_s .- R :226
      acc[1] += (double)rs * (double)t;                // dot(R_s, T_s) :228
      acc[2] += (double)t * (double)t;                 // dot(T_s, T_s)
    }
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    IdrsScal *q = s;
    if (k >= 0 && k < q->s - 1) {                      // f[k+1:s] .-= beta*M[k+1:s,k] :235-237
      for (int i = k + 1; i < q->s; ++i) q->f[i] -= q->beta * idrs_M(q, i, k);
      idrs_solve_c(q, k + 1);                          // c of the next step :186
    }
    if (q->smoothing) {
      q->gamma = tot[1] / tot[2];                      // :228
      q->normR = sqrt(tot[0]);                         // :224 (overwritten by the smoothing pass, :233)
    } else {
      idrs_end_of_step(q, sqrt(tot[0]));               // :224 / :253
    }
  }
};

// ---- smoothing pass :230-233
template <typename T>
struct IdrsSmooth {
  static constexpr int NRED = 1;
  T *Rs, *Xs;
  const T *R, *X;
  IdrsScal *s;
  T gamma;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() { gamma = (T)s->gamma; }
  B200_HD void elem(int64_t i, double *acc) const {
    const T rs_old = Rs[i];
    const T t = rs_old - R[i];
    const T rs = rs_old - gamma * t;                   // :230
    Rs[i] = rs;
    const T xs = Xs[i];
    Xs[i] = xs - gamma * (xs - X[i]);                  // :231
    acc[0] += (double)rs * (double)rs;                 // :233
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const { idrs_end_of_step(s, sqrt(tot[0])); }
};

// ---- W: V = Pl \ R for the Jacobi preconditioner :243-246
template <typename T>
struct IdrsPrecR {
  static constexpr int NRED = 0;
  const T *R, *diag;
  T *V;
  const IdrsScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *) const { V[i] = R[i] / diag[i]; }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- O: omega(Q, R) :70-81
template <typename T>
struct IdrsOmega {
  static constexpr int NRED = 3;
  const T *Q, *R;
  IdrsScal *s;
  B200_HD bool skip() const { return s->done != 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const double t = (double)Q[i], r = (double)R[i];
    acc[0] += r * r;
    acc[1] += t * t;
    acc[2] += t * r;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    const double angle = 0.7071067811865476;           // sqrt(2.)/2 :71
    const double ns = sqrt(tot[0]), nt = sqrt(tot[1]), ts = tot[2];   // :72-74
    const double rho = fabs(ts / (nt * ns));           // :75
    double om = ts / (nt * nt);                        // :76
    if (rho < angle) om = om * angle / rho;            // :77-79
    s->omega = om;
  }
};

struct IdrsOutcome {
  int64_t iters, n_hist;
  double normR, tol;
  int converged, breakdown;
};

// x, b: n values; P: n x s (leading dimension ldp) drawn by the caller; diag: Jacobi Pl or NULL; Plop: callback
// preconditioner (y = Pl \ x) or NULL (Identity when both are NULL).
template <typename T, typename B>
int idrs_run(B &be, const typename B::Op *A, int64_t n, int64_t n_global, T *x, const T *b, int s_dim, const T *P,
             int64_t ldp, const T *diag, double abstol, double reltol, int64_t maxiter, int smoothing, int check_every,
             int64_t hist_cap, double *hist_host, IdrsOutcome *out, const typename B::Op *Plop = nullptr) {
  if (reltol < 0) reltol = sqrt(eps_of<T>());                               // :53
  if (maxiter < 0) maxiter = n_global;                                      // :54
  if (!hist_host) hist_cap = 0;
  if (hist_cap > maxiter) hist_cap = maxiter;
  const int64_t ld = (int64_t)((((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256) / sizeof(T));
  const size_t vb = sizeof(T) * (size_t)ld;
  const size_t hb = ((sizeof(double) * (size_t)(hist_cap > 0 ? hist_cap : 1)) + 255) / 256 * 256;
  const size_t sb = (sizeof(IdrsScal) + 255) / 256 * 256;
  const bool copy_p = ldp != ld;     // the passes index P, U and G with one leading dimension
  void *ws = nullptr;
  int st = be.workspace(vb * (size_t)(2 * s_dim + 4 + (smoothing ? 2 : 0) + (copy_p ? s_dim : 0)) + sb + hb, &ws);
  if (st) return st;
  char *p = (char *)ws;
  T *U = (T *)p; p += vb * s_dim;
  T *G = (T *)p; p += vb * s_dim;
  T *R = (T *)p; p += vb;
  T *Q = (T *)p; p += vb;
  T *V = (T *)p; p += vb;
  T *V2 = (T *)p; p += vb;                 // Pl \ V with a callback preconditioner
  T *Xs = nullptr, *Rs = nullptr;
  if (smoothing) {
    Xs = (T *)p; p += vb;
    Rs = (T *)p; p += vb;
  }
  const T *Pw = P;
  if (copy_p) {
    T *Pc = (T *)p; p += vb * s_dim;
    for (int j = 0; j < s_dim; ++j)
      if ((st = be.copy(Pc + j * ld, P + j * ldp, sizeof(T) * (size_t)n))) return st;
    Pw = Pc;
  }
  IdrsScal *sc = (IdrsScal *)p; p += sb;
  double *hist = hist_cap ? (double *)p : nullptr;

  IdrsScal h;
  memset(&h, 0, sizeof(h));
  h.abstol = abstol;
  h.reltol = reltol;
  h.maxiter = maxiter;
  h.hist = hist;
  h.hist_cap = hist_cap;
  h.s = s_dim;
  h.smoothing = smoothing;
  h.is_f32 = sizeof(T) == 4;
  if ((st = be.to_device(sc, &h, sizeof(h)))) return st;
  if ((st = be.zero(U, vb * s_dim))) return st;                             // :133
  if ((st = be.zero(G, vb * s_dim))) return st;                             // :134

  if ((st = be.apply(A, x, Q))) return st;                                  // A*X :115
  if ((st = be.pass(IdrsInit<T>{b, Q, x, R, Xs, Rs, sc}, n))) return st;

  auto finish_step = [&]() -> int {
    if (!smoothing) return 0;
    return be.pass(IdrsSmooth<T>{Rs, Xs, R, x, sc}, n);
  };

  if (check_every <= 0) check_every = 16;
  int64_t enqueued = 0;
  int step = 0;        // 0-based: 0..s-1 direction steps, s = polynomial step
  for (;;) {
    int done = 0;
    if ((st = be.read_flag(&sc->done, &done))) return st;
    if (done || enqueued >= maxiter) break;
    const int64_t batch = check_every < maxiter - enqueued ? check_every : maxiter - enqueued;
    for (int64_t it = 0; it < batch; ++it) {
      if (step < s_dim) {
        const int k = step;
        if (k == 0 && (st = be.pass(IdrsF<T>{Pw, ld, R, sc}, n))) return st;                         // D
        if (Plop) {                                                                                   // V, split at ldiv! :199
          if ((st = be.pass(IdrsVQ<T>{G, U, R, V, Q, ld, k, sc}, n))) return st;
          if ((st = be.apply(Plop, V, V2))) return st;
          if ((st = be.pass(IdrsUfromVQ<T>{Q, V2, U + k * ld, sc}, n))) return st;
        } else if ((st = be.pass(IdrsNewU<T>{G, R, diag, U, ld, k, sc}, n))) return st;               // V
        if ((st = be.apply(A, U + k * ld, G + k * ld))) return st;                                    // S :202
        // E: dots and updates interleaved; the last pass yields the new column of M
        if (k == 0) {
          if ((st = be.pass(IdrsOrth<T, kIdrsMaxS>{Pw, G, U, ld, k, -1, k, s_dim - k, 1, sc}, n))) return st;
        } else {
          if ((st = be.pass(IdrsOrth<T, 1>{Pw, G, U, ld, k, -1, 0, 1, 0, sc}, n))) return st;        // <P_1, G_k>
          for (int i = 0; i < k; ++i) {
            if (i + 1 < k) {
              if ((st = be.pass(IdrsOrth<T, 1>{Pw, G, U, ld, k, i, i + 1, 1, 0, sc}, n))) return st;
            } else if (s_dim - k == 1) {
              if ((st = be.pass(IdrsOrth<T, 1>{Pw, G, U, ld, k, i, k, 1, 1, sc}, n))) return st;
            } else {
              if ((st = be.pass(IdrsOrth<T, kIdrsMaxS>{Pw, G, U, ld, k, i, k, s_dim - k, 1, sc}, n))) return st;
            }
          }
        }
        if ((st = be.pass(IdrsUpdate<T>{R, x, G + k * ld, U + k * ld, Rs, k, sc}, n))) return st;     // X
        if ((st = finish_step())) return st;
        step += 1;
      } else {
        const T *Vin = R;
        if (Plop) {
          if ((st = be.apply(Plop, R, V))) return st;                                                 // W by callback :243-246
          Vin = V;
        } else if (diag) {
          if ((st = be.pass(IdrsPrecR<T>{R, diag, V, sc}, n))) return st;                             // W
          Vin = V;
        }
        if ((st = be.apply(A, Vin, Q))) return st;                                                   // :248
        if ((st = be.pass(IdrsOmega<T>{Q, R, sc}, n))) return st;                                     // O
        if ((st = be.pass(IdrsUpdate<T>{R, x, Q, Vin == R ? nullptr : V, Rs, -1, sc}, n))) return st; // X'
        if ((st = finish_step())) return st;
        step = 0;
      }
    }
    enqueued += batch;
  }
  if ((st = be.to_host(&h, sc, sizeof(h)))) return st;
  if (smoothing && (st = be.copy(x, Xs, sizeof(T) * (size_t)n))) return st; // copyto!(X, X_s) :170-172
  out->iters = h.iter - 1;
  out->normR = h.normR;
  out->tol = h.tol;
  out->converged = (0 <= h.normR) && (h.normR < h.tol);                     // :168
  out->breakdown = h.breakdown;
  out->n_hist = out->iters < hist_cap ? out->iters : hist_cap;
  if (out->n_hist > 0 && (st = be.to_host(hist_host, hist, sizeof(double) * (size_t)out->n_hist))) return st;
  return 0;
}

}  // namespace b200
