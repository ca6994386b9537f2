// svdl_core.h -- svdl(A; nsv, k, j, tol, reltol, maxiter, method = :ritz, vecs, dolock, v0) of reference
// src/svdl.jl:157-247 (build :353-363, extend! :542-609, thickrestart! :376-404, isconverged :290-350) as fused passes
// (pass_core.h): Golub-Kahan-Lanczos bidiagonalisation with full (double classical Gram-Schmidt) reorthogonalisation of
// the right vectors and thick restart on the Ritz values.
//
// Device side (this header): the Lanczos bases P (m x k) and Q (n x (k+1)), and inside extend! every scalar
// (alpha_j, beta_j, the norms that decide whether a second Gram-Schmidt sweep is needed) stays in device memory: the k - l
// steps of one extension are enqueued back to back.  Per step j:
//   S   q = A' p                                                              :565  (operator application)
//   D   h = Q' q  (15 columns per pass) ; ||q||^2 in the 16th sum             :569-570
//   U   q -= Q h  (16 columns per pass) ; ||q||^2 of the result               :570-571 -> second sweep if ||q|| <= ||q_old||/sqrt(2)
//   D', U'  the second sweep: no-ops unless the device flag says so           :571-573
//   B   beta = ||q|| ; q *= 1/beta                                            :576-577
//   S   p = A q                                                               :584
//   X   p -= beta P[:, j] ; alpha = ||p|| ; then p *= 1/alpha                 :585, :596-597
// Host side, once per outer iteration (the reference does the same on the CPU): svd of the k x k projected matrix
// (one-sided Jacobi, dense_svd below), the convergence test, and the coefficients of the restart; the basis rotations
// Q[:, 1:k] V[:, 1:l] and P[:, 1:k] U[:, 1:l] (:384, :392) are passes again (ConUpdate with host coefficients).
// Not computed: g = A'f - alpha q of :400-401 -- L.beta is overwritten by the first step of the following extend!
// (:561-576) before anything reads it.  method = :harmonic (:424-493): the dense part on the host (svdl_harmonic_dense:
// SVD of the broken-arrow matrix, triangular solve, thin QR), the basis rotations and f -= P (P'f) as passes.
#pragma once
#include <vector>

#include "lobpcg_constraint_core.h"   // ConUpdate / kConBlock: X -= Y coef with host coefficients
#include "pass_core.h"

namespace b200 {

constexpr int kSvdlMaxK = 64;          // Lanczos vectors kept (k of the reference); nsv <= k - 1

struct SvdlScal {
  double h[kSvdlMaxK + 1];             // Q' q
  double dv[kSvdlMaxK + 1], ev[kSvdlMaxK + 1];   // alpha_j, beta_j appended by extend! (:599-600), indexed by column
  double beta, alpha, oldnorm2, qnorm2, inv, thr;
  double sum[16];
  int need2, pad;
};

// ---- D: dots of one chunk of (at most 15) columns of Q with q, and ||q||^2
template <typename T>
struct SvdlDots {
  static constexpr int NRED = 16;
  const T *Qc;                         // first column of the chunk
  int64_t ld;
  int nc, c0;                          // columns in the chunk, index of the first
  const T *q;
  int first, second;                   // first chunk of a sweep (records ||q||^2); pass belongs to the second sweep
  SvdlScal *s;
  B200_HD bool skip() const { return second && s->need2 == 0; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const double qi = (double)q[i];
    B200_UNROLL
    for (int c = 0; c < 15; ++c)
      if (c < nc) acc[c] += (double)Qc[i + c * ld] * qi;
    acc[15] += qi * qi;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    for (int c = 0; c < nc; ++c) s->h[c0 + c] = tot[c];
    if (first && !second) s->oldnorm2 = tot[15];       // oldqnorm = norm(q) :569
  }
};

// ---- U: q -= Q[:, chunk] h[chunk] ; ||q||^2 of the result (the value of the LAST chunk is the norm after the sweep)
template <typename T>
struct SvdlOrthUpd {
  static constexpr int NRED = 1;
  const T *Qc;
  int64_t ld;
  int nc, c0;
  T *q;
  int last, second;
  SvdlScal *s;
  T h[16];
  B200_HD bool skip() const { return second && s->need2 == 0; }
  B200_HD void load() {
    for (int c = 0; c < 16; ++c) h[c] = c < nc ? (T)s->h[c0 + c] : (T)0;
  }
  B200_HD void elem(int64_t i, double *acc) const {
    T t = (T)0;                                         // L.Q * (L.Q'q) :570
    B200_UNROLL
    for (int c = 0; c < 16; ++c)
      if (c < nc) t = t + Qc[i + c * ld] * h[c];
    const T v = q[i] - t;
    q[i] = v;
    acc[0] += (double)v * (double)v;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    if (!last) return;
    s->qnorm2 = tot[0];
    if (!second) s->need2 = sqrt(tot[0]) <= s->thr * sqrt(s->oldnorm2);   // norm(q) <= alpha * oldqnorm :571
  }
};

// ---- plain ||x||^2 (no orthogonalisation: the very first step, and ||f|| of the restart)
template <typename T>
struct SvdlNorm {
  static constexpr int NRED = 1;
  const T *x;
  SvdlScal *s;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const { acc[0] += (double)x[i] * (double)x[i]; }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const { s->qnorm2 = tot[0]; }
};

// scalar steps: beta = ||q|| (:576) / alpha = ||f|| (:396); both leave 1/norm in s->inv for the scale pass
B200_HD void svdl_set_beta(SvdlScal *s) {
  s->beta = sqrt(s->qnorm2);
  s->inv = 1.0 / s->beta;
}
typedef ScalarStep<SvdlScal, svdl_set_beta> SvdlBeta;

// alpha = ||f|| of the restart (:396): recorded as the diagonal entry `col` of the new projected matrix (:402)
struct SvdlRestartAlpha {
  static constexpr int NRED = 0;
  SvdlScal *s;
  int col;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t, double *) const {}
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {
    s->alpha = sqrt(s->qnorm2);
    s->inv = 1.0 / s->alpha;
    s->dv[col] = s->alpha;
  }
};

template <typename T>
struct SvdlScale {                     // x .*= inv(norm) :577, :597, :397, :357, :360
  static constexpr int NRED = 0;
  T *x;
  const SvdlScal *s;
  T inv;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { inv = (T)s->inv; }
  B200_HD void elem(int64_t i, double *) const { x[i] = x[i] * inv; }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- X: p -= beta P[:, j] ; alpha = ||p|| ; records (alpha, beta) as the new entries of B (:585, :596, :599-600)
template <typename T>
struct SvdlAxpyNorm {
  static constexpr int NRED = 1;
  T *p;
  const T *pprev;
  int col;                             // column index the new p will have in P (0-based) = position of alpha on the diagonal
  SvdlScal *s;
  T beta;
  B200_HD bool skip() const { return false; }
  B200_HD void load() { beta = (T)s->beta; }
  B200_HD void elem(int64_t i, double *acc) const {
    const T v = p[i] - beta * pprev[i];
    p[i] = v;
    acc[0] += (double)v * (double)v;
  }
  B200_HD double *sums() const { return s->sum; }
  B200_HD void finish(const double *tot) const {
    s->alpha = sqrt(tot[0]);
    s->inv = 1.0 / s->alpha;
    s->dv[col] = s->alpha;             // push!(L.B.dv, alpha)
    s->ev[col] = s->beta;              // push!(L.B.ev, beta): B[col-1, col]
  }
};

// =====================================================================================================================
// host: dense SVD of a small square matrix by one-sided Jacobi (Hestenes).  A (n x n, column-major) -> U, S (descending),
// V with A = U diag(S) V'.  Stands where the reference calls LAPACK through svd(L.B) (:192; "XXX This can be much
// faster", :67).  Singular vectors are determined up to a common sign per pair, which the algorithm is invariant to.
inline void dense_svd(const std::vector<double> &A, int n, std::vector<double> &U, std::vector<double> &S,
                      std::vector<double> &V) {
  std::vector<double> W(A);                              // columns rotated until mutually orthogonal: W = A V
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < n; ++i) {
          a += W[i + (size_t)p * n] * W[i + (size_t)p * n];
          b += W[i + (size_t)q * n] * W[i + (size_t)q * n];
          c += W[i + (size_t)p * n] * W[i + (size_t)q * n];
        }
        if (c == 0.0) continue;
        off = fmax(off, fabs(c) / sqrt(a * b));
        const double zeta = (b - a) / (2.0 * c);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < n; ++i) {
          const double wp = W[i + (size_t)p * n], wq = W[i + (size_t)q * n];
          W[i + (size_t)p * n] = cs * wp - sn * wq;
          W[i + (size_t)q * n] = sn * wp + cs * wq;
          const double vp = V[i + (size_t)p * n], vq = V[i + (size_t)q * n];
          V[i + (size_t)p * n] = cs * vp - sn * vq;
          V[i + (size_t)q * n] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  std::vector<double> s(n);
  std::vector<int> perm(n);
  for (int j = 0; j < n; ++j) {
    double t = 0;
    for (int i = 0; i < n; ++i) t += W[i + (size_t)j * n] * W[i + (size_t)j * n];
    s[j] = sqrt(t);
    perm[j] = j;
  }
  for (int a = 0; a < n; ++a)                            // selection sort, descending (stable enough for n <= 64)
    for (int b = a + 1; b < n; ++b)
      if (s[perm[b]] > s[perm[a]]) { const int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
  U.assign((size_t)n * n, 0.0);
  S.assign(n, 0.0);
  std::vector<double> Vs((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    const int pj = perm[j];
    S[j] = s[pj];
    for (int i = 0; i < n; ++i) {
      Vs[i + (size_t)j * n] = V[i + (size_t)pj * n];
      U[i + (size_t)j * n] = s[pj] > 0 ? W[i + (size_t)pj * n] / s[pj] : (i == j ? 1.0 : 0.0);
    }
  }
  V.swap(Vs);
}

// host: thin QR of a small column-major rows x cols matrix (rows >= cols) by Householder reflections: Q (rows x cols,
// orthonormal columns), R (cols x cols, upper).  Stands where the reference calls qr(M2) (src/svdl.jl:469); the signs of
// the columns of Q / rows of R are a convention the restart is invariant to.
inline void dense_qr_thin(std::vector<double> Awork, int rows, int cols, std::vector<double> &Q, std::vector<double> &R) {
  std::vector<double> V((size_t)rows * cols, 0.0), tau(cols, 0.0);
  for (int c = 0; c < cols; ++c) {
    double nrm = 0.0;
    for (int i = c; i < rows; ++i) nrm += Awork[i + (size_t)c * rows] * Awork[i + (size_t)c * rows];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    const double a = Awork[c + (size_t)c * rows];
    const double beta = a >= 0 ? -nrm : nrm;
    for (int i = c; i < rows; ++i) V[i + (size_t)c * rows] = Awork[i + (size_t)c * rows];
    V[c + (size_t)c * rows] = a - beta;
    double vn = 0.0;
    for (int i = c; i < rows; ++i) vn += V[i + (size_t)c * rows] * V[i + (size_t)c * rows];
    tau[c] = vn > 0 ? 2.0 / vn : 0.0;
    for (int j = c; j < cols; ++j) {                    // A <- (I - tau v v') A
      double d = 0.0;
      for (int i = c; i < rows; ++i) d += V[i + (size_t)c * rows] * Awork[i + (size_t)j * rows];
      d *= tau[c];
      for (int i = c; i < rows; ++i) Awork[i + (size_t)j * rows] -= d * V[i + (size_t)c * rows];
    }
  }
  R.assign((size_t)cols * cols, 0.0);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i <= j; ++i) R[i + (size_t)j * cols] = Awork[i + (size_t)j * rows];
  Q.assign((size_t)rows * cols, 0.0);
  for (int j = 0; j < cols; ++j) Q[j + (size_t)j * rows] = 1.0;
  for (int c = cols - 1; c >= 0; --c)                   // Q = H_0 H_1 ... applied to the first `cols` columns of I
    for (int j = 0; j < cols; ++j) {
      double d = 0.0;
      for (int i = c; i < rows; ++i) d += V[i + (size_t)c * rows] * Q[i + (size_t)j * rows];
      d *= tau[c];
      for (int i = c; i < rows; ++i) Q[i + (size_t)j * rows] -= d * V[i + (size_t)c * rows];
    }
}

// host: the dense part of harmonicrestart!(A, L, F, j) (reference src/svdl.jl:424-475, :482).  In: B (kk x kk upper,
// column-major), its SVD (U0, S0, V0), beta, j.  Out: Uc (kk x j: L.P * Uc are the new left vectors :473), Qf
// ((kk+1) x (j+1): L.Q * Qf the new right vectors :472), Bnew ((j+1) x (j+1) column-major, its [j, j] entry -- alpha --
// is filled by the caller :482).
inline void svdl_harmonic_dense(const std::vector<double> &Bm, int kk, const std::vector<double> &U0,
                                const std::vector<double> &S0, const std::vector<double> &V0, double beta, int j,
                                std::vector<double> &Uc, std::vector<double> &Qf, std::vector<double> &Bnew) {
  const int m1 = kk + 1;
  std::vector<double> rho(kk);
  for (int i = 0; i < kk; ++i) rho[i] = beta * U0[(kk - 1) + (size_t)i * kk];                    // :431
  std::vector<double> BA((size_t)m1 * m1, 0.0), U2, S2, V2;                                      // [Diagonal(S) rho] padded by a zero row :435
  for (int i = 0; i < kk; ++i) {
    BA[i + (size_t)i * m1] = S0[i];
    BA[i + (size_t)kk * m1] = rho[i];
  }
  dense_svd(BA, m1, U2, S2, V2);                                                                 // :436
  Uc.assign((size_t)kk * j, 0.0);                                                                // U = F0.U * F2.U[:, 1:k] :440
  for (int c = 0; c < j; ++c)
    for (int i = 0; i < kk; ++i) {
      double v = 0.0;
      for (int t = 0; t < kk; ++t) v += U0[i + (size_t)t * kk] * U2[t + (size_t)c * m1];
      Uc[i + (size_t)c * kk] = v;
    }
  std::vector<double> M((size_t)m1 * j, 0.0);                                                    // (blockdiag(F0.V, 1) * F2.V)[:, 1:k] :441-443
  for (int c = 0; c < j; ++c) {
    for (int i = 0; i < kk; ++i) {
      double v = 0.0;
      for (int t = 0; t < kk; ++t) v += V0[i + (size_t)t * kk] * V2[t + (size_t)c * m1];
      M[i + (size_t)c * m1] = v;
    }
    M[kk + (size_t)c * m1] = V2[kk + (size_t)c * m1];                                            // Mend :444
  }
  std::vector<double> r(kk, 0.0);                                                                // r = beta * (L.B \ e_m) :446-462
  bool singular = false;
  for (int i = 0; i < kk; ++i) singular = singular || Bm[i + (size_t)i * kk] == 0.0;
  if (!singular) {
    r[kk - 1] = 1.0;
    for (int i = kk - 1; i >= 0; --i) {
      double acc = r[i];
      for (int t = i + 1; t < kk; ++t) acc -= Bm[i + (size_t)t * kk] * r[t];
      r[i] = acc / Bm[i + (size_t)i * kk];
    }
  } else {                                                                                       // pinv(Matrix(L.B)) * r0 :458
    const double cut = S0[0] * kk * 2.220446049250313e-16;
    for (int i = 0; i < kk; ++i) {
      double v = 0.0;
      for (int t = 0; t < kk; ++t)
        if (S0[t] > cut) v += V0[i + (size_t)t * kk] * U0[(kk - 1) + (size_t)t * kk] / S0[t];
      r[i] = v;
    }
  }
  for (int i = 0; i < kk; ++i) r[i] *= beta;
  std::vector<double> M2((size_t)m1 * (j + 1), 0.0), Rq;                                         // :463-468
  for (int c = 0; c < j; ++c)
    for (int i = 0; i < kk; ++i) M2[i + (size_t)c * m1] = M[i + (size_t)c * m1] + r[i] * M[kk + (size_t)c * m1];
  for (int i = 0; i < kk; ++i) M2[i + (size_t)j * m1] = -r[i];
  M2[kk + (size_t)j * m1] = 1.0;
  dense_qr_thin(M2, m1, j + 1, Qf, Rq);                                                          // :469-470
  const int j1 = j + 1;
  Bnew.assign((size_t)j1 * j1, 0.0);
  // R = R[1:k+1, 1:k] + R[:, k+1] * Mend' :475 ; B = [Diagonal(Sigma) * triu(R') ; 0 ... alpha] :482
  for (int a = 0; a < j; ++a)
    for (int b = a; b < j1; ++b) {
      const double Rba = Rq[b + (size_t)a * j1] + Rq[b + (size_t)j * j1] * M[kk + (size_t)a * m1];   // R[b, a] (0 below the diagonal of Rq)
      Bnew[a + (size_t)b * j1] = S2[a] * Rba;
    }
}

struct SvdlOutcome {
  int64_t iters, mvps, mtvps;
  int converged, kdim;
  double beta;
};

// out[:, 0:l] = In[:, 0:kk] * M[0:kk, 0:l]  (M host, column-major with leading dimension ldm): rotation of a basis
template <typename T, typename B>
int svdl_rotate(B &be, const T *In, int64_t ld, int kk, const double *M, int ldm, int l, T *Out, int64_t ldo,
                int64_t n) {
  int st;
  for (int j0 = 0; j0 < l; j0 += kConBlock) {
    const int bs = l - j0 < kConBlock ? l - j0 : kConBlock;
    for (int j = 0; j < bs; ++j)
      if ((st = be.zero(Out + (int64_t)(j0 + j) * ldo, sizeof(T) * (size_t)n))) return st;
    std::vector<double> coef((size_t)kk * kConBlock, 0.0);            // ConUpdate subtracts: pass -M
    for (int c = 0; c < kk; ++c)
      for (int j = 0; j < bs; ++j) coef[(size_t)c * kConBlock + j] = -M[c + (size_t)(j0 + j) * ldm];
    if ((st = constraint_update<T>(be, In, ld, kk, coef.data(), Out + (int64_t)j0 * ldo, 1, ldo, bs, n))) return st;
  }
  return 0;
}

// extend!(log, A, L, k) :542-609 from l to k; P has l+1 columns and Q has l+1 columns on entry (0-based: p = P[:, l])
template <typename T, typename B>
int svdl_extend(B &be, const typename B::Op *A, const typename B::Op *At, T *P, int64_t ldp, T *Q, int64_t ldq, int64_t m,
                int64_t n, int l, int k, SvdlScal *s, int64_t *mvps, int64_t *mtvps) {
  int st;
  for (int j = l + 1; j <= k; ++j) {                                    // 1-based j of :563: Q has j columns so far
    T *q = Q + (int64_t)j * ldq;
    const T *p = P + (int64_t)(j - 1) * ldp;
    *mtvps += 1;                                                        // :564
    if ((st = be.apply(At, p, q))) return st;                           // :565
    for (int sweep = 0; sweep < 2; ++sweep) {                           // :569-573
      for (int c0 = 0; c0 < j; c0 += 15) {
        const int nc = j - c0 < 15 ? j - c0 : 15;
        if ((st = be.pass(SvdlDots<T>{Q + (int64_t)c0 * ldq, ldq, nc, c0, q, c0 == 0, sweep, s}, n))) return st;
      }
      for (int c0 = 0; c0 < j; c0 += 16) {
        const int nc = j - c0 < 16 ? j - c0 : 16;
        if ((st = be.pass(SvdlOrthUpd<T>{Q + (int64_t)c0 * ldq, ldq, nc, c0, q, c0 + 16 >= j, sweep, s}, n))) return st;
      }
    }
    if ((st = be.scalar(SvdlBeta{s}))) return st;                       // beta = norm(q) :576
    if ((st = be.pass(SvdlScale<T>{q, s}, n))) return st;               // :577
    if (j == k) break;                                                  // :580
    *mvps += 1;                                                         // :582
    T *pn = P + (int64_t)j * ldp;
    if ((st = be.apply(A, q, pn))) return st;                           // :584
    if ((st = be.pass(SvdlAxpyNorm<T>{pn, p, j, s}, m))) return st;     // :585, :596, :599-600
    if ((st = be.pass(SvdlScale<T>{pn, s}, m))) return st;              // :597
  }
  return 0;
}

// A: m x n, At: n x m.  v0: n values (not modified: the reference normalises the caller's vector in place, :357; here
// a copy is).  sigma_host: nsv values.  Uout (m x nsv, ld ldu) / Vout (n x nsv, ld ldv): device, or NULL.
// hist_* (host, may be NULL): ritz maxiter x k, resnorm maxiter x nsv, conv maxiter x nsv (0/1), betas maxiter.
// B_host: k x k column-major, the projected matrix at exit (may be NULL).
template <typename T, typename B>
int svdl_run(B &be, const typename B::Op *A, const typename B::Op *At, int64_t m, int64_t n, const T *v0, int nsv, int k,
             int jkeep, double tol, double reltol, int64_t maxiter, int dolock, double *sigma_host, T *Uout, int64_t ldu,
             T *Vout, int64_t ldv, double *hist_ritz, double *hist_resnorm, int *hist_conv, double *hist_betas,
             double *B_host, SvdlOutcome *out, int method = 0) {
  const int l = nsv;
  const int64_t ldp = (int64_t)((((sizeof(T) * (size_t)(m > 0 ? m : 1)) + 255) / 256 * 256) / sizeof(T));
  const int64_t ldq = (int64_t)((((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256) / sizeof(T));
  const size_t sb = (sizeof(SvdlScal) + 255) / 256 * 256;
  void *ws = nullptr;
  int st = be.workspace(sizeof(T) * (size_t)(2 * ldp * k + 2 * ldq * (k + 1)) + sb, &ws);
  if (st) return st;
  char *w = (char *)ws;
  T *P = (T *)w; w += sizeof(T) * (size_t)ldp * k;
  T *P2 = (T *)w; w += sizeof(T) * (size_t)ldp * k;
  T *Q = (T *)w; w += sizeof(T) * (size_t)ldq * (k + 1);
  T *Q2 = (T *)w; w += sizeof(T) * (size_t)ldq * (k + 1);
  SvdlScal *s = (SvdlScal *)w;
  SvdlScal h;
  memset(&h, 0, sizeof(h));
  h.thr = 0.7071067811865476;                                           // alpha = 1/sqrt(2) :544
  if ((st = be.to_device(s, &h, sizeof(h)))) return st;
  int64_t mvps = 0, mtvps = 0;

  // build(log, A, v0, k) :353-363
  if ((st = be.copy(Q, v0, sizeof(T) * (size_t)n))) return st;
  if ((st = be.pass(SvdlNorm<T>{Q, s}, n))) return st;                  // beta = norm(q) :356
  if ((st = be.scalar(SvdlBeta{s}))) return st;
  if ((st = be.pass(SvdlScale<T>{Q, s}, n))) return st;                 // :357
  if ((st = be.apply(A, Q, P))) return st;                              // p = A*q :358
  if ((st = be.pass(SvdlNorm<T>{P, s}, m))) return st;                  // alpha = norm(p) :359
  if ((st = be.scalar(SvdlBeta{s}))) return st;                         // (alpha lands in s->beta; read back below)
  if ((st = be.pass(SvdlScale<T>{P, s}, m))) return st;                 // :360
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  std::vector<double> Bm((size_t)k * k, 0.0);                           // the projected matrix, dense k x k
  Bm[0] = h.beta;                                                       // Bidiagonal([alpha], [], :U) :361
  int bdim = 1;                                                         // current size of B
  if ((st = svdl_extend<T>(be, A, At, P, ldp, Q, ldq, m, n, 0, k, s, &mvps, &mtvps))) return st;   // :362
  if ((st = be.to_host(&h, s, sizeof(h)))) return st;
  for (int c = 1; c < k; ++c) {
    Bm[c + (size_t)c * k] = h.dv[c];
    Bm[(c - 1) + (size_t)c * k] = h.ev[c];
  }
  bdim = k;
  double beta = h.beta;                                                 // L.beta :607

  std::vector<double> U, S, V;
  int converged = 0;
  int64_t iter = 0;
  for (iter = 1; iter <= maxiter; ++iter) {                             // :188
    dense_svd(Bm, k, U, S, V);                                          // F = svd(L.B) :192
    std::vector<double> Bharm, rho(jkeep, 0.0), coef;
    if (method == 1) {
      // harmonicrestart!(A, L, F, j) :424-493
      std::vector<double> Uc, Qf;
      svdl_harmonic_dense(Bm, k, U, S, V, beta, jkeep, Uc, Qf, Bharm);
      if ((st = svdl_rotate<T>(be, Q, ldq, k + 1, Qf.data(), k + 1, jkeep + 1, Q2, ldq, n))) return st;   // Q = L.Q * Q[:, 1:k+1] :472
      if ((st = svdl_rotate<T>(be, P, ldp, k, Uc.data(), k, jkeep, P2, ldp, m))) return st;               // P = L.P * U[:, 1:k] :473
      { T *t = Q; Q = Q2; Q2 = t; t = P; P = P2; P2 = t; }
      T *f = P + (int64_t)jkeep * ldp;
      if ((st = be.apply(A, Q + (int64_t)jkeep * ldq, f))) return st;                                // f = A * Q[:, k+1] :477
      for (int c0 = 0; c0 < jkeep; c0 += 15) {                                                       // P' f
        const int nc = jkeep - c0 < 15 ? jkeep - c0 : 15;
        if ((st = be.pass(SvdlDots<T>{P + (int64_t)c0 * ldp, ldp, nc, c0, f, c0 == 0, 0, s}, m))) return st;
      }
      for (int c0 = 0; c0 < jkeep; c0 += 16) {                                                       // f -= P * (P' f) :478, ||f||
        const int nc = jkeep - c0 < 16 ? jkeep - c0 : 16;
        SvdlOrthUpd<T> u{P + (int64_t)c0 * ldp, ldp, nc, c0, f, c0 + nc == jkeep, 0, s, {}};
        if ((st = be.pass(u, m))) return st;
      }
      if ((st = be.scalar(SvdlRestartAlpha{s, jkeep}))) return st;                                   // alpha = norm(f) :479
      if ((st = be.pass(SvdlScale<T>{f, s}, m))) return st;                                          // :480
      // (g = A'f - (g.q) q and L.beta = norm(g), :484-488, are dead like :400-401 of the Ritz restart)
    } else {
    // thickrestart!(A, L, F, j) :376-404
    if ((st = svdl_rotate<T>(be, Q, ldq, k, V.data(), k, jkeep, Q2, ldq, n))) return st;               // :384
    if ((st = be.copy(Q2 + (int64_t)jkeep * ldq, Q + (int64_t)k * ldq, sizeof(T) * (size_t)n))) return st;   // :385
    if ((st = svdl_rotate<T>(be, P, ldp, k, U.data(), k, jkeep, P2, ldp, m))) return st;               // :392
    { T *t = Q; Q = Q2; Q2 = t; t = P; P = P2; P2 = t; }
    T *f = P + (int64_t)jkeep * ldp;
    if ((st = be.apply(A, Q + (int64_t)jkeep * ldq, f))) return st;                                // f = A*Q[:, l+1] :390
    coef.assign((size_t)jkeep * kConBlock, 0.0);
    for (int i = 0; i < jkeep; ++i) {
      rho[i] = beta * U[(k - 1) + (size_t)i * k];                                                  // :391
      coef[(size_t)i * kConBlock] = rho[i];
    }
    if ((st = constraint_update<T>(be, P, ldp, jkeep, coef.data(), f, 1, ldp, 1, m))) return st;   // f -= L.P*rho :395
    if ((st = be.pass(SvdlNorm<T>{f, s}, m))) return st;                                           // :396
    if ((st = be.scalar(SvdlRestartAlpha{s, jkeep}))) return st;
    if ((st = be.pass(SvdlScale<T>{f, s}, m))) return st;                                          // :397
    // (g = A'f - alpha*q and L.beta = norm(g), :400-401, are dead: see the header)
    }
    if ((st = svdl_extend<T>(be, A, At, P, ldp, Q, ldq, m, n, jkeep, k, s, &mvps, &mtvps))) return st;   // :201
    if ((st = be.to_host(&h, s, sizeof(h)))) return st;
    // isconverged(L, F, l, tol, reltol) :290-350 uses the NEW L.beta with the OLD F
    const double beta_new = h.beta;
    std::vector<double> dsig(l), delta(l);
    for (int i = 0; i < l; ++i) delta[i] = dsig[i] = beta_new * fabs(U[(k - 1) + (size_t)i * k]);   // :297-300
    if (l > 1) {                                                                                   // :307-341
      double d = INFINITY;
      for (int i = 0; i < l; ++i)
        for (int q = 0; q < i; ++q) d = fmin(d, fabs(S[i] - S[q]));
      for (int i = 0; i < l; ++i)
        if (2 * dsig[i] <= d) delta[i] = fmin(delta[i], dsig[i] * dsig[i] / d);                     // :320-328
    }
    const double thresh = fmax(tol, reltol * S[0]);                                                // :349
    int all = 1;
    std::vector<int> conv(l);
    for (int i = 0; i < l; ++i) {
      conv[i] = delta[i] < thresh;
      all = all && conv[i];
    }
    if (hist_ritz) for (int i = 0; i < k; ++i) hist_ritz[(size_t)(iter - 1) * k + i] = S[i];        // :209
    if (hist_resnorm) for (int i = 0; i < l; ++i) hist_resnorm[(size_t)(iter - 1) * l + i] = delta[i];   // :348
    if (hist_conv) for (int i = 0; i < l; ++i) hist_conv[(size_t)(iter - 1) * l + i] = conv[i];     // :208
    if (hist_betas) hist_betas[iter - 1] = beta_new;                                               // :211
    // the new projected matrix: BrokenArrowBidiagonal([S[1:j]; alpha], rho, []) :402 extended by extend! :599-600
    // (built after the convergence test only because the test needs the old U; the values are the same)
    std::fill(Bm.begin(), Bm.end(), 0.0);
    if (method == 1) {                                                                             // UpperTriangular(...) :482
      for (int b = 0; b <= jkeep; ++b)
        for (int a = 0; a <= b && a < jkeep; ++a) Bm[a + (size_t)b * k] = Bharm[a + (size_t)b * (jkeep + 1)];
    } else {
      for (int i = 0; i < jkeep; ++i) {
        Bm[i + (size_t)i * k] = S[i];
        Bm[i + (size_t)jkeep * k] = rho[i];
      }
    }
    Bm[jkeep + (size_t)jkeep * k] = h.dv[jkeep];                                                   // alpha of the restart
    for (int c = jkeep + 1; c < k; ++c) {
      Bm[c + (size_t)c * k] = h.dv[c];
      Bm[(c - 1) + (size_t)c * k] = h.ev[c];
    }
    if (dolock && method == 0)                                                                     // :214-221
      for (int i = 0; i < l && i < jkeep; ++i)
        if (conv[i]) Bm[i + (size_t)jkeep * k] = 0.0;
    beta = beta_new;
    if (all) {                                                                                     // :222
      converged = 1;
      break;
    }
  }
  if (iter > maxiter) iter = maxiter;
  for (int i = 0; i < l; ++i) sigma_host[i] = S.empty() ? 0.0 : S[i];                               // values = F.S[1:l] :227
  if (Uout && !U.empty() && (st = svdl_rotate<T>(be, P, ldp, k, U.data(), k, l, Uout, ldu, m))) return st;   // L.P*F.U[:,1:l] :231
  if (Vout && !V.empty() && (st = svdl_rotate<T>(be, Q, ldq, k, V.data(), k, l, Vout, ldv, n))) return st;   // L.Q[:,1:k]*F.V[:,1:l] :237
  (void)bdim;
  if (B_host) memcpy(B_host, Bm.data(), sizeof(double) * (size_t)k * k);
  out->iters = maxiter > 0 ? iter : 0;
  out->mvps = mvps;
  out->mtvps = mtvps;
  out->converged = converged;
  out->kdim = k;
  out->beta = beta;
  return 0;
}

}  // namespace b200
