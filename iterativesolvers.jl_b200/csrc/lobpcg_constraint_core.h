// lobpcg_constraint_core.h -- Constraint of reference src/lobpcg.jl:144-224 (standard problem, B = I: BY aliases Y) as
// fused passes (pass_core.h): the deflation X <- X - Y (chol(Y'Y) \ (Y' X)) that lobpcg! applies to the initial block
// (:868, :875) and, inside every step, to the preconditioned active residual block (precond_constr! :564-569).
//
//   G   = Y' X          one pass per column of Y: 16 sums <y_k, X[:, j]>           (:217)   reads 1 + bs columns
//   tmp = gram_chol \ G  host: two triangular solves with the nc x nc factor        (:219)
//   X  -= Y tmp          one pass per chunk of 16 columns of Y                       (:220-221) reads 16 + bs, writes bs
//
// The block X is addressed with a row and a column stride, so the same passes serve a column-major user block
// (rs = 1, cs = ld) and the row-major n x 16 blocks the LOBPCG engine keeps internally (rs = 16, cs = 1).
// The Gram pass re-reads X once per column of Y (nc x (1 + bs) column reads instead of nc + bs): correct first; a
// register-tiled Gram like k_gram is the obvious next step for wide constraints.
#pragma once
#include "pass_core.h"

namespace b200 {

constexpr int kConBlock = 16;     // widest block (LOBPCG's BS)

template <typename T>
struct ConGramCol {               // g_row[j] = <y, X[:, j]>, j < bs
  static constexpr int NRED = kConBlock;
  const T *y;                     // one column of BY (contiguous)
  const T *X;
  int64_t rs, cs;
  int bs;
  double *g_row;                  // kConBlock doubles (device): the result row, also the allreduce buffer
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    const double yv = (double)y[i];
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) acc[j] += yv * (double)X[i * rs + j * cs];
  }
  B200_HD double *sums() const { return g_row; }
  B200_HD void finish(const double *tot) const {
    for (int j = 0; j < kConBlock; ++j) g_row[j] = tot[j];
  }
};

template <typename T>
struct ConUpdate {                // X[i, j] -= sum_{k < kc} Y[i, k] coef[k][j]
  static constexpr int NRED = 0;
  const T *Y;                     // first column of the chunk, column-major with leading dimension ldy
  int64_t ldy;
  int kc;
  T *X;
  int64_t rs, cs;
  int bs;
  T coef[kConBlock][kConBlock];
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *) const {
    T yv[kConBlock];
    B200_UNROLL
    for (int k = 0; k < kConBlock; ++k) yv[k] = k < kc ? Y[i + k * ldy] : (T)0;
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) {
        T t = (T)0;                                    // mul!(X_temp, Y, tmp) :220
        B200_UNROLL
        for (int k = 0; k < kConBlock; ++k)
          if (k < kc) t = t + yv[k] * coef[k][j];
        X[i * rs + j * cs] = X[i * rs + j * cs] - t;   // X .= X .- X_temp :221
      }
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// G (nc x kConBlock, row-major, device g_dev -> host g_host) = BY' X
template <typename T, typename B>
int constraint_gram(B &be, const T *BY, int64_t ldy, int nc, const T *X, int64_t rs, int64_t cs, int bs, int64_t n,
                    double *g_dev, double *g_host) {
  int st;
  for (int k = 0; k < nc; ++k)
    if ((st = be.pass(ConGramCol<T>{BY + (int64_t)k * ldy, X, rs, cs, bs, g_dev + (size_t)k * kConBlock}, n))) return st;
  if (nc > 0 && (st = be.to_host(g_host, g_dev, sizeof(double) * (size_t)nc * kConBlock))) return st;
  return 0;
}

// X -= Y coef  (coef: nc x kConBlock, row-major, host)
template <typename T, typename B>
int constraint_update(B &be, const T *Y, int64_t ldy, int nc, const double *coef, T *X, int64_t rs, int64_t cs, int bs,
                      int64_t n) {
  int st;
  for (int k0 = 0; k0 < nc; k0 += kConBlock) {
    ConUpdate<T> u;
    u.Y = Y + (int64_t)k0 * ldy;
    u.ldy = ldy;
    u.kc = nc - k0 < kConBlock ? nc - k0 : kConBlock;
    u.X = X;
    u.rs = rs;
    u.cs = cs;
    u.bs = bs;
    for (int k = 0; k < kConBlock; ++k)
      for (int j = 0; j < kConBlock; ++j)
        u.coef[k][j] = (k < u.kc && j < bs) ? (T)coef[(size_t)(k0 + k) * kConBlock + j] : (T)0;
    if ((st = be.pass(u, n))) return st;
  }
  return 0;
}

// host: upper Cholesky factor U (nc x nc, column-major, ld = nc) of the Hermitian matrix whose upper triangle is in G
// (column-major); returns 0, or j+1 when the leading minor of order j+1 is not positive definite (PosDefException).
// Pivot rule: LAPACK potrf (the reference's cholesky!, src/lobpcg.jl:181-182) rejects a pivot d only when !(d > 0);
// for a Gram matrix that is singular up to rounding the sign of d is decided by the order of the roundings (FMA
// contraction), so acceptance there is arbitrary and the accepted factor turns constraint_apply into noise.  This
// engine rejects every pivot that is not above the rounding level of its own cancellation,
//     !(d > 4 nc eps G[j][j])      (G[j][j] the diagonal entry before elimination),
// a superset of potrf's failure set: every input potrf rejects is rejected here, and the additional rejections are
// Gram matrices whose Cholesky factor carries no correct digits in its last pivot.  Documented in b200krylov.h.
inline int con_cholesky_upper(double *G, int nc) {
  const double rel = 4.0 * (double)nc * 2.220446049250313e-16;
  for (int j = 0; j < nc; ++j) {
    const double g0 = G[j + (size_t)j * nc];
    double d = g0;
    for (int k = 0; k < j; ++k) d -= G[k + (size_t)j * nc] * G[k + (size_t)j * nc];
    if (!(d > 0.0) || !(d > rel * g0)) return j + 1;
    d = sqrt(d);
    G[j + (size_t)j * nc] = d;
    for (int c = j + 1; c < nc; ++c) {
      double v = G[j + (size_t)c * nc];
      for (int k = 0; k < j; ++k) v -= G[k + (size_t)j * nc] * G[k + (size_t)c * nc];
      G[j + (size_t)c * nc] = v / d;
    }
  }
  return 0;
}

// host: g (nc x kConBlock row-major) <- (U' U) \ g     ldiv!(tmp, gram_chol, gramYBV) :219
inline void con_chol_solve(const double *U, int nc, double *g, int bs) {
  for (int j = 0; j < bs; ++j) {
    for (int i = 0; i < nc; ++i) {                     // U' z = g  (forward)
      double v = g[(size_t)i * kConBlock + j];
      for (int k = 0; k < i; ++k) v -= U[k + (size_t)i * nc] * g[(size_t)k * kConBlock + j];
      g[(size_t)i * kConBlock + j] = v / U[i + (size_t)i * nc];
    }
    for (int i = nc - 1; i >= 0; --i) {                // U t = z   (backward)
      double v = g[(size_t)i * kConBlock + j];
      for (int k = i + 1; k < nc; ++k) v -= U[i + (size_t)k * nc] * g[(size_t)k * kConBlock + j];
      g[(size_t)i * kConBlock + j] = v / U[i + (size_t)i * nc];
    }
  }
}

// the whole application: X <- X - Y (U'U \ (Y' X))     (constr!::Constraint)(X, X_temp) :212-224
template <typename T, typename B>
int constraint_apply(B &be, const T *Y, int64_t ldy, int nc, const double *U, T *X, int64_t rs, int64_t cs, int bs,
                     int64_t n, double *g_dev, double *g_host) {
  if (nc <= 0 || bs <= 0) return 0;                    // :213
  int st;
  if ((st = constraint_gram<T>(be, Y, ldy, nc, X, rs, cs, bs, n, g_dev, g_host))) return st;
  con_chol_solve(U, nc, g_host, bs);
  return constraint_update<T>(be, Y, ldy, nc, g_host, X, rs, cs, bs, n);
}

// U for a fresh constraint: cholesky!(Hermitian(Y' Y)) :178-182.  U: nc x nc column-major (host), out.
template <typename T, typename B>
int constraint_factor(B &be, const T *Y, int64_t ldy, int nc, int64_t n, double *g_dev, double *g_host, double *U) {
  int st;
  for (int c0 = 0; c0 < nc; c0 += kConBlock) {
    const int bs = nc - c0 < kConBlock ? nc - c0 : kConBlock;
    if ((st = constraint_gram<T>(be, Y, ldy, nc, Y + (int64_t)c0 * ldy, 1, ldy, bs, n, g_dev, g_host))) return st;
    for (int k = 0; k < nc; ++k)
      for (int j = 0; j < bs; ++j) U[k + (size_t)(c0 + j) * nc] = g_host[(size_t)k * kConBlock + j];
  }
  return con_cholesky_upper(U, nc) ? -1 : 0;
}

}  // namespace b200
