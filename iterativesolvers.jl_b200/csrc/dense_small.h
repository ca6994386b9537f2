// dense_small.h -- host-side dense kernels for the O(block^3) pieces of LOBPCG and BiCGStab(l):
// Cholesky (cholesky!), symmetric eigen-decomposition (eigen!(Hermitian(A))) and the generalized
// symmetric-definite problem (eigen!(Hermitian(A), Hermitian(B))), sizes <= 3*blocksize <= 48.
// The reference hands these to LAPACK (potrf / syevd / sygvd, reference src/lobpcg.jl:380,615,622);
// here they are plain fp64 C++ (Householder tridiagonalisation + implicit QL), column-major.
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace b200 {
namespace dense {

// In-place upper Cholesky factor of the symmetric positive definite n x n matrix A (column-major,
// leading dimension lda): A = U' U, U stored in the upper triangle.  Returns 0, or k>0 if the leading
// minor of order k is not positive definite (LAPACK potrf convention -> PosDefException).
inline int cholesky_upper(double *A, int n, int lda) {
  for (int j = 0; j < n; ++j) {
    double d = A[j + j * lda];
    for (int k = 0; k < j; ++k) d -= A[k + j * lda] * A[k + j * lda];
    if (!(d > 0.0)) return j + 1;
    d = sqrt(d);
    A[j + j * lda] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[j + i * lda];
      for (int k = 0; k < j; ++k) s -= A[k + j * lda] * A[k + i * lda];
      A[j + i * lda] = s / d;
    }
  }
  return 0;
}

// Symmetric eigen-decomposition: A (n x n, column-major, full storage, destroyed) -> eigenvalues w
// ascending, eigenvectors in the columns of Z (n x n, column-major).  Householder reduction to
// tridiagonal form (tred2) followed by the implicit QL algorithm (tql2).
inline int sym_eig(std::vector<double> &A, int n, std::vector<double> &w, std::vector<double> &Z) {
  auto a = [&](int i, int j) -> double & { return A[(size_t)i + (size_t)j * n]; };
  std::vector<double> d(n), e(n);
  // ---- tred2 (row-oriented on the lower triangle; A is symmetric so a(i,j) == a(j,i))
  for (int i = n - 1; i > 0; --i) {
    const int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += fabs(a(i, k));
      if (scale == 0.0) {
        e[i] = a(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          a(i, k) /= scale;
          h += a(i, k) * a(i, k);
        }
        double f = a(i, l);
        double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        a(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          a(j, i) = a(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += a(j, k) * a(i, k);
          for (int k = j + 1; k <= l; ++k) g += a(k, j) * a(i, k);
          e[j] = g / h;
          f += e[j] * a(i, j);
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = a(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) a(j, k) -= f * e[k] + g * a(i, k);
        }
      }
    } else {
      e[i] = a(i, l);
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += a(i, k) * a(k, j);
        for (int k = 0; k <= l; ++k) a(k, j) -= g * a(k, i);
      }
    }
    d[i] = a(i, i);
    a(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) a(j, i) = a(i, j) = 0.0;
  }
  // ---- tql2
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return 1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = (r = hypot(f, g));
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = a(k, i + 1);
            a(k, i + 1) = s * a(k, i) + c * f;
            a(k, i) = c * a(k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  // sort ascending
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return d[x] < d[y]; });
  w.resize(n);
  Z.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    w[j] = d[idx[j]];
    for (int k = 0; k < n; ++k) Z[(size_t)k + (size_t)j * n] = a(k, idx[j]);
  }
  return 0;
}

// Generalized symmetric-definite problem A z = w B z (LAPACK sygvd, itype 1): B = U'U, C = U^-T A U^-1,
// C y = w y, z = U^-1 y  (so Z' B Z = I).  A, B: n x n column-major, full symmetric storage (B destroyed).
// Returns 0; >0: B's leading minor of that order is not positive definite; <0: QL did not converge.
inline int sym_eig_generalized(std::vector<double> &A, std::vector<double> &B, int n, std::vector<double> &w,
                               std::vector<double> &Z) {
  const int info = cholesky_upper(B.data(), n, n);
  if (info) return info;
  auto U = [&](int i, int j) -> double { return B[(size_t)i + (size_t)j * n]; };
  // C = U^-T A U^-1: first W = U^-T A (solve U' W = A column by column), then C = W U^-1
  std::vector<double> W((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = A[(size_t)i + (size_t)j * n];
      for (int k = 0; k < i; ++k) s -= U(k, i) * W[(size_t)k + (size_t)j * n];
      W[(size_t)i + (size_t)j * n] = s / U(i, i);
    }
  std::vector<double> Cm((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = W[(size_t)i + (size_t)j * n];
      for (int k = 0; k < j; ++k) s -= Cm[(size_t)i + (size_t)k * n] * U(k, j);
      Cm[(size_t)i + (size_t)j * n] = s / U(j, j);
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double v = 0.5 * (Cm[(size_t)i + (size_t)j * n] + Cm[(size_t)j + (size_t)i * n]);
      Cm[(size_t)i + (size_t)j * n] = Cm[(size_t)j + (size_t)i * n] = v;
    }
  std::vector<double> Y;
  if (sym_eig(Cm, n, w, Y)) return -1;
  Z.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = n - 1; i >= 0; --i) {
      double s = Y[(size_t)i + (size_t)j * n];
      for (int k = i + 1; k < n; ++k) s -= U(i, k) * Z[(size_t)k + (size_t)j * n];
      Z[(size_t)i + (size_t)j * n] = s / U(i, i);
    }
  return 0;
}

}  // namespace dense
}  // namespace b200
