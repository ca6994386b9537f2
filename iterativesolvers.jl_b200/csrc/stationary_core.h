// stationary_core.h -- jacobi!, gauss_seidel!, sor!, ssor! for sparse matrices (reference src/stationary_sparse.jl:
// JacobiIterable :203-224, GaussSeidelIterable :247-272, SORIterable :293-336, SSORIterable :358-411, with forward_sub!
// :64-102, backward_sub! :107-143, the OffDiagonal mul! :148-173 and gauss_seidel_multiply! :179-210) as fused passes
// (pass_core.h) over the rows of a device CSR operator.  Beyond SURVEY section 8 (the stationary methods of the
// reference's inventory); single GPU.
//
// The reference's column-oriented in-place sweeps are sequential.  Row i of a sweep, written out:
//   Jacobi           x_i  <- (b_i + sum_{j != i, ascending} a_ij (-x_j)) / a_ii                        :213-217
//   Gauss-Seidel     x_i  <- (b_i + sum_{j > i, asc} a_ij (-x_j) - sum_{j < i, asc} a_ij x_j^new) / a_ii :262-266
//   SOR              n_i  <- w (b_i + sum_{j > i} a_ij (-x_j) - sum_{j < i} a_ij n_j) / a_ii + (1 - w) x_i   :310-318
//   SSOR  forward    t_i  <- as SOR                                                                    :394-400
//         backward   x_i  <- w (b_i + sum_{j < i, DESCENDING} a_ij (-t_j) - sum_{j > i, desc} a_ij x_j^new) / a_ii
//                            + (1 - w) t_i                                                             :402-406
// A row depends on the NEW values of its lower (backward sweep: upper) neighbours only, so the rows are grouped into
// dependency levels once (level(i) = 1 + max level of the neighbours it waits for; host analysis, O(nnz)) and each
// level is one pass over its rows: within a level the rows are independent, every row performs the reference's
// arithmetic in the reference's order, and the result is the sequential sweep's, bit for bit.  A 7-point Laplacian on
// N^3 points has 3N - 2 levels.  Gauss-Seidel writes to a second vector and copies back (the reference's in-place
// trick relies on the sequential order; with concurrent rows an unrelated row of the same level could be overwritten
// before a lower-numbered row has read its old value).
#pragma once
#include <vector>

#include "pass_core.h"

namespace b200 {

enum { ST_JACOBI = 0, ST_GAUSS_SEIDEL = 1, ST_SOR = 2, ST_SSOR = 3 };
// OR-ed into the method: the arithmetic of the reference's dense-matrix methods (src/stationary.jl) instead of the sparse
// ones.  Jacobi :48-70 and Gauss-Seidel :108-127 are the same operations; SOR :167-186 writes the relaxation as
// x_i + w (acc / a_ii - x_i); the backward half of SSOR :248-258 reads BOTH triangles with the values of the forward half
// (it subtracts A[row, col] x[col] before x[col] is updated): a Jacobi-like relaxation, rows independent.
enum { ST_DENSE_ARITHMETIC = 16 };

template <typename T, typename RP>
struct CsrView {                       // rows with ascending column indices
  int64_t n;
  const RP *rowptr;
  const int *colind;
  const T *vals;
};

// one row of a sweep (see the table above).  rows: the rows of the current level (nullptr: row k itself).
template <typename T, typename RP>
struct StRow {
  static constexpr int NRED = 0;
  CsrView<T, RP> A;
  const int *rows, *dpos;              // dpos[i]: position of a_ii in the row
  const T *b, *xold, *xnew;            // xold: read for the triangle that does not create dependencies; xnew: the other one
  const T *mix;                        // SOR / SSOR: the (1 - w) term; nullptr: out_i = acc / a_ii
  T *out;
  T omega, one_minus;
  int backward, jacobi, dense;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t k, double *) const {
    const int64_t i = rows ? rows[k] : k;
    const int64_t r0 = (int64_t)A.rowptr[i], r1 = (int64_t)A.rowptr[i + 1], d = r0 + dpos[i];
    T acc = b[i];
    if (jacobi) {
      for (int64_t p = r0; p < r1; ++p)
        if (p != d) acc = acc + A.vals[p] * (-xold[A.colind[p]]);        // y[row] += nzval * (alpha x[col]), alpha = -1 :160-170
    } else if (!backward) {
      for (int64_t p = d + 1; p < r1; ++p) acc = acc + A.vals[p] * (-xold[A.colind[p]]);   // gauss_seidel_multiply!(-1, U, ...) :183-189
      for (int64_t p = r0; p < d; ++p) acc = acc - A.vals[p] * xnew[A.colind[p]];          // forward_sub! :75-77 / :96-98
    } else {
      for (int64_t p = d - 1; p >= r0; --p) acc = acc + A.vals[p] * (-xold[A.colind[p]]);  // gauss_seidel_multiply!(-1, sL, ...) :202-207
      for (int64_t p = r1 - 1; p > d; --p) acc = acc - A.vals[p] * xnew[A.colind[p]];      // backward_sub! :137-139
    }
    const T dv = A.vals[d];
    if (mix && dense)
      out[i] = mix[i] + omega * (acc / dv - mix[i]);                     // x[col] += w (tmp[col] / A[col, col] - x[col])  stationary.jl:179
    else
      out[i] = mix ? omega * acc / dv + one_minus * mix[i]               // alpha x[col] / nzval[idx] + beta y[col] :93 / :134
                   : acc / dv;                                           // :72 ; ldiv!(x, D, next) :31
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// host analysis: diagonal positions (DiagonalIndices :6-27: a missing or zero diagonal entry is a SingularException ->
// returns row + 1) and the dependency levels of the forward and backward sweeps.
struct StLevels {
  std::vector<int> dpos;
  std::vector<int> rows_f, rows_b;     // rows grouped by level
  std::vector<int64_t> lptr_f, lptr_b; // level l owns rows_x[lptr_x[l] .. lptr_x[l+1])
};
template <typename T, typename RP>
int64_t stationary_analyse(int64_t n, const RP *rowptr, const int *colind, const T *vals, bool want_forward,
                           bool want_backward, StLevels *out) {
  out->dpos.assign((size_t)n, 0);
  for (int64_t i = 0; i < n; ++i) {
    int64_t d = -1;
    for (int64_t p = (int64_t)rowptr[i]; p < (int64_t)rowptr[i + 1]; ++p)
      if (colind[p] == i) { d = p; break; }
    if (d < 0 || vals[d] == (T)0) return i + 1;                          // SingularException(col) :19
    out->dpos[(size_t)i] = (int)(d - (int64_t)rowptr[i]);
  }
  auto group = [&](const std::vector<int> &lev, int nlev, std::vector<int> &rows, std::vector<int64_t> &lptr) {
    lptr.assign((size_t)nlev + 1, 0);
    for (int64_t i = 0; i < n; ++i) lptr[(size_t)lev[(size_t)i] + 1] += 1;
    for (int l = 0; l < nlev; ++l) lptr[(size_t)l + 1] += lptr[(size_t)l];
    rows.assign((size_t)n, 0);
    std::vector<int64_t> fill(lptr.begin(), lptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[(size_t)fill[(size_t)lev[(size_t)i]]++] = (int)i;
  };
  std::vector<int> lev((size_t)n, 0);
  if (want_forward) {
    int nlev = 0;
    for (int64_t i = 0; i < n; ++i) {
      int l = 0;
      for (int64_t p = (int64_t)rowptr[i]; p < (int64_t)rowptr[i] + out->dpos[(size_t)i]; ++p)
        l = std::max(l, lev[(size_t)colind[p]] + 1);
      lev[(size_t)i] = l;
      nlev = std::max(nlev, l + 1);
    }
    group(lev, nlev, out->rows_f, out->lptr_f);
  }
  if (want_backward) {
    int nlev = 0;
    for (int64_t i = n - 1; i >= 0; --i) {
      int l = 0;
      for (int64_t p = (int64_t)rowptr[i] + out->dpos[(size_t)i] + 1; p < (int64_t)rowptr[i + 1]; ++p)
        l = std::max(l, lev[(size_t)colind[p]] + 1);
      lev[(size_t)i] = l;
      nlev = std::max(nlev, l + 1);
    }
    group(lev, nlev, out->rows_b, out->lptr_b);
  }
  return 0;
}

// Exactly `maxiter` iterations (the reference's stationary solvers have no stopping test, :224, :272, :336, :411).
// A: device-visible CSR; lv: the host analysis; dpos_dev / rows_f_dev / rows_b_dev: its arrays in device-visible memory.
template <typename T, typename RP, typename B>
int stationary_run(B &be, const CsrView<T, RP> &A, const StLevels &lv, const int *dpos_dev, const int *rows_f_dev,
                   const int *rows_b_dev, T *x, const T *b, int method, double omega, int64_t maxiter) {
  const int64_t n = A.n;
  const int dense = (method & ST_DENSE_ARITHMETIC) != 0;
  method &= ~ST_DENSE_ARITHMETIC;
  if (maxiter < 0) maxiter = 10;                                         // maxiter::Int = 10
  void *ws = nullptr;
  int st = be.workspace(sizeof(T) * (size_t)(n > 0 ? n : 1), &ws);
  if (st) return st;
  T *next = (T *)ws;
  const T w = (T)omega, omw = (T)1 - (T)omega;                           // one(T) - omega :318, :400, :406
  auto sweep = [&](const std::vector<int64_t> &lptr, const int *rows, const T *xold, const T *xnew, const T *mix, T *out,
                   int backward) -> int {
    for (size_t l = 0; l + 1 < lptr.size(); ++l) {
      const int64_t cnt = lptr[l + 1] - lptr[l];
      StRow<T, RP> r{A, rows + lptr[l], dpos_dev, b, xold, xnew, mix, out, w, omw, backward, 0, dense};
      const int s2 = be.pass(r, cnt);
      if (s2) return s2;
    }
    return 0;
  };
  for (int64_t it = 0; it < maxiter; ++it) {
    if (method == ST_JACOBI) {
      StRow<T, RP> r{A, nullptr, dpos_dev, b, x, x, nullptr, next, w, omw, 0, 1, dense};
      if ((st = be.pass(r, n))) return st;                               // next = D \ (b - (A - D) x) :213-217
      if ((st = be.copy(x, next, sizeof(T) * (size_t)n))) return st;
    } else if (method == ST_GAUSS_SEIDEL) {
      if ((st = sweep(lv.lptr_f, rows_f_dev, x, next, nullptr, next, 0))) return st;   // x <- L \ (-U x + b) :262-266
      if ((st = be.copy(x, next, sizeof(T) * (size_t)n))) return st;
    } else if (method == ST_SOR) {
      if ((st = sweep(lv.lptr_f, rows_f_dev, x, next, x, next, 0))) return st;         // :310-315
      if ((st = be.copy(x, next, sizeof(T) * (size_t)n))) return st;                   // s.x, s.next = s.next, s.x :318
    } else if (!dense) {
      if ((st = sweep(lv.lptr_f, rows_f_dev, x, next, x, next, 0))) return st;         // tmp :394-400
      if ((st = sweep(lv.lptr_b, rows_b_dev, next, x, next, x, 1))) return st;         // x :402-406
    } else {
      // dense SSOR, stationary.jl:227-258: the forward half is the dense SOR sweep; the backward half relaxes every row
      // against the forward half's values (both triangles, descending accumulation order) -- no dependencies
      if ((st = sweep(lv.lptr_f, rows_f_dev, x, next, x, next, 0))) return st;
      StRow<T, RP> r{A, nullptr, dpos_dev, b, next, next, next, x, w, omw, 1, 0, 1};
      if ((st = be.pass(r, n))) return st;
    }
  }
  return 0;
}

}  // namespace b200
