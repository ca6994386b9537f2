/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).
 *
 * Plain-C restatement of the CPU path that IterativeSolvers.jl's cg! runs when handed a
 * SparseMatrixCSC{Float64,Int64}: the serial CSC column-scatter SpMV of Julia's SparseArrays
 * stdlib (external to /root/reference; same loop shape as the in-repo OffDiagonal mul!,
 * reference src/stationary_sparse.jl:160-169) and the unfused vector passes of
 * reference src/cg.jl:43-66 (iterate), :72-100 (PCG iterate), :120-155 (cg_iterator!),
 * :209-242 (cg! driver / history counting).
 *
 * PARITY UNPINNED for SpMV/dot/norm values: the reference ships no golden vectors for them
 *  and Julia is not available in this image; summation order of
 * BLAS ddot/dnrm2 inside the reference is unspecified, so this file uses plain left-to-right
 * sums and the tests carry a stated fp64 tolerance (1e-10 relative).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* ---- SparseArrays mul!(y, A::SparseMatrixCSC, x): y = 0; for col: y[rowval[k]] += nzval[k]*x[col] ---- */
EXPORT void oracle_csc_spmv_f64(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                const int64_t *rowval, const double *nzval, int64_t base,
                                const double *x, double *y) {
  for (int64_t i = 0; i < n_rows; ++i) y[i] = 0.0;
  for (int64_t col = 0; col < n_cols; ++col) {
    const double ax = x[col];
    for (int64_t k = colptr[col] - base; k < colptr[col + 1] - base; ++k)
      y[rowval[k] - base] += nzval[k] * ax;
  }
}

EXPORT void oracle_csc_spmv_f32(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                const int64_t *rowval, const float *nzval, int64_t base,
                                const float *x, float *y) {
  for (int64_t i = 0; i < n_rows; ++i) y[i] = 0.0f;
  for (int64_t col = 0; col < n_cols; ++col) {
    const float ax = x[col];
    for (int64_t k = colptr[col] - base; k < colptr[col + 1] - base; ++k)
      y[rowval[k] - base] += nzval[k] * ax;
  }
}

/* ---- SparseArrays mul!(y, adjoint(A)::Adjoint{SparseMatrixCSC}, x): one gather dot per column,
 *      y[col] = sum_k conj(nzval[k]) * x[rowval[k]]   (real eltypes: adjoint == transpose).
 *      Call sites in the reference: src/qmr.jl:76, src/lsqr.jl:132,172, src/lsmr.jl:118,172. ---- */
EXPORT void oracle_csc_spmv_adj_f64(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                    const int64_t *rowval, const double *nzval, int64_t base,
                                    const double *x, double *y) {
  (void)n_rows;
  for (int64_t col = 0; col < n_cols; ++col) {
    double t = 0.0;
    for (int64_t k = colptr[col] - base; k < colptr[col + 1] - base; ++k)
      t += nzval[k] * x[rowval[k] - base];
    y[col] = t;
  }
}
EXPORT void oracle_csc_spmv_adj_f32(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                    const int64_t *rowval, const float *nzval, int64_t base,
                                    const float *x, float *y) {
  (void)n_rows;
  for (int64_t col = 0; col < n_cols; ++col) {
    float t = 0.0f;
    for (int64_t k = colptr[col] - base; k < colptr[col + 1] - base; ++k)
      t += nzval[k] * x[rowval[k] - base];
    y[col] = t;
  }
}

/* SparseArrays mul!(Y, A, X) on n x bs column-major blocks: the stdlib loops the block column
 * outermost, i.e. re-streams A once per column. */
EXPORT void oracle_csc_spmm_f32(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                const int64_t *rowval, const float *nzval, int64_t base,
                                const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t bs) {
  for (int64_t j = 0; j < bs; ++j)
    oracle_csc_spmv_f32(n_rows, n_cols, colptr, rowval, nzval, base, X + j * ldx, Y + j * ldy);
}
EXPORT void oracle_csc_spmm_f64(int64_t n_rows, int64_t n_cols, const int64_t *colptr,
                                const int64_t *rowval, const double *nzval, int64_t base,
                                const double *X, int64_t ldx, double *Y, int64_t ldy, int64_t bs) {
  for (int64_t j = 0; j < bs; ++j)
    oracle_csc_spmv_f64(n_rows, n_cols, colptr, rowval, nzval, base, X + j * ldx, Y + j * ldy);
}

static double dot_f64(int64_t n, const double *a, const double *b) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static double nrm2_f64(int64_t n, const double *a) { return sqrt(dot_f64(n, a, a)); }

/* Result block shared with the Python side (oracle.py mirrors the layout). */
typedef struct {
  int64_t iters;       /* history.iters                         (src/history.jl:54-66) */
  int64_t mvps;        /* history.mvps                                                  */
  int32_t isconverged; /* converged(iterable) at exit            (src/cg.jl:32,238)      */
  int32_t pad;
  double tol;          /* max(reltol*||r0||, abstol)             (src/cg.jl:141)         */
  double residual;     /* iterable.residual at exit                                      */
} oracle_cg_result;

/*
 * cg!(x, A, b; abstol, reltol, maxiter, Pl, initially_zero, log=true).
 *   jacobi_diag == NULL  -> Pl = Identity()  -> CGIterable  (src/cg.jl:43-66)
 *   jacobi_diag != NULL  -> Pl = JacobiPrec  -> PCGIterable (src/cg.jl:72-100), where
 *                           ldiv!(y,P,x) = y .= x ./ P.diagonal  (test/cg.jl:14-18)
 * resnorm must hold maxiter+1 doubles (src/cg.jl:221); resnorm[k] = residual after step k+1.
 */
EXPORT void oracle_cg_f64(int64_t n, const int64_t *colptr, const int64_t *rowval,
                          const double *nzval, int64_t base, double *x, const double *b,
                          double abstol, double reltol, int64_t maxiter, int32_t initially_zero,
                          const double *jacobi_diag, double *resnorm, oracle_cg_result *out) {
  double *u = (double *)calloc((size_t)n, sizeof(double)); /* u .= 0        src/cg.jl:129 */
  double *r = (double *)malloc((size_t)n * sizeof(double));
  double *c = (double *)malloc((size_t)n * sizeof(double));
  memcpy(r, b, (size_t)n * sizeof(double));               /* copyto!(r, b)  src/cg.jl:130 */
  int64_t mv_products = 0;
  if (!initially_zero) {                                   /* src/cg.jl:133-139 */
    mv_products = 1;
    oracle_csc_spmv_f64(n, n, colptr, rowval, nzval, base, x, c);
    for (int64_t i = 0; i < n; ++i) r[i] -= c[i];
  }
  double residual = nrm2_f64(n, r);                        /* src/cg.jl:140 */
  const double tol = fmax(reltol * residual, abstol);      /* src/cg.jl:141 */
  double prev_residual = 1.0;                              /* one(residual) src/cg.jl:146 */
  double rho = 1.0;                                        /* one(eltype(x)) src/cg.jl:151 */
  int64_t iteration = 0, iters = 0, mvps = mv_products;

  for (;;) {
    if (iteration >= maxiter || residual <= tol) break;    /* done()  src/cg.jl:36,45,74 */
    if (jacobi_diag == NULL) {
      const double beta = (residual * residual) / (prev_residual * prev_residual); /* :50 */
      for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];                   /* :51 */
      oracle_csc_spmv_f64(n, n, colptr, rowval, nzval, base, u, c);               /* :54 */
      const double alpha = (residual * residual) / dot_f64(n, u, c);              /* :55 */
      for (int64_t i = 0; i < n; ++i) x[i] += alpha * u[i];                        /* :58 */
      for (int64_t i = 0; i < n; ++i) r[i] -= alpha * c[i];                        /* :59 */
      prev_residual = residual;                                                    /* :61 */
      residual = nrm2_f64(n, r);                                                   /* :62 */
    } else {
      for (int64_t i = 0; i < n; ++i) c[i] = r[i] / jacobi_diag[i];                /* :79 */
      const double rho_prev = rho;                                                 /* :81 */
      rho = dot_f64(n, c, r);                                                      /* :82 */
      const double beta = rho / rho_prev;                                          /* :85 */
      for (int64_t i = 0; i < n; ++i) u[i] = c[i] + beta * u[i];                   /* :86 */
      oracle_csc_spmv_f64(n, n, colptr, rowval, nzval, base, u, c);               /* :89 */
      const double alpha = rho / dot_f64(n, u, c);                                 /* :90 */
      for (int64_t i = 0; i < n; ++i) x[i] += alpha * u[i];                        /* :93 */
      for (int64_t i = 0; i < n; ++i) r[i] -= alpha * c[i];                        /* :94 */
      residual = nrm2_f64(n, r);                                                   /* :96 */
    }
    iteration += 1;
    iters += 1;                         /* nextiter!(history, mvps=1)  src/cg.jl:231 */
    mvps += 1;
    if (resnorm) resnorm[iters - 1] = residual;            /* push!(:resnorm)  src/cg.jl:232 */
  }
  out->iters = iters;
  out->mvps = mvps;
  out->isconverged = residual <= tol;
  out->pad = 0;
  out->tol = tol;
  out->residual = residual;
  free(u); free(r); free(c);
}

/* Fixed-iteration timing leg for bench.py (cpu_baseline / --impl reference): runs exactly
 * `iters` CGIterable steps (no early exit) on the caller's state; returns nothing useful but x. */
EXPORT void oracle_cg_steps_f64(int64_t n, const int64_t *colptr, const int64_t *rowval,
                                const double *nzval, int64_t base, double *x, double *r, double *u,
                                double *c, double *residual_io, double *prev_residual_io,
                                int64_t iters) {
  double residual = *residual_io, prev_residual = *prev_residual_io;
  for (int64_t it = 0; it < iters; ++it) {
    const double beta = (residual * residual) / (prev_residual * prev_residual);
    for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];
    oracle_csc_spmv_f64(n, n, colptr, rowval, nzval, base, u, c);
    const double alpha = (residual * residual) / dot_f64(n, u, c);
    for (int64_t i = 0; i < n; ++i) x[i] += alpha * u[i];
    for (int64_t i = 0; i < n; ++i) r[i] -= alpha * c[i];
    prev_residual = residual;
    residual = nrm2_f64(n, r);
  }
  *residual_io = residual;
  *prev_residual_io = prev_residual;
}

/* NOT reference behaviour -- a stronger CPU baseline for bench.py only: the same CGIterable steps with every
 * loop spread over the host's threads.  Only valid for a SYMMETRIC matrix: the CSC arrays are then read as the
 * CSR of A, which turns the reference's serial column scatter into a row-parallel gather (the reference itself
 * cannot thread its SpMV: SparseArrays' mul! is serial).  Reductions are OpenMP reductions, so the summation
 * order differs from the serial path (checked to 1e-12 in tests/test_oracle.py). */
EXPORT void oracle_cg_steps_f64_omp(int64_t n, const int64_t *colptr, const int64_t *rowval,
                                    const double *nzval, int64_t base, double *x, double *r, double *u,
                                    double *c, double *residual_io, double *prev_residual_io,
                                    int64_t iters) {
  double residual = *residual_io, prev_residual = *prev_residual_io;
  for (int64_t it = 0; it < iters; ++it) {
    const double beta = (residual * residual) / (prev_residual * prev_residual);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];
    double uc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : uc)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int64_t k = colptr[i] - base; k < colptr[i + 1] - base; ++k) s += nzval[k] * u[rowval[k] - base];
      c[i] = s;
      uc += u[i] * s;
    }
    const double alpha = (residual * residual) / uc;
    double rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * u[i];
      const double ri = r[i] - alpha * c[i];
      r[i] = ri;
      rr += ri * ri;
    }
    prev_residual = residual;
    residual = sqrt(rr);
  }
  *residual_io = residual;
  *prev_residual_io = prev_residual;
}

/*
 * laplace_matrix(T, N, dims) as SparseMatrixCSC{T,Int64} (reference test/laplace_matrix.jl:1-12):
 *   A_1 = D = tridiag(-1, 2, -1);  A_d = kron(A_{d-1}, I_N) + kron(I, D)
 * Unrolling the recursion: index = i_d + N*(i_{d-1} + N*(...)) with the LAST kron factor (D of the
 * newest dimension) fastest; entry (p,q) = 2*dims on the diagonal, -1 where p,q differ by +-1 in
 * exactly one coordinate.  CSC columns hold row indices ascending.  Written directly (no kron) so
 * that 512^3 fits in time and memory; validated against the scipy kron restatement in
 * tests/test_oracle_generators.py.  colptr has n+1 entries; returns nnz.
 */
EXPORT int64_t oracle_laplace_nnz(int64_t N, int32_t dims) {
  int64_t n = 1;
  for (int d = 0; d < dims; ++d) n *= N;
  /* each dimension contributes 2*(N-1)*N^(dims-1) off-diagonals */
  return n + (int64_t)dims * 2 * (N - 1) * (n / N);
}

EXPORT int64_t oracle_laplace_csc_f64(int64_t N, int32_t dims, int64_t base, int64_t *colptr,
                                      int64_t *rowval, double *nzval) {
  int64_t n = 1;
  int64_t stride[8];
  for (int d = 0; d < dims; ++d) { stride[d] = n; n *= N; } /* stride[0]=1 fastest */
  /* pass 1: column counts -> colptr */
  colptr[0] = base;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < n; ++q) {
    int64_t cnt = 1, rem = q;
    for (int d = 0; d < dims; ++d) {
      int64_t c = rem % N; rem /= N;
      cnt += (c > 0) + (c < N - 1);
    }
    colptr[q + 1] = cnt;
  }
  for (int64_t q = 0; q < n; ++q) colptr[q + 1] += colptr[q];
  /* pass 2: fill, ascending row index: -stride[dims-1] ... -stride[0], diag, +stride[0] ... */
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < n; ++q) {
    int64_t coord[8], rem = q;
    for (int d = 0; d < dims; ++d) { coord[d] = rem % N; rem /= N; }
    int64_t k = colptr[q] - base;
    for (int d = dims - 1; d >= 0; --d)
      if (coord[d] > 0) { rowval[k] = q - stride[d] + base; nzval[k] = -1.0; ++k; }
    rowval[k] = q + base; nzval[k] = 2.0 * dims; ++k;
    for (int d = 0; d < dims; ++d)
      if (coord[d] < N - 1) { rowval[k] = q + stride[d] + base; nzval[k] = -1.0; ++k; }
  }
  return colptr[n] - base;
}

EXPORT int32_t oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
