// gen.cu -- host generators of the reference's test matrices (inputs for tests and bench.py).
// laplace_matrix(T, N, dims) of reference test/laplace_matrix.jl:1-12: D = tridiag(-1,2,-1),
// A_d = kron(A_{d-1}, I_N) + kron(I, D)  ==  sum over the dims coordinates of a 1-D second
// difference: diagonal 2*dims, -1 for every +-1 neighbour along one coordinate.  Row/column index
// = sum_d coord_d * N^d.  Entries inside a column (row) are emitted in ascending index order, which
// is the order Julia's sparse kron/+ produce.
#include <math.h>
#include <omp.h>

#include "common.cuh"

using namespace b200;

namespace {
struct Geom {
  int64_t N, n, stride[8];
  int dims;
};
bool make_geom(int64_t N, int dims, Geom *g) {
  if (N < 1 || dims < 1 || dims > 6) return false;
  g->N = N;
  g->dims = dims;
  g->n = 1;
  for (int d = 0; d < dims; ++d) {
    g->stride[d] = g->n;
    g->n *= N;
  }
  return true;
}
inline int row_count(const Geom &g, int64_t q) {
  int c = 1;
  int64_t rem = q;
  for (int d = 0; d < g.dims; ++d) {
    const int64_t x = rem % g.N;
    rem /= g.N;
    c += (x > 0) + (x < g.N - 1);
  }
  return c;
}
template <typename I, typename F>
inline void emit_row(const Geom &g, int64_t q, int64_t base, I *idx, F *val) {
  int64_t coord[8], rem = q;
  for (int d = 0; d < g.dims; ++d) {
    coord[d] = rem % g.N;
    rem /= g.N;
  }
  int k = 0;
  for (int d = g.dims - 1; d >= 0; --d)
    if (coord[d] > 0) { idx[k] = (I)(q - g.stride[d] + base); val[k] = (F)-1; ++k; }
  idx[k] = (I)(q + base); val[k] = (F)(2 * g.dims); ++k;
  for (int d = 0; d < g.dims; ++d)
    if (coord[d] < g.N - 1) { idx[k] = (I)(q + g.stride[d] + base); val[k] = (F)-1; ++k; }
}
}  // namespace

extern "C" {

int64_t b200_gen_laplace_nnz(int64_t N, int dims, int64_t row_begin, int64_t m_local) {
  Geom g;
  if (!make_geom(N, dims, &g) || row_begin < 0 || m_local < 0 || row_begin + m_local > g.n) {
    set_error("b200_gen_laplace_nnz: bad arguments");
    return B200_ERR_INVALID;
  }
  int64_t total = 0;
#pragma omp parallel for reduction(+ : total) schedule(static)
  for (int64_t i = 0; i < m_local; ++i) total += row_count(g, row_begin + i);
  return total;
}

int64_t b200_gen_laplace_csc_i64(int64_t N, int dims, int base, int64_t *colptr, int64_t *rowval, double *nzval) {
  Geom g;
  if (!make_geom(N, dims, &g) || !colptr || !rowval || !nzval) {
    set_error("b200_gen_laplace_csc_i64: bad arguments");
    return B200_ERR_INVALID;
  }
  colptr[0] = base;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < g.n; ++q) colptr[q + 1] = row_count(g, q);
  for (int64_t q = 0; q < g.n; ++q) colptr[q + 1] += colptr[q];
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < g.n; ++q) emit_row<int64_t, double>(g, q, base, rowval + (colptr[q] - base), nzval + (colptr[q] - base));
  return colptr[g.n] - base;
}

// advection_dominated(N, beta) of reference benchmark/advection_diffusion.jl:3-30 as SparseMatrixCSC{Float64,Int64}:
//   A = laplace_matrix(Float64, N, 3) ./ -h^2 + kron(I_{N^2}, spdiagm(-1 => -beta/2h, 1 => beta/2h)),  h = 1/(N+1)
// and the right-hand side b = f(x,y,z) = exp(xyz) sin(pi x) sin(pi y) sin(pi z) on the interior points
// (x fastest).  Values are formed with the same floating-point operations as the reference expression
// (stored value / -(h*h), then + the first-derivative coefficient on the x-neighbours).
int64_t b200_gen_advection_csc_i64(int64_t N, double beta, int base, int64_t *colptr, int64_t *rowval, double *nzval,
                                   double *b) {
  Geom g;
  if (!make_geom(N, 3, &g) || !colptr || !rowval || !nzval) {
    set_error("b200_gen_advection_csc_i64: bad arguments");
    return B200_ERR_INVALID;
  }
  const double h = 1.0 / (double)(N + 1);
  const double mh2 = -(h * h);
  const double lo = -beta / (2 * h), up = beta / (2 * h);
  const double v_diag = 6.0 / mh2, v_off = -1.0 / mh2;
  colptr[0] = base;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < g.n; ++q) colptr[q + 1] = row_count(g, q);
  for (int64_t q = 0; q < g.n; ++q) colptr[q + 1] += colptr[q];
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < g.n; ++q) {
    int64_t *idx = rowval + (colptr[q] - base);
    double *val = nzval + (colptr[q] - base);
    const int64_t cx = q % N, cy = (q / N) % N, cz = q / (N * N);
    int k = 0;
    if (cz > 0) { idx[k] = q - N * N + base; val[k++] = v_off; }
    if (cy > 0) { idx[k] = q - N + base; val[k++] = v_off; }
    if (cx > 0) { idx[k] = q - 1 + base; val[k++] = v_off + up; }       // A[q-1, q]: super-diagonal of row q-1
    idx[k] = q + base; val[k++] = v_diag;
    if (cx < N - 1) { idx[k] = q + 1 + base; val[k++] = v_off + lo; }   // A[q+1, q]: sub-diagonal of row q+1
    if (cy < N - 1) { idx[k] = q + N + base; val[k++] = v_off; }
    if (cz < N - 1) { idx[k] = q + N * N + base; val[k++] = v_off; }
    if (b) {
      const double x = (double)(cx + 1) / (double)(N + 1), y = (double)(cy + 1) / (double)(N + 1),
                   z = (double)(cz + 1) / (double)(N + 1);
      const double pi = 3.141592653589793;
      b[q] = exp(x * y * z) * sin(pi * x) * sin(pi * y) * sin(pi * z);
    }
  }
  return colptr[g.n] - base;
}

int64_t b200_gen_laplace_csr_slab_i32(int64_t N, int dims, int64_t row_begin, int64_t m_local, int32_t *rowptr,
                                      int32_t *colind_global, double *vals) {
  Geom g;
  if (!make_geom(N, dims, &g) || !rowptr || !colind_global || !vals || row_begin < 0 || m_local < 0 ||
      row_begin + m_local > g.n || g.n >= INT32_MAX) {
    set_error("b200_gen_laplace_csr_slab_i32: bad arguments");
    return B200_ERR_INVALID;
  }
  rowptr[0] = 0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m_local; ++i) rowptr[i + 1] = row_count(g, row_begin + i);
  for (int64_t i = 0; i < m_local; ++i) rowptr[i + 1] += rowptr[i];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m_local; ++i) emit_row<int32_t, double>(g, row_begin + i, 0, colind_global + rowptr[i], vals + rowptr[i]);
  return rowptr[m_local];
}

}  // extern "C"
