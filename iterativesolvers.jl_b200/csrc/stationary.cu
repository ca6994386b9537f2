// stationary.cu -- jacobi!, gauss_seidel!, sor!, ssor! on a device CSR operator (reference src/stationary_sparse.jl): the
// level-scheduled sweeps of stationary_core.h on the CUDA backend.  One call = host analysis of the sparsity pattern
// (diagonal positions, dependency levels: the reference's DiagonalIndices and the order its column sweeps impose) followed
// by exactly `maxiter` iterations, one kernel per dependency level.
#include "pass.cuh"
#include "stationary_core.h"

using namespace b200;

namespace {

struct DevInts {
  int *p = nullptr;
  ~DevInts() { if (p) cudaFree(p); }
  int upload(const std::vector<int> &h, cudaStream_t st) {
    if (h.empty()) return B200_OK;
    if (cudaMalloc((void **)&p, sizeof(int) * h.size()) != cudaSuccess) {
      set_error("stationary: cudaMalloc failed");
      return B200_ERR_ALLOC;
    }
    B200_CUDA(cudaMemcpyAsync(p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, st));
    return B200_OK;
  }
};

template <typename T>
int stationary_impl(b200_ctx *ctx, const b200_csr *A, T *x, const T *b, int method, double omega, int64_t maxiter) {
  const int64_t n = A->m_local;
  std::vector<int> rowptr((size_t)n + 1), colind((size_t)A->nnz);
  std::vector<T> vals((size_t)A->nnz);
  B200_CUDA(cudaMemcpyAsync(rowptr.data(), A->rowptr, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaMemcpyAsync(colind.data(), A->colind, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaMemcpyAsync(vals.data(), A->vals, sizeof(T) * (size_t)A->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < n; ++i)
    for (int p = rowptr[(size_t)i] + 1; p < rowptr[(size_t)i + 1]; ++p)
      B200_REQUIRE(colind[(size_t)p - 1] < colind[(size_t)p], "stationary methods need rows with ascending column indices "
                   "(row %lld)", (long long)i);
  StLevels lv;
  const bool fwd = method != ST_JACOBI, bwd = method == ST_SSOR;
  const int64_t sing = stationary_analyse<T, int>(n, rowptr.data(), colind.data(), vals.data(), fwd, bwd, &lv);
  if (sing) {
    set_error("SingularException(%lld): zero or missing diagonal entry (reference src/stationary_sparse.jl:19)", (long long)sing);
    return B200_ERR_BREAKDOWN;
  }
  DevInts dpos, rows_f, rows_b;
  B200_TRY(dpos.upload(lv.dpos, ctx->stream));
  B200_TRY(rows_f.upload(lv.rows_f, ctx->stream));
  B200_TRY(rows_b.upload(lv.rows_b, ctx->stream));
  CudaBackend be{ctx};
  const CsrView<T, int> view{n, A->rowptr, A->colind, (const T *)A->vals};
  const int st = stationary_run<T, int>(be, view, lv, dpos.p, rows_f.p, rows_b.p, x, b, method, omega, maxiter);
  B200_CUDA(cudaStreamSynchronize(ctx->stream));     // the index arrays are freed on return
  return st;
}

}  // namespace

extern "C" {

int b200_stationary(b200_ctx *ctx, const b200_csr *A, void *x_dev, const void *b_dev, int method, double omega,
                    int64_t maxiter) {
  B200_REQUIRE(ctx && A && x_dev && b_dev, "NULL argument");
  B200_REQUIRE(A->ctx == ctx, "operator belongs to another context");
  B200_REQUIRE(ctx->world == 1, "the stationary methods sweep the whole matrix in order: single-GPU contexts only");
  B200_REQUIRE(is_square(A), "this solver needs a square operator");
  B200_REQUIRE(method >= B200_STATIONARY_JACOBI && method <= B200_STATIONARY_SSOR, "unknown stationary method %d", method);
  static_assert(B200_STATIONARY_JACOBI == ST_JACOBI && B200_STATIONARY_GAUSS_SEIDEL == ST_GAUSS_SEIDEL &&
                B200_STATIONARY_SOR == ST_SOR && B200_STATIONARY_SSOR == ST_SSOR, "method codes");
  B200_CUDA(cudaSetDevice(ctx->device));
  return A->dtype == B200_F64 ? stationary_impl<double>(ctx, A, (double *)x_dev, (const double *)b_dev, method, omega, maxiter)
                              : stationary_impl<float>(ctx, A, (float *)x_dev, (const float *)b_dev, method, omega, maxiter);
}

}  // extern "C"
