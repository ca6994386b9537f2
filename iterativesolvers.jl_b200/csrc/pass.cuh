// pass.cuh -- CUDA backend of the fused-pass vocabulary (pass_core.h): one kernel per pass.
//
//   k_pass<P>: grid = multiple of the SM count (stream_grid), 256 threads, grid-stride over the rows so that
//   consecutive threads touch consecutive elements (coalesced 8-byte/4-byte accesses); the NRED sums are
//   reduced warp -> block -> grid in a fixed order (per-block slots; the last block to arrive sums the slots with all
//   its threads in a fixed pattern: run-to-run reproducible), and the finishing thread
//   runs the pass's scalar section.  Passes are HBM-bound streaming kernels: their algorithmic bytes are
//   (vectors read + vectors written) * n * sizeof(T).
#pragma once
#include "blas1.cuh"
#include "csr.cuh"
#include "pass_core.h"

namespace b200 {

constexpr int kPassThreads = 256;

#ifdef __CUDACC__

template <typename P>
__global__ void __launch_bounds__(kPassThreads) k_pass(const P p_in, int64_t n, double *partials,
                                                       unsigned int *ticket, int single) {
  if (p_in.skip()) return;
  P p = p_in;
  p.load();   // the pass's device-resident scalars, read once per thread
  constexpr int NR = P::NRED;
  if constexpr (NR == 0) {
    double dummy[1];
    for (int64_t i = blockIdx.x * (int64_t)kPassThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kPassThreads)
      p.elem(i, dummy);
  } else {
    static_assert(NR <= kPassMaxRed && NR <= kMaxReduceWidth, "too many sums in one pass");
    __shared__ double smem[kPassThreads / 32][NR];
    __shared__ double tot[NR];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kPassThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kPassThreads)
      p.elem(i, acc);
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const double v = warp_sum(acc[j]);
      if (lane == 0) smem[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
      double s = 0.0;
      for (int w = 0; w < kPassThreads / 32; ++w) s += smem[w][threadIdx.x];
      partials[(size_t)blockIdx.x * kMaxReduceWidth + threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // final sum by the whole block in a fixed order: thread t adds slots t, t+256, ... of a column, then the block
    // tree (a single thread walking the ~1200 slots costs ~90 us of exposed L2 latency per pass: measured)
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      double a = 0.0;
      for (unsigned int b = threadIdx.x; b < gridDim.x; b += kPassThreads)
        a += __ldcg(&partials[(size_t)b * kMaxReduceWidth + j]);
      a = warp_sum(a);
      if (lane == 0) smem[warp][j] = a;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
      double s = 0.0;
      for (int w = 0; w < kPassThreads / 32; ++w) s += smem[w][threadIdx.x];
      tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *ticket = 0u;   // re-arm for the next reduction on this stream
      if (single) {
        p.finish(tot);
      } else {
        double *out = p.sums();
        for (int j = 0; j < NR; ++j) out[j] = tot[j];
      }
    }
  }
}

// scalar section of a pass after the cross-GPU allreduce of sums() (multi-GPU contexts), and the
// launch form of ScalarStep
template <typename P>
__global__ void k_pass_finish(P p) {
  if (p.skip()) return;
  p.finish(p.sums());
}

#endif  // __CUDACC__

// What the engines apply: the device CSR operator or a caller-supplied callback (b200_linop).
struct CudaOp {
  const b200_csr *csr = nullptr;
  const b200_linop *fn = nullptr;
};

// The CUDA backend the engines are instantiated with in libb200krylov.so.
struct CudaBackend {
  typedef CudaOp Op;
  b200_ctx *ctx;

  bool single() const { return ctx->world == 1; }

  // y = Op x (b200_csr: halo exchange included on multi-GPU contexts; b200_linop: whatever the callback enqueues)
  int apply(const Op *A, const void *x, void *y) {
    ProfScope prof(ctx, 0);
    if (A->csr) return spmv(ctx, A->csr, x, y);
    ctx->in_callback += 1;
    const int st = A->fn->apply(A->fn->user, x, y, (void *)ctx->stream);
    ctx->in_callback -= 1;
    if (st != 0) {
      set_error("operator / preconditioner callback returned %d", st);
      return B200_ERR_CALLBACK;
    }
    return B200_OK;
  }

#ifdef __CUDACC__
  template <typename P>
  int pass(const P &p, int64_t n) {
    if constexpr (P::NRED == 0) {
      if (n <= 0) return B200_OK;
      ProfScope prof(ctx, 2);
      k_pass<P><<<stream_grid(ctx, n, kPassThreads * 2, 8), kPassThreads, 0, ctx->stream>>>(
          p, n, ctx->red.partials, ctx->red.ticket, 1);
      B200_LAUNCH_CHECK(ctx);
      return B200_OK;
    } else {
      {
        ProfScope prof(ctx, 1);
        // n == 0 (empty slab) still launches one block: the totals (zeros) and the scalar section are needed
        k_pass<P><<<stream_grid(ctx, n, kPassThreads * 2, 8), kPassThreads, 0, ctx->stream>>>(
            p, n, ctx->red.partials, ctx->red.ticket, single() ? 1 : 0);
      }
      B200_LAUNCH_CHECK(ctx);
      if (!single()) {
        B200_TRY(allreduce_sum_dev(ctx, p.sums(), P::NRED));
        k_pass_finish<P><<<1, 1, 0, ctx->stream>>>(p);
        B200_LAUNCH_CHECK(ctx);
      }
      return B200_OK;
    }
  }
  template <typename P>
  int scalar(const P &p) {
    k_pass_finish<P><<<1, 1, 0, ctx->stream>>>(p);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
  }
#endif

  int zero(void *x, size_t bytes) {
    if (bytes) B200_CUDA(cudaMemsetAsync(x, 0, bytes, ctx->stream));
    return B200_OK;
  }
  int copy(void *dst, const void *src, size_t bytes) {
    if (bytes && dst != src) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
  }
  int to_device(void *dst, const void *src_host, size_t bytes) {
    B200_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));   // src_host is usually a stack object
    return B200_OK;
  }
  int to_host(void *dst_host, const void *src, size_t bytes) {
    B200_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  // poll a device-resident int (the solver's done flag)
  int read_flag(const int *flag_dev, int *out) {
    B200_CUDA(cudaMemcpyAsync(ctx->h_flags, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_flags[0];
    return B200_OK;
  }
  // grow-only scratch of the context (valid until the next workspace() call on this context)
  int workspace(size_t bytes, void **out) { return ws_get(ctx, bytes, out); }
};

}  // namespace b200
