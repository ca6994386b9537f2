// cg.cu -- the (P)CG engine: cg!(x, A, b; ...) of reference src/cg.jl:209-242 as fused device kernels.
//
// One CG iteration (reference src/cg.jl:43-66) = three stream-ordered launches, no host round trip:
//   K1  u = r + beta*u                         (beta = residual^2/prev_residual^2, src/cg.jl:50-51)
//   K2  c = A*u  fused with  dot(u,c)          (src/cg.jl:54-55)     <- the HBM-dominant kernel
//   K3  x += alpha*u ; r -= alpha*c ; ||r||^2  (src/cg.jl:58-62)
// All scalars (residual, prev_residual, alpha, beta, tol, iteration, done) live in device memory
// (struct CgScal); the reductions finish on the device (last-block ticket) and the same block does
// the scalar bookkeeping, including the reference's termination test (src/cg.jl:36).  Kernels of
// iterations enqueued after `done` is set return immediately, so the host only polls the flag every
// `check_every` iterations and results do not depend on that period.
// Algorithmic bytes per iteration: nnz*(V+4) + (n+1)*4 + 11*n*V.
// Multi-GPU: the two sums are ncclAllReduce'd (one double each) and a 1-thread kernel does the
// bookkeeping; the halo exchange precedes K2.
#include "blas1.cuh"
#include "spmv_stream.cuh"

using namespace b200;

namespace {

constexpr int kThreads = 256;

struct CgScal {
  double residual;       // it.residual
  double prev_residual;  // it.prev_residual (CGIterable)
  double rho;            // it.rho (PCGIterable)
  double rho_prev;
  double tol;
  double sum;            // scratch: local/global sum of the reduction in flight
  double sum2;
  double dot_uc;
  double abstol, reltol;
  long long iter;        // iterations completed
  long long maxiter;
  long long hist_cap;
  int done;
  int fixed;             // bench: ignore convergence
  int breakdown;
  int pad;
};

// bookkeeping after ||r||^2 is known (src/cg.jl:61-62 + done() :36); single thread
__device__ __forceinline__ void cg_after_norm(CgScal *s, double rr, double *hist, bool pcg) {
  if (!pcg) s->prev_residual = s->residual;
  const double res = sqrt(rr);
  s->residual = res;
  if (hist && s->iter < s->hist_cap) hist[s->iter] = res;
  s->iter += 1;
  if (!(res == res)) s->breakdown = 1;
  const bool conv = !s->fixed && (res <= s->tol);
  s->done = (s->iter >= s->maxiter) || conv || (!s->fixed && s->breakdown);
}

// initial residual norm known (src/cg.jl:140-141)
__device__ __forceinline__ void cg_after_init_norm(CgScal *s, double rr) {
  const double res = sqrt(rr);
  s->residual = res;
  s->prev_residual = 1.0;
  s->rho = 1.0;
  s->rho_prev = 1.0;
  s->tol = fmax(s->reltol * res, s->abstol);
  s->iter = 0;
  s->breakdown = !(res == res);
  const bool conv = !s->fixed && (res <= s->tol);
  s->done = (0 >= s->maxiter) || conv;
}

enum { FIN_NONE = 0, FIN_INIT = 1, FIN_DOT = 2, FIN_NORM = 3, FIN_NORM_PCG = 4, FIN_RHO = 5 };

__device__ __forceinline__ void cg_finish(int kind, CgScal *s, double total, double *hist, bool single_gpu) {
  if (!single_gpu) {  // multi-GPU: leave the local sum for the allreduce + k_cg_scalar
    s->sum = total;
    return;
  }
  switch (kind) {
    case FIN_INIT: cg_after_init_norm(s, total); break;
    case FIN_DOT: s->dot_uc = total; break;
    case FIN_NORM: cg_after_norm(s, total, hist, false); break;
    case FIN_NORM_PCG: cg_after_norm(s, total, hist, true); break;
    case FIN_RHO: s->rho_prev = s->rho; s->rho = total; break;
    default: break;
  }
}

__global__ void k_cg_scalar(int kind, CgScal *s, double *hist) {
  if (kind != FIN_INIT && s->done) return;  // kernels of iterations past `done` did not produce a sum
  cg_finish(kind, s, s->sum, hist, true);
}

// r = b - c (c = A*x) or r = b; u = 0; ||r||^2     (src/cg.jl:129-140)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_cg_init(const T *__restrict__ b, const T *__restrict__ c, int has_c,
                                                      T *__restrict__ r, T *__restrict__ u, int64_t n, CgScal *s,
                                                      double *partials, unsigned int *ticket, int single_gpu) {
  __shared__ double smem[kThreads / 32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    T ri = b[i];
    if (has_c) ri = ri - c[i];
    r[i] = ri;
    u[i] = (T)0;
    acc += (double)ri * (double)ri;
  }
  acc = block_sum<kThreads>(acc, smem);
  double total;
  if (grid_reduce_finish<kThreads>(acc, partials, ticket, smem, &total) && threadIdx.x == 0)
    cg_finish(FIN_INIT, s, total, nullptr, single_gpu);
}

// K1: u = r + beta*u   (CG: beta = residual^2/prev_residual^2 ; PCG: u = c + (rho/rho_prev)*u)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_cg_update_u(const T *__restrict__ r, T *__restrict__ u, int64_t n,
                                                          const CgScal *__restrict__ s, int pcg) {
  if (s->done) return;
  const double beta_d = pcg ? s->rho / s->rho_prev
                            : (s->residual * s->residual) / (s->prev_residual * s->prev_residual);
  const T beta = (T)beta_d;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    // r .+ beta .* u  -- no FMA contraction, as the reference's broadcast computes it
    if constexpr (sizeof(T) == 8) u[i] = __dadd_rn(r[i], __dmul_rn(beta, u[i]));
    else u[i] = __fadd_rn(r[i], __fmul_rn(beta, u[i]));
  }
}

// K2: c = A*u ; sum u.*c
template <typename T, int LPR>
__global__ void __launch_bounds__(kThreads) k_cg_spmv_dot(const int *__restrict__ rowptr,
                                                          const int *__restrict__ colind,
                                                          const T *__restrict__ vals, XView<T> xv, int64_t m,
                                                          T *__restrict__ c, CgScal *s, double *partials,
                                                          unsigned int *ticket, int single_gpu) {
  if (s->done) return;
  __shared__ double smem[kThreads / 32];
  constexpr int ROWS = kThreads / LPR;
  const int sub = threadIdx.x % LPR;
  const int rib = threadIdx.x / LPR;
  double acc = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * ROWS; base < m; base += (int64_t)gridDim.x * ROWS) {
    const int64_t row = base + rib;
    const bool valid = row < m;
    const T ci = row_dot<T, LPR>(rowptr, colind, vals, xv, valid ? row : (m - 1), sub);
    if (valid && sub == 0) {
      c[row] = ci;
      acc += (double)xv.x[row] * (double)ci;
    }
  }
  acc = block_sum<kThreads>(acc, smem);
  double total;
  if (grid_reduce_finish<kThreads>(acc, partials, ticket, smem, &total) && threadIdx.x == 0)
    cg_finish(FIN_DOT, s, total, nullptr, single_gpu);
}

// K2, TMA-streamed form (spmv_stream.cuh): same result contract as k_cg_spmv_dot
template <typename T>
struct CgDotEpi {
  T *__restrict__ c;
  const T *__restrict__ u;
  double acc;
  __device__ __forceinline__ void operator()(int64_t row, T v) {
    c[row] = v;
    acc += (double)u[row] * (double)v;
  }
};
template <typename T, int LPR>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtasPerSm)
    k_cg_spmv_dot_stream(const int *__restrict__ rowptr, const int *__restrict__ colind, const T *__restrict__ vals,
                         XView<T> xv, int64_t m, T *__restrict__ c, CgScal *s, double *partials,
                         unsigned int *ticket, int single_gpu) {
  if (s->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[kStreamThreads / 32];
  CgDotEpi<T> epi{c, xv.x, 0.0};
  spmv_stream_tiles<T, LPR>(rowptr, colind, vals, xv, m, epi, reinterpret_cast<StreamSmem<T> *>(smem_raw));
  const double acc = block_sum<kStreamThreads>(epi.acc, red);
  double total;
  if (grid_reduce_finish<kStreamThreads>(acc, partials, ticket, red, &total) && threadIdx.x == 0)
    cg_finish(FIN_DOT, s, total, nullptr, single_gpu);
}

// K3: x += alpha*u ; r -= alpha*c ; ||r||^2
template <typename T>
__global__ void __launch_bounds__(kThreads) k_cg_update_xr(T *__restrict__ x, T *__restrict__ r,
                                                           const T *__restrict__ u, const T *__restrict__ c,
                                                           int64_t n, CgScal *s, double *hist, double *partials,
                                                           unsigned int *ticket, int pcg, int single_gpu) {
  if (s->done) return;
  __shared__ double smem[kThreads / 32];
  const double alpha_d = pcg ? s->rho / s->dot_uc : (s->residual * s->residual) / s->dot_uc;
  const T alpha = (T)alpha_d;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    T xi, ri;
    if constexpr (sizeof(T) == 8) {
      xi = __dadd_rn(x[i], __dmul_rn(alpha, u[i]));
      ri = __dsub_rn(r[i], __dmul_rn(alpha, c[i]));
    } else {
      xi = __fadd_rn(x[i], __fmul_rn(alpha, u[i]));
      ri = __fsub_rn(r[i], __fmul_rn(alpha, c[i]));
    }
    x[i] = xi;
    r[i] = ri;
    acc += (double)ri * (double)ri;
  }
  acc = block_sum<kThreads>(acc, smem);
  double total;
  if (grid_reduce_finish<kThreads>(acc, partials, ticket, smem, &total) && threadIdx.x == 0)
    cg_finish(pcg ? FIN_NORM_PCG : FIN_NORM, s, total, hist, single_gpu);
}

// PCG: c = r ./ d ; rho = dot(c, r)    (src/cg.jl:79-82, Jacobi ldiv!)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_pcg_precond(const T *__restrict__ d, const T *__restrict__ r,
                                                          T *__restrict__ c, int64_t n, CgScal *s, double *partials,
                                                          unsigned int *ticket, int single_gpu) {
  if (s->done) return;
  __shared__ double smem[kThreads / 32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const T ri = r[i];
    const T ci = ri / d[i];
    c[i] = ci;
    acc += (double)ci * (double)ri;
  }
  acc = block_sum<kThreads>(acc, smem);
  double total;
  if (grid_reduce_finish<kThreads>(acc, partials, ticket, smem, &total) && threadIdx.x == 0)
    cg_finish(FIN_RHO, s, total, nullptr, single_gpu);
}

template <typename T>
struct CgEngine {
  b200_ctx *ctx;
  const b200_csr *A;
  int64_t n;
  T *x, *r, *u, *c;
  const T *b;
  const T *jac;  // NULL => Identity
  CgScal *s;
  double *hist;
  int single;
  int lpr, grid_vec, grid_spmv;

  int after_reduce(int kind) {
    if (single) return B200_OK;
    B200_TRY(allreduce_sum_dev(ctx, &s->sum, 1));
    k_cg_scalar<<<1, 1, 0, ctx->stream>>>(kind, s, hist);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
  }

  int spmv_dot() {
    B200_TRY(halo_exchange(ctx, A, u));
    XView<T> xv = make_xview<T>(A, u);
    if (use_stream(ctx, A)) {
      const int grid = stream_grid_size(ctx, A);
      const size_t smem = sizeof(StreamSmem<T>);
      ProfScope prof(ctx, 0);
#define LAUNCH(L)                                                                                                    \
  do {                                                                                                               \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      B200_CUDA(cudaFuncSetAttribute(k_cg_spmv_dot_stream<T, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                     (int)smem));                                                                    \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    k_cg_spmv_dot_stream<T, L><<<grid, kStreamThreads, smem, ctx->stream>>>(                                        \
        A->rowptr, A->colind, (const T *)A->vals, xv, n, c, s, ctx->red.partials, ctx->red.ticket, single);          \
  } while (0)
      switch (A->stream_lpr) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: LAUNCH(32); break;
      }
#undef LAUNCH
    } else {
    ProfScope prof(ctx, 0);
#define LAUNCH(L)                                                                                               \
  k_cg_spmv_dot<T, L><<<grid_spmv, kThreads, 0, ctx->stream>>>(A->rowptr, A->colind, (const T *)A->vals, xv, n, \
                                                               c, s, ctx->red.partials, ctx->red.ticket, single)
    switch (lpr) {
      case 2: LAUNCH(2); break;
      case 4: LAUNCH(4); break;
      case 8: LAUNCH(8); break;
      case 16: LAUNCH(16); break;
      default: LAUNCH(32); break;
    }
#undef LAUNCH
    }
    B200_LAUNCH_CHECK(ctx);
    return after_reduce(FIN_DOT);
  }

  int iterate() {
    cudaStream_t st = ctx->stream;
    const int pcg = jac != nullptr;
    if (pcg) {
      k_pcg_precond<T><<<grid_vec, kThreads, 0, st>>>(jac, r, c, n, s, ctx->red.partials, ctx->red.ticket, single);
      B200_LAUNCH_CHECK(ctx);
      B200_TRY(after_reduce(FIN_RHO));
    }
    {
      ProfScope prof(ctx, 2);
      k_cg_update_u<T><<<grid_vec, kThreads, 0, st>>>(pcg ? c : r, u, n, s, pcg);
    }
    B200_LAUNCH_CHECK(ctx);
    B200_TRY(spmv_dot());
    {
      ProfScope prof(ctx, 1);
      k_cg_update_xr<T><<<grid_vec, kThreads, 0, st>>>(x, r, u, c, n, s, hist, ctx->red.partials, ctx->red.ticket, pcg, single);
    }
    B200_LAUNCH_CHECK(ctx);
    return after_reduce(pcg ? FIN_NORM_PCG : FIN_NORM);
  }
};

template <typename T>
int cg_solve_impl(b200_ctx *ctx, const b200_csr *A, T *x, const T *b, const b200_cg_opts *o, b200_result *res,
                  double *resnorm_host, int64_t resnorm_cap) {
  cudaStream_t st = ctx->stream;
  const int64_t n = A->m_local;
  const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  const double reltol = o->reltol < 0 ? sqrt(eps) : o->reltol;
  const int64_t maxiter = o->maxiter < 0 ? A->n_global : o->maxiter;
  const int check_every = o->check_every > 0 ? o->check_every : 32;
  const int64_t hist_cap = resnorm_host ? std::min<int64_t>(resnorm_cap, maxiter) : 0;

  // workspace: u, r, c, scalars, history
  const size_t vec_bytes = align_up(sizeof(T) * (size_t)std::max<int64_t>(n, 1), 256);
  const size_t hist_bytes = align_up(sizeof(double) * (size_t)std::max<int64_t>(hist_cap, 1), 256);
  void *ws = nullptr;
  B200_TRY(ws_get(ctx, 3 * vec_bytes + 256 + hist_bytes, &ws));
  char *p = (char *)ws;
  CgEngine<T> e;
  e.ctx = ctx;
  e.A = A;
  e.n = n;
  e.x = x;
  e.b = b;
  e.u = (T *)p; p += vec_bytes;
  e.r = (T *)p; p += vec_bytes;
  e.c = (T *)p; p += vec_bytes;
  e.s = (CgScal *)p; p += 256;
  e.hist = hist_cap ? (double *)p : nullptr;
  e.jac = o->Pl.kind == B200_PREC_JACOBI ? (const T *)o->Pl.diag : nullptr;
  e.single = ctx->world == 1;
  e.lpr = pick_lpr(A->avg_row_nnz);
  e.grid_vec = stream_grid(ctx, n, kThreads * 2, 8);
  e.grid_spmv = stream_grid(ctx, n, kThreads / e.lpr, 8);

  CgScal h;
  memset(&h, 0, sizeof(h));
  h.abstol = o->abstol;
  h.reltol = reltol;
  h.maxiter = maxiter;
  h.hist_cap = hist_cap;
  h.fixed = o->fixed_iterations;
  B200_CUDA(cudaMemcpyAsync(e.s, &h, sizeof(h), cudaMemcpyHostToDevice, st));

  // cg_iterator! (src/cg.jl:120-155)
  int64_t mv_products = 0;
  if (!o->initially_zero) {
    mv_products = 1;
    B200_TRY(spmv(ctx, A, x, e.c));
  }
  k_cg_init<T><<<e.grid_vec, kThreads, 0, st>>>(b, e.c, o->initially_zero ? 0 : 1, e.r, e.u, n, e.s, ctx->red.partials,
                                                 ctx->red.ticket, e.single);
  B200_LAUNCH_CHECK(ctx);
  B200_TRY(e.after_reduce(FIN_INIT));

  // the hot loop (src/cg.jl:229): enqueue check_every iterations, poll the device flag
  int64_t enqueued = 0;
  int *h_done = ctx->h_flags;
  for (;;) {
    B200_CUDA(cudaMemcpyAsync(h_done, &e.s->done, sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (*h_done || enqueued >= maxiter) break;
    const int64_t batch = std::min<int64_t>(check_every, maxiter - enqueued);
    for (int64_t i = 0; i < batch; ++i) B200_TRY(e.iterate());
    enqueued += batch;
  }
  B200_CUDA(cudaMemcpyAsync(&h, e.s, sizeof(h), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (res) {
    res->iters = h.iter;
    res->mvps = mv_products + h.iter;  // history.mvps (src/cg.jl:226-231)
    res->isconverged = h.residual <= h.tol;
    res->status = h.breakdown ? B200_ERR_BREAKDOWN : 0;
    res->tol = h.tol;
    res->residual = h.residual;
    res->n_resnorm = std::min<int64_t>(h.iter, hist_cap);
  }
  if (hist_cap && h.iter > 0) {
    B200_CUDA(cudaMemcpyAsync(resnorm_host, e.hist, sizeof(double) * std::min<int64_t>(h.iter, hist_cap),
                              cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
  }
  return B200_OK;
}

int check_cg_args(b200_ctx *ctx, const b200_csr *A, const void *x, const void *b, const b200_cg_opts *o) {
  B200_REQUIRE(ctx && A && x && b && o, "NULL argument");
  B200_REQUIRE(A->ctx == ctx, "operator belongs to another context");
  B200_REQUIRE(o->Pl.kind == B200_PREC_IDENTITY || (o->Pl.kind == B200_PREC_JACOBI && o->Pl.diag),
               "unsupported preconditioner");
  return B200_OK;
}

}  // namespace

extern "C" {

int b200_cg_solve(b200_ctx *ctx, const b200_csr *A, void *x_dev, const void *b_dev, const b200_cg_opts *opts,
                  b200_result *res, double *resnorm_host, int64_t resnorm_cap) {
  B200_TRY(check_cg_args(ctx, A, x_dev, b_dev, opts));
  B200_CUDA(cudaSetDevice(ctx->device));
  return A->dtype == B200_F64
             ? cg_solve_impl<double>(ctx, A, (double *)x_dev, (const double *)b_dev, opts, res, resnorm_host, resnorm_cap)
             : cg_solve_impl<float>(ctx, A, (float *)x_dev, (const float *)b_dev, opts, res, resnorm_host, resnorm_cap);
}

int b200_cg_solve_host(b200_ctx *ctx, const b200_csr *A, void *x_host, const void *b_host, const b200_cg_opts *opts,
                       b200_result *res, double *resnorm_host, int64_t resnorm_cap) {
  B200_TRY(check_cg_args(ctx, A, x_host, b_host, opts));
  B200_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = dtype_size(A->dtype) * (size_t)A->m_local;
  DevBuf dx, db;
  B200_TRY(dx.alloc(bytes));
  B200_TRY(db.alloc(bytes));
  B200_CUDA(cudaMemcpyAsync(db.p, b_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(cudaMemcpyAsync(dx.p, x_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int s = b200_cg_solve(ctx, A, dx.p, db.p, opts, res, resnorm_host, resnorm_cap);
  if (s != B200_OK) return s;
  B200_CUDA(cudaMemcpyAsync(x_host, dx.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // extern "C"
