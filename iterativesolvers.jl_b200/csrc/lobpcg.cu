// lobpcg.cu -- lobpcg(A, largest, X0; P, tol, maxiter) of reference src/lobpcg.jl:787-839, 865-893 and
// the step functor :692-749, for the standard problem (B = nothing) without constraint -- the path of
// BASELINE.json configs[4] (block = 16, fp32, 3-D Laplacian).
//
// Blocks are n x bs column-major (the reference's layout), device-resident.  Per step (it >= 3):
//   CholQR(R)  : k_gram (R'R, one pass) -> host Cholesky -> k_rdiv (the reference's column sweeps of
//                rdiv! :345-355, done per row in registers)                              (:365-393)
//   AR = A*R   : block SpMM, A streamed ONCE for the 16 columns (the CPU path re-reads A per column) (:124-131)
//   CholQR(P)  : same, AP updated in the same launch                                      (:733)
//   Gram blocks: X'[AR AP R P], R'[AR P], AR'P, P'AP -- 4 launches, each left block is read once for
//                all its right blocks                                                     (:586-605)
//   Rayleigh-Ritz: (3bs x 3bs) generalized symmetric eigenproblem on the host in fp64     (:607-627)
//   update     : ONE launch computes P = R Vr + P Vp, X = X Vx + P, the same for AP/AX, the residual
//                block R = AX - X diag(lambda) and its column norms                        (:629-690, :533-547)
// Soft locking (activeMask, :549-562) gathers the active columns into scratch blocks; with a full mask the
// active blocks alias R/P/AP (no copies).
// The dense contractions are fp32/fp64 FMA on CUDA cores (register-tiled 4x4 per thread): at bs = 16 they
// are HBM-bound (4 flop/byte); routing them through TF32 tensor cores would cost the fp32 parity the
// reference's own tolerance (eps^0.3) does not require but its eigenvalue accuracy does.
#include "blas1.cuh"
#include "dense_small.h"
#include "spmv.cuh"

using namespace b200;

extern "C" int b200_spmm(b200_ctx *ctx, const b200_csr *A, const void *X_dev, int64_t ldx, void *Y_dev, int64_t ldy,
                         int bs);

namespace {

constexpr int kThreads = 256;
constexpr int BSMAX = 16;
constexpr int RC = 64;       // rows per shared-memory chunk in the Gram kernel
constexpr int NRMAX = 4;     // right-hand blocks per Gram launch

// G_b = L' * R_b for b < NR (all n x <=16 column-major, ld = n).  partials: [gridDim][NR*256] doubles.
template <typename T, int NR>
__global__ void __launch_bounds__(kThreads) k_gram(const T *__restrict__ L, int bl, const T *__restrict__ R0,
                                                   const T *__restrict__ R1, const T *__restrict__ R2,
                                                   const T *__restrict__ R3, int br, int64_t n, double *partials,
                                                   unsigned int *ticket, double *__restrict__ out) {
  // staging chunk and the final cross-group reduction share the same shared memory
  constexpr size_t kStageBytes = sizeof(T) * (1 + NR) * BSMAX * RC;
  constexpr size_t kRedBytes = sizeof(double) * 16 * 257;
  __shared__ __align__(16) unsigned char raw[kStageBytes > kRedBytes ? kStageBytes : kRedBytes];
  T(*Ls)[RC] = reinterpret_cast<T(*)[RC]>(raw);
  T(*Rs)[BSMAX][RC] = reinterpret_cast<T(*)[BSMAX][RC]>(raw + sizeof(T) * BSMAX * RC);
  double(*red)[257] = reinterpret_cast<double(*)[257]>(raw);
  __shared__ bool is_last;
  const T *Rp[4] = {R0, R1, R2, R3};
  const int t = threadIdx.x;
  const int g = t >> 4;                 // row group 0..15
  const int ti = (t & 15) >> 2, tj = t & 3;
  double accd[NR][4][4];
#pragma unroll
  for (int b = 0; b < NR; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) accd[b][a][c] = 0.0;

  for (int64_t row0 = (int64_t)blockIdx.x * RC; row0 < n; row0 += (int64_t)gridDim.x * RC) {
    // stage the chunk: thread t loads row (t % RC) of columns t/RC, t/RC+4, ...
    const int rr = t % RC;
    const int64_t grow = row0 + rr;
    for (int j = t / RC; j < BSMAX; j += kThreads / RC) {
      Ls[j][rr] = (j < bl && grow < n) ? L[grow + (int64_t)j * n] : (T)0;
#pragma unroll
      for (int b = 0; b < NR; ++b) Rs[b][j][rr] = (j < br && grow < n) ? Rp[b][grow + (int64_t)j * n] : (T)0;
    }
    __syncthreads();
    T acc[NR][4][4];
#pragma unroll
    for (int b = 0; b < NR; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[b][a][c] = (T)0;
#pragma unroll
    for (int q = 0; q < RC / 16; ++q) {
      const int r = g + 16 * q;
      T l[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) l[a] = Ls[4 * ti + a][r];
#pragma unroll
      for (int b = 0; b < NR; ++b) {
        T rv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rv[c] = Rs[b][4 * tj + c][r];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[b][a][c] += l[a] * rv[c];
      }
    }
#pragma unroll
    for (int b = 0; b < NR; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) accd[b][a][c] += (double)acc[b][a][c];
    __syncthreads();
  }
  // reduce the 16 row groups, one right block at a time
  for (int b = 0; b < NR; ++b) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) red[g][(4 * ti + a) * 16 + 4 * tj + c] = accd[b][a][c];
    __syncthreads();
    double s = 0.0;
    for (int gg = 0; gg < 16; ++gg) s += red[gg][t];
    partials[((size_t)blockIdx.x * NR + b) * 256 + t] = s;
    __syncthreads();
  }
  if (t == 0) {
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int b = 0; b < NR; ++b) {
    double s = 0.0;
    for (unsigned int blk = 0; blk < gridDim.x; ++blk) s += __ldcg(&partials[((size_t)blk * NR + b) * 256 + t]);
    out[b * 256 + t] = s;      // out[b][i*16 + j] = G_b[i][j]
  }
  if (t == 0) *ticket = 0u;
}

// rdiv!(A, U::UpperTriangular) row by row (reference src/lobpcg.jl:345-355), up to 2 blocks per launch
template <typename T>
__global__ void __launch_bounds__(kThreads) k_rdiv(T *__restrict__ X0, T *__restrict__ X1, int nblk, int bs,
                                                   int64_t n, const T *__restrict__ Ufac /* bs x bs col-major */) {
  __shared__ T U[BSMAX][BSMAX];
  for (int q = threadIdx.x; q < BSMAX * BSMAX; q += kThreads) {
    const int i = q % BSMAX, j = q / BSMAX;
    U[i][j] = (i < bs && j < bs) ? Ufac[i + j * bs] : (T)(i == j);
  }
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    for (int blk = 0; blk < nblk; ++blk) {
      T *X = blk == 0 ? X0 : X1;
      T a[BSMAX];
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) a[j] = j < bs ? X[r + (int64_t)j * n] : (T)0;
      a[0] = a[0] / U[0][0];                                         // :347
#pragma unroll
      for (int i = 1; i < BSMAX; ++i) {
        if (i < bs) {
#pragma unroll
          for (int j = 0; j < i; ++j) a[i] = a[i] - a[j] * U[j][i];   // :350
          a[i] = a[i] / U[i][i];                                      // :352
        }
      }
#pragma unroll
      for (int j = 0; j < BSMAX; ++j)
        if (j < bs) X[r + (int64_t)j * n] = a[j];
    }
  }
}

// out[j] (16 per row) = sum_i in[i] * V[i][j]
template <typename T>
__device__ __forceinline__ void row_times_v(const T (&in)[BSMAX], int nin, const T (*V)[BSMAX], T (&out)[BSMAX]) {
#pragma unroll
  for (int i = 0; i < BSMAX; ++i) {
    if (i < nin) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) out[j] += in[i] * V[i][j];
    }
  }
}

struct UpdateArgs {
  void *X, *AX, *P, *AP, *R;          // in/out blocks, sizeX columns
  const void *aR, *aAR, *aP, *aAP;    // active blocks, bs1 / bs2 columns
  int sizeX, bs1, bs2;
  int64_t n;
};

// update_X_P! (:629-690) + residuals! (:533-547) in one pass.  Vx: sizeX x sizeX, Vr: bs1 x sizeX, Vp: bs2 x sizeX
// (row-major [i][j] in the staging buffer, zero padded to 16 x 16); lambda: sizeX.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_update(UpdateArgs a, const T *__restrict__ Vbuf,
                                                     const T *__restrict__ lambda, double *partials,
                                                     unsigned int *ticket, double *__restrict__ norms2) {
  __shared__ T Vx[BSMAX][BSMAX], Vr[BSMAX][BSMAX], Vp[BSMAX][BSMAX];
  __shared__ T lam[BSMAX];
  __shared__ double smem[kThreads / 32][BSMAX];
  __shared__ bool is_last;
  for (int q = threadIdx.x; q < BSMAX * BSMAX; q += kThreads) {
    (&Vx[0][0])[q] = Vbuf[q];
    (&Vr[0][0])[q] = Vbuf[256 + q];
    (&Vp[0][0])[q] = Vbuf[512 + q];
  }
  if (threadIdx.x < BSMAX) lam[threadIdx.x] = threadIdx.x < a.sizeX ? lambda[threadIdx.x] : (T)0;
  __syncthreads();
  T *X = (T *)a.X, *AX = (T *)a.AX, *P = (T *)a.P, *AP = (T *)a.AP, *R = (T *)a.R;
  const T *aR = (const T *)a.aR, *aAR = (const T *)a.aAR, *aP = (const T *)a.aP, *aAP = (const T *)a.aAP;
  const int64_t n = a.n;
  double nrm[BSMAX];
#pragma unroll
  for (int j = 0; j < BSMAX; ++j) nrm[j] = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    T in[BSMAX], pn[BSMAX], xn[BSMAX];
    // ---- block: P = aR Vr + aP Vp ; X = X Vx + P
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) pn[j] = (T)0;
    if (a.bs1 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) in[j] = j < a.bs1 ? aR[r + (int64_t)j * n] : (T)0;
      row_times_v<T>(in, a.bs1, Vr, pn);
    }
    if (a.bs2 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) in[j] = j < a.bs2 ? aP[r + (int64_t)j * n] : (T)0;
      row_times_v<T>(in, a.bs2, Vp, pn);                                     // + aP Vp  (:652-658)
    }
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) in[j] = j < a.sizeX ? X[r + (int64_t)j * n] : (T)0;
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) xn[j] = (T)0;
    row_times_v<T>(in, a.sizeX, Vx, xn);
    if (a.bs1 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) {
        xn[j] = xn[j] + pn[j];                                              // tempX .+ P  :675
        if (j < a.sizeX) P[r + (int64_t)j * n] = pn[j];
      }
    }
#pragma unroll
    for (int j = 0; j < BSMAX; ++j)
      if (j < a.sizeX) X[r + (int64_t)j * n] = xn[j];
    // ---- A block: AP = aAR Vr + aAP Vp ; AX = AX Vx + AP
    T an[BSMAX];
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) pn[j] = (T)0;
    if (a.bs1 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) in[j] = j < a.bs1 ? aAR[r + (int64_t)j * n] : (T)0;
      row_times_v<T>(in, a.bs1, Vr, pn);
    }
    if (a.bs2 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) in[j] = j < a.bs2 ? aAP[r + (int64_t)j * n] : (T)0;
      row_times_v<T>(in, a.bs2, Vp, pn);
    }
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) in[j] = j < a.sizeX ? AX[r + (int64_t)j * n] : (T)0;
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) an[j] = (T)0;
    row_times_v<T>(in, a.sizeX, Vx, an);
    if (a.bs1 > 0) {
#pragma unroll
      for (int j = 0; j < BSMAX; ++j) {
        an[j] = an[j] + pn[j];
        if (j < a.sizeX) AP[r + (int64_t)j * n] = pn[j];
      }
    }
    // ---- residuals!: R = AX - X * Diagonal(lambda)  (:535-536) and column norms (:538-545)
#pragma unroll
    for (int j = 0; j < BSMAX; ++j) {
      if (j < a.sizeX) {
        AX[r + (int64_t)j * n] = an[j];
        const T res = an[j] - xn[j] * lam[j];
        R[r + (int64_t)j * n] = res;
        nrm[j] += (double)res * (double)res;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < BSMAX; ++j) {
    const double v = warp_sum(nrm[j]);
    if (lane == 0) smem[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < BSMAX) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += smem[w][threadIdx.x];
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
  if (threadIdx.x < BSMAX) {
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * kMaxReduceWidth + threadIdx.x]);
    norms2[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// dst[:, k] = src[:, idx[k]]  (update_active! :557-562)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_gather_cols(T *__restrict__ dst, const T *__restrict__ src, int64_t n,
                                                          int bs, const int *__restrict__ idx) {
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads)
    for (int k = 0; k < bs; ++k) dst[r + (int64_t)k * n] = src[r + (int64_t)idx[k] * n];
}

// precond!(R[:,1:bs]) with a Jacobi M: R[:,j] ./= d  (:236-242)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_block_jacobi(T *__restrict__ X, int64_t n, int bs,
                                                           const T *__restrict__ d) {
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    const T di = d[r];
    for (int k = 0; k < bs; ++k) X[r + (int64_t)k * n] = X[r + (int64_t)k * n] / di;
  }
}

template <typename T>
struct Lobpcg {
  b200_ctx *ctx;
  const b200_csr *A;
  int64_t n;
  int sizeX;
  T *X, *AX, *R, *AR, *P, *AP, *gR, *gP, *gAP;   // g*: gather scratch (allocated on first partial mask)
  double *gram_partials, *d_gram;                // device Gram output: NRMAX * 256 doubles
  T *d_small;                                    // V (3*256) + lambda (16) + U (256)
  int *d_idx;
  DevBuf scratch;
  int grid_gram, grid_vec;

  int gram(const T *L, int bl, const T *const *Rb, int nr, int br, double *host_out /* nr*256 */) {
    cudaStream_t st = ctx->stream;
    const T *r0 = Rb[0], *r1 = nr > 1 ? Rb[1] : Rb[0], *r2 = nr > 2 ? Rb[2] : Rb[0], *r3 = nr > 3 ? Rb[3] : Rb[0];
    {
      ProfScope prof(ctx, 1);
      switch (nr) {
        case 1: k_gram<T, 1><<<grid_gram, kThreads, 0, st>>>(L, bl, r0, r1, r2, r3, br, n, gram_partials, ctx->red.ticket, d_gram); break;
        case 2: k_gram<T, 2><<<grid_gram, kThreads, 0, st>>>(L, bl, r0, r1, r2, r3, br, n, gram_partials, ctx->red.ticket, d_gram); break;
        case 3: k_gram<T, 3><<<grid_gram, kThreads, 0, st>>>(L, bl, r0, r1, r2, r3, br, n, gram_partials, ctx->red.ticket, d_gram); break;
        default: k_gram<T, 4><<<grid_gram, kThreads, 0, st>>>(L, bl, r0, r1, r2, r3, br, n, gram_partials, ctx->red.ticket, d_gram); break;
      }
    }
    B200_LAUNCH_CHECK(ctx);
    B200_CUDA(cudaMemcpyAsync(host_out, d_gram, sizeof(double) * 256 * nr, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
  }

  // CholQR (:365-393): blocks[0] is orthonormalised, blocks[1] (A-block) follows if given
  int cholqr(T *blk, T *ablk, int bs) {
    double G[256];
    const T *rb[1] = {blk};
    B200_TRY(gram(blk, bs, rb, 1, bs, G));
    std::vector<double> U((size_t)bs * bs);
    for (int i = 0; i < bs; ++i)
      for (int j = 0; j < bs; ++j) U[i + (size_t)j * bs] = i <= j ? G[i * 16 + j] : G[j * 16 + i];  // Hermitian(gram): upper
    if (dense::cholesky_upper(U.data(), bs, bs)) {
      set_error("PosDefException: CholQR Gram matrix is not positive definite (reference src/lobpcg.jl:380)");
      return B200_ERR_BREAKDOWN;
    }
    std::vector<T> Ut((size_t)bs * bs);
    for (size_t q = 0; q < Ut.size(); ++q) Ut[q] = (T)U[q];
    T *dU = d_small + 3 * 256 + 16;
    B200_CUDA(cudaMemcpyAsync(dU, Ut.data(), sizeof(T) * Ut.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
      ProfScope prof(ctx, 2);
      k_rdiv<T><<<grid_vec, kThreads, 0, ctx->stream>>>(blk, ablk, ablk ? 2 : 1, bs, n, dU);
    }
    B200_LAUNCH_CHECK(ctx);
    B200_CUDA(cudaStreamSynchronize(ctx->stream));   // Ut is a host temporary
    return B200_OK;
  }

  int spmm(const T *Xin, T *Yout, int bs) {
    ProfScope prof(ctx, 0);
    return b200_spmm(ctx, A, Xin, n, Yout, n, bs);
  }
};

template <typename T>
int lobpcg_impl(b200_ctx *ctx, const b200_csr *A, T *X, int64_t ldx, const b200_lobpcg_opts *o,
                b200_lobpcg_result *res, double *lambda_host, double *resnorm_host) {
  cudaStream_t st = ctx->stream;
  const int64_t n = A->m_local;
  const int sizeX = o->blocksize;
  B200_REQUIRE(sizeX >= 1 && sizeX <= BSMAX, "lobpcg: block size %d not in 1..%d", sizeX, BSMAX);
  B200_REQUIRE(ldx == n, "lobpcg: X must be n x blocksize with leading dimension n");
  B200_REQUIRE(sizeX <= n, "X column dimension exceeds the row dimension");                        // :833
  B200_REQUIRE(3 * (int64_t)sizeX <= n, "The LOBPCG algorithms is not stable to use when the matrix size is less "
               "than 3 times the block size. Please use a dense solver instead.");                 // :834
  const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  const double tol = o->tol < 0 ? pow(eps, 0.3) : o->tol;                                           // :751
  const int64_t maxiter = o->maxiter < 0 ? 200 : o->maxiter;
  const T *jac = o->P.kind == B200_PREC_JACOBI ? (const T *)o->P.diag : nullptr;

  Lobpcg<T> L;
  L.ctx = ctx;
  L.A = A;
  L.n = n;
  L.sizeX = sizeX;
  L.X = X;
  L.grid_gram = stream_grid(ctx, n, RC, 4);
  L.grid_vec = stream_grid(ctx, n, kThreads, 4);
  const size_t blk_bytes = align_up(sizeof(T) * (size_t)n * sizeX, 256);
  const size_t small_bytes = align_up(sizeof(T) * (3 * 256 + 16 + 256), 256);
  const size_t gram_bytes = sizeof(double) * ((size_t)L.grid_gram * NRMAX * 256 + NRMAX * 256 + 64);
  void *ws = nullptr;
  B200_TRY(ws_get(ctx, 5 * blk_bytes + small_bytes + gram_bytes + 1024, &ws));
  char *p = (char *)ws;
  L.AX = (T *)p; p += blk_bytes;
  L.R = (T *)p; p += blk_bytes;
  L.AR = (T *)p; p += blk_bytes;
  L.P = (T *)p; p += blk_bytes;
  L.AP = (T *)p; p += blk_bytes;
  L.d_small = (T *)p; p += small_bytes;
  L.gram_partials = (double *)p; p += sizeof(double) * (size_t)L.grid_gram * NRMAX * 256;
  L.d_gram = (double *)p; p += sizeof(double) * NRMAX * 256;
  double *d_norms = (double *)p; p += sizeof(double) * 64;
  L.d_idx = (int *)p;
  L.gR = L.gP = L.gAP = nullptr;

  std::vector<double> ritz(3 * sizeX, 0.0), residuals(sizeX, NAN);                                  // :473-477
  std::vector<char> mask(sizeX, 1);
  int bs = sizeX;
  int64_t iteration = 1;
  int status = B200_OK;

  auto upload_v_and_update = [&](const std::vector<double> &Z, int sub, const std::vector<int> &perm, int bs1,
                                 int bs2, const T *aR, const T *aAR, const T *aP, const T *aAP) -> int {
    // V[1:sub, 1:sizeX] = eigenvectors of the selected Ritz values (:625); split into x / r / p parts
    std::vector<T> Vb(3 * 256 + 16, (T)0);
    for (int j = 0; j < sizeX; ++j) {
      const double *z = &Z[(size_t)perm[j] * sub];
      for (int i = 0; i < sizeX; ++i) Vb[i * 16 + j] = (T)z[i];
      for (int i = 0; i < bs1; ++i) Vb[256 + i * 16 + j] = (T)z[sizeX + i];
      for (int i = 0; i < bs2; ++i) Vb[512 + i * 16 + j] = (T)z[sizeX + bs1 + i];
      Vb[768 + j] = (T)ritz[j];
    }
    B200_CUDA(cudaMemcpyAsync(L.d_small, Vb.data(), sizeof(T) * Vb.size(), cudaMemcpyHostToDevice, st));
    UpdateArgs ua;
    ua.X = L.X; ua.AX = L.AX; ua.P = L.P; ua.AP = L.AP; ua.R = L.R;
    ua.aR = aR; ua.aAR = aAR; ua.aP = aP; ua.aAP = aAP;
    ua.sizeX = sizeX; ua.bs1 = bs1; ua.bs2 = bs2; ua.n = n;
    {
      ProfScope prof(ctx, 1);
      k_update<T><<<L.grid_vec, kThreads, 0, st>>>(ua, L.d_small, L.d_small + 768, ctx->red.partials, ctx->red.ticket, d_norms);
    }
    B200_LAUNCH_CHECK(ctx);
    double nn[BSMAX];
    B200_CUDA(cudaMemcpyAsync(nn, d_norms, sizeof(double) * BSMAX, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < sizeX; ++j) residuals[j] = sqrt(nn[j]);                                     // :545
    return B200_OK;
  };
  auto select = [&](const std::vector<double> &w, int sub, std::vector<int> &perm) {               // :623-624
    perm.resize(sub);
    for (int i = 0; i < sub; ++i) perm[i] = i;
    if (o->largest) std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[a] > w[b]; });
    for (int j = 0; j < sizeX; ++j) ritz[j] = w[perm[j]];
  };

  while (iteration <= maxiter) {                                                                    // :880
    if (iteration == 1) {                                                                           // :695-703
      status = L.cholqr(L.X, nullptr, sizeX);                                                       // ortho_AB_mul_X! :524-532
      if (status) break;
      B200_TRY(L.spmm(L.X, L.AX, sizeX));
      double G[256];
      const T *rb[1] = {L.AX};
      B200_TRY(L.gram(L.X, sizeX, rb, 1, sizeX, G));                                                // XAX :262
      std::vector<double> Am((size_t)sizeX * sizeX), w, Z;
      for (int i = 0; i < sizeX; ++i)
        for (int j = 0; j < sizeX; ++j) Am[i + (size_t)j * sizeX] = i <= j ? G[i * 16 + j] : G[j * 16 + i];
      if (dense::sym_eig(Am, sizeX, w, Z)) { status = B200_ERR_BREAKDOWN; set_error("eigen! did not converge"); break; }
      std::vector<int> perm;
      select(w, sizeX, perm);
      B200_TRY(upload_v_and_update(Z, sizeX, perm, 0, 0, nullptr, nullptr, nullptr, nullptr));      // update_X_P!(0,0)
    } else {
      const bool full = bs == sizeX;
      const bool with_p = iteration > 2;
      T *aR = L.R, *aP = L.P, *aAP = L.AP;
      if (!full) {                                                                                  // update_active! :557-562
        if (!L.gR) {
          B200_TRY(L.scratch.alloc(3 * blk_bytes));
          L.gR = (T *)L.scratch.p;
          L.gP = (T *)((char *)L.scratch.p + blk_bytes);
          L.gAP = (T *)((char *)L.scratch.p + 2 * blk_bytes);
        }
        int idx[BSMAX], k = 0;
        for (int j = 0; j < sizeX; ++j)
          if (mask[j]) idx[k++] = j;
        B200_CUDA(cudaMemcpyAsync(L.d_idx, idx, sizeof(int) * bs, cudaMemcpyHostToDevice, st));
        k_gather_cols<T><<<L.grid_vec, kThreads, 0, st>>>(L.gR, L.R, n, bs, L.d_idx);
        if (with_p) {
          k_gather_cols<T><<<L.grid_vec, kThreads, 0, st>>>(L.gP, L.P, n, bs, L.d_idx);
          k_gather_cols<T><<<L.grid_vec, kThreads, 0, st>>>(L.gAP, L.AP, n, bs, L.d_idx);
        }
        B200_LAUNCH_CHECK(ctx);
        B200_CUDA(cudaStreamSynchronize(st));
        aR = L.gR; aP = L.gP; aAP = L.gAP;
      }
      if (jac) {                                                                                    // precond_constr! :564-569
        k_block_jacobi<T><<<L.grid_vec, kThreads, 0, st>>>(aR, n, bs, jac);
        B200_LAUNCH_CHECK(ctx);
      }
      status = L.cholqr(aR, nullptr, bs);                                                           // :524-532
      if (status) break;
      B200_TRY(L.spmm(aR, L.AR, bs));
      if (with_p) {
        status = L.cholqr(aP, aAP, bs);                                                             // :733
        if (status) break;
      }
      const int n1 = sizeX, n2 = bs, n3 = with_p ? bs : 0, sub = n1 + n2 + n3;
      std::vector<double> gA((size_t)sub * sub, 0.0), gB((size_t)sub * sub, 0.0);
      auto setA = [&](int i, int j, double v) { gA[i + (size_t)j * sub] = v; gA[j + (size_t)i * sub] = v; };
      auto setB = [&](int i, int j, double v) { gB[i + (size_t)j * sub] = v; gB[j + (size_t)i * sub] = v; };
      for (int i = 0; i < n1; ++i) setA(i, i, ritz[i]);                                             // Diagonal(lambda) :289
      for (int i = 0; i < sub; ++i) setB(i, i, 1.0);                                                // I! :315,322,331
      double G[NRMAX * 256];
      {   // X' [AR, R, AP, P]
        const T *rb[4] = {L.AR, aR, aAP, aP};
        B200_TRY(L.gram(L.X, n1, rb, with_p ? 4 : 2, n2, G));
        for (int i = 0; i < n1; ++i)
          for (int j = 0; j < n2; ++j) {
            setA(i, n1 + j, G[i * 16 + j]);                                                         // XAR :265
            setB(i, n1 + j, G[256 + i * 16 + j]);                                                   // XBR :270
            if (with_p) {
              setA(i, n1 + n2 + j, G[512 + i * 16 + j]);                                            // XAP :264
              setB(i, n1 + n2 + j, G[768 + i * 16 + j]);                                            // XBP :269
            }
          }
      }
      {   // R' [AR, P]
        const T *rb[2] = {L.AR, aP};
        B200_TRY(L.gram(aR, n2, rb, with_p ? 2 : 1, n2, G));
        for (int i = 0; i < n2; ++i)
          for (int j = i; j < n2; ++j) setA(n1 + i, n1 + j, G[i * 16 + j]);                         // RAR :266 (upper triangle)
        if (with_p)
          for (int i = 0; i < n2; ++i)
            for (int j = 0; j < n3; ++j) setB(n1 + i, n1 + n2 + j, G[256 + i * 16 + j]);            // RBP :271
      }
      if (with_p) {
        const T *rb1[1] = {aP};
        B200_TRY(L.gram(L.AR, n2, rb1, 1, n3, G));                                                  // RAP = AR' P :267
        for (int i = 0; i < n2; ++i)
          for (int j = 0; j < n3; ++j) setA(n1 + i, n1 + n2 + j, G[i * 16 + j]);
        const T *rb2[1] = {aAP};
        B200_TRY(L.gram(aP, n3, rb2, 1, n3, G));                                                    // PAP :268
        for (int i = 0; i < n3; ++i)
          for (int j = i; j < n3; ++j) setA(n1 + n2 + i, n1 + n2 + j, G[i * 16 + j]);
      }
      std::vector<double> w, Z;
      const int info = dense::sym_eig_generalized(gA, gB, sub, w, Z);                               // :622
      if (info) {
        status = B200_ERR_BREAKDOWN;
        set_error("PosDefException in the Rayleigh-Ritz problem (gramB not positive definite, info=%d)", info);
        break;
      }
      std::vector<int> perm;
      select(w, sub, perm);
      B200_TRY(upload_v_and_update(Z, sub, perm, n2, n3, aR, L.AR, aP, aAP));
    }
    bs = 0;                                                                                         // update_mask! :549-555
    for (int j = 0; j < sizeX; ++j) {
      mask[j] = o->fixed_iterations ? 1 : (residuals[j] > tol);
      bs += mask[j];
    }
    if (bs == 0) break;                                                                             // :885
    iteration += 1;                                                                                 // :886
  }
  bool conv = true;
  for (int j = 0; j < sizeX; ++j) {
    if (lambda_host) lambda_host[j] = ritz[j];
    if (resnorm_host) resnorm_host[j] = residuals[j];
    conv = conv && (residuals[j] <= tol);
  }
  if (res) {
    res->iterations = iteration;                                                                    // :890
    res->converged = conv;
    res->status = status;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  return status;
}

}  // namespace

extern "C" {

int b200_lobpcg_solve(b200_ctx *ctx, const b200_csr *A, void *X_dev, int64_t ldx, const b200_lobpcg_opts *opts,
                      b200_lobpcg_result *res, double *lambda_host, double *resnorm_host) {
  B200_REQUIRE(ctx && A && X_dev && opts, "NULL argument");
  B200_REQUIRE(A->ctx == ctx, "operator belongs to another context");
  B200_REQUIRE(ctx->world == 1, "lobpcg is single-GPU in this version");
  B200_CUDA(cudaSetDevice(ctx->device));
  return A->dtype == B200_F64 ? lobpcg_impl<double>(ctx, A, (double *)X_dev, ldx, opts, res, lambda_host, resnorm_host)
                              : lobpcg_impl<float>(ctx, A, (float *)X_dev, ldx, opts, res, lambda_host, resnorm_host);
}

int b200_dense_sygv_host(int n, const double *A, const double *B, double *w, double *Z) {
  B200_REQUIRE(n >= 1 && n <= 64 && A && w && Z, "bad arguments");
  std::vector<double> Am(A, A + (size_t)n * n), wv, Zv;
  int info;
  if (B) {
    std::vector<double> Bm(B, B + (size_t)n * n);
    info = dense::sym_eig_generalized(Am, Bm, n, wv, Zv);
  } else {
    info = dense::sym_eig(Am, n, wv, Zv);
  }
  if (info) {
    set_error("dense symmetric eigen-solver failed (info=%d)", info);
    return B200_ERR_BREAKDOWN;
  }
  memcpy(w, wv.data(), sizeof(double) * n);
  memcpy(Z, Zv.data(), sizeof(double) * (size_t)n * n);
  return B200_OK;
}

}  // extern "C"
