// spmv_stream.cuh -- the TMA-streamed CSR SpMV body shared by b200_spmv and the fused solver kernels.
//
// Why: with ~7 nonzeros per row a "warp(-slice) per row" kernel is latency-bound: every row costs a
// dependent chain rowptr -> (colind, vals) -> x[col] with one row in flight per lane group
// (measured: 2.3 TB/s = 35 % of the B200's HBM peak).  Here the matrix is consumed as what it is in
// HBM -- three contiguous streams -- and the x gather is served from shared memory where possible:
//
//   * rows are cut into uniform tiles of R = 512/LPR rows; the nonzeros of a tile are one contiguous
//     range of vals/colind, the row pointers one contiguous range of rowptr;
//   * a producer warp (one elected lane) moves the three ranges of tile t+STAGES-1 into a shared-memory
//     ring with cp.async.bulk (the TMA engine; SASS: UBLKCP) while the 16 consumer warps work on tile t;
//     completion is tracked by mbarriers (full[stage]: expect_tx bytes; empty[stage]: one arrive per
//     consumer warp).  The matrix copies carry an L2 evict-first policy so the 12 B/nnz stream does not
//     evict x from L2;
//   * the same stage also receives the x WINDOW [r0-W, r1+W) of the tile by TMA: every column within W
//     of the diagonal (for a 7-point stencil with N <= 512: the +-1 and +-N neighbours, 5 of 7 gathers)
//     is then a shared-memory read; columns outside the window (the +-N^2 planes, halo columns) are
//     global loads that hit L1/L2.  W is picked per operator from the band profile of the matrix
//     (first ncu capture without the window: L2->SM traffic 2.2 GB per SpMV for x alone, L2-bound);
//   * consumers read rowptr/colind/vals from shared memory (no dependent global loads) and issue all x
//     gathers of a row back to back (8 in flight per thread);
//   * the grid is persistent: 1 CTA per SM, tiles interleaved across CTAs (t = blockIdx, +gridDim, ..)
//     so that all resident CTAs sweep neighbouring rows and the far x planes stay in L2.
//
// LPR (lanes per row) = 1 for matrices whose tiles of 512 rows hold <= 4096 nonzeros (the 5/7-point
// stencils), 2/4/../32 for denser rows; operators whose 16-row tiles exceed 4096 nonzeros fall back to
// the sub-warp kernel (spmv.cuh).  With LPR == 1 and fp64 the row sum is accumulated left to right
// with separate multiply and add, i.e. bit-identical to SparseArrays' CSC scatter for a matrix given
// with sorted columns.
#pragma once
#include "spmv.cuh"

namespace b200 {

constexpr int kStreamConsumers = 512;                     // consumer threads (16 warps)
constexpr int kStreamThreads = kStreamConsumers + 32;     // + producer warp
constexpr int kStreamNnzCap = 4096;                       // nonzeros per tile
constexpr int kStreamXwCap = 1536;                        // elements of x staged per tile (R + 2W <= cap)
constexpr int kStreamStages = 3;
constexpr int kStreamCtasPerSm = 1;

template <typename T>
struct alignas(128) StreamStage {
  T val[kStreamNnzCap + 8];
  T xw[kStreamXwCap];
  int col[kStreamNnzCap + 8];
  int rp[kStreamConsumers + 8];
};
template <typename T>
struct StreamSmem {
  StreamStage<T> stage[kStreamStages];
  alignas(8) unsigned long long full[kStreamStages];
  alignas(8) unsigned long long empty[kStreamStages];
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier, L2 cache policy
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         unsigned long long *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// x window of tile [r0, r1): elements [lo, lo+cnt) of the own slab; cnt is a multiple of 4 (16 B)
__device__ __forceinline__ void stream_window(int64_t r0, int64_t r1, int64_t m, int W, int &lo, int &cnt) {
  if (W < 0) {
    lo = 0;
    cnt = 0;
    return;
  }
  const int64_t a = r0 - W > 0 ? r0 - W : 0;
  const int64_t b = r1 + W < m ? r1 + W : m;
  lo = (int)a;
  cnt = (int)((b - a) & ~(int64_t)3);
}

// Runs over all tiles of this CTA.  `epi(row, value)` is called once per row by the lane that owns
// the row result.  Must be called by all kStreamThreads threads of the block.
// W: half-width of the x window (multiple of 4, R + 2W <= kStreamXwCap), or -1 for "no window".
template <typename T, int LPR, typename XV, typename Epi>
__device__ __forceinline__ void spmv_stream_tiles(const int *__restrict__ rowptr, const int *__restrict__ colind,
                                                  const T *__restrict__ vals, const XV &xv, int64_t m, int W,
                                                  Epi &epi, StreamSmem<T> *sm) {
  constexpr int R = kStreamConsumers / LPR;     // rows per tile
  const int tid = threadIdx.x;
  const int64_t ntiles = (m + R - 1) / R;
  if (tid == 0) {
    for (int s = 0; s < kStreamStages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kStreamConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kStreamConsumers) {
    // ------------------------------------------------------------ producer warp
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % kStreamStages;
      const uint32_t ph = (uint32_t)((it / kStreamStages) & 1);
      if (tid == kStreamConsumers) {
        mbar_wait(&sm->empty[s], ph ^ 1u);
        const int64_t r0 = t * R;
        const int64_t r1 = (r0 + R < m) ? (r0 + R) : m;
        const int k0 = __ldg(rowptr + r0), k1 = __ldg(rowptr + r1);
        const int k0a = k0 & ~3;
        const uint32_t cnt = (uint32_t)(((k1 - k0a) + 3) & ~3);
        const uint32_t b_val = cnt * (uint32_t)sizeof(T), b_col = cnt * 4u, b_rp = (uint32_t)(R + 4) * 4u;
        int wlo, wcnt;
        stream_window(r0, r1, m, W, wlo, wcnt);
        const uint32_t b_xw = (uint32_t)wcnt * (uint32_t)sizeof(T);
        StreamStage<T> *st = &sm->stage[s];
        mbar_expect_tx(&sm->full[s], b_val + b_col + b_rp + b_xw);
        bulk_g2s(st->rp, rowptr + r0, b_rp, &sm->full[s], pol_stream);
        bulk_g2s(st->col, colind + k0a, b_col, &sm->full[s], pol_stream);
        if (b_xw) bulk_g2s(st->xw, xv.x + wlo, b_xw, &sm->full[s], pol_keep);
        bulk_g2s(st->val, vals + k0a, b_val, &sm->full[s], pol_stream);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ consumers
    const int sub = tid % LPR, rib = tid / LPR;
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % kStreamStages;
      const uint32_t ph = (uint32_t)((it / kStreamStages) & 1);
      const int64_t r0 = t * R;
      const int64_t r1 = (r0 + R < m) ? (r0 + R) : m;
      int wlo, wcnt;
      stream_window(r0, r1, m, W, wlo, wcnt);
      mbar_wait(&sm->full[s], ph);
      const StreamStage<T> *st = &sm->stage[s];
      const int64_t row = r0 + rib;
      const bool valid = row < m;
      const int k0a = st->rp[0] & ~3;
      int b = 0, e = 0;
      if (valid) {
        b = st->rp[rib] - k0a;
        e = st->rp[rib + 1] - k0a;
      }
      // x[c]: shared-memory window first, L1/L2 otherwise
      auto xget = [&](int c) -> T {
        const unsigned d = (unsigned)(c - wlo);
        return d < (unsigned)wcnt ? st->xw[d] : xv(c);
      };
      T acc = (T)0;
      if constexpr (LPR == 1) {
        // up to 8 gathers in flight; left-to-right, unfused multiply-add (see header comment)
        for (int k = b; k < e; k += 8) {
          T xa[8], va[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool on = k + j < e;
            const int c = on ? st->col[k + j] : wlo;
            va[j] = on ? st->val[k + j] : (T)0;
            xa[j] = on ? xget(c) : (T)0;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (k + j < e) {
              if constexpr (sizeof(T) == 8) acc = __dadd_rn(acc, __dmul_rn(va[j], xa[j]));
              else acc = __fadd_rn(acc, __fmul_rn(va[j], xa[j]));
            }
          }
        }
      } else {
        for (int k = b + sub; k < e; k += 4 * LPR) {
          T xa[4], va[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = k + j * LPR;
            const bool on = kk < e;
            const int c = on ? st->col[kk] : wlo;
            va[j] = on ? st->val[kk] : (T)0;
            xa[j] = on ? xget(c) : (T)0;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) acc += va[j] * xa[j];
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
      }
      if (valid && sub == 0) epi(row, acc);
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&sm->empty[s]);
    }
  }
}

#endif  // __CUDACC__

// true if the TMA-streamed kernel serves this operator (tiles fit, not overridden by the option)
inline bool use_stream(const b200_ctx *ctx, const b200_csr *A) {
  return A->stream_lpr > 0 && ctx->opt_spmv_kernel != 1;
}
inline int stream_grid_size(const b200_ctx *ctx, const b200_csr *A) {
  const int R = kStreamConsumers / A->stream_lpr;
  const int64_t ntiles = (A->m_local + R - 1) / R;
  const int64_t cap = (int64_t)ctx->sm_count * kStreamCtasPerSm;
  return (int)(ntiles < cap ? ntiles : cap);
}
// half-width of the x window for a launch on vector x: the operator's choice, or -1 when the window
// is disabled (option, or x not 16-byte aligned as the bulk copy requires)
inline int stream_window_w(const b200_ctx *ctx, const b200_csr *A, const void *x) {
  if (ctx->opt_stream_window == 0 || A->stream_w <= 0) return -1;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return -1;
  return A->stream_w;
}

}  // namespace b200
