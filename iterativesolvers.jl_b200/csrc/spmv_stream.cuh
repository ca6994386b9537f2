// spmv_stream.cuh -- the TMA-streamed CSR SpMV body shared by b200_spmv and the fused solver kernels.
//
// Why: with ~7 nonzeros per row a "warp(-slice) per row" kernel is latency-bound: every row costs a
// dependent chain rowptr -> (colind, vals) -> x[col] with one row in flight per lane group
// (measured: 2.3 TB/s = 35 % of the B200's HBM peak).  Here the matrix is consumed as what it is in
// HBM -- three contiguous streams -- and only the x gather stays a per-thread load:
//
//   * rows are cut into uniform tiles of R = 512/LPR rows; the nonzeros of a tile are one contiguous
//     range of vals/colind, the row pointers one contiguous range of rowptr;
//   * a producer warp (one elected lane) moves the three ranges of the next tiles into a 4-stage
//     shared-memory ring with cp.async.bulk (the TMA engine; SASS: UBLKCP); completion is tracked by
//     mbarriers (full[stage]: expect_tx bytes; empty[stage]: one arrive per consumer warp of the group
//     that owns the tile).  The copies carry an L2 evict-first policy so the 12 B/nnz stream does not
//     evict x from L2;
//   * consumers read rowptr/colind/vals from shared memory (no dependent global loads) and issue all x
//     gathers of their rows back to back; x comes from L1/L2;
//   * the x gather of a row ends in a DRAM miss roughly once per row (first touch of the leading stencil
//     plane), so a tile's consumer latency is a loaded DRAM round trip (~2500 clk): the first streamed
//     version (one tile in consumer flight per CTA, 512 rows per SM) was bound by exactly that --
//     ncu: DRAM 65 %, L2 44 %, issue 40 %, "long scoreboard" the top stall.  Therefore the 16 consumer
//     warps form TWO groups that work on alternate tiles concurrently and every thread owns TWO rows:
//     1024 rows (7168 gathers) in flight per SM while two more tiles are landing;
//   * the grid is persistent: 1 CTA per SM, tiles interleaved across CTAs (k-th tile of CTA b is
//     b + k*gridDim) so that all resident CTAs sweep neighbouring rows and the x planes stay in L2.
//
// LPR (lanes per row) = 1 for matrices whose tiles of 512 rows hold <= 4096 nonzeros (the 5/7-point
// stencils), 2/4/../32 for denser rows; operators whose 16-row tiles exceed 4096 nonzeros fall back to
// the sub-warp kernel (spmv.cuh).  With LPR == 1 and fp64 the row sum is accumulated left to right
// with separate multiply and add, i.e. bit-identical to SparseArrays' CSC scatter for a matrix given
// with sorted columns.
#pragma once
#include "spmv.cuh"

namespace b200 {

constexpr int kStreamGroupThreads = 256;                  // consumer threads per group (8 warps)
constexpr int kStreamGroups = 2;                          // groups working on alternate tiles
constexpr int kStreamConsumers = kStreamGroupThreads * kStreamGroups;
constexpr int kStreamThreads = kStreamConsumers + 32;     // + producer warp
constexpr int kStreamTileRows = 512;                      // rows per tile at LPR == 1 (2 per thread)
constexpr int kStreamNnzCap = 4096;                       // nonzeros per tile
constexpr int kStreamStages = 4;
constexpr int kStreamCtasPerSm = 1;

template <typename T>
struct alignas(128) StreamStage {
  T val[kStreamNnzCap + 8];
  int col[kStreamNnzCap + 8];
  int rp[kStreamTileRows + 8];
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

// Runs over all tiles of this CTA.  `epi(row, value)` is called once per row by the lane that owns
// the row result.  Must be called by all kStreamThreads threads of the block.
template <typename T, int LPR, typename XV, typename Epi>
__device__ __forceinline__ void spmv_stream_tiles(const int *__restrict__ rowptr, const int *__restrict__ colind,
                                                  const T *__restrict__ vals, const XV &xv, int64_t m, Epi &epi,
                                                  StreamSmem<T> *sm, bool rev = false) {
  constexpr int R = kStreamTileRows / LPR;          // rows per tile
  constexpr int SLOTS = kStreamGroupThreads / LPR;  // row slots per group; each slot owns rows s and s+SLOTS
  const int tid = threadIdx.x;
  const int64_t ntiles = (m + R - 1) / R;
  // `rev`: sweep the tiles from the last to the first.  Consecutive kernels of a solver alternate the sweep
  // direction so that each one starts on the rows the previous kernel touched last (still in the 126 MB L2).
  auto phys = [&](int64_t seq) -> int64_t { return rev ? ntiles - 1 - seq : seq; };
  if (tid == 0) {
    for (int s = 0; s < kStreamStages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kStreamGroupThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kStreamConsumers) {
    // ------------------------------------------------------------ producer warp
    if (tid == kStreamConsumers) {
      const uint64_t pol_stream = policy_evict_first();
      int64_t t = blockIdx.x;
      // bounds of the next tile are fetched one iteration ahead (off the critical path)
      int k0 = 0, k1 = 0;
      if (t < ntiles) {
        const int64_t r0 = phys(t) * R, r1 = (r0 + R < m) ? (r0 + R) : m;
        k0 = __ldg(rowptr + r0);
        k1 = __ldg(rowptr + r1);
      }
      for (int it = 0; t < ntiles; ++it) {
        const int s = it % kStreamStages;
        const uint32_t ph = (uint32_t)((it / kStreamStages) & 1);
        const int64_t r0 = phys(t) * R;
        const int64_t tn = t + gridDim.x;
        int nk0 = 0, nk1 = 0;
        if (tn < ntiles) {
          const int64_t nr0 = phys(tn) * R, nr1 = (nr0 + R < m) ? (nr0 + R) : m;
          nk0 = __ldg(rowptr + nr0);
          nk1 = __ldg(rowptr + nr1);
        }
        mbar_wait(&sm->empty[s], ph ^ 1u);
        const int k0a = k0 & ~3;
        const uint32_t cnt = (uint32_t)(((k1 - k0a) + 3) & ~3);
        const uint32_t b_val = cnt * (uint32_t)sizeof(T), b_col = cnt * 4u, b_rp = (uint32_t)(R + 4) * 4u;
        StreamStage<T> *st = &sm->stage[s];
        mbar_expect_tx(&sm->full[s], b_val + b_col + b_rp);
        bulk_g2s(st->rp, rowptr + r0, b_rp, &sm->full[s], pol_stream);
        bulk_g2s(st->col, colind + k0a, b_col, &sm->full[s], pol_stream);
        bulk_g2s(st->val, vals + k0a, b_val, &sm->full[s], pol_stream);
        t = tn;
        k0 = nk0;
        k1 = nk1;
      }
    }
  } else {
    // ------------------------------------------------------------ consumers: group g takes tiles k = g, g+2, ...
    const int grp = tid / kStreamGroupThreads;
    const int lt = tid % kStreamGroupThreads;
    const int sub = lt % LPR, slot = lt / LPR;
    for (int64_t k = grp;; k += kStreamGroups) {
      const int64_t t = (int64_t)blockIdx.x + k * gridDim.x;
      if (t >= ntiles) break;
      const int s = (int)(k % kStreamStages);
      const uint32_t ph = (uint32_t)((k / kStreamStages) & 1);
      const int64_t r0 = phys(t) * R;
      mbar_wait(&sm->full[s], ph);
      const StreamStage<T> *st = &sm->stage[s];
      const int k0a = st->rp[0] & ~3;
      int b[2], e[2];
      bool valid[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int rib = slot + q * SLOTS;
        valid[q] = r0 + rib < m;
        b[q] = valid[q] ? st->rp[rib] - k0a : 0;
        e[q] = valid[q] ? st->rp[rib + 1] - k0a : 0;
      }
      T acc[2] = {(T)0, (T)0};
      // the epilogue's own per-row operand (e.g. MINRES' v_prev[row]) is requested BEFORE the gathers so that its
      // DRAM latency overlaps theirs; loaded after them it doubled the time a tile occupies its stage
      T pre[2] = {(T)0, (T)0};
      if (sub == 0) {
        if (valid[0]) pre[0] = epi.pre(r0 + slot);
        if (valid[1]) pre[1] = epi.pre(r0 + slot + SLOTS);
      }
      if constexpr (LPR == 1) {
        // 2 rows x up to 8 gathers in flight; left-to-right, unfused multiply-add (see header comment)
        int kk0 = b[0], kk1 = b[1];
        while (kk0 < e[0] || kk1 < e[1]) {
          T xa[2][8];   // the gathers; vals are re-read from shared memory at multiply time (registers)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool on0 = kk0 + j < e[0], on1 = kk1 + j < e[1];
            const int c0 = on0 ? st->col[kk0 + j] : 0;
            const int c1 = on1 ? st->col[kk1 + j] : 0;
            xa[0][j] = on0 ? xv(c0) : (T)0;
            xa[1][j] = on1 ? xv(c1) : (T)0;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (kk0 + j < e[0]) {
              if constexpr (sizeof(T) == 8) acc[0] = __dadd_rn(acc[0], __dmul_rn(st->val[kk0 + j], xa[0][j]));
              else acc[0] = __fadd_rn(acc[0], __fmul_rn(st->val[kk0 + j], xa[0][j]));
            }
            if (kk1 + j < e[1]) {
              if constexpr (sizeof(T) == 8) acc[1] = __dadd_rn(acc[1], __dmul_rn(st->val[kk1 + j], xa[1][j]));
              else acc[1] = __fadd_rn(acc[1], __fmul_rn(st->val[kk1 + j], xa[1][j]));
            }
          }
          kk0 += 8;
          kk1 += 8;
        }
      } else {
        int kk0 = b[0] + sub, kk1 = b[1] + sub;
        while (kk0 < e[0] || kk1 < e[1]) {
          T xa[2][4], va[2][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool on0 = kk0 + j * LPR < e[0], on1 = kk1 + j * LPR < e[1];
            const int c0 = on0 ? st->col[kk0 + j * LPR] : 0;
            const int c1 = on1 ? st->col[kk1 + j * LPR] : 0;
            va[0][j] = on0 ? st->val[kk0 + j * LPR] : (T)0;
            va[1][j] = on1 ? st->val[kk1 + j * LPR] : (T)0;
            xa[0][j] = on0 ? xv(c0) : (T)0;
            xa[1][j] = on1 ? xv(c1) : (T)0;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0] += va[0][j] * xa[0][j];
            acc[1] += va[1][j] * xa[1][j];
          }
          kk0 += 4 * LPR;
          kk1 += 4 * LPR;
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o, LPR);
          acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o, LPR);
        }
      }
      if (sub == 0) {
        if (valid[0]) epi(r0 + slot, acc[0], pre[0]);
        if (valid[1]) epi(r0 + slot + SLOTS, acc[1], pre[1]);
      }
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
  const int R = kStreamTileRows / A->stream_lpr;
  const int64_t ntiles = (A->m_local + R - 1) / R;
  const int64_t cap = (int64_t)ctx->sm_count * kStreamCtasPerSm;
  return (int)(ntiles < cap ? ntiles : cap);
}

}  // namespace b200
