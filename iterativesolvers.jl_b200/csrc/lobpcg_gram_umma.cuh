// lobpcg_gram_umma.cuh -- the eight Gram blocks of one LOBPCG Rayleigh-Ritz step (reference src/lobpcg.jl:586-605
// block_grams_3x3!, computed there as eight mul! calls) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with
// the accumulator in TMEM, fed from shared memory, hand-written for sm_100a.  fp32 blocks only (config #5).
//
//   blocks (row-major n x 16 fp32, 64 B per row):  X, R, AR, P, AP
//   products: X'AR, X'R, R'AR, X'AP, X'P, R'P, AR'P, P'AP      (same order / output format as k_gram_rr_tc<2>)
//
// One MMA per 8 rows.  Every fp32 value a is split a = a_hi + a_lo (both TF32, "3xTF32"); the 8-row slice of the five
// blocks becomes ONE shared-memory operand of 160 columns  [X_h X_l R_h R_l AR_h AR_l P_h P_l AP_h AP_l]  stored
// K-major (K = the 8 rows; measured with tools/umma_probe.cu: kind::tf32 returns zeros for MN-major operands without
// swizzle, so the converter warps transpose 4 x 4 register tiles), of which the MMA reads columns 0..127 as A (left
// factors X, R, AR, P) and columns 32..159 as B (right factors R, AR, P, AP) -- the same bytes serve both operands:
//
//        D (128 x 128, TMEM, fp32)  +=  A^T (128 x 8)  *  B (8 x 128)
//
// so the 32 x 32 tile (left block, right block) of D holds the hi*hi, hi*lo, lo*hi and lo*lo partial Grams of that
// pair side by side; the epilogue adds the four.  Cost per 8 rows: one M128 N128 K8 MMA = 64 tensor cycles
// (B300_MICROARCH "tcgen05 floor"), against 108 cycles of HBM time for the slice (2.5 KB at 23 B/clk/SM): the
// pass is memory-bound instead of pinned at the legacy mma.sync pipe's ceiling (r1: 0.50 of the HBM peak).
//
// Warp roles (512 threads, one CTA per SM, persistent over 64-row stages):
//   warp 4      producer: cp.async.bulk of the five 64-row block slices into a 4-deep raw ring (mbarrier tx-count)
//   warps 6-15  converters: a thread takes 4 rows x 4 columns of one block (4 128-bit loads), splits into hi / lo and
//               stores the transposed tile in the UMMA canonical K-major layout without swizzle (a core matrix is 8
//               columns x 16 B = 4 consecutive rows; 8-column groups 160 B apart, the second K half 3216 B behind the
//               first: with these strides the 128-bit loads and stores are bank-conflict free), fence.proxy.async,
//               arrive on op_full
//   warp 5      one thread issues the tcgen05.mma's (8 per stage); tcgen05.commit releases the operand stage and,
//               every kUmDrain stages, hands the accumulator to the epilogue; two accumulators (2 x 128 TMEM columns)
//   warps 0-3   epilogue: tcgen05.ld of the needed column blocks (warp w owns TMEM lanes 32w..32w+31 = left block w),
//               hi/lo column halves summed into 64 fp32 registers per thread; at the end lanes l and l+16 (hi and lo
//               ROW of the same left column) are combined, per-CTA partials go out in fp64 and the last CTA (ticket)
//               sums them.
//
// Measured alternatives (profiles/r2_summary.md): ncu on this version: DRAM traffic = algorithmic (5.37 GB), shared-memory
// pipe ~50 % LSU wavefronts (conflict-free) + 40 % tensor-core operand reads, tensor pipe 12 %; 1.39 ms per launch = 0.59 of
// the HBM peak (legacy mma.sync kernel: 1.62 ms).  Converters reading the blocks straight from global memory (no bulk-copy
// staging, 4 operand stages) were SLOWER (1.47 ms): the staged version stays.
//
// Accumulation length.  The tensor core adds each K = 8 product sum to the fp32 accumulator with truncation: a chain of c
// MMAs biases a positive sum (the diagonals of R'AR, P'AP) by about -c/2 ulp.  Measured (tools/diag_gram_ortho.py) on
// orthonormal blocks: -2.3e-7 relative with 8 MMAs per hand-over, -2.5e-6 with 64 -- and the legacy mma.sync kernel, whose
// per-warp chains grow with n, -9e-8 at n = 24^3 but -1.9e-6 at 64^3 (at 256^3 it lets lobpcg drift to NEGATIVE Ritz
// values of an SPD matrix within 30 steps).  With the hand-over after every stage (kUmDrain = 1: 8 MMAs, then unbiased fp32
// adds in the epilogue, fp64 across CTAs) the bias is independent of n and `lobpcg` converges where the reference's
// arithmetic does; 64-MMA chains made the natural run of test_lobpcg_fp32_config5_shape end in a PosDefException.
#pragma once
#include "common.cuh"
#include "spmv_stream.cuh"

namespace b200 {

constexpr int kUmRows = 64;                               // rows per stage = 8 MMAs
constexpr int kUmSteps = kUmRows / 8;
constexpr int kUmRawStages = 4;
constexpr int kUmOpStages = 2;
constexpr int kUmGroupStride = 160;                       // SBO: bytes between consecutive 8-column groups (128 B of data)
constexpr int kUmGroups = 20;                             // 5 blocks x (hi, lo) x 2 groups of 8 columns
constexpr int kUmHalfStride = kUmGroups * kUmGroupStride + 16;   // LBO: rows 4..7 of a step start 3216 B behind rows 0..3
constexpr int kUmStepBytes = 2 * kUmHalfStride;           // 6432 B: the operand of one MMA (K = 8 rows)
constexpr int kUmOpBytes = kUmSteps * kUmStepBytes;       // 51456 B per operand stage
constexpr int kUmDrain = 1;                               // stages per accumulator hand-over (64 rows, 8 MMAs): see the note on truncation below
constexpr int kUmEpiWarps = 4, kUmConvWarps = 10;
constexpr int kUmProducerWarp = 4, kUmMmaWarp = 5, kUmConvWarp0 = 6;
constexpr int kUmThreads = (kUmConvWarp0 + kUmConvWarps) * 32;   // 512
constexpr int kUmTmemCols = 256;                          // two 128-column accumulators

struct UmSmem {
  alignas(128) float raw[kUmRawStages][5][kUmRows * 16];  // 80 KB
  alignas(128) unsigned char op[kUmOpStages][kUmOpBytes]; // 100.5 KB
  alignas(8) unsigned long long raw_full[kUmRawStages], raw_empty[kUmRawStages];
  alignas(8) unsigned long long op_full[kUmOpStages], op_empty[kUmOpStages];
  alignas(8) unsigned long long acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int is_last;
};
struct UmArgs {
  const float *blk[5];   // X, R, AR, P, AP
  int64_t n;
  int drain = kUmDrain;  // stages per accumulator hand-over (tests vary it)
};

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): D = F32 [4,6) = 1, A = B = TF32 [7,10) = [10,13) = 2,
// A and B K-major (bits 15, 16 clear), N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t kUmIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#ifdef __CUDACC__
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major without swizzle: start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30) (between the two core matrices along K: rows 0..3 / 4..7), stride byte offset >> 4
// in [32,46) (between 8-column groups along M/N), version 1 in [46,48), layout type 0 = no swizzle in [61,64).
// Element (column m, row k) of the operand: (m / 8) * SBO + (m % 8) * 16 + (k / 4) * LBO + (k % 4) * 4  (tools/umma_probe.cu)
__device__ __forceinline__ uint64_t um_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kUmHalfStride >> 4) << 16) |
         ((uint64_t)(kUmGroupStride >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void um_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(kUmIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void um_commit(unsigned long long *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void um_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"      // in the SAME asm statement: the outputs are not valid before the wait
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// mbarrier wait that turns a protocol bug into a trap (launch failure) instead of a hung GPU: ~1 s of polling
__device__ __forceinline__ void um_wait(unsigned long long *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 24)) __trap();
  }
}
// fp32 -> tf32 bit pattern, round to nearest with ties away from zero (what cvt.rna.tf32.f32 computes for finite values,
// without the Inf / NaN special-casing ptxas wraps around it: two integer operations instead of about nine instructions)
__device__ __forceinline__ uint32_t um_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// product index of (left block lb, right block rb) in the output (order of k_gram_rr_tc<2>), -1 if not needed;
// left blocks X R AR P = 0..3, right blocks R AR P AP = 0..3
__device__ __forceinline__ int um_product(int lb, int rb) {
  //            R   AR   P   AP
  // X          1    0   4    3
  // R          -    2   5    -
  // AR         -    -   6    -
  // P          -    -   -    7
  const int tab[4][4] = {{1, 0, 4, 3}, {-1, 2, 5, -1}, {-1, -1, 6, -1}, {-1, -1, -1, 7}};
  return tab[lb][rb];
}

// out[p * 256 + i * 16 + j] = (left block of product p)' (right block of product p), p < 8.
// partials: gridDim.x * 8 * 256 doubles; ticket: zero on entry, zero on exit.
__global__ void __launch_bounds__(kUmThreads, 1) k_gram_umma(UmArgs a, double *partials, unsigned int *ticket,
                                                             double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char um_smem_raw[];
  UmSmem *sm = reinterpret_cast<UmSmem *>(um_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = a.n;
  const int64_t nchunks = (n + kUmRows - 1) / kUmRows;
  // stages this CTA processes: chunks blockIdx.x, blockIdx.x + gridDim.x, ...
  const int nst = (int)((nchunks > (int64_t)blockIdx.x) ? (nchunks - 1 - blockIdx.x) / gridDim.x + 1 : 0);
  const int drain = a.drain;
  const int ngroups = (nst + drain - 1) / drain;

  if (tid == 0) {
    for (int s = 0; s < kUmRawStages; ++s) {
      mbar_init(&sm->raw_full[s], 1);
      mbar_init(&sm->raw_empty[s], kUmConvWarps);
    }
    for (int s = 0; s < kUmOpStages; ++s) {
      mbar_init(&sm->op_full[s], kUmConvWarps);
      mbar_init(&sm->op_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm->acc_full[s], 1);
      mbar_init(&sm->acc_empty[s], kUmEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kUmMmaWarp) {      // one warp allocates the tensor memory and owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)),
                 "n"(kUmTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm->tmem_base;

  float acc[4][16];              // epilogue warps: [right block][column]
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[rb][c] = 0.f;

  if (warp == kUmProducerWarp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int it = 0; it < nst; ++it) {
        const int s = it % kUmRawStages;
        const uint32_t ph = (uint32_t)((it / kUmRawStages) & 1);
        um_wait(&sm->raw_empty[s], ph ^ 1u);
        const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kUmRows;
        const int rows = (int)((n - r0 < kUmRows) ? (n - r0) : kUmRows);
        const uint32_t bytes = (uint32_t)rows * 16u * (uint32_t)sizeof(float);
        mbar_expect_tx(&sm->raw_full[s], bytes * 5u);
#pragma unroll
        for (int b = 0; b < 5; ++b) bulk_g2s(sm->raw[s][b], a.blk[b] + r0 * 16, bytes, &sm->raw_full[s], pol);
      }
    }
  } else if (warp == kUmMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int t = it % kUmOpStages;
        const uint32_t pht = (uint32_t)((it / kUmOpStages) & 1);
        const int g = it / drain, ab = g & 1;
        if (it % drain == 0) {                       // a new accumulation group: the accumulator must be drained
          um_wait(&sm->acc_empty[ab], (uint32_t)((g >> 1) & 1) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        um_wait(&sm->op_full[t], pht);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(sm->op[t]);
        const uint32_t d = tmem + (uint32_t)(ab * 128);
#pragma unroll
        for (int ks = 0; ks < kUmSteps; ++ks) {
          const uint32_t o = base + (uint32_t)(ks * kUmStepBytes);
          um_mma(d, um_desc(o), um_desc(o + 4u * kUmGroupStride), (it % drain != 0 || ks != 0) ? 1u : 0u);   // B: columns 32..159
        }
        um_commit(&sm->op_empty[t]);                    // arrives when these MMAs have read the operand stage
        if (it % drain == drain - 1 || it == nst - 1) um_commit(&sm->acc_full[ab]);
      }
    }
  } else if (warp >= kUmConvWarp0) {
    // ------------------------------------------------------------------ converters
    // thread <-> (block b, 8-row step ks, row half g, column quad q): rows ks*8 + g*4 .. +3, columns q*4 .. +3 of block b
    const int ct = tid - kUmConvWarp0 * 32;             // 0..319
    const int b = ct >> 6, ks = (ct >> 3) & 7, g = (ct >> 2) & 1, q = ct & 3;
    const int row0 = ks * 8 + g * 4;
    // hi tile of columns q*4.. of block b: operand columns b*32 + q*4 + c -> group b*4 + q/2, row-in-group (q%2)*4 + c;
    // the lo tile sits 16 columns = 2 groups further
    const int dst_off = ks * kUmStepBytes + g * kUmHalfStride + (b * 4 + (q >> 1)) * kUmGroupStride + (q & 1) * 64;
    // software pipeline over the stages: the 128-bit loads of stage it + 1 are issued before this thread waits for the
    // operand slot of stage it, so their latency (and the wait for the producer) hides behind the stores and the fence
    float4 v[4];
    auto load_stage = [&](int it) {
      const int s = it % kUmRawStages;
      const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kUmRows;
      um_wait(&sm->raw_full[s], (uint32_t)((it / kUmRawStages) & 1));
      // the two row halves of a quarter-warp read rows of opposite parity (g = 1 swaps its row pairs): no bank conflict
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = j ^ g;
        v[j] = *reinterpret_cast<const float4 *>(&sm->raw[s][b][(row0 + jj) * 16 + q * 4]);
        if (!(r0 + row0 + jj < n)) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (nst > 0) load_stage(0);
    for (int it = 0; it < nst; ++it) {
      const int s = it % kUmRawStages, t = it % kUmOpStages;
      const uint32_t pht = (uint32_t)((it / kUmOpStages) & 1);
      // hi / lo of the transposed 4 x 4 tile BEFORE the raw stage is released: the conversions consume the loaded
      // registers, so every load has completed when lane 0 arrives on raw_empty.  (Releasing right after issuing the loads
      // let the producer's next bulk copy overwrite the slice under loads still queued behind the operand stores of the
      // other warps -- measured: wrong Gram blocks as soon as a CTA reuses a raw stage; tools/diag_gram_umma2.py.)
      uint4 hi[4], lo[4];                               // [column]: its 4 consecutive rows are one 16-byte K chunk
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float xc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w = g ? v[k ^ 1] : v[k];        // row k of the tile (static register indices)
          xc[k] = c == 0 ? w.x : (c == 1 ? w.y : (c == 2 ? w.z : w.w));
        }
        hi[c].x = um_tf32(xc[0]); hi[c].y = um_tf32(xc[1]); hi[c].z = um_tf32(xc[2]); hi[c].w = um_tf32(xc[3]);
        lo[c].x = um_tf32(xc[0] - __uint_as_float(hi[c].x));
        lo[c].y = um_tf32(xc[1] - __uint_as_float(hi[c].y));
        lo[c].z = um_tf32(xc[2] - __uint_as_float(hi[c].z));
        lo[c].w = um_tf32(xc[3] - __uint_as_float(hi[c].w));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->raw_empty[s]);    // the slice has been consumed
      if (it + 1 < nst) load_stage(it + 1);
      um_wait(&sm->op_empty[t], pht ^ 1u);
      unsigned char *dst = sm->op[t] + dst_off;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4 *>(dst + c * 16) = hi[c];
        *reinterpret_cast<uint4 *>(dst + c * 16 + 2 * kUmGroupStride) = lo[c];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->op_full[t]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0..3 (left block = warp)
    const uint32_t need = warp == 0 ? 0xFu : (warp == 1 ? 0x6u : (warp == 2 ? 0x4u : 0x8u));   // right blocks needed
    for (int g = 0; g < ngroups; ++g) {
      const int ab = g & 1;
      um_wait(&sm->acc_full[ab], (uint32_t)((g >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t row_base = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ab * 128);
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        if ((need >> rb) & 1u) {
          float vh[16], vl[16];
          um_ld16(row_base + (uint32_t)(rb * 32), vh);          // (this row) x (right block rb, hi columns)
          um_ld16(row_base + (uint32_t)(rb * 32 + 16), vl);     // (this row) x (right block rb, lo columns)
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[rb][c] += vh[c] + vl[c];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->acc_empty[ab]);
    }
    // lanes l (hi row of left column l) and l + 16 (lo row of the same column) hold the two halves of G[l][:]
#pragma unroll
    for (int rb = 0; rb < 4; ++rb) {
      const int p = um_product(warp, rb);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float o = __shfl_xor_sync(0xffffffffu, acc[rb][c], 16);
        if (p >= 0 && lane < 16)
          partials[((size_t)blockIdx.x * 8 + p) * 256 + lane * 16 + c] = (double)acc[rb][c] + (double)o;
      }
    }
    __threadfence();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kUmMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kUmTmemCols) : "memory");
  }
  if (tid == 0) {
    __threadfence();
    sm->is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!sm->is_last) return;
  __threadfence();
  // the last CTA sums the per-CTA partials: 2048 outputs over 448 threads, CTAs in a fixed order (deterministic)
  for (int e = tid; e < 8 * 256; e += kUmThreads) {
    double s = 0.0;
    for (unsigned int blk = 0; blk < gridDim.x; ++blk) s += __ldcg(&partials[(size_t)blk * 2048 + e]);
    out[e] = s;
  }
  if (tid == 0) *ticket = 0u;
}
#endif  // __CUDACC__

}  // namespace b200
