// umma_probe.cu -- empirical probe of the tcgen05.mma kind::tf32 shared-memory operand layout (MN-major, no swizzle).
// One CTA.  Region A (bytes [0, REGION)) holds float(i + 1) at float index i; region B is all zero except one float
// (index ob) = 1.0.  D = A^T B then has one nonzero column n*, whose entries are the A values sharing B's k:
// D[m][n*] - 1 = float index of A(m, k*).  Sweeping ob over the region gives the complete (m, k) -> offset map of A
// and the (n, k) -> offset map of B for the descriptor under test.  Not part of the product; run under gpurun.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool wait_bounded(unsigned long long *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}

struct Params { uint32_t lbo, sbo, a_major, b_major, region_floats, b_off_bytes, M, N; };

__global__ void __launch_bounds__(128, 1) k_probe(Params p, int *out /* region_floats x 8 */, int *status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  float *A = reinterpret_cast<float *>(smem);
  float *B = reinterpret_cast<float *>(smem + p.b_off_bytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (p.a_major << 15) | (p.b_major << 16) | ((p.N >> 3) << 17) |
                         ((p.M >> 4) << 24);
  auto desc = [&](uint32_t addr) -> uint64_t {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(p.lbo >> 4) << 16) | ((uint64_t)(p.sbo >> 4) << 32) | (1ull << 46);
  };
  for (uint32_t i = tid; i < p.region_floats; i += 128) A[i] = (float)(i + 1);
  uint32_t phase = 0;
  for (uint32_t ob = 0; ob < p.region_floats; ++ob) {
    for (uint32_t i = tid; i < p.region_floats; i += 128) B[i] = (i == ob) ? 1.0f : 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t da = desc(smem_u32(A)), db = desc(smem_u32(B));
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}\n"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (!wait_bounded(&bar, phase)) { if (tid == 0) *status = 100 + (int)ob; break; }
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread tid <-> TMEM lane tid (= row m of D); scan the N columns for the nonzero one
    int nstar = -1; float val = 0.f; int count = 0;
    for (uint32_t c0 = 0; c0 < p.N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                   "tcgen05.wait::ld.sync.aligned;"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr) : "memory");
      for (int j = 0; j < 16; ++j) {
        const float v = __uint_as_float(r[j]);
        if (v != 0.f) { if (count == 0) { nstar = (int)c0 + j; val = v; } ++count; }
      }
    }
    // out row: [n* seen by m=0, count(m=0), Aoff(m=0), Aoff(m=1), Aoff(m=4), Aoff(m=5), Aoff(m=32), Aoff(m=127)]
    int *o = out + (size_t)ob * 8;
    const int aoff = count ? (int)(val + 0.5f) - 1 : -1;
    if (tid == 0) { o[0] = nstar; o[1] = count; o[2] = aoff; }
    if (tid == 1) o[3] = aoff;
    if (tid == 4) o[4] = aoff;
    if (tid == 5) o[5] = aoff;
    if (tid == 32) o[6] = aoff;
    if (tid == 127) o[7] = aoff;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main(int argc, char **argv) {
  Params p;
  p.lbo = argc > 1 ? atoi(argv[1]) : 128;
  p.sbo = argc > 2 ? atoi(argv[2]) : 160;
  p.a_major = argc > 3 ? atoi(argv[3]) : 1;
  p.b_major = argc > 4 ? atoi(argv[4]) : 1;
  p.region_floats = argc > 5 ? atoi(argv[5]) : 1280;
  p.M = 128; p.N = 128;
  p.b_off_bytes = 16384;
  int *d_out, *d_status;
  cudaMalloc(&d_out, sizeof(int) * p.region_floats * 8);
  cudaMalloc(&d_status, sizeof(int));
  cudaMemset(d_out, 0xff, sizeof(int) * p.region_floats * 8);
  cudaMemset(d_status, 0, sizeof(int));
  const int smem = 32768;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(p, d_out, d_status);
  cudaError_t e = cudaDeviceSynchronize();
  int status = 0;
  cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost);
  printf("# lbo=%u sbo=%u a_major=%u b_major=%u region_floats=%u cuda=%s status=%d\n", p.lbo, p.sbo, p.a_major, p.b_major,
         p.region_floats, cudaGetErrorString(e), status);
  int *h = (int *)malloc(sizeof(int) * p.region_floats * 8);
  cudaMemcpy(h, d_out, sizeof(int) * p.region_floats * 8, cudaMemcpyDeviceToHost);
  printf("# ob nstar count Aoff(m=0) Aoff(1) Aoff(4) Aoff(5) Aoff(32) Aoff(127)\n");
  for (uint32_t ob = 0; ob < p.region_floats; ++ob)
    printf("%u %d %d %d %d %d %d %d %d\n", ob, h[ob * 8], h[ob * 8 + 1], h[ob * 8 + 2], h[ob * 8 + 3], h[ob * 8 + 4], h[ob * 8 + 5],
           h[ob * 8 + 6], h[ob * 8 + 7]);
  return 0;
}
