// peer.cuh -- NVLink peer-memory collectives fused into the solver kernels (one process per GPU).
//
// Every rank owns one "comm buffer" in device memory; all ranks map all buffers with CUDA IPC
// (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, handles all-gathered once over NCCL).  Through NVSwitch
// every peer is one hop away at full NVLink bandwidth, so the two per-iteration collectives of the
// Krylov loops become plain stores/loads on mapped peer pointers inside our own kernels:
//
//   * scalar allreduce (one-shot, all-to-all): the block that finishes a grid-wide reduction writes its
//     partial into slot[parity][rank] of EVERY peer's buffer, fences (system scope), writes a sequence
//     flag to every peer, spins on its own W flags and sums the W slots in rank order -- identical value on
//     every rank, deterministic, ~2 NVLink one-way latencies, no NCCL launch and no extra kernel.  Slots are
//     double-buffered by sequence parity: a rank can be at most one collective ahead of its peers.
//   * halo push: k_halo_push stores x[send_idx] straight into the neighbours' halo segment (which IS the
//     halo part of their extended vector) and then raises a flag; the consumer SpMV spins on the flags of
//     the ranks it receives from before its first gather.  Inside the CG loop a single halo buffer is
//     safe: a neighbour can only start pushing iteration i+1 after the allreduce of iteration i, i.e. after
//     this rank's SpMV of iteration i has completed.
//
// Spins are bounded (kPeerSpinLimit); on timeout an error flag is raised instead of hanging the GPU.
// Buffer layout (bytes):  [0, 64K) slots + flags | [64K, ...) halo segment of this rank
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b200 {

constexpr int kPeerMaxWorld = 8;
constexpr size_t kPeerHeaderBytes = 65536;
constexpr size_t kPeerBufferBytes = (size_t)96 << 20;    // header + up to 96 MB - 64 KB of halo values
constexpr int kPeerArWidth = 64;                        // doubles per allreduce (block of dots in CGS)
constexpr unsigned long long kPeerSpinLimit = 1ull << 28;

// header layout (all 8-byte words)
struct PeerHeader {
  double slot[2][kPeerMaxWorld][kPeerArWidth];          // [parity][source rank][k]
  unsigned long long ar_flag[2][kPeerMaxWorld];         // sequence number last published by `source`
  unsigned long long halo_flag[kPeerMaxWorld];          // halo sequence last pushed by `source`
  unsigned long long error;                             // != 0: a bounded spin timed out somewhere
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "peer header too large");

// what kernels need (passed by value)
struct PeerView {
  int world = 1, rank = 0;
  PeerHeader *hdr[kPeerMaxWorld] = {nullptr};           // hdr[rank] is the local one
};

#ifdef __CUDACC__

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// One-shot allreduce(sum) of `count` doubles held by ONE calling thread (vals[0..count)); every rank
// must call it with the same `seq` (strictly increasing per context).  Result overwrites vals.
__device__ __forceinline__ void peer_allreduce_sum(const PeerView &pv, double *vals, int count,
                                                   unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  for (int q = 0; q < pv.world; ++q)
    for (int k = 0; k < count; ++k) st_relaxed_sys_f64(&pv.hdr[q]->slot[par][pv.rank][k], vals[k]);
  __threadfence_system();
  for (int q = 0; q < pv.world; ++q) st_release_sys(&pv.hdr[q]->ar_flag[par][pv.rank], seq);
  PeerHeader *me = pv.hdr[pv.rank];
  for (int q = 0; q < pv.world; ++q) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(&me->ar_flag[par][q]) != seq) {
      if (++spins > kPeerSpinLimit) {
        me->error = seq;
        break;
      }
    }
  }
  for (int k = 0; k < count; ++k) {
    double s = 0.0;
    for (int q = 0; q < pv.world; ++q) s += ld_relaxed_sys_f64(&me->slot[par][q][k]);   // rank order: same on all ranks
    vals[k] = s;
  }
}

// The same collective for ONE double, issued by a whole warp: lane q talks to rank q -- its slot store, its flag store
// (st.release orders the two for the peer) and its poll run concurrently with the other lanes', so the critical path is one
// NVLink store + one flag latency instead of `world` system-scope releases and `world` polls in sequence (r1: ~20 us of
// every K2 / K3 at 8 GPUs).  `val` is taken from lane 0; every lane returns the sum, formed in rank order (identical on
// all ranks).  All 32 lanes must call.
__device__ __forceinline__ double peer_allreduce_sum_warp(const PeerView &pv, double val, unsigned long long seq) {
  const int lane = (int)(threadIdx.x & 31u);
  val = __shfl_sync(0xffffffffu, val, 0);
  const int par = (int)(seq & 1ull);
  PeerHeader *me = pv.hdr[pv.rank];
  double got = 0.0;
  if (lane < pv.world) {
    PeerHeader *peer = pv.hdr[lane];
    st_relaxed_sys_f64(&peer->slot[par][pv.rank][0], val);
    st_release_sys(&peer->ar_flag[par][pv.rank], seq);
    unsigned long long spins = 0;
    while (ld_acquire_sys(&me->ar_flag[par][lane]) != seq) {
      if (++spins > kPeerSpinLimit) {
        me->error = seq;
        break;
      }
    }
    got = ld_relaxed_sys_f64(&me->slot[par][lane][0]);
  }
  double s = 0.0;
  for (int q = 0; q < pv.world; ++q) s += __shfl_sync(0xffffffffu, got, q);   // rank order: same on all ranks
  return s;
}

// wait until every rank in `from_mask` has pushed halo sequence `seq` (called by one thread per CTA)
__device__ __forceinline__ void peer_wait_halo(const PeerView &pv, unsigned int from_mask, unsigned long long seq) {
  PeerHeader *me = pv.hdr[pv.rank];
  for (int q = 0; q < pv.world; ++q) {
    if (!((from_mask >> q) & 1u)) continue;
    unsigned long long spins = 0;
    while (ld_acquire_sys(&me->halo_flag[q]) < seq) {
      if (++spins > kPeerSpinLimit) {
        me->error = seq;
        break;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace b200
