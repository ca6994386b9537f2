// spmv.cuh -- CSR SpMV building blocks shared by the solver kernels.
//
// Layout in HBM: rowptr int32 (m+1), colind int32 (nnz, local extended index), vals T (nnz), all
// contiguous in row order => a contiguous chunk of rows is a contiguous chunk of the colind/vals
// streams (12 B/nnz in fp64, read exactly once per SpMV); x is gathered through L1/L2.
//
// Algorithmic bytes per SpMV: nnz*(V+4) + (m+1)*4 + 2*m*V.
#pragma once
#include "csr.cuh"

namespace b200 {

#ifdef __CUDACC__

// lanes-per-row selection: smallest power of two >= average row length, in [2, 32]
inline int pick_lpr(double avg_row_nnz) {
  int l = 2;
  while (l < 32 && l < avg_row_nnz) l <<= 1;
  return l;
}

// x gather from the extended vector: own slab or halo buffer
template <typename T>
struct XView {
  const T *__restrict__ x;     // own rows [0, m)
  const T *__restrict__ halo;  // halo values, already shifted by -m (halo_shifted[col] valid for col >= m)
  int m;
  __device__ __forceinline__ T operator()(int col) const { return col < m ? __ldg(x + col) : __ldg(halo + col); }
};
template <typename T>
inline XView<T> make_xview(const b200_csr *A, const void *x_dev, bool peer_halo = false) {
  XView<T> v;
  v.x = (const T *)x_dev;
  const void *h = peer_halo ? A->halo_peer : A->halo;   // which halo buffer the preceding exchange filled
  v.halo = h ? (const T *)h - A->m_local : (const T *)x_dev;
  v.m = (int)A->m_local;
  return v;
}

// One sub-warp of LPR lanes computes (A x)[row]; result valid in all LPR lanes.
template <typename T, int LPR, typename XV>
__device__ __forceinline__ T row_dot(const int *__restrict__ rowptr, const int *__restrict__ colind,
                                     const T *__restrict__ vals, const XV &xv, int64_t row, int sub) {
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  const uint64_t pol = policy_evict_first();
  T acc = (T)0;
  for (int k = b + sub; k < e; k += LPR) {
    const int c = ld_stream<int>(colind + k, pol);
    const T a = ld_stream<T>(vals + k, pol);
    acc += a * xv(c);
  }
#pragma unroll
  for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
  return acc;
}

#endif  // __CUDACC__

}  // namespace b200
