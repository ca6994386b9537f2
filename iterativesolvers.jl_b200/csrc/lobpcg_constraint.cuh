// lobpcg_constraint.cuh -- the device Constraint object (reference src/lobpcg.jl:144-224) shared by
// lobpcg_constraint.cu (C ABI) and lobpcg.cu (the engine applies it to its row-major blocks).
#pragma once
#include <vector>

#include "common.cuh"

struct b200_lobpcg_constraint {
  b200_ctx *ctx = nullptr;
  int dtype = B200_F64;
  int64_t n = 0;            // local rows
  int64_t ld = 0;           // leading dimension of Y (elements)
  int nc = 0, cap = 0;      // columns in use / allocated
  void *Y = nullptr;        // device, column-major n x cap (own copy: the reference keeps C alive the same way)
  void *BY = nullptr;       // generalized problem: B*Y (b200_lobpcg_constraint_create_b); nullptr: BY aliases Y (:163-164)
  b200_linop Bfn{};         // generalized problem: copy of the caller's B descriptor (update! needs B*X, :188-206)
  double *g_dev = nullptr;  // device scratch: cap x 16 doubles
  std::vector<double> U;    // host: upper Cholesky factor of Y'Y, nc x nc column-major
  std::vector<double> g_host;
};

namespace b200 {
// X <- X - Y (U'U \ (Y' X)) on a block with strides (rs, cs) and bs <= 16 columns (dtype of the constraint)
int constraint_apply_block(b200_ctx *ctx, const b200_lobpcg_constraint *c, void *X, int64_t rs, int64_t cs, int bs);
}
