// linop.cuh -- shared helpers of the *_op (callback operator) entry points.
#pragma once
#include "pass.cuh"

namespace b200 {
int check_linop(const b200_linop *A, const char *what);   // qmr.cu
}
