// stubs.cu -- entry points not implemented yet return B200_ERR_UNSUPPORTED (never a CPU fallback).
#include "common.cuh"
using namespace b200;
#define STUB(name) { set_error(name " is not implemented yet in this build"); return B200_ERR_UNSUPPORTED; }
extern "C" {
int b200_lobpcg_solve(b200_ctx *, const b200_csr *, void *, int64_t, const b200_lobpcg_opts *, b200_lobpcg_result *, double *, double *) STUB("b200_lobpcg_solve")
}
