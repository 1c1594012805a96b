// extern "C" surface of the TEST-ONLY cross-check library (first-generation mma.sync kernels).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace muse {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e)); return MUSE_ERR_CUDA; }
  return MUSE_OK;
}
int gemm_mma(const void*, const void*, void*, const float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int attn_fwd_legacy(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int, int, int, int, float, cudaStream_t);
int attn_bwd_legacy(const void*, const void*, const void*, const void*, const void*, const float*, float*, void*, void*, void*, int, int, int, int, int, int, int, int, int, int, int, int, int, float, cudaStream_t);
}  // namespace muse

using namespace muse;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {
const char* xcheck_last_error(void) { return g_err; }
int xcheck_gemm_mma(const void* A, const void* B, void* C, const float* res, int M, int N, int K, int lda, int ldb, int ldc,
                    int a_mn, int b_mn, int epilogue, void* stream) {
  return gemm_mma(A, B, C, res, M, N, K, lda, ldb, ldc, a_mn, b_mn, epilogue, ST(stream));
}
int xcheck_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int nh, int Sq, int Skv,
                    int head_dim, int q_rs, int k_rs, int v_rs, int o_rs, float scale, void* stream) {
  return attn_fwd_legacy(q, k, v, o, lse, B, nh, Sq, Skv, head_dim, q_rs, k_rs, v_rs, o_rs, scale, ST(stream));
}
int xcheck_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, float* dvec,
                    void* dq, void* dk, void* dv, int B, int nh, int Sq, int Skv, int head_dim, int q_rs, int k_rs, int v_rs,
                    int o_rs, int do_rs, int dq_rs, int dk_rs, int dv_rs, float scale, void* stream) {
  return attn_bwd_legacy(q, k, v, o, d_o, lse, dvec, dq, dk, dv, B, nh, Sq, Skv, head_dim, q_rs, k_rs, v_rs, o_rs, do_rs, dq_rs,
                         dk_rs, dv_rs, scale, ST(stream));
}
}
