// Attention for a handful of ragged rows (CUDA cores, one warp per (batch, head, row)).
// MaskGIT sequences are 256 image tokens + 1 class token (S = 257, muse/modeling_transformer.py:1407): the tcgen05
// kernels own 128-row tiles, which leaves exactly ONE query row / key row per (batch, head).  Running those through a
// 64-row tensor-core tile costs as much as a third of the whole backward; a warp per row needs ~70 KFLOP and is ~20x
// cheaper.  Used when (S mod 128) <= kMaxRows, otherwise the mma.sync kernels take the remainder.
#include "common.cuh"

namespace muse {
namespace {

constexpr int HD = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void load_row64(const bf16* p, float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    float t[8];
    load8(p + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c + j] = t[j];
  }
}
__device__ __forceinline__ float dot_row64(const bf16* p, const float (&q)[64]) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    float t[8];
    load8(p + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(t[j], q[c + j], acc);
  }
  return acc;
}
__device__ __forceinline__ void axpy_row64(float a, const bf16* p, float (&acc)[64]) {
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    float t[8];
    load8(p + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c + j] = fmaf(a, t[j], acc[c + j]);
  }
}
// warp-reduce 64 accumulators; lane l ends up holding elements 2l, 2l+1 -> written as one bf16x2
__device__ __forceinline__ void reduce_store_row64(float (&acc)[64], bf16* dst, int lane, float mul) {
  float mine0 = 0.f, mine1 = 0.f;
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    const float s = warp_sum(acc[d]);
    if ((d >> 1) == lane) { if (d & 1) mine1 = s; else mine0 = s; }
  }
  *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pack_bf16(mine0 * mul, mine1 * mul);
}

struct RowArgs {
  const bf16* q; const bf16* k; const bf16* v; const bf16* d_o;
  long long q_rs, k_rs, v_rs, do_rs;
  int B, nh, Sq, Skv, row0, nrows;
  float scale;
};

// forward for query rows [row0, row0 + nrows)
__global__ void __launch_bounds__(128)
attn_fwd_rows_kernel(RowArgs a, bf16* __restrict__ O, long long o_rs, float* __restrict__ LSE) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= a.B * a.nh * a.nrows) return;
  const int qi = a.row0 + w % a.nrows, h = (w / a.nrows) % a.nh, b = w / (a.nrows * a.nh);
  float q[64];
  load_row64(a.q + (static_cast<long long>(b) * a.Sq + qi) * a.q_rs + h * HD, q);
  const float sl2 = a.scale * kLog2e;
  float m = -INFINITY;
  for (int j = lane; j < a.Skv; j += 32)
    m = fmaxf(m, dot_row64(a.k + (static_cast<long long>(b) * a.Skv + j) * a.k_rs + h * HD, q) * sl2);
  m = warp_max(m);
  float l = 0.f, o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
  for (int j = lane; j < a.Skv; j += 32) {
    const float s = dot_row64(a.k + (static_cast<long long>(b) * a.Skv + j) * a.k_rs + h * HD, q) * sl2;
    const float p = bf16_round(exp2f(s - m));  // P is rounded to bf16 before the PV product, like the tensor-core path
    l += exp2f(s - m);
    axpy_row64(p, a.v + (static_cast<long long>(b) * a.Skv + j) * a.v_rs + h * HD, o);
  }
  l = warp_sum(l);
  reduce_store_row64(o, O + (static_cast<long long>(b) * a.Sq + qi) * o_rs + h * HD, lane, 1.f / l);
  if (lane == 0) LSE[(static_cast<long long>(b) * a.nh + h) * a.Sq + qi] = (m + log2f(l)) * kLn2;
}

// dQ (+ D) for query rows [row0, row0 + nrows)
__global__ void __launch_bounds__(128)
attn_bwd_dq_rows_kernel(RowArgs a, const float* __restrict__ LSE, float* __restrict__ Dv, bf16* __restrict__ dQ,
                        long long dq_rs) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= a.B * a.nh * a.nrows) return;
  const int qi = a.row0 + w % a.nrows, h = (w / a.nrows) % a.nh, b = w / (a.nrows * a.nh);
  const long long sidx = (static_cast<long long>(b) * a.nh + h) * a.Sq + qi;
  float q[64], g[64];
  load_row64(a.q + (static_cast<long long>(b) * a.Sq + qi) * a.q_rs + h * HD, q);
  load_row64(a.d_o + (static_cast<long long>(b) * a.Sq + qi) * a.do_rs + h * HD, g);
  const float sl2 = a.scale * kLog2e, lse2 = LSE[sidx] * kLog2e;
  float dsum = 0.f;
  for (int j = lane; j < a.Skv; j += 32) {
    const float p = exp2f(dot_row64(a.k + (static_cast<long long>(b) * a.Skv + j) * a.k_rs + h * HD, q) * sl2 - lse2);
    dsum = fmaf(p, dot_row64(a.v + (static_cast<long long>(b) * a.Skv + j) * a.v_rs + h * HD, g), dsum);
  }
  dsum = warp_sum(dsum);
  if (lane == 0) Dv[sidx] = dsum;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  for (int j = lane; j < a.Skv; j += 32) {
    const bf16* kr = a.k + (static_cast<long long>(b) * a.Skv + j) * a.k_rs + h * HD;
    const float p = exp2f(dot_row64(kr, q) * sl2 - lse2);
    const float dp = dot_row64(a.v + (static_cast<long long>(b) * a.Skv + j) * a.v_rs + h * HD, g);
    axpy_row64(bf16_round(p * (dp - dsum) * a.scale), kr, acc);
  }
  reduce_store_row64(acc, dQ + (static_cast<long long>(b) * a.Sq + qi) * dq_rs + h * HD, lane, 1.f);
}

// dK, dV for key rows [row0, row0 + nrows)
__global__ void __launch_bounds__(128)
attn_bwd_dkdv_rows_kernel(RowArgs a, const float* __restrict__ LSE, const float* __restrict__ Dv,
                          bf16* __restrict__ dK, long long dk_rs, bf16* __restrict__ dV, long long dv_rs) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= a.B * a.nh * a.nrows) return;
  const int kj = a.row0 + w % a.nrows, h = (w / a.nrows) % a.nh, b = w / (a.nrows * a.nh);
  const long long sbase = (static_cast<long long>(b) * a.nh + h) * a.Sq;
  float kk[64], vv[64];
  load_row64(a.k + (static_cast<long long>(b) * a.Skv + kj) * a.k_rs + h * HD, kk);
  load_row64(a.v + (static_cast<long long>(b) * a.Skv + kj) * a.v_rs + h * HD, vv);
  const float sl2 = a.scale * kLog2e;
  float ak[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) ak[d] = 0.f;
  // dK first, dV in a second sweep (keeps the accumulator count at 64 registers)
  for (int i = lane; i < a.Sq; i += 32) {
    const bf16* qr = a.q + (static_cast<long long>(b) * a.Sq + i) * a.q_rs + h * HD;
    const bf16* gr = a.d_o + (static_cast<long long>(b) * a.Sq + i) * a.do_rs + h * HD;
    const float p = exp2f(dot_row64(qr, kk) * sl2 - LSE[sbase + i] * kLog2e);
    const float dp = dot_row64(gr, vv);
    axpy_row64(bf16_round(p * (dp - Dv[sbase + i]) * a.scale), qr, ak);
  }
  reduce_store_row64(ak, dK + (static_cast<long long>(b) * a.Skv + kj) * dk_rs + h * HD, lane, 1.f);
#pragma unroll
  for (int d = 0; d < 64; ++d) ak[d] = 0.f;
  for (int i = lane; i < a.Sq; i += 32) {
    const bf16* qr = a.q + (static_cast<long long>(b) * a.Sq + i) * a.q_rs + h * HD;
    const bf16* gr = a.d_o + (static_cast<long long>(b) * a.Sq + i) * a.do_rs + h * HD;
    const float p = exp2f(dot_row64(qr, kk) * sl2 - LSE[sbase + i] * kLog2e);
    axpy_row64(bf16_round(p), gr, ak);
  }
  reduce_store_row64(ak, dV + (static_cast<long long>(b) * a.Skv + kj) * dv_rs + h * HD, lane, 1.f);
}

}  // namespace

constexpr int kMaxRows = 4;
bool attn_rows_ok(int nrows) { return nrows > 0 && nrows <= kMaxRows; }

static RowArgs make_args(const void* q, const void* k, const void* v, const void* d_o, int B, int nh, int Sq, int Skv,
                         int q_rs, int k_rs, int v_rs, int do_rs, float scale, int row0, int nrows) {
  RowArgs a;
  a.q = reinterpret_cast<const bf16*>(q); a.k = reinterpret_cast<const bf16*>(k); a.v = reinterpret_cast<const bf16*>(v);
  a.d_o = reinterpret_cast<const bf16*>(d_o);
  a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs; a.do_rs = do_rs;
  a.B = B; a.nh = nh; a.Sq = Sq; a.Skv = Skv; a.row0 = row0; a.nrows = nrows; a.scale = scale;
  return a;
}

int attn_fwd_rows(const void* q, const void* k, const void* v, void* o, float* lse, int B, int nh, int Sq, int Skv,
                  int q_rs, int k_rs, int v_rs, int o_rs, float scale, int row0, int nrows, cudaStream_t s) {
  RowArgs a = make_args(q, k, v, nullptr, B, nh, Sq, Skv, q_rs, k_rs, v_rs, 0, scale, row0, nrows);
  attn_fwd_rows_kernel<<<ceil_div(B * nh * nrows, 4), 128, 0, s>>>(a, reinterpret_cast<bf16*>(o), o_rs, lse);
  return check_launch("attn_fwd_rows");
}

int attn_bwd_dq_rows(const void* q, const void* k, const void* v, const void* d_o, const float* lse, float* dvec,
                     void* dq, int B, int nh, int Sq, int Skv, int q_rs, int k_rs, int v_rs, int do_rs, int dq_rs,
                     float scale, int row0, int nrows, cudaStream_t s) {
  RowArgs a = make_args(q, k, v, d_o, B, nh, Sq, Skv, q_rs, k_rs, v_rs, do_rs, scale, row0, nrows);
  attn_bwd_dq_rows_kernel<<<ceil_div(B * nh * nrows, 4), 128, 0, s>>>(a, lse, dvec, reinterpret_cast<bf16*>(dq), dq_rs);
  return check_launch("attn_bwd_dq_rows");
}

int attn_bwd_dkdv_rows(const void* q, const void* k, const void* v, const void* d_o, const float* lse,
                       const float* dvec, void* dk, void* dv, int B, int nh, int Sq, int Skv, int q_rs, int k_rs,
                       int v_rs, int do_rs, int dk_rs, int dv_rs, float scale, int row0, int nrows, cudaStream_t s) {
  RowArgs a = make_args(q, k, v, d_o, B, nh, Sq, Skv, q_rs, k_rs, v_rs, do_rs, scale, row0, nrows);
  attn_bwd_dkdv_rows_kernel<<<ceil_div(B * nh * nrows, 4), 128, 0, s>>>(a, lse, dvec, reinterpret_cast<bf16*>(dk), dk_rs,
                                                                        reinterpret_cast<bf16*>(dv), dv_rs);
  return check_launch("attn_bwd_dkdv_rows");
}

}  // namespace muse
