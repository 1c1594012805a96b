// Attention for a handful of ragged rows (CUDA cores, one 128-thread CTA per (batch, head, row)).
// MaskGIT sequences are 256 image tokens + 1 class token (S = 257, muse/modeling_transformer.py:1407): the tcgen05
// kernels own 128-row tiles, which leaves exactly ONE query row / key row per (batch, head).  Running those through a
// 64-row tensor-core tile cost a third of the whole attention backward; one small CTA per row is ~10x cheaper.
// Phase 1: each thread owns rows of the looped dimension (row-wise 128-byte reads, dot products against the fixed
// row held in shared memory); phase 2: each thread owns one output column (column-wise, coalesced reads).
// Used when (S mod 128) <= kMaxRows, otherwise the mma.sync kernels take the remainder.
#include "common.cuh"

namespace muse {
namespace {

constexpr int HD = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float dot_row64(const bf16* p, const float* q) {
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

// block-wide reductions over 128 threads (4 warps)
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

// out[d] = sum_j wgt[j] * M[j][d] over j in [0, n).  thread = (8-column group dg = tid & 7, row group jg = tid >> 3):
// 16-byte loads, 16 row groups in flight, partial sums combined through shared memory (part: [16][64] floats).
__device__ __forceinline__ void weighted_colsum(const float* wgt, const bf16* M, long long rs, int n, float* part,
                                                bf16* dst, float mul) {
  const int dg = threadIdx.x & 7, jg = threadIdx.x >> 3;
  __syncthreads();  // wgt[] was just written by other threads of the CTA
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (int j = jg; j < n; j += 16) {
    float t[8];
    load8(M + static_cast<long long>(j) * rs + dg * 8, t);
    const float w = wgt[j];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, t[k], acc[k]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) part[jg * 64 + dg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) sum += part[g * 64 + threadIdx.x];
    dst[threadIdx.x] = __float2bfloat16_rn(sum * mul);
  }
}

struct RowArgs {
  const bf16* q; const bf16* k; const bf16* v; const bf16* d_o;
  long long q_rs, k_rs, v_rs, do_rs;
  int B, nh, Sq, Skv, row0, nrows;
  float scale;
};

// forward for query rows [row0, row0 + nrows): grid = (nrows, nh, B)
__global__ void __launch_bounds__(128, 4)
attn_fwd_rows_kernel(RowArgs a, bf16* __restrict__ O, long long o_rs, float* __restrict__ LSE) {
  extern __shared__ float sm[];  // [Skv] scores -> probabilities
  __shared__ float fixed[64], part[16 * 64], red[4];
  const int qi = a.row0 + blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  if (tid < 64) fixed[tid] = __bfloat162float(a.q[(static_cast<long long>(b) * a.Sq + qi) * a.q_rs + h * HD + tid]);
  __syncthreads();
  const bf16* kb = a.k + static_cast<long long>(b) * a.Skv * a.k_rs + h * HD;
  const bf16* vb = a.v + static_cast<long long>(b) * a.Skv * a.v_rs + h * HD;
  const float sl2 = a.scale * kLog2e;
  float mx = -INFINITY;
  for (int j = tid; j < a.Skv; j += 128) {
    const float s = dot_row64(kb + static_cast<long long>(j) * a.k_rs, fixed) * sl2;
    sm[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_max(mx, red);
  float l = 0.f;
  for (int j = tid; j < a.Skv; j += 128) {
    const float p = exp2f(sm[j] - mx);
    l += p;
    sm[j] = bf16_round(p);  // P is rounded to bf16 before the PV product, like the tensor-core path
  }
  l = block_sum(l, red);
  weighted_colsum(sm, vb, a.v_rs, a.Skv, part, O + (static_cast<long long>(b) * a.Sq + qi) * o_rs + h * HD, 1.f / l);
  if (tid == 0) LSE[(static_cast<long long>(b) * a.nh + h) * a.Sq + qi] = (mx + log2f(l)) * kLn2;
}

// dQ (+ D) for query rows [row0, row0 + nrows)
__global__ void __launch_bounds__(128, 4)
attn_bwd_dq_rows_kernel(RowArgs a, const float* __restrict__ LSE, float* __restrict__ Dv, bf16* __restrict__ dQ,
                        long long dq_rs) {
  extern __shared__ float sm[];  // [2][Skv]: p, dp -> ds
  __shared__ float fq[64], fg[64], part[16 * 64], red[4];
  const int qi = a.row0 + blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const long long sidx = (static_cast<long long>(b) * a.nh + h) * a.Sq + qi;
  if (tid < 64) {
    fq[tid] = __bfloat162float(a.q[(static_cast<long long>(b) * a.Sq + qi) * a.q_rs + h * HD + tid]);
    fg[tid] = __bfloat162float(a.d_o[(static_cast<long long>(b) * a.Sq + qi) * a.do_rs + h * HD + tid]);
  }
  __syncthreads();
  const bf16* kb = a.k + static_cast<long long>(b) * a.Skv * a.k_rs + h * HD;
  const bf16* vb = a.v + static_cast<long long>(b) * a.Skv * a.v_rs + h * HD;
  const float sl2 = a.scale * kLog2e, lse2 = LSE[sidx] * kLog2e;
  float* sp = sm;
  float* sdp = sm + a.Skv;
  float dsum = 0.f;
  for (int j = tid; j < a.Skv; j += 128) {
    const float p = exp2f(dot_row64(kb + static_cast<long long>(j) * a.k_rs, fq) * sl2 - lse2);
    const float dp = dot_row64(vb + static_cast<long long>(j) * a.v_rs, fg);
    sp[j] = p;
    sdp[j] = dp;
    dsum = fmaf(p, dp, dsum);
  }
  dsum = block_sum(dsum, red);
  if (tid == 0) Dv[sidx] = dsum;
  for (int j = tid; j < a.Skv; j += 128) sp[j] = bf16_round(sp[j] * (sdp[j] - dsum) * a.scale);
  weighted_colsum(sp, kb, a.k_rs, a.Skv, part, dQ + (static_cast<long long>(b) * a.Sq + qi) * dq_rs + h * HD, 1.f);
}

// dK, dV for key rows [row0, row0 + nrows)
__global__ void __launch_bounds__(128, 4)
attn_bwd_dkdv_rows_kernel(RowArgs a, const float* __restrict__ LSE, const float* __restrict__ Dv,
                          bf16* __restrict__ dK, long long dk_rs, bf16* __restrict__ dV, long long dv_rs) {
  extern __shared__ float sm[];  // [2][Sq]: ds, p
  __shared__ float fk[64], fv[64], part[16 * 64];
  const int kj = a.row0 + blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const long long sbase = (static_cast<long long>(b) * a.nh + h) * a.Sq;
  if (tid < 64) {
    fk[tid] = __bfloat162float(a.k[(static_cast<long long>(b) * a.Skv + kj) * a.k_rs + h * HD + tid]);
    fv[tid] = __bfloat162float(a.v[(static_cast<long long>(b) * a.Skv + kj) * a.v_rs + h * HD + tid]);
  }
  __syncthreads();
  const bf16* qb = a.q + static_cast<long long>(b) * a.Sq * a.q_rs + h * HD;
  const bf16* gb = a.d_o + static_cast<long long>(b) * a.Sq * a.do_rs + h * HD;
  const float sl2 = a.scale * kLog2e;
  float* sds = sm;
  float* spp = sm + a.Sq;
  for (int i = tid; i < a.Sq; i += 128) {
    const float p = exp2f(dot_row64(qb + static_cast<long long>(i) * a.q_rs, fk) * sl2 - LSE[sbase + i] * kLog2e);
    const float dp = dot_row64(gb + static_cast<long long>(i) * a.do_rs, fv);
    sds[i] = bf16_round(p * (dp - Dv[sbase + i]) * a.scale);
    spp[i] = bf16_round(p);
  }
  weighted_colsum(sds, qb, a.q_rs, a.Sq, part, dK + (static_cast<long long>(b) * a.Skv + kj) * dk_rs + h * HD, 1.f);
  weighted_colsum(spp, gb, a.do_rs, a.Sq, part, dV + (static_cast<long long>(b) * a.Skv + kj) * dv_rs + h * HD, 1.f);
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
  if (Skv > 8192) { set_last_error("attn rows: Skv too long"); return MUSE_ERR_UNSUPPORTED; }
  RowArgs a = make_args(q, k, v, nullptr, B, nh, Sq, Skv, q_rs, k_rs, v_rs, 0, scale, row0, nrows);
  attn_fwd_rows_kernel<<<dim3(nrows, nh, B), 128, Skv * sizeof(float), s>>>(a, reinterpret_cast<bf16*>(o), o_rs, lse);
  return check_launch("attn_fwd_rows");
}

int attn_bwd_dq_rows(const void* q, const void* k, const void* v, const void* d_o, const float* lse, float* dvec,
                     void* dq, int B, int nh, int Sq, int Skv, int q_rs, int k_rs, int v_rs, int do_rs, int dq_rs,
                     float scale, int row0, int nrows, cudaStream_t s) {
  if (Skv > 4096) { set_last_error("attn rows: Skv too long"); return MUSE_ERR_UNSUPPORTED; }
  RowArgs a = make_args(q, k, v, d_o, B, nh, Sq, Skv, q_rs, k_rs, v_rs, do_rs, scale, row0, nrows);
  attn_bwd_dq_rows_kernel<<<dim3(nrows, nh, B), 128, 2 * Skv * sizeof(float), s>>>(a, lse, dvec, reinterpret_cast<bf16*>(dq), dq_rs);
  return check_launch("attn_bwd_dq_rows");
}

int attn_bwd_dkdv_rows(const void* q, const void* k, const void* v, const void* d_o, const float* lse,
                       const float* dvec, void* dk, void* dv, int B, int nh, int Sq, int Skv, int q_rs, int k_rs,
                       int v_rs, int do_rs, int dk_rs, int dv_rs, float scale, int row0, int nrows, cudaStream_t s) {
  if (Sq > 4096) { set_last_error("attn rows: Sq too long"); return MUSE_ERR_UNSUPPORTED; }
  RowArgs a = make_args(q, k, v, d_o, B, nh, Sq, Skv, q_rs, k_rs, v_rs, do_rs, scale, row0, nrows);
  attn_bwd_dkdv_rows_kernel<<<dim3(nrows, nh, B), 128, 2 * Sq * sizeof(float), s>>>(
      a, lse, dvec, reinterpret_cast<bf16*>(dk), dk_rs, reinterpret_cast<bf16*>(dv), dv_rs);
  return check_launch("attn_bwd_dkdv_rows");
}

}  // namespace muse
