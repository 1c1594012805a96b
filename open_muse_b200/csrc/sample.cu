// One generate2 decoding step (sample -> confidence -> per-row k-th smallest -> re-mask) as ONE kernel.
// Reference: muse/modeling_transformer.py:1424-1454 + muse/sampling.py:9-15,30-35, which is ~25 ATen launches
// per step (softmax, multinomial = exponential_ + div + argmax, where, gather, where, log, uniform_, log, log,
// mul, add, sort, gather, lt, where ...) over a [B, L, K] fp32 probability tensor.
//
// Randomness enters as PRE-DRAWN noise so the torch generator stream is the reference's own:
//   q_exp [B, L, K] ~ Exp(1)   torch.multinomial(p, 1) == argmax_c(p_c / q_c)   (ATen's n_sample == 1 recipe)
//   u     [B, L]    ~ U(0, 1)  gumbel = -log(-log(u))  (sampling.py:13-15, with the 1e-20 clamps of :9-10)
// Per row b (one CTA, one warp per token in turn):
//   m = max_c x_c ; e_c = exp(x_c - m) ; sampled = first argmax_c e_c / q_c      (softmax normaliser cancels)
//   p_sel = e_sampled / sum_c e_c ; known tokens keep their id and get p_sel = FLT_MAX (:1430-1441)
//   conf = log(max(p_sel,1e-20)) + temperature * gumbel(u)
//   k = max(1, min(#unknown - 1, mask_len))  (:1443-1448);  cut = (k+1)-th smallest conf (sorted[k])
//   next_id = conf < cut ? mask_id : sampled                                       (sampling.py:32-35, :1454)
// "conf_i < sorted[k]"  <=>  #{j : conf_j <= conf_i} <= k, evaluated by counting (L <= 1024): no sort needed.
// Optional classifier-free guidance: x = unc + g * (cond - unc) (:1410-1414), fused into the logit read.
#include <float.h>

#include "common.cuh"

namespace muse {
namespace {

__global__ void __launch_bounds__(1024)
sample_step_kernel(const bf16* __restrict__ logits, const bf16* __restrict__ logits_unc, long long row_stride,
                   long long batch_stride, float guidance, const long long* __restrict__ input_ids,
                   const float* __restrict__ q_exp, const float* __restrict__ u, long long* __restrict__ sampled_out,
                   long long* __restrict__ next_ids, float* __restrict__ conf_out, int L, int K, long long mask_id,
                   int mask_len, float temperature) {
  extern __shared__ float s_conf[];  // [L]
  __shared__ int s_unknown;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_unknown = 0;
  __syncthreads();
  int my_unknown = 0;
  for (int tkn = warp; tkn < L; tkn += nwarps) {
    const bf16* x = logits + b * batch_stride + tkn * row_stride;
    const bf16* xu = logits_unc ? logits_unc + b * batch_stride + tkn * row_stride : nullptr;
    const float* q = q_exp + (static_cast<long long>(b) * L + tkn) * K;
    // pass 1: row max
    float m = -INFINITY;
    for (int c = lane; c < K; c += 32) {
      float v = __bfloat162float(x[c]);
      if (xu) { const float w = __bfloat162float(xu[c]); v = w + guidance * (v - w); }
      m = fmaxf(m, v);
    }
    m = warp_max(m);
    // pass 2: sum of exp, arg-max of exp / q (lowest index on ties)
    float sum = 0.f, best = -1.f;
    int besti = 0;
    float e_best = 0.f;
    for (int c = lane; c < K; c += 32) {
      float v = __bfloat162float(x[c]);
      if (xu) { const float w = __bfloat162float(xu[c]); v = w + guidance * (v - w); }
      const float e = expf(v - m);
      sum += e;
      const float sc = e / q[c];
      if (sc > best) { best = sc; besti = c; e_best = e; }
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      const float oe = __shfl_xor_sync(0xffffffffu, e_best, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; e_best = oe; }
    }
    if (lane == 0) {
      const long long cur = input_ids[static_cast<long long>(b) * L + tkn];
      const bool unknown = (cur == mask_id);
      const long long sid = unknown ? static_cast<long long>(besti) : cur;
      const float p_sel = unknown ? e_best / sum : FLT_MAX;
      const float uu = u[static_cast<long long>(b) * L + tkn];
      const float gum = -logf(fmaxf(-logf(fmaxf(uu, 1e-20f)), 1e-20f));
      const float conf = logf(fmaxf(p_sel, 1e-20f)) + temperature * gum;
      s_conf[tkn] = conf;
      sampled_out[static_cast<long long>(b) * L + tkn] = sid;
      if (conf_out) conf_out[static_cast<long long>(b) * L + tkn] = conf;
      my_unknown += unknown ? 1 : 0;
    }
  }
  if (lane == 0 && my_unknown) atomicAdd(&s_unknown, my_unknown);
  __syncthreads();
  int k = min(s_unknown - 1, mask_len);
  k = max(1, k);
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float ci = s_conf[i];
    int le = 0;
    for (int j = 0; j < L; ++j) le += (s_conf[j] <= ci) ? 1 : 0;
    const long long sid = sampled_out[static_cast<long long>(b) * L + i];
    next_ids[static_cast<long long>(b) * L + i] = (le <= k) ? mask_id : sid;
  }
}

}  // namespace

int sample_step(const void* logits, const void* logits_unc, long long row_stride, long long batch_stride,
                float guidance, const long long* input_ids, const float* q_exp, const float* u,
                long long* sampled, long long* next_ids, float* conf_out, int B, int L, int K, long long mask_id,
                int mask_len, float temperature, cudaStream_t s) {
  if (B <= 0 || L <= 0) return MUSE_OK;
  if (L > 4096) { set_last_error("sample_step: L=%d too long (max 4096)", L); return MUSE_ERR_UNSUPPORTED; }
  sample_step_kernel<<<B, 1024, L * sizeof(float), s>>>(reinterpret_cast<const bf16*>(logits),
                                                        reinterpret_cast<const bf16*>(logits_unc), row_stride,
                                                        batch_stride, guidance, input_ids, q_exp, u, sampled, next_ids,
                                                        conf_out, L, K, mask_id, mask_len, temperature);
  return check_launch("sample_step");
}

}  // namespace muse
