// HBM-bound elementwise / gather kernels of the transformer step:
//   * multi-tensor fp32 -> bf16 weight packing (one launch for all Linear weights of a model)
//   * token + position embedding gather (muse/modeling_transformer.py:942-957) and its backward
//   * GLU: gelu(a) * b (muse/modeling_transformer.py:789-792) forward / backward
//   * plain fp32 -> bf16 cast (gradient of the residual stream entering a dgrad GEMM)
#include "common.cuh"

namespace muse {
namespace {

// ---------------------------------------------------------------- multi-tensor pack
// table[i] = {src fp32 ptr, dst bf16 ptr, numel(multiple of 4), first 1024-element block index}
struct PackEntry {
  const float* src;
  bf16* dst;
  long long numel;
  long long first_block;
};

__global__ void __launch_bounds__(256) pack_bf16_kernel(const PackEntry* __restrict__ table, int n_entries) {
  pdl_enter();
  const long long blk = blockIdx.x;
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {  // last entry with first_block <= blk
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const PackEntry e = table[lo];
  const long long i = ((blk - e.first_block) * 256 + threadIdx.x) * 4;
  if (i < e.numel) {
    const float4 v = *reinterpret_cast<const float4*>(e.src + i);
    uint2 u;
    u.x = pack_bf16(v.x, v.y);
    u.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(e.dst + i) = u;
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n8) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n8) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
}

// ---------------------------------------------------------------- embedding
__global__ void __launch_bounds__(128)
embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                 float* __restrict__ out, int tokens, int S, int H, int vocab) {
  pdl_enter();
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= tokens) return;
  const int lane = threadIdx.x & 31;
  long long id = ids[t];
  if (id < 0 || id >= vocab) id = 0;  // torch raises on OOB ids; host validates, device stays in-bounds
  const float* wr = word + id * H;
  const float* pr = pos ? pos + static_cast<size_t>(t % S) * H : nullptr;  // pos == null: plain gather (ConvEmbed of v2)
  float* o = out + static_cast<size_t>(t) * H;
  for (int c = lane * 4; c < H; c += 128) {
    const float4 a = *reinterpret_cast<const float4*>(wr + c);
    const float4 b = pr ? *reinterpret_cast<const float4*>(pr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(o + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// dword[ids[t]] += dx[t].  Roughly half of the tokens of a training batch are the mask token (id = vocab - 1 in every
// reference config, modeling_transformer.py:1128), so plain atomics serialise T/2 adds on one row.  Each warp therefore
// keeps the hot row's contribution of its 8 tokens in registers, the CTA folds its 8 warps in shared memory, and one
// atomic per column and CTA reaches global memory; all other ids (spread over ~1000 rows) use vector atomics directly.
constexpr int kEmbTokPerWarp = 8;
template <int CH>  // CH float4 chunks per lane: H = CH * 128
__global__ void __launch_bounds__(256)
embed_bwd_word_kernel(const long long* __restrict__ ids, const float* __restrict__ dx, float* __restrict__ dword,
                      int tokens, int H, int vocab, int hot_id) {
  pdl_enter();
  __shared__ float s_hot[CH * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < CH * 128; i += 256) s_hot[i] = 0.f;
  __syncthreads();
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool any_hot = false;
  const int t0 = (blockIdx.x * 8 + warp) * kEmbTokPerWarp;
  for (int k = 0; k < kEmbTokPerWarp; ++k) {
    const int t = t0 + k;
    if (t >= tokens) break;
    const long long id = ids[t];
    if (id < 0 || id >= vocab) continue;
    const float* g = dx + static_cast<size_t>(t) * H;
    if (id == hot_id) {
      any_hot = true;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(g + c * 128 + lane * 4);
        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
      }
    } else {
      float* d = dword + id * H;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        atomicAdd(reinterpret_cast<float4*>(d + c * 128 + lane * 4), *reinterpret_cast<const float4*>(g + c * 128 + lane * 4));
    }
  }
  if (any_hot) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float* h = s_hot + c * 128 + lane * 4;
      atomicAdd(h + 0, acc[c].x); atomicAdd(h + 1, acc[c].y); atomicAdd(h + 2, acc[c].z); atomicAdd(h + 3, acc[c].w);
    }
  }
  __syncthreads();
  if (hot_id >= 0 && hot_id < vocab) {
    float* d = dword + static_cast<size_t>(hot_id) * H;
    for (int i = threadIdx.x; i < CH * 128; i += 256) {
      const float v = s_hot[i];
      if (v != 0.f) atomicAdd(d + i, v);
    }
  }
}

// generic H (any multiple of 4): one warp per token, vector atomics
__global__ void __launch_bounds__(128)
embed_bwd_word_generic_kernel(const long long* __restrict__ ids, const float* __restrict__ dx, float* __restrict__ dword,
                              int tokens, int H, int vocab) {
  pdl_enter();
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= tokens) return;
  const int lane = threadIdx.x & 31;
  long long id = ids[t];
  if (id < 0 || id >= vocab) return;
  const float* g = dx + static_cast<size_t>(t) * H;
  float* d = dword + id * H;
  for (int c = lane * 4; c < H; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(g + c);
    atomicAdd(reinterpret_cast<float4*>(d + c), v);
  }
}

// dpos[s] += sum_b dx[b, s]; block (s, column chunk of 128 floats), threads stride over batch
__global__ void __launch_bounds__(128)
embed_bwd_pos_kernel(const float* __restrict__ dx, float* __restrict__ dpos, int B, int S, int H) {
  pdl_enter();
  const int s = blockIdx.x;
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= H) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += dx[(static_cast<size_t>(b) * S + s) * H + c];
  dpos[static_cast<size_t>(s) * H + c] = acc;  // fixed summation order, plain store: reproducible, no zero fill
}

// ---------------------------------------------------------------- reproducible word-embedding gradient
// dword[v] = sum of dx[t] over the tokens t with ids[t] == v, summed in ascending t whatever the launch order:
// the host passes the stable sort permutation of the ids (`order`) and the segment bounds (`bounds[v]` = first sorted
// position of id v); segments are cut into chunks of kSegChunk tokens, every chunk is summed by one CTA (four thread
// groups take a quarter each, rows in order, then the quarters are added in order), single-chunk segments are stored
// straight to dword[v], longer ones go through per-chunk partial rows that a second kernel adds in chunk order.
constexpr int kSegChunk = 256;

__global__ void __launch_bounds__(1024)
embed_seg_plan_kernel(const long long* __restrict__ bounds, int* __restrict__ chunk_off, int vocab) {
  pdl_enter();
  __shared__ int s_part[1024];
  const int per = ceil_div(vocab, 1024);
  const int v0 = threadIdx.x * per;
  int sum = 0;
  for (int v = v0; v < min(vocab, v0 + per); ++v)
    sum += static_cast<int>((bounds[v + 1] - bounds[v] + kSegChunk - 1) / kSegChunk);
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) { const int t = s_part[i]; s_part[i] = run; run += t; }
    chunk_off[vocab] = run;
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int v = v0; v < min(vocab, v0 + per); ++v) {
    chunk_off[v] = run;
    run += static_cast<int>((bounds[v + 1] - bounds[v] + kSegChunk - 1) / kSegChunk);
  }
}

// ordered sum of `n` rows of H floats: row(i) gives the row pointer; 4 groups x ncol threads, result valid in group 0
template <typename RowFn>
__device__ __forceinline__ void ordered_rows_sum(RowFn row, int n, int H, float* s_q, float* out_row) {
  const int ncol = blockDim.x >> 2;
  const int grp = threadIdx.x / ncol, tx = threadIdx.x % ncol;
  const int per = (n + 3) >> 2;
  const int i0 = grp * per, i1 = min(n, i0 + per);
  for (int c = tx * 4; c < H; c += ncol * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = i0; i < i1; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(row(i) + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(s_q + static_cast<size_t>(grp) * H + c) = acc;
  }
  __syncthreads();
  if (grp == 0) {
    for (int c = tx * 4; c < H; c += ncol * 4) {
      float4 a = *reinterpret_cast<const float4*>(s_q + c);
#pragma unroll
      for (int g = 1; g < 4; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(s_q + static_cast<size_t>(g) * H + c);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      *reinterpret_cast<float4*>(out_row + c) = a;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) embed_seg_chunk_kernel(const long long* __restrict__ order, const long long* __restrict__ bounds,
                                       const int* __restrict__ chunk_off, const float* __restrict__ dx,
                                       float* __restrict__ dword, float* __restrict__ ws, int H, int vocab) {
  pdl_enter();
  extern __shared__ __align__(16) float s_q[];  // [4][H]
  const int g = blockIdx.x;
  if (g >= chunk_off[vocab]) return;
  int lo = 0, hi = vocab - 1;
  while (lo < hi) {  // last v with chunk_off[v] <= g (empty segments share their successor's offset and lose the tie)
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_off[mid] <= g) lo = mid; else hi = mid - 1;
  }
  const int v = lo;
  const int j = g - chunk_off[v];
  const long long p0 = bounds[v] + static_cast<long long>(j) * kSegChunk;
  const int n = static_cast<int>(min(static_cast<long long>(kSegChunk), bounds[v + 1] - p0));
  const bool single = (chunk_off[v + 1] - chunk_off[v]) == 1;
  float* out = single ? dword + static_cast<size_t>(v) * H : ws + static_cast<size_t>(g) * H;
  ordered_rows_sum([&](int i) { return dx + static_cast<size_t>(order[p0 + i]) * H; }, n, H, s_q, out);
}

__global__ void __launch_bounds__(512) embed_seg_final_kernel(const int* __restrict__ chunk_off, const float* __restrict__ ws,
                                       float* __restrict__ dword, int H, int vocab) {
  pdl_enter();
  extern __shared__ __align__(16) float s_q[];
  const int v = blockIdx.x;
  const int c0 = chunk_off[v], nc = chunk_off[v + 1] - c0;
  if (nc == 1) return;  // stored by the chunk kernel
  float* out = dword + static_cast<size_t>(v) * H;
  if (nc == 0) {
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  ordered_rows_sum([&](int i) { return ws + static_cast<size_t>(c0 + i) * H; }, nc, H, s_q, out);
}

// ---------------------------------------------------------------- GLU
// ab: [rows, 2*I] bf16 (a = cols [0,I), b = cols [I,2I)); out [rows, I] bf16.
// Rounding points follow the reference under bf16 autocast: gelu(a) is rounded to bf16 before the product.
__global__ void __launch_bounds__(256)
glu_fwd_kernel(const bf16* __restrict__ ab, bf16* __restrict__ out, long long rows, int I) {
  pdl_enter();
  const int chunks = I / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= rows * chunks) return;
  const long long r = idx / chunks;
  const int c = static_cast<int>(idx % chunks) * 8;
  float a[8], b[8], o[8];
  load8(ab + r * 2 * I + c, a);
  load8(ab + r * 2 * I + I + c, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = bf16_round(gelu_f(a[j])) * b[j];
  store8(out + r * I + c, o);
}

__global__ void __launch_bounds__(256)
glu_bwd_kernel(const bf16* __restrict__ ab, const bf16* __restrict__ dout, bf16* __restrict__ dab, long long rows, int I) {
  pdl_enter();
  const int chunks = I / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= rows * chunks) return;
  const long long r = idx / chunks;
  const int c = static_cast<int>(idx % chunks) * 8;
  float a[8], b[8], d[8], da[8], db[8];
  load8(ab + r * 2 * I + c, a);
  load8(ab + r * 2 * I + I + c, b);
  load8(dout + r * I + c, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float gv, gg;
    gelu_eval(a[j], gv, gg);
    da[j] = d[j] * b[j] * gg;
    db[j] = d[j] * bf16_round(gv);
  }
  store8(dab + r * 2 * I + c, da);
  store8(dab + r * 2 * I + I + c, db);
}

}  // namespace

int pack_bf16(const void* table_dev, int n_entries, long long total_blocks, cudaStream_t s) {
  if (n_entries <= 0 || total_blocks <= 0) return MUSE_OK;
  pdl_launch(static_cast<unsigned>(total_blocks), 256, 0, s)(pack_bf16_kernel, reinterpret_cast<const PackEntry*>(table_dev), n_entries);
  return check_launch("pack_bf16");
}

int cast_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
  if (n <= 0) return MUSE_OK;
  if (n % 8 != 0) { set_last_error("cast_bf16: n=%lld must be a multiple of 8", n); return MUSE_ERR_INVALID; }
  const long long n8 = n / 8;
  pdl_launch(static_cast<unsigned>(ceil_div_ll(n8, 256)), 256, 0, s)(cast_bf16_kernel, src, reinterpret_cast<bf16*>(dst), n8);
  return check_launch("cast_bf16");
}

int embed_fwd(const long long* ids, const float* word, const float* pos, float* out, int B, int S, int H, int vocab,
              cudaStream_t s) {
  if (H % 4 != 0) { set_last_error("embed_fwd: H must be a multiple of 4"); return MUSE_ERR_INVALID; }
  const int tokens = B * S;
  if (tokens <= 0) return MUSE_OK;
  pdl_launch(ceil_div(tokens, 4), 128, 0, s)(embed_fwd_kernel, ids, word, pos, out, tokens, S, H, vocab);
  return check_launch("embed_fwd");
}

int embed_bwd(const long long* ids, const float* dx, float* dword, float* dpos, int B, int S, int H, int vocab,
              cudaStream_t s) {
  if (H % 4 != 0) { set_last_error("embed_bwd: H must be a multiple of 4"); return MUSE_ERR_INVALID; }
  const int tokens = B * S;
  if (tokens <= 0) return MUSE_OK;
  const int hot = vocab - 1;  // the reference's mask_token_id
  const int grid = ceil_div(tokens, 8 * kEmbTokPerWarp);
  if (H == 512) pdl_launch(grid, 256, 0, s)(embed_bwd_word_kernel<4>, ids, dx, dword, tokens, H, vocab, hot);
  else if (H == 768) pdl_launch(grid, 256, 0, s)(embed_bwd_word_kernel<6>, ids, dx, dword, tokens, H, vocab, hot);
  else if (H == 1024) pdl_launch(grid, 256, 0, s)(embed_bwd_word_kernel<8>, ids, dx, dword, tokens, H, vocab, hot);
  else if (H == 128) pdl_launch(grid, 256, 0, s)(embed_bwd_word_kernel<1>, ids, dx, dword, tokens, H, vocab, hot);
  else pdl_launch(ceil_div(tokens, 4), 128, 0, s)(embed_bwd_word_generic_kernel, ids, dx, dword, tokens, H, vocab);
  int rc = check_launch("embed_bwd_word");
  if (rc) return rc;
  if (dpos == nullptr) return MUSE_OK;  // no position table (ConvEmbed of MaskGiTUViT_v2)
  pdl_launch(dim3(S, ceil_div(H, 128)), 128, 0, s)(embed_bwd_pos_kernel, dx, dpos, B, S, H);
  return check_launch("embed_bwd_pos");
}

// workspace: (tokens / kSegChunk + vocab) partial rows of H floats, then vocab + 1 ints of chunk offsets
long long embed_bwd_sorted_workspace_bytes(int tokens, int H, int vocab) {
  return (static_cast<long long>(tokens) / kSegChunk + vocab) * H * 4 + (static_cast<long long>(vocab) + 1) * 4;
}

int embed_bwd_sorted(const long long* order, const long long* bounds, const float* dx, float* dword, float* dpos,
                     void* ws, int B, int S, int H, int vocab, cudaStream_t s) {
  if (H % 4 != 0) { set_last_error("embed_bwd_sorted: H must be a multiple of 4"); return MUSE_ERR_INVALID; }
  const int tokens = B * S;
  if (tokens <= 0 || vocab <= 0) return MUSE_OK;
  const int max_chunks = tokens / kSegChunk + vocab;
  float* partial = reinterpret_cast<float*>(ws);
  int* chunk_off = reinterpret_cast<int*>(partial + static_cast<size_t>(max_chunks) * H);
  int ncol = H / 4;
  if (ncol > 128) ncol = 128;  // 4 groups x ncol threads <= 512 (the kernels' launch bound)
  const size_t smem = static_cast<size_t>(4) * H * sizeof(float);
  if (smem > 48 * 1024) { set_last_error("embed_bwd_sorted: H=%d too wide", H); return MUSE_ERR_UNSUPPORTED; }
  pdl_launch(1, 1024, 0, s)(embed_seg_plan_kernel, bounds, chunk_off, vocab);
  pdl_launch(max_chunks, 4 * ncol, smem, s)(embed_seg_chunk_kernel, order, bounds, chunk_off, dx, dword, partial, H, vocab);
  pdl_launch(vocab, 4 * ncol, smem, s)(embed_seg_final_kernel, chunk_off, partial, dword, H, vocab);
  int rc = check_launch("embed_bwd_sorted");
  if (rc) return rc;
  if (dpos == nullptr) return MUSE_OK;
  pdl_launch(dim3(S, ceil_div(H, 128)), 128, 0, s)(embed_bwd_pos_kernel, dx, dpos, B, S, H);
  return check_launch("embed_bwd_pos");
}

int glu_fwd(const void* ab, void* out, long long rows, int I, cudaStream_t s) {
  if (I % 8 != 0) { set_last_error("glu_fwd: I must be a multiple of 8"); return MUSE_ERR_INVALID; }
  const long long n = rows * (I / 8);
  if (n <= 0) return MUSE_OK;
  pdl_launch(static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s)(glu_fwd_kernel, reinterpret_cast<const bf16*>(ab), reinterpret_cast<bf16*>(out), rows, I);
  return check_launch("glu_fwd");
}

int glu_bwd(const void* ab, const void* dout, void* dab, long long rows, int I, cudaStream_t s) {
  if (I % 8 != 0) { set_last_error("glu_bwd: I must be a multiple of 8"); return MUSE_ERR_INVALID; }
  const long long n = rows * (I / 8);
  if (n <= 0) return MUSE_OK;
  pdl_launch(static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, s)(glu_bwd_kernel, reinterpret_cast<const bf16*>(ab), reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dab), rows, I);
  return check_launch("glu_bwd");
}

}  // namespace muse
