// Row normalisation (LayerNorm without bias / RMSNorm), forward and backward, HBM-bound.
// Reference: muse/modeling_transformer.py:124-137 (LayerNorm -> F.layer_norm, weight only),
// :79-100 (RMSNorm).  Each thread owns exactly one 8-element chunk of a row (16 B of bf16 / 32 B of fp32), a row is
// spread over H/8 threads (8..512), several rows share a CTA when rows are short; every element is read from HBM once
// and written once, row statistics are reduced with shuffles (+ one smem hop when a row spans several warps).
// Fusions folded in (they are separate ATen launches in the reference):
//   * act=1  : exact GELU applied to the input first (mlm_dense -> gelu -> mlm_ln, :980-983)
//   * act=2  : GLU: the input is [rows, 2H] = [a | b] and the normalised value is bf16(gelu(a)) * b, i.e. the
//              FeedForward product (:789-792) feeding mid_mlp_layer_norm (:795-796) without materialising it;
//              backward recomputes it and emits d[a | b] directly (LN backward + GLU backward in one pass)
//   * residual: y = residual + norm(x) (normformer post-attention norm + residual add, :882-884)
//   * backward: dx = norm_bwd(dy) (+ dres), so the residual-stream gradient add is free.
#include "common.cuh"

namespace muse {
namespace {

constexpr int ACT_NONE = 0, ACT_GELU = 1, ACT_GLU = 2;

// ---- warp-per-row kernels (H <= 1024, no GLU): the whole row lives in registers, no block barriers
constexpr int kWarpsPerBlock = 4;

template <typename TX, typename TY, int CH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
norm_fwd_warp_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ res,
                TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
                float eps, int act, int rms) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + static_cast<size_t>(row) * H;
  float v[CH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < H) {
      load8(xr + col, v[c]);
      if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[c][j] = gelu_f(v[c][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[c][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
  const float inv_h = 1.0f / static_cast<float>(H);
  float mean = rms ? 0.f : warp_sum(sum) * inv_h;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < H) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        sq += d * d;
      }
    }
  }
  const float var = warp_sum(sq) * inv_h;
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  TY* yr = y + static_cast<size_t>(row) * H;
  const float* rr = res ? res + static_cast<size_t>(row) * H : nullptr;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < H) {
      float o[8], wv[8];
      if (w) load8(w + col, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * (w ? wv[j] : 1.f);
      if (rr) {
        float r8[8];
        load8(rr + col, r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r8[j];
      }
      store8(yr + col, o);
    }
  }
}

template <typename TDY, typename TX, typename TDX, int CH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
norm_bwd_warp_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                const float* __restrict__ dres, TDX* __restrict__ dx, float* __restrict__ dw, int rows, int H,
                int act, int rms) {
  extern __shared__ float s_dw[];  // [H]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (dw) {
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_dw[i] = 0.f;
    __syncthreads();
  }
  float dw_acc[CH][8];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) dw_acc[c][j] = 0.f;
  const float inv_h = 1.0f / static_cast<float>(H);

  for (int row = blockIdx.x * kWarpsPerBlock + warp; row < rows; row += gridDim.x * kWarpsPerBlock) {
    const TX* xr = x + static_cast<size_t>(row) * H;
    const TDY* dyr = dy + static_cast<size_t>(row) * H;
    const float mean = rms ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    float xh[CH][8], g[CH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < H) {
        float xv[8], dv[8], wv[8];
        load8(xr + col, xv);
        load8(dyr + col, dv);
        if (w) load8(w + col, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = act ? gelu_f(xv[j]) : xv[j];
          xh[c][j] = (a - mean) * rstd;
          dw_acc[c][j] += dv[j] * xh[c][j];
          g[c][j] = dv[j] * (w ? wv[j] : 1.f);
          s1 += g[c][j];
          s2 += g[c][j] * xh[c][j];
        }
      }
    }
    s1 = rms ? 0.f : warp_sum(s1) * inv_h;
    s2 = warp_sum(s2) * inv_h;
    TDX* dxr = dx + static_cast<size_t>(row) * H;
    const float* drr = dres ? dres + static_cast<size_t>(row) * H : nullptr;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < H) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[c][j] - s1 - xh[c][j] * s2);
        if (act) {
          float xv[8];
          load8(xr + col, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= gelu_grad_f(xv[j]);
        }
        if (drr) {
          float r8[8];
          load8(drr + col, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r8[j];
        }
        store8(dxr + col, o);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < H) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&s_dw[col + j], dw_acc[c][j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(&dw[i], s_dw[i]);
  }
}


// ---- CTA-per-row(-group) kernels (H > 1024 or GLU mode)
struct RowMap {
  int tpr;        // threads per row (power of two, >= 8)
  int rpb;        // rows per block
  int threads;    // block size
};

RowMap row_map(int H) {
  int chunks = ceil_div(H, 8);
  int tpr = 8;
  while (tpr < chunks) tpr <<= 1;
  RowMap m;
  m.tpr = tpr;
  m.threads = tpr > 256 ? tpr : 256;
  m.rpb = m.threads / tpr;
  return m;
}

// Sum `v` over the tpr threads of a row. s_red: [rows_per_block][16] scratch; all threads of the block must call.
template <int NV>
__device__ __forceinline__ void row_sum(float (&v)[NV], int tpr, float* s_red, int row_in_block, int tid_in_row) {
  const int w = tpr < 32 ? tpr : 32;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    for (int o = w >> 1; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  if (tpr > 32) {
    const int nw = tpr >> 5, wi = tid_in_row >> 5;
    __syncthreads();  // previous use of s_red finished
    if ((tid_in_row & 31) == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) s_red[(row_in_block * 16 + wi) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = 0.f;
      for (int k = 0; k < nw; ++k) a += s_red[(row_in_block * 16 + k) * NV + i];
      v[i] = a;
    }
  }
}

// v = value to normalise; gv8 = bf16(gelu(a)) and ga8 = gelu'(a) (GLU / GELU modes), b8 = the linear GLU half
template <typename TX>
__device__ __forceinline__ void load_value(const TX* xr, int col, int H, int act, float (&v)[8], float (&gv8)[8],
                                           float (&ga8)[8], float (&b8)[8]) {
  if (act == ACT_GLU) {
    float a8[8];
    load8(xr + col, a8);
    load8(xr + H + col, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gv;
      gelu_eval(a8[j], gv, ga8[j]);
      gv8[j] = bf16_round(gv);
      v[j] = bf16_round(gv8[j] * b8[j]);
    }
  } else {
    load8(xr + col, v);
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gv;
        gelu_eval(v[j], gv, ga8[j]);
        v[j] = gv;
      }
    }
  }
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(512)
norm_fwd_block_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ res,
                TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
                float eps, int act, int rms, int tpr) {
  __shared__ float s_red[16 * 16 * 2];
  const int rpb = blockDim.x / tpr;
  const int rib = threadIdx.x / tpr, tir = threadIdx.x % tpr;
  const int row = blockIdx.x * rpb + rib;
  const int col = tir * 8;
  const bool active = (row < rows) && (col < H);
  const int xs = (act == ACT_GLU) ? 2 * H : H;
  float v[8], gv8[8], ga8[8], b8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (active) load_value(x + static_cast<size_t>(row) * xs, col, H, act, v, gv8, ga8, b8);
  const float inv_h = 1.0f / static_cast<float>(H);
  float s[1] = {0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) s[0] += v[j];
  float mean = 0.f;
  if (!rms) {
    row_sum<1>(s, tpr, s_red, rib, tir);
    mean = s[0] * inv_h;
  }
  float q[1] = {0.f};
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q[0] += d * d; }
  }
  row_sum<1>(q, tpr, s_red, rib, tir);
  const float rstd = rsqrtf(q[0] * inv_h + eps);
  if (active) {
    if (tir == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    float o[8], wv[8];
    if (w) load8(w + col, wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * (w ? wv[j] : 1.f);
    if (res) {
      float r8[8];
      load8(res + static_cast<size_t>(row) * H + col, r8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += r8[j];
    }
    store8(y + static_cast<size_t>(row) * H + col, o);
  }
}

template <typename TDY, typename TX, typename TDX>
__global__ void __launch_bounds__(512)
norm_bwd_block_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                const float* __restrict__ dres, TDX* __restrict__ dx, float* __restrict__ dw, int rows, int H,
                int act, int rms, int tpr) {
  __shared__ float s_red[16 * 16 * 2];
  extern __shared__ float s_dw[];  // [H] (only when several rows share the block)
  const int rpb = blockDim.x / tpr;
  const int rib = threadIdx.x / tpr, tir = threadIdx.x % tpr;
  const int col = tir * 8;
  const bool col_ok = col < H;
  const int xs = (act == ACT_GLU) ? 2 * H : H;
  const float inv_h = 1.0f / static_cast<float>(H);
  float wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wv[j] = 1.f;
  if (w && col_ok) load8(w + col, wv);
  float dw_acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dw_acc[j] = 0.f;

  const int nbatches = ceil_div(rows, rpb);
  for (int batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int row = batch * rpb + rib;
    const bool active = (row < rows) && col_ok;
    float v[8], gv8[8], ga8[8], b8[8], dv[8], xh[8], g[8];
    float s[2] = {0.f, 0.f};
    float mean = 0.f, rstd = 0.f;
    if (active) {
      load_value(x + static_cast<size_t>(row) * xs, col, H, act, v, gv8, ga8, b8);
      load8(dy + static_cast<size_t>(row) * H + col, dv);
      mean = rms ? 0.f : mean_in[row];
      rstd = rstd_in[row];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[j] = (v[j] - mean) * rstd;
        dw_acc[j] += dv[j] * xh[j];
        g[j] = dv[j] * wv[j];
        s[0] += g[j];
        s[1] += g[j] * xh[j];
      }
    }
    row_sum<2>(s, tpr, s_red, rib, tir);
    const float s1 = rms ? 0.f : s[0] * inv_h;
    const float s2 = s[1] * inv_h;
    if (active) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] - s1 - xh[j] * s2);
      if (act == ACT_GLU) {
        float da[8], db[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          da[j] = o[j] * b8[j] * ga8[j];
          db[j] = o[j] * gv8[j];
        }
        store8(dx + static_cast<size_t>(row) * xs + col, da);
        store8(dx + static_cast<size_t>(row) * xs + H + col, db);
      } else {
        if (act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= ga8[j];
        }
        if (dres) {
          float r8[8];
          load8(dres + static_cast<size_t>(row) * H + col, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r8[j];
        }
        store8(dx + static_cast<size_t>(row) * H + col, o);
      }
    }
  }
  if (dw) {
    if (rpb > 1) {
      for (int i = threadIdx.x; i < H; i += blockDim.x) s_dw[i] = 0.f;
      __syncthreads();
      if (col_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&s_dw[col + j], dw_acc[j]);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(&dw[i], s_dw[i]);
    } else if (col_ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&dw[col + j], dw_acc[j]);
    }
  }
}

template <typename TX, typename TY>
int fwd_dispatch(const void* x, const float* w, const float* res, void* y, float* mean, float* rstd, int rows, int H,
                 float eps, int act, int rms, cudaStream_t s) {
  if (H <= 1024 && act != ACT_GLU) {
    const int grid = ceil_div(rows, kWarpsPerBlock);
    const int ch = ceil_div(H, 256);
#define MUSE_NF(CH)                                                                                                \
  norm_fwd_warp_kernel<TX, TY, CH><<<grid, kWarpsPerBlock * 32, 0, s>>>(reinterpret_cast<const TX*>(x), w, res,    \
                                                                         reinterpret_cast<TY*>(y), mean, rstd, rows, \
                                                                         H, eps, act, rms)
    if (ch <= 1) MUSE_NF(1);
    else if (ch <= 2) MUSE_NF(2);
    else MUSE_NF(4);
#undef MUSE_NF
    return check_launch("norm_fwd");
  }
  const RowMap m = row_map(H);
  norm_fwd_block_kernel<TX, TY><<<ceil_div(rows, m.rpb), m.threads, 0, s>>>(
      reinterpret_cast<const TX*>(x), w, res, reinterpret_cast<TY*>(y), mean, rstd, rows, H, eps, act, rms, m.tpr);
  return check_launch("norm_fwd");
}

template <typename TDY, typename TX, typename TDX>
int bwd_dispatch(const void* dy, const void* x, const float* w, const float* mean, const float* rstd,
                 const float* dres, void* dx, float* dw, int rows, int H, int act, int rms, cudaStream_t s) {
  if (H <= 1024 && act != ACT_GLU) {
    int grid = ceil_div(rows, kWarpsPerBlock);
    if (grid > 148 * 8) grid = 148 * 8;
    const int ch = ceil_div(H, 256);
    const size_t smem = dw ? H * sizeof(float) : 0;
#define MUSE_NB(CH)                                                                                          \
  norm_bwd_warp_kernel<TDY, TX, TDX, CH><<<grid, kWarpsPerBlock * 32, smem, s>>>(                            \
      reinterpret_cast<const TDY*>(dy), reinterpret_cast<const TX*>(x), w, mean, rstd, dres,                 \
      reinterpret_cast<TDX*>(dx), dw, rows, H, act, rms)
    if (ch <= 1) MUSE_NB(1);
    else if (ch <= 2) MUSE_NB(2);
    else MUSE_NB(4);
#undef MUSE_NB
    return check_launch("norm_bwd");
  }
  const RowMap m = row_map(H);
  int grid = ceil_div(rows, m.rpb);
  const int cap = 148 * (m.threads > 256 ? 2 : 6);
  if (grid > cap) grid = cap;
  const size_t smem = (dw && m.rpb > 1) ? H * sizeof(float) : 0;
  norm_bwd_block_kernel<TDY, TX, TDX><<<grid, m.threads, smem, s>>>(
      reinterpret_cast<const TDY*>(dy), reinterpret_cast<const TX*>(x), w, mean, rstd, dres,
      reinterpret_cast<TDX*>(dx), dw, rows, H, act, rms, m.tpr);
  return check_launch("norm_bwd");
}

int check_args(const char* who, int H, int act, const void* res_or_dres) {
  if (H % 8 != 0 || H > 4096 || H < 8) {
    set_last_error("%s: H=%d must be a multiple of 8 in [8, 4096]", who, H);
    return MUSE_ERR_UNSUPPORTED;
  }
  if (act < 0 || act > 2) { set_last_error("%s: bad act %d", who, act); return MUSE_ERR_INVALID; }
  if (act == ACT_GLU && res_or_dres) { set_last_error("%s: GLU mode does not take a residual", who); return MUSE_ERR_INVALID; }
  return MUSE_OK;
}

}  // namespace

// dtype codes: 0 = fp32, 1 = bf16.  act: 0 none, 1 GELU(x), 2 GLU (x is [rows, 2H]).
int norm_fwd(const void* x, int x_dt, const float* w, const float* res, void* y, int y_dt, float* mean, float* rstd,
             int rows, int H, float eps, int act, int rms, cudaStream_t s) {
  if (rows <= 0) return MUSE_OK;
  int rc = check_args("norm_fwd", H, act, res);
  if (rc) return rc;
  if (x_dt == 0 && y_dt == 1) return fwd_dispatch<float, bf16>(x, w, res, y, mean, rstd, rows, H, eps, act, rms, s);
  if (x_dt == 1 && y_dt == 1) return fwd_dispatch<bf16, bf16>(x, w, res, y, mean, rstd, rows, H, eps, act, rms, s);
  if (x_dt == 1 && y_dt == 0) return fwd_dispatch<bf16, float>(x, w, res, y, mean, rstd, rows, H, eps, act, rms, s);
  if (x_dt == 0 && y_dt == 0) return fwd_dispatch<float, float>(x, w, res, y, mean, rstd, rows, H, eps, act, rms, s);
  set_last_error("norm_fwd: bad dtype codes %d %d", x_dt, y_dt);
  return MUSE_ERR_INVALID;
}

int norm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* w, const float* mean, const float* rstd,
             const float* dres, void* dx, int dx_dt, float* dw, int rows, int H, int act, int rms, cudaStream_t s) {
  if (rows <= 0) return MUSE_OK;
  int rc = check_args("norm_bwd", H, act, dres);
  if (rc) return rc;
  const int key = dy_dt * 4 + x_dt * 2 + dx_dt;
  switch (key) {
    case 0: return bwd_dispatch<float, float, float>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 1: return bwd_dispatch<float, float, bf16>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 2: return bwd_dispatch<float, bf16, float>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 3: return bwd_dispatch<float, bf16, bf16>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 4: return bwd_dispatch<bf16, float, float>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 5: return bwd_dispatch<bf16, float, bf16>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 6: return bwd_dispatch<bf16, bf16, float>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
    case 7: return bwd_dispatch<bf16, bf16, bf16>(dy, x, w, mean, rstd, dres, dx, dw, rows, H, act, rms, s);
  }
  set_last_error("norm_bwd: bad dtype codes");
  return MUSE_ERR_INVALID;
}

}  // namespace muse
