// Row normalisation (LayerNorm without bias / RMSNorm), forward and backward, HBM-bound.
// Reference: muse/modeling_transformer.py:124-137 (LayerNorm -> F.layer_norm, weight only),
// :79-100 (RMSNorm). One warp owns one row; the row is held in registers between the statistics
// pass and the normalise pass so every element is read from HBM exactly once and written once.
// Fusions folded in (they are separate ATen launches in the reference):
//   * act=1     : the input is passed through exact GELU first (mlm_dense -> gelu -> mlm_ln, :980-983)
//   * residual  : y = residual + norm(x) (normformer post-attention norm + residual add, :882-884)
//   * backward  : dx = norm_bwd(dy) (+ dres_in), so the residual-stream gradient add is free.
#include "common.cuh"

namespace muse {
namespace {

constexpr int kWarpsPerBlock = 4;

template <typename TX, typename TY, int CH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
norm_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ res,
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
norm_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ w,
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

template <typename TX, typename TY>
int fwd_dispatch(const void* x, const float* w, const float* res, void* y, float* mean, float* rstd, int rows, int H,
                 float eps, int act, int rms, cudaStream_t s) {
  const int grid = ceil_div(rows, kWarpsPerBlock);
  const int ch = ceil_div(H, 256);
#define MUSE_NF(CH)                                                                                           \
  norm_fwd_kernel<TX, TY, CH><<<grid, kWarpsPerBlock * 32, 0, s>>>(reinterpret_cast<const TX*>(x), w, res,    \
                                                                    reinterpret_cast<TY*>(y), mean, rstd, rows, \
                                                                    H, eps, act, rms)
  if (ch <= 1) MUSE_NF(1);
  else if (ch <= 2) MUSE_NF(2);
  else if (ch <= 4) MUSE_NF(4);
  else if (ch <= 8) MUSE_NF(8);
  else MUSE_NF(16);
#undef MUSE_NF
  return check_launch("norm_fwd");
}

template <typename TDY, typename TX, typename TDX>
int bwd_dispatch(const void* dy, const void* x, const float* w, const float* mean, const float* rstd,
                 const float* dres, void* dx, float* dw, int rows, int H, int act, int rms, cudaStream_t s) {
  int grid = ceil_div(rows, kWarpsPerBlock);
  if (grid > 148 * 8) grid = 148 * 8;
  const int ch = ceil_div(H, 256);
  const size_t smem = dw ? H * sizeof(float) : 0;
#define MUSE_NB(CH)                                                                                          \
  norm_bwd_kernel<TDY, TX, TDX, CH><<<grid, kWarpsPerBlock * 32, smem, s>>>(                                 \
      reinterpret_cast<const TDY*>(dy), reinterpret_cast<const TX*>(x), w, mean, rstd, dres,                 \
      reinterpret_cast<TDX*>(dx), dw, rows, H, act, rms)
  if (ch <= 1) MUSE_NB(1);
  else if (ch <= 2) MUSE_NB(2);
  else if (ch <= 4) MUSE_NB(4);
  else if (ch <= 8) MUSE_NB(8);
  else MUSE_NB(16);
#undef MUSE_NB
  return check_launch("norm_bwd");
}

}  // namespace

// dtype codes: 0 = fp32, 1 = bf16
int norm_fwd(const void* x, int x_dt, const float* w, const float* res, void* y, int y_dt, float* mean, float* rstd,
             int rows, int H, float eps, int act, int rms, cudaStream_t s) {
  if (rows <= 0) return MUSE_OK;
  if (H % 8 != 0 || H > 4096) {
    set_last_error("norm_fwd: H=%d must be a multiple of 8 and <= 4096", H);
    return MUSE_ERR_UNSUPPORTED;
  }
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
  if (H % 8 != 0 || H > 4096) {
    set_last_error("norm_bwd: H=%d must be a multiple of 8 and <= 4096", H);
    return MUSE_ERR_UNSUPPORTED;
  }
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
