// Forward kernels specific to MaskGiTUViT_v2 (muse/modeling_transformer_v2.py); the GEMMs, attention, GLU and
// cross-entropy are the kernels of the v1 path.  All HBM-bound, token-major [rows = B*h*w, C] activations.
//   * add_norm_mod : prenorm-residual norm (unfused_rms_norm / unfused_layer_norm, :673-738): r' = a + r ; y = norm(r') * w,
//                    optionally followed by the adaLN modulation y * (1 + scale_b) + shift_b (AdaLNModulation :1025-1037)
//                    of the TransformerLayer / GLUFeedForward (:757-792, :926-951).  One pass instead of add, norm, modulate.
//   * dwconv3x3_norm: ResBlock head (:604-612): depthwise 3x3 conv (bf16-rounded like the autocast conv output) + Norm2D.
//   * grn           : GELU + GlobalResponseNorm (:741-751): Gx = ||g||_2 over (h, w) per (image, channel),
//                     Nx = Gx / (mean_c Gx + 1e-6), out = gamma * (g * Nx) + beta + g.
//   * adaln_apply   : in-place fp32 x * (1 + scale_b) + shift_b at the end of a ResBlock.
//   * silu          : SiLU of the conditioning vectors / text states feeding the adaLN and kv mappers.
#include "common.cuh"

namespace muse {
namespace {

constexpr int kWarps = 4;

// a: [rows, H] (TA = bf16 | float), r: fp32 [rows, H] or null, w: fp32 [H] or null,
// ss: fp32, scale of sample b at ss[b * ss_stride + 0 .. H), shift at ss[b * ss_stride + H .. 2H); null -> no modulation
// r_out: fp32 [rows, H] or null ; y: TY [rows, H]
template <typename TA, typename TY, int CH>
__global__ void __launch_bounds__(kWarps * 32)
add_norm_mod_kernel(const TA* __restrict__ a, const float* __restrict__ r, const float* __restrict__ w,
                    const float* __restrict__ ss, long long ss_stride, int rows_per_sample, float* __restrict__ r_out,
                    TY* __restrict__ y, int rows, int H, float eps, int rms) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * H;
  float v[CH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < H) {
      load8(a + base + col, v[c]);
      if (r != nullptr) {
        float rv[8];
        load8(r + base + col, rv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[c][j] += rv[j];
      }
      if (r_out != nullptr) store8(r_out + base + col, v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[c][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
  const float inv_h = 1.0f / static_cast<float>(H);
  const float mean = rms ? 0.f : warp_sum(sum) * inv_h;
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
  const float rstd = rsqrtf(warp_sum(sq) * inv_h + eps);
  const float* sc = ss ? ss + static_cast<size_t>(row / rows_per_sample) * ss_stride : nullptr;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < H) {
      float o[8], wv[8];
      if (w) load8(w + col, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (v[c][j] - mean) * rstd;
        if (w) o[j] *= wv[j];
      }
      if (sc) {
        float s8[8], h8[8];
        load8(sc + col, s8);
        load8(sc + H + col, h8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], 1.f + s8[j], h8[j]);
      }
      store8(y + base + col, o);
    }
  }
}

// x fp32 [B, hh, ww, C]; wk fp32 [9][C] (tap-major); nw fp32 [C] or null; y bf16 [B, hh, ww, C].  One warp per pixel.
template <int CH>
__global__ void __launch_bounds__(kWarps * 32)
dwconv3x3_norm_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ nw,
                      bf16* __restrict__ y, bf16* __restrict__ conv_out, int B, int hh, int ww, int C, float eps, int rms) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long pix = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  if (pix >= static_cast<long long>(B) * hh * ww) return;
  const int px = static_cast<int>(pix % ww);
  const int py = static_cast<int>((pix / ww) % hh);
  const long long b = pix / (static_cast<long long>(ww) * hh);
  float v[CH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    if (col < C) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int iy = py + t / 3 - 1, ix = px + t % 3 - 1;
        if (iy >= 0 && iy < hh && ix >= 0 && ix < ww) {
          float xv[8], kv[8];
          load8(x + ((b * hh + iy) * ww + ix) * C + col, xv);
          load8(wk + t * C + col, kv);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[c][j] = fmaf(xv[j], kv[j], v[c][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[c][j] = bf16_round(v[c][j]);  // the conv output is a bf16 tensor under autocast
        sum += v[c][j];
      }
      if (conv_out != nullptr) store8(conv_out + pix * C + col, v[c]);  // saved for the backward pass
    }
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  const float mean = rms ? 0.f : warp_sum(sum) * inv_c;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) * inv_c + eps);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < C) {
      float o[8], wv[8];
      if (nw) load8(nw + col, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * (nw ? wv[j] : 1.f);
      store8(y + pix * C + col, o);
    }
  }
}

// GRN pass 1: sumsq[b, c] = sum over the image's tokens of bf16(gelu(x))^2.  grid (C/256, B), 128 threads x 2 channels.
__global__ void __launch_bounds__(128)
grn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sumsq, int HW, int C) {
  pdl_enter();
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const int b = blockIdx.y;
  const bf16* p = x + static_cast<size_t>(b) * HW * C + c;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 0; t < HW; ++t) {
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(p + static_cast<size_t>(t) * C));
    const float g0 = bf16_round(gelu_f(v.x)), g1 = bf16_round(gelu_f(v.y));
    s0 = fmaf(g0, g0, s0);
    s1 = fmaf(g1, g1, s1);
  }
  sumsq[static_cast<size_t>(b) * C + c] = s0;
  sumsq[static_cast<size_t>(b) * C + c + 1] = s1;
}

// GRN pass 2 (one CTA per image): nx[c] = sqrt(sumsq[c]) / (mean_c sqrt(sumsq) + 1e-6)
__global__ void __launch_bounds__(256)
grn_finalize_kernel(const float* __restrict__ stat, float* __restrict__ nx_out, int C) {
  pdl_enter();
  __shared__ float s_part[8];
  const float* p = stat + static_cast<size_t>(blockIdx.x) * C;
  float* o = nx_out + static_cast<size_t>(blockIdx.x) * C;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) acc += sqrtf(p[c]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_part[i];
  const float denom = tot / static_cast<float>(C) + 1e-6f;
  for (int c = threadIdx.x; c < C; c += 256) o[c] = sqrtf(p[c]) / denom;
}

// GRN pass 3: out = gamma * (g * nx) + beta + g, g = bf16(gelu(x)); 8 channels per thread
__global__ void __launch_bounds__(256)
grn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ nx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, bf16* __restrict__ out, long long total8, int HW, int C) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int col = static_cast<int>(i % c8) * 8;
  const long long b = i / (static_cast<long long>(c8) * HW);
  float v[8], n8[8], g8[8], b8[8], o[8];
  load8(x + i * 8, v);
  load8(nx + b * C + col, n8);
  load8(gamma + col, g8);
  load8(beta + col, b8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g = bf16_round(gelu_f(v[j]));
    o[j] = fmaf(g8[j], g * n8[j], b8[j]) + g;
  }
  store8(out + i * 8, o);
}

__global__ void __launch_bounds__(256)
adaln_apply_kernel(float* __restrict__ x, const float* __restrict__ ss, long long ss_stride, long long total4,
                   int rows_per_sample, int C) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int col = static_cast<int>(i % c4) * 4;
  const long long b = i / (static_cast<long long>(c4) * rows_per_sample);
  const float* sc = ss + b * ss_stride;
  float4 v = *reinterpret_cast<float4*>(x + i * 4);
  const float4 s = *reinterpret_cast<const float4*>(sc + col);
  const float4 h = *reinterpret_cast<const float4*>(sc + C + col);
  v.x = fmaf(v.x, 1.f + s.x, h.x); v.y = fmaf(v.y, 1.f + s.y, h.y);
  v.z = fmaf(v.z, 1.f + s.z, h.z); v.w = fmaf(v.w, 1.f + s.w, h.w);
  *reinterpret_cast<float4*>(x + i * 4) = v;
}

template <typename TX>
__global__ void __launch_bounds__(256) silu_kernel(const TX* __restrict__ x, bf16* __restrict__ y, long long n8) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(x + i * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = v[j] / (1.f + expf(-v[j]));
  store8(y + i * 8, v);
}

template <typename TA, typename TY>
int launch_add_norm(const void* a, const float* r, const float* w, const float* ss, long long ss_stride,
                    int rows_per_sample, float* r_out, void* y, int rows, int H, float eps, int rms, cudaStream_t s) {
  const int grid = ceil_div(rows, kWarps);
  const TA* ap = reinterpret_cast<const TA*>(a);
  TY* yp = reinterpret_cast<TY*>(y);
#define MUSE_AN(CH) pdl_launch(grid, kWarps * 32, 0, s)(add_norm_mod_kernel<TA, TY, CH>, ap, r, w, ss, ss_stride, rows_per_sample, r_out, yp, rows, H, eps, rms)
  if (H <= 256) MUSE_AN(1);
  else if (H <= 512) MUSE_AN(2);
  else if (H <= 768) MUSE_AN(3);
  else MUSE_AN(4);
#undef MUSE_AN
  return check_launch("add_norm_mod");
}

}  // namespace

int add_norm_mod_fwd(const void* a, int a_dt, const float* r, const float* w, const float* ss, long long ss_stride,
                     int rows_per_sample, float* r_out, void* y, int y_dt, int rows, int H, float eps, int rms,
                     cudaStream_t s) {
  if (rows <= 0) return MUSE_OK;
  if (H % 8 != 0 || H > 1024) { set_last_error("add_norm_mod: H=%d must be a multiple of 8 and <= 1024", H); return MUSE_ERR_UNSUPPORTED; }
  if (ss != nullptr && (rows_per_sample <= 0 || ss_stride % 4 != 0)) { set_last_error("add_norm_mod: bad modulation layout"); return MUSE_ERR_INVALID; }
  if (ss == nullptr) rows_per_sample = 1;
  if (a_dt == 1 && y_dt == 1) return launch_add_norm<bf16, bf16>(a, r, w, ss, ss_stride, rows_per_sample, r_out, y, rows, H, eps, rms, s);
  if (a_dt == 0 && y_dt == 1) return launch_add_norm<float, bf16>(a, r, w, ss, ss_stride, rows_per_sample, r_out, y, rows, H, eps, rms, s);
  if (a_dt == 1 && y_dt == 0) return launch_add_norm<bf16, float>(a, r, w, ss, ss_stride, rows_per_sample, r_out, y, rows, H, eps, rms, s);
  if (a_dt == 0 && y_dt == 0) return launch_add_norm<float, float>(a, r, w, ss, ss_stride, rows_per_sample, r_out, y, rows, H, eps, rms, s);
  set_last_error("add_norm_mod: bad dtype codes %d %d", a_dt, y_dt);
  return MUSE_ERR_INVALID;
}

int dwconv3x3_norm_fwd(const float* x, const float* wk, const float* nw, void* y, void* conv_out, int B, int hh, int ww,
                       int C, float eps, int rms, cudaStream_t s) {
  const long long pixels = static_cast<long long>(B) * hh * ww;
  if (pixels <= 0) return MUSE_OK;
  if (C % 8 != 0 || C > 1024) { set_last_error("dwconv3x3_norm: C=%d must be a multiple of 8 and <= 1024", C); return MUSE_ERR_UNSUPPORTED; }
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(pixels, kWarps));
  bf16* yp = reinterpret_cast<bf16*>(y);
  bf16* cp = reinterpret_cast<bf16*>(conv_out);
  if (C <= 256) pdl_launch(grid, kWarps * 32, 0, s)(dwconv3x3_norm_kernel<1>, x, wk, nw, yp, cp, B, hh, ww, C, eps, rms);
  else if (C <= 512) pdl_launch(grid, kWarps * 32, 0, s)(dwconv3x3_norm_kernel<2>, x, wk, nw, yp, cp, B, hh, ww, C, eps, rms);
  else if (C <= 768) pdl_launch(grid, kWarps * 32, 0, s)(dwconv3x3_norm_kernel<3>, x, wk, nw, yp, cp, B, hh, ww, C, eps, rms);
  else pdl_launch(grid, kWarps * 32, 0, s)(dwconv3x3_norm_kernel<4>, x, wk, nw, yp, cp, B, hh, ww, C, eps, rms);
  return check_launch("dwconv3x3_norm");
}

// x bf16 [B, HW, C] -> out bf16 [B, HW, C]; sumsq_ws / nx_ws fp32 [B, C] (kept by the caller for the backward pass)
int grn_fwd(const void* x, const float* gamma, const float* beta, void* out, float* stat_ws, float* nx_ws, int B, int HW,
            int C, cudaStream_t s) {
  if (B <= 0 || HW <= 0) return MUSE_OK;
  if (C % 8 != 0) { set_last_error("grn: C must be a multiple of 8"); return MUSE_ERR_UNSUPPORTED; }
  const bf16* xp = reinterpret_cast<const bf16*>(x);
  pdl_launch(dim3(ceil_div(C, 256), B), 128, 0, s)(grn_stats_kernel, xp, stat_ws, HW, C);
  int rc = check_launch("grn_stats");
  if (rc) return rc;
  pdl_launch(B, 256, 0, s)(grn_finalize_kernel, stat_ws, nx_ws, C);
  rc = check_launch("grn_finalize");
  if (rc) return rc;
  const long long total8 = static_cast<long long>(B) * HW * (C / 8);
  pdl_launch(static_cast<unsigned>(ceil_div_ll(total8, 256)), 256, 0, s)(grn_apply_kernel, xp, nx_ws, gamma, beta,
                                                                                   reinterpret_cast<bf16*>(out), total8, HW, C);
  return check_launch("grn_apply");
}

int adaln_apply(float* x, const float* ss, long long ss_stride, int B, int rows_per_sample, int C, cudaStream_t s) {
  if (C % 4 != 0 || ss_stride % 4 != 0) { set_last_error("adaln_apply: C and the modulation stride must be multiples of 4"); return MUSE_ERR_UNSUPPORTED; }
  const long long total4 = static_cast<long long>(B) * rows_per_sample * (C / 4);
  if (total4 <= 0) return MUSE_OK;
  pdl_launch(static_cast<unsigned>(ceil_div_ll(total4, 256)), 256, 0, s)(adaln_apply_kernel, x, ss, ss_stride, total4, rows_per_sample, C);
  return check_launch("adaln_apply");
}

int silu_bf16(const void* x, int x_dt, void* y, long long n, cudaStream_t s) {
  if (n <= 0) return MUSE_OK;
  if (n % 8 != 0) { set_last_error("silu: n must be a multiple of 8"); return MUSE_ERR_INVALID; }
  const long long n8 = n / 8;
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(n8, 256));
  if (x_dt == 0) pdl_launch(grid, 256, 0, s)(silu_kernel<float>, reinterpret_cast<const float*>(x), reinterpret_cast<bf16*>(y), n8);
  else pdl_launch(grid, 256, 0, s)(silu_kernel<bf16>, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n8);
  return check_launch("silu");
}

}  // namespace muse
