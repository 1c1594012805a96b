// MaskGitVQGAN encoder / decoder building blocks in fp32, NHWC (muse/modeling_maskgit_vqgan.py):
//   * conv2d 'same' stride-1, k = 3 or 1 (Conv2dSame :33-45) as an implicit GEMM:
//       M = B*H*W output pixels, N = C_out, K = k*k*C_in  (weights pre-packed to [K, C_out])
//     with fused bias, fused residual add (ResnetBlock :82-85) and the nearest x2 upsample of
//     UpsamplingBlock (:146) folded into the input gather (index >> 1) so the 4x larger tensor is never written.
//   * GroupNorm(32, eps=1e-6) + SiLU (:61-79): deterministic statistics pass (per-block partials, fp64 final sum)
//     + one apply pass.
//   * avg_pool2d(2,2) (:112), NCHW <-> NHWC layout changes at the model boundary.
// fp32 SIMT on purpose: token ids must agree with the fp32 reference; each output is one ascending-K
// fma chain (taps row-major, then input channel), so results do not depend on tiling.
// This is the general-shape kernel (3-channel stem / head, odd geometries); the heavy layers run on the tensor cores
// with the same fp32-level accuracy in conv_tc.cu.
#include "common.cuh"

namespace muse {
namespace {

constexpr int KT = 8;  // k per smem stage

template <int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256)
conv2d_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias,
                   const float* __restrict__ res, float* __restrict__ y, int B, int H, int W, int Cin, int Cout,
                   int ksize, int up) {
  pdl_enter();
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ __align__(16) float sA[KT][BM];
  __shared__ __align__(16) float sB[KT][BN];
  const long long M = static_cast<long long>(B) * H * W;
  const int K = ksize * ksize * Cin;
  const long long m0 = static_cast<long long>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // up: 0 plain 'same'; 1 nearest x2 upsample folded into the gather; 2 stride 2 with pad (0,1,0,1) (taming Downsample)
  const int s2 = (up == 2);
  const int Hi = up == 1 ? H / 2 : (s2 ? 2 * H : H), Wi = up == 1 ? W / 2 : (s2 ? 2 * W : W);
  const int pad = s2 ? 0 : ksize / 2;
  const int st = s2 ? 2 : 1;

  // A-loader role: BM rows x 8 k per stage. VEC: one float4 per (row, half); scalar: BM*8/256 elements.
  constexpr int A_PER_THREAD = BM * KT / 256;  // 4 for BM=128
  int a_row[A_PER_THREAD], a_k[A_PER_THREAD];
  int ab[A_PER_THREAD], ay[A_PER_THREAD], ax[A_PER_THREAD];
  bool a_ok[A_PER_THREAD];
  if (VEC) {
    const int r = threadIdx.x / 2;  // requires BM == 128
    a_row[0] = r; a_k[0] = (threadIdx.x & 1) * 4;
  } else {
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
      const int idx = threadIdx.x + i * 256;
      a_row[i] = idx / KT; a_k[i] = idx % KT;
    }
  }
#pragma unroll
  for (int i = 0; i < (VEC ? 1 : A_PER_THREAD); ++i) {
    const long long m = m0 + a_row[i];
    a_ok[i] = m < M;
    const long long mm = a_ok[i] ? m : 0;
    ax[i] = static_cast<int>(mm % W);
    ay[i] = static_cast<int>((mm / W) % H);
    ab[i] = static_cast<int>(mm / (static_cast<long long>(W) * H));
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += KT) {
    __syncthreads();
    if (VEC) {
      const int k = k0 + a_k[0];
      const int tap = k / Cin, ci = k % Cin;
      const int iy = ay[0] * st + tap / ksize - pad, ix = ax[0] * st + tap % ksize - pad;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int Hb = s2 ? Hi : H, Wb = s2 ? Wi : W;  // bounds of the (virtual) input grid the taps index
      if (a_ok[0] && iy >= 0 && iy < Hb && ix >= 0 && ix < Wb) {
        const int sy = up == 1 ? (iy >> 1) : iy, sx = up == 1 ? (ix >> 1) : ix;
        v = *reinterpret_cast<const float4*>(x + ((static_cast<long long>(ab[0]) * Hi + sy) * Wi + sx) * Cin + ci);
      }
      sA[a_k[0] + 0][a_row[0]] = v.x; sA[a_k[0] + 1][a_row[0]] = v.y;
      sA[a_k[0] + 2][a_row[0]] = v.z; sA[a_k[0] + 3][a_row[0]] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < A_PER_THREAD; ++i) {
        const int k = k0 + a_k[i];
        float v = 0.f;
        if (k < K && a_ok[i]) {
          const int tap = k / Cin, ci = k % Cin;
          const int iy = ay[i] * st + tap / ksize - pad, ix = ax[i] * st + tap % ksize - pad;
          const int Hb = s2 ? Hi : H, Wb = s2 ? Wi : W;
          if (iy >= 0 && iy < Hb && ix >= 0 && ix < Wb) {
            const int sy = up == 1 ? (iy >> 1) : iy, sx = up == 1 ? (ix >> 1) : ix;
            v = x[((static_cast<long long>(ab[i]) * Hi + sy) * Wi + sx) * Cin + ci];
          }
        }
        sA[a_k[i]][a_row[i]] = v;
      }
    }
    // B tile: [KT][BN] from wk[k][n]
    for (int idx = threadIdx.x; idx < KT * BN; idx += 256) {
      const int kk = idx / BN, nn = idx % BN;
      const int k = k0 + kk, n = n0 + nn;
      sB[kk][nn] = (k < K && n < Cout) ? wk[static_cast<long long>(k) * Cout + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = sA[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < Cout) {
        float v = acc[i][j];
        if (bias) v += bias[n];
        if (res) v += res[m * Cout + n];
        y[m * Cout + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------- GroupNorm + SiLU (NHWC)
// pass 1: per-(image, channel) moments {sum, sumsq} in fp64 (atomics across pixel chunks)
// pass 2: fold channels into groups -> per-(image, channel) scale = rstd*gamma, shift = beta - mean*scale
// pass 3: y = silu(x * scale + shift)
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, float* __restrict__ partials, int HW, int C, int rows_per_block) {
  pdl_enter();
  // deterministic: fixed thread->data mapping, fixed-order in-block reduction, per-block partials (no atomics)
  extern __shared__ float s_part[];  // [rstep][C][2]
  const int b = blockIdx.y;
  const int c4 = C / 4;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const float* base = x + static_cast<long long>(b) * HW * C;
  const int q = threadIdx.x % c4;          // fixed channel quad per thread (256 % c4 == 0)
  const int lane_r = threadIdx.x / c4;
  const int rstep = blockDim.x / c4;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = r0 + lane_r; r < r1; r += rstep) {
    const float4 v = *reinterpret_cast<const float4*>(base + static_cast<long long>(r) * C + q * 4);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_part[(lane_r * C + q * 4 + j) * 2] = s[j];
    s_part[(lane_r * C + q * 4 + j) * 2 + 1] = ss[j];
  }
  __syncthreads();
  float* out = partials + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * C * 2;  // [image][block][C][2]
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < rstep; ++k) acc += s_part[k * C * 2 + i];
    out[i] = acc;
  }
}

// one warp per (image, group): fp64 sum of the per-block partials [image][block][C][2] in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partials, int nblocks, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ scale_shift, int B, int HW, int C, int groups,
                   float eps) {
  pdl_enter();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= B * groups) return;
  const int lane = threadIdx.x & 31;
  const int b = w / groups, g = w % groups;
  const int cpg = C / groups;
  const float* base = partials + static_cast<long long>(b) * nblocks * C * 2;
  double sum = 0.0, sq = 0.0;
  const int items = nblocks * cpg;
  for (int i = lane; i < items; i += 32) {
    const int blk = i / cpg, k = i % cpg;
    const float2 v = *reinterpret_cast<const float2*>(base + (static_cast<long long>(blk) * C + g * cpg + k) * 2);
    sum += static_cast<double>(v.x);
    sq += static_cast<double>(v.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const double n = static_cast<double>(HW) * cpg;
  const double mean = sum / n;
  double var = sq / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  for (int k = lane; k < cpg; k += 32) {
    const int c = g * cpg + k;
    const float sc = rstd * gamma[c];
    scale_shift[(static_cast<long long>(b) * C + c) * 2] = sc;
    scale_shift[(static_cast<long long>(b) * C + c) * 2 + 1] = beta[c] - static_cast<float>(mean) * sc;
  }
}

// SPLIT: instead of fp32 y, emit the bf16 hi/lo planes (y = hi + lo) the tensor-core convolution consumes (conv_tc.cu)
template <bool SPLIT>
__global__ void __launch_bounds__(256)
gn_apply_silu_kernel(const float* __restrict__ x, const float* __restrict__ scale_shift, float* __restrict__ y,
                     bf16* __restrict__ y_hi, bf16* __restrict__ y_lo, long long total4, int HW, int C, int silu) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int q = static_cast<int>(i % c4);
  const int b = static_cast<int>(i / (static_cast<long long>(c4) * HW));
  const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
  const float4 s0 = *reinterpret_cast<const float4*>(scale_shift + (static_cast<long long>(b) * C + q * 4) * 2);
  const float4 s1 = *reinterpret_cast<const float4*>(scale_shift + (static_cast<long long>(b) * C + q * 4) * 2 + 4);
  float o[4] = {fmaf(v.x, s0.x, s0.y), fmaf(v.y, s0.z, s0.w), fmaf(v.z, s1.x, s1.y), fmaf(v.w, s1.z, s1.w)};
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = silu ? o[j] / (1.f + expf(-o[j])) : o[j];
  if (SPLIT) {
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = bf16_round(o[j]);
    uint2 uh, ul;
    uh.x = pack_bf16(h[0], h[1]); uh.y = pack_bf16(h[2], h[3]);
    ul.x = pack_bf16(o[0] - h[0], o[1] - h[1]); ul.y = pack_bf16(o[2] - h[2], o[3] - h[3]);
    *reinterpret_cast<uint2*>(y_hi + i * 4) = uh;
    if (y_lo) *reinterpret_cast<uint2*>(y_lo + i * 4) = ul;
  } else {
    *reinterpret_cast<float4*>(y + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// 8 channels per thread variant of the SPLIT apply (two 16-byte loads in flight, 16-byte stores per plane)
__global__ void __launch_bounds__(256)
gn_apply_silu_split8_kernel(const float* __restrict__ x, const float* __restrict__ scale_shift, bf16* __restrict__ y_hi,
                            bf16* __restrict__ y_lo, long long total8, int HW, int C, int silu) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int q = static_cast<int>(i % c8);
  const int b = static_cast<int>(i / (static_cast<long long>(c8) * HW));
  float v[8];
  load8(x + i * 8, v);
  const float4* ssp = reinterpret_cast<const float4*>(scale_shift + (static_cast<long long>(b) * C + q * 8) * 2);
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 s2 = ssp[j];  // {scale, shift} of channels 2j, 2j+1
    float o0 = fmaf(v[2 * j], s2.x, s2.y), o1 = fmaf(v[2 * j + 1], s2.z, s2.w);
    if (silu) {
      o0 = o0 / (1.f + expf(-o0));
      o1 = o1 / (1.f + expf(-o1));
    }
    h[2 * j] = bf16_round(o0); h[2 * j + 1] = bf16_round(o1);
    l[2 * j] = o0 - h[2 * j]; l[2 * j + 1] = o1 - h[2 * j + 1];
  }
  store8(y_hi + i * 8, h);
  if (y_lo) store8(y_lo + i * 8, l);  // single-pass bf16 tokenizer mode needs the hi plane only
}

// ---------------------------------------------------------------- pooling / layout
__global__ void __launch_bounds__(256)
avgpool2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int Ho, int Wo, int C) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const int c4 = C / 4;
  const long long total = static_cast<long long>(B) * Ho * Wo * c4;
  if (i >= total) return;
  const int q = static_cast<int>(i % c4);
  const int ox = static_cast<int>((i / c4) % Wo);
  const int oy = static_cast<int>((i / (static_cast<long long>(c4) * Wo)) % Ho);
  const int b = static_cast<int>(i / (static_cast<long long>(c4) * Wo * Ho));
  const int Wi = Wo * 2;
  const float* p = x + ((static_cast<long long>(b) * Ho * 2 + oy * 2) * Wi + ox * 2) * C + q * 4;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 bb = *reinterpret_cast<const float4*>(p + C);
  const float4 c = *reinterpret_cast<const float4*>(p + static_cast<long long>(Wi) * C);
  const float4 d = *reinterpret_cast<const float4*>(p + static_cast<long long>(Wi) * C + C);
  // same association as ATen's avg_pool2d accumulation: ((a + b) + c) + d, then / 4
  *reinterpret_cast<float4*>(y + i * 4) = make_float4(((a.x + bb.x) + c.x + d.x) * 0.25f, ((a.y + bb.y) + c.y + d.y) * 0.25f,
                                                      ((a.z + bb.z) + c.z + d.z) * 0.25f, ((a.w + bb.w) + c.w + d.w) * 0.25f);
}

// out[b, p, c] = in[b, c, p] (to_nhwc) or the inverse; 32x32 smem transpose tiles
__global__ void __launch_bounds__(256)
transpose_cp_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  pdl_enter();
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const float* ip = in + static_cast<long long>(b) * rows * cols;
  float* op = out + static_cast<long long>(b) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) t[j][tx] = ip[static_cast<long long>(r) * cols + c];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) op[static_cast<long long>(c) * rows + r] = t[tx][j];
  }
}

}  // namespace

int conv2d_nhwc(const float* x, const float* wk, const float* bias, const float* res, float* y, int B, int H, int W,
                int Cin, int Cout, int ksize, int upsample2x, cudaStream_t s) {
  if (ksize != 1 && ksize != 3) { set_last_error("conv2d: kernel size %d unsupported (1 or 3)", ksize); return MUSE_ERR_UNSUPPORTED; }
  if (upsample2x == 1 && ((H | W) & 1)) { set_last_error("conv2d: upsample2x needs even output dims"); return MUSE_ERR_INVALID; }
  if (upsample2x == 2 && ksize != 3) { set_last_error("conv2d: the stride-2 mode is the 3x3 taming Downsample"); return MUSE_ERR_INVALID; }
  const long long M = static_cast<long long>(B) * H * W;
  if (M <= 0) return MUSE_OK;
  const bool vec = (Cin % 8 == 0);
  if (Cout > 16) {
    dim3 grid(static_cast<unsigned>(ceil_div_ll(M, 128)), ceil_div(Cout, 128));
    if (vec) pdl_launch(grid, 256, 0, s)(conv2d_nhwc_kernel<8, 8, true>, x, wk, bias, res, y, B, H, W, Cin, Cout, ksize, upsample2x);
    else pdl_launch(grid, 256, 0, s)(conv2d_nhwc_kernel<8, 8, false>, x, wk, bias, res, y, B, H, W, Cin, Cout, ksize, upsample2x);
  } else {
    dim3 grid(static_cast<unsigned>(ceil_div_ll(M, 128)), 1);
    if (vec) pdl_launch(grid, 256, 0, s)(conv2d_nhwc_kernel<8, 1, true>, x, wk, bias, res, y, B, H, W, Cin, Cout, ksize, upsample2x);
    else pdl_launch(grid, 256, 0, s)(conv2d_nhwc_kernel<8, 1, false>, x, wk, bias, res, y, B, H, W, Cin, Cout, ksize, upsample2x);
  }
  return check_launch("conv2d_nhwc");
}

// workspace: partials float [nblocks*B*C*2] with nblocks = ceil(HW / rows_per_block) (see gn_workspace_floats)
static int gn_rows_per_block(int C) { return (256 / (C / 4)) * 64; }

long long gn_workspace_floats(int B, int HW, int C) {
  if (C % 4 != 0 || C > 1024 || 256 % (C / 4) != 0) return -1;
  return static_cast<long long>(ceil_div(HW, gn_rows_per_block(C))) * B * C * 2;
}

// y (fp32) or y_hi / y_lo (bf16 split planes) receive the result: exactly one of the two forms must be given.
int groupnorm_silu_nhwc(const float* x, const float* gamma, const float* beta, float* y, void* y_hi, void* y_lo,
                        float* partials_ws, float* scale_shift_ws, int B, int HW, int C, int groups, float eps,
                        int precomputed_tiles, int apply_silu, cudaStream_t s) {
  if ((y != nullptr) == (y_hi != nullptr) || (y_hi == nullptr && y_lo != nullptr)) {
    set_last_error("groupnorm: give either y (fp32) or y_hi (+ y_lo unless the single-pass bf16 mode is used)");
    return MUSE_ERR_INVALID;
  }
  if (C % groups != 0 || C % 4 != 0 || C > 1024 || 256 % (C / 4) != 0) {
    set_last_error("groupnorm: C=%d groups=%d unsupported (C must divide into groups, C/4 must divide 256)", C, groups);
    return MUSE_ERR_UNSUPPORTED;
  }
  if (B <= 0 || HW <= 0) return MUSE_OK;
  // precomputed_tiles > 0: partials_ws already holds {sum, sumsq} per [image][tile][C], written by the epilogue of the
  // tensor-core convolution that produced x (conv_tc.cu), and the statistics pass over x is skipped.
  int nblocks = precomputed_tiles;
  int rc;
  if (precomputed_tiles <= 0) {
    const int rows_per_block = gn_rows_per_block(C);
    nblocks = ceil_div(HW, rows_per_block);
    const int rstep = 256 / (C / 4);
    pdl_launch(dim3(nblocks, B), 256, rstep * C * 2 * sizeof(float), s)(gn_stats_kernel, x, partials_ws, HW, C, rows_per_block);
    rc = check_launch("gn_stats");
    if (rc) return rc;
  }
  pdl_launch(ceil_div(B * groups, 8), 256, 0, s)(gn_finalize_kernel, partials_ws, nblocks, gamma, beta, scale_shift_ws, B, HW, C,
                                                             groups, eps);
  rc = check_launch("gn_finalize");
  if (rc) return rc;
  const long long total4 = static_cast<long long>(B) * HW * (C / 4);
  const unsigned blocks = static_cast<unsigned>(ceil_div_ll(total4, 256));
  if (y != nullptr)
    pdl_launch(blocks, 256, 0, s)(gn_apply_silu_kernel<false>, x, scale_shift_ws, y, nullptr, nullptr, total4, HW, C, apply_silu);
  else if (C % 8 == 0)
    pdl_launch(static_cast<unsigned>(ceil_div_ll(total4 / 2, 256)), 256, 0, s)(gn_apply_silu_split8_kernel,
        x, scale_shift_ws, reinterpret_cast<bf16*>(y_hi), reinterpret_cast<bf16*>(y_lo), total4 / 2, HW, C, apply_silu);
  else
    pdl_launch(blocks, 256, 0, s)(gn_apply_silu_kernel<true>, x, scale_shift_ws, nullptr, reinterpret_cast<bf16*>(y_hi),
                                                      reinterpret_cast<bf16*>(y_lo), total4, HW, C, apply_silu);
  return check_launch("gn_apply_silu");
}

// Decoder output -> display bytes with the reference's recipe (muse/pipeline_muse.py:245-252, quirk Q10), same fp32 operation
// order so the bytes are identical: t = 2x - 1; t = clip(t, -1, 1); t = (t + 1) / 2; byte = (uint8)(255 * t)  (truncation).
// Input is the decoder's native NHWC tensor, output HWC uint8 per image: the NCHW round trip of the reference
// (decode -> permute(1, 2, 0) -> numpy) and the 4x larger fp32 device->host copy disappear.
__global__ void __launch_bounds__(256)
image_to_uint8_kernel(const float* __restrict__ x, unsigned char* __restrict__ y, long long n4) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
  const float f[4] = {v.x, v.y, v.z, v.w};
  uchar4 o;
  unsigned char* ob = reinterpret_cast<unsigned char*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float t = __fsub_rn(__fmul_rn(2.0f, f[k]), 1.0f);
    t = fminf(fmaxf(t, -1.0f), 1.0f);
    t = __fdiv_rn(__fadd_rn(t, 1.0f), 2.0f);
    ob[k] = static_cast<unsigned char>(static_cast<int>(__fmul_rn(255.0f, t)));  // truncation toward zero, values in [0, 255]
  }
  *reinterpret_cast<uchar4*>(y + i * 4) = o;
}

int image_to_uint8(const float* x, unsigned char* y, long long n, cudaStream_t s) {
  if (n <= 0) return MUSE_OK;
  if (n % 4 != 0) { set_last_error("image_to_uint8: element count must be a multiple of 4"); return MUSE_ERR_INVALID; }
  pdl_launch(static_cast<unsigned>(ceil_div_ll(n / 4, 256)), 256, 0, s)(image_to_uint8_kernel, x, y, n / 4);
  return check_launch("image_to_uint8");
}

int avgpool2_nhwc(const float* x, float* y, int B, int Ho, int Wo, int C, cudaStream_t s) {
  if (C % 4 != 0) { set_last_error("avgpool: C must be a multiple of 4"); return MUSE_ERR_UNSUPPORTED; }
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 4);
  if (total <= 0) return MUSE_OK;
  pdl_launch(static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, s)(avgpool2_nhwc_kernel, x, y, B, Ho, Wo, C);
  return check_launch("avgpool2_nhwc");
}

// in: [B, rows, cols] -> out: [B, cols, rows]
int transpose_batched(const float* in, float* out, int B, int rows, int cols, cudaStream_t s) {
  if (B <= 0 || rows <= 0 || cols <= 0) return MUSE_OK;
  pdl_launch(dim3(ceil_div(cols, 32), ceil_div(rows, 32), B), 256, 0, s)(transpose_cp_kernel, in, out, rows, cols);
  return check_launch("transpose_batched");
}

}  // namespace muse
