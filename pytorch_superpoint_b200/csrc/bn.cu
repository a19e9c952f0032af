// Train-mode BatchNorm (batch statistics) -- the as-written behaviour of the reference, which never calls .eval()
// (models/model_wrap.py:111 is commented out, Val_model_heatmap.py:61-73 has none): every conv -> BN -> ReLU of
// models/unet_parts.py:14-21 / SuperPointNet_gauss2.py:56-69 then normalises with the mean / biased variance of the
// CURRENT batch (the N warped views of one image in the HA export).  The product path folds eval-mode BN into the
// weights (decision 1); this file is the opt-in compatibility path (`SuperPointNet_b200(bn_mode='batch')`):
//   fp32 direct convolution (CUDA cores, fp32 operands and accumulation, no ReLU / pool, fp32 rows)
//   -> per-channel batch statistics (double accumulation)  -> scale / shift
//   -> y = x * scale + shift, ReLU, 2x2 max-pool, fp32 NHWC (+ L2 normalisation for the descriptor head).
// Everything stays fp32: batch normalisation brings every layer -- the logits included -- to unit scale, where bf16
// operands (1-2 % accumulated relative error) show up as 3e-2 in the heat-map; this path is about fidelity to the
// as-written reference, not speed.  Three passes over every activation instead of one fused epilogue: batch
// statistics need the whole layer output before the first normalised value exists -- this is why the product path
// does not do it.
#include "net.cuh"

namespace spb {

// KSxKS "same" convolution, NHWC fp32 in, fp32 rows out; 8x8 pixel tile per CTA, thread = (pixel, quarter of Cout).
// weights fp32 [tap][Cin][CoutPad], CoutPad = 4 * CPT (zero padded), bias [CoutPad].
template <int CPT, int KS>
__global__ void __launch_bounds__(256) conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ wgt,
                                                       const float* __restrict__ bias, float* __restrict__ out, int B,
                                                       int H, int W, int Cin, int Cout) {
  constexpr int HALO = KS - 1, TW = 8 + HALO, CCH = 32, PS = CCH + 1;  // pixel stride 33: no bank conflicts
  __shared__ float s_in[TW * TW * PS];
  const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
  const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = (tile % tiles_x) * 8, y0 = (tile / tiles_x) * 8;
  const int p = threadIdx.x & 63, cog = threadIdx.x >> 6;
  const int px = p & 7, py = p >> 3;
  constexpr int CoutPad = 4 * CPT;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CCH) {
    const int nc = min(CCH, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < TW * TW * nc; i += 256) {
      const int pix = i / nc, c = i % nc;
      const int yy = y0 + pix / TW - HALO / 2, xx = x0 + pix % TW - HALO / 2;
      s_in[pix * PS + c] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                ? __ldg(in + (((long long)n * H + yy) * W + xx) * Cin + c0 + c)
                                : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < KS * KS; ++tap) {
      const float* a = s_in + ((py + tap / KS) * TW + (px + tap % KS)) * PS;
      const float* wrow = wgt + ((long long)tap * Cin + c0) * CoutPad + cog * CPT;
#pragma unroll 2
      for (int ci = 0; ci < nc; ++ci) {
        const float av = a[ci];
        const float4* w4 = reinterpret_cast<const float4*>(wrow + (long long)ci * CoutPad);
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 wv = __ldg(w4 + j);
          acc[4 * j] = fmaf(av, wv.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(av, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(av, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(av, wv.w, acc[4 * j + 3]);
        }
      }
    }
  }
  const int oy = y0 + py, ox = x0 + px;
  if (oy >= H || ox >= W) return;
  float* op = out + (((long long)n * H + oy) * W + ox) * Cout + cog * CPT;
#pragma unroll
  for (int j = 0; j < CPT; ++j)
    if (cog * CPT + j < Cout) op[j] = acc[j] + __ldg(bias + cog * CPT + j);
}

// rows [R,C]: acc[c] += sum x, acc[C + c] += sum x^2 (double).  blockDim = 256 = lanes x cp, cp = pow2 >= C.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ rows, long long R, int C, int cp,
                                                       double* __restrict__ acc) {
  const int c = threadIdx.x % cp, rl = threadIdx.x / cp, lanes = 256 / cp;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  for (long long r = (long long)blockIdx.x * lanes + rl; r < R; r += (long long)gridDim.x * lanes) {
    const double v = (double)__ldg(rows + r * C + c);
    s += v;
    ss += v * v;
  }
  atomicAdd(acc + c, s);
  atomicAdd(acc + C + c, ss);
}

// F.batch_norm(training=True): mean, BIASED variance; scale = gamma / sqrt(var + eps), shift = beta - mean * scale
__global__ void bn_finalize_kernel(const double* __restrict__ acc, long long R, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ mean,
                                   float* __restrict__ var, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[c] / (double)R;
  double v = acc[C + c] / (double)R - m * m;
  if (v < 0.0) v = 0.0;
  const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
  const double sc = g / sqrt(v + (double)eps);
  if (mean) mean[c] = (float)m;
  if (var) var[c] = (float)v;
  scale[c] = (float)sc;
  shift[c] = (float)(b - m * sc);
}

// y = x * scale[c] + shift[c] (+ ReLU) (+ 2x2 max-pool): rows [B,H,W,C] fp32 -> out [B,OH,OW,C] bf16 or fp32
template <bool POOL>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ rows, int B, int H, int W, int C,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       int relu, void* __restrict__ out, int out_f32) {
  const int OH = POOL ? H / 2 : H, OW = POOL ? W / 2 : W;
  const long long total = (long long)B * OH * OW * C;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += gridDim.x * 256LL) {
    const int c = (int)(e % C);
    const long long p = e / C;
    const float sc = __ldg(scale + c), sh = __ldg(shift + c);
    float v;
    if (POOL) {
      const int ox = (int)(p % OW), oy = (int)((p / OW) % OH);
      const long long n = p / ((long long)OW * OH);
      const float* q = rows + (((n * H + 2 * oy) * W) + 2 * ox) * (long long)C + c;
      const float a0 = fmaf(__ldg(q), sc, sh), a1 = fmaf(__ldg(q + C), sc, sh);
      const float a2 = fmaf(__ldg(q + (long long)W * C), sc, sh), a3 = fmaf(__ldg(q + (long long)W * C + C), sc, sh);
      v = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    } else {
      v = fmaf(__ldg(rows + e), sc, sh);
    }
    if (relu) v = fmaxf(v, 0.f);
    if (out_f32)
      static_cast<float*>(out)[e] = v;
    else
      static_cast<__nv_bfloat16*>(out)[e] = __float2bfloat16_rn(v);
  }
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_conv_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int cin, int cout,
                 int cout_pad, int ks, void* stream) {
  SPB_CHECK_ARG(in && w && bias && out && B >= 1 && H >= 1 && W >= 1 && cin >= 1 && cout >= 1 && cout <= cout_pad);
  cudaStream_t st = as_stream(stream);
  const int blocks = B * ceil_div(W, 8) * ceil_div(H, 8);
#define SPB_LAUNCH(CPT, KS) conv_f32_kernel<CPT, KS><<<blocks, 256, 0, st>>>(in, w, bias, out, B, H, W, cin, cout)
  if (ks == 3 && cout_pad == 64) SPB_LAUNCH(16, 3);
  else if (ks == 3 && cout_pad == 128) SPB_LAUNCH(32, 3);
  else if (ks == 3 && cout_pad == 256) SPB_LAUNCH(64, 3);
  else if (ks == 1 && cout_pad == 80) SPB_LAUNCH(20, 1);
  else if (ks == 1 && cout_pad == 256) SPB_LAUNCH(64, 1);
  else {
    set_error("spb_conv_f32: unsupported layer shape ks=%d cout_pad=%d (3x3: 64 / 128 / 256, 1x1: 80 / 256)", ks, cout_pad);
    return SPB_EINVAL;
  }
#undef SPB_LAUNCH
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

size_t spb_bn_batch_stats_workspace_bytes(int C) { return (size_t)2 * (C > 0 ? C : 0) * sizeof(double); }

int spb_bn_batch_stats(const float* rows, long long R, int C, const float* gamma, const float* beta, float eps,
                       float* mean, float* var, float* scale, float* shift, void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(rows && scale && shift && ws && R >= 1 && C >= 1 && C <= 256 && eps >= 0.f);
  if (ws_bytes < spb_bn_batch_stats_workspace_bytes(C) || (reinterpret_cast<uintptr_t>(ws) & 7)) {
    set_error("spb_bn_batch_stats: workspace needs %zu bytes, 8-byte aligned (got %zu)",
              spb_bn_batch_stats_workspace_bytes(C), ws_bytes);
    return SPB_EINVAL;
  }
  cudaStream_t st = as_stream(stream);
  double* acc = static_cast<double*>(ws);
  SPB_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * C * sizeof(double), st));
  int cp = 1;
  while (cp < C) cp <<= 1;
  const int lanes = 256 / cp;
  const long long want = ceil_div(R, (long long)lanes * 64);  // ~64 rows per thread
  const int grid = (int)max(1LL, min(want, (long long)kNumSMs * 16));
  bn_stats_kernel<<<grid, 256, 0, st>>>(rows, R, C, cp, acc);
  SPB_CHECK_LAUNCH();
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, R, C, gamma, beta, eps, mean, var, scale, shift);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_bn_apply(const float* rows, int B, int H, int W, int C, const float* scale, const float* shift, int relu,
                 int pool, void* out, int out_f32, int l2norm, void* stream) {
  SPB_CHECK_ARG(rows && scale && shift && out && B >= 1 && H >= 1 && W >= 1 && C >= 1);
  if (pool && ((H | W) & 1)) {
    set_error("spb_bn_apply: pooling needs even H, W (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  if (l2norm && (!out_f32 || pool)) {
    set_error("spb_bn_apply: l2norm is for un-pooled fp32 rows");
    return SPB_EINVAL;
  }
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)B * (pool ? H / 2 : H) * (pool ? W / 2 : W) * C;
  const int grid = (int)max(1LL, min((long long)ceil_div(total, 256), (long long)kNumSMs * 16));
  if (pool)
    bn_apply_kernel<true><<<grid, 256, 0, st>>>(rows, B, H, W, C, scale, shift, relu, out, out_f32);
  else
    bn_apply_kernel<false><<<grid, 256, 0, st>>>(rows, B, H, W, C, scale, shift, relu, out, out_f32);
  SPB_CHECK_LAUNCH();
  if (l2norm) return l2norm_rows(static_cast<float*>(out), (long long)B * H * W, C, st);
  return SPB_OK;
}

}  // extern "C"
