// CUDA-core bring-up / validation kernels for the SuperPointNet layers (SPB_IMPL_SIMT).
// Same numerics contract as the tcgen05 path -- bf16 activations and weights, fp32 accumulate, bias +
// ReLU + 2x2 max-pool fused in the epilogue -- so the two implementations can be compared tightly
// layer by layer on the GPU.  Not the product path: see conv_tc.cu.
#include "net.cuh"

namespace spb {

// ---- conv1a: 1 -> 64, fp32 image in, NHWC bf16 out (models/SuperPointNet_pretrained.py:56) --------
__global__ void __launch_bounds__(256) conv1a_simt_kernel(const float* __restrict__ x, const float* __restrict__ w9x64,
                                                          const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  __shared__ float s_w[9 * 64];
  __shared__ float s_b[64];
  for (int i = threadIdx.x; i < 9 * 64; i += 256) s_w[i] = w9x64[i];
  if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long total = (long long)B * H * W;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int px = (int)(t % W), py = (int)((t / W) % H);
    const float* img = x + (t / ((long long)W * H)) * (long long)H * W;
    float in[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = py + r - 1, xx = px + s - 1;
        const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (long long)yy * W + xx) : 0.f;
        in[r * 3 + s] = __bfloat162float(__float2bfloat16_rn(v));  // tensor-core path feeds bf16 pixels
      }
    uint4* o = reinterpret_cast<uint4*>(out + t * 64);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a0 = s_b[g * 8 + 2 * j], a1 = s_b[g * 8 + 2 * j + 1];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          a0 = fmaf(in[k], s_w[k * 64 + g * 8 + 2 * j], a0);
          a1 = fmaf(in[k], s_w[k * 64 + g * 8 + 2 * j + 1], a1);
        }
        h[j] = __floats2bfloat162_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
      }
      o[g] = *reinterpret_cast<uint4*>(h);
    }
  }
}

// ---- generic KSxKS conv, NHWC bf16 in; 8x8 pixel tile per CTA, thread = (pixel, quarter of Cout) -----
// weights: bf16 [tap][Cin][CoutPad], CoutPad = 4*CPT.
template <int CPT, int KS>
__global__ void __launch_bounds__(256) conv_simt_kernel(const __nv_bfloat16* __restrict__ in,
                                                        const __nv_bfloat16* __restrict__ wgt,
                                                        const float* __restrict__ bias, void* __restrict__ out,
                                                        int B, int H, int W, int Cin, int Cout, int relu, int pool,
                                                        int out_f32) {
  constexpr int HALO = KS - 1, TW = 8 + HALO, CCH = 32;
  __shared__ __align__(16) __nv_bfloat16 s_in[TW * TW * CCH];
  const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
  const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = (tile % tiles_x) * 8, y0 = (tile / tiles_x) * 8;
  const int p = threadIdx.x & 63, cog = threadIdx.x >> 6;
  const int px = p & 7, py = p >> 3;
  const int CoutPad = 4 * CPT;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CCH) {
    __syncthreads();
    // stage the halo tile for 32 input channels: TW*TW pixels x 64 B, 16-byte pieces
    for (int i = threadIdx.x; i < TW * TW * 4; i += 256) {
      const int pix = i >> 2, q = i & 3;
      const int yy = y0 + pix / TW - HALO / 2, xx = x0 + pix % TW - HALO / 2;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + yy) * W + xx) * Cin + c0 + q * 8));
      *reinterpret_cast<uint4*>(s_in + pix * CCH + q * 8) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < KS * KS; ++tap) {
      const __nv_bfloat16* a = s_in + ((py + tap / KS) * TW + (px + tap % KS)) * CCH;
      const __nv_bfloat16* wrow = wgt + ((long long)tap * Cin + c0) * CoutPad + cog * CPT;
#pragma unroll 4
      for (int ci = 0; ci < CCH; ++ci) {
        const float av = __bfloat162float(a[ci]);
        const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(wrow + (long long)ci * CoutPad);
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j) {
          const float2 wv = __bfloat1622float2(__ldg(w2 + j));
          acc[2 * j] = fmaf(av, wv.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(av, wv.y, acc[2 * j + 1]);
        }
      }
    }
  }
  // epilogue: bias, ReLU, optional 2x2 max-pool via lane shuffles (lane = 4 rows x 8 cols of the tile)
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int co = cog * CPT + j;
    float v = acc[j] + (co < Cout ? __ldg(bias + co) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    if (pool) {
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
    }
    acc[j] = v;
  }
  int oy = y0 + py, ox = x0 + px, OH = H, OW = W;
  bool writer = oy < H && ox < W;
  if (pool) {
    writer = writer && !(px & 1) && !(py & 1);
    oy >>= 1;
    ox >>= 1;
    OH >>= 1;
    OW >>= 1;
  }
  if (!writer) return;
  const long long o = (((long long)n * OH + oy) * OW + ox) * Cout + cog * CPT;
  if (out_f32) {
    float* op = static_cast<float*>(out) + o;
    for (int j = 0; j < CPT; ++j)
      if (cog * CPT + j < Cout) op[j] = acc[j];
  } else {
    __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out) + o;
    for (int j = 0; j < CPT; ++j)
      if (cog * CPT + j < Cout) op[j] = __float2bfloat16_rn(acc[j]);
  }
}

// channel L2 normalisation of NHWC fp32 rows (models/SuperPointNet_pretrained.py:73-74)
__global__ void __launch_bounds__(256) l2norm_rows_kernel(float* __restrict__ d, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    float* p = d + r * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(p[c], p[c], ss);
    const float nrm = sqrtf(warp_sum(ss));
    for (int c = lane; c < C; c += 32) p[c] = __fdiv_rn(p[c], nrm);
  }
}

int conv1a_simt(const LayerDev& L, const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st) {
  const long long total = (long long)B * H * W;
  conv1a_simt_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st>>>(x, L.w1a_f32, L.bias, out, B, H, W);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int conv_simt(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_f32,
              cudaStream_t st) {
  const int blocks = B * ceil_div(W, 8) * ceil_div(H, 8);
  const int cpt = L.cout_pad / 4;
#define SPB_LAUNCH(CPT, KS)                                                                                   \
  conv_simt_kernel<CPT, KS><<<blocks, 256, 0, st>>>(in, L.w_simt, L.bias, out, B, H, W, L.cin, L.cout, L.relu, \
                                                     L.pool, out_f32)
  if (L.ks == 3 && cpt == 16) SPB_LAUNCH(16, 3);
  else if (L.ks == 3 && cpt == 32) SPB_LAUNCH(32, 3);
  else if (L.ks == 3 && cpt == 64) SPB_LAUNCH(64, 3);
  else if (L.ks == 1 && cpt == 20) SPB_LAUNCH(20, 1);
  else if (L.ks == 1 && cpt == 64) SPB_LAUNCH(64, 1);
  else {
    set_error("conv_simt: unsupported layer shape ks=%d cout_pad=%d", L.ks, L.cout_pad);
    return SPB_EINVAL;
  }
#undef SPB_LAUNCH
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int l2norm_rows(float* d, long long rows, int C, cudaStream_t st) {
  l2norm_rows_kernel<<<min(ceil_div(rows, 8), kNumSMs * 8), 256, 0, st>>>(d, rows, C);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb
