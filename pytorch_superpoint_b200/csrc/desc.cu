// (5) Descriptor bilinear sampling at key-points fused with L2 normalisation.
// Replaces models/model_wrap.py:281-299 (sample_desc_from_points): grid = (x/(W/2)-1, y/(H/2)-1)
// computed in float64 then cast to fp32, grid_sample(align_corners=True) on the coarse map (so the
// map is sampled at x*(Wc-1)/W -- the reference's quirk, kept), then desc /= ||desc||_2.
//
// One warp per key-point; lanes stride the D channels.  NHWC (our network's layout) makes the four
// taps 4 contiguous D*4-byte reads; NCHW (the reference layout) is supported for API parity.
#include "common.cuh"

namespace spb {

template <bool NHWC>
__global__ void __launch_bounds__(256) sample_desc_kernel(const float* __restrict__ coarse,
                                                          const float* __restrict__ pts,
                                                          const int* __restrict__ count, int K, int Hc, int Wc,
                                                          int D, int cell, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = count ? min(*count, K) : K;
  if (k >= n) return;
  const double x = (double)pts[3 * k], y = (double)pts[3 * k + 1];
  const double halfw = (double)((float)(Wc * cell)) / 2.0, halfh = (double)((float)(Hc * cell)) / 2.0;
  const float gx = (float)(x / halfw - 1.0), gy = (float)(y / halfh - 1.0);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(Wc - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(Hc - 1));
  const Taps tp = make_taps(ix, iy);
  float ss = 0.f;
  float* o = out + (long long)k * D;
  for (int c = lane; c < D; c += 32) {
    const float v = bilinear(tp, Wc, Hc, [&](int xx, int yy) {
      return NHWC ? __ldg(coarse + ((long long)yy * Wc + xx) * D + c) : __ldg(coarse + ((long long)c * Hc + yy) * Wc + xx);
    });
    o[c] = v;
    ss = __fmaf_rn(v, v, ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  __syncwarp();
  for (int c = lane; c < D; c += 32) o[c] = __fdiv_rn(o[c], nrm);  // zero norm -> NaN, like numpy
}

}  // namespace spb

using namespace spb;

extern "C" int spb_sample_desc_from_points(const float* coarse, int nhwc, const float* pts, const int* count, int K,
                                           int Hc, int Wc, int D, int cell, float* out, void* stream) {
  SPB_CHECK_ARG(coarse && pts && out);
  SPB_CHECK_ARG(K >= 0 && Hc >= 2 && Wc >= 2 && D >= 1 && cell >= 1);
  if (K == 0) return SPB_OK;
  const int blocks = ceil_div(K, 8);
  if (nhwc)
    sample_desc_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(coarse, pts, count, K, Hc, Wc, D, cell, out);
  else
    sample_desc_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(coarse, pts, count, K, Hc, Wc, D, cell, out);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// ---- sub-pixel refinement (NEXT row f-1b): models/model_wrap.py:200-234 + utils/losses.py:48-129 ------------
// One thread per key-point: zero-padded PxP patch centred on the integer key-point, p/(sum p + 1e-6), negatives ->
// 1e-6, log, spatial soft-argmax in pixel coordinates sum(pos*e)/(sum e + 1e-6) (torchgeometry 0.1.x
// SpatialSoftArgmax2d, restated -- the package is not in the image, so this part is parity-unpinned).
namespace spb {
__global__ void __launch_bounds__(128) soft_argmax_kernel(const float* __restrict__ heat, int H, int W,
                                                          const float* __restrict__ pts, const int* __restrict__ count,
                                                          int K, int P, float* __restrict__ dxdy) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int n = count ? min(*count, K) : K;
  if (k >= n) return;
  const int x0 = (int)pts[3 * k] - P / 2, y0 = (int)pts[3 * k + 1] - P / 2;
  float v[49];
  float s = 0.f;
  for (int i = 0; i < P * P; ++i) {
    const int xx = x0 + i % P, yy = y0 + i / P;
    v[i] = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(heat + (long long)yy * W + xx) : 0.f;
    s += v[i];
  }
  const float d = s + 1e-6f;
  float mx = -INFINITY;
  for (int i = 0; i < P * P; ++i) {
    float q = __fdiv_rn(v[i], d);
    if (q < 0.f) q = 1e-6f;
    v[i] = logf(q);
    mx = fmaxf(mx, v[i]);
  }
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int i = 0; i < P * P; ++i) {
    const float e = expf(v[i] - mx);
    se += e;
    sx += (float)(i % P) * e;
    sy += (float)(i / P) * e;
  }
  const float inv = __fdiv_rn(1.f, se + 1e-6f);
  dxdy[2 * k] = sx * inv;
  dxdy[2 * k + 1] = sy * inv;
}
}  // namespace spb

extern "C" int spb_soft_argmax_points(const float* heat, int H, int W, const float* pts, const int* count, int K,
                                      int patch_size, float* dxdy, void* stream) {
  SPB_CHECK_ARG(heat && pts && dxdy && H > 0 && W > 0 && K >= 0);
  SPB_CHECK_ARG(patch_size >= 1 && patch_size <= 7 && (patch_size & 1));
  if (K == 0) return SPB_OK;
  spb::soft_argmax_kernel<<<spb::ceil_div(K, 128), 128, 0, spb::as_stream(stream)>>>(heat, H, W, pts, count, K,
                                                                                     patch_size, dxdy);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}
