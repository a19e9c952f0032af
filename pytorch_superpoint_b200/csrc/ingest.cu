// NEXT row f-4 (image ingest): datasets/Coco.py:165-179 `_read_image` after the JPEG decode --
//   cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA) -> cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) -> float32 / 255
// on the decoded uint8 HxWx3 image, bit-exact against OpenCV (4.x).  Down-scaling (both scale factors >= 1: COCO
// 640x480 -> 320x240, HPatches -> 320x240 / 640x480):
//   * integer scale factors take OpenCV's `resizeAreaFast_` arithmetic: integer box sums, (s + 2) >> 2 for 2x2,
//     round-half-even of s * (1.f / area) otherwise;
//   * everything else `resizeArea_`: per axis a table of (source cell, weight) entries -- a partial cell, full cells,
//     a partial cell, weights relative to the destination cell's width, computed in double and rounded to fp32 -- a
//     row accumulates S * alpha in table order, rows accumulate beta * row, every product and sum rounded to fp32
//     separately (no FMA), round-half-even to uint8 at the end;
//   * the gray value is OpenCV's 15-bit fixed point (R2Y 9798, G2Y 19235, B2Y 3735, + 2^14, >> 15) applied to the
//     channels in MEMORY order (the reference passes imread's BGR data to RGB2GRAY -- kept), then float(g) / 255.
// One thread per destination pixel; the source image is read through L1/L2 (a COCO image is 0.9 MB).
#include "common.cuh"

namespace spb {

// OpenCV computeResizeAreaTab for destination index d
__device__ __forceinline__ void area_cell(int d, int ssize, double scale, int& sfirst, int& sx1, int& sx2, float& ahead,
                                          float& afull, float& atail, bool& head, bool& tail) {
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  sx1 = (int)ceil(fsx1);
  sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  head = (double)sx1 - fsx1 > 1e-3;
  tail = fsx2 - (double)sx2 > 1e-3;
  ahead = (float)(((double)sx1 - fsx1) / cell);
  afull = (float)(1.0 / cell);
  atail = (float)(fmin(fmin(fsx2 - (double)sx2, 1.0), cell) / cell);
  sfirst = sx1 - 1;
}

__device__ __forceinline__ float to_gray01(int c0, int c1, int c2) {
  const int g = (c0 * 9798 + c1 * 19235 + c2 * 3735 + (1 << 14)) >> 15;
  return __fdiv_rn((float)g, 255.f);
}
__device__ __forceinline__ int round_u8(float v) {  // saturate_cast<uchar>(float): round half to even, clamp
  const int r = __float2int_rn(v);
  return min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) ingest_fast_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                          float* __restrict__ out, int H, int W, int iy, int ix) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  int s[3] = {0, 0, 0};
  for (int dy = 0; dy < iy; ++dy) {
    const uint8_t* row = src + ((long long)(y * iy + dy) * Ws + x * ix) * 3;
    for (int dx = 0; dx < ix; ++dx)
      for (int c = 0; c < 3; ++c) s[c] += __ldg(row + dx * 3 + c);
  }
  int v[3];
  if (ix == 2 && iy == 2) {
    for (int c = 0; c < 3; ++c) v[c] = (s[c] + 2) >> 2;
  } else {
    const float scale = __fdiv_rn(1.f, (float)(ix * iy));
    for (int c = 0; c < 3; ++c) v[c] = round_u8(__fmul_rn((float)s[c], scale));
  }
  out[(long long)y * W + x] = to_gray01(v[0], v[1], v[2]);
}

__global__ void __launch_bounds__(256) ingest_area_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                          float* __restrict__ out, int H, int W, double sy, double sx) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  int xf, x1, x2, yf, y1, y2;
  float axh, axm, axt, ayh, aym, ayt;
  bool xh, xt, yh, yt;
  area_cell(x, Ws, sx, xf, x1, x2, axh, axm, axt, xh, xt);
  area_cell(y, Hs, sy, yf, y1, y2, ayh, aym, ayt, yh, yt);
  float sum[3] = {0.f, 0.f, 0.f};
  bool first = true;
  auto row = [&](int sy_, float beta) {  // buf = sum_x S * alpha in table order; sum (+)= beta * buf
    const uint8_t* r = src + (long long)sy_ * Ws * 3;
    float buf[3] = {0.f, 0.f, 0.f};
    auto cell = [&](int sx_, float alpha) {
      for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)__ldg(r + sx_ * 3 + c), alpha));
    };
    if (xh) cell(xf, axh);
    for (int k = x1; k < x2; ++k) cell(k, axm);
    if (xt) cell(x2, axt);
    for (int c = 0; c < 3; ++c)
      sum[c] = first ? __fmul_rn(beta, buf[c]) : __fadd_rn(sum[c], __fmul_rn(beta, buf[c]));
    first = false;
  };
  if (yh) row(yf, ayh);
  for (int k = y1; k < y2; ++k) row(k, aym);
  if (yt) row(y2, ayt);
  out[(long long)y * W + x] = to_gray01(round_u8(sum[0]), round_u8(sum[1]), round_u8(sum[2]));
}

// OpenCV's INTER_AREA when an axis is ENLARGED (scale < 1 on either axis; e.g. KITTI 375x1242 -> 384x1248): the 8-bit
// linear resize with "area" coefficients -- per axis s = floor(d * scale), f = (float)((d + 1) - (s + 1) / scale)
// clipped to [0, 1) by its fractional part, 11-bit fixed-point weights, a horizontal pass into int32 and the vertical
// pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; past the last source column / row only the
// first tap counts.
__device__ __forceinline__ void linear_cell(int d, int ssize, double scale, double inv, int& s0, int& w0, int& w1,
                                            bool& edge) {
  int s = (int)floor(d * scale);
  float f = (float)((double)(d + 1) - (double)(s + 1) * inv);
  f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
  if (s < 0) {
    f = 0.f;
    s = 0;
  }
  edge = s + 1 >= ssize;
  if (s >= ssize - 1) {
    f = 0.f;
    s = ssize - 1;
  }
  s0 = s;
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256) ingest_linear_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                            float* __restrict__ out, int H, int W, double sy, double sx,
                                                            double invy, double invx) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  int x0, a0, a1, y0, b0, b1;
  bool xe, ye;
  linear_cell(x, Ws, sx, invx, x0, a0, a1, xe);
  linear_cell(y, Hs, sy, invy, y0, b0, b1, ye);
  const int x1 = min(x0 + 1, Ws - 1), y1 = min(y0 + 1, Hs - 1);
  const uint8_t* r0 = src + (long long)y0 * Ws * 3;
  const uint8_t* r1 = src + (long long)y1 * Ws * 3;
  int v[3];
  for (int c = 0; c < 3; ++c) {
    const int h0 = xe ? (int)__ldg(r0 + x0 * 3 + c) * 2048 : (int)__ldg(r0 + x0 * 3 + c) * a0 + (int)__ldg(r0 + x1 * 3 + c) * a1;
    const int h1 = xe ? (int)__ldg(r1 + x0 * 3 + c) * 2048 : (int)__ldg(r1 + x0 * 3 + c) * a0 + (int)__ldg(r1 + x1 * 3 + c) * a1;
    v[c] = min(max((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
  }
  out[(long long)y * W + x] = to_gray01(v[0], v[1], v[2]);
}

}  // namespace spb

using namespace spb;

extern "C" int spb_ingest_resize_gray(const uint8_t* bgr, int Hs, int Ws, float* out, int H, int W, void* stream) {
  SPB_CHECK_ARG(bgr && out && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1);
  // OpenCV: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
  const double sx = 1.0 / ((double)W / (double)Ws), sy = 1.0 / ((double)H / (double)Hs);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
  if (sx < 1.0 || sy < 1.0) {  // an axis is enlarged: OpenCV's linear variant, on BOTH axes
    ingest_linear_kernel<<<grid, 256, 0, as_stream(stream)>>>(bgr, Hs, Ws, out, H, W, sy, sx, (double)H / (double)Hs,
                                                              (double)W / (double)Ws);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
  }
  const int ix = (int)lrint(sx), iy = (int)lrint(sy);
  const bool fast = fabs(sx - ix) < 2.220446049250313e-16 && fabs(sy - iy) < 2.220446049250313e-16;
  if (fast)
    ingest_fast_kernel<<<grid, 256, 0, as_stream(stream)>>>(bgr, Hs, Ws, out, H, W, iy, ix);
  else
    ingest_area_kernel<<<grid, 256, 0, as_stream(stream)>>>(bgr, Hs, Ws, out, H, W, sy, sx);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}
