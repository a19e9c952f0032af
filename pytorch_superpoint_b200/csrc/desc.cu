// (5) Descriptor bilinear sampling at key-points fused with L2 normalisation.
// Replaces models/model_wrap.py:281-299 (sample_desc_from_points): grid = (x/(W/2)-1, y/(H/2)-1)
// computed in float64 then cast to fp32, grid_sample(align_corners=True) on the coarse map (so the
// map is sampled at x*(Wc-1)/W -- the reference's quirk, kept), then desc /= ||desc||_2.
//
// One warp per key-point; lanes stride the D channels.  NHWC (our network's layout) makes the four
// taps 4 contiguous D*4-byte reads; NCHW (the reference layout) is supported for API parity.
#include "common.cuh"

namespace spb {

// stride: floats per point in pts (3 for (x, y, conf) rows, 2 for (x, y)); normalize = 0 keeps the raw bilinear
// sample (models/model_utils.py:63-90, the batched deepFEPE-facing variant, does not re-normalise)
template <bool NHWC>
__global__ void __launch_bounds__(256) sample_desc_kernel(const float* __restrict__ coarse,
                                                          const float* __restrict__ pts,
                                                          const int* __restrict__ count, int K, int Hc, int Wc,
                                                          int D, int cell, float* __restrict__ out, int stride,
                                                          int normalize, int zero_pad) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  // blockIdx.y = image of a batch: coarse [B,Hc,Wc,D] / [B,D,Hc,Wc], pts [B,K,stride], count [B], out [B,K,D]
  coarse += (long long)blockIdx.y * Hc * Wc * D;
  pts += (long long)blockIdx.y * K * stride;
  out += (long long)blockIdx.y * K * D;
  if (count) count += blockIdx.y;
  const int n = count ? min(*count, K) : K;
  if (k >= n) {
    if (zero_pad && k < K)  // rows behind the image's key-point count: zero descriptors (never matched)
      for (int c = lane; c < D; c += 32) out[(long long)k * D + c] = 0.f;
    return;
  }
  const double x = (double)pts[stride * k], y = (double)pts[stride * k + 1];
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
  if (!normalize) return;
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  __syncwarp();
  for (int c = lane; c < D; c += 32) o[c] = __fdiv_rn(o[c], nrm);  // zero norm -> NaN, like numpy
}

}  // namespace spb

using namespace spb;

extern "C" int spb_sample_desc_from_points_ex(const float* coarse, int nhwc, const float* pts, int pts_stride,
                                              const int* count, int K, int Hc, int Wc, int D, int cell, int normalize,
                                              float* out, void* stream) {
  SPB_CHECK_ARG(coarse && pts && out);
  SPB_CHECK_ARG(K >= 0 && Hc >= 2 && Wc >= 2 && D >= 1 && cell >= 1 && pts_stride >= 2);
  if (K == 0) return SPB_OK;
  const int blocks = ceil_div(K, 8);
  if (nhwc)
    sample_desc_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(coarse, pts, count, K, Hc, Wc, D, cell, out,
                                                                    pts_stride, normalize, 0);
  else
    sample_desc_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(coarse, pts, count, K, Hc, Wc, D, cell, out,
                                                                     pts_stride, normalize, 0);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// One launch for a batch (export.py:127-145 per image; Val_model_heatmap.py:151-155 is only valid for B = 1, the
// batched semantics are the per-image loop): coarse NHWC [B,Hc,Wc,D], pts [B,K,3], counts [B] -> out [B,K,D], the
// rows behind counts[b] zero filled.
extern "C" int spb_sample_desc_batch(const float* coarse_nhwc, const float* pts, const int* counts, int B, int K, int Hc,
                                     int Wc, int D, int cell, int normalize, float* out, void* stream) {
  SPB_CHECK_ARG(coarse_nhwc && pts && counts && out);
  SPB_CHECK_ARG(B >= 0 && B <= 65535 && K >= 0 && Hc >= 2 && Wc >= 2 && D >= 1 && cell >= 1);
  if (K == 0 || B == 0) return SPB_OK;
  dim3 grid(ceil_div(K, 8), B);
  sample_desc_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(coarse_nhwc, pts, counts, K, Hc, Wc, D, cell, out, 3,
                                                                normalize, 1);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

extern "C" int spb_sample_desc_from_points(const float* coarse, int nhwc, const float* pts, const int* count, int K,
                                           int Hc, int Wc, int D, int cell, float* out, void* stream) {
  return spb_sample_desc_from_points_ex(coarse, nhwc, pts, 3, count, K, Hc, Wc, D, cell, 1, out, stream);
}

// models/model_wrap.py:402-404 (`dense_desc_cpu[i, :, pts[1], pts[0]]`): the columns of the dense descriptor map at the
// integer key-points of every image, one launch for the batch.  dense NCHW [B,D,H,W], pts [B,K,3], counts [B] ->
// out [B,K,D] (rows behind counts[b] untouched).
namespace spb {
__global__ void __launch_bounds__(256) gather_dense_kernel(const float* __restrict__ dense, const float* __restrict__ pts,
                                                           const int* __restrict__ counts, int K, int D, int H, int W,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31, k = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
  if (k >= min(counts[b], K)) return;
  const float* p = pts + ((long long)b * K + k) * 3;
  const int x = min(max((int)p[0], 0), W - 1), y = min(max((int)p[1], 0), H - 1);
  const float* src = dense + (long long)b * D * H * W + (long long)y * W + x;
  float* o = out + ((long long)b * K + k) * D;
  for (int c = lane; c < D; c += 32) o[c] = __ldg(src + (long long)c * H * W);
}
}  // namespace spb

extern "C" int spb_gather_dense_desc(const float* dense_nchw, const float* pts, const int* counts, int B, int K, int D,
                                     int H, int W, float* out, void* stream) {
  SPB_CHECK_ARG(dense_nchw && pts && counts && out);
  SPB_CHECK_ARG(B >= 0 && B <= 65535 && K >= 0 && D >= 1 && H >= 1 && W >= 1);
  if (B == 0 || K == 0) return SPB_OK;
  dim3 grid(spb::ceil_div(K, 8), B);
  spb::gather_dense_kernel<<<grid, 256, 0, spb::as_stream(stream)>>>(dense_nchw, pts, counts, K, D, H, W, out);
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
  heat += (long long)blockIdx.y * H * W;  // blockIdx.y = image of a batch: heat [B,H,W], pts [B,K,3], dxdy [B,K,2]
  pts += (long long)blockIdx.y * K * 3;
  dxdy += (long long)blockIdx.y * K * 2;
  if (count) count += blockIdx.y;
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

// The same for a batch in one launch: heat [B,H,W], pts [B,K,3], counts [B] (device), dxdy [B,K,2].
extern "C" int spb_soft_argmax_points_batch(const float* heat, int B, int H, int W, const float* pts, const int* counts,
                                            int K, int patch_size, float* dxdy, void* stream) {
  SPB_CHECK_ARG(heat && pts && counts && dxdy && H > 0 && W > 0 && K >= 0 && B >= 0 && B <= 65535);
  SPB_CHECK_ARG(patch_size >= 1 && patch_size <= 7 && (patch_size & 1));
  if (K == 0 || B == 0) return SPB_OK;
  dim3 grid(spb::ceil_div(K, 128), B);
  spb::soft_argmax_kernel<<<grid, 128, 0, spb::as_stream(stream)>>>(heat, H, W, pts, counts, K, patch_size, dxdy);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}


// ---- NEXT row f-3: models/model_utils.py:24-60 pred_soft_argmax -------------------------------------------
// Patches as utils/losses.py:16-29,87-98 builds them: box (x-3, y-3, x+3, y+3) around every labelled pixel,
// torchvision roi_pool to 5x5 (a 7-pixel span in 5 bins of 1.4: rows/cols {0,1},{1,2},{2,3,4},{4,5},{5,6}, max over
// each bin, bins clipped to the image), then log (negatives -> 1e-6) and torchgeometry's SpatialSoftArgmax2d in
// pixel coordinates (restated, parity unpinned as in f-1b).  idx [N,4] int64 rows (batch, channel, y, x) = the
// rows of labels.nonzero(); heat [B,1,H,W].  dxdy [N,2] = soft-argmax position (caller subtracts patch//2).
namespace spb {
__global__ void __launch_bounds__(128) roi_soft_argmax_kernel(const float* __restrict__ heat, int H, int W,
                                                              const long long* __restrict__ idx, int N,
                                                              float* __restrict__ dxdy, float* __restrict__ patches) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= N) return;
  const long long b = idx[4 * k];
  const int y = (int)idx[4 * k + 2], x = (int)idx[4 * k + 3];
  const float* img = heat + b * H * W;
  const int lo[5] = {0, 1, 2, 4, 5}, hi[5] = {2, 3, 5, 6, 7};  // floor(p*1.4), ceil((p+1)*1.4)
  float v[25];
  float mx = -INFINITY;
  for (int py = 0; py < 5; ++py)
    for (int px = 0; px < 5; ++px) {
      const int y0 = min(max(y - 3 + lo[py], 0), H), y1 = min(max(y - 3 + hi[py], 0), H);
      const int x0 = min(max(x - 3 + lo[px], 0), W), x1 = min(max(x - 3 + hi[px], 0), W);
      float m = (y1 <= y0 || x1 <= x0) ? 0.f : -INFINITY;  // empty bin -> 0 (torchvision)
      for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) m = fmaxf(m, __ldg(img + (long long)yy * W + xx));
      if (patches) patches[(long long)k * 25 + py * 5 + px] = m;
      const float q = m < 0.f ? 1e-6f : m;
      v[py * 5 + px] = logf(q);
      mx = fmaxf(mx, v[py * 5 + px]);
    }
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int i = 0; i < 25; ++i) {
    const float e = expf(v[i] - mx);
    se += e;
    sx += (float)(i % 5) * e;
    sy += (float)(i / 5) * e;
  }
  const float inv = __fdiv_rn(1.f, se + 1e-6f);
  dxdy[2 * k] = sx * inv;
  dxdy[2 * k + 1] = sy * inv;
}
}  // namespace spb

extern "C" int spb_roi_soft_argmax(const float* heat, int B, int H, int W, const long long* idx, int N, float* dxdy,
                                   float* patches_or_null, void* stream) {
  SPB_CHECK_ARG(heat && dxdy && B >= 1 && H > 0 && W > 0 && N >= 0);
  if (N == 0) return SPB_OK;
  SPB_CHECK_ARG(idx != nullptr);
  spb::roi_soft_argmax_kernel<<<spb::ceil_div(N, 128), 128, 0, spb::as_stream(stream)>>>(heat, H, W, idx, N, dxdy,
                                                                                         patches_or_null);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}


// ---- NEXT row f-4 (dense-descriptor mode of run()): models/model_wrap.py:393-401 ----------------------------
// dense = F.interpolate(coarse, scale_factor=(cell, cell), mode='bilinear')  (align_corners=False: source index
// (dst + 0.5)/cell - 0.5 clamped at 0, second tap clamped to the last cell), then dense / ||dense||_2 over the
// channels.  coarse NHWC [B,Hc,Wc,D] (our network's layout), dense NCHW [B,D,H,W] (the reference's layout).
// Block = 32 consecutive pixels of one row x 8 channel groups: every channel plane is written in 128-byte
// segments; the squared norm is reduced over the channel groups through shared memory.
namespace spb {
__global__ void __launch_bounds__(256) dense_desc_kernel(const float* __restrict__ coarse, float* __restrict__ dense,
                                                         int Hc, int Wc, int D, int cell) {
  __shared__ float ss[8][33];
  const int W = Wc * cell, H = Hc * cell;
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y, b = blockIdx.z, g = threadIdx.y;
  const int xx = x < W ? x : W - 1;
  const float inv = 1.f / (float)cell;
  const float sx = fmaxf(__fsub_rn(__fmul_rn(inv, __fadd_rn((float)xx, 0.5f)), 0.5f), 0.f);
  const float sy = fmaxf(__fsub_rn(__fmul_rn(inv, __fadd_rn((float)y, 0.5f)), 0.5f), 0.f);
  const int x0 = (int)sx, y0 = (int)sy, x1 = min(x0 + 1, Wc - 1), y1 = min(y0 + 1, Hc - 1);
  const float lx = sx - (float)x0, ly = sy - (float)y0, mx = 1.f - lx, my = 1.f - ly;
  const float* c00 = coarse + (((long long)b * Hc + y0) * Wc + x0) * D;
  const float* c01 = coarse + (((long long)b * Hc + y0) * Wc + x1) * D;
  const float* c10 = coarse + (((long long)b * Hc + y1) * Wc + x0) * D;
  const float* c11 = coarse + (((long long)b * Hc + y1) * Wc + x1) * D;
  const int per = D / 8;  // channels per group (D % 8 == 0, per <= 32)
  float v[32];
  float acc = 0.f;
  for (int j = 0; j < per; ++j) {
    const int c = g * per + j;
    v[j] = my * (mx * __ldg(c00 + c) + lx * __ldg(c01 + c)) + ly * (mx * __ldg(c10 + c) + lx * __ldg(c11 + c));
    acc = fmaf(v[j], v[j], acc);
  }
  ss[g][threadIdx.x] = acc;
  __syncthreads();
  float tot = 0.f;
  for (int k = 0; k < 8; ++k) tot += ss[k][threadIdx.x];
  const float nrm = sqrtf(tot);
  if (x < W) {
    float* o = dense + (((long long)b * D + g * per) * H + y) * W + x;
    for (int j = 0; j < per; ++j) o[(long long)j * H * W] = __fdiv_rn(v[j], nrm);
  }
}
}  // namespace spb

extern "C" int spb_dense_desc(const float* coarse_nhwc, int B, int Hc, int Wc, int D, int cell, float* dense_nchw,
                              void* stream) {
  SPB_CHECK_ARG(coarse_nhwc && dense_nchw && B >= 1 && Hc >= 1 && Wc >= 1 && cell >= 1);
  SPB_CHECK_ARG(D >= 8 && D % 8 == 0 && D <= 256 && Hc * cell <= 65535 && B <= 65535);
  dim3 grid(spb::ceil_div(Wc * cell, 32), Hc * cell, B), block(32, 8);
  spb::dense_desc_kernel<<<grid, block, 0, spb::as_stream(stream)>>>(coarse_nhwc, dense_nchw, Hc, Wc, D, cell);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}
