// (1) Batched per-image 3x3 homography warp + valid mask.
// Replaces utils/utils.py:318-351 (inv_warp_image_batch), 286-314 (warp_points), 677-704
// (compute_valid_mask) and the HA branch of datasets/Coco.py:279-284.
//
// HBM-bound gather: the source image (307 KB at 240x320) is read through L1/L2 (it is shared by
// all N warps of an image and stays L2-resident), each thread produces 4 consecutive output
// pixels and issues one 16-byte store, so the N*H*W output stream is fully coalesced.
#include <stdarg.h>

#include "common.cuh"

namespace spb {

static thread_local char g_err[512] = "no error";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
unsigned long long g_launches = 0;

// One thread = 4 consecutive x of one (n, y).  MODE 0 bilinear, 1 nearest, 2 valid mask (no source).
template <int MODE>
__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ img, long long img_stride,
                                                   const float* __restrict__ mats, float* __restrict__ out,
                                                   int N, int H, int W, Lin lx, Lin ly) {
  const int Wq = (W + 3) >> 2;
  const long long total = (long long)N * H * Wq;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(t % Wq);
    const int y = (int)((t / Wq) % H);
    const int n = (int)(t / ((long long)Wq * H));
    const Mat3 M = load_mat(mats + 9 * n);
    const float* src = img + (long long)n * img_stride;
    const float v = ly.at(y);
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = xq * 4 + j;
      float ix, iy;
      warp_coord(M, lx.at(x < W ? x : W - 1), v, W, H, ix, iy);
      if (MODE == 0) {
        const Taps tp = make_taps(ix, iy);
        r[j] = bilinear(tp, W, H, [&](int xx, int yy) { return __ldg(src + (long long)yy * W + xx); });
      } else {
        const float xr = rintf(ix), yr = rintf(iy);
        const bool ok = in_bounds(xr, yr, W, H);
        r[j] = !ok ? 0.f : (MODE == 2 ? 1.f : __ldg(src + (long long)(int)yr * W + (int)xr));
      }
    }
    float* dst = out + ((long long)n * H + y) * W + xq * 4;
    if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
      for (int j = 0; j < 4 && xq * 4 + j < W; ++j) dst[j] = r[j];
    }
  }
}

// HA variant: bilinear warp of ONE image into N views + the nearest-mode valid mask of the same coordinates
// (utils/utils.py:696 uses the same grid and matrices as the image warp, datasets/Coco.py:279-282), packed one
// bit per pixel: bits[(n*H + y)*(W/8) + x/8] bit (x%8).  Needs W % 8 == 0.
// View n belongs to image vm.img(n) and uses the matrix mats[vm.mat(n)] (common.cuh: ViewMap).
// Grid: blockIdx.y = view, blockIdx.x * 256 threads over the (row, 4-pixel group) pairs of that view -- the view's
// matrix is block-uniform and no thread divides 64-bit indices (those divisions were a quarter of the instructions).
// This is the bit-exact CUDA-core form (validation path, SPB_IMPL_SIMT); the product path is ha_fast.cu.
__global__ void __launch_bounds__(256) warp_bits_kernel(const float* __restrict__ img, const float* __restrict__ mats,
                                                        float* __restrict__ out, uint8_t* __restrict__ bits, ViewMap vm,
                                                        int H, int W, Lin lx, Lin ly) {
  const unsigned Wq = (unsigned)W >> 2, per_view = (unsigned)H * Wq;
  const int lane = threadIdx.x & 31, n = blockIdx.y;
  const unsigned t = blockIdx.x * 256u + threadIdx.x;
  const bool in = t < per_view;
  const unsigned tt = in ? t : per_view - 1;
  const int y = (int)(tt / Wq), xq = (int)(tt - (unsigned)y * Wq);
  const Mat3 M = load_mat(mats + 9LL * vm.mat(n));
  const float* src = img + (long long)vm.img(n) * H * W;
  const float v = ly.at(y);
  float r[4];
  unsigned nib = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float ix, iy;
    warp_coord(M, lx.at(xq * 4 + j), v, W, H, ix, iy);
    const Taps tp = make_taps(ix, iy);
    r[j] = bilinear_fast(tp, W, H, [&](int o) { return __ldg(src + o); });
    nib |= (in_bounds(rintf(ix), rintf(iy), W, H) ? 1u : 0u) << j;
  }
  const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
  if (in) {
    const long long row = (long long)n * H + y;
    *reinterpret_cast<float4*>(out + row * W + xq * 4) = make_float4(r[0], r[1], r[2], r[3]);
    if (!(xq & 1)) bits[row * (W >> 3) + (xq >> 1)] = (uint8_t)(nib | (hi << 4));
  }
}

int warp_with_mask_bits(const float* img, const float* mats, float* out, uint8_t* bits, int N, const ViewMap& vm,
                        int H, int W, cudaStream_t st) {
  if (N > 65535) {
    set_error("warp_with_mask_bits: %d views in one launch (max 65535)", N);
    return SPB_EINVAL;
  }
  dim3 grid(ceil_div((long long)H * (W / 4), 256), N);
  warp_bits_kernel<<<grid, 256, 0, st>>>(img, mats, out, bits, vm, H, W, Lin(W), Lin(H));
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// Erosion of the valid masks (utils/utils.py:699-702: cv2.erode with an elliptic structuring element, one
// iteration): out(y, x) = min over the element's pixels of in(y + i - ay, x + j - ax), pixels outside the image
// ignored (cv2's default border value for erosion is +max).  The element arrives as one column span [j1, j2) per row
// (the caller restates cv2.getStructuringElement's formula), so any row-convex element works.
__global__ void __launch_bounds__(256) erode_kernel(const float* __restrict__ in, float* __restrict__ out, int N,
                                                    int H, int W, const int* __restrict__ spans, int kh, int ay,
                                                    int ax) {
  const long long total = (long long)N * H * W;
  for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += gridDim.x * 256LL) {
    const int x = (int)(t % W), y = (int)((t / W) % H);
    const float* img = in + (t / ((long long)W * H)) * H * W;
    float m = INFINITY;
    for (int i = 0; i < kh; ++i) {
      const int yy = y + i - ay;
      if (yy < 0 || yy >= H) continue;
      const int j1 = max(__ldg(spans + 2 * i) - ax + x, 0), j2 = min(__ldg(spans + 2 * i + 1) - ax + x, W);
      for (int xx = j1; xx < j2; ++xx) m = fminf(m, __ldg(img + (long long)yy * W + xx));
    }
    out[t] = m;
  }
}

template <int MODE>
static int launch_warp(const float* img, long long stride, const float* mats, float* out, int N, int H, int W,
                       cudaStream_t st) {
  const long long total = (long long)N * H * ((W + 3) / 4);
  int blocks = ceil_div(total, 256);
  const int cap = kNumSMs * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  warp_kernel<MODE><<<blocks, 256, 0, st>>>(img, stride, mats, out, N, H, W, Lin(W), Lin(H));
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_version(void) { return 100; }
const char* spb_last_error_string(void) { return spb::last_error(); }
unsigned long long spb_launch_count(void) { return __atomic_load_n(&spb::g_launches, __ATOMIC_RELAXED); }

int spb_inv_warp_image_batch(const float* img, long long img_batch_stride, const float* mats, float* out, int N,
                             int H, int W, int mode, void* stream) {
  SPB_CHECK_ARG(img && mats && out);
  SPB_CHECK_ARG(N >= 0 && H >= 2 && W >= 2 && img_batch_stride >= 0);
  SPB_CHECK_ARG(mode == SPB_MODE_BILINEAR || mode == SPB_MODE_NEAREST);
  if (N == 0) return SPB_OK;
  return mode == SPB_MODE_BILINEAR ? launch_warp<0>(img, img_batch_stride, mats, out, N, H, W, as_stream(stream))
                                   : launch_warp<1>(img, img_batch_stride, mats, out, N, H, W, as_stream(stream));
}

int spb_compute_valid_mask(const float* mats, float* mask, int N, int H, int W, void* stream) {
  SPB_CHECK_ARG(mats && mask);
  SPB_CHECK_ARG(N >= 0 && H >= 2 && W >= 2);
  if (N == 0) return SPB_OK;
  return launch_warp<2>(mask /*unused*/, 0, mats, mask, N, H, W, as_stream(stream));
}

int spb_erode(const float* in, float* out, int N, int H, int W, const int* spans, int kh, int anchor_y, int anchor_x,
              void* stream) {
  SPB_CHECK_ARG(in && out && spans && in != out);
  SPB_CHECK_ARG(N >= 0 && H >= 1 && W >= 1 && kh >= 1 && kh <= 256 && anchor_y >= 0 && anchor_y < kh && anchor_x >= 0);
  if (N == 0) return SPB_OK;
  const long long total = (long long)N * H * W;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  erode_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, N, H, W, spans, kh, anchor_y, anchor_x);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
