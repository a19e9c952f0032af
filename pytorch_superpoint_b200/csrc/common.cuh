// Shared helpers for the spb200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spb200.h"

namespace spb {

constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);

#define SPB_CHECK_ARG(cond)                                                            \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      spb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);        \
      return SPB_EINVAL;                                                               \
    }                                                                                  \
  } while (0)

#define SPB_CHECK_CUDA(expr)                                                           \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      spb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SPB_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (bench.py's gpu_launches)
#define SPB_CHECK_LAUNCH()                                   \
  do {                                                       \
    __atomic_fetch_add(&spb::g_launches, 1ull, __ATOMIC_RELAXED); \
    SPB_CHECK_CUDA(cudaGetLastError());                      \
  } while (0)

// ---- coordinates ---------------------------------------------------------------------------------
// Bit-exact restatement of the reference's coordinate pipeline (utils/utils.py:286-314,341-344 +
// grid_sample's align_corners=True un-normalisation); mirrors oracle/oracle.py::warp_grid.
// Every operation is an explicitly rounded fp32 op so that nvcc cannot contract or reorder it.

struct Lin {  // torch.linspace(-1, 1, n)[i]
  float step;
  int n;
  __host__ __device__ Lin() : step(0.f), n(1) {}
  __host__ Lin(int n_) : step(n_ > 1 ? 2.0f / (float)(n_ - 1) : 0.f), n(n_) {}
  __device__ __forceinline__ float at(int i) const {
    if (n == 1) return -1.f;
    return (i < n / 2) ? __fmaf_rn(step, (float)i, -1.f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.f);
  }
};

struct Mat3 {
  float m[9];
};

__device__ __forceinline__ Mat3 load_mat(const float* __restrict__ p) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = __ldg(p + i);
  return r;
}

__device__ __forceinline__ float hrow(const Mat3& M, int r, float u, float v) {
  return __fadd_rn(__fadd_rn(__fmul_rn(M.m[3 * r], u), __fmul_rn(M.m[3 * r + 1], v)), M.m[3 * r + 2]);
}

// source pixel coordinates of normalised point (u, v) under M, for a W x H source
__device__ __forceinline__ void warp_coord(const Mat3& M, float u, float v, int W, int H, float& ix, float& iy) {
  const float X = hrow(M, 0, u, v), Y = hrow(M, 1, u, v), Z = hrow(M, 2, u, v);
  const float gx = __fdiv_rn(X, Z), gy = __fdiv_rn(Y, Z);
  // (g + 1) / 2: halving is exact in binary floating point (and rounds identically in the subnormal range), so the
  // multiplication by 0.5 is bit-identical to the reference's division and saves two IEEE division sequences
  ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(W - 1));
  iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(H - 1));
}

__device__ __forceinline__ bool in_bounds(float xf, float yf, int W, int H) {
  return xf >= 0.f && xf <= (float)(W - 1) && yf >= 0.f && yf <= (float)(H - 1);  // NaN -> false
}

// nearest-mode validity (utils/utils.py:696 on an all-ones image): nearbyint is round-half-even
__device__ __forceinline__ bool nearest_valid(const Mat3& M, float u, float v, int W, int H) {
  float ix, iy;
  warp_coord(M, u, v, W, H, ix, iy);
  return in_bounds(rintf(ix), rintf(iy), W, H);
}

// Bilinear tap set, weights in ATen's order nw, ne, sw, se.
struct Taps {
  float x0, y0, x1, y1, w_nw, w_ne, w_sw, w_se;
};
__device__ __forceinline__ Taps make_taps(float ix, float iy) {
  Taps t;
  t.x0 = floorf(ix);
  t.y0 = floorf(iy);
  t.x1 = __fadd_rn(t.x0, 1.f);
  t.y1 = __fadd_rn(t.y0, 1.f);
  const float ax = __fsub_rn(t.x1, ix), bx = __fsub_rn(ix, t.x0);
  const float ay = __fsub_rn(t.y1, iy), by = __fsub_rn(iy, t.y0);
  t.w_nw = __fmul_rn(ax, ay);
  t.w_ne = __fmul_rn(bx, ay);
  t.w_sw = __fmul_rn(ax, by);
  t.w_se = __fmul_rn(bx, by);
  return t;
}

template <class F>  // F(int x, int y) -> float value of an in-bounds tap
__device__ __forceinline__ float bilinear(const Taps& t, int W, int H, F&& fetch) {
  float acc = 0.f;
  if (in_bounds(t.x0, t.y0, W, H)) acc = __fadd_rn(acc, __fmul_rn(fetch((int)t.x0, (int)t.y0), t.w_nw));
  if (in_bounds(t.x1, t.y0, W, H)) acc = __fadd_rn(acc, __fmul_rn(fetch((int)t.x1, (int)t.y0), t.w_ne));
  if (in_bounds(t.x0, t.y1, W, H)) acc = __fadd_rn(acc, __fmul_rn(fetch((int)t.x0, (int)t.y1), t.w_sw));
  if (in_bounds(t.x1, t.y1, W, H)) acc = __fadd_rn(acc, __fmul_rn(fetch((int)t.x1, (int)t.y1), t.w_se));
  return acc;
}

// The same 4-tap sample with the shared work hoisted: two column and two row validity tests, two float->int
// conversions, one row-base multiply.  Identical arithmetic (same products, same order of the additions, invalid
// taps skipped), so it is bit-identical to bilinear(make_taps(..)).  fetch(int offset) reads element y*W + x.
template <class F>
__device__ __forceinline__ float bilinear_fast(const Taps& t, int W, int H, F&& fetch) {
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  const bool vx0 = t.x0 >= 0.f && t.x0 <= wm, vx1 = t.x1 >= 0.f && t.x1 <= wm;
  const bool vy0 = t.y0 >= 0.f && t.y0 <= hm, vy1 = t.y1 >= 0.f && t.y1 <= hm;
  // clamp before converting so that the offsets of skipped taps stay finite; only valid taps are fetched
  const int xi = (int)fminf(fmaxf(t.x0, -1.f), wm), yi = (int)fminf(fmaxf(t.y0, -1.f), hm);
  const int o = yi * W + xi;
  float acc = 0.f;
  if (vx0 && vy0) acc = __fadd_rn(acc, __fmul_rn(fetch(o), t.w_nw));
  if (vx1 && vy0) acc = __fadd_rn(acc, __fmul_rn(fetch(o + 1), t.w_ne));
  if (vx0 && vy1) acc = __fadd_rn(acc, __fmul_rn(fetch(o + W), t.w_sw));
  if (vx1 && vy1) acc = __fadd_rn(acc, __fmul_rn(fetch(o + W + 1), t.w_se));
  return acc;
}

// View v of a pass -> its source image (relative to the pass's first image) and its matrix (index into the matrix
// arrays the caller passed): explicit per-view index arrays (homography sharding: rotated view blocks, a different
// number of views per image) or arithmetic (vpi views per image; matrices shared by the images or image-major).
struct ViewMap {
  const int* view_image;  // device int[V] or NULL
  const int* view_mat;    // device int[V] or NULL
  int vpi, shared;
  __device__ __forceinline__ int img(int v) const { return view_image ? __ldg(view_image + v) : v / vpi; }
  __device__ __forceinline__ int mat(int v) const {
    return view_mat ? __ldg(view_mat + v) : (shared ? v % vpi : v);
  }
};

// ---- fast coordinate pipeline of the fused HA path (ha_fast.cu) -----------------------------------------------
// The generic operators above evaluate the reference's coordinate expression with one rounding per operation (bit
// exact against the oracle).  Inside the fused HA path the warped views feed a bf16 network and the un-warped
// probabilities a 1e-2 heat-map tolerance, so there the homography is folded ONCE per view (in double) into
//   s = (a0*x + a1*y + a2) / (c0*x + c1*y + c2)        x, y = integer output pixel
// which yields the sample position in a zero-PADDED source frame, already offset by +0.5 for the rounding trick
// (kHaPad border pixels on every side: no per-tap validity tests, coordinates are clamped into the border).
// Only the valid-mask bit stays exact: it is decided from s when s is further than an epsilon from a decision
// boundary and re-evaluated with the exact pipeline otherwise.
constexpr int kHaPad = 2;          // zero border of the padded fp32 source images (pixels, every side)
constexpr int kHeatPadX = 8;       // left/right border of the fp16 probability planes (16-byte aligned rows)
constexpr int kHeatPadY = 2;       // top/bottom border rows of the fp16 probability planes
constexpr unsigned short kHeatMasked = 0xBC00u;  // fp16 -1: "this warped pixel is outside the source image"

struct __align__(16) FoldMat {
  float ax0, ax1, ax2, ay0;
  float ay1, ay2, z0, z1;
  float z2, r0, r1, r2;
};

// padx / pady: position of pixel (0, 0) inside the padded frame
__device__ __forceinline__ FoldMat fold_homography(const float* __restrict__ m, int W, int H, int padx, int pady) {
  const double sx = W > 1 ? 2.0 / (double)(W - 1) : 0.0, sy = H > 1 ? 2.0 / (double)(H - 1) : 0.0;
  double M[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) M[i] = (double)m[i];
  // u = sx*x - 1, v = sy*y - 1  ->  row_i . (u, v, 1) = (M[i0]*sx) x + (M[i1]*sy) y + (M[i2] - M[i0] - M[i1])
  double P[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    P[3 * r] = M[3 * r] * sx;
    P[3 * r + 1] = M[3 * r + 1] * sy;
    P[3 * r + 2] = M[3 * r + 2] - M[3 * r] - M[3 * r + 1];
  }
  const double cx = 0.5 * (double)(W - 1), cy = 0.5 * (double)(H - 1);
  const double ox = (double)padx + 0.5, oy = (double)pady + 0.5;
  FoldMat f;
  f.ax0 = (float)(cx * (P[0] + P[6]) + ox * P[6]);
  f.ax1 = (float)(cx * (P[1] + P[7]) + ox * P[7]);
  f.ax2 = (float)(cx * (P[2] + P[8]) + ox * P[8]);
  f.ay0 = (float)(cy * (P[3] + P[6]) + oy * P[6]);
  f.ay1 = (float)(cy * (P[4] + P[7]) + oy * P[7]);
  f.ay2 = (float)(cy * (P[5] + P[8]) + oy * P[8]);
  f.z0 = (float)P[6];
  f.z1 = (float)P[7];
  f.z2 = (float)P[8];
  // r0: half-width (pixels) of the band around a valid-mask decision boundary inside which the bit is re-evaluated
  // exactly.  The folded and the reference expression differ by a few 1e-4 pixels while the denominator is O(1);
  // cancellation in the denominator amplifies that by (sum of its term magnitudes) / |Z|, so the band grows with
  // it, and a view whose denominator changes sign inside the frame is evaluated exactly everywhere.
  const double zc[4] = {P[8], P[6] * (W - 1) + P[8], P[7] * (H - 1) + P[8], P[6] * (W - 1) + P[7] * (H - 1) + P[8]};
  double zmin = fabs(zc[0]);
  bool flip = false;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    zmin = fmin(zmin, fabs(zc[i]));
    flip = flip || (zc[i] > 0.0) != (zc[0] > 0.0);
  }
  const double zmag = fabs(P[6]) * (W - 1) + fabs(P[7]) * (H - 1) + fabs(P[8]);
  const double base = 2e-3 + 8e-6 * (double)(W > H ? W : H);
  const double amp = (zmin > 0.0 && !flip) ? fmax(1.0, zmag / zmin) : 1e30;
  f.r0 = (float)fmin(base * amp, 1e30);
  if (!(f.r0 == f.r0)) f.r0 = 1e30f;  // NaN matrix
  f.r1 = f.r2 = 0.f;
  return f;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// s in [lo, hi] (already offset by +0.5) -> n = floor(s - 0.5) + 1 as the low bits of a float in [2^23, 2^24) and the
// two linear weights of the taps n - 1 and n.  Ties (s - 0.5 an exact integer) may round either way; both choices
// give the same interpolated value.
struct Tap1 {
  unsigned bits;  // 0x4B000000 + n
  float w0, w1;
};
__device__ __forceinline__ Tap1 tap_axis(float s, float lo, float hi) {
  const float c = fminf(fmaxf(s, lo), hi);  // NaN -> lo
  const float t = __fadd_rn(c, 8388608.f);
  const float d = __fsub_rn(c, __fsub_rn(t, 8388608.f));
  Tap1 r;
  r.bits = __float_as_uint(t);
  r.w1 = d + 0.5f;
  r.w0 = 0.5f - d;
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace spb
