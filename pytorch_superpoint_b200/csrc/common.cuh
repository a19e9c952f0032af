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
