// NEXT row f-2: batched homography sampler for homography adaptation -- the step before the hot path.
// Replaces utils/homographies.py:12-141 (sample_homography_np) as called by datasets/Coco.py:262-276:
//   homographies[i] = inv(sample_homography_np((2,2), shift=-1, **params)), homographies[0] = I,
//   inv_homographies = inverse(homographies)
// The reference builds a scipy truncnorm object per draw (1.4 s per 100 homographies on the CPU) -- at several
// hundred images/s that would be the end-to-end bottleneck.  Here: one thread per homography, Philox counters
// (curand), truncated normals by rejection (|z| <= 2, exact distribution), the same sequence of perturbations
// (perspective, scale candidates, translation, rotation candidates, incl. the reference's quirk that with
// allow_artifacts the last sampled scale and the appended zero angle are never picked), the 8x8 system of
// cv2.getPerspectiveTransform solved in double.  Parity is DISTRIBUTIONAL (different RNG stream), tested against
// draws of the unmodified reference (tests/golden/sampler_draws.npz).
#include <curand_kernel.h>

#include "common.cuh"

namespace spb {

struct SamplerParams {
  int perspective, scaling, rotation, translation, n_scales, n_angles, allow_artifacts;
  double scaling_amplitude, perspective_amplitude_x, perspective_amplitude_y, patch_ratio, max_angle,
      translation_overflow;
};

__device__ __forceinline__ double trunc_normal(curandStatePhilox4_32_10_t* st, double loc, double scale) {
  double z;
  do {
    z = curand_normal_double(st);
  } while (fabs(z) > 2.0);
  return loc + scale * z;
}
__device__ __forceinline__ double uniform_between(curandStatePhilox4_32_10_t* st, double lo, double hi) {
  return lo + (hi - lo) * (1.0 - curand_uniform_double(st));  // [lo, hi) like numpy.random.uniform
}
__device__ __forceinline__ int randint(curandStatePhilox4_32_10_t* st, int n) {  // uniform on 0..n-1
  const int k = (int)((1.0 - curand_uniform_double(st)) * n);
  return k < n ? k : n - 1;
}
__device__ __forceinline__ bool inside_unit(const double (&p)[4][2]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) ok = ok && p[i][0] >= 0.0 && p[i][0] < 1.0 && p[i][1] >= 0.0 && p[i][1] < 1.0;
  return ok;
}

// H with dst ~ H.src for the 4 point pairs: 8x8 Gaussian elimination with partial pivoting (double)
__device__ bool perspective_transform(const double (&s)[4][2], const double (&d)[4][2], double (&H)[9]) {
  double A[8][9];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double x = s[i][0], y = s[i][1], u = d[i][0], v = d[i][1];
    const double r0[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
    const double r1[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
    for (int j = 0; j < 9; ++j) {
      A[i][j] = r0[j];
      A[i + 4][j] = r1[j];
    }
  }
  for (int c = 0; c < 8; ++c) {
    int piv = c;
    for (int r = c + 1; r < 8; ++r)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (fabs(A[piv][c]) < 1e-14) return false;
    if (piv != c)
      for (int j = 0; j < 9; ++j) {
        const double t = A[c][j];
        A[c][j] = A[piv][j];
        A[piv][j] = t;
      }
    for (int r = c + 1; r < 8; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int j = c; j < 9; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int c = 7; c >= 0; --c) {
    double acc = A[c][8];
    for (int j = c + 1; j < 8; ++j) acc -= A[c][j] * H[j];
    H[c] = acc / A[c][c];
  }
  H[8] = 1.0;
  return true;
}

__device__ __forceinline__ bool invert3(const double (&m)[9], double (&o)[9]) {
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
  if (fabs(det) < 1e-300) return false;
  const double r = 1.0 / det;
  o[0] = c0 * r;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * r;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * r;
  o[3] = c1 * r;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * r;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * r;
  o[6] = c2 * r;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * r;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * r;
  return true;
}

// raw != NULL additionally returns the sampled (un-inverted) homography -- what sample_homography_np returns
__global__ void __launch_bounds__(128) sample_homographies_kernel(unsigned long long seed, int total, int per_image,
                                                                 SamplerParams P, float* __restrict__ homographies,
                                                                 float* __restrict__ inv_homographies,
                                                                 float* __restrict__ raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)i, 0, &st);
  double H[9];
  for (int attempt = 0;; ++attempt) {
    const double margin = (1.0 - P.patch_ratio) / 2.0;
    double p[4][2] = {{margin, margin},
                      {margin, margin + P.patch_ratio},
                      {margin + P.patch_ratio, margin + P.patch_ratio},
                      {margin + P.patch_ratio, margin}};
    if (P.perspective) {
      double ax = P.perspective_amplitude_x, ay = P.perspective_amplitude_y;
      if (!P.allow_artifacts) {
        ax = fmin(ax, margin);
        ay = fmin(ay, margin);
      }
      const double dy = trunc_normal(&st, 0.0, ay / 2), dxl = trunc_normal(&st, 0.0, ax / 2),
                   dxr = trunc_normal(&st, 0.0, ax / 2);
      p[0][0] += dxl; p[0][1] += dy;
      p[1][0] += dxl; p[1][1] -= dy;
      p[2][0] += dxr; p[2][1] += dy;
      p[3][0] += dxr; p[3][1] -= dy;
    }
    if (P.scaling) {
      // candidates: scale 1 first, then n_scales truncated normals around 1 (utils/homographies.py:83-84)
      double sc[17];
      const int ns = P.n_scales < 16 ? P.n_scales : 16;
      sc[0] = 1.0;
      for (int k = 0; k < ns; ++k) sc[k + 1] = trunc_normal(&st, 1.0, P.scaling_amplitude / 2);
      const double cx = (p[0][0] + p[1][0] + p[2][0] + p[3][0]) / 4, cy = (p[0][1] + p[1][1] + p[2][1] + p[3][1]) / 4;
      int valid[17], nv = 0;
      for (int k = 0; k <= ns; ++k) {
        if (P.allow_artifacts) {
          if (k < ns) valid[nv++] = k;  // np.arange(n_scales): the LAST sampled scale is never picked (reference quirk)
        } else {
          double q[4][2];
          for (int c = 0; c < 4; ++c) {
            q[c][0] = (p[c][0] - cx) * sc[k] + cx;
            q[c][1] = (p[c][1] - cy) * sc[k] + cy;
          }
          if (inside_unit(q)) valid[nv++] = k;
        }
      }
      if (nv == 0) continue;  // the reference would raise; redraw
      const double s = sc[valid[randint(&st, nv)]];
      for (int c = 0; c < 4; ++c) {
        p[c][0] = (p[c][0] - cx) * s + cx;
        p[c][1] = (p[c][1] - cy) * s + cy;
      }
    }
    if (P.translation) {
      double lo[2] = {1e300, 1e300}, hi[2] = {1e300, 1e300};
      for (int c = 0; c < 4; ++c)
        for (int a = 0; a < 2; ++a) {
          lo[a] = fmin(lo[a], p[c][a]);
          hi[a] = fmin(hi[a], 1.0 - p[c][a]);
        }
      if (P.allow_artifacts)
        for (int a = 0; a < 2; ++a) {
          lo[a] += P.translation_overflow;
          hi[a] += P.translation_overflow;
        }
      const double tx = uniform_between(&st, -lo[0], hi[0]), ty = uniform_between(&st, -lo[1], hi[1]);
      for (int c = 0; c < 4; ++c) {
        p[c][0] += tx;
        p[c][1] += ty;
      }
    }
    if (P.rotation) {
      // candidates: linspace(-max_angle, max_angle, n_angles) then 0 (utils/homographies.py:109-110)
      const int na = P.n_angles;
      const double cx = (p[0][0] + p[1][0] + p[2][0] + p[3][0]) / 4, cy = (p[0][1] + p[1][1] + p[2][1] + p[3][1]) / 4;
      auto angle = [&](int k) {
        if (k >= na) return 0.0;
        return na > 1 ? -P.max_angle + 2.0 * P.max_angle * k / (na - 1) : -P.max_angle;
      };
      auto rotate = [&](double a, double (&q)[4][2]) {  // (pts - c) @ [[cos, -sin], [sin, cos]] + c
        const double cs = cos(a), sn = sin(a);
        for (int c = 0; c < 4; ++c) {
          const double x = p[c][0] - cx, y = p[c][1] - cy;
          q[c][0] = x * cs + y * sn + cx;
          q[c][1] = -x * sn + y * cs + cy;
        }
      };
      int pick;
      if (P.allow_artifacts) {
        pick = randint(&st, na);  // np.arange(n_angles): the appended zero angle is never picked
      } else {
        int nv = 0;
        for (int k = 0; k <= na; ++k) {
          double q[4][2];
          rotate(angle(k), q);
          if (inside_unit(q)) ++nv;
        }
        if (nv == 0) continue;
        int want = randint(&st, nv);
        pick = na;
        for (int k = 0; k <= na; ++k) {
          double q[4][2];
          rotate(angle(k), q);
          if (inside_unit(q) && want-- == 0) {
            pick = k;
            break;
          }
        }
      }
      double q[4][2];
      rotate(angle(pick), q);
      for (int c = 0; c < 4; ++c) {
        p[c][0] = q[c][0];
        p[c][1] = q[c][1];
      }
    }
    // datasets/Coco.py:265: shape (2,2), shift -1 -> both point sets live in [-1,1]^2; cv2 takes float32 points
    double src[4][2] = {{-1, -1}, {-1, 1}, {1, 1}, {1, -1}}, dst[4][2];
    for (int c = 0; c < 4; ++c)
      for (int a = 0; a < 2; ++a) dst[c][a] = (double)(float)(p[c][a] * 2.0 - 1.0);
    if (perspective_transform(src, dst, H) || attempt >= 8) break;
  }
  double Hi[9];
  const bool identity = per_image > 0 && (i % per_image) == 0;  // Coco.py:270: homographies[0] = I
  if (identity || !invert3(H, Hi)) {
    for (int j = 0; j < 9; ++j) Hi[j] = (j % 4 == 0) ? 1.0 : 0.0;
  }
  float hf[9];
  double hd[9], back[9];
  for (int j = 0; j < 9; ++j) {
    hf[j] = (float)Hi[j];
    hd[j] = (double)hf[j];
    homographies[(long long)i * 9 + j] = hf[j];
    if (raw) raw[(long long)i * 9 + j] = (float)H[j];
  }
  // Coco.py:276: inverse of the float32 matrix
  if (!invert3(hd, back))
    for (int j = 0; j < 9; ++j) back[j] = (j % 4 == 0) ? 1.0 : 0.0;
  for (int j = 0; j < 9; ++j) inv_homographies[(long long)i * 9 + j] = (float)back[j];
}

}  // namespace spb

using namespace spb;

extern "C" int spb_sample_ha_homographies(unsigned long long seed, int total, int per_image, int perspective,
                                          int scaling, int rotation, int translation, int n_scales, int n_angles,
                                          double scaling_amplitude, double perspective_amplitude_x,
                                          double perspective_amplitude_y, double patch_ratio, double max_angle,
                                          int allow_artifacts, double translation_overflow, float* homographies,
                                          float* inv_homographies, float* raw_or_null, void* stream) {
  SPB_CHECK_ARG(total >= 0 && per_image >= 0 && homographies && inv_homographies);
  SPB_CHECK_ARG(n_scales >= 1 && n_scales <= 16 && n_angles >= 1 && patch_ratio > 0.0 && patch_ratio <= 1.0);
  if (total == 0) return SPB_OK;
  SamplerParams P;
  P.perspective = perspective;
  P.scaling = scaling;
  P.rotation = rotation;
  P.translation = translation;
  P.n_scales = n_scales;
  P.n_angles = n_angles;
  P.allow_artifacts = allow_artifacts;
  P.scaling_amplitude = scaling_amplitude;
  P.perspective_amplitude_x = perspective_amplitude_x;
  P.perspective_amplitude_y = perspective_amplitude_y;
  P.patch_ratio = patch_ratio;
  P.max_angle = max_angle;
  P.translation_overflow = translation_overflow;
  sample_homographies_kernel<<<ceil_div(total, 128), 128, 0, as_stream(stream)>>>(seed, total, per_image, P,
                                                                                 homographies, inv_homographies,
                                                                                 raw_or_null);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}
