// The non-tensor kernels of the fused homography-adaptation pass (tcgen05 product path):
//   ha_prep_kernel     pads the source images, folds the per-view homographies, paints the borders of the
//                      probability planes
//   warp_views_kernel  (1) N warped views of every image + the exact valid-mask bits
//                      (utils/utils.py:318-351, 677-704; datasets/Coco.py:279-282)
//   agg_heat_kernel    (3) mask * probability, bilinear un-warp and the two sums over the views
//                      (export.py:49-62) from the fp16 probability planes the detector head's epilogue wrote
//                      (conv_tc.cu, OUT == 4: softmax, dustbin drop, depth-to-space and the mask in one store)
//
// Both gathers were issue bound in round 1 (two IEEE divisions, per-tap validity tests, cell-major index
// arithmetic: ~150 and ~250 instructions per pixel, 8 % of the HBM roofline).  Here every view's homography is
// folded once (common.cuh: FoldMat), the frames carry a border so that clamped coordinates replace the per-tap
// tests, floor() is one magic-number addition and the planes are row-major: ~50 / ~65 instructions per pixel.
// A warp works on an 8 x 4 pixel block: with rotations of up to +-90 degrees that keeps the number of cache lines a
// gather instruction touches at 4..9 (a 32 x 1 row: ~20), the quantity the L1 tag stage bounds these kernels by.
// The valid-mask bit is still EXACT: decided from the folded coordinate when that is further than an epsilon from
// the decision boundary, re-evaluated with the reference's rounded fp32 expression (common.cuh: nearest_valid)
// otherwise (a few pixels per view).
#include <cuda_fp16.h>

#include "net.cuh"

namespace spb {

// ------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------
struct PrepArgs {
  const float* imgs;      // [nb,H,W]
  float* src_pad;         // [nb,H+2P,W+2P]
  const float* mats_fwd;  // indexed by ViewMap::mat
  const float* mats_unwarp;
  FoldMat* fold_fwd;      // [V]
  FoldMat* fold_unwarp;   // [V]
  __half* heat16;         // [V,H+2*kHeatPadY,W+2*kHeatPadX]
  ViewMap vm;
  int nb, V, H, W, planes;  // planes: probability planes whose border is painted (V, or 0)
};

__global__ void __launch_bounds__(256) ha_prep_kernel(PrepArgs a) {
  const long long tid = blockIdx.x * 256LL + threadIdx.x, nth = gridDim.x * 256LL;
  // (a) zero-padded copies of the source images
  const int Wp = a.W + 2 * kHaPad, Hp = a.H + 2 * kHaPad;
  const long long npad = (long long)a.nb * Hp * Wp;
  for (long long i = tid; i < npad; i += nth) {
    const int x = (int)(i % Wp) - kHaPad, y = (int)((i / Wp) % Hp) - kHaPad;
    const int b = (int)(i / ((long long)Wp * Hp));
    const bool in = x >= 0 && x < a.W && y >= 0 && y < a.H;
    a.src_pad[i] = in ? __ldg(a.imgs + ((long long)b * a.H + y) * a.W + x) : 0.f;
  }
  // (b) folded homographies of every view
  for (long long v = tid; v < 2LL * a.V; v += nth) {
    const int view = (int)(v >> 1);
    const int mi = a.vm.mat(view);
    if (v & 1)
      a.fold_unwarp[view] = fold_homography(a.mats_unwarp + 9LL * mi, a.W, a.H, kHeatPadX, kHeatPadY);
    else
      a.fold_fwd[view] = fold_homography(a.mats_fwd + 9LL * mi, a.W, a.H, kHaPad, kHaPad);
  }
  // (c) border of the probability planes = "masked": 2 columns left and right of every row, 2 rows above and below
  const int Wh = a.W + 2 * kHeatPadX, Hh = a.H + 2 * kHeatPadY;
  const int per_view = 2 * kHeatPadY * (a.W + 4) + a.H * 4;
  const __half masked = __ushort_as_half(kHeatMasked);
  for (long long i = tid; i < (long long)a.planes * per_view; i += nth) {
    const int view = (int)(i / per_view);
    int r = (int)(i % per_view), x, y;
    if (r < 2 * kHeatPadY * (a.W + 4)) {
      const int row = r / (a.W + 4);
      y = row < kHeatPadY ? row : a.H + row;  // rows 0,1 and H+2,H+3
      x = kHeatPadX - 2 + r % (a.W + 4);
    } else {
      r -= 2 * kHeatPadY * (a.W + 4);
      y = kHeatPadY + (r >> 2);
      x = (r & 2) ? a.W + kHeatPadX + (r & 1) : kHeatPadX - 2 + (r & 1);
    }
    a.heat16[((long long)view * Hh + y) * Wh + x] = masked;
  }
}

int ha_prep(const float* imgs, float* src_pad, const float* mats_fwd, const float* mats_unwarp, FoldMat* fold_fwd,
            FoldMat* fold_unwarp, void* heat16, const ViewMap& vm, int nb, int V, int H, int W, cudaStream_t st,
            int planes) {
  PrepArgs a;
  a.imgs = imgs;
  a.src_pad = src_pad;
  a.mats_fwd = mats_fwd;
  a.mats_unwarp = mats_unwarp;
  a.fold_fwd = fold_fwd;
  a.fold_unwarp = fold_unwarp;
  a.heat16 = static_cast<__half*>(heat16);
  a.vm = vm;
  a.nb = nb;
  a.V = V;
  a.H = H;
  a.W = W;
  a.planes = planes < 0 ? V : planes;
  const long long work = (long long)nb * (H + 2 * kHaPad) * (W + 2 * kHaPad);
  const long long border = (long long)a.planes * (2 * kHeatPadY * (W + 4) + H * 4);
  int blocks = ceil_div(work > border ? work : border, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  ha_prep_kernel<<<blocks, 256, 0, st>>>(a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// (1) warped views + exact valid-mask bits
// ------------------------------------------------------------------------------------------------
struct WarpArgs {
  const float* src_pad;     // [nb,Hp,Wp]
  const FoldMat* fold;      // [V]
  const float* mats_fwd;    // exact re-evaluation of ambiguous mask bits
  float* out;               // [V,H,W]
  uint8_t* bits;            // [V,H,W/8]
  ViewMap vm;
  int H, W;
  Lin lx, ly;
};

// A warp covers 8 x 4 output pixels (not 32 x 1): the sampler draws rotations up to +-90 degrees, under which a
// 32-pixel row of the output maps to a source segment crossing ~20 rows = ~20 cache lines per load instruction, and
// the round-2 ncu capture showed the first version of this kernel bound by the L1 tag stage (l1tex 70-85 %), not by
// issue slots.  An 8 x 4 block touches 4..9 source rows at every angle.
constexpr int kWarpRows = 32;  // rows per CTA: 8 warps x 4 rows, the CTA walks along x

__global__ void __launch_bounds__(256) warp_views_kernel(WarpArgs a) {
  const int view = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = lane & 7, ly = lane >> 3;
  const FoldMat f = a.fold[view];
  const int Wp = a.W + 2 * kHaPad, Hp = a.H + 2 * kHaPad;
  const float* src = a.src_pad + (long long)a.vm.img(view) * Hp * Wp;
  asm volatile("" : "+l"(src));  // opaque base: every tap address is ONE 32 x 32 + 64-bit multiply-add
  // bilinear taps live in [PAD-1, W+PAD] (padded pixel units, +0.5); a valid mask bit needs the sample inside
  // [PAD-0.5, W+PAD-0.5) before the +0.5, i.e. s in [PAD, W+PAD)
  const float clx = kHaPad - 0.5f, chx = (float)a.W + kHaPad + 0.5f;
  const float cly = kHaPad - 0.5f, chy = (float)a.H + kHaPad + 0.5f;
  const float eps = f.r0;  // half-width of the band in which the mask bit is re-evaluated exactly (per view)
  const float ilx = kHaPad + eps, ihx = (float)a.W + kHaPad - eps;  // surely valid inside
  const float ily = kHaPad + eps, ihy = (float)a.H + kHaPad - eps;
  const float olx = kHaPad - eps, ohx = (float)a.W + kHaPad + eps;  // surely invalid outside
  const float oly = kHaPad - eps, ohy = (float)a.H + kHaPad + eps;
  const unsigned kmagic = 0x4B000000u * (unsigned)(Wp + 1);  // (mod 2^32: the offsets below wrap consistently)
  const int Wb = a.W >> 3;
  const int y = blockIdx.x * kWarpRows + warp * 4 + ly;
  const bool row_ok = y < a.H;
  const int yc = row_ok ? y : a.H - 1;
  const float yf = (float)yc;
  const float nx0 = fmaf(f.ax1, yf, f.ax2), ny0 = fmaf(f.ay1, yf, f.ay2), nz0 = fmaf(f.z1, yf, f.z2);
  float* orow = a.out + ((long long)view * a.H + yc) * a.W;
  uint8_t* brow = a.bits + ((long long)view * a.H + yc) * Wb;
  if (blockIdx.x * kWarpRows + warp * 4 >= a.H) return;  // (warp-uniform: no row of this warp exists)
#pragma unroll 1
  for (int xc = 0; xc < a.W; xc += 16) {
    float val[2];
    unsigned word[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int x = xc + 8 * j + lx;  // (W % 8 == 0: a column block of 8 is inside the image or entirely outside)
      const float xf = (float)x;
      const float r = rcp_approx(fmaf(f.z0, xf, nz0));
      const float sx = fmaf(f.ax0, xf, nx0) * r, sy = fmaf(f.ay0, xf, ny0) * r;
      const Tap1 tx = tap_axis(sx, clx, chx), ty = tap_axis(sy, cly, chy);
      const unsigned off = ty.bits * (unsigned)Wp + tx.bits - kmagic;  // south-east tap: row, column >= 2
      const float *p = src + off, *q0 = src + (off - (unsigned)Wp);
      const float top = fmaf(__ldg(q0), tx.w1, __ldg(q0 - 1) * tx.w0);
      const float bot = fmaf(__ldg(p), tx.w1, __ldg(p - 1) * tx.w0);
      val[j] = fmaf(bot, ty.w1, top * ty.w0);
      bool in = sx >= ilx && sx <= ihx && sy >= ily && sy <= ihy;
      const bool maybe = sx >= olx && sx <= ohx && sy >= oly && sy <= ohy;  // NaN -> false -> invalid (as exact)
      const bool live = row_ok && x < a.W;
      if (__any_sync(0xffffffffu, maybe && !in && live)) {
        if (maybe && !in && live)
          in = nearest_valid(load_mat(a.mats_fwd + 9LL * a.vm.mat(view)), a.lx.at(x), a.ly.at(y), a.W, a.H);
      }
      word[j] = __ballot_sync(0xffffffffu, in && live);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int x = xc + 8 * j + lx;
      if (row_ok && x < a.W) {
        orow[x] = val[j];
        if (lx == 0) brow[x >> 3] = (uint8_t)(word[j] >> (ly * 8));  // byte ly of the ballot = 8 pixels of row ly
      }
    }
  }
}

int warp_views(const float* src_pad, const FoldMat* fold, const float* mats_fwd, float* out, uint8_t* bits,
               const ViewMap& vm, int V, int H, int W, cudaStream_t st) {
  if (V > 65535) {
    set_error("warp_views: %d views in one launch (max 65535)", V);
    return SPB_EINVAL;
  }
  WarpArgs a;
  a.src_pad = src_pad;
  a.fold = fold;
  a.mats_fwd = mats_fwd;
  a.out = out;
  a.bits = bits;
  a.vm = vm;
  a.H = H;
  a.W = W;
  a.lx = Lin(W);
  a.ly = Lin(H);
  dim3 grid(ceil_div(H, kWarpRows), V);
  warp_views_kernel<<<grid, 256, 0, st>>>(a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// (3) un-warp + aggregate from the fp16 probability planes
// ------------------------------------------------------------------------------------------------
// heat16[v][y + kHeatPadY][x + kHeatPadX] = fp16 probability of warped pixel (x, y) of view v, or -1 when that
// pixel lies outside the source image (mask 0) or in the border.  One thread per output pixel and view group:
// num += w * p, den += w over the unmasked taps, views n = begin + g, begin + g + G, ... in a fixed order;
// agg_reduce_kernel (heat.cu) then adds the G partial planes in order -> deterministic.
struct AggArgs {
  const __half* heat16;
  const FoldMat* fold;      // un-warp matrices, folded for the heat16 frame
  const int* offs;          // image b owns views [offs[b], offs[b+1])  (NULL: b*N .. (b+1)*N)
  float* part;              // [nb*G,2,H,W]
  int N, H, W, G;
};

constexpr int kAggFoldStage = 128;  // folded matrices staged per round in shared memory

// CTA tile 32 x 8 output pixels, every warp an 8 x 4 block (see warp_views_kernel: the gather is bound by the number
// of cache lines a load instruction touches, and un-warps rotate by up to +-90 degrees).  The folded matrices of the
// CTA's views are staged in shared memory (three broadcast LDS instead of three L1 look-ups per view and thread).
__global__ void __launch_bounds__(256) agg_heat_kernel(AggArgs a) {
  __shared__ float4 s_fold[kAggFoldStage * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7), y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
  const int b = blockIdx.z / a.G, g = blockIdx.z % a.G;
  int n_begin, n_end;
  if (a.offs) {
    n_begin = __ldg(a.offs + b) + g;
    n_end = __ldg(a.offs + b + 1);
  } else {
    n_begin = b * a.N + g;
    n_end = (b + 1) * a.N;
  }
  const int Wh = a.W + 2 * kHeatPadX, Hh = a.H + 2 * kHeatPadY;
  const float clx = kHeatPadX - 0.5f, chx = (float)a.W + kHeatPadX + 0.5f;
  const float cly = kHeatPadY - 0.5f, chy = (float)a.H + kHeatPadY + 0.5f;
  const unsigned kmagic = 0x4B000000u * (unsigned)(Wh + 1);
  const float xf = (float)x, yf = (float)y;
  float num = 0.f, den = 0.f;
  const long long plane_sz = (long long)Hh * Wh;
  const int nviews = n_end > n_begin ? (n_end - n_begin + a.G - 1) / a.G : 0;  // views n_begin + i * G of this CTA
  for (int i0 = 0; i0 < nviews; i0 += kAggFoldStage) {
    const int cnt = min(kAggFoldStage, nviews - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += 256)
      s_fold[e] = __ldg(reinterpret_cast<const float4*>(a.fold + n_begin + (long long)(i0 + e / 3) * a.G) + e % 3);
    __syncthreads();
    // running pointer: the view's plane (the tap offsets carry the magic-number bias, removed by kmagic)
    const unsigned short* plane =
        reinterpret_cast<const unsigned short*>(a.heat16) + (n_begin + (long long)i0 * a.G) * plane_sz;
    const long long plane_step = (long long)a.G * plane_sz;
#pragma unroll 2
    for (int i = 0; i < cnt; ++i, plane += plane_step) {
      const float4 f0 = s_fold[3 * i], f1 = s_fold[3 * i + 1];
      const float z2 = s_fold[3 * i + 2].x;
      const float r = rcp_approx(fmaf(f1.z, xf, fmaf(f1.w, yf, z2)));
      const float sx = fmaf(f0.x, xf, fmaf(f0.y, yf, f0.z)) * r;
      const float sy = fmaf(f0.w, xf, fmaf(f1.x, yf, f1.y)) * r;
      const Tap1 tx = tap_axis(sx, clx, chx), ty = tap_axis(sy, cly, chy);
      const unsigned off = ty.bits * (unsigned)Wh + tx.bits - kmagic;  // south-east tap: row >= 2, column >= 8
      const unsigned short *p = plane + off, *q0 = plane + (off - (unsigned)Wh);
      const float w00 = tx.w0 * ty.w0, w01 = tx.w1 * ty.w0, w10 = tx.w0 * ty.w1, w11 = tx.w1 * ty.w1;
      auto tap = [&](const unsigned short* q, float w) {
        const float pv = __half2float(__ushort_as_half(__ldg(q)));
        if (pv >= 0.f) {
          num = fmaf(pv, w, num);
          den += w;
        }
      };
      tap(q0 - 1, w00);
      tap(q0, w01);
      tap(p - 1, w10);
      tap(p, w11);
    }
  }
  if (x < a.W && y < a.H) {
    float* o = a.part + (long long)blockIdx.z * 2 * a.H * a.W + (long long)y * a.W + x;
    o[0] = num;
    o[(long long)a.H * a.W] = den;
  }
}

int aggregate_heat16(const void* heat16, const FoldMat* fold_unwarp, int nb, int N, int H, int W, float* sums,
                     int accumulate, void* ws, cudaStream_t st, const int* offs) {
  const int G = aggregate_probs_groups(nb, N, H, W);  // (ragged: N = the largest per-image view count)
  AggArgs a;
  a.heat16 = static_cast<const __half*>(heat16);
  a.fold = fold_unwarp;
  a.offs = offs;
  a.part = static_cast<float*>(ws);
  a.N = N;
  a.H = H;
  a.W = W;
  a.G = G;
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), G * nb);
  agg_heat_kernel<<<grid, 256, 0, st>>>(a);
  SPB_CHECK_LAUNCH();
  return reduce_partials(a.part, G, nb, H, W, sums, accumulate, st);
}

}  // namespace spb
