// (3) Heat-map kernels: 65-way cell softmax + dustbin drop + depth-to-space (utils/utils.py:480-525,
// utils/d2s.py:8-25), valid mask (utils/utils.py:677-704), bilinear un-warp and aggregation over
// the N homographies (export.py:49-63).
//
// spb_ha_aggregate is the fused product kernel: per (16x16 output tile, view) it finds the exact
// bounding box of the bilinear taps in the warped frame (block min/max), stages the covering
// cells' logits once in shared memory, turns them into probability*mask, and gathers.  The N
// views are split into G groups over blockIdx.y; each CTA accumulates its views in registers in
// a fixed order and writes one partial; a second tiny kernel adds the G partials in order, so the
// result is deterministic (no atomics).  HBM traffic: N*65*Hc*Wc*4 B read once + 2*H*W*4*(G+1).
#include "net.cuh"

namespace spb {

// ------------------------------------------------------------------------------------------------
// flattenDetection
// ------------------------------------------------------------------------------------------------
// NHWC: one warp per cell, coalesced 260-byte reads.
__global__ void __launch_bounds__(256) flatten_nhwc_kernel(const float* __restrict__ semi, int cs,
                                                           float* __restrict__ heat, int B, int Hc, int Wc) {
  const int lane = threadIdx.x & 31;
  const long long ncell = (long long)B * Hc * Wc;
  const int W = Wc * 8, H = Hc * 8;
  for (long long c = blockIdx.x * 8LL + (threadIdx.x >> 5); c < ncell; c += gridDim.x * 8LL) {
    const float* p = semi + c * cs;
    const float a = p[lane], b = p[lane + 32], d = (lane == 0) ? p[64] : -INFINITY;
    const float mx = warp_max(fmaxf(fmaxf(a, b), d));
    const float ea = expf(a - mx), eb = expf(b - mx), ed = (lane == 0) ? expf(d - mx) : 0.f;
    const float s = warp_sum(ea + eb + ed);
    const int cx = (int)(c % Wc), cy = (int)((c / Wc) % Hc), n = (int)(c / ((long long)Wc * Hc));
    float* o = heat + ((long long)n * H + cy * 8) * W + cx * 8;
    o[(lane >> 3) * W + (lane & 7)] = __fdiv_rn(ea, s);
    o[((lane >> 3) + 4) * W + (lane & 7)] = __fdiv_rn(eb, s);
  }
}

// NCHW (reference layout): one thread per cell, channel loop (coalesced across cells).
__global__ void __launch_bounds__(256) flatten_nchw_kernel(const float* __restrict__ semi, float* __restrict__ heat,
                                                           int B, int Hc, int Wc) {
  const long long plane = (long long)Hc * Wc, ncell = (long long)B * plane;
  const int W = Wc * 8, H = Hc * 8;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
       c += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(c / plane);
    const long long r = c % plane;
    const float* p = semi + (long long)n * 65 * plane + r;
    float mx = -INFINITY;
    for (int k = 0; k < 65; ++k) mx = fmaxf(mx, __ldg(p + k * plane));
    float s = 0.f;
    for (int k = 0; k < 65; ++k) s += expf(__ldg(p + k * plane) - mx);
    const int cx = (int)(r % Wc), cy = (int)(r / Wc);
    float* o = heat + ((long long)n * H + cy * 8) * W + cx * 8;
    for (int k = 0; k < 64; ++k) o[(k >> 3) * W + (k & 7)] = __fdiv_rn(expf(__ldg(p + k * plane) - mx), s);
  }
}

// ------------------------------------------------------------------------------------------------
// combine_heatmap on materialised heat / mask (reference API; sequential sum over n like torch.sum)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const float* __restrict__ heat, const float* __restrict__ mask,
                                                      const float* __restrict__ mats, float* __restrict__ out, int N,
                                                      int H, int W, Lin lx, Lin ly) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float u = lx.at(x), v = ly.at(y);
  float num = 0.f, den = 0.f;
  for (int n = 0; n < N; ++n) {
    const Mat3 M = load_mat(mats + 9 * n);
    float ix, iy;
    warp_coord(M, u, v, W, H, ix, iy);
    const Taps tp = make_taps(ix, iy);
    const float* h = heat + (long long)n * H * W;
    const float* m = mask + (long long)n * H * W;
    num = __fadd_rn(num, bilinear(tp, W, H, [&](int xx, int yy) {
                      return __fmul_rn(__ldg(h + yy * W + xx), __ldg(m + yy * W + xx));
                    }));
    den = __fadd_rn(den, bilinear(tp, W, H, [&](int xx, int yy) { return __ldg(m + yy * W + xx); }));
  }
  out[y * W + x] = __fdiv_rn(num, den);
}

// ------------------------------------------------------------------------------------------------
// fused softmax / d2s / mask / un-warp / aggregate
// ------------------------------------------------------------------------------------------------
constexpr int kAggTile = 16;
constexpr int kAggMaxCells = 64;

struct AggParams {
  const float* semi;
  int cs;
  const float* mu;  // un-warp matrices  (sample['homographies'])
  const float* mf;  // forward matrices  (sample['inv_homographies']): define the valid mask
  int N, H, W, Hc, Wc, G;
  float* part;  // [G,2,H,W]
  Lin lx, ly;
};

// probability*mask and mask of one warped-frame pixel straight from global memory (fallback path)
__device__ __noinline__ float2 prob_mask_global(const AggParams& p, int n, const Mat3& Mf, int xx, int yy) {
  const bool valid = nearest_valid(Mf, p.lx.at(xx), p.ly.at(yy), p.W, p.H);
  if (!valid) return make_float2(0.f, 0.f);
  const float* q = p.semi + (((long long)n * p.Hc + (yy >> 3)) * p.Wc + (xx >> 3)) * p.cs;
  float mx = -INFINITY;
  for (int k = 0; k < 65; ++k) mx = fmaxf(mx, __ldg(q + k));
  float s = 0.f;
  for (int k = 0; k < 65; ++k) s += expf(__ldg(q + k) - mx);
  return make_float2(__fdiv_rn(expf(__ldg(q + ((yy & 7) << 3 | (xx & 7))) - mx), s), 1.f);
}

__global__ void __launch_bounds__(256) ha_aggregate_kernel(AggParams p) {
  __shared__ float s_hm[kAggMaxCells * 64];
  __shared__ float s_m[kAggMaxCells * 64];
  __shared__ int s_red[4][8];
  __shared__ int s_box[5];  // cx0, cy0, cw, ch, mode (0 skip, 1 staged, 2 fallback)

  const int tiles_x = (p.W + kAggTile - 1) / kAggTile;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x = tx * kAggTile + (tid & 15), y = ty * kAggTile + (tid >> 4);
  const bool inside = x < p.W && y < p.H;
  const float u = p.lx.at(inside ? x : 0), v = p.ly.at(inside ? y : 0);
  float num = 0.f, den = 0.f;

  for (int n = blockIdx.y; n < p.N; n += p.G) {
    const Mat3 Mu = load_mat(p.mu + 9 * n);
    const Mat3 Mf = load_mat(p.mf + 9 * n);
    float ix, iy;
    warp_coord(Mu, u, v, p.W, p.H, ix, iy);
    // taps touch [floor(ix), floor(ix)+1] x [floor(iy), floor(iy)+1]; a pixel takes part only if
    // that range meets the image (NaN/inf coordinates fail the comparisons and drop out)
    const bool act = inside && ix > -1.f && ix < (float)p.W && iy > -1.f && iy < (float)p.H;
    const Taps tp = make_taps(ix, iy);
    int bx0 = 1 << 30, by0 = 1 << 30, bx1 = -1, by1 = -1;
    if (act) {
      bx0 = max((int)tp.x0, 0);
      by0 = max((int)tp.y0, 0);
      bx1 = min((int)tp.x1, p.W - 1);
      by1 = min((int)tp.y1, p.H - 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
      by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
      by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0) {
      s_red[0][warp] = bx0;
      s_red[1][warp] = by0;
      s_red[2][warp] = bx1;
      s_red[3][warp] = by1;
    }
    __syncthreads();  // also orders the previous view's gather before this view's staging
    if (tid == 0) {
      int a = 1 << 30, b = 1 << 30, c = -1, d = -1;
      for (int w = 0; w < 8; ++w) {
        a = min(a, s_red[0][w]);
        b = min(b, s_red[1][w]);
        c = max(c, s_red[2][w]);
        d = max(d, s_red[3][w]);
      }
      if (c < 0) {
        s_box[4] = 0;
      } else {
        const int cx0 = a >> 3, cy0 = b >> 3, cw = (c >> 3) - cx0 + 1, ch = (d >> 3) - cy0 + 1;
        s_box[0] = cx0;
        s_box[1] = cy0;
        s_box[2] = cw;
        s_box[3] = ch;
        s_box[4] = (cw * ch <= kAggMaxCells) ? 1 : 2;
      }
    }
    __syncthreads();
    const int mode = s_box[4];
    if (mode == 0) continue;
    const int cx0 = s_box[0], cy0 = s_box[1], cw = s_box[2], ch = s_box[3];

    if (mode == 1) {
      for (int c = warp; c < cw * ch; c += 8) {
        const int cx = cx0 + c % cw, cy = cy0 + c / cw;
        const float* q = p.semi + (((long long)n * p.Hc + cy) * p.Wc + cx) * p.cs;
        const float a = __ldg(q + lane), b = __ldg(q + lane + 32), d = (lane == 0) ? __ldg(q + 64) : -INFINITY;
        const float mx = warp_max(fmaxf(fmaxf(a, b), d));
        const float ea = expf(a - mx), eb = expf(b - mx), ed = (lane == 0) ? expf(d - mx) : 0.f;
        const float s = warp_sum(ea + eb + ed);
        const int px = cx * 8 + (lane & 7), py = cy * 8 + (lane >> 3);
        const float ux = p.lx.at(px);
        const float ma = nearest_valid(Mf, ux, p.ly.at(py), p.W, p.H) ? 1.f : 0.f;
        const float mb = nearest_valid(Mf, ux, p.ly.at(py + 4), p.W, p.H) ? 1.f : 0.f;
        s_hm[c * 64 + lane] = __fmul_rn(__fdiv_rn(ea, s), ma);
        s_hm[c * 64 + lane + 32] = __fmul_rn(__fdiv_rn(eb, s), mb);
        s_m[c * 64 + lane] = ma;
        s_m[c * 64 + lane + 32] = mb;
      }
      __syncthreads();
      if (act) {
        auto idx = [&](int xx, int yy) { return (((yy >> 3) - cy0) * cw + ((xx >> 3) - cx0)) * 64 + ((yy & 7) << 3 | (xx & 7)); };
        num = __fadd_rn(num, bilinear(tp, p.W, p.H, [&](int xx, int yy) { return s_hm[idx(xx, yy)]; }));
        den = __fadd_rn(den, bilinear(tp, p.W, p.H, [&](int xx, int yy) { return s_m[idx(xx, yy)]; }));
      }
    } else if (act) {  // view maps this tile onto more than kAggMaxCells cells: gather from global
      float hm[4] = {0.f, 0.f, 0.f, 0.f}, mm[4] = {0.f, 0.f, 0.f, 0.f};
      const float xs[4] = {tp.x0, tp.x1, tp.x0, tp.x1}, ys[4] = {tp.y0, tp.y0, tp.y1, tp.y1};
      const float ws[4] = {tp.w_nw, tp.w_ne, tp.w_sw, tp.w_se};
      float a = 0.f, b = 0.f;
      for (int k = 0; k < 4; ++k)
        if (in_bounds(xs[k], ys[k], p.W, p.H)) {
          const float2 r = prob_mask_global(p, n, Mf, (int)xs[k], (int)ys[k]);
          hm[k] = r.x;
          mm[k] = r.y;
          a = __fadd_rn(a, __fmul_rn(hm[k], ws[k]));
          b = __fadd_rn(b, __fmul_rn(mm[k], ws[k]));
        }
      num = __fadd_rn(num, a);
      den = __fadd_rn(den, b);
    }
  }
  if (inside) {
    float* o = p.part + (long long)blockIdx.y * 2 * p.H * p.W + (long long)y * p.W + x;
    o[0] = num;
    o[(long long)p.H * p.W] = den;
  }
}

// Bit-exact gather form of the aggregate (validation path SPB_IMPL_SIMT; the product path is ha_fast.cu): softmax
// probabilities (dustbin dropped) as probs[V,Hc,Wc,64] and the valid-mask bits of the warped views,
// one thread per output pixel, views n = begin + g, begin + g + G, ... accumulated in registers in a fixed order.
// blockIdx.z = image * G + group; image b owns the views [b*N, (b+1)*N) of the pass, or [offs[b], offs[b+1]) when
// offs != NULL (homography sharding: a different number of views per image); view n uses the matrix vm.mat(n).
__global__ void __launch_bounds__(256) agg_probs_kernel(const float* __restrict__ probs,
                                                        const uint8_t* __restrict__ bits,
                                                        const float* __restrict__ mu, ViewMap vm, int N, int H,
                                                        int W, int G, float* __restrict__ part, Lin lx, Lin ly,
                                                        const int* __restrict__ offs) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int Wc = W >> 3, Hc = H >> 3;
  const int b = blockIdx.z / G, g = blockIdx.z % G;
  int n_begin = b * N + g, n_end = (b + 1) * N;  // flat view numbers of the pass (probs / bits / vm index)
  if (offs) {
    n_begin = __ldg(offs + b) + g;
    n_end = __ldg(offs + b + 1);
  }
  const float u = lx.at(x), v = ly.at(y);
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  float num = 0.f, den = 0.f;
  for (int n = n_begin; n < n_end; n += G) {
    const Mat3 M = load_mat(mu + 9LL * vm.mat(n));
    float ix, iy;
    warp_coord(M, u, v, W, H, ix, iy);
    const Taps tp = make_taps(ix, iy);
    const float* pn = probs + (long long)n * Hc * Wc * 64;
    const uint8_t* bn = bits + (long long)n * H * Wc;
    // shared work of the four taps hoisted (validity per column / row, one float->int per axis, the cell-major
    // offsets split into a row part and a column part); same products and the same order of additions as
    // tap-by-tap evaluation -> bit-identical sums
    const bool vx0 = tp.x0 >= 0.f && tp.x0 <= wm, vx1 = tp.x1 >= 0.f && tp.x1 <= wm;
    const bool vy0 = tp.y0 >= 0.f && tp.y0 <= hm, vy1 = tp.y1 >= 0.f && tp.y1 <= hm;
    const int x0 = (int)fminf(fmaxf(tp.x0, -1.f), wm), y0 = (int)fminf(fmaxf(tp.y0, -1.f), hm);
    const int x1 = x0 + 1, y1 = y0 + 1;
    const int cx0 = (x0 >> 3) * 64 + (x0 & 7), cx1 = (x1 >> 3) * 64 + (x1 & 7);
    const int ry0 = (y0 >> 3) * Wc * 64 + ((y0 & 7) << 3), ry1 = (y1 >> 3) * Wc * 64 + ((y1 & 7) << 3);
    const int by0 = y0 * Wc, by1 = y1 * Wc;
    float a = 0.f, b = 0.f;
    auto tap = [&](bool ok, int boff, int xx, int poff, float w) {
      if (!ok) return;
      const float m = (float)((__ldg(bn + boff + (xx >> 3)) >> (xx & 7)) & 1);
      const float pr = __ldg(pn + poff);
      a = __fadd_rn(a, __fmul_rn(__fmul_rn(pr, m), w));
      b = __fadd_rn(b, __fmul_rn(m, w));
    };
    tap(vx0 && vy0, by0, x0, ry0 + cx0, tp.w_nw);
    tap(vx1 && vy0, by0, x1, ry0 + cx1, tp.w_ne);
    tap(vx0 && vy1, by1, x0, ry1 + cx0, tp.w_sw);
    tap(vx1 && vy1, by1, x1, ry1 + cx1, tp.w_se);
    num = __fadd_rn(num, a);
    den = __fadd_rn(den, b);
  }
  float* o = part + (long long)blockIdx.z * 2 * H * W + (long long)y * W + x;
  o[0] = num;
  o[(long long)H * W] = den;
}

// sums[b][i] (+)= sum_g part[b][g][i], fixed order; n = elements per image (2*H*W), total = B*n
__global__ void __launch_bounds__(256) agg_reduce_kernel(const float* __restrict__ part, int G, long long n,
                                                         long long total, float* __restrict__ sums, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n, r = i % n;
    float s = accumulate ? sums[i] : 0.f;
    for (int g = 0; g < G; ++g) s = __fadd_rn(s, part[(b * G + g) * n + r]);
    sums[i] = s;
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ sums, float* __restrict__ heat,
                                                       long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    heat[i] = __fdiv_rn(sums[i], sums[n + i]);
}

__global__ void __launch_bounds__(256) finalize_batch_kernel(const float* __restrict__ sums, float* __restrict__ heat,
                                                             long long n, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n, r = i % n;
    heat[i] = __fdiv_rn(sums[b * 2 * n + r], sums[b * 2 * n + n + r]);
  }
}

static int agg_groups(int N, int H, int W) {
  const int tiles = ceil_div(W, kAggTile) * ceil_div(H, kAggTile);
  int G = ceil_div(3LL * kNumSMs * 6, tiles);  // ~3 waves of 6 resident CTAs per SM
  if (G > N) G = N;
  if (G < 1) G = 1;
  return G;
}

int aggregate_probs_groups(int B, int N, int H, int W) {
  const int tiles = ceil_div(W, 32) * ceil_div(H, 8);
  int G = ceil_div(4LL * kNumSMs * 8, (long long)tiles * B);
  if (G > N) G = N;
  if (G < 1) G = 1;
  return G;
}
size_t aggregate_probs_workspace(int B, int N, int H, int W) {
  return (size_t)B * aggregate_probs_groups(B, N, H, W) * 2 * H * W * sizeof(float);
}

// B images x N views each (probs / bits contiguous over b*N + n); sums [B,2,H,W]
int reduce_partials(const float* part, int G, int nb, int H, int W, float* sums, int accumulate, cudaStream_t st) {
  const long long n = 2LL * H * W, total = n * nb;
  agg_reduce_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st>>>(part, G, n, total, sums, accumulate);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int aggregate_probs(const float* probs, const uint8_t* bits, const float* mats_unwarp, const ViewMap& vm, int B, int N,
                    int H, int W, float* sums, int accumulate, void* ws, cudaStream_t st, const int* offs) {
  const int G = aggregate_probs_groups(B, N, H, W);  // (ragged: N = the largest per-image view count)
  float* part = static_cast<float*>(ws);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), G * B);
  agg_probs_kernel<<<grid, 256, 0, st>>>(probs, bits, mats_unwarp, vm, N, H, W, G, part, Lin(W), Lin(H), offs);
  SPB_CHECK_LAUNCH();
  return reduce_partials(part, G, B, H, W, sums, accumulate, st);
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_flatten_detection(const float* semi, int nhwc_cstride, float* heat, int B, int Hc, int Wc, void* stream) {
  SPB_CHECK_ARG(semi && heat && B >= 0 && Hc > 0 && Wc > 0);
  SPB_CHECK_ARG(nhwc_cstride == 0 || nhwc_cstride >= 65);
  if (B == 0) return SPB_OK;
  const long long ncell = (long long)B * Hc * Wc;
  if (nhwc_cstride) {
    int blocks = min(ceil_div(ncell, 8), kNumSMs * 8);
    flatten_nhwc_kernel<<<blocks, 256, 0, as_stream(stream)>>>(semi, nhwc_cstride, heat, B, Hc, Wc);
  } else {
    int blocks = min(ceil_div(ncell, 256), kNumSMs * 8);
    flatten_nchw_kernel<<<blocks, 256, 0, as_stream(stream)>>>(semi, heat, B, Hc, Wc);
  }
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_combine_heatmap(const float* heat, const float* mask, const float* mats_unwarp, float* out, int N, int H,
                        int W, void* stream) {
  SPB_CHECK_ARG(heat && mask && mats_unwarp && out && N >= 1 && H >= 2 && W >= 2);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
  combine_kernel<<<grid, 256, 0, as_stream(stream)>>>(heat, mask, mats_unwarp, out, N, H, W, Lin(W), Lin(H));
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

size_t spb_ha_aggregate_workspace_bytes(int N, int H, int W) {
  if (N < 1 || H < 8 || W < 8) return 0;
  return (size_t)agg_groups(N, H, W) * 2 * H * W * sizeof(float);
}

int spb_ha_aggregate(const float* semi_nhwc, int cstride, const float* mats_unwarp, const float* mats_fwd, int N,
                     int H, int W, float* sums, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(semi_nhwc && mats_unwarp && mats_fwd && sums && ws);
  SPB_CHECK_ARG(N >= 1 && H >= 8 && W >= 8 && (H % 8) == 0 && (W % 8) == 0 && cstride >= 65);
  if (ws_bytes < spb_ha_aggregate_workspace_bytes(N, H, W)) {
    set_error("spb_ha_aggregate: workspace %zu < %zu bytes", ws_bytes, spb_ha_aggregate_workspace_bytes(N, H, W));
    return SPB_ENOMEM;
  }
  AggParams p;
  p.semi = semi_nhwc;
  p.cs = cstride;
  p.mu = mats_unwarp;
  p.mf = mats_fwd;
  p.N = N;
  p.H = H;
  p.W = W;
  p.Hc = H / 8;
  p.Wc = W / 8;
  p.G = agg_groups(N, H, W);
  p.part = static_cast<float*>(ws);
  p.lx = Lin(W);
  p.ly = Lin(H);
  dim3 grid(ceil_div(W, kAggTile) * ceil_div(H, kAggTile), p.G);
  ha_aggregate_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  SPB_CHECK_LAUNCH();
  const long long n = 2LL * H * W;
  agg_reduce_kernel<<<min(ceil_div(n, 256), kNumSMs * 8), 256, 0, as_stream(stream)>>>(p.part, p.G, n, n, sums, accumulate);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_ha_finalize_batch(const float* sums, float* heat, int B, int H, int W, void* stream) {
  SPB_CHECK_ARG(sums && heat && B > 0 && H > 0 && W > 0);
  const long long n = (long long)H * W, total = n * B;
  finalize_batch_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, as_stream(stream)>>>(sums, heat, n, total);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_ha_finalize(const float* sums, float* heat, int H, int W, void* stream) {
  SPB_CHECK_ARG(sums && heat && H > 0 && W > 0);
  const long long n = (long long)H * W;
  finalize_kernel<<<min(ceil_div(n, 256), kNumSMs * 8), 256, 0, as_stream(stream)>>>(sums, heat, n);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
