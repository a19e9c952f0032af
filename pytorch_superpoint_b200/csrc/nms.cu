// (4) Threshold + exact greedy grid NMS + border removal + confidence ordering + top-k.
// Replaces models/model_wrap.py:252-279 (getPtsFromHeatmap), 117-180 (nms_fast) and
// export.py:322-328 (pts[:top_k]).
//
// The reference's sequential greedy sweep (highest confidence first, suppress the (2d+1)^2 window)
// computes the lexicographically-first maximal independent set of the "Chebyshev distance <= d"
// graph under the total order (confidence desc, row-major index asc).  That set is reproduced
// exactly by rounds over two bitmaps held in shared memory:
//     phase 1  an undecided pixel with a KEPT pixel in its window becomes suppressed;
//     phase 2  an undecided pixel that beats every other undecided pixel in its window becomes KEPT.
// Comparisons are integer compares of (fp32 bit pattern, index), so the result is bit-exact.
// One CTA per image (warp-ballot bitmap words, funnel-shift window extraction); survivors are
// compacted in row-major order with their 64-bit sort key and ranked by a second, grid-wide
// count-greater kernel (keys are unique, so rank == final position).
#include "common.cuh"

namespace spb {

constexpr int kNmsThreads = 1024;
constexpr int kMaxDist = 15;  // window bits (2d+1) must fit a 32-bit word

struct NmsParams {
  const float* heat;
  int B, H, W;
  float thr;
  int d, border;
  int ww;      // bitmap words per row incl. one pad word on each side
  int maxk;    // key slots per image
  unsigned long long* keys;  // [B, maxk]
  int* nkept;                // [B]
};

// order-preserving map fp32 -> uint32 (-0 folded onto +0)
__device__ __forceinline__ unsigned ord_bits(float c) {
  unsigned b = __float_as_uint(c);
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ unsigned window_bits(const unsigned* row, int x, int d) {
  // bits x-d .. x+d of a padded bitmap row (row[0] is the left pad word, pixel 0 = bit 0 of row[1])
  const int s = x - d + 32;  // >= 32 - 15 > 0
  const unsigned lo = row[s >> 5], hi = row[(s >> 5) + 1];
  return __funnelshift_r(lo, hi, s & 31) & ((2u << (2 * d)) - 1u);
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(NmsParams p) {
  extern __shared__ unsigned s_bits[];
  const int b = blockIdx.x;
  const int H = p.H, W = p.W, d = p.d, ww = p.ww;
  const int rows = H + 2 * d;
  unsigned* und = s_bits;                // undecided, [rows][ww]
  unsigned* kept = s_bits + rows * ww;   // kept
  __shared__ int s_scan[kNmsThreads / 32];
  __shared__ int s_total;
  const float* heat = p.heat + (long long)b * H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kNmsThreads / 32;
  const int wpr = (W + 31) >> 5;  // data words per row
  const int ntask = H * wpr;

  for (int i = tid; i < 2 * rows * ww; i += kNmsThreads) s_bits[i] = 0u;
  __syncthreads();

  // threshold (model_wrap.py:261): fp32 compare, NaN never passes
  for (int t = warp; t < ntask; t += nwarp) {
    const int y = t / wpr, w = t % wpr, x = w * 32 + lane;
    const bool c = x < W && heat[y * W + x] >= p.thr;
    const unsigned m = __ballot_sync(0xffffffffu, c);
    if (lane == 0) und[(y + d) * ww + w + 1] = m;
  }
  __syncthreads();

  while (true) {
    // phase 1: suppress undecided pixels that see a kept pixel (a kept pixel sees itself and
    // thereby retires from the undecided set one round after it was selected)
    int any = 0;
    for (int t = warp; t < ntask; t += nwarp) {
      const int y = t / wpr, w = t % wpr, x = w * 32 + lane;
      const unsigned word = und[(y + d) * ww + w + 1];
      if (word == 0u) continue;
      bool still = (word >> lane) & 1u;
      if (still) {
        for (int r = 0; r <= 2 * d; ++r)
          if (window_bits(kept + (y + r) * ww, x, d)) {
            still = false;
            break;
          }
      }
      const unsigned m = __ballot_sync(0xffffffffu, still);
      if (lane == 0) und[(y + d) * ww + w + 1] = m;
      any |= (m != 0u);
    }
    if (!__syncthreads_or(any)) break;
    // phase 2: select local winners among the undecided
    for (int t = warp; t < ntask; t += nwarp) {
      const int y = t / wpr, w = t % wpr, x = w * 32 + lane;
      const unsigned word = und[(y + d) * ww + w + 1];
      if (word == 0u) continue;
      bool win = (word >> lane) & 1u;
      if (win) {
        const unsigned mine = ord_bits(heat[y * W + x]);
        for (int r = 0; r <= 2 * d && win; ++r) {
          const int yy = y + r - d;
          unsigned bits = window_bits(und + (y + r) * ww, x, d);
          if (r == d) bits &= ~(1u << d);
          while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            const int xx = x - d + j;
            const unsigned other = ord_bits(__ldg(heat + yy * W + xx));
            const bool greater = other > mine;
            if (greater || (other == mine && (yy < y || (yy == y && xx < x)))) {
              win = false;
              break;
            }
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, win);
      if (lane == 0 && m) kept[(y + d) * ww + w + 1] |= m;
    }
    __syncthreads();
  }

  // border removal AFTER suppression (model_wrap.py:274-278), row-major compaction, sort keys
  int mycount = 0;
  for (int t = warp; t < ntask; t += nwarp) {  // pass 1: count per warp
    const int y = t / wpr, w = t % wpr, x = w * 32 + lane;
    const bool k = ((kept[(y + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    mycount += __popc(__ballot_sync(0xffffffffu, k));
  }
  // tasks are strided over warps, so a plain per-warp prefix would not be row-major; compaction
  // order does not matter for the result (keys are unique and ranked afterwards), only counts do.
  if (lane == 0) s_scan[warp] = mycount;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < nwarp; ++i) {
      const int c = s_scan[i];
      s_scan[i] = acc;
      acc += c;
    }
    s_total = acc;
    p.nkept[b] = acc < p.maxk ? acc : p.maxk;
  }
  __syncthreads();
  int base = s_scan[warp];
  unsigned long long* keys = p.keys + (long long)b * p.maxk;
  for (int t = warp; t < ntask; t += nwarp) {
    const int y = t / wpr, w = t % wpr, x = w * 32 + lane;
    const bool k = ((kept[(y + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      const float c = heat[y * W + x];
      const unsigned bits = ord_bits(c);
      if (slot < p.maxk) keys[slot] = ((unsigned long long)bits << 32) | (unsigned)(y * W + x);
    }
    base += __popc(m);
  }
}

// rank = number of keys greater than mine; final order = key descending
// (confidence desc, then larger row-major index first: model_wrap.py:177-178 then 271-272 with stable sorts)
__global__ void __launch_bounds__(256) rank_scatter_kernel(const float* __restrict__ heat, int H, int W,
                                                           const unsigned long long* __restrict__ keys,
                                                           const int* __restrict__ nkept, int maxk, int limit,
                                                           int capacity, float* __restrict__ pts,
                                                           int* __restrict__ counts, int* __restrict__ total) {
  __shared__ unsigned long long s_k[256];
  const int b = blockIdx.y;
  const int K = nkept[b];
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int c = K;
    if (limit > 0 && c > limit) c = limit;
    if (c > capacity) c = capacity;
    counts[b] = c;
    if (total) total[b] = K;
  }
  if (blockIdx.x * 256 >= K) return;
  const unsigned long long* kb = keys + (long long)b * maxk;
  const unsigned long long mine = j < K ? kb[j] : 0ull;
  int rank = 0;
  for (int base = 0; base < K; base += 256) {
    __syncthreads();
    s_k[threadIdx.x] = base + threadIdx.x < K ? kb[base + threadIdx.x] : 0ull;
    __syncthreads();
    const int lim = min(256, K - base);
    for (int i = 0; i < lim; ++i) rank += s_k[i] > mine;
  }
  if (j < K) {
    int c = K;
    if (limit > 0 && c > limit) c = limit;
    if (c > capacity) c = capacity;
    if (rank < c) {
      const unsigned idx = (unsigned)(mine & 0xffffffffull);
      float* o = pts + ((long long)b * capacity + rank) * 3;
      o[0] = (float)(idx % W);
      o[1] = (float)(idx / W);
      o[2] = heat[(long long)b * H * W + idx];
    }
  }
}

static int nms_maxk(int H, int W, int d) {
  return ceil_div(H, d + 1) * ceil_div(W, d + 1);  // kept points are pairwise > d apart (Chebyshev)
}
static size_t nms_smem(int H, int W, int d) {
  const int ww = ceil_div(W, 32) + 2;
  return (size_t)2 * (H + 2 * d) * ww * sizeof(unsigned);
}

}  // namespace spb

using namespace spb;

extern "C" {

size_t spb_nms_workspace_bytes(int B, int H, int W, int nms_dist) {
  if (B < 1 || H < 1 || W < 1 || nms_dist < 0) return 0;
  return (size_t)B * nms_maxk(H, W, nms_dist) * sizeof(unsigned long long) + (size_t)B * sizeof(int) + 256;
}

int spb_get_pts_from_heatmap(const float* heat, int B, int H, int W, float conf_thresh, int nms_dist,
                             int border_remove, int top_k, int capacity, float* pts, int* counts, int* total,
                             void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(heat && pts && counts && ws);
  SPB_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && capacity >= 1 && border_remove >= 0);
  SPB_CHECK_ARG(nms_dist >= 0 && nms_dist <= kMaxDist);
  SPB_CHECK_ARG((long long)H * W < (1LL << 31));
  if (ws_bytes < spb_nms_workspace_bytes(B, H, W, nms_dist)) {
    set_error("spb_get_pts_from_heatmap: workspace %zu < %zu bytes", ws_bytes, spb_nms_workspace_bytes(B, H, W, nms_dist));
    return SPB_ENOMEM;
  }
  const size_t smem = nms_smem(H, W, nms_dist);
  if (smem > 200 * 1024) {
    set_error("spb_get_pts_from_heatmap: image %dx%d needs %zu B of bitmap shared memory (> 200 KB)", H, W, smem);
    return SPB_EINVAL;
  }
  SPB_CHECK_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  NmsParams p;
  p.heat = heat;
  p.B = B;
  p.H = H;
  p.W = W;
  p.thr = conf_thresh;
  p.d = nms_dist;
  p.border = border_remove;
  p.ww = ceil_div(W, 32) + 2;
  p.maxk = nms_maxk(H, W, nms_dist);
  p.keys = static_cast<unsigned long long*>(ws);
  p.nkept = reinterpret_cast<int*>(p.keys + (size_t)B * p.maxk);
  cudaStream_t st = as_stream(stream);
  nms_kernel<<<B, kNmsThreads, smem, st>>>(p);
  SPB_CHECK_LAUNCH();
  dim3 grid(ceil_div(p.maxk, 256), B);
  rank_scatter_kernel<<<grid, 256, 0, st>>>(heat, H, W, p.keys, p.nkept, p.maxk, top_k, capacity, pts, counts, total);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
