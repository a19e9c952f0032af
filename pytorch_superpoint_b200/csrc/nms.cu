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
// Comparisons are integer compares of (order-preserving fp32 key, index), so the result is bit-exact.
//
// One thread-block CLUSTER (up to 8 CTAs) per image: each CTA owns a band of rows -- its slice of the
// two bitmaps and (when it fits) the fp32 keys of band + halo live in its shared memory; after each
// phase the d boundary rows are pushed into the neighbours' halo rows through distributed shared
// memory and the cluster barrier separates the phases.  Bitmap words come from warp ballots, windows
// from funnel shifts.  Survivors are compacted with their 64-bit sort key and ranked by a second,
// grid-wide count-greater kernel (keys are unique, so rank == final position).
#include <cstdlib>

#include "common.cuh"

namespace spb {

constexpr int kNmsThreads = 1024;  // threads per CTA for small batches; large batches run 512 / 256 (more CTAs per SM)
constexpr int kMaxDist = 15;  // window bits (2d+1) must fit a 32-bit word
constexpr int kMaxCluster = 8;

struct NmsParams {
  const float* heat;
  int B, H, W;
  float thr;
  int d, border;
  int ww;      // bitmap words per row incl. one pad word on each side
  int rb;      // rows per band (the last band may be shorter, never shorter than d)
  int keys_smem;
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

__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a word into the same shared-memory variable of another CTA of the cluster
__device__ __forceinline__ void dsmem_store(const void* local, unsigned rank, unsigned v) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) nms_kernel(NmsParams p) {
  extern __shared__ unsigned s_dyn[];
  __shared__ int s_flag[kMaxCluster];
  __shared__ int s_cnt[kMaxCluster];
  __shared__ int s_scan[THREADS / 32];
  const unsigned crank = cluster_rank(), csize = cluster_size();
  const int b = blockIdx.x / csize;
  const int H = p.H, W = p.W, d = p.d, ww = p.ww;
  const int y0 = crank * p.rb, y1 = min(H, y0 + p.rb), nr = max(y1 - y0, 0);
  const int lrows = p.rb + 2 * d;                 // local bitmap rows: [0,d) top halo, [d,d+nr) own, then bottom halo
  unsigned* und = s_dyn;                          // [lrows][ww]
  unsigned* kept = s_dyn + lrows * ww;
  unsigned* keys_s = s_dyn + 2 * lrows * ww;      // [lrows][W] order-preserving keys (band + halo) when keys_smem
  const float* heat = p.heat + (long long)b * H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = THREADS / 32;
  const int wpr = (W + 31) >> 5;  // data words per row
  const int ntask = nr * wpr;

  for (int i = tid; i < 2 * lrows * ww; i += THREADS) s_dyn[i] = 0u;
  if (p.keys_smem)
    for (int i = tid; i < lrows * W; i += THREADS) {
      const int yy = y0 - d + i / W;
      keys_s[i] = (yy >= 0 && yy < H) ? ord_bits(__ldg(heat + (long long)yy * W + i % W)) : 0u;
    }
  cluster_barrier();  // every CTA of the cluster is running and has cleared its bitmaps

  auto key_at = [&](int yy, int xx) -> unsigned {
    return p.keys_smem ? keys_s[(yy - y0 + d) * W + xx] : ord_bits(__ldg(heat + (long long)yy * W + xx));
  };
  // push my d first / last own rows of a bitmap into the neighbours' halo rows
  auto push_halo = [&](unsigned* bm) {
    if (d == 0 || nr == 0) return;
    const int n = d * ww;
    if (crank > 0)  // my first d rows -> bottom halo of the band above (which is always a full band)
      for (int i = tid; i < n; i += THREADS) dsmem_store(bm + (d + p.rb) * ww + i, crank - 1, bm[d * ww + i]);
    if (crank + 1 < csize && y1 < H)  // my last d rows -> top halo of the band below
      for (int i = tid; i < n; i += THREADS) dsmem_store(bm + i, crank + 1, bm[nr * ww + i]);
  };

  // threshold (model_wrap.py:261): fp32 compare, NaN never passes
  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool c = x < W && heat[(long long)y * W + x] >= p.thr;
    const unsigned m = __ballot_sync(0xffffffffu, c);
    if (lane == 0) und[(ly + d) * ww + w + 1] = m;
  }
  __syncthreads();
  push_halo(und);
  cluster_barrier();

  while (true) {
    // phase 1: suppress undecided pixels that see a kept pixel (a kept pixel sees itself and
    // thereby retires from the undecided set one round after it was selected)
    int any = 0;
    for (int t = warp; t < ntask; t += nwarp) {
      const int ly = t / wpr, w = t % wpr, x = w * 32 + lane;
      const unsigned word = und[(ly + d) * ww + w + 1];
      if (word == 0u) continue;
      bool still = (word >> lane) & 1u;
      if (still) {
        for (int r = 0; r <= 2 * d; ++r)
          if (window_bits(kept + (ly + r) * ww, x, d)) {
            still = false;
            break;
          }
      }
      const unsigned m = __ballot_sync(0xffffffffu, still);
      if (lane == 0) und[(ly + d) * ww + w + 1] = m;
      any |= (m != 0u);
    }
    any = __syncthreads_or(any);
    push_halo(und);
    if (tid < (int)csize) dsmem_store(&s_flag[crank], tid, (unsigned)any);
    cluster_barrier();
    int go = 0;
    for (unsigned i = 0; i < csize; ++i) go |= s_flag[i];
    if (!go) break;
    // phase 2: select local winners among the undecided
    for (int t = warp; t < ntask; t += nwarp) {
      const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
      const unsigned word = und[(ly + d) * ww + w + 1];
      if (word == 0u) continue;
      bool win = (word >> lane) & 1u;
      if (win) {
        const unsigned mine = key_at(y, x);
        for (int r = 0; r <= 2 * d && win; ++r) {
          const int yy = y + r - d;
          unsigned bits = window_bits(und + (ly + r) * ww, x, d);
          if (r == d) bits &= ~(1u << d);
          while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            const int xx = x - d + j;
            const unsigned other = key_at(yy, xx);
            if (other > mine || (other == mine && (yy < y || (yy == y && xx < x)))) {
              win = false;
              break;
            }
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, win);
      if (lane == 0 && m) kept[(ly + d) * ww + w + 1] |= m;
    }
    __syncthreads();
    push_halo(kept);
    cluster_barrier();
  }

  // border removal AFTER suppression (model_wrap.py:274-278), compaction with sort keys
  int mycount = 0;
  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool k = ((kept[(ly + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    mycount += __popc(__ballot_sync(0xffffffffu, k));
  }
  if (lane == 0) s_scan[warp] = mycount;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < nwarp; ++i) {
      const int c = s_scan[i];
      s_scan[i] = acc;
      acc += c;
    }
    for (unsigned i = 0; i < csize; ++i) dsmem_store(&s_cnt[crank], i, (unsigned)acc);
  }
  cluster_barrier();
  int cta_base = 0, total = 0;
  for (unsigned i = 0; i < csize; ++i) {
    if (i < crank) cta_base += s_cnt[i];
    total += s_cnt[i];
  }
  if (crank == 0 && tid == 0) p.nkept[b] = total < p.maxk ? total : p.maxk;
  int base = cta_base + s_scan[warp];
  unsigned long long* keys = p.keys + (long long)b * p.maxk;
  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool k = ((kept[(ly + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < p.maxk) keys[slot] = ((unsigned long long)key_at(y, x) << 32) | (unsigned)(y * W + x);
    }
    base += __popc(m);
  }
  cluster_barrier();  // nobody leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------
// The same rounds with SEPARABLE window maxima (the default): per round
//   (A) every pixel of a row that can matter gets the leftmost maximum of the undecided keys in [x-d, x+d]
//       (key + column offset, shared by the 2d+1 output rows that look at this row);
//   (B) an undecided pixel wins iff the topmost maximum over the rows y-d .. y+d of (A) is itself -- the total order
//       (key desc, row-major index asc) again, so the selected set is identical to nms_kernel's phase 2;
//   (C) the winners of the round are dilated by d on the BITMAP (funnel shifts, one thread per word) and cleared
//       from the undecided set together with everything they suppress -- nms_kernel's phase 1 without a per-pixel
//       loop.
// A dense candidate field (a random-init network: 34 k of 76.8 k pixels above the threshold) costs ~90 instructions
// per pixel and round instead of a divergent scan of up to (2d+1)^2 neighbours per candidate.
// ------------------------------------------------------------------------------------------------
template <int THREADS, int DT>  // DT: compile-time nms distance (the tap loops unroll, their loads overlap) or 0 = p.d
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) nms2_kernel(NmsParams p) {
  extern __shared__ unsigned s_dyn[];
  __shared__ int s_flag[kMaxCluster];
  __shared__ int s_cnt[kMaxCluster];
  __shared__ int s_scan[THREADS / 32];
  const unsigned crank = cluster_rank(), csize = cluster_size();
  const int b = blockIdx.x / csize;
  const int H = p.H, W = p.W, d = DT ? DT : p.d, ww = p.ww;
  const int y0 = crank * p.rb, y1 = min(H, y0 + p.rb), nr = max(y1 - y0, 0);
  const int lrows = p.rb + 2 * d;  // local rows: [0,d) top halo, [d,d+nr) own, then bottom halo
  unsigned* und = s_dyn;           // [lrows][ww] undecided
  unsigned* kept = und + lrows * ww;
  unsigned* newk = kept + lrows * ww;   // winners of the current round
  unsigned* hdil = newk + lrows * ww;   // newk dilated horizontally
  unsigned* rk = hdil + lrows * ww;     // [lrows][W] leftmost maximum of the undecided keys in [x-d, x+d] (0: none)
  unsigned char* rj = reinterpret_cast<unsigned char*>(rk + lrows * W);  // [lrows][W] its window offset 0 .. 2d
  unsigned* keys_s = reinterpret_cast<unsigned*>(rj + (((size_t)lrows * W + 15) & ~(size_t)15));  // optional
  const float* heat = p.heat + (long long)b * H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = THREADS / 32;
  const int wpr = (W + 31) >> 5;
  const int ntask = nr * wpr, nall = lrows * wpr;

  for (int i = tid; i < 4 * lrows * ww; i += THREADS) s_dyn[i] = 0u;
  if (p.keys_smem)
    for (int i = tid; i < lrows * W; i += THREADS) {
      const int yy = y0 - d + i / W;
      keys_s[i] = (yy >= 0 && yy < H) ? ord_bits(__ldg(heat + (long long)yy * W + i % W)) : 0u;
    }
  cluster_barrier();

  // key of local row lr (0 .. lrows-1), column xx; only called where an undecided bit is set (inside the image)
  auto key_l = [&](int lr, int xx) -> unsigned {
    return p.keys_smem ? keys_s[lr * W + xx] : ord_bits(__ldg(heat + (long long)(y0 - d + lr) * W + xx));
  };
  auto push_halo = [&](unsigned* bm) {
    if (d == 0 || nr == 0) return;
    const int n = d * ww;
    if (crank > 0)
      for (int i = tid; i < n; i += THREADS) dsmem_store(bm + (d + p.rb) * ww + i, crank - 1, bm[d * ww + i]);
    if (crank + 1 < csize && y1 < H)
      for (int i = tid; i < n; i += THREADS) dsmem_store(bm + i, crank + 1, bm[nr * ww + i]);
  };

  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool c = x < W && heat[(long long)y * W + x] >= p.thr;
    const unsigned m = __ballot_sync(0xffffffffu, c);
    if (lane == 0) und[(ly + d) * ww + w + 1] = m;
  }
  __syncthreads();
  push_halo(und);
  cluster_barrier();

  while (true) {
    // (A) row maxima for every local row / column some own undecided pixel looks at
    if (DT == 4) {
      // a warp takes 24 output columns of a row plus 4 columns on either side: every lane loads ITS masked key once
      // and the leftmost maximum over [x-4, x+4] comes from four shuffle steps (windows of 2, 4, 8, 9 lanes)
      for (int i = tid; i < nall; i += THREADS) {  // who looks at a row: undecided bits of the OWN rows within d
        const int lr = i / wpr, w = i % wpr;
        unsigned v = 0u;
        for (int q = max(lr - d, d); q <= min(lr + d, d + nr - 1); ++q) v |= und[q * ww + w + 1];
        hdil[lr * ww + w + 1] = v;  // (hdil is free until (C))
      }
      __syncthreads();
      constexpr int OUT = 32 - 2 * 4;
      const int spr = (W + OUT - 1) / OUT;
      for (int t = warp; t < lrows * spr; t += nwarp) {
        const int lr = t / spr, s0 = (t % spr) * OUT;  // first output column of the strip
        const int start = s0 - 4 + 32;                 // bit position of lane 0's column in the padded bitmap row
        const int wi = start >> 5, sh = start & 31;
        const unsigned* vrow = hdil + lr * ww;
        const unsigned vneed = __funnelshift_r(vrow[wi], wi + 1 < ww ? vrow[wi + 1] : 0u, sh);
        if (((vneed >> 4) & ((1u << OUT) - 1u)) == 0u) continue;  // nobody looks at these columns of this row
        const unsigned* urow = und + lr * ww;
        const unsigned ubits = __funnelshift_r(urow[wi], wi + 1 < ww ? urow[wi + 1] : 0u, sh);
        unsigned k = ((ubits >> lane) & 1u) ? key_l(lr, s0 - 4 + lane) : 0u;
        unsigned o = (unsigned)lane;
        const unsigned k0 = k;
#pragma unroll
        for (int st = 1; st <= 4; st <<= 1) {  // strict compares: the LEFT operand (smaller columns) keeps equal keys
          const unsigned k2 = __shfl_down_sync(0xffffffffu, k, st), o2 = __shfl_down_sync(0xffffffffu, o, st);
          if (k2 > k) {
            k = k2;
            o = o2;
          }
        }
        const unsigned k8 = __shfl_down_sync(0xffffffffu, k0, 8);
        if (k8 > k) {
          k = k8;
          o = (unsigned)lane + 8u;
        }
        const int x = s0 + lane;  // lane i < 24 holds the result for the window [s0 + i - 4, s0 + i + 4]
        if (lane < OUT && x < W) {
          rk[lr * W + x] = k;
          rj[lr * W + x] = (unsigned char)(o - (unsigned)lane);
        }
      }
    } else {
      for (int t = warp; t < nall; t += nwarp) {
        const int lr = t / wpr, w = t % wpr, x = w * 32 + lane;
        unsigned need = 0u;  // undecided bits of OWN rows within d rows of lr, this word
        for (int q = max(lr - d, d); q <= min(lr + d, d + nr - 1); ++q) need |= und[q * ww + w + 1];
        if (need == 0u) continue;
        unsigned best = 0u, bj = 0u;
        if (x < W) {
          unsigned rest = window_bits(und + lr * ww, x, d);
          while (rest) {
            const int j = __ffs(rest) - 1;
            rest &= rest - 1;
            const unsigned kk = key_l(lr, x - d + j);
            if (kk > best) {  // strict: the leftmost of equal keys stays
              best = kk;
              bj = (unsigned)j;
            }
          }
          rk[lr * W + x] = best;
          rj[lr * W + x] = (unsigned char)bj;
        }
      }
    }
    __syncthreads();
    // (B) winners: the topmost row maximum over y-d .. y+d is this pixel itself
    for (int t = warp; t < ntask; t += nwarp) {
      const int ly = t / wpr, w = t % wpr, x = w * 32 + lane;
      const unsigned word = und[(ly + d) * ww + w + 1];
      unsigned m = 0u;
      if (word != 0u) {
        bool win = (word >> lane) & 1u;
        if (win) {
          unsigned best = 0u;
          int br = 0, bj = 0;
          if (DT) {
            unsigned k[2 * DT + 1];
#pragma unroll
            for (int r = 0; r <= 2 * DT; ++r) k[r] = rk[(ly + r) * W + x];
#pragma unroll
            for (int r = 0; r <= 2 * DT; ++r)
              if (k[r] > best) {  // strict: the topmost of equal keys stays
                best = k[r];
                br = r;
              }
            bj = br == DT ? (int)rj[(ly + DT) * W + x] : -1;  // (only the centre row can make this pixel the winner)
          } else {
            for (int r = 0; r <= 2 * d; ++r) {
              const unsigned kk = rk[(ly + r) * W + x];
              if (kk > best) {
                best = kk;
                br = r;
                bj = rj[(ly + r) * W + x];
              }
            }
          }
          win = br == d && bj == d;  // (best is then this pixel's own key)
        }
        m = __ballot_sync(0xffffffffu, win);
      }
      if (lane == 0) {
        newk[(ly + d) * ww + w + 1] = m;
        if (m) kept[(ly + d) * ww + w + 1] |= m;
      }
    }
    __syncthreads();
    push_halo(newk);
    cluster_barrier();
    // (C) clear the winners and everything within d of a winner from the undecided set (bitmap dilation)
    for (int i = tid; i < nall; i += THREADS) {
      const int lr = i / wpr, w = i % wpr;
      const unsigned* row = newk + lr * ww + w + 1;
      const unsigned cur = row[0], prev = row[-1], next = row[1];
      unsigned o = cur;
      if (cur | prev | next)
        for (int s = 1; s <= d; ++s) o |= __funnelshift_l(prev, cur, s) | __funnelshift_r(cur, next, s);
      hdil[lr * ww + w + 1] = o;
    }
    __syncthreads();
    int any = 0;
    for (int i = tid; i < ntask; i += THREADS) {
      const int ly = i / wpr, w = i % wpr;
      unsigned u = und[(ly + d) * ww + w + 1];
      if (u) {
        unsigned v = 0u;
        for (int r = 0; r <= 2 * d; ++r) v |= hdil[(ly + r) * ww + w + 1];
        u &= ~v;
        und[(ly + d) * ww + w + 1] = u;
        any |= (u != 0u);
      }
    }
    any = __syncthreads_or(any);
    push_halo(und);
    if (tid < (int)csize) dsmem_store(&s_flag[crank], tid, (unsigned)any);
    cluster_barrier();
    int go = 0;
    for (unsigned i = 0; i < csize; ++i) go |= s_flag[i];
    if (!go) break;
  }

  // border removal AFTER suppression (model_wrap.py:274-278), compaction with sort keys
  int mycount = 0;
  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool k = ((kept[(ly + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    mycount += __popc(__ballot_sync(0xffffffffu, k));
  }
  if (lane == 0) s_scan[warp] = mycount;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < nwarp; ++i) {
      const int c = s_scan[i];
      s_scan[i] = acc;
      acc += c;
    }
    for (unsigned i = 0; i < csize; ++i) dsmem_store(&s_cnt[crank], i, (unsigned)acc);
  }
  cluster_barrier();
  int cta_base = 0, total = 0;
  for (unsigned i = 0; i < csize; ++i) {
    if (i < crank) cta_base += s_cnt[i];
    total += s_cnt[i];
  }
  if (crank == 0 && tid == 0) p.nkept[b] = total < p.maxk ? total : p.maxk;
  int base = cta_base + s_scan[warp];
  unsigned long long* keys = p.keys + (long long)b * p.maxk;
  for (int t = warp; t < ntask; t += nwarp) {
    const int ly = t / wpr, w = t % wpr, x = w * 32 + lane, y = y0 + ly;
    const bool k = ((kept[(ly + d) * ww + w + 1] >> lane) & 1u) && x >= p.border && x < W - p.border &&
                   y >= p.border && y < H - p.border;
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < p.maxk) keys[slot] = ((unsigned long long)key_l(ly + d, x) << 32) | (unsigned)(y * W + x);
    }
    base += __popc(m);
  }
  cluster_barrier();  // nobody leaves while a peer may still write into its shared memory
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

// cluster size: the largest C <= 8 whose last band still has at least max(d,1) rows
static int nms_cluster(int H, int d, int& rb) {
  const int need = d > 1 ? d : 1;
  for (int c = kMaxCluster; c >= 1; --c) {
    rb = ceil_div(H, c);
    const int last = H - (c - 1) * rb;
    if (last >= need && rb >= need) return c;
  }
  rb = H;
  return 1;
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
  return spb_get_pts_from_heatmap_ex(heat, B, H, W, conf_thresh, nms_dist, border_remove, top_k, capacity, pts, counts,
                                     total, ws, ws_bytes, 0, stream);
}

int spb_get_pts_from_heatmap_ex(const float* heat, int B, int H, int W, float conf_thresh, int nms_dist,
                                int border_remove, int top_k, int capacity, float* pts, int* counts, int* total,
                                void* ws, size_t ws_bytes, int cta_threads, void* stream) {
  SPB_CHECK_ARG(heat && pts && counts && ws);
  SPB_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && capacity >= 1 && border_remove >= 0);
  SPB_CHECK_ARG(nms_dist >= 0 && nms_dist <= kMaxDist);
  SPB_CHECK_ARG(cta_threads == 0 || cta_threads == 256 || cta_threads == 512 || cta_threads == 1024);
  SPB_CHECK_ARG((long long)H * W < (1LL << 31));
  if (ws_bytes < spb_nms_workspace_bytes(B, H, W, nms_dist)) {
    set_error("spb_get_pts_from_heatmap: workspace %zu < %zu bytes", ws_bytes, spb_nms_workspace_bytes(B, H, W, nms_dist));
    return SPB_ENOMEM;
  }
  NmsParams p;
  p.heat = heat;
  p.B = B;
  p.H = H;
  p.W = W;
  p.thr = conf_thresh;
  p.d = nms_dist;
  p.border = border_remove;
  p.ww = ceil_div(W, 32) + 2;
  const int C = nms_cluster(H, nms_dist, p.rb);
  const size_t lrows = (size_t)p.rb + 2 * nms_dist;
  // algo 1 (default): separable window maxima (nms2_kernel); 0: per-candidate window scans (nms_kernel).
  int algo = 1;
  if (const char* e = getenv("SPB_NMS_ALGO")) algo = atoi(e) != 0;
  const size_t key_bytes = lrows * W * sizeof(unsigned);
  const size_t cap = 200 * 1024;
  size_t bm_bytes = 2 * lrows * p.ww * sizeof(unsigned);
  if (algo == 1) {  // + newk, hdil bitmaps, row maxima (key + offset)
    const size_t need = 4 * lrows * p.ww * sizeof(unsigned) + key_bytes + ((lrows * W + 15) & ~(size_t)15);
    if (need > cap)
      algo = 0;  // a band of this image does not fit with the row maxima: the scan kernel needs less
    else
      bm_bytes = need;
  }
  if (bm_bytes > cap) {
    set_error("spb_get_pts_from_heatmap: image %dx%d needs %zu B of bitmap shared memory per CTA (> 200 KB)", H, W, bm_bytes);
    return SPB_EINVAL;
  }
  // the keys of band + halo live in shared memory when two CTAs of this size still fit on an SM (or, for the scan
  // kernel, whenever they fit at all)
  const bool one_cta_per_sm = (long long)B * C <= kNumSMs || cta_threads == 1024;
  p.keys_smem = (bm_bytes + key_bytes <= (algo == 1 && !one_cta_per_sm ? 100 * 1024 : cap)) ? 1 : 0;
  const size_t smem = bm_bytes + (p.keys_smem ? key_bytes : 0);
  p.maxk = nms_maxk(H, W, nms_dist);
  p.keys = static_cast<unsigned long long*>(ws);
  p.nkept = reinterpret_cast<int*>(p.keys + (size_t)B * p.maxk);
  cudaStream_t st = as_stream(stream);
  // One CTA of 1024 threads per SM holds 148 / C images at a time; when the batch has more CTAs than that, smaller
  // CTAs (2 or 4 per SM, registers and shared memory permitting) keep other images running while one waits at its
  // cluster barrier.  SPB_NMS_THREADS overrides (A/B measurements).
  int threads = kNmsThreads;
  if (cta_threads == 1024 || cta_threads == 512 || cta_threads == 256) {
    threads = cta_threads;  // the caller knows better (a launch that runs BESIDE other kernels: 1024)
  } else {
    if ((long long)B * C > kNumSMs && smem <= 100 * 1024) threads = 512;
    if ((long long)B * C > 2 * kNumSMs && smem <= 50 * 1024) threads = 256;
  }
  if (const char* e = getenv("SPB_NMS_THREADS")) {
    const int t = atoi(e);
    if (t == 256 || t == 512 || t == 1024) threads = t;
  }
  void (*kern)(NmsParams);
  if (algo == 1 && nms_dist == 4)  // the reference's distance: unrolled taps
    kern = threads == 1024 ? nms2_kernel<1024, 4> : threads == 512 ? nms2_kernel<512, 4> : nms2_kernel<256, 4>;
  else if (algo == 1)
    kern = threads == 1024 ? nms2_kernel<1024, 0> : threads == 512 ? nms2_kernel<512, 0> : nms2_kernel<256, 0>;
  else
    kern = threads == 1024 ? nms_kernel<1024> : threads == 512 ? nms_kernel<512> : nms_kernel<256>;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * C);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  SPB_CHECK_LAUNCH();
  dim3 grid(ceil_div(p.maxk, 256), B);
  rank_scatter_kernel<<<grid, 256, 0, st>>>(heat, H, W, p.keys, p.nkept, p.maxk, top_k, capacity, pts, counts, total);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
