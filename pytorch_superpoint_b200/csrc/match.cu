// NEXT row f-1: two-way nearest-neighbour descriptor matching.
// Replaces PointTracker.nn_match_two_way, models/model_wrap.py:437-480:
//   dmat = sqrt(2 - 2*clip(desc1^T desc2, -1, 1)); idx = argmin over rows, idx2 = argmin over columns;
//   keep i if dmat[i, idx[i]] < nn_thresh and idx2[idx[i]] == i.
// One pass over the K1 x K2 score matrix, never materialised: 64x64 tiles of fp32 FMAs (descriptors are fp32 and the
// distances feed comparisons, so no bf16 here), each tile folds its row / column minima into 64-bit keys
// (distance bits << 32 | index) with atomicMin -- distances are >= 0, so the integer order is the float order and
// ties resolve to the LOWEST index exactly like np.argmin.  A second single-block kernel applies the threshold and
// the mutual check and compacts the survivors in ascending idx1 order (ballot + prefix count).
#include "common.cuh"

namespace spb {

constexpr int kMT = 64;  // tile edge
constexpr int kMK = 16;  // depth of one shared-memory stage

// POINT_MAJOR false: desc1 [D, K1], desc2 [D, K2] row-major (the reference's M x N arrays);
// POINT_MAJOR true: desc [K, D] rows (what the descriptor sampler writes), blockIdx.z = pair of a batch
// (desc1 / desc2 [P,K,D], keys [P][K1 + K2]).  Keys pre-filled with ~0.
template <bool POINT_MAJOR>
__global__ void __launch_bounds__(256) match_tile_kernel(const float* __restrict__ d1, const float* __restrict__ d2,
                                                         int K1, int K2, int D, unsigned long long* __restrict__ rowkey,
                                                         unsigned long long* __restrict__ colkey) {
  __shared__ float sa[kMK][kMT + 4], sb[kMK][kMT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * kMT, j0 = blockIdx.x * kMT;
  if (POINT_MAJOR) {
    d1 += (long long)blockIdx.z * K1 * D;
    d2 += (long long)blockIdx.z * K2 * D;
    rowkey += (long long)blockIdx.z * (K1 + K2);
    colkey += (long long)blockIdx.z * (K1 + K2);
  }
  float acc[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += kMK) {
    for (int e = threadIdx.x; e < kMK * kMT; e += 256) {
      if (POINT_MAJOR) {  // 16 consecutive threads read the 16 channels of one point: 64-byte segments
        const int k = e % kMK, c = e / kMK;
        sa[k][c] = (k0 + k < D && i0 + c < K1) ? __ldg(d1 + (long long)(i0 + c) * D + k0 + k) : 0.f;
        sb[k][c] = (k0 + k < D && j0 + c < K2) ? __ldg(d2 + (long long)(j0 + c) * D + k0 + k) : 0.f;
      } else {
        const int k = e / kMT, c = e % kMT;
        sa[k][c] = (k0 + k < D && i0 + c < K1) ? __ldg(d1 + (long long)(k0 + k) * K1 + i0 + c) : 0.f;
        sb[k][c] = (k0 + k < D && j0 + c < K2) ? __ldg(d2 + (long long)(k0 + k) * K2 + j0 + c) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = sa[k][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = sb[k][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
  // distances, then minima: per thread over its 4x4 block, then atomics (<= 4 row + 4 column keys per thread)
  unsigned long long rk[4], ck[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) rk[r] = ck[r] = ~0ull;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
      if (i < K1 && j < K2) {
        const float s = fminf(fmaxf(acc[r][c], -1.f), 1.f);
        const float d = __fsqrt_rn(__fsub_rn(2.f, __fmul_rn(2.f, s)));
        const unsigned long long bits = (unsigned long long)__float_as_uint(d) << 32;
        rk[r] = min(rk[r], bits | (unsigned)j);
        ck[c] = min(ck[c], bits | (unsigned)i);
      }
    }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    // the 16 threads of a row group are 16 consecutive lanes: fold them before touching memory
    unsigned long long v = rk[r];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int i = i0 + ty * 4 + r;
    if (tx == 0 && i < K1 && v != ~0ull) atomicMin(rowkey + i, v);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = j0 + tx * 4 + c;
    if (j < K2 && ck[c] != ~0ull) atomicMin(colkey + j, ck[c]);
  }
}

// blockIdx.x = pair of a batch (0 for the single-pair entry): keys advance by key_stride, outputs by out_stride
__global__ void __launch_bounds__(1024) match_select_kernel(const unsigned long long* __restrict__ rowkey,
                                                            const unsigned long long* __restrict__ colkey, int K1,
                                                            float thr, int* __restrict__ idx1, int* __restrict__ idx2,
                                                            float* __restrict__ score, int* __restrict__ count,
                                                            long long key_stride, long long out_stride) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  rowkey += blockIdx.x * key_stride;
  colkey += blockIdx.x * key_stride;
  idx1 += blockIdx.x * out_stride;
  idx2 += blockIdx.x * out_stride;
  score += blockIdx.x * out_stride;
  count += blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < K1; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    int j = 0;
    float sc = 0.f;
    if (i < K1) {
      const unsigned long long rk = rowkey[i];
      j = (int)(unsigned)rk;
      sc = __uint_as_float((unsigned)(rk >> 32));
      keep = sc < thr && (int)(unsigned)colkey[j] == i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (keep) {
      const int o = off + __popc(m & ((1u << lane) - 1));
      idx1[o] = i;
      idx2[o] = j;
      score[o] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_cnt[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

}  // namespace spb

using namespace spb;

extern "C" {

size_t spb_nn_match_workspace_bytes(int K1, int K2) {
  if (K1 < 0 || K2 < 0) return 0;
  return ((size_t)K1 + (size_t)K2) * sizeof(unsigned long long) + 256;
}

int spb_nn_match_two_way(const float* desc1, int K1, const float* desc2, int K2, int D, float nn_thresh, int* idx1,
                         int* idx2, float* score, int* count, void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(count && K1 >= 0 && K2 >= 0 && D >= 1);
  SPB_CHECK_ARG(nn_thresh >= 0.f);  // models/model_wrap.py:455-456 raises ValueError
  cudaStream_t st = as_stream(stream);
  if (K1 == 0 || K2 == 0) {  // models/model_wrap.py:453-454: no matches
    SPB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    return SPB_OK;
  }
  SPB_CHECK_ARG(desc1 && desc2 && idx1 && idx2 && score && ws);
  if (ws_bytes < spb_nn_match_workspace_bytes(K1, K2)) {
    set_error("spb_nn_match_two_way: workspace %zu < %zu bytes", ws_bytes, spb_nn_match_workspace_bytes(K1, K2));
    return SPB_ENOMEM;
  }
  unsigned long long* rowkey =
      reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  unsigned long long* colkey = rowkey + K1;
  SPB_CHECK_CUDA(cudaMemsetAsync(rowkey, 0xff, ((size_t)K1 + K2) * sizeof(unsigned long long), st));
  dim3 grid(ceil_div(K2, kMT), ceil_div(K1, kMT));
  match_tile_kernel<false><<<grid, 256, 0, st>>>(desc1, desc2, K1, K2, D, rowkey, colkey);
  SPB_CHECK_LAUNCH();
  match_select_kernel<<<1, 1024, 0, st>>>(rowkey, colkey, K1, nn_thresh, idx1, idx2, score, count, 0, 0);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// export.py:147-151 for a batch of image pairs without a host round trip: desc1 / desc2 [P,K,D] in the sampler's
// point-major layout (rows behind an image's key-point count are zero: distance sqrt(2) to everything, so they never
// pass nn_thresh < sqrt(2) nor win a mutual check against a real descriptor below it).  idx1 / idx2 / score [P,K],
// count [P] (device).
size_t spb_nn_match_pairs_workspace_bytes(int P, int K) {
  if (P < 0 || K < 0) return 0;
  return (size_t)P * 2 * K * sizeof(unsigned long long) + 256;
}

int spb_nn_match_pairs(const float* desc1, const float* desc2, int P, int K, int D, float nn_thresh, int* idx1,
                       int* idx2, float* score, int* count, void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(count && P >= 0 && P <= 65535 && K >= 0 && D >= 1);
  SPB_CHECK_ARG(nn_thresh >= 0.f && nn_thresh < 1.41421f);  // (zero padding relies on it)
  cudaStream_t st = as_stream(stream);
  if (P == 0) return SPB_OK;
  if (K == 0) {
    SPB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * P, st));
    return SPB_OK;
  }
  SPB_CHECK_ARG(desc1 && desc2 && idx1 && idx2 && score && ws);
  if (ws_bytes < spb_nn_match_pairs_workspace_bytes(P, K)) {
    set_error("spb_nn_match_pairs: workspace %zu < %zu bytes", ws_bytes, spb_nn_match_pairs_workspace_bytes(P, K));
    return SPB_ENOMEM;
  }
  unsigned long long* rowkey =
      reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  SPB_CHECK_CUDA(cudaMemsetAsync(rowkey, 0xff, (size_t)P * 2 * K * sizeof(unsigned long long), st));
  dim3 grid(ceil_div(K, kMT), ceil_div(K, kMT), P);
  match_tile_kernel<true><<<grid, 256, 0, st>>>(desc1, desc2, K, K, D, rowkey, rowkey + K);
  SPB_CHECK_LAUNCH();
  match_select_kernel<<<P, 1024, 0, st>>>(rowkey, rowkey + K, K, nn_thresh, idx1, idx2, score, count, 2LL * K, K);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
