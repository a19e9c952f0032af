// (2) 64 -> 64 channel 3x3 convolution with the three COLUMN taps stacked along the UMMA N dimension (N = 192).
// Replaces conv2a / conv2b (and conv1b when the first block is not fused) of
// models/SuperPointNet_pretrained.py:57-61, SuperPointNet_gauss2.py:53-56 / unet_parts.py:14-21,41-44.
//
// Why: an M128 x N64 x K16 tcgen05.mma reads 4 KB of A and 2 KB of B from shared memory for 32 cycles of tensor
// work -- 192 B/clk against the 128 B/clk the SM can deliver, so the shifted-window kernel (conv_tc.cu) tops out
// at 2/3 of the tensor peak for Cout = 64 (measured: tools/ubench_mma.cu, 48.0 cycles per MMA) and at ~55 % with
// the rest of the kernel's shared-memory traffic.  Here ONE A view serves three taps:
//     D[m, (s, co)] = sum_{r, ci} A_r[m, ci] * W[co][r][s][ci]          (3 x 4 MMAs of M128 x N192 x K16 per tile)
// where row m = (ry, wx) is the input pixel (y0 - 1 + ry + r, x0 - 1 + wx) of a 16-pixel-wide window, so that
//     out(y0 + ry, x0 + ox) = D[(ry, ox), s=0] + D[(ry, ox + 1), s=1] + D[(ry, ox + 2), s=2]      ox = 0..13.
// The A operand is read 12 times per tile instead of 36 (10 KB per 96-cycle MMA = 107 B/clk: full tensor rate);
// the price is 14 useful columns out of 16 (87.5 %) and re-aligning the s=1 / s=2 partial sums by one / two TMEM
// lanes in the epilogue (two shuffles per channel).  Every SHFL is one wavefront of the shared-memory pipe the tensor
// core reads its operands through (ncu r01: 512 of ~2000 wavefronts per tile), so the net gain over the shifted-
// window kernel is modest.  The alternative of re-aligning INSIDE tensor memory with tcgen05.shift.down (row i <-
// row i+1 within each 32-lane quadrant, 8 columns per instruction) was measured and rejected: 24 shifts per tile
// cost 15.6 cycles each, serialised with the MMAs (tools/ubench_mma.cu "s3 tile": 1526 vs 1152 cycles per tile).
//   * tile = 8 rows x 14 columns of outputs; A = ONE TMA box of 10 rows x 16 pixels x 64 channels (zero filled
//     outside the image = the conv padding); tap row r is the same block shifted by r*16 pixels (2 KB, keeps the
//     1024-byte swizzle atoms aligned); UMMA row m = ry*16 + wx, 8-row groups are 1 KB apart (SBO);
//   * TMEM lane m sits in quadrant ry/2 at lane (ry%2)*16 + wx: the column shift (lane+1, +2) never leaves the
//     quadrant for ox < 14, and both partners of the 2x2 max-pool (lane^1, lane^16) are in the same warp;
//   * weights [r][(s, co)][ci] stay resident in shared memory (72 KB), accumulators double buffered in TMEM
//     (2 x 192 columns), two MMA issuer warps alternate tiles with a token hand-over (see conv_tc.cu).
// Warps: 0 TMA producer, 1 and 18 MMA issuers, 2..17 epilogue in two groups of 8 that alternate tiles.
#include "tc_common.cuh"

namespace spb {

#ifndef SPB_S3_NA
#define SPB_S3_NA 4
#endif
constexpr int kS3Rows = 8, kS3Cols = 14, kS3Win = 16, kS3NA = SPB_S3_NA;

struct S3Smem {
  uint8_t b[3][24576];         // per tap row: 192 rows (s, co) x 128 B, SWIZZLE_128B
  uint8_t a[kS3NA][20480];     // 10 rows x 16 pixels x 128 B, SWIZZLE_128B
  uint8_t stage[2][2][14336];  // [epilogue group][buffer]: output tile 8 x 14 pixels x 128 B (pooled: 4 x 7),
                               // SWIZZLE_128B, TMA store source
  float bias[64];
  uint64_t bars[3 + 2 * kS3NA + 6];  // b_full[3], a_full[NA], a_empty[NA], acc_full[2], acc_empty[2], token[2]
  uint32_t tmem_slot;
};

struct S3Args {
  const float* bias;
  __nv_bfloat16* out;
  int B, H, W, tiles_x, tiles_y, ntiles;
};

#ifndef SPB_S3_TOKEN_R
#define SPB_S3_TOKEN_R 0  // the issue token is handed over behind this tap row (0, 1 and 2 measure the same)
#endif
constexpr int kS3EpiWarps = 16, kS3Threads = 64 + 32 * kS3EpiWarps + 32;

template <bool POOL>
__global__ void __launch_bounds__(kS3Threads, 1) conv_s3_kernel(const __grid_constant__ CUtensorMap tm_act,
                                                         const __grid_constant__ CUtensorMap tm_wgt,
                                                         const __grid_constant__ CUtensorMap tm_out, S3Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S3Smem& sm = *reinterpret_cast<S3Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(sm.bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a_full = [&](int s) { return bar0 + 8u * (3 + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (3 + kS3NA + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (3 + 2 * kS3NA + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (3 + 2 * kS3NA + 2 + s); };
  auto token = [&](int s) { return bar0 + 8u * (3 + 2 * kS3NA + 4 + s); };

  if (threadIdx.x < 64) sm.bias[threadIdx.x] = a.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(b_full(s), 1);
    for (int s = 0; s < kS3NA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kS3EpiWarps / 2);
      mbar_init(token(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;
  const int tiles_per_view = a.tiles_x * a.tiles_y;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (int r = 0; r < 3; ++r) {
        mbar_expect_tx(b_full(r), 24576);
        tma_load_2d(smem_u32(sm.b[r]), &tm_wgt, b_full(r), 0, r * 192);
      }
      int it = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
        const int n = t / tiles_per_view, rem = t % tiles_per_view;
        const int y0 = (rem / a.tiles_x) * kS3Rows, x0 = (rem % a.tiles_x) * kS3Cols;
        const int sa = it % kS3NA, pa = (it / kS3NA) & 1;
        mbar_wait(a_empty(sa), pa ^ 1);
        mbar_expect_tx(a_full(sa), 20480);
        tma_load_4d(smem_u32(sm.a[sa]), &tm_act, a_full(sa), 0, x0 - 1, y0 - 1, n);
      }
    }
  } else if (warp == 1 || warp == 2 + kS3EpiWarps) {
    // =========================== MMA issuers (alternate tiles, token hand-over after tap row 0) ===========
    const int q = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = make_idesc(128, 192);
    constexpr uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t d_tmem = tmem_base + (uint32_t)q * 192;
    int k = 0;
    for (int t = blockIdx.x + q * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++k) {
      const int it = 2 * k + q;
      const int sa = it % kS3NA, pa = (it / kS3NA) & 1;
      mbar_wait(acc_empty(q), (k & 1) ^ 1);
      mbar_wait(a_full(sa), pa);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sm.a[sa]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (k == 0) {
          mbar_wait(b_full(r), 0);
          tc_fence_after();
        }
        const uint64_t adesc = desc_hi | (uint64_t)(((a0 + r * (kS3Win * 128)) >> 4) | (1u << 16));
        const uint64_t bdesc = desc_hi | (uint64_t)((smem_u32(sm.b[r]) >> 4) | (1u << 16));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (r | kk) != 0 ? 1u : 0u);
          if (r == SPB_S3_TOKEN_R) mbar_arrive(token(q ^ 1));
          if (r == 2) {
            tc_commit(a_empty(sa));
            tc_commit(acc_full(q));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (warps 2..17): two groups of 8 warps alternate tiles ===========
    // One warp's chain per tile (barrier wait -> tcgen05.ld -> shuffle-add -> pack -> pool -> stage -> fence ->
    // named barrier -> TMA store) is ~1400 cycles of latency against 1152 MMA cycles per tile, so a single group
    // processing every tile in lock-step was what the tensor pipe waited for (measured: the kernel WITHOUT its
    // MMAs ran at 70 % of the full kernel's time).  Group g owns the tiles of parity g = accumulator stage g, two
    // staging buffers of its own and named barrier 1+g; each warp converts 32 channels of its TMEM quadrant.
    const int grp = (warp - 2) >> 3, quad = warp & 3, chalf = ((warp - 2) >> 2) & 1;
    const int ry = 2 * quad + (lane >> 4), ox = lane & 15;  // tile row / window column of this thread's TMEM lane
    const bool store_thread = threadIdx.x == 64 + grp * 256;
    // row of the staging tile this thread owns (the TMA store clips what lies outside the image)
    const bool owner = ox < kS3Cols && (!POOL || !(lane & 17));  // pooled: even column, even row of the window
    const int p = POOL ? (ry >> 1) * (kS3Cols / 2) + (ox >> 1) : ry * kS3Cols + ox;
    const float* bz = sm.bias + chalf * 32;  // (registers are the scarcer resource at 608 threads: 4 LDS.128 per batch)
    int k = 0;  // this group's tile counter
    for (int t = blockIdx.x + grp * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++k) {
      const int n = t / tiles_per_view, rem = t % tiles_per_view;
      const int y0 = (rem / a.tiles_x) * kS3Rows, x0 = (rem % a.tiles_x) * kS3Cols;
      mbar_wait(acc_full(grp), k & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)grp * 192 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      uint32_t pk[16];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {  // two batches of 16 channels (48 live accumulator registers at a time)
        uint32_t c[3][16];
#pragma unroll
        for (int s = 0; s < 3; ++s) tc_ld_nowait<16>(taddr + s * 64 + cb * 16, c[s]);
        tc_wait_ld();
        if (cb == 1) {  // everything this warp needs has left the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(grp));
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bz + cb * 16 + j);
          const float bj[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float c1 = __shfl_down_sync(0xffffffffu, __uint_as_float(c[1][j + e]), 1);
            const float c2 = __shfl_down_sync(0xffffffffu, __uint_as_float(c[2][j + e]), 2);
            f[j + e] = ((__uint_as_float(c[0][j + e]) + c1) + c2) + bj[e];
          }
        }
        s3_pack16<POOL>(f, pk + cb * 8);
      }
      // Stage the tile in shared memory (SWIZZLE_128B rows = pixels) and let ONE TMA tensor store write it:
      // per-lane 16-byte global stores at a 128-byte stride cost one L1 wavefront per lane and competed with the
      // tensor core's operand reads for the shared-memory pipe (78 % L1/TEX busy, ncu r01).
      uint8_t* stg = sm.stage[grp][k & 1];
      if (owner) {
        uint8_t* srow = stg + p * 128;
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<uint4*>(srow + (((chalf * 4 + v) ^ (p & 7)) * 16)) =
              make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the group's previous store (its other staging buffer) must have drained before anyone passes this
      // barrier and starts writing that buffer for the group's next tile
      if (store_thread) tma_store_wait_read<0>();
      named_bar_sync(1 + grp, 256);
      if (store_thread) {
        if (POOL)
          tma_store_4d(&tm_out, smem_u32(stg), 0, x0 >> 1, y0 >> 1, n);
        else
          tma_store_4d(&tm_out, smem_u32(stg), 0, x0, y0, n);
        tma_store_commit();
      }
    }
    if (store_thread) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_s3 = 1;  // 0: 64->64 layers fall back to the shifted-window kernel (A/B tests)
extern "C" int spb_conv_tc_set_s3(int on) {
  g_s3 = on ? 1 : 0;
  return SPB_OK;
}
bool conv_s3_enabled() { return g_s3 != 0; }

// w_s3: bf16 [r][s][co][ci] = 576 rows x 64, K-major; one TMA box = one tap row (192 rows)
int conv_s3_prepare(LayerDev& L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv_s3: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  L.tmap_w3 = malloc(sizeof(CUtensorMap) + 64);
  if (!L.tmap_w3) return SPB_ENOMEM;
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, 576};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 192};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w_s3, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_s3: cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return SPB_ECUDA;
  }
  memcpy(L.tmap_w3, &tm, sizeof(tm));
  return SPB_OK;
}

template <bool POOL>
static int launch_s3(const LayerDev& L, const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W,
                     cudaStream_t st) {
  CUtensorMap tm_act, tm_wgt, tm_out;
  memcpy(&tm_wgt, L.tmap_w3, sizeof(tm_wgt));
  int rc = encode_act_map(&tm_act, in, B, H, W, 64, kS3Win, kS3Rows + 2);
  if (rc != SPB_OK) return rc;
  rc = POOL ? encode_act_map(&tm_out, out, B, H / 2, W / 2, 64, kS3Cols / 2, kS3Rows / 2)
            : encode_act_map(&tm_out, out, B, H, W, 64, kS3Cols, kS3Rows);
  if (rc != SPB_OK) return rc;
  S3Args a;
  a.bias = L.bias;
  a.out = out;
  a.B = B;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, kS3Cols);
  a.tiles_y = ceil_div(H, kS3Rows);
  a.ntiles = B * a.tiles_x * a.tiles_y;
  const int smem = (int)sizeof(S3Smem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv_s3_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv_s3_kernel<POOL><<<grid, kS3Threads, smem, st>>>(tm_act, tm_wgt, tm_out, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int conv_s3(const LayerDev& L, const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st) {
  if (!L.tmap_w3 || L.cin != 64 || L.cout != 64 || L.ks != 3 || !L.relu) {
    set_error("conv_s3: layer is not a 64->64 3x3 ReLU convolution with prepared weights");
    return SPB_EINVAL;
  }
  if (L.pool && ((H | W) & 1)) {
    set_error("conv_s3: pooled layer needs even H, W (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  return L.pool ? launch_s3<true>(L, in, out, B, H, W, st) : launch_s3<false>(L, in, out, B, H, W, st);
}

}  // namespace spb
