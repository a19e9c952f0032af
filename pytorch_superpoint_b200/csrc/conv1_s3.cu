// EXPERIMENTAL, OFF BY DEFAULT (spb_conv_tc_set_fuse1(2) selects it; parity-tested, bit-equal to the conv1a kernel +
// conv_s3): conv1a + conv1b + 2x2 max-pool in one kernel, conv1b with the three COLUMN taps stacked along the UMMA N
// dimension (N = 192) -- the first VGG block of models/SuperPointNet_pretrained.py:55-58,
// SuperPointNet_gauss2.py:50-52 / unet_parts.py:14-21.
// Measured 15 % SLOWER than conv1_fused_kernel on the whole forward (8.70 vs 7.56 ms per 800 views 240x320): the
// stacked epilogue + conv1a epilogue + im2col build execute ~4700 warp instructions per 112-pixel tile, as many issue
// cycles as the 1250 MMA cycles they were meant to hide behind, and tcgen05.mma issue blocks while the pipe is busy
// (profiles/r01_conv1_s3_probe.txt).  Kept as the measured record of that design point.
//
// conv1_fused_kernel (conv_tc.cu) issues conv1b as 36 M128 x N64 x K16 MMAs per 16x8 tile, and an N = 64 MMA is
// bound by its shared-memory operand reads (48 cycles instead of 32, tools/ubench_mma.cu): 16 cycles per output
// pixel at best.  conv_s3.cu showed that 12 N = 192 MMAs over a 16-pixel-wide window run at the full tensor rate;
// this kernel gives the fused block the same MMA shape:
//   * tile = 8 rows x 14 columns of conv1b outputs (4 x 7 pooled); conv1b's A operand = conv1a's output for the
//     10 x 16 halo pixels of the window (row p = hy*16 + hx, 128 B each, SWIZZLE_128B), written by stage 1;
//   * conv1a for those 160 pixels = two M128 x N64 x K16 MMAs over a K=16 im2col tile (9 taps + 2 bias slots; a
//     halo pixel outside the image gets an all-zero row: conv1b zero-pads conv1a's OUTPUT) into ONE 128-column
//     accumulator (TMEM: 2 x 192 columns for conv1b + 128 for conv1a = 512); pixels 128..159 are TMEM lanes 0..31
//     of the second MMA, so the two stage-1 warps of quadrant 0 convert twice as many rows as the others;
//   * the zero-padded fp32 patch under a halo block (12 rows x 18 columns) arrives by TMA as a 12 x 20 box whose
//     first column is (x0 - 2) rounded down to a multiple of 4 (TMA: 16-byte aligned innermost start);
//   * conv1b: per tap row r one A view (the halo block shifted by r*16 pixels) against 192 weight rows (s, co);
//     the epilogue re-aligns the s = 1, 2 partial sums by one / two TMEM lanes with shuffles (see conv_s3.cu),
//     adds the bias, packs bf16, ReLU, 2x2 max-pool inside the warp, stages 28 pooled pixels x 128 B and ONE TMA
//     tensor store writes them (clipped at the image border).
// Warps: 0 (also loads the weights) and 1 MMA issuers, 2..17 conv1b epilogue in two groups of 8 that alternate
// tiles (a warp converts 32 channels of its TMEM quadrant) and, two tiles ahead, build the im2col tile of their next
// tile, 18..25 conv1a epilogue (two warps per quadrant, 32 channels each).  832 threads -> 72 registers per thread.
// MMA order: conv1a(i+1) is issued by the warp that will issue conv1b(i), before it; see the issuer section.
#include "tc_common.cuh"

namespace spb {

constexpr int kC1EpiWarps = 16, kC1Threads = 64 + 32 * kC1EpiWarps + 256;
constexpr int kC1Rows = 8, kC1Cols = 14, kC1Win = 16, kC1Halo = (kC1Rows + 2) * kC1Win;  // 160 halo pixels

struct C1S3Smem {
  uint8_t b[3][24576];         // conv1b weights per tap row: 192 rows (s, co) x 128 B, SWIZZLE_128B (TMA)
  uint8_t a2[2][20480];        // conv1b A operand: 10 x 16 halo pixels x 128 B, SWIZZLE_128B, written by stage 1
  uint8_t ostage[2][2][4096];  // [epilogue group][buffer]: 4 x 7 pooled pixels x 128 B, SWIZZLE_128B (TMA store)
  uint8_t a1[2][8192];         // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices (rows >= 160 stay zero)
  uint8_t b1[2048];            // conv1a weights [64][16]
  float patch[4][256];         // zero-padded 12 x 20 fp32 patch (rows y0-2.., cols ((x0-2) & ~3)..), TMA
  float bias2[64];
  uint64_t bars[24];  // b_full[3] a1_full[2] a1_empty[2] acc1_full acc1_empty[2] a2_full[2] a2_empty[2] acc2_full[2] acc2_empty[2] token[2] patch_full[4]
  uint32_t tmem_slot;
};

struct C1S3Args {
  const __nv_bfloat16* w1a;  // [64][16] (taps 0..8, bias head / tail in slots 9, 10)
  const float* bias2;
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;  // optional timeline probe (CTA 0): [64 tiles][16 slots] of clock64, or NULL
};

__global__ void __launch_bounds__(kC1Threads, 1) conv1_s3_kernel(const __grid_constant__ CUtensorMap tm_w2,
                                                                 const __grid_constant__ CUtensorMap tm_out,
                                                                 const __grid_constant__ CUtensorMap tm_x, C1S3Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  C1S3Smem& sm = *reinterpret_cast<C1S3Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x;
  const uint32_t bar0 = smem_u32(sm.bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a1_full = [&](int s) { return bar0 + 8u * (3 + s); };
  auto a1_empty = [&](int s) { return bar0 + 8u * (5 + s); };
  const uint32_t acc1_full = bar0 + 8u * 7;
  auto acc1_empty = [&](int s) { return bar0 + 8u * (8 + s); };  // by tile parity: each issuer waits every phase of its own
  auto a2_full = [&](int s) { return bar0 + 8u * (10 + s); };
  auto a2_empty = [&](int s) { return bar0 + 8u * (12 + s); };
  auto acc2_full = [&](int s) { return bar0 + 8u * (14 + s); };
  auto acc2_empty = [&](int s) { return bar0 + 8u * (16 + s); };
  auto token = [&](int s) { return bar0 + 8u * (18 + s); };
  auto patch_full = [&](int s) { return bar0 + 8u * (20 + s); };

  if (threadIdx.x < 64) {
    sm.bias2[threadIdx.x] = a.bias2[threadIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(a.w1a + threadIdx.x * 16);
    const int n = threadIdx.x;
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + (n & 7) * 16) = src[0];
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + 128 + (n & 7) * 16) = src[1];
  }
  for (int i = threadIdx.x; i < (int)(sizeof(sm.a1) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(sm.a1)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(b_full(s), 1);
    mbar_init(acc1_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a1_full(s), 256);
      mbar_init(a1_empty(s), 1);
      mbar_init(acc1_empty(s), 8);
      mbar_init(a2_full(s), 256);
      mbar_init(a2_empty(s), 1);
      mbar_init(acc2_full(s), 1);
      mbar_init(acc2_empty(s), kC1EpiWarps / 2);
      mbar_init(token(s), 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(patch_full(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;
  // TMEM columns: conv1b accumulators [0,192) [192,384); conv1a accumulator [384,448) (pixels 0..127), [448,512) (128..159)
  const int tiles_per_view = a.tiles_x * a.tiles_y;
#define C1_PROBE(tile, slot)                                                                          \
  if (a.dbg && blockIdx.x == 0 && (tile) < 64 && lane == 0 && (warp <= 2 || warp == 10 || warp == 18)) \
  a.dbg[(tile) * 16 + (slot)] = clock64()

  if (warp <= 1) {
    // =========================== MMA issuers (warp 0 also loads the conv1b weights first) ===========================
    // tcgen05.mma issue BLOCKS while the tensor pipe is busy (measured: issuing the 12 MMAs of one conv1b takes their
    // 1150 execution cycles), so whatever an issuer warp does after its conv1b starts a tile late.  Issuer q owns
    // conv1b of the tiles of parity q; conv1a(i+1) is issued by the issuer of tile i BEFORE its conv1b(i) -- that
    // warp is idle while the other one is blocked in conv1b(i-1), so conv1a(i+1) slips in between the other warp's
    // MMAs as soon as the conv1a epilogue has read conv1a(i) out of the single conv1a accumulator, and the conv1a
    // epilogue of tile i+1 has the whole of conv1b(i) to produce conv1b(i+1)'s operand.  The two conv1b's are kept
    // in order by a token handed over behind the first tap row.
    const int q = warp;
    if (q == 0 && lane == 0) {
      for (int r = 0; r < 3; ++r) {
        mbar_expect_tx(b_full(r), 24576);
        tma_load_2d(smem_u32(sm.b[r]), &tm_w2, b_full(r), 0, r * 192);
      }
    }
    __syncwarp();
    constexpr uint32_t idesc1 = make_idesc(128, 64), idesc2 = make_idesc(128, 192);
    constexpr uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t b1desc = make_desc_noswz(smem_u32(sm.b1), 128, 256);
    const uint32_t acc1 = tmem_base + 384, acc2 = tmem_base + (uint32_t)q * 192;
    const uint32_t a2 = smem_u32(sm.a2[q]);
    auto issue_conv1a = [&](int buf) {  // operand a1[buf]; the caller has waited for it and for the accumulator
      const uint64_t a1desc = make_desc_noswz(smem_u32(sm.a1[buf]), 128, 256);
      tc_fence_after();
      if (elect_one()) {
        tc_mma(acc1, a1desc, b1desc, idesc1, 0u);
        tc_mma(acc1 + 64, a1desc + (uint64_t)(4096 >> 4), b1desc, idesc1, 0u);
        tc_commit(acc1_full);
        tc_commit(a1_empty(buf));
      }
      __syncwarp();
    };
    if (q == 0 && (int)blockIdx.x < a.ntiles) {  // conv1a(0)
      mbar_wait(a1_full(0), 0);
      issue_conv1a(0);
    }
    int k = 0;
    for (int t = blockIdx.x + q * G; t < a.ntiles; t += 2 * G, ++k) {
      const int it = 2 * k + q;
      C1_PROBE(it, 0);
      if (t + G < a.ntiles) {  // conv1a(it+1): needs conv1a(it) read out (k-th phase of acc1_empty(q)) and its im2col tile
        mbar_wait(acc1_empty(q), k & 1);
        mbar_wait(a1_full(q ^ 1), ((it + 1) >> 1) & 1);
        issue_conv1a(q ^ 1);
      }
      C1_PROBE(it, 1);
      mbar_wait(acc2_empty(q), (k & 1) ^ 1);
      mbar_wait(a2_full(q), k & 1);
      C1_PROBE(it, 2);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      C1_PROBE(it, 3);
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (k == 0) {
          mbar_wait(b_full(r), 0);
          tc_fence_after();
        }
        const uint64_t adesc = desc_hi | (uint64_t)(((a2 + r * (kC1Win * 128)) >> 4) | (1u << 16));
        const uint64_t bdesc = desc_hi | (uint64_t)((smem_u32(sm.b[r]) >> 4) | (1u << 16));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(acc2, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc2, (r | kk) != 0 ? 1u : 0u);
          if (r == 0) mbar_arrive(token(q ^ 1));
          if (r == 2) {
            tc_commit(a2_empty(q));
            tc_commit(acc2_full(q));
          }
        }
        __syncwarp();
      }
      C1_PROBE(it, 4);
    }
  } else if (warp >= 2 + kC1EpiWarps) {
    // =========================== conv1a epilogue: ReLU + bf16 into conv1b's A operand (warps 18..25) =========
    const int quad = warp & 3;                           // TMEM lane quadrant of this warp
    const int chalf = (warp - 2 - kC1EpiWarps) >> 2;     // which 32 of the 64 conv1a channels this warp converts
    const uint32_t taddr = tmem_base + 384 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += G, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      // the two barriers are polled by two different lanes at once (a satisfied wait still costs ~200 cycles)
      C1_PROBE(it, 6);
      if (lane == 0) mbar_wait(acc1_full, it & 1);
      if (lane == 1) mbar_wait(a2_empty(buf), ph ^ 1);
      __syncwarp();
      C1_PROBE(it, 7);
      tc_fence_after();
      // bias and the out-of-image zeroing are already in the accumulator (see build): ReLU + bf16 only.
      // Pixels quad*32 + lane first; the quadrant-0 warps then also convert pixels 128 + lane (TMEM lanes 0..31 of the
      // second MMA) and release the accumulator after that second read.
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 0 || quad == 0) {
          uint32_t v[32];
          tc_ld_nowait<32>(taddr + h * 64, v);
          tc_wait_ld();
          if (h == 1 || quad != 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc1_empty(buf));
          }
          const int p = h * 128 + quad * 32 + lane;
          uint8_t* srow = sm.a2[buf] + p * 128;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[e] = relu_pack_bf16(v[j + 2 * e], v[j + 2 * e + 1]);
            const int chunk = (chalf * 4 + j / 8) ^ (p & 7);  // SWIZZLE_128B
            *reinterpret_cast<uint4*>(srow + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a2_full(buf));
      C1_PROBE(it, 8);
    }
  } else {
    // =========================== conv1b epilogue + im2col build (warps 2..17): two groups of 8 warps ===========
    // Group g owns the tiles of parity g: accumulator stage g, im2col buffer g, staging buffers [g][.], named
    // barrier 1+g.  Per tile i it first builds the im2col tile of tile i+2 (the zero-padded fp32 patch under that
    // tile's halo block arrives by TMA, issued by the group's first thread one iteration earlier), then converts
    // conv1b(i).
    const int grp = (warp - 2) >> 3, quad = warp & 3, chalf = ((warp - 2) >> 2) & 1;
    const int gtid = threadIdx.x - 64 - grp * 256;          // 0..255 inside the group; < 160: halo pixel to build
    const int ry = 2 * quad + (lane >> 4), ox = lane & 15;  // tile row / window column of this thread's TMEM lane
    const bool store_thread = gtid == 0;
    const bool owner = ox < kC1Cols && !(lane & 17);        // even column, even row of the 2x2 window
    const int p = (ry >> 1) * (kC1Cols / 2) + (ox >> 1);    // pooled pixel of the tile (4 rows x 7 columns)
    auto coords = [&](int t, int& n, int& y0, int& x0) {
      n = t / tiles_per_view;
      const int rem = t - n * tiles_per_view, ty = rem / a.tiles_x;
      y0 = ty * kC1Rows;
      x0 = (rem - ty * a.tiles_x) * kC1Cols;
    };
    auto issue_patch = [&](int it, int t) {  // (one thread) patch of this CTA's it-th tile = global tile t
      int n, y0, x0;
      coords(t, n, y0, x0);
      mbar_expect_tx(patch_full(it & 3), 960);
      // 16-byte aligned start column: (x0 - 2) rounded down to a multiple of 4 (two's complement: -2 -> -4)
      tma_load_3d(smem_u32(sm.patch[it & 3]), &tm_x, patch_full(it & 3), (x0 - 2) & ~3, y0 - 2, n);
    };
    auto build = [&](int it, int y0, int x0) {  // patch -> K=16 im2col rows of the 160 halo pixels of tile it
      if (gtid < kC1Halo) {
        const int hy = gtid >> 4, hx = gtid & 15;
        const float* pp = sm.patch[it & 3] + hy * 20 + hx + ((x0 - 2) & 3);  // tap (0,0) of halo pixel gtid
        // conv1b zero-pads conv1a's OUTPUT: a halo pixel outside the image must come out as exactly 0, so its
        // whole im2col row is zeroed (its taps can reach inside the image), INCLUDING the two bias slots
        const int iy = y0 - 1 + hy, ix = x0 - 1 + hx;
        const bool inside = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        uint32_t pk[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int k0 = 2 * k, k1 = 2 * k + 1;
          const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[(k0 / 3) * 20 + k0 % 3], pp[(k1 / 3) * 20 + k1 % 3]);
          pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[2 * 20 + 2], 1.f);  // slot 9: bias head
        pk[4] = *reinterpret_cast<const uint32_t*>(&hh);
        if (!inside) pk[0] = pk[1] = pk[2] = pk[3] = pk[4] = 0u;
        const int h = gtid >> 7, rr = gtid & 127;
        uint8_t* row = sm.a1[grp] + h * 4096 + (rr >> 3) * 256 + (rr & 7) * 16;
        *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(row + 128) = make_uint4(pk[4], inside ? 0x00003F80u : 0u, 0, 0);  // slot 10: bias tail
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a1_full(grp));
    };
    int t = blockIdx.x + grp * G;
    int n = 0, y0 = 0, x0 = 0;
    if (t < a.ntiles) {
      coords(t, n, y0, x0);
      if (store_thread) {
        issue_patch(grp, t);
        if (t + 2 * G < a.ntiles) issue_patch(grp + 2, t + 2 * G);
      }
      if (lane == 0) mbar_wait(patch_full(grp), 0);
      __syncwarp();
      build(grp, y0, x0);
      named_bar_sync(1 + grp, 256);  // (patch buffer grp is re-filled at the top of the first iteration)
    }
    int k = 0;  // this group's tile counter
    for (; t < a.ntiles; t += 2 * G, ++k) {
      const int it = 2 * k + grp;
      int n2 = 0, y2 = 0, x2 = 0;
      C1_PROBE(it, 9);
      // patch of tile it+4 into the buffer of tile it, a whole iteration before it is needed (an L2/HBM round trip
      // is ~1000 cycles): every thread of the group finished build(it) before the group's last named barrier
      if (store_thread && t + 4 * G < a.ntiles) issue_patch(it + 4, t + 4 * G);
      if (t + 2 * G < a.ntiles) {
        coords(t + 2 * G, n2, y2, x2);
        if (lane == 0) mbar_wait(a1_empty(grp), k & 1);                               // conv1a(it) has read a1[grp]
        if (lane == 1) mbar_wait(patch_full((it + 2) & 3), ((it + 2) >> 2) & 1);
        __syncwarp();
        build(it + 2, y2, x2);
      }
      uint8_t* stg = sm.ostage[grp][k & 1];
      uint8_t* srow = stg + p * 128;
      C1_PROBE(it, 10);
      mbar_wait(acc2_full(grp), k & 1);
      C1_PROBE(it, 11);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)grp * 192 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {  // four batches of 8 channels (24 live accumulator registers at a time)
        uint32_t c[3][8];
#pragma unroll
        for (int s = 0; s < 3; ++s) tc_ld_nowait<8>(taddr + s * 64 + cb * 8, c[s]);
        tc_wait_ld();
        if (cb == 3) {  // everything this warp needs has left the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty(grp));
        }
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(sm.bias2 + chalf * 32 + cb * 8 + j);
          const float bj[4] = {bb.x, bb.y, bb.z, bb.w};
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float c1 = __shfl_down_sync(0xffffffffu, __uint_as_float(c[1][j + e]), 1);
            const float c2 = __shfl_down_sync(0xffffffffu, __uint_as_float(c[2][j + e]), 2);
            f[e] = ((__uint_as_float(c[0][j + e]) + c1) + c2) + bj[e];
          }
          s3_pack4<true>(f, pk + j / 2);
        }
        if (owner)
          *reinterpret_cast<uint4*>(srow + (((chalf * 4 + cb) ^ (p & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the group's previous store (its other staging buffer) must have drained before anyone passes this barrier
      // and starts writing that buffer for the group's next tile
      if (store_thread) tma_store_wait_read<0>();
      named_bar_sync(1 + grp, 256);
      if (store_thread) {
        tma_store_4d(&tm_out, smem_u32(stg), 0, x0 >> 1, y0 >> 1, n);  // clipped at the image border
        tma_store_commit();
      }
      C1_PROBE(it, 12);
      n = n2;
      y0 = y2;
      x0 = x2;
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
int conv1_s3(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
             cudaStream_t st) {
  if (!L1b.tmap_w3 || ((H | W) & 1) || (W & 3)) {
    set_error("conv1_s3: needs prepared N=192 weights, even H and W %% 4 == 0 (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  CUtensorMap tm_w2, tm_out, tm_x;
  memcpy(&tm_w2, L1b.tmap_w3, sizeof(tm_w2));
  C1S3Args a;
  a.w1a = L1a.w_tc;
  a.bias2 = L1b.bias;
  a.V = V;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, kC1Cols);
  a.tiles_y = ceil_div(H, kC1Rows);
  a.ntiles = V * a.tiles_x * a.tiles_y;
  a.dbg = conv1_probe_buf();
  const int rc = encode_act_map(&tm_out, out, V, H / 2, W / 2, 64, kC1Cols / 2, kC1Rows / 2);
  if (rc != SPB_OK) return rc;
  {
    // fp32 views [V,H,W]: box = the 12 x 20 patch under a halo block, zero filled outside the image
    EncodeTiledFn enc = get_encode();
    if (!enc) {
      set_error("conv1_s3: cuTensorMapEncodeTiled not available from the driver");
      return SPB_ECUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)V};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {20, 12, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv1_s3: cuTensorMapEncodeTiled(views %dx%dx%d) failed with CUresult %d", V, H, W, (int)r);
      return SPB_ECUDA;
    }
  }
  const int smem = (int)sizeof(C1S3Smem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_s3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv1_s3_kernel<<<grid, kC1Threads, smem, st>>>(tm_w2, tm_out, tm_x, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb
