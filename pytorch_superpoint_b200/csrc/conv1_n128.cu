// (2a') conv1a + conv1b + 2x2 max-pool in one kernel, conv1b with TWO of the three column taps stacked along the UMMA
// N dimension -- the first VGG block of models/SuperPointNet_pretrained.py:55-58, SuperPointNet_gauss2.py:50-52 /
// unet_parts.py:14-21.  Selected with spb_conv_tc_set_fuse1(3).
//
// The middle ground between conv1_fused_kernel (conv_tc.cu: 36 N = 64 MMAs per tile, bound by the shared-memory
// operand reads at 48 cycles per MMA) and conv1_s3_kernel (conv1_s3.cu: 12 N = 192 MMAs, full tensor rate, but two
// re-alignment shuffles per channel, one conv1a accumulator and 832 threads -- issue bound).  Per tap row r and
// 16-channel K step:
//     MMA_a  M128 x N128 x K16   A = halo block shifted by r rows,            B = [W(r, s=0); W(r, s=1)]   64 cycles
//     MMA_b  M128 x N64  x K16   A = the same block shifted by 2 more PIXELS, B = W(r, s=2)                48 cycles
// MMA_b accumulates into the columns of the s = 0 half, so that
//     out(ry, ox) = D[(ry, ox), 0..63] + D[(ry, ox + 1), 64..127]           ox = 0..13 of the 16-pixel window:
// ONE re-alignment shuffle per channel, 1344 MMA cycles per 112-pixel tile (12.0 per pixel against 16.1 at best for
// the shifted-window block), TMEM 2 x 128 columns for conv1b + 2 x 128 for conv1a (double buffered like
// conv1_fused_kernel).  Warps: 0/1 MMA issuers alternating tiles, 2..9 conv1b epilogue, 10..17 conv1a epilogue,
// 18..22 im2col build (one thread per halo pixel) -- build and conversion are separate warps because together they
// were a ~2000-cycle serial chain per tile that the MMA issuers waited for (profiles/r01_conv1_n128_probe.txt).
// Tile geometry, patch TMA and im2col rows as in conv1_s3.cu (8 x 14 outputs, 10 x 16 halo pixels, row p = hy*16+hx).
// The s = 2 view of the last tile row reads two rows past the halo block (the next buffer): they only feed the
// window columns 14, 15, which are never stored.
#include "tc_common.cuh"

namespace spb {

constexpr int kN1Rows = 8, kN1Cols = 14, kN1Win = 16, kN1Halo = (kN1Rows + 2) * kN1Win;  // 160 halo pixels
constexpr int kN1Threads = 576 + kN1Halo;  // 2 issuer + 8 conv1b-epilogue + 8 conv1a-epilogue + 5 im2col warps

struct C1N128Smem {
  uint8_t b[3][24576];         // conv1b weights per tap row: 192 rows (s, co) x 128 B, SWIZZLE_128B (TMA)
  uint8_t a2[2][20480];        // conv1b A operand: 10 x 16 halo pixels x 128 B, SWIZZLE_128B, written by stage 1
  uint8_t ostage[2][4][1024];  // [buffer][TMEM quadrant]: one pooled row of 7 pixels x 128 B, SWIZZLE_128B (TMA store
                               // source); also what the s = 2 view reads past the end of a2[1]
  uint8_t a1[2][8192];         // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices (rows >= 160 stay zero)
  uint8_t b1[2048];            // conv1a weights [64][16]
  float patch[4][256];         // zero-padded 12 x 20 fp32 patch (rows y0-2.., cols ((x0-2) & ~3)..), TMA
  int4 origin[4];              // (x0, y0, column offset of x0-2 inside the patch, -) of the patch's tile
  float bias2[64];
  uint64_t bars[23];  // b_full[3] a1_full[2] acc1_full[2] acc1_empty[2] a2_full[2] a2_empty[2] acc2_full[2] acc2_empty[2] token[2] patch_full[4]
  uint32_t tmem_slot;
};

struct C1N128Args {
  const __nv_bfloat16* w1a;  // [64][16] (taps 0..8, bias head / tail in slots 9, 10)
  const float* bias2;
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;  // optional timeline probe (CTA 0): [64 tiles][16 slots] of clock64, or NULL
};

__global__ void __launch_bounds__(kN1Threads, 1) conv1_n128_kernel(const __grid_constant__ CUtensorMap tm_w2,
                                                                   const __grid_constant__ CUtensorMap tm_out,
                                                                   const __grid_constant__ CUtensorMap tm_x,
                                                                   C1N128Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  C1N128Smem& sm = *reinterpret_cast<C1N128Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x;
  const uint32_t bar0 = smem_u32(sm.bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a1_full = [&](int s) { return bar0 + 8u * (3 + s); };
  auto acc1_full = [&](int s) { return bar0 + 8u * (5 + s); };
  auto acc1_empty = [&](int s) { return bar0 + 8u * (7 + s); };
  auto a2_full = [&](int s) { return bar0 + 8u * (9 + s); };
  auto a2_empty = [&](int s) { return bar0 + 8u * (11 + s); };
  auto acc2_full = [&](int s) { return bar0 + 8u * (13 + s); };
  auto acc2_empty = [&](int s) { return bar0 + 8u * (15 + s); };
  auto token = [&](int s) { return bar0 + 8u * (17 + s); };
  auto patch_full = [&](int s) { return bar0 + 8u * (19 + s); };

  if (threadIdx.x < 64) {
    sm.bias2[threadIdx.x] = a.bias2[threadIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(a.w1a + threadIdx.x * 16);
    const int n = threadIdx.x;
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + (n & 7) * 16) = src[0];
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + 128 + (n & 7) * 16) = src[1];
  }
  // a1 rows >= 160 are never written again; ostage is read as A rows 160, 161 by the s = 2 view of a2[1]
  for (int i = threadIdx.x; i < (int)((sizeof(sm.ostage) + sizeof(sm.a1)) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(sm.ostage)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(b_full(s), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a1_full(s), kN1Halo);
      mbar_init(acc1_full(s), 1);
      mbar_init(acc1_empty(s), 8);
      mbar_init(a2_full(s), 256);
      mbar_init(a2_empty(s), 1);
      mbar_init(acc2_full(s), 1);
      mbar_init(acc2_empty(s), 8);
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
  // TMEM columns: conv1b accumulators [0,128) [128,256); conv1a accumulator j at 256 + 128 j (pixels 0..127 | 128..159)
  const int tiles_per_view = a.tiles_x * a.tiles_y;
#define N1_PROBE(tile, slot)                                                                                       \
  if (a.dbg && blockIdx.x == 0 && (tile) < 64 && lane == 0 && (warp <= 2 || warp == 10 || warp == 18))             \
  a.dbg[(tile) * 16 + (slot)] = clock64()

  if (warp <= 1) {
    // =========================== MMA issuers (warp 0 also loads the conv1b weights first) ===========================
    // Issuer q owns the tiles of parity q: conv1a(j) into acc1[q], conv1b(j) into acc2[q].  The MMA queue is FIFO
    // across the two warps, so conv1a(j+2) must be queued BEFORE the other issuer starts conv1b(j+1), or stage 1
    // gets its accumulator a whole conv1b late (measured: 650 idle cycles per tile): the token is handed over only
    // after conv1a(j+2), whose two barriers are waited for before the last tap row of conv1b(j) is issued.
    // MMA order  ... conv1b(j) conv1a(j+2) | conv1b(j+1) conv1a(j+3) | ...
    const int q = warp;
    if (q == 0 && lane == 0) {
      for (int r = 0; r < 3; ++r) {
        mbar_expect_tx(b_full(r), 24576);
        tma_load_2d(smem_u32(sm.b[r]), &tm_w2, b_full(r), 0, r * 192);
      }
    }
    __syncwarp();
    constexpr uint32_t idesc64 = make_idesc(128, 64), idesc128 = make_idesc(128, 128);
    constexpr uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t b1desc = make_desc_noswz(smem_u32(sm.b1), 128, 256);
    const uint64_t a1desc = make_desc_noswz(smem_u32(sm.a1[q]), 128, 256);
    const uint32_t acc1 = tmem_base + 256 + (uint32_t)q * 128, acc2 = tmem_base + (uint32_t)q * 128;
    const uint32_t a2 = smem_u32(sm.a2[q]);
    auto wait_conv1a = [&](int k) {  // this issuer's k-th tile: accumulator read out, im2col tile built
      mbar_wait(acc1_empty(q), (k & 1) ^ 1);
      mbar_wait(a1_full(q), k & 1);
    };
    auto issue_conv1a = [&]() {
      tc_fence_after();
      if (elect_one()) {
        tc_mma(acc1, a1desc, b1desc, idesc64, 0u);
        tc_mma(acc1 + 64, a1desc + (uint64_t)(4096 >> 4), b1desc, idesc64, 0u);
        tc_commit(acc1_full(q));
      }
      __syncwarp();
    };
    // prologue: conv1a(0) by issuer 0, conv1a(1) by issuer 1 (both before any conv1b)
    if ((int)blockIdx.x + q * G < a.ntiles) {
      wait_conv1a(0);
      issue_conv1a();
    }
    int k = 0;
    for (int t = blockIdx.x + q * G; t < a.ntiles; t += 2 * G, ++k) {
      const bool has_next = t + 2 * G < a.ntiles;
      N1_PROBE(2 * k + q, 0);
      mbar_wait(acc2_empty(q), (k & 1) ^ 1);
      N1_PROBE(2 * k + q, 1);
      mbar_wait(a2_full(q), k & 1);
      N1_PROBE(2 * k + q, 2);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      N1_PROBE(2 * k + q, 3);
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (k == 0) {
          mbar_wait(b_full(r), 0);
          tc_fence_after();
        }
        if (r == 2 && has_next) wait_conv1a(k + 1);  // (the queue holds two tap rows of work meanwhile)
        const uint32_t arow = a2 + r * (kN1Win * 128);
        const uint64_t adesc_a = desc_hi | (uint64_t)((arow >> 4) | (1u << 16));
        const uint64_t adesc_b = desc_hi | (uint64_t)(((arow + 2 * 128) >> 4) | (1u << 16));  // two pixels further
        const uint64_t bdesc_a = desc_hi | (uint64_t)((smem_u32(sm.b[r]) >> 4) | (1u << 16));
        const uint64_t bdesc_b = desc_hi | (uint64_t)(((smem_u32(sm.b[r]) + 128 * 128) >> 4) | (1u << 16));  // rows s = 2
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            tc_mma(acc2, adesc_a + (uint64_t)(kk * 2), bdesc_a + (uint64_t)(kk * 2), idesc128, (r | kk) != 0 ? 1u : 0u);
            tc_mma(acc2, adesc_b + (uint64_t)(kk * 2), bdesc_b + (uint64_t)(kk * 2), idesc64, 1u);
          }
          if (r == 2) {
            tc_commit(a2_empty(q));
            tc_commit(acc2_full(q));
          }
        }
        __syncwarp();
      }
      N1_PROBE(2 * k + q, 4);
      if (has_next) issue_conv1a();
      if (lane == 0) mbar_arrive(token(q ^ 1));
      __syncwarp();
      N1_PROBE(2 * k + q, 5);
    }
  } else if (warp >= 18) {
    // =========================== im2col build (warps 18..22, one thread per halo pixel) ===========================
    // Thread bsid turns the 3x3 neighbourhood of halo pixel bsid in the fp32 patch into one K=16 im2col row.  The
    // patch of tile j arrives by TMA (four buffers, issued four tiles ahead by the first build thread once every
    // build thread has finished with the buffer's previous tile).  a1[j & 1] is free once conv1a(j-2) has executed.
    const int bsid = threadIdx.x - 576;  // 0..159
    const int hy = bsid >> 4, hx = bsid & 15;
    const int h = bsid >> 7, rr = bsid & 127;
    uint8_t* const row0 = sm.a1[0] + h * 4096 + (rr >> 3) * 256 + (rr & 7) * 16;
    const int tpv = tiles_per_view;
    const int dgn = G / tpv, dgr = G % tpv;
    const int dgy = dgr / a.tiles_x, dgx = dgr % a.tiles_x;
    int pn = (int)blockIdx.x / tpv, pty = ((int)blockIdx.x % tpv) / a.tiles_x, ptx = ((int)blockIdx.x % tpv) % a.tiles_x;
    auto issue_patch = [&](int j) {  // (thread 576 only) patch of this CTA's j-th tile = where the walker stands
      const int x0 = ptx * kN1Cols, y0 = pty * kN1Rows;
      const int xs = (x0 - 2) & ~3;  // 16-byte aligned start column (two's complement: -2 -> -4)
      sm.origin[j & 3] = make_int4(x0, y0, x0 - 2 - xs, 0);
      mbar_expect_tx(patch_full(j & 3), 960);
      tma_load_3d(smem_u32(sm.patch[j & 3]), &tm_x, patch_full(j & 3), xs, y0 - 2, pn);
      ptx += dgx;
      if (ptx >= a.tiles_x) { ptx -= a.tiles_x; ++pty; }
      pty += dgy;
      if (pty >= a.tiles_y) { pty -= a.tiles_y; ++pn; }
      pn += dgn;
    };
    if (bsid == 0)
      for (int j = 0; j < 4; ++j)
        if ((int)blockIdx.x + j * G < a.ntiles) issue_patch(j);
    int j = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += G, ++j) {
      const int buf = j & 1;
      if (lane == 0) mbar_wait(patch_full(j & 3), (j >> 2) & 1);
      if (lane == 1 && j >= 2) mbar_wait(acc1_full(buf), ((j - 2) >> 1) & 1);  // conv1a(j-2) has read a1[buf]
      __syncwarp();
      const int4 org = sm.origin[j & 3];
      const float* pp = sm.patch[j & 3] + hy * 20 + hx + org.z;  // tap (0,0) of halo pixel bsid
      // conv1b zero-pads conv1a's OUTPUT: a halo pixel outside the image must come out as exactly 0, so its
      // whole im2col row is zeroed (its taps can reach inside the image), INCLUDING the two bias slots
      const int iy = org.y - 1 + hy, ix = org.x - 1 + hx;
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
      uint8_t* row = row0 + buf * 8192;
      *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(row + 128) = make_uint4(pk[4], inside ? 0x00003F80u : 0u, 0, 0);  // slot 10: bias tail
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a1_full(buf));
      N1_PROBE(j, 7);
      // every build thread is done with patch buffer j & 3 -> refill it with the patch of tile j+4
      named_bar_sync(6, kN1Halo);
      if (bsid == 0 && t + 4 * G < a.ntiles) issue_patch(j + 4);
    }
  } else if (warp >= 10) {
    // =========================== conv1a epilogue (warps 10..17): ReLU + bf16 into conv1b's A operand =========
    const int quad = warp & 3;           // TMEM lane quadrant of this warp
    const int chalf = (warp - 10) >> 2;  // which 32 of the 64 conv1a channels this warp converts
    int it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += G, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      N1_PROBE(it, 6);
      // the two barriers are polled by two different lanes at once (a satisfied wait still costs ~200 cycles)
      if (lane == 0) mbar_wait(acc1_full(buf), ph);
      if (lane == 1) mbar_wait(a2_empty(buf), ph ^ 1);
      __syncwarp();
      N1_PROBE(it, 8);
      tc_fence_after();
      // bias and the out-of-image zeroing are already in the accumulator (see the build): ReLU + bf16 only.
      // Pixels quad*32 + lane first; the quadrant-0 warps then also convert pixels 128 + lane (TMEM lanes 0..31 of the
      // second MMA) and release the accumulator after that second read.
      const uint32_t taddr = tmem_base + 256 + buf * 128 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
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
      N1_PROBE(it, 9);
    }
  } else {
    // =========================== conv1b epilogue (warps 2..9): re-align, bias, ReLU, 2x2 max-pool, bf16 ===========
    // Two warps per TMEM quadrant, 32 channels each; window column ox of tile row ry sits in TMEM lane
    // (ry % 2) * 16 + ox of quadrant ry / 2, so the s = 1 partner (lane + 1) and both 2x2 pool partners (lane ^ 1,
    // lane ^ 16) are in the same warp.  A quadrant = two tile rows = ONE pooled row of 7 pixels: its two warps stage
    // 7 x 128 B (SWIZZLE_128B), meet at a 64-thread named barrier and one TMA tensor store writes the row, clipped
    // at the image border.
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    const int ox = lane & 15;
    const bool issuer = chalf == 0 && lane == 0;
    const bool owner = ox < kN1Cols && !(lane & 17);  // even column, even row of the 2x2 window
    const int p = ox >> 1;                            // pooled pixel of the quadrant's row (7 columns)
    const float* bz = sm.bias2 + chalf * 32;
    int it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += G, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      const int n = t / tiles_per_view, rem = t - n * tiles_per_view, ty = rem / a.tiles_x;
      const int y0 = ty * kN1Rows, x0 = (rem - ty * a.tiles_x) * kN1Cols;
      uint8_t* stg = sm.ostage[buf][quad];
      N1_PROBE(it, 10);
      mbar_wait(acc2_full(buf), ph);
      N1_PROBE(it, 11);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 128 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      uint32_t pk[16];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {  // two batches of 16 channels
        uint32_t c0[16], c1[16];
        tc_ld_nowait<16>(taddr + cb * 16, c0);       // s = 0 and s = 2 taps, aligned
        tc_ld_nowait<16>(taddr + 64 + cb * 16, c1);  // s = 1 tap, one window column to the right
        tc_wait_ld();
        if (cb == 1) {  // everything this warp needs has left the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty(buf));
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bz + cb * 16 + j);
          const float bj[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(c1[j + e]), 1);
            f[j + e] = (__uint_as_float(c0[j + e]) + s1) + bj[e];
          }
        }
        s3_pack16<true>(f, pk + cb * 8);
      }
      if (owner) {
        uint8_t* srow = stg + p * 128;
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<uint4*>(srow + (((chalf * 4 + v) ^ (p & 7)) * 16)) =
              make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
      }
      N1_PROBE(it, 12);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the previous tile's store (other staging buffer) must have drained before the pair passes this barrier
      // and goes on to overwrite that buffer
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(1 + quad, 64);
      if (issuer) {
        tma_store_4d(&tm_out, smem_u32(stg), 0, x0 >> 1, (y0 >> 1) + quad, n);  // clipped at the image border
        tma_store_commit();
      }
      N1_PROBE(it, 13);
    }
    if (issuer) tma_store_wait_all();
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
int conv1_n128(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
               cudaStream_t st) {
  if (!L1b.tmap_w3 || ((H | W) & 1) || (W & 3)) {
    set_error("conv1_n128: needs prepared column-stacked weights, even H and W %% 4 == 0 (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  CUtensorMap tm_w2, tm_out, tm_x;
  memcpy(&tm_w2, L1b.tmap_w3, sizeof(tm_w2));
  C1N128Args a;
  a.w1a = L1a.w_tc;
  a.bias2 = L1b.bias;
  a.V = V;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, kN1Cols);
  a.tiles_y = ceil_div(H, kN1Rows);
  a.ntiles = V * a.tiles_x * a.tiles_y;
  a.dbg = conv1_probe_buf();
  const int rc = encode_act_map(&tm_out, out, V, H / 2, W / 2, 64, kN1Cols / 2, 1);
  if (rc != SPB_OK) return rc;
  {
    // fp32 views [V,H,W]: box = the 12 x 20 patch under a halo block, zero filled outside the image
    EncodeTiledFn enc = get_encode();
    if (!enc) {
      set_error("conv1_n128: cuTensorMapEncodeTiled not available from the driver");
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
      set_error("conv1_n128: cuTensorMapEncodeTiled(views %dx%dx%d) failed with CUresult %d", V, H, W, (int)r);
      return SPB_ECUDA;
    }
  }
  const int smem = (int)sizeof(C1N128Smem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_n128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv1_n128_kernel<<<grid, kN1Threads, smem, st>>>(tm_w2, tm_out, tm_x, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb
