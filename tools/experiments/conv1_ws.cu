// EXPERIMENT, NOT PART OF libspb200.so (round 2; measured slower than conv1_fused_kernel and removed from the build).
// Correct: bit-identical to conv1_fused_kernel on every shape tried (odd tile counts, fewer tiles than CTAs).
// Measured on 800 views 240x320, whole detector forward (tools/ab_fuse1.py): 9.2-9.4 ms against 7.65 ms.
// Why (profiles/r02_conv1_ws_probe_*.txt): sharing a weight block between two tiles forces the MMA stream to switch
// accumulators -- with every MMA (fill / lastuse alternating: ~75 cycles per MMA inside the kernel, 90 with plain
// MMAs in the same order) or every four MMAs (four collector buffers: better, still +13 %).  In isolation the same
// sequences run at 45 / 48 cycles (tools/ubench_mma_ws.cu): the cost appears only next to the TMEM reads of stage 1
// and the epilogue, i.e. every accumulator switch is a TMEM round trip of the 32 KB tile that competes with them.
// To build it again: add it to SOURCES in build.py, declare conv1_ws in net.cuh and dispatch on g_fuse1 in conv_tc.cu.
//
// (2a'') conv1a + conv1b + 2x2 max-pool in one kernel, conv1b issued as WEIGHT-STATIONARY tile pairs -- the first VGG
// block of models/SuperPointNet_pretrained.py:55-58, SuperPointNet_gauss2.py:50-52 / unet_parts.py:14-21.
//
// conv1_fused_kernel (conv_tc.cu) sits on the shared-memory bandwidth of the SM: 273 KB of shared-memory traffic per
// 128-pixel tile at 128 B/clk = 2130 cycles against 2116 measured, 216 KB of it the operand reads of the 36
// M128 x N64 x K16 MMAs (4 KB of A + 2 KB of B each).  `tcgen05.mma.ws` keeps the B operand in a collector buffer of
// the tensor core (`.collector::b0::fill` / `::lastuse`): a B-resident MMA reads only its A operand (40-42 cycles
// instead of 48, bit-identical accumulators -- tools/ubench_mma_ws.cu, profiles/r02_ubench_mma_ws.txt).  Here the
// MMAs of TWO tiles are interleaved per (tap, K step), so every weight block is read once per pair:
//     fill:    acc2[j]   += A(tile j,   tap, k) . W(tap, k)      reads 4 KB + 2 KB
//     lastuse: acc2[j+1] += A(tile j+1, tap, k) . W(tap, k)      reads 4 KB
// Everything else is the shifted-window block: per 16x8 output tile the fp32 patch under the 18x10 halo block arrives
// by TMA, 8 stage-1 warps build the K=16 im2col tile (bias folded in), two M128 N64 K16 MMAs produce conv1a for the
// 180 halo pixels in TMEM, the same warps apply ReLU + bf16 and write conv1b's SWIZZLE_128B A operand, 8 epilogue
// warps pool in registers, stage and TMA-store.  Differences forced by the pairing:
//   * FOUR conv1b A buffers and FOUR conv1b accumulators (a pair in flight while stage 1 fills the next pair):
//     TMEM = 4 x 64 + 2 x 128 (conv1a) = 512 columns, shared memory 201 KB;
//   * conv1a has its OWN issuer warp (blocking waits for the im2col tile and the drained accumulator): the conv1b
//     issuers never poll for it.  A conv1a MMA that lands between a fill and its lastuse only costs that pair its
//     re-use (the collector refetches B; accumulators stay exact -- tools/ubench_mma_ws.cu, variant W6);
//   * two conv1b issuer warps alternate PAIRS with a token handed over before the last taps (one is through its
//     ~200-cycle barrier waits when the other runs out of MMAs), each with its own collector buffer (b0 / b1).
// Warps: 0/1 conv1b issuers (0 loads the weights, 1 allocates TMEM), 2..9 conv1b epilogue, 10..17 stage 1,
// 18 conv1a issuer.
#include "tc_common.cuh"

namespace spb {

struct Conv1WsSmem {
  uint8_t b2[9][8192];         // conv1b weights, 9 taps x [64 rows x 128 B], SWIZZLE_128B (TMA)
  uint8_t a2[4][23552];        // conv1b A operand: 18x10 halo pixels x 128 B, SWIZZLE_128B, written by stage 1
  uint8_t ostage[2][4][1024];  // pooled output of one TMEM quadrant: 2 x 4 pixels x 128 B, SWIZZLE_128B (TMA store)
  uint8_t a1[2][8192];         // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices
  uint8_t b1[2048];            // conv1a weights [64][16]
  float patch[4][320];         // zero-padded 20x16 fp32 image patch under the halo block, TMA
  int2 origin[4];
  float bias2[64];
  uint64_t bars[9 + 2 + 2 + 2 + 4 + 4 + 4 + 4 + 4 + 2];
  uint32_t tmem_slot;
};

struct Conv1WsArgs {
  const __nv_bfloat16* w1a;  // [64][16]
  const float* bias2;
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;  // optional timeline probe (CTA 0): [64 tiles][16 slots] of clock64, or NULL
  int plain;       // A/B: issue the pairs with plain MMAs (no collector re-use), everything else unchanged
};
// slots: 0 issuer: pair loop top, 1 operands ready, 2 token, 3 issued | 4 conv1a issuer: ready, 5 issued |
// 6 stage 1: iteration top, 7 built next, 8 waits done, 9 a2 written | 10 epilogue: top, 11 acc ready, 12 stored
#define WS_PROBE(j, slot, cond) \
  if (a.dbg && blockIdx.x == 0 && (j) < 64 && (cond)) a.dbg[(j) * 16 + (slot)] = clock64()

// collector buffer BUF = 0..3.  Per tap the four K steps of the first tile fill b0..b3, the four K steps of the second
// tile use them: every accumulator sees runs of four consecutive MMAs (alternating the accumulator with every MMA
// was measured at ~90 cycles per MMA inside the kernel), every weight block is still read once per pair.
#define SPB_WS_MMA(BUF, OP)                                                                                        \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"                                                        \
               "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b" #BUF "::" #OP " [%0], %1, %2, %3, p;\n}" ::"r"(d), \
               "l"(ad), "l"(bd), "r"(idesc), "r"(acc)                                                              \
               : "memory")
template <int BUF>
__device__ __forceinline__ void tc_mma_ws_fill(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (BUF == 0) SPB_WS_MMA(0, fill);
  else if (BUF == 1) SPB_WS_MMA(1, fill);
  else if (BUF == 2) SPB_WS_MMA(2, fill);
  else SPB_WS_MMA(3, fill);
}
template <int BUF>
__device__ __forceinline__ void tc_mma_ws_last(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (BUF == 0) SPB_WS_MMA(0, lastuse);
  else if (BUF == 1) SPB_WS_MMA(1, lastuse);
  else if (BUF == 2) SPB_WS_MMA(2, lastuse);
  else SPB_WS_MMA(3, lastuse);
}

__global__ void __launch_bounds__(608, 1) conv1_ws_kernel(const __grid_constant__ CUtensorMap tm_w2,
                                                          const __grid_constant__ CUtensorMap tm_out,
                                                          const __grid_constant__ CUtensorMap tm_x, Conv1WsArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Conv1WsSmem& sm = *reinterpret_cast<Conv1WsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(sm.bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a1_full = [&](int s) { return bar0 + 8u * (9 + s); };
  auto acc1_full = [&](int s) { return bar0 + 8u * (11 + s); };
  auto acc1_empty = [&](int s) { return bar0 + 8u * (13 + s); };
  auto a2_full = [&](int s) { return bar0 + 8u * (15 + s); };
  auto a2_empty = [&](int s) { return bar0 + 8u * (19 + s); };
  auto acc2_full = [&](int s) { return bar0 + 8u * (23 + s); };
  auto acc2_empty = [&](int s) { return bar0 + 8u * (27 + s); };
  auto patch_full = [&](int s) { return bar0 + 8u * (31 + s); };
  auto token = [&](int s) { return bar0 + 8u * (35 + s); };

  if (threadIdx.x < 64) {
    sm.bias2[threadIdx.x] = a.bias2[threadIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(a.w1a + threadIdx.x * 16);
    const int n = threadIdx.x;
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + (n & 7) * 16) = src[0];
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + 128 + (n & 7) * 16) = src[1];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 9; ++s) mbar_init(b_full(s), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a1_full(s), 256);
      mbar_init(acc1_full(s), 1);
      mbar_init(acc1_empty(s), 8);
      mbar_init(token(s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(a2_full(s), 256);
      mbar_init(a2_empty(s), 1);
      mbar_init(acc2_full(s), 1);
      mbar_init(acc2_empty(s), 8);
      mbar_init(patch_full(s), 1);
    }
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
  // TMEM columns: conv1b accumulator of tile j at 64 (j & 3); conv1a accumulator of tile j at 256 + 128 (j & 1)
  const int tiles_per_view = a.tiles_x * a.tiles_y;
  // this CTA's tiles: t_j = blockIdx.x + j gridDim.x, j = 0 .. nt - 1
  const int nt = (int)blockIdx.x < a.ntiles ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp <= 1) {
    // =========================== conv1b issuers: pairs of tiles, alternating ===========================
    // Issuer q owns the pairs of parity q: tiles 4m + 2q, 4m + 2q + 1.  The token is handed over before the last two
    // taps, so the other issuer -- already through its barrier waits -- queues its MMAs right behind.
    const int q = warp;
    if (q == 0 && lane == 0) {
      for (int tap = 0; tap < 9; ++tap) {
        mbar_expect_tx(b_full(tap), 8192);
        tma_load_2d(smem_u32(sm.b2[tap]), &tm_w2, b_full(tap), tap * 64, 0);
      }
    }
    __syncwarp();
    constexpr uint32_t idesc = make_idesc(128, 64);
    constexpr uint64_t hi_a2 = ((uint64_t)(1280 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    constexpr uint64_t hi_b2 = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    int k = 0;  // this issuer's pair counter
    for (int j = 2 * q; j < nt; j += 4, ++k) {
      const bool two = j + 1 < nt;
      const int s0 = j & 3, s1 = s0 + 1;
      const uint32_t ph = (j >> 2) & 1;  // (both tiles of a pair share buffers' phase: s0 is even)
      WS_PROBE(j, 0, lane == 0);
      // the (up to) four operand barriers of the pair are polled by four lanes at once
      if (lane == 0) mbar_wait(acc2_empty(s0), ph ^ 1);
      if (lane == 1) mbar_wait(a2_full(s0), ph);
      if (two && lane == 2) mbar_wait(acc2_empty(s1), ph ^ 1);
      if (two && lane == 3) mbar_wait(a2_full(s1), ph);
      __syncwarp();
      WS_PROBE(j, 1, lane == 0);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      WS_PROBE(j, 2, lane == 0);
      tc_fence_after();
      const uint32_t d0 = tmem_base + 64 * s0, d1 = tmem_base + 64 * s1;
      const uint32_t a20 = smem_u32(sm.a2[s0]), a21 = smem_u32(sm.a2[s1 & 3]);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        if (k == 0) mbar_wait(b_full(tap), 0);
        const uint32_t tap_off = ((tap / 3) * 10 + (tap % 3)) * 128;
        const uint64_t ad0 = hi_a2 | (uint64_t)(((a20 + tap_off) >> 4) | (1u << 16));
        const uint64_t ad1 = hi_a2 | (uint64_t)(((a21 + tap_off) >> 4) | (1u << 16));
        const uint64_t bdesc = hi_b2 | (uint64_t)((smem_u32(sm.b2[tap]) >> 4) | (1u << 16));
        if (elect_one()) {
          const uint32_t acc0 = tap != 0 ? 1u : 0u;
          if (!two || a.plain) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(d0, ad0 + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk ? 1u : acc0);
            if (two) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma(d1, ad1 + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk ? 1u : acc0);
            }
          } else {
            tc_mma_ws_fill<0>(d0, ad0, bdesc, idesc, acc0);
            tc_mma_ws_fill<1>(d0, ad0 + 2, bdesc + 2, idesc, 1u);
            tc_mma_ws_fill<2>(d0, ad0 + 4, bdesc + 4, idesc, 1u);
            tc_mma_ws_fill<3>(d0, ad0 + 6, bdesc + 6, idesc, 1u);
            tc_mma_ws_last<0>(d1, ad1, bdesc, idesc, acc0);
            tc_mma_ws_last<1>(d1, ad1 + 2, bdesc + 2, idesc, 1u);
            tc_mma_ws_last<2>(d1, ad1 + 4, bdesc + 4, idesc, 1u);
            tc_mma_ws_last<3>(d1, ad1 + 6, bdesc + 6, idesc, 1u);
          }
          if (tap == 8) mbar_arrive(token(q ^ 1));  // (the issuers share the four collector buffers: no interleaving)
          if (tap == 8) {
            tc_commit(a2_empty(s0));
            tc_commit(acc2_full(s0));
            if (two) {
              tc_commit(a2_empty(s1));
              tc_commit(acc2_full(s1));
            }
          }
        }
        __syncwarp();
      }
      WS_PROBE(j, 3, lane == 0);
    }
  } else if (warp == 18) {
    // =========================== conv1a issuer ===========================
    constexpr uint32_t idesc = make_idesc(128, 64);
    const uint64_t b1desc = make_desc_noswz(smem_u32(sm.b1), 128, 256);
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1, ph = (j >> 1) & 1;
      if (lane == 0) mbar_wait(a1_full(s), ph);          // stage 1 built the im2col tile of tile j
      if (lane == 1) mbar_wait(acc1_empty(s), ph ^ 1);   // ... and drained the accumulator of tile j - 2
      __syncwarp();
      WS_PROBE(j, 4, lane == 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a1desc = make_desc_noswz(smem_u32(sm.a1[s]), 128, 256);
        const uint32_t acc1 = tmem_base + 256 + s * 128;
        tc_mma(acc1, a1desc, b1desc, idesc, 0u);
        tc_mma(acc1 + 64, a1desc + (uint64_t)(4096 >> 4), b1desc, idesc, 0u);
        tc_commit(acc1_full(s));
      }
      __syncwarp();
      WS_PROBE(j, 5, lane == 0);
    }
  } else if (warp >= 10 && warp <= 17) {
    // =========================== stage 1: im2col build + conv1a epilogue into conv1b's A operand =========
    const int sid = threadIdx.x - 320;           // 0..255
    const int quad = warp & 3;                   // TMEM lane quadrant of this warp
    const int chalf = (warp - 10) >> 2;          // which 32 of the 64 conv1a channels this warp converts
    const int lrow = quad * 32 + lane;           // TMEM lane; the thread owns halo pixels lrow and lrow + 128
    const int tpv = tiles_per_view;
    const int dgn = (int)gridDim.x / tpv, dgr = (int)gridDim.x % tpv;
    const int dgy = dgr / a.tiles_x, dgx = dgr % a.tiles_x;
    int pn = (int)blockIdx.x / tpv, pty = ((int)blockIdx.x % tpv) / a.tiles_x, ptx = ((int)blockIdx.x % tpv) % a.tiles_x;
    auto issue_patch = [&](int j) {  // patch of this CTA's j-th tile = the tile the walker stands on; then advance
      if (warp == 10) {
        if (elect_one()) {
          sm.origin[j & 3] = make_int2(ptx * 8, pty * 16);
          mbar_expect_tx(patch_full(j & 3), 1280);
          tma_load_3d(smem_u32(sm.patch[j & 3]), &tm_x, patch_full(j & 3), ptx * 8 - 4, pty * 16 - 2, pn);
        }
        __syncwarp();
        ptx += dgx;
        if (ptx >= a.tiles_x) { ptx -= a.tiles_x; ++pty; }
        pty += dgy;
        if (pty >= a.tiles_y) { pty -= a.tiles_y; ++pn; }
        pn += dgn;
      }
    };
    auto build = [&](int j) {  // patch -> K=16 im2col rows of the 180 halo pixels of this CTA's j-th tile
      const int buf = j & 1;   // (the caller has waited for patch_full(j & 3))
      if (sid < 180) {
        const int2 org = sm.origin[j & 3];
        const float* pp = sm.patch[j & 3] + (sid / 10) * 16 + sid % 10 + 2;  // tap (0,0) of halo pixel sid
        // conv1b zero-pads conv1a's OUTPUT: a halo pixel outside the image must come out as exactly 0, so its
        // whole im2col row is zeroed (its taps can reach inside the image), INCLUDING the two bias slots
        const int hy = org.y - 1 + sid / 10, hx = org.x - 1 + sid % 10;
        const bool inside = hy >= 0 && hy < a.H && hx >= 0 && hx < a.W;
        uint32_t pk[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int k0 = 2 * k, k1 = 2 * k + 1;
          const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[(k0 / 3) * 16 + k0 % 3], pp[(k1 / 3) * 16 + k1 % 3]);
          pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[2 * 16 + 2], 1.f);  // slot 9: bias head
        pk[4] = *reinterpret_cast<const uint32_t*>(&hh);
        if (!inside) pk[0] = pk[1] = pk[2] = pk[3] = pk[4] = 0u;
        const int h = sid >> 7, rr = sid & 127;
        uint8_t* row = sm.a1[buf] + h * 4096 + (rr >> 3) * 256 + (rr & 7) * 16;
        *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(row + 128) = make_uint4(pk[4], inside ? 0x00003F80u : 0u, 0, 0);  // slot 10: bias tail
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a1_full(buf));
    };
    if (nt > 0) {
      issue_patch(0);
      if (nt > 1) issue_patch(1);
      if (nt > 2) issue_patch(2);
      mbar_wait(patch_full(0), 0);
      build(0);
    }
    for (int it = 0; it < nt; ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;     // conv1a accumulator / im2col buffer of tile it
      const int b2 = it & 3, ph2 = (it >> 2) & 1;     // conv1b A buffer of tile it
      WS_PROBE(it, 6, threadIdx.x == 320);
      if (it + 1 < nt) {
        // im2col of tile it+1 FIRST: conv1a(it+1) then has a whole iteration to get through the MMA queue.  Its
        // buffer a1[(it+1)&1] was read by conv1a(it-1), which completed before acc1_full(it-1) that iteration it-1
        // of this loop waited for.
        mbar_wait(patch_full((it + 1) & 3), ((it + 1) >> 2) & 1);
        build(it + 1);
        if (it + 3 < nt) issue_patch(it + 3);   // patch of tile it+3
      }
      WS_PROBE(it, 7, threadIdx.x == 320);
      if (lane == 0) mbar_wait(acc1_full(buf), ph);
      if (lane == 1) mbar_wait(a2_empty(b2), ph2 ^ 1);
      __syncwarp();
      WS_PROBE(it, 8, threadIdx.x == 320);
      tc_fence_after();
      uint32_t v[2][32];
      const uint32_t taddr = tmem_base + 256 + buf * 128 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      tc_ld_nowait<32>(taddr, v[0]);        // halo pixels lrow        (M half 0)
      tc_ld_nowait<32>(taddr + 64, v[1]);   // halo pixels lrow + 128  (M half 1)
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc1_empty(buf));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = h * 128 + lrow;
        if (p < 180) {  // bias and the out-of-image zeroing are already in the accumulator (see build)
          uint8_t* srow = sm.a2[b2] + p * 128;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pk[q] = relu_pack_bf16(v[h][j + 2 * q], v[h][j + 2 * q + 1]);
            const int chunk = (chalf * 4 + j / 8) ^ (p & 7);  // SWIZZLE_128B
            *reinterpret_cast<uint4*>(srow + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a2_full(b2));
      WS_PROBE(it, 9, threadIdx.x == 320);
    }
  } else if (warp >= 2 && warp <= 9) {
    // =========================== conv1b epilogue (warps 2..9): bias + ReLU + 2x2 max-pool + bf16 ===========
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    const bool owner = !(lane & 9);                          // even column, even row of the 2x2 window
    const int p = (lane >> 4) * 4 + ((lane & 7) >> 1);       // pooled pixel of the quadrant (2 rows x 4 columns)
    const bool issuer = chalf == 0 && lane == 0;
    float bz[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bz[j] = sm.bias2[chalf * 32 + j];
    int t = blockIdx.x;
    for (int it = 0; it < nt; ++it, t += gridDim.x) {
      const int buf = it & 1, b2 = it & 3, ph2 = (it >> 2) & 1;
      const int n = t / tiles_per_view, rem = t % tiles_per_view;
      const int y0 = (rem / a.tiles_x) * 16 + 4 * quad, x0 = (rem % a.tiles_x) * 8;
      WS_PROBE(it, 10, threadIdx.x == 64);
      mbar_wait(acc2_full(b2), ph2);
      WS_PROBE(it, 11, threadIdx.x == 64);
      tc_fence_after();
      const uint32_t taddr = tmem_base + b2 * 64 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      uint32_t v[32];
      tc_ld_nowait<32>(taddr, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty(b2));
      uint32_t pk[16];
      epi_pack32_regs<true>(v, bz, pk);
      uint8_t* stg = sm.ostage[buf][quad];
      if (owner) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(stg + p * 128 + (((chalf * 4 + q4) ^ (p & 7)) * 16)) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the previous tile's store (other staging buffer) must have drained before the pair passes this barrier
      // and goes on to overwrite that buffer
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(3 + quad, 64);
      if (issuer) {
        tma_store_4d(&tm_out, smem_u32(stg), 0, x0 >> 1, y0 >> 1, n);  // clipped at the image border
        tma_store_commit();
      }
      WS_PROBE(it, 12, threadIdx.x == 64);
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

int conv1_ws(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
             cudaStream_t st, int plain) {
  CUtensorMap tm_w2;
  memcpy(&tm_w2, L1b.tmap_w, sizeof(tm_w2));
  Conv1WsArgs a;
  a.w1a = L1a.w_tc;
  a.bias2 = L1b.bias;
  a.V = V;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, 8);
  a.tiles_y = ceil_div(H, 16);
  a.ntiles = V * a.tiles_x * a.tiles_y;
  a.dbg = conv1_probe_buf();
  a.plain = plain;
  CUtensorMap tm_out, tm_x;
  {
    const int rc = encode_act_map(&tm_out, out, V, H / 2, W / 2, 64, 4, 2);
    if (rc != SPB_OK) return rc;
    // fp32 views [V,H,W]: box = the 16 x 20 patch under a halo block, zero filled outside the image
    EncodeTiledFn enc = get_encode();
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)V};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {16, 20, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv1_ws: cuTensorMapEncodeTiled(views %dx%dx%d) failed with CUresult %d", V, H, W, (int)r);
      return SPB_ECUDA;
    }
  }
  const int smem = (int)sizeof(Conv1WsSmem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv1_ws_kernel<<<grid, 608, smem, st>>>(tm_w2, tm_out, tm_x, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb
