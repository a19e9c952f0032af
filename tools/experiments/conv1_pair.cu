// EXPERIMENT, NOT PART OF libspb200.so (round 2; measured slower than conv1_fused_kernel and removed from the build).
// Correct: bit-identical to conv1_fused_kernel on every shape tried (800 x 240x320, 100 x 384x1248, odd tile counts
// 1 x 48x40 and 7 x 80x72, fewer tiles than CTAs) on the first run.
// Measured on 800 views 240x320, whole detector forward (tools/ab_fuse1.py 2 800): 9.9-12.0 ms against 7.6 ms
// (variant 0 release.cluster arrives + acquire.cluster waits 12.0; 1 relaxed remote arrives 9.9; 2 cta-scope waits
// 12.2; 3 both 10.5); 100 x 384x1248: 8.2 against 5.9 ms.
// Why (profiles/r02_conv1_pair_probe.txt): the period per tile goes from 2115 to 3290 cycles.  (1) A remote
// mbarrier arrive blocks the arriving lane for ~400-500 cycles (release.cluster: ~1000): the follower's stage 1
// needs three per tile and its conv1a epilogue takes 1418 cycles instead of 357 -- the follower's stage 1 becomes a
// 3200-cycle serial loop with no waiting in it, the leader's issuers wait ~3000 cycles per tile pair for its a2_full.
// (2) While pair MMAs execute, the OTHER shared-memory work of both SMs crawls: the leader's im2col build (9 LDS +
// 2 STS per thread) takes 1064 cycles instead of 228, its conv1a epilogue 535 instead of 357, although its own
// arrivals are local -- a pair MMA keeps two SMs' operand fetches in lock-step (and ships each SM's half of B to the
// peer), and the LSU traffic this fused block needs next to the MMAs (~50 KB per tile + the TMEM reads) loses more
// than the halved weight reads (36 KB per tile) give back.  The >= 128-channel layers (conv_tc2.cu) have no such
// traffic next to their MMAs, which is why the same technique wins there.
// To build it again: copy it to csrc/, add it to SOURCES in build.py, declare conv1_pair in net.cuh, let
// spb_conv_tc_set_fuse1 accept 2 and dispatch on it in net.cu's forward().  SPB_C1P_VARIANT (env): bit 0 relaxed
// remote arrives, bit 1 cta-scope issuer waits polled by one lane, bit 2 probe the follower CTA instead of the leader.
//
// The fused first block (conv1a + conv1b + 2x2 max-pool, see conv1_fused_kernel in conv_tc.cu) on CTA PAIRS:
// tcgen05.mma.cta_group::2, M = 256 over two SMs, issued by the leader CTA of a 2-CTA cluster.
// Why: the single-CTA block sits on the SM's shared-memory bandwidth (273 KB per 128-pixel tile at 128 B/clk), and
// 74 KB of that is the conv1b weights every tile re-reads.  In a pair each CTA supplies only HALF of every weight
// block (32 of the 64 rows): 4 KB of A + 1 KB of B per MMA instead of 4 + 2 (43 instead of 48 cycles per MMA,
// tools/ubench_mma2.cu), without switching accumulators the way weight-stationary tile pairs had to.
// Everything but the MMA stream stays CTA-local: each CTA builds the im2col / conv1a activations of ITS tile
// (stage-1 warps), converts and stores ITS 128 accumulator rows (epilogue warps).  Cross-CTA protocol:
//   * barriers the tensor core completes (acc1_full, a2_empty, acc2_full) are signalled in both CTAs by multicast
//     tcgen05.commit;
//   * barriers the issuers wait for (a1_full, acc1_empty, a2_full, acc2_empty) live in the LEADER: one arrival per
//     producer warp, local for the leader's warps, a remote release-arrive (mapa) for the follower's -- no relay warp;
//   * the weights arrive by pair TMA (cta_group::2) completing on the leader's mbarrier.
// Tile t = blockIdx.x + j * gridDim.x as in the single-CTA kernel (gridDim.x even, so the pair owns tiles 2P, 2P+1);
// both CTAs run while the LEADER's tile exists, a follower tile beyond the end is a dummy (view index V: TMA zero
// fills the patch and clips the store).
#include <cstdlib>

#include "tc_common.cuh"

namespace spb {

struct Conv1PairSmem {
  uint8_t b2[9][4096];      // conv1b weights, 9 taps x this CTA's [32 rows x 128 B], SWIZZLE_128B (TMA)
  uint8_t a2[2][23552];     // conv1b A operand: 18x10 halo pixels x 128 B, SWIZZLE_128B, written by stage-1
  uint8_t ostage[2][4][1024];  // pooled output of one TMEM quadrant: 2 x 4 pixels x 128 B, SWIZZLE_128B (TMA store source)
  uint8_t a1[2][8192];      // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices
  uint8_t b1[1024];         // conv1a weights, this CTA's [32][16]
  float patch[4][320];      // zero-padded 20x16 fp32 image patch under the halo block, TMA
  int2 origin[4];           // (x0, y0) of the tile each patch belongs to
  float bias2[64];
  uint64_t bars[9 + 14 + 2 + 4];  // b_full[9], a1_full[2], acc1_full[2], acc1_empty[2], a2_full[2], a2_empty[2], acc2_full[2], acc2_empty[2], token[2], patch_full[4]
  uint32_t tmem_slot;
};

struct Conv1PairArgs {
  const __nv_bfloat16* w1a;     // [64][16]
  const float* bias2;
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;               // optional timeline probe (CTA 0): [64 iterations][16 slots] of clock64, or NULL
  int variant;                  // experiment bits: 1 relaxed remote arrives, 2 cta-scope issuer waits polled by one lane
};

__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

#ifndef SPB_C1P_TAP
#define SPB_C1P_TAP 8  // conv1a of the issuer's next tile pair is issued behind this tap of conv1b
#endif

__global__ void __launch_bounds__(576, 1) conv1_pair_kernel(const __grid_constant__ CUtensorMap tm_w2,
                                                            const __grid_constant__ CUtensorMap tm_out,
                                                            const __grid_constant__ CUtensorMap tm_x, Conv1PairArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Conv1PairSmem& sm = *reinterpret_cast<Conv1PairSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t bar0 = smem_u32(sm.bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a1_full = [&](int s) { return bar0 + 8u * (9 + s); };
  auto acc1_full = [&](int s) { return bar0 + 8u * (11 + s); };
  auto acc1_empty = [&](int s) { return bar0 + 8u * (13 + s); };
  auto a2_full = [&](int s) { return bar0 + 8u * (15 + s); };
  auto a2_empty = [&](int s) { return bar0 + 8u * (17 + s); };
  auto acc2_full = [&](int s) { return bar0 + 8u * (19 + s); };
  auto acc2_empty = [&](int s) { return bar0 + 8u * (21 + s); };
  auto token = [&](int s) { return bar0 + 8u * (23 + s); };
  auto patch_full = [&](int s) { return bar0 + 8u * (25 + s); };
  // one arrival per producer warp on the LEADER's barrier (the caller has ordered the warp's writes: proxy fence /
  // tcgen05 fence + __syncwarp)
  auto arrive_leader = [&](uint32_t bar) {
    if (rank == 0)
      mbar_arrive(bar);
    else if (a.variant & 1)
      mbar_arrive_remote_relaxed(bar, 0);
    else
      mbar_arrive_remote(bar, 0);
  };
  auto wait_pair = [&](uint32_t bar, uint32_t parity) {  // issuer side of a barrier both CTAs arrive on
    if (a.variant & 2) {
      if (lane == 0) mbar_wait(bar, parity);
      __syncwarp();
    } else {
      mbar_wait_cluster(bar, parity);
    }
  };
  // both CTAs of a pair run while the leader's tile exists: t < nt2 with ntiles rounded up to even
  const int nt2 = (a.ntiles + 1) & ~1;
  const int pc = (a.variant >> 2) & 1;  // probed CTA

  if (threadIdx.x < 64) sm.bias2[threadIdx.x] = a.bias2[threadIdx.x];
  if (threadIdx.x < 32) {  // this CTA's 32 rows of the conv1a weights
    const uint4* src = reinterpret_cast<const uint4*>(a.w1a + ((int)rank * 32 + threadIdx.x) * 16);
    const int n = threadIdx.x;
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + (n & 7) * 16) = src[0];
    *reinterpret_cast<uint4*>(sm.b1 + (n >> 3) * 256 + 128 + (n & 7) * 16) = src[1];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 9; ++s) mbar_init(b_full(s), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a1_full(s), 16);
      mbar_init(acc1_full(s), 1);
      mbar_init(acc1_empty(s), 16);
      mbar_init(a2_full(s), 16);
      mbar_init(a2_empty(s), 1);
      mbar_init(acc2_full(s), 1);
      mbar_init(acc2_empty(s), 16);
      mbar_init(token(s), 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(patch_full(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;
  // TMEM columns: conv1b accumulators [0,64) [64,128); conv1a accumulators buffer j at 128 + 128 j (two M halves)
  const int tiles_per_view = a.tiles_x * a.tiles_y;

  if (warp == 0 && lane == 0) {  // both CTAs: own half of every conv1b weight block, completing on the leader
    for (int tap = 0; tap < 9; ++tap) {
      if (rank == 0) mbar_expect_tx(b_full(tap), 8192);
      tma2_load_2d(smem_u32(sm.b2[tap]), &tm_w2, mapa_rank0(b_full(tap)), tap * 64, (int)rank * 32);
    }
  }
  if (warp <= 1) {
    if (rank == 0) {
      // =========================== MMA issuers (leader CTA) ===========================
      // Issuer q owns the tile pairs of parity q: conv1a(j) into acc1[q], conv1b(j) into acc2[q]; the two issuers take
      // turns on conv1b through a token handed over before the last two taps.  MMA order:
      // ... conv1b(j) | conv1a(j+2) | conv1b(j+1) | conv1a(j+3) ...
      __syncwarp();
      const int q = warp;
      constexpr uint32_t idesc = make_idesc(256, 64);
      constexpr uint64_t hi_a2 = ((uint64_t)(1280 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      constexpr uint64_t hi_b2 = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t b1desc = make_desc_noswz(smem_u32(sm.b1), 128, 256);
      const uint64_t a1desc = make_desc_noswz(smem_u32(sm.a1[q]), 128, 256);
      const uint32_t acc1 = tmem_base + 128 + q * 128, acc2 = tmem_base + q * 64;
      const uint32_t a2 = smem_u32(sm.a2[q]);
      auto issue_conv1a = [&](int k) {  // this issuer's k-th tile pair
        wait_pair(acc1_empty(q), (k & 1) ^ 1);
        wait_pair(a1_full(q), k & 1);
        tc_fence_after();
        if (elect_one()) {
          tc_mma2(acc1, a1desc, b1desc, idesc, 0u);
          tc_mma2(acc1 + 64, a1desc + (uint64_t)(4096 >> 4), b1desc, idesc, 0u);
          tc_commit2(acc1_full(q));
        }
        __syncwarp();
      };
#define SPB_PROBE(slot) \
  if (a.dbg && (int)blockIdx.x == pc && it < 64 && lane == 0) a.dbg[it * 16 + (slot)] = clock64()
      int k = 0;
      if ((int)(blockIdx.x + q * gridDim.x) < a.ntiles) issue_conv1a(0);
      for (int t = blockIdx.x + q * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++k) {
        const int it = 2 * k + q;
        SPB_PROBE(0);
        wait_pair(acc2_empty(q), (k & 1) ^ 1);
        SPB_PROBE(1);
        wait_pair(a2_full(q), k & 1);
        SPB_PROBE(2);
        if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
        SPB_PROBE(3);
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (k == 0) mbar_wait_cluster(b_full(tap), 0);
          const uint32_t tap_off = ((tap / 3) * 10 + (tap % 3)) * 128;
          const uint64_t adesc = hi_a2 | (uint64_t)(((a2 + tap_off) >> 4) | (1u << 16));
          const uint64_t bdesc = hi_b2 | (uint64_t)((smem_u32(sm.b2[tap]) >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma2(acc2, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (tap | kk) != 0 ? 1u : 0u);
            if (tap == 6) mbar_arrive(token(q ^ 1));
            if (tap == 8) {
              tc_commit2(a2_empty(q));
              tc_commit2(acc2_full(q));
            }
          }
          __syncwarp();
          if (tap == SPB_C1P_TAP && t + 2 * (int)gridDim.x < a.ntiles) issue_conv1a(k + 1);
        }
        SPB_PROBE(4);
      }
    }
  } else if (warp >= 10) {
    // =========================== stage 1 (both CTAs): im2col build + conv1a epilogue into conv1b's A operand ======
    const int sid = threadIdx.x - 320;           // 0..255
    const int quad = warp & 3;                   // TMEM lane quadrant of this warp
    const int chalf = (warp - 10) >> 2;          // which 32 of the 64 conv1a channels this warp converts
    const int lrow = quad * 32 + lane;           // TMEM lane; the thread owns halo pixels lrow and lrow + 128
    // warp 10 walks this CTA's tile sequence incrementally (see conv1_fused_kernel); a dummy tile lands on view V
    const int tpv = tiles_per_view;
    const int dgn = (int)gridDim.x / tpv, dgr = (int)gridDim.x % tpv;
    const int dgy = dgr / a.tiles_x, dgx = dgr % a.tiles_x;
    int pn = (int)blockIdx.x / tpv, pty = ((int)blockIdx.x % tpv) / a.tiles_x, ptx = ((int)blockIdx.x % tpv) % a.tiles_x;
    auto issue_patch = [&](int j) {
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
    auto build = [&](int j, int pit) {  // patch -> K=16 im2col rows of the 180 halo pixels of this CTA's j-th tile
      auto probe = [&](int slot) {
        if (a.dbg && (int)blockIdx.x == pc && pit >= 0 && pit < 64 && threadIdx.x == 320) a.dbg[pit * 16 + slot] = clock64();
      };
      const int buf = j & 1;  // (the caller has waited for patch_full(j & 3))
      probe(13);
      if (sid < 180) {
        const int2 org = sm.origin[j & 3];
        const float* pp = sm.patch[j & 3] + (sid / 10) * 16 + sid % 10 + 2;  // tap (0,0) of halo pixel sid
        // conv1b zero-pads conv1a's OUTPUT: a halo pixel outside the image gets an all-zero im2col row, bias slots
        // included
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
      probe(14);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_leader(a1_full(buf));
      probe(15);
    };
    int it = 0;
    if ((int)blockIdx.x < nt2) {
      issue_patch(0);
      if ((int)(blockIdx.x + gridDim.x) < nt2) issue_patch(1);
      if ((int)(blockIdx.x + 2 * gridDim.x) < nt2) issue_patch(2);
      mbar_wait(patch_full(0), 0);
      build(0, -1);
    }
    for (int t = blockIdx.x; t < nt2; t += gridDim.x, ++it) {
#define SPB_PROBE1(slot) \
  if (a.dbg && (int)blockIdx.x == pc && it < 64 && threadIdx.x == 320) a.dbg[it * 16 + (slot)] = clock64()
      SPB_PROBE1(5);
      const int buf = it & 1, ph = (it >> 1) & 1;
      const bool has_next = t + (int)gridDim.x < nt2;
      if (has_next) {
        mbar_wait(patch_full((it + 1) & 3), ((it + 1) >> 2) & 1);
        build(it + 1, it);
        if (t + 3 * (int)gridDim.x < nt2) issue_patch(it + 3);
      }
      SPB_PROBE1(6);
      if (lane == 0) mbar_wait(acc1_full(buf), ph);
      if (lane == 1) mbar_wait(a2_empty(buf), ph ^ 1);
      __syncwarp();
      SPB_PROBE1(8);
      tc_fence_after();
      uint32_t v[2][32];
      const uint32_t taddr = tmem_base + 128 + buf * 128 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      tc_ld_nowait<32>(taddr, v[0]);        // halo pixels lrow        (M half 0)
      tc_ld_nowait<32>(taddr + 64, v[1]);   // halo pixels lrow + 128  (M half 1)
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(acc1_empty(buf));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = h * 128 + lrow;
        if (p < 180) {  // bias and the out-of-image zeroing are already in the accumulator (see build)
          uint8_t* srow = sm.a2[buf] + p * 128;
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
      __syncwarp();
      if (lane == 0) arrive_leader(a2_full(buf));
      SPB_PROBE1(9);
    }
  } else {
    // =========================== conv1b epilogue (warps 2..9, both CTAs): bias + ReLU + 2x2 max-pool + bf16 =======
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    const bool owner = !(lane & 9);                          // even column, even row of the 2x2 window
    const int p = (lane >> 4) * 4 + ((lane & 7) >> 1);       // pooled pixel of the quadrant (2 rows x 4 columns)
    const bool issuer = chalf == 0 && lane == 0;
    float bz[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bz[j] = sm.bias2[chalf * 32 + j];
    int it = 0;
    for (int t = blockIdx.x; t < nt2; t += gridDim.x, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      const int n = t / tiles_per_view, rem = t % tiles_per_view;  // a dummy tile has n == V: the store is clipped
      const int y0 = (rem / a.tiles_x) * 16 + 4 * quad, x0 = (rem % a.tiles_x) * 8;
      if (a.dbg && (int)blockIdx.x == pc && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 10] = clock64();
      mbar_wait(acc2_full(buf), ph);
      if (a.dbg && (int)blockIdx.x == pc && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 11] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 64 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      uint32_t v[32];
      tc_ld_nowait<32>(taddr, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(acc2_empty(buf));
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
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(3 + quad, 64);
      if (issuer) {
        tma_store_4d(&tm_out, smem_u32(stg), 0, x0 >> 1, y0 >> 1, n);  // clipped at the image border
        tma_store_commit();
      }
      if (a.dbg && (int)blockIdx.x == pc && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 12] = clock64();
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees tensor memory) while the pair's MMAs / remote arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int conv1_pair(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
               cudaStream_t st) {
  if (!conv1_fused_supports(H, W)) {
    set_error("conv1_pair: needs even H and W %% 4 == 0 (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv1_pair: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  Conv1PairArgs a;
  a.w1a = L1a.w_tc;
  a.bias2 = L1b.bias;
  a.V = V;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, 8);
  a.tiles_y = ceil_div(H, 16);
  a.ntiles = V * a.tiles_x * a.tiles_y;
  a.dbg = conv1_probe_buf();
  a.variant = getenv("SPB_C1P_VARIANT") ? atoi(getenv("SPB_C1P_VARIANT")) : 0;
  CUtensorMap tm_w2, tm_out, tm_x;
  {  // conv1b weights [64][9*64] K-major: one box = HALF of the rows of one tap block
    cuuint64_t dims[2] = {(cuuint64_t)9 * 64, (cuuint64_t)L1b.cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)9 * 64 * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(&tm_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L1b.w_tc, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv1_pair: cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
      return SPB_ECUDA;
    }
  }
  {
    const int rc = encode_act_map(&tm_out, out, V, H / 2, W / 2, 64, 4, 2);
    if (rc != SPB_OK) return rc;
    // fp32 views [V,H,W]: box = the 16 x 20 patch under a halo block, zero filled outside the image
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)V};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {16, 20, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv1_pair: cuTensorMapEncodeTiled(views %dx%dx%d) failed with CUresult %d", V, H, W, (int)r);
      return SPB_ECUDA;
    }
  }
  const int smem = (int)sizeof(Conv1PairSmem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = 2 * ((a.ntiles + 1) / 2);
  if (grid > kNumSMs) grid = kNumSMs & ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(576);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv1_pair_kernel, tm_w2, tm_out, tm_x, a));
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // namespace spb
