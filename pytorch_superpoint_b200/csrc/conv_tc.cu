// (2) bf16 implicit-GEMM convolution on tcgen05 / TMEM fed by TMA  (sm_100a).
// Replaces the nn.Conv2d(+BatchNorm2d eval, ReLU, MaxPool2d) chain of
// models/SuperPointNet_pretrained.py:56-73, SuperPointNet_gauss2.py:53-63 / unet_parts.py:14-21,41-44.
//
// GEMM view per CTA tile:  D[128 pixels, Cout] = sum_{tap, cin-chunk} A_tap[128 px, 64] . W_tap[64, Cout]
//   * activations are NHWC bf16, so one pixel's 64 channels are exactly one 128-byte SWIZZLE_128B row;
//   * the output tile is 16 rows x 8 columns of pixels (M = 128); every 8-row UMMA core-matrix group
//     is one image row, groups are SBO apart -- therefore the 9 filter taps are 9 *shifted views* of ONE
//     halo block (18 x 10 pixels) that TMA loads once per (tile, 64-channel chunk): the A descriptor's
//     start address moves by (r*10 + s)*128 B, SBO stays 10*128 B.  Out-of-image halo pixels are zero
//     filled by TMA (negative / overflowing box coordinates) = the conv's zero padding;
//   * weights are K-major rows [Cout][tap*Cin + ci]; small layers keep all 9 (x chunks) blocks
//     resident in shared memory for the life of the persistent CTA, big ones stream them through a ring;
//   * accumulators live in TMEM (double buffered), one elected thread issues tcgen05.mma, four
//     epilogue warps read TMEM with tcgen05.ld and apply bias + ReLU + 2x2 max-pool (lane shuffles)
//     + bf16 pack (or fp32 / channel-L2-normalised output for the two heads).
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include "tc_common.cuh"

namespace spb {

// ------------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------------
// OUT_: 0 = ReLU + bf16 NHWC (encoder / head 3x3), 1 = raw fp32 rows staged through shared memory for coalesced
// stores (convPb: 65 channels), 2 = fp32 rows L2-normalised over the channels (convDb), 4 = convPb inside the HA pass:
// softmax probabilities (dustbin dropped) written as fp16 planes with the depth-to-space indexing and the valid mask
// applied (consumed by ha_fast.cu).
// NT_: output tiles per accumulator stage.  With streamed weights (128 -> 128) every weight block that arrives in shared
// memory is used for TWO tiles (8 MMAs instead of 4): the layer was bound by L2 -> SM weight traffic (288 KB per tile).
template <int CIN_, int NPAD_, int KS_, bool HALO_, bool BRES_, int NA_, int NB_, bool POOL_, int OUT_, int NT_ = 1>
struct ConvCfg {
  static constexpr int CIN = CIN_, NPAD = NPAD_, KS = KS_, NA = NA_, OUT = OUT_, NT = NT_;
  static constexpr bool POOL = POOL_;
  static constexpr int BW = (NPAD_ % 32 == 0) ? 32 : 16;  // TMEM columns per epilogue batch
  static constexpr int EPI_WARPS = OUT_ == 0 ? 8 : 4;     // bf16 layers: two warps per TMEM lane quadrant, half the columns each
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;  // + the second MMA issuer (last warp)
  static constexpr int STAGE_BYTES = OUT_ == 1 ? 4 * 32 * 65 * 4 : 0;
  // bf16 layers with >= 128 output channels stage each warp's 32 pixels x 64 channels (4 KB, SWIZZLE_128B) and
  // write it with one TMA tensor store: per-lane 16-byte global stores at a 256/512-byte stride cost one L1
  // wavefront per lane, in the shared-memory pipe the tensor core reads its operands through
  static constexpr bool TMA_OUT = OUT_ == 0 && NPAD_ >= 128;
  static constexpr int OSTAGE_BYTES = TMA_OUT ? 8 * 4096 : 0;
  static constexpr bool HALO = HALO_ && KS_ > 1, BRES = BRES_;
  static constexpr int TAPS = KS * KS, CHUNKS = CIN / 64;
  static constexpr int TH = 16, TW = 8;
  static constexpr int P = HALO ? TW + KS - 1 : TW;  // pixels per shared-memory row of an A block
  static constexpr int R = HALO ? TH + KS - 1 : TH;
  static constexpr int A_BYTES = P * R * 128;
  static constexpr int A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_BYTES = NPAD * 128;
  static constexpr int NBLK = TAPS * CHUNKS;
  static constexpr int NBS = BRES ? NBLK : NB_;
  static constexpr int TILE_COLS = NPAD <= 64 ? 64 : (NPAD <= 128 ? 128 : 256);  // TMEM columns of one tile
  static constexpr int ACC_STRIDE = NT * TILE_COLS;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static_assert(TMEM_COLS <= 512 && (NT == 1 || (HALO_ && KS_ > 1 && OUT_ == 0 && NPAD_ >= 128)), "tile pairs: halo + TMA-store layers");
  static constexpr int NBAR = 2 * NA + 2 * NBS + 4 + 2;  // + the two issue tokens
  static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)NA * A_STRIDE + (size_t)NBS * B_BYTES + OSTAGE_BYTES +
                                 NPAD * 4 + NBAR * 8 + 16 + STAGE_BYTES;
  static_assert(B_BYTES % 1024 == 0, "weight block must keep 1024-byte alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct ConvArgs {
  const float* bias;
  void* out;
  const void* aux;  // OUT == 4: valid-mask bits [B, 8H, W] of the warped views
  int B, H, W, cout;
  int tiles_x, tiles_y, ntiles;  // ntiles = real tiles; super-tile s covers tiles s*NT .. s*NT+NT-1 (missing ones are
  int nst;                       // dummies in image B: TMA zero-fills their loads and clips their stores)
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act,
                                                                const __grid_constant__ CUtensorMap tm_wgt,
                                                                const __grid_constant__ CUtensorMap tm_out, ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)C::NA * C::A_STRIDE;
  uint8_t* sOut = sB + (size_t)C::NBS * C::B_BYTES;  // 1024-byte aligned (A_STRIDE and B_BYTES are multiples of 1 KB)
  float* sBias = reinterpret_cast<float*>(sOut + C::OSTAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + C::NPAD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
  float* sStage = reinterpret_cast<float*>(tmem_slot + 4);  // OUT == 1 only
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (C::NA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * C::NA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * C::NA + C::NBS + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * C::NA + 2 * C::NBS + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * C::NA + 2 * C::NBS + 2 + s); };
  auto token = [&](int s) { return bar0 + 8u * (2 * C::NA + 2 * C::NBS + 4 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < C::NPAD; i += blockDim.x) sBias[i] = a.bias[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < C::NBS; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), C::EPI_WARPS);
      mbar_init(token(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      bool first = true;
      const int tpv = a.tiles_x * a.tiles_y;
      for (int st = blockIdx.x; st < a.nst; st += gridDim.x) {
        const int t = st * C::NT;  // (NT == 1: the tile itself)
        const int n = t / tpv, rem = t % tpv;
        const int y0 = (rem / a.tiles_x) * C::TH, x0 = (rem % a.tiles_x) * C::TW;
        for (int kc = 0; kc < C::CHUNKS; ++kc) {
          if (C::HALO) {
#pragma unroll
            for (int u = 0; u < C::NT; ++u) {
              const int tu = t + u, nu = tu < a.ntiles ? tu / tpv : a.B, ru = tu < a.ntiles ? tu % tpv : 0;
              mbar_wait(a_empty(sa), pa ^ 1);
              mbar_expect_tx(a_full(sa), C::A_BYTES);
              tma_load_4d(smem_u32(sA + (size_t)sa * C::A_STRIDE), &tm_act, a_full(sa), kc * 64,
                          (ru % a.tiles_x) * C::TW - C::KS / 2, (ru / a.tiles_x) * C::TH - C::KS / 2, nu);
              if (++sa == C::NA) { sa = 0; pa ^= 1; }
            }
          }
          for (int tap = 0; tap < C::TAPS; ++tap) {
            if (!C::HALO) {
              mbar_wait(a_empty(sa), pa ^ 1);
              mbar_expect_tx(a_full(sa), C::A_BYTES);
              tma_load_4d(smem_u32(sA + (size_t)sa * C::A_STRIDE), &tm_act, a_full(sa), kc * 64,
                          x0 + tap % C::KS - C::KS / 2, y0 + tap / C::KS - C::KS / 2, n);
              if (++sa == C::NA) { sa = 0; pa ^= 1; }
            }
            if (C::BRES) {
              if (first) {
                const int blk = kc * C::TAPS + tap;
                mbar_expect_tx(b_full(blk), C::B_BYTES);
                tma_load_2d(smem_u32(sB + (size_t)blk * C::B_BYTES), &tm_wgt, b_full(blk), tap * C::CIN + kc * 64, 0);
              }
            } else {
              mbar_wait(b_empty(sb), pb ^ 1);
              mbar_expect_tx(b_full(sb), C::B_BYTES);
              tma_load_2d(smem_u32(sB + (size_t)sb * C::B_BYTES), &tm_wgt, b_full(sb), tap * C::CIN + kc * 64, 0);
              if (++sb == C::NBS) { sb = 0; pb ^= 1; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1 || warp == 2 + C::EPI_WARPS) {
    // =========================== MMA issuers ===========================
    // TWO issuer warps alternate tiles (issuer q owns the tiles of parity q and the TMEM accumulator q).  The tensor
    // pipe's queue is only a few MMAs deep, so with a single issuer the barrier waits between two tiles (a few
    // hundred cycles of dependent shared-memory round trips) left the pipe idle ~1/3 of the time at N = 64.
    // Now issuer q^1 finishes all waits for its tile while q is still issuing, then blocks on a token that q hands
    // over BEFORE its last two tap groups: the other tile's MMAs queue up behind them and the pipe never drains.
    // (Order between the two issuers' MMAs is irrelevant: different accumulators; every tcgen05.commit tracks the
    // MMAs of its own thread only.)  Inside an issuer the whole warp walks the pipeline converged and one elected
    // lane issues; descriptors are a per-stage base plus compile-time constants.
    const int q = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = make_idesc(128, C::NPAD);
    constexpr uint64_t desc_hi_a = ((uint64_t)((C::P * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    constexpr uint64_t desc_hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    constexpr int A_PER_TILE = (C::HALO ? C::CHUNKS : C::CHUNKS * C::TAPS) * C::NT;  // A stages per super-tile
    // the token is handed over after this (chunk, tap) group has been issued
    constexpr int PASS_KC = C::TAPS >= 3 ? C::CHUNKS - 1 : C::CHUNKS - 2;
    constexpr int PASS_TAP = C::TAPS >= 3 ? C::TAPS - 3 : 0;
    static_assert(PASS_KC >= 0, "token hand-over point");
    auto advance = [](int& s, int& p, int n, int ring) {
      s += n;
      while (s >= ring) { s -= ring; p ^= 1; }
    };
    const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
    int sa = 0, pa = 0, sb = 0, pb = 0;
    if (q == 1) {
      advance(sa, pa, A_PER_TILE, C::NA);
      if (!C::BRES) advance(sb, pb, C::NBLK, C::NBS);
    }
    const uint32_t d_tmem = tmem_base + (uint32_t)q * C::ACC_STRIDE;
    int k = 0;  // this issuer's tile counter
    for (int t = blockIdx.x + q * gridDim.x; t < a.nst; t += 2 * gridDim.x, ++k) {
      const bool first = k == 0;
      mbar_wait(acc_empty(q), (k & 1) ^ 1);
      mbar_wait(a_full(sa), pa);  // first operand block of the tile (waited again, for free, in the loop)
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int kc = 0; kc < C::CHUNKS; ++kc) {
        int sau[C::NT];  // the A stages of this chunk, one per tile of the super-tile
        if (C::HALO) {
          int s2 = sa, p2 = pa;
#pragma unroll
          for (int u = 0; u < C::NT; ++u) {
            mbar_wait(a_full(s2), p2);
            sau[u] = s2;
            if (++s2 == C::NA) { s2 = 0; p2 ^= 1; }
          }
        }
#pragma unroll
        for (int tap = 0; tap < C::TAPS; ++tap) {
          if (!C::HALO) mbar_wait(a_full(sa), pa);
          const int slot = C::BRES ? kc * C::TAPS + tap : sb;
          if (C::BRES) {
            if (first) mbar_wait(b_full(slot), 0);  // weights stay resident after the first tile
          } else {
            mbar_wait(b_full(slot), pb);
          }
          tc_fence_after();
          const uint32_t tap_off = C::HALO ? ((tap / C::KS) * C::P + (tap % C::KS)) * 128 : 0;
          const uint64_t bdesc = desc_hi_b | (uint64_t)(((sB_u32 + (uint32_t)slot * C::B_BYTES) >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < C::NT; ++u) {
              const int su = C::HALO ? sau[u] : sa;
              const uint64_t adesc =
                  desc_hi_a | (uint64_t)(((sA_u32 + (uint32_t)su * C::A_STRIDE + tap_off) >> 4) | (1u << 16));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma(d_tmem + u * C::TILE_COLS, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                       (kc | tap | kk) != 0 ? 1u : 0u);
            }
            if (!C::BRES) tc_commit(b_empty(sb));
            if (!C::HALO) tc_commit(a_empty(sa));
            if (C::HALO && tap == C::TAPS - 1) {
#pragma unroll
              for (int u = 0; u < C::NT; ++u) tc_commit(a_empty(sau[u]));
            }
            if (kc == C::CHUNKS - 1 && tap == C::TAPS - 1) tc_commit(acc_full(q));
            if (kc == PASS_KC && tap == PASS_TAP) mbar_arrive(token(q ^ 1));
          }
          __syncwarp();
          if (!C::BRES) {
            if (++sb == C::NBS) { sb = 0; pb ^= 1; }
          }
          if (!C::HALO) {
            if (++sa == C::NA) { sa = 0; pa ^= 1; }
          }
        }
        if (C::HALO) advance(sa, pa, C::NT, C::NA);
      }
      // skip the other issuer's tile in both rings
      advance(sa, pa, A_PER_TILE, C::NA);
      if (!C::BRES) advance(sb, pb, C::NBLK, C::NBS);
    }
  } else {
    // =========================== epilogue (warps 2..5, or 2..9 for the bf16 layers) ===========================
    // TMEM -> registers in batches of BW columns, software pipelined (the next batch's tcgen05.ld is in
    // flight while this one is processed); the accumulator is handed back to the MMA warp as soon as the
    // last batch has landed in registers.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int m = quad * 32 + lane, py = m >> 3, px = m & 7;
    constexpr int COLS = C::NPAD / (C::EPI_WARPS / 4);  // accumulator columns this warp converts
    constexpr int BW = C::BW, NBATCH = COLS / BW;
    const int col0 = ((warp - 2) >> 2) * COLS;
    int as = 0, pacc = 0;
    const int tpv = a.tiles_x * a.tiles_y;
    for (int sti = blockIdx.x; sti < a.nst; sti += gridDim.x) {
      mbar_wait(acc_full(as), pacc);
      tc_fence_after();
#pragma unroll 1
     for (int u = 0; u < C::NT; ++u) {  // the tiles of this accumulator stage
      const int t = sti * C::NT + u;
      const int n = t < a.ntiles ? t / tpv : a.B, rem = t < a.ntiles ? t % tpv : 0;  // dummy tile: image B (clipped)
      const int y0 = (rem / a.tiles_x) * C::TH, x0 = (rem % a.tiles_x) * C::TW;
      const int y = y0 + py, x = x0 + px;
      const uint32_t taddr =
          tmem_base + (uint32_t)as * C::ACC_STRIDE + (uint32_t)u * C::TILE_COLS + ((uint32_t)(quad * 32) << 16);
      bool valid = y < a.H && x < a.W;
      long long opix;
      if (C::POOL) {
        valid = valid && !(px & 1) && !(py & 1);
        opix = ((long long)n * (a.H >> 1) + (y >> 1)) * (a.W >> 1) + (x >> 1);
      } else {
        opix = ((long long)n * a.H + y) * a.W + x;
      }
      float norm = 1.f;
      if (C::OUT == 2) {  // first pass over the accumulator: squared norm of the 256-channel row
        float ss = 0.f;
        for (int c0 = 0; c0 < C::NPAD; c0 += 16) {
          uint32_t v[16];
          tc_ld_nowait<16>(taddr + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float f = __uint_as_float(v[j]) + sBias[c0 + j];
            ss = fmaf(f, f, ss);
          }
        }
        norm = sqrtf(ss);
      }
      if (C::OUT == 4) {
        // convPb in the HA path: 65-way softmax of the cell in registers (each thread holds its whole row), dustbin
        // dropped -- utils/utils.py:514-523 fused here.
        uint32_t w[80];
#pragma unroll
        for (int b = 0; b < 5; ++b) tc_ld_nowait<16>(taddr + b * 16, *reinterpret_cast<uint32_t(*)[16]>(&w[b * 16]));
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(as));
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 65; ++j) {
          const float f = __uint_as_float(w[j]) + sBias[j];
          w[j] = __float_as_uint(f);
          mx = fmaxf(mx, f);
        }
        // The logits come from bf16 tensor-core products (parity is a 1e-2 tolerance on the heat-map, not bit-exact),
        // so the softmax uses the fast exponential and ONE reciprocal per cell: the IEEE expf/div sequences were
        // ~1800 instructions per thread and made this epilogue the layer's critical path.
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 65; ++j) {
          const float e = __expf(__uint_as_float(w[j]) - mx);
          w[j] = __float_as_uint(e);
          sum += e;
        }
        const float rs = __frcp_rn(sum);
        {
          // depth-to-space + mask + fp16 in the store (utils/d2s.py:8-25, export.py:52): row r of the 8x8 cell is
          // 16 bytes of plane row 8y + r; the 8 lanes of a cell row of the tile write 128 contiguous bytes.
          // A pixel of the warped view that lies outside the source image (mask bit 0) is stored as kHeatMasked.
          if (valid && t < a.ntiles) {
            const int Wh = a.W * 8 + 2 * kHeatPadX, Hh = a.H * 8 + 2 * kHeatPadY;
            const uint8_t* mb = static_cast<const uint8_t*>(a.aux) + ((long long)n * (a.H * 8) + y * 8) * a.W + x;
            __half* o = static_cast<__half*>(a.out) + ((long long)n * Hh + y * 8 + kHeatPadY) * Wh + kHeatPadX + x * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned m = __ldg(mb + r * a.W);
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(__uint_as_float(w[r * 8 + 2 * q]) * rs,
                                                    __uint_as_float(w[r * 8 + 2 * q + 1]) * rs);
                const unsigned b2 = (m >> (2 * q)) & 3u;
                const unsigned sel = (b2 & 1u) * 0xFFFFu | (b2 >> 1) * 0xFFFF0000u;
                pk[q] = (*reinterpret_cast<const uint32_t*>(&h) & sel) | ((kHeatMasked * 0x10001u) & ~sel);
              }
              *reinterpret_cast<uint4*>(o + (long long)r * Wh) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          continue;
        }
      }
      uint32_t v[2][BW];
      tc_ld_nowait<BW>(taddr + col0, v[0]);
#pragma unroll
      for (int b = 0; b < NBATCH; ++b) {
        tc_wait_ld();
        if (b + 1 < NBATCH) {
          tc_ld_nowait<BW>(taddr + col0 + (b + 1) * BW, v[(b + 1) & 1]);
        } else if (u == C::NT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(as));
        }
        const int c0 = col0 + b * BW;
        if constexpr (C::TMA_OUT) {
          // two 32-channel batches fill one 128-byte row per pixel of this warp's staging tile; then ONE TMA store
          uint8_t* stg = sOut + (warp - 2) * 4096;
          const int p = C::POOL ? (lane >> 4) * 4 + ((lane & 7) >> 1) : lane;  // staging row of this lane's pixel
          uint32_t pk[16];
          epi_pack32<C::POOL>(v[b & 1], sBias + c0, pk);
          if ((b & 1) == 0) {  // the previous store out of this tile must have finished reading it
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          if (!C::POOL || !(lane & 9)) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(stg + p * 128 + ((((b & 1) * 4 + q4) ^ (p & 7)) * 16)) =
                  make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
          if (b & 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              const int yq = y0 + 4 * quad;
              if (C::POOL)
                tma_store_4d(&tm_out, smem_u32(stg), c0 - 32, x0 >> 1, yq >> 1, n);
              else
                tma_store_4d(&tm_out, smem_u32(stg), c0 - 32, x0, yq, n);
              tma_store_commit();
            }
          }
          continue;
        } else if constexpr (C::OUT == 0) {
          epi_bf16_32<C::POOL>(v[b & 1], sBias + c0, valid, static_cast<__nv_bfloat16*>(a.out) + opix * a.cout + c0);
          continue;
        }
        float f[BW];
#pragma unroll
        for (int j = 0; j < BW; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(sBias + c0 + j);
          f[j] = __uint_as_float(v[b & 1][j]) + bb.x;
          f[j + 1] = __uint_as_float(v[b & 1][j + 1]) + bb.y;
          f[j + 2] = __uint_as_float(v[b & 1][j + 2]) + bb.z;
          f[j + 3] = __uint_as_float(v[b & 1][j + 3]) + bb.w;
        }
        if (C::OUT == 0) {
#pragma unroll
          for (int j = 0; j < BW; ++j) {
            f[j] = fmaxf(f[j], 0.f);
            if (C::POOL) {
              f[j] = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 1));
              f[j] = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 8));
            }
          }
          if (valid) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(a.out) + opix * a.cout + c0;
#pragma unroll
            for (int j = 0; j < BW; j += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
                pk[q] = *reinterpret_cast<const uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        } else if (C::OUT == 1) {
          // stage the warp's 32 pixels x cout floats exactly as they lie in global memory (row stride = cout)
          float* st = sStage + quad * (32 * 65) + lane * a.cout + c0;
#pragma unroll
          for (int j = 0; j < BW; ++j)
            if (c0 + j < a.cout) st[j] = f[j];
        } else {
          if (valid) {
            float* o = static_cast<float*>(a.out) + opix * a.cout + c0;
#pragma unroll
            for (int j = 0; j < BW; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(__fdiv_rn(f[j], norm), __fdiv_rn(f[j + 1], norm),
                                                              __fdiv_rn(f[j + 2], norm), __fdiv_rn(f[j + 3], norm));
          }
        }
      }
      if (C::OUT == 1) {
        // coalesced copy-out: the warp owns 4 image rows x 8 pixels; each row's 8*cout floats are contiguous
        __syncwarp();
        const int seg = 8 * a.cout;
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
          const int yy = y0 + quad * 4 + r;
          if (yy < a.H && x0 < a.W) {
            float* o = static_cast<float*>(a.out) + (((long long)n * a.H + yy) * a.W + x0) * a.cout;
            const float* src = sStage + quad * (32 * 65) + r * seg;
            const int lim = min(seg, (a.W - x0) * a.cout);
            for (int i = lane; i < lim; i += 32) o[i] = src[i];
          }
        }
        __syncwarp();
      }
     }  // tiles of the stage
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
    if (C::TMA_OUT && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// conv1a (1 -> 64, 3x3) on the tensor cores: 128 consecutive pixels per tile, the 9 taps of every pixel are
// gathered straight from the fp32 image into a K=16 im2col row (bf16, no-swizzle K-major core matrices:
// 8 rows x 16 B, K chunks LBO = 128 B apart, 8-row groups SBO = 256 B apart), ONE tcgen05.mma (M128 N64 K16)
// per tile, epilogue bias + ReLU + bf16 and a fully contiguous 16 KB NHWC store.  HBM-write bound.
// Warps 0..3: im2col build of tile i+1, then epilogue of tile i (double-buffered A and TMEM); warp 4: MMA.
// ------------------------------------------------------------------------------------------------
struct Conv1aSmem {
  uint8_t stage[2][16384];  // output tiles, 128 rows x 128 B, SWIZZLE_128B (TMA store source)
  uint8_t a[2][4096];       // im2col tiles
  uint8_t b[2048];          // weights [64][16] in core-matrix layout
  float bias[64];
  uint64_t bars[6];         // a_full[2], acc_full[2], acc_empty[2]
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(160, 4) conv1a_tc_kernel(const __grid_constant__ CUtensorMap tm_out,
                                                           const float* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ w16,
                                                           const float* __restrict__ bias, long long npix, int H, int W,
                                                           int ntiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Conv1aSmem& sm = *reinterpret_cast<Conv1aSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(sm.bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (4 + s); };

  if (threadIdx.x < 64) {
    sm.bias[threadIdx.x] = bias[threadIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(w16 + threadIdx.x * 16);  // weight row n: two 16-byte K chunks
    const int n = threadIdx.x;
    *reinterpret_cast<uint4*>(sm.b + (n >> 3) * 256 + (n & 7) * 16) = src[0];
    *reinterpret_cast<uint4*>(sm.b + (n >> 3) * 256 + 128 + (n & 7) * 16) = src[1];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 128);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sm.b written by the generic proxy, read by UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(128, 64);
    const uint64_t bdesc = make_desc_noswz(smem_u32(sm.b), 128, 256);
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc_empty(buf), ph ^ 1);
      mbar_wait(a_full(buf), ph);
      tc_fence_after();
      if (elect_one()) {
        tc_mma(tmem_base + buf * 64, make_desc_noswz(smem_u32(sm.a[buf]), 128, 256), bdesc, idesc, 0u);
        tc_commit(acc_full(buf));
      }
      __syncwarp();
    }
  } else {
    const int r = threadIdx.x;  // im2col row / TMEM lane / output row of this thread
    auto build = [&](int t, int buf) {
      const long long p = (long long)t * 128 + r;
      uint32_t pk[5] = {0, 0, 0, 0, 0};
      if (p < npix) {
        const int px = (int)(p % W), py = (int)((p / W) % H);
        const float* img = x + (p / ((long long)W * H)) * (long long)H * W;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int yy = py + k / 3 - 1, xx = px + k % 3 - 1;
          v[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (long long)yy * W + xx) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
          pk[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[8], 1.f);  // slot 9 = 1: bias head (folded into the MMA)
        pk[4] = *reinterpret_cast<const uint32_t*>(&h);
      }
      uint8_t* row = sm.a[buf] + (r >> 3) * 256 + (r & 7) * 16;
      *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(row + 128) = make_uint4(pk[4], p < npix ? 0x00003F80u : 0u, 0, 0);  // slot 10 = 1: bias tail
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a_full(buf));
    };
    int it = 0;
    if (blockIdx.x < ntiles) build(blockIdx.x, 0);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      if (t + (int)gridDim.x < ntiles) build(t + gridDim.x, buf ^ 1);
      // the TMA store that last read stage[buf] (two tiles ago) must have drained before we overwrite it
      if (threadIdx.x == 0) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
      mbar_wait(acc_full(buf), ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 64 + ((uint32_t)(warp * 32) << 16);
      uint8_t* srow = sm.stage[buf] + r * 128;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        uint32_t v[32];
        tc_ld_nowait<32>(taddr + b * 32, v);
        tc_wait_ld();
        if (b == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(buf));
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) pk[q] = relu_pack_bf16(v[j + 2 * q], v[j + 2 * q + 1]);  // bias is in the MMA
          const int chunk = (b * 4 + j / 8) ^ (r & 7);  // SWIZZLE_128B: 16-byte chunk XOR (row mod 8)
          *reinterpret_cast<uint4*>(srow + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        tma_store_2d(&tm_out, smem_u32(sm.stage[buf]), 0, t * 128);  // rows beyond npix are clipped by TMA
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// conv1a + conv1b + 2x2 max-pool in ONE kernel (the first VGG block, 44 % of the network's FLOPs).
// conv1a's 64-channel output never touches HBM: per 16x8 output tile, the four "stage-1" warps gather the
// 9 taps of the 18x10 halo pixels from the fp32 view into a K=16 im2col tile, two tcgen05.mma (M128 N64 K16)
// produce conv1a for the 180 halo pixels in TMEM, the same warps apply bias + ReLU (and zero the pixels that
// lie outside the image -- conv1b pads conv1a's OUTPUT with zeros), pack bf16 and write the rows straight into
// conv1b's A operand in shared memory in the SWIZZLE_128B layout the shifted-window descriptors expect.
// conv1b then runs exactly as in conv_tc_kernel (36 MMAs, weights resident, double-buffered TMEM, pooled
// epilogue).  Warps: 0 weight loader, 1 MMA issuer, 2..9 conv1b epilogue, 10..17 stage-1 (in both groups two
// warps share a TMEM lane quadrant, each converting one half of the 64 channels).
// The MMA order is  conv1a(i+1), conv1b(i), conv1a(i+2), ...  so stage-1 always works one tile ahead.
// ------------------------------------------------------------------------------------------------
struct Conv1FusedSmem {
  uint8_t b2[9][8192];      // conv1b weights, 9 taps x [64 rows x 128 B], SWIZZLE_128B (TMA)
  uint8_t a2[2][23552];     // conv1b A operand: 18x10 halo pixels x 128 B, SWIZZLE_128B, written by stage-1
  uint8_t ostage[2][4][1024];  // pooled output of one TMEM quadrant: 2 x 4 pixels x 128 B, SWIZZLE_128B (TMA store source)
  uint8_t a1[2][8192];      // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices
  uint8_t b1[2048];         // conv1a weights [64][16]
  float patch[4][320];      // zero-padded 20x16 fp32 image patch under the halo block (rows y0-2.., cols x0-4..), TMA
  int2 origin[4];           // (x0, y0) of the tile each patch belongs to
  float bias1[64], bias2[64];
  uint64_t bars[9 + 14 + 2 + 4];  // b_full[9], a1_full[2], acc1_full[2], acc1_empty[2], a2_full[2], a2_empty[2], acc2_full[2], acc2_empty[2], token[2], patch_full[4]
  uint32_t tmem_slot;
};

#ifndef SPB_C1A_TAP
#define SPB_C1A_TAP 8  // conv1a of the issuer's next tile is issued behind this tap of conv1b
#endif
// Measured alternatives for getting conv1a's accumulator to stage 1 sooner (it waits ~750 cycles for it): issuing
// conv1a behind tap 3 or 6 (-10..14 %), and a THIRD issuer warp that issues every conv1a the moment its im2col tile
// is ready (+-0 %): the tensor pipe executes in issue order behind a queue of conv1b MMAs either way, and the block
// already sits at 84 % of the shared-memory operand-read bound.
struct Conv1FusedArgs {
  const __nv_bfloat16* w1a;     // [64][16]
  const float *bias1, *bias2;
  __nv_bfloat16* out;           // [V,H/2,W/2,64]
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;               // optional timeline probe (CTA 0): [64 iterations][16 slots] of clock64, or NULL
};

__global__ void __launch_bounds__(576, 1) conv1_fused_kernel(const __grid_constant__ CUtensorMap tm_w2,
                                                             const __grid_constant__ CUtensorMap tm_out,
                                                             const __grid_constant__ CUtensorMap tm_x, Conv1FusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Conv1FusedSmem& sm = *reinterpret_cast<Conv1FusedSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

  if (threadIdx.x < 64) {
    sm.bias1[threadIdx.x] = a.bias1[threadIdx.x];
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
  // TMEM columns: conv1b accumulators [0,64) [64,128); conv1a accumulators buffer j at 128 + 128 j (two M halves)
  const int tiles_per_view = a.tiles_x * a.tiles_y;

  if (warp <= 1) {
    // =========================== MMA issuers (warp 0 also loads the conv1b weights first) ===========================
    // Issuer q owns the tiles of parity q: conv1a(j) into acc1[q], conv1b(j) into acc2[q].  The two issuers take
    // turns on conv1b through a token handed over before the last two taps (see conv_tc_kernel), so one of them is
    // always through its barrier waits when the other runs out of MMAs.  Global MMA order stays
    // ... conv1b(j) | conv1a(j+2) | conv1b(j+1) | conv1a(j+3) ...  : stage 1 works one tile ahead of conv1b.
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
    const uint64_t b1desc = make_desc_noswz(smem_u32(sm.b1), 128, 256);
    const uint64_t a1desc = make_desc_noswz(smem_u32(sm.a1[q]), 128, 256);
    const uint32_t acc1 = tmem_base + 128 + q * 128, acc2 = tmem_base + q * 64;
    const uint32_t a2 = smem_u32(sm.a2[q]);
    auto issue_conv1a = [&](int k) {  // this issuer's k-th tile
      mbar_wait(acc1_empty(q), (k & 1) ^ 1);
      mbar_wait(a1_full(q), k & 1);
      tc_fence_after();
      if (elect_one()) {
        tc_mma(acc1, a1desc, b1desc, idesc, 0u);
        tc_mma(acc1 + 64, a1desc + (uint64_t)(4096 >> 4), b1desc, idesc, 0u);
        tc_commit(acc1_full(q));
      }
      __syncwarp();
    };
#define SPB_PROBE(slot)                                                                       \
  if (a.dbg && blockIdx.x == 0 && it < 64 && lane == 0) a.dbg[it * 16 + (slot)] = clock64()
    int k = 0;
    if ((int)(blockIdx.x + q * gridDim.x) < a.ntiles) issue_conv1a(0);
    for (int t = blockIdx.x + q * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++k) {
      const int it = 2 * k + q;
      SPB_PROBE(0);
      mbar_wait(acc2_empty(q), (k & 1) ^ 1);
      SPB_PROBE(1);
      mbar_wait(a2_full(q), k & 1);
      SPB_PROBE(2);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      SPB_PROBE(3);
      tc_fence_after();
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        if (k == 0) mbar_wait(b_full(tap), 0);
        const uint32_t tap_off = ((tap / 3) * 10 + (tap % 3)) * 128;
        const uint64_t adesc = hi_a2 | (uint64_t)(((a2 + tap_off) >> 4) | (1u << 16));
        const uint64_t bdesc = hi_b2 | (uint64_t)((smem_u32(sm.b2[tap]) >> 4) | (1u << 16));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(acc2, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (tap | kk) != 0 ? 1u : 0u);
          if (tap == 6) mbar_arrive(token(q ^ 1));
          if (tap == 8) {
            tc_commit(a2_empty(q));
            tc_commit(acc2_full(q));
          }
        }
        __syncwarp();
        if (tap == SPB_C1A_TAP && t + 2 * (int)gridDim.x < a.ntiles) issue_conv1a(k + 1);
      }
      SPB_PROBE(4);
    }
  } else if (warp >= 10) {
    // =========================== stage 1: im2col build + conv1a epilogue into conv1b's A operand =========
    const int sid = threadIdx.x - 320;           // 0..255
    const int quad = warp & 3;                   // TMEM lane quadrant of this warp
    const int chalf = (warp - 10) >> 2;          // which 32 of the 64 conv1a channels this warp converts
    const int lrow = quad * 32 + lane;           // TMEM lane; the thread owns halo pixels lrow and lrow + 128
    // The zero-padded fp32 patch under tile j's halo block arrives by TMA (box 16 x 20 of the [V,H,W] view tensor
    // at (x0-4, y0-2): the innermost start must be 16-byte aligned, so the 12 columns needed sit at offsets 2..13;
    // out-of-image elements are zero filled) two tiles ahead, issued by thread 320 together with the tile's
    // origin; nobody in stage 1 computes a global address or waits at a CTA barrier.  Four patch buffers: the TMA
    // for tile j+2 overwrites the buffer of tile j-2, which every stage-1 thread finished reading before it arrived
    // on a1_full(j-2) -- and thread 320 has since waited for conv1a(j-1), which needed all those arrivals.
    // Warp 10 walks the tile sequence of this CTA incrementally (no divisions) and one elected lane issues.
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
    auto build = [&](int j, int pit) {  // patch -> K=16 im2col rows of the 180 halo pixels of this CTA's j-th tile
      auto probe = [&](int slot) {
        if (a.dbg && blockIdx.x == 0 && pit >= 0 && pit < 64 && threadIdx.x == 320) a.dbg[pit * 16 + slot] = clock64();
      };
      const int buf = j & 1;  // (the caller has waited for patch_full(j & 3))
      probe(13);
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
      probe(14);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a1_full(buf));
      probe(15);
    };
    int it = 0;
    if ((int)blockIdx.x < a.ntiles) {
      issue_patch(0);
      if ((int)(blockIdx.x + gridDim.x) < a.ntiles) issue_patch(1);
      if ((int)(blockIdx.x + 2 * gridDim.x) < a.ntiles) issue_patch(2);
      mbar_wait(patch_full(0), 0);
      build(0, -1);
    }
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
#define SPB_PROBE1(slot)                                                                                 \
  if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 320) a.dbg[it * 16 + (slot)] = clock64()
      SPB_PROBE1(5);
      const int buf = it & 1, ph = (it >> 1) & 1;
      const bool has_next = t + (int)gridDim.x < a.ntiles;
      if (has_next) {
        // im2col of tile it+1 FIRST, and without waiting for anything conv1a(it) related: conv1a(it+1) then has a
        // whole iteration to get through the MMA queue before its accumulator is needed
        mbar_wait(patch_full((it + 1) & 3), ((it + 1) >> 2) & 1);
        build(it + 1, it);
        if (t + 3 * (int)gridDim.x < a.ntiles) issue_patch(it + 3);   // patch of tile it+3
      }
      SPB_PROBE1(6);
      // the two barriers epi1 depends on are polled by two different lanes at once (a satisfied mbarrier wait still
      // costs a ~200-cycle shared-memory round trip under load)
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
      if (lane == 0) mbar_arrive(acc1_empty(buf));
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
      mbar_arrive(a2_full(buf));
      SPB_PROBE1(9);
    }
  } else {
    // =========================== conv1b epilogue (warps 2..9): bias + ReLU + 2x2 max-pool + bf16 ===========
    // The two warps of a TMEM quadrant (32 channels each) fill the 8 pooled pixels x 128 B of that quadrant in a
    // SWIZZLE_128B staging tile and ONE TMA tensor store writes it: 8-lane 16-byte global stores cost a wavefront
    // per lane in the pipe the tensor core reads its operands through; so do bias LDS -> biases live in registers.
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    const bool owner = !(lane & 9);                          // even column, even row of the 2x2 window
    const int p = (lane >> 4) * 4 + ((lane & 7) >> 1);       // pooled pixel of the quadrant (2 rows x 4 columns)
    const bool issuer = chalf == 0 && lane == 0;
    float bz[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bz[j] = sm.bias2[chalf * 32 + j];
    int it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      const int n = t / tiles_per_view, rem = t % tiles_per_view;
      const int y0 = (rem / a.tiles_x) * 16 + 4 * quad, x0 = (rem % a.tiles_x) * 8;
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 10] = clock64();
      mbar_wait(acc2_full(buf), ph);
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 11] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 64 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      uint32_t v[32];
      tc_ld_nowait<32>(taddr, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty(buf));
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
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 12] = clock64();
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
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_weight_map(CUtensorMap* tm, const LayerDev& L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  const cuuint64_t K = (cuuint64_t)L.ks * L.ks * L.cin;
  cuuint64_t dims[2] = {K, (cuuint64_t)L.cout_pad};
  cuuint64_t strides[1] = {K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)L.cout_pad};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w_tc, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return SPB_ECUDA;
  }
  return SPB_OK;
}

int encode_act_map(CUtensorMap* tm, const void* in, int B, int H, int W, int C, int boxw, int boxh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)boxw, (cuuint32_t)boxh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled(activations %dx%dx%dx%d box %dx%d) failed with CUresult %d", B, H, W, C,
              boxw, boxh, (int)r);
    return SPB_ECUDA;
  }
  return SPB_OK;
}

int conv1a_tc(const LayerDev& L, const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st) {
  const long long npix = (long long)B * H * W;
  const int ntiles = ceil_div(npix, 128);
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv1a_tc: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, (cuuint64_t)npix};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv1a_tc: cuTensorMapEncodeTiled(output) failed with CUresult %d", (int)r);
    return SPB_ECUDA;
  }
  const int smem = (int)sizeof(Conv1aSmem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = kNumSMs * 4;
  if (grid > ntiles) grid = ntiles;
  conv1a_tc_kernel<<<grid, 160, smem, st>>>(tm, x, L.w_tc, L.bias, npix, H, W, ntiles);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

static long long* g_conv1_dbg = nullptr;
extern "C" int spb_conv1_fused_set_probe(long long* dev_buf) {  // debug: timeline probe buffer [64][16] or NULL
  g_conv1_dbg = dev_buf;
  return SPB_OK;
}
long long* conv1_probe_buf() { return g_conv1_dbg; }
// conv1a fused into conv1b's kernel: 0 two kernels (A/B tests), 1 the fused block; set_fuse1(-1) restores the default.
// (Three other fused designs were built, measured slower and removed from the library: column-stacked N = 192 and
// N = 128 + 64 in round 1, weight-stationary tile pairs -- tools/experiments/conv1_ws.cu -- and the block on CTA pairs
// (cta_group::2) -- tools/experiments/conv1_pair.cu -- in round 2)
constexpr int kFuse1Default = 1;
static int g_fuse1 = kFuse1Default;
extern "C" int spb_conv_tc_set_fuse1(int on) {
  g_fuse1 = on < 0 ? kFuse1Default : on > 1 ? 1 : on;
  return SPB_OK;
}
bool conv1_fused_enabled() { return g_fuse1 != 0; }

bool conv1_fused_supports(int H, int W) { return !((H | W) & 1) && !(W & 3); }

int conv1_fused(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
                cudaStream_t st) {
  if (!conv1_fused_supports(H, W)) {
    set_error("conv1_fused: needs even H and W %% 4 == 0 (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  CUtensorMap tm_w2;
  memcpy(&tm_w2, L1b.tmap_w, sizeof(tm_w2));
  Conv1FusedArgs a;
  a.w1a = L1a.w_tc;
  a.bias1 = L1a.bias;
  a.bias2 = L1b.bias;
  a.out = out;
  a.V = V;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, 8);
  a.tiles_y = ceil_div(H, 16);
  a.ntiles = V * a.tiles_x * a.tiles_y;
  a.dbg = g_conv1_dbg;
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
      set_error("conv1_fused: cuTensorMapEncodeTiled(views %dx%dx%d) failed with CUresult %d", V, H, W, (int)r);
      return SPB_ECUDA;
    }
  }
  const int smem = (int)sizeof(Conv1FusedSmem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv1_fused_kernel<<<grid, 576, smem, st>>>(tm_w2, tm_out, tm_x, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int conv_tc_prepare(LayerDev& L) {
  L.tmap_w = malloc(sizeof(CUtensorMap) + 64);
  if (!L.tmap_w) return SPB_ENOMEM;
  CUtensorMap tm;
  const int rc = encode_weight_map(&tm, L);
  if (rc != SPB_OK) return rc;
  memcpy(L.tmap_w, &tm, sizeof(tm));
  return SPB_OK;
}

static int g_pair = 1;  // 128 -> 128 layers: two tiles per accumulator stage / weight pass (0: one, for A/B tests)
extern "C" int spb_conv_tc_set_pair(int on) {
  g_pair = on ? 1 : 0;
  return SPB_OK;
}
static int g_halo_mode = 1;  // 1: one halo block + shifted descriptors; 0: per-tap reload (debug/A-B)
extern "C" int spb_conv_tc_set_halo(int on) {
  g_halo_mode = on ? 1 : 0;
  return SPB_OK;
}

template <class C>
static int launch(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode,
                  cudaStream_t st, const void* aux) {
  CUtensorMap tm_act, tm_wgt, tm_out;
  memcpy(&tm_wgt, L.tmap_w, sizeof(tm_wgt));
  int rc = encode_act_map(&tm_act, in, B, H, W, L.cin, C::P, C::R);
  if (rc != SPB_OK) return rc;
  if (C::TMA_OUT) {  // one warp's quadrant: 4 rows x 8 pixels (pooled: 2 x 4) x 64 channels
    rc = C::POOL ? encode_act_map(&tm_out, out, B, H / 2, W / 2, L.cout, 4, 2)
                 : encode_act_map(&tm_out, out, B, H, W, L.cout, 8, 4);
    if (rc != SPB_OK) return rc;
  } else {
    tm_out = tm_act;
  }
  ConvArgs a;
  a.bias = L.bias;
  a.out = out;
  a.aux = aux;
  a.B = B;
  a.H = H;
  a.W = W;
  a.cout = L.cout;
  if (L.pool != (C::POOL ? 1 : 0) || L.relu != (C::OUT == 0 ? 1 : 0) || out_mode != C::OUT ||
      ((C::OUT == 1 || C::OUT == 4) && L.cout != 65) || (C::OUT == 4 && !aux)) {
    set_error("conv_tc: kernel configuration does not match the layer (pool %d relu %d out_mode %d)", L.pool, L.relu,
              out_mode);
    return SPB_EINVAL;
  }
  a.tiles_x = ceil_div(W, C::TW);
  a.tiles_y = ceil_div(H, C::TH);
  a.ntiles = B * a.tiles_x * a.tiles_y;
  a.nst = (a.ntiles + C::NT - 1) / C::NT;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int grid = a.nst < kNumSMs ? a.nst : kNumSMs;
  conv_tc_kernel<C><<<grid, C::THREADS, C::SMEM, st>>>(tm_act, tm_wgt, tm_out, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int conv_tc(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode,
            cudaStream_t st, const void* aux) {
  if (!L.tmap_w) {
    set_error("conv_tc: layer has no weight tensor map");
    return SPB_EINVAL;
  }
  if (L.pool && ((H | W) & 1)) {
    set_error("conv_tc: pooled layer needs even H, W (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  if (L.w_s3 && conv_s3_enabled() && out_mode == 0)
    return conv_s3(L, in, static_cast<__nv_bfloat16*>(out), B, H, W, st);
  if (g_halo_mode) {  // >= 128 output channels: CTA pairs (conv_tc2.cu) unless switched off
    bool handled = false;
    const int rc = conv_tc2(L, in, out, B, H, W, out_mode, st, &handled);
    if (rc != SPB_OK || handled) return rc;
  }
  const bool halo = g_halo_mode != 0;
#define SPB_CONV(CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT) \
  launch<ConvCfg<CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT>>(L, in, out, B, H, W, out_mode, st, aux)
#define SPB_CONV2(CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT) \
  launch<ConvCfg<CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT, 2>>(L, in, out, B, H, W, out_mode, st, aux)
  if (L.ks == 3 && L.cin == 64 && L.cout_pad == 64) {
    if (L.pool) return halo ? SPB_CONV(64, 64, 3, true, true, 4, 0, true, 0) : SPB_CONV(64, 64, 3, false, true, 6, 0, true, 0);
    return halo ? SPB_CONV(64, 64, 3, true, true, 4, 0, false, 0) : SPB_CONV(64, 64, 3, false, true, 6, 0, false, 0);
  }
  if (L.ks == 3 && L.cin == 64 && L.cout_pad == 128)
    return halo ? SPB_CONV(64, 128, 3, true, true, 2, 0, false, 0) : SPB_CONV(64, 128, 3, false, true, 3, 0, false, 0);
  if (L.ks == 3 && L.cin == 128 && L.cout_pad == 128) {
    // tile pairs: every streamed weight block feeds two tiles (g_pair == 0: one tile per stage, for A/B tests)
    if (halo && g_pair) return L.pool ? SPB_CONV2(128, 128, 3, true, false, 4, 6, true, 0) : SPB_CONV2(128, 128, 3, true, false, 4, 6, false, 0);
    if (L.pool) return halo ? SPB_CONV(128, 128, 3, true, false, 4, 6, true, 0) : SPB_CONV(128, 128, 3, false, false, 6, 6, true, 0);
    return halo ? SPB_CONV(128, 128, 3, true, false, 4, 6, false, 0) : SPB_CONV(128, 128, 3, false, false, 6, 6, false, 0);
  }
  if (L.ks == 3 && L.cin == 128 && L.cout_pad == 256)
    return halo ? SPB_CONV(128, 256, 3, true, false, 2, 4, false, 0) : SPB_CONV(128, 256, 3, false, false, 3, 4, false, 0);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 80 && out_mode == 1) return SPB_CONV(256, 80, 1, false, true, 6, 0, false, 1);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 80 && out_mode == 4) return SPB_CONV(256, 80, 1, false, true, 6, 0, false, 4);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 256 && out_mode == 2) return SPB_CONV(256, 256, 1, false, true, 4, 0, false, 2);
#undef SPB_CONV
#undef SPB_CONV2
  set_error("conv_tc: unsupported layer shape ks=%d cin=%d cout_pad=%d", L.ks, L.cin, L.cout_pad);
  return SPB_EINVAL;
}

}  // namespace spb
