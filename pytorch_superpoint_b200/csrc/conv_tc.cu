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
#include <cuda.h>

#include "net.cuh"

namespace spb {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must abort the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("spb200 conv_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BW>
__device__ __forceinline__ void tc_ld_nowait(uint32_t taddr, uint32_t (&v)[BW]);
template <>
__device__ __forceinline__ void tc_ld_nowait<16>(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tc_ld_nowait<32>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 fp32 accumulator columns of one pixel -> + bias -> bf16x2 pack -> ReLU -> optional 2x2 max-pool with PACKED
// lane shuffles (lane = 4 tile rows x 8 tile columns: partners are lane^1 and lane^8; max commutes with the
// monotone bf16 rounding, so pooling after the pack is exact) -> one 64-byte store by the lane that owns the output.
template <bool POOL>
__device__ __forceinline__ void epi_bf16_32(const uint32_t (&v)[32], const float* __restrict__ bias, bool valid,
                                            __nv_bfloat16* __restrict__ o) {
  uint32_t pk[16];
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + j);
    __nv_bfloat162 h0 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y), zero);
    __nv_bfloat162 h1 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j + 2]) + bb.z, __uint_as_float(v[j + 3]) + bb.w), zero);
    uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
    if (POOL) {
#pragma unroll
      for (int sh = 1; sh <= 8; sh <<= 3) {
        const uint32_t p0 = __shfl_xor_sync(0xffffffffu, u0, sh), p1 = __shfl_xor_sync(0xffffffffu, u1, sh);
        h0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u0), *reinterpret_cast<const __nv_bfloat162*>(&p0));
        h1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u1), *reinterpret_cast<const __nv_bfloat162*>(&p1));
        u0 = *reinterpret_cast<const uint32_t*>(&h0);
        u1 = *reinterpret_cast<const uint32_t*>(&h1);
      }
    }
    pk[j / 2] = u0;
    pk[j / 2 + 1] = u1;
  }
  if (valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(o + q * 8) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------------
// OUT_: 0 = ReLU + bf16 NHWC (encoder / head 3x3), 1 = raw fp32 rows staged through shared memory for coalesced
// stores (convPb: 65 channels), 2 = fp32 rows L2-normalised over the channels (convDb).
template <int CIN_, int NPAD_, int KS_, bool HALO_, bool BRES_, int NA_, int NB_, bool POOL_, int OUT_>
struct ConvCfg {
  static constexpr int CIN = CIN_, NPAD = NPAD_, KS = KS_, NA = NA_, OUT = OUT_;
  static constexpr bool POOL = POOL_;
  static constexpr int BW = (NPAD_ % 32 == 0) ? 32 : 16;  // TMEM columns per epilogue batch
  static constexpr int EPI_WARPS = OUT_ == 0 ? 8 : 4;     // bf16 layers: two warps per TMEM lane quadrant, half the columns each
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;  // + the second MMA issuer (last warp)
  static constexpr int STAGE_BYTES = OUT_ == 1 ? 4 * 32 * 65 * 4 : 0;
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
  static constexpr int ACC_STRIDE = NPAD <= 64 ? 64 : (NPAD <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int NBAR = 2 * NA + 2 * NBS + 4 + 2;  // + the two issue tokens
  static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)NA * A_STRIDE + (size_t)NBS * B_BYTES + NPAD * 4 +
                                 NBAR * 8 + 16 + STAGE_BYTES;
  static_assert(B_BYTES % 1024 == 0, "weight block must keep 1024-byte alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct ConvArgs {
  const float* bias;
  void* out;
  int B, H, W, cout;
  int tiles_x, tiles_y, ntiles;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act,
                                                         const __grid_constant__ CUtensorMap tm_wgt, ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)C::NA * C::A_STRIDE;
  float* sBias = reinterpret_cast<float*>(sB + (size_t)C::NBS * C::B_BYTES);
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
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int n = t / (a.tiles_x * a.tiles_y), rem = t % (a.tiles_x * a.tiles_y);
        const int y0 = (rem / a.tiles_x) * C::TH, x0 = (rem % a.tiles_x) * C::TW;
        for (int kc = 0; kc < C::CHUNKS; ++kc) {
          if (C::HALO) {
            mbar_wait(a_empty(sa), pa ^ 1);
            mbar_expect_tx(a_full(sa), C::A_BYTES);
            tma_load_4d(smem_u32(sA + (size_t)sa * C::A_STRIDE), &tm_act, a_full(sa), kc * 64, x0 - C::KS / 2,
                        y0 - C::KS / 2, n);
            if (++sa == C::NA) { sa = 0; pa ^= 1; }
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
    constexpr int A_PER_TILE = C::HALO ? C::CHUNKS : C::CHUNKS * C::TAPS;
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
    for (int t = blockIdx.x + q * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++k) {
      const bool first = k == 0;
      mbar_wait(acc_empty(q), (k & 1) ^ 1);
      mbar_wait(a_full(sa), pa);  // first operand block of the tile (waited again, for free, in the loop)
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int kc = 0; kc < C::CHUNKS; ++kc) {
        if (C::HALO) mbar_wait(a_full(sa), pa);
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
          const uint64_t adesc = desc_hi_a | (uint64_t)(((sA_u32 + (uint32_t)sa * C::A_STRIDE + tap_off) >> 4) | (1u << 16));
          const uint64_t bdesc = desc_hi_b | (uint64_t)(((sB_u32 + (uint32_t)slot * C::B_BYTES) >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                     (kc | tap | kk) != 0 ? 1u : 0u);
            if (!C::BRES) tc_commit(b_empty(sb));
            if (!C::HALO) tc_commit(a_empty(sa));
            if (C::HALO && tap == C::TAPS - 1) tc_commit(a_empty(sa));
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
        if (C::HALO) {
          if (++sa == C::NA) { sa = 0; pa ^= 1; }
        }
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
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
      const int n = t / (a.tiles_x * a.tiles_y), rem = t % (a.tiles_x * a.tiles_y);
      const int y0 = (rem / a.tiles_x) * C::TH, x0 = (rem % a.tiles_x) * C::TW;
      const int y = y0 + py, x = x0 + px;
      mbar_wait(acc_full(as), pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)as * C::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
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
      if (C::OUT == 3) {
        // convPb in the HA path: 65-way softmax of the cell in registers (each thread holds its whole row),
        // dustbin dropped, probabilities written as [.,64] fp32 -- utils/utils.py:514-523 fused here.
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
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 65; ++j) {
          const float e = expf(__uint_as_float(w[j]) - mx);
          w[j] = __float_as_uint(e);
          sum += e;
        }
        if (valid) {
          float* o = static_cast<float*>(a.out) + opix * 64;
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(o + j) =
                make_float4(__fdiv_rn(__uint_as_float(w[j]), sum), __fdiv_rn(__uint_as_float(w[j + 1]), sum),
                            __fdiv_rn(__uint_as_float(w[j + 2]), sum), __fdiv_rn(__uint_as_float(w[j + 3]), sum));
        }
        if (++as == 2) { as = 0; pacc ^= 1; }
        continue;
      }
      uint32_t v[2][BW];
      tc_ld_nowait<BW>(taddr + col0, v[0]);
#pragma unroll
      for (int b = 0; b < NBATCH; ++b) {
        tc_wait_ld();
        if (b + 1 < NBATCH) {
          tc_ld_nowait<BW>(taddr + col0 + (b + 1) * BW, v[(b + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(as));
        }
        const int c0 = col0 + b * BW;
        if constexpr (C::OUT == 0) {
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
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
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
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// two fp32 accumulators -> ReLU -> packed bf16x2 (pack first, then one packed max)
__device__ __forceinline__ uint32_t relu_pack_bf16(uint32_t a, uint32_t b) {
  const __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b)),
                                   __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}

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
    if (threadIdx.x == 0) tma_store_wait_read<0>();
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
  uint8_t a1[2][8192];      // conv1a im2col: 256 rows x 32 B, no-swizzle core matrices
  uint8_t b1[2048];         // conv1a weights [64][16]
  float patch[2][20 * 12];  // zero-padded 20x12 fp32 image patch under the halo block (rows y0-2.., cols x0-2..)
  unsigned mrow[2][16];     // valid-mask bits of the tile's 16 rows x 8 pixels (homography-warp mode)
  float bias1[64], bias2[64];
  uint64_t bars[9 + 14 + 2];  // b_full[9], a1_full[2], acc1_full[2], acc1_empty[2], a2_full[2], a2_empty[2], acc2_full[2], acc2_empty[2], token[2]
  uint32_t tmem_slot;
};

struct Conv1FusedArgs {
  const float* x;               // [V,H,W] fp32 views -- or, in warp mode (mats != NULL), the SOURCE images [V/vpi,H,W]
  const float* mats;            // warp mode: forward homographies (view n uses mats[n] or mats[n % vpi] when shared)
  uint8_t* bits;                // warp mode: valid-mask bits out, [V,H,W/8]
  int vpi, shared_mats;
  Lin lx, ly;
  const __nv_bfloat16* w1a;     // [64][16]
  const float *bias1, *bias2;
  __nv_bfloat16* out;           // [V,H/2,W/2,64]
  int V, H, W, tiles_x, tiles_y, ntiles;
  long long* dbg;               // optional timeline probe (CTA 0): [64 iterations][16 slots] of clock64, or NULL
};

__global__ void __launch_bounds__(576, 1) conv1_fused_kernel(const __grid_constant__ CUtensorMap tm_w2, Conv1FusedArgs a) {
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

  if (threadIdx.x < 32) sm.mrow[threadIdx.x >> 4][threadIdx.x & 15] = 0u;
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
      }
      SPB_PROBE(4);
      if (t + 2 * (int)gridDim.x < a.ntiles) issue_conv1a(k + 1);
    }
  } else if (warp >= 10) {
    // =========================== stage 1: im2col build + conv1a epilogue into conv1b's A operand =========
    const int sid = threadIdx.x - 320;           // 0..255
    const int quad = warp & 3;                   // TMEM lane quadrant of this warp
    const int chalf = (warp - 10) >> 2;          // which 32 of the 64 conv1a channels this warp converts
    const int lrow = quad * 32 + lane;           // TMEM lane; the thread owns halo pixels lrow and lrow + 128
    // Element sid of the 20x12 patch of tile t (zero outside the image).  In warp mode the pixel of view n is
    // produced here from the source image -- the homography warp (utils/utils.py:318-351, datasets/Coco.py:279)
    // is fused into conv1's producer and the warped view never exists in memory; the sign bit of the returned
    // word carries the nearest-mode validity of the same coordinates (utils/utils.py:696) for the mask bits.
    struct PatchVal {
      float v;
      unsigned valid;
    };
    auto patch_load = [&](int t) -> PatchVal {
      PatchVal r{0.f, 0u};
      if (sid >= 240) return r;
      const int n = t / tiles_per_view, rem = t % tiles_per_view;
      const int yy = (rem / a.tiles_x) * 16 - 2 + sid / 12, xx = (rem % a.tiles_x) * 8 - 2 + sid % 12;
      if (!(yy >= 0 && yy < a.H && xx >= 0 && xx < a.W)) return r;
      if (a.mats == nullptr) {
        r.v = __ldg(a.x + ((long long)n * a.H + yy) * a.W + xx);
        return r;
      }
      const Mat3 M = load_mat(a.mats + 9 * (a.shared_mats ? n % a.vpi : n));
      const float* src = a.x + (long long)(n / a.vpi) * a.H * a.W;
      float ix, iy;
      warp_coord(M, a.lx.at(xx), a.ly.at(yy), a.W, a.H, ix, iy);
      const Taps tp = make_taps(ix, iy);
      r.v = bilinear(tp, a.W, a.H, [&](int sx, int sy) { return __ldg(src + (long long)sy * a.W + sx); });
      r.valid = in_bounds(rintf(ix), rintf(iy), a.W, a.H) ? 1u : 0u;
      return r;
    };
    auto build = [&](PatchVal pv, int t, int buf, int pit) {  // patch -> K=16 im2col rows of the 180 halo pixels of tile t
      auto probe = [&](int slot) {
        if (a.dbg && blockIdx.x == 0 && pit >= 0 && pit < 64 && threadIdx.x == 320) a.dbg[pit * 16 + slot] = clock64();
      };
      const int y0 = ((t % tiles_per_view) / a.tiles_x) * 16, x0 = ((t % tiles_per_view) % a.tiles_x) * 8;
      if (sid < 240) {
        sm.patch[buf][sid] = pv.v;
        const int pr = sid / 12 - 2, pc = sid % 12 - 2;  // position inside the 16x8 tile
        if (a.bits && pv.valid && pr >= 0 && pr < 16 && pc >= 0 && pc < 8) atomicOr(&sm.mrow[buf][pr], 1u << pc);
      }
      named_bar_sync(2, 256);
      probe(13);
      if (a.bits && sid < 16) {
        if (y0 + sid < a.H && x0 < a.W)
          a.bits[((long long)(t / tiles_per_view) * a.H + y0 + sid) * (a.W >> 3) + (x0 >> 3)] = (uint8_t)sm.mrow[buf][sid];
        sm.mrow[buf][sid] = 0u;
      }
      if (sid < 180) {
        const float* pp = sm.patch[buf] + (sid / 10) * 12 + sid % 10;  // tap (0,0) of halo pixel sid
        // conv1b zero-pads conv1a's OUTPUT: a halo pixel outside the image must come out as exactly 0, so its
        // whole im2col row is zeroed (its taps can reach inside the image), INCLUDING the two bias slots
        const int hy = y0 - 1 + sid / 10, hx = x0 - 1 + sid % 10;
        const bool inside = hy >= 0 && hy < a.H && hx >= 0 && hx < a.W;
        uint32_t pk[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int k0 = 2 * k, k1 = 2 * k + 1;
          const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[(k0 / 3) * 12 + k0 % 3], pp[(k1 / 3) * 12 + k1 % 3]);
          pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const __nv_bfloat162 hh = __floats2bfloat162_rn(pp[2 * 12 + 2], 1.f);  // slot 9: bias head
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
    PatchVal pv{0.f, 0u};
    if ((int)blockIdx.x < a.ntiles) {
      build(patch_load(blockIdx.x), blockIdx.x, 0, -1);
      if ((int)(blockIdx.x + gridDim.x) < a.ntiles) pv = patch_load(blockIdx.x + gridDim.x);
    }
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
#define SPB_PROBE1(slot)                                                                                 \
  if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 320) a.dbg[it * 16 + (slot)] = clock64()
      SPB_PROBE1(5);
      if (t + (int)gridDim.x < a.ntiles) {
        build(pv, t + gridDim.x, (it + 1) & 1, it);                                // im2col of tile it+1
        if (t + 2 * (int)gridDim.x < a.ntiles) pv = patch_load(t + 2 * gridDim.x);  // prefetch for tile it+2
      }
      SPB_PROBE1(6);
      const int buf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc1_full(buf), ph);
      SPB_PROBE1(7);
      mbar_wait(a2_empty(buf), ph ^ 1);
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
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    const int m = quad * 32 + lane, py = m >> 3, px = m & 7;
    int it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1, ph = (it >> 1) & 1;
      const int n = t / tiles_per_view, rem = t % tiles_per_view;
      const int y = (rem / a.tiles_x) * 16 + py, x = (rem % a.tiles_x) * 8 + px;
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 10] = clock64();
      mbar_wait(acc2_full(buf), ph);
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 11] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 64 + chalf * 32 + ((uint32_t)(quad * 32) << 16);
      const bool valid = y < a.H && x < a.W && !(px & 1) && !(py & 1);
      const long long opix = ((long long)n * (a.H >> 1) + (y >> 1)) * (a.W >> 1) + (x >> 1);
      uint32_t v[32];
      tc_ld_nowait<32>(taddr, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty(buf));
      epi_bf16_32<true>(v, sm.bias2 + chalf * 32, valid, a.out + opix * 64 + chalf * 32);
      if (a.dbg && blockIdx.x == 0 && it < 64 && threadIdx.x == 64) a.dbg[it * 16 + 12] = clock64();
    }
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

static int encode_act_map(CUtensorMap* tm, const void* in, int B, int H, int W, int C, int boxw, int boxh) {
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
static int g_fuse1 = 1;  // conv1a fused into conv1b's kernel (0: two kernels, for A/B tests)
extern "C" int spb_conv_tc_set_fuse1(int on) {
  g_fuse1 = on ? 1 : 0;
  return SPB_OK;
}
bool conv1_fused_enabled() { return g_fuse1 != 0; }
// Homography warp + valid-mask bits produced inside conv1's stage 1 instead of by warp_bits_kernel.  Bit-identical
// results, but the 4-tap gathers lengthen stage 1 (the critical producer of the fused block), so it is OFF by default.
static int g_fuse_warp = 0;
extern "C" int spb_conv_tc_set_fuse_warp(int on) {
  g_fuse_warp = on ? 1 : 0;
  return SPB_OK;
}
bool conv1_warp_fused_enabled() { return g_fuse1 != 0 && g_fuse_warp != 0; }

int conv1_fused(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
                const ViewSource* vs, cudaStream_t st) {
  if ((H | W) & 1) {
    set_error("conv1_fused: pooled block needs even H, W (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  CUtensorMap tm_w2;
  memcpy(&tm_w2, L1b.tmap_w, sizeof(tm_w2));
  Conv1FusedArgs a;
  a.x = x;
  a.mats = vs ? vs->mats : nullptr;
  a.bits = vs ? vs->bits : nullptr;
  a.vpi = vs ? vs->vpi : 1;
  a.shared_mats = vs ? vs->shared_mats : 0;
  a.lx = Lin(W);
  a.ly = Lin(H);
  if (vs && (W & 7)) {
    set_error("conv1_fused: warp mode needs W %% 8 == 0 (got %d)", W);
    return SPB_EINVAL;
  }
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
  const int smem = (int)sizeof(Conv1FusedSmem) + 1024;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv1_fused_kernel<<<grid, 576, smem, st>>>(tm_w2, a);
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

static int g_halo_mode = 1;  // 1: one halo block + shifted descriptors; 0: per-tap reload (debug/A-B)
extern "C" int spb_conv_tc_set_halo(int on) {
  g_halo_mode = on ? 1 : 0;
  return SPB_OK;
}

template <class C>
static int launch(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode,
                  cudaStream_t st) {
  CUtensorMap tm_act, tm_wgt;
  memcpy(&tm_wgt, L.tmap_w, sizeof(tm_wgt));
  int rc = encode_act_map(&tm_act, in, B, H, W, L.cin, C::P, C::R);
  if (rc != SPB_OK) return rc;
  ConvArgs a;
  a.bias = L.bias;
  a.out = out;
  a.B = B;
  a.H = H;
  a.W = W;
  a.cout = L.cout;
  if (L.pool != (C::POOL ? 1 : 0) || L.relu != (C::OUT == 0 ? 1 : 0) || out_mode != C::OUT ||
      ((C::OUT == 1 || C::OUT == 3) && L.cout != 65)) {
    set_error("conv_tc: kernel configuration does not match the layer (pool %d relu %d out_mode %d)", L.pool, L.relu,
              out_mode);
    return SPB_EINVAL;
  }
  a.tiles_x = ceil_div(W, C::TW);
  a.tiles_y = ceil_div(H, C::TH);
  a.ntiles = B * a.tiles_x * a.tiles_y;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  conv_tc_kernel<C><<<grid, C::THREADS, C::SMEM, st>>>(tm_act, tm_wgt, a);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int conv_tc(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode,
            cudaStream_t st) {
  if (!L.tmap_w) {
    set_error("conv_tc: layer has no weight tensor map");
    return SPB_EINVAL;
  }
  if (L.pool && ((H | W) & 1)) {
    set_error("conv_tc: pooled layer needs even H, W (got %dx%d)", H, W);
    return SPB_EINVAL;
  }
  const bool halo = g_halo_mode != 0;
#define SPB_CONV(CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT) \
  launch<ConvCfg<CIN, NPAD, KS, HALO, BRES, NA, NB, POOL, OUT>>(L, in, out, B, H, W, out_mode, st)
  if (L.ks == 3 && L.cin == 64 && L.cout_pad == 64) {
    if (L.pool) return halo ? SPB_CONV(64, 64, 3, true, true, 4, 0, true, 0) : SPB_CONV(64, 64, 3, false, true, 6, 0, true, 0);
    return halo ? SPB_CONV(64, 64, 3, true, true, 4, 0, false, 0) : SPB_CONV(64, 64, 3, false, true, 6, 0, false, 0);
  }
  if (L.ks == 3 && L.cin == 64 && L.cout_pad == 128)
    return halo ? SPB_CONV(64, 128, 3, true, true, 3, 0, false, 0) : SPB_CONV(64, 128, 3, false, true, 4, 0, false, 0);
  if (L.ks == 3 && L.cin == 128 && L.cout_pad == 128) {
    if (L.pool) return halo ? SPB_CONV(128, 128, 3, true, false, 4, 6, true, 0) : SPB_CONV(128, 128, 3, false, false, 6, 6, true, 0);
    return halo ? SPB_CONV(128, 128, 3, true, false, 4, 6, false, 0) : SPB_CONV(128, 128, 3, false, false, 6, 6, false, 0);
  }
  if (L.ks == 3 && L.cin == 128 && L.cout_pad == 256)
    return halo ? SPB_CONV(128, 256, 3, true, false, 3, 4, false, 0) : SPB_CONV(128, 256, 3, false, false, 4, 4, false, 0);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 80 && out_mode == 1) return SPB_CONV(256, 80, 1, false, true, 6, 0, false, 1);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 80 && out_mode == 3) return SPB_CONV(256, 80, 1, false, true, 6, 0, false, 3);
  if (L.ks == 1 && L.cin == 256 && L.cout_pad == 256 && out_mode == 2) return SPB_CONV(256, 256, 1, false, true, 4, 0, false, 2);
#undef SPB_CONV
  set_error("conv_tc: unsupported layer shape ks=%d cin=%d cout_pad=%d", L.ks, L.cin, L.cout_pad);
  return SPB_EINVAL;
}

}  // namespace spb
