// (2) The >= 128-output-channel 3x3 layers on CTA PAIRS: tcgen05.mma.cta_group::2 (M = 256 over two SMs).
// Same implicit GEMM as conv_tc.cu (one TMA halo block per tile and 64-channel chunk, 9 shifted SWIZZLE_128B
// descriptors, weights resident or streamed, two alternating MMA issuer warps, staged TMA-store epilogue), but the
// two CTAs of a cluster process their tiles in lock-step with ONE MMA stream issued by the leader CTA:
//   * each CTA supplies its own 128 rows of A and holds only HALF of every weight block (N/2 rows);
//   * per M128-equivalent MMA an SM now reads 4 KB of A + N*16 B of B instead of N*32 B: for N = 128 that is 48
//     shared-memory wavefronts per 64-cycle MMA instead of 64 -- the single-CTA kernel saturates the operand-read
//     pipe with the MMAs alone, so every other shared-memory byte (streamed weight blocks, halo blocks, epilogue
//     staging) added time (ncu r01: 69 % tensor-active for the 128 -> 128 layers, matching the wavefront sum);
//   * the streamed weight traffic from L2 and its shared-memory writes halve as well.
// Protocol: barriers the MMA completes (a_empty, b_empty, acc_full) are signalled in BOTH CTAs by multicast
// tcgen05.commit; barriers the MMA waits for have a "peer" twin in the leader that the follower CTA's two otherwise
// idle issuer warps forward to with remote mbarrier arrives (mapa): warp 1 relays a_full / b_full in consumption
// order, warp 10 relays acc_empty.  Producer, epilogue and TMA stores are CTA-local.
// Warps (both CTAs): 0 TMA producer, 1 and 10 MMA issuers (leader) / relays (follower), 2..9 epilogue.
#include "tc_common.cuh"

namespace spb {

template <int CIN_, int NPAD_, bool BRES_, int NA_, int NB_, bool POOL_, int NT_>
struct Cfg2 {
  static constexpr int CIN = CIN_, NPAD = NPAD_, NA = NA_, NT = NT_;
  static constexpr bool BRES = BRES_, POOL = POOL_;
  static constexpr int CHUNKS = CIN / 64, TAPS = 9, P = 10, R = 18, TH = 16, TW = 8;
  static constexpr int A_BYTES = P * R * 128, A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int BH_ROWS = NPAD / 2, B_BYTES = BH_ROWS * 128;  // this CTA's half of a weight block
  static constexpr int NBLK = TAPS * CHUNKS, NBS = BRES ? NBLK : NB_;
  static constexpr int TILE_COLS = NPAD, ACC_STRIDE = NT * TILE_COLS, TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int OSTAGE_BYTES = 8 * 4096;
  static constexpr int NBAR = 3 * NA + 3 * NBS + 8;
  static constexpr int THREADS = 352;
  static constexpr size_t SMEM = 1024 + (size_t)NA * A_STRIDE + (size_t)NBS * B_BYTES + OSTAGE_BYTES + NPAD * 4 + NBAR * 8 + 16;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  static_assert(B_BYTES % 1024 == 0 && SMEM <= 227 * 1024, "shared memory");
  static_assert(NA + NBS <= 32, "one relay lane per ring slot");
};

struct Conv2Args {
  const float* bias;
  int B, H, W, tiles_x, tiles_y, ntiles, nst;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_act,
                                                                 const __grid_constant__ CUtensorMap tm_wgt,
                                                                 const __grid_constant__ CUtensorMap tm_out, Conv2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)C::NA * C::A_STRIDE;
  uint8_t* sOut = sB + (size_t)C::NBS * C::B_BYTES;
  float* sBias = reinterpret_cast<float*>(sOut + C::OSTAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + C::NPAD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (C::NA + s); };
  auto peer_a = [&](int s) { return bar0 + 8u * (2 * C::NA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (3 * C::NA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (3 * C::NA + C::NBS + s); };
  auto peer_b = [&](int s) { return bar0 + 8u * (3 * C::NA + 2 * C::NBS + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (3 * C::NA + 3 * C::NBS + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (3 * C::NA + 3 * C::NBS + 2 + s); };
  auto peer_acc = [&](int s) { return bar0 + 8u * (3 * C::NA + 3 * C::NBS + 4 + s); };
  auto token = [&](int s) { return bar0 + 8u * (3 * C::NA + 3 * C::NBS + 6 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tpv = a.tiles_x * a.tiles_y;

  for (int i = threadIdx.x; i < C::NPAD; i += blockDim.x) sBias[i] = a.bias[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(peer_a(s), 1);
    }
    for (int s = 0; s < C::NBS; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
      mbar_init(peer_b(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 8);
      mbar_init(peer_acc(s), 1);
      mbar_init(token(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // super-tile j of this CTA: index 2*(cluster + j*nclusters) + rank; both CTAs run while the LEADER's index < nst
  auto lead_index = [&](int j) { return 2 * (cluster + j * nclusters); };
  auto advance = [](int& s, int& p, int n, int ring) {
    s += n;
    while (s >= ring) { s -= ring; p ^= 1; }
  };

  if (warp == 0) {
    // =========================== TMA producer (both CTAs: own halo blocks, own half of the weights) ==========
    if (lane == 0) {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      for (int j = 0; lead_index(j) < a.nst; ++j) {
        const int t0 = (lead_index(j) + (int)rank) * C::NT;
        for (int kc = 0; kc < C::CHUNKS; ++kc) {
#pragma unroll
          for (int u = 0; u < C::NT; ++u) {
            const int tu = t0 + u, nu = tu < a.ntiles ? tu / tpv : a.B, ru = tu < a.ntiles ? tu % tpv : 0;
            mbar_wait(a_empty(sa), pa ^ 1);
            if (rank == 0) mbar_expect_tx(a_full(sa), 2 * C::A_BYTES);  // both CTAs' blocks complete on the leader
            tma2_load_4d(smem_u32(sA + (size_t)sa * C::A_STRIDE), &tm_act, mapa_rank0(a_full(sa)), kc * 64,
                         (ru % a.tiles_x) * C::TW - 1, (ru / a.tiles_x) * C::TH - 1, nu);
            if (++sa == C::NA) { sa = 0; pa ^= 1; }
          }
          for (int tap = 0; tap < C::TAPS; ++tap) {
            if (C::BRES) {
              if (j == 0) {
                const int blk = kc * C::TAPS + tap;
                if (rank == 0) mbar_expect_tx(b_full(blk), 2 * C::B_BYTES);
                tma2_load_2d(smem_u32(sB + (size_t)blk * C::B_BYTES), &tm_wgt, mapa_rank0(b_full(blk)),
                             tap * C::CIN + kc * 64, (int)rank * C::BH_ROWS);
              }
            } else {
              mbar_wait(b_empty(sb), pb ^ 1);
              if (rank == 0) mbar_expect_tx(b_full(sb), 2 * C::B_BYTES);
              tma2_load_2d(smem_u32(sB + (size_t)sb * C::B_BYTES), &tm_wgt, mapa_rank0(b_full(sb)), tap * C::CIN + kc * 64,
                           (int)rank * C::BH_ROWS);
              if (++sb == C::NBS) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if ((warp == 1 || warp == 10) && rank == 0) {
    // =========================== MMA issuers (leader CTA): issuer q owns the super-tiles of parity q ==========
    const int q = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = make_idesc(256, C::NPAD);
    constexpr uint64_t desc_hi_a = ((uint64_t)((C::P * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    constexpr uint64_t desc_hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    constexpr int A_PER_ST = C::CHUNKS * C::NT;
    const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
    int sa = 0, pa = 0, sb = 0, pb = 0;
    if (q == 1) {
      advance(sa, pa, A_PER_ST, C::NA);
      if (!C::BRES) advance(sb, pb, C::NBLK, C::NBS);
    }
    const uint32_t d_tmem = tmem_base + (uint32_t)q * C::ACC_STRIDE;
    int k = 0;
    for (int j = q; lead_index(j) < a.nst; j += 2, ++k) {
      const bool first = k == 0;
      mbar_wait(acc_empty(q), (k & 1) ^ 1);
      mbar_wait(peer_acc(q), (k & 1) ^ 1);
      mbar_wait(a_full(sa), pa);
      if (q == 1 || k > 0) mbar_wait(token(q), (q == 0 ? k - 1 : k) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int kc = 0; kc < C::CHUNKS; ++kc) {
        int sau[C::NT];
        {
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
          const int slot = C::BRES ? kc * C::TAPS + tap : sb;
          if (C::BRES) {
            if (first) mbar_wait(b_full(slot), 0);
          } else {
            mbar_wait(b_full(slot), pb);
          }
          tc_fence_after();
          const uint32_t tap_off = ((tap / 3) * C::P + (tap % 3)) * 128;
          const uint64_t bdesc = desc_hi_b | (uint64_t)(((sB_u32 + (uint32_t)slot * C::B_BYTES) >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < C::NT; ++u) {
              const uint64_t adesc =
                  desc_hi_a | (uint64_t)(((sA_u32 + (uint32_t)sau[u] * C::A_STRIDE + tap_off) >> 4) | (1u << 16));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma2(d_tmem + u * C::TILE_COLS, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                        (kc | tap | kk) != 0 ? 1u : 0u);
            }
            if (!C::BRES) tc_commit2(b_empty(sb));
            if (tap == C::TAPS - 1) {
#pragma unroll
              for (int u = 0; u < C::NT; ++u) tc_commit2(a_empty(sau[u]));
            }
            if (kc == C::CHUNKS - 1 && tap == C::TAPS - 1) tc_commit2(acc_full(q));
            if (kc == C::CHUNKS - 1 && tap == C::TAPS - 3) mbar_arrive(token(q ^ 1));
          }
          __syncwarp();
          if (!C::BRES) {
            if (++sb == C::NBS) { sb = 0; pb ^= 1; }
          }
        }
        advance(sa, pa, C::NT, C::NA);
      }
      advance(sa, pa, A_PER_ST, C::NA);
      if (!C::BRES) advance(sb, pb, C::NBLK, C::NBS);
    }
  } else if (warp == 1 || warp == 10) {
    // =========================== follower CTA: forward acc_empty to the leader (one lane per accumulator stage) ======
    // (Operand barriers need no relay: both CTAs' TMA loads complete on the leader's mbarrier.  A software relay of
    // a_full / b_full was measured first: ~1500 cycles per stage for the remote release-arrive, 2x slower layers.)
    const int steps = 2 * cluster < a.nst ? (a.nst - 2 * cluster + 2 * nclusters - 1) / (2 * nclusters) : 0;
    uint32_t local_bar = 0, peer_bar = 0;
    int uses = 0;
    if (warp == 10 && lane < 2) {
      local_bar = acc_empty(lane);
      peer_bar = peer_acc(lane);
      uses = (steps + 1 - lane) / 2;
    }
    for (int n = 0; n < uses; ++n) {
      mbar_wait(local_bar, n & 1);
      mbar_arrive_remote(peer_bar, 0);
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 2..9, both CTAs) ===========================
    const int quad = warp & 3;
    constexpr int COLS = C::NPAD / 2, NBATCH = COLS / 32;
    const int col0 = ((warp - 2) >> 2) * COLS;
    uint8_t* stg = sOut + (warp - 2) * 4096;
    const int p = C::POOL ? (lane >> 4) * 4 + ((lane & 7) >> 1) : lane;
    for (int j = 0; lead_index(j) < a.nst; ++j) {
      const int as = j & 1, pacc = (j >> 1) & 1;
      mbar_wait(acc_full(as), pacc);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < C::NT; ++u) {
        const int t = (lead_index(j) + (int)rank) * C::NT + u;
        const int n = t < a.ntiles ? t / tpv : a.B, rem = t < a.ntiles ? t % tpv : 0;
        const int y0 = (rem / a.tiles_x) * C::TH, x0 = (rem % a.tiles_x) * C::TW;
        const uint32_t taddr =
            tmem_base + (uint32_t)as * C::ACC_STRIDE + (uint32_t)u * C::TILE_COLS + ((uint32_t)(quad * 32) << 16);
        uint32_t v[2][32];
        tc_ld_nowait<32>(taddr + col0, v[0]);
#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
          tc_wait_ld();
          if (b + 1 < NBATCH) {
            tc_ld_nowait<32>(taddr + col0 + (b + 1) * 32, v[(b + 1) & 1]);
          } else if (u == C::NT - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(as));
          }
          const int c0 = col0 + b * 32;
          uint32_t pk[16];
          epi_pack32<C::POOL>(v[b & 1], sBias + c0, pk);
          if ((b & 1) == 0) {
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
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees tensor memory) while the pair's MMAs / remote arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_cg2 = 1;  // 0: single-CTA kernels (conv_tc.cu) for these layers (A/B tests)
extern "C" int spb_conv_tc_set_cg2(int on) {  // 0 off, 1 the layers where the pair wins, 2 every supported layer
  g_cg2 = on;
  return SPB_OK;
}
bool conv_tc2_enabled() { return g_cg2 != 0; }

template <class C>
static int launch2(const LayerDev& L, const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv_tc2: cuTensorMapEncodeTiled not available from the driver");
    return SPB_ECUDA;
  }
  CUtensorMap tm_act, tm_wgt, tm_out;
  int rc = encode_act_map(&tm_act, in, B, H, W, L.cin, C::P, C::R);
  if (rc != SPB_OK) return rc;
  rc = C::POOL ? encode_act_map(&tm_out, out, B, H / 2, W / 2, L.cout, 4, 2) : encode_act_map(&tm_out, out, B, H, W, L.cout, 8, 4);
  if (rc != SPB_OK) return rc;
  {  // weights [cout_pad][9*cin] K-major: one box = HALF of the rows of one (tap, chunk) block
    const cuuint64_t K = (cuuint64_t)9 * L.cin;
    cuuint64_t dims[2] = {K, (cuuint64_t)L.cout_pad};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)C::BH_ROWS};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(&tm_wgt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w_tc, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tc2: cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
      return SPB_ECUDA;
    }
  }
  Conv2Args a;
  a.bias = L.bias;
  a.B = B;
  a.H = H;
  a.W = W;
  a.tiles_x = ceil_div(W, C::TW);
  a.tiles_y = ceil_div(H, C::TH);
  a.ntiles = B * a.tiles_x * a.tiles_y;
  a.nst = (a.ntiles + C::NT - 1) / C::NT;
  SPB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  int grid = 2 * ((a.nst + 1) / 2);
  if (grid > kNumSMs) grid = kNumSMs & ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<C>, tm_act, tm_wgt, tm_out, a));
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

// returns SPB_OK and sets *handled when the layer has a CTA-pair kernel
int conv_tc2(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode, cudaStream_t st,
             bool* handled) {
  *handled = false;
  if (!g_cg2 || out_mode != 0 || L.ks != 3 || !L.relu) return SPB_OK;
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  *handled = true;
  // (64 -> 128, weights resident: measured 0.407 ms per 800 views on one CTA, 0.427 on a CTA pair, 0.409 on a CTA
  // pair with tile pairs -- the layer is not operand-read bound; kept instantiated for the tests behind g_cg2 == 2)
  if (L.cin == 64 && L.cout_pad == 128 && !L.pool && g_cg2 == 2)
    return launch2<Cfg2<64, 128, true, 3, 0, false, 1>>(L, in, o, B, H, W, st);
  if (L.cin == 128 && L.cout_pad == 128)
    return L.pool ? launch2<Cfg2<128, 128, false, 4, 8, true, 2>>(L, in, o, B, H, W, st)
                  : launch2<Cfg2<128, 128, false, 4, 8, false, 2>>(L, in, o, B, H, W, st);
  if (L.cin == 128 && L.cout_pad == 256 && !L.pool) return launch2<Cfg2<128, 256, false, 2, 6, false, 1>>(L, in, o, B, H, W, st);
  *handled = false;
  return SPB_OK;
}

}  // namespace spb
