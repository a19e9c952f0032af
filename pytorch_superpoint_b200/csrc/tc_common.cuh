// PTX wrappers and small device helpers shared by the tcgen05 / TMEM / TMA kernels (conv_tc.cu, conv_s3.cu).
#pragma once
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
#ifdef SPB_MBAR_HINT_NS  // A/B: let the hardware suspend the waiting thread for up to this many ns per try
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"((uint32_t)SPB_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    if (ok) return;
    // (a __nanosleep here was measured: 20 / 100 ns back-off costs 1.5 % of throughput, no power-cap relief)
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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
__device__ __forceinline__ void tc_ld_nowait<8>(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
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
__device__ __forceinline__ void epi_pack32(const uint32_t (&v)[32], const float* __restrict__ bias, uint32_t (&pk)[16]) {
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
}
// same with the 32 biases already in registers (an LDS is a wavefront of the pipe the tensor core reads through)
template <bool POOL>
__device__ __forceinline__ void epi_pack32_regs(const uint32_t (&v)[32], const float (&bz)[32], uint32_t (&pk)[16]) {
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    __nv_bfloat162 h0 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j]) + bz[j], __uint_as_float(v[j + 1]) + bz[j + 1]), zero);
    __nv_bfloat162 h1 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j + 2]) + bz[j + 2], __uint_as_float(v[j + 3]) + bz[j + 3]), zero);
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
}
template <bool POOL>
__device__ __forceinline__ void epi_bf16_32(const uint32_t (&v)[32], const float* __restrict__ bias, bool valid,
                                            __nv_bfloat16* __restrict__ o) {
  uint32_t pk[16];
  epi_pack32<POOL>(v, bias, pk);
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

__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// all of this thread's bulk stores have COMPLETED (written, not just read out of shared memory): the last thing a
// storing thread does before the kernel's final barrier
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// two fp32 accumulators -> ReLU -> packed bf16x2 (pack first, then one packed max)
__device__ __forceinline__ uint32_t relu_pack_bf16(uint32_t a, uint32_t b) {
  const __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b)),
                                   __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}


// ------------------------------------------------------------------------------------------------
// CTA pairs (tcgen05 cta_group::2): cluster barrier, remote mbarrier arrive, pair TMA loads, pair MMA / commit
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the completion bytes are signalled on the
// LEADER's mbarrier (cluster address), so one barrier phase covers both CTAs' halves without a software relay
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local_addr) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(0));
  return remote;
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// acquire at cluster scope: the barrier is arrived on by the peer CTA (mbar_arrive_remote)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("spb200 conv pair: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// host: driver entry point for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();  // conv_tc.cu
// bf16 NHWC activations [B,H,W,C] as a 4-D tensor map with a (64 ch, boxw px, boxh rows, 1) SWIZZLE_128B box
int encode_act_map(CUtensorMap* tm, const void* in, int B, int H, int W, int C, int boxw, int boxh);

// 4 fp32 channels of one pixel (bias added) -> bf16x2 -> ReLU -> optional 2x2 max-pool (partners lane^1, lane^16)
template <bool POOL>
__device__ __forceinline__ void s3_pack4(const float* f, uint32_t* pk) {
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
  __nv_bfloat162 h0 = __hmax2(__floats2bfloat162_rn(f[0], f[1]), zero);
  __nv_bfloat162 h1 = __hmax2(__floats2bfloat162_rn(f[2], f[3]), zero);
  uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
  if (POOL) {
#pragma unroll
    for (int sh = 1; sh <= 16; sh <<= 4) {
      const uint32_t p0 = __shfl_xor_sync(0xffffffffu, u0, sh), p1 = __shfl_xor_sync(0xffffffffu, u1, sh);
      h0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u0), *reinterpret_cast<const __nv_bfloat162*>(&p0));
      h1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u1), *reinterpret_cast<const __nv_bfloat162*>(&p1));
      u0 = *reinterpret_cast<const uint32_t*>(&h0);
      u1 = *reinterpret_cast<const uint32_t*>(&h1);
    }
  }
  pk[0] = u0;
  pk[1] = u1;
}
// 16 channels: four groups of 4
template <bool POOL>
__device__ __forceinline__ void s3_pack16(const float (&f)[16], uint32_t* pk) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) s3_pack4<POOL>(f + j, pk + j / 2);
}

}  // namespace spb
