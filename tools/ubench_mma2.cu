// Micro-benchmark (not product code): tcgen05.mma with cta_group::2 (M = 256 across a CTA pair, each CTA supplies its
// 128 rows of A and HALF of B) -- does sharing B lift the N = 64 shared-memory operand-read ceiling (48 -> 40 cycles)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma2.bin tools/ubench_mma2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (clock64() - t0 > 2000000000LL) {
      printf("timeout block %d\n", blockIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;          // 32 KB halo block (this CTA's 128 rows)
  uint8_t* sB = smem + 32768;  // N/2 rows x 128 B (this CTA's half of B)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (32768 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 1) {
    long long t0 = 0;
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(256, N);
      const uint64_t hi_a = ((uint64_t)((10 * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (N <= 128 ? (it & 1) * N : 0);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t off = ((tap / 3) * 10 + tap % 3) * 128;
          const uint64_t ad = hi_a | (uint64_t)(((a0 + off) >> 4) | (1u << 16));
          const uint64_t bd = hi_b | (uint64_t)((b0 >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accf = (tap | k) != 0;
              asm volatile(
                  "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                  "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(accf)
                  : "memory");
            }
          }
          __syncwarp();
        }
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    if (rank == 0 && lane == 0) out[blockIdx.x / 2] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N>
static void run(int grid) {
  long long* d;
  cudaMalloc(&d, 8 * 74);
  cudaMemset(d, 0, 8 * 74);
  const int smem = 1024 + 32768 + 16384, iters = 2000;
  cudaFuncSetAttribute(bench2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench2<N><<<grid, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d grid=%d: CUDA error %s\n", N, grid, cudaGetErrorString(e));
    exit(1);
  }
  long long h[74];
  cudaMemcpy(h, d, 8 * (grid / 2), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < grid / 2; ++i) m += (double)h[i];
  m /= grid / 2;
  printf("cta_group::2 M256 N=%3d grid=%3d: %6.1f cycles per MMA (per SM: 128 x %d x 16; ideal %d, cta_group::1 SS: %d)\n", N,
         grid, m / (iters * 36.0), N, N / 2, N == 64 ? 48 : N / 2);
  cudaFree(d);
}

int main() {
  for (int grid : {2, 148}) {
    run<64>(grid);
    run<128>(grid);
    run<192>(grid);
    run<256>(grid);
  }
  return 0;
}
