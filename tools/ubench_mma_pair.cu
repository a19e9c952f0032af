// Micro-benchmark (not product code; written at the end of round 1, NOT yet run on a GPU): the MMA sequence of
// conv1_n128_kernel -- per tap row and K step one M128 x N128 x K16 MMA followed by one M128 x N64 x K16 MMA that
// accumulates into the first 64 of the columns the N = 128 MMA just wrote, its A view shifted by two pixels (256 B)
// inside the 1024-byte SWIZZLE_128B atoms.  The kernel's probe showed 1825 cycles for the 2 + 24 MMAs of a tile
// instead of 96 + 1344; the variants below separate the possible causes:
//   V0  the kernel's sequence                          (N128 -> cols 0..127, N64 shifted A -> cols 0..63)
//   V1  N64 with the UNSHIFTED A view                  (isolates the 256-byte shift)
//   V2  N64 into its own columns 128..191              (isolates the accumulator dependency between the two shapes)
//   V3  both                                           (neither the shift nor the dependency)
//   V4  24 N128 MMAs only / V5  24 N64 MMAs only       (the 64 / 48-cycle references)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma_pair.bin tools/ubench_mma_pair.cu && tools/ubench_mma_pair.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
               "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int V>
__global__ void __launch_bounds__(64, 1) bench(long long* res, int tiles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;          // 24 KB: 10 x 16 halo pixels x 128 B (+ overrun)
  uint8_t* sB = smem + 24576;  // 3 x 192 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (24576 + 73728) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    constexpr uint32_t idesc64 = make_idesc(128, 64), idesc128 = make_idesc(128, 128);
    constexpr uint64_t hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (t & 1) * 256;  // two accumulator stages, as in the kernel
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const uint32_t arow = a0 + r * 2048;
        const uint64_t ad_a = hi | (uint64_t)((arow >> 4) | (1u << 16));
        const uint64_t ad_b = hi | (uint64_t)(((arow + ((V == 1 || V == 3) ? 0 : 256)) >> 4) | (1u << 16));
        const uint64_t bd_a = hi | (uint64_t)(((b0 + r * 24576) >> 4) | (1u << 16));
        const uint64_t bd_b = hi | (uint64_t)(((b0 + r * 24576 + 16384) >> 4) | (1u << 16));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (V == 4) {
              tc_mma_ss(d, ad_a + 2 * kk, bd_a + 2 * kk, idesc128, (r | kk) != 0);
              tc_mma_ss(d + 128, ad_a + 2 * kk, bd_a + 2 * kk, idesc128, (r | kk) != 0);
            } else if (V == 5) {
              tc_mma_ss(d, ad_b + 2 * kk, bd_b + 2 * kk, idesc64, (r | kk) != 0);
              tc_mma_ss(d + 64, ad_b + 2 * kk, bd_b + 2 * kk, idesc64, (r | kk) != 0);
            } else {
              tc_mma_ss(d, ad_a + 2 * kk, bd_a + 2 * kk, idesc128, (r | kk) != 0);
              if (V >= 2)
                tc_mma_ss(d + 128, ad_b + 2 * kk, bd_b + 2 * kk, idesc64, (r | kk) != 0);
              else
                tc_mma_ss(d, ad_b + 2 * kk, bd_b + 2 * kk, idesc64, 1u);
            }
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (lane == 0) res[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int V>
static void run(const char* name, int ideal, int tiles, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  const int smem = 1024 + 24576 + 73728;
  cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<V><<<grid, 64, smem>>>(d, tiles);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < grid; ++i) m += (double)h[i];
  printf("%-52s grid=%3d: %7.1f cycles per tile (ideal %d)\n", name, grid, m / grid / tiles, ideal);
  cudaFree(d);
}

int main() {
  const int tiles = 2000;
  for (int grid : {1, 148}) {
    run<0>("V0 N128 + N64(shifted A) into the same columns", 1344, tiles, grid);
    run<1>("V1 N128 + N64(unshifted A) into the same columns", 1344, tiles, grid);
    run<2>("V2 N128 + N64(shifted A) into its own columns", 1344, tiles, grid);
    run<3>("V3 N128 + N64(unshifted A) into its own columns", 1344, tiles, grid);
    run<4>("V4 24 x N128", 1536, tiles, grid);
    run<5>("V5 24 x N64", 1152, tiles, grid);
  }
  return 0;
}
