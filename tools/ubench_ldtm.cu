// Micro-benchmark (not product code): tcgen05.ld (TMEM -> registers) throughput per SM for the 32x32b shape,
// x16 / x32 / x64 columns per instruction, 4 / 8 / 16 warps, 4 loads in flight per warp before each wait.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) a ^= v[i];
  return a;
}
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i];
  return a;
}
template <int X, int INFLIGHT>
__global__ void __launch_bounds__(544, 1) k(long long* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp >= 1 && warp <= nwarps) {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t r[INFLIGHT];
#pragma unroll
      for (int j = 0; j < INFLIGHT; ++j) r[j] = ld<X>(taddr + ((it * INFLIGHT + j) * X) % 448);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < INFLIGHT; ++j) acc ^= r[j];
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) out[1] = 1;
    if (warp == 1 && lane == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
template <int X, int F>
void run(int nw) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int iters = 2000;
  k<X, F><<<1, 544>>>(d, iters, nw);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("x%-2d inflight %d warps %2d: %s  %.1f cyc per load per warp -> %.0f B/clk/SM\n", X, F, nw, cudaGetErrorString(e),
         (double)h / (iters * F), (double)nw * iters * F * X * 128.0 / h);
  cudaFree(d);
}
int main() {
  for (int nw : {4, 8, 16}) {
    run<16, 3>(nw);
    run<16, 6>(nw);
    run<32, 1>(nw);
    run<32, 2>(nw);
    run<32, 4>(nw);
  }
  return 0;
}
