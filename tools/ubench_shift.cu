// Micro-benchmark (not product code): semantics and cost of tcgen05.shift.down on sm_100a -- can TMEM rows be
// re-aligned for free instead of with warp shuffles (which cost one shared-memory wavefront each)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_shift.bin tools/ubench_shift.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) shift_test(uint32_t* out, long long* cyc, int nshift) {
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  // columns 0..15 of every lane: value = lane * 100 + column
  for (int c = 0; c < 16; ++c) {
    const uint32_t v = threadIdx.x * 100 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < nshift; ++i)
        asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem) : "memory");  // columns 0..7, all lanes
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (lane == 0) cyc[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 16; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[threadIdx.x * 16 + c] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  uint32_t* d;
  long long* c;
  cudaMalloc(&d, 128 * 16 * 4);
  cudaMalloc(&c, 8);
  for (int ns : {1, 2, 64, 1024}) {
    shift_test<<<1, 128>>>(d, c, ns);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    uint32_t h[128 * 16];
    long long hc;
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    printf("nshift=%d: %lld cycles total, %.1f per shift\n", ns, hc, (double)hc / ns);
    if (ns <= 2) {
      for (int l : {0, 1, 2, 15, 16, 31, 32, 33, 63, 64, 126, 127})
        printf("  lane %3d: col0=%u col7=%u col8=%u col15=%u\n", l, h[l * 16], h[l * 16 + 7], h[l * 16 + 8], h[l * 16 + 15]);
    }
  }
  return 0;
}
