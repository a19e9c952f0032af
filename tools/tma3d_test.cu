// Stand-alone check (not product code) of the fp32 3-D TMA box load used for conv1's image patch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int n, int bytes) {
  __shared__ __align__(1024) float patch[512];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(patch)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(n) : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok && clock64() - t0 < 200000000LL)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  if (threadIdx.x == 0 && !ok) printf("TIMEOUT\n");
  for (int i = threadIdx.x; i < 320; i += blockDim.x) out[i] = patch[i];
}
int main(int argc, char** argv) {
  const int BOX0 = argc > 1 ? atoi(argv[1]) : 12, PROMO = argc > 2 ? atoi(argv[2]) : 1, X0 = argc > 3 ? atoi(argv[3]) : -2;
  const int V = 3, H = 32, W = 32;
  std::vector<float> h(V * H * W);
  for (int i = 0; i < V * H * W; ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 320 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)V};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)BOX0, 20, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, PROMO ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  for (int tc = 0; tc < 2; ++tc) {
    const int x = tc ? 20 : X0, y = tc ? 15 : -2, n = 1;
    k<<<1, 256>>>(tm, o, x, y, n, BOX0 * 80);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case %d: %s\n", tc, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    float ho[320];
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < 20 && BOX0 * 20 <= 320; ++r2)
      for (int c = 0; c < BOX0; ++c) {
        const int yy = y + r2, xx = x + c;
        const float want = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)((n * H + yy) * W + xx) : 0.f;
        if (ho[r2 * BOX0 + c] != want) ++bad;
      }
    printf("  mismatches: %d (patch[0]=%g patch[26]=%g)\n", bad, ho[0], ho[26]);
  }
  return 0;
}
