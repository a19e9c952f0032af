// Micro-benchmark (not product code): operand re-use of tcgen05.mma for the Cout = 64 layers.
// An M128 x N64 x K16 MMA reads 4 KB of A + 2 KB of B from shared memory for 32 cycles of tensor work; the SM
// delivers 128 B/clk, so the plain form runs at 48 cycles (tools/ubench_mma.cu).  PTX has two ways of keeping an
// operand in the tensor core's collector between MMAs:
//   tcgen05.mma.ws ... .collector::bN::fill / ::use     weight-stationary: B stays, A streams
//   tcgen05.mma    ... .collector::a::fill / ::use      A stays, B streams
// Variants (36 MMAs per "tile", like one conv1b tile):
//   W0  plain SS                                              (48 cycles per MMA reference)
//   W1  .ws, B filled every MMA (no reuse)                    (cost of the .ws form itself)
//   W2  .ws, B fill + 1 use  (two tiles share every weight block)
//   W3  .ws, B fill + 3 uses (four tiles share every weight block)
//   W4  A fill + 2 uses with three different B blocks         (one A window, three column taps)
// plus a bit-exact comparison of the accumulators of W2 / W4 against W0 on pseudo-random operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma_ws.bin tools/ubench_mma_ws.cu
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
#define MMA_FN(name, text)                                                                                         \
  __device__ __forceinline__ void name(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {       \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n" text " [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(ad),    \
                 "l"(bd), "r"(idesc), "r"(acc)                                                                     \
                 : "memory");                                                                                      \
  }
MMA_FN(mma_ss, "tcgen05.mma.cta_group::1.kind::f16")
MMA_FN(mma_a_fill, "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill")
MMA_FN(mma_a_use, "tcgen05.mma.cta_group::1.kind::f16.collector::a::use")
MMA_FN(mma_a_last, "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse")
MMA_FN(mma_ws, "tcgen05.mma.ws.cta_group::1.kind::f16")
MMA_FN(mma_ws_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill")
MMA_FN(mma_ws_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use")
MMA_FN(mma_ws_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse")

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// smem: A tiles 4 x 16 KB (128 rows x 128 B, SWIZZLE_128B atoms), B blocks 9 x 8 KB (64 rows x 128 B)
constexpr int kABytes = 4 * 16384, kBBytes = 9 * 8192;

template <int V>
__global__ void __launch_bounds__(128, 1) bench(long long* res, uint32_t* dump, int tiles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x) {
    // small integers as bf16 pairs: every product and sum is exact in fp32, so any two orders agree bit for bit
    const uint32_t h = (uint32_t)i * 2654435761u;
    const int lo = (int)((h >> 8) % 7) - 3, hi = (int)((h >> 16) % 7) - 3;
    auto bf = [](int v) -> uint32_t { return (__float_as_uint((float)v) >> 16) & 0xffffu; };
    reinterpret_cast<uint32_t*>(smem)[i] = bf(lo) | (bf(hi) << 16);
  }
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
    constexpr uint32_t idesc = make_idesc(128, 64);
    constexpr uint64_t hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    auto adesc = [&](int tile, int kk) { return (hi | (uint64_t)(((a0 + tile * 16384) >> 4) | (1u << 16))) + 2 * kk; };
    auto bdesc = [&](int tap, int kk) { return (hi | (uint64_t)(((b0 + tap * 8192) >> 4) | (1u << 16))) + 2 * kk; };
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (tap | kk) != 0;
            if (V == 0) {  // four tiles, plain
#pragma unroll
              for (int u = 0; u < 4; ++u) mma_ss(tmem + 64 * u, adesc(u, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 1) {
#pragma unroll
              for (int u = 0; u < 4; ++u) mma_ws(tmem + 64 * u, adesc(u, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 2) {
#pragma unroll
              for (int u = 0; u < 4; u += 2) {
                mma_ws_fill(tmem + 64 * u, adesc(u, kk), bdesc(tap, kk), idesc, acc);
                mma_ws_last(tmem + 64 * (u + 1), adesc(u + 1, kk), bdesc(tap, kk), idesc, acc);
              }
            } else if (V == 3) {
              mma_ws_fill(tmem, adesc(0, kk), bdesc(tap, kk), idesc, acc);
              mma_ws_use(tmem + 64, adesc(1, kk), bdesc(tap, kk), idesc, acc);
              mma_ws_use(tmem + 128, adesc(2, kk), bdesc(tap, kk), idesc, acc);
              mma_ws_last(tmem + 192, adesc(3, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 4) {  // one A window, three weight blocks (taps tap, tap+1, tap+2 mod 9), + a plain fourth
              mma_a_fill(tmem, adesc(0, kk), bdesc(tap, kk), idesc, acc);
              mma_a_use(tmem + 64, adesc(0, kk), bdesc((tap + 1) % 9, kk), idesc, acc);
              mma_a_last(tmem + 128, adesc(0, kk), bdesc((tap + 2) % 9, kk), idesc, acc);
              mma_ss(tmem + 192, adesc(3, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 6) {  // W2 with a foreign plain MMA (other operands, other columns) between fill and lastuse
              mma_ws_fill(tmem, adesc(0, kk), bdesc(tap, kk), idesc, acc);
              mma_ss(tmem + 128, adesc(2, kk), bdesc((tap + 4) % 9, kk), idesc, acc);
              mma_ws_last(tmem + 64, adesc(1, kk), bdesc(tap, kk), idesc, acc);
              mma_ss(tmem + 192, adesc(3, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 7) {  // reference for W6, plain
              mma_ss(tmem, adesc(0, kk), bdesc(tap, kk), idesc, acc);
              mma_ss(tmem + 128, adesc(2, kk), bdesc((tap + 4) % 9, kk), idesc, acc);
              mma_ss(tmem + 64, adesc(1, kk), bdesc(tap, kk), idesc, acc);
              mma_ss(tmem + 192, adesc(3, kk), bdesc(tap, kk), idesc, acc);
            } else if (V == 5) {  // reference for W4's accumulators, plain
              mma_ss(tmem, adesc(0, kk), bdesc(tap, kk), idesc, acc);
              mma_ss(tmem + 64, adesc(0, kk), bdesc((tap + 1) % 9, kk), idesc, acc);
              mma_ss(tmem + 128, adesc(0, kk), bdesc((tap + 2) % 9, kk), idesc, acc);
              mma_ss(tmem + 192, adesc(3, kk), bdesc(tap, kk), idesc, acc);
            }
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (lane == 0) res[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (dump && blockIdx.x == 0) {  // accumulators: 128 lanes x 256 columns
    for (int c = 0; c < 256; c += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + c + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) dump[(warp * 32 + lane) * 256 + c + j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static uint32_t* g_dump[8];

template <int V>
static void run(const char* name, int tiles, int grid, bool keep) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  uint32_t* dump = nullptr;
  if (keep) {
    cudaMalloc(&dump, 128 * 256 * 4);
    cudaMemset(dump, 0, 128 * 256 * 4);
  }
  const int smem = 1024 + kABytes + kBBytes;
  cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<V><<<grid, 128, smem>>>(d, dump, tiles);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < grid; ++i) m += (double)h[i];
  printf("%-64s grid=%3d: %7.2f cycles per MMA (%d MMAs)\n", name, grid, m / grid / tiles / 144.0, tiles * 144);
  if (keep) {
    g_dump[V] = (uint32_t*)malloc(128 * 256 * 4);
    cudaMemcpy(g_dump[V], dump, 128 * 256 * 4, cudaMemcpyDeviceToHost);
    cudaFree(dump);
  }
  cudaFree(d);
}

static void compare(int a, int b, const char* what) {
  long long bad = 0, nz = 0;
  for (int i = 0; i < 128 * 256; ++i) {
    bad += g_dump[a][i] != g_dump[b][i];
    nz += g_dump[a][i] != 0;
  }
  printf("%s: %lld of %d accumulator words differ (%lld non-zero in the reference)\n", what, bad, 128 * 256, nz);
}

int main() {
  // correctness first: ONE tile (accumulators are overwritten by the first MMA of a tile, so tiles = 1 is the product)
  run<0>("W0 plain SS, 4 tiles x 36", 1, 1, true);
  run<1>("W1 .ws no reuse", 1, 1, true);
  run<2>("W2 .ws fill + lastuse", 1, 1, true);
  run<3>("W3 .ws fill + 2 use + lastuse", 1, 1, true);
  run<4>("W4 A fill / use / lastuse x 3 weight blocks", 1, 1, true);
  run<5>("W5 reference for W4 (plain)", 1, 1, true);
  run<6>("W6 .ws fill / foreign plain MMA / lastuse", 1, 1, true);
  run<7>("W7 reference for W6 (plain)", 1, 1, true);
  compare(7, 6, "W6 vs W7 (a plain MMA between fill and lastuse)");
  compare(0, 1, "W1 vs W0");
  compare(0, 2, "W2 vs W0");
  compare(0, 3, "W3 vs W0");
  compare(5, 4, "W4 vs W5");
  const int tiles = 500;
  for (int grid : {1, 148}) {
    run<0>("W0 plain SS", tiles, grid, false);
    run<1>("W1 .ws, B filled by every MMA", tiles, grid, false);
    run<2>("W2 .ws, B fill + lastuse (2 tiles per weight block)", tiles, grid, false);
    run<3>("W3 .ws, B fill + use + use + lastuse (4 tiles per weight block)", tiles, grid, false);
    run<4>("W4 A fill + use + lastuse (3 weight blocks per A) + 1 plain", tiles, grid, false);
    run<6>("W6 .ws fill / plain / lastuse / plain", tiles, grid, false);
  }
  return 0;
}
