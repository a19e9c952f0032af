// Micro-benchmark (not product code): tcgen05.mma issue rate for the conv tile shapes (M128, N in {64..256},
// K16 bf16, A and B from SWIZZLE_128B shared memory, or A from TMEM), tcgen05.ld rate, and the MMA rate while
// epilogue-like warps hammer TMEM loads / shared-memory stores.  Grounds the N=64 ceiling used in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma.bin tools/ubench_mma.cu && tools/ubench_mma.bin
#include <cstdint>
#include <cstdio>
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
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
               "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Res {
  long long mma_cycles, ld_cycles;
};

// mode bit0: A from TMEM; bit1: 8 warps loop tcgen05.ld while the MMAs run; bit2: those warps also store 16 B/lane
// to shared memory per loaded 32 columns (the conv1a -> A-operand write-back of the fused kernel)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// MODE bit3: alternate two accumulators MMA by MMA (two independent accumulation chains)
template <int N, int MODE>
__global__ void __launch_bounds__(320, 1) bench(Res* res, int iters, int pitch) {
  constexpr int mode = MODE;
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                   // 32 KB: halo block
  uint8_t* sB = smem + 32768;           // N x 128 B
  uint8_t* sW = sB + 256 * 128;         // 32 KB scratch for store traffic
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    stop = 0;
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
  if (warp == 1) {
    if (iters > 0) {
      constexpr uint32_t idesc = make_idesc(128, N);
      const uint64_t hi_a = ((uint64_t)((pitch * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + ((N <= 128 || (mode & 16)) ? (it & 1) * N : 0);
        if (mode & 16) {  // conv_s3 tile: 12 MMAs then 24 row shifts of the partial-sum blocks
          if (elect_one()) {
#pragma unroll
            for (int g = 0; g < 12; ++g) {
              const uint32_t off = ((g / 4) * pitch) * 128;
              const uint64_t ad = hi_a | (uint64_t)(((a0 + off) >> 4) | (1u << 16));
              const uint64_t bd = hi_b | (uint64_t)((b0 >> 4) | (1u << 16));
              tc_mma_ss(d, ad + 2 * (g & 3), bd + 2 * (g & 3), idesc, g != 0);
            }
            if (mode & 32) {
#pragma unroll
              for (int c8 = 64; c8 < 192; c8 += 8) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(d + c8) : "memory");
#pragma unroll
              for (int c8 = 128; c8 < 192; c8 += 8) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(d + c8) : "memory");
            }
          }
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t off = ((tap / 3) * pitch + tap % 3) * 128;
          const uint64_t ad = hi_a | (uint64_t)(((a0 + off) >> 4) | (1u << 16));
          const uint64_t bd = hi_b | (uint64_t)((b0 >> 4) | (1u << 16));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t dd = (mode & 8) ? tmem + (k & 1) * N : d;
              const uint32_t accf = (mode & 8) ? (tap | (k >> 1)) != 0 : (tap | k) != 0;
              if (mode & 1)
                tc_mma_ts(dd, tmem + 256 + tap * 8 + k * 8, bd + 2 * k, idesc, accf);
              else
                tc_mma_ss(dd, ad + 2 * k, bd + 2 * k, idesc, accf);
            }
          }
          __syncwarp();
        }
      }
      if (elect_one()) tc_commit(smem_u32(&bar[0]));
      __syncwarp();
      mbar_wait(smem_u32(&bar[0]), 0);
      if (lane == 0) {
        res[blockIdx.x].mma_cycles = clock64() - t0;
        stop = 1;
      }
    }
  } else if (warp >= 2 && (mode & 2)) {
    const int quad = warp & 3;
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (iters > 0 ? !stop : n < 4096) {
      uint32_t v[32];
      tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + ((n & 3) * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (mode & 4) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(sW + ((warp - 2) * 32 + lane) * 128 + ((j / 8) ^ (lane & 7)) * 16) =
              make_uint4(v[j] ^ v[j + 1], v[j + 2] ^ v[j + 3], v[j + 4] ^ v[j + 5], v[j + 6] ^ v[j + 7]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
      ++n;
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) res[blockIdx.x].ld_cycles = -1;
    if (warp == 2 && lane == 0) res[blockIdx.x].ld_cycles = n ? (t1 - t0) * 1000 / n : 0;  // milli-cycles per x32 load per warp
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N, int MODE>
static void run(const char* name, int iters, int pitch, int grid) {
  const int mode = MODE;
  Res* d;
  cudaMalloc(&d, sizeof(Res) * grid);
  cudaMemset(d, 0, sizeof(Res) * grid);
  const int smem = 1024 + 32768 + 256 * 128 + 32768;
  cudaFuncSetAttribute(bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, MODE><<<grid, 320, smem>>>(d, iters, pitch);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s N=%d mode=%d: CUDA error %s\n", name, N, mode, cudaGetErrorString(e));
    exit(1);
  }
  Res h[148];
  cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
  double mm = 0, ld = 0;
  for (int i = 0; i < grid; ++i) {
    mm += (double)h[i].mma_cycles;
    ld += (double)h[i].ld_cycles;
  }
  mm /= grid;
  ld /= grid;
  if (iters > 0)
    printf("%-34s N=%3d grid=%3d: %7.1f cyc/MMA (ideal %3d)  -> %5.1f%% of tensor peak;  ld %.1f cyc per x32 per warp\n", name, N,
           grid, mm / (iters * 36.0), N / 2, 100.0 * (N / 2) / (mm / (iters * 36.0)), ld / 1000.0);
  else
    printf("%-34s ld only grid=%3d: %.1f cyc per tcgen05.ld.32x32b.x32 per warp (8 warps => %.1f B/cyc/SM)\n", name, grid,
           ld / 1000.0, 8 * 4096.0 / (ld / 1000.0));
  cudaFree(d);
}

int main() {
  const int it = 2000;
  for (int grid : {1, 148}) {
    run<64, 0>("SS (A,B smem) pitch10", it, 10, grid);
    run<128, 0>("SS (A,B smem) pitch10", it, 10, grid);
    run<192, 0>("SS (A,B smem) pitch10", it, 10, grid);
    run<256, 0>("SS (A,B smem) pitch10", it, 10, grid);
    run<64, 0>("SS pitch8", it, 8, grid);
    run<192, 0>("SS pitch32 (4x32 tile)", it, 32, grid);
    run<64, 8>("SS two accumulation chains", it, 10, grid);
    run<128, 8>("SS two accumulation chains", it, 10, grid);
    run<64, 1>("TS (A tmem)", it, 10, grid);
    run<128, 1>("TS (A tmem)", it, 10, grid);
    run<64, 9>("TS two chains", it, 10, grid);
    run<64, 2>("SS + 8 warps tcgen05.ld", it, 10, grid);
    run<64, 6>("SS + 8 warps ld + st.shared", it, 10, grid);
    run<192, 6>("SS + 8 warps ld + st.shared", it, 32, grid);
    run<192, 16>("s3 tile: 12 MMAs (x3 = per 36)", it, 16, grid);
    run<192, 48>("s3 tile: 12 MMAs + 24 shifts (x3)", it, 16, grid);
    run<64, 2>("ld only", 0, 10, grid);
    run<64, 6>("ld + st.shared only", 0, 10, grid);
  }
  return 0;
}
