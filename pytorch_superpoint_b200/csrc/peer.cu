// The one exchange of the sharded HA path (export.py:61-63 `torch.sum(...)/torch.sum(...)` over ALL homographies) as a
// fused reduce + divide over NVLink peer memory -- no collective library kernel on the data path.
// Every rank keeps its partial planes sums[B,2,H,W] (sum heat*mask, sum mask over ITS views) in a cudaMalloc'ed
// buffer that the other ranks of the box map through CUDA IPC.  The owner of image b then runs ONE kernel that loads
// the partial planes of image b from every peer (P2P loads through NVSwitch), adds them in rank order (deterministic,
// independent of which rank computes) and divides: reduce-scatter + export.py:63 in a single pass, the summed planes
// never exist in memory.  Synchronisation is two monotonic counters per (peer, slot) pushed into the consumer's
// memory with system-scope stores: `ready` (the producer finished the partial planes of step k) and `consumed` (the
// owner finished reading them, the producer may overwrite the slot two steps later).  Waiting happens in a
// one-CTA kernel IN FRONT of the consumer kernel, so the consumer starts with clean caches and never spins itself.
#include "common.cuh"

namespace spb {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const float* p[kMaxPeers];
};
struct FlagPtrs {
  unsigned int* p[kMaxPeers];
};

__global__ void peer_signal_kernel(FlagPtrs f, int n, unsigned int value) {
  if ((int)threadIdx.x < n && f.p[threadIdx.x]) {
    __threadfence_system();  // everything this stream wrote before is visible system-wide first
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[threadIdx.x]), "r"(value) : "memory");
  }
}

// thread t waits until flags[t] >= value (skip[t] != 0: not waited for).  Bounded: a protocol bug aborts, not hangs.
__global__ void peer_wait_kernel(const unsigned int* flags, int n, unsigned int value, unsigned int skip_mask) {
  if ((int)threadIdx.x < n && !((skip_mask >> threadIdx.x) & 1u)) {
    const long long t0 = clock64();
    unsigned int v;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
      if ((int)(v - value) >= 0) break;  // counters are monotonic (wrap-safe compare)
      if (clock64() - t0 > 20000000000LL) {
        printf("spb200 peer_wait: timed out waiting for peer %d (have %u, want %u)\n", (int)threadIdx.x, v, value);
        __trap();
      }
      __nanosleep(200);
    }
  }
}

// heat[b] = (sum_p sums_p[b0+b][0]) / (sum_p sums_p[b0+b][1]), peers added in rank order; float4 per thread
__global__ void __launch_bounds__(256) finalize_peers_kernel(PeerPtrs pp, int world, long long b0, int nb, long long hw,
                                                             float* __restrict__ heat) {
  const long long n4 = hw >> 2;  // (H * W is a multiple of 64: H, W multiples of 8)
  const long long total = (long long)nb * n4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long b = i / n4, j = i % n4;
    const long long off = ((b0 + b) * 2) * hw + j * 4;
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f), den = num;
    for (int p = 0; p < world; ++p) {
      const float4 a = *reinterpret_cast<const float4*>(pp.p[p] + off);
      const float4 d = *reinterpret_cast<const float4*>(pp.p[p] + off + hw);
      num.x += a.x; num.y += a.y; num.z += a.z; num.w += a.w;
      den.x += d.x; den.y += d.y; den.z += d.z; den.w += d.w;
    }
    float4 o;
    o.x = __fdiv_rn(num.x, den.x);
    o.y = __fdiv_rn(num.y, den.y);
    o.z = __fdiv_rn(num.z, den.z);
    o.w = __fdiv_rn(num.w, den.w);
    *reinterpret_cast<float4*>(heat + b * hw + j * 4) = o;
  }
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_peer_alloc(void** ptr, size_t bytes) {
  SPB_CHECK_ARG(ptr && bytes > 0);
  SPB_CHECK_CUDA(cudaMalloc(ptr, bytes));
  SPB_CHECK_CUDA(cudaMemset(*ptr, 0, bytes));
  SPB_CHECK_CUDA(cudaDeviceSynchronize());
  return SPB_OK;
}

int spb_peer_free(void* ptr) {
  if (ptr) SPB_CHECK_CUDA(cudaFree(ptr));
  return SPB_OK;
}

int spb_peer_export(void* ptr, unsigned char* handle64) {
  SPB_CHECK_ARG(ptr && handle64);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  SPB_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return SPB_OK;
}

int spb_peer_open(const unsigned char* handle64, void** ptr) {
  SPB_CHECK_ARG(handle64 && ptr);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SPB_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SPB_OK;
}

int spb_peer_close(void* ptr) {
  if (ptr) SPB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return SPB_OK;
}

int spb_peer_signal(unsigned int* const* flags, int n, unsigned int value, void* stream) {
  SPB_CHECK_ARG(flags && n >= 1 && n <= kMaxPeers);
  FlagPtrs f;
  for (int i = 0; i < kMaxPeers; ++i) f.p[i] = i < n ? flags[i] : nullptr;
  peer_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(f, n, value);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_peer_wait(const unsigned int* flags, int n, unsigned int value, unsigned int skip_mask, void* stream) {
  SPB_CHECK_ARG(flags && n >= 1 && n <= kMaxPeers);
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(flags, n, value, skip_mask);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

int spb_ha_finalize_peers(const float* const* peer_sums, int world, int b0, int nb, float* heat, int H, int W,
                          void* stream) {
  SPB_CHECK_ARG(peer_sums && world >= 1 && world <= kMaxPeers && b0 >= 0 && nb >= 1 && heat && H >= 1 && W >= 1);
  SPB_CHECK_ARG(((long long)H * W) % 4 == 0);
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < world ? peer_sums[i] : nullptr;
  for (int i = 0; i < world; ++i) SPB_CHECK_ARG(pp.p[i] != nullptr);
  const long long hw = (long long)H * W, total = (long long)nb * (hw >> 2);
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  finalize_peers_kernel<<<grid, 256, 0, as_stream(stream)>>>(pp, world, b0, nb, hw, heat);
  SPB_CHECK_LAUNCH();
  return SPB_OK;
}

}  // extern "C"
