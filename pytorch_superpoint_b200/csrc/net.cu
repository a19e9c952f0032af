// (2) SuperPointNet forward (encoder + detector/descriptor heads) and the whole-HA entry point.
// models/SuperPointNet_pretrained.py:46-76, SuperPointNet_gauss2.py:43-70 (+unet_parts.py:10-48),
// SuperPointNet.py:154-224 -- eval-mode BatchNorm is folded into the conv weights by the caller.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "net.cuh"

namespace spb {

struct LayerSpec {
  int cin, cout, ks, relu, pool;
};
static const LayerSpec kSpec[kNumLayers] = {
    {1, 64, 3, 1, 0},    {64, 64, 3, 1, 1},    {64, 64, 3, 1, 0},   {64, 64, 3, 1, 1},
    {64, 128, 3, 1, 0},  {128, 128, 3, 1, 1},  {128, 128, 3, 1, 0}, {128, 128, 3, 1, 0},
    {128, 256, 3, 1, 0}, {256, 65, 1, 0, 0},   {128, 256, 3, 1, 0}, {256, 256, 1, 0, 0}};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline void* align_ptr(void* p) {  // workspaces carry 1 KB of slack for this
  return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// The workspace-size entry points take no network handle, so the layout cannot depend on a net's implementation
// switch: once any net of the process has been put on the CUDA-core validation path (which runs conv1a and conv1b
// as two kernels and needs the full-resolution intermediate), every workspace is sized for it.
static bool g_simt_used = false;
static inline bool first_block_fused(int H, int W) {
  return !g_simt_used && conv1_fused_enabled() && conv1_fused_supports(H, W);
}

struct Workspace {  // carve-up of the caller's buffer for a batch of B views at H x W
  __nv_bfloat16 *buf0, *buf1;
  size_t bytes;
  Workspace(void* base, int B, int H, int W) {
    // buf0 holds conv1a's full-resolution output only when the first block runs as two kernels; with the fused
    // block (the default) its largest tenant is conv2a's H/2 x W/2 output -- 4x smaller (800 views at 240x320:
    // 2 GB instead of 7.9 GB)
    const size_t full = first_block_fused(H, W) ? (size_t)(H / 2) * (W / 2) : (size_t)H * W;
    const size_t b0 = align_up((size_t)B * full * 64 * 2, 1024);
    const size_t b1 = align_up((size_t)B * (H / 2) * (W / 2) * 64 * 2, 1024);
    char* p = static_cast<char*>(base);
    buf0 = reinterpret_cast<__nv_bfloat16*>(p);
    buf1 = reinterpret_cast<__nv_bfloat16*>(p + b0);
    bytes = b0 + b1;
  }
};

// ---- optional per-stage CUDA-event timing (events are recorded on the launch stream) -------------------
struct StageTimer {
  Net* net;
  ProfRec* r;
  cudaStream_t st;
  StageTimer(const Net* n, int stage, cudaStream_t s) : net(const_cast<Net*>(n)), r(nullptr), st(s) {
    if (!net->profile) return;
    if (net->usedrecs == net->nrecs) {
      if (net->nrecs == net->caprecs) {
        const int cap = net->caprecs ? net->caprecs * 2 : 256;
        ProfRec* nr = static_cast<ProfRec*>(realloc(net->recs, cap * sizeof(ProfRec)));
        if (!nr) return;
        net->recs = nr;
        net->caprecs = cap;
      }
      ProfRec& x = net->recs[net->nrecs];
      if (cudaEventCreate(&x.a) != cudaSuccess || cudaEventCreate(&x.b) != cudaSuccess) return;
      net->nrecs++;
    }
    r = &net->recs[net->usedrecs++];
    r->stage = stage;
    cudaEventRecord(r->a, st);
  }
  ~StageTimer() {
    if (r) cudaEventRecord(r->b, st);
  }
};

__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ semi, float* __restrict__ probs,
                                                           long long rows) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    const float* p = semi + r * 65;
    const float a = p[lane], b = p[lane + 32], d = (lane == 0) ? p[64] : -INFINITY;
    const float mx = warp_max(fmaxf(fmaxf(a, b), d));
    const float ea = expf(a - mx), eb = expf(b - mx), ed = (lane == 0) ? expf(d - mx) : 0.f;
    const float s = warp_sum(ea + eb + ed);
    probs[r * 64 + lane] = __fdiv_rn(ea, s);
    probs[r * 64 + lane + 32] = __fdiv_rn(eb, s);
  }
}

static int run_layer(const Net* net, int li, const void* in, void* out, int B, int H, int W, int out_mode,
                     cudaStream_t st) {
  const LayerDev& L = net->layer[li];
  StageTimer tm(net, li, st);
  if (li == 0)
    return net->impl == SPB_IMPL_TCGEN05
               ? conv1a_tc(L, static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), B, H, W, st)
               : conv1a_simt(L, static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), B, H, W, st);
  if (net->impl == SPB_IMPL_TCGEN05)
    return conv_tc(L, static_cast<const __nv_bfloat16*>(in), out, B, H, W, out_mode, st);
  int rc = conv_simt(L, static_cast<const __nv_bfloat16*>(in), out, B, H, W, out_mode != 0, st);
  if (rc == SPB_OK && out_mode == 2) rc = l2norm_rows(static_cast<float*>(out), (long long)B * H * W, L.cout, st);
  return rc;
}

// semi: logits [B,Hc,Wc,65] (or NULL); probs: softmax probabilities [B,Hc,Wc,64] (or NULL); desc optional;
// heat16 (+ bits): fp16 probability planes of the HA pass (conv_tc out_mode 4), tcgen05 only
static int forward(const Net* net, const float* x, int B, int H, int W, float* semi, float* probs, float* desc,
                   void* ws, cudaStream_t st, void* heat16 = nullptr, const uint8_t* bits = nullptr) {
  Workspace w(ws, B, H, W);
  int rc;
#define RUN(li, in, out, h, wd, mode)                                     \
  if ((rc = run_layer(net, li, in, out, B, h, wd, mode, st)) != SPB_OK) return rc
  if (net->impl == SPB_IMPL_TCGEN05 && first_block_fused(H, W)) {
    // conv1a + conv1b + pool in one kernel: the 64-channel full-resolution activation never leaves the SM
    StageTimer tm(net, 1, st);
    if ((rc = conv1_fused(net->layer[0], net->layer[1], x, w.buf1, B, H, W, st)) != SPB_OK) return rc;
  } else {
    RUN(0, x, w.buf0, H, W, 0);
    RUN(1, w.buf0, w.buf1, H, W, 0);  // + pool -> H/2
  }
  RUN(2, w.buf1, w.buf0, H / 2, W / 2, 0);
  RUN(3, w.buf0, w.buf1, H / 2, W / 2, 0);  // + pool -> H/4
  RUN(4, w.buf1, w.buf0, H / 4, W / 4, 0);
  RUN(5, w.buf0, w.buf1, H / 4, W / 4, 0);  // + pool -> H/8
  RUN(6, w.buf1, w.buf0, H / 8, W / 8, 0);
  RUN(7, w.buf0, w.buf1, H / 8, W / 8, 0);
  RUN(8, w.buf1, w.buf0, H / 8, W / 8, 0);
  if (semi) RUN(9, w.buf0, semi, H / 8, W / 8, 1);
  if (probs) {  // validation path: logits (semi) first, then a softmax kernel
    if (!semi) {
      set_error("forward: fp32 probabilities need a semi buffer (they are the CUDA-core validation path's form)");
      return SPB_EINVAL;
    }
    const long long rows = (long long)B * (H / 8) * (W / 8);
    softmax_rows_kernel<<<min(ceil_div(rows, 8), kNumSMs * 8), 256, 0, st>>>(semi, probs, rows);
    SPB_CHECK_LAUNCH();
  }
  if (heat16) {
    if (net->impl != SPB_IMPL_TCGEN05 || !bits) {
      set_error("forward: fp16 probability planes are produced by the tcgen05 path only");
      return SPB_EINVAL;
    }
    StageTimer tm(net, 9, st);
    if ((rc = conv_tc(net->layer[9], w.buf0, heat16, B, H / 8, W / 8, 4, st, bits)) != SPB_OK) return rc;
  }
  if (desc) {
    RUN(10, w.buf1, w.buf0, H / 8, W / 8, 0);
    RUN(11, w.buf0, desc, H / 8, W / 8, 2);
  }
#undef RUN
  return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

static int net_fill(Net* net, const float* const* weights, const float* const* biases) {
  int dev = 0;
  cudaDeviceProp prop;
  SPB_CHECK_CUDA(cudaGetDevice(&dev));
  SPB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("spb_net_create: device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major,
              prop.minor);
    return SPB_EARCH;
  }
  net->impl = SPB_IMPL_TCGEN05;
  net->ha_overlap = 0;
  for (int li = 0; li < kNumLayers; ++li) {
    SPB_CHECK_ARG(weights[li] && biases[li]);
    LayerDev& L = net->layer[li];
    const LayerSpec& S = kSpec[li];
    L.cin = S.cin;
    L.cout = S.cout;
    L.ks = S.ks;
    L.relu = S.relu;
    L.pool = S.pool;
    L.cout_pad = (S.cout + 15) / 16 * 16;
    const int taps = S.ks * S.ks, K = taps * S.cin;
    std::vector<__nv_bfloat16> ws((size_t)taps * S.cin * L.cout_pad, __float2bfloat16(0.f));
    const int Kpad = (li == 0) ? 16 : K;  // conv1a: 9 taps padded to one UMMA K step
    std::vector<__nv_bfloat16> wt((size_t)L.cout_pad * Kpad, __float2bfloat16(0.f));
    std::vector<float> b(L.cout_pad, 0.f);
    const float* w = weights[li];  // OIHW
    for (int co = 0; co < S.cout; ++co) {
      b[co] = biases[li][co];
      for (int ci = 0; ci < S.cin; ++ci)
        for (int t = 0; t < taps; ++t) {
          const __nv_bfloat16 v = __float2bfloat16_rn(w[((size_t)co * S.cin + ci) * taps + t]);
          ws[((size_t)t * S.cin + ci) * L.cout_pad + co] = v;
          wt[(size_t)co * Kpad + (size_t)t * S.cin + ci] = v;
        }
    }
    if (li == 0) {
      // tensor-core conv1a folds its bias into the MMA: im2col slots 9 and 10 carry the constant 1 and the
      // weight rows carry the bias split into a bf16 head and a bf16 residual (fp32-accurate to ~2^-17)
      for (int co = 0; co < S.cout; ++co) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(b[co]);
        wt[(size_t)co * Kpad + 9] = hi;
        wt[(size_t)co * Kpad + 10] = __float2bfloat16_rn(b[co] - __bfloat162float(hi));
      }
    }
    SPB_CHECK_CUDA(cudaMalloc(&L.bias, b.size() * sizeof(float)));
    SPB_CHECK_CUDA(cudaMemcpy(L.bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    SPB_CHECK_CUDA(cudaMalloc(&L.w_simt, ws.size() * 2));
    SPB_CHECK_CUDA(cudaMemcpy(L.w_simt, ws.data(), ws.size() * 2, cudaMemcpyHostToDevice));
    SPB_CHECK_CUDA(cudaMalloc(&L.w_tc, wt.size() * 2));
    SPB_CHECK_CUDA(cudaMemcpy(L.w_tc, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
    if (li == 0) {
      std::vector<float> w1(9 * 64);
      for (int t = 0; t < 9; ++t)
        for (int co = 0; co < 64; ++co) w1[t * 64 + co] = __bfloat162float(ws[(size_t)t * 64 + co]);
      SPB_CHECK_CUDA(cudaMalloc(&L.w1a_f32, w1.size() * sizeof(float)));
      SPB_CHECK_CUDA(cudaMemcpy(L.w1a_f32, w1.data(), w1.size() * sizeof(float), cudaMemcpyHostToDevice));
    } else {
      int rc = conv_tc_prepare(L);
      if (rc != SPB_OK) return rc;
      if (S.ks == 3 && S.cin == 64 && S.cout == 64) {
        std::vector<__nv_bfloat16> w3((size_t)576 * 64);
        for (int r = 0; r < 3; ++r)
          for (int sx = 0; sx < 3; ++sx)
            for (int co = 0; co < 64; ++co)
              for (int ci = 0; ci < 64; ++ci)
                w3[((size_t)(r * 3 + sx) * 64 + co) * 64 + ci] = wt[(size_t)co * Kpad + (size_t)(r * 3 + sx) * 64 + ci];
        SPB_CHECK_CUDA(cudaMalloc(&L.w_s3, w3.size() * 2));
        SPB_CHECK_CUDA(cudaMemcpy(L.w_s3, w3.data(), w3.size() * 2, cudaMemcpyHostToDevice));
        if ((rc = conv_s3_prepare(L)) != SPB_OK) return rc;
      }
    }
  }
  return SPB_OK;
}

int spb_net_destroy(spb_net* h);

int spb_net_create(spb_net** out, const float* const* weights, const float* const* biases) {
  SPB_CHECK_ARG(out && weights && biases);
  *out = nullptr;
  Net* net = static_cast<Net*>(calloc(1, sizeof(Net)));
  if (!net) return SPB_ENOMEM;
  const int rc = net_fill(net, weights, biases);
  if (rc != SPB_OK) {  // free what the failed construction had allocated (device buffers, tensor maps)
    spb_net_destroy(reinterpret_cast<spb_net*>(net));
    return rc;
  }
  *out = reinterpret_cast<spb_net*>(net);
  return SPB_OK;
}

int spb_net_destroy(spb_net* h) {
  if (!h) return SPB_OK;
  Net* net = reinterpret_cast<Net*>(h);
  for (int li = 0; li < kNumLayers; ++li) {
    LayerDev& L = net->layer[li];
    cudaFree(L.bias);
    cudaFree(L.w_simt);
    cudaFree(L.w_tc);
    cudaFree(L.w1a_f32);
    cudaFree(L.w_s3);
    free(L.tmap_w);
    free(L.tmap_w3);
  }
  for (int i = 0; i < net->nrecs; ++i) {
    cudaEventDestroy(net->recs[i].a);
    cudaEventDestroy(net->recs[i].b);
  }
  free(net->recs);
  if (net->aux) {
    cudaStreamDestroy(net->aux);
    cudaEventDestroy(net->ev_begin);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(net->ev_warp[i]);
      cudaEventDestroy(net->ev_fwd[i]);
      cudaEventDestroy(net->ev_agg[i]);
    }
  }
  free(net);
  return SPB_OK;
}

int spb_net_set_impl(spb_net* h, int impl) {
  SPB_CHECK_ARG(h && (impl == SPB_IMPL_TCGEN05 || impl == SPB_IMPL_SIMT));
  if (impl == SPB_IMPL_SIMT) g_simt_used = true;  // (sticky: see first_block_fused)
  reinterpret_cast<Net*>(h)->impl = impl;
  return SPB_OK;
}

size_t spb_net_workspace_bytes(int B, int H, int W, int want_desc) {
  (void)want_desc;
  if (B < 1 || H < 8 || W < 8) return 0;
  Workspace w(nullptr, B, H, W);
  return w.bytes + 1024;
}

int spb_net_forward(spb_net* h, const float* x, int B, int H, int W, float* semi, float* desc, void* ws,
                    size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(h && x && semi && ws);
  SPB_CHECK_ARG(B >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  if (ws_bytes < spb_net_workspace_bytes(B, H, W, desc != nullptr)) {
    set_error("spb_net_forward: workspace %zu < %zu bytes", ws_bytes, spb_net_workspace_bytes(B, H, W, desc != nullptr));
    return SPB_ENOMEM;
  }
  return forward(reinterpret_cast<Net*>(h), x, B, H, W, semi, nullptr, desc, align_ptr(ws), as_stream(stream));
}

int spb_net_layer(spb_net* h, int layer, const void* in, void* out, int B, int H, int W, void* stream) {
  SPB_CHECK_ARG(h && in && out && layer >= 0 && layer < kNumLayers && B >= 1 && H >= 1 && W >= 1);
  const int mode = layer == 9 ? 1 : (layer == 11 ? 2 : 0);
  return run_layer(reinterpret_cast<Net*>(h), layer, in, out, B, H, W, mode, as_stream(stream));
}

// ---- whole HA step ----------------------------------------------------------------------------------
// One pass = V views (nb images) through: prep -> warp (+ mask bits) -> detector net -> gather-aggregate.
// tcgen05 (product): ha_fast.cu kernels, the detector head writes masked fp16 probability planes.
// SIMT (validation): the bit-exact CUDA-core warp / aggregate kernels around logits + a softmax kernel.
struct HaLayout {
  size_t warp, bits, semi, probs, agg, net, warp2, bits2, src_pad, fold, total;
};
static size_t heat16_bytes(size_t V, int H, int W) {
  return V * (size_t)(H + 2 * kHeatPadY) * (size_t)(W + 2 * kHeatPadX) * 2;
}
// V views of nb images (at most max_views per image); two = a second set of warp / bits / probability buffers
// (overlapped passes of spb_ha_heatmap_sums_batch)
static HaLayout ha_layout(int nb, size_t V, int max_views, int H, int W, bool simt, bool two) {
  const size_t cells = (size_t)(H / 8) * (W / 8);
  HaLayout L;
  size_t o = 0;
  L.warp = o;
  o += align_up(V * H * W * 4, 1024);
  L.bits = o;
  o += align_up(V * H * (W / 8), 1024);
  L.semi = o;  // logits (SIMT validation path) / second probability buffer (overlapped passes)
  o += (simt || two) ? align_up(simt ? V * cells * 65 * 4 : heat16_bytes(V, H, W), 1024) : 0;
  L.probs = o;  // fp32 [V,cells,64] (SIMT) or the fp16 planes (tcgen05)
  o += align_up(simt ? V * cells * 64 * 4 : heat16_bytes(V, H, W), 1024);
  L.agg = o;
  o += align_up(aggregate_probs_workspace(nb, max_views, H, W), 1024);
  L.net = o;
  o += spb_net_workspace_bytes((int)V, H, W, 0);
  L.warp2 = o;
  o += two ? align_up(V * H * W * 4, 1024) : 0;
  L.bits2 = o;
  o += two ? align_up(V * H * (W / 8), 1024) : 0;
  L.src_pad = o;
  o += align_up((size_t)nb * (H + 2 * kHaPad) * (W + 2 * kHaPad) * 4, 1024);
  L.fold = o;  // [2 sets][2][V] FoldMat
  o += align_up(4 * V * sizeof(FoldMat), 1024);
  L.total = o + 1024;
  return L;
}

// Views per pass: bounded by BYTES (views x H x W), not by a count -- 800 views at 240x320 (61 M pixels: 0.25 GB of
// fp32 views + 2 GB of half-resolution bf16 activations), 200 at 480x640, 128 at 384x1248.
static int ha_pass_views(const Net* net, int H, int W) {
  if (net && net->ha_chunk > 0) return net->ha_chunk;
  const long long cap = 800LL * 240 * 320 / ((long long)H * W);
  return cap < 1 ? 1 : (int)cap;
}

static void ha_plan(const Net* net, int B, int N, int H, int W, int& Bg, int& n) {
  const int cap = ha_pass_views(net, H, W);
  if (cap >= N) {
    n = N;
    Bg = cap / N;
    if (Bg > B) Bg = B;
  } else {
    n = cap;
    Bg = 1;
  }
}

struct PassBufs {
  float* warped;
  uint8_t* bits;
  void* probs;     // fp16 planes (tcgen05) or fp32 [V,cells,64] (SIMT)
  float* semi;     // SIMT only
  FoldMat* fold;   // [2][V]: forward, un-warp
  float* src_pad;
  void* agg_ws;
  void* net_ws;
};

// imgs: first image of the pass; mats_*: the arrays vm.mat() indexes; offs: NULL (nb x n views) or device int[nb+1]
static int ha_warp_stage(Net* net, const PassBufs& pb, const float* imgs, const float* mats_unwarp,
                         const float* mats_fwd, const ViewMap& vm, int nb, int V, int H, int W, cudaStream_t st) {
  int rc;
  if (net->impl == SPB_IMPL_TCGEN05) {
    StageTimer tm(net, 12, st);
    rc = ha_prep(imgs, pb.src_pad, mats_fwd, mats_unwarp, pb.fold, pb.fold + V, pb.probs, vm, nb, V, H, W, st);
    if (rc != SPB_OK) return rc;
    return warp_views(pb.src_pad, pb.fold, mats_fwd, pb.warped, pb.bits, vm, V, H, W, st);
  }
  StageTimer tm(net, 12, st);
  return warp_with_mask_bits(imgs, mats_fwd, pb.warped, pb.bits, V, vm, H, W, st);
}
static int ha_net_stage(Net* net, const PassBufs& pb, int V, int H, int W, cudaStream_t st) {
  if (net->impl == SPB_IMPL_TCGEN05)
    return forward(net, pb.warped, V, H, W, nullptr, nullptr, nullptr, pb.net_ws, st, pb.probs, pb.bits);
  return forward(net, pb.warped, V, H, W, pb.semi, static_cast<float*>(pb.probs), nullptr, pb.net_ws, st);
}
static int ha_agg_stage(Net* net, const PassBufs& pb, const float* mats_unwarp, const ViewMap& vm, int nb,
                        int max_views, int V, int H, int W, float* sums, int accumulate, const int* offs,
                        cudaStream_t st) {
  StageTimer tm(net, 13, st);
  if (net->impl == SPB_IMPL_TCGEN05)
    return aggregate_heat16(pb.probs, pb.fold + V, nb, max_views, H, W, sums, accumulate, pb.agg_ws, st, offs);
  return aggregate_probs(static_cast<const float*>(pb.probs), pb.bits, mats_unwarp, vm, nb, max_views, H, W, sums,
                         accumulate, pb.agg_ws, st, offs);
}

size_t spb_ha_workspace_bytes_batch(spb_net* h, int B, int N, int H, int W) {
  if (!h || B < 1 || N < 1 || H < 8 || W < 8) return 0;
  Net* net = reinterpret_cast<Net*>(h);
  int Bg, n;
  ha_plan(net, B, N, H, W, Bg, n);
  return ha_layout(Bg, (size_t)Bg * n, n, H, W, net->impl != SPB_IMPL_TCGEN05, net->ha_overlap != 0).total;
}

size_t spb_ha_workspace_bytes(int N, int H, int W) {
  if (N < 1 || H < 8 || W < 8) return 0;
  return ha_layout(1, (size_t)N, N, H, W, true, true).total;  // an upper bound for every setting of a single image
}

int spb_ha_pass_views(spb_net* h, int H, int W) {
  if (H < 8 || W < 8) return 0;
  return ha_pass_views(reinterpret_cast<Net*>(h), H, W);
}

int spb_ha_heatmap_sums_batch(spb_net* h, const float* imgs, const float* mats_unwarp, const float* mats_fwd,
                              int shared_mats, int B, int N, int H, int W, float* sums, int accumulate, void* ws,
                              size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(h && imgs && mats_unwarp && mats_fwd && sums && ws);
  SPB_CHECK_ARG(B >= 1 && N >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  Net* net = reinterpret_cast<Net*>(h);
  int Bg, n;
  ha_plan(net, B, N, H, W, Bg, n);
  const bool tc = net->impl == SPB_IMPL_TCGEN05;
  const HaLayout L = ha_layout(Bg, (size_t)Bg * n, n, H, W, !tc, net->ha_overlap != 0);
  if (ws_bytes < L.total) {
    set_error("spb_ha_heatmap_sums: workspace %zu < %zu bytes", ws_bytes, L.total);
    return SPB_ENOMEM;
  }
  char* base = static_cast<char*>(align_ptr(ws));
  cudaStream_t st = as_stream(stream);
  // the passes: Bg images x n views each (view chunks of ONE image only when it has more views than a pass holds;
  // per-image matrix sets are only contiguous over a whole image, so view chunks imply bg == 1)
  struct Pass {
    int b0, bg, n0, nn;
  };
  const int passes_b = ceil_div(B, Bg), passes_n = ceil_div(N, n), P = passes_b * passes_n;
  auto pass = [&](int c) {
    Pass p;
    p.b0 = (c / passes_n) * Bg;
    p.bg = (B - p.b0 < Bg) ? B - p.b0 : Bg;
    p.n0 = (c % passes_n) * n;
    p.nn = (N - p.n0 < n) ? N - p.n0 : n;
    return p;
  };
  auto moff = [&](const Pass& p) { return shared_mats ? 9LL * p.n0 : 9LL * ((long long)p.b0 * N + p.n0); };
  const size_t Vmax = (size_t)Bg * n;
  auto bufs = [&](int c) {  // pass buffers alternate by pass parity in the overlapped mode
    const int k = (net->ha_overlap && tc) ? (c & 1) : 0;
    PassBufs pb;
    pb.warped = reinterpret_cast<float*>(base + (k ? L.warp2 : L.warp));
    pb.bits = reinterpret_cast<uint8_t*>(base + (k ? L.bits2 : L.bits));
    pb.probs = base + ((k && tc) ? L.semi : L.probs);
    pb.semi = reinterpret_cast<float*>(base + L.semi);
    pb.fold = reinterpret_cast<FoldMat*>(base + L.fold) + (size_t)k * 2 * Vmax;
    pb.src_pad = reinterpret_cast<float*>(base + L.src_pad);
    pb.agg_ws = base + L.agg;
    pb.net_ws = base + L.net;
    return pb;
  };
  auto vmap = [&](const Pass& p) {
    ViewMap vm;
    vm.view_image = nullptr;
    vm.view_mat = nullptr;
    vm.vpi = p.nn;
    vm.shared = shared_mats;
    return vm;
  };
  auto do_warp = [&](int c, cudaStream_t s) {
    const Pass p = pass(c);
    return ha_warp_stage(net, bufs(c), imgs + (long long)p.b0 * H * W, mats_unwarp + moff(p), mats_fwd + moff(p),
                         vmap(p), p.bg, p.bg * p.nn, H, W, s);
  };
  auto do_agg = [&](int c, cudaStream_t s) {
    const Pass p = pass(c);
    return ha_agg_stage(net, bufs(c), mats_unwarp + moff(p), vmap(p), p.bg, p.nn, p.bg * p.nn, H, W,
                        sums + (long long)p.b0 * 2 * H * W, accumulate || p.n0 > 0, nullptr, s);
  };
  int rc;
  if (!(tc && net->ha_overlap && P > 1)) {
    for (int c = 0; c < P; ++c) {
      const Pass p = pass(c);
      if ((rc = do_warp(c, st)) != SPB_OK) return rc;
      if ((rc = ha_net_stage(net, bufs(c), p.bg * p.nn, H, W, st)) != SPB_OK) return rc;
      if ((rc = do_agg(c, st)) != SPB_OK) return rc;
    }
    return SPB_OK;
  }
  // Overlapped passes (opt-in, spb_net_set_ha_overlap): the image warp and the aggregate leave the tensor cores
  // idle, and a 256-thread CTA of either fits into the registers the persistent conv kernels leave free on every SM.
  // Measured in round 1: +1.5 % at best under the power cap -- off by default.
  // Side stream:  warp(1) | agg(0) warp(2) | agg(1) warp(3) | ...   main stream: warp(0), then the detector net of
  // every pass.  Buffers alternate by pass parity (the padded source images are shared: the side stream is ordered,
  // and the main stream's only use of them is warp(0) before anything else); the hazards are
  //   net(c)    after warp(c)                      (ev_warp)    and after agg(c-2) released probs[c&1]   (ev_agg)
  //   agg(c)    after net(c)                       (ev_fwd)
  //   warp(c+2) after net(c) read warped[c&1] and agg(c) read bits[c&1]   (side-stream order behind agg(c))
  // and the sums of one image are accumulated by agg(c), agg(c+1), ... in side-stream order, as before.
  if (!net->aux) {
    SPB_CHECK_CUDA(cudaStreamCreateWithFlags(&net->aux, cudaStreamNonBlocking));
    SPB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_begin, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      SPB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_warp[i], cudaEventDisableTiming));
      SPB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_fwd[i], cudaEventDisableTiming));
      SPB_CHECK_CUDA(cudaEventCreateWithFlags(&net->ev_agg[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t sx = net->aux;
  if ((rc = do_warp(0, st)) != SPB_OK) return rc;
  SPB_CHECK_CUDA(cudaEventRecord(net->ev_begin, st));  // everything the caller queued before this call + warp(0)
  SPB_CHECK_CUDA(cudaStreamWaitEvent(sx, net->ev_begin, 0));
  if ((rc = do_warp(1, sx)) != SPB_OK) return rc;
  SPB_CHECK_CUDA(cudaEventRecord(net->ev_warp[1], sx));
  for (int c = 0; c < P; ++c) {
    const Pass p = pass(c);
    if (c >= 1) SPB_CHECK_CUDA(cudaStreamWaitEvent(st, net->ev_warp[c & 1], 0));
    if (c >= 2) SPB_CHECK_CUDA(cudaStreamWaitEvent(st, net->ev_agg[c & 1], 0));
    if ((rc = ha_net_stage(net, bufs(c), p.bg * p.nn, H, W, st)) != SPB_OK) return rc;
    SPB_CHECK_CUDA(cudaEventRecord(net->ev_fwd[c & 1], st));
    SPB_CHECK_CUDA(cudaStreamWaitEvent(sx, net->ev_fwd[c & 1], 0));
    if ((rc = do_agg(c, sx)) != SPB_OK) return rc;
    SPB_CHECK_CUDA(cudaEventRecord(net->ev_agg[c & 1], sx));
    if (c + 2 < P) {
      if ((rc = do_warp(c + 2, sx)) != SPB_OK) return rc;
      SPB_CHECK_CUDA(cudaEventRecord(net->ev_warp[c & 1], sx));
    }
  }
  SPB_CHECK_CUDA(cudaStreamWaitEvent(st, net->ev_agg[(P - 1) & 1], 0));  // the call completes in the caller's stream
  return SPB_OK;
}

// Homography sharding: a flat list of V views with a different number of views per image and explicit matrix
// indices (ONE pass: the caller cuts a batch into passes of at most spb_ha_pass_views views, sharding.plan_batch).
size_t spb_ha_workspace_bytes_ragged(int B, int V, int max_views_per_image, int H, int W) {
  if (B < 1 || V < 1 || max_views_per_image < 1 || H < 8 || W < 8) return 0;
  return ha_layout(B, (size_t)V, max_views_per_image, H, W, false, false).total;
}

int spb_ha_heatmap_sums_ragged(spb_net* h, const float* imgs, const float* mats_unwarp, const float* mats_fwd,
                               const int* view_image, const int* view_mat, const int* view_offsets, int B, int V,
                               int max_views_per_image, int H, int W, float* sums, int accumulate, void* ws,
                               size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(h && imgs && mats_unwarp && mats_fwd && view_image && view_offsets && sums && ws);
  SPB_CHECK_ARG(B >= 1 && V >= 1 && max_views_per_image >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  Net* net = reinterpret_cast<Net*>(h);
  if (net->impl != SPB_IMPL_TCGEN05) {
    set_error("spb_ha_heatmap_sums_ragged: tcgen05 implementation only");
    return SPB_EINVAL;
  }
  if (ws_bytes < spb_ha_workspace_bytes_ragged(B, V, max_views_per_image, H, W)) {
    set_error("spb_ha_heatmap_sums_ragged: workspace %zu < %zu bytes", ws_bytes,
              spb_ha_workspace_bytes_ragged(B, V, max_views_per_image, H, W));
    return SPB_ENOMEM;
  }
  const HaLayout L = ha_layout(B, (size_t)V, max_views_per_image, H, W, false, false);
  char* base = static_cast<char*>(align_ptr(ws));
  PassBufs pb;
  pb.warped = reinterpret_cast<float*>(base + L.warp);
  pb.bits = reinterpret_cast<uint8_t*>(base + L.bits);
  pb.probs = base + L.probs;
  pb.semi = nullptr;
  pb.fold = reinterpret_cast<FoldMat*>(base + L.fold);
  pb.src_pad = reinterpret_cast<float*>(base + L.src_pad);
  pb.agg_ws = base + L.agg;
  pb.net_ws = base + L.net;
  ViewMap vm;
  vm.view_image = view_image;
  vm.view_mat = view_mat;  // NULL: the matrices are listed in view order
  vm.vpi = 1;
  vm.shared = 0;
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = ha_warp_stage(net, pb, imgs, mats_unwarp, mats_fwd, vm, B, V, H, W, st)) != SPB_OK) return rc;
  if ((rc = ha_net_stage(net, pb, V, H, W, st)) != SPB_OK) return rc;
  return ha_agg_stage(net, pb, mats_unwarp, vm, B, max_views_per_image, V, H, W, sums, accumulate, view_offsets, st);
}

// The image side of the HA branch on its own (datasets/Coco.py:279-282): V warped views + their valid-mask bits with
// the kernels of the fused pass (folded coordinates: views within ~1e-3 of spb_inv_warp_image_batch, mask bits exact).
size_t spb_ha_warp_views_workspace_bytes(int nb, int V, int H, int W) {
  if (nb < 1 || V < 1 || H < 8 || W < 8) return 0;
  return align_up((size_t)nb * (H + 2 * kHaPad) * (W + 2 * kHaPad) * 4, 1024) + align_up(2 * (size_t)V * sizeof(FoldMat), 1024) +
         align_up(heat16_bytes(1, H, W), 1024) + 1024;
}

int spb_ha_warp_views(const float* imgs, const float* mats_fwd, const int* view_image, const int* view_mat, int nb,
                      int V, int H, int W, float* views, uint8_t* bits, void* ws, size_t ws_bytes, void* stream) {
  SPB_CHECK_ARG(imgs && mats_fwd && view_image && views && bits && ws);
  SPB_CHECK_ARG(nb >= 1 && V >= 1 && H >= 8 && W >= 8 && W % 8 == 0);
  if (ws_bytes < spb_ha_warp_views_workspace_bytes(nb, V, H, W)) {
    set_error("spb_ha_warp_views: workspace %zu < %zu bytes", ws_bytes, spb_ha_warp_views_workspace_bytes(nb, V, H, W));
    return SPB_ENOMEM;
  }
  char* base = static_cast<char*>(align_ptr(ws));
  float* src_pad = reinterpret_cast<float*>(base);
  base += align_up((size_t)nb * (H + 2 * kHaPad) * (W + 2 * kHaPad) * 4, 1024);
  FoldMat* fold = reinterpret_cast<FoldMat*>(base);
  base += align_up(2 * (size_t)V * sizeof(FoldMat), 1024);
  ViewMap vm;
  vm.view_image = view_image;
  vm.view_mat = view_mat;
  vm.vpi = 1;
  vm.shared = 0;
  cudaStream_t st = as_stream(stream);
  // (the un-warp fold and the plane border of ha_prep are not needed here: V = 0 planes, forward matrices twice)
  int rc = ha_prep(imgs, src_pad, mats_fwd, mats_fwd, fold, fold + V, base, vm, nb, V, H, W, st, /*planes=*/0);
  if (rc != SPB_OK) return rc;
  return warp_views(src_pad, fold, mats_fwd, views, bits, vm, V, H, W, st);
}

int spb_ha_heatmap_sums(spb_net* h, const float* img, const float* mats_unwarp, const float* mats_fwd, int N, int H,
                        int W, float* sums, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  return spb_ha_heatmap_sums_batch(h, img, mats_unwarp, mats_fwd, 1, 1, N, H, W, sums, accumulate, ws, ws_bytes, stream);
}

int spb_net_profile(spb_net* h, int on) {
  SPB_CHECK_ARG(h);
  Net* net = reinterpret_cast<Net*>(h);
  net->profile = on ? 1 : 0;
  net->usedrecs = 0;
  return SPB_OK;
}

int spb_net_profile_read(spb_net* h, float* ms, int* counts) {
  SPB_CHECK_ARG(h && ms && counts);
  Net* net = reinterpret_cast<Net*>(h);
  for (int i = 0; i < kNumStages; ++i) {
    ms[i] = 0.f;
    counts[i] = 0;
  }
  for (int i = 0; i < net->usedrecs; ++i) {
    ProfRec& r = net->recs[i];
    SPB_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    SPB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.stage] += t;
    counts[r.stage]++;
  }
  net->usedrecs = 0;
  return SPB_OK;
}

int spb_net_set_ha_overlap(spb_net* h, int on) {
  SPB_CHECK_ARG(h);
  reinterpret_cast<Net*>(h)->ha_overlap = on ? 1 : 0;
  return SPB_OK;
}

int spb_net_set_ha_chunk(spb_net* h, int views) {
  SPB_CHECK_ARG(h && views >= 0);
  reinterpret_cast<Net*>(h)->ha_chunk = views;
  return SPB_OK;
}

}  // extern "C"
