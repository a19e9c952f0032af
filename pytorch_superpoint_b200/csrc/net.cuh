// Network state shared by net.cu (C ABI), conv_simt.cu and conv_tc.cu.
#pragma once
#include "common.cuh"

namespace spb {

constexpr int kNumLayers = 12;  // conv1a,1b,2a,2b,3a,3b,4a,4b,Pa,Pb,Da,Db

struct LayerDev {
  int cin, cout, ks, relu, pool;
  int cout_pad;                // multiple of 16 (65 -> 80): SIMT quarter split and UMMA N
  float* bias;                 // fp32 [cout_pad], zero padded
  __nv_bfloat16* w_simt;       // bf16 [tap][cin][cout_pad]
  __nv_bfloat16* w_tc;         // bf16 [cout_pad][tap*cin]   (K-major rows for the UMMA B operand)
  float* w1a_f32;              // conv1a only: bf16-rounded weights as fp32 [9][64]
  void* tmap_w;                // host copy of the CUtensorMap for w_tc (conv_tc.cu), 128 B
  __nv_bfloat16* w_s3;         // 64->64 3x3 layers: bf16 [r][s][co][ci] (conv_s3.cu: column taps stacked along N)
  void* tmap_w3;               // host copy of the CUtensorMap for w_s3
};

constexpr int kNumStages = 16;  // 0..11 layers, 12 warp, 13 aggregate

struct ProfRec {
  int stage;
  cudaEvent_t a, b;
};

struct Net {
  LayerDev layer[kNumLayers];
  int impl;
  int ha_chunk;  // views per pass in spb_ha_heatmap_sums (0 = all)
  int profile;   // record CUDA events around every stage launch (bench / roofline only)
  ProfRec* recs;
  int nrecs, caprecs, usedrecs;
  // spb_ha_heatmap_sums_batch with more than one pass: image warp of pass c+1 and the aggregate of pass c-1 run on
  // this stream, under the detector net of pass c (created on first use)
  int ha_overlap;  // 0 (default) / 1
  cudaStream_t aux;
  cudaEvent_t ev_begin, ev_warp[2], ev_fwd[2], ev_agg[2];
};

// conv_simt.cu
int conv1a_simt(const LayerDev& L, const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st);
int conv_simt(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_f32, cudaStream_t st);
int l2norm_rows(float* d, long long rows, int C, cudaStream_t st);

// warp.cu / heat.cu pieces used by the fused HA path (CUDA-core validation path, SPB_IMPL_SIMT)
int warp_with_mask_bits(const float* img, const float* mats, float* out, uint8_t* bits, int N, const ViewMap& vm,
                        int H, int W, cudaStream_t st);
int aggregate_probs_groups(int B, int N, int H, int W);
size_t aggregate_probs_workspace(int B, int N, int H, int W);
int aggregate_probs(const float* probs, const uint8_t* bits, const float* mats_unwarp, const ViewMap& vm, int B, int N,
                    int H, int W, float* sums, int accumulate, void* ws, cudaStream_t st, const int* offs = nullptr);
int reduce_partials(const float* part, int G, int nb, int H, int W, float* sums, int accumulate, cudaStream_t st);

// ha_fast.cu: the non-tensor kernels of the tcgen05 HA pass (folded homographies, padded frames, fp16 planes)
int ha_prep(const float* imgs, float* src_pad, const float* mats_fwd, const float* mats_unwarp, FoldMat* fold_fwd,
            FoldMat* fold_unwarp, void* heat16, const ViewMap& vm, int nb, int V, int H, int W, cudaStream_t st,
            int planes = -1);
int warp_views(const float* src_pad, const FoldMat* fold, const float* mats_fwd, float* out, uint8_t* bits,
               const ViewMap& vm, int V, int H, int W, cudaStream_t st);
int aggregate_heat16(const void* heat16, const FoldMat* fold_unwarp, int nb, int N, int H, int W, float* sums,
                     int accumulate, void* ws, cudaStream_t st, const int* offs);

// conv_tc.cu (tcgen05 / TMEM / TMA)
int conv1a_tc(const LayerDev& L, const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st);
bool conv1_fused_enabled();
bool conv1_fused_supports(int H, int W);  // even H, W % 4 == 0 (TMA row pitch)
int conv1_fused(const LayerDev& L1a, const LayerDev& L1b, const float* x, __nv_bfloat16* out, int V, int H, int W,
                cudaStream_t st);
int conv_tc_prepare(LayerDev& L);  // builds the weight tensor map after w_tc is uploaded
// out_mode: 0 bf16 NHWC, 1 fp32 rows (cout channels), 2 fp32 rows L2-normalised, 4 = convPb in the HA pass: softmax
// probabilities (dustbin dropped) as fp16 planes
// [B, 8H + 2*kHeatPadY, 8W + 2*kHeatPadX] (softmax, dustbin drop, depth-to-space; aux = the valid-mask bits
// [B,8H,W] of the warped views: masked pixels are stored as kHeatMasked)
int conv_tc(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode, cudaStream_t st,
            const void* aux = nullptr);


// conv_tc2.cu (CTA pairs, tcgen05.mma.cta_group::2, for the >= 128-output-channel 3x3 layers)
int conv_tc2(const LayerDev& L, const __nv_bfloat16* in, void* out, int B, int H, int W, int out_mode, cudaStream_t st,
             bool* handled);

// conv_s3.cu (64 -> 64, N = 192 column-tap stacking)
bool conv_s3_enabled();
int conv_s3_prepare(LayerDev& L);  // after w_s3 is uploaded
int conv_s3(const LayerDev& L, const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st);
long long* conv1_probe_buf();  // timeline probe buffer set with spb_conv1_fused_set_probe, or NULL
}  // namespace spb
