/* spb200 -- C ABI of the B200-native SuperPoint homography-adaptation / extraction hot path.
 *
 * One shared library (libspb200.so, sm_100a only).  Every entry point is `extern "C"`, takes plain
 * pointers and sizes (no torch types), returns 0 on success or a negative SPB_E* code and never
 * throws; `spb_last_error_string()` describes the last failure on the calling thread.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the comment says HOST;
 *   - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream); kernels are only
 *     enqueued, never synchronised; nothing is allocated after `spb_net_create`;
 *   - images are row-major fp32 [H,W]; homographies are row-major fp32 [N,3,3] acting on the
 *     reference's normalised coordinates (u,v in [-1,1], `torch.linspace(-1,1,W)` per pixel);
 *   - the caller owns every buffer, including workspaces (sizes from the *_workspace_bytes calls).
 *
 * Each declaration cites the reference interface (file:line under eric-yyjau/pytorch-superpoint)
 * that it replaces.  INTEGRATION.md shows the reference-side ctypes binding.
 */
#ifndef SPB200_H
#define SPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_OK 0
#define SPB_EINVAL (-1)   /* bad argument (null pointer, unsupported size) */
#define SPB_ECUDA (-2)    /* a CUDA runtime/driver call failed              */
#define SPB_ENOMEM (-3)   /* workspace too small / allocation failed        */
#define SPB_EARCH (-4)    /* device is not sm_100                           */

#define SPB_MODE_BILINEAR 0
#define SPB_MODE_NEAREST 1

#define SPB_IMPL_TCGEN05 0 /* bf16 implicit-GEMM on tcgen05/TMEM fed by TMA (product path)        */
#define SPB_IMPL_SIMT 1    /* CUDA-core bring-up/validation kernels with the same bf16 numerics  */

int spb_version(void);
const char* spb_last_error_string(void);
/* number of CUDA kernels this library has launched in this process (all entry points) */
unsigned long long spb_launch_count(void);

/* ---- (1) homography warp ------------------------------------------------------------------------
 * utils/utils.py:318-351 `inv_warp_image_batch(img, mat_homo_inv, device, mode)` incl. warp_points
 * (286-314) and F.grid_sample(align_corners=True, zero padding).
 *   out[n](p) = img[n](M[n].p).  img_batch_stride = elements between consecutive source images;
 *   0 broadcasts ONE image to all N warps (the HA case `img.repeat(N,1,1,1)`, datasets/Coco.py:279,
 *   never materialised). */
int spb_inv_warp_image_batch(const float* img, long long img_batch_stride, const float* mats, float* out,
                             int N, int H, int W, int mode, void* stream);

/* utils/utils.py:677-704 `compute_valid_mask(image_shape, inv_homography, device, erosion_radius=0)`:
 * nearest-mode warp of an all-ones image -> mask[N,H,W] in {0,1} (fp32, like the reference). */
int spb_compute_valid_mask(const float* mats, float* mask, int N, int H, int W, void* stream);

/* utils/utils.py:699-702: the erosion of those masks (`cv2.erode(mask, getStructuringElement(MORPH_ELLIPSE,
 * (2r, 2r)), iterations=1)`, datasets/Coco.py:281-283 `valid_border_margin`).  Generic form: the structuring element
 * is given as one column span [j1, j2) per row (spans: DEVICE int[kh][2]) with its anchor; pixels outside the image
 * are ignored (cv2's default erosion border).  in and out [N,H,W] fp32, out != in. */
int spb_erode(const float* in, float* out, int N, int H, int W, const int* spans, int kh, int anchor_y, int anchor_x,
              void* stream);

/* ---- (3) heat-map ---------------------------------------------------------------------------------
 * utils/utils.py:480-525 `flattenDetection(semi, tensor=True)` + utils/d2s.py:8-25 DepthToSpace(8):
 * softmax over 65 channels, drop the dustbin, pixel-shuffle.  semi is NCHW fp32 [B,65,Hc,Wc]
 * (the reference layout) when nhwc_cstride == 0, else NHWC with that channel stride (>= 65).
 * heat is [B,8Hc,8Wc]. */
int spb_flatten_detection(const float* semi, int nhwc_cstride, float* heat, int B, int Hc, int Wc, void* stream);

/* export.py:49-63 `combine_heatmap(heatmap, inv_homographies, mask_2D, device)` on materialised
 * heat/mask [N,H,W]: out[H,W] = sum_n unwarp_n(heat_n*mask_n) / sum_n unwarp_n(mask_n). */
int spb_combine_heatmap(const float* heat, const float* mask, const float* mats_unwarp, float* out,
                        int N, int H, int W, void* stream);

/* The fused form of utils.py:480-525 + utils.py:677-704 + export.py:49-62 used by the HA path:
 * semi (NHWC fp32, channel stride cstride) of N warped views -> per-cell softmax, dustbin drop,
 * depth-to-space, analytic valid mask (from mats_fwd), bilinear un-warp (mats_unwarp) and the two
 * sums over the N views.  sums[2,H,W] = {sum heat*mask, sum mask}; accumulate != 0 adds to the
 * existing content (used when the homographies of one image arrive in several chunks).
 * ws must hold spb_ha_aggregate_workspace_bytes(N,H,W). */
size_t spb_ha_aggregate_workspace_bytes(int N, int H, int W);
int spb_ha_aggregate(const float* semi_nhwc, int cstride, const float* mats_unwarp, const float* mats_fwd,
                     int N, int H, int W, float* sums, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* export.py:63 the final division (after the NCCL all-reduce of `sums` when N is sharded). */
int spb_ha_finalize(const float* sums, float* heat, int H, int W, void* stream);
int spb_ha_finalize_batch(const float* sums /*[B,2,H,W]*/, float* heat /*[B,H,W]*/, int B, int H, int W, void* stream);

/* ---- the exchange of the sharded HA path over NVLink peer memory (replaces nn.DataParallel's gather, export.py:265, and
 * fuses the reduction over the ranks with export.py:61-63).  One process per GPU of ONE box: every rank keeps its
 * partial planes sums[B,2,H,W] in a buffer from spb_peer_alloc (cudaMalloc, zeroed), exports a 64-byte CUDA IPC handle
 * that the others open (spb_peer_open: the peer's buffer mapped into this process).
 * spb_ha_finalize_peers: heat[b] = (sum_p peer_sums[p][b0+b][0]) / (sum_p peer_sums[p][b0+b][1]) for b < nb, the
 *   `world` partial planes loaded from peer memory and added in rank order (reduce-scatter + division in one kernel).
 * spb_peer_signal: after everything enqueued on `stream` so far, store `value` to the n flag addresses (peer memory;
 *   NULL entries are skipped) with system-scope release.  spb_peer_wait: hold `stream` until the n LOCAL flags are
 *   >= value (bit t of skip_mask set: flag t is not waited for).  Counters are monotonic per (peer, slot). */
int spb_peer_alloc(void** ptr, size_t bytes);
int spb_peer_free(void* ptr);
int spb_peer_export(void* ptr, unsigned char* handle64);
int spb_peer_open(const unsigned char* handle64, void** ptr);
int spb_peer_close(void* ptr);
int spb_peer_signal(unsigned int* const* flags, int n, unsigned int value, void* stream);
int spb_peer_wait(const unsigned int* flags, int n, unsigned int value, unsigned int skip_mask, void* stream);
int spb_ha_finalize_peers(const float* const* peer_sums, int world, int b0, int nb, float* heat, int H, int W,
                          void* stream);

/* ---- (4) threshold + greedy grid NMS + border removal + top-k -----------------------------------
 * models/model_wrap.py:252-279 `getPtsFromHeatmap` incl. `nms_fast` (117-180) and export.py:322-328
 * (`pts[:top_k]`).  heat [B,H,W].  For image b, pts[b, j, :] = (x, y, conf) for j < counts[b],
 * ordered like the reference (confidence descending; ties: larger row-major index first), cut to
 * min(top_k, capacity) rows (top_k <= 0: no cut).  Exact: identical index sets and order to the
 * reference run with a stable argsort.  total[b] (may be NULL) receives the un-cut survivor count. */
size_t spb_nms_workspace_bytes(int B, int H, int W, int nms_dist);
int spb_get_pts_from_heatmap(const float* heat, int B, int H, int W, float conf_thresh, int nms_dist,
                             int border_remove, int top_k, int capacity, float* pts, int* counts, int* total,
                             void* ws, size_t ws_bytes, void* stream);
/* The same with the CTA size of the NMS kernel chosen by the caller: 0 = automatic (1024 threads; 512 when the batch
 * has more clusters than SMs, so that a second image runs on an SM while the first waits at its cluster barrier --
 * best when the launch has the GPU to itself), 1024 = for a launch that runs BESIDE other kernels (the exporter's side
 * stream under the convolutions of the next batch: smaller CTAs take the SMs away from them for longer). */
int spb_get_pts_from_heatmap_ex(const float* heat, int B, int H, int W, float conf_thresh, int nms_dist,
                                int border_remove, int top_k, int capacity, float* pts, int* counts, int* total,
                                void* ws, size_t ws_bytes, int cta_threads, void* stream);

/* ---- (5) descriptor sampling -----------------------------------------------------------------------
 * models/model_wrap.py:281-299 `sample_desc_from_points(coarse_desc, pts)`: bilinear
 * grid_sample(align_corners=True) of the coarse descriptor map at x*(Wc-1)/W-style coordinates
 * (the reference's normalisation quirk is kept), then L2 normalisation.
 *   coarse: fp32, NCHW [D,Hc,Wc] when nhwc == 0 (reference layout) or NHWC [Hc,Wc,D] when nhwc != 0.
 *   pts: [K,3] rows (x, y, conf) as produced by spb_get_pts_from_heatmap; count: DEVICE int* with the
 *   number of valid rows (NULL: all K rows).  out: [K,D] (row k = descriptor of point k). */
int spb_sample_desc_from_points(const float* coarse, int nhwc, const float* pts, const int* count, int K,
                                int Hc, int Wc, int D, int cell, float* out, void* stream);

/* Same with the point row stride (2 = (x, y), 3 = (x, y, conf)) and the final L2 normalisation optional: the batched
 * variant models/model_utils.py:63-90 (SuperPointNet_process.sample_desc_from_points) returns the raw samples. */
int spb_sample_desc_from_points_ex(const float* coarse, int nhwc, const float* pts, int pts_stride, const int* count,
                                   int K, int Hc, int Wc, int D, int cell, int normalize, float* out, void* stream);

/* One launch for a batch (the per-image loop of export.py:127-145; Val_model_heatmap.py:151-155 itself is only valid
 * for B = 1): coarse NHWC [B,Hc,Wc,D], pts [B,K,3], counts [B] (device) -> out [B,K,D], rows behind counts[b] zero. */
int spb_sample_desc_batch(const float* coarse_nhwc, const float* pts, const int* counts, int B, int K, int Hc, int Wc,
                          int D, int cell, int normalize, float* out, void* stream);

/* NEXT row f-3 -- models/model_utils.py:24-60 `SuperPointNet_process.pred_soft_argmax`: 5x5 roi_pool patches of the
 * 7x7 window around every labelled pixel (utils/losses.py:16-29,87-98 + torchvision roi_pool bin rule), log, spatial
 * soft-argmax in pixel coordinates (torchgeometry restated: parity unpinned).  heat [B,1,H,W] fp32, idx [N,4] int64
 * rows (batch, 0, y, x) as produced by labels.nonzero(); dxdy [N,2]; patches_or_null [N,25]. */
int spb_roi_soft_argmax(const float* heat, int B, int H, int W, const long long* idx, int N, float* dxdy,
                        float* patches_or_null, void* stream);

/* NEXT row f-4 (dense-descriptor mode) -- models/model_wrap.py:393-401: bilinear x`cell` up-sampling of the coarse
 * descriptors (F.interpolate, align_corners=False) fused with the channel L2 normalisation.
 * coarse NHWC [B,Hc,Wc,D] fp32 (D % 8 == 0, D <= 256) -> dense NCHW [B,D,Hc*cell,Wc*cell] fp32. */
int spb_dense_desc(const float* coarse_nhwc, int B, int Hc, int Wc, int D, int cell, float* dense_nchw, void* stream);

/* models/model_wrap.py:402-404: the dense map's columns at the integer key-points of every image (one launch):
 * dense NCHW [B,D,H,W], pts [B,K,3], counts [B] (device) -> out [B,K,D]; rows behind counts[b] are left untouched. */
int spb_gather_dense_desc(const float* dense_nchw, const float* pts, const int* counts, int B, int K, int D, int H, int W,
                          float* out, void* stream);

/* NEXT row f-1b -- models/model_wrap.py:200-234 `soft_argmax_points` (+ utils/losses.py:48-129): per key-point
 * soft-argmax offset over the patch_size x patch_size (odd, <= 7) heat-map patch.  dxdy [K,2] fp32; the
 * refined point is (x + dx - patch_size/2, y + dy - patch_size/2).  torchgeometry's SpatialSoftArgmax2d is
 * restated from its published definition (package not vendored by the reference): parity unpinned. */
int spb_soft_argmax_points(const float* heat, int H, int W, const float* pts, const int* count, int K,
                           int patch_size, float* dxdy, void* stream);
/* a batch in one launch: heat [B,H,W], pts [B,K,3], counts [B] (device), dxdy [B,K,2] */
int spb_soft_argmax_points_batch(const float* heat, int B, int H, int W, const float* pts, const int* counts, int K,
                                 int patch_size, float* dxdy, void* stream);

/* NEXT row f-1 -- models/model_wrap.py:437-480 `PointTracker.nn_match_two_way`: desc1 [D,K1], desc2 [D,K2] fp32
 * (unit-norm columns, row-major like the reference's M x N arrays).  L2 distance sqrt(2 - 2*clip(d1.d2)), row and
 * column argmin (ties -> lowest index, as np.argmin), keep i if score < nn_thresh and the match is mutual.
 * Outputs (device): idx1/idx2/score [>= min(K1,K2)] in ascending idx1 order, *count = L.  nn_thresh < 0 is
 * SPB_EINVAL (the reference raises ValueError); K1 == 0 or K2 == 0 gives count 0. */
size_t spb_nn_match_workspace_bytes(int K1, int K2);
int spb_nn_match_two_way(const float* desc1, int K1, const float* desc2, int K2, int D, float nn_thresh, int* idx1,
                         int* idx2, float* score, int* count, void* ws, size_t ws_bytes, void* stream);

/* The same for P image pairs in one launch and without a host round trip (export.py:147-151 per pair): desc1 / desc2
 * [P,K,D] in the descriptor sampler's point-major layout, rows behind an image's key-point count zero filled (a zero
 * row is sqrt(2) away from every unit descriptor: it can neither pass nn_thresh < sqrt(2) nor beat a real match).
 * idx1 / idx2 / score [P,K], count [P] (all device). */
size_t spb_nn_match_pairs_workspace_bytes(int P, int K);
int spb_nn_match_pairs(const float* desc1, const float* desc2, int P, int K, int D, float nn_thresh, int* idx1,
                       int* idx2, float* score, int* count, void* ws, size_t ws_bytes, void* stream);

/* NEXT row f-2 -- utils/homographies.py:12-141 `sample_homography_np` as called by datasets/Coco.py:262-276:
 * `total` homographies (image b owns [b*per_image, (b+1)*per_image); index 0 of every image is the identity when
 * per_image > 0), homographies = inv(sampled) and inv_homographies = inverse(homographies), both [total,3,3] fp32
 * device; raw_or_null optionally receives what sample_homography_np itself returns.  Same perturbation sequence and
 * parameter meaning as the reference (shape (2,2), shift -1); Philox counter RNG -> DISTRIBUTIONAL parity only. */
int spb_sample_ha_homographies(unsigned long long seed, int total, int per_image, int perspective, int scaling,
                               int rotation, int translation, int n_scales, int n_angles, double scaling_amplitude,
                               double perspective_amplitude_x, double perspective_amplitude_y, double patch_ratio,
                               double max_angle, int allow_artifacts, double translation_overflow,
                               float* homographies, float* inv_homographies, float* raw_or_null, void* stream);

/* NEXT row f-4 (image ingest) -- datasets/Coco.py:165-179 `_read_image` behind the JPEG decode:
 * `cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA)` -> `cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)` (applied to
 * imread's BGR memory order, as the reference does) -> `astype(float32) / 255`.  bgr: DEVICE uint8 [Hs,Ws,3]; out
 * fp32 [H,W] in [0,1].  Bit-exact against OpenCV 4.x: down-scaling (Hs >= H and Ws >= W) through its box-sum
 * arithmetic for integer factors and its per-axis cell tables otherwise; when an axis is enlarged (KITTI 375x1242 ->
 * 384x1248) through its 11-bit fixed-point linear variant with "area" coefficients. */
int spb_ingest_resize_gray(const uint8_t* bgr, int Hs, int Ws, float* out, int H, int W, void* stream);

/* ---- (2) network ------------------------------------------------------------------------------------
 * models/SuperPointNet_pretrained.py:46-76, SuperPointNet_gauss2.py:43-70, SuperPointNet.py:154-224.
 * weights/biases: HOST pointers, 12 layers in the order conv1a,1b,2a,2b,3a,3b,4a,4b,Pa,Pb,Da,Db;
 * weights fp32 OIHW with eval-mode BatchNorm already folded in (w*g/sqrt(v+eps)), biases likewise. */
typedef struct spb_net spb_net;
int spb_net_create(spb_net** net, const float* const* weights, const float* const* biases);
int spb_net_destroy(spb_net* net);
int spb_net_set_impl(spb_net* net, int impl); /* SPB_IMPL_* (default TCGEN05) */
/* views processed per pass by spb_ha_heatmap_sums* (0 = the default: bounded by bytes, 800 x 240 x 320 pixels of
 * views per pass); bounds the workspace */
int spb_net_set_ha_chunk(spb_net* net, int views);
/* spb_ha_heatmap_sums_batch with more than one pass: run the image warp of the next pass and the aggregate of the
 * previous one on an internal side stream under the detector net of the current pass (results are bit-identical,
 * the call still completes in the caller's stream).  Default 0 = strictly sequential passes: under the power cap the
 * overlap measured +1.5 % at best. */
int spb_net_set_ha_overlap(spb_net* net, int on);
/* Measurement hooks (bench.py roofline): record CUDA events on the launch stream around every stage
 * (0..11 = layers conv1a..convDb, 12 = image warp, 13 = fused aggregate) while on != 0;
 * spb_net_profile_read synchronises those events, returns summed milliseconds and launch counts per
 * stage (arrays of 16) and clears the log. */
int spb_net_profile(spb_net* net, int on);
int spb_net_profile_read(spb_net* net, float* ms, int* counts);
size_t spb_net_workspace_bytes(int B, int H, int W, int want_desc);
/* x [B,1,H,W] fp32 (H, W multiples of 8) -> semi NHWC fp32 [B,H/8,W/8,65];
 * desc NHWC fp32 [B,H/8,W/8,256], L2-normalised over channels (NULL: detector head only). */
int spb_net_forward(spb_net* net, const float* x, int B, int H, int W, float* semi, float* desc,
                    void* ws, size_t ws_bytes, void* stream);
/* Debug/validation: run ONE layer (index 0..11) on NHWC bf16 input (fp32 [B,H,W] for layer 0);
 * output NHWC bf16 (fp32 for Pb/Db; Db rows L2-normalised), 2x2 max-pool applied where the network has one. */
int spb_net_layer(spb_net* net, int layer, const void* in, void* out, int B, int H, int W, void* stream);

/* ---- train-mode BatchNorm (opt-in): the as-written reference never calls .eval() (models/model_wrap.py:111,
 * Val_model_heatmap.py:61-73), so its conv -> BN -> ReLU blocks (models/unet_parts.py:14-21,
 * SuperPointNet_gauss2.py:56-69) use the statistics of the current batch.  The host walks the layers with the RAW
 * conv weights (no folded BN), everything in fp32: convolution -> batch statistics -> normalise + ReLU + pool.
 * spb_conv_f32: "same" convolution on CUDA cores, in NHWC fp32 [B,H,W,cin], w fp32 [ks*ks][cin][cout_pad] (zero
 *   padded), bias [cout_pad], out fp32 rows [B*H*W, cout]; (ks, cout_pad) in 3x3: 64 / 128 / 256, 1x1: 80 / 256.
 * spb_bn_batch_stats: rows [R,C] -> mean, BIASED variance (F.batch_norm(training=True)) and scale = gamma /
 *   sqrt(var + eps), shift = beta - mean * scale (gamma / beta NULL: 1 / 0; mean / var may be NULL); C <= 256.
 * spb_bn_apply: y = x * scale[c] + shift[c], optional ReLU, optional 2x2 max-pool; out = bf16 NHWC (out_f32 == 0) or
 *   fp32 rows; l2norm != 0: channel L2 normalisation of the fp32 rows (the descriptor head). */
int spb_conv_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int cin, int cout,
                 int cout_pad, int ks, void* stream);
size_t spb_bn_batch_stats_workspace_bytes(int C);
int spb_bn_batch_stats(const float* rows, long long R, int C, const float* gamma, const float* beta, float eps,
                       float* mean, float* var, float* scale, float* shift, void* ws, size_t ws_bytes, void* stream);
int spb_bn_apply(const float* rows, int B, int H, int W, int C, const float* scale, const float* shift, int relu,
                 int pool, void* out, int out_f32, int l2norm, void* stream);

/* ---- the whole HA step for one image (datasets/Coco.py:262-284 + export.py:309-310) -------------
 * img [H,W] fp32; mats_fwd = sample['inv_homographies'] (warps the image, defines the valid mask),
 * mats_unwarp = sample['homographies'] (un-warps the heat-map); both [N,3,3] for the N views THIS
 * caller owns (all of them on one GPU, a shard under homography sharding).
 * Does: warp N views -> detector network -> fused aggregate into sums[2,H,W]. */
size_t spb_ha_workspace_bytes(int N, int H, int W);
int spb_ha_heatmap_sums(spb_net* net, const float* img, const float* mats_unwarp, const float* mats_fwd,
                        int N, int H, int W, float* sums, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* Same for a batch of B images in one set of launches (export.py's per-image loop, batched so that a shard of
 * 12-13 views per image still fills the GPU): imgs [B,H,W]; matrices [N,3,3] shared by all images
 * (shared_mats != 0) or [B,N,3,3]; sums [B,2,H,W].  Views are processed in passes of at most
 * spb_ha_pass_views(net, H, W) views. */
size_t spb_ha_workspace_bytes_batch(spb_net* net, int B, int N, int H, int W);
int spb_ha_heatmap_sums_batch(spb_net* net, const float* imgs, const float* mats_unwarp, const float* mats_fwd,
                              int shared_mats, int B, int N, int H, int W, float* sums, int accumulate, void* ws,
                              size_t ws_bytes, void* stream);

/* The image side of the HA branch on its own -- datasets/Coco.py:279-282 (`inv_warp_image_batch` of the repeated image
 * + `compute_valid_mask`) with the kernels of the fused pass: imgs [nb,H,W]; view v is image view_image[v] (device
 * int[V]) warped by mats_fwd[view_mat[v]] (view_mat NULL: mats_fwd[v]).  views [V,H,W] fp32 (folded coordinates:
 * within ~1e-3 of spb_inv_warp_image_batch), bits [V,H,W/8]: bit (x % 8) of byte x / 8 = the valid mask of pixel x,
 * EXACTLY spb_compute_valid_mask.  W % 8 == 0. */
size_t spb_ha_warp_views_workspace_bytes(int nb, int V, int H, int W);
int spb_ha_warp_views(const float* imgs, const float* mats_fwd, const int* view_image, const int* view_mat, int nb,
                      int V, int H, int W, float* views, uint8_t* bits, void* ws, size_t ws_bytes, void* stream);

/* Homography sharding (pytorch_superpoint_b200/sharding.py; replaces nn.DataParallel, export.py:265): a rank's share
 * of 100 views over 8 ranks is 12 or 13 per image and the larger shares rotate over the images of a batch, so a pass
 * is a flat list of V views: view v belongs to image view_image[v] (device int[V], grouped by image in ascending
 * order, relative to imgs), uses the matrices mats_*[view_mat[v]] (device int[V]; NULL: the matrices are listed in
 * view order) and image b owns the views [view_offsets[b], view_offsets[b+1]) (device int[B+1]); sums [B,2,H,W].
 * ONE pass: V is bounded by the workspace the caller provides; spb_ha_pass_views(net, H, W) is the number of views
 * the library itself puts into a pass (bounded by bytes: 800 at 240x320, 200 at 480x640, 128 at 384x1248; or the
 * spb_net_set_ha_chunk value). */
int spb_ha_pass_views(spb_net* net, int H, int W);
size_t spb_ha_workspace_bytes_ragged(int B, int V, int max_views_per_image, int H, int W);
int spb_ha_heatmap_sums_ragged(spb_net* net, const float* imgs, const float* mats_unwarp, const float* mats_fwd,
                               const int* view_image, const int* view_mat, const int* view_offsets, int B, int V,
                               int max_views_per_image, int H, int W, float* sums, int accumulate, void* ws,
                               size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPB200_H */
