/*
 * ihpr_b200.h -- C-ABI of the B200-native integral-regression (3D soft-argmax) hot path.
 *
 * Every entry point replaces a Python call site of the reference (the reference has no FFI layer of
 * its own: the seam is the three callables of common/nets/loss.py and main/model.py; INTEGRATION.md
 * shows the ctypes binding a maintainer adds).  Citations are relative to /root/reference/.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers unless the
 *     function name ends in _host.
 *   - heat is the head's output (B, J*D, H, W), NCHW-contiguous, channel c = j*D + d
 *     (common/nets/loss.py:16,18).  Joint-volume r = b*J + j is the contiguous run of
 *     N = D*H*W elements starting at r*N.  dtype: IHPR_F32 or IHPR_BF16.
 *   - coords are (B, J, 3) fp32 in (x, y, z) order, 0-based voxel units (loss.py:28-32).
 *   - stats are (B, J, 2) fp32: {m = max_i h_i, l = sum_i exp(h_i - m)}; logsumexp = m + ln l.
 *     The backward recomputes the softmax from heat + stats; the softmax is never materialised.
 *   - the caller owns every buffer, including the workspace (size from ihpr_workspace_bytes, must be
 *     zero-filled ONCE before first use; the kernels leave its tickets zeroed again).  The device entry points allocate
 *     nothing, free nothing and keep no pointer after they return.  The one exception is ihpr_integral_l1_fwd_bwd_host,
 *     which takes HOST buffers and therefore owns (and caches, per device) the device staging buffers, streams and events
 *     it copies through; ihpr_host_release frees them.  The ticket layout
 *     depends on B*J: re-zero a workspace before reusing it with a different B*J.  A workspace must not
 *     be shared by launches that may run concurrently (one per stream).
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the device
 *     except the *_host entry point, which is synchronous by contract.
 *   - return 0 on success, a negative IHPR_E* code otherwise; ihpr_last_error() gives the text
 *     (thread-local).  No entry point ever falls back to a CPU implementation.
 *   - NaN policy is the reference's: a row of all -inf gives NaN, +inf or NaN inputs propagate NaN
 *     to that row's outputs.
 *   - thread-safe: the reference calls its criterion from N Python threads, one per GPU
 *     (common/nets/balanced_parallel.py:149-173).  Mutable state is per thread (last error, launch count, kernel variant:
 *     two threads may hold different variants) or guarded by a mutex (the *_host entry's per-device staging cache, the
 *     per-device kernel-choice cache filled by the first call of ihpr_integral_l1_fwd_bwd at a new shape).
 */
#ifndef IHPR_B200_H_
#define IHPR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IHPR_VERSION 100            /* major*10000 + minor*100 + patch */

#define IHPR_OK        0
#define IHPR_EINVAL   -1            /* bad shape / null pointer / misaligned buffer / small workspace */
#define IHPR_EARCH    -2            /* device is not sm_100 */
#define IHPR_ECUDA    -3            /* CUDA runtime / launch failure; see ihpr_last_error() */

#define IHPR_F32       0
#define IHPR_BF16      1

int         ihpr_version(void);
const char *ihpr_last_error(void);

/* Bytes of caller-owned, zero-initialised scratch the kernels need for this problem size. */
size_t ihpr_workspace_bytes(int B, int J, int D, int H, int W);

/* soft_argmax(heatmaps, joint_num) -> (B,J,3)            replaces common/nets/loss.py:13-34
 * (called from main/test.py:65,72, main/main.py:57, main/mpii_s3d.py:70, main/up3d_s3d.py:67).
 * stats may be NULL (inference, torch.no_grad path of main/test.py:53). */
int ihpr_softargmax3d_fwd(const void *heat, int dtype, int B, int J, int D, int H, int W,
                          float *coords, float *stats,
                          void *workspace, size_t workspace_bytes, void *stream);

/* autograd backward of soft_argmax (implicit in the reference: loss.backward(), main/train.py:71):
 * grad_heat[i] = p_i * sum_c grad_coords_c * (c(i) - coords_c), written in heat's dtype. */
int ihpr_softargmax3d_bwd(const void *heat, int dtype, int B, int J, int D, int H, int W,
                          const float *coords, const float *stats, const float *grad_coords,
                          void *grad_heat, void *stream);

/* JointLocationLoss.forward(heatmap_out, gt_coord, gt_vis, gt_have_depth) -> scalar
 *                                                         replaces common/nets/loss.py:36-52
 * (called from main/train.py:67 through DataParallelCriterion, common/base.py:71).
 * gt (B,J,3), vis (B,J,1), have_depth (B,1), all fp32.  Writes loss[0], coords and stats. */
int ihpr_integral_l1_fwd(const void *heat, int dtype, int B, int J, int D, int H, int W,
                         const float *gt, const float *vis, const float *have_depth,
                         float *loss, float *coords, float *stats,
                         void *workspace, size_t workspace_bytes, void *stream);

/* backward of the above: grad_out is a DEVICE pointer to the scalar d(objective)/d(loss)
 * (autograd's incoming gradient; 1.0 for loss.backward()).  Fuses
 * g_c = grad_out * sign(coord_c - gt_c) * vis * w_c / (3*B*J)   (loss.py:49-52)
 * with the soft-argmax backward; writes grad_heat in heat's dtype. */
int ihpr_integral_l1_bwd(const void *heat, int dtype, int B, int J, int D, int H, int W,
                         const float *coords, const float *stats,
                         const float *gt, const float *vis, const float *have_depth,
                         const float *grad_out, void *grad_heat, void *stream);

/* Training step in ONE launch: JointLocationLoss forward (as ihpr_integral_l1_fwd) plus d loss / d heat for an
 * upstream gradient of 1, i.e. what main/train.py:67-71 (criterion + loss.backward()) produces.  Each
 * joint-volume is streamed from HBM once and re-read from L2 (DRAM traffic 2 N s instead of 3 N s) by a cooperative
 * launch whose partner CTAs wait for each other (K5); variant 7 selects K5c instead, which keeps the joint-volume in
 * the shared memory of a thread-block cluster between the two passes (exactly 2 N s of DRAM traffic, but clusters of
 * 8 / 16 CTAs leave 28 / 36 of a B200's 148 SMs idle, so it is slower and not the default).  Falls back to the
 * two-kernel sequence for small batches and for shapes only the scalar kernels handle -- same results to rounding
 * either way.  With variant 0 (auto) the choice between the one-launch form and K1 + K2 is MEASURED: the first eager call on a
 * device at a new shape runs both forms (2 + 3 extra passes, ONE cudaEventSynchronize on the caller's stream -- the only
 * synchronisation a device entry point ever makes) and every later call at that shape runs the faster one; a call on a stream
 * under CUDA-graph capture never measures (it uses the one-launch form until an eager call has decided).  IHPR_CALIBRATE=0
 * or variant 8 keep the static rule (the one-launch form whenever it applies); variant 9 forces K1 + K2. */
int ihpr_integral_l1_fwd_bwd(const void *heat, int dtype, int B, int J, int D, int H, int W,
                             const float *gt, const float *vis, const float *have_depth,
                             float *loss, float *coords, float *stats, void *grad_heat,
                             void *workspace, size_t workspace_bytes, void *stream);

/* grad_heat *= *grad_out (device scalar), a no-op launch when *grad_out == 1: turns the unit-gradient result of
 * ihpr_integral_l1_fwd_bwd into autograd's answer for an arbitrary upstream gradient.  n = B*J*D*H*W elements. */
int ihpr_scale_grad(void *grad_heat, int dtype, size_t n, const float *grad_out, void *stream);

/* JointLocationLoss.forward (common/nets/loss.py:49-52) on coordinates that already exist -- the companion of
 * ihpr_head_softargmax_fwd, which produces coords without a heat-map:
 *   loss = mean_{b,j} (|dx| + |dy| + |dz| * have_depth[b]) * vis[b,j] / 3.   One launch, fixed summation order. */
int ihpr_integral_l1_from_coords(const float *coords, const float *gt, const float *vis, const float *have_depth,
                                 int B, int J, float *loss, void *stream);

/* Training-sample preparation of DatasetLoader.__getitem__ (data/dataset.py:84-152) for a whole batch on the device.
 *
 * ihpr_augment_patches -- generate_patch_image (dataset.py:201-221: optional horizontal flip, cv2.warpAffine with
 * INTER_LINEAR and a constant zero border, BGR -> RGB, float32), the colour scale + clip of dataset.py:92-93 and the
 * ToTensor + Normalize transform of common/base.py:93-95, i.e. out = (clip(patch * color_scale, 0, 255) - mean) / std
 * (the reference does NOT divide by 255: ToTensor only rescales uint8 input).  The warp restates OpenCV's fixed-point
 * algorithm for uint8 images: patches are bit-identical to the reference's.
 *   images (B, Hs, Ws, 3) uint8 BGR as cv2.imread returns them, padded to a common Hs x Ws; sizes (B,2) int32 = valid
 *   rows, cols of each image; trans (B,6) double = the forward 2x3 patch transform of gen_trans_from_patch_cv
 *   (dataset.py:229-257) computed for the ALREADY MIRRORED box centre when do_flip[b] != 0; color_scale (B,3) in RGB
 *   order; pixel_mean / pixel_std: HOST pointers to 3 floats (main/config.py:31-32); out (B,3,out_h,out_w) fp32, or
 *   (B,out_h,out_w,3) when channels_last != 0.  All other pointers are device pointers.
 *
 * ihpr_augment_joints -- dataset.py:96-135,148-149: mirror x and swap left/right joints when flipped (flip_perm: the
 * pair swaps as one permutation, nullable), trans_point2d, depth / (bbox3d_depth / 2 * scale) -> [0, 1), visibility
 * *= inside the patch and the depth range, coordinates scaled from the (in_h, in_w) patch to the (out_h, out_w, depth_dim)
 * heat-map volume.  joint_img (B,J,3) double: x, y in source-image pixels, z root-relative depth (mm); joint_vis (B,J)
 * double; scale (B) double = the augmentation scale; gt_coord (B,J,3) / gt_vis (B,J) fp32 out, ready for
 * ihpr_integral_l1_fwd*.  fp64 arithmetic as in the reference's numpy (agreement ~1e-12 before the final fp32 cast). */
int ihpr_augment_patches(const unsigned char *images, const int *sizes, int B, int Hs, int Ws,
                         const double *trans, const int *do_flip, const float *color_scale,
                         const float *pixel_mean, const float *pixel_std,
                         int out_h, int out_w, float *out, int channels_last, void *stream);
int ihpr_augment_joints(const double *joint_img, const double *joint_vis, const int *sizes,
                        const double *trans, const double *scale, const int *do_flip, const int *flip_perm,
                        int B, int J, int in_h, int in_w, int out_h, int out_w, int depth_dim, double bbox3d_depth,
                        float *gt_coord, float *gt_vis, void *stream);

/* Test-time post-processing of the (B, J, 3) soft-argmax result in one launch, all buffers on the device:
 *   1. flip-test merge, main/test.py:67-76 -- coords_flipped (nullable = no flip test) is the soft_argmax of the
 *      mirrored image: x' = W - x - 1, joint j takes the flipped pass's joint flip_perm[j] (nullable = identity;
 *      the reference's pairwise swaps as a permutation), result (coords + flipped') / 2   -> merged_out;
 *   2. warp_coord_to_original, common/utils/pose_utils.py:68-75 -- x / W * bbox_w + bbox_x, y / H * bbox_h + bbox_y,
 *      (z / D * 2 - 1) * bbox3d_depth / 2 + center_cam.z                                      -> pixel_out;
 *   3. pixel2cam, pose_utils.py:14-20 -- ((x - c) / f) * z                                    -> cam_out,
 *   4. minus the root joint's camera coordinate when root_idx >= 0, data/Human36M/Human36M.py:226-228.
 * bbox (B,4) = x, y, w, h; center_cam (B,3); focal, princpt (B,2); bbox3d_depth = cfg.bbox_3d_shape[0]
 * (main/config.py:29, 2000 mm).  Each of the three outputs (B, J, 3) is optional (NULL); outputs must not alias
 * inputs.  fp32 arithmetic in the reference's operation order (the reference mixes fp32 and fp64 in numpy:
 * agreement is to fp32 rounding, ~1e-6 relative). */
int ihpr_coords_to_camera(const float *coords, const float *coords_flipped, const int *flip_perm,
                          int B, int J, int D, int H, int W,
                          const float *bbox, const float *center_cam, const float *focal, const float *princpt,
                          float bbox3d_depth, int root_idx,
                          float *merged_out, float *pixel_out, float *cam_out, void *stream);

/* HeadNet.final_layer (1x1 conv with bias, main/model.py:14-20,42) fused with soft_argmax (loss.py:13-34), forward:
 * coords of the heat-map  W x + b  without ever writing the heat-map (tcgen05 / TMEM GEMM with a soft-argmax epilogue).
 * x_nhwc: (B, H, W, K) bf16, i.e. the (B, K, H, W) activations in channels_last memory format; weight: (J*D, K) bf16
 * row-major (= Conv2d weight (J*D, K, 1, 1)); bias: (J*D) fp32.  Needs K % 64 == 0, K <= 256, D in {32, 64, 128},
 * W % 32 == 0, H*W % 256 == 0.  Replaces main/test.py:62-65 (model forward tail + soft_argmax) under no_grad. */
int ihpr_head_softargmax_fwd(const void *x_nhwc, const void *weight, const float *bias,
                             int B, int K, int J, int D, int H, int W,
                             float *coords, float *stats, void *stream);

/* Backward companion of ihpr_head_softargmax_fwd for training with the integral L1 loss: recomputes the heat-map
 * tile by tile on the tensor cores and writes d loss / d heat-map (bf16, (B, J*D, H, W) contiguous) from the coords /
 * stats the forward produced; grad_out is the device scalar d objective / d loss.  dbias_partial (may be NULL) receives
 * (B, 4, J*D) fp32 partial sums of the unrounded gradient: d loss / d bias = their sum over the first two axes.
 * This entry is for callers that want d loss / d heat-map itself; training uses ihpr_head_integral_l1_bwd_params below,
 * which goes all the way to dW / dX / dbias without storing this gradient either. */
int ihpr_head_integral_l1_bwd(const void *x_nhwc, const void *weight, const float *bias,
                              int B, int K, int J, int D, int H, int W,
                              const float *coords, const float *stats,
                              const float *gt, const float *vis, const float *have_depth,
                              const float *grad_out, void *grad_heat, float *dbias_partial, void *stream);

/* The complete backward of final_layer + soft_argmax + JointLocationLoss (main/model.py:14-20,42 under main/train.py:67-71)
 * with NOTHING heat-map-sized stored: d loss / d weight (J*D, K) fp32, d loss / d bias (J*D) fp32 and d loss / d x
 * (B, H, W, K) bf16 -- the channels_last layout of x -- come straight out of two tensor-core kernels (csrc/head_fused_bwd.cu):
 * each recomputes heat-map tiles in TMEM, turns them into bf16 gradient tiles in shared memory and feeds those to a second
 * tcgen05.mma (dW: summed over pixels per channel tile, then over the batch in fixed order; dX: summed over all channel
 * tiles per pixel tile).  No library GEMM, no atomics: results are bit-reproducible.  Any of dx_nhwc / dweight / dbias may
 * be NULL (not needed).  coords / stats are what ihpr_head_softargmax_fwd produced; grad_out is the device scalar
 * d objective / d loss.  workspace: ihpr_head_bwd_workspace_bytes bytes, 256-byte aligned, no initialisation needed.
 * Same shape limits as ihpr_head_softargmax_fwd.  3-4 launches (constants, dW kernel + batch reduction, dX kernel). */
size_t ihpr_head_bwd_workspace_bytes(int B, int K, int J, int D, int H, int W);
int ihpr_head_integral_l1_bwd_params(const void *x_nhwc, const void *weight, const float *bias,
                                     int B, int K, int J, int D, int H, int W,
                                     const float *coords, const float *stats,
                                     const float *gt, const float *vis, const float *have_depth,
                                     const float *grad_out, void *dx_nhwc, float *dweight, float *dbias,
                                     void *workspace, size_t workspace_bytes, void *stream);

/* The last block of HeadNet.deconv_layers at inference (main/model.py:22-38, the third ConvTranspose2d(256, 256, k4, s2, p1,
 * bias=False) + BatchNorm2d in eval mode + ReLU, 32 x 32 -> 64 x 64) as one tensor-core kernel (csrc/deconv_bn_relu.cu): four
 * sub-pixel phases, each a GEMM over 2 x 2 taps x C_in whose shifted A operand is a zero-filled TMA box of the input; BatchNorm
 * folded to a per-channel scale / shift applied in fp32 to the accumulator, ReLU, bf16.
 *   ihpr_deconv_bn_relu_prepare -- once per set of parameters (1 launch): re-lays the ConvTranspose2d weight (C_in, C_out, 4, 4)
 *     bf16 contiguous into 16 K-major (phase, tap) planes and folds gamma / beta / running_mean / running_var (C_out fp32 each)
 *     and eps into scale / shift, all inside `workspace` (ihpr_deconv_bn_relu_workspace_bytes bytes -- 2 MiB for 256 x 256 --
 *     256-byte aligned, no initialisation needed);
 *   ihpr_deconv_bn_relu -- per batch (1 launch): x_nhwc (B, Hin, Win, C_in) bf16 (the channels_last layout of (B, C_in, Hin, Win))
 *     -> y_nhwc (B, 2 Hin, 2 Win, C_out) bf16, exactly the operand ihpr_head_softargmax_fwd reads; `prepared` is a workspace
 *     ihpr_deconv_bn_relu_prepare filled on the same stream (or earlier).
 * Needs C_out == 256, C_in % 64 == 0 and Win == 32 with Hin % 8 == 0 or Win == 16 with Hin % 16 == 0 -- the second and the third
 * block of the reference head (16 x 16 -> 32 x 32 -> 64 x 64) for its 256 x 256 input.  Inference only (running statistics; no backward). */
size_t ihpr_deconv_bn_relu_workspace_bytes(int Cin, int Cout);
int ihpr_deconv_bn_relu_prepare(const void *weight, const float *gamma, const float *beta,
                                const float *running_mean, const float *running_var, float eps,
                                int Cin, int Cout, void *workspace, size_t workspace_bytes, void *stream);
int ihpr_deconv_bn_relu(const void *x_nhwc, const void *prepared, int B, int Cin, int Cout, int Hin, int Win,
                        void *y_nhwc, void *stream);

/* The same block in TRAINING (main/model.py:22-38 under main/train.py:64-71): ConvTranspose2d(256, 256, k4, s2, p1, bias=False) +
 * BatchNorm2d with BATCH statistics + ReLU, forward and backward (csrc/deconv_bn_relu.cu modes kTrain / kDgrad, csrc/bn_train.cu).
 *   ihpr_deconv_bn_relu_train_fwd (4 launches): re-lays `weight` (C_in, C_out, 4, 4) bf16; the tensor-core GEMM stores the raw
 *     convolution output y_raw_nhwc (B, 2 Hin, 2 Win, 256) bf16 and accumulates per-channel sums of y, y^2 in its epilogue; a finalize
 *     launch reduces them in a fixed order (fp64) into saved[0:256] = mean, saved[256:512] = 1 / sqrt(var + eps) (biased variance),
 *     saved[512:768] = gamma * rstd, saved[768:1024] = beta - mean * scale, and updates running_mean / running_var in place as
 *     torch.nn.BatchNorm2d does (momentum, unbiased variance; both may be NULL); one streaming pass writes
 *     out_nhwc = max(0, y_raw * scale + shift) bf16 -- the operand ihpr_head_softargmax_fwd reads.
 *   ihpr_deconv_bn_relu_train_bwd (4-6 launches): dout_nhwc = d loss / d out (bf16 NHWC) -> dgamma, dbeta (256 fp32 each, fixed-order
 *     fp64 reduction), dy_raw_nhwc = d loss / d y_raw (bf16 NHWC; the ReLU mask is recomputed from y_raw) and, when dx_nhwc is not NULL,
 *     dx_nhwc = d loss / d x (B, Hin, Win, 256) bf16 by one tensor-core GEMM (a stride-2 convolution of dy_raw as 16 zero-filled TMA taps).
 *     The weight gradient is ihpr_deconv_wgrad's (x against dy_raw).
 * `workspace`: ihpr_deconv_train_workspace_bytes bytes, 256-byte aligned, no initialisation; may be shared by forward and backward
 * calls on one stream.  Same shape rules as ihpr_deconv_bn_relu plus C_in == 256.  Deterministic for a given device. */
size_t ihpr_deconv_train_workspace_bytes(int Cin, int Cout);
int ihpr_deconv_bn_relu_train_fwd(const void *x_nhwc, const void *weight, const float *gamma, const float *beta,
                                  float *running_mean, float *running_var, float momentum, float eps,
                                  int B, int Cin, int Cout, int Hin, int Win,
                                  void *y_raw_nhwc, void *out_nhwc, float *saved,
                                  void *workspace, size_t workspace_bytes, void *stream);
int ihpr_deconv_bn_relu_train_bwd(const void *dout_nhwc, const void *y_raw_nhwc, const void *weight, const float *saved,
                                  int B, int Cin, int Cout, int Hin, int Win,
                                  void *dy_raw_nhwc, float *dgamma, float *dbeta, void *dx_nhwc,
                                  void *workspace, size_t workspace_bytes, void *stream);

/* The weight gradient of the same block (csrc/deconv_wgrad.cu, K11): dweight (C_in, C_out, 4, 4) fp32 (the layout of
 * ConvTranspose2d.weight) = sum over the batch of x^T . dy per tap, 16 tensor-core GEMMs whose contraction runs over the pixels (both operands
 * are the pixel-major TMA boxes of the NHWC tensors, read as MN-major UMMA operands); the batch is split over the SMs and the fp32 partials are
 * added in a fixed order (deterministic).  x_nhwc (B, Hin, Win, 256) bf16, dy_nhwc (B, 2 Hin, 2 Win, 256) bf16 = dy_raw_nhwc of
 * ihpr_deconv_bn_relu_train_bwd.  2 launches.  `workspace`: ihpr_deconv_wgrad_workspace_bytes bytes (64 MiB), 256-byte aligned. */
size_t ihpr_deconv_wgrad_workspace_bytes(int Cin, int Cout);
int ihpr_deconv_wgrad(const void *x_nhwc, const void *dy_nhwc, int B, int Cin, int Cout, int Hin, int Win,
                      float *dweight, void *workspace, size_t workspace_bytes, void *stream);

/* One reference training step of the path with HOST buffers (what a CPU caller of
 * JointLocationLoss + backward, main/train.py:67-71, holds): copies heat to the device in
 * `slices` batch slices pipelined over internal streams, runs forward + backward and copies
 * loss / coords / grad_heat back.  Synchronous.  grad_heat_host may be NULL (loss only).
 * Pinned host buffers are required for the copies to overlap; pageable ones still work. */
int ihpr_integral_l1_fwd_bwd_host(const void *heat_host, int dtype, int B, int J, int D, int H, int W,
                                  const float *gt_host, const float *vis_host, const float *have_depth_host,
                                  float grad_out, float *loss_host, float *coords_host,
                                  void *grad_heat_host, int device, int slices);

/* Frees the device buffers / streams the *_host entry point caches for `device`. */
int ihpr_host_release(int device);

/* Tuning / introspection (does not change results beyond rounding; the setting belongs to the CALLING THREAD): kernel variant 0 = auto,
 * 1 = TMA-bulk shared-memory ring, 2 = direct 128-bit global loads; for ihpr_integral_l1_fwd_bwd
 * 7 = cluster-resident K5c where it applies, 8 = the one-launch form whenever it applies (no measurement), 9 = always the
 * two-kernel sequence; for ihpr_head_integral_l1_bwd_params 3 = both kernels on SM pairs (tcgen05.mma.cta_group::2; bit-identical results,
 * slower on B200: profiles/r02_head_bench_pairs.txt); for the Python autograd wrapper 6 = the comparison arm (K4 heat-map gradient +
 * library GEMMs). */
int ihpr_set_variant(int variant);
int ihpr_get_variant(void);
/* Number of kernels the LAST call on this thread launched (for bench.py's gpu_launches). */
int ihpr_last_launch_count(void);
/* What the last ihpr_integral_l1_fwd_bwd on this thread ran: 1 = one launch (K5 / K5c), 2 = K1 then K2, 0 = nothing yet;
 * +16 when that call was the one that measured both forms for its (device, shape). */
int ihpr_last_path_choice(void);

#ifdef __cplusplus
}
#endif
#endif /* IHPR_B200_H_ */
