/*
 * ctagan.h -- C ABI of libctagan.so: the sm_100a kernels behind the CTA-GAN hot path.
 *
 * Plain pointers + sizes, no torch types.  Every function returns 0 on success or a non-zero code
 * (ctagan_last_error() gives the message); nothing throws, nothing allocates device memory: the caller
 * owns all buffers (activations, packed weights, workspaces) and passes the cudaStream_t to enqueue on
 * (as void*).  All kernels are capturable in CUDA graphs.
 *
 * Layouts: activations are NHWC ("pixels x channels"), element type `dtype` (CTAGAN_F32 / CTAGAN_BF16);
 * 1-channel images/flows at the module boundary are NCHW == NHWC.  Statistics, losses, weight gradients
 * and master weights are fp32 (accumulators of reductions are fp64).
 *
 * The reference has no native FFI (it is pure PyTorch); each entry point below replaces the ATen/cuDNN
 * call that the cited reference line dispatches (paths relative to the upstream tree).
 */
#ifndef CTAGAN_H_
#define CTAGAN_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTAGAN_F32 0
#define CTAGAN_BF16 1

#define CTAGAN_ACT_NONE 0
#define CTAGAN_ACT_RELU 1
#define CTAGAN_ACT_LRELU 2 /* LeakyReLU(0.2) */
#define CTAGAN_ACT_TANH 3

#define CTAGAN_OK 0
#define CTAGAN_ERR_ARG 1
#define CTAGAN_ERR_CUDA 2
#define CTAGAN_ERR_UNSUPPORTED 3

int ctagan_version(void);
const char *ctagan_last_error(void);
/* 0 if device `dev` is compute capability 10.x, else CTAGAN_ERR_UNSUPPORTED (no other-arch fallback). */
int ctagan_check_device(int dev);

/* A dedicated CUDA stream (cudaStreamNonBlocking; priority: 0 = default, -1 = highest the device offers).  The host side uses
 * these for the branches of an iteration (generator chains, discriminator updates, weight-gradient lanes) instead of streams
 * from PyTorch's pool, which hands the same 32 streams out round-robin: two "different" streams of a long-lived process may be
 * one CUDA stream, which silently serialises branches and entangles their memory reuse. */
int ctagan_stream_create(int priority, void **stream_out);
int ctagan_stream_destroy(void *stream);

/*
 * One "gather convolution" geometry covers Conv2d fwd/dgrad and ConvTranspose2d fwd/dgrad:
 *   y[n,oh,ow,co] = act( bias[co] + sum_{kh,kw,ci} x[n,ih,iw,ci] * wp[co,kh,kw,ci] )
 *   ih*dil = oh*stride + kh - pad_h   (tap skipped unless divisible and 0 <= ih < Hi), same for iw.
 * Conv2d(s,p) fwd: stride=s, dil=1, pad=p.   Conv2d dgrad / ConvTranspose2d fwd: stride=1, dil=s, pad=K-1-p with
 * flipped+transposed packed weights (ctagan_pack_weights mode 1).  Reflection padding is materialised by the
 * producer (ctagan_norm_act_pad), so such convs run with pad=0 on the padded buffer.
 */
typedef struct {
  int32_t N, Hi, Wi, Ci; /* input  NHWC */
  int32_t Ho, Wo, Co;    /* output NHWC */
  int32_t KH, KW;
  int32_t stride, dil;
  int32_t pad_h, pad_w;
  int32_t act;   /* epilogue activation, CTAGAN_ACT_* */
  int32_t dtype; /* element type of x, wp, y */
  int32_t gy_margin; /* wgrad only: gy carries an all-zero border of this many pixels (a hint: it is skipped) */
} ctagan_conv_geom;

/* Replaces nn.Conv2d / nn.ConvTranspose2d forward and the input-gradient half of their backward
 * (Model/CycleGan.py:11,15,28,36,51,59,78-94; trainer/layers.py:83,280,293).  bias may be NULL (fp32[Co] otherwise).
 * engine: 0 = auto, 1 = CUDA-core kernels (fp32-accumulate FFMA; specialised ones for 1-2 channel layers), 2 = force the
 * tcgen05 kernel (error if ineligible), 3 = generic CUDA-core implicit-GEMM kernel only (cross-check). */
int ctagan_conv_gather(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y,
                       int engine, void *stream);
/* The same with caller-provided scratch: the layers with 1-2 OUTPUT channels (7x7 tail of Model/CycleGan.py:58-60, the 2-channel flow
 * head of trainer/reg.py) then run on the tensor cores in two steps (one GEMM per 128 input positions into per-tap planes, then a
 * gather of taps values per output), which needs ctagan_conv_gather_workspace_bytes bytes (0 for every other geometry: the call is
 * then identical to ctagan_conv_gather and workspace may be NULL). */
size_t ctagan_conv_gather_workspace_bytes(const ctagan_conv_geom *g, int engine);
int ctagan_conv_gather_ws(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, void *workspace,
                          size_t workspace_bytes, int engine, void *stream);

/* ctagan_conv_gather with the InstanceNorm statistics fused into the epilogue (tcgen05 engine only; CTAGAN_ERR_UNSUPPORTED
 * otherwise -- query with ctagan_conv_gather_engine).  stats_out[N][Co][2] receives (mean, rstd) of the fp32 convolution output.
 * The reduction is DETERMINISTIC: every CTA stores the column sums of its tile in its own slot of stat_scratch
 * (ctagan_conv_gather_stats_scratch_bytes bytes, any content, 8-byte aligned) and the last CTA of an (image, column tile) to arrive
 * adds the slots in tile order; stat_tickets counts the arrivals: N * ceil(Co/32) uint32, ZERO on entry (zero again on return). */
size_t ctagan_conv_gather_stats_scratch_bytes(const ctagan_conv_geom *g, int engine);
int ctagan_conv_gather_stats(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, uint32_t *stat_tickets,
                             void *stat_scratch, size_t stat_scratch_bytes, float *stats_out, int engine, void *stream);
/* engine ctagan_conv_gather would use: 1 CUDA-core generic, 2 tcgen05, 4 CUDA-core specialised (1-2 channel layers) */
int ctagan_conv_gather_engine(const ctagan_conv_geom *g, int engine);

/* Grouped launches (tcgen05 engine only).  The batch is `groups` consecutive, equally sized image groups; group k is convolved
 * with the weights in slot `slot[k]` of one packed buffer wp[slots][Co][taps][Ci] (bias[slots][Co]).  This is how the two generators
 * (and the two discriminators) of a CycleGAN iteration -- same architecture, different weights, independent inputs
 * (trainer/CycTrainer.py:144-157: netG_A2B(real_A) beside netG_B2A(real_B), then netG_B2A(fake_B) beside netG_A2B(fake_A)) -- run as
 * ONE launch per layer: at batch 1 a single network fills only part of the chip and its kernels are latency-bound, so two
 * problems per launch cost about the same time as one.  stat_tickets / stat_scratch / stats_out as in ctagan_conv_gather_stats
 * (all NULL / 0: no statistics).
 * CTAGAN_ERR_UNSUPPORTED when the geometry does not run on the tcgen05 engine (ask ctagan_conv_gather_grouped_supported first). */
#define CTAGAN_MAX_GROUPS 4
typedef struct {
  int32_t groups;
  int32_t slot[CTAGAN_MAX_GROUPS];
} ctagan_conv_groups;
int ctagan_conv_gather_grouped_supported(const ctagan_conv_geom *g, const ctagan_conv_groups *gr);
int ctagan_conv_gather_grouped(const ctagan_conv_geom *g, const ctagan_conv_groups *gr, const void *x, const void *wp, const float *bias,
                               void *y, uint32_t *stat_tickets, void *stat_scratch, size_t stat_scratch_bytes, float *stats_out, void *stream);
/* one weight (and optional bias) gradient per group: dw[groups][Co][Ci][KH][KW], db[groups][Co]; workspace as for ctagan_conv_wgrad
 * (0 bytes == unsupported geometry) */
size_t ctagan_conv_wgrad_grouped_workspace_bytes(const ctagan_conv_geom *g, int groups);
int ctagan_conv_wgrad_grouped(const ctagan_conv_geom *g, int groups, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                              size_t workspace_bytes, void *stream);

/* Weight gradient (and optional bias gradient) of the same geometry:
 *   dw[a,b,kh,kw] (+)= sum_{n,oh,ow} gy[n,oh,ow,a] * gx[n,ih,iw,b],  db[a] = sum gy[n,oh,ow,a]
 * with (ih,iw) from (oh,ow,kh,kw) as above; g->Co == A (channels of gy), g->Ci == B (channels of gx); dw is fp32 in
 * PyTorch's [A][B][KH][KW] order and is OVERWRITTEN (db too, may be NULL) -- or, with accumulate != 0, ADDED TO (the second use of a
 * network inside one backward pass, e.g. the cycle pass of CycTrainer.py:153-157, lands in the same gradient buffer as the first:
 * what autograd's AccumulateGrad does, without a temporary and an add kernel).  Conv2d: gy=dy, gx=x.  ConvTranspose2d: gy=x, gx=dy.
 * `accumulate` is a bit set: CTAGAN_WGRAD_ACCUMULATE (1) as above; CTAGAN_WGRAD_PACKED (2): dw is stored as [A][KH][KW][B] (the
 * channels-last order of the same logical [A][B][KH][KW] tensor, i.e. what `p.grad` is when it is a torch.channels_last view): the
 * tensor-core kernel then writes whole 16-byte rows instead of a 4-byte scatter with stride KH*KW.  Not available for the 1-2 channel
 * layers (CTAGAN_ERR_UNSUPPORTED).
 * Replaces the weight-gradient half of cudnn/ATen convolution_backward for the layers cited above. */
#define CTAGAN_WGRAD_ACCUMULATE 1
#define CTAGAN_WGRAD_PACKED 2
int ctagan_conv_wgrad(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db,
                      void *workspace, size_t workspace_bytes, int engine, int accumulate, void *stream);
/* Scratch bytes ctagan_conv_wgrad needs for this geometry/engine: the per-CTA / per-split partial sums of every engine's split
 * reduction (each CTA stores its partial result in its own row, a second kernel adds the rows in order: no floating-point atomics,
 * the same inputs give the same bits). */
size_t ctagan_conv_wgrad_workspace_bytes(const ctagan_conv_geom *g, int engine);

/* fp32 master weights W[O][I][KH][KW] -> packed `dtype` weights.
 * mode 0: wp[O][kh][kw][I] = W[O][I][kh][kw];  mode 1: wp[I][kh][kw][O] = W[O][I][KH-1-kh][KW-1-kw]. */
int ctagan_pack_weights(const float *w, void *wp, int O, int I, int KH, int KW, int mode, int dtype, void *stream);

/* The same for many weights in ONE launch (all layers of a network x both layouts, after every optimiser step).
 * `items` is a host array; it is copied into the kernel parameters. */
typedef struct {
  const float *w;
  void *wp;
  int32_t O, I, KH, KW, mode;
} ctagan_pack_item;
int ctagan_pack_weights_multi(const ctagan_pack_item *items, int n_items, int dtype, void *stream);

/* f3: the optimiser step of a whole parameter group in ONE launch -- torch.optim.Adam(betas=(0.5, 0.999)) of trainer/CycTrainer.py:67-73
 * (RegTrainer.py:97-101, HdTrainer.py:101-105, p2pTrainer.py:62-63) with the arithmetic of PyTorch's fused Adam kernel, operation for
 * operation -- that also emits both packed copies of every convolution weight (ctagan_pack_weights modes 0 and 1, `packed_dtype`) from
 * the tile it has just updated.  An item is one parameter tensor viewed as [O][I][KH][KW] (1-D tensors: O = numel, I = KH = KW = 1);
 * wp0 / wp1 may be NULL.  items_dev / tile_start_dev: the table in DEVICE memory (ctagan_adam_pack_tiles fills the host copy of
 * tile_start[n_items + 1]; total_tiles = tile_start[n_items]); lr_dev: the learning rate; step_dev: the step counter as a float
 * (as torch's state['step']; incremented by the kernel when advance_step != 0 -- a step may be split into several launches over disjoint
 * parts of the group, e.g. an early one for the layers whose gradients are final before the backward pass ends: only the last one
 * advances the counter, and it must be ordered after the others); ticket_dev: one zeroed uint32 per launch in flight.  Capturable in
 * CUDA graphs. */
typedef struct {
  float *p;
  const float *g;
  float *m;
  float *v;
  void *wp0;
  void *wp1;
  int32_t O, I, KH, KW;
  int32_t g_packed; /* 0: g is [O][I][KH][KW] like p; 1: g is [O][KH][KW][I] (CTAGAN_WGRAD_PACKED) */
  int32_t reserved;
} ctagan_adam_item;
size_t ctagan_adam_pack_smem_bytes(const ctagan_adam_item *items_host, int n_items);
int ctagan_adam_pack_tiles(const ctagan_adam_item *items_host, int n_items, int *tile_start_host);
int ctagan_adam_pack_multi(const ctagan_adam_item *items_dev, const int *tile_start_dev, int n_items, int total_tiles, size_t smem_bytes,
                           const float *lr_dev, float *step_dev, uint32_t *ticket_dev, float beta1, float beta2, float eps, int packed_dtype,
                           int advance_step, void *stream);

/* InstanceNorm2d statistics (affine=False, eps=1e-5, biased variance; Model/CycleGan.py:12,16,29,37,52,82,86,90,
 * trainer/layers.py:14): x[N][HW][C] -> stats[N][C][2] = (mean, rstd) fp32.  acc: caller scratch of
 * ctagan_instnorm_stats_scratch_doubles doubles (per-block partial sums, added in block order). */
size_t ctagan_instnorm_stats_scratch_doubles(int N, int HW, int C, int dtype);
int ctagan_instnorm_stats(const void *x, float *stats, double *acc, int N, int HW, int C, int dtype, void *stream);

/* Fused InstanceNorm-apply + activation + residual add + reflection pad (one pass):
 *   out[n,hp,wp,c] = act((x[n,h,w,c]-mean)*rstd) + res[n,h+res_pad,w+res_pad,c],  (h,w) = reflect(hp-pad, wp-pad)
 * stats==NULL: no normalisation; res==NULL: no residual.  res has spatial size (H+2*res_pad, W+2*res_pad).
 * Replaces InstanceNorm2d+ReLU/LeakyReLU+ReflectionPad2d+residual add (Model/CycleGan.py:10-21,27-30). */
int ctagan_norm_act_pad(const void *x, const float *stats, const void *res, int res_pad, void *out, int N, int H, int W, int C, int pad,
                        int act, int dtype, void *stream);

/* Backward of the above.  gout is the gradient w.r.t. `out` (padded, size H+2*pad); x is the saved raw conv output
 * (or, when stats==NULL, any tensor with the sign of the pre-activation, e.g. the post-activation output).
 *   g  = (fold_reflect(gout) + addend) * act'(.)          (addend optional, unpadded; act' from x, stats)
 *   dx = rstd*(g - mean(g) - xhat*mean(g*xhat))   (stats!=NULL)   or   g   (stats==NULL)
 * The residual branch of the forward receives fold_reflect(gout) itself (call with stats=NULL, act=NONE).
 * dx is written as [N][H+2*out_pad][W+2*out_pad][C] with a zero margin of out_pad pixels: with out_pad = K-1-p the stride-1
 * input-gradient convolution that consumes it becomes a plain VALID convolution (the tcgen05 engine's native form).
 * g_out (optional, [N][H][W][C]): also receives fold_reflect(gout) + addend, i.e. the gradient that continues along the skip
 * connection of a residual block -- one launch then serves both consumers of a block's output gradient.
 * bf16 maps of up to 64x64 pixels per image run as ONE kernel (thread-block clusters of 8 CTAs per 16 channels, per-channel sums
 * combined through distributed shared memory, operands read once and kept in registers; acc / scratch are not used); otherwise two
 * kernels: reduce (every block stores its partial sums in its own row of `scratch`,
 * ctagan_norm_act_pad_bwd_scratch_doubles doubles; the last block of an image adds the rows in order into acc), then apply.
 * acc: N*C*2 + N doubles (only with stats); the last N are the arrival tickets and must be ZERO on entry (they are zero again on
 * return): acc_is_zero != 0 promises that, otherwise they are cleared here.  ctagan_norm_act_pad_bwd_launches tells which path runs. */
size_t ctagan_norm_act_pad_bwd_scratch_doubles(int has_stats, int N, int H, int W, int C, int dtype);
int ctagan_norm_act_pad_bwd(const void *gout, const void *x, const float *stats, const void *addend, void *dx, void *g_out,
                            double *acc, int acc_is_zero, double *scratch, int N, int H, int W, int C, int pad, int act, int out_pad,
                            int dtype, void *stream);
int ctagan_norm_act_pad_bwd_launches(int has_stats, int H, int W, int C, int dtype);

/* Pointwise activation backward for conv-epilogue activations: dx = gy * act'(y) computed from the OUTPUT y
 * (relu/lrelu: sign(y); tanh: 1-y^2).  n elements. */
int ctagan_act_bwd(const void *gy, const void *y, void *dx, int64_t n, int act, int dtype, void *stream);

/* MaxPool2d(2) (trainer/layers.py:172) on NHWC; bwd routes to the first maximum in window scan order (ATen rule). */
int ctagan_maxpool2_fwd(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream);
/* gx = scatter(gy) + addend (addend optional: the skip-connection gradient of trainer/reg.py:94) */
int ctagan_maxpool2_bwd(const void *gy, const void *x, const void *addend, void *gx, int N, int H, int W, int C, int dtype, void *stream);

/* F.interpolate(x, 2x, mode='bilinear', align_corners=False) + torch.cat([up, skip], 1) (trainer/reg.py:93-94):
 * x[N][H][W][C1], skip[N][2H][2W][C2] -> out[N][2H][2W][C1+C2].  bwd: gout -> gx (gskip is a strided read of gout). */
int ctagan_upsample2x_cat_fwd(const void *x, const void *skip, void *out, int N, int H, int W, int C1, int C2, int dtype, void *stream);
int ctagan_upsample2x_cat_bwd(const void *gout, void *gx, void *gskip, int N, int H, int W, int C1, int C2, int dtype, void *stream);

/* Channel slice copy / concat helper: dst[n,p,dst_off:dst_off+C] = src[n,p,0:C]  (torch.cat, trainer/reg.py:77, p2pTrainer.py:131). */
int ctagan_copy_channels(const void *src, void *dst, int64_t pixels, int C, int src_stride, int src_off, int dst_stride, int dst_off,
                         int dtype, void *stream);

/* Transformer_2D.forward (trainer/transformer.py:11-31): bilinear grid_sample(align_corners=True, padding_mode="border") of
 * src[B][C][H][W] at (i + flow[b,0,i,j], j + flow[b,1,i,j]); fp32 in/out (module boundary).  bwd produces gsrc AND gflow (either
 * may be NULL).  A block stages the source window of its 16x64 output tile (+-8 pixels) in shared memory; rows move as 16-byte
 * vectors (flow / out / gout / gflow 16-byte aligned).  gsrc is a scatter-add: it is accumulated in 64-bit fixed point (scaled by
 * max|gout|), so it does not depend on the order of the additions; workspace: ctagan_warp_bwd_workspace_bytes (only with gsrc). */
int ctagan_warp_fwd(const float *src, const float *flow, float *out, int B, int C, int H, int W, void *stream);
size_t ctagan_warp_bwd_workspace_bytes(int B, int C, int H, int W);
int ctagan_warp_bwd(const float *gout, const float *src, const float *flow, float *gsrc, float *gflow, void *workspace,
                    size_t workspace_bytes, int B, int C, int H, int W, void *stream);

/* Fused single-pass losses; `loss` is one fp32 on the device, acc a scratch of CTAGAN_LOSS_ACC_DOUBLES doubles (ticket + one partial
 * sum per block, added in block order by the last block to arrive: deterministic, one launch, no host sync).  fp32 tensors.
 * l1: torch.nn.L1Loss (CycTrainer.py:77);  mse_const: torch.nn.MSELoss vs a broadcast constant (CycTrainer.py:76,83-84);
 * smooth: smooothing_loss (trainer/utils.py:165-173);  masked_l1: HdTrainer.py:726-735. */
#define CTAGAN_LOSS_ACC_DOUBLES 1024
int ctagan_l1_fwd(const float *a, const float *b, float *loss, double *acc, int64_t n, void *stream);
int ctagan_l1_bwd(const float *a, const float *b, const float *gloss, float *ga, int64_t n, void *stream);
/* mse_const: the constant is `target`, or -- when target_dev != NULL -- the fp32 value target_dev points to on the device (the
 * reference keeps its targets as (1,1) device tensors, CycTrainer.py:83-84: no host read, graph-capturable). */
int ctagan_mse_const_fwd(const float *p, float target, const float *target_dev, float *loss, double *acc, int64_t n, void *stream);
int ctagan_mse_const_bwd(const float *p, float target, const float *target_dev, const float *gloss, float *gp, int64_t n, void *stream);
int ctagan_smooth_fwd(const float *flow, float *loss, double *acc, int B, int C, int H, int W, void *stream);
int ctagan_smooth_bwd(const float *flow, const float *gloss, float *gflow, int B, int C, int H, int W, void *stream);
int ctagan_masked_l1_fwd(const float *warped, const float *b1, const float *b2, float *loss, double *acc, int64_t n, void *stream);
int ctagan_masked_l1_bwd(const float *warped, const float *b1, const float *b2, const float *gloss, float *gwarped, int64_t n, void *stream);

/* Global average pool of a (N, HW, C) map -> (N, C) fp32 (Model/CycleGan.py:103; Model/HdGan.py:279,288) and its backward. */
int ctagan_plane_mean_fwd(const void *x, float *out, int N, int HW, int C, int dtype, void *stream);
int ctagan_plane_mean_bwd(const float *gout, void *gx, int N, int HW, int C, int dtype, void *stream);

/* Module boundary: fp32 NCHW <-> internal NHWC of `dtype` (cast fused). */
int ctagan_nchw_to_nhwc(const float *src, void *dst, int N, int C, int64_t HW, int dtype, void *stream);
int ctagan_nhwc_to_nchw(const void *src, float *dst, int N, int C, int64_t HW, int dtype, void *stream);

/* torch.cat([a, b], 1) of two 1-channel fp32 images into one 2-channel NHWC `dtype` tensor (trainer/reg.py:77) and the
 * split of its gradient (a or b may be NULL).  n = pixels (N*H*W). */
int ctagan_interleave2(const float *a, const float *b, void *dst, int64_t n, int dtype, void *stream);
int ctagan_deinterleave2(const void *src, float *a, float *b, int64_t n, int dtype, void *stream);

/* ---- f1: input pipeline on the GPU (batched; the reference does this per slice on the CPU inside its DataLoader worker) ----
 * hu_to_unit: read_dicom (trainer/datasets.py:74-82) and the raw half of read_ori_w (:62-65): v = raw + add, negative -> 0,
 *   (v / 4095 - 0.5) / 0.5 (double arithmetic, stored fp32).  add = 0 for stored pixel values, 1024 for HU (SimpleITK reads HU).
 * hu_window: the display-window half of read_ori_w (:45-56; centre 50 / width 400 there): trunc((hu - win_min) * 255 / (win_max -
 *   win_min)) clipped to [0, 255], / 255, (x - 0.5) / 0.5.
 * resize_nearest: Resize (trainer/utils.py:13-30) = F.interpolate(size) with the default 'nearest' mode, [B][Hs][Ws] -> [B][Hd][Wd].
 * affine_nearest: the resampling of RandomAffine(fillcolor=-1) (trainer/CycTrainer.py:91-95): PIL's nearest-neighbour affine transform in
 *   16.16 fixed point; inv_matrix[B][6] (device, fp64) holds the inverse affine matrix (a, b, c, d, e, f) of every slice, the random
 *   parameters themselves are drawn on the host exactly as torchvision draws them. */
int ctagan_hu_to_unit(const int16_t *raw, float *out, int64_t n, int add, void *stream);
int ctagan_hu_window(const int16_t *hu, float *out, int64_t n, double center, double width, void *stream);
int ctagan_resize_nearest(const float *src, float *dst, int B, int Hs, int Ws, int Hd, int Wd, void *stream);
int ctagan_affine_nearest(const float *src, float *dst, int B, int H, int W, const double *inv_matrix, float fill, void *stream);

/* ---- f2: evaluation metrics of test() on the GPU (trainer/CycTrainer.py:286-330, :362-398), batched, no host round trip per slice ----
 * fake / real: [B][H][W] fp32 in [-1, 1].  Per slice: the display window (to_windowdata, :34-57, centre wc / width ww), the
 * 0.3-threshold masks, and for both image pairs the reference compares (the thresholded windowed pair and the masked raw pair)
 * MAE, PSNR, SSIM (skimage compare_ssim defaults: 7x7 uniform window, sample covariance, data range 2) and UQI:
 *   out[B][8] = (MAEw, PSNRw, SSIMw, UQIw, MAE, PSNR, SSIM, UQI), fp64.  Deterministic (ordered partial sums).
 * scratch: ctagan_eval_metrics_scratch_doubles doubles.  LPIPS (a third-party network) is not computed.
 * to_dicom_i16: (x + 1) * 0.5 * 4095 truncated to int16, the pixel data written back to DICOM (:337-341). */
size_t ctagan_eval_metrics_scratch_doubles(int B, int H, int W);
int ctagan_eval_metrics(const float *fake, const float *real, double *out, double *scratch, int B, int H, int W, double wc, double ww,
                        void *stream);
int ctagan_to_dicom_i16(const float *x, int16_t *out, int64_t n, void *stream);

/* dtype conversion (fp32 <-> activation dtype), n elements. */
int ctagan_cast(const void *src, int src_dtype, void *dst, int dst_dtype, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTAGAN_H_ */
