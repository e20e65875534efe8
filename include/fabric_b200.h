/* fabric_b200 -- C ABI of the B200-native BiDateNet hot path.
 *
 * The reference (granularai/fabric) has no FFI: its hot path is the Python nn.Module surface of
 * models/bidate_model.py / models/unet_parts.py and the loss callables of utils/metrics.py
 * (SURVEY.md section 8b).  This header is the C boundary that surface binds to in this repo: every entry
 * point names the reference construct it replaces.  Plain pointers and sizes only; all pointers are DEVICE
 * pointers unless the name ends in _host; `stream` is a cudaStream_t passed as void*.  Nothing here
 * allocates device memory: scratch is passed in.  Every function returns 0 on success or a negative
 * fb_status; fabric_b200_last_error() returns a thread-local message.  sm_100a only: on any other device every
 * compute entry point returns FB_ERR_ARCH (there is no fallback path).
 *
 * Activation layout ("NHWC5"): bf16, channel-innermost, dims [G][B][H][W][C]; G is the date group
 * (2 inside the weight-shared encoder, which runs date 1 and date 2 as one launch; 1 in the decoder).
 */
#ifndef FABRIC_B200_H_
#define FABRIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FB_OK = 0,
  FB_ERR_SHAPE = -1,   /* unsupported / inconsistent shape */
  FB_ERR_ALIGN = -2,   /* pointer not 16-byte aligned */
  FB_ERR_ARCH = -3,    /* device is not sm_100 */
  FB_ERR_LAUNCH = -4,  /* CUDA launch / driver error (message has cudaGetErrorString) */
  FB_ERR_ARG = -5      /* null pointer or bad enum */
} fb_status;

int fabric_b200_version(void);
const char* fabric_b200_last_error(void);
/* number of SMs of the current device (persistent grids are sized from it); <0 on error */
int fabric_b200_sm_count(void);

/* ---- layout packing ------------------------------------------------------------------------------------- */

/* NCHW fp32 [B][C][H][W] -> NHWC bf16 [B][H][W][Cpad], channels C..Cpad-1 zero.  Entry of
 * BiDateNet.forward (models/bidate_model.py:22-23,29; inputs come from train.py:83-84). */
int fabric_b200_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C, int Cpad, int H, int W,
                                           void* stream);
/* Raw input: NCHW uint16 [B][C][H][W] (Sentinel-2 digital numbers, C <= 16) -> NHWC bf16 [B][H][W][16] with the per-band
 * z-score (v - mean[c]) * inv_std[c] of city_loader (utils/dataloaders.py:94-99) applied on the fly, so the host hands
 * over the rasters as stored (half the bytes of the reference's fp32 patches at train.py:189-190). */
int fabric_b200_pack_nchw_u16_to_nhwc_bf16(const uint16_t* src, void* dst, const float* mean, const float* inv_std, int B,
                                           int C, int H, int W, void* stream);
/* Training patches with the loader's augmentation fused (onera_siamese_loader, utils/dataloaders.py:152-163): src NCHW
 * [B][C][S][S] (src_dtype 0 = fp32, 1 = uint16 with mean / inv_std as above, both nullable for fp32), aug int32 [B][3] =
 * (rot90 quarter turns, flip rows, flip columns) per sample -> NHWC bf16 [B][S][S][16] of the augmented patch.  The same
 * `aug` rows are applied to the labels int64 [B][S][S] by fabric_b200_augment_labels (out of place). */
int fabric_b200_pack_nchw_aug(const void* src, int src_dtype, void* dst, const int* aug, const float* mean,
                              const float* inv_std, int B, int C, int S, void* stream);
int fabric_b200_augment_labels(const int64_t* src, int64_t* dst, const int* aug, int B, int S, void* stream);
/* NHWC bf16 [B][H][W][C] -> NCHW fp32 [B][C][H][W] (standalone block outputs / tests). */
int fabric_b200_unpack_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W, void* stream);

/* nn.Conv2d weight [Cout][Cin][3][3] fp32 -> bf16 GEMM operand.
 *   mode 0 (forward): dst[Cout][9][CinPad], k = tap*CinPad + ci, tap = ky*3+kx
 *   mode 1 (dgrad)  : dst[Cin][9][Cout],    taps flipped, so that dX = conv3x3(dY, dst)            */
int fabric_b200_pack_conv3x3_weight(const float* w, void* dst, int Cout, int Cin, int CinPad, int mode, void* stream);
/* mode 0 with w[co] * scale[co] (scale fp32 [Cout], nullable): eval-mode BatchNorm scale folded into the weights, used
 * together with fb_conv3x3_desc.shift_in_acc */
int fabric_b200_pack_conv3x3_weight_scaled(const float* w, const float* scale, void* dst, int Cout, int Cin, int CinPad,
                                           int mode, void* stream);

/* ---- 3x3 convolution (tcgen05 implicit GEMM) ---------------------------------------------------------------- */

typedef struct {
  int n_tile;      /* 0 = auto, else 64 / 128 / 256 */
  int halo;        /* -1 = auto, 0 = per-tap TMA loads, 1 = halo tile shared by the nine taps */
  int a_stages;    /* 0 = auto */
  int b_stages;    /* 0 = auto */
  int b_resident;  /* -1 = auto, 0 / 1 */
  int grid;        /* 0 = auto (#SMs rounded to a multiple of the N tiles) */
  int ctas;        /* 0 = auto, 1 = one CTA per 128-pixel tile, 2 = CTA pair (tcgen05 cta_group::2, M = 256) */
  int epi_warps;   /* 0 = auto, 4 or 8 epilogue warps (8 only for n_tile <= 128) */
  int occupancy;   /* 0 / 1 = one CTA per SM; 2 = two (n_tile 64, 4 epilogue warps, halo, resident weights, <= 112 KB smem) */
} fb_conv_tuning;

typedef struct {
  int G, B, H, W;      /* NHWC5 dims of input and output */
  int Cin;             /* padded input channels: 16 or a multiple of 64 */
  int Cout;            /* multiple of 64 */
  int relu;            /* apply max(.,0) after scale/shift */
  int store_main;      /* write y (0 only together with head_out) */
  const void* x;       /* bf16 [G][B][H][W][Cin] */
  const void* w;       /* bf16 [Cout][9][Cin] (fabric_b200_pack_conv3x3_weight) */
  void* y;             /* bf16 [G][B][H][W][Cout] */
  const float* scale;  /* [Cout] or NULL (=1): y = acc*scale + shift */
  const float* shift;  /* [Cout] or NULL (=0) */
  void* pool_out;      /* NULL or bf16 [G][B][H/2][W/2][Cout]: fused nn.MaxPool2d(2) of y (unet_parts.py:40) */
  float* stats_ws;     /* NULL or fp32 workspace of fabric_b200_conv3x3_stats_ws_floats(): per-CTA partial
                          (sum, sum of squares) of y per channel and date group, for BatchNorm batch moments */
  const float* head_w; /* NULL or fp32 [2][64]: fused outconv 1x1 (unet_parts.py:86), needs Cout == 64 */
  const float* head_b; /* fp32 [2] */
  float* head_out;     /* fp32 NCHW [G*B][2][H][W] */
  void* prod_out;      /* NULL or bf16 [B][H][W][prod_channels]: fused relu(y[date 1] * y[date 0]) written into
                          channels [0, Cout) -- the skip half of the decoder input (bidate_model.py:35-38); G == 2 */
  int prod_channels;
  int shift_in_acc;    /* 1: the accumulator starts at shift[] (scale must be NULL, i.e. folded into w): y = acc + shift.
                          The epilogue is then convert (+ReLU) + store; the epilogue warps re-prime TMEM after each tile */
  fb_conv_tuning tune;
  /* Fused BatchNorm-backward reduce, for the data-gradient launch whose output is dL/da of a training-mode
   * BatchNorm2d + ReLU (unet_parts.py:14-15 as seen by autograd): bnbwd_z = that layer's pre-activation z, bf16
   * [G][B][H][W][Cout]; bnbwd_coef = fp32 [4][G][Cout] (scale, shift, mean, invstd from fabric_b200_bn_finalize).  The
   * epilogue stores dy = relu'(z*scale+shift) * acc instead of acc and writes per-CTA (sum dy, sum dy*z) partials into
   * stats_ws (same layout and size as the moment partials), which fabric_b200_bn_bwd_from_partials finishes -- the
   * stand-alone reduce pass over (dL/da, z) disappears.  Raw-accumulator epilogue only (no scale / shift / relu / pool /
   * head / product). */
  const void* bnbwd_z;
  const float* bnbwd_coef;
} fb_conv3x3_desc;

/* y = conv3x3(x, w) * scale + shift (+ReLU) (+pool) (+BN moment partials) (+1x1 head).
 * Replaces nn.Conv2d(in,out,3,padding=1) and the BatchNorm2d(eval)/ReLU that follow it
 * (models/unet_parts.py:13-18); with mode-1 weights it is also the data gradient of that conv. */
int fabric_b200_conv3x3(const fb_conv3x3_desc* d, void* stream);
/* The launch plan fabric_b200_conv3x3 would use on a device with `sms` SMs and `smem_optin` bytes of opt-in shared memory
 * per block (B200: 148, 232448).  Pure host arithmetic, no CUDA call: pointers in `d` only need to be NULL / non-NULL. */
typedef struct {
  int n_tile, ck, halo, grid, smem_bytes, ctas, epi_warps, a_stages, b_stages, b_resident, out_bufs, total_units;
  int pool_tma, prod_tma, ctas_per_sm;
  int reg_stats;   /* 1: per-channel sums accumulate in registers across the CTA's tiles (training instantiation) */
} fb_conv3x3_plan;
int fabric_b200_conv3x3_plan(const fb_conv3x3_desc* d, int sms, int smem_optin, fb_conv3x3_plan* out);
/* grid the launch above will use, and the stats workspace size (floats) for it */
int fabric_b200_conv3x3_grid(const fb_conv3x3_desc* d);
int64_t fabric_b200_conv3x3_stats_ws_floats(const fb_conv3x3_desc* d);

/* ---- BatchNorm helpers --------------------------------------------------------------------------------------- */

/* Eval-mode BatchNorm2d folded with the conv bias into the conv epilogue (unet_parts.py:13-14):
 *   scale = gamma / sqrt(running_var + eps), shift = (conv_bias - running_mean) * scale + beta */
int fabric_b200_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                             const float* running_var, const float* conv_bias, float eps, float* scale, float* shift,
                             int C, void* stream);

/* ---- decoder input ---------------------------------------------------------------------------------------------- */

/* One kernel for   torch.cat([relu(s_d2*s_d1), pad(upsample_bilinear_x2(low))], 1)
 * (models/bidate_model.py:35-38 + models/unet_parts.py:56-58,65-78):
 *   out[b][y][x][0:Cs]      = skip[0][b][y][x][:] * skip[1][b][y][x][:]
 *   out[b][y][x][Cs:Cs+Cl]  = bilinear(align_corners=True) of low, zero outside the padded window
 * low has low_groups = 2 (low = product of the two dates, up1) or 1 (previous decoder stage).
 * skip == NULL: only the upsampled channels are written (the skip half came from fabric_b200_conv3x3's prod_out). */
int fabric_b200_build_up_input(const void* skip, const void* low, void* out, int B, int H, int W, int Cs, int h, int w,
                               int Cl, int low_groups, void* stream);

/* outconv (unet_parts.py:83-90): NHWC bf16 [B][H][W][64] -> NCHW fp32 logits [B][2][H][W] */
int fabric_b200_outconv(const void* x, const float* w, const float* b, float* logits, int B, int H, int W, int C,
                        void* stream);

/* ---- training: BatchNorm batch statistics ----------------------------------------------------------------- */

/* nn.BatchNorm2d in training mode (unet_parts.py:14,17), finishing the moment partials the conv epilogue wrote
 * (stats_ws of fabric_b200_conv3x3, launched with `grid` CTAs and N tile `n_tile`).  Per date group g:
 *   mean_g, var_g (biased) -> scale[g][c] = gamma*invstd, shift[g][c] = beta - mean*scale  (conv bias cancels),
 *   running_mean/var updated with momentum (unbiased variance, conv bias added to the mean), group 0 first then
 *   group 1 (the reference runs date 1, then date 2); num_batches_tracked += G. */
int fabric_b200_bn_finalize(const float* stats_ws, int grid, int n_tile, int C, int G, int64_t count_per_group,
                            const float* conv_bias, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* scale,
                            float* shift, float* mean, float* invstd, void* stream);
/* a = relu(z*scale[g] + shift[g]) bf16 NHWC5, optional fused MaxPool2d(2) copy, optional fused
 * relu(a[date 1] * a[date 0]) written into channels [0, C) of prod_out [B][H][W][prod_channels] (G == 2) */
int fabric_b200_bn_apply_relu(const void* z, const float* scale, const float* shift, void* a, void* pool_out, void* prod_out,
                              int prod_channels, int G, int B, int H, int W, int C, void* stream);

/* The last decoder BatchNorm + ReLU (up4, 64 channels) and `outconv` (unet_parts.py:83-90, bidate_model.py:38-39) in ONE
 * pass: a = relu(z*scale+shift) stored as bf16 [B][H][W][64] (a may be NULL: not stored -- the backward entry point below
 * recomputes it from z) and logits = W a + b as fp32 NCHW [B][2][H][W]. */
int fabric_b200_bn_apply_relu_head(const void* z, const float* scale, const float* shift, void* a, const float* head_w,
                                   const float* head_b, float* logits, int B, int H, int W, int C, void* stream);
/* ... and their backward, also fused: du = W^T dlogits is never materialised (dy = relu'(.) * du is recomputed from
 * dlogits in both passes), the activation needed for the head's weight gradient is recomputed from z.  Phase 1: partials
 * of (sum dy, sum dy*xhat), dW, db into the first sms*2*258 floats of ws; phase 2: dgamma, dbeta, dw [2][64], db [2] and
 * dz = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)); phase 3 = both.  Replaces fabric_b200_outconv_bwd +
 * fabric_b200_bn_relu_bwd for that layer (three passes over 64-channel full-resolution tensors become two over z).
 * count_scale / grad_scale: exact-global mode, as in fabric_b200_bn_relu_bwd_phase (the caller all-reduces the phase-1
 * partials; grad_scale = 1/world then also applies to dw / db). */
int64_t fabric_b200_bn_head_bwd_ws_floats(void);
int fabric_b200_bn_head_bwd(int phase, const float* dlogits, const void* z, const float* scale, const float* shift,
                            const float* mean, const float* invstd, const float* gamma, const float* head_w, void* dz,
                            float* dgamma, float* dbeta, float* dw, float* db, float* ws, int B, int H, int W, int C,
                            float count_scale, float grad_scale, void* stream);

/* ---- training: losses (utils/metrics.py) --------------------------------------------------------------------- */

/* kind: 0 tversky(alpha,beta) :130-171, 1 dice :51-83, 2 jaccard :86-119, 3 focal(gamma) :19-48,
 * 4 softmax cross entropy (working stand-in for the reference's broken `bce`, utils/helpers.py:303-304).
 * logits NCHW fp32 [B][2][H][W], labels int64 [B][H][W]; label_ndim 3 or 4 selects the reference's `dims`
 * behaviour (3-D labels reduce over batch and H only).  Writes the scalar loss and dL/dlogits. */
int64_t fabric_b200_seg_loss_ws_floats(int B, int H, int W);
int fabric_b200_seg_loss_fwd_bwd(int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                                 const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                                 float* ws, void* stream);

/* The same in two phases, for the exact-global loss of a data-parallel job (the reference computes the loss on the batch
 * nn.DataParallel gathered, train.py:91-92; tversky / dice / jaccard are ratios of batch sums, so per-rank losses do not
 * add up): phase 1 writes the per-column sums [6][W] (I_c, P_c, T_c for c = 0, 1) at ws + fabric_b200_seg_loss_sums_offset();
 * the caller all-reduces those 6*W floats; phase 2 turns them into the loss and dL/dlogits.  Focal / CE are plain means:
 * phase 1 does nothing and `mean_scale` (1/world) scales their loss and gradient in phase 2. */
int64_t fabric_b200_seg_loss_sums_offset(int B, int H, int W);
int fabric_b200_seg_loss_phase(int phase, int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                               const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                               float* ws, float mean_scale, void* stream);

/* ---- training: backward ----------------------------------------------------------------------------------------- */

/* outconv backward: du = dlogits^T W (bf16 NHWC), dw [2][C], db [2] */
int64_t fabric_b200_outconv_bwd_ws_floats(int C);
int fabric_b200_outconv_bwd(const float* dlogits, const void* u, const float* w, void* du, float* dw, float* db, float* ws,
                            int B, int H, int W, int C, void* stream);

/* BatchNorm(train) + ReLU backward, with the adjoints of the ops that consume the activation fused into the read:
 *   dy = relu'(z*scale+shift) * [ ga * (mul_other ? a[other date] : 1)  +  maxpool-unpool(gp) ]
 *   dz = gamma*invstd * (dy - mean(dy) - xhat*mean(dy*xhat)),   dgamma = sum dy*xhat,  dbeta = sum dy
 * ga: bf16 [ga_groups][B][H][W][ga_channels] (first C channels used; ga_groups 1 = shared by both dates),
 * mul_other: product fusion relu(d2*d1) adjoint (bidate_model.py:35-38); gp: bf16 [G][B][H/2][W/2][C] gradient of
 * the pooled copy, routed to the first maximum of each 2x2 window like nn.MaxPool2d. */
int64_t fabric_b200_bn_bwd_ws_floats(int G, int C);
int fabric_b200_bn_relu_bwd(const void* z, const void* a, const void* ga, int ga_groups, int ga_channels, int mul_other,
                            const void* gp, const float* scale, const float* shift, const float* mean, const float* invstd,
                            const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws, int G, int B, int H, int W,
                            int C, void* stream);

/* The same in two phases, for exact-global BatchNorm (SyncBN == the reference's single-device statistics over the whole
 * batch): phase 1 writes the (sum dy, sum dy*xhat) partials into the first fabric_b200_bn_bwd_partial_floats() floats of ws;
 * the caller all-reduces them; phase 2 finishes with the element count multiplied by `count_scale` (world size) and
 * dgamma / dbeta multiplied by `grad_scale` (1/world: the later SUM all-reduce of the gradients restores the global value). */
int64_t fabric_b200_bn_bwd_partial_floats(int G, int C);
int fabric_b200_bn_relu_bwd_phase(int phase, const void* z, const void* a, const void* ga, int ga_groups, int ga_channels,
                                  int mul_other, const void* gp, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws,
                                  int G, int B, int H, int W, int C, float count_scale, float grad_scale, void* stream);

/* BatchNorm(train) + ReLU backward when the producer of dL/da already did the reduce: `dy` = relu'(.) * dL/da (bf16) and
 * `partial` = per-CTA (sum dy, sum dy*z) in the conv epilogue's layout [grid][2][n_tile][2], both written by the
 * data-gradient launch of fabric_b200_conv3x3 with bnbwd_z set.  Finishes dgamma / dbeta and writes
 * dz = gamma*invstd * (dy - mean(dy) - xhat*mean(dy*xhat)) in ONE pass over (dy, z).  coef_ws: G*3*C floats.
 * count_scale / grad_scale as in fabric_b200_bn_relu_bwd_phase (exact-global mode all-reduces `partial` first). */
int fabric_b200_bn_bwd_from_partials(const void* z, const void* dy, const float* partial, int grid, int n_tile,
                                     const float* mean, const float* invstd, const float* gamma, void* dz, float* dgamma,
                                     float* dbeta, float* coef_ws, int G, int B, int H, int W, int C, float count_scale,
                                     float grad_scale, void* stream);

/* adjoint of the upsample half of fabric_b200_build_up_input: dcat [B][H][W][Cs+Cl] -> dlow [B][h][w][Cl] */
int fabric_b200_up_input_bwd(const void* dcat, void* dlow, int B, int H, int W, int Cs, int h, int w, int Cl, void* stream);

/* weight gradient of the 3x3 conv on tcgen05: ws[split][Ca][9][Cb] += sum_pix p[pix][ca] * q[pix+tap][cb] */
typedef struct {
  int G, B, H, W;
  int Ca;          /* channels of p = dL/d(conv output) = Cout, multiple of 64 */
  int Cb;          /* padded channels of q = conv input: 16 or a multiple of 64 */
  const void* p;   /* bf16 [G][B][H][W][Ca] */
  const void* q;   /* bf16 [G][B][H][W][Cb] */
  float* ws;       /* fp32 [splits][Ca][9][Cb], fabric_b200_conv3x3_wgrad_ws_floats() */
  int splits;      /* 0 = auto */
  int wide;        /* 0 = three N=64 MMAs per K step, 1 = one N=192 MMA (overlapping N atoms), 2 = second form: the filter
                      row goes through an 18-row Q halo tile, 32-channel Q chunks, three N=96 MMAs per K step (falls back
                      to 1 for Cb = 16 and maps of 8 rows or fewer), 3 = 2 where it measured faster (Ca >= 128 and
                      Cb <= 128), else 1; 4 = 1 with the halo-P tile forced (Ca == 64, H > 8: ONE 18-row P tile per pixel
                      tile serves all three filter rows; modes 1 and 3 choose it by themselves from 16384 pixel tiles up) */
} fb_wgrad_desc;
/* the launch plan on a device with `sms` SMs / `smem_optin` bytes of opt-in shared memory (pure host arithmetic) */
typedef struct {
  int form;        /* 1 = filter row through the P tile, 2 = through the Q halo tile, 3 = form 1 with ONE 18-row P halo tile for
                      all three filter rows (Ca == 64) */
  int grid, items, splits, stages, smem_bytes, tiles_total;
} fb_wgrad_plan;
int fabric_b200_conv3x3_wgrad_plan(const fb_wgrad_desc* d, int sms, int smem_optin, fb_wgrad_plan* out);
int64_t fabric_b200_conv3x3_wgrad_ws_floats(const fb_wgrad_desc* d);
int fabric_b200_conv3x3_wgrad_splits(const fb_wgrad_desc* d);
int fabric_b200_conv3x3_wgrad(const fb_wgrad_desc* d, void* stream);
/* sum the split-K partials and transpose to nn.Conv2d layout: dw [Cout][Cin][3][3] fp32 */
int fabric_b200_wgrad_reduce(const float* ws, int splits, int Cout, int Cin, int CinPad, float* dw, void* stream);
/* the same after an operand-swapped fabric_b200_conv3x3_wgrad call (p = conv input, q = dL/dz, Ca = Cin, Cb = Cout: used
   when dL/dz has 64 channels and the input >= 128, so that the input's channels fill the 128 MMA rows):
   ws [splits][Cin][9][Cout] holds dW'[ci][tap'][co] = dW[co][8 - tap'][ci] */
int fabric_b200_wgrad_reduce_swapped(const float* ws, int splits, int Cout, int Cin, float* dw, void* stream);

/* ---- full-scene inference (reference utils/inference.py, utils/dataloaders.py:94-99, train.py:96-106) -- */

/* Tile gather: scene [C][H][W] (scene_dtype 0 = fp32, 1 = uint16 raw Sentinel-2 digital numbers) -> packed tiles
 * bf16 [N][p][p][Cpad]; tile n starts at (origins[2n], origins[2n+1]) = (row, col).  With mean / inv_std (per band)
 * the z-score (v - mean[c]) * inv_std[c] of city_loader (dataloaders.py:94-99) is applied on the fly.  Replaces
 * _get_patches' host-side extract/vstack (inference.py:134-181) + the NCHW transpose + the layout pack. */
int fabric_b200_gather_tiles(const void* scene, int scene_dtype, const int* origins, void* dst, const float* mean,
                             const float* inv_std, int N, int C, int Cpad, int H, int W, int p, void* stream);
/* torch.max(logits, 1) indices (train.py:96,138,199; ties -> class 0) as uint8 mask (nullable), and with labels
 * (nullable) the confusion counts of the positive class accumulated into counts[4] = (TP, FP, FN, TN) -- what
 * sklearn prfs(average='binary') is computed from at train.py:103-106. */
int fabric_b200_argmax_metrics(const float* logits, const int64_t* labels, uint8_t* mask, uint64_t* counts, int B, int H, int W,
                               void* stream);
/* _get_bands (inference.py:184-236): write tiles [first, first+count) of masks uint8 [N][p][p] into canvas uint8
 * [H][W] at their origins.  The caller launches one call per tile class (grid, last column, last row, corner) in that
 * order, which reproduces the reference's "later write wins" rule deterministically. */
int fabric_b200_scatter_tiles(const uint8_t* masks, const int* origins, uint8_t* canvas, int first, int count, int p, int H,
                              int W, void* stream);

/* Plain SGD p -= lr * grad_scale * g over many tensors in ONE launch (reference train.py:55,95:
 * optim.SGD(model.parameters(), lr), no momentum / weight decay).  `chunks` is a DEVICE array of n_chunks records
 * {float* p; const float* g; int32 n; int32 pad} (16-byte records, <= 65536 elements each): one CTA per record.
 * With grad_scale = 1/world and g pointing into the all-reduced bucket this is the optimizer step fused behind the
 * single NCCL all-reduce. */
int fabric_b200_sgd_step(const void* chunks, int n_chunks, float lr, float grad_scale, void* stream);

/* The whole optimizer side of one training step in ONE launch (reference train.py:95 `optimizer.step()` plus what
 * nn.DataParallel's per-forward replica broadcast implies for the BatchNorm buffers).  `chunks` is a DEVICE array of
 * n_chunks 64-byte records, one CTA each (<= 65536 elements):
 *   { float* p; const float* g; bf16* wf; bf16* wd; int32 n, mode, off, Cout, Cin, CinPad, pad, pad }
 *   mode 0: p -= lr * grad_scale * g
 *   mode 1: p *= stats_scale             (running statistics living in the SUM-all-reduced bucket)
 *   mode 2: mode 0 for elements [off, off+n) of a conv weight [Cout][Cin][3][3], and the same new values written as bf16
 *           into its packed copies wf[Cout][9][CinPad] (fabric_b200_pack_conv3x3_weight mode 0) and wd[Cin][9][Cout]
 *           (mode 1), so that the next step launches no pack kernels. */
int fabric_b200_train_step_update(const void* chunks, int n_chunks, float lr, float grad_scale, float stats_scale,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FABRIC_B200_H_ */
