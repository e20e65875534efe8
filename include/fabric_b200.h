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
/* NHWC bf16 [B][H][W][C] -> NCHW fp32 [B][C][H][W] (standalone block outputs / tests). */
int fabric_b200_unpack_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W, void* stream);

/* nn.Conv2d weight [Cout][Cin][3][3] fp32 -> bf16 GEMM operand.
 *   mode 0 (forward): dst[Cout][9][CinPad], k = tap*CinPad + ci, tap = ky*3+kx
 *   mode 1 (dgrad)  : dst[Cin][9][Cout],    taps flipped, so that dX = conv3x3(dY, dst)            */
int fabric_b200_pack_conv3x3_weight(const float* w, void* dst, int Cout, int Cin, int CinPad, int mode, void* stream);

/* ---- 3x3 convolution (tcgen05 implicit GEMM) ---------------------------------------------------------------- */

typedef struct {
  int n_tile;      /* 0 = auto, else 64 / 128 / 256 */
  int halo;        /* -1 = auto, 0 = per-tap TMA loads, 1 = halo tile shared by the nine taps */
  int a_stages;    /* 0 = auto */
  int b_stages;    /* 0 = auto */
  int b_resident;  /* -1 = auto, 0 / 1 */
  int grid;        /* 0 = auto (#SMs rounded to a multiple of the N tiles) */
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
  fb_conv_tuning tune;
} fb_conv3x3_desc;

/* y = conv3x3(x, w) * scale + shift (+ReLU) (+pool) (+BN moment partials) (+1x1 head).
 * Replaces nn.Conv2d(in,out,3,padding=1) and the BatchNorm2d(eval)/ReLU that follow it
 * (models/unet_parts.py:13-18); with mode-1 weights it is also the data gradient of that conv. */
int fabric_b200_conv3x3(const fb_conv3x3_desc* d, void* stream);
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
 * low has low_groups = 2 (low = product of the two dates, up1) or 1 (previous decoder stage). */
int fabric_b200_build_up_input(const void* skip, const void* low, void* out, int B, int H, int W, int Cs, int h, int w,
                               int Cl, int low_groups, void* stream);

/* outconv (unet_parts.py:83-90): NHWC bf16 [B][H][W][64] -> NCHW fp32 logits [B][2][H][W] */
int fabric_b200_outconv(const void* x, const float* w, const float* b, float* logits, int B, int H, int W, int C,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FABRIC_B200_H_ */
