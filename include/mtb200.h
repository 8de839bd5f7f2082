/*
 * mtb200.h -- C ABI of libmtb200.so: the B200-native (sm_100a) hot path of MultiTalent's 3D U-Net training step and
 * sliding-window predictor.
 *
 * The reference (MIC-DKFZ/MultiTalent, a fork of nnU-Net V1) is 100 % Python over PyTorch library ops: it has no FFI
 * of its own.  Each entry point below therefore cites the reference *call site* (file:line under /root/reference)
 * whose arithmetic it replaces; the Python host side (multitalent_b200/) binds these through ctypes and mirrors the
 * reference's module / trainer interface on top.  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative mtb200_status on failure; the message of the last failure on the
 *    calling thread is returned by mtb200_last_error().  Nothing here aborts, exits or throws across the boundary.
 *  - pointers are raw device pointers unless the name says host; the library owns NO device memory: workspaces are
 *    allocated by the caller (PyTorch's caching allocator) and passed in.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All launches are asynchronous.
 *  - activations are channels-last NDHWC ("voxel-major") with a channel stride `ldc` and a channel offset `coff`, so a
 *    producer can write straight into one half of a concatenated skip buffer (replaces torch.cat, generic_UNet.py:392).
 *  - dtype enum: 0 = float32 (T0 parity mode), 1 = bfloat16, 2 = float16 (T1 production modes).
 */
#ifndef MTB200_H
#define MTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTB200_VERSION 100

typedef enum {
  MTB200_OK = 0,
  MTB200_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  MTB200_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed (launch error etc.) */
  MTB200_ERR_UNSUPPORTED = -3, /* valid request, not supported by this kernel (caller should pick another path) */
} mtb200_status;

typedef enum { MTB200_F32 = 0, MTB200_BF16 = 1, MTB200_F16 = 2 } mtb200_dtype;

#define MTB200_MAX_TAPS 32
#define MTB200_MAX_GROUPS 8

/* One "tap-table" convolution problem:
 *     out[b, o*os + ooff_g, co] (+)= bias[co] + sum_{t in group g} sum_ci  W[widx_t][co][ci] * f(in[b, o*is + off_t, ci])
 * for o over the logical output grid (Do,Ho,Wo); f = optional on-load transform  v = x*scale+shift; v>0 ? v : v*slope
 * (InstanceNorm apply + LeakyReLU of the PRODUCER fused into the consumer's load); reads outside the input volume are 0
 * AFTER the transform (the reference zero-pads the normalised tensor, generic_UNet.py:46,245).
 * With the right tap table this one form covers Conv3d forward (stride 1 / strided), its data gradient (stride 1: flipped
 * taps; strided: 8 parity groups), ConvTranspose3d(kernel==stride) forward (8 groups of 1 tap, os=2) and its data
 * gradient (8 taps, is=2), and the 1x1x1 heads.  Weights are packed [n_widx][Cout][Cin] ("K-major").            */
typedef struct {
  const void* in;       /* NDHWC activations, dtype `dtype` */
  void* out;            /* NDHWC, dtype `dtype` */
  const void* w;        /* packed weights [n_widx][Cout][Cin], dtype `wdtype` */
  const float* bias;    /* [Cout] or NULL */
  const float* xform;   /* [B][Cin][4] = {scale, shift, slope, unused} or NULL (input already normalised) */
  double* stats;        /* [B][Cout][2] running {sum, sum of squares} of the ROUNDED outputs, or NULL */
  int32_t dtype, wdtype;
  int32_t B;
  int32_t Di, Hi, Wi, in_ldc, in_coff, Cin;       /* full input volume */
  int32_t Dof, Hof, Wof, out_ldc, out_coff, Cout; /* full output volume */
  int32_t Do, Ho, Wo;                             /* logical output grid (per group) */
  int32_t is[3], os[3];                           /* input / output coordinate strides */
  int32_t ngroups;
  int32_t group_tap_begin[MTB200_MAX_GROUPS + 1];
  int32_t group_ooff[MTB200_MAX_GROUPS][3];
  int32_t ntaps;
  int32_t tap_off[MTB200_MAX_TAPS][3];
  int32_t tap_widx[MTB200_MAX_TAPS];
  int32_t accumulate;   /* 1: out += result (gradient accumulation into a skip buffer) */
  /* Optional fused InstanceNorm-backward reduction (data-gradient launches): when `red` is non-NULL and the kernel the
   * problem is dispatched to supports it (line-streaming and pointwise tcgen05 kernels -- ask mtb200_last_kernel()), the
   * epilogue treats its FINAL output tile g = d(loss)/d(activation) of the producing layer, whose RAW conv output is
   * red_y (NDHWC, same dtype / grid as `out`), and accumulates
   *     red[b][c][0] += sum_v dv,   red[b][c][1] += sum_v dv * xhat,     dv = g * lrelu'(scale*y + shift), xhat = (y-mean)*rstd
   * with red_xform [B][Cout][4] = {scale, shift, slope, -} and red_meanrstd [B][Cout][2] -- exactly what
   * mtb200_in_bwd_reduce computes in a separate pass over g and y (InstanceNorm3d + LeakyReLU backward, the autograd of
   * generic_UNet.py:63-70).  Kernels without support ignore the fields. */
  const void* red_y;
  const float* red_xform;
  const float* red_meanrstd;
  double* red;
  int32_t red_ldc, red_coff;
  int32_t impl;         /* 0 = auto, 1 = CUDA-core FFMA kernel, 2 = tcgen05 kernels (fastest applicable of the three),
                           3 = tcgen05 per-tap kernel only, 4 = tcgen05 plane-streaming kernel only,
                           5 = tcgen05 line-streaming kernel (dy taps merged into N) only,
                           6 = tcgen05 pointwise (1x1x1) streaming kernel only,
                           7 = tcgen05 group-merged lattice kernel (stride-2 dgrad / ConvTranspose fwd) only */
  /* PLANAR concatenation (the two halves of a decoder input as two compact tensors instead of interleaved channels: a
   * consumer of ONE half then reads whole 128-byte lines; interleaved 32-channel halves cost twice the DRAM traffic).
   * in_split > 0: `in` is [2][B][D][H][W][in_ldc]; input channel c lives in half c / in_split at channel c % in_split
   * (Cin == 2 * in_split, in_coff == 0).  out_split: the same for `out` (data gradient of such a layer).  Only the
   * line-streaming tensor-core kernel implements it (MTB200_ERR_UNSUPPORTED otherwise). */
  int32_t in_split, out_split;
} mtb200_conv_params;

/* Weight-gradient of the same tap-table problem:
 *     dw[widx_t][co][ci] += sum_{b,o} dy[b, o*os + ooff_g, co] * f(x[b, o*is + off_t, ci])      (fp32, atomically)
 * `x`/`xform`/in_* describe the forward input, `dy`/out_* the gradient w.r.t. the forward output.               */
typedef struct {
  const void* x;
  const void* dy;
  float* dw;            /* [n_widx][Cout][Cin] fp32, caller zero-initialises */
  const float* xform;   /* as in mtb200_conv_params */
  int32_t dtype;
  int32_t B;
  int32_t Di, Hi, Wi, in_ldc, in_coff, Cin;
  int32_t Dof, Hof, Wof, out_ldc, out_coff, Cout;
  int32_t Do, Ho, Wo;
  int32_t is[3], os[3];
  int32_t ngroups;
  int32_t group_tap_begin[MTB200_MAX_GROUPS + 1];
  int32_t group_ooff[MTB200_MAX_GROUPS][3];
  int32_t ntaps;
  int32_t tap_off[MTB200_MAX_TAPS][3];
  int32_t tap_widx[MTB200_MAX_TAPS];
  int32_t impl;
  int32_t in_split;     /* planar halves of x, as mtb200_conv_params::in_split */
} mtb200_wgrad_params;

/* ---- library / error handling -------------------------------------------------------------------------------- */
int mtb200_version(void);
const char* mtb200_last_error(void);
/* 1 if the tcgen05 path was compiled in and the current device is sm_100; 0 otherwise. */
int mtb200_has_tcgen05(void);
/* Name of the CUDA kernel family the calling thread's most recent entry point launched (e.g. "conv_line_umma",
 * "wgrad_taps_umma", "conv_taps_ffma") -- what the dispatch inside mtb200_conv_taps / mtb200_wgrad_taps picked.
 * Measurement only (bench.py groups its per-launch timings by it); "" before the first launch. */
const char* mtb200_last_kernel(void);

/* ---- a1/a2/a4/a5: conv -> (stats) ; replaces nn.Conv3d / nn.ConvTranspose3d calls at
 *      generic_UNet.py:57,66 (conv in ConvDropoutNormNonlin), :335-336 + :391 (tu), :350-351 + :394 (seg_outputs),
 *      custom_modules/conv_blocks.py:161-172,188-199 (BasicResidualBlock convs), generic_modular_UNet.py:235-251      */
int mtb200_conv_taps(const mtb200_conv_params* p, void* stream);
/* autograd of the above w.r.t. the weights (replaces cuDNN wgrad behind loss.backward(),
 * MultiTalent_Trainer_DDP.py:350,362) */
int mtb200_wgrad_taps(const mtb200_wgrad_params* p, void* stream);
/* ---- a1 (first layer): Conv3d(1 -> Cout, 3x3x3, stride 1, padding 1) on the single-channel patch
 *      (generic_UNet.py:46 with input_channels = 1, conv_blocks_context.0.blocks.0) and its weight gradient, with the
 *      GEMM K dimension = the 27 taps (im2col tile built in shared memory).  `x` = [B][D][H][W] voxels of `dtype`
 *      (bf16 / fp16), `x_stride` elements apart (1 = compact volume, ldc = channel 0 of an NDHWC buffer).
 *      `w` = packed [27][Cout_p][Cin_p] (only ci = 0 is read); `dw` the same shape in fp32 (atomically accumulated).
 *      out / dy: NDHWC slices [B*D*H*W][ldc] + coff of Cout_p in {16, 32, 64} channels; stats as in mtb200_conv_taps. */
int mtb200_conv_c1_fwd(const void* x, int64_t x_stride, const void* w, int32_t Cin_p, const float* bias, void* out,
                       int32_t out_ldc, int32_t out_coff, int32_t Cout_p, double* stats, int32_t dtype, int32_t B,
                       int32_t D, int32_t H, int32_t W, void* stream);
int mtb200_conv_c1_wgrad(const void* x, int64_t x_stride, const void* dy, int32_t dy_ldc, int32_t dy_coff,
                         int32_t Cout_p, float* dw, int32_t Cin_p, int32_t dtype, int32_t B, int32_t D, int32_t H,
                         int32_t W, void* stream);
/* column sums: out[c] (+)= sum_rows m[row*ldc + coff + c]   (bias gradients), fp32 out */
int mtb200_colsum(const void* m, int32_t dtype, int64_t rows, int32_t ldc, int32_t coff, int32_t C, float* out,
                  void* stream);

/* ---- a1: InstanceNorm3d(eps, affine) + LeakyReLU; replaces generic_UNet.py:63-64,70 / conv_blocks.py:173-186 ------ */
/* stats [B][C][2] doubles {sum, sumsq} over `nvox` voxels -> xform [B][C][4] {scale=gamma*rstd, shift=beta-mean*scale,
 * slope, 0} and meanrstd [B][C][2].  gamma/beta are [C] (C = padded channel count; padded entries must be 0).    */
int mtb200_in_finalize(const double* stats, const float* gamma, const float* beta, int32_t B, int32_t C, int64_t nvox,
                       float eps, float slope, float* xform, float* meanrstd, void* stream);
/* per-(b,c) sum / sum-of-squares of an NDHWC tensor (used when the producer had no stats epilogue) */
int mtb200_in_stats(const void* y, int32_t dtype, int32_t B, int64_t nvox, int32_t ldc, int32_t coff, int32_t C,
                    double* stats, void* stream);
/* out = f(y) with f from xform (materialise the normalised activation); optional residual:
 * out = lrelu_slope2( f(y) + g(res) ) where g is res's own transform (identity if res_xform NULL)  -- the tail of
 * BasicResidualBlock.forward (conv_blocks.py:205-213) */
int mtb200_norm_act(const void* y, int32_t in_ldc, int32_t in_coff, void* out, int32_t out_ldc, int32_t out_coff,
                    int32_t dtype, int32_t B, int64_t nvox, int32_t C, const float* xform, const void* res,
                    int32_t res_ldc, int32_t res_coff, const float* res_xform, float slope2, void* stream);
/* backward of act = lrelu(gamma*xhat+beta): pass 1 accumulates red[B][C][2] doubles {sum dv, sum dv*xhat} */
int mtb200_in_bwd_reduce(const void* dact, int32_t d_ldc, int32_t d_coff, const void* y, int32_t y_ldc, int32_t y_coff,
                         int32_t dtype, int32_t B, int64_t nvox, int32_t C, const float* xform, const float* meanrstd,
                         double* red, void* stream);
/* pass 2: dy = rstd*gamma*(dv - mean(dv) - xhat*mean(dv*xhat)); also dgamma[c] += sum_b red[..][1], dbeta += red[..][0]
 * (done once by the first block).  dy may alias dact. */
int mtb200_in_bwd_apply(const void* dact, int32_t d_ldc, int32_t d_coff, const void* y, int32_t y_ldc, int32_t y_coff,
                        void* dy, int32_t dy_ldc, int32_t dy_coff, int32_t dtype, int32_t B, int64_t nvox, int32_t C,
                        const float* xform, const float* meanrstd, const float* gamma, const double* red,
                        float* dgamma, float* dbeta, void* stream);
/* dv = dact * lrelu'(f(y))  only (no norm): used for the second LeakyReLU of a residual block */
int mtb200_lrelu_bwd(const void* dact, const void* act, void* dv, int32_t dtype, int64_t n, float slope, void* stream);
/* backward of the tail of BasicResidualBlock.forward (custom_modules/conv_blocks.py:205-213, `out += residual;
 * nonlin2(out)`): dv = dact * lrelu'(act) is written (acc = 0) or added (acc = 1) to the gradient slices of the two
 * summands; NDHWC slices [nrows][ldc] + coff of C channels each; either destination may be NULL. */
int mtb200_residual_bwd(const void* dact, int32_t d_ldc, int32_t d_coff, const void* act, int32_t a_ldc, int32_t a_coff,
                        void* dst0, int32_t ldc0, int32_t coff0, int32_t acc0, void* dst1, int32_t ldc1, int32_t coff1,
                        int32_t acc1, int32_t dtype, int64_t nrows, int32_t C, float slope, void* stream);

/* ---- a11: generic nnU-Net loss, softmax + cross entropy + soft Dice; replaces DC_and_CE_loss.forward
 *      (training/loss_functions/dice_loss.py:488-545, :155-195, :100-152; crossentropy.py:4-11) and its autograd ------- */
/* pass 1: stats[b][c][3] doubles += {sum p_c*[y==c], sum p_c, sum [y==c]}, ce_sum[b] += sum -log p_y over the voxels of
 * sample b; p = softmax over the C real classes of the NDHWC logits [B][nvox][ldc]; target = float label map [B][nvox]. */
int mtb200_dcce_stats(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t Cp, const float* target, int32_t B,
                      int64_t nvox, double* stats, double* ce_sum, void* stream);
/* pass 2: dz[b][v][k] = gscale[0] * ( ce_weight * (p_k - [y==k]) + p_k * (G_k - sum_c G_c p_c) ),
 * G_c = coef[b][c][0] * [y==c] + coef[b][c][1]  (the Dice term's d/dp, computed by the caller from the pooled stats) */
int mtb200_dcce_bwd(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t Cp, const float* target, int32_t B,
                    int64_t nvox, const float* coef, float ce_weight, const float* gscale, void* dz, int32_t dz_ldc,
                    void* stream);

/* ---- a9/a10: MultiTalent multi-head loss (sigmoid + BCE + pooled soft Dice); replaces the python loop at
 *      MultiTalent_Trainer_DDP.py:567-594 (stats), :596-606 (Dice), and its autograd ------------------------------ */
/* pass 1: stats[b][j][4] doubles += {sum bce, sum sigma*y, sum sigma, sum y} over the voxels of sample b for every
 * channel j whose bit is set in valid_mask[b]; y = bit j of pos_mask[label].  logits NDHWC [B][nvox][ldc]; target
 * float32 label map [B][nvox] (integer-valued, MultiTalent_Trainer_DDP.py:580-584).
 * `hard` (may be NULL): [B][C][2] doubles += {sum [z > 0] * y, sum [z > 0]} -- the thresholded-prediction counts of
 * run_online_evaluation (MultiTalent_Trainer_DDP.py:372-397: tp = hard[0], fp = hard[1] - hard[0], fn = sum y - hard[0])
 * out of the same pass. */
int mtb200_mt_loss_stats(const void* logits, int32_t dtype, int32_t ldc, int32_t C, const float* target, int32_t B,
                         int64_t nvox, const uint64_t* valid_mask, const uint64_t* pos_mask, int32_t n_labels,
                         double* stats, double* hard, void* stream);
/* tiny on-device finalize for one scale: local stats [B][C][4] + pooled {tp, sigma+y} of ALL ranks
 * pooled[B][C][2] (= sum over ranks of local {tp, sum sigma + sum y}; NULL = derive from the local stats, world 1)
 * -> losses[3] += w*{ce - dc, ce, dc}; coef[B][C][4] = {w/nvox, w*W*2/D, w*W*2*TP/D^2, valid} for pass 2 */
int mtb200_mt_loss_finalize(const double* stats, const double* pooled, const uint64_t* valid_mask, int32_t B, int32_t C,
                            int64_t nvox, float weight, float world_size, float* losses, float* coef, void* stream);
/* pass 2: dlogits[b][v][j] = gscale * ( coef0*(sigma - y) - sigma*(1-sigma)*(y*coef1 - coef2) ) for valid (b,j), else 0 */
int mtb200_mt_loss_bwd(const void* logits, int32_t dtype, int32_t ldc, int32_t C, const float* target, int32_t B,
                       int64_t nvox, const uint64_t* pos_mask, int32_t n_labels, const float* coef, const float* gscale,
                       void* dlogits, int32_t d_ldc, void* stream);

/* a5 + a9 backward in one pass (generic_UNet.py:349-351 backward + MultiTalent_Trainer_DDP.py:567-606 backward): pass 2 of
 * the loss, the head's data gradient and the head's weight gradient of ONE 1x1x1 segmentation head (no bias).  A sample of
 * a partially labelled dataset supervises a contiguous run of output channels; `win_c0[b]` (a multiple of 8) is the first
 * channel of the 16-channel window that contains every supervised channel of sample b (coef[b][j][3] != 0 only inside it).
 *   d[b][v][j]   = mtb200_mt_loss_bwd's formula, rounded to `dtype`, for j in the window of b (never written to memory)
 *   dx[b][v][ci] (+)= sum_j d[b][v][j] * W[j][ci]            W: `w_swap`, packed [Cin][Cout] (mtb200_pack_weights, swap_io)
 *   dw[j][ci]    += sum_{b,v} d[b][v][j] * x[b][v][ci]        dw: packed fp32 [Cout][Cin]
 * 16-bit tensors only, Cin (padded) 32 or 64; MTB200_ERR_UNSUPPORTED otherwise (the caller runs the three separate
 * passes). */
#define MTB200_MAX_HEAD_BATCH 16
typedef struct {
  const void* logits;        /* [B][nvox][z_ldc] */
  const float* target;       /* [B][nvox] label ids */
  const float* coef;         /* [B][C8][4] from mtb200_mt_loss_finalize */
  const float* gscale;       /* device scalar (upstream gradient x loss scale) or NULL */
  const uint64_t* pos_mask;  /* [n_labels] label -> bitmask of positive output channels */
  const void* x;             /* head input (materialised activation) [B][nvox][x_ldc], channels x_coff .. x_coff + Cin */
  const void* w_swap;        /* [Cin][Cout], dtype */
  const void* w_fwd;         /* [Cout][Cin], dtype: needed when `logits` is NULL (deferred head: the window of the logits is
                              * recomputed from x on the tensor core, rounded to dtype -- the values a stored tensor holds) */
  void* dx;                  /* [B][nvox][dx_ldc], channels dx_coff .. dx_coff + Cin */
  float* dw;                 /* [Cout][Cin] */
  int64_t nvox;
  int32_t dtype, B, z_ldc, C8, n_labels, x_ldc, x_coff, Cin, Cout, dx_ldc, dx_coff;
  int32_t accumulate;        /* 1: dx += (another consumer already wrote its share) */
  int32_t win_c0[MTB200_MAX_HEAD_BATCH];
} mtb200_head_bwd_params;
int mtb200_head_bwd_fused(const mtb200_head_bwd_params* p, void* stream);
/* Forward of a DEFERRED head under the MultiTalent loss: loss pass 1 (mtb200_mt_loss_stats' sums, same arithmetic) computed
 * straight from the head's input -- logits window = x . W^T on the tensor core, rounded to dtype, never written to memory.
 * stats[B][C8][4] += {sum bce, sum sigma*y, sum sigma, sum y}; hard (may be NULL): [B][C8][2] += {sum [z>0] y, sum [z>0]}. */
typedef struct {
  const void* x;             /* head input [B][nvox][x_ldc], channels x_coff .. x_coff + Cin */
  const void* w_fwd;         /* [Cout][Cin], dtype (mtb200_pack_weights, forward layout) */
  const float* target;       /* [B][nvox] label ids */
  const uint64_t* valid_mask;/* [B] supervised output channels */
  const uint64_t* pos_mask;  /* [n_labels] */
  double* stats;
  double* hard;
  int64_t nvox;
  int32_t dtype, B, C8, n_labels, x_ldc, x_coff, Cin, Cout;
  int32_t win_c0[MTB200_MAX_HEAD_BATCH];
} mtb200_head_fwd_params;
int mtb200_head_fwd_stats(const mtb200_head_fwd_params* p, void* stream);
/* Inference (a5 + a14/a15): head -> non-linearity -> x weight x Gaussian -> scatter-add of ONE tile into the sliding-window
 * accumulators, the logits never stored (generic_UNet.py:349-351 + neural_network.py:374-394, 531-589):
 *   z[v][c] = round_dtype(bias[c] + sum_ci W[c][ci] x[v][ci]);  f = sigmoid (nonlin 1) / softmax over the C classes (2) / id (0)
 *   acc[c][x0 + d][y0 + h][z0 + w] += weight * gauss[d][h][w] * f(z)[src(d, h, w)][c],   src = the flipped voxel (`flip` bits
 *   0 / 1 / 2 = w / h / d) of the tile that went through the network;  nb[...] += gauss[d][h][w]  (nb may be NULL).
 * Tiles of consecutive launches may overlap (stream order).  16-bit, Cin (padded) 32 or 64, at most 48 padded classes. */
typedef struct {
  const void* x;          /* head input of the tile [pd][ph][pw][x_ldc], channels x_coff .. x_coff + Cin */
  const void* w_fwd;      /* [Cout][Cin], dtype */
  const float* bias;      /* [Cout] or NULL */
  const float* gauss;     /* [pd][ph][pw] or NULL */
  float* acc;             /* [C][X][Y][Z] */
  float* nb;              /* [X][Y][Z] or NULL */
  float weight;
  int32_t dtype, x_ldc, x_coff, Cin, Cout, C, pd, ph, pw, flip, nonlin, X, Y, Z, x0, y0, z0;
} mtb200_head_agg_params;
int mtb200_head_aggregate(const mtb200_head_agg_params* p, void* stream);

/* ---- a14/a15/a16: sliding-window predictor; replaces neural_network.py:374-394 (tile loop + host numpy accumulate),
 *      :531-589 (mirror TTA), :405 (normalise), :415-417 (threshold) ------------------------------------------------ */
/* tile[0][d][h][w][c] = vol[c][x0+fd(d)][y0+fh(h)][z0+fw(w)] with optional flips (bit0: W, bit1: H, bit2: D);
 * vol is float32 [Cin][X][Y][Z] (the reference's (c,x,y,z) numpy contract), tile NDHWC of `dtype` with ldc channels */
int mtb200_sw_gather_tile(const float* vol, int32_t Cin, int32_t X, int32_t Y, int32_t Z, int32_t x0, int32_t y0,
                          int32_t z0, int32_t pd, int32_t ph, int32_t pw, int32_t flip, void* tile, int32_t dtype,
                          int32_t ldc, void* stream);
/* acc[c][x0+d][y0+h][z0+w] += weight * gauss[d][h][w] * nonlin(logits[fd(d)][fh(h)][fw(w)][:])[c],  c < C;
 * if nb != NULL also nb[x0+d][..] += gauss[d][h][w]  (one weight volume instead of the reference's 47 copies).
 * gauss may be NULL (= 1).  apply_sigmoid = the inference non-linearity (`inference_apply_nonlin`, neural_network.py:
 * 531-586): 0 accumulates the raw values, 1 = sigmoid (MultiTalent, MultiTalent_Trainer_DDP.py:46), 2 = softmax over
 * the C channels (softmax_helper, nnUNetTrainerV2.py:162 -- the fine-tuning trainers). */
int mtb200_sw_aggregate(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t pd, int32_t ph, int32_t pw,
                        int32_t flip, const float* gauss, float weight, int32_t apply_sigmoid, float* acc, float* nb,
                        int32_t X, int32_t Y, int32_t Z, int32_t x0, int32_t y0, int32_t z0, void* stream);
/* acc[c][v] /= nb[v] in place; seg[v] = class_order[last c with prob > 0.5] else 0 (float32, neural_network.py:415-417)
 * or argmax if class_order == NULL (seg then holds the channel index as float). */
int mtb200_sw_finalize(float* acc, const float* nb, int32_t C, int64_t nvox, const float* class_order, float* seg,
                       void* stream);
/* The same for a SLAB of the volume: `acc`, `nb`, `seg` point at the slab's first voxel, the classes of `acc` are
 * `class_stride` voxels apart (the whole volume), `nvox` voxels are finalised.  Lets the predictor normalise, threshold and
 * ship the planes no later tile can touch while the remaining tiles are still being computed. */
int mtb200_sw_finalize_slab(float* acc, const float* nb, int32_t C, int64_t class_stride, int64_t nvox,
                            const float* class_order, float* seg, void* stream);

/* ---- a12: clip_grad_norm_(12) + SGD(nesterov) on a flat fp32 parameter arena; replaces
 *      MultiTalent_Trainer_DDP.py:351-353 / nnUNetTrainerV2.py:166-170 ---------------------------------------------- */
int mtb200_sumsq(const float* g, int64_t n, double* out /* [1], caller zeroes */, void* stream);
/* coef = min(1, max_norm/(sqrt(sumsq)*inv_scale + 1e-6)) * inv_scale; g' = g*coef + wd*p; buf = first ? g' : m*buf+g';
 * p -= lr*(g' + m*buf).  skip the whole update if sumsq is not finite (GradScaler.step semantics).  `dyn_scale` (device,
 * may be NULL): the current dynamic loss scale; inv_scale is divided by dyn_scale[0] on the device (GradScaler.unscale_,
 * MultiTalent_Trainer_DDP.py:351) so that fp16 training needs no host synchronisation. */
int mtb200_sgd_step(float* p, const float* g, float* buf, int64_t n, const double* sumsq, float inv_scale,
                    float max_norm, float lr, float momentum, float weight_decay, int32_t first_step,
                    const float* dyn_scale, void* stream);
/* torch.cuda.amp.GradScaler.update (MultiTalent_Trainer_DDP.py:354) on the device: state = {scale, growth_tracker,
 * found_inf of this step, skipped steps}; sumsq not finite -> scale *= backoff_factor, tracker = 0; else tracker += 1 and
 * scale *= growth_factor every `growth_interval` clean steps. */
int mtb200_loss_scale_update(const double* sumsq, float* state, float growth_factor, float backoff_factor,
                             int32_t growth_interval, void* stream);

/* ---- layout plumbing ------------------------------------------------------------------------------------------- */
/* reference weight [Cout][Cin][kd][kh][kw] fp32 (Conv3d) or [Cin][Cout][kd][kh][kw] (ConvTranspose3d, transposed=1)
 * -> packed [ntap][Cout_p][Cin_p] of `wdtype`, zero padded.  swap_io=1 packs the transpose [ntap][Cin_p][Cout_p]
 * (operand of the data-gradient problem).  split/split_p describe a concatenated input whose two halves are padded
 * separately: logical input channel ci >= split lives at packed index ci - split + split_p (split = 0: no split). */
int mtb200_pack_weights(const float* w, int32_t Cout, int32_t Cin, int32_t ntap, int32_t transposed, int32_t swap_io,
                        void* packed, int32_t wdtype, int32_t Cout_p, int32_t Cin_p, int32_t split, int32_t split_p,
                        void* stream);
/* One launch for every convolution of a network (the optimizer step invalidates all packed copies at once).  `descs` is
 * a DEVICE array of `n` descriptors sorted by blk_begin; descriptor i owns thread blocks [blk_begin_i, blk_begin_{i+1}),
 * (Cout_p/16) * (Cin_p/16) of them, `total_blocks` in all.  Either destination may be NULL; ntap <= 32; Cout_p and Cin_p
 * are multiples of 16.  Same element semantics as mtb200_pack_weights. */
typedef struct {
  const float* w;
  void* packed;        /* [ntap][Cout_p][Cin_p] */
  void* packed_swap;   /* [ntap][Cin_p][Cout_p] */
  int32_t Cout, Cin, ntap, transposed, Cout_p, Cin_p, split, split_p;
  int32_t blk_begin, reserved;
} mtb200_pack_desc;
int mtb200_pack_weights_batched(const mtb200_pack_desc* descs, int32_t n, int32_t total_blocks, int32_t wdtype,
                                void* stream);
/* packed fp32 gradient [ntap][Cout_p][Cin_p] -> reference layout, grad (+)= scale * dw */
int mtb200_unpack_wgrad(const float* dw, int32_t Cout, int32_t Cin, int32_t ntap, int32_t transposed, int32_t Cout_p,
                        int32_t Cin_p, int32_t split, int32_t split_p, float scale, int32_t accumulate, float* grad,
                        void* stream);
/* The same for every layer of a step in ONE launch: grad += dw per descriptor.  Layer i owns blocks
 * [blk_begin_i, blk_begin_{i+1}) of MTB200_UNPACK_CHUNK reference-layout elements each; `descs` lives in device memory. */
#define MTB200_UNPACK_CHUNK 4096
typedef struct {
  const float* dw; /* packed [ntap][Cout_p][Cin_p] fp32 */
  float* grad;     /* reference layout ([Cout][Cin][taps] or, transposed, [Cin][Cout][taps]) */
  int32_t Cout, Cin, ntap, transposed, Cout_p, Cin_p, split, split_p;
  int32_t blk_begin, reserved;
} mtb200_unpack_desc;
int mtb200_unpack_wgrad_batched(const mtb200_unpack_desc* descs, int32_t n, int32_t total_blocks, void* stream);
/* NCDHW fp32 [B][C][nvox] <-> NDHWC `dtype` [B][nvox][ldc] (+coff); padded channels are written as 0 */
int mtb200_ncdhw_to_ndhwc(const float* src, int32_t B, int32_t C, int64_t nvox, void* dst, int32_t dtype, int32_t ldc,
                          int32_t coff, int32_t Cp, void* stream);
int mtb200_ndhwc_to_ncdhw(const void* src, int32_t dtype, int32_t ldc, int32_t coff, int32_t B, int32_t C, int64_t nvox,
                          float* dst, void* stream);

/* ---- the steps either side of the hot path on device-resident volumes ------------------------ */
/* patch crop + pad out of a preprocessed case [C][X][Y][Z] fp32: dst[c][i][j][k] = src[c][lb + (i,j,k)] where that lies
 * inside the case, else pad_values[c] (edge_mode 0: np.pad 'constant') or the nearest edge voxel (edge_mode 1: 'edge');
 * replaces the slice + np.pad pairs of DataLoader3D.generate_train_batch, training/dataloading/dataset_loading.py:
 * 340-378 (data channels padded with `pad_kwargs_data`, the label channel with -1). */
int mtb200_crop_pad(const float* src, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t lbx, int32_t lby, int32_t lbz,
                    float* dst, int32_t pd, int32_t ph, int32_t pw, int32_t edge_mode, const float* pad_values,
                    void* stream);
/* nearest-neighbour resize of NC label volumes [X][Y][Z] -> [X2][Y2][Z2] with skimage's pixel-centre convention
 * (src index = floor((o + 0.5) * in / out), clamped); replaces resize_segmentation(order 0) inside
 * downsample_seg_for_ds_transform2, training/data_augmentation/downsampling.py:87-104 (deep-supervision targets). */
int mtb200_resize_nearest(const float* src, int64_t NC, int32_t X, int32_t Y, int32_t Z, float* dst, int32_t X2,
                          int32_t Y2, int32_t Z2, void* stream);
/* probability volume [C][X][Y][Z] fp32 -> original grid [X2][Y2][Z2] with per-axis interpolation order (0 nearest,
 * 1 linear; "separate z" = order 0 along the low-resolution axis), pixel-centre coordinates, edge clamping; optional
 * resampled probabilities (`prob`, fp32 or fp16 as the reference's npz) and the label map seg[v] = class_order[last c with
 * p_c > 0.5] (class_order NULL: argmax); replaces resample_data_or_seg + the threshold loop of
 * save_segmentation_nifti_from_softmax, inference/segmentation_export.py:77-123, preprocessing/preprocessing.py:109-197. */
int mtb200_resample_probs(const float* src, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t X2, int32_t Y2, int32_t Z2,
                          int32_t order_x, int32_t order_y, int32_t order_z, void* prob, int32_t prob_is_f16,
                          const float* class_order, uint8_t* seg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTB200_H */
