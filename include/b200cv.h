/* b200cv.h -- C ABI of libb200cv.so, the B200 (sm_100a) hot path behind the Python model/loss
 * API of cv-core/MIT-Driverless-CV-TrainingInfra.
 *
 * The reference has no native interface for this path: every entry point below replaces a
 * PyTorch (ATen/cuDNN) call site of the reference, cited as <file>:<line> relative to the
 * reference tree.  The Python binding a maintainer would add is a ctypes.CDLL stub (see
 * INTEGRATION.md and mit-driverless-cv-traininginfra_b200/b200cv/lib.py).
 *
 * Conventions
 *   - every function returns 0 on success; >0 is a cudaError_t, <0 a library code (below);
 *     b200cv_last_error() returns a thread-local message.  Nothing throws, nothing exits.
 *   - all pointers are DEVICE pointers unless a name ends in _host; the library never allocates
 *     or frees tensor memory and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - activations are NHWC bf16 with the channel count padded by b200cv_pad_channels();
 *     "packed" weights are bf16 [rows][taps][channels] made by b200cv_pack_weights().
 *   - fp32-parity mode ("split"): an fp32 value v is stored as THREE bf16 numbers p0 = bf16(v),
 *     p1 = bf16(v - p0), p2 = bf16(v - p0 - p1) -- 3 x 8 = 24 mantissa bits, fp32's own precision.  A split
 *     activation row is [p0(C) | p1(C) | p2(C)]: piece j lies j * `lo` elements after p0 (lo == C for a whole
 *     tensor, the total width of the buffer for a channel slice of a concat buffer).  Entry points that take a
 *     `*_lo` piece stride treat EVERY activation operand of the call as split when it is non-zero; 0 selects the
 *     default bf16 mode.  Convolutions run six tensor-core passes (all piece products x_i * w_j with i + j <= 2,
 *     fp32 accumulation in tensor memory; the dropped products are below 2^-24 of the result), element-wise
 *     kernels compute on the summed pieces in fp32 and re-split what they store.  This is the mode whose results
 *     match the reference's fp32 arithmetic (CVC-YOLOv3/models.py:59-69 is fp32 end to end): losses to 1e-3 and
 *     better, gradients as closely as the reference's own fp32 rounding allows (tests/test_gpu_fp32_mode.py).
 */
#ifndef B200CV_H_
#define B200CV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CV_OK 0
#define B200CV_ERR_ARG (-1)     /* bad argument / unsupported shape            */
#define B200CV_ERR_DRIVER (-2)  /* driver entry point or tensor-map encode failed */
#define B200CV_ERR_DEVICE (-3)  /* a kernel reported a pipeline timeout        */

#define B200CV_ACT_NONE 0
#define B200CV_ACT_LEAKY 1
#define B200CV_ACT_RELU 2

#define B200CV_DT_BF16 0
#define B200CV_DT_F32 1

/* One entry of a per-channel statistics matrix (see b200cv_conv_args.stats): 16 bytes, zero = empty. */
typedef struct b200cv_stat {
  int64_t w1, w2;
} b200cv_stat;

/* ---- library ------------------------------------------------------------------------- */
const char* b200cv_version(void);
const char* b200cv_last_error(void);
/* Reads (and clears) the per-device error word the tcgen05/TMA pipelines write on a bounded-wait
 * timeout.  Synchronises `stream`.  Test/diagnostic use only -- never on the hot path. */
int b200cv_check_device_error(void* stream);
/* 16 for c<=16, 32 for c<=32, otherwise the next multiple of 64. */
int b200cv_pad_channels(int c);
/* Number of bf16 pieces of a split (fp32-parity mode) value: 3. */
int b200cv_split_pieces(void);

/* ---- layout ---------------------------------------------------------------------------- */
/* NCHW fp32 [N,C,H,W] -> NHWC bf16 [N,H,W,Cpad] (extra channels zero).
 * Replaces the implicit layout of `imgs.to(device)` CVC-YOLOv3/train.py:60, RektNet/train_eval.py:60. */
int b200cv_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int Cpad,
                                 void* stream);
/* the same into / out of a split tensor [N,H,W,3*Cpad] (fp32-parity mode). */
int b200cv_nchw_f32_to_nhwc_split(const float* src, void* dst, int N, int C, int H, int W, int Cpad, void* stream);
int b200cv_nhwc_split_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int Cpad, void* stream);
/* NHWC bf16 [N,H,W,Cpad] (first C channels) -> NCHW fp32 [N,C,H,W]. */
int b200cv_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int Cpad,
                                 void* stream);
/* Explicit im2col for few-channel inputs (the 3-channel image layers: CVC-YOLOv3/models.py conv_0,
 * RektNet/keypoint_net.py:17): NCHW fp32 x -> bf16 patch matrix [N*OH*OW][Kp], k = (r*S+s)*C + c, zero padded.
 * The convolution then runs as a 1x1 conv over the patch matrix (Cin = Kp). */
int b200cv_im2col_nchw_f32(const float* x, void* patches, int N, int C, int H, int W, int R, int S, int stride,
                           int pad, int dil, int Kp, void* stream);
/* The image layers WITHOUT a patch matrix (bf16 mode): k x k (3 or 7), stride 1, pad (k-1)/2, 3-channel fp32 NCHW x,
 * 16 or 32 filters.  Replaces im2col + 1x1 conv for CVC-YOLOv3/models.py:59-69 (conv_0) and RektNet/keypoint_net.py:17
 * (the 7x7 stem): a CTA stages the halo of an 8x32-pixel tile in shared memory and the warps gather their mma.sync
 * fragments from it.  w_flat = the FLAT pack [Cout][Kp] (k = tap*3 + c, see b200cv_pack_entry.transpose == 2).
 *   fwd:   y[N,H,W,y_ld] bf16 = act(conv(x, w) * scale[c] + shift[c]) (scale / shift may be null); `stats` (or null) gets
 *          the per-channel sum / sum of squares of the stored values, [stats_parts][2*Cout] b200cv_stat.
 *   wgrad: dw_flat[Cout][Kp] fp32 += dy^T * patches (the flat gradient layout of b200cv_unpack_wgrad_multi).
 * b200cv_conv_image_supported() tells whether a layer shape takes this path (1) or the im2col path (0). */
int b200cv_conv_image_supported(int C, int R, int S, int stride, int pad, int dil, int Cout);
int b200cv_conv_image_fwd(const float* x, const void* w_flat, int N, int C, int H, int W, int R, int S, int pad,
                          int dil, int Cout, int Kp, void* y, int64_t y_ld, const float* scale, const float* shift,
                          int act, float slope, void* stats, int stats_parts, void* stream);
int b200cv_conv_image_wgrad(const float* x, const void* dy, int64_t dy_ld, int N, int C, int H, int W, int R, int S,
                            int pad, int dil, int Cout, int Kp, float* dw_flat, void* stream);
/* wgrad with the BatchNorm + activation backward of the SAME layer applied on the fly: `da` = dL/da of the layer's
 * activation, dy = BN'(da, y) is formed in registers (the constants, d(gamma), d(beta) and coef exactly as
 * b200cv_bn_bwd_stats_apply computes them from `partials`) and never written -- an image layer has no data gradient,
 * so nothing else reads it.  Replaces b200cv_bn_bwd_stats_apply + b200cv_conv_image_wgrad for conv_0 / the stem. */
int b200cv_conv_image_wgrad_bn(const float* x, const void* da, int64_t da_ld, const void* y, int64_t y_ld,
                               const void* partials, int nparts, int64_t count, const float* gamma, const float* scale,
                               const float* shift, const float* mean, const float* rstd, int act, float slope,
                               float* coef, float* dgamma, float* dbeta, int N, int C, int H, int W, int R, int S,
                               int pad, int dil, int Cout, int Kp, float* dw_flat, void* stream);
/* fp32-parity mode: split patch matrix [N*OH*OW][p0(Kp) | p1(Kp) | p2(Kp)]. */
int b200cv_im2col_nchw_f32_split(const float* x, void* patches, int N, int C, int H, int W, int R, int S, int stride,
                                 int pad, int dil, int Kp, void* stream);
/* OIHW fp32 conv weight -> packed bf16.
 *   transpose == 0: [O][R*S][Ipad]      (forward operand;  K index = tap*Ipad + i)
 *   transpose == 1: [I][R*S][Opad]      (data-gradient operand; K index = tap*Opad + o)
 * nn.Conv2d weights: CVC-YOLOv3/models.py:59-65, RektNet/keypoint_net.py:17,25, resnet.py:12-18. */
int b200cv_pack_weights(const float* w_oihw, void* dst, int O, int I, int R, int S, int Ipad,
                        int Opad, int transpose, void* stream);
/* packed fp32 gradient [O][R*S][Ipad] -> OIHW fp32 (the .grad layout torch.optim expects). */
int b200cv_unpack_wgrad(const float* dw_packed, float* dw_oihw, int O, int I, int R, int S,
                        int Ipad, void* stream);

/* Multi-tensor forms: one launch packs (or un-packs) every conv weight of a network.  `table_dev` is a
 * DEVICE array of n entries; pack: src = OIHW fp32 weight, dst = packed bf16; unpack: src = packed fp32
 * gradient [O][R*S][Ipad], dst = OIHW fp32 gradient (transpose/Opad ignored). */
typedef struct b200cv_pack_entry {
  const float* src;
  void* dst;
  int32_t O, I, RS, Ipad, Opad, transpose;
} b200cv_pack_entry;
/* entry.transpose == 3 (3x3 filters, bf16 mode): the depth-to-space operand of b200cv_conv_dgrad_d2s,
 * dst[(a*2+b)*I + i][(u*2+v)*Opad + o] = w[o][i][r(a,u)][s(b,v)] or 0, r(0,0)=1, r(1,0)=2, r(1,1)=0.
 * entry.transpose == 2 selects the FLAT layout used with b200cv_im2col_nchw_f32: packed row = [Ipad] with
 * k = tap*I + i (Ipad >= R*S*I); the matching gradient row is un-packed the same way.
 * transpose | 8 (also for b200cv_pack_weights) = fp32-parity (split) pack: every innermost run of W channels becomes
 * [w0(W) | w1(W) | w2(W)], the three bf16 pieces of the fp32 weight. */
int b200cv_pack_weights_multi(const void* table_dev, int n, void* stream);
int b200cv_unpack_wgrad_multi(const void* table_dev, int n, void* stream);

/* ---- convolution (tcgen05 implicit GEMM) ------------------------------------------------- */
typedef struct b200cv_conv_args {
  /* input activation: NHWC bf16 [N,H,W,Cin], Cin == b200cv_pad_channels(true Cin) */
  int32_t N, H, W, Cin;
  int32_t Cout; /* true number of output channels */
  int32_t R, S, stride, pad, dil;
  const void* x;
  const void* w; /* packed bf16, see b200cv_pack_weights */
  /* output [N,OH,OW,Cout] addressed by element strides (NHWC slices, NCHW, ... all work) */
  void* y;
  int32_t y_dtype; /* B200CV_DT_* */
  int64_t y_sn, y_sh, y_sw, y_sc;
  /* fused epilogue:  y = act(acc*scale[c] + shift[c] + residual)   (see res_after_act) */
  const float* scale;   /* [Cout] or NULL */
  const float* shift;   /* [Cout] or NULL (conv bias / folded BN) */
  const void* residual; /* bf16 or NULL, addressed like y with its own strides */
  int64_t r_sn, r_sh, r_sw, r_sc;
  int32_t act;
  float slope;
  int32_t res_after_act; /* 1: y = act(acc*scale+shift) + residual (darknet shortcut after the activation) */
  /* per-channel [sum(y) | sum(y*y)] over N*OH*OW, ADDED into stats[stats_parts][2*Cout] (zero it first;
   * CTA b adds into row b % stats_parts); NULL = off.  Cout <= 1024 when stats are requested.
   * Every entry of a statistics matrix is a b200cv_stat: two int64 words of a fixed-point sum
   * (value = w1 * 2^-20 + w2 * 2^-70) added with INTEGER atomics, so the totals -- and the whole forward pass --
   * are bit-reproducible whatever the CTA scheduling (csrc/stat_acc.cuh). */
  void* stats;
  int32_t stats_parts;
  /* b200cv_conv_dgrad only -- fused first pass of the BatchNorm backward of the layer that PRODUCED the
   * activation whose gradient this call writes (b200cv_bn_bwd_reduce folded into the epilogue): with
   * G = the stored output (bf16), z = bn_y*bn_scale+bn_shift, dz = G*act'(z) the epilogue ADDS
   * [sum dz | sum dz*(bn_y-bn_mean)*bn_rstd] per channel into bn_sums[bn_parts][2*Cout] (zero it first).
   * bn_y: NHWC bf16 rows with pitch bn_y_ld laid out like the output.  Needs the row-major bf16 output form
   * (stride-1 gradients); otherwise the call fails with B200CV_ERR_ARG.  bn_sums == NULL = off. */
  void* bn_sums; /* b200cv_stat [bn_parts][2*Cout] */
  int32_t bn_parts;
  const void* bn_y;
  int64_t bn_y_ld;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_rstd;
  int32_t bn_act;
  float bn_slope;
  /* fp32-parity ("split") mode -- see the top of this header.  x_lo != 0: x is a split tensor (row =
   * [p0(Cin) | p1(Cin) | p2(Cin)], x_lo == Cin, pixel pitch 3*Cin) and w a split pack (K run of a tap =
   * [w0(Cin) | w1(Cin) | w2(Cin)]); the GEMM accumulates the six products x_i*w_j, i + j <= 2, in fp32.  y_lo != 0:
   * the bf16 output is written split with piece stride y_lo (direct-store epilogue).  r_lo: the residual is split.
   * All three are 0 in the default bf16 mode. */
  int64_t x_lo, y_lo, r_lo;
} b200cv_conv_args;

/* y = conv2d(x, w).  nn.Conv2d forward: CVC-YOLOv3/models.py:59-65,320-321;
 * RektNet/keypoint_net.py:59,64, resnet.py:22-26. */
int b200cv_conv_fwd(const b200cv_conv_args* a, void* stream);

/* Data gradient.  Here x = dY (NHWC bf16 [N,OH,OW,Cin=pad(Cout_fwd)]), w = transposed pack
 * [Cin_fwd][R*S][pad(Cout_fwd)], Cout = Cin_fwd, (H,W) = spatial size of dY, and (out_h,out_w) the
 * spatial size of dX; stride/pad/dil are the FORWARD conv's.  residual adds an existing gradient
 * (shortcut / route fan-out).  autograd of nn.Conv2d: CVC-YOLOv3/train.py:70. */
int b200cv_conv_dgrad(const b200cv_conv_args* a, int out_h, int out_w, void* stream);
/* The 3x3 stride-2 pad-1 data gradient as ONE launch (instead of one per output-parity class): a GEMM over the dy
 * grid whose 4*C output columns are the 2x2 block of input pixels each dy pixel owns, stored depth-to-space by 3-D TMA
 * stores.  `a->w` = the `transpose == 3` pack of b200cv_pack_weights_multi ([4*Cin_fwd][4*pad(Cout_fwd)]).  Needs an
 * even output whose width is a multiple of 8 and >= 64, C = a->Cout a padded multiple of 32, a contiguous NHWC bf16 y and no residual /
 * fused epilogue (B200CV_ERR_ARG otherwise -- the caller then takes b200cv_conv_dgrad).  CVC-YOLOv3 down-sampling
 * convs, models.py:59-65 with stride=2. */
int b200cv_conv_dgrad_d2s(const b200cv_conv_args* a, int out_h, int out_w, void* stream);

/* Weight gradient: dw_packed[o][tap][i] += sum_pixels dY[pix][o] * X[pix+tap][i]  (fp32, atomically
 * accumulated; zero it first).  x: NHWC bf16 [N,H,W,Cin]; dy: NHWC bf16 [N,OH,OW,dy_ld].
 * fp32-parity mode: x_lo == Cin (x rows are [p0|p1|p2], pitch 3*Cin) and dy_lo = piece stride of dy (dy_ld covers
 * all pieces); 0, 0 in the bf16 mode.  The gradient is fp32 either way.
 * autograd of nn.Conv2d: CVC-YOLOv3/train.py:70, RektNet/train_eval.py:71. */
int b200cv_conv_wgrad(const void* x, const void* dy, float* dw_packed, int N, int H, int W, int Cin,
                      int Cout, int dy_ld, int R, int S, int stride, int pad, int dil, int64_t x_lo, int64_t dy_lo,
                      void* stream);

/* ---- BatchNorm (training mode) + activation, NHWC bf16 rows with explicit pitch -------------- */
/* From the conv epilogue's [sum | sumsq] produce scale/shift for the apply pass, the saved
 * mean/rstd for backward, and update running stats (momentum, unbiased var) -- nn.BatchNorm2d in
 * train(): CVC-YOLOv3/models.py:67, RektNet/keypoint_net.py:19, resnet.py:13,16,20.
 * conv_bias (optional) is the bias of a conv whose bias was folded out (it cancels in train-mode BN). */
int b200cv_bn_finalize(const void* stats, int stats_parts, int64_t count, const float* gamma, const float* beta,
                       const float* conv_bias, float eps, float momentum, float* running_mean,
                       float* running_var, float* scale, float* shift, float* save_mean, float* save_rstd,
                       int C, void* stream);
/* out = act(y*scale+shift [+ y2*scale2+shift2]) [+ post]   (LeakyReLU/ReLU: models.py:69-71,
 * resnet.py:24-26; `post` = darknet shortcut add models.py:325-327). */
int b200cv_bn_apply_act(const void* y, int64_t y_ld, const float* scale, const float* shift, const void* y2,
                        int64_t y2_ld, const float* scale2, const float* shift2, const void* post,
                        int64_t post_ld, void* out, int64_t out_ld, int64_t rows, int C, int act, float slope,
                        void* stream);
/* backward pass 1: partials[p][c] = sum dz, partials[p][C+c] = sum dz*xhat over the rows block p owns
 * (p < nparts, every row of partials is written), with dz = da*act'(z).  z is recomputed from
 * y*scale+shift, or its sign taken from `aout` (saved activation output) when given. */
int b200cv_bn_bwd_reduce(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                         int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                         const float* rstd, void* partials, int nparts, int64_t rows, int C, int act, float slope,
                         void* stream);
/* coef = [gamma*rstd | sum_dz/M | sum_dz_xhat/M] from the nparts partial rows; dgamma/dbeta written if non-NULL. */
int b200cv_bn_bwd_finalize(const void* partials, int nparts, const float* gamma, const float* rstd, int64_t count,
                           float* coef,
                           float* dgamma, float* dbeta, int C, void* stream);
/* backward pass 2: dy = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)). */
int b200cv_bn_bwd_apply(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                        int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                        const float* rstd, const float* coef, void* dy, int64_t dy_ld, int64_t rows, int C,
                        int act, float slope, void* stream);
/* Two BatchNorms under one activation, out = act(bn_A(yA) + bn_B(yB)) (RektNet/resnet.py:21-27: shortcut BN + second
 * conv's BN, one ReLU): both layers see dz = da * act'(aout).  One pass for both layers instead of one per layer;
 * partials / coef / dy as in b200cv_bn_bwd_reduce / b200cv_bn_bwd_apply, one set per layer. */
int b200cv_bn_bwd_reduce2(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, const void* yA,
                          int64_t yA_ld, const void* yB, int64_t yB_ld, const float* meanA, const float* rstdA,
                          const float* meanB, const float* rstdB, void* partialsA, void* partialsB, int nparts,
                          int64_t rows, int C, int act, float slope, void* stream);
int b200cv_bn_bwd_apply2(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, const void* yA, int64_t yA_ld,
                         const void* yB, int64_t yB_ld, const float* meanA, const float* rstdA, const float* meanB,
                         const float* rstdB, const float* coefA, const float* coefB, void* dyA, int64_t dyA_ld,
                         void* dyB, int64_t dyB_ld, int64_t rows, int C, int act, float slope, void* stream);
/* Fused forms used by the Darknet engine (one launch instead of finalize + apply): every block first folds the
 * [parts][2C] partial statistics into the per-channel constants in shared memory, block 0 also publishes them
 * (scale/shift/save_mean/save_rstd, running statistics; dgamma/dbeta/coef) exactly like b200cv_bn_finalize /
 * b200cv_bn_bwd_finalize, then the block streams its rows like b200cv_bn_apply_act / b200cv_bn_bwd_apply. */
int b200cv_bn_stats_apply_act(const void* stats, int stats_parts, int64_t count, const float* gamma,
                              const float* beta, const float* conv_bias, float eps, float momentum,
                              float* running_mean, float* running_var, float* scale, float* shift,
                              float* save_mean, float* save_rstd, const void* y, int64_t y_ld, const void* post,
                              int64_t post_ld, void* out, int64_t out_ld, int64_t rows, int C, int act,
                              float slope, void* stream);
int b200cv_bn_bwd_stats_apply(const void* partials, int nparts, int64_t count, const float* gamma, float* coef,
                              float* dgamma, float* dbeta, const void* da, int64_t da_ld, const void* y,
                              int64_t y_ld, const float* scale, const float* shift, const float* mean,
                              const float* rstd, void* dy, int64_t dy_ld, int64_t rows, int C, int act,
                              float slope, void* stream);
/* dz = da * act'(aout) for an activation that does not follow a BatchNorm. */
int b200cv_act_bwd(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, void* dz, int64_t dz_ld,
                   int64_t rows, int C, int act, float slope, void* stream);
/* channel-slice copy (route/concat, models.py:322-324) or accumulate (gradient fan-in). */
int b200cv_copy_slice(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows, int C,
                      int accumulate, void* stream);
/* out[c] += sum over rows (conv bias gradient). */
int b200cv_col_sum(const void* x, int64_t ld, int64_t rows, int C, float* out, void* stream);
/* 2x2 max-pool, stride 2, or stride 1 with a ZERO pad right/bottom (models.py:74-84). */
int b200cv_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int stride, void* stream);
int b200cv_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int stride,
                          void* stream);
/* nearest x2 upsample (models.py:86-88). */
int b200cv_upsample2x_fwd(const void* x, void* y, int64_t y_ld, int N, int H, int W, int C, void* stream);
int b200cv_upsample2x_bwd(const void* dy, int64_t dy_ld, void* dx, int N, int H, int W, int C, int accumulate,
                          void* stream);

/* ---- fp32-parity ("split") forms of the element-wise kernels (csrc/split_ops.cu) ----------------------------
 * Same arithmetic as the bf16 entry points above on split rows: piece j of every activation operand of a call lies
 * j * `lo` elements after piece 0 (row pitch >= 2*lo + C).  b200cv_bn_finalize / b200cv_bn_bwd_finalize are shared. */
int b200cv_split_bn_apply_act(const void* y, int64_t y_ld, const float* scale, const float* shift, const void* y2,
                              int64_t y2_ld, const float* scale2, const float* shift2, const void* post,
                              int64_t post_ld, void* out, int64_t out_ld, int64_t rows, int C, int64_t lo, int act,
                              float slope, void* stream);
int b200cv_split_bn_bwd_reduce(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                               int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                               const float* rstd, void* partials, int nparts, int64_t rows, int C, int64_t lo, int act,
                               float slope, void* stream);
int b200cv_split_bn_bwd_apply(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                              int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                              const float* rstd, const float* coef, void* dy, int64_t dy_ld, int64_t rows, int C,
                              int64_t lo, int act, float slope, void* stream);
int b200cv_split_act_bwd(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, void* dz, int64_t dz_ld,
                         int64_t rows, int C, int64_t lo, int act, float slope, void* stream);
/* channel-slice copy / accumulate between split buffers with their own lo offsets (route concat, gradient fan-in). */
int b200cv_split_copy_slice(const void* src, int64_t src_ld, int64_t src_lo, void* dst, int64_t dst_ld,
                            int64_t dst_lo, int64_t rows, int C, int accumulate, void* stream);
/* fp32 rows [rows][src_ld] -> split rows, and back (head gradients, module boundaries). */
int b200cv_split_from_f32(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t dst_lo, int64_t rows,
                          int C, void* stream);
int b200cv_split_to_f32(const void* src, int64_t src_ld, int64_t src_lo, float* dst, int64_t dst_ld, int64_t rows,
                        int C, void* stream);
/* whole split tensors [N,H,W,3C] (lo == C). */
int b200cv_split_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int stride, void* stream);
int b200cv_split_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int stride,
                                void* stream);
int b200cv_split_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, int accumulate, void* stream);

/* ---- YOLO head ------------------------------------------------------------------------------ */
/* Target assignment (CVC-YOLOv3/utils/utils.py:195-275 build_targets, :163-193 bbox_iou).
 * targets [B,T,5] = (cls,cx,cy,w,h) normalised, zero rows = padding; anchors_scaled [A,2] = anchor/stride.
 * Outputs: owner int32 [B,A,Gh,Gw] (winning target index, -1 none), ign u8 [Gh,Gw], rec float [B,T,8]
 * (gi,gj,best,label as int32 bits; tx,ty,tw,th), counts int32 [2] = (N_mask, N_conf_false). */
int b200cv_yolo_targets(const float* targets, const float* anchors_scaled, int B, int T, int A, int Gh, int Gw,
                        float ignore_thres, int32_t* owner, uint8_t* ign, float* rec, int32_t* counts,
                        void* stream);
/* Loss sums and/or gradient of the head logits (models.py:150-155,172-211).  logits fp32 addressed by
 * element strides (channel = a*(5+C)+attr).  sums double[6] (x,y,w,h,obj,noobj un-normalised, atomically
 * accumulated; NULL = skip).  dlogits (NULL = skip): bf16 or fp32; channel-contiguous layouts get all
 * d_channels channels written (zeros for class/pad channels), strided fp32 layouts must be zero-filled by
 * the caller.  gscale: device scalar multiplied into the gradient (upstream grad), NULL = 1. */
int b200cv_yolo_loss(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc, int B, int A,
                     int C, int Gh, int Gw, const int32_t* owner, const uint8_t* ign, const float* rec, int T,
                     const int32_t* counts, float xy_loss, float wh_loss, float obj_loss, float noobj_loss,
                     double* sums, void* dlogits, int dl_dtype, int64_t d_sb, int64_t d_sy, int64_t d_sx,
                     int64_t d_sc, int d_channels, const float* gscale, void* stream);
/* Engine form of the same loss/gradient in two kernels: b200cv_yolo_loss_cells runs one thread per anchor
 * cell and writes compact fp32 cell gradients dcell[pixel][cell_ld] (index a*5+attr, every cell written);
 * b200cv_yolo_expand_dlogits streams them into the dense NHWC head gradient (exact zeros for class/pad
 * channels).  sums / gscale as in b200cv_yolo_loss. */
int b200cv_yolo_loss_cells(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc, int B, int A,
                           int C, int Gh, int Gw, const int32_t* owner, const uint8_t* ign, const float* rec, int T,
                           const int32_t* counts, float xy_loss, float wh_loss, float obj_loss, float noobj_loss,
                           double* sums, float* dcell, int cell_ld, const float* gscale, void* stream);
int b200cv_yolo_expand_dlogits(const float* dcell, int cell_ld, void* dlogits, int dl_dtype, int64_t d_ld,
                               int d_channels, int64_t npix, int A, int C, void* stream);
/* out7[0] += total; out7[1..6] += (x,y,w,h,obj,noobj) -- the tuple order of models.py:211,338. */
int b200cv_yolo_loss_finalize(const double* sums, const int32_t* counts, float xy_loss, float wh_loss,
                              float obj_loss, float noobj_loss, float* out7, void* stream);
/* eval decode (models.py:150-169,213-220) into out[b, row_offset + (a*Gh+gy)*Gw+gx, :5+C]. */
int b200cv_yolo_decode(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc, int B, int A,
                       int C, int Gh, int Gw, const float* anchors_scaled, float stride, float* out,
                       int64_t out_batch_stride, int64_t row_offset, void* stream);
/* the eight dense tensors build_targets() returns (u8,u8,f32 x5,u8). */
int b200cv_yolo_targets_dense(const int32_t* owner, const uint8_t* ign, const float* rec, int B, int T, int A,
                              int C, int Gh, int Gw, uint8_t* mask, uint8_t* conf_mask, float* tx, float* ty,
                              float* tw, float* th, float* tconf, uint8_t* tcls, void* stream);

/* ---- RektNet head ----------------------------------------------------------------------------- */
/* flat_softmax + soft_argmax (RektNet/keypoint_net.py:46-56): logits/hm [rows=B*K][H*W] fp32, pts [rows][2];
 * vx[W], vy[H] are the linspace coordinate tables. */
int b200cv_kpt_softmax_argmax(const float* logits, const float* vx, const float* vy, float* hm, float* pts,
                              int rows, int H, int W, void* stream);
/* CrossRatioLoss.forward (RektNet/cross_ratio_loss.py:20-63).  loss_type 0 l2_softargmax, 1 l2_heatmap,
 * 2 l1_softargmax.  loss3 = (location, geo, total); ubar float[18] = batch-mean unit vectors (saved for
 * backward); ws = one double of scratch. */
int b200cv_kpt_loss(const float* hm, const float* pts, const float* thm, const float* tpts, int B, int K, int HW,
                    int loss_type, int include_geo, float gamma_h, float gamma_v, double* ws, float* loss3,
                    float* ubar, void* stream);
/* Fused backward: loss gradients (g_loc, g_geo device scalars) + optional upstream d_hm/d_pts ->
 * soft-argmax -> softmax Jacobian -> dlogits as NHWC bf16 [B*H*W][16] (ld == 16), or split rows
 * [p0(16) | p1(16) | p2(16)] for ld == 48 (fp32-parity mode). */
int b200cv_kpt_head_bwd(const float* hm, const float* thm, const float* pts, const float* tpts, const float* ubar,
                        const float* vx, const float* vy, const float* g_loc, const float* g_geo,
                        const float* d_hm_up, const float* d_pts_up, int B, int K, int H, int W, int loss_type,
                        int include_geo, float gamma_h, float gamma_v, void* dlogits, int ld, void* stream);
/* Un-fused CrossRatioLoss backward for tensors that did not come from KeypointNet: d_pts [B,K,2] and,
 * if d_hm != NULL, d_hm [B,K,H,W] (= 2 g (hm - thm) / B for l2_heatmap, else 0). */
int b200cv_kpt_loss_bwd(const float* hm, const float* thm, const float* pts, const float* tpts, const float* ubar,
                        const float* g_loc, const float* g_geo, int B, int K, int H, int W, int loss_type,
                        int include_geo, float gamma_h, float gamma_v, float* d_pts, float* d_hm, void* stream);

/* ---- detection post-processing and the detect -> RektNet joint ----------------------- */
/* Confidence filter + greedy top-k NMS, one image per CTA, no host round trip.
 * Replaces CVC-YOLOv3/detect.py:84-90 (rows with conf > conf_thres, (cx,cy,w,h) -> corners) and
 * CVC-YOLOv3/utils/nms.py:4-61 (visit in descending score, drop boxes with IoU > nms_thres, at most top_k
 * candidates).  det = Darknet eval output [B][rows][row_len] fp32 (row = cx,cy,w,h,conf,cls... for box_format 0;
 * x1,y1,x2,y2,score,... for box_format 1 = the arguments of nms() itself), images det_batch_stride elements apart;
 * conf_thres = -inf keeps every row as a candidate.  Equal scores are visited later-row-first (= the reversed STABLE ascending sort;
 * the reference's unstable sort leaves that order open).  Outputs, in visiting order and zero / -1 filled past
 * counts[b]: boxes [B][top_k][4] (x1,y1,x2,y2), scores [B][top_k], det_rows int32 [B][top_k] (row of `det`),
 * counts int32 [B].  top_k <= 512. */
int b200cv_detect_nms(const float* det, int64_t det_batch_stride, int B, int rows, int row_len, int box_format,
                      float conf_thres, float nms_thres, int top_k, float* boxes, float* scores, int32_t* det_rows,
                      int32_t* counts, void* stream);
/* counts int32 [B] -> offsets int32 [B+1] (exclusive scan, offsets[B] = number of crops) and src int32 [n][2] =
 * (image, slot) of crop n; src must hold B*top_k rows.  B <= 1024. */
int b200cv_detect_compact(const int32_t* counts, int B, int top_k, int32_t* offsets, int32_t* src, void* stream);
/* Crop + `cv2.resize(crop, (out_w,out_h))` (8-bit INTER_LINEAR, bit-exact with OpenCV's fixed-point kernel) +
 * HWC->CHW + /255.0 -> fp32: RektNet/utils.py:73-76 (prep_image), RektNet/detect.py:32-34.
 * frames u8 [B][H][W][3]; crop n takes boxes[src[n][0]][src[n][1]] (network-input corners), maps it to the frame by
 * x / ratio - pad (CVC-YOLOv3/detect.py:93-96) with geom = (ratio, pad_w, pad_h) per image (geom_stride floats
 * apart, 0 = shared), floors / ceils it to a non-empty rectangle inside the frame.
 * out fp32 [n_crops][3][out_h][out_w]; rects int32 [n_crops][4] = x0,y0,x1,y1 (exclusive). */
int b200cv_crop_resize_u8(const uint8_t* frames, int B, int H, int W, const float* boxes, int top_k,
                          const int32_t* src, int n_crops, const float* geom, int geom_stride, int out_w, int out_h,
                          float* out, int32_t* rects, void* stream);

/* Letterbox front end of CVC-YOLOv3/detect.py:62-72 (and validate.py / datasets.py's scale step): pad the frame to the
 * network's aspect ratio with `fill` (torchvision pad, fill=127), resize with PIL BILINEAR (Pillow's 8-bit two-pass
 * resampler, bit-exact), to_tensor (/255).  frames u8 [B][H][W][3]; the padded image (H+2*pad_h) x (W+2*pad_w) is
 * resampled to out_h x out_w with host-made tables (Resample.c precompute_coeffs + normalize_coeffs_8bpc, see
 * b200cv/preprocess.py): per output column hx_min/hx_cnt int32 [out_w] and hx_k int32 [out_w][ksize_h] (22-bit fixed
 * point), per output row vy_*; NULL tables = that pass is the identity.  out fp32 [B][3][out_h][out_w];
 * reverse_channels != 0 writes input channel c to plane 2-c (BGR frames -> RGB planes). */
int b200cv_letterbox_u8(const uint8_t* frames, int B, int H, int W, int pad_w, int pad_h, int fill,
                        int reverse_channels, const int32_t* hx_min, const int32_t* hx_cnt, const int32_t* hx_k,
                        int ksize_h, const int32_t* vy_min, const int32_t* vy_cnt, const int32_t* vy_k, int ksize_v,
                        int out_w, int out_h, float* out, void* stream);
/* Tile-and-scale input pipeline (CVC-YOLOv3/utils/datasets.py:143-159 + utils/utils.py:321-426):
 * frames u8 [B][H][W][3] -> scale_image (PIL LANCZOS resize to new_w x new_h) -> pad `fill` up to the patch size ->
 * crop one patch per frame -> to_tensor, out fp32 [B][3][out_h][out_w].  off_x / off_y int32 [B]: the patch origin in
 * the SCALED frame (rounded left/top of get_patch minus the pad; may be negative = inside the pad).  Tables as for
 * b200cv_letterbox_u8 but indexed by scaled-frame column / row, made for the LANCZOS filter (b200cv/tiler.py); NULL =
 * that axis is not rescaled. */
int b200cv_tile_scale_u8(const uint8_t* frames, int B, int H, int W, int new_w, int new_h, int fill,
                         const int32_t* off_x, const int32_t* off_y, const int32_t* hx_min, const int32_t* hx_cnt,
                         const int32_t* hx_k, int ksize_h, const int32_t* vy_min, const int32_t* vy_cnt,
                         const int32_t* vy_k, int ksize_v, int out_w, int out_h, float* out, void* stream);
/* Per-image detection metric on the NMS output: CVC-YOLOv3/validate.py:98-128 (target boxes from the
 * normalised labels, bbox_iou with the +1 convention, greedy matching in score order at iou_thres) and
 * utils/utils.py:58-119 (average_precision, compute_ap).  boxes [B][top_k][4] / counts [B] from b200cv_detect_nms;
 * targets [B][T][5] = (cls,cx,cy,w,h) normalised, rows with a non-positive box number are padding; width/height =
 * network input size.  Outputs per image: ap, recall, precision (0 when skipped), valid int32 (0 = the reference
 * skips the image: no detection or no label), correct u8 [B][top_k] (true-positive flag per kept detection).
 * Single class, like the reference.  T <= 4096. */
int b200cv_detect_match_ap(const float* boxes, const int32_t* counts, int B, int top_k, const float* targets, int T,
                           float width, float height, float iou_thres, float* ap, float* recall, float* precision,
                           int32_t* valid, uint8_t* correct, void* stream);

/* ---- optimizer steps ------------------------------------------------------------------ */
/* One launch per param group.  table: device int64 [n_chunks][5] = {param*, grad*, state1*, state2*, count} (fp32
 * arrays; a chunk is processed by one CTA, keep count around 16K).
 * Adam = torch.optim.Adam(lr, betas, eps, weight_decay) as called at CVC-YOLOv3/train.py:181, RektNet/train_eval.py:263:
 * state1 = exp_avg, state2 = exp_avg_sq, step_size = lr / (1 - beta1^t), bias_correction2_sqrt = sqrt(1 - beta2^t);
 * the betas are doubles because torch forms 1 - beta in double before rounding it to fp32. */
int b200cv_adam_step_multi(const int64_t* table, int n_chunks, float step_size, double beta1, double beta2, float eps,
                           float weight_decay, float bias_correction2_sqrt, void* stream);
/* SGD = torch.optim.SGD(lr, momentum, weight_decay) of CVC-YOLOv3/train.py:185: state1 = momentum_buffer (NULL when
 * momentum == 0), first_step != 0 initialises the buffer with the gradient. */
int b200cv_sgd_step_multi(const int64_t* table, int n_chunks, float lr, float momentum, float weight_decay,
                          int first_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CV_H_ */
