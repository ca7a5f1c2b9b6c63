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

/* ---- library ------------------------------------------------------------------------- */
const char* b200cv_version(void);
const char* b200cv_last_error(void);
/* Reads (and clears) the per-device error word the tcgen05/TMA pipelines write on a bounded-wait
 * timeout.  Synchronises `stream`.  Test/diagnostic use only -- never on the hot path. */
int b200cv_check_device_error(void* stream);
/* 16 for c<=16, 32 for c<=32, otherwise the next multiple of 64. */
int b200cv_pad_channels(int c);

/* ---- layout ---------------------------------------------------------------------------- */
/* NCHW fp32 [N,C,H,W] -> NHWC bf16 [N,H,W,Cpad] (extra channels zero).
 * Replaces the implicit layout of `imgs.to(device)` CVC-YOLOv3/train.py:60, RektNet/train_eval.py:60. */
int b200cv_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int Cpad,
                                 void* stream);
/* NHWC bf16 [N,H,W,Cpad] (first C channels) -> NCHW fp32 [N,C,H,W]. */
int b200cv_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int Cpad,
                                 void* stream);
/* OIHW fp32 conv weight -> packed bf16.
 *   transpose == 0: [O][R*S][Ipad]      (forward operand;  K index = tap*Ipad + i)
 *   transpose == 1: [I][R*S][Opad]      (data-gradient operand; K index = tap*Opad + o)
 * nn.Conv2d weights: CVC-YOLOv3/models.py:59-65, RektNet/keypoint_net.py:17,25, resnet.py:12-18. */
int b200cv_pack_weights(const float* w_oihw, void* dst, int O, int I, int R, int S, int Ipad,
                        int Opad, int transpose, void* stream);
/* packed fp32 gradient [O][R*S][Ipad] -> OIHW fp32 (the .grad layout torch.optim expects). */
int b200cv_unpack_wgrad(const float* dw_packed, float* dw_oihw, int O, int I, int R, int S,
                        int Ipad, void* stream);

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
  /* fused epilogue:  y = act(acc*scale[c] + shift[c] + residual)  */
  const float* scale;   /* [Cout] or NULL */
  const float* shift;   /* [Cout] or NULL (conv bias / folded BN) */
  const void* residual; /* bf16 or NULL, addressed like y with its own strides */
  int64_t r_sn, r_sh, r_sw, r_sc;
  int32_t act;
  float slope;
  /* per-channel [sum(y) | sum(y*y)] over N*OH*OW, atomically ADDED into stats[2*Cout]; NULL = off */
  float* stats;
} b200cv_conv_args;

/* y = conv2d(x, w).  nn.Conv2d forward: CVC-YOLOv3/models.py:59-65,320-321;
 * RektNet/keypoint_net.py:59,64, resnet.py:22-26. */
int b200cv_conv_fwd(const b200cv_conv_args* a, void* stream);

/* Data gradient.  Here x = dY (NHWC bf16 [N,OH,OW,Cin=pad(Cout_fwd)]), w = transposed pack
 * [Cin_fwd][R*S][pad(Cout_fwd)], Cout = Cin_fwd, (H,W) = spatial size of dY, and (out_h,out_w) the
 * spatial size of dX; stride/pad/dil are the FORWARD conv's.  residual adds an existing gradient
 * (shortcut / route fan-out).  autograd of nn.Conv2d: CVC-YOLOv3/train.py:70. */
int b200cv_conv_dgrad(const b200cv_conv_args* a, int out_h, int out_w, void* stream);

/* Weight gradient: dw_packed[o][tap][i] += sum_pixels dY[pix][o] * X[pix+tap][i]  (fp32, atomically
 * accumulated; zero it first).  x: NHWC bf16 [N,H,W,Cin]; dy: NHWC bf16 [N,OH,OW,dy_ld].
 * autograd of nn.Conv2d: CVC-YOLOv3/train.py:70, RektNet/train_eval.py:71. */
int b200cv_conv_wgrad(const void* x, const void* dy, float* dw_packed, int N, int H, int W, int Cin,
                      int Cout, int dy_ld, int R, int S, int stride, int pad, int dil, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CV_H_ */
