// Internal (non-ABI) declarations shared between the translation units of libb200cv.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200cv.h"

namespace b200cv {

struct StatAcc;  // stat_acc.cuh: 2 x int64 fixed-point accumulator of one per-channel statistic

// ---- error plumbing (api.cpp) -------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> error code
int* device_error_word();            // per-device int the kernels report pipeline timeouts into
int sm_count();

#define B200CV_CHECK_ARG(cond, ...)                                   \
  do {                                                                \
    if (!(cond)) return ::b200cv::set_error(B200CV_ERR_ARG, __VA_ARGS__); \
  } while (0)

// ---- tensor maps (tmap.cpp) ---------------------------------------------------
// bf16 NHWC activation, im2col mode. Returns 0 on success.
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                          int64_t stride_w_elems, int64_t stride_h_elems, int64_t stride_n_elems,
                          int lower_w, int lower_h, int upper_w, int upper_h, int trav_w, int trav_h,
                          int channels_per_pixel, int pixels_per_column);
// bf16 row-major 2-D matrix [rows][cols] with row pitch `ld` elements; box = [box_rows][box_cols].
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows, int box_cols);

// bf16 3-D tensor [d2][d1][d0] (d0 contiguous) with element strides stride1 / stride2; box = [1][box1][box0].
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                      int64_t stride2, int box0, int box1);

// fp32 row-major 2-D matrix (the packed weight gradient): box = [box_rows][box_cols], swizzle = box row bytes.
int make_tmap_2d_f32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                     int box_cols);

// ---- implicit-GEMM convolution core (conv_igemm.cu) ----------------------------
constexpr int kMaxTaps = 64;
// fp32-parity (split) mode: an fp32 value is stored as kSplitPieces bf16 numbers (hi, mid, lo: 3 x 8 = 24 mantissa
// bits, i.e. fp32's own precision), piece j of a row lying j * lo elements after the hi piece.
constexpr int kSplitPieces = 3;

struct IgemmParams {
  // GEMM view: D[m, n] = sum_{tap, c} A_tap[m, c] * B[n, tap_k[tap] + c]
  int M_total;   // output pixels (rows of D)
  int OHW, OW;   // pixels per image / per row of the traversal grid
  int lower_w, lower_h, trav_w, trav_h;
  int num_taps, cblocks;  // k-iterations = num_taps * cblocks
  int Cout;               // valid columns of D
  int num_m_tiles, num_n_tiles;
  int tile_m;  // 128, or 256 = two sub-tiles sharing the weight tiles (see conv_igemm.cu, kM2)
  // epilogue: v = acc*scale[n] + shift[n] (+ residual) -> act -> store; stats on the stored value
  void* out;
  int out_fp32;  // 0: bf16, 1: fp32
  int vec_ok;    // 16-byte vector stores allowed (o_sc == 1 and everything 16B aligned)
  long long o_sn, o_sh, o_sw, o_sc;
  const __nv_bfloat16* res;
  int res_after_act;
  int res_vec_ok;  // residual rows are channel-contiguous and 16B aligned
  long long r_sn, r_sh, r_sw, r_sc;
  const float* scale;
  const float* shift;
  int act;  // 0 none, 1 leaky(slope), 2 relu
  float slope;
  StatAcc* stats;  // [stats_parts][2*Cout]: sum, sum of squares (added, see stat_acc.cuh), or null
  int stats_parts;
  int* err;
  // fused BN-backward reduction (dgrad, staged epilogue only): sums of dz and dz*(y-mean)*rstd, see b200cv.h
  StatAcc* bn_sums;
  int bn_parts;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_rstd;
  float bn_neg;  // act'(z) for z <= 0
  int stages;                 // pipeline stages of this launch (set by launch_igemm)
  int y_slots, y_slots_log2;  // y tiles per epilogue warp of the fused BN-backward reduction (power of two)
  int stg_slots;              // staging tiles per epilogue warp of the wide staged epilogue (1 or 2)
  // depth-to-space output of the one-launch stride-2 data gradient (b200cv_conv_dgrad_d2s): GEMM row m = pixel
  // (n, q, p) of the dy grid, column (a, b, c) -> dx[n, 2q+a, 2p+b, c]; d2s_c2 = 2*C columns per output row (0 = off),
  // tmO is then the 3-D map {2C, OW, 2*N*OH} with a box of d2s_g rows
  int d2s_c2;
  // statistics of the depth-to-space form are per CHANNEL (column & stat_mask); rows of the statistics matrices hold
  // 2*stat_cols entries.  run_igemm sets stat_cols = Cout, stat_mask = -1 otherwise.
  int stat_cols, stat_mask;
  int d2s_g;  // rows per depth-to-space store: the largest power of two <= 32 that divides OW
  int res_iters;  // residual added by the tensor core: extra k-iterations D += I[:, k-slice] * R[k-slice rows, :] (0 = off)
  int dbg;  // B200CV_DBG bits (bring-up timing experiments only): 1 no stores, 2 no stats, 4 no TMEM read
  // fp32-parity (split) mode: lo halves of the output / residual lie this many elements after the hi halves (0 = off)
  long long out_lo, res_lo;
  short tap_w[kMaxTaps];
  short tap_h[kMaxTaps];
  short tap_c[kMaxTaps];  // channel offset of the A operand (0, or the lo half of a split activation)
  int tap_k[kMaxTaps];
};

// A: im2col tensor map over the activation; B: 2-D map over the packed weights.
// kc = channels per k-block (16/32/64); block_n in {16,32,64,128,256}.
// tmO: 2-D map over a row-major bf16 output [M][Cout] with box {epilogue chunk, 32 rows} (staged TMA-store
// epilogue), or null for the generic (direct-store) epilogue.
// tmI / tmR (both or neither): 128x128 bf16 identity and the row-major residual [M][ld] for p.res_iters > 0.
// tmY: map over the BN input y laid out like the output (same box as tmO) for p.bn_sums != null.
int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO, const CUtensorMap* tmI,
                 const CUtensorMap* tmR, const CUtensorMap* tmY, const IgemmParams& p, int kc, int block_n,
                 cudaStream_t stream);
const void* device_identity128();    // bf16 [128][128] identity matrix (library-owned, per device)

// Launch with (optionally) programmatic dependent launch: the kernel must call ptx::pdl_wait() before its first
// global-memory access (see ptx.cuh).  Off unless B200CV_PDL=1 (measured: no gain inside a CUDA graph).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
// Channel padding rule for NHWC bf16 activations: 16, 32, or a multiple of 64.
inline int pad_channels(int c) { return c <= 16 ? 16 : (c <= 32 ? 32 : round_up(c, 64)); }
inline int kc_for(int cpad) { return cpad < 64 ? cpad : 64; }

}  // namespace b200cv
