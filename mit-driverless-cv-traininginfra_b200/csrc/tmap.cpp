// TMA tensor-map construction.  The driver entry points are resolved at run time through the
// CUDA runtime so that the library links (and `ctypes.CDLL` loads) on machines without libcuda.
#include <cstring>
#include <mutex>

#include "internal.h"

namespace b200cv {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeTiledFn g_tiled = nullptr;
EncodeIm2colFn g_im2col = nullptr;
int g_driver_version = 0;
std::once_flag g_once;

void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&g_driver_version);
}

CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  if (inner_bytes >= 128) return CU_TENSOR_MAP_SWIZZLE_128B;
  if (inner_bytes >= 64) return CU_TENSOR_MAP_SWIZZLE_64B;
  return CU_TENSOR_MAP_SWIZZLE_32B;
}

}  // namespace

int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                          int64_t stride_w_elems, int64_t stride_h_elems, int64_t stride_n_elems,
                          int lower_w, int lower_h, int upper_w, int upper_h, int trav_w, int trav_h,
                          int channels_per_pixel, int pixels_per_column) {
  std::call_once(g_once, resolve);
  if (!g_im2col) return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeIm2col not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)stride_w_elems * 2, (cuuint64_t)stride_h_elems * 2,
                           (cuuint64_t)stride_n_elems * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)trav_w, (cuuint32_t)trav_h, 1};
  CUresult r = g_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                        strides, lower, upper, (cuuint32_t)channels_per_pixel,
                        (cuuint32_t)pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_for_bytes(channels_per_pixel * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200CV_ERR_DRIVER,
                     "cuTensorMapEncodeIm2col failed (%d): N%d H%d W%d C%d lo(%d,%d) up(%d,%d) "
                     "trav(%d,%d) cpp%d ppc%d",
                     (int)r, N, H, W, C, lower_w, lower_h, upper_w, upper_h, trav_w, trav_h,
                     channels_per_pixel, pixels_per_column);
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (the same
  // workaround CUTLASS applies): clear bit 21 of the second descriptor word.
  if (g_driver_version <= 13010) {
    const int64_t bytes = (int64_t)N * stride_n_elems * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  }
  return 0;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows, int box_cols) {
  std::call_once(g_once, resolve);
  if (!g_tiled) return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle_for_bytes(box_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200CV_ERR_DRIVER,
                     "cuTensorMapEncodeTiled failed (%d): rows%lld cols%lld ld%lld box(%d,%d)", (int)r,
                     (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                      int64_t stride2, int box0, int box1) {
  std::call_once(g_once, resolve);
  if (!g_tiled) return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * 2, (cuuint64_t)stride2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box0 * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeTiled(3d) failed (%d): dims %lld %lld %lld box(%d,%d)", (int)r,
                     (long long)d0, (long long)d1, (long long)d2, box0, box1);
  return 0;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                     int box_cols) {
  std::call_once(g_once, resolve);
  if (!g_tiled) return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_cols * 4),
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200CV_ERR_DRIVER, "cuTensorMapEncodeTiled(f32) failed (%d): rows%lld cols%lld ld%lld box(%d,%d)",
                     (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
  return 0;
}

}  // namespace b200cv
