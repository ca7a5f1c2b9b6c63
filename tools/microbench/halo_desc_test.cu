// Feasibility test (measurement aid, not product code) for the halo-tile implicit GEMM planned in DESIGN.md §8:
// can the nine taps of a 3x3 convolution be nine tcgen05 A-operand descriptors into ONE TMA-tiled, swizzled halo patch?
//
//   tile   = 16 image rows x 8 columns = 128 output pixels (one 8-row UMMA core-matrix group = 8 pixels of one row)
//   patch  = 18 rows x 16 columns x C channels, one tiled 4-D TMA load (start at (r0-1, c0-1), zero fill outside)
//   tap    = descriptor start patch + (dr*16 + dc) * C*2 bytes, SBO = 16 * C*2 bytes (a multiple of the swizzle atom),
//            base-offset field (bits 49..51) = (start >> 7) & 7 because the column shift un-aligns the start
//
// For C in {64, 32, 16} (128B / 64B / 32B swizzle), with and without the base-offset field, at an interior and at an
// edge tile, every tap's 128 x 64 product is compared EXACTLY (small-integer data) with a CPU reference; the table of
// mismatch counts says which encoding the hardware expects.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o halo_desc_test halo_desc_test.cu && ./halo_desc_test
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mit-driverless-cv-traininginfra_b200/csrc/ptx.cuh"

using namespace b200cv;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int BN = 64;   // output channels of the test GEMM
constexpr int PW = 16;   // patch width (8 + 2, padded to 16 pixels)
constexpr int PH = 18;   // patch height (16 + 2)
constexpr int IMG_H = 20, IMG_W = 24;

__device__ __forceinline__ void tma_load_tiled_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c, int w, int h,
                                                  int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(ptx::smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

template <int C>
__global__ void __launch_bounds__(128)
halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int r0, int c0,
            int use_base_offset, float* __restrict__ out, int* __restrict__ err) {
  constexpr int kRowBytes = C * 2;
  constexpr int kLayout = C == 64 ? 2 : (C == 32 ? 4 : 6);  // 128B / 64B / 32B swizzle
  constexpr int kPatchBytes = PH * PW * kRowBytes;
  constexpr int kWTile = BN * kRowBytes;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* patch = smem;                                        // 1024-byte aligned
  uint8_t* wt = smem + ((kPatchBytes + 1023) / 1024) * 1024;    // nine [BN][C] weight tiles, each 1024-aligned
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_load, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<BN>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(&bar_load, kPatchBytes + 9 * kWTile);
    tma_load_tiled_4d(patch, &tmX, &bar_load, 0, c0 - 1, r0 - 1, 0);
    for (int t = 0; t < 9; ++t) ptx::tma_load_2d(wt + t * ((kWTile + 1023) / 1024) * 1024, &tmW, &bar_load, t * C, 0);
  }
  ptx::mbar_wait(&bar_load, 0, err, 1);
  __syncthreads();
  constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, 0, 0);
  uint32_t mma_phase = 0;
  for (int t = 0; t < 9; ++t) {
    if (threadIdx.x == 0) {
      ptx::tc_fence_after();
      const int dr = t / 3, dc = t % 3;
      const uint32_t sa = ptx::smem_u32(patch) + (dr * PW + dc) * kRowBytes;
      const uint32_t sb = ptx::smem_u32(wt + t * ((kWTile + 1023) / 1024) * 1024);
      uint64_t adesc = ptx::make_smem_desc(sa, 16, PW * kRowBytes, kLayout);
      if (use_base_offset) adesc |= static_cast<uint64_t>((sa >> 7) & 7) << 49;
      const uint64_t bdesc = ptx::make_smem_desc(sb, 16, 8 * kRowBytes, kLayout);
#pragma unroll
      for (int k = 0; k < C / 16; ++k) ptx::umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      ptx::umma_commit(&bar_mma);
    }
    ptx::mbar_wait(&bar_mma, mma_phase, err, 2);
    mma_phase ^= 1;
    ptx::tc_fence_after();
    // warp w reads TMEM lanes 32w..32w+31 (= output rows), 64 columns
    const int row = warp * 32 + lane;
    float* o = out + ((size_t)t * 128 + row) * BN;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + half * 32, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[half * 32 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<BN>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMapSwizzle swz(int bytes) {
  return bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

template <int C>
static void run(EncodeTiledFn enc) {
  // small-integer data: every product and sum is exact in bf16 x bf16 -> fp32
  std::vector<__nv_bfloat16> hx((size_t)IMG_H * IMG_W * C), hw((size_t)BN * 9 * C);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1234 + C);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 9 - 4); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 5 - 2); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dout;
  int* derr;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dout, (size_t)9 * 128 * BN * 4));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)IMG_W, (cuuint64_t)IMG_H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)IMG_W * C * 2, (cuuint64_t)IMG_H * IMG_W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, PW, PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz(C * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("C=%d: encode X failed (%d)\n", C, (int)r); return; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * C, (cuuint64_t)BN};
    cuuint64_t strides[1] = {(cuuint64_t)9 * C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz(C * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("C=%d: encode W failed (%d)\n", C, (int)r); return; }
  }
  const int smem_bytes = 1024 + ((PH * PW * C * 2 + 1023) / 1024) * 1024 + 9 * ((BN * C * 2 + 1023) / 1024) * 1024;
  CK(cudaFuncSetAttribute(halo_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int tiles[2][2] = {{2, 8}, {4, 16}};  // interior; bottom/right edge (rows 3..20, columns 15..30 of 20 x 24)
  const int edge0[2] = {0, 0};                // top/left edge
  std::vector<float> got((size_t)9 * 128 * BN);
  for (int bo = 1; bo >= 0; --bo) {
    for (int ti = 0; ti < 3; ++ti) {
      const int r0 = ti < 2 ? tiles[ti][0] : edge0[0], c0 = ti < 2 ? tiles[ti][1] : edge0[1];
      CK(cudaMemset(derr, 0, 4));
      CK(cudaMemset(dout, 0xff, got.size() * 4));
      halo_kernel<C><<<1, 128, smem_bytes>>>(tmX, tmW, r0, c0, bo, dout, derr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C=%d base_offset=%d tile(%d,%d): kernel failed: %s\n", C, bo, r0, c0, cudaGetErrorString(e)); exit(2); }
      int herr = 0;
      CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
      printf("C=%2d base_offset=%d tile(r0=%d,c0=%2d) timeout=%d  mismatches per tap:", C, bo, r0, c0, herr);
      int total = 0;
      for (int t = 0; t < 9; ++t) {
        const int dr = t / 3, dc = t % 3;
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int ih = r0 + m / 8 + dr - 1, iw = c0 + m % 8 + dc - 1;
          for (int n = 0; n < BN; ++n) {
            float ref = 0.f;
            if (ih >= 0 && ih < IMG_H && iw >= 0 && iw < IMG_W)
              for (int c = 0; c < C; ++c) ref += fx[((size_t)ih * IMG_W + iw) * C + c] * fw[((size_t)n * 9 + t) * C + c];
            if (got[((size_t)t * 128 + m) * BN + n] != ref) ++bad;
          }
        }
        printf(" %4d", bad);
        total += bad;
      }
      printf("  -> %s\n", total == 0 ? "EXACT" : "wrong");
    }
  }
  cudaFree(dx); cudaFree(dw); cudaFree(dout); cudaFree(derr);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  run<64>(enc);
  run<32>(enc);
  run<16>(enc);
  return 0;
}
