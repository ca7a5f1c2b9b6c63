// Micro-benchmark (measurement aid, not product code): sustained TMA load throughput of ONE SM as a function of
// the box size and of the number of issuing threads.  Every CTA streams 2-D boxes [R rows][64 bf16] (128-byte
// rows, SWIZZLE_128B) from an L2-resident matrix with D loads in flight per issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_bench tma_bench.cu && ./tma_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int D>
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int iters,
                                                  int producers, int rows_per_cta, int store, void* /*unused*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (su32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < D; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[w][i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < producers && lane == 0) {
    const uint32_t bytes = box_rows * 128;
    uint8_t* base = smem + warp * D * 32768;
    const int row0 = blockIdx.x * rows_per_cta;
    for (int i = 0; i < iters; ++i) {
      const int slot = i % D;
      if (i >= D) {  // wait for the load issued D iterations ago
        const uint32_t parity = ((i / D) - 1) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}"
                       : "=r"(ok) : "r"(su32(&bar[warp][slot])), "r"(parity) : "memory");
      }
      const int r = row0 + ((i * box_rows + warp * 4096) % (rows_per_cta - box_rows));
      if (!store) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar[warp][slot])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(base + slot * 32768)), "l"((uint64_t)&tm), "r"(su32(&bar[warp][slot])), "r"(0), "r"(r) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"((uint64_t)&tm), "r"(su32(base + slot * 32768)), "r"(0), "r"(r) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&bar[warp][slot])) : "memory");
      }
    }
    if (store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else {
      for (int s = 0; s < D && s < iters; ++s) {  // drain
        const int i = iters - 1 - s;
        const int slot = i % D;
        const uint32_t parity = (i / D) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}"
                       : "=r"(ok) : "r"(su32(&bar[warp][slot])), "r"(parity) : "memory");
      }
    }
  }
  __syncthreads();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int ctas = 148, rows_per_cta = 1024;  // 148 * 1024 rows * 128 B = 19 MB: L2 resident
  const int64_t rows = (int64_t)ctas * rows_per_cta;
  void* buf;
  CK(cudaMalloc(&buf, rows * 128));
  CK(cudaMemset(buf, 0, rows * 128));
  constexpr int D = 3;
  auto kern = tma_kernel<D>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("mode  producers box_rows box_KB  ns/op(per SM)  GB/s per SM   chip TB/s\n");
  for (int store = 0; store < 2; ++store)
    for (int producers = 1; producers <= 2; ++producers)
      for (int box_rows : {8, 16, 32, 64, 128, 256}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {64, (cuuint64_t)rows};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int iters = 4000;
        const int smem = 1024 + producers * D * 32768;
        kern<<<ctas, 128, smem>>>(tm, box_rows, 200, producers, rows_per_cta, store, nullptr);
        CK(cudaEventRecord(e0));
        kern<<<ctas, 128, smem>>>(tm, box_rows, iters, producers, rows_per_cta, store, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)iters * producers;
        const double ns_per_op = ms * 1e6 / ops;
        const double gbs = box_rows * 128.0 / ns_per_op;
        printf("%-5s %9d %8d %6.1f %14.1f %12.1f %11.2f\n", store ? "store" : "load", producers, box_rows,
               box_rows * 128 / 1024.0, ns_per_op, gbs, gbs * ctas / 1000.0);
      }
  return 0;
}
