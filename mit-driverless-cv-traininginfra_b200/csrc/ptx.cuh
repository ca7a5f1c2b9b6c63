// Thin inline-PTX wrappers for the sm_100a features the conv engine uses:
// mbarrier, TMA (tiled + im2col loads, tiled stores), tcgen05 (alloc / mma / commit / ld).
// Everything here is device-only and header-only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200cv {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// A kernel launched with the programmatic-stream-serialization attribute may start (barrier init, TMEM allocation,
// descriptor prefetch) while its predecessor drains; pdl_wait() blocks until the predecessor has completed and its
// memory operations are visible, and must precede EVERY global-memory access.  No-ops for ordinary launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline must not hang the GPU box, and it must not go unnoticed either.  On timeout the
// error word is set (b200cv_check_device_error names the stage) and the kernel TRAPS: the launch -- and every later
// CUDA call of the process -- fails loudly instead of feeding garbage activations to the optimizer.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// im2col-mode load of `pixelsPerColumn` consecutive output pixels starting at the
// base pixel (w,h,n); (off_w, off_h) is the filter-tap offset added to every pixel.
__device__ __forceinline__ void tma_load_im2col_4d(void* smem, const CUtensorMap* m, uint64_t* bar,
                                                   int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
      "h"(off_w), "h"(off_h)
      : "memory");
}
// Predicated forms for a CONVERGED warp (`leader` = elect_one() taken once): the operands stay warp-uniform, so
// ptxas keeps them in uniform registers instead of broadcasting them from the one active lane before every
// instruction (ELECT + 4 x R2UR.BROADCAST per UTMALDG: ~150 cycles per TMA issued from a single-lane branch).
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t leader, uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %2, 0;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t leader, void* smem, const CUtensorMap* m, uint64_t* bar,
                                                  int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n\t}\n" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_elect(uint32_t leader, void* smem, const CUtensorMap* m,
                                                         uint64_t* bar, int c, int w, int h, int n, uint16_t off_w,
                                                         uint16_t off_h) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %9, 0;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}\n" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
      "h"(off_w), "h"(off_h), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(uint32_t leader, void* smem, const CUtensorMap* m, uint64_t* bar,
                                                  int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %6, 0;\n\t"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(leader)
      : "memory");
}
// stores / reductions of the lane with leader != 0 (bulk groups are per thread: commit / wait in the same lane)
__device__ __forceinline__ void tma_store_2d_elect(uint32_t leader, const CUtensorMap* m, const void* smem, int c0,
                                                   int c1) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %4, 0;\n\t"
      "@pe cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n\t"
      "@pe cp.async.bulk.commit_group;\n\t}\n" ::"l"(reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d_elect(uint32_t leader, const CUtensorMap* m, const void* smem, int c0,
                                                   int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
      "@pe cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n\t}\n" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_elect(uint32_t leader, const CUtensorMap* m, const void* smem,
                                                        int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %4, 0;\n\t"
      "@pe cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];\n\t"
      "@pe cp.async.bulk.commit_group;\n\t}\n" ::"l"(reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tma_commit_elect(uint32_t leader) {
  asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %0, 0;\n\t@pe cp.async.bulk.commit_group;\n\t}\n" ::"r"(leader)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
// 3-D tiled store: elements whose coordinates fall outside the tensor (negative or past the end) are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// fp32 (or any map-typed) reduce-add of a shared-memory tile into global memory, done by the TMA unit / L2
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same, executed by a CONVERGED warp: one elected lane issues.  Issuing from inside a single-lane branch makes
// ptxas move every operand from vector to uniform registers with ELECT + R2UR.BROADCAST in front of each UTCHMMA
// (~200 cycles per MMA measured: the 3x3 layers were bound by the issue rate of this one thread, 950 cycles per
// k-iteration for 512 cycles of tensor work); with warp-uniform operands they stay in uniform registers.
__device__ __forceinline__ void umma_bf16_elect(uint32_t leader, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "setp.ne.b32 pe, %5, 0;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// (`leader` = elect_one() taken ONCE by the warp: tcgen05.commit tracks the MMAs of the thread that executes it)
__device__ __forceinline__ void umma_commit_elect(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "setp.ne.b32 pe, %1, 0;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (tcgen05): start>>4 [0,14) | LBO>>4 [16,30) |
// SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64).
// Layout types: 0 none, 2 = 128B swizzle, 4 = 64B swizzle, 6 = 32B swizzle.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace b200cv
