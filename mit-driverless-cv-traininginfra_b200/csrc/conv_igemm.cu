// Implicit-GEMM convolution core for sm_100a.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) + tap][c] * B[n][tap_k[tap] + c]
//
// * A (NHWC bf16 activations) is fetched by TMA in im2col mode: one instruction brings the
//   128 consecutive output pixels of a tile (crossing rows and images, zero-filling the halo)
//   for one filter tap and one block of KC channels into 128B/64B/32B-swizzled shared memory.
// * B (packed bf16 weights, K-major) is fetched by a tiled TMA load.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a TMEM accumulator;
//   two accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// * Persistent: grid = #SMs, static round-robin tile schedule.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..9 = epilogue
//   (two warps per TMEM lane quarter, each taking every other 32-column chunk of the accumulator:
//   with one warp per scheduler every latency of the epilogue was exposed and short-K layers were
//   epilogue-bound).  Per-channel BatchNorm statistics are column sums over the 32 rows of a warp
//   (shared-memory transpose, 16 columns at a time) accumulated in a warp-private shared row for the
//   whole CTA lifetime -- no shared-memory atomics -- and flushed to global once.
//
// The same kernel serves forward convolution, stride-1 data-gradient (flipped taps, transposed
// weights) and the four parity classes of a stride-2 data-gradient; only the tap table, the
// bounding-box corners and the output strides differ (see conv_api.cu).
#include <cuda_bf16.h>

#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"
#include "stat_acc.cuh"

namespace b200cv {

namespace {

constexpr int kBlockM = 128;
constexpr int kEpiWarp0 = 2;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);  // 320
constexpr int kSmemMax = 227 * 1024;                    // dynamic shared memory of one CTA alone on an SM
constexpr int kSmemMax2 = 113 * 1024;                   // ... of each of two co-resident CTAs
constexpr int kScratchLd = 17;                          // [32 rows][16 columns + 1] fp32 transpose tile
constexpr int kMaxStages = 8;
constexpr int kMaxYSlots = 4;

// kTma: the output is a plain row-major bf16 matrix [M][ld] -> the epilogue stages 32x32 (or 32x16) bf16 tiles in
// swizzled shared memory and writes them with TMA bulk stores; BN statistics are read back from the staged
// tile.  Otherwise ("generic": fp32 / strided / ragged outputs) rows are stored directly from registers.
// kMode 2 = kTma plus the fused BN-backward reduction (a 2-deep ring of TMA-loaded y tiles per epilogue warp).
// kM2: the tile is 256 output pixels = two 128-row sub-tiles that share every B (weight) tile -- for GEMMs whose N is
// only 128 wide this cuts the L2->SM bytes per FLOP by a quarter (A 32 KB + B 16 KB per 512 tensor cycles, the
// ratio of a 128x256 tile).  Epilogue warps 2..5 take sub-tile 0, warps 6..9 sub-tile 1.
template <int KC, int BN, int kMode, bool kM2 = false>
struct Cfg {
  static_assert(!kM2 || (BN == 128 && kMode >= 1), "the 256-row tile exists for BN == 128 with the staged epilogue");
  static constexpr int kSub = kM2 ? 2 : 1;           // 128-row sub-tiles per tile
  static constexpr int kTileM = 128 * kSub;
  static constexpr int kAccCols = BN * kSub;         // TMEM columns of one accumulator stage
  static constexpr bool kTma = kMode >= 1;
  static constexpr bool kBnRed = kMode == 2;
  static constexpr bool kWide = BN >= 128;            // staged epilogue works on 64-column (128-byte) blocks
  static constexpr int kStgTile = kWide ? 4096 : 2048;  // bytes of one staging / y tile
  static constexpr int kASubBytes = kBlockM * KC * 2;
  static constexpr int kABytes = kSub * kASubBytes;
  static constexpr int kBBytes = BN * KC * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // Small stages (<= 24 KB: the 16/32/64-channel layers) are bound by the per-k-iteration overhead of the single
  // producer / MMA threads (~150 cycles per TMA instruction issued), not by bandwidth: run two CTAs per SM (each
  // with a shallower ring) so that two producers and two MMA issuers work per SM.
  static constexpr int kTmemCols = (2 * kAccCols) < 32 ? 32 : (2 * kAccCols);
  static constexpr int kCtasPerSm = (kStageBytes <= 24 * 1024 && kTmemCols <= 256) ? 2 : 1;
  static constexpr int kRowBytes = KC * 2;               // 32 / 64 / 128
  static constexpr int kLayout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  static constexpr int kSBO = 8 * kRowBytes;             // 8-row swizzle atom pitch
  static constexpr int kChunk = BN >= 32 ? 32 : 16;      // epilogue column chunk
  static constexpr int kNumChunks = BN / kChunk;         // 1..8
  static constexpr int kChunksPerWarp = (kNumChunks + 1) / 2;
  static constexpr int kAccPerWarp = kChunksPerWarp * kChunk;  // columns a warp keeps statistics for
  // barriers (8 B each): full/empty (8 stages max), tfull/tempty, kMaxYSlots per epilogue warp for the y ring; + tmem ptr
  static constexpr int kBarBytes = (2 * kMaxStages + 4 + kMaxYSlots * kEpiWarps) * 8 + 16;
  // generic: per-warp [sum | sumsq] rows + fp32 transpose scratch; TMA: one 1 KB-aligned staging tile per warp
  static constexpr int kStatBytes = kTma ? 0 : kEpiWarps * 2 * kAccPerWarp * 4;
  static constexpr int kScratchBytes = kTma ? kEpiWarps * kStgTile : kEpiWarps * 32 * kScratchLd * 4;
  // shared memory of a launch with `y_slots` y tiles per epilogue warp (0 unless kBnRed); the pipeline gets
  // whatever is left (stages are a RUN-TIME parameter: short-K layers trade stages for a deeper y ring)
  // stg_slots: staging tiles per epilogue warp (wide staged epilogue: 1, or 2 for short-K tiles, see launch_one)
  static constexpr int extra_bytes(int y_slots, int stg_slots = 1) {
    return 1024 /*align slack*/ + 2048 /*barrier block + align*/ + kStatBytes + kScratchBytes * stg_slots +
           kEpiWarps * y_slots * kStgTile;
  }
  static constexpr int stages_for(int y_slots, int stg_slots = 1) {
    const int n = ((kCtasPerSm == 2 ? kSmemMax2 : kSmemMax) - extra_bytes(y_slots, stg_slots)) / kStageBytes;
    return n > kMaxStages ? kMaxStages : n;
  }
  static_assert(kBarBytes <= 1024, "barrier block");
  static_assert(kStageBytes % 256 == 0, "stage alignment");
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == 1) return v > 0.f ? v : v * slope;
  if (act == 2) return v > 0.f ? v : 0.f;
  return v;
}

// One step of the keep-half butterfly: lanes whose BIT is set keep s[N..2N) and give s[0..N) away (and vice
// versa); after the steps N = 4, 2, 1 every lane of the 8-lane group holds the group total of ONE column.
template <int N, int BIT>
__device__ __forceinline__ void keep_half_step(float (&s)[8], float (&t)[8], int lane) {
  const bool up = (lane & BIT) != 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float send_s = up ? s[k] : s[k + N], keep_s = up ? s[k + N] : s[k];
    const float send_t = up ? t[k] : t[k + N], keep_t = up ? t[k + N] : t[k];
    s[k] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, BIT);
    t[k] = keep_t + __shfl_xor_sync(0xffffffffu, send_t, BIT);
  }
}

template <int KC, int BN, int kMode, bool kM2>
__global__ void __launch_bounds__(kThreads, Cfg<KC, BN, kMode, kM2>::kCtasPerSm)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmI,
             const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY,
             const __grid_constant__ IgemmParams p) {
  using C = Cfg<KC, BN, kMode, kM2>;
  constexpr bool kTma = C::kTma;
  constexpr bool kBnRed = C::kBnRed;
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B (128B-swizzle atom) by pointer arithmetic so the shared state space stays provable
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  const int nstages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + nstages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* y_bar = tempty_bar + 2;  // [kEpiWarps][kMaxYSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_bar + kMaxYSlots * kEpiWarps);
  // stage ring (multiple of 1 KB) | 1 KB barrier block | staging tiles (TMA) or stats rows + scratch (generic)
  uint8_t* s_extra = reinterpret_cast<uint8_t*>(full_bar) + C::kBarBytes;
  s_extra += (1024u - (ptx::smem_u32(s_extra) & 1023u)) & 1023u;  // staging tiles: swizzle-atom aligned
  float* s_stats = reinterpret_cast<float*>(s_extra);
  float* s_scratch = s_stats + kEpiWarps * 2 * C::kAccPerWarp;

  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kiters = p.num_taps * p.cblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (kTma) ptx::prefetch_tmap(&tmO);
    if (p.res_iters) {
      ptx::prefetch_tmap(&tmI);
      ptx::prefetch_tmap(&tmR);
    }
    for (int i = 0; i < nstages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], kEpiWarps);  // one arrive per epilogue warp
    }
    if (kBnRed) {
      ptx::prefetch_tmap(&tmY);
      for (int i = 0; i < kMaxYSlots * kEpiWarps; ++i) ptx::mbar_init(&y_bar[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  if (!kTma && p.stats)
    for (int i = threadIdx.x; i < kEpiWarps * 2 * C::kAccPerWarp; i += kThreads) s_stats[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();  // everything above touched only shared / tensor memory

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    {  // converged warp; the elected lane issues (ptx::*_elect keep the operands in uniform registers)
      const uint32_t leader = ptx::elect_one() ? 1u : 0u;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles;
        const int nt = tile - mt * p.num_n_tiles;
        const int m0 = mt * C::kTileM;
        const int n_img = m0 / p.OHW;
        const int rem = m0 - n_img * p.OHW;
        const int pr = rem / p.OW;
        const int qc = rem - pr * p.OW;
        const int cw = p.lower_w + qc * p.trav_w;
        const int ch = p.lower_h + pr * p.trav_h;
        // second 128-row sub-tile (kM2): its own base pixel; past the last image the loads are zero-filled
        const int m1 = m0 + kBlockM;
        const int n_img1 = m1 / p.OHW;
        const int rem1 = m1 - n_img1 * p.OHW;
        const int pr1 = rem1 / p.OW;
        const int cw1 = p.lower_w + (rem1 - pr1 * p.OW) * p.trav_w;
        const int ch1 = p.lower_h + pr1 * p.trav_h;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const uint16_t ow = static_cast<uint16_t>(p.tap_w[tap]);
          const uint16_t oh = static_cast<uint16_t>(p.tap_h[tap]);
          const int kb = p.tap_k[tap];
          const int ca = p.tap_c[tap];  // 0, or the lo half of a split activation (fp32-parity mode)
          for (int cb = 0; cb < p.cblocks; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 1);
            uint8_t* sa = stage_base + stage * C::kStageBytes;
            uint8_t* sb = sa + C::kABytes;
            ptx::mbar_expect_tx_elect(leader, &full_bar[stage], C::kStageBytes);
            ptx::tma_load_im2col_4d_elect(leader, sa, &tmA, &full_bar[stage], ca + cb * KC, cw, ch, n_img, ow, oh);
            if constexpr (kM2)
              ptx::tma_load_im2col_4d_elect(leader, sa + C::kASubBytes, &tmA, &full_bar[stage], ca + cb * KC, cw1, ch1, n_img1, ow,
                                      oh);
            ptx::tma_load_2d_elect(leader, sb, &tmB, &full_bar[stage], kb + cb * KC, nt * BN);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        // residual by tensor core: D += I[:, j*KC .. +KC) * R[m0 + j*KC .. +KC, n-tile]; the residual rows land
        // as an MN-major B operand, 64 channels (128 bytes) per slab
        if constexpr (BN % 64 == 0) {
          for (int j = 0; j < p.res_iters; ++j) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 1);
            uint8_t* sa = stage_base + stage * C::kStageBytes;
            uint8_t* sb = sa + C::kABytes;
            ptx::mbar_expect_tx_elect(leader, &full_bar[stage], C::kStageBytes);
            ptx::tma_load_2d_elect(leader, sa, &tmI, &full_bar[stage], j * KC, 0);
#pragma unroll
            for (int sl = 0; sl < BN / 64; ++sl)
              ptx::tma_load_2d_elect(leader, sb + sl * KC * 128, &tmR, &full_bar[stage], nt * BN + sl * 64, m0 + j * KC);
            if constexpr (kM2) {  // residual rows of sub-tile 1 go where its A tile would be (same size: BN == 128)
#pragma unroll
              for (int sl = 0; sl < BN / 64; ++sl)
                ptx::tma_load_2d_elect(leader, sa + C::kASubBytes + sl * KC * 128, &tmR, &full_bar[stage], nt * BN + sl * 64,
                                 m1 + j * KC);
            }
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t leader = ptx::elect_one() ? 1u : 0u;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      {  // the whole warp walks the loop (converged); one elected lane issues, see ptx::umma_bf16_elect
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
        for (int it = 0; it < kiters; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase, p.err, 3);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(stage_base + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint64_t adesc = ptx::make_smem_desc(sa, 16, C::kSBO, C::kLayout);
          const uint64_t bdesc = ptx::make_smem_desc(sb, 16, C::kSBO, C::kLayout);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the >>4 field
            ptx::umma_bf16_elect(leader, d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
            if constexpr (kM2) {
              const uint64_t adesc1 = ptx::make_smem_desc(sa + C::kASubBytes, 16, C::kSBO, C::kLayout);
              ptx::umma_bf16_elect(leader, d_tmem + BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
            }
          }
          ptx::umma_commit_elect(leader, &empty_bar[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (BN % 64 == 0) {
          constexpr uint32_t idesc_res = ptx::make_idesc_bf16(kBlockM, BN, 0, 1);  // B = residual rows, MN-major
          for (int j = 0; j < p.res_iters; ++j) {
            ptx::mbar_wait(&full_bar[stage], phase, p.err, 3);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(stage_base + stage * C::kStageBytes);
            const uint32_t sb = sa + C::kABytes;
            const uint64_t adesc = ptx::make_smem_desc(sa, 16, C::kSBO, C::kLayout);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              // MN-major B: LBO = distance between 64-channel slabs, SBO = 8 rows of 128 bytes
              const uint64_t bdesc = ptx::make_smem_desc(sb + k * 16 * 128, KC * 128, 8 * 128, 2);
              ptx::umma_bf16_elect(leader, d_tmem, adesc + 2 * k, bdesc, idesc_res, 1u);
              if constexpr (kM2) {
                const uint64_t bdesc1 =
                    ptx::make_smem_desc(sa + C::kASubBytes + k * 16 * 128, KC * 128, 8 * 128, 2);
                ptx::umma_bf16_elect(leader, d_tmem + BN, adesc + 2 * k, bdesc1, idesc_res, 1u);
              }
            }
            ptx::umma_commit_elect(leader, &empty_bar[stage]);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        ptx::umma_commit_elect(leader, &tfull_bar[acc]);
      }
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - kEpiWarp0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int half = ew >> 2;      // which chunks of the accumulator: half, half + 2, ...
    if constexpr (kTma && C::kWide) {
      // ---- staged epilogue, wide form (BN >= 128): a warp owns 64-column blocks (two TMEM chunks) so that every
      // TMA store / y-tile load moves full 128-byte rows (32 rows x 128 B = 4 KB, SWIZZLE_128B).  With 32-column
      // blocks the fused BN-backward reduction of the HBM-bound 1x1 gradients read y in 64-byte pieces and ran at
      // half the DRAM efficiency.
      constexpr int kBlocks = BN / 64;        // 2 or 4 blocks per tile
      // blocks per warp: block = half + 2 * bi; with 256-row tiles `half` selects the sub-tile and a warp takes all
      constexpr int kBPW = kM2 ? kBlocks : kBlocks / 2;
      auto block_of = [&](int bi) { return kM2 ? bi : half + 2 * bi; };
      const int sub_m = kM2 ? half * kBlockM : 0;   // row offset of this warp's sub-tile
      const int sub_c = kM2 ? half * BN : 0;        // TMEM column offset of its accumulator
      // staging tiles: slot s of warp ew at s_extra + (s * kEpiWarps + ew) * 4 KB.  With two slots (short-K tiles: the
      // 1x1 layers, where the epilogue is the critical path) a block is staged while the bulk store of the previous
      // one is still reading its tile; with one slot every block waited for that store first.
      uint8_t* const stg0 = s_extra + ew * 4096;
      const int stg_slots = p.stg_slots;
      uint32_t stg_count = 0;
      uint8_t* stg = stg0;
      const int sw_w = lane & 7;              // writer: row = lane, 16-byte piece j -> j ^ (row & 7)
      const int rq = lane & 7, rg = lane >> 3;  // reader: piece rq, rows rg + 4 i (i = 0..7)
      const int my_col = 8 * rq + 4 * ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1);  // first of the 2 columns owned
      float acc_s[kBPW][2], acc_t[kBPW][2];
#pragma unroll
      for (int i = 0; i < kBPW; ++i) acc_s[i][0] = acc_s[i][1] = acc_t[i][0] = acc_t[i][1] = 0.f;
      StatAcc* const stat_base = kBnRed ? p.bn_sums : p.stats;
      const int stat_parts = kBnRed ? p.bn_parts : p.stats_parts;
      StatAcc* stats_row =
          stat_base ? stat_base + static_cast<long long>(blockIdx.x % stat_parts) * 2 * p.stat_cols : nullptr;
      // The epilogue warps that own the same columns (the four lane quarters of one half; all eight with 256-row tiles)
      // first add their partial sums in shared memory, in warp order, so that ONE integer add per column and CTA goes
      // to global memory: with 148 CTAs x 4-8 warps hitting the same few addresses the atomics of the last wave
      // showed up as a 5-8 % tail on the 3x3 data gradients.  All eight warps flush at the same tiles.
      bool store_pending = false;
      auto flush = [&](int nt_) {
        if (store_pending) {  // the parking area is this warp's staging tile: its last bulk store must have read it
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
          store_pending = false;
        }
        float* park = reinterpret_cast<float*>(stg0);
#pragma unroll
        for (int bi = 0; bi < kBPW; ++bi) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            park[((bi * 2 + e) * 2 + 0) * 32 + lane] = acc_s[bi][e];
            park[((bi * 2 + e) * 2 + 1) * 32 + lane] = acc_t[bi][e];
            acc_s[bi][e] = acc_t[bi][e] = 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        const int g0 = kM2 ? 0 : 4 * half;
        if (ew == g0) {
#pragma unroll
          for (int bi = 0; bi < kBPW; ++bi) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float ssum = 0.f, tsum = 0.f;
              for (int w = g0; w < g0 + (kM2 ? 8 : 4); ++w) {
                const float* pk = reinterpret_cast<const float*>(s_extra + w * 4096);
                ssum += pk[((bi * 2 + e) * 2 + 0) * 32 + lane];
                tsum += pk[((bi * 2 + e) * 2 + 1) * 32 + lane];
              }
              const int n = nt_ * BN + block_of(bi) * 64 + my_col + e;
              if (n < p.Cout) {
                const int ch = n & p.stat_mask;  // depth-to-space: four column groups share a channel
                stat_add(stats_row + ch, ssum);
                stat_add(stats_row + p.stat_cols + ch, kBnRed ? tsum * __ldg(p.bn_rstd + ch) : tsum);
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      };
      // ---- fused BN backward: ring of TMA-loaded y tiles (4 KB each), `yslots` blocks ahead
      const int yslots = p.y_slots, ylog = p.y_slots_log2;
      uint8_t* const ybuf = s_extra + stg_slots * kEpiWarps * 4096 + ew * yslots * 4096;
      uint64_t* const ybar = y_bar + kMaxYSlots * ew;
      auto block_valid = [&](int tile, int bi) { return bi < kBPW && (tile % p.num_n_tiles) * BN + block_of(bi) * 64 < p.Cout; };
      auto next_block = [&](int& tile, int& bi) {
        do {
          if (++bi >= kBPW) {
            bi = 0;
            tile += gridDim.x;
          }
        } while (tile < num_tiles && !block_valid(tile, bi));
      };
      const uint32_t lead0 = lane == 0 ? 1u : 0u;  // issuing lane of this warp's TMA traffic (converged issue)
      auto issue_y = [&](int tile, int bi, int slot) {
        const int mt = tile / p.num_n_tiles;
        const int nt = tile - mt * p.num_n_tiles;
        ptx::mbar_expect_tx_elect(lead0, &ybar[slot], 4096);
        const int ncol0 = nt * BN + block_of(bi) * 64, mrow0 = mt * C::kTileM + sub_m + quarter * 32;
        if (p.d2s_c2) {  // y lies like the depth-to-space output: the same row groups as the stores
          const int cls_a = ncol0 / p.d2s_c2, col_in = ncol0 - cls_a * p.d2s_c2;
          int nq = mrow0 / p.OW, p0 = mrow0 - nq * p.OW;
          for (int r0 = 0; r0 < 32; r0 += p.d2s_g) {
            ptx::tma_load_3d_elect(lead0, ybuf + slot * 4096 + r0 * 128, &tmY, &ybar[slot], col_in, p0, 2 * nq + cls_a);
            p0 += p.d2s_g;
            if (p0 >= p.OW) { p0 = 0; ++nq; }
          }
        } else {
          ptx::tma_load_2d_elect(lead0, ybuf + slot * 4096, &tmY, &ybar[slot], ncol0, mrow0);
        }
      };
      uint32_t ycount = 0;
      int pf_tile = blockIdx.x, pf_bi = 0;
      if (kBnRed) {
        if (pf_tile < num_tiles && !block_valid(pf_tile, pf_bi)) next_block(pf_tile, pf_bi);
        for (int d = 0; d < yslots && pf_tile < num_tiles; ++d) {
          issue_y(pf_tile, pf_bi, d);
          next_block(pf_tile, pf_bi);
        }
      }
      int stat_nt = -1;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles;
        const int nt = tile - mt * p.num_n_tiles;
        if (stat_base && nt != stat_nt) {
          if (stat_nt >= 0) flush(stat_nt);
          stat_nt = nt;
        }
        const int m_warp = mt * C::kTileM + sub_m + quarter * 32;
        const int m = m_warp + lane;
        const bool row_ok = m < p.M_total;
        long long r_row = 0;
        if (p.res) {
          const int mm = row_ok ? m : 0;
          const int n_img = mm / p.OHW;
          const int rem = mm - n_img * p.OHW;
          const int pr = rem / p.OW;
          const int qc = rem - pr * p.OW;
          r_row = n_img * p.r_sn + pr * p.r_sh + qc * p.r_sw;
        }
        ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, 4);
        ptx::tc_fence_after();
        const uint32_t t_row =
            tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::kAccCols + sub_c;

#pragma unroll
        for (int bi = 0; bi < kBPW; ++bi) {
          const int b0 = block_of(bi) * 64;
          const int nb_base = nt * BN + b0;
          if (nb_base >= p.Cout || (p.dbg & 4)) break;  // warp-uniform
          // the bulk store that last used this staging slot must have finished READING it
          stg = stg0 + (stg_count & (stg_slots - 1)) * (kEpiWarps * 4096);
          ++stg_count;
          if (store_pending) {
            if (lane == 0) {
              if (stg_slots == 2) ptx::tma_store_wait_read<1>();
              else ptx::tma_store_wait_read<0>();
            }
            __syncwarp();
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int n_base = nb_base + 32 * cc;
            float v[32];
            {
              uint32_t r[32];
              ptx::tmem_ld_32x32(t_row + b0 + 32 * cc, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            }
            if (n_base < p.Cout) {
              const bool full = n_base + 32 <= p.Cout;
              if (p.scale) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  if (full || n_base + j + 4 <= p.Cout) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n_base + j));
                    v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
                  }
                }
              }
              if (p.shift) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  if (full || n_base + j + 4 <= p.Cout) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.shift + n_base + j));
                    v[j] += s4.x; v[j + 1] += s4.y; v[j + 2] += s4.z; v[j + 3] += s4.w;
                  }
                }
              }
              const float neg = p.act == 1 ? p.slope : (p.act == 2 ? 0.f : 1.f);
              if (p.res) {
                if (p.res_after_act) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
                }
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + r_row + n_base);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  if (full || n_base + j + 8 <= p.Cout) {
                    const uint4 q = __ldg(rp + (j >> 3));
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 f = __bfloat1622float2(h[e]);
                      v[j + 2 * e] += f.x;
                      v[j + 2 * e + 1] += f.y;
                    }
                  }
                }
                if (!p.res_after_act && p.act != 0) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
                }
              } else if (p.act != 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
              }
            }
            if (!row_ok || n_base >= p.Cout) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * cc + j) ^ sw_w) * 16)) = pk;
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (!(p.dbg & 1)) {
            if (p.d2s_c2) {
              // depth-to-space: this 64-column block belongs to output rows 2q + a.  Stored in groups of d2s_g
              // rows (d2s_g divides OW, so a group never crosses the end of a dy row -- a TMA store must not start
              // at a negative coordinate); the swizzle is a function of the shared-memory address, so a group at
              // any row offset of the staging tile is read back correctly
              const int cls_a = nb_base / p.d2s_c2, col_in = nb_base - cls_a * p.d2s_c2;
              int nq = m_warp / p.OW, p0 = m_warp - nq * p.OW;
              for (int r0 = 0; r0 < 32; r0 += p.d2s_g) {
                ptx::tma_store_3d_elect(lead0, &tmO, stg + r0 * 128, col_in, p0, 2 * nq + cls_a);
                p0 += p.d2s_g;
                if (p0 >= p.OW) { p0 = 0; ++nq; }
              }
              ptx::tma_commit_elect(lead0);
            } else {
              ptx::tma_store_2d_elect(lead0, &tmO, stg, nb_base, m_warp);  // store + commit
            }
          }
          store_pending = true;
          if ((kBnRed || p.stats) && !(p.dbg & 2)) {
            float s[8], t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = t[k] = 0.f;
            if constexpr (kBnRed) {
              // dz = G * act'(y*scale+shift) from the STORED gradient tile and the TMA-loaded y tile
              const int yslot = ycount & (yslots - 1);
              ptx::mbar_wait(&ybar[yslot], (ycount >> ylog) & 1, p.err, 5);
              const uint8_t* ytile = ybuf + yslot * 4096;
              const int ncol = (nb_base + 8 * rq) & p.stat_mask;  // channel of the first of this lane's 8 columns
              float sc[8], sh[8], mu[8];
              if (nb_base + 8 * rq + 8 <= p.Cout) {
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
                  const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.bn_scale + ncol + k));
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bn_shift + ncol + k));
                  const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.bn_mean + ncol + k));
                  sc[k] = a4.x; sc[k + 1] = a4.y; sc[k + 2] = a4.z; sc[k + 3] = a4.w;
                  sh[k] = b4.x; sh[k + 1] = b4.y; sh[k + 2] = b4.z; sh[k + 3] = b4.w;
                  mu[k] = c4.x; mu[k + 1] = c4.y; mu[k + 2] = c4.z; mu[k + 3] = c4.w;
                }
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) sc[k] = sh[k] = mu[k] = 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = rg + 4 * i;
                const int off = row * 128 + ((rq ^ (row & 7)) * 16);
                const uint4 ug = *reinterpret_cast<const uint4*>(stg + off);
                const uint4 uy = *reinterpret_cast<const uint4*>(ytile + off);
                const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
                const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&uy);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 g2 = __bfloat1622float2(hg[e]);
                  const float2 y2 = __bfloat1622float2(hy[e]);
                  const float dz0 = g2.x * (fmaf(y2.x, sc[2 * e], sh[2 * e]) > 0.f ? 1.f : p.bn_neg);
                  const float dz1 = g2.y * (fmaf(y2.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f ? 1.f : p.bn_neg);
                  s[2 * e] += dz0;
                  s[2 * e + 1] += dz1;
                  t[2 * e] = fmaf(dz0, y2.x - mu[2 * e], t[2 * e]);
                  t[2 * e + 1] = fmaf(dz1, y2.y - mu[2 * e + 1], t[2 * e + 1]);
                }
              }
            } else {
              // column sums / sums of squares of the STORED (bf16-rounded) tile
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = rg + 4 * i;
                const uint4 u = *reinterpret_cast<const uint4*>(stg + row * 128 + ((rq ^ (row & 7)) * 16));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  s[2 * e] += f.x;
                  s[2 * e + 1] += f.y;
                  t[2 * e] = fmaf(f.x, f.x, t[2 * e]);
                  t[2 * e + 1] = fmaf(f.y, f.y, t[2 * e + 1]);
                }
              }
            }
            // keep-half butterfly over the 4 row-group lanes (lane bits 4, 3): two columns per lane remain
            keep_half_step<4, 16>(s, t, lane);
            keep_half_step<2, 8>(s, t, lane);
            acc_s[bi][0] += s[0];
            acc_s[bi][1] += s[1];
            acc_t[bi][0] += t[0];
            acc_t[bi][1] += t[1];
            __syncwarp();  // all lanes are done reading: the tiles may be overwritten / refilled
            if constexpr (kBnRed) {
              const int yslot = ycount & (yslots - 1);
              ++ycount;
              if (pf_tile < num_tiles) {
                issue_y(pf_tile, pf_bi, yslot);
                next_block(pf_tile, pf_bi);
              }
            }
          }
        }
        // accumulator drained: hand the TMEM stage back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (stat_base && stat_nt >= 0) flush(stat_nt);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
      __syncwarp();
    } else if constexpr (kTma) {
      // ---- staged epilogue: registers -> swizzled smem tile -> TMA bulk store; statistics from the staged tile
      constexpr int RB = C::kChunk * 2;  // bytes per staged row: 64 (SWIZZLE_64B) or 32 (SWIZZLE_32B)
      constexpr int NJ = RB / 16;        // 16-byte pieces per row
      uint8_t* stg = s_extra + ew * 2048;
      const int sw_w = RB == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);  // writer: row = lane
      // reader of the statistics: lane = (q: 16-byte piece, g: row group); rows g + (32/NR)*i, piece q
      constexpr int NR = RB == 64 ? 4 : 2;                // rows per reader lane
      const int rq = RB == 64 ? (lane & 3) : (lane & 1);
      const int rg = RB == 64 ? (lane >> 2) : (lane >> 1);
      const int sw_r = RB == 64 ? ((rg >> 1) & 3) : ((rg >> 2) & 1);
      const uint8_t* rd0 = stg + rg * RB + ((rq ^ sw_r) * 16);
      const int my_col = RB == 64 ? (8 * rq + rg) : (8 * rq + (rg >> 1));  // column this lane ends up owning
      const bool col_owner = RB == 64 ? true : ((lane & 2) == 0);
      float acc_s[C::kChunksPerWarp], acc_t[C::kChunksPerWarp];
#pragma unroll
      for (int i = 0; i < C::kChunksPerWarp; ++i) acc_s[i] = acc_t[i] = 0.f;
      // statistics target: forward [sum | sumsq] (p.stats) or, fused BN backward, [sum dz | sum dz*xhat] (p.bn_sums)
      StatAcc* const stat_base = kBnRed ? p.bn_sums : p.stats;
      const int stat_parts = kBnRed ? p.bn_parts : p.stats_parts;
      StatAcc* stats_row =
          stat_base ? stat_base + static_cast<long long>(blockIdx.x % stat_parts) * 2 * p.Cout : nullptr;
      // ---- fused BN backward: ring of two TMA-loaded y tiles, one chunk ahead of the one being processed
      const int yslots = p.y_slots, ylog = p.y_slots_log2;  // power of two
      uint8_t* const ybuf = s_extra + kEpiWarps * 2048 + ew * yslots * 2048;
      uint64_t* const ybar = y_bar + kMaxYSlots * ew;
      const uint32_t y_off = static_cast<uint32_t>(rd0 - stg);  // same (row, piece) as in the staging tile
      auto chunk_valid = [&](int tile, int ci) {
        const int c0 = (half + 2 * ci) * C::kChunk;
        return ci < C::kChunksPerWarp && c0 < BN && (tile % p.num_n_tiles) * BN + c0 < p.Cout;
      };
      auto next_chunk = [&](int& tile, int& ci) {  // in processing order; tile >= num_tiles = none left
        do {
          if (++ci >= C::kChunksPerWarp) {
            ci = 0;
            tile += gridDim.x;
          }
        } while (tile < num_tiles && !chunk_valid(tile, ci));
      };
      const uint32_t lead0 = lane == 0 ? 1u : 0u;  // issuing lane of this warp's TMA traffic (converged issue)
      auto issue_y = [&](int tile, int ci, int slot) {
        const int mt = tile / p.num_n_tiles;
        const int nt = tile - mt * p.num_n_tiles;
        ptx::mbar_expect_tx_elect(lead0, &ybar[slot], 32 * RB);
        ptx::tma_load_2d_elect(lead0, ybuf + slot * 2048, &tmY, &ybar[slot], nt * BN + (half + 2 * ci) * C::kChunk,
                               mt * kBlockM + quarter * 32);
      };
      uint32_t ycount = 0;  // chunks processed by this warp: slot = ycount % yslots, phase = (ycount / yslots) & 1
      int pf_tile = blockIdx.x, pf_ci = 0;  // next chunk whose y tile has not been requested yet
      if (kBnRed) {
        if (pf_tile < num_tiles && !chunk_valid(pf_tile, pf_ci)) next_chunk(pf_tile, pf_ci);
        for (int d = 0; d < yslots && pf_tile < num_tiles; ++d) {
          issue_y(pf_tile, pf_ci, d);
          next_chunk(pf_tile, pf_ci);
        }
      }
      int stat_nt = -1;
      bool store_pending = false;
      // the four lane-quarter warps of one half own the same columns: add their partial sums in shared memory (warp
      // order) and send ONE integer add per column and CTA to global memory (see the wide form above)
      auto flush_narrow = [&](int nt_) {
        if (store_pending) {
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
          store_pending = false;
        }
        float* park = reinterpret_cast<float*>(stg);
#pragma unroll
        for (int ci = 0; ci < C::kChunksPerWarp; ++ci) {
          park[(ci * 2 + 0) * 32 + lane] = acc_s[ci];
          park[(ci * 2 + 1) * 32 + lane] = acc_t[ci];
          acc_s[ci] = acc_t[ci] = 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (ew == 4 * half) {
#pragma unroll
          for (int ci = 0; ci < C::kChunksPerWarp; ++ci) {
            float ssum = 0.f, tsum = 0.f;
            for (int w = 4 * half; w < 4 * half + 4; ++w) {
              const float* pk = reinterpret_cast<const float*>(s_extra + w * 2048);
              ssum += pk[(ci * 2 + 0) * 32 + lane];
              tsum += pk[(ci * 2 + 1) * 32 + lane];
            }
            const int n = nt_ * BN + (half + 2 * ci) * C::kChunk + my_col;
            if (col_owner && (half + 2 * ci) < C::kNumChunks && n < p.Cout) {
              stat_add(stats_row + n, ssum);
              stat_add(stats_row + p.Cout + n, kBnRed ? tsum * __ldg(p.bn_rstd + n) : tsum);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      };
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles;
        const int nt = tile - mt * p.num_n_tiles;
        if (stat_base && nt != stat_nt) {
          if (stat_nt >= 0) flush_narrow(stat_nt);
          stat_nt = nt;
        }
        const int m_warp = mt * kBlockM + quarter * 32;  // first output row of this warp
        const int m = m_warp + lane;
        const bool row_ok = m < p.M_total;
        long long r_row = 0;
        if (p.res) {
          const int mm = row_ok ? m : 0;
          const int n_img = mm / p.OHW;
          const int rem = mm - n_img * p.OHW;
          const int pr = rem / p.OW;
          const int qc = rem - pr * p.OW;
          r_row = n_img * p.r_sn + pr * p.r_sh + qc * p.r_sw;
        }
        ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, 4);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;

#pragma unroll
        for (int ci = 0; ci < C::kChunksPerWarp; ++ci) {
          const int c0 = (half + 2 * ci) * C::kChunk;
          const int n_base = nt * BN + c0;
          if (c0 >= BN || n_base >= p.Cout || (p.dbg & 4)) break;  // warp-uniform
          float v[C::kChunk];
          if constexpr (C::kChunk == 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_row + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {
            uint32_t r[16];
            ptx::tmem_ld_32x16(t_row + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          // columns >= Cout of a ragged last chunk are clipped by the TMA store; keep their loads in bounds
          const bool full = n_base + C::kChunk <= p.Cout;
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 4) {
              if (full || n_base + j + 4 <= p.Cout) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n_base + j));
                v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
              }
            }
          }
          if (p.shift) {
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 4) {
              if (full || n_base + j + 4 <= p.Cout) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.shift + n_base + j));
                v[j] += s4.x; v[j + 1] += s4.y; v[j + 2] += s4.z; v[j + 3] += s4.w;
              }
            }
          }
          const float neg = p.act == 1 ? p.slope : (p.act == 2 ? 0.f : 1.f);
          if (p.res) {
            if (p.res_after_act) {
#pragma unroll
              for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
            }
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + r_row + n_base);
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 8) {
              if (full || n_base + j + 8 <= p.Cout) {
                const uint4 q = __ldg(rp + (j >> 3));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  v[j + 2 * e] += f.x;
                  v[j + 2 * e + 1] += f.y;
                }
              }
            }
            if (!p.res_after_act && p.act != 0) {
#pragma unroll
              for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
            }
          } else if (p.act != 0) {
#pragma unroll
            for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
          }
          if (!row_ok) {  // rows past M_total are clipped by the store and must not be counted
#pragma unroll
            for (int j = 0; j < C::kChunk; ++j) v[j] = 0.f;
          }
          // the previous bulk store of this warp must have finished READING the staging tile
          if (store_pending) {
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(stg + lane * RB + ((j ^ sw_w) * 16)) = pk;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (!(p.dbg & 1)) ptx::tma_store_2d_elect(lead0, &tmO, stg, n_base, m_warp);  // store + commit
          store_pending = true;
          if constexpr (kBnRed) {
            // fused BN-backward reduction: dz = G * act'(y*scale+shift) from the STORED gradient tile and the
            // TMA-loaded y tile; per column sum dz and sum dz*(y-mean) (rstd is applied at the flush)
            const int yslot = ycount & (yslots - 1);
            ptx::mbar_wait(&ybar[yslot], (ycount >> ylog) & 1, p.err, 5);
            const uint8_t* yrd = ybuf + yslot * 2048 + y_off;
            const int ncol = n_base + 8 * rq;
            float sc[8], sh[8], mu[8];
            if (ncol + 8 <= p.Cout) {
#pragma unroll
              for (int k = 0; k < 8; k += 4) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.bn_scale + ncol + k));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bn_shift + ncol + k));
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.bn_mean + ncol + k));
                sc[k] = a4.x; sc[k + 1] = a4.y; sc[k + 2] = a4.z; sc[k + 3] = a4.w;
                sh[k] = b4.x; sh[k + 1] = b4.y; sh[k + 2] = b4.z; sh[k + 3] = b4.w;
                mu[k] = c4.x; mu[k + 1] = c4.y; mu[k + 2] = c4.z; mu[k + 3] = c4.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) sc[k] = sh[k] = mu[k] = 0.f;
            }
            float s[8], t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = t[k] = 0.f;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              const uint4 ug = *reinterpret_cast<const uint4*>(rd0 + i * (32 / NR) * RB);
              const uint4 uy = *reinterpret_cast<const uint4*>(yrd + i * (32 / NR) * RB);
              const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
              const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&uy);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 g2 = __bfloat1622float2(hg[e]);
                const float2 y2 = __bfloat1622float2(hy[e]);
                const float dz0 = g2.x * (fmaf(y2.x, sc[2 * e], sh[2 * e]) > 0.f ? 1.f : p.bn_neg);
                const float dz1 = g2.y * (fmaf(y2.y, sc[2 * e + 1], sh[2 * e + 1]) > 0.f ? 1.f : p.bn_neg);
                s[2 * e] += dz0;
                s[2 * e + 1] += dz1;
                t[2 * e] = fmaf(dz0, y2.x - mu[2 * e], t[2 * e]);
                t[2 * e + 1] = fmaf(dz1, y2.y - mu[2 * e + 1], t[2 * e + 1]);
              }
            }
            keep_half_step<4, 16>(s, t, lane);
            keep_half_step<2, 8>(s, t, lane);
            keep_half_step<1, 4>(s, t, lane);
            if (RB == 32) {
              s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
              t[0] += __shfl_xor_sync(0xffffffffu, t[0], 2);
            }
            acc_s[ci] += s[0];
            acc_t[ci] += t[0];
            ++ycount;
            __syncwarp();  // both tiles are free again: refill the y slot with the chunk `yslots` ahead
            if (pf_tile < num_tiles) {
              issue_y(pf_tile, pf_ci, yslot);
              next_chunk(pf_tile, pf_ci);
            }
          } else if (p.stats && !(p.dbg & 2)) {
            // column sums of the STORED (bf16-rounded) tile: each lane adds NR rows of one 16-byte piece, a
            // keep-half butterfly over the row-group lanes leaves one column per lane
            float s[8], t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = t[k] = 0.f;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              const uint4 u = *reinterpret_cast<const uint4*>(rd0 + i * (32 / NR) * RB);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                s[2 * e] += f.x;
                s[2 * e + 1] += f.y;
                t[2 * e] = fmaf(f.x, f.x, t[2 * e]);
                t[2 * e + 1] = fmaf(f.y, f.y, t[2 * e + 1]);
              }
            }
            keep_half_step<4, 16>(s, t, lane);
            keep_half_step<2, 8>(s, t, lane);
            keep_half_step<1, 4>(s, t, lane);
            if (RB == 32) {  // 16 row-group lanes: one more (full) step
              s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
              t[0] += __shfl_xor_sync(0xffffffffu, t[0], 2);
            }
            acc_s[ci] += s[0];
            acc_t[ci] += t[0];
            __syncwarp();  // all lanes are done reading before the tile is overwritten
          }
        }
        // accumulator drained: hand the TMEM stage back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (stat_base && stat_nt >= 0) flush_narrow(stat_nt);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
      __syncwarp();
    } else {
    float* scratch = s_scratch + ew * 32 * kScratchLd;
    float* my_sum = s_stats + ew * 2 * C::kAccPerWarp;  // [kAccPerWarp sums | kAccPerWarp sums of squares]
    float* my_sq = my_sum + C::kAccPerWarp;
    StatAcc* stats_row = p.stats ? p.stats + static_cast<long long>(blockIdx.x % p.stats_parts) * 2 * p.Cout : nullptr;
    int stat_nt = -1;  // n-tile the warp-private statistics belong to
    // adds the warp-private partial sums of n-tile `nt` to this CTA's row of the global partials and clears them
    auto flush_stats = [&](int nt) {
      for (int i = lane; i < C::kAccPerWarp; i += 32) {
        const int ci = i / C::kChunk;
        const int n = nt * BN + (half + 2 * ci) * C::kChunk + (i - ci * C::kChunk);
        const float s1 = my_sum[i], s2 = my_sq[i];
        if (n < p.Cout && (s1 != 0.f || s2 != 0.f)) {
          stat_add(stats_row + n, s1);
          stat_add(stats_row + p.Cout + n, s2);
        }
        my_sum[i] = 0.f;
        my_sq[i] = 0.f;
      }
      __syncwarp();
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles;
      const int nt = tile - mt * p.num_n_tiles;
      if (p.stats && nt != stat_nt) {
        if (stat_nt >= 0) flush_stats(stat_nt);
        stat_nt = nt;
      }
      const int m = mt * kBlockM + quarter * 32 + lane;
      const bool row_ok = m < p.M_total;
      long long o_row = 0, r_row = 0;
      {
        const int mm = row_ok ? m : 0;
        const int n_img = mm / p.OHW;
        const int rem = mm - n_img * p.OHW;
        const int pr = rem / p.OW;
        const int qc = rem - pr * p.OW;
        o_row = n_img * p.o_sn + pr * p.o_sh + qc * p.o_sw;
        r_row = n_img * p.r_sn + pr * p.r_sh + qc * p.r_sw;
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, 4);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;

#pragma unroll 1
      for (int ci = 0; ci < C::kChunksPerWarp; ++ci) {
        const int c0 = (half + 2 * ci) * C::kChunk;
        const int n_base = nt * BN + c0;
        if (c0 >= BN || n_base >= p.Cout) break;  // warp-uniform
        if (p.dbg & 4) continue;
        float v[C::kChunk];
        if constexpr (C::kChunk == 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
          uint32_t r[16];
          ptx::tmem_ld_32x16(t_row + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
        const bool fast = p.vec_ok && (n_base + C::kChunk <= p.Cout) && (!p.res || p.res_vec_ok);
        if (fast) {
          // branch-free per element: every branch costs an exposed latency in these warps
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n_base + j));
              v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
            }
          }
          if (p.shift) {
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.shift + n_base + j));
              v[j] += s4.x; v[j + 1] += s4.y; v[j + 2] += s4.z; v[j + 3] += s4.w;
            }
          }
          const float neg = p.act == 1 ? p.slope : (p.act == 2 ? 0.f : 1.f);
          if (p.res) {
            if (p.res_after_act) {
#pragma unroll
              for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
            }
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + (row_ok ? r_row : 0) + n_base);
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 8) {
              const uint4 q = __ldg(rp + (j >> 3));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[j + 2 * e] += f.x;
                v[j + 2 * e + 1] += f.y;
              }
            }
            if (p.res_lo) {  // split residual: add the lower pieces too
#pragma unroll
              for (int pc = 1; pc < kSplitPieces; ++pc) {
                const uint4* rl = reinterpret_cast<const uint4*>(p.res + (row_ok ? r_row : 0) + pc * p.res_lo + n_base);
#pragma unroll
                for (int j = 0; j < C::kChunk; j += 8) {
                  const uint4 q = __ldg(rl + (j >> 3));
                  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h[e]);
                    v[j + 2 * e] += f.x;
                    v[j + 2 * e + 1] += f.y;
                  }
                }
              }
            }
            if (!p.res_after_act && p.act != 0) {
#pragma unroll
              for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
            }
          } else if (p.act != 0) {
#pragma unroll
            for (int j = 0; j < C::kChunk; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * neg;
          }
          if (!row_ok) {  // rows past M_total: nothing stored, nothing counted
#pragma unroll
            for (int j = 0; j < C::kChunk; ++j) v[j] = 0.f;
          }
          if (p.out_fp32) {
            if (row_ok) {
              float* o = reinterpret_cast<float*>(p.out) + o_row + n_base;
#pragma unroll
              for (int j = 0; j < C::kChunk; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else if (p.out_lo) {
            // fp32-parity mode: v = sum of kSplitPieces bf16 pieces, out_lo elements apart; statistics see that sum
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + o_row + n_base;
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 8) {
              float rem[8], tot[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { rem[e] = v[j + e]; tot[e] = 0.f; }
#pragma unroll
              for (int pc = 0; pc < kSplitPieces; ++pc) {
                uint4 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  h[e] = __floats2bfloat162_rn(rem[2 * e], rem[2 * e + 1]);
                  const float2 f = __bfloat1622float2(h[e]);
                  rem[2 * e] -= f.x;
                  rem[2 * e + 1] -= f.y;
                }
                if (row_ok) *reinterpret_cast<uint4*>(o + pc * p.out_lo + j) = pk;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j + e] -= rem[e];  // = the stored value (exact: the pieces do not overlap)
              (void)tot;
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + o_row + n_base;
#pragma unroll
            for (int j = 0; j < C::kChunk; j += 8) {
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
                const float2 f = __bfloat1622float2(h[e]);  // statistics see the stored (rounded) value
                v[j + 2 * e] = f.x;
                v[j + 2 * e + 1] = f.y;
              }
              if (row_ok && !(p.dbg & 1)) *reinterpret_cast<uint4*>(o + j) = pk;
            }
          }
        } else {
          // generic path: ragged channel tail, strided / unaligned outputs (NCHW heads, odd pitches)
#pragma unroll
          for (int j = 0; j < C::kChunk; ++j) {
            const int n = n_base + j;
            float x = v[j];
            if (n < p.Cout && row_ok) {
              if (p.scale) x *= __ldg(p.scale + n);
              if (p.shift) x += __ldg(p.shift + n);
              float rv = 0.f;
              if (p.res) {
                rv = __bfloat162float(p.res[r_row + n * p.r_sc]);
                if (p.res_lo)
                  for (int pc = 1; pc < kSplitPieces; ++pc)
                    rv += __bfloat162float(p.res[r_row + pc * p.res_lo + n * p.r_sc]);
              }
              if (p.res && !p.res_after_act) x += rv;
              x = apply_act(x, p.act, p.slope);
              if (p.res && p.res_after_act) x += rv;
              if (p.out_fp32) {
                reinterpret_cast<float*>(p.out)[o_row + n * p.o_sc] = x;
              } else if (p.out_lo) {
                float rem = x;
                for (int pc = 0; pc < kSplitPieces; ++pc) {
                  const __nv_bfloat16 hb = __float2bfloat16_rn(rem);
                  reinterpret_cast<__nv_bfloat16*>(p.out)[o_row + pc * p.out_lo + n * p.o_sc] = hb;
                  rem -= __bfloat162float(hb);
                }
                x -= rem;
              } else {
                const __nv_bfloat16 hb = __float2bfloat16_rn(x);
                reinterpret_cast<__nv_bfloat16*>(p.out)[o_row + n * p.o_sc] = hb;
                x = __bfloat162float(hb);
              }
            } else {
              x = 0.f;
            }
            v[j] = x;
          }
        }
        // per-channel sum / sum of squares over the 32 rows of this warp: transpose 16 columns at a time
        // through shared memory; lane l sums column (l & 15) over rows 16*(l >> 4) .. +15 with independent
        // partial sums, the two half-warps are combined with one shuffle
        if (p.stats && !(p.dbg & 2)) {
#pragma unroll
          for (int hh = 0; hh < C::kChunk / 16; ++hh) {
#pragma unroll
            for (int j = 0; j < 16; ++j) scratch[lane * kScratchLd + j] = v[hh * 16 + j];
            __syncwarp();
            const float* col = scratch + (lane >> 4) * 16 * kScratchLd + (lane & 15);
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int r = 0; r < 16; r += 2) {
              const float x0 = col[r * kScratchLd], x1 = col[(r + 1) * kScratchLd];
              a0 += x0;
              a1 += x1;
              b0 = fmaf(x0, x0, b0);
              b1 = fmaf(x1, x1, b1);
            }
            float s1 = a0 + a1, s2 = b0 + b1;
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (lane < 16) {
              const int i = ci * C::kChunk + hh * 16 + lane;
              my_sum[i] += s1;
              my_sq[i] += s2;
            }
            __syncwarp();
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.stats && stat_nt >= 0) flush_stats(stat_nt);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int KC, int BN, int kMode, bool kM2 = false>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmI,
               const CUtensorMap& tmR, const CUtensorMap& tmY, const IgemmParams& p, cudaStream_t stream) {
  using C = Cfg<KC, BN, kMode, kM2>;
  static_assert(C::stages_for(C::kBnRed ? 1 : 0) >= 2, "pipeline too shallow");
  IgemmParams q = p;
  // y ring depth of the fused BN-backward reduction: tiles with few k-iterations (1x1 layers) finish a chunk in
  // less than one memory latency, so they need several y tiles in flight and can spare pipeline stages
  q.y_slots = 0;
  q.y_slots_log2 = 0;
  if (C::kBnRed) {
    const int kit = p.num_taps * p.cblocks + p.res_iters;
    static const int forced = getenv("B200CV_Y_SLOTS") ? atoi(getenv("B200CV_Y_SLOTS")) : 0;  // tuning aid
    q.y_slots = forced ? forced : (kit >= 16 ? 1 : 4);
    while (q.y_slots > 1 && C::stages_for(q.y_slots) < (forced ? 2 : 3)) q.y_slots >>= 1;
    q.y_slots_log2 = q.y_slots == 4 ? 2 : (q.y_slots == 2 ? 1 : 0);
  }
  // second staging tile per epilogue warp for short-K tiles of the wide staged epilogue (1x1 layers: the pipeline
  // needs few stages there and the epilogue is the critical path)
  q.stg_slots = 1;
  {
    static const int forced = getenv("B200CV_STG_SLOTS") ? atoi(getenv("B200CV_STG_SLOTS")) : 0;
    const int kit = p.num_taps * p.cblocks + p.res_iters;
    if (C::kTma && C::kWide && (forced ? forced == 2 : (kit <= 8 && !C::kBnRed))) {
      // trade y-ring depth for the second staging tile, never below 3 pipeline stages
      int ys = q.y_slots;
      while (ys > 1 && C::stages_for(ys, 2) < 3) ys >>= 1;
      if (C::stages_for(ys, 2) >= 3 && (!C::kBnRed || ys >= 2 || q.y_slots == 1)) {
        q.stg_slots = 2;
        q.y_slots = ys;
        q.y_slots_log2 = ys == 4 ? 2 : (ys == 2 ? 1 : 0);
      }
    }
  }
  q.stages = C::stages_for(q.y_slots, q.stg_slots);
  const int smem_bytes = C::extra_bytes(q.y_slots, q.stg_slots) + q.stages * C::kStageBytes;
  constexpr int kSmemAttr = C::kCtasPerSm == 2 ? kSmemMax2 : kSmemMax;
  static bool configured = false;  // benign race: attribute set is idempotent
  auto kern = igemm_kernel<KC, BN, kMode, kM2>;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttr);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "igemm smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = sm_count() * C::kCtasPerSm;
  if (grid > tiles) grid = tiles;
  if (grid < 1) return 0;
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmO, tmI, tmR, tmY, q);
  if (le != cudaSuccess) return set_error(static_cast<int>(le), "igemm launch: %s", cudaGetErrorString(le));
  return check_launch("igemm_kernel");
}

}  // namespace

int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO, const CUtensorMap* tmI,
                 const CUtensorMap* tmR, const CUtensorMap* tmY, const IgemmParams& p, int kc, int block_n,
                 cudaStream_t stream) {
  if (p.bn_sums && (!tmO || !tmY))
    return set_error(B200CV_ERR_ARG, "igemm: the fused BN-backward reduction needs the staged (TMA-store) epilogue");
  if (p.res_iters && (!tmI || !tmR || block_n % 64 != 0))
    return set_error(B200CV_ERR_ARG, "igemm: tensor-core residual needs identity/residual maps and block_n %% 64 == 0");
  const CUtensorMap& mI = tmI ? *tmI : tmA;  // unused copies when the feature is off
  const CUtensorMap& mR = tmR ? *tmR : tmA;
  if (p.tile_m == 256) {  // two 128-row sub-tiles per tile: BN == 128, KC == 64, staged epilogue only
    if (kc != 64 || block_n != 128 || !tmO)
      return set_error(B200CV_ERR_ARG, "igemm: the 256-row tile needs kc 64, block_n 128 and the staged epilogue");
    return p.bn_sums ? launch_one<64, 128, 2, true>(tmA, tmB, *tmO, mI, mR, *tmY, p, stream)
                     : launch_one<64, 128, 1, true>(tmA, tmB, *tmO, mI, mR, tmA, p, stream);
  }
#define B200CV_IGEMM_CASE(KC_, BN_)                                                        \
  if (kc == KC_ && block_n == BN_)                                                         \
    return !tmO ? launch_one<KC_, BN_, 0>(tmA, tmB, tmA, mI, mR, tmA, p, stream)           \
                : (p.bn_sums ? launch_one<KC_, BN_, 2>(tmA, tmB, *tmO, mI, mR, *tmY, p, stream) \
                             : launch_one<KC_, BN_, 1>(tmA, tmB, *tmO, mI, mR, tmA, p, stream));
  B200CV_IGEMM_CASE(64, 256)
  B200CV_IGEMM_CASE(64, 128)
  B200CV_IGEMM_CASE(64, 64)
  B200CV_IGEMM_CASE(64, 32)
  B200CV_IGEMM_CASE(64, 16)
  B200CV_IGEMM_CASE(32, 256)
  B200CV_IGEMM_CASE(32, 128)
  B200CV_IGEMM_CASE(32, 64)
  B200CV_IGEMM_CASE(32, 32)
  B200CV_IGEMM_CASE(32, 16)
  B200CV_IGEMM_CASE(16, 256)
  B200CV_IGEMM_CASE(16, 128)
  B200CV_IGEMM_CASE(16, 64)
  B200CV_IGEMM_CASE(16, 32)
  B200CV_IGEMM_CASE(16, 16)
#undef B200CV_IGEMM_CASE
  return set_error(B200CV_ERR_ARG, "igemm: unsupported tile kc=%d block_n=%d", kc, block_n);
}

}  // namespace b200cv
