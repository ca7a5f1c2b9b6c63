// Weight-gradient convolution for sm_100a:
//
//   dW[o][tap][i] += sum_{pix} dY[pix][o] * X[pix + tap][i]
//
// GEMM view per tap: M = output channels (128-row tile), N = input channels (BNW-column tile),
// K = output pixels.  Both operands are "MN-major" for tcgen05 (the contiguous dimension in
// memory is the channel, K = pixel is the strided one), which is exactly how TMA lays an
// NHWC tile down: 64 pixel rows x (<=128 bytes of channels), swizzled.
//   A = dY tile  : tiled 2-D TMA over [pixels][dy_ld], one or two 64-channel slabs
//   B = X tile   : im2col TMA (same map family as the forward pass), one per filter tap
// One dY tile is reused for T taps (T accumulators in TMEM), split-K over the pixel range; the epilogue
// stages 32x32 fp32 tiles in swizzled shared memory and lets the TMA unit reduce-add them into the packed
// gradient (cp.reduce.async.bulk.tensor .add): the per-thread red.global.add.v4 version spent ~40 LSU cycles
// per warp instruction (32 rows = 32 lines) in a non-overlapped tail of every CTA.  CTAs that share a pixel
// range (all output-channel tiles / tap groups of one k-split) are adjacent in the grid so that range is
// fetched from HBM once and served from L2 to the others.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace b200cv {
namespace {

constexpr int kWThreads = 256;   // warp 0, 6, 7: TMA producers; warp 1: MMA issuer; warps 2..5: epilogue
constexpr int kWProducers = 3;
constexpr int kMaxWStages = 8;

struct WgradParams {
  int M_pix, OHW, OW, OH;
  int lower_w, lower_h, trav_w, trav_h;
  int Cout, Cin, RS, Ipad;  // Ipad: row pitch of dW in channels (== Cin)
  int num_o_tiles, num_i_tiles, num_tap_groups, ksplit;
  int T;         // taps per CTA
  int a_slabs;   // 1 or 2 64-channel slabs of dY
  int stages;
  int kb_total;  // ceil(M_pix / 64)
  float* dw;
  int* err;
  // fp32-parity (split) mode: npass = 6 passes per k-block over the piece pairs (dY_i, X_j), i + j <= 2, smallest
  // first; a_lo / b_lo = piece stride in dY / X rows (npass = 1, offsets 0 in the bf16 mode)
  int npass, a_lo, b_lo;
  short tap_w[kMaxTaps];
  short tap_h[kMaxTaps];
};

// CB: channels per B slab (16/32/64), BNW: N tile (CB, or 128 = two 64-channel slabs), kKB: pixels per k-block.
// kKB = 64 for the multi-tap (3x3) layers; the single-tap layers (1x1 convolutions, the flat image layers) were bound
// by the ISSUE rate of their one producer thread -- one TMA instruction per ~166 ns whatever its size, 2-4 of them
// per 64-pixel block against 130 ns of tensor work -- and take 128- or 256-pixel blocks.
template <int CB, int BNW, int kKB>
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ CUtensorMap tmDW, const __grid_constant__ WgradParams p) {
  constexpr int kBSlabs = BNW / CB;            // 1 or 2
  constexpr int kASlabBytes = kKB * 128;       // 64 pixels x 64 channels bf16
  constexpr int kBSlabBytes = kKB * CB * 2;
  constexpr int kBTapBytes = kBSlabs * kBSlabBytes;
  constexpr int kBRow = CB * 2;
  constexpr int kBLayout = CB == 64 ? 2 : (CB == 32 ? 4 : 6);
  constexpr int kChunk = BNW >= 32 ? 32 : 16;

  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B (128B-swizzle atom) by pointer arithmetic so the shared state space stays provable
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_bytes = p.a_slabs * kASlabBytes;
  const int stage_bytes = p.a_slabs * kASlabBytes + p.T * kBTapBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxWStages;
  uint64_t* done_bar = empty_bar + kMaxWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  // per-epilogue-warp staging tile [32 rows][kChunk fp32], 1 KB aligned (swizzle atom)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(tmem_slot + 4);
  s_stage += (1024u - (ptx::smem_u32(s_stage) & 1023u)) & 1023u;

  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode (k-split major): blockIdx.x = ((ks * num_o_tiles + o_tile) * num_i_tiles + i_tile) * num_tap_groups + tg
  int b = blockIdx.x;
  const int tg = b % p.num_tap_groups; b /= p.num_tap_groups;
  const int it = b % p.num_i_tiles; b /= p.num_i_tiles;
  const int ot = b % p.num_o_tiles; b /= p.num_o_tiles;
  const int ks = b;
  const int kb_per = (p.kb_total + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * kb_per;
  const int kb1 = min(p.kb_total, kb0 + kb_per);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
    const int np = p.T < kWProducers ? p.T : kWProducers;  // active producers, each arrives once per stage
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full_bar[i], np);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();  // everything above touched only shared / tensor memory

  if (nkb > 0) {
    if (warp == 0 || warp >= 6) {
      // ---- TMA producers.  A single thread issuing all (2 + 2T) loads of a k-block (plus two integer divisions
      // for the pixel coordinates) took ~0.9 us per k-block and starved the tensor core (42 % active): the taps are
      // now split over up to three single-thread producers and the coordinates advance incrementally.
      const int pj = warp == 0 ? 0 : warp - 5;  // producer index 0..2: taps t with t % 3 == pj (+ dY for pj == 0)
      if (pj < p.T) {  // converged warp; the elected lane issues (ptx::*_elect)
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        int my_taps = 0;
        uint16_t tw[3], th[3];
        int tt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int t = pj + kWProducers * k;
          tt[k] = t;
          if (t < p.T) {
            tw[k] = static_cast<uint16_t>(p.tap_w[tg * p.T + t]);
            th[k] = static_cast<uint16_t>(p.tap_h[tg * p.T + t]);
            ++my_taps;
          } else {
            tw[k] = th[k] = 0;
          }
        }
        const uint32_t my_bytes = (pj == 0 ? a_bytes : 0) + my_taps * kBTapBytes;
        int m0 = kb0 * kKB;
        int n_img = m0 / p.OHW;
        int pr = (m0 - n_img * p.OHW) / p.OW;
        int qc = m0 - n_img * p.OHW - pr * p.OW;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int cw = p.lower_w + qc * p.trav_w;
          const int ch = p.lower_h + pr * p.trav_h;
          for (int ps = 0; ps < p.npass; ++ps) {
            // piece pair of this pass: (2,0) (0,2) (1,1) (1,0) (0,1) (0,0) packed two bits each
            const int a_off = p.npass == 1 ? 0 : ((0x052 >> (2 * ps)) & 3) * p.a_lo;
            const int b_off = p.npass == 1 ? 0 : ((0x118 >> (2 * ps)) & 3) * p.b_lo;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 11);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            ptx::mbar_expect_tx_elect(leader, &full_bar[stage], my_bytes);
            if (pj == 0)
              for (int sl = 0; sl < p.a_slabs; ++sl)
                ptx::tma_load_2d_elect(leader, sa + sl * kASlabBytes, &tmDY, &full_bar[stage], a_off + ot * 128 + sl * 64, m0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (tt[k] < p.T) {
#pragma unroll
                for (int sl = 0; sl < kBSlabs; ++sl)
                  ptx::tma_load_im2col_4d_elect(leader, sb + tt[k] * kBTapBytes + sl * kBSlabBytes, &tmX, &full_bar[stage],
                                          b_off + it * BNW + sl * CB, cw, ch, n_img, tw[k], th[k]);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          // next k-block: 64 output pixels further in (image, row, column) order
          m0 += kKB;
          qc += kKB;
          while (qc >= p.OW) { qc -= p.OW; ++pr; }
          while (pr >= p.OH) { pr -= p.OH; ++n_img; }
        }
      }
    } else if (warp == 1) {
      {  // converged warp, one elected lane issues (ptx::umma_bf16_elect: keeps the operands in uniform registers)
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        // The T tap tiles of a stage are consecutive BNW-channel slabs of ONE MN-major B operand (slab pitch =
        // kBSlabBytes), and their accumulators are consecutive TMEM columns: a single MMA covers up to 256 columns
        // = several taps.  Issuing one N = BNW MMA per tap made the 32/64-channel layers MMA-issue-bound
        // (36 tiny MMAs per k-block from one thread).
        const int ntot = p.T * BNW;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nkb * p.npass; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase, p.err, 12);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k) {
            // MN-major: LBO = distance between channel slabs, SBO = 8 pixel rows.  With a single dY slab (Cout <= 64)
            // the upper 64 rows of the M = 128 operand alias the lower ones (LBO 0): their products are never stored,
            // and the operand must not reach past the stage (the A region is a_slabs wide).
            const uint64_t adesc = ptx::make_smem_desc(sa + k * 16 * 128, p.a_slabs == 2 ? kASlabBytes : 0, 8 * 128, 2);
#pragma unroll 1
            for (int n0 = 0; n0 < ntot; n0 += 256) {
              const int n = ntot - n0 < 256 ? ntot - n0 : 256;
              const uint64_t bdesc = ptx::make_smem_desc(sb + (n0 / CB) * kBSlabBytes + k * 16 * kBRow, kBSlabBytes,
                                                         8 * kBRow, kBLayout);
              ptx::umma_bf16_elect(leader, tmem_base + n0, adesc, bdesc, ptx::make_idesc_bf16(128, n, 1, 1),
                                   (kb | k) != 0);
            }
          }
          ptx::umma_commit_elect(leader, &empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_elect(leader, done_bar);
      }
      __syncwarp();
    } else {
      const int quarter = warp & 3;
      constexpr int RB = kChunk * 4;  // bytes per staged row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
      constexpr int NJ = RB / 16;
      // Staging: once done_bar has fired every TMA load has landed and every MMA has read its operands, so the whole
      // pipeline ring is free -- each epilogue warp takes a quarter of it as a ring of 4 KB tiles and keeps all its
      // reduce-adds in flight (with the single 4 KB tile of the first version every chunk waited ~0.4 us for the
      // previous reduce to finish reading it: ~5 us of non-overlapped tail per CTA).
      const int ring_slots = max(1, (p.stages * stage_bytes / 4) / 4096);
      uint8_t* const ring = (p.stages * stage_bytes / 4 >= 4096) ? smem + (warp - 2) * ring_slots * 4096
                                                                 : s_stage + (warp - 2) * 4096;
      int slot = 0;
      const int sw = RB == 128 ? (lane & 7) : ((lane >> 1) & 3);
      const int o_row = ot * 128 + quarter * 32;
      ptx::mbar_wait(done_bar, 0, p.err, 13);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      bool pending = false;
      for (int t = 0; t < p.T; ++t) {
        const int tap = tg * p.T + t;
#pragma unroll 1
        for (int c0 = 0; c0 < BNW; c0 += kChunk) {
          float v[kChunk];
          if constexpr (kChunk == 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_row + t * BNW + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {
            uint32_t r[16];
            ptx::tmem_ld_32x16(t_row + t * BNW + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          const int i0 = it * BNW + c0;
          if (o_row < p.Cout && i0 < p.Cin) {  // warp-uniform
            if (pending && slot == 0) {  // ring wrapped: the earlier reduces must have finished reading their tiles
              if (lane == 0) ptx::tma_store_wait_read<0>();
              __syncwarp();
            }
            uint8_t* stg = ring + slot * 4096;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
              *reinterpret_cast<float4*>(stg + lane * RB + ((j ^ sw) * 16)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            // rows >= Cout are clipped; issued + committed by lane 0 of the converged warp
            ptx::tma_reduce_add_2d_elect(lane == 0 ? 1u : 0u, &tmDW, stg, tap * p.Ipad + i0, o_row);
            pending = true;
            if (++slot == ring_slots) slot = 0;
          }
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

template <int CB, int BNW, int kKB>
int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmDW, WgradParams& p,
                 cudaStream_t stream) {
  constexpr int kBTapBytes = (BNW / CB) * kKB * CB * 2;
  const int stage_bytes = p.a_slabs * kKB * 128 + p.T * kBTapBytes;
  p.stages = std::min(kMaxWStages, (204 * 1024) / stage_bytes);
  if (p.stages < 2) return set_error(B200CV_ERR_ARG, "wgrad: stage too large (%d bytes)", stage_bytes);
  // align slack | stage ring | barriers + tmem slot | align slack + 4 staging tiles of 4 KB
  const int smem = 1024 + p.stages * stage_bytes + (2 * kMaxWStages + 1) * 8 + 32 + 1024 + 4 * 4096;
  auto kern = wgrad_kernel<CB, BNW, kKB>;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "wgrad smem attr: %s", cudaGetErrorString(e));
    configured_smem = 227 * 1024;
  }
  const int grid = p.num_o_tiles * p.num_i_tiles * p.num_tap_groups * p.ksplit;
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kWThreads), (size_t)smem, stream, tmDY, tmX, tmDW, p);
  if (le != cudaSuccess) return set_error(static_cast<int>(le), "wgrad launch: %s", cudaGetErrorString(le));
  return check_launch("wgrad_kernel");
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_conv_wgrad(const void* x, const void* dy, float* dw_packed, int N, int H, int W,
                                 int Cin, int Cout, int dy_ld, int R, int S, int stride, int pad, int dil,
                                 int64_t x_lo, int64_t dy_lo, void* stream) {
  B200CV_CHECK_ARG(x && dy && dw_packed, "conv_wgrad: null pointer");
  B200CV_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cout > 0, "conv_wgrad: empty shape");
  B200CV_CHECK_ARG(Cin == pad_channels(Cin), "conv_wgrad: Cin=%d is not a padded channel count", Cin);
  B200CV_CHECK_ARG(dy_ld >= Cout && dy_ld % 8 == 0, "conv_wgrad: dy_ld=%d must be >= Cout and a multiple of 8",
                   dy_ld);
  B200CV_CHECK_ARG(R * S <= kMaxTaps && stride >= 1 && stride <= 8 && dil >= 1 && pad >= 0,
                   "conv_wgrad: unsupported filter");
  B200CV_CHECK_ARG((x_lo == 0) == (dy_lo == 0) && (x_lo == 0 || (x_lo == Cin && dy_lo % 8 == 0 && dy_ld >= (kSplitPieces - 1) * dy_lo + Cout)),
                   "conv_wgrad: fp32-parity mode needs both operands split (x_lo == Cin, dy_lo + Cout <= dy_ld)");
  const int xmul = x_lo ? kSplitPieces : 1;  // stored channels per logical channel of x
  const int OH = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int OW = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  B200CV_CHECK_ARG(OH > 0 && OW > 0, "conv_wgrad: empty output");

  WgradParams p{};
  p.OHW = OH * OW;
  p.OW = OW;
  p.OH = OH;
  p.M_pix = N * p.OHW;
  p.lower_w = -pad;
  p.lower_h = -pad;
  p.trav_w = stride;
  p.trav_h = stride;
  p.Cout = Cout;
  p.Cin = Cin;
  p.RS = R * S;
  p.Ipad = Cin;
  p.dw = dw_packed;
  p.err = device_error_word();
  p.npass = x_lo ? 6 : 1;
  p.a_lo = (int)dy_lo;
  p.b_lo = (int)x_lo;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      p.tap_w[r * S + s] = (short)(s * dil);
      p.tap_h[r * S + s] = (short)(r * dil);
    }
  const int cb = Cin < 64 ? Cin : 64;
  const int bnw = Cin < 64 ? Cin : (Cin % 128 == 0 ? 128 : 64);
  p.num_o_tiles = (Cout + 127) / 128;
  p.num_i_tiles = Cin / bnw;
  p.a_slabs = (Cout > 64 && (dy_lo ? dy_lo : dy_ld) > 64) ? 2 : 1;
  // taps per CTA: the largest divisor of RS whose accumulators fit the 512 TMEM columns
  int T = 1;
  for (int t = 1; t <= p.RS; ++t)
    if (p.RS % t == 0 && t * bnw <= 512 && t <= 9) T = t;
  if (bnw == 128 && T > 3) T = 3;  // keep >= 3 pipeline stages in shared memory
  p.T = T;
  p.num_tap_groups = p.RS / T;
  // pixels per k-block: 64 with several taps per CTA; single-tap layers take bigger blocks (see wgrad_kernel)
  static const bool no_big_kb = getenv("B200CV_WGRAD_KB64") != nullptr;
  const int kKB = (p.RS == 1 && !no_big_kb) ? (cb == 64 ? 128 : 256) : 64;
  p.kb_total = (p.M_pix + kKB - 1) / kKB;
  const int base_ctas = p.num_o_tiles * p.num_i_tiles * p.num_tap_groups;
  // split-K so that the grid fills (at most) two full waves of one CTA per SM: rounding the split UP left a
  // third, nearly empty wave (e.g. 312 CTAs on 148 SMs)
  int ksplit = std::max(1, (2 * sm_count()) / base_ctas);
  // short K ranges (1x1 layers at 13^2 / 26^2): one wave of longer CTAs halves the per-CTA fixed cost
  // (prologue, pipeline fill, epilogue) that dominates them
  static const int one_wave_below = getenv("B200CV_WGRAD_1WAVE") ? atoi(getenv("B200CV_WGRAD_1WAVE")) : 24;
  if (p.kb_total / ksplit < one_wave_below && base_ctas <= sm_count()) ksplit = std::max(1, sm_count() / base_ctas);
  ksplit = std::min(ksplit, std::max(1, p.kb_total / 4));
  p.ksplit = ksplit;

  CUtensorMap tmDY, tmX;
  int rc = make_tmap_2d_bf16(&tmDY, dy, p.M_pix, dy_ld, dy_ld, kKB, 64);
  if (rc) return rc;
  rc = make_tmap_im2col_bf16(&tmX, x, N, H, W, xmul * Cin, xmul * Cin, (int64_t)W * xmul * Cin,
                             (int64_t)H * W * xmul * Cin, -pad, -pad, pad - (S - 1) * dil, pad - (R - 1) * dil, stride,
                             stride, cb, kKB);
  if (rc) return rc;
  CUtensorMap tmDW;
  rc = make_tmap_2d_f32(&tmDW, dw_packed, Cout, (int64_t)p.RS * p.Ipad, (int64_t)p.RS * p.Ipad, 32,
                        bnw >= 32 ? 32 : 16);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kKB == 64) {
    if (cb == 64 && bnw == 128) return launch_wgrad<64, 128, 64>(tmDY, tmX, tmDW, p, st);
    if (cb == 64 && bnw == 64) return launch_wgrad<64, 64, 64>(tmDY, tmX, tmDW, p, st);
    if (cb == 32) return launch_wgrad<32, 32, 64>(tmDY, tmX, tmDW, p, st);
    if (cb == 16) return launch_wgrad<16, 16, 64>(tmDY, tmX, tmDW, p, st);
  } else {
    if (cb == 64 && bnw == 128) return launch_wgrad<64, 128, 128>(tmDY, tmX, tmDW, p, st);
    if (cb == 64 && bnw == 64) return launch_wgrad<64, 64, 128>(tmDY, tmX, tmDW, p, st);
    if (cb == 32) return launch_wgrad<32, 32, 256>(tmDY, tmX, tmDW, p, st);
    if (cb == 16) return launch_wgrad<16, 16, 256>(tmDY, tmX, tmDW, p, st);
  }
  return set_error(B200CV_ERR_ARG, "conv_wgrad: unsupported channel tile");
}
