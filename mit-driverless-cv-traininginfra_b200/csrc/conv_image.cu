// The image layers (CVC-YOLOv3/models.py:59-69 conv_0: 3 -> 32, 3x3; RektNet/keypoint_net.py:17 stem: 3 -> 16, 7x7):
// k x k, stride 1, "same" padding, 3-channel fp32 NCHW input, NHWC bf16 output.
//
// 0.9 % of a Darknet-53 step's FLOPs but 1.0 ms of it as "explicit im2col + 1x1 implicit GEMM": the [M][32] bf16
// patch matrix (709 MB at 416^2 bs64) was written once and read twice (forward, weight gradient).  Here the patch
// matrix never exists: a CTA stages the (8 + k - 1) x (32 + k - 1) x 3 halo of an 8 x 32-pixel tile in shared
// memory as bf16 and every warp gathers its mma.sync A (forward) / B (weight gradient) fragments from it.  K = 27
// or 147 is far too small for a tcgen05 pipeline to pay (one 128 x 32 x 32 MMA per tile); with warp-level
// mma.m16n8k16 the tensor work is a few percent of the kernel.  What the kernels move is 709 MB written (forward) or
// read (weight gradient) per 0.27 / 0.24 ms: instruction-issue bound at about half of the HBM rate (fragment
// gathers are 16-bit shared-memory loads), 2x faster than the three launches they replace.
//
// Numerics are those of the implicit-GEMM path: image values and weights rounded to bf16, fp32 accumulation,
// bf16 output, BatchNorm statistics of the STORED values accumulated per CTA and added once (stat_acc.cuh).
#include <algorithm>

#include "internal.h"
#include "stat_acc.cuh"

namespace b200cv {
namespace {

constexpr int kThreads = 256;  // 8 warps: warp w owns row w of the tile
constexpr int kTileH = 8;
constexpr int kTileW = 32;     // two m16 pixel groups per warp

template <int R, int COUT>
struct ImgCfg {
  static constexpr int kTaps = R * R;
  static constexpr int kK = 3 * kTaps;             // 27 / 147
  static constexpr int kSteps = (kK + 15) / 16;    // mma k-steps of the forward GEMM
  static constexpr int kNT = COUT / 8;             // n8 tiles of the forward GEMM
  static constexpr int kNT8 = (kK + 7) / 8;        // n8 tiles of the weight-gradient GEMM (N = K of the filter)
  static constexpr int kMT = COUT / 16;            // m16 tiles of the weight-gradient GEMM
  static constexpr int kHH = kTileH + R - 1;
  static constexpr int kHW = kTileW + R - 1;
  static constexpr int kHWp = kHW + 2;             // row pitch (elements)
  static constexpr int kPlane = kHH * kHWp + 4;    // plane pitch (elements): planes land on different banks
  static constexpr int kHaloElems = 3 * kPlane;
  static constexpr int kOutPitch = COUT + 8;       // staged output / dy rows: 16-byte multiples, conflict-free
  static constexpr int kKoffN = kSteps * 16;
  static_assert(COUT == 16 || COUT == 32, "image layers have 16 or 32 output channels");
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack2(const __nv_bfloat16* p0, const __nv_bfloat16* p1) {
  const uint32_t lo = *reinterpret_cast<const unsigned short*>(p0);
  const uint32_t hi = *reinterpret_cast<const unsigned short*>(p1);
  return lo | (hi << 16);
}

// halo tile of (n, y0, x0): bf16, planar [c][row][col], zero outside the image.  The fp32 image values of the NEXT
// tile are fetched into registers while the current tile is computed (halo_fetch) and converted / stored at the top of
// the next iteration (halo_commit): with 16 resident warps per SM the loads would otherwise be fully exposed.
template <class C>
struct HaloRegs {
  static constexpr int kN = (3 * C::kHH * C::kHW + kThreads - 1) / kThreads;
  float v[kN];
  // per-thread constants (the same halo elements every tile)
  int rc[kN];    // (row << 16) | col inside the halo, or -1 past its end
  int goff[kN];  // (c*H + row)*W + col: global offset relative to the tile's top-left halo element
  int soff[kN];  // element offset in the shared-memory halo

  __device__ __forceinline__ void init(int H, int W) {
#pragma unroll
    for (int i = 0; i < kN; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      const int col = idx % C::kHW;
      const int q = idx / C::kHW;
      const int row = q % C::kHH;
      const int c = q / C::kHH;
      rc[i] = c < 3 ? ((row << 16) | col) : -1;
      goff[i] = (c * H + row) * W + col;
      soff[i] = c * C::kPlane + row * C::kHWp + col;
    }
  }
};

// Tile walk without divisions: (n, ty, tx) advanced by the grid size in mixed radix.
struct TileWalk {
  int n, ty, tx;
  int dn, dty, dtx, tiles_x, tiles_y;
  __device__ __forceinline__ void start(int first, int step, int tx_count, int ty_count) {
    tiles_x = tx_count; tiles_y = ty_count;
    tx = first % tiles_x; ty = (first / tiles_x) % tiles_y; n = first / (tiles_x * tiles_y);
    dtx = step % tiles_x; dty = (step / tiles_x) % tiles_y; dn = step / (tiles_x * tiles_y);
  }
  __device__ __forceinline__ void advance() {
    tx += dtx;
    int carry = tx >= tiles_x;
    tx -= carry ? tiles_x : 0;
    ty += dty + carry;
    carry = ty >= tiles_y;
    ty -= carry ? tiles_y : 0;
    n += dn + carry;
  }
};

// the fp32 image values of tile `t` -> registers (32-bit offsets: the entry points check 3*H*W < 2^31)
template <class C, int R>
__device__ __forceinline__ void halo_fetch(HaloRegs<C>& h, const float* __restrict__ x, const TileWalk& t, int N, int H,
                                           int W) {
  constexpr int pad = (R - 1) / 2;
  if (t.n >= N) return;
  const int gy0 = t.ty * kTileH - pad, gx0 = t.tx * kTileW - pad;
  const int base = t.n * 3 * H * W + gy0 * W + gx0;  // 32-bit: the entry points check N*3*H*W < 2^31
#pragma unroll
  for (int i = 0; i < HaloRegs<C>::kN; ++i) {
    const int r = h.rc[i];
    const int gy = gy0 + (r >> 16), gx = gx0 + (r & 0xffff);
    const bool ok = r >= 0 && static_cast<unsigned>(gy) < static_cast<unsigned>(H) &&
                    static_cast<unsigned>(gx) < static_cast<unsigned>(W);
    h.v[i] = ok ? __ldg(x + (base + h.goff[i])) : 0.f;
  }
}

template <class C>
__device__ __forceinline__ void halo_commit(const HaloRegs<C>& h, __nv_bfloat16* halo) {
#pragma unroll
  for (int i = 0; i < HaloRegs<C>::kN; ++i)
    if (h.rc[i] >= 0) halo[h.soff[i]] = __float2bfloat16_rn(h.v[i]);
}

// offset of filter element k = tap*3 + c inside the halo, relative to the output pixel's top-left halo element
template <class C, int R>
__device__ __forceinline__ void fill_koff(int* koff) {
  for (int k = threadIdx.x; k < C::kKoffN; k += kThreads) {
    int off = 0;
    if (k < C::kK) {
      const int tap = k / 3, c = k - 3 * tap;
      off = c * C::kPlane + (tap / R) * C::kHWp + (tap % R);
    }
    koff[k] = off;  // padded k: any valid element (its weight is zero)
  }
}

// ------------------------------------------------------------------------------------------------ forward
// kStats: training pass (raw conv output + BatchNorm statistics); otherwise the inference pass with the folded
// per-channel affine and the activation in the epilogue.
template <int R, int COUT, bool kStats>
__global__ void __launch_bounds__(kThreads, R == 3 ? 3 : 2)
conv_image_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wflat, int Kp, int N, int H,
                      int W, __nv_bfloat16* __restrict__ y, long long y_ld, const float* __restrict__ scale,
                      const float* __restrict__ shift, int act, float slope, StatAcc* stats, int stats_parts) {
  using C = ImgCfg<R, COUT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* outs = halo + ((C::kHaloElems + 7) & ~7);            // [256 pixels][kOutPitch]
  int* koff = reinterpret_cast<int*>(outs + kTileH * kTileW * C::kOutPitch);
  float* red = reinterpret_cast<float*>(koff + C::kKoffN);             // [8 warps][2 * COUT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  fill_koff<C, R>(koff);

  // B fragments (weights) stay in registers for the CTA lifetime: B[k][n] = wflat[n][k]
  uint32_t bf[C::kSteps][C::kNT][2];
#pragma unroll
  for (int ks = 0; ks < C::kSteps; ++ks)
#pragma unroll
    for (int nt = 0; nt < C::kNT; ++nt) {
      const __nv_bfloat16* wp = wflat + (long long)(nt * 8 + g) * Kp + ks * 16 + 2 * t;
      bf[ks][nt][0] = *reinterpret_cast<const uint32_t*>(wp);
      bf[ks][nt][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
    }
  float sc[kStats ? 1 : C::kNT][2], sh[kStats ? 1 : C::kNT][2];
  if constexpr (!kStats) {
#pragma unroll
    for (int nt = 0; nt < C::kNT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + 2 * t + e;
        sc[nt][e] = scale ? __ldg(scale + col) : 1.f;
        sh[nt][e] = shift ? __ldg(shift + col) : 0.f;
      }
  }
  const float neg = act == 1 ? slope : (act == 2 ? 0.f : 1.f);
  float ssum[kStats ? C::kNT : 1][2], ssq[kStats ? C::kNT : 1][2];
#pragma unroll
  for (int nt = 0; nt < (kStats ? C::kNT : 1); ++nt) ssum[nt][0] = ssum[nt][1] = ssq[nt][0] = ssq[nt][1] = 0.f;
  // halo offsets of this lane's filter elements: registers for the 3x3 layer, shared memory for the 7x7 stem
  constexpr bool kKoffRegs = C::kSteps <= 2;
  __syncthreads();  // koff complete
  int ko[kKoffRegs ? C::kSteps : 1][4];
  if constexpr (kKoffRegs) {
#pragma unroll
    for (int ks = 0; ks < C::kSteps; ++ks) {
      const int k0 = ks * 16 + 2 * t;
      ko[ks][0] = koff[k0]; ko[ks][1] = koff[k0 + 1]; ko[ks][2] = koff[k0 + 8]; ko[ks][3] = koff[k0 + 9];
    }
  }

  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  HaloRegs<C> pre;
  pre.init(H, W);
  TileWalk cur, nxt;
  cur.start(blockIdx.x, gridDim.x, tiles_x, tiles_y);
  halo_fetch<C, R>(pre, x, cur, N, H, W);
  for (; cur.n < N; cur = nxt) {
    const int n = cur.n, y0 = cur.ty * kTileH, x0 = cur.tx * kTileW;
    nxt = cur;
    nxt.advance();
    halo_commit<C>(pre, halo);  // every warp passed the barrier after its last halo read of the previous tile
    __syncthreads();            // halo complete; the previous tile's staged outputs have been stored
    halo_fetch<C, R>(pre, x, nxt, N, H, W);
    const bool row_ok = y0 + warp < H;
    // both 16-pixel groups of this warp's row: all fragment gathers are independent (latency, not issue, bound)
    float acc[2][C::kNT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < C::kNT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    const __nv_bfloat16* p0 = halo + warp * C::kHWp + g;  // pixel g of group 0; +8: pixel g + 8; +16: group 1
#pragma unroll
    for (int ks = 0; ks < C::kSteps; ++ks) {
      int o0, o1, o2, o3;
      if constexpr (kKoffRegs) {
        o0 = ko[ks][0]; o1 = ko[ks][1]; o2 = ko[ks][2]; o3 = ko[ks][3];
      } else {
        const int k0 = ks * 16 + 2 * t;
        o0 = koff[k0]; o1 = koff[k0 + 1]; o2 = koff[k0 + 8]; o3 = koff[k0 + 9];
      }
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const __nv_bfloat16* pm = p0 + mt * 16;
        a[mt][0] = pack2(pm + o0, pm + o1);
        a[mt][1] = pack2(pm + 8 + o0, pm + 8 + o1);
        a[mt][2] = pack2(pm + o2, pm + o3);
        a[mt][3] = pack2(pm + 8 + o2, pm + 8 + o3);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < C::kNT; ++nt) mma_bf16(acc[mt][nt], a[mt], bf[ks][nt][0], bf[ks][nt][1]);
    }
    // epilogue: affine / activation (inference) or statistics of the stored values (training); bf16 staged rows
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = mt * 16 + g + 8 * half;
        const bool ok = row_ok && x0 + px < W;
        __nv_bfloat16* orow = outs + (warp * kTileW + px) * C::kOutPitch;
#pragma unroll
        for (int nt = 0; nt < C::kNT; ++nt) {
          float v0 = acc[mt][nt][2 * half], v1 = acc[mt][nt][2 * half + 1];
          if constexpr (!kStats) {
            v0 = v0 * sc[nt][0] + sh[nt][0];
            v1 = v1 * sc[nt][1] + sh[nt][1];
            v0 = v0 > 0.f ? v0 : v0 * neg;
            v1 = v1 > 0.f ? v1 : v1 * neg;
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          *reinterpret_cast<__nv_bfloat162*>(orow + nt * 8 + 2 * t) = h;
          if constexpr (kStats) {
            if (ok) {
              const float2 f = __bfloat1622float2(h);
              ssum[nt][0] += f.x; ssq[nt][0] += f.x * f.x;
              ssum[nt][1] += f.y; ssq[nt][1] += f.y * f.y;
            }
          }
        }
      }
    __syncthreads();
    // coalesced stores: a tile row is kTileW * COUT contiguous bf16 in global memory
    constexpr int kVecPerPix = COUT / 8;
    __nv_bfloat16* ytile = y + (((long long)n * H + y0) * W + x0) * y_ld;
    const int ld32 = static_cast<int>(y_ld);  // offsets inside a tile fit 32 bits
#pragma unroll
    for (int i = 0; i < kTileH * kTileW * kVecPerPix / kThreads; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      const int v = idx % kVecPerPix;
      const int pix = idx / kVecPerPix;
      const int px = pix % kTileW, row = pix / kTileW;
      if (y0 + row < H && x0 + px < W) {
        const uint4 q = *reinterpret_cast<const uint4*>(outs + pix * C::kOutPitch + v * 8);
        *reinterpret_cast<uint4*>(ytile + ((row * W + px) * ld32 + v * 8)) = q;
      }
    }
  }
  if constexpr (kStats) {
    // lanes with the same t hold the same columns: fold over g, then over the warps in warp order
#pragma unroll
    for (int nt = 0; nt < C::kNT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) {
          ssum[nt][e] += __shfl_xor_sync(0xffffffffu, ssum[nt][e], m);
          ssq[nt][e] += __shfl_xor_sync(0xffffffffu, ssq[nt][e], m);
        }
        if (g == 0) {
          red[warp * 2 * COUT + nt * 8 + 2 * t + e] = ssum[nt][e];
          red[warp * 2 * COUT + COUT + nt * 8 + 2 * t + e] = ssq[nt][e];
        }
      }
    __syncthreads();
    if (threadIdx.x < 2 * COUT) {
      float s = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) s += red[w * 2 * COUT + threadIdx.x];
      StatAcc* row = stats + (long long)(blockIdx.x % stats_parts) * 2 * COUT;
      stat_add(row + threadIdx.x, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[o][k] = sum over pixels dy[pixel][o] * patch[pixel][k]: M = COUT, N = K of the filter, K = pixels.
// kBn: `dy` is dL/da of the layer's activation and the BatchNorm + activation backward is applied on the fly
// (dz = da * act'(y*scale+shift), dy = g*dz + A*y + B with the per-channel constants of b200cv_bn_bwd_stats_apply,
// rounded to bf16 like the stored dy of the two-pass form): the image layers have no data gradient, so dy has no
// other reader and the 709 MB tensor (416^2 bs64) is never written.
struct ImgBnBwd {
  const __nv_bfloat16* y;   // conv output (BN input), laid out like dy
  long long y_ld;
  const float* scale;       // forward affine of the BN: z = y*scale + shift
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const StatAcc* partials;  // [nparts][2*COUT]: sum dz | sum dz*xhat
  int nparts;
  float count;
  float neg;                // act'(z) for z <= 0
  float* coef;              // [3*COUT] g | k1 | k2 (or null), written by block 0 like bn_bwd_stats_apply
  float* dgamma;
  float* dbeta;
};

template <int R, int COUT, bool kBn>
__global__ void __launch_bounds__(kThreads, R == 3 ? (kBn ? 2 : 3) : 1)
conv_image_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, long long dy_ld, int N,
                        int H, int W, float* __restrict__ dw, int Kp, const ImgBnBwd bn) {
  using C = ImgCfg<R, COUT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* dys = halo + ((C::kHaloElems + 7) & ~7);             // [256 pixels][kOutPitch]
  int* koff = reinterpret_cast<int*>(dys + kTileH * kTileW * C::kOutPitch);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  fill_koff<C, R>(koff);
  // per-thread BatchNorm-backward constants of the 8 channels this thread converts (vector v of every pixel)
  float bsc[kBn ? 8 : 1], bsh[kBn ? 8 : 1], bg[kBn ? 8 : 1], bA[kBn ? 8 : 1], bB[kBn ? 8 : 1];
  if constexpr (kBn) {
    const int c0 = (threadIdx.x % (COUT / 8)) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float s1 = stat_fold(bn.partials, bn.nparts, 2 * COUT, c);
      const float s2 = stat_fold(bn.partials, bn.nparts, 2 * COUT, COUT + c);
      const float rstd = __ldg(bn.rstd + c), mean = __ldg(bn.mean + c);
      const float gg = __ldg(bn.gamma + c) * rstd;
      const float k1 = s1 / bn.count, k2 = s2 / bn.count;
      bg[j] = gg;
      bA[j] = -gg * k2 * rstd;
      bB[j] = -gg * k1 - bA[j] * mean;
      bsc[j] = __ldg(bn.scale + c);
      bsh[j] = __ldg(bn.shift + c);
      if (blockIdx.x == 0 && threadIdx.x < COUT / 8) {
        if (bn.coef) {
          bn.coef[c] = gg;
          bn.coef[COUT + c] = k1;
          bn.coef[2 * COUT + c] = k2;
        }
        if (bn.dbeta) bn.dbeta[c] = s1;
        if (bn.dgamma) bn.dgamma[c] = s2;
      }
    }
  }
  float acc[C::kMT][C::kNT8][4];
#pragma unroll
  for (int mt = 0; mt < C::kMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < C::kNT8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  constexpr int kVecPerPix = COUT / 8;
  constexpr int kDyVecs = kTileH * kTileW * kVecPerPix / kThreads;  // 16-byte dy vectors per thread and tile
  HaloRegs<C> pre;
  pre.init(H, W);
  uint4 dpre[kDyVecs], ypre[kBn ? kDyVecs : 1];
  auto dy_fetch = [&](const TileWalk& t) {
    if (t.n >= N) return;
    const long long tile0 = ((long long)t.n * H + t.ty * kTileH) * W + t.tx * kTileW;
    const __nv_bfloat16* dtile = dy + tile0 * dy_ld;
    const __nv_bfloat16* ytile = kBn ? bn.y + tile0 * bn.y_ld : nullptr;
#pragma unroll
    for (int i = 0; i < kDyVecs; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      const int v = idx % kVecPerPix;
      const int pix = idx / kVecPerPix;
      const int px = pix % kTileW, row = pix / kTileW;
      dpre[i] = make_uint4(0u, 0u, 0u, 0u);  // pixels outside the image contribute nothing
      if constexpr (kBn) ypre[i] = make_uint4(0u, 0u, 0u, 0u);
      if (t.ty * kTileH + row < H && t.tx * kTileW + px < W) {
        dpre[i] = __ldg(reinterpret_cast<const uint4*>(dtile + ((row * W + px) * static_cast<int>(dy_ld) + v * 8)));
        if constexpr (kBn)
          ypre[i] = __ldg(reinterpret_cast<const uint4*>(ytile + ((row * W + px) * static_cast<int>(bn.y_ld) + v * 8)));
      } else if constexpr (kBn) {
        dpre[i].x = 0x7fc07fc0u;  // marks an outside pixel (bf16 NaN pair): its dy must be exactly zero
      }
    }
  };
  TileWalk cur, nxt;
  cur.start(blockIdx.x, gridDim.x, tiles_x, tiles_y);
  halo_fetch<C, R>(pre, x, cur, N, H, W);
  dy_fetch(cur);
  for (; cur.n < N; cur = nxt) {
    nxt = cur;
    nxt.advance();
    __syncthreads();  // every warp is done with the previous tile's buffers
    halo_commit<C>(pre, halo);
#pragma unroll
    for (int i = 0; i < kDyVecs; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      uint4 q = dpre[i];
      if constexpr (kBn) {
        if (q.x == 0x7fc07fc0u) {
          q = make_uint4(0u, 0u, 0u, 0u);
        } else {
          const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&dpre[i]);
          const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&ypre[i]);
          __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 g2 = __bfloat1622float2(hg[e]);
            const float2 y2 = __bfloat1622float2(hy[e]);
            const float dz0 = g2.x * (fmaf(y2.x, bsc[2 * e], bsh[2 * e]) > 0.f ? 1.f : bn.neg);
            const float dz1 = g2.y * (fmaf(y2.y, bsc[2 * e + 1], bsh[2 * e + 1]) > 0.f ? 1.f : bn.neg);
            ho[e] = __floats2bfloat162_rn(fmaf(bg[2 * e], dz0, fmaf(bA[2 * e], y2.x, bB[2 * e])),
                                          fmaf(bg[2 * e + 1], dz1, fmaf(bA[2 * e + 1], y2.y, bB[2 * e + 1])));
          }
        }
      }
      *reinterpret_cast<uint4*>(dys + (idx / kVecPerPix) * C::kOutPitch + (idx % kVecPerPix) * 8) = q;
    }
    __syncthreads();
    halo_fetch<C, R>(pre, x, nxt, N, H, W);
    dy_fetch(nxt);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {  // 16 pixels of this warp's row per k-step
      // A = dy^T through ldmatrix.trans: stored rows are pixels (k), 8 channels (m) per 16-byte row
      uint32_t a[C::kMT][4];
#pragma unroll
      for (int mt = 0; mt < C::kMT; ++mt) {
        const int j = lane >> 3, r = lane & 7;
        const __nv_bfloat16* src =
            dys + (warp * kTileW + ks * 16 + (j >> 1) * 8 + r) * C::kOutPitch + mt * 16 + (j & 1) * 8;
        const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(src));
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                     : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3])
                     : "r"(addr));
      }
      // B[k = pixel][n = filter element]: pixels 2t, 2t+1 (and + 8) of this group, filter element nt*8 + g
      const __nv_bfloat16* p0 = halo + warp * C::kHWp + ks * 16 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < C::kNT8; ++nt) {
        const int kk = nt * 8 + g;
        uint32_t b0 = 0u, b1 = 0u;
        if (kk < C::kK) {
          const int o = koff[kk];
          b0 = pack2(p0 + o, p0 + o + 1);
          b1 = pack2(p0 + o + 8, p0 + o + 9);
        }
#pragma unroll
        for (int mt = 0; mt < C::kMT; ++mt) mma_bf16(acc[mt][nt], a[mt], b0, b1);
      }
    }
  }
  // CTA reduction in shared memory (warp order), then one fp32 add per element and CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);  // [COUT][kNT8 * 8]: the tile buffers are free now
  constexpr int kCols = C::kNT8 * 8;
  for (int i = threadIdx.x; i < COUT * kCols; i += kThreads) red[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < kThreads / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < C::kMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < C::kNT8; ++nt) {
          float* r0 = red + (mt * 16 + g) * kCols + nt * 8 + 2 * t;
          r0[0] += acc[mt][nt][0];
          r0[1] += acc[mt][nt][1];
          r0[8 * kCols] += acc[mt][nt][2];
          r0[8 * kCols + 1] += acc[mt][nt][3];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < COUT * kCols; i += kThreads) {
    const int o = i / kCols, k = i - o * kCols;
    if (k < C::kK && red[i] != 0.f) atomicAdd(dw + (long long)o * Kp + k, red[i]);
  }
}

template <int R, int COUT>
size_t image_smem_bytes() {
  using C = ImgCfg<R, COUT>;
  const size_t tile = (size_t)((C::kHaloElems + 7) & ~7) * 2 + (size_t)kTileH * kTileW * C::kOutPitch * 2 +
                      (size_t)C::kKoffN * 4 + (size_t)(kThreads / 32) * 2 * COUT * 4;
  const size_t red = (size_t)COUT * C::kNT8 * 8 * 4;
  return std::max(tile, red);
}

int image_grid(int N, int H, int W, int per_sm) {
  const long long tiles = (long long)N * ((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW);
  return (int)std::min<long long>(tiles, (long long)sm_count() * per_sm);
}

template <int R, int COUT, bool kStats>
int launch_fwd(const float* x, const void* w, int Kp, int N, int H, int W, void* y, long long y_ld,
               const float* scale, const float* shift, int act, float slope, void* stats, int parts,
               cudaStream_t st) {
  auto kern = conv_image_fwd_kernel<R, COUT, kStats>;
  const size_t smem = image_smem_bytes<R, COUT>();
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "conv_image_fwd smem attr: %s", cudaGetErrorString(e));
  }
  kern<<<image_grid(N, H, W, R == 3 ? 3 : 2), kThreads, smem, st>>>(x, static_cast<const __nv_bfloat16*>(w), Kp, N, H, W,
                                                       static_cast<__nv_bfloat16*>(y), y_ld, scale, shift, act, slope,
                                                       static_cast<StatAcc*>(stats), parts > 0 ? parts : 1);
  return check_launch("conv_image_fwd");
}

template <int R, int COUT, bool kBn>
int launch_wgrad(const float* x, const void* dy, long long dy_ld, int N, int H, int W, float* dw, int Kp,
                 const ImgBnBwd& bn, cudaStream_t st) {
  auto kern = conv_image_wgrad_kernel<R, COUT, kBn>;
  const size_t smem = image_smem_bytes<R, COUT>();
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "conv_image_wgrad smem attr: %s", cudaGetErrorString(e));
  }
  kern<<<image_grid(N, H, W, kBn ? 2 : 3), kThreads, smem, st>>>(x, static_cast<const __nv_bfloat16*>(dy), dy_ld, N, H,
                                                                 W, dw, Kp, bn);
  return check_launch("conv_image_wgrad");
}

bool image_shape_ok(int C, int R, int S, int pad, int dil, int Cout) {
  return C == 3 && R == S && (R == 3 || R == 7) && dil == 1 && pad == (R - 1) / 2 && (Cout == 16 || Cout == 32);
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_conv_image_supported(int C, int R, int S, int stride, int pad, int dil, int Cout) {
  return stride == 1 && image_shape_ok(C, R, S, pad, dil, Cout) ? 1 : 0;
}

extern "C" int b200cv_conv_image_fwd(const float* x, const void* w_flat, int N, int C, int H, int W, int R, int S,
                                     int pad, int dil, int Cout, int Kp, void* y, int64_t y_ld, const float* scale,
                                     const float* shift, int act, float slope, void* stats, int stats_parts,
                                     void* stream) {
  B200CV_CHECK_ARG(x && w_flat && y && N > 0 && H > 0 && W > 0, "conv_image_fwd: bad args");
  B200CV_CHECK_ARG(image_shape_ok(C, R, S, pad, dil, Cout),
                   "conv_image_fwd: only 3-channel 3x3 / 7x7 stride-1 same-padding layers with 16 or 32 filters");
  B200CV_CHECK_ARG(Kp >= ((3 * R * S + 15) / 16) * 16 && Kp % 8 == 0, "conv_image_fwd: Kp=%d too small", Kp);
  B200CV_CHECK_ARG(3LL * N * H * W < (1LL << 31) && (long long)N * ((H + 7) / 8) * ((W + 31) / 32) < (1LL << 31),
                   "conv_image_fwd: image too large for 32-bit tile arithmetic");
  B200CV_CHECK_ARG(y_ld >= Cout && y_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(w_flat) & 3) == 0,
                   "conv_image_fwd: y must be 16-byte aligned rows of >= Cout bf16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200CV_CHECK_ARG(!stats || (!scale && !shift && act == 0),
                   "conv_image_fwd: statistics (training) and a folded affine / activation (inference) are exclusive");
#define B200CV_IMG_FWD(R_, CO_)                                                                                     \
  return stats ? launch_fwd<R_, CO_, true>(x, w_flat, Kp, N, H, W, y, y_ld, scale, shift, act, slope, stats,       \
                                           stats_parts, st)                                                         \
               : launch_fwd<R_, CO_, false>(x, w_flat, Kp, N, H, W, y, y_ld, scale, shift, act, slope, stats,      \
                                            stats_parts, st)
  if (R == 3 && Cout == 32) B200CV_IMG_FWD(3, 32);
  if (R == 3 && Cout == 16) B200CV_IMG_FWD(3, 16);
  if (R == 7 && Cout == 32) B200CV_IMG_FWD(7, 32);
  B200CV_IMG_FWD(7, 16);
#undef B200CV_IMG_FWD
}

namespace {
int image_wgrad_dispatch(const float* x, const void* dy, int64_t dy_ld, int N, int C, int H, int W, int R, int S, int pad,
                         int dil, int Cout, int Kp, float* dw_flat, const ImgBnBwd* bn, void* stream) {
  B200CV_CHECK_ARG(x && dy && dw_flat && N > 0 && H > 0 && W > 0, "conv_image_wgrad: bad args");
  B200CV_CHECK_ARG(image_shape_ok(C, R, S, pad, dil, Cout),
                   "conv_image_wgrad: only 3-channel 3x3 / 7x7 stride-1 same-padding layers with 16 or 32 filters");
  B200CV_CHECK_ARG(Kp >= 3 * R * S, "conv_image_wgrad: Kp=%d too small", Kp);
  B200CV_CHECK_ARG(3LL * N * H * W < (1LL << 31) && (long long)N * ((H + 7) / 8) * ((W + 31) / 32) < (1LL << 31),
                   "conv_image_wgrad: image too large for 32-bit tile arithmetic");
  B200CV_CHECK_ARG(dy_ld >= Cout && dy_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
                   "conv_image_wgrad: dy must be 16-byte aligned rows of >= Cout bf16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ImgBnBwd none{};
#define B200CV_IMG_WG(R_, CO_)                                                                       \
  return bn ? launch_wgrad<R_, CO_, true>(x, dy, dy_ld, N, H, W, dw_flat, Kp, *bn, st)               \
            : launch_wgrad<R_, CO_, false>(x, dy, dy_ld, N, H, W, dw_flat, Kp, none, st)
  if (R == 3 && Cout == 32) B200CV_IMG_WG(3, 32);
  if (R == 3 && Cout == 16) B200CV_IMG_WG(3, 16);
  if (R == 7 && Cout == 32) B200CV_IMG_WG(7, 32);
  B200CV_IMG_WG(7, 16);
#undef B200CV_IMG_WG
}
}  // namespace

extern "C" int b200cv_conv_image_wgrad(const float* x, const void* dy, int64_t dy_ld, int N, int C, int H, int W,
                                       int R, int S, int pad, int dil, int Cout, int Kp, float* dw_flat,
                                       void* stream) {
  return image_wgrad_dispatch(x, dy, dy_ld, N, C, H, W, R, S, pad, dil, Cout, Kp, dw_flat, nullptr, stream);
}

extern "C" int b200cv_conv_image_wgrad_bn(const float* x, const void* da, int64_t da_ld, const void* y, int64_t y_ld,
                                          const void* partials, int nparts, int64_t count, const float* gamma,
                                          const float* scale, const float* shift, const float* mean,
                                          const float* rstd, int act, float slope, float* coef, float* dgamma,
                                          float* dbeta, int N, int C, int H, int W, int R, int S, int pad, int dil,
                                          int Cout, int Kp, float* dw_flat, void* stream) {
  B200CV_CHECK_ARG(y && partials && nparts > 0 && count > 0 && gamma && scale && shift && mean && rstd,
                   "conv_image_wgrad_bn: incomplete BatchNorm arguments");
  B200CV_CHECK_ARG(y_ld >= Cout && y_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                   "conv_image_wgrad_bn: y must be 16-byte aligned rows of >= Cout bf16");
  ImgBnBwd bn{};
  bn.y = static_cast<const __nv_bfloat16*>(y);
  bn.y_ld = y_ld;
  bn.scale = scale; bn.shift = shift; bn.mean = mean; bn.rstd = rstd; bn.gamma = gamma;
  bn.partials = static_cast<const StatAcc*>(partials);
  bn.nparts = nparts;
  bn.count = static_cast<float>(count);
  bn.neg = act == B200CV_ACT_LEAKY ? slope : (act == B200CV_ACT_RELU ? 0.f : 1.f);
  bn.coef = coef; bn.dgamma = dgamma; bn.dbeta = dbeta;
  return image_wgrad_dispatch(x, da, da_ld, N, C, H, W, R, S, pad, dil, Cout, Kp, dw_flat, &bn, stream);
}
