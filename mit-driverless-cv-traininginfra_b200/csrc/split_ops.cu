// Element-wise kernels of the fp32-parity ("split", bf16x3) mode.
//
// An fp32 value v is stored as kSplitPieces = 3 bf16 numbers p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1)
// (3 x 8 = 24 mantissa bits: fp32's own precision); a split row is [p0(C) | p1(C) | p2(C)] with piece j lying
// j * lo elements after p0.  Every kernel here sums the pieces into fp32,
// computes exactly what its bf16 counterpart in elementwise.cu computes, and re-splits what it stores.  This mode
// exists for parity with the reference's fp32 arithmetic (CVC-YOLOv3/models.py:59-69), not for speed: the kernels
// are plain grid-stride loops, one thread per 8 channels of one row.
#include <cuda_bf16.h>

#include <algorithm>

#include "internal.h"
#include "stat_acc.cuh"

namespace b200cv {
namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld_split8(const bf16* p, long long lo, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 0.f;
#pragma unroll
  for (int pc = kSplitPieces - 1; pc >= 0; --pc) {  // smallest piece first
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + pc * lo));
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __bfloat1622float2(ha[i]);
      f[2 * i] += x.x;
      f[2 * i + 1] += x.y;
    }
  }
}
__device__ __forceinline__ void st_split8(bf16* p, long long lo, const float (&f)[8]) {
  float rem[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rem[i] = f[i];
#pragma unroll
  for (int pc = 0; pc < kSplitPieces; ++pc) {
    uint4 a;
    __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ha[i] = __floats2bfloat162_rn(rem[2 * i], rem[2 * i + 1]);
      const float2 x = __bfloat1622float2(ha[i]);
      rem[2 * i] -= x.x;
      rem[2 * i + 1] -= x.y;
    }
    *reinterpret_cast<uint4*>(p + pc * lo) = a;
  }
}
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float neg_slope(int act, float slope) {
  return act == B200CV_ACT_LEAKY ? slope : (act == B200CV_ACT_RELU ? 0.f : 1.f);
}

int grid_of(long long items) {
  return (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, (long long)sm_count() * 16));
}
bool ok_split(const void* p, long long ld, long long lo, int C) {
  return p && C > 0 && C % 8 == 0 && lo >= C && lo % 8 == 0 && ld >= (kSplitPieces - 1) * lo + C && ld % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// out = act(y*scale+shift [+ y2*scale2+shift2]) [+ post]
__global__ void split_bn_apply_kernel(const bf16* __restrict__ y, long long y_ld, const float* __restrict__ scale,
                                      const float* __restrict__ shift, const bf16* __restrict__ y2, long long y2_ld,
                                      const float* __restrict__ scale2, const float* __restrict__ shift2,
                                      const bf16* __restrict__ post, long long post_ld, bf16* __restrict__ out,
                                      long long out_ld, long long rows, int C, long long lo, int act, float slope) {
  const int vpr = C >> 3;
  const float neg = neg_slope(act, slope);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * vpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float v[8], sc[8], sh[8], z[8];
    ld_split8(y + r * y_ld + c0, lo, v);
    ld8f(scale + c0, sc);
    ld8f(shift + c0, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = v[j] * sc[j] + sh[j];
    if (y2) {
      ld_split8(y2 + r * y2_ld + c0, lo, v);
      ld8f(scale2 + c0, sc);
      ld8f(shift2 + c0, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += v[j] * sc[j] + sh[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = z[j] > 0.f ? z[j] : z[j] * neg;
    if (post) {
      ld_split8(post + r * post_ld + c0, lo, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += v[j];
    }
    st_split8(out + r * out_ld + c0, lo, z);
  }
}

// dz = da * act'(z), z from y*scale+shift or the sign of aout.  kApply: dy = g*(dz - k1 - xhat*k2) with coef = [g|k1|k2];
// otherwise sums[c] += dz, sums[C+c] += dz*xhat (block-level integer accumulation, then merged into the global matrix)
template <bool kApply>
__global__ void __launch_bounds__(256)
split_bn_bwd_kernel(const bf16* __restrict__ da, long long da_ld, const bf16* __restrict__ y, long long y_ld,
                    const bf16* __restrict__ aout, long long aout_ld, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ coef, bf16* __restrict__ dy, long long dy_ld, StatAcc* sums, int nparts,
                    long long rows, int C, long long lo, int act, float slope) {
  extern __shared__ StatAcc s_sum[];  // [2C] (reduce only)
  if (!kApply) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i].w1 = s_sum[i].w2 = 0;
    __syncthreads();
  }
  const int vpr = C >> 3;
  const float neg = neg_slope(act, slope);
  // a thread keeps ONE channel group for its whole life so that the reduction can stay in registers
  const int groups = blockDim.x / vpr;  // rows handled per block iteration
  const int cv = threadIdx.x % vpr, rg = threadIdx.x / vpr;
  const int c0 = cv << 3;
  if (rg < groups) {
    float mu[8], rs[8], sc[8], sh[8], g[8], k1[8], k2[8], s1[8], s2[8];
    ld8f(mean + c0, mu);
    ld8f(rstd + c0, rs);
    if (!aout) {
      ld8f(scale + c0, sc);
      ld8f(shift + c0, sh);
    }
    if (kApply) {
      ld8f(coef + c0, g);
      ld8f(coef + C + c0, k1);
      ld8f(coef + 2 * C + c0, k2);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    for (long long r = (long long)blockIdx.x * groups + rg; r < rows; r += (long long)gridDim.x * groups) {
      float a[8], v[8], zs[8];
      ld_split8(da + r * da_ld + c0, lo, a);
      ld_split8(y + r * y_ld + c0, lo, v);
      if (aout) {
        ld_split8(aout + r * aout_ld + c0, lo, zs);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) zs[j] = v[j] * sc[j] + sh[j];
      }
      if (kApply) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = a[j] * (zs[j] > 0.f ? 1.f : neg);
          o[j] = g[j] * (dz - k1[j] - (v[j] - mu[j]) * rs[j] * k2[j]);
        }
        st_split8(dy + r * dy_ld + c0, lo, o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = a[j] * (zs[j] > 0.f ? 1.f : neg);
          s1[j] += dz;
          s2[j] = fmaf(dz, (v[j] - mu[j]) * rs[j], s2[j]);
        }
      }
    }
    if (!kApply) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        stat_add(&s_sum[c0 + j], s1[j]);
        stat_add(&s_sum[C + c0 + j], s2[j]);
      }
    }
  }
  if (!kApply) {
    __syncthreads();
    StatAcc* row = sums + (long long)(blockIdx.x % nparts) * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) stat_merge(row + i, s_sum[i]);
  }
}

__global__ void split_act_bwd_kernel(const bf16* __restrict__ da, long long da_ld, const bf16* __restrict__ aout,
                                     long long aout_ld, bf16* __restrict__ dz, long long dz_ld, long long rows, int C,
                                     long long lo, int act, float slope) {
  const int vpr = C >> 3;
  const float neg = neg_slope(act, slope);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * vpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float a[8], z[8];
    ld_split8(da + r * da_ld + c0, lo, a);
    ld_split8(aout + r * aout_ld + c0, lo, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= z[j] > 0.f ? 1.f : neg;
    st_split8(dz + r * dz_ld + c0, lo, a);
  }
}

__global__ void split_copy_slice_kernel(const bf16* __restrict__ src, long long s_ld, long long s_lo,
                                        bf16* __restrict__ dst, long long d_ld, long long d_lo, long long rows, int C,
                                        int accumulate) {
  const int vpr = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * vpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float a[8];
    ld_split8(src + r * s_ld + c0, s_lo, a);
    if (accumulate) {
      float b[8];
      ld_split8(dst + r * d_ld + c0, d_lo, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
    st_split8(dst + r * d_ld + c0, d_lo, a);
  }
}

// fp32 rows <-> split rows (head gradients, layout conversions at the module boundary)
__global__ void split_from_f32_kernel(const float* __restrict__ src, long long s_ld, bf16* __restrict__ dst,
                                      long long d_ld, long long d_lo, long long rows, int C) {
  const int vpr = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * vpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float a[8];
    ld8f(src + r * s_ld + c0, a);
    st_split8(dst + r * d_ld + c0, d_lo, a);
  }
}
__global__ void split_to_f32_kernel(const bf16* __restrict__ src, long long s_ld, long long s_lo,
                                    float* __restrict__ dst, long long d_ld, long long rows, int C) {
  const int vpr = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * vpr;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float a[8];
    ld_split8(src + r * s_ld + c0, s_lo, a);
    float* o = dst + r * d_ld + c0;
    *reinterpret_cast<float4*>(o) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(a[4], a[5], a[6], a[7]);
  }
}

// 2x2 max-pool, stride 2, or stride 1 with a ZERO pad right/bottom (CVC-YOLOv3/models.py:74-84); whole split tensors
// [N,H,W,3C].  kBwd: dx = dy routed to the first maximum of the window (the argmax rule of elementwise.cu).
template <bool kBwd>
__global__ void split_maxpool_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ out,
                                     int N, int H, int W, int C, int stride, int OH, int OW) {
  const int vpr = C >> 3;
  const long long ld = (long long)kSplitPieces * C, lo = C;
  const int GH = kBwd ? H : OH, GW = kBwd ? W : OW;
  const long long total = (long long)N * GH * GW * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int gw = (int)(r % GW); r /= GW;
    const int gh = (int)(r % GH);
    const long long n = r / GH;
    if (!kBwd) {
      float m[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
      for (int dh = 0; dh < 2; ++dh)
        for (int dw = 0; dw < 2; ++dw) {
          const int ih = gh * stride + dh, iw = gw * stride + dw;
          float v[8];
          if (ih < H && iw < W) ld_split8(x + ((n * H + ih) * W + iw) * ld + c0, lo, v);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;  // the zero pad of the stride-1 variant
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
        }
      st_split8(out + ((n * OH + gh) * OW + gw) * ld + c0, lo, m);
    } else {
      // input pixel (gh, gw) receives dy of every window in which it is the FIRST maximum (scan order dh, dw)
      float me[8], acc[8];
      ld_split8(x + ((n * H + gh) * W + gw) * ld + c0, lo, me);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int wh = 0; wh < 2; ++wh)
        for (int ww = 0; ww < 2; ++ww) {
          // window (oh, ow) covers input rows oh*stride + {0,1}
          const int num_h = gh - wh, num_w = gw - ww;
          if (num_h < 0 || num_w < 0 || num_h % stride || num_w % stride) continue;
          const int oh = num_h / stride, ow = num_w / stride;
          if (oh >= OH || ow >= OW) continue;
          float best[8];
          int arg[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = -1; }
          for (int dh = 0; dh < 2; ++dh)
            for (int dw = 0; dw < 2; ++dw) {
              const int ih = oh * stride + dh, iw = ow * stride + dw;
              float v[8];
              if (ih < H && iw < W) ld_split8(x + ((n * H + ih) * W + iw) * ld + c0, lo, v);
              else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (v[j] > best[j]) { best[j] = v[j]; arg[j] = dh * 2 + dw; }
            }
          float g[8];
          ld_split8(dy + ((n * OH + oh) * OW + ow) * ld + c0, lo, g);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (arg[j] == wh * 2 + ww) acc[j] += g[j];
        }
      (void)me;
      st_split8(out + ((n * H + gh) * W + gw) * ld + c0, lo, acc);
    }
  }
}

// dx[n,h,w] (+)= sum of the 2x2 block of dy (nearest x2 upsample backward); whole split tensors
__global__ void split_upsample_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W,
                                          int C, int accumulate) {
  const int vpr = C >> 3;
  const long long ld = (long long)kSplitPieces * C, lo = C;
  const long long total = (long long)N * H * W * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const long long n = r / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dh = 0; dh < 2; ++dh)
      for (int dw = 0; dw < 2; ++dw) {
        float v[8];
        ld_split8(dy + ((n * 2 * H + 2 * h + dh) * (2LL * W) + 2 * w + dw) * ld + c0, lo, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    bf16* o = dx + ((n * H + h) * W + w) * ld + c0;
    if (accumulate) {
      float b[8];
      ld_split8(o, lo, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += b[j];
    }
    st_split8(o, lo, acc);
  }
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_split_bn_apply_act(const void* y, int64_t y_ld, const float* scale, const float* shift,
                                         const void* y2, int64_t y2_ld, const float* scale2, const float* shift2,
                                         const void* post, int64_t post_ld, void* out, int64_t out_ld, int64_t rows,
                                         int C, int64_t lo, int act, float slope, void* stream) {
  B200CV_CHECK_ARG(ok_split(y, y_ld, lo, C) && ok_split(out, out_ld, lo, C) && scale && shift && rows > 0,
                   "split_bn_apply_act: bad args");
  B200CV_CHECK_ARG(!y2 || (ok_split(y2, y2_ld, lo, C) && scale2 && shift2), "split_bn_apply_act: bad second branch");
  B200CV_CHECK_ARG(!post || ok_split(post, post_ld, lo, C), "split_bn_apply_act: bad residual");
  split_bn_apply_kernel<<<grid_of(rows * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)y, y_ld, scale, shift, (const bf16*)y2, y2_ld, scale2, shift2, (const bf16*)post, post_ld,
      (bf16*)out, out_ld, rows, C, lo, act, slope);
  return check_launch("split_bn_apply_act");
}

static int check_bwd(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout, int64_t aout_ld,
                     const float* scale, const float* shift, const float* mean, const float* rstd, int64_t rows, int C,
                     int64_t lo) {
  B200CV_CHECK_ARG(ok_split(da, da_ld, lo, C) && ok_split(y, y_ld, lo, C) && mean && rstd && rows > 0,
                   "split_bn_bwd: bad args");
  B200CV_CHECK_ARG(aout ? ok_split(aout, aout_ld, lo, C) : (scale && shift), "split_bn_bwd: need aout or scale/shift");
  B200CV_CHECK_ARG(C / 8 <= 256, "split_bn_bwd: C=%d too large", C);
  return 0;
}

extern "C" int b200cv_split_bn_bwd_reduce(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                                          int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                                          const float* rstd, void* partials, int nparts, int64_t rows, int C,
                                          int64_t lo, int act, float slope, void* stream) {
  if (int rc = check_bwd(da, da_ld, y, y_ld, aout, aout_ld, scale, shift, mean, rstd, rows, C, lo)) return rc;
  B200CV_CHECK_ARG(partials && nparts > 0, "split_bn_bwd_reduce: null partials");
  const int groups = 256 / (C / 8);
  const int grid = (int)std::max<long long>(1, std::min<long long>((rows + groups - 1) / groups, sm_count() * 8LL));
  split_bn_bwd_kernel<false><<<grid, 256, 2 * C * sizeof(StatAcc), static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)da, da_ld, (const bf16*)y, y_ld, (const bf16*)aout, aout_ld, scale, shift, mean, rstd, nullptr,
      nullptr, 0, static_cast<StatAcc*>(partials), nparts, rows, C, lo, act, slope);
  return check_launch("split_bn_bwd_reduce");
}

extern "C" int b200cv_split_bn_bwd_apply(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                                         int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                                         const float* rstd, const float* coef, void* dy, int64_t dy_ld, int64_t rows,
                                         int C, int64_t lo, int act, float slope, void* stream) {
  if (int rc = check_bwd(da, da_ld, y, y_ld, aout, aout_ld, scale, shift, mean, rstd, rows, C, lo)) return rc;
  B200CV_CHECK_ARG(coef && ok_split(dy, dy_ld, lo, C), "split_bn_bwd_apply: bad args");
  const int groups = 256 / (C / 8);
  const int grid = (int)std::max<long long>(1, std::min<long long>((rows + groups - 1) / groups, sm_count() * 16LL));
  split_bn_bwd_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)da, da_ld, (const bf16*)y, y_ld, (const bf16*)aout, aout_ld, scale, shift, mean, rstd, coef,
      (bf16*)dy, dy_ld, nullptr, 1, rows, C, lo, act, slope);
  return check_launch("split_bn_bwd_apply");
}

extern "C" int b200cv_split_act_bwd(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, void* dz,
                                    int64_t dz_ld, int64_t rows, int C, int64_t lo, int act, float slope,
                                    void* stream) {
  B200CV_CHECK_ARG(ok_split(da, da_ld, lo, C) && ok_split(aout, aout_ld, lo, C) && ok_split(dz, dz_ld, lo, C) &&
                       rows > 0,
                   "split_act_bwd: bad args");
  split_act_bwd_kernel<<<grid_of(rows * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)da, da_ld, (const bf16*)aout, aout_ld, (bf16*)dz, dz_ld, rows, C, lo, act, slope);
  return check_launch("split_act_bwd");
}

extern "C" int b200cv_split_copy_slice(const void* src, int64_t src_ld, int64_t src_lo, void* dst, int64_t dst_ld,
                                       int64_t dst_lo, int64_t rows, int C, int accumulate, void* stream) {
  B200CV_CHECK_ARG(ok_split(src, src_ld, src_lo, C) && ok_split(dst, dst_ld, dst_lo, C) && rows > 0,
                   "split_copy_slice: bad args");
  split_copy_slice_kernel<<<grid_of(rows * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)src, src_ld, src_lo, (bf16*)dst, dst_ld, dst_lo, rows, C, accumulate);
  return check_launch("split_copy_slice");
}

extern "C" int b200cv_split_from_f32(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t dst_lo,
                                     int64_t rows, int C, void* stream) {
  B200CV_CHECK_ARG(src && src_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && src_ld >= C &&
                       ok_split(dst, dst_ld, dst_lo, C) && rows > 0,
                   "split_from_f32: bad args");
  split_from_f32_kernel<<<grid_of(rows * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_ld, (bf16*)dst, dst_ld, dst_lo, rows, C);
  return check_launch("split_from_f32");
}

extern "C" int b200cv_split_to_f32(const void* src, int64_t src_ld, int64_t src_lo, float* dst, int64_t dst_ld,
                                   int64_t rows, int C, void* stream) {
  B200CV_CHECK_ARG(dst && dst_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && dst_ld >= C &&
                       ok_split(src, src_ld, src_lo, C) && rows > 0,
                   "split_to_f32: bad args");
  split_to_f32_kernel<<<grid_of(rows * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)src, src_ld, src_lo, dst, dst_ld, rows, C);
  return check_launch("split_to_f32");
}

extern "C" int b200cv_split_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int stride,
                                           void* stream) {
  B200CV_CHECK_ARG(ok_split(x, kSplitPieces * C, C, C) && ok_split(y, kSplitPieces * C, C, C) && (stride == 1 || stride == 2),
                   "split_maxpool_fwd: bad args");
  const int OH = stride == 2 ? H / 2 : H, OW = stride == 2 ? W / 2 : W;
  split_maxpool_kernel<false><<<grid_of((long long)N * OH * OW * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, nullptr, (bf16*)y, N, H, W, C, stride, OH, OW);
  return check_launch("split_maxpool_fwd");
}

extern "C" int b200cv_split_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C,
                                           int stride, void* stream) {
  B200CV_CHECK_ARG(ok_split(x, kSplitPieces * C, C, C) && ok_split(dy, kSplitPieces * C, C, C) && ok_split(dx, kSplitPieces * C, C, C) &&
                       (stride == 1 || stride == 2),
                   "split_maxpool_bwd: bad args");
  const int OH = stride == 2 ? H / 2 : H, OW = stride == 2 ? W / 2 : W;
  split_maxpool_kernel<true><<<grid_of((long long)N * H * W * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, (const bf16*)dy, (bf16*)dx, N, H, W, C, stride, OH, OW);
  return check_launch("split_maxpool_bwd");
}

extern "C" int b200cv_split_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, int accumulate,
                                           void* stream) {
  B200CV_CHECK_ARG(ok_split(dy, kSplitPieces * C, C, C) && ok_split(dx, kSplitPieces * C, C, C) && N > 0 && H > 0 && W > 0,
                   "split_upsample_bwd: bad args");
  split_upsample_bwd_kernel<<<grid_of((long long)N * H * W * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)dy, (bf16*)dx, N, H, W, C, accumulate);
  return check_launch("split_upsample_bwd");
}
