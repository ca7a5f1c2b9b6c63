// Bandwidth-bound NHWC bf16 kernels around the convolutions: training-mode BatchNorm
// (finalize / apply+activation / backward reductions / backward apply), 2x2 max-pool,
// nearest x2 upsample, channel-slice copy and accumulate, per-channel column sums.
// All tensors are [rows = N*H*W][channels] with an explicit row pitch, so a "tensor" may be a
// channel slice of a wider concat buffer.  Threads move 8 channels (16 bytes) at a time.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "stat_acc.cuh"

namespace b200cv {
namespace {

struct bf16x8 {
  uint4 raw;
};
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& v) {
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  if (act == B200CV_ACT_LEAKY) return z > 0.f ? z : z * slope;
  if (act == B200CV_ACT_RELU) return z > 0.f ? z : 0.f;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
  if (act == B200CV_ACT_LEAKY) return z > 0.f ? 1.f : slope;
  if (act == B200CV_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

int ew_grid(long long work_items, int block) {
  const long long g = (work_items + block - 1) / block;
  return (int)std::max<long long>(1, std::min<long long>(g, (long long)sm_count() * 8));
}

// ------------------------------------------------------------------ BN finalize
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const StatAcc* __restrict__ stats, int parts, float count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps, float momentum,
                   float* running_mean, float* running_var, float* __restrict__ scale, float* __restrict__ shift,
                   float* __restrict__ save_mean, float* __restrict__ save_rstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s1 = stat_fold(stats, parts, 2 * C, c);
  const float s2 = stat_fold(stats, parts, 2 * C, C + c);
  const float mean = s1 / count;
  float var = s2 / count - mean * mean;
  var = var > 0.f ? var : 0.f;
  const float rstd = rsqrtf(var + eps);
  const float g = gamma[c];
  scale[c] = g * rstd;
  shift[c] = beta[c] - mean * g * rstd;
  save_mean[c] = mean;
  save_rstd[c] = rstd;
  if (running_mean) {
    const float m_full = mean + (conv_bias ? conv_bias[c] : 0.f);  // bias is folded out of the conv
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m_full;
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// ------------------------------------------------------------------ streaming helpers (4 channels / thread)
// The BN passes are pure streaming kernels.  A thread owns FOUR fixed channels (8-byte accesses; a warp still
// moves 256 contiguous bytes per instruction) so that the per-channel constants fit in few registers and
// 5-6 blocks of 256 threads stay resident per SM: the 8-channel version needed 112-146 registers, ran one or
// two blocks per SM and was latency-bound at ~3.5 TB/s.
__device__ __forceinline__ uint2 ldg_stream8(const __nv_bfloat16* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg8(__nv_bfloat16* p, const uint2& v) { *reinterpret_cast<uint2*>(p) = v; }
__device__ __forceinline__ void unpack4(const uint2& r, float (&f)[4]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  uint2 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(f[0], f[1]);
  h[1] = __floats2bfloat162_rn(f[2], f[3]);
  return r;
}
__device__ __forceinline__ void load4f(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
// width-generic forms (V = 4 or 8 channels per thread)
template <int V> struct VecT;
template <> struct VecT<4> { typedef uint2 type; };
template <> struct VecT<8> { typedef uint4 type; };
template <int V> __device__ __forceinline__ typename VecT<V>::type ldg_streamV(const __nv_bfloat16* p);
template <> __device__ __forceinline__ uint2 ldg_streamV<4>(const __nv_bfloat16* p) { return ldg_stream8(p); }
template <> __device__ __forceinline__ uint4 ldg_streamV<8>(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void unpackV(const uint2& r, float (&f)[4]) { unpack4(r, f); }
__device__ __forceinline__ void unpackV(const uint4& r, float (&f)[8]) { unpack8(r, f); }
__device__ __forceinline__ void storeV(__nv_bfloat16* p, const float (&f)[4]) { stg8(p, pack4(f)); }
__device__ __forceinline__ void storeV(__nv_bfloat16* p, const float (&f)[8]) { stg16(p, pack8(f)); }
template <int V> __device__ __forceinline__ void loadVf(const float* p, float (&v)[V]) {
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + j));
    v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
  }
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
constexpr int kEwThreads = 256;
constexpr int kEwUnroll = 4;      // rows in flight per thread
constexpr int kEwBlocksPerSm = 6;

// grid of a row-streaming kernel: blocks walk rows with stride gridDim * rows_per_pass
int stream_grid(long long rows, int C, int V = 4) {
  const int rpp = kEwThreads / (C / V);
  const long long passes = (rows + (long long)rpp * kEwUnroll - 1) / ((long long)rpp * kEwUnroll);
  return (int)std::max<long long>(1, std::min<long long>(passes, (long long)sm_count() * kEwBlocksPerSm));
}

// ------------------------------------------------------------------ BN apply (+second branch, +residual)
//   out = act( y*scale + shift  [+ y2*scale2 + shift2] ) [+ post]
struct ApplyArgs {
  const __nv_bfloat16* y; long long y_ld;
  const float* scale; const float* shift;
  const __nv_bfloat16* y2; long long y2_ld;
  const float* scale2; const float* shift2;
  const __nv_bfloat16* post; long long post_ld;
  __nv_bfloat16* out; long long out_ld;
  long long rows; int C; int act; float slope;
};
template <bool kTwo, bool kPost>
__global__ void __launch_bounds__(kEwThreads, 4) bn_apply_kernel(const ApplyArgs a) {
  const int vpr = a.C >> 2;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv << 2;
  if (r0 >= rpp) return;
  float sc[4], sh[4], sc2[4], sh2[4];
  load4f(a.scale + c0, sc);
  load4f(a.shift + c0, sh);
  if (kTwo) {
    load4f(a.scale2 + c0, sc2);
    load4f(a.shift2 + c0, sh2);
  }
  const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
  const long long stride = (long long)gridDim.x * rpp;
  for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
    uint2 qy[kEwUnroll], q2[kEwUnroll], qp[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        qy[u] = ldg_stream8(a.y + rr * a.y_ld + c0);
        if (kTwo) q2[u] = ldg_stream8(a.y2 + rr * a.y2_ld + c0);
        if (kPost) qp[u] = ldg_stream8(a.post + rr * a.post_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        float v[4], z[4];
        unpack4(qy[u], v);
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = v[j] * sc[j] + sh[j];
        if (kTwo) {
          float w[4];
          unpack4(q2[u], w);
#pragma unroll
          for (int j = 0; j < 4; ++j) z[j] += w[j] * sc2[j] + sh2[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = z[j] > 0.f ? z[j] : z[j] * neg;
        if (kPost) {
          float w[4];
          unpack4(qp[u], w);
#pragma unroll
          for (int j = 0; j < 4; ++j) z[j] += w[j];
        }
        stg8(a.out + rr * a.out_ld + c0, pack4(z));
      }
    }
  }
}

// ---- fused: fold the partial statistics (bn_finalize) in the prologue of every block, then apply
struct FusedFwdArgs {
  const StatAcc* stats; int parts; float count;
  const float* gamma; const float* beta; const float* conv_bias; float eps, momentum;
  float* running_mean; float* running_var;
  float* scale; float* shift; float* save_mean; float* save_rstd;
};
template <bool kPost, int V>
__global__ void __launch_bounds__(kEwThreads, V == 8 ? 2 : 4) bn_stats_apply_kernel(const FusedFwdArgs f, const ApplyArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_ss[];  // [scale | shift]
  const int C = a.C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s1 = stat_fold(f.stats, f.parts, 2 * C, c);
    const float s2 = stat_fold(f.stats, f.parts, 2 * C, C + c);
    const float mean = s1 / f.count;
    float var = s2 / f.count - mean * mean;
    var = var > 0.f ? var : 0.f;
    const float rstd = rsqrtf(var + f.eps);
    const float g = f.gamma[c];
    const float sc = g * rstd, sh = f.beta[c] - mean * g * rstd;
    s_ss[c] = sc;
    s_ss[C + c] = sh;
    if (blockIdx.x == 0) {
      f.scale[c] = sc;
      f.shift[c] = sh;
      f.save_mean[c] = mean;
      f.save_rstd[c] = rstd;
      if (f.running_mean) {
        const float m_full = mean + (f.conv_bias ? f.conv_bias[c] : 0.f);
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * m_full;
        const float unbiased = f.count > 1.f ? var * f.count / (f.count - 1.f) : var;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
      }
    }
  }
  __syncthreads();
  const int vpr = C / V;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv * V;
  if (r0 >= rpp) return;
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    sc[j] = s_ss[c0 + j];
    sh[j] = s_ss[C + c0 + j];
  }
  const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
  const long long stride = (long long)gridDim.x * rpp;
  for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
    typename VecT<V>::type qy[kEwUnroll], qp[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      const long long rw = a.rows - 1 - rr;  // walk the rows backwards: the producer wrote the tail last (L2-hot)
      if (rr < a.rows) {
        qy[u] = ldg_streamV<V>(a.y + rw * a.y_ld + c0);
        if (kPost) qp[u] = ldg_streamV<V>(a.post + rw * a.post_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      const long long rw = a.rows - 1 - rr;  // walk the rows backwards: the producer wrote the tail last (L2-hot)
      if (rr < a.rows) {
        float v[V], z[V];
        unpackV(qy[u], v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          z[j] = v[j] * sc[j] + sh[j];
          z[j] = z[j] > 0.f ? z[j] : z[j] * neg;
        }
        if (kPost) {
          float w[V];
          unpackV(qp[u], w);
#pragma unroll
          for (int j = 0; j < V; ++j) z[j] += w[j];
        }
        storeV(a.out + rw * a.out_ld + c0, z);
      }
    }
  }
}

// ------------------------------------------------------------------ BN backward
//   dz = da * act'(z);  pass 1: sums[c] += dz, sums[C+c] += dz * xhat  (xhat = (y-mean)*rstd)
//   pass 2: dy = g * (dz - k1 - xhat*k2)   with coef = [g | k1 | k2]
// z is recomputed as y*scale+shift; `aout` (the saved activation output) can be given instead when z is
// not recomputable from one branch alone (only its sign is used).
struct BwdArgs {
  const __nv_bfloat16* da; long long da_ld;
  const __nv_bfloat16* y; long long y_ld;
  const __nv_bfloat16* aout; long long aout_ld;   // optional: sign source for act'
  const float* scale; const float* shift;         // of this BN (to recompute z when aout == null)
  const float* mean; const float* rstd;
  StatAcc* sums;                                  // pass 1 out: [nparts][2C], ADDED to (caller zeroes)
  int nparts;
  const float* coef;                              // pass 2 in  [3C]: g, k1, k2
  __nv_bfloat16* dy; long long dy_ld;             // pass 2 out
  long long rows; int C; int act; float slope;
};

template <bool kAout, int V>
__global__ void __launch_bounds__(kEwThreads, V == 8 ? 2 : 4) bn_bwd_reduce_kernel(const BwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ StatAcc s_acc[];  // [2C]
  const int C = a.C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i].w1 = s_acc[i].w2 = 0;
  __syncthreads();
  const int vpr = C / V;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv * V;
  if (r0 < rpp) {
    float sc[V], sh[V], mean[V];
    if (!kAout) {
      loadVf<V>(a.scale + c0, sc);
      loadVf<V>(a.shift + c0, sh);
    }
    loadVf<V>(a.mean + c0, mean);
    const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
    float s1[V], s2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s1[j] = s2[j] = 0.f;
    const long long stride = (long long)gridDim.x * rpp;
    for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
      typename VecT<V>::type qda[kEwUnroll], qy[kEwUnroll], qa[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long rr = r + u * stride;
        if (rr < a.rows) {
          qda[u] = ldg_streamV<V>(a.da + rr * a.da_ld + c0);
          qy[u] = ldg_streamV<V>(a.y + rr * a.y_ld + c0);
          if (kAout) qa[u] = ldg_streamV<V>(a.aout + rr * a.aout_ld + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        if (r + u * stride < a.rows) {
          float da[V], y[V], zs[V];
          unpackV(qda[u], da);
          unpackV(qy[u], y);
          if (kAout) {
            unpackV(qa[u], zs);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) zs[j] = y[j] * sc[j] + sh[j];
          }
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float dz = da[j] * (zs[j] > 0.f ? 1.f : neg);
            s1[j] += dz;
            s2[j] = fmaf(dz, y[j] - mean[j], s2[j]);
          }
        }
      }
    }
    float rstd[V];
    loadVf<V>(a.rstd + c0, rstd);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      stat_add(&s_acc[c0 + j], s1[j]);
      stat_add(&s_acc[C + c0 + j], s2[j] * rstd[j]);
    }
  }
  __syncthreads();
  StatAcc* row = a.sums + (long long)(blockIdx.x % a.nparts) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) stat_merge(row + i, s_acc[i]);
}

// coef[c] = gamma*rstd ; coef[C+c] = sum_dz/M ; coef[2C+c] = sum_dz_xhat/M ; also dgamma/dbeta
// block = 32 channels x 8 part-lanes: partial rows are summed 8 at a time, then folded through shared memory
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const StatAcc* __restrict__ partials, int nparts, const float* __restrict__ gamma,
                       const float* __restrict__ rstd, float count, float* __restrict__ coef, float* dgamma,
                       float* dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s1 = stat_fold(partials, nparts, 2 * C, c);
    const float s2 = stat_fold(partials, nparts, 2 * C, C + c);
    coef[c] = gamma[c] * rstd[c];
    coef[C + c] = s1 / count;
    coef[2 * C + c] = s2 / count;
    if (dbeta) dbeta[c] = s1;
    if (dgamma) dgamma[c] = s2;
  }
}

//   dy = g*(dz - k1 - xhat*k2) = g*dz + A*y + B   with A = -g*k2*rstd, B = -g*k1 - A*mean
template <bool kAout>
__global__ void __launch_bounds__(kEwThreads, 4) bn_bwd_apply_kernel(const BwdArgs a) {
  const int C = a.C;
  const int vpr = C >> 2;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv << 2;
  if (r0 >= rpp) return;
  float sc[4], sh[4], g[4], A[4], B[4];
  if (!kAout) {
    load4f(a.scale + c0, sc);
    load4f(a.shift + c0, sh);
  }
  {
    float k1[4], k2[4], mean[4], rstd[4];
    load4f(a.coef + c0, g);
    load4f(a.coef + C + c0, k1);
    load4f(a.coef + 2 * C + c0, k2);
    load4f(a.mean + c0, mean);
    load4f(a.rstd + c0, rstd);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A[j] = -g[j] * k2[j] * rstd[j];
      B[j] = -g[j] * k1[j] - A[j] * mean[j];
    }
  }
  const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
  const long long stride = (long long)gridDim.x * rpp;
  for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
    uint2 qda[kEwUnroll], qy[kEwUnroll], qa[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        qda[u] = ldg_stream8(a.da + rr * a.da_ld + c0);
        qy[u] = ldg_stream8(a.y + rr * a.y_ld + c0);
        if (kAout) qa[u] = ldg_stream8(a.aout + rr * a.aout_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        float da[4], y[4], zs[4], o[4];
        unpack4(qda[u], da);
        unpack4(qy[u], y);
        if (kAout) {
          unpack4(qa[u], zs);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) zs[j] = y[j] * sc[j] + sh[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dz = da[j] * (zs[j] > 0.f ? 1.f : neg);
          o[j] = fmaf(g[j], dz, fmaf(A[j], y[j], B[j]));
        }
        stg8(a.dy + rr * a.dy_ld + c0, pack4(o));
      }
    }
  }
}

// ---- two BatchNorms under one activation: out = act(bn_A(yA) + bn_B(yB))  (RektNet/resnet.py:21-27: the shortcut
// BN and the second conv's BN meet in one ReLU).  Both see dz = da * act'(out): one pass reads da and the saved output
// once for the two layers instead of once per layer (4 tensor reads instead of 6 in the reduction, 4 reads + 2 writes
// instead of 6 + 2 in the apply pass).
struct Bwd2Args {
  const __nv_bfloat16* da; long long da_ld;
  const __nv_bfloat16* aout; long long aout_ld;
  const __nv_bfloat16* yA; long long yA_ld;
  const __nv_bfloat16* yB; long long yB_ld;
  const float* meanA; const float* rstdA;
  const float* meanB; const float* rstdB;
  StatAcc* sumsA; StatAcc* sumsB; int nparts;      // reduce: [nparts][2C] each, ADDED to
  const float* coefA; const float* coefB;          // apply: [3C] each (g, k1, k2)
  __nv_bfloat16* dyA; long long dyA_ld;
  __nv_bfloat16* dyB; long long dyB_ld;
  long long rows; int C; int act; float slope;
};

__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_reduce2_kernel(const Bwd2Args a) {
  extern __shared__ StatAcc s_acc2[];  // [3C]: sum dz | sum dz*(yA-meanA)*rstdA | sum dz*(yB-meanB)*rstdB
  const int C = a.C;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_acc2[i].w1 = s_acc2[i].w2 = 0;
  __syncthreads();
  const int vpr = C >> 2;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv << 2;
  if (r0 < rpp) {
    float meanA[4], meanB[4];
    load4f(a.meanA + c0, meanA);
    load4f(a.meanB + c0, meanB);
    const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, sA[4] = {0.f, 0.f, 0.f, 0.f}, sB[4] = {0.f, 0.f, 0.f, 0.f};
    const long long stride = (long long)gridDim.x * rpp;
    for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
      uint2 qda[kEwUnroll], qa[kEwUnroll], qyA[kEwUnroll], qyB[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long rr = r + u * stride;
        if (rr < a.rows) {
          qda[u] = ldg_stream8(a.da + rr * a.da_ld + c0);
          qa[u] = ldg_stream8(a.aout + rr * a.aout_ld + c0);
          qyA[u] = ldg_stream8(a.yA + rr * a.yA_ld + c0);
          qyB[u] = ldg_stream8(a.yB + rr * a.yB_ld + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        if (r + u * stride < a.rows) {
          float da[4], zs[4], yA[4], yB[4];
          unpack4(qda[u], da);
          unpack4(qa[u], zs);
          unpack4(qyA[u], yA);
          unpack4(qyB[u], yB);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dz = da[j] * (zs[j] > 0.f ? 1.f : neg);
            s1[j] += dz;
            sA[j] = fmaf(dz, yA[j] - meanA[j], sA[j]);
            sB[j] = fmaf(dz, yB[j] - meanB[j], sB[j]);
          }
        }
      }
    }
    float rstdA[4], rstdB[4];
    load4f(a.rstdA + c0, rstdA);
    load4f(a.rstdB + c0, rstdB);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      stat_add(&s_acc2[c0 + j], s1[j]);
      stat_add(&s_acc2[C + c0 + j], sA[j] * rstdA[j]);
      stat_add(&s_acc2[2 * C + c0 + j], sB[j] * rstdB[j]);
    }
  }
  __syncthreads();
  const long long part = (long long)(blockIdx.x % a.nparts) * 2 * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    stat_merge(a.sumsA + part + i, s_acc2[i]);
    stat_merge(a.sumsB + part + i, s_acc2[i]);
    stat_merge(a.sumsA + part + C + i, s_acc2[C + i]);
    stat_merge(a.sumsB + part + C + i, s_acc2[2 * C + i]);
  }
}

__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_apply2_kernel(const Bwd2Args a) {
  const int C = a.C;
  const int vpr = C >> 2;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv << 2;
  if (r0 >= rpp) return;
  // dy = g*dz + A*y + B per layer, A = -g*k2*rstd, B = -g*k1 - A*mean (see bn_bwd_apply_kernel)
  float gA[4], AA[4], BA[4], gB[4], AB[4], BB[4];
  {
    float k1[4], k2[4], mean[4], rstd[4];
    load4f(a.coefA + c0, gA);
    load4f(a.coefA + C + c0, k1);
    load4f(a.coefA + 2 * C + c0, k2);
    load4f(a.meanA + c0, mean);
    load4f(a.rstdA + c0, rstd);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      AA[j] = -gA[j] * k2[j] * rstd[j];
      BA[j] = -gA[j] * k1[j] - AA[j] * mean[j];
    }
    load4f(a.coefB + c0, gB);
    load4f(a.coefB + C + c0, k1);
    load4f(a.coefB + 2 * C + c0, k2);
    load4f(a.meanB + c0, mean);
    load4f(a.rstdB + c0, rstd);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      AB[j] = -gB[j] * k2[j] * rstd[j];
      BB[j] = -gB[j] * k1[j] - AB[j] * mean[j];
    }
  }
  const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
  const long long stride = (long long)gridDim.x * rpp;
  for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
    uint2 qda[kEwUnroll], qa[kEwUnroll], qyA[kEwUnroll], qyB[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        qda[u] = ldg_stream8(a.da + rr * a.da_ld + c0);
        qa[u] = ldg_stream8(a.aout + rr * a.aout_ld + c0);
        qyA[u] = ldg_stream8(a.yA + rr * a.yA_ld + c0);
        qyB[u] = ldg_stream8(a.yB + rr * a.yB_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      if (rr < a.rows) {
        float da[4], zs[4], yA[4], yB[4], oA[4], oB[4];
        unpack4(qda[u], da);
        unpack4(qa[u], zs);
        unpack4(qyA[u], yA);
        unpack4(qyB[u], yB);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dz = da[j] * (zs[j] > 0.f ? 1.f : neg);
          oA[j] = fmaf(gA[j], dz, fmaf(AA[j], yA[j], BA[j]));
          oB[j] = fmaf(gB[j], dz, fmaf(AB[j], yB[j], BB[j]));
        }
        stg8(a.dyA + rr * a.dyA_ld + c0, pack4(oA));
        stg8(a.dyB + rr * a.dyB_ld + c0, pack4(oB));
      }
    }
  }
}

// ---- fused: fold the backward partial sums (bn_bwd_finalize) in the prologue of every block, then apply
struct FusedBwdArgs {
  const StatAcc* partials; int nparts; float count;
  const float* gamma; float* coef; float* dgamma; float* dbeta;
};
template <int V>
__global__ void __launch_bounds__(kEwThreads, V == 8 ? 2 : 4) bn_bwd_stats_apply_kernel(const FusedBwdArgs f, const BwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_gab[];  // [g | A | B]
  const int C = a.C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s1 = stat_fold(f.partials, f.nparts, 2 * C, c);
    const float s2 = stat_fold(f.partials, f.nparts, 2 * C, C + c);
    const float rstd = a.rstd[c], mean = a.mean[c];
    const float g = f.gamma[c] * rstd;
    const float k1 = s1 / f.count, k2 = s2 / f.count;
    const float A = -g * k2 * rstd;
    s_gab[c] = g;
    s_gab[C + c] = A;
    s_gab[2 * C + c] = -g * k1 - A * mean;
    if (blockIdx.x == 0) {
      if (f.coef) {
        f.coef[c] = g;
        f.coef[C + c] = k1;
        f.coef[2 * C + c] = k2;
      }
      if (f.dbeta) f.dbeta[c] = s1;
      if (f.dgamma) f.dgamma[c] = s2;
    }
  }
  __syncthreads();
  const int vpr = C / V;
  const int rpp = kEwThreads / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  const int c0 = cv * V;
  if (r0 >= rpp) return;
  float sc[V], sh[V], g[V], A[V], B[V];
  loadVf<V>(a.scale + c0, sc);
  loadVf<V>(a.shift + c0, sh);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = s_gab[c0 + j];
    A[j] = s_gab[C + c0 + j];
    B[j] = s_gab[2 * C + c0 + j];
  }
  const float neg = a.act == B200CV_ACT_LEAKY ? a.slope : (a.act == B200CV_ACT_RELU ? 0.f : 1.f);
  const long long stride = (long long)gridDim.x * rpp;
  for (long long r = (long long)blockIdx.x * rpp + r0; r < a.rows; r += stride * kEwUnroll) {
    typename VecT<V>::type qda[kEwUnroll], qy[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      const long long rw = a.rows - 1 - rr;  // walk the rows backwards: the producer wrote the tail last (L2-hot)
      if (rr < a.rows) {
        qda[u] = ldg_streamV<V>(a.da + rw * a.da_ld + c0);
        qy[u] = ldg_streamV<V>(a.y + rw * a.y_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long rr = r + u * stride;
      const long long rw = a.rows - 1 - rr;  // walk the rows backwards: the producer wrote the tail last (L2-hot)
      if (rr < a.rows) {
        float da[V], y[V], o[V];
        unpackV(qda[u], da);
        unpackV(qy[u], y);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float z = y[j] * sc[j] + sh[j];
          const float dz = da[j] * (z > 0.f ? 1.f : neg);
          o[j] = fmaf(g[j], dz, fmaf(A[j], y[j], B[j]));
        }
        storeV(a.dy + rw * a.dy_ld + c0, o);
      }
    }
  }
}

// ------------------------------------------------------------------ activation-only backward (no BN)
//   dz = da * act'(aout)
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ da, long long da_ld,
                               const __nv_bfloat16* __restrict__ aout, long long aout_ld,
                               __nv_bfloat16* __restrict__ dz, long long dz_ld, long long rows, int C, int act,
                               float slope) {
  const int vpr = C >> 3;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    float g[8], z[8];
    unpack8(ldg16(da + r * da_ld + c0), g);
    unpack8(ldg16(aout + r * aout_ld + c0), z);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act_grad(z[j], act, slope);
    stg16(dz + r * dz_ld + c0, pack8(g));
  }
}

// ------------------------------------------------------------------ slice copy / accumulate
__global__ void copy_slice_kernel(const __nv_bfloat16* __restrict__ src, long long s_ld,
                                  __nv_bfloat16* __restrict__ dst, long long d_ld, long long rows, int C,
                                  int accumulate) {
  const int vpr = C >> 3;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c0 = (int)(i - r * vpr) << 3;
    uint4 v = ldg16(src + r * s_ld + c0);
    if (accumulate) {
      float a[8], b[8];
      unpack8(v, a);
      unpack8(*reinterpret_cast<const uint4*>(dst + r * d_ld + c0), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      v = pack8(a);
    }
    stg16(dst + r * d_ld + c0, v);
  }
}

// ------------------------------------------------------------------ column sums (conv bias gradient)
__global__ void __launch_bounds__(256) col_sum_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                      long long rows, int C, float* __restrict__ out) {
  extern __shared__ float s_colsum[];
  float* s_acc = s_colsum;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int cv = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  if (r0 < rpp) {
    float s[8] = {0};
    for (long long r = (long long)blockIdx.x * rpp + r0; r < rows; r += (long long)gridDim.x * rpp) {
      float v[8];
      unpack8(ldg16(x + r * ld + (cv << 3)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(cv << 3) + j], s[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, s_acc[i]);
}

// ------------------------------------------------------------------ 2x2 max-pool (stride 2, or stride 1 with
// a ZERO pad on the right/bottom edge -- nn.ZeroPad2d((0,1,0,1)) + MaxPool2d(2,1), CVC-YOLOv3/models.py:74-84)
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                   int W, int C, int stride, int OH, int OW) {
  const int vpr = C >> 3;
  const long long total = (long long)N * OH * OW * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long long n = r / OH;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int ih = oh * stride + dh, iw = ow * stride + dw;
        float v[8];
        if (ih < H && iw < W) unpack8(ldg16(x + ((n * H + ih) * W + iw) * C + c0), v);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;  // zero padding takes part in the max
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    stg16(y + ((n * OH + oh) * OW + ow) * C + c0, pack8(m));
  }
}
// Gather form: every input pixel sums the gradients of the windows in which it is the FIRST maximum
// (scan order dh, dw -- the element PyTorch's max_pool2d backward routes to).
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int stride, int OH,
                                   int OW) {
  const int vpr = C >> 3;
  const long long total = (long long)N * H * W * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int iw = (int)(r % W); r /= W;
    const int ih = (int)(r % H);
    const long long n = r / H;
    float g[8] = {0};
    float me[8];
    unpack8(ldg16(x + ((n * H + ih) * W + iw) * C + c0), me);
    for (int oh = (ih - 1 + stride - 1) / stride; oh * stride <= ih; ++oh) {
      if (oh < 0 || oh >= OH) continue;
      for (int ow = (iw - 1 + stride - 1) / stride; ow * stride <= iw; ++ow) {
        if (ow < 0 || ow >= OW) continue;
        const int my_pos = (ih - oh * stride) * 2 + (iw - ow * stride);
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
          if (pos == my_pos) continue;
          const int jh = oh * stride + (pos >> 1), jw = ow * stride + (pos & 1);
          float v[8];
          if (jh < H && jw < W) unpack8(ldg16(x + ((n * H + jh) * W + jw) * C + c0), v);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (pos < my_pos ? v[j] >= me[j] : v[j] > me[j]) win[j] = false;
        }
        float d[8];
        unpack8(ldg16(dy + ((n * OH + oh) * OW + ow) * C + c0), d);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (win[j]) g[j] += d[j];
      }
    }
    stg16(dx + ((n * H + ih) * W + iw) * C + c0, pack8(g));
  }
}

// Stride-2 windows on even H, W do not overlap: one thread per WINDOW reads its four inputs and dy once and writes the
// four dx values (dy to the first maximum, zeros elsewhere) -- the gather form above re-read three neighbours and dy
// for every input pixel (0.55 ms of the 2.6 ms yolov3-tiny step at 416x416 bs16).
__global__ void maxpool_bwd_s2_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                      __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int OH, int OW) {
  const int vpr = C >> 3;
  const long long total = (long long)N * OH * OW * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long long n = r / OH;
    const long long base = ((n * H + 2 * oh) * W + 2 * ow) * C + c0;
    float v[4][8], d[8];
    unpack8(ldg16(x + base), v[0]);
    unpack8(ldg16(x + base + C), v[1]);
    unpack8(ldg16(x + base + (long long)W * C), v[2]);
    unpack8(ldg16(x + base + (long long)W * C + C), v[3]);
    unpack8(ldg16(dy + ((n * OH + oh) * OW + ow) * C + c0), d);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int arg = 0;
      float best = v[0][j];
#pragma unroll
      for (int pos = 1; pos < 4; ++pos)
        if (v[pos][j] > best) { best = v[pos][j]; arg = pos; }  // first maximum in scan order
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) o[pos][j] = pos == arg ? d[j] : 0.f;
    }
    stg16(dx + base, pack8(o[0]));
    stg16(dx + base + C, pack8(o[1]));
    stg16(dx + base + (long long)W * C, pack8(o[2]));
    stg16(dx + base + (long long)W * C + C, pack8(o[3]));
  }
}

// ------------------------------------------------------------------ nearest x2 upsample
__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                      long long y_ld, int N, int H, int W, int C) {
  const int vpr = C >> 3;
  const int OH = 2 * H, OW = 2 * W;
  const long long total = (long long)N * OH * OW * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long long n = r / OH;
    stg16(y + ((n * OH + oh) * OW + ow) * y_ld + c0, ldg16(x + ((n * H + (oh >> 1)) * W + (ow >> 1)) * C + c0));
  }
}
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_ld,
                                      __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int accumulate) {
  const int vpr = C >> 3;
  const int OH = 2 * H, OW = 2 * W;
  const long long total = (long long)N * H * W * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vpr) << 3;
    long long r = i / vpr;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const long long n = r / H;
    float g[8] = {0};
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        float v[8];
        unpack8(ldg16(dy + ((n * OH + 2 * h + dh) * OW + 2 * w + dw) * dy_ld + c0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += v[j];
      }
    __nv_bfloat16* d = dx + ((n * H + h) * W + w) * C + c0;
    if (accumulate) {
      float b[8];
      unpack8(*reinterpret_cast<const uint4*>(d), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += b[j];
    }
    stg16(d, pack8(g));
  }
}

// channels per thread of the streaming BN kernels (B200CV_EW_VEC=4|8 overrides; tuning aid)
int ew_vec(int C) {
  static const int forced = getenv("B200CV_EW_VEC") ? atoi(getenv("B200CV_EW_VEC")) : 0;
  const int v = forced == 4 || forced == 8 ? forced : 4;  // measured: 8 is no faster (5.0-5.6 TB/s either way)
  return (v == 8 && C % 8 == 0 && C / 8 <= kEwThreads) ? 8 : 4;
}

bool ok_vec(const void* p, long long ld, int C) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0 && C % 8 == 0 && C > 0;
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;
typedef __nv_bfloat16 bf16;

extern "C" int b200cv_bn_finalize(const void* stats, int stats_parts, int64_t count, const float* gamma, const float* beta,
                                  const float* conv_bias, float eps, float momentum, float* running_mean,
                                  float* running_var, float* scale, float* shift, float* save_mean,
                                  float* save_rstd, int C, void* stream) {
  B200CV_CHECK_ARG(stats && gamma && beta && scale && shift && save_mean && save_rstd && C > 0 && count > 0 &&
                       stats_parts > 0,
                   "bn_finalize: bad args");
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const StatAcc*>(stats), stats_parts, (float)count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift,
      save_mean, save_rstd, C);
  return check_launch("bn_finalize");
}

extern "C" int b200cv_bn_apply_act(const void* y, int64_t y_ld, const float* scale, const float* shift,
                                   const void* y2, int64_t y2_ld, const float* scale2, const float* shift2,
                                   const void* post, int64_t post_ld, void* out, int64_t out_ld, int64_t rows,
                                   int C, int act, float slope, void* stream) {
  B200CV_CHECK_ARG(ok_vec(y, y_ld, C) && ok_vec(out, out_ld, C) && scale && shift && rows > 0,
                   "bn_apply_act: bad args");
  B200CV_CHECK_ARG(!y2 || (ok_vec(y2, y2_ld, C) && scale2 && shift2), "bn_apply_act: bad second branch");
  B200CV_CHECK_ARG(!post || ok_vec(post, post_ld, C), "bn_apply_act: bad residual");
  ApplyArgs a{(const bf16*)y, y_ld, scale, shift, (const bf16*)y2, y2_ld, scale2, shift2,
              (const bf16*)post, post_ld, (bf16*)out, out_ld, rows, C, act, slope};
  B200CV_CHECK_ARG(C / 4 <= kEwThreads, "bn_apply_act: C too large");
  {
    const int grid = stream_grid(rows, C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y2 && post) bn_apply_kernel<true, true><<<grid, kEwThreads, 0, st>>>(a);
    else if (y2) bn_apply_kernel<true, false><<<grid, kEwThreads, 0, st>>>(a);
    else if (post) bn_apply_kernel<false, true><<<grid, kEwThreads, 0, st>>>(a);
    else bn_apply_kernel<false, false><<<grid, kEwThreads, 0, st>>>(a);
  }
  return check_launch("bn_apply_act");
}

static int fill_bwd(BwdArgs& a, const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                    int64_t aout_ld, const float* scale, const float* shift, const float* mean, const float* rstd,
                    int64_t rows, int C, int act, float slope) {
  B200CV_CHECK_ARG(ok_vec(da, da_ld, C) && ok_vec(y, y_ld, C) && mean && rstd && rows > 0, "bn_bwd: bad args");
  B200CV_CHECK_ARG(aout ? ok_vec(aout, aout_ld, C) : (scale && shift), "bn_bwd: need aout or scale/shift");
  B200CV_CHECK_ARG(C / 4 <= kEwThreads, "bn_bwd: C=%d too large", C);
  a = BwdArgs{};
  a.da = (const bf16*)da; a.da_ld = da_ld; a.y = (const bf16*)y; a.y_ld = y_ld;
  a.aout = (const bf16*)aout; a.aout_ld = aout_ld; a.scale = scale; a.shift = shift;
  a.mean = mean; a.rstd = rstd; a.rows = rows; a.C = C; a.act = act; a.slope = slope;
  return 0;
}

extern "C" int b200cv_bn_bwd_reduce(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                                    int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                                    const float* rstd, void* partials, int nparts, int64_t rows, int C, int act,
                                    float slope, void* stream) {
  BwdArgs a;
  if (int rc = fill_bwd(a, da, da_ld, y, y_ld, aout, aout_ld, scale, shift, mean, rstd, rows, C, act, slope))
    return rc;
  B200CV_CHECK_ARG(partials != nullptr && nparts > 0, "bn_bwd_reduce: null partials");
  a.sums = static_cast<StatAcc*>(partials);  // [nparts][2C], zeroed by the caller: block b ADDS its sums to row b % nparts
  a.nparts = nparts;
  const int V = ew_vec(C);
  const int grid = stream_grid(rows, C, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (V == 8) {
    if (aout) launch_pdl(bn_bwd_reduce_kernel<true, 8>, dim3(grid), dim3(kEwThreads), 2 * C * sizeof(StatAcc), st, a);
    else launch_pdl(bn_bwd_reduce_kernel<false, 8>, dim3(grid), dim3(kEwThreads), 2 * C * sizeof(StatAcc), st, a);
  } else {
    if (aout) launch_pdl(bn_bwd_reduce_kernel<true, 4>, dim3(grid), dim3(kEwThreads), 2 * C * sizeof(StatAcc), st, a);
    else launch_pdl(bn_bwd_reduce_kernel<false, 4>, dim3(grid), dim3(kEwThreads), 2 * C * sizeof(StatAcc), st, a);
  }
  return check_launch("bn_bwd_reduce");
}

extern "C" int b200cv_bn_bwd_finalize(const void* partials, int nparts, const float* gamma, const float* rstd,
                                      int64_t count, float* coef, float* dgamma, float* dbeta, int C, void* stream) {
  B200CV_CHECK_ARG(partials && nparts > 0 && gamma && rstd && coef && C > 0 && count > 0, "bn_bwd_finalize: bad args");
  bn_bwd_finalize_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const StatAcc*>(partials), nparts, gamma, rstd, (float)count, coef, dgamma, dbeta, C);
  return check_launch("bn_bwd_finalize");
}

extern "C" int b200cv_bn_bwd_apply(const void* da, int64_t da_ld, const void* y, int64_t y_ld, const void* aout,
                                   int64_t aout_ld, const float* scale, const float* shift, const float* mean,
                                   const float* rstd, const float* coef, void* dy, int64_t dy_ld, int64_t rows,
                                   int C, int act, float slope, void* stream) {
  BwdArgs a;
  if (int rc = fill_bwd(a, da, da_ld, y, y_ld, aout, aout_ld, scale, shift, mean, rstd, rows, C, act, slope))
    return rc;
  B200CV_CHECK_ARG(coef && ok_vec(dy, dy_ld, C), "bn_bwd_apply: bad args");
  a.coef = coef; a.dy = (bf16*)dy; a.dy_ld = dy_ld;
  {
    const int grid = stream_grid(rows, C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (aout) bn_bwd_apply_kernel<true><<<grid, kEwThreads, 0, st>>>(a);
    else bn_bwd_apply_kernel<false><<<grid, kEwThreads, 0, st>>>(a);
  }
  return check_launch("bn_bwd_apply");
}

static int fill_bwd2(Bwd2Args& a, const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, const void* yA,
                     int64_t yA_ld, const void* yB, int64_t yB_ld, const float* meanA, const float* rstdA,
                     const float* meanB, const float* rstdB, int64_t rows, int C, int act, float slope) {
  B200CV_CHECK_ARG(ok_vec(da, da_ld, C) && ok_vec(aout, aout_ld, C) && ok_vec(yA, yA_ld, C) && ok_vec(yB, yB_ld, C) &&
                       meanA && rstdA && meanB && rstdB && rows > 0,
                   "bn_bwd2: bad args");
  B200CV_CHECK_ARG(C / 4 <= kEwThreads && kEwThreads % (C / 4) == 0, "bn_bwd2: unsupported C=%d", C);
  a = Bwd2Args{};
  a.da = (const bf16*)da; a.da_ld = da_ld; a.aout = (const bf16*)aout; a.aout_ld = aout_ld;
  a.yA = (const bf16*)yA; a.yA_ld = yA_ld; a.yB = (const bf16*)yB; a.yB_ld = yB_ld;
  a.meanA = meanA; a.rstdA = rstdA; a.meanB = meanB; a.rstdB = rstdB;
  a.rows = rows; a.C = C; a.act = act; a.slope = slope;
  return 0;
}

extern "C" int b200cv_bn_bwd_reduce2(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, const void* yA,
                                     int64_t yA_ld, const void* yB, int64_t yB_ld, const float* meanA,
                                     const float* rstdA, const float* meanB, const float* rstdB, void* partialsA,
                                     void* partialsB, int nparts, int64_t rows, int C, int act, float slope,
                                     void* stream) {
  Bwd2Args a;
  if (int rc = fill_bwd2(a, da, da_ld, aout, aout_ld, yA, yA_ld, yB, yB_ld, meanA, rstdA, meanB, rstdB, rows, C, act,
                         slope))
    return rc;
  B200CV_CHECK_ARG(partialsA && partialsB && nparts > 0, "bn_bwd_reduce2: null partials");
  a.sumsA = static_cast<StatAcc*>(partialsA); a.sumsB = static_cast<StatAcc*>(partialsB); a.nparts = nparts;
  bn_bwd_reduce2_kernel<<<stream_grid(rows, C), kEwThreads, 3 * C * sizeof(StatAcc), static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("bn_bwd_reduce2");
}

extern "C" int b200cv_bn_bwd_apply2(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, const void* yA,
                                    int64_t yA_ld, const void* yB, int64_t yB_ld, const float* meanA,
                                    const float* rstdA, const float* meanB, const float* rstdB, const float* coefA,
                                    const float* coefB, void* dyA, int64_t dyA_ld, void* dyB, int64_t dyB_ld,
                                    int64_t rows, int C, int act, float slope, void* stream) {
  Bwd2Args a;
  if (int rc = fill_bwd2(a, da, da_ld, aout, aout_ld, yA, yA_ld, yB, yB_ld, meanA, rstdA, meanB, rstdB, rows, C, act,
                         slope))
    return rc;
  B200CV_CHECK_ARG(coefA && coefB && ok_vec(dyA, dyA_ld, C) && ok_vec(dyB, dyB_ld, C), "bn_bwd_apply2: bad args");
  a.coefA = coefA; a.coefB = coefB;
  a.dyA = (bf16*)dyA; a.dyA_ld = dyA_ld; a.dyB = (bf16*)dyB; a.dyB_ld = dyB_ld;
  bn_bwd_apply2_kernel<<<stream_grid(rows, C), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("bn_bwd_apply2");
}

extern "C" int b200cv_bn_stats_apply_act(const void* stats, int stats_parts, int64_t count, const float* gamma,
                                         const float* beta, const float* conv_bias, float eps, float momentum,
                                         float* running_mean, float* running_var, float* scale, float* shift,
                                         float* save_mean, float* save_rstd, const void* y, int64_t y_ld,
                                         const void* post, int64_t post_ld, void* out, int64_t out_ld, int64_t rows,
                                         int C, int act, float slope, void* stream) {
  B200CV_CHECK_ARG(stats && gamma && beta && scale && shift && save_mean && save_rstd && C > 0 && count > 0 &&
                       stats_parts > 0,
                   "bn_stats_apply_act: bad statistics args");
  B200CV_CHECK_ARG(ok_vec(y, y_ld, C) && ok_vec(out, out_ld, C) && rows > 0, "bn_stats_apply_act: bad args");
  B200CV_CHECK_ARG(!post || ok_vec(post, post_ld, C), "bn_stats_apply_act: bad residual");
  B200CV_CHECK_ARG(C / 4 <= kEwThreads, "bn_stats_apply_act: C too large");
  FusedFwdArgs f{static_cast<const StatAcc*>(stats), stats_parts, (float)count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var,
                 scale, shift, save_mean, save_rstd};
  ApplyArgs a{(const bf16*)y, y_ld, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
              (const bf16*)post, post_ld, (bf16*)out, out_ld, rows, C, act, slope};
  const int V = ew_vec(C);
  const int grid = stream_grid(rows, C, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (V == 8) {
    if (post) launch_pdl(bn_stats_apply_kernel<true, 8>, dim3(grid), dim3(kEwThreads), smem, st, f, a);
    else launch_pdl(bn_stats_apply_kernel<false, 8>, dim3(grid), dim3(kEwThreads), smem, st, f, a);
  } else {
    if (post) launch_pdl(bn_stats_apply_kernel<true, 4>, dim3(grid), dim3(kEwThreads), smem, st, f, a);
    else launch_pdl(bn_stats_apply_kernel<false, 4>, dim3(grid), dim3(kEwThreads), smem, st, f, a);
  }
  return check_launch("bn_stats_apply_act");
}

extern "C" int b200cv_bn_bwd_stats_apply(const void* partials, int nparts, int64_t count, const float* gamma,
                                         float* coef, float* dgamma, float* dbeta, const void* da, int64_t da_ld,
                                         const void* y, int64_t y_ld, const float* scale, const float* shift,
                                         const float* mean, const float* rstd, void* dy, int64_t dy_ld,
                                         int64_t rows, int C, int act, float slope, void* stream) {
  BwdArgs a;
  if (int rc = fill_bwd(a, da, da_ld, y, y_ld, nullptr, 0, scale, shift, mean, rstd, rows, C, act, slope)) return rc;
  B200CV_CHECK_ARG(partials && nparts > 0 && gamma && count > 0 && ok_vec(dy, dy_ld, C),
                   "bn_bwd_stats_apply: bad args");
  a.dy = (bf16*)dy; a.dy_ld = dy_ld;
  FusedBwdArgs f{static_cast<const StatAcc*>(partials), nparts, (float)count, gamma, coef, dgamma, dbeta};
  const int V = ew_vec(C);
  const int grid = stream_grid(rows, C, V);
  const size_t smem = 3 * (size_t)C * sizeof(float);
  if (V == 8) launch_pdl(bn_bwd_stats_apply_kernel<8>, dim3(grid), dim3(kEwThreads), smem, static_cast<cudaStream_t>(stream), f, a);
  else launch_pdl(bn_bwd_stats_apply_kernel<4>, dim3(grid), dim3(kEwThreads), smem, static_cast<cudaStream_t>(stream), f, a);
  return check_launch("bn_bwd_stats_apply");
}

extern "C" int b200cv_act_bwd(const void* da, int64_t da_ld, const void* aout, int64_t aout_ld, void* dz,
                              int64_t dz_ld, int64_t rows, int C, int act, float slope, void* stream) {
  B200CV_CHECK_ARG(ok_vec(da, da_ld, C) && ok_vec(aout, aout_ld, C) && ok_vec(dz, dz_ld, C) && rows > 0,
                   "act_bwd: bad args");
  act_bwd_kernel<<<ew_grid(rows * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)da, da_ld, (const bf16*)aout, aout_ld, (bf16*)dz, dz_ld, rows, C, act, slope);
  return check_launch("act_bwd");
}

extern "C" int b200cv_copy_slice(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows, int C,
                                 int accumulate, void* stream) {
  B200CV_CHECK_ARG(ok_vec(src, src_ld, C) && ok_vec(dst, dst_ld, C) && rows > 0, "copy_slice: bad args");
  copy_slice_kernel<<<ew_grid(rows * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)src, src_ld, (bf16*)dst, dst_ld, rows, C, accumulate);
  return check_launch("copy_slice");
}

extern "C" int b200cv_col_sum(const void* x, int64_t ld, int64_t rows, int C, float* out, void* stream) {
  B200CV_CHECK_ARG(ok_vec(x, ld, C) && out && rows > 0 && C / 8 <= 256, "col_sum: bad args");
  const int vpr = C / 8;
  const int rpp = 256 / vpr;
  const long long passes = (rows + rpp - 1) / rpp;
  const int grid = (int)std::max<long long>(1, std::min<long long>(passes, (long long)sm_count() * 4));
  col_sum_kernel<<<grid, 256, C * sizeof(float), static_cast<cudaStream_t>(stream)>>>((const bf16*)x, ld, rows, C, out);
  return check_launch("col_sum");
}

extern "C" int b200cv_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int stride, void* stream) {
  B200CV_CHECK_ARG(ok_vec(x, C, C) && ok_vec(y, C, C) && (stride == 1 || stride == 2), "maxpool_fwd: bad args");
  const int OH = stride == 2 ? H / 2 : H, OW = stride == 2 ? W / 2 : W;
  maxpool_fwd_kernel<<<ew_grid((long long)N * OH * OW * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, (bf16*)y, N, H, W, C, stride, OH, OW);
  return check_launch("maxpool_fwd");
}

extern "C" int b200cv_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int stride,
                                     void* stream) {
  B200CV_CHECK_ARG(ok_vec(x, C, C) && ok_vec(dy, C, C) && ok_vec(dx, C, C) && (stride == 1 || stride == 2),
                   "maxpool_bwd: bad args");
  const int OH = stride == 2 ? H / 2 : H, OW = stride == 2 ? W / 2 : W;
  if (stride == 2 && H % 2 == 0 && W % 2 == 0) {
    maxpool_bwd_s2_kernel<<<ew_grid((long long)N * OH * OW * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (const bf16*)x, (const bf16*)dy, (bf16*)dx, N, H, W, C, OH, OW);
    return check_launch("maxpool_bwd");
  }
  maxpool_bwd_kernel<<<ew_grid((long long)N * H * W * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, (const bf16*)dy, (bf16*)dx, N, H, W, C, stride, OH, OW);
  return check_launch("maxpool_bwd");
}

extern "C" int b200cv_upsample2x_fwd(const void* x, void* y, int64_t y_ld, int N, int H, int W, int C, void* stream) {
  B200CV_CHECK_ARG(ok_vec(x, C, C) && ok_vec(y, y_ld, C), "upsample_fwd: bad args");
  upsample2x_fwd_kernel<<<ew_grid((long long)N * 4 * H * W * (C / 8), 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>((const bf16*)x, (bf16*)y, y_ld, N, H, W, C);
  return check_launch("upsample_fwd");
}

extern "C" int b200cv_upsample2x_bwd(const void* dy, int64_t dy_ld, void* dx, int N, int H, int W, int C,
                                     int accumulate, void* stream) {
  B200CV_CHECK_ARG(ok_vec(dy, dy_ld, C) && ok_vec(dx, C, C), "upsample_bwd: bad args");
  upsample2x_bwd_kernel<<<ew_grid((long long)N * H * W * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const bf16*)dy, dy_ld, (bf16*)dx, N, H, W, C, accumulate);
  return check_launch("upsample_bwd");
}
