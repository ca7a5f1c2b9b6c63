// RektNet head: spatial softmax + soft-argmax (RektNet/keypoint_net.py:46-56), CrossRatioLoss
// (RektNet/cross_ratio_loss.py:20-63) and ONE fused backward kernel that goes from the loss
// straight to the gradient of the head-conv logits (location term + collinearity term ->
// soft-argmax -> softmax Jacobian), written as the NHWC bf16 operand the conv dgrad/wgrad read.
//
// The geometric term of the reference is a B x B tensordot (every pair of samples), so
//   mean_{i,j}(1 - a_i . b_j) = 1 - mean(a) . mean(b):
// it collapses to dot products of batch-mean unit vectors; d/da_i = -mean(b)/B.
#include <cuda_bf16.h>

#include <algorithm>

#include "internal.h"

namespace b200cv {
namespace {

constexpr int kKpt = 7;
// unit vectors p[i]-p[j]:     v53    v31    v10    v64    v42    v20    h21    h43    h65
__constant__ int c_vi[9] = {5, 3, 1, 6, 4, 2, 2, 4, 6};
__constant__ int c_vj[9] = {3, 1, 0, 4, 2, 0, 1, 3, 5};
// the six (1 - a.b) terms: (v31,v53) (v10,v31) (v64,v42) (v42,v20) | (h43,h21) (h65,h43)
__constant__ int c_ta[6] = {1, 2, 3, 4, 7, 8};
__constant__ int c_tb[6] = {0, 1, 4, 5, 6, 7};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += s_red[w];
  return t;
}
template <int NT>
__device__ __forceinline__ float block_max(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t = fmaxf(t, s_red[w]);
  return t;
}

// ---------------------------------------------------------------- softmax + soft-argmax
// one block per (b,k) row of H*W logits; the row lives in registers between the passes
constexpr int kSmThreads = 256;
constexpr int kSmPer = 32;  // up to 8192 pixels per heat-map
__global__ void __launch_bounds__(kSmThreads)
kpt_softmax_argmax_kernel(const float* __restrict__ logits, const float* __restrict__ vx,
                          const float* __restrict__ vy, float* __restrict__ hm, float* __restrict__ pts, int H, int W) {
  __shared__ float s_red[kSmThreads / 32];
  const int HW = H * W;
  const float* z = logits + (long long)blockIdx.x * HW;
  float v[kSmPer];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < kSmPer; ++j) {
    const int i = threadIdx.x + j * kSmThreads;
    v[j] = i < HW ? z[i] : -INFINITY;
    m = fmaxf(m, v[j]);
  }
  m = block_max<kSmThreads>(m, s_red);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kSmPer; ++j) {
    const int i = threadIdx.x + j * kSmThreads;
    v[j] = i < HW ? expf(v[j] - m) : 0.f;
    s += v[j];
  }
  s = block_sum<kSmThreads>(s, s_red);
  const float inv = 1.f / s;
  float ex = 0.f, ey = 0.f;
  float* o = hm + (long long)blockIdx.x * HW;
#pragma unroll
  for (int j = 0; j < kSmPer; ++j) {
    const int i = threadIdx.x + j * kSmThreads;
    if (i < HW) {
      const float p = v[j] * inv;
      o[i] = p;
      const int h = i / W, w = i - h * W;
      ex += p * __ldg(vx + w);
      ey += p * __ldg(vy + h);
    }
  }
  ex = block_sum<kSmThreads>(ex, s_red);
  ey = block_sum<kSmThreads>(ey, s_red);
  if (threadIdx.x == 0) {
    pts[2 * blockIdx.x] = ex;
    pts[2 * blockIdx.x + 1] = ey;
  }
}

// ---------------------------------------------------------------- loss forward
__global__ void __launch_bounds__(256)
kpt_hm_sqerr_kernel(const float* __restrict__ hm, const float* __restrict__ thm, long long n, double* acc) {
  __shared__ float s_red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* a = reinterpret_cast<const float4*>(hm);
  const float4* b = reinterpret_cast<const float4*>(thm);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) s += (hm[i] - thm[i]) * (hm[i] - thm[i]);
  s = block_sum<256>(s, s_red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)s);
}

// single block: point-based location loss, batch-mean unit vectors, geometric loss, the 3 outputs
__global__ void __launch_bounds__(256)
kpt_loss_points_kernel(const float* __restrict__ pts, const float* __restrict__ tpts, int B, int K, int loss_type,
                       int include_geo, float gamma_h, float gamma_v, const double* hm_sqerr,
                       float* __restrict__ loss3, float* __restrict__ ubar) {
  __shared__ float s_red[8];
  __shared__ float s_u[18];
  float loc = 0.f;
  if (loss_type != 1) {
    for (int i = threadIdx.x; i < B * K * 2; i += blockDim.x) {
      const float d = pts[i] - tpts[i];
      loc += loss_type == 0 ? d * d : fabsf(d);
    }
  }
  loc = block_sum<256>(loc, s_red);
  if (loss_type == 1) loc = (float)(*hm_sqerr);
  loc /= (float)B;
  float geo = 0.f;
  if (include_geo) {
    for (int e = 0; e < 9; ++e) {
      float ux = 0.f, uy = 0.f;
      for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* p = pts + (long long)b * K * 2;
        const float dx = p[2 * c_vi[e]] - p[2 * c_vj[e]];
        const float dy = p[2 * c_vi[e] + 1] - p[2 * c_vj[e] + 1];
        const float nrm = fmaxf(sqrtf(dx * dx + dy * dy), 1e-12f);  // F.normalize eps
        ux += dx / nrm;
        uy += dy / nrm;
      }
      ux = block_sum<256>(ux, s_red);
      uy = block_sum<256>(uy, s_red);
      if (threadIdx.x == 0) {
        s_u[2 * e] = ux / (float)B;
        s_u[2 * e + 1] = uy / (float)B;
      }
    }
    __syncthreads();
    float t[6];
    for (int q = 0; q < 6; ++q)
      t[q] = 1.f - (s_u[2 * c_ta[q]] * s_u[2 * c_tb[q]] + s_u[2 * c_ta[q] + 1] * s_u[2 * c_tb[q] + 1]);
    geo = gamma_h * (t[4] + t[5]) / 2.f + gamma_v * (t[0] + t[1] + t[2] + t[3]) / 4.f;
    if (threadIdx.x < 18) ubar[threadIdx.x] = s_u[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    loss3[0] = loc;
    loss3[1] = geo;
    loss3[2] = loc + geo;
  }
}

// ---------------------------------------------------------------- fused backward
// block = one image.  dpts (location + geometric + upstream) is derived in-block, then two sweeps
// over the K heat-maps of the image: (1) s_k = sum_i p_ki g_ki, (2) dlogit_ki = p_ki (g_ki - s_k)
// with g_ki = dhm_ki + dpx_k x_i + dpy_k y_i, written as NHWC bf16 rows of `ld` channels.
struct KptBwdArgs {
  const float* hm; const float* thm; const float* pts; const float* tpts; const float* ubar;
  const float* vx; const float* vy;
  const float* g_loc; const float* g_geo;       // device scalars (null = 0)
  const float* d_hm_up; const float* d_pts_up;  // optional upstream gradients of hm / pts
  int B, K, H, W, loss_type, include_geo;
  float gamma_h, gamma_v;
  __nv_bfloat16* dlogits; int ld;
};
// d(loss)/d(points) of image b into s_dp[K*2]: location term + collinearity term + upstream gradient.
// Must be called by the whole block; ends with a __syncthreads().
__device__ __forceinline__ void point_grads(const KptBwdArgs& a, int b, float gl, float gg, float* s_dp) {
  const int K = a.K;
  const float invB = 1.f / (float)a.B;
  if (threadIdx.x < K * 2) {
    const int idx = b * K * 2 + threadIdx.x;
    float d = a.d_pts_up ? a.d_pts_up[idx] : 0.f;
    const float diff = a.pts[idx] - a.tpts[idx];
    if (a.loss_type == 0) d += gl * 2.f * diff * invB;
    else if (a.loss_type == 2) d += gl * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) * invB;
    s_dp[threadIdx.x] = d;
  }
  __syncthreads();
  if (a.include_geo && threadIdx.x == 0 && gg != 0.f) {
    const float* p = a.pts + (long long)b * K * 2;
    for (int e = 0; e < 9; ++e) {
      // gradient w.r.t. the unit vector u_e of this sample: -coef * mean(partner) / B
      float gx = 0.f, gy = 0.f;
      for (int q = 0; q < 6; ++q) {
        const float coef = (q < 4 ? a.gamma_v / 4.f : a.gamma_h / 2.f) * invB * gg;
        if (c_ta[q] == e) { gx -= coef * a.ubar[2 * c_tb[q]]; gy -= coef * a.ubar[2 * c_tb[q] + 1]; }
        if (c_tb[q] == e) { gx -= coef * a.ubar[2 * c_ta[q]]; gy -= coef * a.ubar[2 * c_ta[q] + 1]; }
      }
      const float dx = p[2 * c_vi[e]] - p[2 * c_vj[e]];
      const float dy = p[2 * c_vi[e] + 1] - p[2 * c_vj[e] + 1];
      const float n = sqrtf(dx * dx + dy * dy);
      float vx_, vy_;
      if (n > 1e-12f) {  // d normalize: (g - u (u.g)) / |v|
        const float ux = dx / n, uy = dy / n;
        const float dot = ux * gx + uy * gy;
        vx_ = (gx - ux * dot) / n;
        vy_ = (gy - uy * dot) / n;
      } else {  // clamped branch: u = v / eps
        vx_ = gx / 1e-12f;
        vy_ = gy / 1e-12f;
      }
      s_dp[2 * c_vi[e]] += vx_;  s_dp[2 * c_vi[e] + 1] += vy_;
      s_dp[2 * c_vj[e]] -= vx_;  s_dp[2 * c_vj[e] + 1] -= vy_;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) kpt_head_bwd_kernel(const KptBwdArgs a) {
  __shared__ float s_red[8];
  __shared__ float s_dp[kKpt * 2];
  __shared__ float s_s[kKpt];
  const int b = blockIdx.x;
  const int K = a.K, HW = a.H * a.W;
  const float gl = a.g_loc ? *a.g_loc : 0.f;
  const float gg = a.g_geo ? *a.g_geo : 0.f;
  const float invB = 1.f / (float)a.B;
  point_grads(a, b, gl, gg, s_dp);
  const float hm_coef = a.loss_type == 1 ? gl * 2.f * invB : 0.f;
  const float* P = a.hm + (long long)b * K * HW;
  const float* Tm = a.thm ? a.thm + (long long)b * K * HW : nullptr;
  const float* U = a.d_hm_up ? a.d_hm_up + (long long)b * K * HW : nullptr;
  for (int k = 0; k < K; ++k) {
    const float dpx = s_dp[2 * k], dpy = s_dp[2 * k + 1];
    float s = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float p = P[k * HW + i];
      const int h = i / a.W, w = i - h * a.W;
      float g = dpx * __ldg(a.vx + w) + dpy * __ldg(a.vy + h);
      if (hm_coef != 0.f) g += hm_coef * (p - Tm[k * HW + i]);
      if (U) g += U[k * HW + i];
      s += p * g;
    }
    s = block_sum<256>(s, s_red);
    if (threadIdx.x == 0) s_s[k] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int h = i / a.W, w = i - h * a.W;
    const float x = __ldg(a.vx + w), y = __ldg(a.vy + h);
    float o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = 0.f;
#pragma unroll
    for (int k = 0; k < kKpt; ++k) {
      if (k < K) {
        const float p = P[k * HW + i];
        float g = s_dp[2 * k] * x + s_dp[2 * k + 1] * y;
        if (hm_coef != 0.f) g += hm_coef * (p - Tm[k * HW + i]);
        if (U) g += U[k * HW + i];
        o[k] = p * (g - s_s[k]);
      }
    }
    __nv_bfloat16* d = a.dlogits + ((long long)b * HW + i) * a.ld;
    uint4 pk[2];
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int k = 0; k < 8; ++k) hh[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
    *reinterpret_cast<uint4*>(d) = pk[0];
    *reinterpret_cast<uint4*>(d + 8) = pk[1];
    if (a.ld == 16 * kSplitPieces) {  // fp32-parity (split) rows: [p0(16) | p1(16) | p2(16)]
      float rem[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 f = __bfloat1622float2(hh[k]);
        rem[2 * k] = o[2 * k] - f.x;
        rem[2 * k + 1] = o[2 * k + 1] - f.y;
      }
#pragma unroll
      for (int pc = 1; pc < kSplitPieces; ++pc) {
        uint4 pl[2];
        __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(pl);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          ll[k] = __floats2bfloat162_rn(rem[2 * k], rem[2 * k + 1]);
          const float2 f = __bfloat1622float2(ll[k]);
          rem[2 * k] -= f.x;
          rem[2 * k + 1] -= f.y;
        }
        *reinterpret_cast<uint4*>(d + 16 * pc) = pl[0];
        *reinterpret_cast<uint4*>(d + 16 * pc + 8) = pl[1];
      }
    }
  }
}


// Un-fused form of the loss backward, for heat-maps / points that did not come from KeypointNet:
// d_pts [B,K,2] and (l2_heatmap only) d_hm = 2 g (hm - thm) / B.
__global__ void __launch_bounds__(256) kpt_loss_bwd_kernel(const KptBwdArgs a, float* d_pts, float* d_hm) {
  __shared__ float s_dp[kKpt * 2 + 2];
  const int b = blockIdx.x;
  const float gl = a.g_loc ? *a.g_loc : 0.f;
  const float gg = a.g_geo ? *a.g_geo : 0.f;
  point_grads(a, b, gl, gg, s_dp);
  if (threadIdx.x < a.K * 2) d_pts[b * a.K * 2 + threadIdx.x] = s_dp[threadIdx.x];
  if (d_hm) {
    const float coef = a.loss_type == 1 ? gl * 2.f / (float)a.B : 0.f;
    const long long n = (long long)a.K * a.H * a.W;
    const float* P = a.hm + b * n;
    const float* T = a.thm ? a.thm + b * n : nullptr;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) d_hm[b * n + i] = T ? coef * (P[i] - T[i]) : 0.f;
  }
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_kpt_softmax_argmax(const float* logits, const float* vx, const float* vy, float* hm, float* pts,
                                         int rows, int H, int W, void* stream) {
  B200CV_CHECK_ARG(logits && vx && vy && hm && pts && rows > 0 && H > 0 && W > 0, "kpt_softmax_argmax: bad args");
  B200CV_CHECK_ARG(H * W <= kSmThreads * kSmPer, "kpt_softmax_argmax: heat-map larger than %d pixels",
                   kSmThreads * kSmPer);
  kpt_softmax_argmax_kernel<<<rows, kSmThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, vx, vy, hm, pts, H, W);
  return check_launch("kpt_softmax_argmax");
}

extern "C" int b200cv_kpt_loss(const float* hm, const float* pts, const float* thm, const float* tpts, int B, int K,
                               int HW, int loss_type, int include_geo, float gamma_h, float gamma_v, double* ws,
                               float* loss3, float* ubar, void* stream) {
  B200CV_CHECK_ARG(pts && tpts && loss3 && ubar && ws && B > 0 && K > 0, "kpt_loss: bad args");
  B200CV_CHECK_ARG(loss_type >= 0 && loss_type <= 2, "kpt_loss: unknown loss type %d", loss_type);
  B200CV_CHECK_ARG(!include_geo || K == kKpt, "kpt_loss: the geometric term needs %d keypoints", kKpt);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (loss_type == 1) {
    B200CV_CHECK_ARG(hm && thm && HW > 0, "kpt_loss: l2_heatmap needs hm and target_hm");
    B200CV_CHECK_ARG(((reinterpret_cast<uintptr_t>(hm) | reinterpret_cast<uintptr_t>(thm)) & 15) == 0,
                     "kpt_loss: heat-maps must be 16-byte aligned");
    cudaMemsetAsync(ws, 0, sizeof(double), st);
    const long long n = (long long)B * K * HW;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, (long long)sm_count() * 8));
    kpt_hm_sqerr_kernel<<<grid, 256, 0, st>>>(hm, thm, n, ws);
  }
  kpt_loss_points_kernel<<<1, 256, 0, st>>>(pts, tpts, B, K, loss_type, include_geo, gamma_h, gamma_v, ws, loss3, ubar);
  return check_launch("kpt_loss");
}

extern "C" int b200cv_kpt_head_bwd(const float* hm, const float* thm, const float* pts, const float* tpts,
                                   const float* ubar, const float* vx, const float* vy, const float* g_loc,
                                   const float* g_geo, const float* d_hm_up, const float* d_pts_up, int B, int K, int H,
                                   int W, int loss_type, int include_geo, float gamma_h, float gamma_v, void* dlogits,
                                   int ld, void* stream) {
  B200CV_CHECK_ARG(hm && pts && tpts && vx && vy && dlogits, "kpt_head_bwd: null pointer");
  B200CV_CHECK_ARG(B > 0 && K > 0 && K <= kKpt && H > 0 && W > 0 && ld >= 16 && ld % 8 == 0, "kpt_head_bwd: bad shape");
  B200CV_CHECK_ARG(ld == 16 || ld == 16 * kSplitPieces, "kpt_head_bwd: dlogits rows are 16 channels (48 = split rows)");
  B200CV_CHECK_ARG(loss_type != 1 || thm, "kpt_head_bwd: l2_heatmap needs target_hm");
  B200CV_CHECK_ARG(!include_geo || (K == kKpt && ubar), "kpt_head_bwd: geometric term needs 7 keypoints and ubar");
  KptBwdArgs a{hm, thm, pts, tpts, ubar, vx, vy, g_loc, g_geo, d_hm_up, d_pts_up, B, K, H, W, loss_type, include_geo,
               gamma_h, gamma_v, static_cast<__nv_bfloat16*>(dlogits), ld};
  kpt_head_bwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("kpt_head_bwd");
}

extern "C" int b200cv_kpt_loss_bwd(const float* hm, const float* thm, const float* pts, const float* tpts,
                                   const float* ubar, const float* g_loc, const float* g_geo, int B, int K, int H, int W,
                                   int loss_type, int include_geo, float gamma_h, float gamma_v, float* d_pts,
                                   float* d_hm, void* stream) {
  B200CV_CHECK_ARG(pts && tpts && d_pts && B > 0 && K > 0 && K <= kKpt, "kpt_loss_bwd: bad args");
  B200CV_CHECK_ARG(!d_hm || hm, "kpt_loss_bwd: d_hm needs hm");
  B200CV_CHECK_ARG(!include_geo || (K == kKpt && ubar), "kpt_loss_bwd: geometric term needs 7 keypoints and ubar");
  KptBwdArgs a{hm, thm, pts, tpts, ubar, nullptr, nullptr, g_loc, g_geo, nullptr, nullptr, B, K, H, W, loss_type,
               include_geo, gamma_h, gamma_v, nullptr, 0};
  kpt_loss_bwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, d_pts, d_hm);
  return check_launch("kpt_loss_bwd");
}
