// Multi-tensor optimizer steps: the `optimizer.step()` that follows backward in
// CVC-YOLOv3/train.py:84,180-187 (torch.optim.Adam / SGD with weight_decay, momentum) and RektNet/train_eval.py:70,263
// (Adam).  One launch updates every parameter of a param group: the host passes a chunk table
// {param, grad, state1, state2, count} (device int64 [n][5]); HBM-bound streaming, 28 B (Adam) / 20 B (SGD with
// momentum) per parameter.
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace b200cv {
namespace {

struct OptChunk {
  float* p;
  const float* g;
  float* s1;
  float* s2;
  long long n;
};
static_assert(sizeof(OptChunk) == 40, "chunk table rows are 5 x int64 (ABI)");

template <typename F>
__device__ __forceinline__ void for_each4(const OptChunk& c, F f) {
  const long long n = c.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) |
                     reinterpret_cast<uintptr_t>(c.s1) | reinterpret_cast<uintptr_t>(c.s2)) & 15) == 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) f(i * 4, 4);
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) f(i, 1);
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) f(i, 1);
  }
}

// torch.optim.Adam (amsgrad=False, maximize=False), single-tensor formulation of torch 2.x:
//   g += wd*p; m = lerp(m, g, 1-b1); v = v*b2 + (1-b2)*g*g; p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_multi_kernel(const OptChunk* __restrict__ table, float step_size, float beta1, float omb1, float beta2,
                  float omb2, float eps, float weight_decay, float bc2_sqrt) {
  const OptChunk c = table[blockIdx.x];
  auto upd = [&](float& p, float g, float& m, float& v) {
    if (weight_decay != 0.f) g = g + weight_decay * p;
    m = m + (g - m) * omb1;
    v = v * beta2 + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
  };
  for_each4(c, [&](long long i, int w) {
    if (w == 4) {
      float4 p = *reinterpret_cast<float4*>(c.p + i);
      const float4 g = __ldcs(reinterpret_cast<const float4*>(c.g + i));
      float4 m = *reinterpret_cast<float4*>(c.s1 + i);
      float4 v = *reinterpret_cast<float4*>(c.s2 + i);
      upd(p.x, g.x, m.x, v.x);
      upd(p.y, g.y, m.y, v.y);
      upd(p.z, g.z, m.z, v.z);
      upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(c.p + i) = p;
      *reinterpret_cast<float4*>(c.s1 + i) = m;
      *reinterpret_cast<float4*>(c.s2 + i) = v;
    } else {
      float p = c.p[i], m = c.s1[i], v = c.s2[i];
      upd(p, c.g[i], m, v);
      c.p[i] = p;
      c.s1[i] = m;
      c.s2[i] = v;
    }
  });
}

// torch.optim.SGD (dampening=0, nesterov=False): g += wd*p; buf = first ? g : mu*buf + g; p -= lr*buf
__global__ void __launch_bounds__(256)
sgd_multi_kernel(const OptChunk* __restrict__ table, float lr, float momentum, float weight_decay, int first_step) {
  const OptChunk c = table[blockIdx.x];
  const bool has_buf = c.s1 != nullptr;
  auto upd = [&](float& p, float g, float& buf) {
    if (weight_decay != 0.f) g = g + weight_decay * p;
    if (has_buf) {
      buf = first_step ? g : momentum * buf + g;
      g = buf;
    }
    p = p - lr * g;
  };
  for_each4(c, [&](long long i, int w) {
    if (w == 4) {
      float4 p = *reinterpret_cast<float4*>(c.p + i);
      const float4 g = __ldcs(reinterpret_cast<const float4*>(c.g + i));
      float4 m = has_buf ? *reinterpret_cast<float4*>(c.s1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      upd(p.x, g.x, m.x);
      upd(p.y, g.y, m.y);
      upd(p.z, g.z, m.z);
      upd(p.w, g.w, m.w);
      *reinterpret_cast<float4*>(c.p + i) = p;
      if (has_buf) *reinterpret_cast<float4*>(c.s1 + i) = m;
    } else {
      float p = c.p[i], m = has_buf ? c.s1[i] : 0.f;
      upd(p, c.g[i], m);
      c.p[i] = p;
      if (has_buf) c.s1[i] = m;
    }
  });
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_adam_step_multi(const int64_t* table, int n_chunks, float step_size, double beta1, double beta2,
                                      float eps, float weight_decay, float bias_correction2_sqrt, void* stream) {
  B200CV_CHECK_ARG(table || n_chunks == 0, "adam_step_multi: null table");
  B200CV_CHECK_ARG(bias_correction2_sqrt > 0.f, "adam_step_multi: bias_correction2_sqrt must be > 0");
  if (n_chunks <= 0) return B200CV_OK;
  adam_multi_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptChunk*>(table), step_size, (float)beta1, (float)(1.0 - beta1), (float)beta2,
      (float)(1.0 - beta2), eps, weight_decay, bias_correction2_sqrt);  // 1 - beta in double, like torch's scalars
  return check_launch("adam_step_multi");
}

extern "C" int b200cv_sgd_step_multi(const int64_t* table, int n_chunks, float lr, float momentum,
                                     float weight_decay, int first_step, void* stream) {
  B200CV_CHECK_ARG(table || n_chunks == 0, "sgd_step_multi: null table");
  if (n_chunks <= 0) return B200CV_OK;
  sgd_multi_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptChunk*>(table), lr, momentum, weight_decay, first_step);
  return check_launch("sgd_step_multi");
}
