// YOLO head: target assignment, multi-part loss (forward + gradient w.r.t. the raw head logits)
// and eval-mode decode.  Compiled with -fmad=false: grid indices and anchor IoUs must reproduce
// the reference's fp32 arithmetic exactly (CVC-YOLOv3/utils/utils.py:195-275, :163-193).
//
// Target assignment never materialises the reference's eight dense [B,A,G,G] tensors on the hot
// path; it produces
//   rec[b,t]   : per-target record (cell, best anchor, tx ty tw th, label)
//   owner[b,a,gj,gi] : index t of the target that owns the cell (last writer in (b,t) order wins,
//                      like index_put_ on CPU), -1 = no object
//   ign[gj,gi] : 1 if ANY target of ANY image has an anchor IoU > thresh in that cell -- the
//                reference clears conf_mask[:, :, gj, gi] for all images and anchors
//   counts     : N_m = #mask cells, N_f = #(conf_mask - mask) cells  (the loss means' divisors)
// The dense tensors are produced only by b200cv_yolo_targets_dense (the public build_targets()).
#include <cuda_bf16.h>

#include <algorithm>

#include "internal.h"

namespace b200cv {
namespace {

struct __align__(16) YoloRec {
  int gi, gj, best, label;
  float tx, ty, tw, th;
};
static_assert(sizeof(YoloRec) == 32, "YoloRec must be 32 bytes (ABI: rec is float[B*T*8])");

__global__ void yolo_targets_kernel(const float* __restrict__ targets, const float* __restrict__ anchors, int B,
                                    int T, int A, int Gh, int Gw, float thres, int* __restrict__ owner,
                                    unsigned char* __restrict__ ign, YoloRec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  const float* row = targets + (size_t)i * 5;
  const float* row0 = targets + (size_t)b * T * 5;
  const float s = (((row[0] + row[1]) + row[2]) + row[3]) + row[4];
  const bool valid = s > 0.f;
  const float* src = valid ? row : row0;  // padding rows become copies of row 0 (utils.py:223-228)
  const float gx = src[1] * (float)Gw;
  const float gy = src[2] * (float)Gh;
  const float gw = src[3] * (float)Gw;
  const float gh = src[4] * (float)Gh;
  const int gi = (int)gx;  // .long(): truncation toward zero
  const int gj = (int)gy;
  // anchor IoU in the corner-format branch of bbox_iou with the "+1 pixel" convention
  float best_iou = -1.f;
  int best = 0;
  bool over = false;
  const float b1_area = ((gw - 0.f) + 1.f) * ((gh - 0.f) + 1.f);
  for (int a = 0; a < A; ++a) {
    const float aw = anchors[2 * a], ah = anchors[2 * a + 1];
    const float ix = fmaxf((fminf(gw, aw) - 0.f) + 1.f, 0.f);
    const float iy = fmaxf((fminf(gh, ah) - 0.f) + 1.f, 0.f);
    const float inter = ix * iy;
    const float b2_area = ((aw - 0.f) + 1.f) * ((ah - 0.f) + 1.f);
    const float iou = inter / (((b1_area + b2_area) - inter) + 1e-12f);
    if (iou > thres) over = true;
    if (iou > best_iou) {  // strict: first maximum wins, like torch.argmax
      best_iou = iou;
      best = a;
    }
  }
  YoloRec r;
  r.gi = gi;
  r.gj = gj;
  r.best = best;
  r.label = (int)row[0];  // the row's own label (0 for padding rows), utils.py:271
  r.tx = gx - (float)gi;
  r.ty = gy - (float)gj;
  r.tw = logf(gw / anchors[2 * best] + 1e-16f);
  r.th = logf(gh / anchors[2 * best + 1] + 1e-16f);
  const bool inside = gi >= 0 && gi < Gw && gj >= 0 && gj < Gh;  // the reference raises IndexError otherwise
  if (!inside) r.best = -1;
  rec[i] = r;
  if (inside) {
    atomicMax(&owner[((b * A + best) * Gh + gj) * Gw + gi], t);
    if (over) ign[gj * Gw + gi] = 1;
  }
}

__global__ void yolo_counts_kernel(const int* __restrict__ owner, const unsigned char* __restrict__ ign,
                                   const YoloRec* __restrict__ rec, int B, int T, int A, int Gh, int Gw,
                                   int* __restrict__ counts) {
  __shared__ int s_nm, s_win_ign, s_S;
  if (threadIdx.x == 0) { s_nm = 0; s_win_ign = 0; s_S = 0; }
  __syncthreads();
  int nm = 0, wi = 0, S = 0;
  for (int i = threadIdx.x; i < B * T; i += blockDim.x) {
    const YoloRec r = rec[i];
    if (r.best < 0) continue;
    const int b = i / T, t = i - b * T;
    if (owner[((b * A + r.best) * Gh + r.gj) * Gw + r.gi] == t) {
      ++nm;
      if (ign[r.gj * Gw + r.gi]) ++wi;
    }
  }
  for (int i = threadIdx.x; i < Gh * Gw; i += blockDim.x) S += ign[i] ? 1 : 0;
  atomicAdd(&s_nm, nm);
  atomicAdd(&s_win_ign, wi);
  atomicAdd(&s_S, S);
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long total = (long long)B * A * Gh * Gw;
    counts[0] = s_nm;
    counts[1] = (int)(total - (long long)B * A * s_S - (s_nm - s_win_ign));
  }
}

__global__ void fill_i32_kernel(int* p, int v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// loss terms (CVC-YOLOv3/models.py:150-155,199-211).  BCE is nn.BCELoss on fp32 PROBABILITIES:
// sigma first, then log clamped at -100; backward (sigma - t)/max(sigma(1-sigma),1e-12)*sigma(1-sigma).
struct LossConsts {
  float xy, wh, obj, noobj;
};
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

struct CellInfo {
  bool m, cf;      // object cell / no-object (conf_false) cell
  float t[4];      // tx ty tw th
};
__device__ __forceinline__ CellInfo load_cell(const int* owner, const unsigned char* ign, const YoloRec* rec, int T,
                                              int b, int a, int gy, int gx, int A, int Gh, int Gw) {
  CellInfo c;
  const int own = owner[((b * A + a) * Gh + gy) * Gw + gx];
  c.m = own >= 0;
  c.cf = !c.m && !ign[gy * Gw + gx];
  if (c.m) {
    const YoloRec r = rec[b * T + own];
    c.t[0] = r.tx; c.t[1] = r.ty; c.t[2] = r.tw; c.t[3] = r.th;
  } else {
    c.t[0] = c.t[1] = c.t[2] = c.t[3] = 0.f;
  }
  return c;
}
// returns d(loss)/d(logit) for attribute `attr` (0..4) and accumulates the un-normalised loss sums
__device__ __forceinline__ float attr_term(int attr, float z, const CellInfo& c, const LossConsts& k, float inv_nm,
                                           float inv_nf, float (&acc)[6]) {
  if (attr < 2) {
    if (!c.m) return 0.f;
    const float s = sigmoidf_(z);
    const float d = s - (attr == 0 ? c.t[0] : c.t[1]);
    if (attr == 0) acc[0] += d * d; else acc[1] += d * d;
    return k.xy * 2.f * inv_nm * d * (s * (1.f - s));
  }
  if (attr < 4) {
    if (!c.m) return 0.f;
    const float d = z - (attr == 2 ? c.t[2] : c.t[3]);
    if (attr == 2) acc[2] += d * d; else acc[3] += d * d;
    return k.wh * 2.f * inv_nm * d;
  }
  if (!c.m && !c.cf) return 0.f;
  const float s = sigmoidf_(z);
  const float ds = s * (1.f - s);
  if (c.m) {
    acc[4] += -fmaxf(logf(s), -100.f);
    return k.obj * inv_nm * ((s - 1.f) / fmaxf(ds, 1e-12f)) * ds;
  }
  acc[5] += -fmaxf(logf(1.f - s), -100.f);
  return k.noobj * inv_nf * (s / fmaxf(ds, 1e-12f)) * ds;
}

__device__ __forceinline__ void block_accumulate(float (&acc)[6], double* sums) {
  __shared__ float s_part[6][8];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)s_part[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}

// Channel-contiguous layout (the engine's): logits fp32 [pixels][z_ld], dlogits [pixels][d_ld] with
// channel = a*(5+C)+attr.  One thread owns 8 consecutive channels of one pixel: 16/32-byte stores,
// loads only where an x,y,w,h,conf channel falls into its range (class logits are never read: the
// class term has weight 0, models.py:205).
template <typename OutT>
__global__ void __launch_bounds__(256)
yolo_loss_nhwc_kernel(const float* __restrict__ z, long long z_ld, OutT* __restrict__ dl, long long d_ld, int d_ch,
                      int B, int A, int C, int Gh, int Gw, const int* __restrict__ owner,
                      const unsigned char* __restrict__ ign, const YoloRec* __restrict__ rec, int T,
                      const int* __restrict__ counts, LossConsts k, double* sums, const float* gscale) {
  // The grid is sized so that (threads per block) % vpp == 0: a thread keeps the SAME 8 channels for its whole
  // life, classifies them once, and only walks over pixels -- no per-item divisions by (5+C), and the threads
  // that own nothing but class/pad channels just stream zeros.
  const int nattr = 5 + C;
  const int nch = A * nattr;
  const int vpp = d_ch >> 3;
  const int tpb = (blockDim.x / vpp) * vpp;  // active threads per block
  float acc[6] = {0, 0, 0, 0, 0, 0};
  if ((int)threadIdx.x < tpb) {
    const int v = threadIdx.x % vpp;
    const int c0 = v << 3;
    int an[8], at[8];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = c0 + j;
      an[j] = ch < nch ? ch / nattr : -1;
      at[j] = ch < nch ? ch - an[j] * nattr : 99;
      any |= at[j] < 5;
    }
    const int pix_per_block = tpb / vpp;
    const int npix = B * Gh * Gw;
    const float g = gscale ? *gscale : 1.f;
    const float inv_nm = 1.f / (float)counts[0];
    const float inv_nf = 1.f / (float)counts[1];
    const int GG = Gh * Gw;
    for (int pix = blockIdx.x * pix_per_block + threadIdx.x / vpp; pix < npix; pix += gridDim.x * pix_per_block) {
      float out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (any) {
        const int b = pix / GG;
        const int rem = pix - b * GG;
        const int gy = rem / Gw;
        const int gx = rem - gy * Gw;
        int a_cached = -1;
        CellInfo cell;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (at[j] < 5) {
            if (an[j] != a_cached) {
              cell = load_cell(owner, ign, rec, T, b, an[j], gy, gx, A, Gh, Gw);
              a_cached = an[j];
            }
            if (cell.m || (at[j] == 4 && cell.cf))
              out[j] = g * attr_term(at[j], __ldg(z + (long long)pix * z_ld + c0 + j), cell, k, inv_nm, inv_nf, acc);
          }
        }
      }
      if (dl) {
        OutT* o = dl + (long long)pix * d_ld + c0;
        if constexpr (sizeof(OutT) == 2) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(out[2 * j], out[2 * j + 1]);
          *reinterpret_cast<uint4*>(o) = pk;
        } else {
          *reinterpret_cast<float4*>(o) = make_float4(out[0], out[1], out[2], out[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(out[4], out[5], out[6], out[7]);
        }
      }
    }
  }
  if (sums) block_accumulate(acc, sums);
}

// Arbitrary strides (the public YOLOLayer on an NCHW fp32 `sample`): one thread per (b,a,gy,gx) cell,
// writes only the five non-zero gradient channels (the caller zero-fills dlogits first).
__global__ void __launch_bounds__(256)
yolo_loss_strided_kernel(const float* __restrict__ z, long long z_sb, long long z_sy, long long z_sx, long long z_sc,
                         float* __restrict__ dl, long long d_sb, long long d_sy, long long d_sx, long long d_sc,
                         long long d_sa, int B, int A, int C, int Gh, int Gw, const int* __restrict__ owner,
                         const unsigned char* __restrict__ ign, const YoloRec* __restrict__ rec, int T,
                         const int* __restrict__ counts, LossConsts k, double* sums, const float* gscale) {
  const int nattr = 5 + C;
  const long long total = (long long)B * A * Gh * Gw;
  const float g = gscale ? *gscale : 1.f;
  const float inv_nm = 1.f / (float)counts[0];
  const float inv_nf = 1.f / (float)counts[1];
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % Gw);
    long long r = i / Gw;
    const int gy = (int)(r % Gh); r /= Gh;
    const int a = (int)(r % A);
    const int b = (int)(r / A);
    const CellInfo cell = load_cell(owner, ign, rec, T, b, a, gy, gx, A, Gh, Gw);
    if (!cell.m && !cell.cf) {
      if (dl && d_sa == 5)  // compact cell gradients are not pre-zeroed by the caller
        for (int attr = 0; attr < 5; ++attr) dl[b * d_sb + gy * d_sy + gx * d_sx + (long long)a * 5 + attr] = 0.f;
      continue;
    }
    const long long zoff = b * z_sb + gy * z_sy + gx * z_sx + (long long)a * nattr * z_sc;
    const long long doff = b * d_sb + gy * d_sy + gx * d_sx + (long long)a * d_sa;
    if (!cell.m && dl && d_sa == 5)
      for (int attr = 0; attr < 4; ++attr) dl[doff + attr] = 0.f;
    for (int attr = cell.m ? 0 : 4; attr < 5; ++attr) {
      const float d = attr_term(attr, z[zoff + attr * z_sc], cell, k, inv_nm, inv_nf, acc);
      if (dl) dl[doff + attr * d_sc] = d * g;
    }
  }
  if (sums) block_accumulate(acc, sums);
}

// Dense head gradient from the compact per-cell gradients: row = pixel, d_ch channels, channel a*(5+C)+attr takes
// dcell[pix][a*5+attr] for attr < 5 and an exact zero otherwise.  Pure streaming: a thread owns 8 fixed channels.
template <typename OutT>
__global__ void __launch_bounds__(256)
yolo_expand_kernel(const float* __restrict__ dcell, int cell_ld, OutT* __restrict__ dl, long long d_ld, int d_ch,
                   int npix, int A, int C) {
  const int nattr = 5 + C;
  const int nch = A * nattr;
  const int vpp = d_ch >> 3;
  const int tpb = (blockDim.x / vpp) * vpp;
  if ((int)threadIdx.x >= tpb) return;
  const int c0 = (threadIdx.x % vpp) << 3;
  int src[8];
  bool any = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c0 + j;
    const int a = ch < nch ? ch / nattr : 0;
    const int attr = ch < nch ? ch - a * nattr : 99;
    src[j] = attr < 5 ? a * 5 + attr : -1;
    any |= src[j] >= 0;
  }
  const int pix_per_block = tpb / vpp;
  for (int pix = blockIdx.x * pix_per_block + threadIdx.x / vpp; pix < npix; pix += gridDim.x * pix_per_block) {
    float out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (any) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (src[j] >= 0) out[j] = __ldg(dcell + (long long)pix * cell_ld + src[j]);
    }
    OutT* o = dl + (long long)pix * d_ld + c0;
    if constexpr (sizeof(OutT) == 2) {
      uint4 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(out[2 * j], out[2 * j + 1]);
      *reinterpret_cast<uint4*>(o) = pk;
    } else {
      *reinterpret_cast<float4*>(o) = make_float4(out[0], out[1], out[2], out[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(out[4], out[5], out[6], out[7]);
    }
  }
}

// out7[0] += total, out7[1..6] += (x, y, w, h, obj, noobj)  -- the order of models.py:211
__global__ void yolo_loss_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ counts,
                                          LossConsts k, float* __restrict__ out7) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float nm = (float)counts[0], nf = (float)counts[1];
  const float lx = k.xy * ((float)sums[0] / nm);
  const float ly = k.xy * ((float)sums[1] / nm);
  const float lw = k.wh * ((float)sums[2] / nm);
  const float lh = k.wh * ((float)sums[3] / nm);
  const float lobj = k.obj * ((float)sums[4] / nm);
  const float lnoobj = k.noobj * ((float)sums[5] / nf);
  const float total = lx + ly + lw + lh + lnoobj + lobj;  // + 0 * class term
  out7[0] += total;
  out7[1] += lx; out7[2] += ly; out7[3] += lw; out7[4] += lh; out7[5] += lobj; out7[6] += lnoobj;
}

// ---------------------------------------------------------------------------------------------
// eval-mode decode (models.py:150-169,213-220): rows ordered (a, gy, gx); boxes scaled by stride.
__global__ void yolo_decode_kernel(const float* __restrict__ z, long long z_sb, long long z_sy, long long z_sx,
                                   long long z_sc, int B, int A, int C, int Gh, int Gw,
                                   const float* __restrict__ anchors, float stride, float* __restrict__ out,
                                   long long out_sb, long long row0) {
  const int nattr = 5 + C;
  const long long total = (long long)B * A * Gh * Gw * nattr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int attr = (int)(i % nattr);
    long long r = i / nattr;
    const int gx = (int)(r % Gw); r /= Gw;
    const int gy = (int)(r % Gh); r /= Gh;
    const int a = (int)(r % A);
    const int b = (int)(r / A);
    const float v = z[b * z_sb + gy * z_sy + gx * z_sx + ((long long)a * nattr + attr) * z_sc];
    float o;
    if (attr == 0) o = (sigmoidf_(v) + (float)gx) * stride;
    else if (attr == 1) o = (sigmoidf_(v) + (float)gy) * stride;
    else if (attr == 2) o = (expf(v) * anchors[2 * a]) * stride;
    else if (attr == 3) o = (expf(v) * anchors[2 * a + 1]) * stride;
    else o = sigmoidf_(v);
    out[b * out_sb + (row0 + ((long long)a * Gh + gy) * Gw + gx) * nattr + attr] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// dense expansion for the public build_targets(): mask, conf_mask (u8), tx ty tw th tconf (f32), tcls (u8)
__global__ void yolo_dense_cells_kernel(const int* __restrict__ owner, const unsigned char* __restrict__ ign,
                                        const YoloRec* __restrict__ rec, int B, int T, int A, int Gh, int Gw,
                                        unsigned char* mask, unsigned char* conf_mask, float* tx, float* ty, float* tw,
                                        float* th, float* tconf) {
  const long long total = (long long)B * A * Gh * Gw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % Gw);
    long long r = i / Gw;
    const int gy = (int)(r % Gh); r /= Gh;
    const int b = (int)(r / A);
    const int own = owner[i];
    const bool m = own >= 0;
    mask[i] = m;
    conf_mask[i] = m || !ign[gy * Gw + gx];
    float v[4] = {0, 0, 0, 0};
    if (m) {
      const YoloRec q = rec[b * T + own];
      v[0] = q.tx; v[1] = q.ty; v[2] = q.tw; v[3] = q.th;
    }
    tx[i] = v[0]; ty[i] = v[1]; tw[i] = v[2]; th[i] = v[3];
    tconf[i] = m ? 1.f : 0.f;
  }
}
__global__ void yolo_dense_cls_kernel(const YoloRec* __restrict__ rec, int B, int T, int A, int C, int Gh, int Gw,
                                      unsigned char* tcls) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const YoloRec r = rec[i];
  if (r.best < 0 || r.label < 0 || r.label >= C) return;
  const int b = i / T;
  tcls[((((long long)b * A + r.best) * Gh + r.gj) * Gw + r.gi) * C + r.label] = 1;  // every (b,t) sets its bit
}

int grid1d(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, (long long)sm_count() * 8));
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_yolo_targets(const float* targets, const float* anchors_scaled, int B, int T, int A, int Gh,
                                   int Gw, float ignore_thres, int32_t* owner, uint8_t* ign, float* rec,
                                   int32_t* counts, void* stream) {
  B200CV_CHECK_ARG(targets && anchors_scaled && owner && ign && rec && counts, "yolo_targets: null pointer");
  B200CV_CHECK_ARG(B > 0 && T > 0 && A > 0 && Gh > 0 && Gw > 0, "yolo_targets: empty shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long cells = (long long)B * A * Gh * Gw;
  fill_i32_kernel<<<grid1d(cells, 256), 256, 0, st>>>(owner, -1, cells);
  cudaMemsetAsync(ign, 0, (size_t)Gh * Gw, st);
  yolo_targets_kernel<<<(B * T + 127) / 128, 128, 0, st>>>(targets, anchors_scaled, B, T, A, Gh, Gw, ignore_thres,
                                                          owner, ign, reinterpret_cast<YoloRec*>(rec));
  yolo_counts_kernel<<<1, 256, 0, st>>>(owner, ign, reinterpret_cast<const YoloRec*>(rec), B, T, A, Gh, Gw, counts);
  return check_launch("yolo_targets");
}

extern "C" int b200cv_yolo_loss(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc, int B,
                                int A, int C, int Gh, int Gw, const int32_t* owner, const uint8_t* ign,
                                const float* rec, int T, const int32_t* counts, float xy_loss, float wh_loss,
                                float obj_loss, float noobj_loss, double* sums, void* dlogits, int dl_dtype,
                                int64_t d_sb, int64_t d_sy, int64_t d_sx, int64_t d_sc, int d_channels,
                                const float* gscale, void* stream) {
  B200CV_CHECK_ARG(logits && owner && ign && rec && counts, "yolo_loss: null pointer");
  B200CV_CHECK_ARG(sums || dlogits, "yolo_loss: nothing to compute");
  B200CV_CHECK_ARG(B > 0 && A > 0 && C >= 0 && Gh > 0 && Gw > 0 && T > 0, "yolo_loss: empty shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const LossConsts k{xy_loss, wh_loss, obj_loss, noobj_loss};
  const YoloRec* r = reinterpret_cast<const YoloRec*>(rec);
  const int nch = A * (5 + C);
  const bool z_nhwc = z_sc == 1 && z_sy == (int64_t)Gw * z_sx && z_sb == (int64_t)Gh * z_sy;
  const bool d_nhwc = !dlogits || (d_sc == 1 && d_sy == (int64_t)Gw * d_sx && d_sb == (int64_t)Gh * d_sy &&
                                   d_channels % 8 == 0 && d_channels >= nch && d_channels <= d_sx &&
                                   (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0 && d_sx % 8 == 0);
  if (z_nhwc && d_nhwc) {
    const int dch = dlogits ? d_channels : round_up(nch, 8);
    B200CV_CHECK_ARG(dch / 8 <= 256, "yolo_loss: more than 2048 head channels");
    const long long work = (long long)B * Gh * Gw * (dch / 8);
    if (!dlogits || dl_dtype == B200CV_DT_BF16)
      yolo_loss_nhwc_kernel<__nv_bfloat16><<<grid1d(work, 256), 256, 0, st>>>(
          logits, z_sx, static_cast<__nv_bfloat16*>(dlogits), d_sx, dch, B, A, C, Gh, Gw, owner, ign, r, T, counts, k,
          sums, gscale);
    else
      yolo_loss_nhwc_kernel<float><<<grid1d(work, 256), 256, 0, st>>>(logits, z_sx, static_cast<float*>(dlogits), d_sx,
                                                                      dch, B, A, C, Gh, Gw, owner, ign, r, T, counts,
                                                                      k, sums, gscale);
    return check_launch("yolo_loss_nhwc");
  }
  B200CV_CHECK_ARG(!dlogits || dl_dtype == B200CV_DT_F32, "yolo_loss: strided dlogits must be fp32 (zero-filled)");
  yolo_loss_strided_kernel<<<grid1d((long long)B * A * Gh * Gw, 256), 256, 0, st>>>(
      logits, z_sb, z_sy, z_sx, z_sc, static_cast<float*>(dlogits), d_sb, d_sy, d_sx, d_sc, (long long)(5 + C) * d_sc,
      B, A, C, Gh, Gw, owner, ign, r, T, counts, k, sums, gscale);
  return check_launch("yolo_loss_strided");
}

// Two-kernel form of the head gradient used by the engine: (1) one thread per anchor cell (massively parallel,
// hides the owner -> record -> logit load chain) writes compact fp32 cell gradients [pixels][cell_ld >= 5A];
// (2) a streaming kernel expands them into the dense NHWC dlogits rows the head conv's dgrad/wgrad read.
extern "C" int b200cv_yolo_loss_cells(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc,
                                      int B, int A, int C, int Gh, int Gw, const int32_t* owner, const uint8_t* ign,
                                      const float* rec, int T, const int32_t* counts, float xy_loss, float wh_loss,
                                      float obj_loss, float noobj_loss, double* sums, float* dcell, int cell_ld,
                                      const float* gscale, void* stream) {
  B200CV_CHECK_ARG(logits && owner && ign && rec && counts && (sums || dcell), "yolo_loss_cells: null pointer");
  B200CV_CHECK_ARG(B > 0 && A > 0 && C >= 0 && Gh > 0 && Gw > 0 && T > 0 && (!dcell || cell_ld >= 5 * A),
                   "yolo_loss_cells: bad shape");
  const LossConsts k{xy_loss, wh_loss, obj_loss, noobj_loss};
  yolo_loss_strided_kernel<<<grid1d((long long)B * A * Gh * Gw, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, z_sb, z_sy, z_sx, z_sc, dcell, (long long)Gh * Gw * cell_ld, (long long)Gw * cell_ld, cell_ld, 1, 5, B,
      A, C, Gh, Gw, owner, ign, reinterpret_cast<const YoloRec*>(rec), T, counts, k, sums, gscale);
  return check_launch("yolo_loss_cells");
}

extern "C" int b200cv_yolo_expand_dlogits(const float* dcell, int cell_ld, void* dlogits, int dl_dtype, int64_t d_ld,
                                          int d_channels, int64_t npix, int A, int C, void* stream) {
  B200CV_CHECK_ARG(dcell && dlogits && npix > 0 && npix < (1ll << 31) && A > 0 && cell_ld >= 5 * A,
                   "yolo_expand_dlogits: bad args");
  B200CV_CHECK_ARG(d_channels % 8 == 0 && d_channels >= A * (5 + C) && d_channels <= d_ld && d_channels / 8 <= 256 &&
                       (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0 && d_ld % 8 == 0,
                   "yolo_expand_dlogits: bad dlogits layout");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid1d(npix * (d_channels / 8), 256);
  if (dl_dtype == B200CV_DT_BF16)
    yolo_expand_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dcell, cell_ld, static_cast<__nv_bfloat16*>(dlogits), d_ld,
                                                            d_channels, (int)npix, A, C);
  else
    yolo_expand_kernel<float><<<grid, 256, 0, st>>>(dcell, cell_ld, static_cast<float*>(dlogits), d_ld, d_channels,
                                                    (int)npix, A, C);
  return check_launch("yolo_expand_dlogits");
}

extern "C" int b200cv_yolo_loss_finalize(const double* sums, const int32_t* counts, float xy_loss, float wh_loss,
                                         float obj_loss, float noobj_loss, float* out7, void* stream) {
  B200CV_CHECK_ARG(sums && counts && out7, "yolo_loss_finalize: null pointer");
  yolo_loss_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, counts, LossConsts{xy_loss, wh_loss, obj_loss, noobj_loss}, out7);
  return check_launch("yolo_loss_finalize");
}

extern "C" int b200cv_yolo_decode(const float* logits, int64_t z_sb, int64_t z_sy, int64_t z_sx, int64_t z_sc, int B,
                                  int A, int C, int Gh, int Gw, const float* anchors_scaled, float stride, float* out,
                                  int64_t out_batch_stride, int64_t row_offset, void* stream) {
  B200CV_CHECK_ARG(logits && anchors_scaled && out && B > 0 && A > 0 && Gh > 0 && Gw > 0, "yolo_decode: bad args");
  const long long total = (long long)B * A * Gh * Gw * (5 + C);
  yolo_decode_kernel<<<grid1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, z_sb, z_sy, z_sx, z_sc, B, A, C, Gh, Gw, anchors_scaled, stride, out, out_batch_stride, row_offset);
  return check_launch("yolo_decode");
}

extern "C" int b200cv_yolo_targets_dense(const int32_t* owner, const uint8_t* ign, const float* rec, int B, int T,
                                         int A, int C, int Gh, int Gw, uint8_t* mask, uint8_t* conf_mask, float* tx,
                                         float* ty, float* tw, float* th, float* tconf, uint8_t* tcls, void* stream) {
  B200CV_CHECK_ARG(owner && ign && rec && mask && conf_mask && tx && ty && tw && th && tconf && tcls,
                   "yolo_targets_dense: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long cells = (long long)B * A * Gh * Gw;
  const YoloRec* r = reinterpret_cast<const YoloRec*>(rec);
  yolo_dense_cells_kernel<<<grid1d(cells, 256), 256, 0, st>>>(owner, ign, r, B, T, A, Gh, Gw, mask, conf_mask, tx, ty,
                                                              tw, th, tconf);
  cudaMemsetAsync(tcls, 0, (size_t)cells * C, st);
  yolo_dense_cls_kernel<<<(B * T + 127) / 128, 128, 0, st>>>(r, B, T, A, C, Gh, Gw, tcls);
  return check_launch("yolo_targets_dense");
}
