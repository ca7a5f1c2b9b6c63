// Detection post-processing and the detect -> RektNet joint (BASELINE config 5):
//   * detect_nms_kernel: confidence filter + (cx,cy,w,h) -> corners (CVC-YOLOv3/detect.py:84-90) and the greedy
//     top-k NMS of CVC-YOLOv3/utils/nms.py:4-61, one CTA per image, no host round trip;
//   * detect_compact_kernel: per-image keep counts -> a flat crop list;
//   * crop_resize_kernel: `cv2.resize(frame[y0:y1, x0:x1], (w,h))` (8-bit INTER_LINEAR, the fixed-point algorithm of
//     OpenCV's resize.cpp) + HWC->CHW + /255.0 of RektNet/utils.py:73-76, RektNet/detect.py:32-34.
// Compiled with -fmad=false: box corners, areas, IoUs and the interpolation tables must reproduce the reference's
// fp32/fp64 arithmetic exactly (kept-index sets and crop bytes are compared bit-for-bit).
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace b200cv {
namespace {

constexpr int kNmsThreads = 512;
constexpr int kNmsMaxK = 512;  // top_k limit: the suppression bit matrix is kNmsMaxK x kNmsMaxK/32 words of smem

// Order-preserving map float -> uint32 (any sign), so that (score, row) pairs compare as one 64-bit integer.
__device__ __forceinline__ unsigned long long nms_key(float conf, int row) {
  const unsigned u = __float_as_uint(conf);
  const unsigned m = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)m << 32) | (unsigned)row;
}
__device__ __forceinline__ float nms_key_score(unsigned long long k) {
  const unsigned m = (unsigned)(k >> 32);
  return __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}

// One CTA per image.  Order of visit = (score descending, row descending): the reversed stable ascending sort.
__global__ void __launch_bounds__(kNmsThreads)
detect_nms_kernel(const float* __restrict__ det, long long batch_stride, int rows, int row_len, float conf_thres,
                  float nms_thres, int top_k, int corners, float* __restrict__ boxes_out, float* __restrict__ scores_out,
                  int* __restrict__ rows_out, int* __restrict__ counts) {
  __shared__ unsigned hist[256];  // radix-select histogram; reused as the keep list (shorts) after the selection
  __shared__ unsigned long long keys[kNmsMaxK];
  __shared__ float4 box[kNmsMaxK];
  __shared__ float area[kNmsMaxK];
  __shared__ unsigned supp[kNmsMaxK][kNmsMaxK / 32];
  static_assert(sizeof(hist) >= kNmsMaxK * sizeof(short), "keep list must fit in the histogram");
  short* keep = reinterpret_cast<short*>(hist);
  __shared__ unsigned s_cnt, s_need, s_nkeep;
  __shared__ unsigned long long s_prefix;

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* d = det + (size_t)b * batch_stride;

  // ---- candidates: conf > thres
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  unsigned mine = 0;
  for (int r = tid; r < rows; r += kNmsThreads) mine += d[(size_t)r * row_len + 4] > conf_thres;
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  const int n_cand = (int)s_cnt;
  const int n_sel = n_cand < top_k ? n_cand : top_k;

  // ---- the top_k-th largest key by MSB-first radix select (keys are unique: the row is part of the key)
  unsigned long long thresh_key = 0;
  if (n_cand > top_k) {
    if (tid == 0) {
      s_prefix = 0;
      s_need = (unsigned)top_k;
    }
    unsigned long long known = 0;  // mask of the key bits already fixed
    for (int pass = 7; pass >= 0; --pass) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned need = s_need;
      const int shift = pass * 8;
      for (int r = tid; r < rows; r += kNmsThreads) {
        const float c = d[(size_t)r * row_len + 4];
        if (c > conf_thres) {
          const unsigned long long k = nms_key(c, r);
          if ((k & known) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
      }
      __syncthreads();
      if (tid < 32) {
        // lane l owns digits 255-8l .. 248-8l (descending); suffix counts by a shuffle scan over the lanes
        unsigned h[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          h[j] = hist[255 - 8 * tid - j];
          s += h[j];
        }
        unsigned incl = s;
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += v;
        }
        unsigned above = incl - s;  // keys with a larger digit than this lane's
        if (above < need && need <= incl) {  // exactly one lane: the digit of the need-th largest key is here
          for (int j = 0; j < 8; ++j) {
            if (need <= above + h[j]) {
              s_prefix = prefix | ((unsigned long long)(255 - 8 * tid - j) << shift);
              s_need = need - above;
              break;
            }
            above += h[j];
          }
        }
      }
      known |= 0xffull << shift;
      __syncthreads();
    }
    thresh_key = s_prefix;
  }

  // ---- gather the selected keys, sort them descending (bitonic, padded with zeros)
  int P = 1;
  while (P < n_sel) P <<= 1;
  for (int i = tid; i < P; i += kNmsThreads) keys[i] = 0;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int r = tid; r < rows; r += kNmsThreads) {
    const float c = d[(size_t)r * row_len + 4];
    if (c > conf_thres) {
      const unsigned long long k = nms_key(c, r);
      if (k >= thresh_key) {
        const unsigned slot = atomicAdd(&s_cnt, 1u);
        if (slot < (unsigned)kNmsMaxK) keys[slot] = k;
      }
    }
  }
  __syncthreads();
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += kNmsThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c = keys[l];
          const bool desc = (i & k2) == 0;
          if (desc ? (a < c) : (a > c)) {
            keys[i] = c;
            keys[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- corner boxes and areas (detect.py:86-89, nms.py:19-23)
  for (int i = tid; i < n_sel; i += kNmsThreads) {
    const int r = (int)(unsigned)(keys[i] & 0xffffffffull);
    const float* q = d + (size_t)r * row_len;
    float4 bx;
    if (corners) {
      bx = make_float4(q[0], q[1], q[2], q[3]);
    } else {
      const float hw = q[2] / 2.f, hh = q[3] / 2.f;
      bx.x = q[0] - hw;
      bx.y = q[1] - hh;
      bx.z = q[0] + hw;
      bx.w = q[1] + hh;
    }
    box[i] = bx;
    area[i] = (bx.z - bx.x) * (bx.w - bx.y);
  }
  __syncthreads();

  // ---- suppression bits: supp[i] bit j (j > i) = box i, once kept, removes box j  (nms.py:43-60)
  const int words = (n_sel + 31) >> 5;
  for (int item = tid; item < n_sel * words; item += kNmsThreads) {
    const int i = item / words, w = item - i * words;
    const float4 bi = box[i];
    const float ai = area[i];
    unsigned bits = 0;
    const int j0 = w << 5;
    for (int t = 0; t < 32; ++t) {
      const int j = j0 + t;
      if (j > i && j < n_sel) {
        const float4 bj = box[j];
        const float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);
        const float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
        const float iw = fmaxf(xx2 - xx1, 0.f), ih = fmaxf(yy2 - yy1, 0.f);
        const float inter = iw * ih;
        const float uni = (area[j] - inter) + ai;
        const float iou = inter / uni;
        if (!(iou <= nms_thres)) bits |= 1u << t;  // NaN is suppressed, like IoU.le(overlap)
      }
    }
    supp[i][w] = bits;
  }
  __syncthreads();

  // ---- greedy scan in score order by one warp (lane l holds removed-word l; kNmsMaxK/32 = 16 words)
  if (tid < 32) {
    unsigned removed = 0;
    unsigned nk = 0;
    for (int i = 0; i < n_sel; ++i) {
      const unsigned rw = __shfl_sync(0xffffffffu, removed, i >> 5);
      if (!((rw >> (i & 31)) & 1u)) {
        if (tid == 0) keep[nk] = (short)i;
        ++nk;
        if (tid < words) removed |= supp[i][tid];
      }
    }
    if (tid == 0) s_nkeep = nk;
  }
  __syncthreads();

  const int nk = (int)s_nkeep;
  for (int s = tid; s < top_k; s += kNmsThreads) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int r = -1;
    if (s < nk) {
      const int i = keep[s];
      bx = box[i];
      sc = nms_key_score(keys[i]);
      r = (int)(unsigned)(keys[i] & 0xffffffffull);
    }
    reinterpret_cast<float4*>(boxes_out)[(size_t)b * top_k + s] = bx;
    scores_out[(size_t)b * top_k + s] = sc;
    rows_out[(size_t)b * top_k + s] = r;
  }
  if (tid == 0) counts[b] = nk;
}

// counts[B] -> offsets[B+1] (exclusive scan; offsets[B] = total) and src[n] = (image, slot) of crop n.
__global__ void detect_compact_kernel(const int* __restrict__ counts, int B, int top_k, int* __restrict__ offsets,
                                      int2* __restrict__ src) {
  __shared__ int s_off[1025];
  const int tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      s_off[b] = acc;
      acc += counts[b];
    }
    s_off[B] = acc;
  }
  __syncthreads();
  for (int b = tid; b <= B; b += blockDim.x) offsets[b] = s_off[b];
  for (int b = 0; b < B; ++b) {
    const int n = s_off[b + 1] - s_off[b];
    for (int s = tid; s < n; s += blockDim.x) src[s_off[b] + s] = make_int2(b, s);
  }
}

// resize.cpp's table entry for destination index d: source index and the two 11-bit weights.
__device__ __forceinline__ void linear_tab(int d, int src, int dst, bool clamp, int* s_idx, short* a0, short* a1) {
  const double scale = (double)src / (double)dst;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int i = (int)floorf(f);
  f = f - (float)i;
  if (clamp && i < 0) {
    i = 0;
    f = 0.f;
  }
  if (clamp && i >= src - 1) {
    i = src - 1;
    f = 0.f;
  }
  *s_idx = i;
  // saturate_cast<short>(float) = round half to even
  *a0 = (short)__float2int_rn((1.f - f) * 2048.f);
  *a1 = (short)__float2int_rn(f * 2048.f);
}

constexpr int kCropMaxDim = 256;

// One CTA per crop.  frames u8 [B][H][W][3]; out fp32 [n][3][dh][dw]; rects int32 [n][4] = x0,y0,x1,y1.
__global__ void __launch_bounds__(256)
crop_resize_kernel(const unsigned char* __restrict__ frames, int H, int W, const float* __restrict__ boxes, int top_k,
                   const int2* __restrict__ src, const float* __restrict__ geom, int geom_stride, int dw, int dh,
                   float* __restrict__ out, int* __restrict__ rects) {
  __shared__ int sx[kCropMaxDim], sy[kCropMaxDim];
  __shared__ short ax0[kCropMaxDim], ax1[kCropMaxDim], ay0[kCropMaxDim], ay1[kCropMaxDim];
  __shared__ float lut[256];
  __shared__ int rect[4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int2 bs = src[n];
  if (tid == 0) {
    const float* bx = boxes + ((size_t)bs.x * top_k + bs.y) * 4;
    const float* g = geom + (size_t)bs.x * geom_stride;  // ratio, pad_w, pad_h (per image, or shared when stride 0)
    const float x0 = bx[0] / g[0] - g[1], y0 = bx[1] / g[0] - g[2];
    const float x1 = bx[2] / g[0] - g[1], y1 = bx[3] / g[0] - g[2];
    // float -> int with clamping done in float first (boxes can be far outside the frame or NaN)
    const float fx0 = fminf(fmaxf(floorf(x0), 0.f), (float)(W - 1));
    const float fy0 = fminf(fmaxf(floorf(y0), 0.f), (float)(H - 1));
    const int ix0 = (fx0 == fx0) ? (int)fx0 : 0, iy0 = (fy0 == fy0) ? (int)fy0 : 0;
    const float fx1 = fminf(fmaxf(ceilf(x1), (float)(ix0 + 1)), (float)W);
    const float fy1 = fminf(fmaxf(ceilf(y1), (float)(iy0 + 1)), (float)H);
    rect[0] = ix0;
    rect[1] = iy0;
    rect[2] = (fx1 == fx1) ? (int)fx1 : ix0 + 1;
    rect[3] = (fy1 == fy1) ? (int)fy1 : iy0 + 1;
  }
  lut[tid] = (float)((double)tid / 255.0);  // (u8 / 255.0).astype(float32)
  __syncthreads();
  const int x0 = rect[0], y0 = rect[1];
  const int cw = rect[2] - x0, ch = rect[3] - y0;
  if (tid < 4) rects[(size_t)n * 4 + tid] = rect[tid];
  for (int d = tid; d < dw; d += 256) linear_tab(d, cw, dw, true, &sx[d], &ax0[d], &ax1[d]);
  for (int d = tid; d < dh; d += 256) linear_tab(d, ch, dh, false, &sy[d], &ay0[d], &ay1[d]);
  __syncthreads();
  const unsigned char* f = frames + ((size_t)bs.x * H + y0) * W * 3 + (size_t)x0 * 3;
  const size_t rowb = (size_t)W * 3;
  float* o = out + (size_t)n * 3 * dh * dw;
  const int plane = dh * dw;
  const bool same = (cw == dw && ch == dh);
  const bool half = (cw == 2 * dw && ch == 2 * dh);  // resize.cpp: exact 2x decimation runs the INTER_AREA kernel
  for (int e = tid; e < 3 * plane; e += 256) {
    const int c = e / plane, p = e - c * plane;
    const int dy = p / dw, dx = p - dy * dw;
    int v;
    if (same) {
      v = f[dy * rowb + dx * 3 + c];
    } else if (half) {
      const unsigned char* q = f + (size_t)(2 * dy) * rowb + (2 * dx) * 3 + c;
      v = (q[0] + q[3] + q[rowb] + q[rowb + 3] + 2) >> 2;
    } else {
      const int xa = sx[dx], xb = min(xa + 1, cw - 1);
      const int ya = min(max(sy[dy], 0), ch - 1), yb = min(max(sy[dy] + 1, 0), ch - 1);
      const int a0 = ax0[dx], a1 = ax1[dx];
      const unsigned char* r0 = f + (size_t)ya * rowb + c;
      const unsigned char* r1 = f + (size_t)yb * rowb + c;
      const int S0 = r0[xa * 3] * a0 + r0[xb * 3] * a1;
      const int S1 = r1[xa * 3] * a0 + r1[xb * 3] * a1;
      v = ((((int)ay0[dy] * (S0 >> 4)) >> 16) + (((int)ay1[dy] * (S1 >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
    }
    o[e] = lut[v];
  }
}


// Per-image detection metric of CVC-YOLOv3/validate.py:80-128 + utils/utils.py:58-119 (average_precision, compute_ap)
// on the NMS output (already in descending-score order; the reference's two re-sorts by -conf are stable no-ops).
// One CTA per image: threads find each detection's best target (bbox_iou with the "+1 pixel" convention, first maximum
// wins), thread 0 runs the greedy matching and the precision/recall sweep.  Single class, like the reference.
constexpr int kApMaxT = 4096;

__global__ void __launch_bounds__(128)
detect_match_ap_kernel(const float* __restrict__ boxes, const int* __restrict__ counts, int top_k,
                       const float* __restrict__ targets, int T, float width, float height, float iou_thres,
                       float* __restrict__ ap_out, float* __restrict__ r_out, float* __restrict__ p_out,
                       int* __restrict__ valid, unsigned char* __restrict__ correct_out) {
  __shared__ float best_iou[kNmsMaxK];
  __shared__ int best_t[kNmsMaxK];
  __shared__ unsigned char detected[kApMaxT];
  __shared__ unsigned char correct[kNmsMaxK];
  __shared__ float prec[kNmsMaxK + 2];
  __shared__ int s_ngt;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = counts[b];
  const float* tg = targets + (size_t)b * T * 5;
  if (tid == 0) s_ngt = 0;
  for (int t = tid; t < T; t += blockDim.x) detected[t] = 0;
  __syncthreads();
  // labels[(labels[:, 1:5] <= 0).sum(dim=1) == 0]: rows whose four box numbers are all positive
  int mine = 0;
  for (int t = tid; t < T; t += blockDim.x) {
    const float* q = tg + (size_t)t * 5;
    mine += (q[1] > 0.f && q[2] > 0.f && q[3] > 0.f && q[4] > 0.f);
  }
  if (mine) atomicAdd(&s_ngt, mine);
  __syncthreads();
  const int n_gt = s_ngt;
  for (int i = tid; i < top_k; i += blockDim.x) correct_out[(size_t)b * top_k + i] = 0;
  if (n == 0 || n_gt == 0) {  // validate.py:95-96 (no detection) and :117-118 (no label): the image is skipped
    if (tid == 0) {
      ap_out[b] = 0.f;
      r_out[b] = 0.f;
      p_out[b] = 0.f;
      valid[b] = 0;
    }
    return;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    const float4 d = reinterpret_cast<const float4*>(boxes)[(size_t)b * top_k + i];
    const float d_area = ((d.z - d.x) + 1.f) * ((d.w - d.y) + 1.f);
    float bi = -1.f;
    int bt = 0;
    bool first = true;
    for (int t = 0; t < T; ++t) {
      const float* q = tg + (size_t)t * 5;
      if (!(q[1] > 0.f && q[2] > 0.f && q[3] > 0.f && q[4] > 0.f)) continue;
      // xywh2xyxy (utils.py:121-126), then *= width / height (validate.py:105-106)
      const float tx1 = (q[1] - q[3] / 2.f) * width, ty1 = (q[2] - q[4] / 2.f) * height;
      const float tx2 = (q[1] + q[3] / 2.f) * width, ty2 = (q[2] + q[4] / 2.f) * height;
      const float iw = fmaxf((fminf(d.z, tx2) - fmaxf(d.x, tx1)) + 1.f, 0.f);
      const float ih = fmaxf((fminf(d.w, ty2) - fmaxf(d.y, ty1)) + 1.f, 0.f);
      const float inter = iw * ih;
      const float t_area = ((tx2 - tx1) + 1.f) * ((ty2 - ty1) + 1.f);
      const float iou = inter / (((d_area + t_area) - inter) + 1e-12f);
      if (first || iou > bi) {  // torch.argmax: first maximum
        bi = iou;
        bt = t;
        first = false;
      }
    }
    best_iou[i] = bi;
    best_t[i] = bt;
  }
  __syncthreads();
  if (tid == 0) {
    // greedy matching in score order (validate.py:123-127)
    for (int i = 0; i < n; ++i) {
      const int t = best_t[i];
      const bool ok = best_iou[i] > iou_thres && !detected[t];
      correct[i] = ok;
      if (ok) detected[t] = 1;
    }
    // average_precision (utils.py:58-89): cumulative TP / FP, recall = tpc / n_gt, precision = tpc / (tpc + fpc)
    const float ngt = (float)n_gt;  // n_gt + 1e-16 == n_gt
    float tpc = 0.f, fpc = 0.f;
    prec[0] = 0.f;
    for (int i = 0; i < n; ++i) {
      tpc += correct[i] ? 1.f : 0.f;
      fpc += correct[i] ? 0.f : 1.f;
      prec[i + 1] = tpc / (tpc + fpc);
    }
    prec[n + 1] = 0.f;
    const float r = tpc / ngt, p = tpc / (tpc + fpc);
    // compute_ap (utils.py:91-119): precision envelope from the right, area under the steps of the recall curve
    for (int i = n + 1; i > 0; --i) prec[i - 1] = fmaxf(prec[i - 1], prec[i]);
    double ap = 0.0;
    float prev_rec = 0.f, tp2 = 0.f;
    for (int k = 0; k <= n; ++k) {  // mrec[k+1] vs mrec[k]; mrec = [0, recall..., 1]
      float rec_next;
      if (k < n) {
        tp2 += correct[k] ? 1.f : 0.f;
        rec_next = tp2 / ngt;
      } else {
        rec_next = 1.f;
      }
      if (rec_next != prev_rec) ap += (double)((rec_next - prev_rec) * prec[k + 1]);
      prev_rec = rec_next;
    }
    ap_out[b] = (float)ap;
    r_out[b] = r;
    p_out[b] = p;
    valid[b] = 1;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) correct_out[(size_t)b * top_k + i] = correct[i];
}


// Letterbox front end of CVC-YOLOv3/detect.py:62-72: pad to the network's aspect ratio with 127 (torchvision pad),
// PIL BILINEAR resize (Pillow's 8-bit two-pass convolution resampler, src/libImaging/Resample.c: 22-bit fixed-point
// coefficients, horizontal pass rounded to u8, then the vertical pass), to_tensor (/255).  The coefficient tables depend
// only on the geometry and come from the host (b200cv/preprocess.py), so the kernel is pure integer arithmetic and
// bit-exact with Pillow.  One thread per output pixel, three channels; the padded image is never materialised.
constexpr int kPilBits = 22;

__global__ void __launch_bounds__(256)
letterbox_kernel(const unsigned char* __restrict__ frames, int H, int W, int pad_w, int pad_h, int fill, int reverse,
                 const int* __restrict__ hx_min, const int* __restrict__ hx_cnt, const int* __restrict__ hx_k, int ksh,
                 const int* __restrict__ vy_min, const int* __restrict__ vy_cnt, const int* __restrict__ vy_k, int ksv,
                 int out_w, int out_h, float* __restrict__ out) {
  __shared__ float lut[256];
  lut[threadIdx.x] = (float)threadIdx.x / 255.f;  // to_tensor: uint8 -> float32, .div(255)
  __syncthreads();
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= out_w * out_h) return;
  const int yy = e / out_w, xx = e - yy * out_w;
  const unsigned char* f = frames + (size_t)b * H * W * 3;
  const int x0 = hx_k ? hx_min[xx] : xx, xn = hx_k ? hx_cnt[xx] : 1;
  const int y0 = vy_k ? vy_min[yy] : yy, yn = vy_k ? vy_cnt[yy] : 1;
  const int* kh = hx_k ? hx_k + (size_t)xx * ksh : nullptr;
  const int* kv = vy_k ? vy_k + (size_t)yy * ksv : nullptr;
  int accv[3] = {1 << (kPilBits - 1), 1 << (kPilBits - 1), 1 << (kPilBits - 1)};
  int last[3] = {0, 0, 0};
  for (int r = 0; r < yn; ++r) {
    const int fy = y0 + r - pad_h;
    const bool row_in = fy >= 0 && fy < H;
    int hv[3];
    if (kh) {
      int acc[3] = {1 << (kPilBits - 1), 1 << (kPilBits - 1), 1 << (kPilBits - 1)};
      for (int x = 0; x < xn; ++x) {
        const int fx = x0 + x - pad_w;
        const int k = kh[x];
        if (row_in && fx >= 0 && fx < W) {
          const unsigned char* q = f + ((size_t)fy * W + fx) * 3;
          acc[0] += q[0] * k;
          acc[1] += q[1] * k;
          acc[2] += q[2] * k;
        } else {
          acc[0] += fill * k;
          acc[1] += fill * k;
          acc[2] += fill * k;
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) hv[c] = min(max(acc[c] >> kPilBits, 0), 255);
    } else {
      const int fx = x0 - pad_w;
      if (row_in && fx >= 0 && fx < W) {
        const unsigned char* q = f + ((size_t)fy * W + fx) * 3;
        hv[0] = q[0];
        hv[1] = q[1];
        hv[2] = q[2];
      } else {
        hv[0] = hv[1] = hv[2] = fill;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      last[c] = hv[c];
      if (kv) accv[c] += hv[c] * kv[r];
    }
  }
  const size_t plane = (size_t)out_h * out_w;
  float* o = out + (size_t)b * 3 * plane + e;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v = kv ? min(max(accv[c] >> kPilBits, 0), 255) : last[c];
    o[(size_t)(reverse ? 2 - c : c) * plane] = lut[v];
  }
}

// Tile-and-scale input pipeline (CVC-YOLOv3/utils/datasets.py:143-159): scale_image (PIL LANCZOS resize of the whole
// frame) -> pad 127 up to the patch size -> crop patch `patch_index` -> to_tensor, for a batch of frames in one launch.
// Output pixel (xx, yy) of image b is pixel (off_x[b] + xx, off_y[b] + yy) of the scaled frame (offsets = rounded patch
// origin minus the pad): inside the scaled frame it is Pillow's two-pass 8-bit resample (horizontal pass rounded to u8,
// then vertical; coefficient tables from the host, b200cv/tiler.py), outside it is the fill value.  Neither the scaled
// frame nor the padded one is materialised.
__global__ void __launch_bounds__(256)
tile_scale_kernel(const unsigned char* __restrict__ frames, int H, int W, int new_w, int new_h, int fill,
                  const int* __restrict__ off_x, const int* __restrict__ off_y, const int* __restrict__ hx_min,
                  const int* __restrict__ hx_cnt, const int* __restrict__ hx_k, int ksh,
                  const int* __restrict__ vy_min, const int* __restrict__ vy_cnt, const int* __restrict__ vy_k, int ksv,
                  int out_w, int out_h, float* __restrict__ out) {
  __shared__ float lut[256];
  lut[threadIdx.x] = (float)threadIdx.x / 255.f;  // to_tensor
  __syncthreads();
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= out_w * out_h) return;
  const int yy = e / out_w, xx = e - yy * out_w;
  const int sx = off_x[b] + xx, sy = off_y[b] + yy;  // position in the scaled frame
  const size_t plane = (size_t)out_h * out_w;
  float* o = out + (size_t)b * 3 * plane + e;
  if (sx < 0 || sx >= new_w || sy < 0 || sy >= new_h) {
    o[0] = o[plane] = o[2 * plane] = lut[fill];
    return;
  }
  const unsigned char* f = frames + (size_t)b * H * W * 3;
  const int x0 = hx_k ? hx_min[sx] : sx, xn = hx_k ? hx_cnt[sx] : 1;
  const int y0 = vy_k ? vy_min[sy] : sy, yn = vy_k ? vy_cnt[sy] : 1;
  const int* kh = hx_k ? hx_k + (size_t)sx * ksh : nullptr;
  const int* kv = vy_k ? vy_k + (size_t)sy * ksv : nullptr;
  int accv[3] = {1 << (kPilBits - 1), 1 << (kPilBits - 1), 1 << (kPilBits - 1)};
  int last[3] = {0, 0, 0};
  for (int r = 0; r < yn; ++r) {
    const unsigned char* row = f + (size_t)(y0 + r) * W * 3;
    int hv[3];
    if (kh) {
      int acc[3] = {1 << (kPilBits - 1), 1 << (kPilBits - 1), 1 << (kPilBits - 1)};
      for (int x = 0; x < xn; ++x) {
        const unsigned char* q = row + (size_t)(x0 + x) * 3;
        const int k = kh[x];
        acc[0] += q[0] * k;
        acc[1] += q[1] * k;
        acc[2] += q[2] * k;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) hv[c] = min(max(acc[c] >> kPilBits, 0), 255);
    } else {
      const unsigned char* q = row + (size_t)x0 * 3;
      hv[0] = q[0];
      hv[1] = q[1];
      hv[2] = q[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      last[c] = hv[c];
      if (kv) accv[c] += hv[c] * kv[r];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v = kv ? min(max(accv[c] >> kPilBits, 0), 255) : last[c];
    o[(size_t)c * plane] = lut[v];
  }
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_tile_scale_u8(const uint8_t* frames, int B, int H, int W, int new_w, int new_h, int fill,
                                    const int32_t* off_x, const int32_t* off_y, const int32_t* hx_min,
                                    const int32_t* hx_cnt, const int32_t* hx_k, int ksize_h, const int32_t* vy_min,
                                    const int32_t* vy_cnt, const int32_t* vy_k, int ksize_v, int out_w, int out_h,
                                    float* out, void* stream) {
  if (B == 0) return B200CV_OK;
  B200CV_CHECK_ARG(frames && off_x && off_y && out && B > 0 && H > 0 && W > 0 && new_w > 0 && new_h > 0 && out_w > 0 &&
                       out_h > 0 && fill >= 0 && fill <= 255,
                   "tile_scale_u8: bad args");
  B200CV_CHECK_ARG(hx_k ? (hx_min && hx_cnt && ksize_h > 0) : (W == new_w),
                   "tile_scale_u8: no horizontal tables although the frame is rescaled");
  B200CV_CHECK_ARG(vy_k ? (vy_min && vy_cnt && ksize_v > 0) : (H == new_h),
                   "tile_scale_u8: no vertical tables although the frame is rescaled");
  const dim3 grid((out_w * out_h + 255) / 256, B);
  tile_scale_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, H, W, new_w, new_h, fill, off_x, off_y, hx_min, hx_cnt, hx_k, ksize_h, vy_min, vy_cnt, vy_k, ksize_v,
      out_w, out_h, out);
  return check_launch("tile_scale_u8");
}

extern "C" int b200cv_detect_nms(const float* det, int64_t det_batch_stride, int B, int rows, int row_len,
                                 int box_format, float conf_thres, float nms_thres, int top_k, float* boxes,
                                 float* scores, int32_t* det_rows, int32_t* counts, void* stream) {
  if (B == 0) return B200CV_OK;
  B200CV_CHECK_ARG(det && boxes && scores && det_rows && counts, "detect_nms: null pointer");
  B200CV_CHECK_ARG(B >= 0 && rows >= 0 && row_len >= 5, "detect_nms: bad shape B=%d rows=%d row_len=%d", B, rows,
                   row_len);
  B200CV_CHECK_ARG(top_k >= 1 && top_k <= kNmsMaxK, "detect_nms: top_k=%d outside [1,%d]", top_k, kNmsMaxK);
  B200CV_CHECK_ARG(box_format == 0 || box_format == 1, "detect_nms: box_format must be 0 (centre) or 1 (corners)");
  B200CV_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "detect_nms: boxes must be 16-byte aligned");
  detect_nms_kernel<<<B, kNmsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      det, det_batch_stride, rows, row_len, conf_thres, nms_thres, top_k, box_format, boxes, scores, det_rows,
      counts);
  return check_launch("detect_nms");
}

extern "C" int b200cv_detect_compact(const int32_t* counts, int B, int top_k, int32_t* offsets, int32_t* src,
                                     void* stream) {
  B200CV_CHECK_ARG(counts && offsets && src, "detect_compact: null pointer");
  B200CV_CHECK_ARG(B >= 0 && B <= 1024, "detect_compact: B=%d outside [0,1024]", B);
  detect_compact_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(counts, B, top_k, offsets,
                                                                         reinterpret_cast<int2*>(src));
  return check_launch("detect_compact");
}

extern "C" int b200cv_crop_resize_u8(const uint8_t* frames, int B, int H, int W, const float* boxes, int top_k,
                                     const int32_t* src, int n_crops, const float* geom, int geom_stride, int out_w,
                                     int out_h, float* out, int32_t* rects, void* stream) {
  if (n_crops <= 0) return B200CV_OK;
  B200CV_CHECK_ARG(frames && boxes && src && geom && out && rects, "crop_resize_u8: null pointer");
  B200CV_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "crop_resize_u8: bad frame shape");
  B200CV_CHECK_ARG(out_w >= 1 && out_w <= kCropMaxDim && out_h >= 1 && out_h <= kCropMaxDim,
                   "crop_resize_u8: output size %dx%d outside [1,%d]", out_w, out_h, kCropMaxDim);
  B200CV_CHECK_ARG(geom_stride == 0 || geom_stride >= 3, "crop_resize_u8: geom_stride must be 0 or >= 3");
  if (n_crops <= 0) return B200CV_OK;
  crop_resize_kernel<<<n_crops, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, H, W, boxes, top_k, reinterpret_cast<const int2*>(src), geom, geom_stride, out_w, out_h, out, rects);
  return check_launch("crop_resize_u8");
}

extern "C" int b200cv_detect_match_ap(const float* boxes, const int32_t* counts, int B, int top_k,
                                      const float* targets, int T, float width, float height, float iou_thres,
                                      float* ap, float* recall, float* precision, int32_t* valid, uint8_t* correct,
                                      void* stream) {
  if (B == 0) return B200CV_OK;
  B200CV_CHECK_ARG(boxes && counts && targets && ap && recall && precision && valid && correct,
                   "detect_match_ap: null pointer");
  B200CV_CHECK_ARG(top_k >= 1 && top_k <= kNmsMaxK, "detect_match_ap: top_k=%d outside [1,%d]", top_k, kNmsMaxK);
  B200CV_CHECK_ARG(T >= 0 && T <= kApMaxT, "detect_match_ap: T=%d outside [0,%d]", T, kApMaxT);
  B200CV_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "detect_match_ap: boxes must be 16-byte aligned");
  detect_match_ap_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, counts, top_k, targets, T, width, height, iou_thres, ap, recall, precision, valid, correct);
  return check_launch("detect_match_ap");
}

extern "C" int b200cv_letterbox_u8(const uint8_t* frames, int B, int H, int W, int pad_w, int pad_h, int fill,
                                   int reverse_channels, const int32_t* hx_min, const int32_t* hx_cnt,
                                   const int32_t* hx_k, int ksize_h, const int32_t* vy_min, const int32_t* vy_cnt,
                                   const int32_t* vy_k, int ksize_v, int out_w, int out_h, float* out, void* stream) {
  if (B == 0) return B200CV_OK;
  B200CV_CHECK_ARG(frames && out, "letterbox_u8: null pointer");
  B200CV_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_w > 0 && out_h > 0 && pad_w >= 0 && pad_h >= 0,
                   "letterbox_u8: bad shape");
  B200CV_CHECK_ARG(fill >= 0 && fill <= 255, "letterbox_u8: fill must be a byte");
  B200CV_CHECK_ARG(hx_k ? (hx_min && hx_cnt && ksize_h > 0) : (W + 2 * pad_w == out_w),
                   "letterbox_u8: no horizontal tables although the padded width differs from the output width");
  B200CV_CHECK_ARG(vy_k ? (vy_min && vy_cnt && ksize_v > 0) : (H + 2 * pad_h == out_h),
                   "letterbox_u8: no vertical tables although the padded height differs from the output height");
  const dim3 grid((out_w * out_h + 255) / 256, B);
  letterbox_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, H, W, pad_w, pad_h, fill, reverse_channels, hx_min, hx_cnt, hx_k, ksize_h, vy_min, vy_cnt, vy_k, ksize_v,
      out_w, out_h, out);
  return check_launch("letterbox_u8");
}
