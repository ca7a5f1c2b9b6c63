// C-ABI convolution entry points: turn a conv description into the tap table / bounding box /
// output strides of the implicit-GEMM core (conv_igemm.cu), plus the weight/layout pack kernels.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "stat_acc.cuh"

namespace b200cv {

namespace {

int pick_block_n(int cout, int num_m_tiles, int kiters, bool bn_reduce) {
  int bn = 16;
  while (bn < cout && bn < 256) bn *= 2;
  // Short-K data gradients with the fused BatchNorm-backward reduction (the 1x1 layers): 128-wide tiles leave room
  // for a deeper y ring in shared memory (measured 7.64 -> 7.52 ms over the dgrads of a Darknet-53 step); override
  // with B200CV_SHORTK_BN=<width> (applies to every short-K GEMM).
  static const int shortk_bn = getenv("B200CV_SHORTK_BN") ? atoi(getenv("B200CV_SHORTK_BN")) : 0;
  if (shortk_bn ? (kiters <= 8 && bn > shortk_bn) : (bn_reduce && kiters <= 8 && bn > 128)) return shortk_bn ? shortk_bn : 128;
  // Prefer 128-wide tiles when 256-wide ones leave the last wave mostly empty.
  if (bn == 256) {
    const int sms = sm_count();
    auto waste = [&](int b) {
      const long long tiles = (long long)num_m_tiles * ((cout + b - 1) / b);
      const long long waves = (tiles + sms - 1) / sms;
      return (double)(waves * sms - tiles) / (double)(waves * sms);
    };
    if (cout % 256 != 0 && cout % 128 == 0) bn = 128;
    else if (waste(256) > 0.25 && waste(128) < waste(256)) bn = 128;
  }
  return bn;
}

struct Geometry {
  // activation tensor the im2col map walks
  int N, H, W, C;
  int lower_w, lower_h, upper_w, upper_h, trav_w, trav_h;
  int OHt, OWt;  // traversal grid (rows of D per image = OHt*OWt)
};

// fp32-parity mode: every tap becomes six k-passes over the split operands x = x0 + x1 + x2, w = w0 + w1 + w2 (bf16
// pieces of decreasing magnitude): all products x_i * w_j with i + j <= 2, smallest first; what is dropped
// (i + j >= 3) is below 2^-24 of the result, fp32's own rounding.  A packed row holds [w0(Cin) | w1(Cin) | w2(Cin)]
// per tap.
int expand_split_taps(IgemmParams& p, int cin) {
  static const int kPassA[6] = {2, 0, 1, 1, 0, 0};
  static const int kPassW[6] = {0, 2, 1, 0, 1, 0};
  const int n = p.num_taps;
  if (6 * n > kMaxTaps) return set_error(B200CV_ERR_ARG, "conv: %d taps do not fit the fp32-parity mode", n);
  for (int t = n - 1; t >= 0; --t) {
    const short tw = p.tap_w[t], th = p.tap_h[t];
    const int k3 = kSplitPieces * p.tap_k[t];
    for (int ps = 5; ps >= 0; --ps) {
      const int i = ps * n + t;
      p.tap_w[i] = tw;
      p.tap_h[i] = th;
      p.tap_c[i] = (short)(kPassA[ps] * cin);
      p.tap_k[i] = k3 + kPassW[ps] * cin;
    }
  }
  p.num_taps = 6 * n;
  return 0;
}

int run_igemm(const Geometry& g, const void* act, const void* wpk, int64_t w_rows, int64_t w_cols,
              IgemmParams& p, cudaStream_t stream, const void* bn_y = nullptr, int64_t bn_y_ld = 0,
              bool split_in = false) {
  const int kc = kc_for(g.C);
  if (split_in) {
    if (int rc = expand_split_taps(p, g.C)) return rc;
    w_cols *= kSplitPieces;
  }
  const int cmul = split_in ? kSplitPieces : 1;  // channels of a stored activation row per logical channel
  p.cblocks = g.C / kc;
  p.OHW = g.OHt * g.OWt;
  p.OW = g.OWt;
  p.M_total = g.N * p.OHW;
  p.lower_w = g.lower_w;
  p.lower_h = g.lower_h;
  p.trav_w = g.trav_w;
  p.trav_h = g.trav_h;
  p.num_m_tiles = (p.M_total + 127) / 128;
  const int bn = pick_block_n(p.Cout, p.num_m_tiles, p.num_taps * p.cblocks, p.bn_sums != nullptr);
  p.num_n_tiles = (p.Cout + bn - 1) / bn;
  p.err = device_error_word();
  if (!p.d2s_c2) {
    p.stat_cols = p.Cout;
    p.stat_mask = -1;
  }
  {
    static const int dbg = getenv("B200CV_DBG") ? atoi(getenv("B200CV_DBG")) : 0;
    p.dbg = dbg;
  }
  CUtensorMap tmA, tmB;
  int rc = make_tmap_im2col_bf16(&tmA, act, g.N, g.H, g.W, cmul * g.C, cmul * g.C, (int64_t)g.W * cmul * g.C,
                                 (int64_t)g.H * g.W * cmul * g.C, g.lower_w, g.lower_h, g.upper_w, g.upper_h,
                                 g.trav_w, g.trav_h, kc, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, wpk, w_rows, w_cols, w_cols, bn, kc);
  if (rc) return rc;
  // Residual added by the tensor core (D += I * R as extra k-iterations) when the epilogue is a pure "+ residual":
  // no scale / activation, the residual a row-major bf16 matrix [M][ld] walked exactly like the output rows.
  CUtensorMap tmI, tmR;
  const CUtensorMap* pI = nullptr;
  const CUtensorMap* pR = nullptr;
  static const bool no_res_mma = getenv("B200CV_NO_RES_MMA") != nullptr;
  if (p.res && !no_res_mma && !split_in && !p.res_lo && !p.out_lo && bn % 64 == 0 && !p.scale && p.act == 0 &&
      p.res_vec_ok && p.r_sc == 1 &&
      p.r_sw % 8 == 0 && p.r_sh == (long long)p.OW * p.r_sw && p.r_sn == (long long)p.OHW * p.r_sw) {
    const void* ident = device_identity128();
    if (!ident) return set_error(B200CV_ERR_DEVICE, "igemm: identity matrix allocation failed");
    rc = make_tmap_2d_bf16(&tmI, ident, 128, 128, 128, 128, kc);
    if (rc) return rc;
    // residual columns past the tensor are zero-filled by TMA (the residual has >= Cout channels)
    const long long res_cols = std::min<long long>(p.r_sw, (p.Cout + 7) / 8 * 8);
    rc = make_tmap_2d_bf16(&tmR, p.res, p.M_total, res_cols, p.r_sw, kc, 64);
    if (rc) return rc;
    pI = &tmI;
    pR = &tmR;
    p.res_iters = 128 / kc;
    p.res = nullptr;  // nothing left for the epilogue to add
  }
  // Staged TMA-store epilogue when the output is a plain row-major bf16 matrix [M][ld] (forward convs and
  // stride-1 data gradients); strided (parity scatter, NCHW), fp32 or unaligned outputs take the generic one.
  const long long ld = p.o_sw;
  static const bool no_tma_store = getenv("B200CV_NO_TMA_STORE") != nullptr;
  const bool tma_out = !no_tma_store && !p.out_fp32 && !p.out_lo && !p.res_lo && p.vec_ok && p.o_sc == 1 &&
                       ld >= p.Cout && ld % 8 == 0 &&
                       p.o_sh == (long long)p.OW * ld && p.o_sn == (long long)p.OHW * ld && p.Cout % 8 == 0 &&
                       (!p.res || p.res_vec_ok);
  // 256-row tiles for 128-wide GEMMs (data gradients of the 128-channel 3x3 layers, forward 64->128): the weight
  // tile is shared by two sub-tiles.  Only when there are enough tiles left to fill the machine twice.
  static const bool no_m2 = getenv("B200CV_NO_M2") != nullptr;
  p.tile_m = 128;
  // (not for short K loops: the HBM-bound 1x1 layers lose -- 35 -> 52 us for 256->128 @52x52)
  if (tma_out && !no_m2 && !p.d2s_c2 && kc == 64 && bn == 128 && p.num_taps * p.cblocks >= 8 &&
      (p.M_total + 255) / 256 * p.num_n_tiles >= 2 * sm_count()) {
    p.tile_m = 256;
    p.num_m_tiles = (p.M_total + 255) / 256;
  }
  if (p.d2s_c2) {
    // depth-to-space output (b200cv_conv_dgrad_d2s): {2C, OW, 2*N*OH} with strides 2C (pixel pair) / out_w*C (row)
    if (bn < 128) return set_error(B200CV_ERR_ARG, "conv_dgrad_d2s: needs the wide staged epilogue (4*C >= 128)");
    CUtensorMap tmO, tmY;
    rc = make_tmap_3d_bf16(&tmO, p.out, p.d2s_c2, p.OW, 2LL * g.N * g.OHt, p.d2s_c2, (long long)p.OW * p.d2s_c2, 64,
                           p.d2s_g);
    if (rc) return rc;
    if (p.bn_sums) {  // y of the layer whose dL/da this is: a contiguous NHWC tensor like the output
      rc = make_tmap_3d_bf16(&tmY, bn_y, p.d2s_c2, p.OW, 2LL * g.N * g.OHt, p.d2s_c2, (long long)p.OW * p.d2s_c2, 64,
                             p.d2s_g);
      if (rc) return rc;
    }
    return launch_igemm(tmA, tmB, &tmO, nullptr, nullptr, p.bn_sums ? &tmY : nullptr, p, kc, bn, stream);
  }
  if (tma_out) {
    CUtensorMap tmO, tmY;
    const int ebox = bn >= 128 ? 64 : (bn >= 32 ? 32 : 16);  // epilogue block width (conv_igemm.cu: kWide)
    rc = make_tmap_2d_bf16(&tmO, p.out, p.M_total, p.Cout, ld, 32, ebox);
    if (rc) return rc;
    if (p.bn_sums) {
      rc = make_tmap_2d_bf16(&tmY, bn_y, p.M_total, p.Cout, bn_y_ld, 32, ebox);
      if (rc) return rc;
    }
    return launch_igemm(tmA, tmB, &tmO, pI, pR, p.bn_sums ? &tmY : nullptr, p, kc, bn, stream);
  }
  if (p.bn_sums)
    return set_error(B200CV_ERR_ARG, "conv_dgrad: the fused BN-backward reduction needs a row-major bf16 output");
  return launch_igemm(tmA, tmB, nullptr, pI, pR, nullptr, p, kc, bn, stream);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void fill_epilogue(IgemmParams& p, const b200cv_conv_args* a, int64_t out_off, int64_t res_off,
                   int64_t mul_h, int64_t mul_w) {
  const int esz = a->y_dtype == B200CV_DT_F32 ? 4 : 2;
  p.out = static_cast<char*>(a->y) + out_off * esz;
  p.out_fp32 = a->y_dtype == B200CV_DT_F32;
  p.o_sn = a->y_sn;
  p.o_sh = a->y_sh * mul_h;
  p.o_sw = a->y_sw * mul_w;
  p.o_sc = a->y_sc;
  const int vec = 16 / esz;
  p.vec_ok = a->y_sc == 1 && aligned16(p.out) && p.o_sn % vec == 0 && p.o_sh % vec == 0 &&
             p.o_sw % vec == 0;
  p.res = a->residual ? static_cast<const __nv_bfloat16*>(a->residual) + res_off : nullptr;
  p.r_sn = a->r_sn;
  p.r_sh = a->r_sh * mul_h;
  p.r_sw = a->r_sw * mul_w;
  p.r_sc = a->r_sc;
  p.res_vec_ok = p.res && a->r_sc == 1 && aligned16(p.res) && p.r_sn % 8 == 0 && p.r_sh % 8 == 0 && p.r_sw % 8 == 0;
  p.scale = a->scale;
  p.shift = a->shift;
  p.act = a->act;
  p.res_after_act = a->res_after_act;
  p.slope = a->slope;
  p.stats = static_cast<StatAcc*>(a->stats);
  p.stats_parts = a->stats_parts > 0 ? a->stats_parts : 1;
  p.out_lo = a->y_lo;
  p.res_lo = p.res ? a->r_lo : 0;
}

int validate_common(const b200cv_conv_args* a) {
  B200CV_CHECK_ARG(a != nullptr, "conv: null args");
  B200CV_CHECK_ARG(a->x && a->w && a->y, "conv: null tensor pointer");
  B200CV_CHECK_ARG(a->N > 0 && a->H > 0 && a->W > 0 && a->Cout > 0, "conv: empty shape");
  B200CV_CHECK_ARG(a->Cin == pad_channels(a->Cin), "conv: Cin=%d is not a padded channel count", a->Cin);
  B200CV_CHECK_ARG(a->R >= 1 && a->S >= 1 && a->R * a->S <= kMaxTaps, "conv: filter %dx%d unsupported",
                   a->R, a->S);
  B200CV_CHECK_ARG(a->stride >= 1 && a->stride <= 8 && a->dil >= 1 && a->pad >= 0, "conv: bad stride/dil/pad");
  B200CV_CHECK_ARG(aligned16(a->x) && aligned16(a->w), "conv: x/w must be 16-byte aligned");
  B200CV_CHECK_ARG(a->y_dtype == B200CV_DT_BF16 || a->y_dtype == B200CV_DT_F32, "conv: bad y_dtype");
  B200CV_CHECK_ARG(!a->stats || a->Cout <= 1024, "conv: statistics support at most 1024 output channels");
  B200CV_CHECK_ARG(a->x_lo == 0 || a->x_lo == a->Cin, "conv: x_lo must be 0 or Cin (whole split tensors)");
  B200CV_CHECK_ARG(a->y_lo == 0 || (a->y_dtype == B200CV_DT_BF16 && a->y_lo % 8 == 0),
                   "conv: a split output is bf16 with a 16-byte aligned piece stride");
  B200CV_CHECK_ARG(a->r_lo == 0 || a->r_lo % 8 == 0, "conv: r_lo must be a multiple of 8");
  B200CV_CHECK_ARG(!(a->bn_sums && (a->x_lo || a->y_lo)), "conv: no fused BN-backward reduction in the fp32-parity mode");
  return 0;
}

}  // namespace
}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_conv_fwd(const b200cv_conv_args* a, void* stream) {
  if (int rc = validate_common(a)) return rc;
  const int OH = (a->H + 2 * a->pad - a->dil * (a->R - 1) - 1) / a->stride + 1;
  const int OW = (a->W + 2 * a->pad - a->dil * (a->S - 1) - 1) / a->stride + 1;
  B200CV_CHECK_ARG(OH > 0 && OW > 0, "conv_fwd: empty output");
  B200CV_CHECK_ARG(a->pad <= 127 && (a->R - 1) * a->dil <= 255, "conv_fwd: pad/dilation out of TMA range");
  Geometry g;
  g.N = a->N; g.H = a->H; g.W = a->W; g.C = a->Cin;
  g.lower_w = -a->pad; g.lower_h = -a->pad;
  g.upper_w = a->pad - (a->S - 1) * a->dil;
  g.upper_h = a->pad - (a->R - 1) * a->dil;
  g.trav_w = a->stride; g.trav_h = a->stride;
  g.OHt = OH; g.OWt = OW;
  IgemmParams p{};
  p.Cout = a->Cout;
  p.num_taps = a->R * a->S;
  for (int r = 0; r < a->R; ++r)
    for (int s = 0; s < a->S; ++s) {
      const int t = r * a->S + s;
      p.tap_w[t] = (short)(s * a->dil);
      p.tap_h[t] = (short)(r * a->dil);
      p.tap_k[t] = t * a->Cin;
    }
  fill_epilogue(p, a, 0, 0, 1, 1);
  return run_igemm(g, a->x, a->w, a->Cout, (int64_t)a->R * a->S * a->Cin, p,
                   static_cast<cudaStream_t>(stream), nullptr, 0, a->x_lo != 0);
}

extern "C" int b200cv_conv_dgrad(const b200cv_conv_args* a, int out_h, int out_w, void* stream) {
  if (int rc = validate_common(a)) return rc;
  B200CV_CHECK_ARG(out_h > 0 && out_w > 0, "conv_dgrad: empty output");
  const int s = a->stride;
  B200CV_CHECK_ARG(s == 1 || a->dil == 1, "conv_dgrad: stride>1 with dilation>1 unsupported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // One launch per output parity class (a single class when stride == 1).
  for (int pa = 0; pa < s; ++pa) {
    for (int pb = 0; pb < s; ++pb) {
      const int OHt = (out_h - pa + s - 1) / s;
      const int OWt = (out_w - pb + s - 1) / s;
      if (OHt <= 0 || OWt <= 0) continue;
      // taps r with (pa + pad - r*dil) divisible by s contribute dY[i + (pa + pad - r*dil)/s]
      int dr[kMaxTaps], rr[kMaxTaps], nr = 0, dsv[kMaxTaps], ss[kMaxTaps], ns = 0;
      for (int r = 0; r < a->R; ++r) {
        const int num = pa + a->pad - r * a->dil;
        if (((num % s) + s) % s == 0) { dr[nr] = num / s; rr[nr] = r; ++nr; }
      }
      for (int q = 0; q < a->S; ++q) {
        const int num = pb + a->pad - q * a->dil;
        if (((num % s) + s) % s == 0) { dsv[ns] = num / s; ss[ns] = q; ++ns; }
      }
      B200CV_CHECK_ARG(nr > 0 && ns > 0, "conv_dgrad: parity class (%d,%d) has no taps", pa, pb);
      const int lo_h = *std::min_element(dr, dr + nr);
      const int lo_w = *std::min_element(dsv, dsv + ns);
      Geometry g;
      g.N = a->N; g.H = a->H; g.W = a->W; g.C = a->Cin;
      g.lower_h = lo_h; g.lower_w = lo_w;
      g.upper_h = lo_h + OHt - a->H;
      g.upper_w = lo_w + OWt - a->W;
      g.trav_h = 1; g.trav_w = 1;
      g.OHt = OHt; g.OWt = OWt;
      B200CV_CHECK_ARG(g.lower_h >= -128 && g.lower_h <= 127 && g.upper_h >= -128 && g.upper_h <= 127 &&
                           g.lower_w >= -128 && g.lower_w <= 127 && g.upper_w >= -128 && g.upper_w <= 127,
                       "conv_dgrad: bounding box out of TMA range");
      IgemmParams p{};
      p.Cout = a->Cout;
      p.num_taps = nr * ns;
      int t = 0;
      for (int i = 0; i < nr; ++i)
        for (int j = 0; j < ns; ++j, ++t) {
          p.tap_h[t] = (short)(dr[i] - lo_h);
          p.tap_w[t] = (short)(dsv[j] - lo_w);
          p.tap_k[t] = (rr[i] * a->S + ss[j]) * a->Cin;
        }
      fill_epilogue(p, a, pa * a->y_sh + pb * a->y_sw, pa * a->r_sh + pb * a->r_sw, s, s);
      if (a->bn_sums) {
        B200CV_CHECK_ARG(s == 1, "conv_dgrad: the fused BN-backward reduction needs stride 1");
        B200CV_CHECK_ARG(a->bn_y && a->bn_scale && a->bn_shift && a->bn_mean && a->bn_rstd && a->bn_parts > 0 &&
                             a->bn_y_ld % 8 == 0 && aligned16(a->bn_y),
                         "conv_dgrad: incomplete bn_* arguments");
        p.bn_sums = static_cast<StatAcc*>(a->bn_sums);
        p.bn_parts = a->bn_parts;
        p.bn_scale = a->bn_scale;
        p.bn_shift = a->bn_shift;
        p.bn_mean = a->bn_mean;
        p.bn_rstd = a->bn_rstd;
        p.bn_neg = a->bn_act == B200CV_ACT_LEAKY ? a->bn_slope : (a->bn_act == B200CV_ACT_RELU ? 0.f : 1.f);
      }
      int rc = run_igemm(g, a->x, a->w, a->Cout, (int64_t)a->R * a->S * a->Cin, p, st, a->bn_y, a->bn_y_ld,
                         a->x_lo != 0);
      if (rc) return rc;
    }
  }
  return 0;
}

// Stride-2 3x3 pad-1 data gradient as ONE GEMM over the dy grid (instead of four parity-class launches with the
// direct-store epilogue): row m = dy pixel (n, q, p); the four input pixels (2q+a, 2p+b) it "owns" are the column
// blocks (a, b, c) of a 4*C wide output; K = the 2x2 dy neighbourhood (q+u, p+v) x Cout.  Input row 2q (a = 0) sees
// filter row r = 1 at u = 0; input row 2q+1 (a = 1) sees r = 2 at u = 0 and r = 0 at u = 1 -- the fourth (a, u)
// combination is a zero block of the operand written by the `transpose == 3` pack (7 of 16 blocks are zero: these
// layers are HBM-bound, 16/9 of the FLOPs is free).  The epilogue is the staged TMA store with a 3-D map.
extern "C" int b200cv_conv_dgrad_d2s(const b200cv_conv_args* a, int out_h, int out_w, void* stream) {
  if (int rc = validate_common(a)) return rc;
  const int C = a->Cout;  // channels of dx
  B200CV_CHECK_ARG(a->R == 3 && a->S == 3 && a->stride == 2 && a->pad == 1 && a->dil == 1,
                   "conv_dgrad_d2s: 3x3 stride-2 pad-1 layers only");
  B200CV_CHECK_ARG(out_h == 2 * a->H && out_w == 2 * a->W && a->W >= 32 && a->W % 4 == 0,
                   "conv_dgrad_d2s: needs an even %dx%d output whose width is a multiple of 8, at least 64", out_h, out_w);
  B200CV_CHECK_ARG(C == pad_channels(C) && C % 32 == 0, "conv_dgrad_d2s: C=%d must be a padded multiple of 32", C);
  B200CV_CHECK_ARG(!a->residual && !a->scale && !a->shift && a->act == 0 && !a->stats && !a->x_lo && !a->y_lo,
                   "conv_dgrad_d2s: plain bf16 gradient only (no residual / affine epilogue / split)");
  B200CV_CHECK_ARG(a->y_dtype == B200CV_DT_BF16 && a->y_sc == 1 && a->y_sw == C && a->y_sh == (int64_t)out_w * C &&
                       a->y_sn == (int64_t)out_h * out_w * C && aligned16(a->y),
                   "conv_dgrad_d2s: y must be a contiguous NHWC bf16 tensor");
  Geometry g;
  g.N = a->N; g.H = a->H; g.W = a->W; g.C = a->Cin;
  g.lower_h = 0; g.lower_w = 0; g.upper_h = 0; g.upper_w = 0;  // taps reach one pixel past the bottom / right edge
  g.trav_h = 1; g.trav_w = 1;
  g.OHt = a->H; g.OWt = a->W;
  IgemmParams p{};
  p.Cout = 4 * C;
  p.num_taps = 4;
  for (int u = 0; u < 2; ++u)
    for (int v = 0; v < 2; ++v) {
      p.tap_h[2 * u + v] = (short)u;
      p.tap_w[2 * u + v] = (short)v;
      p.tap_k[2 * u + v] = (2 * u + v) * a->Cin;
    }
  fill_epilogue(p, a, 0, 0, 1, 1);
  p.d2s_c2 = 2 * C;
  p.stat_cols = C;
  p.stat_mask = C - 1;
  if (a->bn_sums) {  // fused first pass of the BatchNorm backward of the layer that produced this activation
    B200CV_CHECK_ARG((C & (C - 1)) == 0, "conv_dgrad_d2s: the fused BN-backward reduction needs a power-of-two C");
    B200CV_CHECK_ARG(a->bn_y && a->bn_scale && a->bn_shift && a->bn_mean && a->bn_rstd && a->bn_parts > 0 &&
                         a->bn_y_ld == C && aligned16(a->bn_y),
                     "conv_dgrad_d2s: incomplete bn_* arguments (bn_y must be contiguous NHWC)");
    p.bn_sums = static_cast<StatAcc*>(a->bn_sums);
    p.bn_parts = a->bn_parts;
    p.bn_scale = a->bn_scale;
    p.bn_shift = a->bn_shift;
    p.bn_mean = a->bn_mean;
    p.bn_rstd = a->bn_rstd;
    p.bn_neg = a->bn_act == B200CV_ACT_LEAKY ? a->bn_slope : (a->bn_act == B200CV_ACT_RELU ? 0.f : 1.f);
  }
  p.d2s_g = 32;
  while (a->W % p.d2s_g) p.d2s_g >>= 1;
  return run_igemm(g, a->x, a->w, 4LL * C, 4LL * a->Cin, p, static_cast<cudaStream_t>(stream), a->bn_y, a->bn_y_ld);
}

// ------------------------------------------------------------------------------------------
// layout / packing kernels (bandwidth-trivial next to the convolutions)
namespace b200cv {
namespace {

// one thread per pixel: C coalesced plane reads (consecutive threads = consecutive pixels), then the
// pixel's Cpad bf16 channels go out as 16-byte stores
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                             int C, long long HW, int Cpad, long long total_pix) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_pix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW;
    const long long hw = i - n * HW;
    const float* s = src + n * C * HW + hw;
    __nv_bfloat16* d = dst + i * Cpad;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j) < C ? __ldg(s + (long long)(c0 + j) * HW) : 0.f;
      uint4 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>(d + c0) = pk;
    }
  }
}

__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                             int C, long long HW, int Cpad, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    const long long nc = i / HW;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    dst[i] = __bfloat162float(src[(n * HW + hw) * Cpad + c]);
  }
}

__device__ __forceinline__ void put_packed(__nv_bfloat16* dst, long long row, int col, int W, bool split, float v);

__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int O,
                                    int I, int RS, int Ipad, int Opad, int transpose, long long total) {
  const bool split = (transpose & 8) != 0;
  transpose &= 7;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    const int W = transpose ? Opad : Ipad;
    const int col = (int)(i % W);
    const long long r = i / W;
    if (!transpose) {  // dst[o][t][ip]
      const int t = (int)(r % RS);
      const int o = (int)(r / RS);
      if (col < I) v = w[((long long)o * I + col) * RS + t];
    } else {  // dst[i][t][op]
      const int t = (int)(r % RS);
      const int ii = (int)(r / RS);
      if (col < O) v = w[((long long)col * I + ii) * RS + t];
    }
    put_packed(dst, r, col, W, split, v);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int I, int RS,
                                    int Ipad, long long total) {
  // dst[o][i][t] = src[o][t][i]
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(j % RS);
    const long long r = j / RS;
    const int i = (int)(r % I);
    const long long o = r / I;
    dst[j] = src[(o * RS + t) * Ipad + i];
  }
}


// ---- multi-tensor variants: one launch for every conv of a network (grid.y = table entry) ----------
struct PackEntry {          // mirrors b200cv_pack_entry (include/b200cv.h)
  const float* src;         // OIHW fp32 (pack) / packed fp32 gradient (unpack)
  void* dst;                // packed bf16 (pack) / OIHW fp32 gradient (unpack)
  int O, I, RS, Ipad, Opad, transpose;
};
// Layout changes of the weights are transposes of small matrices: every block stages one contiguous piece of the
// source in shared memory so that BOTH the global reads and the global writes are coalesced (the first version
// gathered with a stride of RS floats and ran at ~1 TB/s).
constexpr int kPackSmemFloats = 9600;  // 37.5 KB: one [I*RS] row (I*RS <= 9600) or a 32 x (32*RS+1) tile (RS <= 9)

// One packed element: row-major [row][W] bf16, or -- fp32-parity (split) pack -- [row][w0(W) | w1(W) | w2(W)].
__device__ __forceinline__ void put_packed(__nv_bfloat16* dst, long long row, int col, int W, bool split, float v) {
  const long long b = row * (split ? kSplitPieces * W : W) + col;
  if (!split) {
    dst[b] = __float2bfloat16_rn(v);
    return;
  }
#pragma unroll
  for (int j = 0; j < kSplitPieces; ++j) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    dst[b + j * W] = h;
    v -= __bfloat162float(h);
  }
}

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackEntry* __restrict__ table) {
  __shared__ float sm[kPackSmemFloats];
  PackEntry e = table[blockIdx.y];
  const bool split = (e.transpose & 8) != 0;  // b200cv_pack_entry: transpose | 8 = split pack
  e.transpose &= 7;
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(e.dst);
  const int tid = threadIdx.x, nt = blockDim.x;
  if (e.transpose == 3) {
    // depth-to-space operand of b200cv_conv_dgrad_d2s: dst[(a*2+b)*I + i][(u*2+v)*Opad + o] (3x3 filters)
    const long long total = 16LL * e.I * e.Opad;
    for (long long idx = blockIdx.x * (long long)nt + tid; idx < total; idx += (long long)gridDim.x * nt) {
      const int col = (int)(idx % (4 * e.Opad));
      const int row = (int)(idx / (4 * e.Opad));
      const int uv = col / e.Opad, o = col - uv * e.Opad;
      const int ab = row / e.I, i = row - ab * e.I;
      const int r = (ab >> 1) == 0 ? ((uv >> 1) == 0 ? 1 : -1) : ((uv >> 1) == 0 ? 2 : 0);
      const int c = (ab & 1) == 0 ? ((uv & 1) == 0 ? 1 : -1) : ((uv & 1) == 0 ? 2 : 0);
      float v = 0.f;
      if (r >= 0 && c >= 0 && o < e.O) v = e.src[(((long long)o * e.I + i) * 3 + r) * 3 + c];
      dst[idx] = __float2bfloat16_rn(v);
    }
  } else if (e.transpose == 2) {
    // "flat" pack [O][Ipad] with k = tap*I + i (whole filter in one padded K run); tiny (image layers only)
    const long long total = (long long)e.O * e.Ipad;
    for (long long i = blockIdx.x * (long long)nt + tid; i < total; i += (long long)gridDim.x * nt) {
      const int k = (int)(i % e.Ipad);
      const int o = (int)(i / e.Ipad);
      float v = 0.f;
      if (k < e.RS * e.I) v = e.src[((long long)o * e.I + (k % e.I)) * e.RS + (k / e.I)];
      put_packed(dst, o, k, e.Ipad, split, v);
    }
  } else if (!e.transpose) {
    // dst[o][t][ip] <- src[o][i][t]: one output channel (I*RS contiguous floats) per block iteration.  No
    // per-element division: a warp walks one tap row at a time (the first version spent most of its time in
    // integer divisions by RS*Ipad and ran at 1 TB/s).
    const int row = e.I * e.RS;
    const bool staged = row <= kPackSmemFloats;
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int o = blockIdx.x; o < e.O; o += gridDim.x) {
      const float* sp = e.src + (long long)o * row;
      if (staged) {
        for (int j = tid; j < row; j += nt) sm[j] = sp[j];
        __syncthreads();
      }
      for (int t = warp; t < e.RS; t += nw) {
        const long long prow = (long long)o * e.RS + t;
        if (!split && (e.Ipad & 1) == 0) {  // two channels per lane: 4-byte stores
          __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dst + prow * e.Ipad);
          for (int ip = 2 * lane; ip < e.Ipad; ip += 64) {
            const float v0 = ip < e.I ? (staged ? sm[ip * e.RS + t] : sp[ip * e.RS + t]) : 0.f;
            const float v1 = ip + 1 < e.I ? (staged ? sm[(ip + 1) * e.RS + t] : sp[(ip + 1) * e.RS + t]) : 0.f;
            d2[ip >> 1] = __floats2bfloat162_rn(v0, v1);
          }
        } else {
          for (int ip = lane; ip < e.Ipad; ip += 32) {
            float v = 0.f;
            if (ip < e.I) v = staged ? sm[ip * e.RS + t] : sp[ip * e.RS + t];
            put_packed(dst, prow, ip, e.Ipad, split, v);
          }
        }
      }
      if (staged) __syncthreads();
    }
  } else {
    // dst[i][t][op] <- src[op][i][t]: 32 (input) x 32 (output) channel tiles
    const int ld = 32 * e.RS + 1;
    const int tiles_i = (e.I + 31) / 32, tiles_o = (e.Opad + 31) / 32;
    if (32 * ld > kPackSmemFloats) {  // huge filters: per-element gather
      const long long total = (long long)e.I * e.RS * e.Opad;
      for (long long i = blockIdx.x * (long long)nt + tid; i < total; i += (long long)gridDim.x * nt) {
        const int op = (int)(i % e.Opad);
        const long long r = i / e.Opad;
        const int t = (int)(r % e.RS);
        const int ii = (int)(r / e.RS);
        put_packed(dst, r, op, e.Opad, split, op < e.O ? e.src[((long long)op * e.I + ii) * e.RS + t] : 0.f);
      }
      return;
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int tile = blockIdx.x; tile < tiles_i * tiles_o; tile += gridDim.x) {
      const int i0 = (tile / tiles_o) * 32, o0 = (tile % tiles_o) * 32;
      const int ni = min(32, e.I - i0);
      const int run = ni * e.RS;  // contiguous source floats of one output channel inside this tile
      // load: warp w takes output channels w, w + 8, ...: `run` contiguous floats each (coalesced)
      for (int ol = warp; ol < 32; ol += nw) {
        const bool ok = o0 + ol < e.O;
        const float* sp = e.src + ((long long)(o0 + ol) * e.I + i0) * e.RS;
        for (int k = lane; k < run; k += 32) sm[ol * ld + k] = ok ? sp[k] : 0.f;
      }
      __syncthreads();
      // store: row k = il*RS + t of the tile is 32 consecutive output channels; two rows per warp instruction
      // (lanes 0-15 / 16-31), two channels per lane
      if (!split) {
        const int half = lane >> 4, ol2 = 2 * (lane & 15);
        for (int k = 2 * warp + half; k < run; k += 2 * nw) {
          if (o0 + ol2 < e.Opad) {  // Opad is even: the pair is in or out together
            const __nv_bfloat162 h = __floats2bfloat162_rn(sm[ol2 * ld + k], sm[(ol2 + 1) * ld + k]);
            *reinterpret_cast<__nv_bfloat162*>(dst + ((long long)i0 * e.RS + k) * e.Opad + o0 + ol2) = h;
          }
        }
      } else {
        for (int k = warp; k < run; k += nw)
          if (o0 + lane < e.Opad) put_packed(dst, (long long)i0 * e.RS + k, o0 + lane, e.Opad, split, sm[lane * ld + k]);
      }
      __syncthreads();
    }
  }
}
__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const PackEntry* __restrict__ table) {
  __shared__ float sm[kPackSmemFloats];
  const PackEntry e = table[blockIdx.y];
  float* dst = static_cast<float*>(e.dst);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ld = e.I + 1;
  if (e.transpose == 2 || e.RS * ld > kPackSmemFloats) {
    const long long total = (long long)e.O * e.I * e.RS;
    for (long long j = blockIdx.x * (long long)nt + tid; j < total; j += (long long)gridDim.x * nt) {
      const int t = (int)(j % e.RS);
      const long long r = j / e.RS;
      const int i = (int)(r % e.I);
      const long long o = r / e.I;
      dst[j] = e.transpose == 2 ? e.src[o * e.Ipad + t * e.I + i] : e.src[(o * e.RS + t) * e.Ipad + i];
    }
    return;
  }
  // dst[o][i][t] = src[o][t][i]: one output channel per block iteration through shared memory
  for (int o = blockIdx.x; o < e.O; o += gridDim.x) {
    const float* sp = e.src + (long long)o * e.RS * e.Ipad;
    for (int j = tid; j < e.RS * e.I; j += nt) {
      const int t = j / e.I, i = j - t * e.I;
      sm[t * ld + i] = sp[(long long)t * e.Ipad + i];
    }
    __syncthreads();
    float* dp = dst + (long long)o * e.I * e.RS;
    for (int j = tid; j < e.I * e.RS; j += nt) {
      const int i = j / e.RS, t = j - i * e.RS;
      dp[j] = sm[t * ld + i];
    }
    __syncthreads();
  }
}

// explicit im2col for the 3-channel input layers: NCHW fp32 image -> bf16 patch matrix [N*OH*OW][Kp],
// k = (r*S+s)*C + c.  One thread per output pixel (plane reads coalesce across threads); R,S,C are
// compile-time so the patch lives in registers.  The 32 rows of a warp are contiguous in the patch matrix: each lane
// parks its row in shared memory (row pitch + 16 bytes: conflict-free 16-byte stores) and the warp then streams the
// 32*Kp*2 bytes out as fully coalesced 16-byte stores (per-lane row stores touched 32 different rows per
// instruction: 0.435 -> 0.33 ms for the 384-byte rows of the RektNet 7x7 stem at 80x80 bs256; rows of <= 64 bytes are
// stored directly, staging them was 15 % slower).
template <int R, int S, int C>
__global__ void __launch_bounds__(256)
im2col_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W, int stride,
                   int pad, int dil, int OH, int OW, int Kp) {
  constexpr int K = R * S * C;
  constexpr int K8 = (K + 7) / 8 * 8;
  extern __shared__ uint4 im2col_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = Kp >> 3;        // 16-byte chunks per row
  const int pitch = chunks + 1;      // in 16-byte units
  const bool staged = chunks > 4;
  uint4* wbuf = im2col_smem + (size_t)warp * 32 * pitch;
  const long long total = (long long)N * OH * OW;
  const long long step = (long long)gridDim.x * blockDim.x;
  // whole warps iterate together (the tail warp keeps its inactive lanes for the cooperative copy)
  for (long long m0 = blockIdx.x * (long long)blockDim.x + warp * 32; m0 < total; m0 += step) {
    const long long m = m0 + lane;
    if (m < total) {
      const int ow = (int)(m % OW);
      const long long t = m / OW;
      const int oh = (int)(t % OH);
      const long long n = t / OH;
      const float* xn = x + n * C * H * W;
      float v[K8];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ih = oh * stride - pad + r * dil;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int iw = ow * stride - pad + s * dil;
          const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
          for (int c = 0; c < C; ++c) v[(r * S + s) * C + c] = ok ? __ldg(xn + ((long long)c * H + ih) * W + iw) : 0.f;
        }
      }
#pragma unroll
      for (int k = K; k < K8; ++k) v[k] = 0.f;
      // short rows (<= 64 bytes: the 3x3 stem) go straight out -- measured faster than staging them
      uint4* row = staged ? wbuf + lane * pitch : reinterpret_cast<uint4*>(out + m * Kp);
#pragma unroll
      for (int k0 = 0; k0 < K8; k0 += 8) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[k0 + 2 * j], v[k0 + 2 * j + 1]);
        row[k0 >> 3] = pk;
      }
      for (int j = K8 >> 3; j < chunks; ++j) row[j] = make_uint4(0, 0, 0, 0);
    }
    if (!staged) continue;
    __syncwarp();
    const long long rows = total - m0 < 32 ? total - m0 : 32;
    uint4* dst = reinterpret_cast<uint4*>(out + m0 * Kp);
    const int n16 = (int)rows * chunks;
    for (int q = lane; q < n16; q += 32) {
      const int r = q / chunks, j = q - r * chunks;
      dst[q] = wbuf[r * pitch + j];
    }
    __syncwarp();
  }
}
// generic fallback (any R,S,C)
__global__ void im2col_nchw_generic_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C,
                                           int H, int W, int R, int S, int stride, int pad, int dil, int OH, int OW,
                                           int Kp) {
  const long long total = (long long)N * OH * OW;
  const int K = R * S * C;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(m % OW);
    const long long t = m / OW;
    const int oh = (int)(t % OH);
    const long long n = t / OH;
    const float* xn = x + n * C * H * W;
    __nv_bfloat16* o = out + m * Kp;
    for (int k = 0; k < Kp; ++k) {
      float val = 0.f;
      if (k < K) {
        const int c = k % C, rs = k / C;
        const int ih = oh * stride - pad + (rs / S) * dil, iw = ow * stride - pad + (rs % S) * dil;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(xn + ((long long)c * H + ih) * W + iw);
      }
      o[k] = __float2bfloat16_rn(val);
    }
  }
}

// fp32-parity (split) forms of the layout kernels: plain loops, one thread per element
__global__ void nchw_f32_to_nhwc_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C,
                                              long long HW, int Cpad, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long pix = i / Cpad;
    const long long n = pix / HW, hw = pix - n * HW;
    const float v = c < C ? __ldg(src + (n * C + c) * HW + hw) : 0.f;
    put_packed(dst, pix, c, Cpad, true, v);
  }
}
__global__ void nhwc_split_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C,
                                              long long HW, int Cpad, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    const long long nc = i / HW;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    const __nv_bfloat16* p = src + (n * HW + hw) * kSplitPieces * Cpad + c;
    float v = 0.f;
#pragma unroll
    for (int j = kSplitPieces - 1; j >= 0; --j) v += __bfloat162float(p[j * Cpad]);
    dst[i] = v;
  }
}
__global__ void im2col_nchw_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C,
                                         int H, int W, int R, int S, int stride, int pad, int dil, int OH, int OW,
                                         int Kp) {
  const long long total = (long long)N * OH * OW * Kp;
  const int K = R * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const long long m = i / Kp;
    const int ow = (int)(m % OW);
    const long long t = m / OW;
    const int oh = (int)(t % OH);
    const long long n = t / OH;
    float val = 0.f;
    if (k < K) {
      const int c = k % C, rs = k / C;
      const int ih = oh * stride - pad + (rs / S) * dil, iw = ow * stride - pad + (rs % S) * dil;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(x + ((n * C + c) * H + ih) * W + iw);
    }
    put_packed(out, m, k, Kp, true, val);
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  return (int)std::max<long long>(1, std::min(g, cap));
}

}  // namespace
}  // namespace b200cv

extern "C" int b200cv_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W,
                                            int Cpad, void* stream) {
  B200CV_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "nchw->nhwc: bad args");
  const long long pix = (long long)N * H * W;
  nchw_f32_to_nhwc_bf16_kernel<<<grid_for(pix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), C, (long long)H * W, Cpad, pix);
  return check_launch("nchw_f32_to_nhwc_bf16");
}

extern "C" int b200cv_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W,
                                            int Cpad, void* stream) {
  B200CV_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "nhwc->nchw: bad args");
  const long long total = (long long)N * C * H * W;
  nhwc_bf16_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, C, (long long)H * W, Cpad, total);
  return check_launch("nhwc_bf16_to_nchw_f32");
}

extern "C" int b200cv_nchw_f32_to_nhwc_split(const float* src, void* dst, int N, int C, int H, int W, int Cpad,
                                             void* stream) {
  B200CV_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "nchw->nhwc(split): bad args");
  const long long total = (long long)N * H * W * Cpad;
  nchw_f32_to_nhwc_split_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), C, (long long)H * W, Cpad, total);
  return check_launch("nchw_f32_to_nhwc_split");
}

extern "C" int b200cv_nhwc_split_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int Cpad,
                                             void* stream) {
  B200CV_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "nhwc(split)->nchw: bad args");
  const long long total = (long long)N * C * H * W;
  nhwc_split_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, C, (long long)H * W, Cpad, total);
  return check_launch("nhwc_split_to_nchw_f32");
}

extern "C" int b200cv_im2col_nchw_f32_split(const float* x, void* patches, int N, int C, int H, int W, int R, int S,
                                            int stride, int pad, int dil, int Kp, void* stream) {
  B200CV_CHECK_ARG(x && patches && N > 0 && C > 0 && H > 0 && W > 0 && R > 0 && S > 0 && stride > 0 && dil > 0,
                   "im2col(split): bad args");
  B200CV_CHECK_ARG(Kp >= R * S * C && Kp % 8 == 0, "im2col(split): Kp=%d must be a multiple of 8 and >= R*S*C", Kp);
  const int OH = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int OW = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  B200CV_CHECK_ARG(OH > 0 && OW > 0, "im2col(split): empty output");
  const long long total = (long long)N * OH * OW * Kp;
  im2col_nchw_split_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(patches), N, C, H, W, R, S, stride, pad, dil, OH, OW, Kp);
  return check_launch("im2col_nchw_split");
}

extern "C" int b200cv_pack_weights(const float* w, void* dst, int O, int I, int R, int S, int Ipad,
                                   int Opad, int transpose, void* stream) {
  B200CV_CHECK_ARG(w && dst && O > 0 && I > 0 && R > 0 && S > 0 && Ipad >= I && Opad >= O,
                   "pack_weights: bad args");
  const long long total = (transpose & 7) ? (long long)I * R * S * Opad : (long long)O * R * S * Ipad;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(dst), O, I, R * S, Ipad, Opad, transpose, total);
  return check_launch("pack_weights");
}

extern "C" int b200cv_unpack_wgrad(const float* src, float* dst, int O, int I, int R, int S, int Ipad,
                                   void* stream) {
  B200CV_CHECK_ARG(src && dst && O > 0 && I > 0 && Ipad >= I, "unpack_wgrad: bad args");
  const long long total = (long long)O * I * R * S;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, I, R * S, Ipad, total);
  return check_launch("unpack_wgrad");
}

extern "C" int b200cv_pack_weights_multi(const void* table_dev, int n, void* stream) {
  B200CV_CHECK_ARG(table_dev && n > 0 && n <= 65535, "pack_weights_multi: bad args");
  static_assert(sizeof(PackEntry) == sizeof(b200cv_pack_entry), "pack entry ABI mismatch");
  pack_weights_multi_kernel<<<dim3(256, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const PackEntry*>(table_dev));
  return check_launch("pack_weights_multi");
}

extern "C" int b200cv_unpack_wgrad_multi(const void* table_dev, int n, void* stream) {
  B200CV_CHECK_ARG(table_dev && n > 0 && n <= 65535, "unpack_wgrad_multi: bad args");
  unpack_wgrad_multi_kernel<<<dim3(256, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const PackEntry*>(table_dev));
  return check_launch("unpack_wgrad_multi");
}

extern "C" int b200cv_im2col_nchw_f32(const float* x, void* patches, int N, int C, int H, int W, int R, int S,
                                      int stride, int pad, int dil, int Kp, void* stream) {
  B200CV_CHECK_ARG(x && patches && N > 0 && C > 0 && H > 0 && W > 0 && R > 0 && S > 0 && stride > 0 && dil > 0,
                   "im2col: bad args");
  B200CV_CHECK_ARG(Kp >= R * S * C && Kp % 8 == 0, "im2col: Kp=%d must be a multiple of 8 and >= R*S*C", Kp);
  const int OH = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int OW = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  B200CV_CHECK_ARG(OH > 0 && OW > 0, "im2col: empty output");
  const long long total = (long long)N * OH * OW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(patches);
  const int grid = grid_for(total, 256);
  // staging: 8 warps x 32 rows x (Kp*2 + 16) bytes
  const size_t stage_bytes = Kp / 8 > 4 ? (size_t)8 * 32 * (Kp / 8 + 1) * 16 : 0;
  auto launch = [&](auto kern) -> int {
    static size_t configured = 48 * 1024;  // per instantiation; benign race: the attribute set is idempotent
    if (stage_bytes > configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
      if (e != cudaSuccess) return set_error(static_cast<int>(e), "im2col smem attr: %s", cudaGetErrorString(e));
      configured = stage_bytes;
    }
    kern<<<grid, 256, stage_bytes, st>>>(x, o, N, H, W, stride, pad, dil, OH, OW, Kp);
    return check_launch("im2col_nchw");
  };
  if (stage_bytes <= 200 * 1024 && R == 3 && S == 3 && C == 3) return launch(im2col_nchw_kernel<3, 3, 3>);
  if (stage_bytes <= 200 * 1024 && R == 7 && S == 7 && C == 3) return launch(im2col_nchw_kernel<7, 7, 3>);
  if (stage_bytes <= 200 * 1024 && R == 3 && S == 3 && C == 1) return launch(im2col_nchw_kernel<3, 3, 1>);
  else
    im2col_nchw_generic_kernel<<<grid, 256, 0, st>>>(x, o, N, C, H, W, R, S, stride, pad, dil, OH, OW, Kp);
  return check_launch("im2col_nchw");
}
