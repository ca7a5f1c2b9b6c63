// Native bring-up harness for the tcgen05/TMA convolution kernels.  Calls libb200cv.so through
// its C ABI only and checks every case against a plain CPU loop nest over the same
// bf16-rounded operands.  Test infrastructure -- not part of the product path.
//
//   conv_harness            run all cases
//   conv_harness <substr>   run the cases whose name contains <substr>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/b200cv.h"

static const char* g_filter = nullptr;
static int g_fail = 0, g_run = 0;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint32_t g_seed = 12345;
static float frand() {  // uniform [-1, 1)
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Case {
  std::string name;
  int N, H, W, Cin_true, Cout, R, S, stride, pad, dil;
  bool fp32_out = false, nchw_out = false, affine = false, residual = false, stats = false;
  int act = 0;
};

static int pad_c(int c) { return b200cv_pad_channels(c); }

template <typename T>
static T* dalloc(size_t n) {
  T* p;
  CK(cudaMalloc(&p, n * sizeof(T)));
  CK(cudaMemset(p, 0, n * sizeof(T)));
  return p;
}

static bool report(const std::string& name, double max_err, double max_ref, double tol_rel, int nbad,
                   int rc, int derr) {
  const bool ok = rc == 0 && derr == 0 && nbad == 0;
  printf("%-44s %s  max_err=%.4g max_ref=%.4g tol=%.3g bad=%d rc=%d dev=%d %s\n", name.c_str(),
         ok ? "PASS" : "FAIL", max_err, max_ref, tol_rel, nbad, rc, derr,
         (rc || derr) ? b200cv_last_error() : "");
  fflush(stdout);
  ++g_run;
  if (!ok) ++g_fail;
  return ok;
}

// ------------------------------------------------------------------------------------------
static void run_fwd(const Case& c) {
  if (g_filter && c.name.find(g_filter) == std::string::npos) return;
  const int Cin = pad_c(c.Cin_true);
  const int OH = (c.H + 2 * c.pad - c.dil * (c.R - 1) - 1) / c.stride + 1;
  const int OW = (c.W + 2 * c.pad - c.dil * (c.S - 1) - 1) / c.stride + 1;
  const size_t nx = (size_t)c.N * c.H * c.W * Cin;
  const size_t nw = (size_t)c.Cout * c.R * c.S * Cin;
  const int Cld = c.nchw_out ? c.Cout : pad_c(c.Cout);
  const size_t ny = (size_t)c.N * OH * OW * Cld;
  std::vector<float> x(nx, 0.f), w(nw, 0.f), scale(c.Cout, 1.f), shift(c.Cout, 0.f), res(ny, 0.f);
  for (size_t i = 0; i < nx; ++i)
    if ((int)(i % Cin) < c.Cin_true) x[i] = bf(frand());
  for (size_t i = 0; i < nw; ++i)
    if ((int)(i % Cin) < c.Cin_true) w[i] = bf(frand() * 0.25f);
  if (c.affine)
    for (int i = 0; i < c.Cout; ++i) { scale[i] = 0.5f + 0.5f * frand(); shift[i] = frand(); }
  if (c.residual)
    for (size_t i = 0; i < ny; ++i) res[i] = bf(frand());

  std::vector<__nv_bfloat16> xb(nx), wb(nw), rb(ny);
  for (size_t i = 0; i < nx; ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < nw; ++i) wb[i] = __float2bfloat16_rn(w[i]);
  for (size_t i = 0; i < ny; ++i) rb[i] = __float2bfloat16_rn(res[i]);
  auto* dx = dalloc<__nv_bfloat16>(nx);
  auto* dw = dalloc<__nv_bfloat16>(nw);
  auto* dr = dalloc<__nv_bfloat16>(ny);
  CK(cudaMemcpy(dx, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, wb.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr, rb.data(), ny * 2, cudaMemcpyHostToDevice));
  float *dscale = dalloc<float>(c.Cout), *dshift = dalloc<float>(c.Cout), *dstats = dalloc<float>(8 * c.Cout);  // b200cv_stat [2*Cout]
  CK(cudaMemcpy(dscale, scale.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dshift, shift.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  void* dy = c.fp32_out ? (void*)dalloc<float>(ny) : (void*)dalloc<__nv_bfloat16>(ny);

  b200cv_conv_args a;
  memset(&a, 0, sizeof(a));
  a.N = c.N; a.H = c.H; a.W = c.W; a.Cin = Cin; a.Cout = c.Cout;
  a.R = c.R; a.S = c.S; a.stride = c.stride; a.pad = c.pad; a.dil = c.dil;
  a.x = dx; a.w = dw; a.y = dy;
  a.y_dtype = c.fp32_out ? B200CV_DT_F32 : B200CV_DT_BF16;
  if (c.nchw_out) { a.y_sn = (int64_t)c.Cout * OH * OW; a.y_sc = (int64_t)OH * OW; a.y_sh = OW; a.y_sw = 1; }
  else { a.y_sn = (int64_t)OH * OW * Cld; a.y_sh = (int64_t)OW * Cld; a.y_sw = Cld; a.y_sc = 1; }
  if (c.affine) { a.scale = dscale; a.shift = dshift; }
  if (c.residual) { a.residual = dr; a.r_sn = a.y_sn; a.r_sh = a.y_sh; a.r_sw = a.y_sw; a.r_sc = a.y_sc; }
  a.act = c.act; a.slope = 0.1f;
  if (c.stats) a.stats = dstats;
  int rc = b200cv_conv_fwd(&a, nullptr);
  int derr = b200cv_check_device_error(nullptr);
  CK(cudaDeviceSynchronize());

  std::vector<float> y(ny, 0.f);
  if (c.fp32_out) CK(cudaMemcpy(y.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
  else {
    std::vector<__nv_bfloat16> yb(ny);
    CK(cudaMemcpy(yb.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) y[i] = __bfloat162float(yb[i]);
  }
  std::vector<float> hstats(2 * c.Cout, 0.f);
  {
    std::vector<b200cv_stat> raw(2 * c.Cout);
    CK(cudaMemcpy(raw.data(), dstats, 2 * c.Cout * sizeof(b200cv_stat), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 2 * c.Cout; ++i)
      hstats[i] = (float)((double)raw[i].w1 / 1048576.0 + (double)raw[i].w2 * 0x1p-70);
  }

  double max_err = 0, max_ref = 0;
  int nbad = 0;
  std::vector<double> ssum(c.Cout, 0.0), ssq(c.Cout, 0.0);
  const double tol = c.fp32_out ? 2e-3 : 1.2e-2;
  for (int n = 0; n < c.N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow)
        for (int o = 0; o < c.Cout; ++o) {
          double acc = 0;
          for (int r = 0; r < c.R; ++r) {
            const int ih = oh * c.stride - c.pad + r * c.dil;
            if (ih < 0 || ih >= c.H) continue;
            for (int s = 0; s < c.S; ++s) {
              const int iw = ow * c.stride - c.pad + s * c.dil;
              if (iw < 0 || iw >= c.W) continue;
              const float* xp = &x[(((size_t)n * c.H + ih) * c.W + iw) * Cin];
              const float* wp = &w[(((size_t)o * c.R + r) * c.S + s) * Cin];
              for (int ci = 0; ci < c.Cin_true; ++ci) acc += (double)xp[ci] * wp[ci];
            }
          }
          const size_t yi = c.nchw_out ? (((size_t)n * c.Cout + o) * OH + oh) * OW + ow
                                       : (((size_t)n * OH + oh) * OW + ow) * Cld + o;
          double v = acc * scale[o] + shift[o];
          if (c.residual) v += res[yi];
          if (c.act == 1) v = v > 0 ? v : 0.1 * v;
          if (c.act == 2) v = v > 0 ? v : 0;
          const double got = y[yi];
          ssum[o] += got;
          ssq[o] += got * got;
          const double err = fabs(got - v);
          if (fabs(v) > max_ref) max_ref = fabs(v);
          if (err > max_err) max_err = err;
          if (err > tol * (fabs(v) + 1.0)) {
            if (nbad < 6)
              printf("   mismatch n=%d oh=%d ow=%d o=%d got=%.5f want=%.5f\n", n, oh, ow, o, got, v);
            ++nbad;
          }
        }
  if (c.stats) {
    for (int o = 0; o < c.Cout; ++o) {
      const double e1 = fabs(hstats[o] - ssum[o]), e2 = fabs(hstats[c.Cout + o] - ssq[o]);
      if (e1 > 1e-3 * (fabs(ssum[o]) + 10.0) || e2 > 1e-3 * (ssq[o] + 10.0)) {
        if (nbad < 6) printf("   stats mismatch o=%d sum %.4f vs %.4f  sq %.4f vs %.4f\n", o, hstats[o], ssum[o],
                             hstats[c.Cout + o], ssq[o]);
        ++nbad;
      }
    }
  }
  report("fwd/" + c.name, max_err, max_ref, tol, nbad, rc, derr);
  cudaFree(dx); cudaFree(dw); cudaFree(dr); cudaFree(dscale); cudaFree(dshift); cudaFree(dstats); cudaFree(dy);
}

// ------------------------------------------------------------------------------------------
// dgrad: dX[n,h,w,ci] = sum_{r,s,o} dY[n,(h+pad-r*dil)/st,(w+pad-s*dil)/st,o] * W[o][r][s][ci]
static void run_dgrad(const Case& c) {
  if (g_filter && ("dgrad/" + c.name).find(g_filter) == std::string::npos) return;
  const int Cin_f = c.Cin_true;              // forward input channels = dX channels
  const int Cxp = pad_c(Cin_f);              // dX channel pitch
  const int Cop = pad_c(c.Cout);             // dY channel pitch
  const int OH = (c.H + 2 * c.pad - c.dil * (c.R - 1) - 1) / c.stride + 1;
  const int OW = (c.W + 2 * c.pad - c.dil * (c.S - 1) - 1) / c.stride + 1;
  const size_t ndy = (size_t)c.N * OH * OW * Cop;
  const size_t nwt = (size_t)Cin_f * c.R * c.S * Cop;
  const size_t ndx = (size_t)c.N * c.H * c.W * Cxp;
  std::vector<float> dy(ndy, 0.f), wt(nwt, 0.f), res(ndx, 0.f);
  for (size_t i = 0; i < ndy; ++i)
    if ((int)(i % Cop) < c.Cout) dy[i] = bf(frand());
  for (size_t i = 0; i < nwt; ++i)
    if ((int)(i % Cop) < c.Cout) wt[i] = bf(frand() * 0.25f);
  if (c.residual)
    for (size_t i = 0; i < ndx; ++i) res[i] = bf(frand());
  std::vector<__nv_bfloat16> dyb(ndy), wtb(nwt), rb(ndx);
  for (size_t i = 0; i < ndy; ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
  for (size_t i = 0; i < nwt; ++i) wtb[i] = __float2bfloat16_rn(wt[i]);
  for (size_t i = 0; i < ndx; ++i) rb[i] = __float2bfloat16_rn(res[i]);
  auto* ddy = dalloc<__nv_bfloat16>(ndy);
  auto* dwt = dalloc<__nv_bfloat16>(nwt);
  auto* dres = dalloc<__nv_bfloat16>(ndx);
  auto* ddx = dalloc<__nv_bfloat16>(ndx);
  CK(cudaMemcpy(ddy, dyb.data(), ndy * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwt, wtb.data(), nwt * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dres, rb.data(), ndx * 2, cudaMemcpyHostToDevice));

  b200cv_conv_args a;
  memset(&a, 0, sizeof(a));
  a.N = c.N; a.H = OH; a.W = OW; a.Cin = Cop; a.Cout = Cin_f;
  a.R = c.R; a.S = c.S; a.stride = c.stride; a.pad = c.pad; a.dil = c.dil;
  a.x = ddy; a.w = dwt; a.y = ddx; a.y_dtype = B200CV_DT_BF16;
  a.y_sn = (int64_t)c.H * c.W * Cxp; a.y_sh = (int64_t)c.W * Cxp; a.y_sw = Cxp; a.y_sc = 1;
  if (c.residual) { a.residual = dres; a.r_sn = a.y_sn; a.r_sh = a.y_sh; a.r_sw = a.y_sw; a.r_sc = 1; }
  int rc = b200cv_conv_dgrad(&a, c.H, c.W, nullptr);
  int derr = b200cv_check_device_error(nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> dxb(ndx);
  CK(cudaMemcpy(dxb.data(), ddx, ndx * 2, cudaMemcpyDeviceToHost));

  double max_err = 0, max_ref = 0;
  int nbad = 0;
  const double tol = 1.2e-2;
  for (int n = 0; n < c.N; ++n)
    for (int h = 0; h < c.H; ++h)
      for (int w = 0; w < c.W; ++w)
        for (int ci = 0; ci < Cin_f; ++ci) {
          double acc = 0;
          for (int r = 0; r < c.R; ++r) {
            const int nh = h + c.pad - r * c.dil;
            if (nh < 0 || nh % c.stride) continue;
            const int oh = nh / c.stride;
            if (oh >= OH) continue;
            for (int s = 0; s < c.S; ++s) {
              const int nw_ = w + c.pad - s * c.dil;
              if (nw_ < 0 || nw_ % c.stride) continue;
              const int ow = nw_ / c.stride;
              if (ow >= OW) continue;
              const float* dp = &dy[(((size_t)n * OH + oh) * OW + ow) * Cop];
              const float* wp = &wt[(((size_t)ci * c.R + r) * c.S + s) * Cop];
              for (int o = 0; o < c.Cout; ++o) acc += (double)dp[o] * wp[o];
            }
          }
          const size_t xi = (((size_t)n * c.H + h) * c.W + w) * Cxp + ci;
          if (c.residual) acc += res[xi];
          const double got = __bfloat162float(dxb[xi]);
          const double err = fabs(got - acc);
          if (fabs(acc) > max_ref) max_ref = fabs(acc);
          if (err > max_err) max_err = err;
          if (err > tol * (fabs(acc) + 1.0)) {
            if (nbad < 6) printf("   mismatch n=%d h=%d w=%d ci=%d got=%.5f want=%.5f\n", n, h, w, ci, got, acc);
            ++nbad;
          }
        }
  report("dgrad/" + c.name, max_err, max_ref, tol, nbad, rc, derr);
  cudaFree(ddy); cudaFree(dwt); cudaFree(dres); cudaFree(ddx);
}

// ------------------------------------------------------------------------------------------
static void run_wgrad(const Case& c) {
  if (g_filter && ("wgrad/" + c.name).find(g_filter) == std::string::npos) return;
  const int Cin = pad_c(c.Cin_true);
  const int Cop = pad_c(c.Cout);
  const int OH = (c.H + 2 * c.pad - c.dil * (c.R - 1) - 1) / c.stride + 1;
  const int OW = (c.W + 2 * c.pad - c.dil * (c.S - 1) - 1) / c.stride + 1;
  const size_t nx = (size_t)c.N * c.H * c.W * Cin;
  const size_t ndy = (size_t)c.N * OH * OW * Cop;
  const size_t ndw = (size_t)c.Cout * c.R * c.S * Cin;
  std::vector<float> x(nx, 0.f), dy(ndy, 0.f);
  for (size_t i = 0; i < nx; ++i)
    if ((int)(i % Cin) < c.Cin_true) x[i] = bf(frand());
  for (size_t i = 0; i < ndy; ++i)
    if ((int)(i % Cop) < c.Cout) dy[i] = bf(frand() * 0.25f);
  std::vector<__nv_bfloat16> xb(nx), dyb(ndy);
  for (size_t i = 0; i < nx; ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < ndy; ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
  auto* dx = dalloc<__nv_bfloat16>(nx);
  auto* ddy = dalloc<__nv_bfloat16>(ndy);
  auto* ddw = dalloc<float>(ndw);
  CK(cudaMemcpy(dx, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy, dyb.data(), ndy * 2, cudaMemcpyHostToDevice));
  int rc = b200cv_conv_wgrad(dx, ddy, ddw, c.N, c.H, c.W, Cin, c.Cout, Cop, c.R, c.S, c.stride, c.pad, c.dil,
                             0, 0, nullptr);
  int derr = b200cv_check_device_error(nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<float> dw(ndw);
  CK(cudaMemcpy(dw.data(), ddw, ndw * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  int nbad = 0;
  const double tol = 2e-3;
  for (int o = 0; o < c.Cout; ++o)
    for (int r = 0; r < c.R; ++r)
      for (int s = 0; s < c.S; ++s)
        for (int ci = 0; ci < Cin; ++ci) {
          double acc = 0;
          if (ci < c.Cin_true)
            for (int n = 0; n < c.N; ++n)
              for (int oh = 0; oh < OH; ++oh) {
                const int ih = oh * c.stride - c.pad + r * c.dil;
                if (ih < 0 || ih >= c.H) continue;
                for (int ow = 0; ow < OW; ++ow) {
                  const int iw = ow * c.stride - c.pad + s * c.dil;
                  if (iw < 0 || iw >= c.W) continue;
                  acc += (double)dy[(((size_t)n * OH + oh) * OW + ow) * Cop + o] *
                         x[(((size_t)n * c.H + ih) * c.W + iw) * Cin + ci];
                }
              }
          const double got = dw[(((size_t)o * c.R + r) * c.S + s) * Cin + ci];
          const double err = fabs(got - acc);
          if (fabs(acc) > max_ref) max_ref = fabs(acc);
          if (err > max_err) max_err = err;
          if (err > tol * (fabs(acc) + 1.0)) {
            if (nbad < 6) printf("   mismatch o=%d r=%d s=%d ci=%d got=%.5f want=%.5f\n", o, r, s, ci, got, acc);
            ++nbad;
          }
        }
  report("wgrad/" + c.name, max_err, max_ref, tol, nbad, rc, derr);
  cudaFree(dx); cudaFree(ddy); cudaFree(ddw);
}


// ------------------------------------------------------------------------------------------
// Timing mode: real Darknet-53 / RektNet layer shapes, CUDA-event timed, no CPU check.
static void bench_case(const char* name, int N, int H, int W, int Cin_true, int Cout, int R, int stride, int pad,
                       int dil) {
  if (getenv("B200CV_BENCH_N")) N = atoi(getenv("B200CV_BENCH_N"));  // e.g. an L2-resident problem
  const int Cin = pad_c(Cin_true), Cop = pad_c(Cout);
  const int OH = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int OW = (W + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const size_t nx = (size_t)N * H * W * Cin, ny = (size_t)N * OH * OW * Cop, nw = (size_t)Cop * R * R * Cin;
  auto* dx = dalloc<__nv_bfloat16>(nx);
  auto* dy = dalloc<__nv_bfloat16>(ny);
  auto* dw = dalloc<__nv_bfloat16>(nw);
  auto* dwt = dalloc<__nv_bfloat16>((size_t)Cin * R * R * Cop);
  auto* dgw = dalloc<float>(nw);
  auto* dstats = dalloc<float>(8 * Cout);  // b200cv_stat [2*Cout]
  // non-trivial contents so the power draw is realistic
  {
    std::vector<__nv_bfloat16> h(std::max(nx, std::max(ny, nw)));
    for (auto& v : h) v = __float2bfloat16_rn(frand());
    CK(cudaMemcpy(dx, h.data(), nx * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy, h.data(), ny * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, h.data(), nw * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dwt, h.data(), nw * 2, cudaMemcpyHostToDevice));
  }
  b200cv_conv_args f;
  memset(&f, 0, sizeof(f));
  f.N = N; f.H = H; f.W = W; f.Cin = Cin; f.Cout = Cout; f.R = R; f.S = R; f.stride = stride; f.pad = pad; f.dil = dil;
  f.x = dx; f.w = dw; f.y = dy; f.y_dtype = B200CV_DT_BF16;
  f.y_sn = (int64_t)OH * OW * Cop; f.y_sh = (int64_t)OW * Cop; f.y_sw = Cop; f.y_sc = 1;
  f.stats = dstats;
  b200cv_conv_args g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.H = OH; g.W = OW; g.Cin = Cop; g.Cout = Cin_true; g.R = R; g.S = R; g.stride = stride; g.pad = pad; g.dil = dil;
  g.x = dy; g.w = dwt; g.y = dx; g.y_dtype = B200CV_DT_BF16;
  g.y_sn = (int64_t)H * W * Cin; g.y_sh = (int64_t)W * Cin; g.y_sw = Cin; g.y_sc = 1;
  const double flop = 2.0 * N * OH * OW * (double)Cout * R * R * Cin_true;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms[3] = {0, 0, 0};
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 10;
    int rc = 0;
    for (int i = 0; i < 3 + iters; ++i) {
      if (i == 3) CK(cudaEventRecord(e0));
      if (mode == 0) rc |= b200cv_conv_fwd(&f, nullptr);
      if (mode == 1) rc |= b200cv_conv_dgrad(&g, H, W, nullptr);
      if (mode == 2) rc |= b200cv_conv_wgrad(dx, dy, dgw, N, H, W, Cin, Cout, Cop, R, R, stride, pad, dil, 0, 0, nullptr);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms[mode], e0, e1));
    ms[mode] /= iters;
    if (rc) printf("  rc=%d %s\n", rc, b200cv_last_error());
  }
  int derr = b200cv_check_device_error(nullptr);
  printf("%-28s N%d %dx%d c%d->%d k%d s%d | fwd %.3f ms %.0f TF | dgrad %.3f ms %.0f TF | wgrad %.3f ms %.0f TF | dev=%d\n",
         name, N, H, W, Cin_true, Cout, R, stride, ms[0], flop / ms[0] * 1e-9, ms[1], flop / ms[1] * 1e-9, ms[2],
         flop / ms[2] * 1e-9, derr);
  fflush(stdout);
  cudaFree(dx); cudaFree(dy); cudaFree(dw); cudaFree(dwt); cudaFree(dgw); cudaFree(dstats);
}

static int g_only = -1, g_case = 0;
#define bench_case(...) do { if (g_only < 0 || g_only == g_case) bench_case(__VA_ARGS__); ++g_case; } while (0)
static void run_bench() {
  bench_case("dk53 3x3 128->256 @52", 64, 52, 52, 128, 256, 3, 1, 1, 1);
  bench_case("dk53 3x3 256->512 @26", 64, 26, 26, 256, 512, 3, 1, 1, 1);
  bench_case("dk53 3x3 512->1024 @13", 64, 13, 13, 512, 1024, 3, 1, 1, 1);
  bench_case("dk53 3x3 64->128 @104", 64, 104, 104, 64, 128, 3, 1, 1, 1);
  bench_case("dk53 3x3 32->64 @208", 64, 208, 208, 32, 64, 3, 1, 1, 1);
  bench_case("dk53 1x1 256->128 @52", 64, 52, 52, 256, 128, 1, 1, 0, 1);
  bench_case("dk53 1x1 512->256 @26", 64, 26, 26, 512, 256, 1, 1, 0, 1);
  bench_case("dk53 1x1 1024->512 @13", 64, 13, 13, 1024, 512, 1, 1, 0, 1);
  bench_case("dk53 3x3 3->32 @416", 64, 416, 416, 3, 32, 3, 1, 1, 1);
  bench_case("dk53 3x3s2 32->64 @416", 64, 416, 416, 32, 64, 3, 2, 1, 1);
  bench_case("dk53 3x3s2 128->256 @104", 64, 104, 104, 128, 256, 3, 2, 1, 1);
  bench_case("dk53 1x1 1024->255 @13", 64, 13, 13, 1024, 255, 1, 1, 0, 1);
  bench_case("rekt 3x3d2 64->128 @80", 256, 80, 80, 64, 128, 3, 1, 2, 2);
  bench_case("rekt 3x3 128->128 @80", 256, 80, 80, 128, 128, 3, 1, 1, 1);
  bench_case("rekt 3x3 16->16 @80", 256, 80, 80, 16, 16, 3, 1, 1, 1);
  bench_case("rekt 7x7 3->16 @80", 256, 80, 80, 3, 16, 7, 1, 3, 1);
  bench_case("rekt 3x3d2 16->32 @80", 256, 80, 80, 16, 32, 3, 1, 2, 2);
  bench_case("rekt 3x3 32->32 @80", 256, 80, 80, 32, 32, 3, 1, 1, 1);
  bench_case("rekt 3x3d2 32->64 @80", 256, 80, 80, 32, 64, 3, 1, 2, 2);
  bench_case("rekt 3x3 64->64 @80", 256, 80, 80, 64, 64, 3, 1, 1, 1);
  bench_case("rekt 1x1 16->32 @80", 256, 80, 80, 16, 32, 1, 1, 0, 1);
  bench_case("dk53 1x1 64->32 @208", 64, 208, 208, 64, 32, 1, 1, 0, 1);
  bench_case("dk53 1x1 128->64 @104", 64, 104, 104, 128, 64, 1, 1, 0, 1);
  bench_case("dk53 3x3s2 64->128 @208", 64, 208, 208, 64, 128, 3, 2, 1, 1);
}

static Case mk(const char* name, int N, int H, int W, int Cin, int Cout, int R, int stride, int pad, int dil) {
  Case c;
  c.name = name; c.N = N; c.H = H; c.W = W; c.Cin_true = Cin; c.Cout = Cout; c.R = R; c.S = R;
  c.stride = stride; c.pad = pad; c.dil = dil;
  return c;
}

int main(int argc, char** argv) {
  if (argc > 1) g_filter = argv[1];
  printf("%s\n", b200cv_version());
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (g_filter && strcmp(g_filter, "bench") == 0) {
    if (argc > 2) g_only = atoi(argv[2]);
    run_bench();
    return 0;
  }

  std::vector<Case> fwd;
  fwd.push_back(mk("1x1_c64_o64", 2, 12, 12, 64, 64, 1, 1, 0, 1));
  fwd.push_back(mk("1x1_c128_o128", 2, 13, 13, 128, 128, 1, 1, 0, 1));
  fwd.push_back(mk("3x3_c64_o128", 2, 13, 13, 64, 128, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_c128_o256", 3, 13, 13, 128, 256, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_c64_o512_ntiles", 2, 10, 10, 64, 512, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_s2_c64_o128", 2, 16, 16, 64, 128, 3, 2, 1, 1));
  fwd.push_back(mk("3x3_s2_odd_c64_o64", 2, 13, 13, 64, 64, 3, 2, 1, 1));
  fwd.push_back(mk("3x3_c32_o64", 2, 13, 13, 32, 64, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_c16_o32", 2, 13, 13, 16, 32, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_c3_o32", 2, 16, 16, 3, 32, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_dil2_c16_o16", 2, 20, 20, 16, 16, 3, 1, 2, 2));
  fwd.push_back(mk("3x3_dil2_c64_o128", 2, 20, 20, 64, 128, 3, 1, 2, 2));
  fwd.push_back(mk("7x7_c3_o16", 2, 20, 20, 3, 16, 7, 1, 3, 1));
  { Case c = mk("1x1_c128_o7_nchw_f32", 2, 20, 20, 128, 7, 1, 1, 0, 1); c.fp32_out = true; c.nchw_out = true; c.affine = true; fwd.push_back(c); }
  { Case c = mk("1x1_c256_o18_f32_bias", 2, 13, 13, 256, 18, 1, 1, 0, 1); c.fp32_out = true; c.affine = true; fwd.push_back(c); }
  { Case c = mk("1x1_c512_o255_f32_bias", 2, 13, 13, 512, 255, 1, 1, 0, 1); c.fp32_out = true; c.affine = true; fwd.push_back(c); }
  { Case c = mk("3x3_c64_o128_stats", 4, 13, 13, 64, 128, 3, 1, 1, 1); c.stats = true; fwd.push_back(c); }
  { Case c = mk("3x3_c64_o64_affine_leaky_res", 2, 13, 13, 64, 64, 3, 1, 1, 1); c.affine = true; c.act = 1; c.residual = true; fwd.push_back(c); }
  { Case c = mk("3x3_c128_o256_big_stats", 8, 52, 52, 128, 256, 3, 1, 1, 1); c.stats = true; fwd.push_back(c); }
  fwd.push_back(mk("1x1_c1024_o512", 4, 13, 13, 1024, 512, 1, 1, 0, 1));
  // several tiles per CTA on the small-stage configurations (multi-lane TMA producers, ring wraps, ragged last tile)
  fwd.push_back(mk("3x3_c16_o16_many_tiles", 33, 80, 79, 16, 16, 3, 1, 1, 1));
  fwd.push_back(mk("3x3_dil2_c32_o64_many_tiles", 17, 80, 80, 32, 64, 3, 1, 2, 2));
  fwd.push_back(mk("1x1_c64_o32_many_tiles", 9, 104, 104, 64, 32, 1, 1, 0, 1));
  { Case c = mk("3x3_c64_o64_many_tiles_stats_res", 9, 80, 80, 64, 64, 3, 1, 1, 1); c.stats = true; c.affine = true; c.act = 1; c.residual = true; fwd.push_back(c); }
  for (auto& c : fwd) run_fwd(c);

  std::vector<Case> dg;
  dg.push_back(mk("1x1_c64_o128", 2, 13, 13, 64, 128, 1, 1, 0, 1));
  dg.push_back(mk("3x3_c64_o128", 2, 13, 13, 64, 128, 3, 1, 1, 1));
  dg.push_back(mk("3x3_c128_o256", 2, 13, 13, 128, 256, 3, 1, 1, 1));
  dg.push_back(mk("3x3_s2_c64_o128", 2, 16, 16, 64, 128, 3, 2, 1, 1));
  dg.push_back(mk("3x3_s2_odd_c64_o128", 2, 13, 13, 64, 128, 3, 2, 1, 1));
  dg.push_back(mk("3x3_dil2_c16_o16", 2, 20, 20, 16, 16, 3, 1, 2, 2));
  dg.push_back(mk("1x1_c256_o18", 2, 13, 13, 256, 18, 1, 1, 0, 1));
  { Case c = mk("3x3_c64_o64_residual", 2, 13, 13, 64, 64, 3, 1, 1, 1); c.residual = true; dg.push_back(c); }
  dg.push_back(mk("3x3_c16_o32_many_tiles", 17, 80, 80, 16, 32, 3, 1, 1, 1));
  dg.push_back(mk("3x3_s2_c32_o64_many_tiles", 9, 104, 104, 32, 64, 3, 2, 1, 1));
  { Case c = mk("3x3_c64_o64_residual_many_tiles", 9, 80, 80, 64, 64, 3, 1, 1, 1); c.residual = true; dg.push_back(c); }
  for (auto& c : dg) run_dgrad(c);

  std::vector<Case> wg;
  wg.push_back(mk("1x1_c64_o128", 2, 13, 13, 64, 128, 1, 1, 0, 1));
  wg.push_back(mk("1x1_c128_o64", 2, 13, 13, 128, 64, 1, 1, 0, 1));
  wg.push_back(mk("3x3_c64_o128", 2, 13, 13, 64, 128, 3, 1, 1, 1));
  wg.push_back(mk("3x3_c128_o256", 4, 26, 26, 128, 256, 3, 1, 1, 1));
  wg.push_back(mk("3x3_s2_c64_o128", 2, 16, 16, 64, 128, 3, 2, 1, 1));
  wg.push_back(mk("3x3_c32_o64", 2, 13, 13, 32, 64, 3, 1, 1, 1));
  wg.push_back(mk("3x3_c16_o32", 2, 13, 13, 16, 32, 3, 1, 1, 1));
  wg.push_back(mk("3x3_c3_o32", 2, 16, 16, 3, 32, 3, 1, 1, 1));
  wg.push_back(mk("3x3_dil2_c16_o16", 2, 20, 20, 16, 16, 3, 1, 2, 2));
  wg.push_back(mk("7x7_c3_o16", 2, 20, 20, 3, 16, 7, 1, 3, 1));
  wg.push_back(mk("1x1_c256_o18", 2, 13, 13, 256, 18, 1, 1, 0, 1));
  wg.push_back(mk("1x1_c128_o7", 2, 20, 20, 128, 7, 1, 1, 0, 1));
  for (auto& c : wg) run_wgrad(c);

  printf("== %d cases, %d failed\n", g_run, g_fail);
  return g_fail ? 1 : 0;
}
