// Library-level plumbing of libb200cv.so: thread-local error message, per-device error word,
// cached device properties.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "internal.h"

namespace b200cv {

namespace {
thread_local char g_err[512] = "";
std::mutex g_mu;
int g_sm_count[64] = {0};
int* g_err_word[64] = {nullptr};
void* g_identity[64] = {nullptr};
}  // namespace

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

bool pdl_enabled() {
  // opt-in: measured 1 % SLOWER on the Darknet-53 step (graph replay already hides launch gaps; early CTAs of the
  // dependent kernel only take resources from the tail of its predecessor)
  static const bool on = getenv("B200CV_PDL") && atoi(getenv("B200CV_PDL")) != 0;
  return on;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_sm_count[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = n > 0 ? n : 148;
  }
  return g_sm_count[dev];
}

int* device_error_word() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_err_word[dev]) {
    // 4 bytes of library-owned immutable-lifetime state per device (see ownership rule in b200cv.h)
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(int));
    g_err_word[dev] = p;
  }
  return g_err_word[dev];
}

const void* device_identity128() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_identity[dev]) {
    // 32 KB of library-owned immutable state per device: the A operand of the "residual by tensor core" trick
    static unsigned short host[128 * 128];
    for (int i = 0; i < 128 * 128; ++i) host[i] = (i / 128 == i % 128) ? 0x3F80 : 0;  // bf16 1.0
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(host)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(p, host, sizeof(host), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    g_identity[dev] = p;
  }
  return g_identity[dev];
}

}  // namespace b200cv

extern "C" {

const char* b200cv_version(void) { return "b200cv 0.1 (sm_100a)"; }

const char* b200cv_last_error(void) { return b200cv::g_err; }

int b200cv_pad_channels(int c) { return b200cv::pad_channels(c); }

int b200cv_check_device_error(void* stream) {
  int* w = b200cv::device_error_word();
  if (!w) return b200cv::set_error(B200CV_ERR_DEVICE, "no device error word");
  int h = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(&h, w, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess)
    return b200cv::set_error(static_cast<int>(e), "check_device_error: %s", cudaGetErrorString(e));
  if (h != 0) {
    cudaMemsetAsync(w, 0, sizeof(int), s);
    cudaStreamSynchronize(s);
    return b200cv::set_error(B200CV_ERR_DEVICE, "pipeline wait timed out (role code %d)", h);
  }
  return 0;
}

int b200cv_split_pieces(void) { return b200cv::kSplitPieces; }

}  // extern "C"
