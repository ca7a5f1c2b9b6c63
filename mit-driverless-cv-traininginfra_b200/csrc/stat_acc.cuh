// Order-independent accumulation of per-channel statistics (BatchNorm sums, BN-backward sums).
//
// Many CTAs add their fp32 partial sums into the same [parts][2C] matrix.  With fp32 atomics the result depends on
// the arrival order, and a last-bit difference of a batch mean is amplified by a deep LeakyReLU network (sign flips)
// into percent-level run-to-run noise of the loss parts.  Here every partial is converted EXACTLY into a 128-bit
// fixed-point number held in two 64-bit words and added with integer atomics, which commute: the total -- and with
// it the whole forward pass -- is bit-reproducible whatever the scheduling.
//
//   v = w1 * 2^-20 + w2 * 2^-70        |v| < 2^31 per partial, exact for ulp(v) >= 2^-70
//
// (w1 = trunc(v * 2^20), w2 = the remainder * 2^70: at most 24 significant bits, so both conversions are exact.)
// Up to 2^11 partials may be added per entry without overflow.  A non-finite or out-of-range partial poisons the
// entry (w1 pinned near INT64_MIN), and the readers turn a poisoned entry into NaN, like a float sum would.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200cv {

struct StatAcc {
  long long w1, w2;
};
static_assert(sizeof(StatAcc) == 16, "StatAcc layout");

constexpr long long kStatPoison = (1LL << 62);

__device__ __forceinline__ void stat_add(StatAcc* p, float v) {
  if (v == 0.f) return;
  const double d = static_cast<double>(v);
  if (!(fabs(d) < 2147483648.0)) {  // also catches NaN
    atomicExch(reinterpret_cast<unsigned long long*>(&p->w1), 0x8000000000000000ull);
    return;
  }
  const long long a = static_cast<long long>(d * 1048576.0);  // truncation toward zero
  const double r = d - static_cast<double>(a) * (1.0 / 1048576.0);
  const long long b = __double2ll_rn(r * 0x1p70);
  if (a) atomicAdd(reinterpret_cast<unsigned long long*>(&p->w1), static_cast<unsigned long long>(a));
  if (b) atomicAdd(reinterpret_cast<unsigned long long*>(&p->w2), static_cast<unsigned long long>(b));
}

// Adds a block-level accumulator (shared memory) into a global one: integers all the way, poison preserved.
__device__ __forceinline__ void stat_merge(StatAcc* g, const StatAcc& s) {
  if (s.w1 > kStatPoison || s.w1 < -kStatPoison) {
    atomicExch(reinterpret_cast<unsigned long long*>(&g->w1), 0x8000000000000000ull);
    return;
  }
  if (s.w1) atomicAdd(reinterpret_cast<unsigned long long*>(&g->w1), static_cast<unsigned long long>(s.w1));
  if (s.w2) atomicAdd(reinterpret_cast<unsigned long long*>(&g->w2), static_cast<unsigned long long>(s.w2));
}

// Sum of `parts` rows of a [parts][row_entries] matrix at column `idx`, as a float.
__device__ __forceinline__ float stat_fold(const StatAcc* base, int parts, long long row_entries, long long idx) {
  long long w1 = 0, w2 = 0;
  bool bad = false;
  for (int p = 0; p < parts; ++p) {
    const longlong2 q = __ldg(reinterpret_cast<const longlong2*>(base + p * row_entries + idx));
    bad |= (q.x > kStatPoison) || (q.x < -kStatPoison);
    w1 += q.x;
    w2 += q.y;
  }
  if (bad) return __int_as_float(0x7fc00000);
  return static_cast<float>(static_cast<double>(w1) * (1.0 / 1048576.0) + static_cast<double>(w2) * 0x1p-70);
}

}  // namespace b200cv
