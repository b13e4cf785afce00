// Shared helpers for the pixelpick_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/pixelpick_b200.h"

namespace pp {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;  // kernels launched by this library (bench.py's gpu_launches)

#define PP_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::pp::set_error(__VA_ARGS__);             \
      return PP_ERR_INVALID_ARG;                \
    }                                           \
  } while (0)

#define PP_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::pp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return PP_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define PP_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    ++::pp::g_launches;                                                                 \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::pp::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return PP_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// ---- ordering key -------------------------------------------------------------------------
// Maps a float score to a uint32 whose ASCENDING order is the selection order:
//   largest  : descending score, NaN first (torch.topk treats NaN as the largest value)
//   !largest : ascending score, NaN last
// -0.0 and +0.0 compare equal (canonicalised), every NaN payload compares equal.
__host__ __device__ inline uint32_t ord_key(float s, bool largest) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  s = s + 0.0f;  // -0.0 -> +0.0 (IEEE: not an identity, the compiler keeps it)
  u = __float_as_uint(s);
#else
  s = s + 0.0f;
  union { float f; uint32_t u; } cv; cv.f = s; u = cv.u;
#endif
  if (s != s) u = 0xFFFFFFFFu;
  else u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return largest ? ~u : u;
}

// Level-0 bucket of the radix select: a LINEAR quantisation of the score instead of the key's leading bits.  Every
// acquisition score lives in [0, ln C] (entropy) or [0, 1] (least confidence, margin), where the leading 11 float bits
// resolve poorly: [0.5, 1) is only 4 buckets, so the "largest" strategies left ~25 % of an image in the boundary bucket
// for the single-CTA radix tail.  2047 linear buckets over [0, 4) put < 1 % there (used for largest-first selections:
// entropy, least confidence).  Any monotone map keeps the select
// exact (ascending bucket <=> ascending ordering key, ties in a bucket are resolved on the full key afterwards); out of
// range scores are clamped (still monotone, just coarse).  NaN: bucket 0 for largest (first), 2047 otherwise (last).
__host__ __device__ inline uint32_t ord_key(float s, bool largest);
__host__ __device__ inline uint32_t bucket0(float s, bool largest) {
  // smallest-first selections (margin, random) pick values near 0, where the float's own exponent bits already spread the
  // candidates over many buckets: keep the key's leading 11 bits there (measured: the linear map cost them 3 %)
  if (!largest) return ord_key(s, false) >> 21;
  if (s != s) return largest ? 0u : 2047u;
  // round(clamp(s, 0, 4) * 511.5) read out of the mantissa after adding 2^23 (one FFMA + one mask; a float->int
  // conversion runs on the quarter-rate pipe and the scoring kernel is close to issue-bound).  Monotone in s.
  const float c = s < 0.f ? 0.f : (s > 4.f ? 4.f : s);
  const float t = c * 511.5f + 8388608.0f;
#ifdef __CUDA_ARCH__
  const uint32_t q = __float_as_uint(t) & 0x7FFFFFu;
#else
  union { float f; uint32_t u; } cv; cv.f = t; const uint32_t q = cv.u & 0x7FFFFFu;
#endif
  return largest ? 2047u - q : q;
}

__host__ __device__ inline float ord_key_inv(uint32_t k, bool largest) {
  uint32_t u = largest ? ~k : k;
  if (u == 0xFFFFFFFFu) u = 0x7FC00000u;
  else u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } cv; cv.u = u; return cv.f;
#endif
}

// Bilinear source index/weights for one axis, align_corners=True — restates ATen's
// compute_source_index_and_lambda / guard_index_and_lambda (aten/native/UpSample.h), which is what
// F.interpolate(..., mode='bilinear', align_corners=True) runs (deeplab.py:49,55,58).
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_ac(int o, int in_size, int out_size, float scale) {
  Lerp r;
  if (in_size == out_size) {
    r.i0 = r.i1 = o;
    r.l0 = 1.f;
    r.l1 = 0.f;
    return r;
  }
  const float real = scale * (float)o;
  int i0 = (int)real;
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l1 = real - (float)i0;
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  r.i0 = i0;
  r.i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  r.l1 = l1;
  r.l0 = 1.f - l1;
  return r;
}
// ATen area_pixel_compute_scale(align_corners=True)
static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pp
