// T path - the ResNet stem's max-pool (nn.MaxPool2d(3, 2, 1), resnet_models.py:116 / resnet_backbone.py:56) on NHWC bf16,
// forward (+ the argmax code of every output, one byte) and backward.  HBM-bound: a thread owns 8 channels (one 16-byte
// vector) of one pixel.  Forward: 9 taps, first maximum in row-major window order wins, NaN propagates (ATen's rule).
// Backward in GATHER form - every input pixel looks at the <= 4 windows that contain it and takes the gradient of those whose
// argmax code points back at it - so no atomics and no zero-fill of the 4x larger input gradient.
#include "pp_common.cuh"

namespace pp {

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                                              int Ho, int Wo, __nv_bfloat16* __restrict__ y,
                                                              uint8_t* __restrict__ code, int64_t total) {
  const int cv = C / 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    bool first = true;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)n * H + iy) * W + ix) * C) + c8);
        float v[8];
        unpack8(q, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // ATen: if (val > maxval || isnan(val)) -> take it; the first tap always initialises
          if (first || v[e] > best[e] || v[e] != v[e]) { best[e] = v[e]; arg[e] = ky * 3 + kx; }
        }
        first = false;
      }
    }
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) ow[e] = (__float_as_uint(best[2 * e]) >> 16) | (__float_as_uint(best[2 * e + 1]) & 0xFFFF0000u);
    reinterpret_cast<uint4*>(y + (((int64_t)n * Ho + oy) * Wo + ox) * C)[c8] = o;
    if (code) {
      uint2 cc;
      cc.x = (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
      cc.y = (uint32_t)arg[4] | ((uint32_t)arg[5] << 8) | ((uint32_t)arg[6] << 16) | ((uint32_t)arg[7] << 24);
      reinterpret_cast<uint2*>(code + (((int64_t)n * Ho + oy) * Wo + ox) * C)[c8] = cc;
    }
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ code,
                                                              int N, int H, int W, int C, int Ho, int Wo,
                                                              __nv_bfloat16* __restrict__ dx, int64_t total) {
  const int cv = C / 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // windows (oy, ox) that contain (iy, ix): 2 oy - 1 <= iy <= 2 oy + 1
    const int oy_lo = iy / 2, oy_hi = (iy + 1) / 2, ox_lo = ix / 2, ox_hi = (ix + 1) / 2;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      if (oy >= Ho) continue;
      const int ky = iy - (2 * oy - 1);
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        if (ox >= Wo) continue;
        const int want = ky * 3 + (ix - (2 * ox - 1));
        const int64_t o = (((int64_t)n * Ho + oy) * Wo + ox) * C;
        const uint2 cc = __ldg(reinterpret_cast<const uint2*>(code + o) + c8);
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(dy + o) + c8);
        float g[8];
        unpack8(q, g);
        const uint32_t cw[2] = {cc.x, cc.y};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if ((int)((cw[e >> 2] >> (8 * (e & 3))) & 0xFFu) == want) acc[e] += g[e];
      }
    }
    uint4 o4;
    __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
    for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    reinterpret_cast<uint4*>(dx + (((int64_t)n * H + iy) * W + ix) * C)[c8] = o4;
  }
}

static inline int pool_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace pp

extern "C" {

int pp_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* y, unsigned char* code, void* stream) {
  using namespace pp;
  PP_CHECK_ARG(x && y, "pp_maxpool3x3s2_fwd: null pointer");
  PP_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pp_maxpool3x3s2_fwd: bad shape (C must be a multiple of 8)");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)N * Ho * Wo * (C / 8);
  maxpool3x3s2_fwd_kernel<<<pool_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(y), code, total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_maxpool3x3s2_bwd(const void* dy, const unsigned char* code, int N, int H, int W, int C, void* dx, void* stream) {
  using namespace pp;
  PP_CHECK_ARG(dy && code && dx, "pp_maxpool3x3s2_bwd: null pointer");
  PP_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pp_maxpool3x3s2_bwd: bad shape (C must be a multiple of 8)");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)N * H * W * (C / 8);
  maxpool3x3s2_bwd_kernel<<<pool_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), code, N, H, W, C, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(dx), total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"
