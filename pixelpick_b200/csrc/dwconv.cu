// T path, MobileNetV2 encoder — depthwise 3x3 convolution, NHWC bf16, forward / data gradient / weight gradient.
// Reference: InvertedResidual's `nn.Conv2d(hidden, hidden, 3, stride, 0, dilation, groups=hidden, bias=False)`
// (mobilenet_v2.py:33-35,46-48), which always runs on the explicitly pre-padded tensor of fixed_padding
// (mobilenet_v2.py:15-21,60-66) — so the kernels implement a VALID convolution (no implicit padding):
//   Ho = (Hi - 2*dil - 1) / stride + 1.
// Depthwise convs have 9 MACs per loaded element: they are HBM-bound.  One thread owns 8 consecutive channels
// (one 16-byte vector); the forward slides a window along x so each input vector is loaded once per row it feeds.
// Weights are read as the module's fp32 [C][1][3][3] tensor directly; the weight gradient is written in that layout.
#include "pp_common.cuh"

namespace pp {

constexpr int kDwThreads = 256;
constexpr int kDwTX = 4;  // outputs per thread along x (forward)

__device__ __forceinline__ void dw_unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 dw_pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct DwParams {
  const __nv_bfloat16* x;   // [N][Hi][Wi][C]
  const float* w;           // [C][9]
  const __nv_bfloat16* dy;  // [N][Ho][Wo][C]   (dgrad / wgrad)
  __nv_bfloat16* y;         // forward output [N][Ho][Wo][C]; dgrad output [N][Hi][Wi][C]
  float* dw;                // wgrad output [C][9] (zeroed by the launcher)
  int N, Hi, Wi, Ho, Wo, C, stride, dil;
};

// weights staged once per CTA in shared memory, transposed to [tap][C] so a thread reads its 8 channels of a tap as
// two conflict-free 16-byte loads (keeping all 72 weights in registers capped occupancy at 8 warps / SM)
__device__ __forceinline__ void dw_stage_weights(const float* __restrict__ w, int C, float* sw) {
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) {
    const int c = i / 9, t = i - c * 9;
    sw[t * C + c] = __ldg(w + i);
  }
  __syncthreads();
}
__device__ __forceinline__ void dw_tap_weights(const float* sw, int C, int tap, int c0, float (&wv)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(sw + tap * C + c0);
  const float4 b = *reinterpret_cast<const float4*>(sw + tap * C + c0 + 4);
  wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w;
  wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
}

template <int S, int D>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_fwd_kernel(const DwParams p) {
  extern __shared__ float sw[];
  dw_stage_weights(p.w, p.C, sw);
  constexpr int NC = (kDwTX - 1) * S + 2 * D + 1;  // input columns feeding kDwTX outputs
  const int groups = p.C >> 3;
  const int xtiles = (p.Wo + kDwTX - 1) / kDwTX;
  const int64_t total = (int64_t)p.N * p.Ho * xtiles * groups;
  for (int64_t i = (int64_t)blockIdx.x * kDwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kDwThreads) {
    const int g = (int)(i % groups);
    int64_t t = i / groups;
    const int xt = (int)(t % xtiles);
    t /= xtiles;
    const int yo = (int)(t % p.Ho);
    const int n = (int)(t / p.Ho);
    const int xo0 = xt * kDwTX;
    float acc[kDwTX][8];
#pragma unroll
    for (int a = 0; a < kDwTX; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    const int xi0 = xo0 * S;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float wr[3][8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) dw_tap_weights(sw, p.C, ky * 3 + kx, g * 8, wr[kx]);
      const int yi = yo * S + ky * D;
      const __nv_bfloat16* row = p.x + (((int64_t)n * p.Hi + yi) * p.Wi + xi0) * p.C + g * 8;
      uint4 v[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        v[c] = make_uint4(0u, 0u, 0u, 0u);
        if (xi0 + c < p.Wi) v[c] = __ldg(reinterpret_cast<const uint4*>(row + (int64_t)c * p.C));
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float f[8];
        dw_unpack8(v[c], f);
#pragma unroll
        for (int a = 0; a < kDwTX; ++a)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            if (a * S + kx * D == c) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[a][j] = fmaf(f[j], wr[kx][j], acc[a][j]);
            }
      }
    }
    __nv_bfloat16* o = p.y + (((int64_t)n * p.Ho + yo) * p.Wo + xo0) * p.C + g * 8;
#pragma unroll
    for (int a = 0; a < kDwTX; ++a)
      if (xo0 + a < p.Wo) *reinterpret_cast<uint4*>(o + (int64_t)a * p.C) = dw_pack8(acc[a]);
  }
}

// data gradient (gather form): dX[n,yi,xi,c] = sum_{ky,kx} W[c,ky,kx] * dY[n,(yi-ky*D)/S,(xi-kx*D)/S,c]
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_dgrad_kernel(const DwParams p) {
  extern __shared__ float sw[];
  dw_stage_weights(p.w, p.C, sw);
  const int groups = p.C >> 3;
  const int64_t total = (int64_t)p.N * p.Hi * p.Wi * groups;
  const int S = p.stride, D = p.dil;
  for (int64_t i = (int64_t)blockIdx.x * kDwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kDwThreads) {
    const int g = (int)(i % groups);
    int64_t t = i / groups;
    const int xi = (int)(t % p.Wi);
    t /= p.Wi;
    const int yi = (int)(t % p.Hi);
    const int n = (int)(t / p.Hi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* base = p.dy + (int64_t)n * p.Ho * p.Wo * p.C + g * 8;
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = yi - ky * D;
      const int yo = ty / S;
      const bool oky = ty >= 0 && (ty - yo * S) == 0 && yo < p.Ho;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = xi - kx * D;
        const int xo = tx / S;
        const bool okx = tx >= 0 && (tx - xo * S) == 0 && xo < p.Wo;
        ok[ky * 3 + kx] = oky && okx;
        v[ky * 3 + kx] = make_uint4(0u, 0u, 0u, 0u);
        if (oky && okx) v[ky * 3 + kx] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)yo * p.Wo + xo) * p.C));
      }
    }
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      if (!ok[tp]) continue;
      float f[8], wv[8];
      dw_unpack8(v[tp], f);
      dw_tap_weights(sw, p.C, tp, g * 8, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wv[j], acc[j]);
    }
    *reinterpret_cast<uint4*>(p.y + (((int64_t)n * p.Hi + yi) * p.Wi + xi) * p.C + g * 8) = dw_pack8(acc);
  }
}

// weight gradient: dW[c,ky,kx] = sum_{n,yo,xo} dY[n,yo,xo,c] * X[n,yo*S+ky*D,xo*S+kx*D,c]
// thread -> fixed 8-channel group, a contiguous run of output pixels per block (neighbouring pixels share input
// columns: L1 reuse); 72 fp32 accumulators per thread, block reduction through shared memory, one atomic per block.
__global__ void __launch_bounds__(kDwThreads) dwconv_wgrad_kernel(const DwParams p, int64_t px_per_block) {
  __shared__ float sh[kDwThreads * 8];
  const int groups = p.C >> 3;
  const int rows_per_block = kDwThreads / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int S = p.stride, D = p.dil;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * px_per_block;
  const int64_t m1 = (m0 + px_per_block < M) ? m0 + px_per_block : M;
  if (r < rows_per_block) {
    for (int64_t m = m0 + r; m < m1; m += rows_per_block) {
      const int xo = (int)(m % p.Wo);
      const int64_t t = m / p.Wo;
      const int yo = (int)(t % p.Ho);
      const int n = (int)(t / p.Ho);
      float d[8];
      dw_unpack8(__ldg(reinterpret_cast<const uint4*>(p.dy + m * p.C + g * 8)), d);
      const __nv_bfloat16* xb = p.x + (((int64_t)n * p.Hi + (int64_t)yo * S) * p.Wi + (int64_t)xo * S) * p.C + g * 8;
      uint4 v[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          v[ky * 3 + kx] = __ldg(reinterpret_cast<const uint4*>(xb + ((int64_t)ky * D * p.Wi + kx * D) * p.C));
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        float f[8];
        dw_unpack8(v[tp], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[tp][j] = fmaf(f[j], d[j], acc[tp][j]);
      }
    }
  }
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = acc[tp][j];
    __syncthreads();
    for (int t = threadIdx.x; t < groups * 8; t += kDwThreads) {
      const int gg = t >> 3, j = t & 7;
      float s = 0.f;
      for (int rr = 0; rr < rows_per_block; ++rr) s += sh[(rr * groups + gg) * 8 + j];
      atomicAdd(p.dw + (size_t)(gg * 8 + j) * 9 + tp, s);
    }
  }
}

static inline int dw_grid(int64_t total) {
  int64_t b = (total + kDwThreads - 1) / kDwThreads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

static int dw_check(const char* what, const void* a, const void* b, const void* c, int N, int Hi, int Wi, int C, int stride,
                    int dil, int* Ho, int* Wo) {
  PP_CHECK_ARG(a && b && c, "%s: null pointer", what);
  PP_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && C >= 8 && C % 8 == 0 && C <= 2048, "%s: bad shape (C=%d must be a multiple of 8, <= 2048)",
               what, C);
  PP_CHECK_ARG((stride == 1 || stride == 2) && (dil == 1 || dil == 2 || dil == 4), "%s: stride=%d dil=%d unsupported", what,
               stride, dil);
  PP_CHECK_ARG(Hi >= 2 * dil + 1 && Wi >= 2 * dil + 1, "%s: input %dx%d smaller than the dilated kernel", what, Hi, Wi);
  *Ho = (Hi - 2 * dil - 1) / stride + 1;
  *Wo = (Wi - 2 * dil - 1) / stride + 1;
  return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C, int stride, int dil,
                     void* stream) {
  DwParams p{};
  int rc = dw_check("pp_dwconv3x3_fwd", x, w, y, N, Hi, Wi, C, stride, dil, &p.Ho, &p.Wo);
  if (rc != PP_OK) return rc;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.w = w; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.N = N; p.Hi = Hi; p.Wi = Wi; p.C = C; p.stride = stride; p.dil = dil;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)N * p.Ho * ((p.Wo + kDwTX - 1) / kDwTX) * (C / 8);
  const int grid = dw_grid(total);
  const size_t sm = (size_t)C * 9 * sizeof(float);  // <= 72 KB at C = 2048
  if (sm > 48 * 1024) {
    static bool attr = false;
    if (!attr) {
      PP_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      PP_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      PP_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      PP_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      PP_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      attr = true;
    }
  }
  if (stride == 1 && dil == 1) dwconv_fwd_kernel<1, 1><<<grid, kDwThreads, sm, st>>>(p);
  else if (stride == 2 && dil == 1) dwconv_fwd_kernel<2, 1><<<grid, kDwThreads, sm, st>>>(p);
  else if (stride == 1 && dil == 2) dwconv_fwd_kernel<1, 2><<<grid, kDwThreads, sm, st>>>(p);
  else if (stride == 1 && dil == 4) dwconv_fwd_kernel<1, 4><<<grid, kDwThreads, sm, st>>>(p);
  else if (stride == 2 && dil == 2) dwconv_fwd_kernel<2, 2><<<grid, kDwThreads, sm, st>>>(p);
  else {
    set_error("pp_dwconv3x3_fwd: stride=%d dil=%d unsupported", stride, dil);
    return PP_ERR_INVALID_ARG;
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream) {
  DwParams p{};
  int rc = dw_check("pp_dwconv3x3_dgrad", dy, w, dx, N, Hi, Wi, C, stride, dil, &p.Ho, &p.Wo);
  if (rc != PP_OK) return rc;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.w = w; p.y = reinterpret_cast<__nv_bfloat16*>(dx);
  p.N = N; p.Hi = Hi; p.Wi = Wi; p.C = C; p.stride = stride; p.dil = dil;
  const size_t sm = (size_t)C * 9 * sizeof(float);
  if (sm > 48 * 1024) {
    static bool attr = false;
    if (!attr) {
      PP_CUDA(cudaFuncSetAttribute(dwconv_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      attr = true;
    }
  }
  dwconv_dgrad_kernel<<<dw_grid((int64_t)N * Hi * Wi * (C / 8)), kDwThreads, sm, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream) {
  DwParams p{};
  int rc = dw_check("pp_dwconv3x3_wgrad", x, dy, dw, N, Hi, Wi, C, stride, dil, &p.Ho, &p.Wo);
  if (rc != PP_OK) return rc;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.dw = dw;
  p.N = N; p.Hi = Hi; p.Wi = Wi; p.C = C; p.stride = stride; p.dil = dil;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PP_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * 9 * sizeof(float), st));
  const int64_t M = (int64_t)N * p.Ho * p.Wo;
  const int rows_per_block = kDwThreads / (C / 8);
  // enough blocks to fill the machine, each with a contiguous run of >= 8 rounds of pixels
  int64_t blocks = 148 * 8;
  int64_t per = (M + blocks - 1) / blocks;
  const int64_t min_per = (int64_t)rows_per_block * 8;
  if (per < min_per) per = min_per;
  per = (per + rows_per_block - 1) / rows_per_block * rows_per_block;
  blocks = (M + per - 1) / per;
  dwconv_wgrad_kernel<<<(int)blocks, kDwThreads, 0, st>>>(p, per);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"
