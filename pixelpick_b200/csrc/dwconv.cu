// T path, MobileNetV2 encoder — depthwise 3x3 convolution, NHWC bf16, forward / data gradient / weight gradient.
// Reference: InvertedResidual's `nn.Conv2d(hidden, hidden, 3, stride, 0, dilation, groups=hidden, bias=False)`
// (mobilenet_v2.py:33-35,46-48), which always runs on the explicitly pre-padded tensor of fixed_padding
// (mobilenet_v2.py:15-21,60-66) — so the kernels implement a VALID convolution (no implicit padding):
//   Ho = (Hi - 2*dil - 1) / stride + 1.
//
// 9 MACs per loaded element: HBM-bound in principle, but a naive "9 loads per output" kernel is bound by the
// L1/shared-memory data path (128 B/clk/SM) long before HBM — first versions of these kernels (16-byte channel
// vectors, weights in shared memory, one output per thread) ran at 1.0-1.3 TB/s.  So every kernel here reuses loaded
// vectors from REGISTERS: a thread owns 4 channels (8-byte vectors, the 36 weights of those channels live in
// registers for the whole kernel) and a strip of 8 (fwd/dgrad) or 4 (wgrad) pixels along x, sliding the 3x3 window
// over the columns it loaded: 3.75 loads per output instead of 9 (+9 weight reads).
#include "pp_common.cuh"

namespace pp {

constexpr int kDwThreads = 256;
constexpr int kDwTX = 8;   // outputs per thread along x (fwd, dgrad stride 1)
constexpr int kDwTXW = 4;  // pixels per thread along x (wgrad)

__device__ __forceinline__ void dw_unpack4(const uint2& v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xFFFF0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xFFFF0000u);
}
__device__ __forceinline__ uint2 dw_pack4(const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
__device__ __forceinline__ uint2 dw_ld(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

struct DwParams {
  const __nv_bfloat16* in;   // fwd: x [N][Hin][Win][C]; dgrad: dy; wgrad: x
  const __nv_bfloat16* in2;  // wgrad: dy [N][Hout][Wout][C]
  const float* w;            // [C][9] (fwd / dgrad)
  __nv_bfloat16* out;        // fwd: y [N][Hout][Wout][C]; dgrad: dx
  float* dw;                 // wgrad output [C][9] (zeroed by the launcher)
  int N, Hin, Win, Hout, Wout, C;
  const float* scale;        // fwd, inference: y = act(conv * scale[c] + shift[c]) (folded BatchNorm), or null
  const float* shift;
  int act;                   // 0 none, 1 ReLU, 2 ReLU6
  int off;                   // dgrad (stride 1): input coordinate = output - off + tap * dil  (off = 2 * dil)
  int flip;                  // dgrad: weights used flipped (tap 8 - t)
  // work decomposition: blockIdx.y = chunk of nq channel quads (<= 64); lanes = 256 / nq threads walk the CTA's
  // contiguous range of items; item = (tile, row in tile, x strip) with tiles of TYB x 2^TXB_log2 pixels of the OUTPUT
  // plane so vertically adjacent strips (which share 2 of their 3 input rows) run in the same CTA -> L1 hits
  int nq, TYB_log2, TXB_log2, tiles_y, tiles_x;
};

struct DwItem {
  int n, y, x0;
};
// item index -> (image, row, first column); strip_log2 = log2(pixels per strip)
__device__ __forceinline__ DwItem dw_item(const DwParams& p, uint32_t gi, int strip_log2) {
  const int xs_log2 = p.TXB_log2 - strip_log2;  // strips per tile row
  const int ipt_log2 = p.TYB_log2 + xs_log2;    // items per tile
  const uint32_t tile = gi >> ipt_log2, local = gi & ((1u << ipt_log2) - 1u);
  const uint32_t per_img = (uint32_t)p.tiles_y * (uint32_t)p.tiles_x;
  const uint32_t n = tile / per_img;
  const uint32_t r = tile - n * per_img;
  const uint32_t ty = r / (uint32_t)p.tiles_x;
  const uint32_t tx = r - ty * (uint32_t)p.tiles_x;
  DwItem it;
  it.n = (int)n;
  it.y = (int)((ty << p.TYB_log2) + (local >> xs_log2));
  it.x0 = (int)((tx << p.TXB_log2) + ((local & ((1u << xs_log2) - 1u)) << strip_log2));
  return it;
}
__device__ __forceinline__ uint32_t dw_num_items(const DwParams& p, int strip_log2) {
  return ((uint32_t)p.N * (uint32_t)p.tiles_y * (uint32_t)p.tiles_x) << (p.TYB_log2 + p.TXB_log2 - strip_log2);
}

// thread -> (channel quad of the chunk, lane); returns false for the idle remainder threads
__device__ __forceinline__ bool dw_thread(const DwParams& p, int& c0, int& lane, int& lanes) {
  const int Q = p.C >> 2;
  const int q0 = blockIdx.y * p.nq;
  const int nq = (Q - q0) < p.nq ? (Q - q0) : p.nq;
  lanes = kDwThreads / nq;
  const int ql = threadIdx.x % nq;
  lane = threadIdx.x / nq;
  c0 = (q0 + ql) * 4;
  return lane < lanes;
}

// ------------------------------------------------------------------------------------------------------
// forward (and, with off / flip, the stride-1 data gradient: a correlation of dy with the flipped kernel)
//   out[y, x] = sum_{ky,kx} w[ky,kx] * in[y*S - off + ky*D, x*S - off + kx*D]      (zero outside the input)
// ------------------------------------------------------------------------------------------------------
template <int S, int D>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_fwd_kernel(const DwParams p) {
  constexpr int NC = (kDwTX - 1) * S + 2 * D + 1;  // input columns feeding kDwTX outputs
  int c0, lane, lanes;
  if (!dw_thread(p, c0, lane, lanes)) return;
  float wr[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) wr[t][j] = __ldg(p.w + (size_t)(c0 + j) * 9 + (p.flip ? 8 - t : t));
  float esc[4] = {1.f, 1.f, 1.f, 1.f}, esf[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.scale) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      esc[j] = __ldg(p.scale + c0 + j);
      esf[j] = __ldg(p.shift + c0 + j);
    }
  }
  const uint32_t total = dw_num_items(p, 3);
  const uint32_t per_cta = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = blockIdx.x * per_cta;
  const uint32_t end = (begin + per_cta < total) ? begin + per_cta : total;
  for (uint32_t gi = begin + lane; gi < end; gi += lanes) {
    const DwItem it = dw_item(p, gi, 3);
    if (it.y >= p.Hout || it.x0 >= p.Wout) continue;
    float acc[kDwTX][4];
#pragma unroll
    for (int a = 0; a < kDwTX; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
    const int xi0 = it.x0 * S - p.off;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yi = it.y * S - p.off + ky * D;
      if (yi < 0 || yi >= p.Hin) continue;
      const __nv_bfloat16* row = p.in + (((int64_t)it.n * p.Hin + yi) * p.Win + xi0) * p.C + c0;
      uint2 v[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        v[c] = make_uint2(0u, 0u);
        if (xi0 + c >= 0 && xi0 + c < p.Win) v[c] = dw_ld(row + (int64_t)c * p.C);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float f[4];
        dw_unpack4(v[c], f);
#pragma unroll
        for (int a = 0; a < kDwTX; ++a)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            if (a * S + kx * D == c) {
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[a][j] = fmaf(f[j], wr[ky * 3 + kx][j], acc[a][j]);
            }
      }
    }
    if (p.scale) {
#pragma unroll
      for (int a = 0; a < kDwTX; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y = fmaf(acc[a][j], esc[j], esf[j]);
          if (p.act) y = fmaxf(y, 0.f);
          if (p.act == 2) y = fminf(y, 6.f);
          acc[a][j] = y;
        }
    }
    __nv_bfloat16* o = p.out + (((int64_t)it.n * p.Hout + it.y) * p.Wout + it.x0) * p.C + c0;
#pragma unroll
    for (int a = 0; a < kDwTX; ++a)
      if (it.x0 + a < p.Wout) *reinterpret_cast<uint2*>(o + (int64_t)a * p.C) = dw_pack4(acc[a]);
  }
}

// ------------------------------------------------------------------------------------------------------
// data gradient, stride 2 (dilation 1): dx[yi, xi] = sum_{ky,kx} w[ky,kx] * dy[(yi-ky)/2, (xi-kx)/2] over the taps with
// even (yi-ky), (xi-kx).  A thread owns a strip of 2 x 2-pixel quads of dx: rows 2a, 2a+1, columns 2b0 .. 2b0+3 — all
// of their taps read the 2 x 3 block dy[a-1..a][b0-1..b0+1]: 6 loads for 8 outputs.
//   dx(2a  , 2b  ) = w00 dy[a,b] + w02 dy[a,b-1] + w20 dy[a-1,b] + w22 dy[a-1,b-1]
//   dx(2a  , 2b+1) = w01 dy[a,b] + w21 dy[a-1,b]
//   dx(2a+1, 2b  ) = w10 dy[a,b] + w12 dy[a,b-1]
//   dx(2a+1, 2b+1) = w11 dy[a,b]
// Here in = dy [N][Hin][Win], out = dx [N][Hout][Wout]; items are (quad row a, strip of 2 quads).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_dgrad_s2_kernel(const DwParams p) {
  int c0, lane, lanes;
  if (!dw_thread(p, c0, lane, lanes)) return;
  float wr[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) wr[t][j] = __ldg(p.w + (size_t)(c0 + j) * 9 + t);
  const uint32_t total = dw_num_items(p, 1);  // tiles count QUADS; strip = 2 quads
  const uint32_t per_cta = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = blockIdx.x * per_cta;
  const uint32_t end = (begin + per_cta < total) ? begin + per_cta : total;
  for (uint32_t gi = begin + lane; gi < end; gi += lanes) {
    const DwItem it = dw_item(p, gi, 1);
    const int a = it.y, b0 = it.x0;  // quad coordinates
    if (2 * a >= p.Hout || 2 * b0 >= p.Wout) continue;
    float d[2][3][4];  // dy[a-1+r][b0-1+c]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int yy = a - 1 + r;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int xx = b0 - 1 + c;
        uint2 v = make_uint2(0u, 0u);
        if (yy >= 0 && yy < p.Hin && xx >= 0 && xx < p.Win)
          v = dw_ld(p.in + (((int64_t)it.n * p.Hin + yy) * p.Win + xx) * p.C + c0);
        dw_unpack4(v, d[r][c]);
      }
    }
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {  // quad b = b0 + qd uses columns c = qd (b-1) and qd + 1 (b)
      float o00[4], o01[4], o10[4], o11[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float cur = d[1][qd + 1][j], left = d[1][qd][j], up = d[0][qd + 1][j], upleft = d[0][qd][j];
        o00[j] = wr[0][j] * cur + wr[2][j] * left + wr[6][j] * up + wr[8][j] * upleft;
        o01[j] = wr[1][j] * cur + wr[7][j] * up;
        o10[j] = wr[3][j] * cur + wr[5][j] * left;
        o11[j] = wr[4][j] * cur;
      }
      const int yi = 2 * a, xi = 2 * (b0 + qd);
      __nv_bfloat16* o = p.out + (((int64_t)it.n * p.Hout + yi) * p.Wout + xi) * p.C + c0;
      if (xi < p.Wout) {
        *reinterpret_cast<uint2*>(o) = dw_pack4(o00);
        if (yi + 1 < p.Hout) *reinterpret_cast<uint2*>(o + (int64_t)p.Wout * p.C) = dw_pack4(o10);
      }
      if (xi + 1 < p.Wout) {
        *reinterpret_cast<uint2*>(o + p.C) = dw_pack4(o01);
        if (yi + 1 < p.Hout) *reinterpret_cast<uint2*>(o + (int64_t)p.Wout * p.C + p.C) = dw_pack4(o11);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// weight gradient: dW[c,ky,kx] = sum_{n,yo,xo} dY[n,yo,xo,c] * X[n,yo*S+ky*D,xo*S+kx*D,c]
// in = x, in2 = dy; a thread keeps the 36 accumulators of its 4 channels over its whole share of the pixels (strips of
// 4 along x sharing their input columns), then the CTA reduces through shared memory: one atomic per (channel, tap).
// ------------------------------------------------------------------------------------------------------
template <int S, int D>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_wgrad_kernel(const DwParams p) {
  constexpr int NC = (kDwTXW - 1) * S + 2 * D + 1;
  __shared__ float sh[kDwThreads * 4];
  int c0, lane, lanes;
  const bool active = dw_thread(p, c0, lane, lanes);
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  if (active) {
    const uint32_t total = dw_num_items(p, 2);
    const uint32_t per_cta = (total + gridDim.x - 1) / gridDim.x;
    const uint32_t begin = blockIdx.x * per_cta;
    const uint32_t end = (begin + per_cta < total) ? begin + per_cta : total;
    for (uint32_t gi = begin + lane; gi < end; gi += lanes) {
      const DwItem it = dw_item(p, gi, 2);
      if (it.y >= p.Hout || it.x0 >= p.Wout) continue;
      float d[kDwTXW][4];
      const __nv_bfloat16* dyp = p.in2 + (((int64_t)it.n * p.Hout + it.y) * p.Wout + it.x0) * p.C + c0;
#pragma unroll
      for (int a = 0; a < kDwTXW; ++a) {
        uint2 v = make_uint2(0u, 0u);  // zero gradient for the strip's out-of-range tail
        if (it.x0 + a < p.Wout) v = dw_ld(dyp + (int64_t)a * p.C);
        dw_unpack4(v, d[a]);
      }
      const int xi0 = it.x0 * S;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yi = it.y * S + ky * D;
        const __nv_bfloat16* row = p.in + (((int64_t)it.n * p.Hin + yi) * p.Win + xi0) * p.C + c0;
        uint2 v[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          v[c] = make_uint2(0u, 0u);
          if (xi0 + c < p.Win) v[c] = dw_ld(row + (int64_t)c * p.C);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float f[4];
          dw_unpack4(v[c], f);
#pragma unroll
          for (int a = 0; a < kDwTXW; ++a)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if (a * S + kx * D == c) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[ky * 3 + kx][j] = fmaf(f[j], d[a][j], acc[ky * 3 + kx][j]);
              }
        }
      }
    }
  }
  // block reduction: threads of one channel quad sit nq apart
  const int Q = p.C >> 2;
  const int q0 = blockIdx.y * p.nq;
  const int nq = (Q - q0) < p.nq ? (Q - q0) : p.nq;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) sh[threadIdx.x * 4 + j] = active ? acc[tp][j] : 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < nq * 4; t += kDwThreads) {
      const int qq = t >> 2, j = t & 3;
      float s = 0.f;
      for (int rr = 0; rr < lanes; ++rr) s += sh[(rr * nq + qq) * 4 + j];
      atomicAdd(p.dw + (size_t)((q0 + qq) * 4 + j) * 9 + tp, s);
    }
  }
}

// tiling of a plane of Hp x Wp units (pixels, or 2x2 quads for the stride-2 data gradient)
static void dw_tiling(DwParams& p, int Hp, int Wp, int* nchunks) {
  const int Q = p.C / 4;
  *nchunks = (Q + 63) / 64;
  p.nq = (Q + *nchunks - 1) / *nchunks;
  p.TYB_log2 = 3;
  p.TXB_log2 = p.nq <= 8 ? 6 : (p.nq <= 32 ? 5 : 4);  // fewer channels -> more lanes -> wider tiles
  p.tiles_y = (Hp + (1 << p.TYB_log2) - 1) >> p.TYB_log2;
  p.tiles_x = (Wp + (1 << p.TXB_log2) - 1) >> p.TXB_log2;
}

static int dw_check(const char* what, const void* a, const void* b, const void* c, int N, int Hi, int Wi, int C, int stride,
                    int dil, int* Ho, int* Wo) {
  PP_CHECK_ARG(a && b && c, "%s: null pointer", what);
  PP_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && C >= 8 && C % 8 == 0 && C <= 4096,
               "%s: bad shape (C=%d must be a multiple of 8, <= 4096)", what, C);
  PP_CHECK_ARG((stride == 1 && (dil == 1 || dil == 2 || dil == 4)) || (stride == 2 && dil == 1),
               "%s: stride=%d dil=%d unsupported (stride 1: dil 1/2/4; stride 2: dil 1)", what, stride, dil);
  PP_CHECK_ARG(Hi >= 2 * dil + 1 && Wi >= 2 * dil + 1, "%s: input %dx%d smaller than the dilated kernel", what, Hi, Wi);
  PP_CHECK_ARG((int64_t)N * (Hi + 8) * (Wi + 64) * (C / 4) < (1ll << 31), "%s: tensor too large for 32-bit item indices", what);
  *Ho = (Hi - 2 * dil - 1) / stride + 1;
  *Wo = (Wi - 2 * dil - 1) / stride + 1;
  return PP_OK;
}

static dim3 dw_grid(const DwParams& p, int nchunks, int strip_log2, int ctas_per_sm) {
  const int64_t items = ((int64_t)p.N * p.tiles_y * p.tiles_x) << (p.TYB_log2 + p.TXB_log2 - strip_log2);
  const int lanes = kDwThreads / p.nq;
  int64_t gx = (items + lanes - 1) / lanes;  // at least one item per lane
  const int64_t cap = (int64_t)148 * ctas_per_sm / nchunks;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)nchunks);
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C, int stride, int dil,
                     void* stream) {
  return pp_dwconv3x3_fwd_bnact(x, w, nullptr, nullptr, 0, y, N, Hi, Wi, C, stride, dil, stream);
}

int pp_dwconv3x3_fwd_bnact(const void* x, const float* w, const float* scale, const float* shift, int act, void* y, int N,
                           int Hi, int Wi, int C, int stride, int dil, void* stream) {
  PP_CHECK_ARG((scale == nullptr) == (shift == nullptr) && act >= 0 && act <= 2, "pp_dwconv3x3_fwd_bnact: scale/shift come in pairs, act 0..2");
  DwParams p{};
  p.scale = scale; p.shift = shift; p.act = act;
  int Ho, Wo;
  int rc = dw_check("pp_dwconv3x3_fwd", x, w, y, N, Hi, Wi, C, stride, dil, &Ho, &Wo);
  if (rc != PP_OK) return rc;
  p.in = reinterpret_cast<const __nv_bfloat16*>(x); p.w = w; p.out = reinterpret_cast<__nv_bfloat16*>(y);
  p.N = N; p.Hin = Hi; p.Win = Wi; p.Hout = Ho; p.Wout = Wo; p.C = C; p.off = 0; p.flip = 0;
  int nchunks;
  dw_tiling(p, Ho, Wo, &nchunks);
  const dim3 grid = dw_grid(p, nchunks, 3, 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (stride == 2) dwconv_fwd_kernel<2, 1><<<grid, kDwThreads, 0, st>>>(p);
  else if (dil == 1) dwconv_fwd_kernel<1, 1><<<grid, kDwThreads, 0, st>>>(p);
  else if (dil == 2) dwconv_fwd_kernel<1, 2><<<grid, kDwThreads, 0, st>>>(p);
  else dwconv_fwd_kernel<1, 4><<<grid, kDwThreads, 0, st>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream) {
  DwParams p{};
  int Ho, Wo;
  int rc = dw_check("pp_dwconv3x3_dgrad", dy, w, dx, N, Hi, Wi, C, stride, dil, &Ho, &Wo);
  if (rc != PP_OK) return rc;
  p.in = reinterpret_cast<const __nv_bfloat16*>(dy); p.w = w; p.out = reinterpret_cast<__nv_bfloat16*>(dx);
  p.N = N; p.Hin = Ho; p.Win = Wo; p.Hout = Hi; p.Wout = Wi; p.C = C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int nchunks;
  if (stride == 1) {
    // correlation of dy (zero outside) with the flipped kernel, origin shifted by 2 * dil
    p.off = 2 * dil; p.flip = 1;
    dw_tiling(p, Hi, Wi, &nchunks);
    const dim3 grid = dw_grid(p, nchunks, 3, 8);
    if (dil == 1) dwconv_fwd_kernel<1, 1><<<grid, kDwThreads, 0, st>>>(p);
    else if (dil == 2) dwconv_fwd_kernel<1, 2><<<grid, kDwThreads, 0, st>>>(p);
    else dwconv_fwd_kernel<1, 4><<<grid, kDwThreads, 0, st>>>(p);
  } else {
    dw_tiling(p, (Hi + 1) / 2, (Wi + 1) / 2, &nchunks);  // planes of 2x2 quads
    const dim3 grid = dw_grid(p, nchunks, 1, 8);
    dwconv_dgrad_s2_kernel<<<grid, kDwThreads, 0, st>>>(p);
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream) {
  DwParams p{};
  int Ho, Wo;
  int rc = dw_check("pp_dwconv3x3_wgrad", x, dy, dw, N, Hi, Wi, C, stride, dil, &Ho, &Wo);
  if (rc != PP_OK) return rc;
  p.in = reinterpret_cast<const __nv_bfloat16*>(x); p.in2 = reinterpret_cast<const __nv_bfloat16*>(dy); p.dw = dw;
  p.N = N; p.Hin = Hi; p.Win = Wi; p.Hout = Ho; p.Wout = Wo; p.C = C;
  int nchunks;
  dw_tiling(p, Ho, Wo, &nchunks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PP_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * 9 * sizeof(float), st));
  const dim3 grid = dw_grid(p, nchunks, 2, 2);  // one resident wave: the per-CTA reduction + atomics are the fixed cost
  if (stride == 2) dwconv_wgrad_kernel<2, 1><<<grid, kDwThreads, 0, st>>>(p);
  else if (dil == 1) dwconv_wgrad_kernel<1, 1><<<grid, kDwThreads, 0, st>>>(p);
  else if (dil == 2) dwconv_wgrad_kernel<1, 2><<<grid, kDwThreads, 0, st>>>(p);
  else dwconv_wgrad_kernel<1, 4><<<grid, kDwThreads, 0, st>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"
