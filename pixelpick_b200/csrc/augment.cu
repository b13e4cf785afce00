// Input pipeline on the device: the reference's joint geometric augmentation, datasets/base_dataset.py:48-127 -
// random scale (image: PIL BILINEAR, label map: PIL NEAREST, query / human-label masks: torch nearest) -> pad to the crop size
// -> random crop -> horizontal flip - plus TF.to_tensor + TF.normalize (base_dataset.py:183), for a whole batch in ONE launch.
// The random draws stay on the host (Python's `random`, same call order as the reference); the host also builds the per-axis
// resampling tables in double precision exactly as Pillow does (Resample.c:precompute_coeffs / normalize_coeffs_8bpc, 22-bit
// fixed point; Geometry.c's running-sum nearest indices), so the kernel is integer arithmetic on those tables and the result
// equals PIL's two-pass resample (horizontal pass rounded to uint8, then vertical) bit for bit.
// One thread = one output pixel: it evaluates the horizontal pass for each of the <= 5 source rows its vertical filter needs.
#include "pp_common.cuh"

namespace pp {

constexpr int kAugHdr = 20;
constexpr int kPrecisionBits = 32 - 8 - 2;

struct AugParams {
  const uint8_t* x;   // [B][H][W][3]
  const uint8_t* y;   // [B][H][W] or null
  const uint8_t* q;   // [B][H][W] or null (0 / 255 or 0 / 1: any non-zero = labelled)
  const uint8_t* lq;  // [B][H][W] or null
  int B, H, W, crop_h, crop_w;
  const int32_t* header;  // [B][kAugHdr]
  const int32_t* tables;  // packed per-sample tables, offsets in the header
  float mean[3], stdv[3];
  int mean_val[3], ignore_index;
  float* x_out;    // [B][3][crop_h][crop_w] (or null)
  uint8_t* x_u8;   // [B][crop_h][crop_w][3]: the augmented image before to_tensor / normalize (or null)
  uint8_t* y_out;  // [B][crop_h][crop_w]
  uint8_t* q_out;
  uint8_t* lq_out;
};

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) augment_geometric_kernel(const AugParams p) {
  const int b = blockIdx.z;
  const int oy = blockIdx.y;
  const int ox = blockIdx.x * 256 + threadIdx.x;
  if (ox >= p.crop_w) return;
  const int32_t* h = p.header + (size_t)b * kAugHdr;
  const int h_rs = h[0], w_rs = h[1], start_h = h[2], start_w = h[3], flip = h[4], ksx = h[5], ksy = h[6];
  const int32_t* t = p.tables;
  const int px = (flip ? p.crop_w - 1 - ox : ox) + start_w;  // hflip is the LAST step: output column ox shows crop column cw-1-ox
  const int py = oy + start_h;
  const size_t o = ((size_t)b * p.crop_h + oy) * p.crop_w + ox;
  const size_t plane = (size_t)p.crop_h * p.crop_w;
  const bool inside = py < h_rs && px < w_rs;  // else: the pad region (right / bottom)
  // ---- image: vertical pass over horizontally resampled rows ----
  int rgb[3] = {p.mean_val[0], p.mean_val[1], p.mean_val[2]};
  if (inside) {
    const uint8_t* img = p.x + (size_t)b * p.H * p.W * 3;
    const int xmin = t[h[11] + px], xn = t[h[12] + px];
    const int32_t* kx = t + h[13] + (size_t)px * ksx;
    const int ymin = t[h[14] + py], yn = t[h[15] + py];
    const int32_t* ky = t + h[16] + (size_t)py * ksy;
    const bool need_h = w_rs != p.W, need_v = h_rs != p.H;
    int acc[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
    const int r_lo = need_v ? ymin : py, r_n = need_v ? yn : 1;
    for (int r = 0; r < r_n; ++r) {
      const uint8_t* row = img + (size_t)(r_lo + r) * p.W * 3;
      int v[3];
      if (need_h) {
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int c = 0; c < xn; ++c) {
          const uint8_t* s = row + (size_t)(xmin + c) * 3;
          const int k = kx[c];
          a0 += (int)s[0] * k;
          a1 += (int)s[1] * k;
          a2 += (int)s[2] * k;
        }
        v[0] = clip8(a0); v[1] = clip8(a1); v[2] = clip8(a2);  // the horizontal pass is stored as uint8 by PIL
      } else {
        const uint8_t* s = row + (size_t)px * 3;
        v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
      }
      if (need_v) {
        const int k = ky[r];
        acc[0] += v[0] * k; acc[1] += v[1] * k; acc[2] += v[2] * k;
      } else {
        rgb[0] = v[0]; rgb[1] = v[1]; rgb[2] = v[2];
      }
    }
    if (need_v) { rgb[0] = clip8(acc[0]); rgb[1] = clip8(acc[1]); rgb[2] = clip8(acc[2]); }
  }
  if (p.x_out) {
#pragma unroll
    for (int c = 0; c < 3; ++c)  // TF.to_tensor (/ 255) then TF.normalize ((x - mean) / std), float32 as torchvision
      p.x_out[((size_t)b * 3 + c) * plane + (size_t)oy * p.crop_w + ox] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)rgb[c], 255.0f), p.mean[c]), p.stdv[c]);
  }
  if (p.x_u8) {
#pragma unroll
    for (int c = 0; c < 3; ++c) p.x_u8[o * 3 + c] = (uint8_t)rgb[c];
  }
  // ---- label map: PIL NEAREST tables; masks: torch nearest tables ----
  if (p.y) {
    int v = p.ignore_index;
    if (inside) v = p.y[((size_t)b * p.H + t[h[8] + py]) * p.W + t[h[7] + px]];
    p.y_out[o] = (uint8_t)v;
  }
  if (p.q || p.lq) {
    const int sy = inside ? t[h[10] + py] : 0, sx = inside ? t[h[9] + px] : 0;
    if (p.q) p.q_out[o] = inside ? (p.q[((size_t)b * p.H + sy) * p.W + sx] ? 1 : 0) : 0;
    if (p.lq) p.lq_out[o] = inside ? p.lq[((size_t)b * p.H + sy) * p.W + sx] : (uint8_t)p.ignore_index;
  }
}

}  // namespace pp

static int augment_geometric_launch(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                                    const int32_t* header, const int32_t* tables, int crop_h, int crop_w, const float* mean3,
                                    const float* std3, const int* mean_val3, int ignore_index, float* x_out, uint8_t* x_u8_out,
                                    uint8_t* y_out, uint8_t* q_out, uint8_t* lq_out, void* stream) {
  using namespace pp;
  PP_CHECK_ARG(x && header && tables && mean_val3 && (x_out || x_u8_out) && (!x_out || (mean3 && std3)),
               "pp_augment_geometric: null pointer");
  PP_CHECK_ARG((!y || y_out) && (!q || q_out) && (!lq || lq_out), "pp_augment_geometric: an input map without its output");
  PP_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && crop_h > 0 && crop_h <= 65535 && crop_w > 0, "pp_augment_geometric: bad shape");
  AugParams p;
  p.x = x; p.y = y; p.q = q; p.lq = lq;
  p.B = B; p.H = H; p.W = W; p.crop_h = crop_h; p.crop_w = crop_w;
  p.header = header; p.tables = tables;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? mean3[c] : 0.f;
    p.stdv[c] = std3 ? std3[c] : 1.f;
    p.mean_val[c] = mean_val3[c];
  }
  p.ignore_index = ignore_index;
  p.x_out = x_out; p.x_u8 = x_u8_out; p.y_out = y_out; p.q_out = q_out; p.lq_out = lq_out;
  dim3 grid((unsigned)((crop_w + 255) / 256), (unsigned)crop_h, (unsigned)B);
  augment_geometric_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

extern "C" int pp_augment_geometric(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                                    const int32_t* header, const int32_t* tables, int crop_h, int crop_w, const float* mean3,
                                    const float* std3, const int* mean_val3, int ignore_index, float* x_out, uint8_t* y_out,
                                    uint8_t* q_out, uint8_t* lq_out, void* stream) {
  PP_CHECK_ARG(x_out && mean3 && std3, "pp_augment_geometric: null pointer");
  return augment_geometric_launch(x, y, q, lq, B, H, W, header, tables, crop_h, crop_w, mean3, std3, mean_val3, ignore_index, x_out,
                                  nullptr, y_out, q_out, lq_out, stream);
}

extern "C" int pp_augment_geometric_u8(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                                       const int32_t* header, const int32_t* tables, int crop_h, int crop_w,
                                       const int* mean_val3, int ignore_index, uint8_t* x_u8_out, uint8_t* y_out, uint8_t* q_out,
                                       uint8_t* lq_out, void* stream) {
  PP_CHECK_ARG(x_u8_out, "pp_augment_geometric_u8: null pointer");
  return augment_geometric_launch(x, y, q, lq, B, H, W, header, tables, crop_h, crop_w, nullptr, nullptr, mean_val3, ignore_index,
                                  nullptr, x_u8_out, y_out, q_out, lq_out, stream);
}

// =====================================================================================================================
// Photometric augmentation, datasets/base_dataset.py:129-141 - RandomApply([ColorJitter(0.8, 0.8, 0.8, 0.2)], p = 0.8) ->
// RandomGrayscale(0.2) -> GaussianBlur (cv2, kernel 10 % of the shorter side, p = 0.5) - on the uint8 crop, then
// TF.to_tensor + TF.normalize.  The reference runs torchvision's PIL path and OpenCV; every step below restates THEIR integer /
// float arithmetic, so the output equals the libraries' bit for bit (oracle/augment_oracle.py is checked against Pillow on all
// 2^24 colours and against cv2; tests/test_augment_gpu.py checks this kernel against the oracle):
//   brightness / contrast / saturation = ImageEnhance: Image.blend(degenerate, image, factor) with Blend.c's float arithmetic
//       (in1 + alpha * (in2 - in1) in fp32, truncated; clipped outside [0, 1]); degenerate = black / the rounded mean of the
//       L image / the L image, L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16;
//   hue = Convert.c rgb2hsv (fp32 quotients, the hue folded in double and rounded to fp32 before the x255), a uint8 shift with
//       wrap-around, hsv2rgb (double products, round half away);
//   grayscale = the L image in all three channels;
//   blur = OpenCV's bit-exact uint8 Gaussian: 8-bit fixed-point taps that sum to 256 (built on the host), horizontal pass
//       exact in Q8.8, vertical pass exact in Q16.16, (v + 2^15) >> 16, BORDER_REFLECT_101.
// The four colour-jitter steps run in a per-image random order; contrast needs the image-wide mean of the state just before it,
// so a first launch reduces that (recomputing the steps before it), a second applies everything point-wise.
// =====================================================================================================================
namespace pp {

constexpr int kPhotoHdr = 16;  // int32 per image: jitter_on, op[4], brightness, contrast, saturation (fp32 bits), hue shift, gray_on, blur_on

struct PhotoParams {
  const uint8_t* x;  // [B][H][W][3]
  int B, H, W, ksize;
  const int32_t* hdr;   // [B][kPhotoHdr]
  const int32_t* taps;  // [B][ksize] (may be null when no image is blurred)
  float mean[3], stdv[3];
  unsigned long long* lsum;  // [B]
  uint8_t* pt;               // [B][H][W][3] point-wise result
  uint16_t* hq;              // [B][H][W][3] horizontal blur pass, Q8.8
  float* out;                // [B][3][H][W] normalised (or null)
  uint8_t* out_u8;           // [B][H][W][3] (or null)
};

__device__ __forceinline__ int photo_blend(int deg, int v, float a) {  // Blend.c
  if (a == 0.f) return deg;
  if (a == 1.f) return v;
  const float t = __fadd_rn((float)deg, __fmul_rn(a, (float)(v - deg)));
  if (a >= 0.f && a <= 1.f) return (int)t;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}
__device__ __forceinline__ int photo_l(const int (&c)[3]) { return (c[0] * 19595 + c[1] * 38470 + c[2] * 7471 + 0x8000) >> 16; }

__device__ __forceinline__ void photo_rgb2hsv(int (&c)[3]) {  // Convert.c:rgb2hsv_row
  const int r = c[0], g = c[1], b = c[2];
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  int uh = 0, us = 0;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float s = __fdiv_rn(cr, (float)maxc);
    const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
    else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
    const double v = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);  // in [5/6, 11/6]: fmod(v, 1) = v - floor(v), exact
    const float hf = (float)(v - floor(v));
    uh = (int)__dmul_rn((double)hf, 255.0);
    us = (int)__dmul_rn((double)s, 255.0);
    uh = uh < 0 ? 0 : (uh > 255 ? 255 : uh);
    us = us < 0 ? 0 : (us > 255 ? 255 : us);
  }
  c[0] = uh; c[1] = us; c[2] = maxc;
}
__device__ __forceinline__ void photo_hsv2rgb(int (&c)[3]) {  // Convert.c:hsv2rgb
  const int h = c[0], s = c[1], v = c[2];
  if (s == 0) { c[0] = c[1] = c[2] = v; return; }
  const double h6 = __ddiv_rn(__dmul_rn((double)h, 6.0), 255.0);
  const int i = (int)floor(h6);
  const float f = (float)__dsub_rn(h6, (double)i);
  const float fs = (float)__ddiv_rn((double)s, 255.0);
  const double vd = (double)v, fsd = (double)fs, fd = (double)f;
  int p = (int)round(__dmul_rn(vd, __dsub_rn(1.0, fsd)));
  int q = (int)round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, fd))));
  int t = (int)round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, __dsub_rn(1.0, fd)))));
  p = p < 0 ? 0 : (p > 255 ? 255 : p);
  q = q < 0 ? 0 : (q > 255 ? 255 : q);
  t = t < 0 ? 0 : (t > 255 ? 255 : t);
  switch (i % 6) {
    case 0: c[0] = v; c[1] = t; c[2] = p; break;
    case 1: c[0] = q; c[1] = v; c[2] = p; break;
    case 2: c[0] = p; c[1] = v; c[2] = t; break;
    case 3: c[0] = p; c[1] = q; c[2] = v; break;
    case 4: c[0] = t; c[1] = p; c[2] = v; break;
    default: c[0] = v; c[1] = p; c[2] = q; break;
  }
}

// the colour-jitter steps of one pixel in the image's order; STOP: return at the contrast step (-> true) with the state before it
template <bool STOP>
__device__ __forceinline__ bool photo_jitter(int (&c)[3], const int32_t* h, int mean_l) {
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int op = h[1 + k];
    if (op == 0) {
      const float a = __int_as_float(h[5]);
#pragma unroll
      for (int e = 0; e < 3; ++e) c[e] = photo_blend(0, c[e], a);
    } else if (op == 1) {
      if (STOP) return true;
      const float a = __int_as_float(h[6]);
#pragma unroll
      for (int e = 0; e < 3; ++e) c[e] = photo_blend(mean_l, c[e], a);
    } else if (op == 2) {
      const float a = __int_as_float(h[7]);
      const int l = photo_l(c);
#pragma unroll
      for (int e = 0; e < 3; ++e) c[e] = photo_blend(l, c[e], a);
    } else {
      photo_rgb2hsv(c);
      c[0] = (c[0] + h[8]) & 255;
      photo_hsv2rgb(c);
    }
  }
  return false;
}

__global__ void __launch_bounds__(256) photo_mean_kernel(const PhotoParams p) {
  const int b = blockIdx.y;
  const int32_t* h = p.hdr + (size_t)b * kPhotoHdr;
  if (!h[0]) return;
  const int n = p.H * p.W;
  const uint8_t* img = p.x + (size_t)b * n * 3;
  unsigned int acc = 0;  // <= 255 * pixels per thread
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    int c[3] = {img[(size_t)i * 3], img[(size_t)i * 3 + 1], img[(size_t)i * 3 + 2]};
    photo_jitter<true>(c, h, 0);
    acc += (unsigned int)photo_l(c);
  }
  unsigned long long v = acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  __shared__ unsigned long long sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(p.lsum + b, t);
  }
}

__global__ void __launch_bounds__(256) photo_point_kernel(const PhotoParams p) {
  const int b = blockIdx.y;
  const int32_t* h = p.hdr + (size_t)b * kPhotoHdr;
  const int n = p.H * p.W;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const size_t o = ((size_t)b * n + i) * 3;
  int c[3] = {p.x[o], p.x[o + 1], p.x[o + 2]};
  if (h[0]) {
    // ImageStat.Stat(L).mean[0] = sum / count in double; ImageEnhance.Contrast: int(mean + 0.5)
    const int mean_l = (int)((double)p.lsum[b] / (double)n + 0.5);
    photo_jitter<false>(c, h, mean_l);
  }
  if (h[9]) c[0] = c[1] = c[2] = photo_l(c);
  p.pt[o] = (uint8_t)c[0]; p.pt[o + 1] = (uint8_t)c[1]; p.pt[o + 2] = (uint8_t)c[2];
}

__device__ __forceinline__ int reflect101(int v, int n) {
  if (v < 0) v = -v;
  if (v >= n) v = 2 * n - 2 - v;
  return v;
}

__global__ void __launch_bounds__(256) photo_blur_h_kernel(const PhotoParams p) {
  const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * 256 + threadIdx.x;
  const int32_t* h = p.hdr + (size_t)b * kPhotoHdr;
  if (!h[10] || p.ksize <= 0 || x >= p.W) return;
  const int32_t* k = p.taps + (size_t)b * p.ksize;
  const int r = p.ksize >> 1;
  const uint8_t* row = p.pt + ((size_t)b * p.H + y) * p.W * 3;
  int a0 = 0, a1 = 0, a2 = 0;
  for (int t = 0; t < p.ksize; ++t) {
    const uint8_t* s = row + (size_t)reflect101(x + t - r, p.W) * 3;
    const int kk = k[t];
    a0 += kk * s[0]; a1 += kk * s[1]; a2 += kk * s[2];
  }
  uint16_t* d = p.hq + (((size_t)b * p.H + y) * p.W + x) * 3;
  d[0] = (uint16_t)a0; d[1] = (uint16_t)a1; d[2] = (uint16_t)a2;  // <= 255 * 256
}

__global__ void __launch_bounds__(256) photo_finish_kernel(const PhotoParams p) {
  const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * 256 + threadIdx.x;
  const int32_t* h = p.hdr + (size_t)b * kPhotoHdr;
  if (x >= p.W) return;
  const size_t pix = ((size_t)b * p.H + y) * p.W + x;
  int c[3];
  if (h[10] && p.ksize > 0) {
    const int32_t* k = p.taps + (size_t)b * p.ksize;
    const int r = p.ksize >> 1;
    unsigned int a0 = 1u << 15, a1 = 1u << 15, a2 = 1u << 15;
    for (int t = 0; t < p.ksize; ++t) {
      const uint16_t* s = p.hq + (((size_t)b * p.H + reflect101(y + t - r, p.H)) * p.W + x) * 3;
      const unsigned int kk = (unsigned int)k[t];
      a0 += kk * s[0]; a1 += kk * s[1]; a2 += kk * s[2];
    }
    c[0] = (int)(a0 >> 16); c[1] = (int)(a1 >> 16); c[2] = (int)(a2 >> 16);
  } else {
    c[0] = p.pt[pix * 3]; c[1] = p.pt[pix * 3 + 1]; c[2] = p.pt[pix * 3 + 2];
  }
  if (p.out_u8) {
    p.out_u8[pix * 3] = (uint8_t)c[0]; p.out_u8[pix * 3 + 1] = (uint8_t)c[1]; p.out_u8[pix * 3 + 2] = (uint8_t)c[2];
  }
  if (p.out) {
    const size_t plane = (size_t)p.H * p.W;
#pragma unroll
    for (int e = 0; e < 3; ++e)
      p.out[((size_t)b * 3 + e) * plane + (size_t)y * p.W + x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)c[e], 255.0f), p.mean[e]), p.stdv[e]);
  }
}

static size_t photo_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace pp

extern "C" int pp_augment_photometric_workspace_bytes(int B, int H, int W, size_t* bytes) {
  PP_CHECK_ARG(bytes && B > 0 && H > 0 && W > 0, "pp_augment_photometric_workspace_bytes: bad args");
  const size_t n = (size_t)B * H * W * 3;
  *bytes = pp::photo_align((size_t)B * sizeof(unsigned long long)) + pp::photo_align(n) + pp::photo_align(n * sizeof(uint16_t));
  return PP_OK;
}

extern "C" int pp_augment_photometric(const uint8_t* x, int B, int H, int W, const int32_t* header, const int32_t* blur_taps,
                                      int ksize, const float* mean3, const float* std3, void* workspace, size_t workspace_bytes,
                                      float* x_out, uint8_t* x_u8_out, void* stream) {
  using namespace pp;
  PP_CHECK_ARG(x && header && workspace && (x_out || x_u8_out) && (!x_out || (mean3 && std3)), "pp_augment_photometric: null pointer");
  PP_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0 && (int64_t)H * W < (1ll << 30), "pp_augment_photometric: bad shape");
  PP_CHECK_ARG(ksize >= 0 && (ksize == 0 || (blur_taps && (ksize & 1) && ksize / 2 < H && ksize / 2 < W)),
               "pp_augment_photometric: blur kernel size %d (odd, radius smaller than the image, taps given)", ksize);
  size_t need = 0;
  pp_augment_photometric_workspace_bytes(B, H, W, &need);
  PP_CHECK_ARG(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0,
               "pp_augment_photometric: workspace too small or misaligned");
  PhotoParams p;
  p.x = x; p.B = B; p.H = H; p.W = W; p.ksize = ksize; p.hdr = header; p.taps = blur_taps;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? mean3[c] : 0.f;
    p.stdv[c] = std3 ? std3[c] : 1.f;
  }
  const size_t n = (size_t)B * H * W * 3;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  p.lsum = reinterpret_cast<unsigned long long*>(ws);
  p.pt = ws + photo_align((size_t)B * sizeof(unsigned long long));
  p.hq = reinterpret_cast<uint16_t*>(p.pt + photo_align(n));
  p.out = x_out; p.out_u8 = x_u8_out;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PP_CUDA(cudaMemsetAsync(p.lsum, 0, (size_t)B * sizeof(unsigned long long), st));
  const int npx = H * W;
  int bx = (npx + 256 * 8 - 1) / (256 * 8);
  if (bx > 1024) bx = 1024;
  photo_mean_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, st>>>(p);
  PP_LAUNCH_CHECK();
  photo_point_kernel<<<dim3((unsigned)((npx + 255) / 256), (unsigned)B), 256, 0, st>>>(p);
  PP_LAUNCH_CHECK();
  const dim3 g2((unsigned)((W + 255) / 256), (unsigned)H, (unsigned)B);
  if (ksize > 0) {
    photo_blur_h_kernel<<<g2, 256, 0, st>>>(p);
    PP_LAUNCH_CHECK();
  }
  photo_finish_kernel<<<g2, 256, 0, st>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
