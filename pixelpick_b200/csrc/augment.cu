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
  float* x_out;    // [B][3][crop_h][crop_w]
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
#pragma unroll
  for (int c = 0; c < 3; ++c)  // TF.to_tensor (/ 255) then TF.normalize ((x - mean) / std), float32 as torchvision
    p.x_out[((size_t)b * 3 + c) * plane + (size_t)oy * p.crop_w + ox] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)rgb[c], 255.0f), p.mean[c]), p.stdv[c]);
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

extern "C" int pp_augment_geometric(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                                    const int32_t* header, const int32_t* tables, int crop_h, int crop_w, const float* mean3,
                                    const float* std3, const int* mean_val3, int ignore_index, float* x_out, uint8_t* y_out,
                                    uint8_t* q_out, uint8_t* lq_out, void* stream) {
  using namespace pp;
  PP_CHECK_ARG(x && header && tables && mean3 && std3 && mean_val3 && x_out, "pp_augment_geometric: null pointer");
  PP_CHECK_ARG((!y || y_out) && (!q || q_out) && (!lq || lq_out), "pp_augment_geometric: an input map without its output");
  PP_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && crop_h > 0 && crop_h <= 65535 && crop_w > 0, "pp_augment_geometric: bad shape");
  AugParams p;
  p.x = x; p.y = y; p.q = q; p.lq = lq;
  p.B = B; p.H = H; p.W = W; p.crop_h = crop_h; p.crop_w = crop_w;
  p.header = header; p.tables = tables;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3[c];
    p.stdv[c] = std3[c];
    p.mean_val[c] = mean_val3[c];
  }
  p.ignore_index = ignore_index;
  p.x_out = x_out; p.y_out = y_out; p.q_out = q_out; p.lq_out = lq_out;
  dim3 grid((unsigned)((crop_w + 255) / 256), (unsigned)crop_h, (unsigned)B);
  augment_geometric_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
